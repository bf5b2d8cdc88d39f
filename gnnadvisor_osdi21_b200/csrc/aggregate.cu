// aggregate.cu -- the neighbour-group aggregation kernel of libgnna_b200.so (sm_100a).
//
// Replaces the five SIMT kernels of the reference (GNNAdvisor/GNNConv/GNNAdvisor_kernel.cu:
// SAG :186-259, GCN fwd :324-415, GCN bwd :478-552, GIN fwd :620-689, GIN bwd :749-814).  They
// are the same computation with three edge-weight rules, so this is ONE kernel template:
//     out[part2node[g], :] (+)= scale * sum_{k in [part_ptr[g], part_ptr[g+1])} w_k * X[col_idx[k], :]
//     SAG: w=1, scale=1      GCN: w=fl(deg[src]*deg[nid]), scale=1      GIN: w=1, scale=eps
//
// What is different from the reference design (and why, for B200):
//   * The reference keeps ONE neighbour row in flight per warp and accumulates through a
//     shared-memory read-modify-write (latency bound).  Here a "sub-warp" of
//     LPR lanes (LPR = dim_worker rounded to a power of two, the reference's dimension-worker
//     knob) owns one neighbour-group; the 32/LPR sub-warps of a warp own different groups, so no
//     lane idles when dim < 32*VEC.  Each sub-warp issues U (=8/KCH) independent 128-bit row
//     loads before it consumes any of them, accumulators live in registers.  With 32 resident
//     warps that is >100 KB of gathers in flight per SM, enough to cover HBM3e/L2 latency.
//   * Neighbour ids (and the GCN degree of each neighbour) are fetched by the sub-warp
//     cooperatively (coalesced), then broadcast with __shfl_sync -- no shared memory at all.
//   * Summation order inside a group is the reference's (serial, CSR order, product rounded
//     before the add: __fmul_rn/__fadd_rn never contract to FMA), so a node with a single
//     group is bit-identical to the reference.  Groups of one node are merged with ONE
//     vectorised `red.global.add.v4.f32` per 16 bytes (sm_90+) instead of two atomicExch per
//     float (kernel.cu:12-17); a group that covers its node's whole adjacency list (checked
//     against row_ptr, which the reference passes but never reads) uses a plain store.
//   * dim is tiled over gridDim.y when one sub-warp pass (LPR*KCH*VEC elements) does not cover
//     it, so any dim works (3703 for citeseer GIN) with bounded registers.
//   * 64-bit row offsets: num_nodes*dim may exceed 2^31 (the reference's PackedTensorAccessor32
//     cannot).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"
#include "gather.cuh"

// Launch bounds: at most 512 threads per CTA (warp_per_block <= 16) and at least 2 CTAs of that size
// per SM, i.e. a 64-register budget.  The min-blocks argument matters: with __launch_bounds__(512)
// alone ptxas aims for 32 registers (full occupancy) and, to get there, re-serialises the row loads
// (one LDG.128 in flight, seen in SASS); with the 64-register budget it keeps all eight in flight.
#ifndef GNNA_LB
#define GNNA_LB 512
#endif
// Measured on B200 (tools/sweep_dims.py, profiles/r01_sweep_*): 3 CTAs (42 registers, 48 resident warps)
// is 5-10 % faster than 2 for sub-warps of <= 16 lanes; full-warp rows (LPR = 32) and wide rows
// (KCH >= 2) are best with the 64-register budget.
#ifndef GNNA_MIN_CTAS
#define GNNA_MIN_CTAS 3
#endif
#ifndef GNNA_WIDE_MIN_CTAS
#define GNNA_WIDE_MIN_CTAS 2
#endif
#ifndef GNNA_BF16_NARROW   // 1: 128-bit bf16 rows (VEC = 8) with one chunk per lane also get the 3-CTA / 40-register budget
#define GNNA_BF16_NARROW 0
#endif
// GNNA_CHAIN=1 shortens the dependent-load chain of one neighbour-group (table -> ids -> rows ... -> row_ptr):
//   * the ids of up to 32 neighbours (a whole group at the default partSize) are fetched in ONE round trip
//     instead of one round trip per 8 or 16 neighbours;
//   * the loads the flush needs (row_ptr[src], row_ptr[src+1], degrees[src]) are issued before the gather
//     instead of after it.
// Narrow and bf16 rows are latency-bound (neither L2 bytes nor L1 wavefronts saturate), so the chain length is
// what sets their speed.  GNNA_CHAIN=0 is the previous kernel (kept for A/B runs, tools/ab_chain.py).
#ifndef GNNA_CHAIN
#define GNNA_CHAIN 1
#endif
#ifndef GNNA_CHAIN_IDS     // 32 ids per round trip
#define GNNA_CHAIN_IDS GNNA_CHAIN
#endif
#ifndef GNNA_CHAIN_HOIST   // flush loads issued before the gather
#define GNNA_CHAIN_HOIST GNNA_CHAIN
#endif

namespace gnna {

// ------------------------------------------------------------------------------------------
// fp32 output: plain vector store (group owns its row) or vector reduction (row shared)
// ------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void store_or_red(float *p, const float (&a)[V], bool own) {
    if (own) {
        if constexpr (V % 4 == 0) {
#pragma unroll
            for (int i = 0; i < V; i += 4) *reinterpret_cast<float4 *>(p + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
        } else if constexpr (V == 2) {
            *reinterpret_cast<float2 *>(p) = make_float2(a[0], a[1]);
        } else {
            p[0] = a[0];
        }
    } else {
        if constexpr (V % 4 == 0) {
#pragma unroll
            for (int i = 0; i < V; i += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "f"(a[i]), "f"(a[i + 1]),
                             "f"(a[i + 2]), "f"(a[i + 3])
                             : "memory");
        } else if constexpr (V == 2) {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a[0]), "f"(a[1]) : "memory");
        } else {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a[0]) : "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
//   T    element type of X (float / __nv_bfloat16); out is always fp32
//   VEC  elements per load (VEC*sizeof(T) in {16,8,4,2}); dim % VEC == 0
//   LPR  lanes per neighbour row (sub-warp width, power of two)
//   KCH  vector chunks per lane inside one d-tile (d-tile = LPR*KCH*VEC elements)
//   WEIGHTED  GCN: per-neighbour weight fl(deg[src]*deg[nid])
// ------------------------------------------------------------------------------------------
// flags of aggregate_kernel
enum : int {
    F_SCALE = 1,      // multiply the group sum by `scale` before the merge (GIN: eps, kernel.cu:686)
    F_ROWSCALE = 2,   // multiply the group sum by degrees[src] (GCN on pre-scaled features, see prescale_rows)
    F_ACCUMULATE = 4, // `out` already holds partial sums (edges split over several CSRs): never a plain store
};

template <typename T, int VEC, int LPR, int KCH, bool WEIGHTED>
__global__ void __launch_bounds__(GNNA_LB, (KCH >= 2 || LPR == 32 || (VEC >= 8 && !GNNA_BF16_NARROW)) ? GNNA_WIDE_MIN_CTAS : GNNA_MIN_CTAS)
aggregate_kernel(const T *__restrict__ X, float *__restrict__ out,
                 const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                 const float *__restrict__ degrees,
                 const int32_t *__restrict__ part_ptr, const int32_t *__restrict__ part2node,
                 long long num_parts, int dim, int ldx, float scale, int flags, const GateParams gate)
{
    // dim = row stride of `out` (% VEC == 0, 16-byte aligned rows); ldx = row stride of X in elements (>= dim)
    constexpr int S = 32 / LPR;                      // neighbour-groups per warp
    // neighbour ids fetched per lane per batch: 32 neighbours per batch (8 at most per lane); the exact-GCN
    // variant also carries a weight per id and keeps the short batch (it is about rounding, not speed)
    constexpr int IPL_SHORT = (LPR >= 8) ? 1 : 8 / LPR;
    constexpr int IPL_LONG = (LPR >= 32) ? 1 : (32 / LPR > 8 ? 8 : 32 / LPR);
    constexpr int IPL = (GNNA_CHAIN_IDS && !WEIGHTED) ? IPL_LONG : IPL_SHORT;
    constexpr int B = LPR * IPL;                     // neighbours per batch (>= 8)
    constexpr int U = (KCH >= 8) ? 1 : 8 / KCH;      // neighbour rows in flight per sub-warp
    static_assert(B % U == 0, "batch must be a multiple of the unroll");
    constexpr unsigned FULL = 0xffffffffu;

    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR;
    const int l = lane % LPR;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long g = warp_global * S + sub;
    const int nchunks = ldx / VEC;
    const int chunk0 = blockIdx.y * (LPR * KCH) + l;   // this lane's first chunk; then += LPR

    const bool gvalid = g < num_parts;
    int src = 0, beg = 0, end = 0;
    if (gvalid) {
        src = ldg_stream(part2node + g);
        beg = ldg_stream(part_ptr + g);
        end = ldg_stream(part_ptr + g + 1);
    }
    // Sharded path, exchange fused into the aggregation (common.h: GateParams): the rows of a peer are still crossing
    // NVLink when this kernel starts.  A CTA whose groups gather a peer's rows waits for that peer's flag -- set by the
    // peer's push after its last remote store (release, system scope) -- before it touches them; CTAs are dispatched in
    // group order and the segments are ordered by arrival, so the SMs aggregate what is there while the rest lands.
    // The wait sits HERE, after the table entries of the group were requested (they are this rank's own static data), so
    // its round trip overlaps theirs; CTAs of the rank's own segment skip it without touching memory.
    if (gate.nseg > 0) {
        const long long per_cta = (long long)(blockDim.x >> 5) * S;
        const long long g0 = (long long)blockIdx.x * per_cta;
        const long long g1 = g0 + per_cta < num_parts ? g0 + per_cta : num_parts;
        bool need = false;
        for (int sgm = 0; sgm < gate.nseg; sgm++)
            need |= gate.peer[sgm] >= 0 && g0 < gate.bounds[sgm + 1] && g1 > gate.bounds[sgm];
        if (need) {                                  // CTA-uniform
            if (threadIdx.x == 0) {
                const unsigned step = *reinterpret_cast<const volatile unsigned *>(gate.step_ptr);
                for (int sgm = 0; sgm < gate.nseg; sgm++) {
                    if (gate.peer[sgm] < 0 || g0 >= gate.bounds[sgm + 1] || g1 <= gate.bounds[sgm]) continue;
                    const volatile unsigned *f = gate.flags + gate.peer[sgm];
                    const long long t0 = clock64();
                    while (*f < step)                // relaxed polling; the acquire that orders the row loads follows
                        if (clock64() - t0 > 8000000000LL) { atomicExch(gate.error_word, 2u); break; }
                }
                asm volatile("fence.acquire.sys;" ::: "memory");
            }
            __syncthreads();
        }
    }
    const int len = max(end - beg, 0);               // end <= beg: empty group, contributes nothing
    const int maxlen = __reduce_max_sync(FULL, len); // warp-uniform trip count
    if (maxlen == 0) return;
    // neighbours every sub-warp of this warp still has: batches below this bound need no predicates
    const int minlen = __reduce_min_sync(FULL, len);

    float src_norm = 0.f;
    if (WEIGHTED && gvalid) src_norm = __ldg(degrees + src);

    float acc[KCH][VEC];
    const char *lane_base[KCH];
    const int row_bytes = ldx * (int)sizeof(T);
#pragma unroll
    for (int k = 0; k < KCH; k++) {
        // a lane whose chunk lies beyond the row end reads chunk 0 instead (same cache line as lane 0, no
        // extra traffic): its sums are never stored, and the fast path below needs no per-lane predicate
        const int c = chunk0 + k * LPR;
        lane_base[k] = reinterpret_cast<const char *>(X) + (size_t)(c < nchunks ? c : 0) * (VEC * sizeof(T));
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[k][v] = 0.f;
    }

#if GNNA_CHAIN_IDS || GNNA_CHAIN_HOIST
    // what the flush needs, requested now: the answers arrive while the rows are gathered
    bool own = false;
    float mul = (flags & F_SCALE) ? scale : 1.f;
    {
        int r0 = 0, r1 = 0;
        if (GNNA_CHAIN_HOIST && len > 0) {
            r0 = ldg_keep(row_ptr + src);
            r1 = ldg_keep(row_ptr + src + 1);
            if (flags & F_ROWSCALE) mul = ldg_keep_f(degrees + src);
        }
        // cooperative, coalesced fetch of the first batch's neighbour ids (issued before the compare below waits)
        int nid[IPL];
        float wgt[IPL];
        auto fetch_ids = [&](int base) {
#pragma unroll
            for (int i = 0; i < IPL; i++) {
                const int n = base + i * LPR + l;
                nid[i] = -1;
                wgt[i] = 0.f;
                if (n < len) {
                    nid[i] = ldg_stream(col_idx + beg + n);
                    if (WEIGHTED) wgt[i] = __fmul_rn(src_norm, __ldg(degrees + nid[i]));
                }
            }
        };
        fetch_ids(0);
        // plain store when this group is the node's whole adjacency list, else vector reduction
        if (GNNA_CHAIN_HOIST) own = !(flags & F_ACCUMULATE) && (beg == r0) && (end == r1);
        for (int base = 0; base < maxlen; base += B) {
            if (base > 0) fetch_ids(base);
            // steps whose U neighbours every sub-warp of the warp still has need no predicates
#pragma unroll
            for (int j0 = 0; j0 < B; j0 += U) {
                if (base + j0 + U <= minlen) {
                    batch_step<T, VEC, LPR, KCH, U, IPL, WEIGHTED, false>(lane_base, row_bytes, nchunks, chunk0, j0, nid, wgt, acc);
                } else {
                    if (base + j0 >= maxlen) break;
                    batch_step<T, VEC, LPR, KCH, U, IPL, WEIGHTED, true>(lane_base, row_bytes, nchunks, chunk0, j0, nid, wgt, acc);
                }
            }
        }
        if (!GNNA_CHAIN_HOIST && len > 0) {
            own = !(flags & F_ACCUMULATE) && (beg == __ldg(row_ptr + src)) && (end == __ldg(row_ptr + src + 1));
            if (flags & F_ROWSCALE) mul = __ldg(degrees + src);
        }
    }
#else
    for (int base = 0; base < maxlen; base += B) {
        // cooperative, coalesced fetch of this batch's neighbour ids (+ GCN weights in exact mode)
        int nid[IPL];
        float wgt[IPL];
#pragma unroll
        for (int i = 0; i < IPL; i++) {
            const int n = base + i * LPR + l;
            nid[i] = -1;
            wgt[i] = 0.f;
            if (n < len) {
                nid[i] = ldg_stream(col_idx + beg + n);
                if (WEIGHTED) wgt[i] = __fmul_rn(src_norm, __ldg(degrees + nid[i]));
            }
        }
        if (base + B <= minlen) {
#pragma unroll
            for (int j0 = 0; j0 < B; j0 += U)
                batch_step<T, VEC, LPR, KCH, U, IPL, WEIGHTED, false>(lane_base, row_bytes, nchunks, chunk0, j0, nid, wgt, acc);
        } else {
            const int cnt = min(B, maxlen - base);   // warp-uniform
#pragma unroll
            for (int j0 = 0; j0 < B; j0 += U) {
                if (j0 >= cnt) break;
                batch_step<T, VEC, LPR, KCH, U, IPL, WEIGHTED, true>(lane_base, row_bytes, nchunks, chunk0, j0, nid, wgt, acc);
            }
        }
    }

#endif

    if (len > 0) {
#if !(GNNA_CHAIN_IDS || GNNA_CHAIN_HOIST)
        // plain store when this group is the node's whole adjacency list, else vector reduction
        const bool own = !(flags & F_ACCUMULATE) && (beg == __ldg(row_ptr + src)) && (end == __ldg(row_ptr + src + 1));
        float mul = (flags & F_SCALE) ? scale : 1.f;
        if (flags & F_ROWSCALE) mul = __ldg(degrees + src);
#endif
        float *orow = out + (long long)src * dim;
#pragma unroll
        for (int k = 0; k < KCH; k++) {
            const int c = chunk0 + k * LPR;
            if (c < nchunks) {
                if (flags & (F_SCALE | F_ROWSCALE)) {
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[k][v] = __fmul_rn(mul, acc[k][v]);
                }
                store_or_red<VEC>(orow + (long long)c * VEC, acc[k], own);
            }
        }
    }
}

// Xs[i, 0:dim] = (degrees ? degrees[i] : 1) * X[i, 0:dim], Xs[i, dim:ldx] = 0.
// The pre-pass of the aggregation: GCN pre-scale and/or re-pack of rows whose width is not a multiple
// of four floats to a 16-byte aligned stride (so the gather can use 128-bit loads).  N*dim elements,
// i.e. 2N/E of the gather traffic.
template <int VEC>
__global__ void __launch_bounds__(256)
repack_rows_kernel(const float *X, float *Xs, const float *__restrict__ degrees,
                   long long num_nodes, int dim, int ldx)   // X == Xs is allowed (in-place pre-scale): no __restrict__, no ld.nc on X
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    if constexpr (VEC == 4) {   // dim == ldx, both 16-byte aligned
        const int cpr = dim / 4;
        const long long total = num_nodes * cpr;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const float n = degrees ? __ldg(degrees + i / cpr) : 1.f;
            float4 v = reinterpret_cast<const float4 *>(X)[i];
            v.x = __fmul_rn(n, v.x); v.y = __fmul_rn(n, v.y); v.z = __fmul_rn(n, v.z); v.w = __fmul_rn(n, v.w);
            reinterpret_cast<float4 *>(Xs)[i] = v;
        }
    } else {
        const long long total = num_nodes * ldx;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const long long r = i / ldx;
            const int c = (int)(i - r * ldx);
            float v = 0.f;
            if (c < dim) {
                v = X[r * dim + c];
                if (degrees) v = __fmul_rn(__ldg(degrees + r), v);
            }
            Xs[i] = v;
        }
    }
}

// out[i, 0:dim] = Ys[i, 0:dim] for a padded Ys [N, ld]: the inverse re-pack of the OUTPUT when dim % 4 != 0
// (the kernel then merges groups with 16-byte vector reductions into the aligned scratch instead of
// one scalar reduction per float -- measured 1.6x on the 41-wide layer of the Reddit GCN).
__global__ void __launch_bounds__(256)
unpack_rows_kernel(const float *__restrict__ Ys, float *__restrict__ out, long long num_nodes, int dim, int ld)
{
    const long long stride = (long long)gridDim.x * blockDim.x, total = num_nodes * dim;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / dim;
        out[i] = Ys[r * ld + (i - r * dim)];
    }
}

// Xb[i, 0:dim] = bf16( (degrees ? degrees[i] : 1) * X[i, 0:dim] ), Xb[i, dim:ldb] = 0, ldb % 8 == 0.
// The producer pass of the mixed-precision layers (gnna_forward_mixed / gnna_backward_mixed): halves the bytes the
// gather moves and pads odd widths (41, 47 classes) to whole 16-byte chunks in the same pass.  One thread writes
// one 16-byte chunk (8 bf16); reads are float4 when the row is 16-byte aligned.
__global__ void __launch_bounds__(256)
scale_rows_bf16_kernel(const float *__restrict__ X, __nv_bfloat16 *__restrict__ Xb, const float *__restrict__ degrees,
                       long long num_rows, int dim, int ldb)
{
    const int cpr = ldb / 8;
    const long long total = num_rows * cpr, stride = (long long)gridDim.x * blockDim.x;
    const bool aligned = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / cpr;
        const int c0 = (int)(i - r * cpr) * 8;
        const float n = degrees ? __ldg(degrees + r) : 1.f;
        const float *xr = X + r * dim;
        float f[8];
        if (aligned && c0 + 8 <= dim) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(xr + c0));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(xr + c0 + 4));
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int v = 0; v < 8; v++) f[v] = (c0 + v < dim) ? __ldg(xr + c0 + v) : 0.f;
        }
        uint32_t w[4];
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(__fmul_rn(n, f[2 * v]), __fmul_rn(n, f[2 * v + 1]));
            w[v] = *reinterpret_cast<const uint32_t *>(&h);
        }
        reinterpret_cast<uint4 *>(Xb)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ------------------------------------------------------------------------------------------
// host side: geometry choice + dispatch
// ------------------------------------------------------------------------------------------
static int pow2_floor(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }
static int pow2_ceil(int x) { int p = 1; while (p < x) p *= 2; return p; }

struct Geometry {
    int vec, lpr, kch, wpb, gy;
    long long gx;
};

// dim_worker keeps the reference's meaning (lanes that work on one neighbour row, param.py:93-103)
// but is rounded down to a power of two and never exceeds what the row needs; the lanes the
// reference would leave idle work on other groups.  dim_worker <= 0: as many lanes as the row has
// 16-byte chunks (one chunk per lane), i.e. fully coalesced 128-bit loads.
// `dim` here is the row stride the kernel loads with (ldx): fp32 rows are always re-packed to a
// multiple of four floats first, so fp32 uses 128-bit loads only; bf16 rows pick the widest vector
// that divides them.
static Geometry choose_geometry(int elem_bytes, int dim, long long num_parts, int dim_worker, int warp_per_block)
{
    Geometry g;
    const int maxvec = 16 / elem_bytes;
    g.vec = maxvec;
    if (elem_bytes == 4 && dim % 4 != 0) dim = (dim + 7) / 8 * 8;
    while (g.vec > 1 && dim % g.vec != 0) g.vec /= 2;
    const int nchunks = dim / g.vec;
    const int need = pow2_ceil(nchunks) > 32 ? 32 : pow2_ceil(nchunks);
    g.lpr = need;
    if (dim_worker > 0) {
        int want = pow2_floor(dim_worker > 32 ? 32 : dim_worker);
        if (want < g.lpr) g.lpr = want;
    } else if (need == 16) {
        g.lpr = 8;          // library's own choice for 9..16 chunks per row (fp32 D = 36..64): two chunks per lane,
    }                       // twice the groups in flight per warp -- 7 % faster when HBM-bound, equal when L2-bound
                            // (profiles/r01_params_*.txt); wider rows were not measured that way and keep one chunk per lane
    int per_lane = (nchunks + g.lpr - 1) / g.lpr;
    const int max_kch = g.vec >= 8 ? 2 : 4;        // accumulators: KCH*VEC <= 16 registers
    g.kch = per_lane >= 4 ? 4 : (per_lane >= 2 ? 2 : 1);
    if (g.kch > max_kch) g.kch = max_kch;
    g.gy = (nchunks + g.lpr * g.kch - 1) / (g.lpr * g.kch);
    g.wpb = warp_per_block <= 0 ? 4 : (warp_per_block > GNNA_LB / 32 ? GNNA_LB / 32 : warp_per_block);
    const int S = 32 / g.lpr;
    const long long per_block = (long long)g.wpb * S;
    g.gx = (num_parts + per_block - 1) / per_block;
    return g;
}

template <typename T, int VEC, int LPR, int KCH, bool W>
static cudaError_t launch(const Geometry &g, cudaStream_t st, const void *X, float *out, const int32_t *row_ptr,
                          const int32_t *col_idx, const float *deg, const int32_t *pp, const int32_t *pn,
                          long long P, int dim, int ldx, float scale, int flags, const GateParams *gate)
{
    dim3 grid((unsigned)g.gx, (unsigned)g.gy, 1), block(g.wpb * 32, 1, 1);
    GateParams gp;
    if (gate) gp = *gate;
    else gp.nseg = 0;
    aggregate_kernel<T, VEC, LPR, KCH, W><<<grid, block, 0, st>>>(reinterpret_cast<const T *>(X), out, row_ptr, col_idx,
                                                                  deg, pp, pn, P, dim, ldx, scale, flags, gp);
    return cudaGetLastError();
}

template <typename T, int VEC, int LPR, bool W, typename... A>
static cudaError_t dispatch_kch(const Geometry &g, A... a)
{
    switch (g.kch) {
        case 1: return launch<T, VEC, LPR, 1, W>(g, a...);
        case 2: return launch<T, VEC, LPR, 2, W>(g, a...);
        default:
            if constexpr (VEC >= 8) return launch<T, VEC, LPR, 2, W>(g, a...);   // never chosen (max_kch)
            else return launch<T, VEC, LPR, 4, W>(g, a...);
    }
}

template <typename T, int VEC, bool W, typename... A>
static cudaError_t dispatch_lpr(const Geometry &g, A... a)
{
    switch (g.lpr) {
        case 1: return dispatch_kch<T, VEC, 1, W>(g, a...);
        case 2: return dispatch_kch<T, VEC, 2, W>(g, a...);
        case 4: return dispatch_kch<T, VEC, 4, W>(g, a...);
        case 8: return dispatch_kch<T, VEC, 8, W>(g, a...);
        case 16: return dispatch_kch<T, VEC, 16, W>(g, a...);
        default: return dispatch_kch<T, VEC, 32, W>(g, a...);
    }
}

template <typename T, bool W, typename... A>
static cudaError_t dispatch_vec(const Geometry &g, A... a)
{
    constexpr int MAXV = 16 / (int)sizeof(T);
    if (g.vec == MAXV) return dispatch_lpr<T, MAXV, W>(g, a...);
    if (g.vec == MAXV / 2) return dispatch_lpr<T, MAXV / 2, W>(g, a...);
    if constexpr (MAXV >= 8) {
        if (g.vec == MAXV / 4) return dispatch_lpr<T, MAXV / 4, W>(g, a...);
    }
    return dispatch_lpr<T, 1, W>(g, a...);
}

// GCN rounding mode.  The reference computes fl(fl(n_i*n_j) * t_j) per edge, which costs one scattered
// 4-byte gather of degrees[nid] per edge -- measured on B200 that gather is 1/3 of the kernel's L1TEX
// wavefronts (profiles/r01_v1_*).  The default path therefore uses the algebraically identical
//     out_i = n_i * sum_j (n_j * t_j)
// i.e. one pre-scale pass over the [N, D] features (0.4 % of the gather traffic) and a weight-free
// gather; each term differs from the reference's by at most 2 roundings (~1.2e-7 relative, well
// inside the 1e-4 parity bar).  GNNA_GCN_EXACT=1 selects the reference's per-edge rounding.
static int g_gcn_exact = -1;   // -1: not decided yet (environment), 0/1: set
bool gcn_exact_mode()
{
    if (g_gcn_exact < 0) {
        const char *e = getenv("GNNA_GCN_EXACT");
        g_gcn_exact = (e && e[0] == '1') ? 1 : 0;
    }
    return g_gcn_exact == 1;
}

// Xs[i, 0:dim] = (degrees ? degrees[i] : 1) * X[i, 0:dim]; columns dim..ldx-1 zero.  X == Xs allowed when ldx == dim.
static int g_staged = -1;
static bool staged_mode()
{
    if (g_staged < 0) {
        const char *e = getenv("GNNA_STAGED");
        g_staged = (e && e[0] == '1') ? 1 : 0;
    }
    return g_staged == 1;
}

// tables of at most this many groups take the single-launch path (GNNA_SMALL_PARTS; 0 switches it off).  16 K groups
// of <= 32 neighbours are <= 0.5 M row reads: below that the general path's four stream operations cost more than its kernel.
static long long g_small_parts = -1;
static long long small_parts_limit()
{
    if (g_small_parts < 0) {
        const char *e = getenv("GNNA_SMALL_PARTS");
        g_small_parts = e ? atoll(e) : 16384;
        if (g_small_parts < 0) g_small_parts = 0;
    }
    return g_small_parts;
}

int repack_rows(const float *X, float *Xs, const float *degrees, int64_t num_nodes, int dim, int ldx, cudaStream_t stream)
{
    if (num_nodes == 0 || dim == 0) return GNNA_OK;
    const bool v4 = (ldx == dim) && (dim % 4 == 0) && ((((uintptr_t)X | (uintptr_t)Xs) & 15) == 0);
    const long long items = v4 ? (long long)num_nodes * (dim / 4) : (long long)num_nodes * ldx;
    long long blocks = (items + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    if (v4) repack_rows_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(X, Xs, degrees, num_nodes, dim, ldx);
    else repack_rows_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(X, Xs, degrees, num_nodes, dim, ldx);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

int prescale_rows(const float *X, float *Xs, const float *degrees, int64_t num_nodes, int dim, cudaStream_t stream)
{
    return repack_rows(X, Xs, degrees, num_nodes, dim, dim, stream);
}

int scale_rows_bf16(const float *X, void *Xb, const float *degrees, int64_t num_rows, int dim, int ldb, cudaStream_t stream)
{
    if (num_rows == 0 || dim == 0) return GNNA_OK;
    GNNA_REQUIRE(ldb >= dim && ldb % 8 == 0, "scale_rows_bf16: ldb %d must be a multiple of 8 and >= dim %d", ldb, dim);
    GNNA_REQUIRE(((uintptr_t)Xb & 15) == 0, "scale_rows_bf16: output not 16-byte aligned");
    const long long items = (long long)num_rows * (ldb / 8);
    long long blocks = (items + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    scale_rows_bf16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(X, reinterpret_cast<__nv_bfloat16 *>(Xb), degrees, num_rows, dim, ldb);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

// stream-ordered scratch from a PRIVATE pool per device: freed blocks up to 1 GiB stay in the pool for the next call
// (no OS round trip per aggregation), anything beyond goes back to the driver at the next synchronisation, and the
// process-wide default pool -- which other libraries and torch's cudaMallocAsync backend share -- is left alone.
static int scratch_alloc(float **p, size_t bytes, cudaStream_t stream)
{
    static cudaMemPool_t pools[64] = {nullptr};
    int dev = 0;
    GNNA_CUDA_CHECK(cudaGetDevice(&dev));
    GNNA_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        GNNA_CUDA_CHECK(cudaMemPoolCreate(&pools[dev], &props));
        unsigned long long keep = 1ULL << 30;
        cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    GNNA_CUDA_CHECK(cudaMallocFromPoolAsync((void **)p, bytes, pools[dev], stream));
    return GNNA_OK;
}

// mode: MODE_SAG / MODE_GCN / MODE_GIN, or MODE_GCN_PRESCALED (X already holds n_j * t_j)
// ldx: row stride of X in elements (0 = dim).
int aggregate(int mode, int elem_bytes, const void *X, void *out,
              const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
              const int32_t *part_ptr, const int32_t *part2node,
              int64_t num_nodes, int dim, int64_t num_parts,
              int part_size, int dim_worker, int warp_per_block, cudaStream_t stream, int ldx, int64_t num_rows_x,
              bool accumulate, const GateParams *gate)
{
    (void)part_size;  // group length is read from part_ptr; any table with sorted groups works
    GNNA_REQUIRE(mode >= MODE_SAG && mode <= MODE_GCN_PRESCALED, "aggregate: bad mode %d", mode);
    GNNA_REQUIRE(elem_bytes == 4 || elem_bytes == 2, "aggregate: element size %d not supported", elem_bytes);
    GNNA_REQUIRE(num_nodes >= 0 && dim >= 0 && num_parts >= 0, "aggregate: negative size");
    if (num_nodes == 0 || dim == 0) return GNNA_OK;
    GNNA_REQUIRE(X && out, "aggregate: null feature pointer");
    const bool gcn = (mode == MODE_GCN || mode == MODE_GCN_PRESCALED);
    GNNA_REQUIRE(!gcn || degrees, "aggregate: GCN mode needs degrees");
    if (ldx <= 0) ldx = dim;
    if (num_rows_x <= 0) num_rows_x = num_nodes;     // rows of X (>= num_nodes when X carries halo rows)
    GNNA_REQUIRE(ldx >= dim, "aggregate: ldx %d < dim %d", ldx, dim);

    GNNA_REQUIRE(num_parts == 0 || (row_ptr && col_idx && part_ptr && part2node), "aggregate: null index pointer");

    // launch-bound graphs (Cora, citeseer): ONE kernel that owns rows, writes every row once -- no zero-fill, no
    // pre-scale pass, no scratch (aggregate_small.cu)
    if (elem_bytes == 4 && !accumulate && !gate && num_parts > 0 && num_parts <= small_parts_limit() && num_nodes <= 8 * small_parts_limit()) {
        const int rc = aggregate_small(mode, (const float *)X, (float *)out, col_idx, degrees, eps, part_ptr, part2node,
                                       (long long)num_nodes, (long long)num_parts, dim, ldx, gcn_exact_mode(), stream);
        if (rc != GNNA_ERR_UNSUPPORTED) return rc;
    }

    // fp32 pre-pass into stream-ordered scratch buffers when it pays:
    //   * default GCN rounding: pre-scale rows by degrees (no per-edge degree gather afterwards)
    //   * rows not 16-byte aligned (dim % 4 != 0): re-pack X to a stride of whole 32-byte sectors so the
    //     gather uses LDG.128, and aggregate into an equally padded output that is un-packed at the end
    float *scratch = nullptr, *out_scratch = nullptr;
    float *final_out = reinterpret_cast<float *>(out);
    const int out_dim = dim;
    auto release = [&]() {
        if (scratch) cudaFreeAsync(scratch, stream);
        if (out_scratch) cudaFreeAsync(out_scratch, stream);
    };
    if (elem_bytes == 4 && num_parts > 0) {
        const bool want_scale = (mode == MODE_GCN && !gcn_exact_mode());
        const bool want_pad = (ldx % 4 != 0) || ((uintptr_t)X & 15) || ((uintptr_t)out & 15);
        if (want_scale || want_pad) {
            const int new_ld = (!want_pad) ? dim : (dim + 7) / 8 * 8;
            int rc = scratch_alloc(&scratch, sizeof(float) * (size_t)num_rows_x * (size_t)new_ld, stream);
            if (rc != GNNA_OK) return rc;
            rc = repack_rows((const float *)X, scratch, want_scale ? degrees : nullptr, num_rows_x, dim, new_ld, stream);
            if (rc != GNNA_OK) { release(); return rc; }
            X = scratch;
            ldx = new_ld;
            if (want_scale) mode = MODE_GCN_PRESCALED;
            if (want_pad) {
                rc = scratch_alloc(&out_scratch, sizeof(float) * (size_t)num_nodes * (size_t)new_ld, stream);
                if (rc != GNNA_OK) { release(); return rc; }
                out = out_scratch;
                dim = new_ld;
            }
        }
    }

    if (elem_bytes == 2 && ldx != dim && num_parts > 0) {
        // bf16 rows padded to whole 16-byte chunks (scale_rows_bf16): the kernel writes whole chunks, so it
        // aggregates into an equally padded fp32 scratch that is un-packed at the end
        GNNA_REQUIRE(ldx % 8 == 0 && !accumulate, "aggregate: padded bf16 rows need ldx %% 8 == 0 and no accumulation");
        const int rc = scratch_alloc(&out_scratch, sizeof(float) * (size_t)num_nodes * (size_t)ldx, stream);
        if (rc != GNNA_OK) return rc;
        out = out_scratch;
        dim = ldx;
    }

    // rows shared by several groups are merged with reductions, rows without neighbours stay zero
    if (!accumulate) {
        cudaError_t me = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)num_nodes * (size_t)dim, stream);
        if (me != cudaSuccess) { release(); return fail(GNNA_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(me)); }
    }
    if (num_parts == 0) return GNNA_OK;

    const Geometry g = choose_geometry(elem_bytes, ldx, num_parts, dim_worker, warp_per_block);
    if (g.gx > 0x7fffffffLL || g.gy > 65535) {
        release();
        return fail(GNNA_ERR_INVALID, "aggregate: launch too large (%lld x %d CTAs)", g.gx, g.gy);
    }
    const float scale = (mode == MODE_GIN) ? eps : 1.0f;
    int flags = 0;
    if (mode == MODE_GIN) flags |= F_SCALE;
    if (mode == MODE_GCN_PRESCALED) flags |= F_ROWSCALE;
    if (accumulate) flags |= F_ACCUMULATE;
    const bool weighted = (mode == MODE_GCN);
    float *o = reinterpret_cast<float *>(out);
    cudaError_t e;
    // opt-in: persistent kernel with the integer streams staged through TMA bulk copies (aggregate_staged.cu)
    if (staged_mode() && !gate && elem_bytes == 4 && !weighted && ldx == dim && g.gy == 1) {
        const int rc = aggregate_staged((const float *)X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                        (long long)num_parts, 0x7fffffffffffffffLL, dim, part_size, scale, flags, stream);
        if (rc != GNNA_ERR_UNSUPPORTED) {
            e = cudaSuccess;
            if (rc != GNNA_OK) { release(); return rc; }
            goto finish;
        }
    }
    // run-based software-pipelined kernel (aggregate_runs.cu) where the library's rule or the caller selects it
    if (runs_mode() != 0 && !weighted && !gate) {
        const int rc = aggregate_runs(elem_bytes, X, o, row_ptr, col_idx, degrees, part_ptr, part2node, (long long)num_nodes,
                                      (long long)num_parts, dim, ldx, scale, flags & (F_SCALE | F_ROWSCALE), stream);
        if (rc != GNNA_ERR_UNSUPPORTED) {
            e = cudaSuccess;
            if (rc != GNNA_OK) { release(); return rc; }
            goto finish;
        }
    }
    if (elem_bytes == 4) {
        if (weighted)
            e = dispatch_lpr<float, 4, true>(g, stream, X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                             (long long)num_parts, dim, ldx, scale, flags, gate);
        else
            e = dispatch_lpr<float, 4, false>(g, stream, X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                             (long long)num_parts, dim, ldx, scale, flags, gate);
    } else {
        if (weighted)
            e = dispatch_vec<__nv_bfloat16, true>(g, stream, X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                                  (long long)num_parts, dim, ldx, scale, flags, gate);
        else
            e = dispatch_vec<__nv_bfloat16, false>(g, stream, X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                                  (long long)num_parts, dim, ldx, scale, flags, gate);
    }
    count_launch(1);
finish:
    if (e == cudaSuccess && out_scratch) {
        const long long total = (long long)num_nodes * out_dim;
        long long blocks = (total + 255) / 256;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        unpack_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(out_scratch, final_out, num_nodes, out_dim, dim);
        e = cudaGetLastError();
        count_launch(1);
    }
    release();
    if (e != cudaSuccess) return fail(GNNA_ERR_CUDA, "aggregate launch: %s", cudaGetErrorString(e));
    return GNNA_OK;
}

}  // namespace gnna

extern "C" int64_t gnna_set_small_parts(int64_t limit)
{
    const long long prev = gnna::small_parts_limit();
    gnna::g_small_parts = limit < 0 ? 0 : limit;
    return prev;
}

extern "C" int gnna_set_staged(int on)
{
    const int prev = gnna::staged_mode() ? 1 : 0;
    gnna::g_staged = on ? 1 : 0;
    return prev;
}

extern "C" int gnna_set_gcn_exact(int on)
{
    const int prev = gnna::gcn_exact_mode() ? 1 : 0;
    gnna::g_gcn_exact = on ? 1 : 0;
    return prev;
}

extern "C" int gnna_query_launch(int elem_bytes, int dim, int64_t num_parts, int dim_worker, int warp_per_block,
                                 gnna_launch_info *info)
{
    if (!info || dim <= 0 || (elem_bytes != 4 && elem_bytes != 2))
        return gnna::fail(GNNA_ERR_INVALID, "gnna_query_launch: bad argument");
    gnna::Geometry g = gnna::choose_geometry(elem_bytes, dim, num_parts, dim_worker, warp_per_block);
    info->vec_width = g.vec;
    info->lanes_per_row = g.lpr;
    info->chunks_per_lane = g.kch;
    info->warps_per_block = g.wpb;
    info->groups_per_warp = 32 / g.lpr;
    info->grid_x = g.gx;
    info->grid_y = g.gy;
    info->kernels = 1;
    return GNNA_OK;
}
