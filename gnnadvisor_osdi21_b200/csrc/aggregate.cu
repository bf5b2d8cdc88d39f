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

#include "common.h"

// Launch bounds: at most 512 threads per CTA (warp_per_block <= 16) and at least 2 CTAs of that size
// per SM, i.e. a 64-register budget.  The min-blocks argument matters: with __launch_bounds__(512)
// alone ptxas aims for 32 registers (full occupancy) and, to get there, re-serialises the row loads
// (one LDG.128 in flight, seen in SASS); with the 64-register budget it keeps all eight in flight.
#ifndef GNNA_LB
#define GNNA_LB 512
#endif
#ifndef GNNA_MIN_CTAS
#define GNNA_MIN_CTAS 2
#endif

namespace gnna {

// ------------------------------------------------------------------------------------------
// vector loads of VEC elements of T through the read-only path, unpacked to fp32
// ------------------------------------------------------------------------------------------
template <int BYTES> struct Raw;
template <> struct Raw<16> { uint4 v; };
template <> struct Raw<8> { uint2 v; };
template <> struct Raw<4> { uint32_t v; };
template <> struct Raw<2> { uint16_t v; };

// Predicated read-only loads as volatile asm: the U*KCH loads of one batch must be ISSUED before
// the first is consumed (that is the memory-level parallelism this kernel lives on).  Written as
// `if (ok) x = __ldg(p); else x = 0;` the compiler sinks every load next to its use and keeps a
// single row in flight (seen in SASS: the same destination registers reused back to back).
__device__ __forceinline__ void ldg_raw(Raw<16> &r, const void *p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
                 "@p ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "=r"(r.v.x), "=r"(r.v.y), "=r"(r.v.z), "=r"(r.v.w) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ldg_raw(Raw<8> &r, const void *p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t"
                 "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
                 "@p ld.global.nc.v2.b32 {%0, %1}, [%2];\n\t}"
                 : "=r"(r.v.x), "=r"(r.v.y) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ldg_raw(Raw<4> &r, const void *p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
                 "mov.b32 %0, 0;\n\t"
                 "@p ld.global.nc.b32 %0, [%1];\n\t}"
                 : "=r"(r.v) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ void ldg_raw(Raw<2> &r, const void *p, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
                 "mov.b16 %0, 0;\n\t"
                 "@p ld.global.nc.b16 %0, [%1];\n\t}"
                 : "=h"(r.v) : "l"(p), "r"((int)ok));
}

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// fp32
__device__ __forceinline__ void unpack(const Raw<16> &r, float (&f)[4], float) {
    f[0] = __uint_as_float(r.v.x); f[1] = __uint_as_float(r.v.y); f[2] = __uint_as_float(r.v.z); f[3] = __uint_as_float(r.v.w);
}
__device__ __forceinline__ void unpack(const Raw<8> &r, float (&f)[2], float) {
    f[0] = __uint_as_float(r.v.x); f[1] = __uint_as_float(r.v.y);
}
__device__ __forceinline__ void unpack(const Raw<4> &r, float (&f)[1], float) { f[0] = __uint_as_float(r.v); }
// bf16
__device__ __forceinline__ void unpack(const Raw<16> &r, float (&f)[8], __nv_bfloat16) {
    f[0] = bf16lo(r.v.x); f[1] = bf16hi(r.v.x); f[2] = bf16lo(r.v.y); f[3] = bf16hi(r.v.y);
    f[4] = bf16lo(r.v.z); f[5] = bf16hi(r.v.z); f[6] = bf16lo(r.v.w); f[7] = bf16hi(r.v.w);
}
__device__ __forceinline__ void unpack(const Raw<8> &r, float (&f)[4], __nv_bfloat16) {
    f[0] = bf16lo(r.v.x); f[1] = bf16hi(r.v.x); f[2] = bf16lo(r.v.y); f[3] = bf16hi(r.v.y);
}
__device__ __forceinline__ void unpack(const Raw<4> &r, float (&f)[2], __nv_bfloat16) { f[0] = bf16lo(r.v); f[1] = bf16hi(r.v); }
__device__ __forceinline__ void unpack(const Raw<2> &r, float (&f)[1], __nv_bfloat16) { f[0] = bf16lo(r.v); }

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
template <typename T, int VEC, int LPR, int KCH, bool WEIGHTED>
__global__ void __launch_bounds__(GNNA_LB, GNNA_MIN_CTAS)
aggregate_kernel(const T *__restrict__ X, float *__restrict__ out,
                 const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                 const float *__restrict__ degrees,
                 const int32_t *__restrict__ part_ptr, const int32_t *__restrict__ part2node,
                 long long num_parts, int dim, float scale, int apply_scale)
{
    constexpr int S = 32 / LPR;                      // neighbour-groups per warp
    constexpr int IPL = (LPR >= 8) ? 1 : 8 / LPR;    // neighbour ids fetched per lane per batch
    constexpr int B = LPR * IPL;                     // neighbours per batch (>= 8)
    constexpr int U = (KCH >= 8) ? 1 : 8 / KCH;      // neighbour rows in flight per sub-warp
    static_assert(B % U == 0, "batch must be a multiple of the unroll");
    using RawT = Raw<VEC * (int)sizeof(T)>;
    constexpr unsigned FULL = 0xffffffffu;

    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR;
    const int l = lane % LPR;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long g = warp_global * S + sub;
    const int nchunks = dim / VEC;
    const int chunk0 = blockIdx.y * (LPR * KCH) + l;   // this lane's first chunk; then += LPR

    const bool gvalid = g < num_parts;
    int src = 0, beg = 0, end = 0;
    if (gvalid) {
        src = __ldg(part2node + g);
        beg = __ldg(part_ptr + g);
        end = __ldg(part_ptr + g + 1);
    }
    const int len = max(end - beg, 0);               // end <= beg: empty group, contributes nothing
    const int maxlen = __reduce_max_sync(FULL, len); // warp-uniform trip count
    if (maxlen == 0) return;

    float src_norm = 0.f;
    if (WEIGHTED && gvalid) src_norm = __ldg(degrees + src);

    float acc[KCH][VEC];
#pragma unroll
    for (int k = 0; k < KCH; k++)
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[k][v] = 0.f;

    for (int base = 0; base < maxlen; base += B) {
        // cooperative, coalesced fetch of this batch's neighbour ids (+ GCN weights)
        int nid[IPL];
        float wgt[IPL];
#pragma unroll
        for (int i = 0; i < IPL; i++) {
            const int n = base + i * LPR + l;
            nid[i] = -1;
            wgt[i] = 0.f;
            if (n < len) {
                nid[i] = __ldg(col_idx + beg + n);
                if (WEIGHTED) wgt[i] = __fmul_rn(src_norm, __ldg(degrees + nid[i]));
            }
        }
        const int cnt = min(B, maxlen - base);       // warp-uniform
#pragma unroll
        for (int j0 = 0; j0 < B; j0 += U) {
            if (j0 >= cnt) break;
            RawT raw[U][KCH];
            float w[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = j0 + u;                // compile-time after unrolling
                const int nj = __shfl_sync(FULL, nid[j / LPR], j % LPR, LPR);
                if (WEIGHTED) w[u] = __shfl_sync(FULL, wgt[j / LPR], j % LPR, LPR);
                const T *row = X + (long long)nj * dim;
#pragma unroll
                for (int k = 0; k < KCH; k++) {
                    const int c = chunk0 + k * LPR;
                    ldg_raw(raw[u][k], row + (long long)c * VEC, nj >= 0 && c < nchunks);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int k = 0; k < KCH; k++) {
                    float f[VEC];
                    unpack(raw[u][k], f, T());
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        if (WEIGHTED) acc[k][v] = __fadd_rn(acc[k][v], __fmul_rn(w[u], f[v]));
                        else acc[k][v] = __fadd_rn(acc[k][v], f[v]);
                    }
                }
            }
        }
    }

    if (len > 0) {
        // plain store when this group is the node's whole adjacency list, else vector reduction
        const bool own = (beg == __ldg(row_ptr + src)) && (end == __ldg(row_ptr + src + 1));
        float *orow = out + (long long)src * dim;
#pragma unroll
        for (int k = 0; k < KCH; k++) {
            const int c = chunk0 + k * LPR;
            if (c < nchunks) {
                if (apply_scale) {
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[k][v] = __fmul_rn(scale, acc[k][v]);
                }
                store_or_red<VEC>(orow + (long long)c * VEC, acc[k], own);
            }
        }
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
static Geometry choose_geometry(int elem_bytes, int dim, long long num_parts, int dim_worker, int warp_per_block)
{
    Geometry g;
    const int maxvec = 16 / elem_bytes;
    g.vec = maxvec;
    while (g.vec > 1 && dim % g.vec != 0) g.vec /= 2;
    const int nchunks = dim / g.vec;
    const int need = pow2_ceil(nchunks) > 32 ? 32 : pow2_ceil(nchunks);
    g.lpr = need;
    if (dim_worker > 0) {
        int want = pow2_floor(dim_worker > 32 ? 32 : dim_worker);
        if (want < g.lpr) g.lpr = want;
    }
    int per_lane = (nchunks + g.lpr - 1) / g.lpr;
    const int max_kch = g.vec >= 8 ? 2 : 4;        // accumulators: KCH*VEC <= 16 registers
    g.kch = per_lane >= 4 ? 4 : (per_lane >= 2 ? 2 : 1);
    if (g.kch > max_kch) g.kch = max_kch;
    g.gy = (nchunks + g.lpr * g.kch - 1) / (g.lpr * g.kch);
    g.wpb = warp_per_block <= 0 ? 8 : (warp_per_block > GNNA_LB / 32 ? GNNA_LB / 32 : warp_per_block);
    const int S = 32 / g.lpr;
    const long long per_block = (long long)g.wpb * S;
    g.gx = (num_parts + per_block - 1) / per_block;
    return g;
}

template <typename T, int VEC, int LPR, int KCH, bool W>
static cudaError_t launch(const Geometry &g, cudaStream_t st, const void *X, float *out, const int32_t *row_ptr,
                          const int32_t *col_idx, const float *deg, const int32_t *pp, const int32_t *pn,
                          long long P, int dim, float scale, int apply_scale)
{
    dim3 grid((unsigned)g.gx, (unsigned)g.gy, 1), block(g.wpb * 32, 1, 1);
    aggregate_kernel<T, VEC, LPR, KCH, W><<<grid, block, 0, st>>>(reinterpret_cast<const T *>(X), out, row_ptr, col_idx,
                                                                  deg, pp, pn, P, dim, scale, apply_scale);
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

int aggregate(int mode, int elem_bytes, const void *X, void *out,
              const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
              const int32_t *part_ptr, const int32_t *part2node,
              int64_t num_nodes, int dim, int64_t num_parts,
              int part_size, int dim_worker, int warp_per_block, cudaStream_t stream)
{
    (void)part_size;  // group length is read from part_ptr; any table with sorted groups works
    GNNA_REQUIRE(mode >= MODE_SAG && mode <= MODE_GIN, "aggregate: bad mode %d", mode);
    GNNA_REQUIRE(elem_bytes == 4 || elem_bytes == 2, "aggregate: element size %d not supported", elem_bytes);
    GNNA_REQUIRE(num_nodes >= 0 && dim >= 0 && num_parts >= 0, "aggregate: negative size");
    if (num_nodes == 0 || dim == 0) return GNNA_OK;
    GNNA_REQUIRE(X && out, "aggregate: null feature pointer");
    GNNA_REQUIRE(mode != MODE_GCN || degrees, "aggregate: GCN mode needs degrees");

    // rows shared by several groups are merged with reductions, rows without neighbours stay zero
    GNNA_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)num_nodes * (size_t)dim, stream));
    if (num_parts == 0) return GNNA_OK;
    GNNA_REQUIRE(row_ptr && col_idx && part_ptr && part2node, "aggregate: null index pointer");

    const Geometry g = choose_geometry(elem_bytes, dim, num_parts, dim_worker, warp_per_block);
    GNNA_REQUIRE(g.gx <= 0x7fffffffLL, "aggregate: too many neighbour groups for one launch (%lld CTAs)", g.gx);
    GNNA_REQUIRE(g.gy <= 65535, "aggregate: dim %d needs %d d-tiles (> 65535)", dim, g.gy);
    const float scale = (mode == MODE_GIN) ? eps : 1.0f;
    const int apply_scale = (mode == MODE_GIN) ? 1 : 0;
    float *o = reinterpret_cast<float *>(out);
    cudaError_t e;
    if (elem_bytes == 4) {
        if (mode == MODE_GCN)
            e = dispatch_vec<float, true>(g, stream, X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                          (long long)num_parts, dim, scale, apply_scale);
        else
            e = dispatch_vec<float, false>(g, stream, X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                           (long long)num_parts, dim, scale, apply_scale);
    } else {
        if (mode == MODE_GCN)
            e = dispatch_vec<__nv_bfloat16, true>(g, stream, X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                                  (long long)num_parts, dim, scale, apply_scale);
        else
            e = dispatch_vec<__nv_bfloat16, false>(g, stream, X, o, row_ptr, col_idx, degrees, part_ptr, part2node,
                                                   (long long)num_parts, dim, scale, apply_scale);
    }
    if (e != cudaSuccess) return fail(GNNA_ERR_CUDA, "aggregate launch: %s", cudaGetErrorString(e));
    count_launch(1);
    return GNNA_OK;
}

}  // namespace gnna

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
