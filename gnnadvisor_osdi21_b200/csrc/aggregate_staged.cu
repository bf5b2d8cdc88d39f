// aggregate_staged.cu -- persistent variant of the aggregation kernel with the neighbour-group table and the
// column indices streamed through TMA bulk copies (cp.async.bulk -> SASS UBLKCP) into a shared-memory ring.
//
// Same computation, same in-group summation order and same flush as aggregate.cu; what changes is where the
// integer streams come from.  In aggregate.cu every sub-warp walks a three-level dependent chain per group
// (part2node/part_ptr -> col_idx -> rows) and hides it with occupancy.  Here a CTA is persistent (grid =
// SMs x resident CTAs) and owns TILE_G = 64 consecutive groups at a time; a producer warp keeps a 3-stage ring
// full: per tile one bulk copy each for part_ptr[g0 .. g0+64] and part2node[g0 .. g0+64), then -- once those have
// landed and the tile's edge range is known -- one bulk copy of col_idx[beg .. end) (16-byte aligned span).
// Consumers (4 warps) wait on the stage's mbarriers, read their groups' ids from shared memory and run the same
// batched 128-bit row gather; when a warp is done with a stage it arrives on the stage's "empty" barrier.
// Tiles the ring cannot hold (a group longer than part_size, the last tile, a non-monotone slice) are processed
// with direct loads.  Precondition for the staged tiles: groups are non-empty (build_part never emits an empty
// group), so the 16-byte rounding of a tile's index span stays inside col_idx for every tile but the last.
//
// Selected with GNNA_STAGED=1 / gnna_set_staged(1).  Opt-in, not the default: measured on B200
// (profiles/r01_staged_vs_default.txt) it is 7 % (Reddit look-alike, L2-bound) to 17 % (ogbn-products look-alike,
// HBM-bound) SLOWER than aggregate.cu at D=64.  The integer streams are 1.5 % of the bytes and 48 resident warps
// per SM already hide their dependent chain; the ring shortens that chain but its shared memory and the
// 64-register consumers cap a SM at 24 gathering warps (6 CTAs x 4), and the gather is what needs the
// parallelism.  A variant with 8 consumer warps under a 45-register cap spills and is slower still.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.h"
#include "gather.cuh"

namespace gnna {

constexpr int ST_TILE_G = 64;        // groups per tile (multiple of 4: 16-byte aligned table slices)
constexpr int ST_STAGES = 3;
constexpr int ST_CONSUMER_WARPS = 4;
constexpr int ST_THREADS = (ST_CONSUMER_WARPS + 1) * 32;

__device__ __forceinline__ uint32_t st_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void st_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "ST_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra ST_DONE;\n\t"
        "bra ST_WAIT;\n\t"
        "ST_DONE:\n\t}" ::"r"(st_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void st_mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"(st_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}" ::"r"(st_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void st_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(st_smem_u32(bar))
                 : "memory");
}

struct StageLayout {
    int pp_off, pn_off, ci_off, stage_bytes, ci_cap;   // byte offsets inside one stage; ci_cap in ints
};

__host__ __device__ inline StageLayout stage_layout(int part_size)
{
    StageLayout L;
    L.pp_off = 0;                                   // part_ptr slice: TILE_G + 4 ints
    L.pn_off = (ST_TILE_G + 4) * 4;                 // part2node slice: TILE_G ints
    L.ci_off = L.pn_off + ST_TILE_G * 4;            // col_idx span: TILE_G * part_size (+ alignment slack) ints
    L.ci_cap = ST_TILE_G * part_size + 8;
    L.stage_bytes = (L.ci_off + L.ci_cap * 4 + 127) / 128 * 128;
    return L;
}

// slot of the control words of a stage (ints): [0] mode of the tile: 0 = staged, 1 = direct loads, 2 = no more tiles
template <int LPR>
__global__ void __launch_bounds__(ST_THREADS, 6)
aggregate_staged_kernel(const float *__restrict__ X, float *__restrict__ out,
                        const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                        const float *__restrict__ degrees,
                        const int32_t *__restrict__ part_ptr, const int32_t *__restrict__ part2node,
                        long long num_parts, long long num_edges, int dim, int part_size, float scale, int flags)
{
    constexpr int VEC = 4, KCH = 1;
    constexpr int S = 32 / LPR;                          // groups per warp pass
    constexpr int IPL = (LPR >= 8) ? 1 : 8 / LPR;
    constexpr int B = LPR * IPL;
    constexpr int U = 8;
    constexpr int GROUPS_PER_PASS = ST_CONSUMER_WARPS * S;
    constexpr int PASSES = ST_TILE_G / GROUPS_PER_PASS;  // 8 (LPR 16) ... 1 (LPR 2)
    static_assert(ST_TILE_G % GROUPS_PER_PASS == 0, "tile must be a whole number of passes");
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char st_smem[];
    const StageLayout L = stage_layout(part_size);
    uint64_t *bars = reinterpret_cast<uint64_t *>(st_smem);           // [stage][3]: full_tab, full_idx, empty
    int *ctrl = reinterpret_cast<int *>(st_smem + ST_STAGES * 3 * 8);  // [stage][4]: mode, idx base (edge id of s_ci[0])
    unsigned char *stages = st_smem + 256;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long num_tiles = (num_parts + ST_TILE_G - 1) / ST_TILE_G;

    if (tid == 0) {
        for (int s = 0; s < ST_STAGES; s++) {
            st_mbar_init(bars + s * 3 + 0, 1);                   // table slices landed (producer's expect_tx arrive)
            st_mbar_init(bars + s * 3 + 1, 1);                   // index span landed / tile published
            st_mbar_init(bars + s * 3 + 2, ST_CONSUMER_WARPS);   // every consumer warp is done with the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == ST_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ producer (one lane)
        if (lane == 0) {
            int it = 0;
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
                const int s = it % ST_STAGES;
                const uint32_t ph = (it / ST_STAGES) & 1;
                unsigned char *st = stages + (size_t)s * L.stage_bytes;
                int *c = ctrl + s * 4;
                st_mbar_wait(bars + s * 3 + 2, ph ^ 1);          // stage is free
                const long long g0 = tile * ST_TILE_G;
                const bool whole = (g0 + ST_TILE_G + 4 <= num_parts + 1);     // table slices stay inside the arrays
                int mode = 1;
                if (whole) {
                    st_mbar_expect_tx(bars + s * 3 + 0, (ST_TILE_G + 4) * 4 + ST_TILE_G * 4);
                    st_bulk_g2s(st + L.pp_off, part_ptr + g0, (ST_TILE_G + 4) * 4, bars + s * 3 + 0);
                    st_bulk_g2s(st + L.pn_off, part2node + g0, ST_TILE_G * 4, bars + s * 3 + 0);
                    st_mbar_wait(bars + s * 3 + 0, ph);
                    const int *spp = reinterpret_cast<const int *>(st + L.pp_off);
                    const int beg = spp[0], end = spp[ST_TILE_G];
                    const long long abeg = beg & ~3;
                    const long long aend = ((long long)end + 3) & ~3LL;
                    // monotone table (every group inside [beg, end)) is checked by the consumers per group
                    if (end >= beg && aend - abeg <= L.ci_cap && aend <= num_edges && beg >= 0) {
                        const uint32_t bytes = (uint32_t)(aend - abeg) * 4;
                        mode = 0;
                        c[0] = 0;                                // control words first: the arrive below releases them
                        c[1] = (int)abeg;
                        if (bytes > 0) {
                            st_mbar_expect_tx(bars + s * 3 + 1, bytes);
                            st_bulk_g2s(st + L.ci_off, col_idx + abeg, bytes, bars + s * 3 + 1);
                        } else {
                            st_mbar_arrive(bars + s * 3 + 1);
                        }
                    }
                } else {
                    st_mbar_arrive(bars + s * 3 + 0);            // keep the table barrier's phase in step
                }
                if (mode == 1) {
                    c[0] = 1;
                    c[1] = 0;
                    st_mbar_arrive(bars + s * 3 + 1);            // publish a direct-load tile (release: c[0] is visible)
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int sub = lane / LPR, l = lane % LPR;
    const int nchunks = dim / VEC;
    const char *lane_base[1] = {reinterpret_cast<const char *>(X) + (size_t)(l < nchunks ? l : 0) * 16};
    const int row_bytes = dim * 4;
    int it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
        const int s = it % ST_STAGES;
        const uint32_t ph = (it / ST_STAGES) & 1;
        unsigned char *st = stages + (size_t)s * L.stage_bytes;
        const int *c = ctrl + s * 4;
        st_mbar_wait(bars + s * 3 + 0, ph);
        st_mbar_wait(bars + s * 3 + 1, ph);
        const int mode = *reinterpret_cast<const volatile int *>(c);
        const int idx_base = *reinterpret_cast<const volatile int *>(c + 1);
        const int *spp = reinterpret_cast<const int *>(st + L.pp_off);
        const int *spn = reinterpret_cast<const int *>(st + L.pn_off);
        const int *sci = reinterpret_cast<const int *>(st + L.ci_off);
        const long long g0 = tile * ST_TILE_G;
        const int t_beg = mode == 0 ? spp[0] : 0, t_end = mode == 0 ? spp[ST_TILE_G] : 0;

#pragma unroll 1
        for (int pass = 0; pass < PASSES; pass++) {
            const int gi = pass * GROUPS_PER_PASS + warp * S + sub;      // group inside the tile
            const long long g = g0 + gi;
            const bool gvalid = g < num_parts;
            int src = 0, beg = 0, end = 0;
            bool in_smem = false;
            if (gvalid) {
                if (mode == 0) {
                    src = spn[gi];
                    beg = spp[gi];
                    end = spp[gi + 1];
                    in_smem = (beg >= t_beg) && (end <= t_end);          // a non-monotone table falls back to global loads
                } else {
                    src = ldg_stream(part2node + g);
                    beg = ldg_stream(part_ptr + g);
                    end = ldg_stream(part_ptr + g + 1);
                }
            }
            const int len = max(end - beg, 0);
            const int maxlen = __reduce_max_sync(FULL, len);
            if (maxlen == 0) continue;
            const int minlen = __reduce_min_sync(FULL, len);
            float acc[KCH][VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[0][v] = 0.f;
            for (int base = 0; base < maxlen; base += B) {
                int nid[IPL];
                float wgt[IPL];
#pragma unroll
                for (int q = 0; q < IPL; q++) {
                    const int n = base + q * LPR + l;
                    nid[q] = -1;
                    wgt[q] = 0.f;
                    if (n < len) nid[q] = in_smem ? sci[beg - idx_base + n] : ldg_stream(col_idx + beg + n);
                }
                if (base + B <= minlen) {
#pragma unroll
                    for (int j0 = 0; j0 < B; j0 += U)
                        batch_step<float, VEC, LPR, KCH, U, IPL, false, false>(lane_base, row_bytes, nchunks, l, j0, nid, wgt, acc);
                } else {
                    const int cnt = min(B, maxlen - base);
#pragma unroll
                    for (int j0 = 0; j0 < B; j0 += U) {
                        if (j0 >= cnt) break;
                        batch_step<float, VEC, LPR, KCH, U, IPL, false, true>(lane_base, row_bytes, nchunks, l, j0, nid, wgt, acc);
                    }
                }
            }
            if (len > 0 && l < nchunks) {
                const bool own = !(flags & 4) && (beg == __ldg(row_ptr + src)) && (end == __ldg(row_ptr + src + 1));
                float mul = (flags & 1) ? scale : 1.f;
                if (flags & 2) mul = __ldg(degrees + src);
                if (flags & 3) {
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[0][v] = __fmul_rn(mul, acc[0][v]);
                }
                float *o = out + (long long)src * dim + (long long)l * VEC;
                if (own) {
                    *reinterpret_cast<float4 *>(o) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
                } else {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(acc[0][0]), "f"(acc[0][1]),
                                 "f"(acc[0][2]), "f"(acc[0][3])
                                 : "memory");
                }
            }
        }
        __syncwarp();
        if (lane == 0) st_mbar_arrive(bars + s * 3 + 2);     // this warp no longer reads the stage
    }
}

// returns GNNA_OK if launched, GNNA_ERR_UNSUPPORTED if this shape has no staged variant (caller uses aggregate.cu)
int aggregate_staged(const float *X, float *out, const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                     const int32_t *part_ptr, const int32_t *part2node, long long num_parts, long long num_edges,
                     int dim, int part_size, float scale, int flags, cudaStream_t stream)
{
    if (dim % 4 != 0 || dim > 128 || dim < 8 || part_size < 1 || part_size > 64 || num_edges < 0) return GNNA_ERR_UNSUPPORTED;
    if ((((uintptr_t)part_ptr | (uintptr_t)part2node | (uintptr_t)col_idx | (uintptr_t)X | (uintptr_t)out) & 15) != 0)
        return GNNA_ERR_UNSUPPORTED;               // bulk copies and 128-bit accesses need 16-byte aligned bases
    const int nchunks = dim / 4;
    int lpr = 2;
    while (lpr < nchunks) lpr *= 2;
    const StageLayout L = stage_layout(part_size);
    const int smem = 256 + ST_STAGES * L.stage_bytes;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const long long tiles = (num_parts + ST_TILE_G - 1) / ST_TILE_G;
    int per_sm = (220 * 1024) / (smem + 1024);
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) return GNNA_ERR_UNSUPPORTED;
    long long grid = (long long)sms * per_sm;
    if (grid > tiles) grid = tiles;
#define GNNA_ST_LAUNCH(LPRV)                                                                                         \
    do {                                                                                                             \
        auto k = aggregate_staged_kernel<LPRV>;                                                                      \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                  \
        if (e != cudaSuccess) return fail(GNNA_ERR_CUDA, "staged: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
        k<<<(unsigned)grid, ST_THREADS, smem, stream>>>(X, out, row_ptr, col_idx, degrees, part_ptr, part2node,      \
                                                        num_parts, num_edges, dim, part_size, scale, flags);        \
    } while (0)
    switch (lpr) {
        case 2: GNNA_ST_LAUNCH(2); break;
        case 4: GNNA_ST_LAUNCH(4); break;
        case 8: GNNA_ST_LAUNCH(8); break;
        case 16: GNNA_ST_LAUNCH(16); break;
        default: GNNA_ST_LAUNCH(32); break;
    }
#undef GNNA_ST_LAUNCH
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

}  // namespace gnna
