// fused_gemm.cu -- aggregate -> X*W in ONE kernel: the neighbour gather feeds a tcgen05 tensor-core tile.
//
// Reference pattern being fused (GNNAdvisor/GNNConv/GNNAdvisor_kernel.cu): spmm_forward_cuda_gin
// :559-617 = aggregation kernel, then `torch::mm(tmp, weight)` (:605) re-reading the aggregated
// features from HBM.  BASELINE.json's north star asks for this fusion where the update is a true dense
// GEMM (hidden dims >= 64) and bf16 operands are acceptable: config "bf16 ... fused aggregate+X*W
// tensor-core tile".  The fp32 operators (gnna_forward_gin_f32 ...) keep the cuBLAS SGEMM.
//
// One CTA owns TILE_M = 128 consecutive destination rows:
//   1. gather: the CTA's neighbour-groups (found by binary search in part2node) are cut into contiguous
//      runs, one per sub-warp; a sub-warp sums consecutive groups of the same row in registers (same
//      batched 128-bit gather as aggregate.cu) and adds the row total into an fp32 tile in shared
//      memory (red.shared.add.f32 once per row and run, not per group);
//   2. the fp32 tile is scaled (eps or degrees[i]), written out as the aggregated features (GIN's backward
//      needs them, gnn_conv.py:109) and converted to bf16 in the canonical K-major no-swizzle UMMA
//      layout (8x16-byte core matrices); W was converted into the B operand the same way;
//   3. one elected thread issues DIN/16 `tcgen05.mma.cta_group::1.kind::f16` (M=128, N=round_up(dout,16),
//      K=16 each), accumulator in TMEM, completion through `tcgen05.commit` -> mbarrier;
//   4. epilogue: all eight warps read their TMEM lanes with `tcgen05.ld.32x32b.x16` and store out rows.
// Tensor-pipe utilisation is necessarily tiny (2*128*64*64 flops per tile against ~16 MB of gathers):
// the point of the fusion is the HBM round trip of the aggregated features it removes, not flops.
#include <cuda_bf16.h>

#include "common.h"
#include "gather.cuh"

#ifndef GNNA_CHAIN
#define GNNA_CHAIN 1   // 1: software-pipelined gather (table + ids of the next group prefetched); 0: the first version
#endif

namespace gnna {

#ifndef GNNA_FUSED_MIN_CTAS
#define GNNA_FUSED_MIN_CTAS 2     // CTAs of 512 threads per SM the register budget is set for (2: 64 registers, 3: 40)
#endif
constexpr int TILE_M = 128;
constexpr int FUSED_THREADS = 512;   // 16 warps; 2 CTAs per SM (64-register budget) keep 32 warps gathering

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (r, k) of an [R x K] bf16 operand lives at
//   (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2,   SBO = (K/8)*128      (8 rows x 16 bytes = one core matrix)
__device__ __forceinline__ uint32_t umma_offset(int r, int k, int K) {
    return (uint32_t)((r >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type 0 = SWIZZLE_NONE [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A and B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// First index of the sorted array a[0..n) whose value is >= key, found by ONE WARP: 32 probes per round trip, so a table
// of 3.7 M groups takes 5 dependent loads instead of the 22 of a binary search (the tile's CTA waits for this).
__device__ __forceinline__ long long warp_lower_bound(const int32_t *__restrict__ a, long long n, long long key, int lane)
{
    long long lo = 0, hi = n;                              // the answer lies in [lo, hi]
    while (hi > lo) {
        const long long step = (hi - lo + 31) / 32;        // >= 1
        const long long pos = lo + (long long)lane * step;
        const bool ge = (pos >= hi) || ((long long)__ldg(a + pos) >= key);
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        if (m == 0) { lo += 31 * step + 1; continue; }     // every probe is below the key
        const int f = __ffs(m) - 1;                        // first probe at or above the key
        const long long at = lo + (long long)f * step;
        hi = at < hi ? at : hi;
        if (f > 0) lo += (long long)(f - 1) * step + 1;    // the probe before it was below
        else hi = lo;
    }
    return lo;
}

template <typename T, int DIN>
__global__ void __launch_bounds__(FUSED_THREADS, GNNA_FUSED_MIN_CTAS)
fused_aggregate_gemm_kernel(const T *__restrict__ X, const float *__restrict__ W, float *__restrict__ out,
                            float *__restrict__ x_agg, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                            const float *__restrict__ degrees, const int32_t *__restrict__ part_ptr,
                            const int32_t *__restrict__ part2node, long long num_nodes, long long num_parts,
                            int dout, int npad, float scale, int flags)
{
    constexpr int VEC = 16 / (int)sizeof(T);            // elements per 128-bit load
    constexpr int LPR = DIN / VEC;                      // lanes per row: 16 (fp32, 64) ... must be <= 32
    static_assert(LPR >= 4 && LPR <= 32, "DIN/VEC must be a sub-warp width");
    constexpr int NSUB = FUSED_THREADS / LPR;           // sub-warps per CTA
    constexpr int U = 8;
#if GNNA_CHAIN
    constexpr int B = 32;                               // ids of a whole group (default partSize) in one round trip
#else
    constexpr int B = (LPR >= 8) ? LPR : 8;
#endif
    constexpr int IPL = B / LPR;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(1024) unsigned char smem[];
    float *s_tile = reinterpret_cast<float *>(smem);                                  // [TILE_M][DIN] fp32
    unsigned char *s_a = smem + TILE_M * DIN * 4;                                     // bf16 A operand
    unsigned char *s_b = s_a + TILE_M * DIN * 2;                                      // bf16 B operand [npad x DIN]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_b + (size_t)npad * DIN * 2);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 1);
    long long *s_range = reinterpret_cast<long long *>(s_bar + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row0 = (long long)blockIdx.x * TILE_M;

    // ---- setup: group range of this tile, TMEM, barrier, B operand, zero the fp32 tile
    if (warp == 1 || warp == 2) {              // first group whose node >= row0 / >= row0 + TILE_M: two warps, side by side
        const long long v = warp_lower_bound(part2node, num_parts, warp == 1 ? row0 : row0 + TILE_M, lane);
        if (lane == 0) s_range[warp - 1] = v;
    }
    if (tid == 0) {
        s_range[2] = (long long)__ldg(row_ptr + num_nodes);   // entries of col_idx: bound of the speculative id prefetch
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        int cols = npad < 32 ? 32 : npad;      // power of two >= 32 (npad is 16, 32, 64, 128 or 256)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int i = tid; i < DIN * npad; i += FUSED_THREADS) {
        const int k = i / npad, n = i - k * npad;
        const float w = (n < dout) ? __ldg(W + (long long)k * dout + n) : 0.f;
        *reinterpret_cast<__nv_bfloat16 *>(s_b + umma_offset(n, k, DIN)) = __float2bfloat16_rn(w);
    }
    for (int i = tid; i < TILE_M * DIN; i += FUSED_THREADS) s_tile[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    // ---- 1. gather: contiguous runs of groups per sub-warp, row totals into the fp32 tile
    {
        const int sub = tid / LPR, l = tid % LPR;
        const long long g_begin = s_range[0], g_end = s_range[1];
        const long long total = g_end - g_begin;
        const long long chunk = (total + NSUB - 1) / NSUB;
        const long long my_begin = g_begin + sub * chunk;
        const long long my_end = my_begin + chunk < g_end ? my_begin + chunk : g_end;
        const char *lane_base[1] = {reinterpret_cast<const char *>(X) + (size_t)l * 16};
        const int row_bytes = DIN * (int)sizeof(T);
        float acc[1][VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[0][v] = 0.f;
        int cur = -1;
        auto flush = [&]() {
            if (cur >= 0) {
                float *dst = s_tile + (size_t)(cur - row0) * DIN + l * VEC;
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    atomicAdd(dst + v, acc[0][v]);
                    acc[0][v] = 0.f;
                }
            }
        };
#if GNNA_CHAIN
        // Software pipeline over the run: the table entries AND the neighbour ids of group i+1 are requested before the
        // rows of group i are gathered, so the dependent chain of a group is its row loads only.  Groups of a run are
        // contiguous in col_idx (part_ptr[g+1] ends group g and begins group g+1), so the ids of the next group start
        // at `end`; they are fetched speculatively (bounded by the length of col_idx) and masked by the group's length
        // once the table entry arrived.
        const int e_limit = (int)s_range[2];
        int src_n = -1, beg_n = 0, end_n = 0;
        int nid_n[IPL];
        auto prefetch_ids = [&](int from) {
#pragma unroll
            for (int q = 0; q < IPL; q++) {
                const long long k = (long long)from + q * LPR + l;   // 64-bit: `from` may sit within 32 of INT_MAX
                nid_n[q] = (k < e_limit) ? ldg_stream(col_idx + k) : -1;
            }
        };
#pragma unroll
        for (int q = 0; q < IPL; q++) nid_n[q] = -1;
        if (my_begin < my_end) {
            src_n = ldg_stream(part2node + my_begin);
            beg_n = ldg_stream(part_ptr + my_begin);
            end_n = ldg_stream(part_ptr + my_begin + 1);
            prefetch_ids(beg_n);
        }
        for (long long i = 0; i < chunk; i++) {            // warp-uniform trip count
            const long long g = my_begin + i;
            const bool gvalid = g < my_end;
            const int src = gvalid ? src_n : -1, beg = beg_n, end = end_n;
            const int len = gvalid ? max(end - beg, 0) : 0;
            int nid[IPL];
            float wgt[IPL];
#pragma unroll
            for (int q = 0; q < IPL; q++) {
                nid[q] = (q * LPR + l < len) ? nid_n[q] : -1;
                wgt[q] = 0.f;
            }
            if (g + 1 < my_end) {                          // request group i+1
                src_n = ldg_stream(part2node + g + 1);
                beg_n = end;
                end_n = ldg_stream(part_ptr + g + 2);
                prefetch_ids(end);
            }
            if (gvalid && src != cur) { flush(); cur = src; }
            const int maxlen = __reduce_max_sync(FULL, len);
            const int minlen = __reduce_min_sync(FULL, len);
            for (int base = 0; base < maxlen; base += B) {
                if (base > 0) {                            // groups longer than one batch (partSize > 32): direct loads
#pragma unroll
                    for (int q = 0; q < IPL; q++) {
                        const int n = base + q * LPR + l;
                        nid[q] = (n < len) ? ldg_stream(col_idx + beg + n) : -1;
                    }
                }
#pragma unroll
                for (int j0 = 0; j0 < B; j0 += U) {
                    if (base + j0 + U <= minlen) {
                        batch_step<T, VEC, LPR, 1, U, IPL, false, false>(lane_base, row_bytes, LPR, l, j0, nid, wgt, acc);
                    } else {
                        if (base + j0 >= maxlen) break;
                        batch_step<T, VEC, LPR, 1, U, IPL, false, true>(lane_base, row_bytes, LPR, l, j0, nid, wgt, acc);
                    }
                }
            }
        }
#else
        for (long long i = 0; i < chunk; i++) {            // warp-uniform trip count
            const long long g = my_begin + i;
            const bool gvalid = g < my_end;
            int src = -1, beg = 0, end = 0;
            if (gvalid) {
                src = ldg_stream(part2node + g);
                beg = ldg_stream(part_ptr + g);
                end = ldg_stream(part_ptr + g + 1);
            }
            const int len = max(end - beg, 0);
            if (gvalid && src != cur) { flush(); cur = src; }
            const int maxlen = __reduce_max_sync(FULL, len);
            const int minlen = __reduce_min_sync(FULL, len);
            for (int base = 0; base < maxlen; base += B) {
                int nid[IPL];
                float wgt[IPL];
#pragma unroll
                for (int q = 0; q < IPL; q++) {
                    const int n = base + q * LPR + l;
                    nid[q] = (n < len) ? ldg_stream(col_idx + beg + n) : -1;
                    wgt[q] = 0.f;
                }
                if (base + B <= minlen) {
#pragma unroll
                    for (int j0 = 0; j0 < B; j0 += U)
                        batch_step<T, VEC, LPR, 1, U, IPL, false, false>(lane_base, row_bytes, LPR, l, j0, nid, wgt, acc);
                } else {
                    const int cnt = min(B, maxlen - base);
#pragma unroll
                    for (int j0 = 0; j0 < B; j0 += U) {
                        if (j0 >= cnt) break;
                        batch_step<T, VEC, LPR, 1, U, IPL, false, true>(lane_base, row_bytes, LPR, l, j0, nid, wgt, acc);
                    }
                }
            }
        }
#endif
        flush();
    }
    __syncthreads();

    // ---- 2. scale, write the aggregated features, build the bf16 A operand
    for (int i = tid; i < TILE_M * DIN; i += FUSED_THREADS) {
        const int m = i / DIN, k = i - m * DIN;
        const long long row = row0 + m;
        float v = s_tile[i];
        if (row < num_nodes) {
            if (flags & 1) v = __fmul_rn(scale, v);
            if (flags & 2) v = __fmul_rn(__ldg(degrees + row), v);
            if (x_agg) x_agg[row * DIN + k] = v;
        }
        *reinterpret_cast<__nv_bfloat16 *>(s_a + umma_offset(m, k, DIN)) = __float2bfloat16_rn(v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> async proxy (UMMA)
    __syncthreads();

    // ---- 3. MMA: D[128 x npad] (TMEM) = A[128 x DIN] * B[npad x DIN]^T
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TILE_M, npad);
            const uint32_t sbo = (DIN / 8) * 128;
#pragma unroll
            for (int ks = 0; ks < DIN / 16; ks++) {
                const uint64_t adesc = umma_desc(smem_u32(s_a) + ks * 256, 128, sbo);
                const uint64_t bdesc = umma_desc(smem_u32(s_b) + ks * 256, 128, sbo);
                const uint32_t accum = ks > 0 ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base),
                    "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(s_bar))
                         : "memory");
        }
        __syncwarp();
    }

    // ---- 4. epilogue: TMEM -> registers -> out
    mbar_wait(s_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int q = warp & 3;                          // TMEM lanes 32q .. 32q+31 are the only ones this warp may read
        constexpr int WARP_SETS = FUSED_THREADS / 128;   // warps with the same q split the columns between them
        const int parts = (npad / 16) < WARP_SETS ? (npad / 16) : WARP_SETS;
        const int h = warp >> 2;
        if (h < parts) {
            const int per = ((npad / 16 + parts - 1) / parts) * 16;
            const int c_begin = h * per, c_end = min(c_begin + per, npad);
            const long long row = row0 + q * 32 + lane;
            for (int c = c_begin; c < c_end; c += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < num_nodes) {
                    float *o = out + row * dout + c;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (c + j < dout) o[j] = __uint_as_float(r[j]);
                }
            }
        }
    }

    // ---- 5. release TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        int cols = npad < 32 ? 32 : npad;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols));
    }
}

static int fused_smem_bytes(int din, int npad) { return TILE_M * din * 4 + TILE_M * din * 2 + npad * din * 2 + 64; }

template <typename T, int DIN>
static int launch_fused(const void *X, const float *W, float *out, float *x_agg, const int32_t *row_ptr,
                        const int32_t *col_idx, const float *degrees,
                        const int32_t *part_ptr, const int32_t *part2node, int64_t num_nodes, int64_t num_parts, int dout,
                        int npad, float scale, int flags, cudaStream_t stream)
{
    const int smem = fused_smem_bytes(DIN, npad);
    if (smem > 227 * 1024)
        return fail(GNNA_ERR_UNSUPPORTED, "aggregate_gemm_fused: din %d x dout %d needs %d bytes of shared memory", DIN, dout, smem);
    auto kern = fused_aggregate_gemm_kernel<T, DIN>;
    GNNA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long tiles = (num_nodes + TILE_M - 1) / TILE_M;
    kern<<<(unsigned)tiles, FUSED_THREADS, smem, stream>>>(reinterpret_cast<const T *>(X), W, out, x_agg, row_ptr, col_idx, degrees,
                                                          part_ptr, part2node, (long long)num_nodes, (long long)num_parts,
                                                          dout, npad, scale, flags);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

}  // namespace gnna

using namespace gnna;

extern "C" int gnna_aggregate_gemm_fused_bf16(int mode, const void *X, int x_is_bf16, const float *W, float eps,
                                              float *out, float *x_agg,
                                              const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                                              const int32_t *part_ptr, const int32_t *part2node,
                                              int64_t num_nodes, int din, int dout, int64_t num_parts,
                                              int part_size, int dim_worker, int warp_per_block, void *stream_)
{
    (void)part_size; (void)dim_worker; (void)warp_per_block;
    cudaStream_t stream = (cudaStream_t)stream_;
    GNNA_REQUIRE(mode == MODE_SAG || mode == MODE_GIN || mode == MODE_GCN_PRESCALED,
                 "aggregate_gemm_fused: mode %d not supported (0 SAG, 2 GIN, 3 GCN on pre-scaled features)", mode);
    GNNA_REQUIRE(num_nodes >= 0 && num_parts >= 0, "aggregate_gemm_fused: negative size");
    if (num_nodes == 0) return GNNA_OK;
    GNNA_REQUIRE(X && W && out && row_ptr && (num_parts == 0 || (col_idx && part_ptr && part2node)), "aggregate_gemm_fused: null pointer");
    GNNA_REQUIRE(mode != MODE_GCN_PRESCALED || degrees, "aggregate_gemm_fused: GCN mode needs degrees");
    GNNA_REQUIRE(dout >= 1 && dout <= 256, "aggregate_gemm_fused: dout %d out of range (1..256)", dout);
    int npad = 16;
    while (npad < dout) npad *= 2;                       // 16, 32, 64, 128, 256: a legal UMMA N and a TMEM column count
    const float scale = (mode == MODE_GIN) ? eps : 1.f;
    const int flags = (mode == MODE_GIN ? 1 : 0) | (mode == MODE_GCN_PRESCALED ? 2 : 0);
#define GNNA_FUSED_CASE(TYPE, DIN)                                                                               \
    return launch_fused<TYPE, DIN>(X, W, out, x_agg, row_ptr, col_idx, degrees, part_ptr, part2node, num_nodes, num_parts, \
                                   dout, npad, scale, flags, stream)
    if (!x_is_bf16) {
        if (din == 64) GNNA_FUSED_CASE(float, 64);
        if (din == 128) GNNA_FUSED_CASE(float, 128);
        if (din == 32) GNNA_FUSED_CASE(float, 32);
    } else {
        if (din == 64) GNNA_FUSED_CASE(__nv_bfloat16, 64);
        if (din == 128) GNNA_FUSED_CASE(__nv_bfloat16, 128);
        if (din == 256) GNNA_FUSED_CASE(__nv_bfloat16, 256);
    }
#undef GNNA_FUSED_CASE
    return fail(GNNA_ERR_UNSUPPORTED, "aggregate_gemm_fused: din %d (%s) has no fused tile; use the unfused operators", din,
                x_is_bf16 ? "bf16" : "fp32");
}
