// aggregate_runs.cu -- run-based, software-pipelined variant of the aggregation kernel.
//
// Same computation as aggregate.cu (reference: GNNAdvisor/GNNConv/GNNAdvisor_kernel.cu SAG :186-259, GCN :324-415,
// :478-552, GIN :620-689, :749-814) and the same in-group summation order; what changes is the scheduling:
//   * aggregate.cu: a sub-warp owns ONE neighbour-group and walks its dependent chain (table -> ids -> rows), the
//     chain is hidden by 32-48 resident warps;
//   * here a sub-warp owns a RUN of R consecutive groups and software-pipelines it: the table entries AND the
//     neighbour ids of group i+1 are requested before the rows of group i are gathered (groups are contiguous in
//     col_idx, so the ids of group i+1 start where group i ends and can be fetched speculatively, bounded by the
//     length of col_idx, and masked by the group's length once its table entry arrived).  The dependent chain per
//     group is its row loads only.  Consecutive groups of the same node are merged in registers, so a node with k
//     groups costs ceil(k/R)+1 vector reductions instead of k.
// The flush is always a vector reduction (`red.global.add.v4.f32`) into the zeroed output.
// Supported: one 16-byte chunk per lane (dim <= 128 fp32 / 256 bf16 with 128-bit rows), unweighted modes (SAG, GIN,
// pre-scaled GCN).  Everything else stays on aggregate.cu.  gnna_set_runs(R) / GNNA_RUNS=R: R > 0 forces it, 0 switches
// it off, -1 (default) lets the library choose (auto_runs below, from profiles/r01_v8_ab_runs.txt).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"
#include "gather.cuh"

#ifndef GNNA_RUNS_MIN_CTAS
#define GNNA_RUNS_MIN_CTAS 4      // x 256 threads: 64 registers, 32 resident warps
#endif
#ifndef GNNA_RUNS_DEFAULT
#define GNNA_RUNS_DEFAULT (-1)
#endif

namespace gnna {

constexpr int RUNS_THREADS = 256;

// flags: 1 = multiply by `scale` (GIN), 2 = multiply by degrees[node] (GCN on pre-scaled rows)
template <typename T, int VEC, int LPR>
__global__ void __launch_bounds__(RUNS_THREADS, GNNA_RUNS_MIN_CTAS)
aggregate_runs_kernel(const T *__restrict__ X, float *__restrict__ out, const int32_t *__restrict__ row_ptr,
                      const int32_t *__restrict__ col_idx, const float *__restrict__ degrees,
                      const int32_t *__restrict__ part_ptr, const int32_t *__restrict__ part2node,
                      long long num_nodes, long long num_parts, int run, int dim, int ldx, float scale, int flags)
{
    constexpr int S = 32 / LPR;                    // runs per warp
    constexpr int B = 32;                          // ids per round trip
    constexpr int IPL = B / LPR;
    constexpr int U = 8;                           // neighbour rows in flight per sub-warp
    constexpr unsigned FULL = 0xffffffffu;

    const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long my_begin = (warp_global * S + sub) * run;
    const long long my_end = my_begin + run < num_parts ? my_begin + run : num_parts;
    const int nchunks = ldx / VEC;
    const int row_bytes = ldx * (int)sizeof(T);
    // a lane beyond the row end reads chunk 0 (same line as lane 0) and never stores
    const char *lane_base[1] = {reinterpret_cast<const char *>(X) + (size_t)(l < nchunks ? l : 0) * 16};

    const int e_limit = __ldg(row_ptr + num_nodes);   // entries of col_idx: bound of the speculative id prefetch

    float acc[1][VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[0][v] = 0.f;

    int cur = -1;
    float mul = (flags & 1) ? scale : 1.f;
    auto flush = [&]() {
        if (cur >= 0 && l < nchunks) {
            float *dst = out + (long long)cur * dim + l * VEC;
            if (flags & 3) {
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[0][v] = __fmul_rn(mul, acc[0][v]);
            }
#pragma unroll
            for (int v = 0; v < VEC; v += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + v), "f"(acc[0][v]), "f"(acc[0][v + 1]),
                             "f"(acc[0][v + 2]), "f"(acc[0][v + 3])
                             : "memory");
        }
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[0][v] = 0.f;
    };

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
    for (int i = 0; i < run; i++) {                    // warp-uniform trip count
        const long long g = my_begin + i;
        const bool gvalid = g < my_end;
        const int src = gvalid ? src_n : -1, beg = beg_n, end = end_n;
        const int len = gvalid ? max(end - beg, 0) : 0;   // end <= beg: empty group (kernel.cu:383)
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
        if (gvalid && src != cur) {
            flush();
            cur = src;
            if (flags & 2) mul = ldg_keep_f(degrees + src);   // arrives while the node's rows are gathered
        }
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
                    batch_step<T, VEC, LPR, 1, U, IPL, false, false>(lane_base, row_bytes, nchunks, l, j0, nid, wgt, acc);
                } else {
                    if (base + j0 >= maxlen) break;
                    batch_step<T, VEC, LPR, 1, U, IPL, false, true>(lane_base, row_bytes, nchunks, l, j0, nid, wgt, acc);
                }
            }
        }
    }
    flush();
}

static int g_runs = -2;   // -2: not decided yet (environment), -1: auto, 0: off, R > 0: forced
int runs_mode()
{
    if (g_runs < -1) {
        const char *e = getenv("GNNA_RUNS");
        g_runs = e ? atoi(e) : GNNA_RUNS_DEFAULT;
        if (g_runs < -1) g_runs = -1;
        if (g_runs > 64) g_runs = 64;
    }
    return g_runs;
}

// The library's own choice (measured on B200, profiles/r01_v8_ab_runs.txt, r01_v9_ab_auto.txt).
//   * dense graphs (>= 4 groups per node, e.g. the Reddit look-alike with 16: groups are merged in registers and nearly
//     every group is full, so the speculative id prefetch is never wasted): every bf16 width (12-23 % faster; the
//     128-byte-row gather reaches the L2 port limit) and fp32 rows that are not exactly one or two 128-byte lines
//     (4, 12, 32 chunks: 3-6 %; at 8 and 16 chunks the default kernel already sits at the L2 port limit);
//   * sparse graphs (ogbn-products look-alike: 2 groups per node, HBM-bound): bf16 rows of >= 6 chunks only (3-6 %);
//     narrow rows lose up to 15 % there (prefetch wasted on partial groups) and fp32 is within +-3 %.
static int auto_runs(int elem_bytes, int nchunks, long long num_nodes, long long num_parts)
{
    const bool dense = num_parts >= 4 * num_nodes;
    if (elem_bytes == 2) {
        if (dense) return nchunks >= 8 ? 4 : 8;
        return nchunks >= 6 ? 4 : 0;
    }
    return (dense && nchunks != 8 && nchunks != 16) ? 4 : 0;
}

template <typename T, int VEC, int LPR>
static int launch_runs(const void *X, float *out, const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                       const int32_t *part_ptr, const int32_t *part2node, long long num_nodes, long long num_parts, int run,
                       int dim, int ldx, float scale, int flags, cudaStream_t stream)
{
    constexpr int S = 32 / LPR;
    const long long runs = (num_parts + run - 1) / run;
    const long long per_block = (long long)(RUNS_THREADS / 32) * S;
    const long long blocks = (runs + per_block - 1) / per_block;
    if (blocks > 0x7fffffffLL) return fail(GNNA_ERR_INVALID, "aggregate_runs: launch too large (%lld CTAs)", blocks);
    aggregate_runs_kernel<T, VEC, LPR><<<(unsigned)blocks, RUNS_THREADS, 0, stream>>>(
        reinterpret_cast<const T *>(X), out, row_ptr, col_idx, degrees, part_ptr, part2node, num_nodes, num_parts, run, dim, ldx,
        scale, flags);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

// `out` must already be zeroed (or hold partial sums).  GNNA_ERR_UNSUPPORTED when the shape has no such variant.
int aggregate_runs(int elem_bytes, const void *X, float *out, const int32_t *row_ptr, const int32_t *col_idx,
                   const float *degrees, const int32_t *part_ptr, const int32_t *part2node, long long num_nodes,
                   long long num_parts, int dim, int ldx, float scale, int flags, cudaStream_t stream)
{
    int run = runs_mode();
    if (run == 0) return GNNA_ERR_UNSUPPORTED;
    const int vec = 16 / elem_bytes;
    if (ldx % vec != 0 || dim % 4 != 0 || (((uintptr_t)X | (uintptr_t)out) & 15) || !row_ptr) return GNNA_ERR_UNSUPPORTED;
    const int nchunks = ldx / vec;
    int lpr = 4;
    while (lpr < nchunks) lpr *= 2;
    if (lpr > 32) return GNNA_ERR_UNSUPPORTED;
    if (run < 0) run = auto_runs(elem_bytes, nchunks, num_nodes, num_parts);
    if (run <= 0) return GNNA_ERR_UNSUPPORTED;
#define GNNA_RUNS_CASE(TYPE, VEC, LPR)                                                                                   \
    return launch_runs<TYPE, VEC, LPR>(X, out, row_ptr, col_idx, degrees, part_ptr, part2node, num_nodes, num_parts, run, \
                                       dim, ldx, scale, flags, stream)
    if (elem_bytes == 4) {
        switch (lpr) {
            case 4: GNNA_RUNS_CASE(float, 4, 4);
            case 8: GNNA_RUNS_CASE(float, 4, 8);
            case 16: GNNA_RUNS_CASE(float, 4, 16);
            default: GNNA_RUNS_CASE(float, 4, 32);
        }
    } else {
        switch (lpr) {
            case 4: GNNA_RUNS_CASE(__nv_bfloat16, 8, 4);
            case 8: GNNA_RUNS_CASE(__nv_bfloat16, 8, 8);
            case 16: GNNA_RUNS_CASE(__nv_bfloat16, 8, 16);
            default: GNNA_RUNS_CASE(__nv_bfloat16, 8, 32);
        }
    }
#undef GNNA_RUNS_CASE
}

}  // namespace gnna

// run length the library would use for a call of this shape under the current setting (0 = csrc/aggregate.cu)
extern "C" int gnna_query_runs(int elem_bytes, int row_elems, int64_t num_nodes, int64_t num_parts)
{
    if ((elem_bytes != 4 && elem_bytes != 2) || row_elems <= 0) return 0;
    const int vec = 16 / elem_bytes;
    if (row_elems % vec != 0) return 0;
    const int nchunks = row_elems / vec;
    if (nchunks > 32) return 0;
    const int run = gnna::runs_mode();
    return run < 0 ? gnna::auto_runs(elem_bytes, nchunks, num_nodes, num_parts) : run;
}

extern "C" int gnna_set_runs(int run)
{
    const int prev = gnna::runs_mode();
    gnna::g_runs = run < 0 ? -1 : (run > 64 ? 64 : run);
    return prev;
}
