// aggregate_small.cu -- single-launch aggregation for graphs whose whole aggregation is launch-latency bound
// (Cora, citeseer: BASELINE.json configs #1/#2; 10^4 edges, a 10 us kernel).
//
// The general path (aggregate.cu) costs four stream operations per call on such a graph -- zero-fill of the output
// (rows shared by several groups are merged with reductions), the GCN pre-scale pass into a scratch buffer
// (cudaMallocFromPoolAsync + kernel) and the gather -- 21 us for a 13 us kernel on the Cora look-alike.
// Here ONE kernel does everything and nothing is accumulated into memory:
//   * a sub-warp owns a destination ROW, not a group: the sub-warp that holds the FIRST group of a row (part2node[g-1] !=
//     part2node[g]; the table lists groups in node order, GNNAdvisor.cpp:233-249) walks all the groups of that row in
//     table order, the others retire at once.  The group table keeps its meaning -- group g is col_idx[part_ptr[g] ..
//     part_ptr[g+1]), an empty or negative range contributes nothing (kernel.cu:383, so F6 tables behave as in the
//     reference) -- only the merge changes: in registers, in ascending group order (the oracle's order), not by atomics;
//   * every row is WRITTEN exactly once with plain stores -- rows without any group are zero-filled by the sub-warp of the
//     next row that has one (and the last one fills the tail) -- so there is no memset and no reduction;
//   * GCN needs no pre-scaled copy: the per-neighbour degree is gathered (10^4 extra 4-byte loads are free here) and the
//     arithmetic is the general path's, rounding for rounding: default mode sum_j fl(n_j x_j) per group, fl(n_i * group)
//     added to the row; exact mode (gnna_set_gcn_exact) fl(fl(n_i n_j) x_j) as kernel.cu:389,403;
//   * any width up to 512 floats, 128-bit loads when rows are 16-byte aligned, scalar otherwise (7 or 6 classes): no re-pack.
// Work per sub-warp is proportional to the row's degree, which is why the library only takes this path for small tables
// (aggregate.cu: small_path()).  No reference counterpart beyond the kernels it replaces (kernel.cu:186-259, 324-415, 620-689).
#include "common.h"

namespace gnna {

constexpr int SMALL_KMAX = 4;     // vector chunks per lane
constexpr int SMALL_UNROLL = 4;   // neighbour rows in flight per sub-warp

enum : int { SM_SCALE = 1, SM_ROWSCALE = 2, SM_WEIGHTED = 4, SM_PRESCALE_EDGE = 8 };

template <int VEC>
__device__ __forceinline__ void load_row(const float *p, float (&v)[VEC])
{
    if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_row(float *p, const float (&v)[VEC])
{
    if constexpr (VEC == 4) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else p[0] = v[0];
}

// lpr: lanes per row (power of two <= 32); chunks: dim / VEC; a lane owns chunks lane, lane+lpr, ... (<= SMALL_KMAX of them)
template <int VEC>
__global__ void __launch_bounds__(128)
aggregate_small_kernel(const float *__restrict__ X, float *__restrict__ out, const int32_t *__restrict__ col_idx,
                       const float *__restrict__ degrees, const int32_t *__restrict__ part_ptr,
                       const int32_t *__restrict__ part2node, long long num_parts, long long num_nodes, int dim, int ldx,
                       int lpr, float scale, int flags)
{
    const long long sub = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / lpr;
    const int lane = threadIdx.x & (lpr - 1);
    if (sub >= num_parts) return;
    const int row = __ldg(part2node + sub);
    const int prev = sub > 0 ? __ldg(part2node + sub - 1) : -1;
    if (prev == row) return;                                   // not the first group of its row
    const int chunks = dim / VEC;
    // rows without any group between the previous group's row and this one: all zero
    float zero[VEC];
#pragma unroll
    for (int i = 0; i < VEC; i++) zero[i] = 0.f;
    for (long long r = (long long)prev + 1; r < row; r++)
        for (int c = lane; c < chunks; c += lpr) store_row<VEC>(out + r * dim + c * VEC, zero);

    float acc[SMALL_KMAX][VEC];
#pragma unroll
    for (int k = 0; k < SMALL_KMAX; k++)
#pragma unroll
        for (int i = 0; i < VEC; i++) acc[k][i] = 0.f;
    const float n_i = (flags & (SM_ROWSCALE | SM_WEIGHTED)) ? __ldg(degrees + row) : 1.f;

    long long g = sub;
    for (; g < num_parts && __ldg(part2node + g) == row; g++) {
        const int beg = __ldg(part_ptr + g), end = __ldg(part_ptr + g + 1);
        if (end <= beg) continue;                              // kernel.cu:383: an empty (or F6) group adds nothing
        float ga[SMALL_KMAX][VEC];
#pragma unroll
        for (int k = 0; k < SMALL_KMAX; k++)
#pragma unroll
            for (int i = 0; i < VEC; i++) ga[k][i] = 0.f;
        for (int e = beg; e < end; e += SMALL_UNROLL) {
            int nid[SMALL_UNROLL];
            float w[SMALL_UNROLL];
#pragma unroll
            for (int u = 0; u < SMALL_UNROLL; u++) {
                nid[u] = (e + u < end) ? __ldg(col_idx + e + u) : -1;
                w[u] = 1.f;
            }
            if (flags & (SM_WEIGHTED | SM_PRESCALE_EDGE)) {
#pragma unroll
                for (int u = 0; u < SMALL_UNROLL; u++)
                    if (nid[u] >= 0) {
                        const float n_j = __ldg(degrees + nid[u]);
                        w[u] = (flags & SM_WEIGHTED) ? __fmul_rn(n_i, n_j) : n_j;
                    }
            }
#pragma unroll
            for (int k = 0; k < SMALL_KMAX; k++) {
                const int c = lane + k * lpr;
                if (c < chunks) {
                    float v[SMALL_UNROLL][VEC];
#pragma unroll
                    for (int u = 0; u < SMALL_UNROLL; u++)
                        if (nid[u] >= 0) load_row<VEC>(X + (long long)nid[u] * ldx + c * VEC, v[u]);
#pragma unroll
                    for (int u = 0; u < SMALL_UNROLL; u++)
                        if (nid[u] >= 0) {
#pragma unroll
                            for (int i = 0; i < VEC; i++) {
                                const float t = (flags & (SM_WEIGHTED | SM_PRESCALE_EDGE)) ? __fmul_rn(w[u], v[u][i]) : v[u][i];
                                ga[k][i] = __fadd_rn(ga[k][i], t);
                            }
                        }
                }
            }
        }
        // merge the group into the row exactly as the general path's flush scales it before its reduction
#pragma unroll
        for (int k = 0; k < SMALL_KMAX; k++)
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                float t = ga[k][i];
                if (flags & SM_SCALE) t = __fmul_rn(scale, t);
                if (flags & SM_ROWSCALE) t = __fmul_rn(n_i, t);
                acc[k][i] = __fadd_rn(acc[k][i], t);
            }
    }
#pragma unroll
    for (int k = 0; k < SMALL_KMAX; k++) {
        const int c = lane + k * lpr;
        if (c < chunks) store_row<VEC>(out + (long long)row * dim + c * VEC, acc[k]);
    }
    if (g >= num_parts)                                        // this was the last row with a group: zero the tail
        for (long long r = (long long)row + 1; r < num_nodes; r++)
            for (int c = lane; c < chunks; c += lpr) store_row<VEC>(out + r * dim + c * VEC, zero);
}

// mode as aggregate(): MODE_SAG / MODE_GCN / MODE_GIN / MODE_GCN_PRESCALED.  Returns GNNA_ERR_UNSUPPORTED when the shape
// does not fit (the caller then takes the general path).
int aggregate_small(int mode, const float *X, float *out, const int32_t *col_idx, const float *degrees, float eps,
                    const int32_t *part_ptr, const int32_t *part2node, long long num_nodes, long long num_parts, int dim, int ldx,
                    bool exact_gcn, cudaStream_t stream)
{
    if (num_parts <= 0 || dim <= 0) return GNNA_ERR_UNSUPPORTED;
    const bool v4 = (dim % 4 == 0) && (ldx % 4 == 0) && ((((uintptr_t)X | (uintptr_t)out) & 15) == 0);
    const int chunks = v4 ? dim / 4 : dim;
    int lpr = 1;
    while (lpr < chunks && lpr < 32) lpr <<= 1;
    if (chunks > lpr * SMALL_KMAX) return GNNA_ERR_UNSUPPORTED;
    int flags = 0;
    if (mode == MODE_GIN) flags |= SM_SCALE;
    if (mode == MODE_GCN_PRESCALED) flags |= SM_ROWSCALE;
    if (mode == MODE_GCN) flags |= exact_gcn ? SM_WEIGHTED : (SM_ROWSCALE | SM_PRESCALE_EDGE);
    const long long threads = num_parts * lpr;
    const long long blocks = (threads + 127) / 128;
    if (blocks > 0x7fffffffLL) return GNNA_ERR_UNSUPPORTED;
    if (v4)
        aggregate_small_kernel<4><<<(unsigned)blocks, 128, 0, stream>>>(X, out, col_idx, degrees, part_ptr, part2node, num_parts,
                                                                      num_nodes, dim, ldx, lpr, eps, flags);
    else
        aggregate_small_kernel<1><<<(unsigned)blocks, 128, 0, stream>>>(X, out, col_idx, degrees, part_ptr, part2node, num_parts,
                                                                      num_nodes, dim, ldx, lpr, eps, flags);
    GNNA_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return GNNA_OK;
}

}  // namespace gnna
