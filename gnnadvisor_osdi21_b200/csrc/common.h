// Shared helpers of libgnna_b200.so (error text, launch counter, argument checks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/gnna_b200.h"

namespace gnna {

char *error_buffer();                       // thread-local, 512 bytes
int fail(int code, const char *fmt, ...);   // formats into error_buffer(), returns code
void count_launch(int n = 1);               // thread-local launch counter

#define GNNA_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::gnna::fail(GNNA_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                       \
    } while (0)

#define GNNA_REQUIRE(cond, ...)                                       \
    do {                                                              \
        if (!(cond)) return ::gnna::fail(GNNA_ERR_INVALID, __VA_ARGS__); \
    } while (0)

enum Mode { MODE_SAG = 0, MODE_GCN = 1, MODE_GIN = 2, MODE_GCN_PRESCALED = 3 };

// Gate of the fused exchange+aggregation kernel of the sharded path (aggregate.cu: aggregate_kernel, halo.cu): the group
// table is the concatenation of one segment per OWNER of the neighbours; a CTA whose groups lie in a peer's segment waits
// (acquire, system scope, bounded) until that peer's rows of the current step have landed in this rank's buffer.
// nseg == 0: no gate (every single-GPU launch).
constexpr int GATE_MAX_SEGS = 16;
struct GateParams {
    long long bounds[GATE_MAX_SEGS + 1];   // segment s = groups [bounds[s], bounds[s+1])
    int peer[GATE_MAX_SEGS];               // rank whose flag gates segment s; -1: this rank's own rows, no wait
    int nseg;
    const unsigned *flags;                 // flags[q] >= step  <=>  rank q's rows of this step have arrived
    const unsigned *step_ptr;
    unsigned *error_word;
};

// One aggregation launch (aggregate.cu).  elem: 4 = fp32, 2 = bf16.
int aggregate(int mode, int elem_bytes, const void *X, void *out,
              const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
              const int32_t *part_ptr, const int32_t *part2node,
              int64_t num_nodes, int dim, int64_t num_parts,
              int part_size, int dim_worker, int warp_per_block, cudaStream_t stream,
              int ldx = 0, int64_t num_rows_x = 0, bool accumulate = false, const GateParams *gate = nullptr);

// persistent TMA-staged variant (aggregate_staged.cu); GNNA_ERR_UNSUPPORTED when the shape has no such variant
int aggregate_staged(const float *X, float *out, const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                     const int32_t *part_ptr, const int32_t *part2node, long long num_parts, long long num_edges,
                     int dim, int part_size, float scale, int flags, cudaStream_t stream);

// run-based software-pipelined variant (aggregate_runs.cu); `out` already zeroed; flags: 1 = scale, 2 = degrees[node];
// GNNA_ERR_UNSUPPORTED when switched off (gnna_set_runs(0), the default) or the shape has no such variant
int aggregate_runs(int elem_bytes, const void *X, float *out, const int32_t *row_ptr, const int32_t *col_idx,
                   const float *degrees, const int32_t *part_ptr, const int32_t *part2node, long long num_nodes,
                   long long num_parts, int dim, int ldx, float scale, int flags, cudaStream_t stream);
int runs_mode();

// single-launch row-owned variant for launch-bound graphs (aggregate_small.cu); GNNA_ERR_UNSUPPORTED when the shape does not fit
int aggregate_small(int mode, const float *X, float *out, const int32_t *col_idx, const float *degrees, float eps,
                    const int32_t *part_ptr, const int32_t *part2node, long long num_nodes, long long num_parts, int dim, int ldx,
                    bool exact_gcn, cudaStream_t stream);

bool gcn_exact_mode();   // GNNA_GCN_EXACT / gnna_set_gcn_exact

// Tall-skinny dense products on the tensor cores (gemm_tf32x3.cu: tcgen05 kind::tf32, 3xTF32 split, fp32-grade accuracy).
// Row-major C[m,n] = op(A) * B; GNNA_ERR_UNSUPPORTED when the shape is not one of the two contractions it is built for (the
// caller then uses cuBLAS).  row_scale (non-transposed A only, may be null): C[i,:] *= row_scale[i] in the epilogue.
int gemm_tf32x3(cudaStream_t st, bool ta, bool tb, int64_t m, int64_t n, int64_t k, const float *A, const float *B, float *C,
                const float *row_scale);
bool tc_gemm_enabled();  // GNNA_TC_GEMM / gnna_set_tc_gemm (default on)

// Xs[i,:] = degrees[i] * X[i,:]  (X == Xs allowed)
int prescale_rows(const float *X, float *Xs, const float *degrees, int64_t num_nodes, int dim, cudaStream_t stream);

// Xb[i, 0:dim] = bf16(degrees[i] * X[i, :]) (degrees may be NULL: plain conversion), zero-padded to ldb (% 8 == 0) columns
int scale_rows_bf16(const float *X, void *Xb, const float *degrees, int64_t num_rows, int dim, int ldb, cudaStream_t stream);

}  // namespace gnna
