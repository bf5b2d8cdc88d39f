// ops.cu -- C-ABI entry points: error plumbing, the three aggregation calls and the four layer
// operators (aggregation + the dense products the reference does with torch::mm).
//
// Reference host launchers replaced here (GNNAdvisor/GNNConv/GNNAdvisor_kernel.cu):
//   SAG_cuda :110-184, spmm_forward_cuda :267-322, spmm_backward_cuda :422-476,
//   spmm_forward_cuda_gin :559-617, spmm_backward_cuda_gin :696-747.
// The dense products stay library SGEMMs (cuBLAS, fp32, TF32 off) exactly as torch::mm is in the
// reference; every one of them has min(din, dout) <= 64 at fp32 in the reference's configurations
// and is bandwidth-bound, so the aggregation kernel is where the time goes.
#include <cublas_v2.h>

#include <mutex>
#include <string.h>

#include "common.h"

namespace gnna {

static thread_local char g_err[512];
static thread_local long long g_launches = 0;

char *error_buffer() { return g_err; }

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches += n; }

// one cuBLAS handle per device, created on first use
static std::mutex g_handle_mutex;
static cublasHandle_t g_handles[64] = {nullptr};

static int get_handle(cublasHandle_t *h)
{
    int dev = 0;
    GNNA_CUDA_CHECK(cudaGetDevice(&dev));
    GNNA_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_handle_mutex);
    if (!g_handles[dev]) {
        cublasStatus_t s = cublasCreate(&g_handles[dev]);
        if (s != CUBLAS_STATUS_SUCCESS) {
            g_handles[dev] = nullptr;
            return fail(GNNA_ERR_CUBLAS, "cublasCreate failed (%d)", (int)s);
        }
        cublasSetMathMode(g_handles[dev], CUBLAS_DEFAULT_MATH);    // plain fp32 SGEMM, TF32 stays off (as torch::mm)
        // a workspace of our own: without one cuBLAS allocates stream-ordered memory behind the call, which a stream
        // capture (main.py --cuda_graph, the multi-GPU step graphs) must not depend on
        void *ws = nullptr;
        const size_t ws_bytes = 32u << 20;
        if (cudaMalloc(&ws, ws_bytes) == cudaSuccess) cublasSetWorkspace(g_handles[dev], ws, ws_bytes);
        else (void)cudaGetLastError();
    }
    *h = g_handles[dev];
    return GNNA_OK;
}

// C[m,n] = op(A) * op(B), all row-major.  Row-major C = A*B is column-major C^T = B^T * A^T.
// row_scale / scaled: the caller would like C[i,:] *= row_scale[i]; *scaled says whether this call did it (only the
// tensor-core path has that epilogue, otherwise the caller runs its pre-scale pass).
static int sgemm_rm(cudaStream_t st, bool ta, bool tb, int64_t m, int64_t n, int64_t k,
                    const float *A, const float *B, float *C, const float *row_scale = nullptr, bool *scaled = nullptr)
{
    if (scaled) *scaled = false;
    if (m == 0 || n == 0) return GNNA_OK;
    if (tc_gemm_enabled() && k > 0) {      // the two tall-skinny contractions of a layer: tcgen05, 3xTF32 (gemm_tf32x3.cu)
        const int rc = gemm_tf32x3(st, ta, tb, m, n, k, A, B, C, (!ta && scaled) ? row_scale : nullptr);
        if (rc != GNNA_ERR_UNSUPPORTED) {
            if (rc == GNNA_OK && scaled && row_scale && !ta) *scaled = true;
            return rc;
        }
    }
    cublasHandle_t h;
    int rc = get_handle(&h);
    if (rc != GNNA_OK) return rc;
    cublasSetStream(h, st);
    const float one = 1.f, zero = 0.f;
    const int64_t lda = ta ? m : k, ldb = tb ? k : n;
    if (k == 0) {
        GNNA_CUDA_CHECK(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)m * (size_t)n, st));
        return GNNA_OK;
    }
    cublasStatus_t s = cublasSgemm_64(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N,
                                      n, m, k, &one, B, ldb, A, lda, &zero, C, n);
    if (s != CUBLAS_STATUS_SUCCESS) return fail(GNNA_ERR_CUBLAS, "cublasSgemm failed (%d)", (int)s);
    return GNNA_OK;
}

}  // namespace gnna

using namespace gnna;

extern "C" int gnna_abi_version(void) { return 1; }
extern "C" const char *gnna_last_error(void) { return error_buffer(); }

extern "C" int64_t gnna_launch_count(int reset)
{
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" int gnna_sag_f32(const float *X, float *out, const int32_t *row_ptr, const int32_t *col_idx,
                            const int32_t *part_ptr, const int32_t *part2node,
                            int64_t num_nodes, int dim, int64_t num_parts,
                            int part_size, int dim_worker, int warp_per_block, void *stream)
{
    return aggregate(MODE_SAG, 4, X, out, row_ptr, col_idx, nullptr, 1.f, part_ptr, part2node, num_nodes, dim,
                     num_parts, part_size, dim_worker, warp_per_block, (cudaStream_t)stream);
}

extern "C" int gnna_gcn_aggregate_f32(const float *X, float *out, const int32_t *row_ptr, const int32_t *col_idx,
                                      const float *degrees, const int32_t *part_ptr, const int32_t *part2node,
                                      int64_t num_nodes, int dim, int64_t num_parts,
                                      int part_size, int dim_worker, int warp_per_block, void *stream)
{
    return aggregate(MODE_GCN, 4, X, out, row_ptr, col_idx, degrees, 1.f, part_ptr, part2node, num_nodes, dim,
                     num_parts, part_size, dim_worker, warp_per_block, (cudaStream_t)stream);
}

extern "C" int gnna_gin_aggregate_f32(const float *X, float *out, const int32_t *row_ptr, const int32_t *col_idx,
                                      float eps, const int32_t *part_ptr, const int32_t *part2node,
                                      int64_t num_nodes, int dim, int64_t num_parts,
                                      int part_size, int dim_worker, int warp_per_block, void *stream)
{
    return aggregate(MODE_GIN, 4, X, out, row_ptr, col_idx, nullptr, eps, part_ptr, part2node, num_nodes, dim,
                     num_parts, part_size, dim_worker, warp_per_block, (cudaStream_t)stream);
}

extern "C" int gnna_aggregate_bf16(int mode, const void *X_bf16, float *out_f32, const int32_t *row_ptr,
                                   const int32_t *col_idx, const float *degrees, float eps,
                                   const int32_t *part_ptr, const int32_t *part2node,
                                   int64_t num_nodes, int dim, int64_t num_parts,
                                   int part_size, int dim_worker, int warp_per_block, void *stream)
{
    return aggregate(mode, 2, X_bf16, out_f32, row_ptr, col_idx, degrees, eps, part_ptr, part2node, num_nodes, dim,
                     num_parts, part_size, dim_worker, warp_per_block, (cudaStream_t)stream);
}

extern "C" int gnna_aggregate_f32_ex(int mode, const float *X, int64_t num_src_rows, float *out, int64_t num_dst_rows,
                                     const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
                                     const int32_t *part_ptr, const int32_t *part2node, int dim, int64_t num_parts,
                                     int part_size, int dim_worker, int warp_per_block, void *stream)
{
    GNNA_REQUIRE(mode >= MODE_SAG && mode <= MODE_GIN, "gnna_aggregate_f32_ex: bad mode %d", mode);
    GNNA_REQUIRE(num_src_rows >= num_dst_rows, "gnna_aggregate_f32_ex: fewer source rows than destination rows");
    return aggregate(mode, 4, X, out, row_ptr, col_idx, degrees, eps, part_ptr, part2node, num_dst_rows, dim,
                     num_parts, part_size, dim_worker, warp_per_block, (cudaStream_t)stream, dim, num_src_rows);
}

extern "C" int gnna_aggregate_part_f32_ex(int mode, int accumulate, const float *X, int64_t num_src_rows, float *out,
                                          int64_t num_dst_rows, const int32_t *row_ptr, const int32_t *col_idx,
                                          const float *degrees, float eps, const int32_t *part_ptr,
                                          const int32_t *part2node, int dim, int64_t num_parts,
                                          int part_size, int dim_worker, int warp_per_block, void *stream)
{
    GNNA_REQUIRE(mode == MODE_SAG || mode == MODE_GIN || mode == MODE_GCN_PRESCALED,
                 "gnna_aggregate_part_f32_ex: mode %d not supported (0, 2, 3)", mode);
    GNNA_REQUIRE(dim % 4 == 0 && (((uintptr_t)X | (uintptr_t)out) & 15) == 0, "gnna_aggregate_part_f32_ex: dim %% 4 != 0 or unaligned");
    return aggregate(mode, 4, X, out, row_ptr, col_idx, degrees, eps, part_ptr, part2node, num_dst_rows, dim,
                     num_parts, part_size, dim_worker, warp_per_block, (cudaStream_t)stream, dim, num_src_rows,
                     accumulate != 0);
}

// The sharded step with the halo exchange FUSED into the aggregation (dist.ShardedGraph.aggregate_overlapped): ONE launch
// over the concatenation of the per-owner sub-CSRs of a rank; the CTAs of a peer's segment wait inside the kernel for that
// peer's flag in the halo control block (csrc/halo.cu) while the CTAs ahead of them aggregate what has already landed.
// out is zero-filled, every group is merged with reductions.  mode 0 SAG, 2 GIN, 3 GCN on pre-scaled rows; dim % 4 == 0.
extern "C" int gnna_aggregate_gated_f32(int mode, const float *X, int64_t num_src_rows, float *out, int64_t num_dst_rows,
                                        const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
                                        const int32_t *part_ptr, const int32_t *part2node, int dim, int64_t num_parts,
                                        const int64_t *seg_bounds_host, const int32_t *seg_peer_host, int num_segs,
                                        void *my_ctrl, int part_size, int dim_worker, int warp_per_block, void *stream)
{
    GNNA_REQUIRE(mode == MODE_SAG || mode == MODE_GIN || mode == MODE_GCN_PRESCALED,
                 "gnna_aggregate_gated_f32: mode %d not supported (0, 2, 3)", mode);
    GNNA_REQUIRE(dim % 4 == 0 && (((uintptr_t)X | (uintptr_t)out) & 15) == 0, "gnna_aggregate_gated_f32: dim %% 4 != 0 or unaligned");
    GNNA_REQUIRE(num_segs >= 1 && num_segs <= GATE_MAX_SEGS && seg_bounds_host && seg_peer_host && my_ctrl,
                 "gnna_aggregate_gated_f32: bad segment description");
    GNNA_REQUIRE(seg_bounds_host[0] == 0 && seg_bounds_host[num_segs] == num_parts, "gnna_aggregate_gated_f32: segments do not cover the table");
    GateParams gp;
    memset(&gp, 0, sizeof(gp));
    for (int s = 0; s < num_segs; s++) {
        GNNA_REQUIRE(seg_bounds_host[s] <= seg_bounds_host[s + 1], "gnna_aggregate_gated_f32: segment bounds not ascending");
        gp.bounds[s] = seg_bounds_host[s];
        gp.peer[s] = seg_peer_host[s];
    }
    gp.bounds[num_segs] = seg_bounds_host[num_segs];
    gp.nseg = num_segs;
    unsigned *ctrl = (unsigned *)my_ctrl;          // layout: halo.cu (flags [0..15], error word [48], step [49])
    gp.flags = ctrl;
    gp.error_word = ctrl + 48;
    gp.step_ptr = ctrl + 49;
    GNNA_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)num_dst_rows * (size_t)dim, (cudaStream_t)stream));
    if (num_parts == 0) return GNNA_OK;
    return aggregate(mode, 4, X, out, row_ptr, col_idx, degrees, eps, part_ptr, part2node, num_dst_rows, dim,
                     num_parts, part_size, dim_worker, warp_per_block, (cudaStream_t)stream, dim, num_src_rows,
                     true, &gp);
}

extern "C" int gnna_prescale_rows_f32(const float *X, float *Xs, const float *degrees, int64_t num_rows, int dim, void *stream)
{
    GNNA_REQUIRE(num_rows >= 0 && dim >= 0, "gnna_prescale_rows_f32: negative size");
    GNNA_REQUIRE(num_rows == 0 || dim == 0 || (X && Xs && degrees), "gnna_prescale_rows_f32: null pointer");
    return prescale_rows(X, Xs, degrees, num_rows, dim, (cudaStream_t)stream);
}

// Row-major C[m,n] = op(A) * op(B) on the library's cuBLAS handle (fp32, TF32 off): the dense products of the layer
// operators for callers that compose a layer themselves (the sharded layers do: product, halo exchange, aggregation).
extern "C" int gnna_sgemm_f32(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k,
                              const float *A, const float *B, float *C, void *stream)
{
    GNNA_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gnna_sgemm_f32: negative size");
    GNNA_REQUIRE(m == 0 || n == 0 || (C && (k == 0 || (A && B))), "gnna_sgemm_f32: null pointer");
    return sgemm_rm((cudaStream_t)stream, trans_a != 0, trans_b != 0, m, n, k, A, B, C);
}

#define GNNA_TRY(expr)             \
    do {                           \
        int _rc = (expr);          \
        if (_rc != GNNA_OK) return _rc; \
    } while (0)

extern "C" int gnna_forward_f32(const float *X, const float *W, float *T_ws, float *out,
                                const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                                const int32_t *part_ptr, const int32_t *part2node,
                                int64_t num_nodes, int din, int dout, int64_t num_parts,
                                int part_size, int dim_worker, int warp_per_block, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    GNNA_REQUIRE(X && W && T_ws && out, "gnna_forward_f32: null pointer");
    GNNA_REQUIRE(degrees, "gnna_forward_f32: null degrees");
    const bool prescale = !gcn_exact_mode();
    bool scaled = false;
    GNNA_TRY(sgemm_rm(st, false, false, num_nodes, dout, din, X, W, T_ws, prescale ? degrees : nullptr, &scaled));   // kernel.cu:280
    int mode = MODE_GCN;
    if (prescale) {            // T is our own scratch: rows scaled by n_j in the product's epilogue, else in place afterwards
        if (!scaled) GNNA_TRY(prescale_rows(T_ws, T_ws, degrees, num_nodes, dout, st));
        mode = MODE_GCN_PRESCALED;
    }
    return aggregate(mode, 4, T_ws, out, row_ptr, col_idx, degrees, 1.f, part_ptr, part2node, num_nodes, dout,
                     num_parts, part_size, dim_worker, warp_per_block, st);               // kernel.cu:282-313
}

extern "C" int gnna_backward_f32(const float *d_out, const float *X, const float *W, float *G_ws,
                                 float *d_input, float *d_weight,
                                 const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                                 const int32_t *part_ptr, const int32_t *part2node,
                                 int64_t num_nodes, int din, int dout, int64_t num_parts,
                                 int part_size, int dim_worker, int warp_per_block, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    GNNA_REQUIRE(d_out && X && W && G_ws && d_weight, "gnna_backward_f32: null pointer");
    GNNA_TRY(aggregate(MODE_GCN, 4, d_out, G_ws, row_ptr, col_idx, degrees, 1.f, part_ptr, part2node, num_nodes, dout,
                       num_parts, part_size, dim_worker, warp_per_block, st));            // kernel.cu:436-463
    if (d_input) GNNA_TRY(sgemm_rm(st, false, true, num_nodes, din, dout, G_ws, W, d_input));   // :472
    return sgemm_rm(st, true, false, din, dout, num_nodes, X, G_ws, d_weight);            // :473
}

// ---- mixed-precision GCN layer: fp32 dense products, bf16 GATHERED rows, fp32 accumulation and outputs --------
// (BASELINE.json config "Reddit GCN 2-layer D=64 bf16"; no reference counterpart, the reference is fp32-only)
static int round_up8(int x) { return (x + 7) / 8 * 8; }

extern "C" int gnna_scale_rows_bf16(const float *X, void *Xb, const float *degrees, int64_t num_rows, int dim, int ldb,
                                    void *stream)
{
    GNNA_REQUIRE(num_rows >= 0 && dim >= 0, "gnna_scale_rows_bf16: negative size");
    GNNA_REQUIRE(num_rows == 0 || dim == 0 || (X && Xb), "gnna_scale_rows_bf16: null pointer");
    return scale_rows_bf16(X, Xb, degrees, num_rows, dim, ldb, (cudaStream_t)stream);
}

extern "C" int gnna_aggregate_bf16_ex(int mode, const void *X_bf16, int ldx, float *out_f32, const int32_t *row_ptr,
                                      const int32_t *col_idx, const float *degrees, float eps,
                                      const int32_t *part_ptr, const int32_t *part2node,
                                      int64_t num_nodes, int dim, int64_t num_parts,
                                      int part_size, int dim_worker, int warp_per_block, void *stream)
{
    GNNA_REQUIRE(ldx >= dim, "gnna_aggregate_bf16_ex: ldx %d < dim %d", ldx, dim);
    return aggregate(mode, 2, X_bf16, out_f32, row_ptr, col_idx, degrees, eps, part_ptr, part2node, num_nodes, dim,
                     num_parts, part_size, dim_worker, warp_per_block, (cudaStream_t)stream, ldx);
}

extern "C" int gnna_forward_mixed(const float *X, const float *W, float *T_ws, void *Tb_ws, float *out,
                                  const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                                  const int32_t *part_ptr, const int32_t *part2node,
                                  int64_t num_nodes, int din, int dout, int64_t num_parts,
                                  int part_size, int dim_worker, int warp_per_block, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    GNNA_REQUIRE(X && W && T_ws && Tb_ws && out && degrees, "gnna_forward_mixed: null pointer");
    const int ld = round_up8(dout);
    GNNA_TRY(sgemm_rm(st, false, false, num_nodes, dout, din, X, W, T_ws));               // kernel.cu:280
    GNNA_TRY(scale_rows_bf16(T_ws, Tb_ws, degrees, num_nodes, dout, ld, st));             // Tb_j = bf16(n_j * T_j)
    return aggregate(MODE_GCN_PRESCALED, 2, Tb_ws, out, row_ptr, col_idx, degrees, 1.f, part_ptr, part2node, num_nodes,
                     dout, num_parts, part_size, dim_worker, warp_per_block, st, ld);     // out_i = n_i * sum_j Tb_j
}

extern "C" int gnna_backward_mixed(const float *d_out, const float *X, const float *W, void *Gb_ws, float *G_ws,
                                   float *d_input, float *d_weight,
                                   const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                                   const int32_t *part_ptr, const int32_t *part2node,
                                   int64_t num_nodes, int din, int dout, int64_t num_parts,
                                   int part_size, int dim_worker, int warp_per_block, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    GNNA_REQUIRE(d_out && X && W && Gb_ws && G_ws && d_weight && degrees, "gnna_backward_mixed: null pointer");
    const int ld = round_up8(dout);
    GNNA_TRY(scale_rows_bf16(d_out, Gb_ws, degrees, num_nodes, dout, ld, st));
    GNNA_TRY(aggregate(MODE_GCN_PRESCALED, 2, Gb_ws, G_ws, row_ptr, col_idx, degrees, 1.f, part_ptr, part2node, num_nodes,
                       dout, num_parts, part_size, dim_worker, warp_per_block, st, ld)); // kernel.cu:436-463
    if (d_input) GNNA_TRY(sgemm_rm(st, false, true, num_nodes, din, dout, G_ws, W, d_input));   // :472
    return sgemm_rm(st, true, false, din, dout, num_nodes, X, G_ws, d_weight);            // :473
}

// mixed-precision GIN layer: the gathered matrices (X forward, Pm = dOut*W^T backward) travel as bf16
extern "C" int gnna_forward_gin_mixed(const float *X, const float *W, float eps, void *Xb_ws, float *out, float *x_agg,
                                      const int32_t *row_ptr, const int32_t *col_idx,
                                      const int32_t *part_ptr, const int32_t *part2node,
                                      int64_t num_nodes, int din, int dout, int64_t num_parts,
                                      int part_size, int dim_worker, int warp_per_block, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    GNNA_REQUIRE(X && W && Xb_ws && out && x_agg, "gnna_forward_gin_mixed: null pointer");
    const int ld = round_up8(din);
    GNNA_TRY(scale_rows_bf16(X, Xb_ws, nullptr, num_nodes, din, ld, st));
    GNNA_TRY(aggregate(MODE_GIN, 2, Xb_ws, x_agg, row_ptr, col_idx, nullptr, eps, part_ptr, part2node, num_nodes, din,
                       num_parts, part_size, dim_worker, warp_per_block, st, ld));        // kernel.cu:572-603
    return sgemm_rm(st, false, false, num_nodes, dout, din, x_agg, W, out);               // :605
}

extern "C" int gnna_backward_gin_mixed(const float *d_out, const float *x_agg, const float *W, float eps,
                                       float *Pm_ws, void *Pmb_ws, float *d_input, float *d_weight,
                                       const int32_t *row_ptr, const int32_t *col_idx,
                                       const int32_t *part_ptr, const int32_t *part2node,
                                       int64_t num_nodes, int din, int dout, int64_t num_parts,
                                       int part_size, int dim_worker, int warp_per_block, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    GNNA_REQUIRE(d_out && x_agg && W && d_weight && (!d_input || (Pm_ws && Pmb_ws)), "gnna_backward_gin_mixed: null pointer");
    GNNA_TRY(sgemm_rm(st, true, false, din, dout, num_nodes, x_agg, d_out, d_weight));    // kernel.cu:710
    if (!d_input) return GNNA_OK;
    const int ld = round_up8(din);
    GNNA_TRY(sgemm_rm(st, false, true, num_nodes, din, dout, d_out, W, Pm_ws));           // :711
    GNNA_TRY(scale_rows_bf16(Pm_ws, Pmb_ws, nullptr, num_nodes, din, ld, st));
    return aggregate(MODE_GIN, 2, Pmb_ws, d_input, row_ptr, col_idx, nullptr, eps, part_ptr, part2node, num_nodes, din,
                     num_parts, part_size, dim_worker, warp_per_block, st, ld);           // :712-738
}

extern "C" int gnna_forward_gin_f32(const float *X, const float *W, float eps, float *out, float *x_agg,
                                    const int32_t *row_ptr, const int32_t *col_idx,
                                    const int32_t *part_ptr, const int32_t *part2node,
                                    int64_t num_nodes, int din, int dout, int64_t num_parts,
                                    int part_size, int dim_worker, int warp_per_block, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    GNNA_REQUIRE(X && W && out && x_agg, "gnna_forward_gin_f32: null pointer");
    GNNA_TRY(aggregate(MODE_GIN, 4, X, x_agg, row_ptr, col_idx, nullptr, eps, part_ptr, part2node, num_nodes, din,
                       num_parts, part_size, dim_worker, warp_per_block, st));            // kernel.cu:572-603
    return sgemm_rm(st, false, false, num_nodes, dout, din, x_agg, W, out);               // :605
}

extern "C" int gnna_backward_gin_f32(const float *d_out, const float *x_agg, const float *W, float eps,
                                     float *Pm_ws, float *d_input, float *d_weight,
                                     const int32_t *row_ptr, const int32_t *col_idx,
                                     const int32_t *part_ptr, const int32_t *part2node,
                                     int64_t num_nodes, int din, int dout, int64_t num_parts,
                                     int part_size, int dim_worker, int warp_per_block, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    GNNA_REQUIRE(d_out && x_agg && W && d_weight && (!d_input || Pm_ws), "gnna_backward_gin_f32: null pointer");
    GNNA_TRY(sgemm_rm(st, true, false, din, dout, num_nodes, x_agg, d_out, d_weight));    // kernel.cu:710
    if (!d_input) return GNNA_OK;                                                          // caller needs no input gradient
    GNNA_TRY(sgemm_rm(st, false, true, num_nodes, din, dout, d_out, W, Pm_ws));           // :711
    return aggregate(MODE_GIN, 4, Pm_ws, d_input, row_ptr, col_idx, nullptr, eps, part_ptr, part2node, num_nodes, din,
                     num_parts, part_size, dim_worker, warp_per_block, st);               // :712-738
}
