/*
 * gnna_oracle.c -- CPU restatement of the GNNAdvisor hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker*: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (gnnadvisor_osdi21_b200/) never links, imports or calls anything in oracle/.
 *
 * Every function restates, in plain C, what one reference function computes and cites
 * the reference lines it follows (paths relative to /root/reference/GNNAdvisor/GNNConv/).
 * Float arithmetic is done in fp32 with un-fused multiply/add (build with
 * -ffp-contract=off) because the reference kernels use __fmaf_rn(a, b, 0), i.e. a
 * separately rounded product (GNNAdvisor_kernel.cu:389,403,530,542).
 *
 * Parity pinning (see oracle/README.md):
 *   - build_part: pinned bit-exactly against the reference's own build_part compiled
 *     from /root/reference (oracle/_ref, tests/golden/build_part_*.npz).
 *   - float kernels: the reference ships no golden vectors; pinned against
 *     outputs of the reference CUDA kernels themselves run on a B200
 *     (tests/golden/refgpu_*.npz, produced by oracle/make_golden_refgpu.py).
 *
 * Summation order: inside one neighbour-group the reference adds neighbours serially in
 * CSR order (kernel.cu:383-406); groups of one node are merged with atomics in arbitrary
 * order (kernel.cu:409-413).  The oracle merges groups in ascending group order, which is
 * one of the orders the reference can produce.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* build_part  (GNNAdvisor.cpp:210-251)                                                 */
/* ------------------------------------------------------------------------------------ */

/* Pass 1, GNNAdvisor.cpp:219-227: number of neighbour-groups. */
API int64_t oracle_count_parts(int part_size, const int32_t *indptr, int64_t num_nodes)
{
    int64_t num_parts = 0;
    for (int64_t i = 0; i < num_nodes; i++) {
        int degree = indptr[i + 1] - indptr[i];
        int this_parts = (degree % part_size == 0) ? degree / part_size : degree / part_size + 1;
        num_parts += this_parts;
    }
    return num_parts;
}

/*
 * Pass 2, GNNAdvisor.cpp:229-249, INCLUDING both defects of the reference:
 *   F5  the tables are float32 tensors (torch::zeros(n), :229-230), so offsets above 2^24
 *       are rounded;
 *   F6  the terminal partPtr[P] is only written when the LAST node has >= 1 neighbour
 *       (:246-247); otherwise it keeps its zero initialisation.
 * part_ptr has num_parts+1 entries, part2node num_parts entries.
 */
API void oracle_build_part_f32(int part_size, const int32_t *indptr, int64_t num_nodes,
                               float *part_ptr, float *part2node, int64_t num_parts)
{
    memset(part_ptr, 0, sizeof(float) * (size_t)(num_parts + 1));
    memset(part2node, 0, sizeof(float) * (size_t)num_parts);
    int64_t c = 0;
    for (int64_t i = 0; i < num_nodes; i++) {
        int degree = indptr[i + 1] - indptr[i];
        int this_parts = (degree % part_size == 0) ? degree / part_size : degree / part_size + 1;
        for (int pid = 0; pid < this_parts; pid++) {
            int part_beg = indptr[i] + pid * part_size;
            int part_end = part_beg + part_size < indptr[i + 1] ? part_beg + part_size : indptr[i + 1];
            part_ptr[c] = (float)part_beg;          /* :244 */
            part2node[c++] = (float)i;              /* :245 */
            if (i == num_nodes - 1 && part_end == indptr[i + 1])
                part_ptr[c] = (float)part_end;      /* :246-247 */
        }
    }
}

/* The caller's cast `partPtr.int()` (GNNA_main.py:109-110) applied to the float tables. */
API void oracle_build_part_i32_compat(int part_size, const int32_t *indptr, int64_t num_nodes,
                                      int32_t *part_ptr, int32_t *part2node, int64_t num_parts)
{
    float *fp = (float *)malloc(sizeof(float) * (size_t)(num_parts + 1));
    float *fn = (float *)malloc(sizeof(float) * (size_t)(num_parts ? num_parts : 1));
    oracle_build_part_f32(part_size, indptr, num_nodes, fp, fn, num_parts);
    for (int64_t k = 0; k <= num_parts; k++) part_ptr[k] = (int32_t)fp[k];
    for (int64_t k = 0; k < num_parts; k++) part2node[k] = (int32_t)fn[k];
    free(fp);
    free(fn);
}

/* Integer-exact table: same enumeration, int32 storage, terminal always = indptr[N]
 * (what the reference intends; differs from it only where F5/F6 corrupt the table). */
API void oracle_build_part_i32_exact(int part_size, const int32_t *indptr, int64_t num_nodes,
                                     int32_t *part_ptr, int32_t *part2node, int64_t num_parts)
{
    int64_t c = 0;
    for (int64_t i = 0; i < num_nodes; i++) {
        int degree = indptr[i + 1] - indptr[i];
        int this_parts = (degree + part_size - 1) / part_size;
        for (int pid = 0; pid < this_parts; pid++) {
            part_ptr[c] = indptr[i] + pid * part_size;
            part2node[c++] = (int32_t)i;
        }
    }
    part_ptr[num_parts] = indptr[num_nodes];
}

/* degrees = sqrt(max(deg, 1)) in float32  (GNNAdvisor/dataset.py:11-18,121-122) */
API void oracle_degrees(const int32_t *indptr, int64_t num_nodes, float *degrees)
{
    for (int64_t i = 0; i < num_nodes; i++) {
        int d = indptr[i + 1] - indptr[i];
        degrees[i] = sqrtf((float)(d > 0 ? d : 1));
    }
}

/* ------------------------------------------------------------------------------------ */
/* Aggregation kernels, one "warp" == one neighbour-group                               */
/* ------------------------------------------------------------------------------------ */

enum { MODE_SAG = 0, MODE_GCN = 1, MODE_GIN = 2 };

/*
 * One group, exactly the per-warp body of the reference kernels:
 *   SAG  kernel.cu:212-258   partial[d] += X[nid][d]
 *   GCN  kernel.cu:350-414 (fwd) == :504-551 (bwd)
 *        w = fl(n_src * n_nid) ; partial[d] += fl(w * X[nid][d])
 *   GIN  kernel.cu:646-688 (fwd) == :775-813 (bwd)   partial[d] += X[nid][d] ; eps at write-back
 * then out[src][d] += partial[d]  (atomicAdd_F, :12-17) -- eps*partial for GIN (:686,:811).
 * A group with part_end <= part_beg does nothing at all (the loop at :383 never runs and the
 * write-back adds whatever is in shared memory; with zero iterations the reference adds
 * uninitialised smem -- see oracle/README.md "empty groups": no table produced by build_part
 * on a graph whose last node is non-isolated contains one, and for the F6 table the CUDA
 * product treats it as "adds nothing", which is what the oracle does).
 */
/* Columns are processed in blocks of GB so that the running sums of a block live in registers (the timed CPU baseline
 * spends its time here).  Every output element is still its own serial chain over the group's neighbours in CSR order --
 * partial[d] = fl(partial[d] + fl(w * x[d])) -- so blocking over d changes no bit of the result. */
#define GB 64
#define PF 6                                 /* neighbours ahead whose rows are requested (a hint: no effect on results) */
static inline void prefetch_row(const float *row, int64_t n)
{
    for (int64_t d = 0; d < n; d += 16) __builtin_prefetch(row + d, 0, 3);
}
static void group_body(int mode, float *restrict out, const float *restrict X, const int32_t *restrict col_idx,
                       const float *restrict degrees, float eps, int64_t dim,
                       int32_t src, int64_t beg, int64_t end, float *restrict partial)
{
    if (end <= beg) return;
    const float src_norm = (mode == MODE_GCN) ? degrees[src] : 1.0f;
    for (int64_t d0 = 0; d0 < dim; d0 += GB) {
        const int64_t dn = dim - d0 < GB ? dim - d0 : GB;
        float acc[GB];
        for (int64_t d = 0; d < GB; d++) acc[d] = 0.0f;
        if (mode == MODE_GCN) {
            for (int64_t k = beg; k < end; k++) {
                const int32_t nid = col_idx[k];
                const float w = src_norm * degrees[nid];
                const float *row = X + (int64_t)nid * dim + d0;
                if (k + PF < end) prefetch_row(X + (int64_t)col_idx[k + PF] * dim + d0, dn);
                if (dn == GB) {
                    for (int64_t d = 0; d < GB; d++) {
                        float prod = w * row[d];
                        acc[d] = acc[d] + prod;
                    }
                } else {
                    for (int64_t d = 0; d < dn; d++) {
                        float prod = w * row[d];
                        acc[d] = acc[d] + prod;
                    }
                }
            }
        } else {
            for (int64_t k = beg; k < end; k++) {
                const float *row = X + (int64_t)col_idx[k] * dim + d0;
                if (k + PF < end) prefetch_row(X + (int64_t)col_idx[k + PF] * dim + d0, dn);
                if (dn == GB) {
                    for (int64_t d = 0; d < GB; d++) acc[d] = acc[d] + row[d];
                } else {
                    for (int64_t d = 0; d < dn; d++) acc[d] = acc[d] + row[d];
                }
            }
        }
        for (int64_t d = 0; d < dn; d++) partial[d0 + d] = acc[d];
    }
    float *orow = out + (int64_t)src * dim;
    if (mode == MODE_GIN)
        for (int64_t d = 0; d < dim; d++) orow[d] = orow[d] + eps * partial[d];
    else
        for (int64_t d = 0; d < dim; d++) orow[d] = orow[d] + partial[d];
}

/*
 * Whole aggregation, literal form: zero the output (torch::zeros[_like], kernel.cu:121,
 * 282,436,572,712) then run every group in ascending group order on one thread.
 */
API void oracle_aggregate(int mode, const float *X, float *out,
                          const int32_t *col_idx, const float *degrees, float eps,
                          const int32_t *part_ptr, const int32_t *part2node,
                          int64_t num_nodes, int64_t dim, int64_t num_parts)
{
    memset(out, 0, sizeof(float) * (size_t)(num_nodes * dim));
    float *partial = (float *)malloc(sizeof(float) * (size_t)(dim ? dim : 1));
    for (int64_t w = 0; w < num_parts; w++)
        group_body(mode, out, X, col_idx, degrees, eps, dim,
                   part2node[w], part_ptr[w], part_ptr[w + 1], partial);
    free(partial);
}

/*
 * Same result bit for bit, on all host threads: part2node is non-decreasing (build_part
 * emits groups in node order), so the groups are cut into contiguous runs at node
 * boundaries and each thread owns whole nodes; inside a node the groups are still merged
 * in ascending order.  Used for the timed CPU baseline (bench.py cpu_baseline /
 * --impl reference).  Returns 0, or -1 if part2node is not sorted (then nothing is done).
 */
API int oracle_aggregate_mt(int mode, const float *X, float *out,
                            const int32_t *col_idx, const float *degrees, float eps,
                            const int32_t *part_ptr, const int32_t *part2node,
                            int64_t num_nodes, int64_t dim, int64_t num_parts, int num_threads)
{
    for (int64_t w = 1; w < num_parts; w++)
        if (part2node[w] < part2node[w - 1]) return -1;
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
    const int64_t CH = 256;                      /* groups per work chunk before node-aligning */
    int64_t nchunks = (num_parts + CH - 1) / CH;
#pragma omp parallel
    {
        float *partial = (float *)malloc(sizeof(float) * (size_t)(dim ? dim : 1));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < num_nodes * dim; i++) out[i] = 0.0f;
#pragma omp for schedule(dynamic, 4)
        for (int64_t c = 0; c < nchunks; c++) {
            int64_t w0 = c * CH, w1 = w0 + CH < num_parts ? w0 + CH : num_parts;
            /* a chunk owns every group of the nodes whose FIRST group lies inside it */
            while (w0 < w1 && w0 > 0 && part2node[w0] == part2node[w0 - 1]) w0++;
            if (w0 == w1) continue;              /* all continuation groups: an earlier chunk owns them */
            while (w1 < num_parts && part2node[w1] == part2node[w1 - 1]) w1++;
            for (int64_t w = w0; w < w1; w++)
                group_body(mode, out, X, col_idx, degrees, eps, dim,
                           part2node[w], part_ptr[w], part_ptr[w + 1], partial);
        }
        free(partial);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Dense products (torch::mm -> cuBLAS SGEMM in the reference)                          */
/* ------------------------------------------------------------------------------------ */

/*
 * C[m,n] = op(A)[m,k] * op(B)[k,n], row-major, op = transpose when the flag is set.
 * cuBLAS' summation order is unspecified, so the oracle accumulates in double and rounds
 * once: the centre of the tolerance band every fp32 summation order falls in.
 * Call sites restated: kernel.cu:280 (X*W), :472 (G*W^T), :473 (X^T*G), :605 (S*W),
 * :710 (S^T*dOut), :711 (dOut*W^T).
 */
API void oracle_mm(const float *A, int trans_a, const float *B, int trans_b, float *C,
                   int64_t m, int64_t k, int64_t n, int num_threads)
{
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
    int64_t lda = trans_a ? m : k, ldb = trans_b ? k : n;
#pragma omp parallel
    {
        double *acc = (double *)malloc(sizeof(double) * (size_t)(n ? n : 1));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < m; i++) {
            for (int64_t j = 0; j < n; j++) acc[j] = 0.0;
            for (int64_t p = 0; p < k; p++) {
                double a = trans_a ? A[p * lda + i] : A[i * lda + p];
                if (trans_b)
                    for (int64_t j = 0; j < n; j++) acc[j] += a * (double)B[j * ldb + p];
                else
                    for (int64_t j = 0; j < n; j++) acc[j] += a * (double)B[p * ldb + j];
            }
            for (int64_t j = 0; j < n; j++) C[i * n + j] = (float)acc[j];
        }
        free(acc);
    }
}

/* ------------------------------------------------------------------------------------ */
/* The five operators of the extension surface (GNNAdvisor.cpp:253-263)                 */
/* ------------------------------------------------------------------------------------ */

/* SAG(X, ...) -> out           kernel.cu:110-184 */
API void oracle_SAG(const float *X, float *out, const int32_t *col_idx,
                    const int32_t *part_ptr, const int32_t *part2node,
                    int64_t num_nodes, int64_t dim, int64_t num_parts)
{
    oracle_aggregate(MODE_SAG, X, out, col_idx, NULL, 1.0f, part_ptr, part2node, num_nodes, dim, num_parts);
}

/* forward(X, W, ...) -> [out]  kernel.cu:267-322: T = X*W (:280), out = Ahat*T.
 * T is caller-provided scratch [num_nodes, dout]. */
API void oracle_forward(const float *X, const float *W, float *T, float *out,
                        const int32_t *col_idx, const float *degrees,
                        const int32_t *part_ptr, const int32_t *part2node,
                        int64_t num_nodes, int64_t din, int64_t dout, int64_t num_parts)
{
    oracle_mm(X, 0, W, 0, T, num_nodes, din, dout, 0);
    oracle_aggregate(MODE_GCN, T, out, col_idx, degrees, 1.0f, part_ptr, part2node, num_nodes, dout, num_parts);
}

/* backward(dOut, X, W, ...) -> [dX, dW]  kernel.cu:422-476:
 * G = Ahat*dOut (:436-463), dX = G*W^T (:472), dW = X^T*G (:473).  G is scratch [N,dout]. */
API void oracle_backward(const float *d_out, const float *X, const float *W, float *G,
                         float *d_input, float *d_weight,
                         const int32_t *col_idx, const float *degrees,
                         const int32_t *part_ptr, const int32_t *part2node,
                         int64_t num_nodes, int64_t din, int64_t dout, int64_t num_parts)
{
    oracle_aggregate(MODE_GCN, d_out, G, col_idx, degrees, 1.0f, part_ptr, part2node, num_nodes, dout, num_parts);
    oracle_mm(G, 0, W, 1, d_input, num_nodes, dout, din, 0);
    oracle_mm(X, 1, G, 0, d_weight, din, num_nodes, dout, 0);
}

/* forward_gin(X, W, eps, ...) -> [out, X_agg]  kernel.cu:559-617:
 * S = eps * A*X (:572-603), out = S*W (:605). */
API void oracle_forward_gin(const float *X, const float *W, float eps, float *out, float *x_agg,
                            const int32_t *col_idx,
                            const int32_t *part_ptr, const int32_t *part2node,
                            int64_t num_nodes, int64_t din, int64_t dout, int64_t num_parts)
{
    oracle_aggregate(MODE_GIN, X, x_agg, col_idx, NULL, eps, part_ptr, part2node, num_nodes, din, num_parts);
    oracle_mm(x_agg, 0, W, 0, out, num_nodes, din, dout, 0);
}

/* backward_gin(dOut, X_agg, W, eps, ...) -> [dX, dW]  kernel.cu:696-747:
 * dW = S^T*dOut (:710), Pm = dOut*W^T (:711), dX = eps * A*Pm (:712-738). Pm is scratch [N,din]. */
API void oracle_backward_gin(const float *d_out, const float *x_agg, const float *W, float eps,
                             float *Pm, float *d_input, float *d_weight,
                             const int32_t *col_idx,
                             const int32_t *part_ptr, const int32_t *part2node,
                             int64_t num_nodes, int64_t din, int64_t dout, int64_t num_parts)
{
    oracle_mm(x_agg, 1, d_out, 0, d_weight, din, num_nodes, dout, 0);
    oracle_mm(d_out, 0, W, 1, Pm, num_nodes, dout, din, 0);
    oracle_aggregate(MODE_GIN, Pm, d_input, col_idx, NULL, eps, part_ptr, part2node, num_nodes, din, num_parts);
}

API int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
