/*
 * gnna_b200.h -- C ABI of the B200-native GNNAdvisor aggregation runtime (libgnna_b200.so).
 *
 * This is the drop-in boundary for ONE path of YukeWang96/GNNAdvisor_OSDI21: the
 * neighbour-partitioned SpMM aggregation behind the `GNNAdvisor` extension module
 * (reference: GNNAdvisor/GNNConv/GNNAdvisor.cpp:253-263 exports SAG, forward, backward,
 * forward_gin, backward_gin, build_part).  Each entry point below names the reference
 * function it replaces.  Signatures carry only plain pointers and sizes; the Python module
 * `GNNAdvisor` (gnnadvisor_osdi21_b200/compat/GNNAdvisor.py) is a thin ctypes binding that
 * allocates outputs with torch and passes raw device pointers + the current CUDA stream.
 *
 * Conventions
 *   - all feature matrices are row-major contiguous; `f32` = float, `bf16` = __nv_bfloat16 bits
 *   - index arrays are int32 (the reference's IntTensor contract, GNNA_main.py:107-110)
 *   - pointers are DEVICE pointers unless the parameter name ends in `_host`
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what
 *     the reference launches on, GNNAdvisor_kernel.cu:149,298)
 *   - every call is asynchronous w.r.t. the host (like the reference) and returns GNNA_OK or a
 *     negative error code; gnna_last_error() gives the message (the reference printf()s and
 *     exit(-1)s on a launch error, kernel.cu:177-181 -- we return the error instead)
 *   - part_size / dim_worker / warp_per_block keep the reference's meaning (param.py:27-29):
 *     neighbours per group / lanes that cooperate on one neighbour row / warps per CTA.
 *     dim_worker <= 0 or warp_per_block <= 0 selects the built-in B200 choice (4 warps per CTA; one 16-byte
 *     chunk per lane, two for rows of 9..16 chunks).
 *   - a group with part_ptr[w+1] <= part_ptr[w] contributes nothing (kernel.cu:383).
 */
#ifndef GNNA_B200_H
#define GNNA_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define GNNA_API __attribute__((visibility("default")))
#else
#define GNNA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GNNA_OK 0
#define GNNA_ERR_INVALID (-1)   /* bad argument (null pointer, non-positive size, ...) */
#define GNNA_ERR_CUDA (-2)      /* CUDA runtime / launch error */
#define GNNA_ERR_CUBLAS (-3)    /* cuBLAS error in one of the dense products */
#define GNNA_ERR_UNSUPPORTED (-4)

/* Library identification and error text (thread-local). */
GNNA_API int gnna_abi_version(void);
GNNA_API const char *gnna_last_error(void);

/* ---- neighbour-group table ---------------------------------------------------------------
 * replaces build_part(int partSize, Tensor indptr)      GNNAdvisor.cpp:210-251
 * Pass 1 (:219-227): number of groups P = sum_i ceil(deg_i / part_size).                  */
GNNA_API int64_t gnna_count_parts_host(int part_size, const int32_t *indptr_host, int64_t num_nodes);

/* Pass 2 (:229-249) on the host, multi-threaded.  part_ptr_host has P+1 entries, part2node_host P.
 * compat != 0 reproduces the reference bit for bit after the caller's `.int()` cast
 * (float32 round-trip of every entry = F5, terminal left 0 when the last node is isolated = F6);
 * compat == 0 writes the integer-exact table the reference intends.                        */
GNNA_API int gnna_build_part_host(int part_size, const int32_t *indptr_host, int64_t num_nodes,
                         int32_t *part_ptr_host, int32_t *part2node_host, int64_t num_parts,
                         int compat);

/* Same table built on the device (scan + binary-search expand), integer-exact.
 * `num_parts_out_host` receives P (one 8-byte D2H copy, synchronises `stream`); call once with
 * part_ptr == NULL to size the outputs, then again with buffers of P+1 / P entries.        */
GNNA_API int gnna_build_part_device(int part_size, const int32_t *indptr, int64_t num_nodes,
                           int32_t *part_ptr, int32_t *part2node, int64_t *num_parts_out_host,
                           void *workspace, int64_t workspace_bytes, void *stream);
GNNA_API int64_t gnna_build_part_workspace_bytes(int64_t num_nodes);

/* degrees[i] = sqrtf(max(deg_i, 1))                    GNNAdvisor/dataset.py:11-18,121-122 */
GNNA_API int gnna_degrees(const int32_t *indptr, int64_t num_nodes, float *degrees, void *stream);

/* ---- aggregation kernels -------------------------------------------------------------------
 * One launch each; `out` is fully written (no pre-zeroing needed by the caller).
 *
 * replaces SAG_cuda + SAG_cuda_kernel                  GNNAdvisor_kernel.cu:110-259
 *   out[i,:] = sum_{j in N(i)} X[j,:]                                                      */
GNNA_API int gnna_sag_f32(const float *X, float *out,
                 const int32_t *row_ptr, const int32_t *col_idx,
                 const int32_t *part_ptr, const int32_t *part2node,
                 int64_t num_nodes, int dim, int64_t num_parts,
                 int part_size, int dim_worker, int warp_per_block, void *stream);

/* replaces spmm_forward_cuda_kernel / spmm_backward_cuda_kernel   kernel.cu:324-415, 478-552
 *   out[i,:] = sum_{j in N(i)} degrees[i]*degrees[j] * X[j,:]   (rounding: see gnna_set_gcn_exact)  */
GNNA_API int gnna_gcn_aggregate_f32(const float *X, float *out,
                           const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                           const int32_t *part_ptr, const int32_t *part2node,
                           int64_t num_nodes, int dim, int64_t num_parts,
                           int part_size, int dim_worker, int warp_per_block, void *stream);

/* replaces spmm_forward_cuda_kernel_gin / spmm_backward_cuda_kernel_gin  kernel.cu:620-689, 749-814
 *   out[i,:] = sum over groups g of node i of fl( eps * sum_{j in g} X[j,:] )               */
GNNA_API int gnna_gin_aggregate_f32(const float *X, float *out,
                           const int32_t *row_ptr, const int32_t *col_idx, float eps,
                           const int32_t *part_ptr, const int32_t *part2node,
                           int64_t num_nodes, int dim, int64_t num_parts,
                           int part_size, int dim_worker, int warp_per_block, void *stream);

/* Rectangular form used by the sharded (multi-GPU) path: X has num_src_rows rows (a rank's own rows
 * followed by its halo rows), out has num_dst_rows <= num_src_rows rows; col_idx indexes X, part2node
 * indexes out; degrees (GCN) has num_src_rows entries.  mode: 0 SAG, 1 GCN, 2 GIN.  No reference
 * counterpart (the reference is single-GPU).                                                   */
GNNA_API int gnna_aggregate_f32_ex(int mode, const float *X, int64_t num_src_rows, float *out, int64_t num_dst_rows,
                                   const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
                                   const int32_t *part_ptr, const int32_t *part2node, int dim, int64_t num_parts,
                                   int part_size, int dim_worker, int warp_per_block, void *stream);

/* Same, for a graph whose edges are split over several CSRs (per-owner sub-shards of the overlapped
 * multi-GPU step): accumulate != 0 ADDS into `out` (no zero fill, always reductions).  mode 3 = GCN on
 * features already scaled by degrees[j] (gnna_prescale_rows_f32); requires dim % 4 == 0.             */
GNNA_API int gnna_aggregate_part_f32_ex(int mode, int accumulate, const float *X, int64_t num_src_rows, float *out,
                                        int64_t num_dst_rows, const int32_t *row_ptr, const int32_t *col_idx,
                                        const float *degrees, float eps, const int32_t *part_ptr,
                                        const int32_t *part2node, int dim, int64_t num_parts,
                                        int part_size, int dim_worker, int warp_per_block, void *stream);

/* The sharded step with the halo exchange FUSED into the aggregation: ONE launch over the concatenation of a rank's
 * per-owner sub-CSRs (segment s = groups [seg_bounds[s], seg_bounds[s+1]), gathering rows owned by rank seg_peer[s]; -1 =
 * the rank's own rows).  The CTAs of a peer's segment wait INSIDE the kernel (ld.acquire.sys on that peer's flag in the halo
 * control block `my_ctrl`, bounded) until the peer's push has delivered this step's rows, while the CTAs ahead of them
 * aggregate what has already landed: the NVLink transfer overlaps the gather segment by segment with no kernel boundary
 * in between.  out is zero-filled and accumulated with reductions.  mode 0, 2, 3; dim % 4 == 0.  No reference counterpart. */
GNNA_API int gnna_aggregate_gated_f32(int mode, const float *X, int64_t num_src_rows, float *out, int64_t num_dst_rows,
                                      const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
                                      const int32_t *part_ptr, const int32_t *part2node, int dim, int64_t num_parts,
                                      const int64_t *seg_bounds_host, const int32_t *seg_peer_host, int num_segs,
                                      void *my_ctrl, int part_size, int dim_worker, int warp_per_block, void *stream);

/* Xs[i,:] = degrees[i] * X[i,:] (X == Xs allowed): the pre-scale pass of the default GCN path, exposed
 * so a producer can hand pre-scaled rows to gnna_aggregate_part_f32_ex(mode 3) / the halo push.          */
GNNA_API int gnna_prescale_rows_f32(const float *X, float *Xs, const float *degrees, int64_t num_rows, int dim, void *stream);

/* Row-major C[m,n] = op(A) * op(B), fp32 (cuBLAS SGEMM, TF32 off): the library call the reference makes through
 * torch::mm (GNNAdvisor_kernel.cu:280,472,473,605,710,711), for callers that compose a layer from its parts --
 * the sharded layers run product, halo exchange and aggregation as separate steps.                        */
GNNA_API int gnna_sgemm_f32(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k,
                            const float *A, const float *B, float *C, void *stream);

/* bf16-storage variant: neighbour rows are gathered as bf16 (half the gather bytes), summed in
 * fp32 and written as fp32.  An extension: the reference is fp32-only.
 * mode: 0 = SAG, 1 = GCN (degrees, per-edge weights), 2 = GIN (eps), 3 = GCN on features the caller
 * already scaled by degrees[j] (out_i = degrees[i] * sum_j X[j]: no per-edge degree gather).     */
GNNA_API int gnna_aggregate_bf16(int mode, const void *X_bf16, float *out_f32,
                        const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
                        const int32_t *part_ptr, const int32_t *part2node,
                        int64_t num_nodes, int dim, int64_t num_parts,
                        int part_size, int dim_worker, int warp_per_block, void *stream);

/* ---- the four layer operators (aggregation + the dense products around it) -----------------
 * Dense products are fp32 cuBLAS SGEMM (no TF32), as torch::mm is in the reference.
 *
 * replaces spmm_forward_cuda          kernel.cu:267-322   T = X*W ; out = Ahat*T
 *   X [N,din], W [din,dout], T_ws [N,dout] scratch, out [N,dout]                            */
GNNA_API int gnna_forward_f32(const float *X, const float *W, float *T_ws, float *out,
                     const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                     const int32_t *part_ptr, const int32_t *part2node,
                     int64_t num_nodes, int din, int dout, int64_t num_parts,
                     int part_size, int dim_worker, int warp_per_block, void *stream);

/* replaces spmm_backward_cuda         kernel.cu:422-476   G = Ahat*dOut ; dX = G*W^T ; dW = X^T*G
 *   d_out [N,dout], G_ws [N,dout] scratch, d_input [N,din] (may be NULL: skip it), d_weight [din,dout] */
GNNA_API int gnna_backward_f32(const float *d_out, const float *X, const float *W, float *G_ws,
                      float *d_input, float *d_weight,
                      const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                      const int32_t *part_ptr, const int32_t *part2node,
                      int64_t num_nodes, int din, int dout, int64_t num_parts,
                      int part_size, int dim_worker, int warp_per_block, void *stream);

/* replaces spmm_forward_cuda_gin      kernel.cu:559-617   S = eps*A*X ; out = S*W
 *   x_agg [N,din] receives S (saved for backward, gnn_conv.py:109), out [N,dout]            */
GNNA_API int gnna_forward_gin_f32(const float *X, const float *W, float eps, float *out, float *x_agg,
                         const int32_t *row_ptr, const int32_t *col_idx,
                         const int32_t *part_ptr, const int32_t *part2node,
                         int64_t num_nodes, int din, int dout, int64_t num_parts,
                         int part_size, int dim_worker, int warp_per_block, void *stream);

/* replaces spmm_backward_cuda_gin     kernel.cu:696-747   dW = S^T*dOut ; Pm = dOut*W^T ; dX = eps*A*Pm
 *   Pm_ws [N,din] scratch, d_input [N,din] (may be NULL: only dW is computed), d_weight [din,dout] */
GNNA_API int gnna_backward_gin_f32(const float *d_out, const float *x_agg, const float *W, float eps,
                          float *Pm_ws, float *d_input, float *d_weight,
                          const int32_t *row_ptr, const int32_t *col_idx,
                          const int32_t *part_ptr, const int32_t *part2node,
                          int64_t num_nodes, int din, int dout, int64_t num_parts,
                          int part_size, int dim_worker, int warp_per_block, void *stream);

/* ---- mixed-precision layers (BASELINE.json config "Reddit GCN 2-layer D=64 bf16"; the reference is fp32-only) ----
 * Same operators as gnna_forward_f32 / gnna_backward_f32 (spmm_forward_cuda kernel.cu:267-322, spmm_backward_cuda
 * :422-476) with the GATHERED matrix stored as bf16: dense products, accumulation and every output stay fp32.
 *   forward : T = X*W (SGEMM) ; Tb_j = bf16(n_j*T_j) ; out_i = n_i * sum_j Tb_j
 *   backward: Gb_j = bf16(n_j*dOut_j) ; G_i = n_i * sum_j Gb_j ; dX = G*W^T (d_input may be NULL) ; dW = X^T*G
 * Tb_ws / Gb_ws: bf16 scratch [N, round_up(dout, 8)] (16-byte aligned); T_ws / G_ws: fp32 scratch [N, dout].  */
GNNA_API int gnna_forward_mixed(const float *X, const float *W, float *T_ws, void *Tb_ws, float *out,
                                const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                                const int32_t *part_ptr, const int32_t *part2node,
                                int64_t num_nodes, int din, int dout, int64_t num_parts,
                                int part_size, int dim_worker, int warp_per_block, void *stream);
GNNA_API int gnna_backward_mixed(const float *d_out, const float *X, const float *W, void *Gb_ws, float *G_ws,
                                 float *d_input, float *d_weight,
                                 const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                                 const int32_t *part_ptr, const int32_t *part2node,
                                 int64_t num_nodes, int din, int dout, int64_t num_parts,
                                 int part_size, int dim_worker, int warp_per_block, void *stream);
/* The GIN pair the same way (spmm_forward_cuda_gin kernel.cu:559-617, spmm_backward_cuda_gin :696-747): the gathered
 * matrices (X forward; Pm = dOut*W^T backward) are converted to bf16 first.  Xb_ws / Pmb_ws: bf16 scratch
 * [N, round_up(din, 8)]; x_agg, Pm_ws, d_input (may be NULL: only dW), d_weight as in the fp32 pair.             */
GNNA_API int gnna_forward_gin_mixed(const float *X, const float *W, float eps, void *Xb_ws, float *out, float *x_agg,
                                    const int32_t *row_ptr, const int32_t *col_idx,
                                    const int32_t *part_ptr, const int32_t *part2node,
                                    int64_t num_nodes, int din, int dout, int64_t num_parts,
                                    int part_size, int dim_worker, int warp_per_block, void *stream);
GNNA_API int gnna_backward_gin_mixed(const float *d_out, const float *x_agg, const float *W, float eps,
                                     float *Pm_ws, void *Pmb_ws, float *d_input, float *d_weight,
                                     const int32_t *row_ptr, const int32_t *col_idx,
                                     const int32_t *part_ptr, const int32_t *part2node,
                                     int64_t num_nodes, int din, int dout, int64_t num_parts,
                                     int part_size, int dim_worker, int warp_per_block, void *stream);
/* Building blocks of the four above.  scale_rows: Xb[i, 0:dim] = bf16(degrees[i] * X[i, :]) (degrees NULL: plain
 * conversion), columns dim..ldb-1 zero, ldb % 8 == 0.  aggregate_bf16_ex: gnna_aggregate_bf16 on rows of stride ldx
 * elements (ldx == dim, or ldx % 8 == 0 for padded rows).                                              */
GNNA_API int gnna_scale_rows_bf16(const float *X, void *Xb_bf16, const float *degrees, int64_t num_rows, int dim, int ldb,
                                  void *stream);
GNNA_API int gnna_aggregate_bf16_ex(int mode, const void *X_bf16, int ldx, float *out_f32,
                                    const int32_t *row_ptr, const int32_t *col_idx, const float *degrees, float eps,
                                    const int32_t *part_ptr, const int32_t *part2node,
                                    int64_t num_nodes, int dim, int64_t num_parts,
                                    int part_size, int dim_worker, int warp_per_block, void *stream);

/* ---- fused aggregate -> X*W on the tensor cores (csrc/fused_gemm.cu) -----------------------------------
 * out = (c_i * sum_{j in N(i)} X[j,:]) * W in ONE kernel: a CTA gathers 128 destination rows into shared
 * memory, converts them to bf16 and multiplies by W (bf16) with tcgen05.mma, fp32 accumulation in TMEM.
 * Fuses what the reference does as aggregation kernel + torch::mm (spmm_forward_cuda_gin, kernel.cu:559-617)
 * for the bf16 configuration of BASELINE.json; the fp32 operators above stay bit-faithful and unfused.
 *   mode 0: c_i = 1 (SAG)   2: c_i = eps (GIN forward)   3: c_i = degrees[i], X already scaled by degrees[j] (GCN)
 *   X [N, din] fp32 (x_is_bf16 = 0) or bf16 (1); W [din, dout] fp32 (rounded to bf16 inside); out [N, dout] fp32;
 *   x_agg [N, din] fp32 receives the aggregated features (may be NULL).  Supported: din in {32, 64, 128} (fp32 X)
 *   or {64, 128, 256} (bf16 X), 1 <= dout <= 256; GNNA_ERR_UNSUPPORTED otherwise.                      */
GNNA_API int gnna_aggregate_gemm_fused_bf16(int mode, const void *X, int x_is_bf16, const float *W, float eps,
                                            float *out, float *x_agg,
                                            const int32_t *row_ptr, const int32_t *col_idx, const float *degrees,
                                            const int32_t *part_ptr, const int32_t *part2node,
                                            int64_t num_nodes, int din, int dout, int64_t num_parts,
                                            int part_size, int dim_worker, int warp_per_block, void *stream);

/* ---- NVLink-native halo exchange for the sharded path (csrc/halo.cu; no reference counterpart) ------
 * Device memory that other processes on the node can map (CUDA IPC): alloc returns a zeroed buffer and
 * its 64-byte handle; peers open / close it; the owner frees it.                                   */
GNNA_API int gnna_ipc_alloc(int64_t bytes, void **ptr, unsigned char *handle64);
GNNA_API int gnna_ipc_open(const unsigned char *handle64, void **ptr);
GNNA_API int gnna_ipc_close(void *ptr);
GNNA_API int gnna_ipc_free(void *ptr);

/* One exchange step = begin_step, push, wait, (aggregate), ack.  The step number (1, 2, 3, ...) lives in the
 * control block and is read on the device, so a step captured in a CUDA graph can be replayed.
 * begin_step: first thing on the compute stream: step += 1.
 * push: ONE persistent kernel (GNNA_PUSH_CTAS CTAs, default 96) copies, peer after peer in ring order
 * (my_rank+1, my_rank+2, ...), rows x_local[send_idx[send_begin[p] .. send_begin[p+1])] into rows
 * peer_dst_row0[p].. of peer p's mapped feature buffer (128-bit stores over NVLink) and then stores the step
 * into flags[my_rank] of p's control block (release, system scope).  Must be ordered after begin_step.
 * wait: returns (on the stream) once the flag of every peer in `peer_mask` (bit q = rank q; 0 = all peers)
 * in MY control block reached the step.
 * ack: after the aggregation that read the halo rows: stores the step into acks[my_rank] of every peer, so
 * the buffer of this step parity may be overwritten at step+2 (push waits for it).
 * Control block: 64 uint32 in IPC memory: [0,16) flags, [16,32) acks, [32,48) scratch, [48] error word
 * (non-zero: a bounded wait timed out), [49] step.  *_host arrays have `world` entries.              */
GNNA_API int gnna_halo_begin_step(void *my_ctrl, void *stream);
GNNA_API int gnna_halo_push_f32(const float *x_local, const int64_t *send_idx, const int32_t *send_begin_host,
                                void *const *peer_feature_base_host, void *const *peer_ctrl_host,
                                const int64_t *peer_dst_row0_host, void *my_ctrl,
                                int world, int my_rank, int dim, void *stream);

/* gnna_halo_push_f32 for DENSE halos: a peer whose bit (by rank) is set in dense_mask asked for ALL n_local rows of this
 * rank (its block for this rank is the rank's local rows as they lie in memory) and gets them as ONE device-to-device
 * cudaMemcpyAsync over NVLink on the copy engines -- no SM, no index list; peers not in the mask are served by the push
 * kernel from their send lists.  Same flags / acknowledgements / step counter as gnna_halo_push_f32.          */
GNNA_API int gnna_halo_push_ce(const float *x_local, int64_t n_local, const int64_t *send_idx, const int32_t *send_begin_host,
                               void *const *peer_feature_base_host, void *const *peer_ctrl_host,
                               const int64_t *peer_dst_row0_host, void *my_ctrl,
                               int world, int my_rank, int dim, uint32_t dense_mask, void *stream);
GNNA_API int gnna_halo_wait(void *my_ctrl, int world, int my_rank, uint32_t peer_mask, void *stream);
GNNA_API int gnna_halo_ack(void *const *peer_ctrl_host, void *my_ctrl, int world, int my_rank, void *stream);

/* ---- vertex reordering ----------------------------------------------------------------------
 * replaces the python module `rabbit` (rabbit_module/src/reorder.cpp:235-295, rabbit_order.hpp:393-673):
 * Rabbit Order community-based renumbering.  Host code on all host threads, deterministic (the same permutation for
 * any number of threads; the reference's optimistic parallel merges give a different one every run).  The edge list may
 * be directed and may contain duplicates and self loops (it is symmetrised internally, like the reference does).
 * perm_old_to_new_host[v] = new id of vertex v (a permutation of 0..num_nodes-1).
 * _ex: `window` = vertices whose merges are evaluated concurrently against one state (csrc/reorder.cu); 1 = the
 * sequential algorithm, <= 0 = GNNA_RABBIT_WINDOW or the built-in choice min(16384, num_nodes / 64).             */
GNNA_API int gnna_rabbit_reorder_host(const int32_t *src_host, const int32_t *dst_host, int64_t num_edges,
                                      int64_t num_nodes, int32_t *perm_old_to_new_host);
GNNA_API int gnna_rabbit_reorder_host_ex(const int32_t *src_host, const int32_t *dst_host, int64_t num_edges,
                                         int64_t num_nodes, int32_t *perm_old_to_new_host, int64_t window);

/* ---- dataset path on the host (csrc/dataset.cu) ------------------------------------------------
 * replaces the loader of GNNAdvisor/dataset.py: the per-line Python loop over a whitespace `src dst` text file
 * (:62-72) and scipy's coo_matrix(...).tocsr() (:108-111), both on all host threads.
 * gnna_edge_text_scan: an upper bound of the number of edges in the file (its line count), to size the buffers.
 * gnna_edge_text_parse: the edges in file order; lines that are blank or start with '#' / '%' are skipped, tokens after
 * the second are ignored, anything else that is not two integers is an error naming the line (the reference raises
 * ValueError); *num_nodes_host = largest id + 1 (:72).                                                              */
GNNA_API int gnna_edge_text_scan(const char *path_host, int64_t *max_edges_host);
GNNA_API int gnna_edge_text_parse(const char *path_host, int64_t *src_host, int64_t *dst_host, int64_t capacity,
                                  int64_t *num_edges_host, int64_t *num_nodes_host);
/* COO -> CSR exactly as scipy builds it for the reference (dataset.py:108-122): duplicate edges merged, the columns of a
 * row ascending, self loops kept; an endpoint outside [0, num_nodes) is an error (scipy raises too).  row_ptr_host has
 * num_nodes + 1 entries, col_idx_host room for num_edges; *nnz_host = edges left after merging (< 2^31).            */
GNNA_API int gnna_csr_from_edges_host(const int64_t *src_host, const int64_t *dst_host, int64_t num_edges,
                                      int64_t num_nodes, int32_t *row_ptr_host, int32_t *col_idx_host, int64_t *nnz_host);

/* ---- introspection (for tests, bench.py and the tuner) ------------------------------------
 * Launch geometry the library would use for a given call: fills lanes_per_row, chunks_per_lane,
 * vec_width, warps_per_block, grid_x, grid_y.  Returns GNNA_OK.                             */
typedef struct gnna_launch_info {
    int vec_width;        /* elements per load: 4, 2 or 1 (fp32); 8, 4, 2, 1 (bf16) */
    int lanes_per_row;    /* lanes cooperating on one neighbour row (dim_worker, power of two) */
    int chunks_per_lane;  /* vector chunks each lane owns inside one d-tile */
    int warps_per_block;
    int groups_per_warp;  /* 32 / lanes_per_row */
    int64_t grid_x;
    int grid_y;           /* d-tiles */
    int kernels;          /* kernel launches this aggregation issues (memset node excluded) */
} gnna_launch_info;
GNNA_API int gnna_query_launch(int elem_bytes, int dim, int64_t num_parts, int dim_worker, int warp_per_block,
                      gnna_launch_info *info);

/* Workload generation: pairs [start, start+count) of the counter-based pair stream that defines the synthetic look-alike
 * graphs (gnnadvisor_osdi21_b200/graph.py: stream_pairs -- splitmix64 of the pair index, R-MAT (kind 0, thresholds scaled by
 * 2^32, `bits` levels) or uniform (kind 1)); dropped pairs (self loops, ids >= num_nodes) come back as -1/-1.  Bit for bit what
 * the torch definition gives on the CPU, as one kernel.  No reference counterpart (the reference reads dataset files).   */
GNNA_API int gnna_stream_pairs(int64_t start, int64_t count, uint64_t seed_mix, int64_t num_nodes, int kind, int bits,
                               uint64_t t_a, uint64_t t_ab, uint64_t t_abc, int64_t *src_out, int64_t *dst_out, void *stream);

/* Measurement infrastructure: read `bytes` of `buf` (a buffer that fits in L2) `passes` times with 128-bit loads that
 * bypass L1 -- mode 0 a coalesced stream, mode 1 randomly ordered 256-byte rows (the D=64 fp32 gather's pattern).  Timed
 * by the caller with CUDA events, it gives bench.py the L2 -> SM bandwidth of the box: the roof of the aggregation when
 * the feature matrix is L2-resident.  threads_per_sm (<= 0: 1536) come in CTAs of block_threads (128 / 256 / 512).  `bytes` is
 * rounded down to whole batches of the grid; *bytes_per_pass = what one pass reads.  `sink`: any 4 writable device bytes.  No reference counterpart.                                       */
GNNA_API int gnna_probe_l2_read(const void *buf, int64_t bytes, int passes, int mode, int threads_per_sm, int block_threads,
                                void *sink, int64_t *bytes_per_pass, void *stream);

/* GCN rounding: 0 (default) = out_i = n_i * sum_j (n_j * x_j): one pre-scale pass over the features,
 * then a weight-free gather (no per-edge degrees[nid] gather; each term within 2 roundings of the
 * reference's).  1 = the reference's per-edge fl(fl(n_i*n_j) * x_j) (kernel.cu:389,403), bit-identical
 * to it for every node whose neighbours fit one group.  Also settable with GNNA_GCN_EXACT=1 in the
 * environment.  Returns the previous setting.                                                  */
GNNA_API int gnna_set_gcn_exact(int on);

/* Group tables of at most `limit` groups (default 16384, GNNA_SMALL_PARTS in the environment; 0 = never) are aggregated by
 * ONE kernel launch in which a sub-warp owns a destination row and writes it once -- no zero-fill, no pre-scale pass, no
 * scratch allocation (csrc/aggregate_small.cu): the launch-latency-bound graphs (Cora, citeseer).  Same group-table
 * semantics and rounding as the general path.  Returns the previous limit.                                    */
GNNA_API int64_t gnna_set_small_parts(int64_t limit);

/* Which dense products run on the tensor cores (csrc/gemm_tf32x3.cu: tcgen05 kind::tf32 with the 3xTF32 split, fp32-grade
 * accuracy) instead of cuBLAS' SIMT SGEMM (what torch::mm is in the reference).  Candidates are the two tall-skinny
 * contractions of a layer: X*W (>= 8192 rows, K >= 256) and X^T*G (reduced over >= 8192 rows, >= 256 columns), n <= 128.
 * 0 = none; 1 = both, lockstep kernel; 2 = both, warp-specialised kernel; 3 (default) = the measured winner per product:
 * X^T*G on the warp-specialised kernel, X*W on cuBLAS.  Also GNNA_TC_GEMM in the environment.  Returns the previous mode. */
GNNA_API int gnna_set_tc_gemm(int mode);

/* 1 = use the persistent kernel that streams the group table and the column indices through TMA bulk copies
 * (cp.async.bulk) into a shared-memory ring (csrc/aggregate_staged.cu) where it applies (fp32, dim % 4 == 0,
 * 8 <= dim <= 128, part_size <= 64); 0 (default) = the occupancy-driven kernel (csrc/aggregate.cu).  Also
 * GNNA_STAGED=1 in the environment.  Returns the previous setting.  Measured 7-17 % slower than the default on
 * B200 (the index streams are 1.5 % of the bytes), hence opt-in.                                       */
GNNA_API int gnna_set_staged(int on);

/* The run-based software-pipelined kernel (csrc/aggregate_runs.cu; unweighted modes, 128-bit rows of at most 32
 * chunks): a sub-warp owns R consecutive neighbour-groups, prefetches the table entries and ids of the next group
 * while it gathers the rows of the current one, and merges groups of one node in registers.
 * R > 0 = use it with runs of R groups wherever it applies; 0 = never (always csrc/aggregate.cu);
 * -1 (default) = the library chooses from its measurements (aggregate_runs.cu auto_runs): on graphs with >= 4 groups
 * per node every bf16 width and the fp32 widths that are not whole 128-byte lines; on sparser graphs bf16 rows of >= 6
 * chunks.  Also GNNA_RUNS=R in the environment.  Returns the previous setting.                                  */
GNNA_API int gnna_set_runs(int run);
/* The run length a call of this shape would get under the current setting (0 = csrc/aggregate.cu): row_elems = row
 * stride of the gathered matrix in elements (after padding: fp32 rows are padded to a multiple of 4, bf16 rows of the
 * mixed operators to a multiple of 8).  Host-only, no GPU needed.                                            */
GNNA_API int gnna_query_runs(int elem_bytes, int row_elems, int64_t num_nodes, int64_t num_parts);

/* Number of kernels this library has launched on this thread since the last reset
 * (bench.py's "gpu_launches" is read from here, not guessed). */
GNNA_API int64_t gnna_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* GNNA_B200_H */
