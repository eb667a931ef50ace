/*
 * cnn_graph_b200 -- C ABI of the B200-native Chebyshev graph-convolution hot path.
 *
 * The reference (xu-wang11/cnn_graph) has no FFI: its hot path sits behind a
 * Python name-binding surface (getattr(self, 'chebyshev5'), getattr(filter,
 * 'cheby_conv'), the RNN-cell protocol).  This header is the boundary a
 * maintainer of the reference binds with ctypes from those Python methods
 * (INTEGRATION.md shows the stubs).  Every entry point cites the reference
 * function it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - all tensors are dense, contiguous, float32 unless stated; "dev" pointers
 *     are CUDA device pointers owned by the caller, "host" pointers are CPU;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success, non-zero on error; cg_last_error()
 *     returns a thread-local message for the last failure;
 *   - no function allocates device memory behind the caller's back except
 *     cg_graph_create (the packed operator, freed by cg_graph_destroy);
 *     scratch is passed in by the caller, sized by the *_workspace_bytes query;
 *   - thread-compatible: no global state besides the per-thread error string.
 *
 * Tensor layouts (reference layouts, unchanged):
 *   x   [N, M, Fin]      lib/models.py:193   (N signals, M vertices, Fin features)
 *   W   [Fin*K, Fout]    lib/models.py:222   row = fin*K + k  (fin-major)
 *   y   [N, M, Fout]     lib/models.py:224
 *   Xt  [K, M, C]        lib/graph.py:248    Chebyshev basis of X [M, C]
 */
#ifndef CNN_GRAPH_B200_H
#define CNN_GRAPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CG_ABI_VERSION 1

/* ---- status ------------------------------------------------------------ */
enum {
    CG_OK = 0,
    CG_ERR_ARG = 1,      /* bad argument (shape, null pointer, unsupported size)   */
    CG_ERR_CUDA = 2,     /* CUDA runtime error; message carries cudaGetErrorString */
    CG_ERR_WORKSPACE = 3 /* workspace too small                                     */
};

int cg_abi_version(void);
const char *cg_last_error(void);

/* ---- rescaled Laplacian handle ---------------------------------------- */
/* Opaque packed operator: L~ and its transpose, CSR + ELL, resident in HBM.
 * Replaces the per-call COO -> tf.SparseTensor -> tf.sparse_reorder staging of
 * lib/models.py:198-201 and lib/filter.py:66-70.  The caller passes the
 * ALREADY RESCALED matrix (graph.rescale_L, lib/graph.py:232-238) as host CSR
 * with sorted column indices. */
typedef struct cg_graph cg_graph_t;

int cg_graph_create(cg_graph_t **out, int M, int64_t nnz, const int32_t *host_indptr,
                    const int32_t *host_indices, const float *host_values);
int cg_graph_destroy(cg_graph_t *g);
/* info[0]=M, [1]=nnz, [2]=max row length of L~, [3]=max row length of L~^T,
 * [4]=1 if the operator fits the shared-memory (on-chip) kernels, else 0     */
int cg_graph_info(const cg_graph_t *g, int64_t info[5]);

/* ---- Chebyshev basis ---------------------------------------------------- */
/* graph.chebyshev(L, X, K)  lib/graph.py:241-258:
 *   Xt[0] = X, Xt[1] = L~ X, Xt[k] = 2 L~ Xt[k-1] - Xt[k-2].
 * dev_X [M, C], dev_Xt [K, M, C].  transpose != 0 applies L~^T instead.
 * flags: CG_FILTER_* (kernel selection, see below).                           */
int cg_cheb_basis(const cg_graph_t *g, int transpose, const float *dev_X, float *dev_Xt,
                  int64_t C, int K, int flags, void *stream);

/* One step of the same recurrence on caller-owned slabs: out[0:rows] = alpha * (L~ X1)[0:rows] - X0[0:rows]
 * (dev_X0 may be NULL).  dev_X1 has all M rows of the (padded) operator; used by the row-partitioned
 * recurrence of config C5, which exchanges the halo rows of X1 between two steps.              */
int cg_cheb_step(const cg_graph_t *g, int transpose, const float *dev_X1, const float *dev_X0, float *dev_out,
                 int rows, int64_t C, float alpha, void *stream);
/* The same step restricted to a list of row tiles (row-partitioned runs compute the tiles that need no halo row while
 * the halo is in flight, the others after).  cg_cheb_step_tile_rows: rows per tile for slabs of C columns on this side
 * of the operator, 0 when the tiled step does not apply (then use cg_cheb_step).  dev_tiles [ntiles] int32 on the device. */
int cg_cheb_step_tile_rows(const cg_graph_t *g, int transpose, int64_t C);
int cg_cheb_step_tiles(const cg_graph_t *g, int transpose, const float *dev_X1, const float *dev_X0, float *dev_out, int rows,
                       int64_t C, float alpha, const int32_t *dev_tiles, int ntiles, void *stream);
/* Halo rows of a row-partitioned slab read from the owners' buffers over NVLink (peer / symmetric memory, SURVEY.md
 * 8(e)(2)): dev_dst [nhalo][C] <- peer[src_rank[i]][slab_offset + src_row[i] * C ...].  dev_peer_ptrs: device array of
 * world float* (the ranks' buffer bases, e.g. _SymmetricMemory.buffer_ptrs_dev); slab_offset in elements.             */
int cg_halo_pull(const void *dev_peer_ptrs, const int32_t *dev_src_rank, const int32_t *dev_src_row, int64_t slab_offset,
                 float *dev_dst, int64_t nhalo, int C, void *stream);


/* ---- Chebyshev filter (chebyshev5 / chebyshev2 / cheby_conv) ---------- */
/* Forward: lib/models.py:192-224, lib/graph_conv.py:144-176, lib/filter.py:45-95
 *   y[n,m,fo] = sum_{fin,k} (T_k(L~) x)[n,m,fin] * W[fin*K+k, fo]
 * Backward (TF autodiff of the above, lib/graph_model.py:296):
 *   dx[n,m,fin] = sum_{k,fo} (T_k(L~^T) gy)[n,m,fo] * W[fin*K+k, fo]
 *   dW[fin*K+k, fo] = sum_{n,m} (T_k(L~) x)[n,m,fin] * gy[n,m,fo]
 * dev_dx may be NULL (first layer / chebyshev2, which has no x-gradient,
 * lib/models.py:183).  dev_dW is overwritten (not accumulated).
 * `flags`: CG_FILTER_* bits.                                                 */
enum {
    CG_FILTER_DEFAULT = 0,
    CG_FILTER_FORCE_STREAMING = 1, /* never use the on-chip (SMEM-resident) kernels */
    CG_FILTER_FORCE_ONCHIP = 2,    /* fail with CG_ERR_ARG if the on-chip kernels do not fit */
    CG_FILTER_NO_FUSED = 4,        /* never use the fused recurrence+contraction (tcgen05) kernel */
    CG_FILTER_FORCE_FUSED = 8,     /* fail with CG_ERR_ARG if the fused kernel does not support the shape */
    CG_FILTER_NO_CLENSHAW = 16,    /* input gradient by the forward-form fused kernel on L~^T, not the adjoint recurrence */
    CG_FILTER_STACK_PLANES = 32    /* the saved basis (fwd_ex stack_out / bwd_ex saved_stack) is the fused kernel's bf16 hi|mid
                                      operand plane image [2][K][ceil(N*M/128)][Fin/8][128][8]; only where
                                      cg_cheb_filter_stack_planes() returns 1.  Pass the same flag to both calls and to
                                      cg_cheb_filter_stack_bytes (rows are padded to whole chunks of 128). */
};
size_t cg_cheb_filter_fwd_workspace_bytes(const cg_graph_t *g, int N, int Fin, int Fout, int K, int flags);
size_t cg_cheb_filter_bwd_workspace_bytes(const cg_graph_t *g, int N, int Fin, int Fout, int K,
                                          int need_dx, int flags);
int cg_cheb_filter_fwd(const cg_graph_t *g, const float *dev_x, const float *dev_W, float *dev_y,
                       int N, int Fin, int Fout, int K, void *dev_workspace, size_t workspace_bytes,
                       int flags, void *stream);
int cg_cheb_filter_bwd(const cg_graph_t *g, const float *dev_x, const float *dev_W, const float *dev_gy,
                       float *dev_dx, float *dev_dW, int N, int Fin, int Fout, int K,
                       void *dev_workspace, size_t workspace_bytes, int flags, void *stream);

/* Training variants.  The forward pass can leave the Chebyshev basis X_k = T_k(L~) x behind
 * (dev_stack_out, [K, N, M, Fin] float32, sample-major) so that the backward pass forms
 * dW = X_k^T gy without repeating the recurrence -- the B200 answer to TF keeping the whole
 * concat/transposed stack alive for autodiff (lib/models.py:207-220).
 * cg_cheb_filter_stack_bytes returns the size of that buffer, or 0 when the shape cannot use it
 * (then pass NULL).  dev_saved_stack in the backward may be NULL (the basis is recomputed). */
size_t cg_cheb_filter_stack_bytes(const cg_graph_t *g, int N, int Fin, int Fout, int K, int flags);
/* 1 when the saved basis of this shape can use the CG_FILTER_STACK_PLANES format (fused forward kernel + the
 * plane-streaming weight-gradient kernel), else 0. */
int cg_cheb_filter_stack_planes(const cg_graph_t *g, int N, int Fin, int Fout, int K, int flags);
int cg_cheb_filter_fwd_ex(const cg_graph_t *g, const float *dev_x, const float *dev_W, float *dev_y,
                          float *dev_stack_out, int N, int Fin, int Fout, int K, void *dev_workspace,
                          size_t workspace_bytes, int flags, void *stream);
int cg_cheb_filter_bwd_ex(const cg_graph_t *g, const float *dev_x, const float *dev_W, const float *dev_gy,
                          const float *dev_saved_stack, float *dev_dx, float *dev_dW, int N, int Fin, int Fout,
                          int K, void *dev_workspace, size_t workspace_bytes, int flags, void *stream);

/* ---- bias + activation (b1relu / b1tanh / b2relu) ---------------------- */
/* lib/models.py:226-247.  bias_kind: 0 none (fork b1relu), 1 per filter
 * [F] (upstream b1relu, b1tanh), 2 per vertex and filter [M, F] (b2relu).
 * act: 0 identity, 1 relu, 2 tanh.
 * Backward takes the forward OUTPUT y (relu: mask y > 0; tanh: 1 - y^2).
 * dev_dbias may be NULL; when given it is overwritten.                       */
int cg_bias_act_fwd(const float *dev_x, const float *dev_bias, float *dev_y, int N, int M, int F,
                    int bias_kind, int act, void *stream);
int cg_bias_act_bwd(const float *dev_y, const float *dev_gy, float *dev_gx, float *dev_dbias,
                    int N, int M, int F, int bias_kind, int act, void *stream);

/* ---- permuted pooling (mpool1 / apool1) -------------------------------- */
/* lib/models.py:249-266: max / mean over p consecutive (permuted) vertices.
 * kind: 1 max, 2 avg.  dev_argmax [N, M/p, F] uint8 (max only, may be NULL on
 * avg): index in 0..p-1 of the FIRST maximal element (TF MaxPoolGrad routing). */
int cg_pool_fwd(const float *dev_x, float *dev_y, uint8_t *dev_argmax, int N, int M, int F, int p,
                int kind, void *stream);
int cg_pool_bwd(const float *dev_gy, const uint8_t *dev_argmax, float *dev_gx, int N, int M, int F,
                int p, int kind, void *stream);

/* ---- fused bias + activation + pooling -------------------------------- */
/* The brelu -> pool tail of a cgcnn layer (lib/models.py:226-266 as sequenced by the model's
 * _inference) in one pass: dev_y [N, M/p, F] = pool_p(act(x + bias)).  dev_aux [N, M/p, F] uint8:
 * max pooling -- index of the first maximal activated value; avg pooling -- bit q set when
 * activated element q is > 0.  The backward takes the pooled gradient, the pooled OUTPUT and aux
 * and writes the gradient of x (and of the bias, overwritten, when dev_dbias is non-NULL).
 * p <= 8; avg pooling after tanh is not supported (use the separate entry points).        */
int cg_bias_act_pool_fwd(const float *dev_x, const float *dev_bias, float *dev_y, uint8_t *dev_aux, int N,
                         int M, int F, int p, int bias_kind, int act, int kind, void *stream);
int cg_bias_act_pool_bwd(const float *dev_gy, const float *dev_y, const uint8_t *dev_aux, float *dev_gx,
                         float *dev_dbias, int N, int M, int F, int p, int bias_kind, int act, int kind,
                         void *stream);

/* ---- first-layer fusion ------------------------------------------------------------------------------------ */
/* Scalar-input layer (Fin = 1) followed by bias (per filter or none) + relu + max pooling of 4 (lib/models.py:226-257
 * after :192-224): weight and bias gradient straight from the gradient of the POOLED output -- the pooling-backward
 * pass and the 4x larger filter-output gradient are never formed.  dev_stack: the fp32 basis [K][N][M] that
 * cg_cheb_filter_fwd_ex left behind (no CG_FILTER_STACK_PLANES); dev_g_pooled / dev_y_pooled / dev_aux [N][M/4][Fout]:
 * gradient of the pooled output, pooled output and argmax bytes of cg_bias_act_pool_fwd; dev_dW [K][Fout]; dev_db
 * [Fout] or NULL.  No input gradient (first layer).                                                                 */
int cg_cheb_dw_pooled_supported(const cg_graph_t *g, int N, int Fout, int K, int p, int act, int kind, int bias_kind);
/* Forward of that layer without the [N, M, Fout] filter output: recurrence (the fp32 basis [K][N][M] is left in
 * dev_stack_out for cg_cheb_dw_pooled), then the contraction with bias + relu + max pooling of 4 in its epilogue.
 * dev_y_pooled / dev_aux [N][M/4][Fout] as written by cg_bias_act_pool_fwd.  bias_kind: 0 (dev_bias NULL) or 1.  */
int cg_cheb_first_layer_fwd_supported(const cg_graph_t *g, int N, int Fout, int K, int bias_kind);
int cg_cheb_first_layer_fwd(const cg_graph_t *g, const float *dev_x, const float *dev_W, const float *dev_bias,
                            float *dev_stack_out, float *dev_y_pooled, uint8_t *dev_aux, int N, int Fout, int K, void *stream);
size_t cg_cheb_dw_pooled_workspace_bytes(const cg_graph_t *g, int N, int Fout, int K);
int cg_cheb_dw_pooled(const cg_graph_t *g, const float *dev_stack, const float *dev_g_pooled, const float *dev_y_pooled,
                      const uint8_t *dev_aux, float *dev_dW, float *dev_db, int N, int Fout, int K, void *workspace,
                      size_t workspace_bytes, void *stream);

/* ---- contractions over a caller-owned stack (row-partitioned filter, config C5) ------------------------- */
/* lib/models.py:218-223 on a basis the caller built slab by slab (cg_cheb_step + halo exchange): dev_stack holds K
 * slabs [R][F] that are slab_stride ELEMENTS apart (>= R*F; the partition keeps halo rows behind every slab).
 *   cg_cheb_contract, transposed == 0:  y[r, fo] = sum_{k,f}  stack_k[r, f]  W[f*K + k, fo]   (slabs of Fin columns)
 *   cg_cheb_contract, transposed != 0:  y[r, f]  = sum_{k,fo} stack_k[r, fo] W[f*K + k, fo]   (slabs of Fout columns:
 *                                        the input gradient from Z_k = T_k(L~^T) gy)
 *   cg_cheb_contract_dw:                dW[f*K + k, fo] = sum_r stack_k[r, f] gy[r, fo]
 * Tensor-core GEMMs (bf16 hi+mid split, fp32 accumulation); scratch from cg_cheb_contract_workspace_bytes.        */
size_t cg_cheb_contract_workspace_bytes(int64_t R, int Fin, int Fout, int K);
int cg_cheb_contract(const float *dev_stack, int64_t slab_stride, const float *dev_W, float *dev_y, int64_t R, int Fin,
                     int Fout, int K, int transposed, void *workspace, size_t workspace_bytes, void *stream);
int cg_cheb_contract_dw(const float *dev_stack, int64_t slab_stride, const float *dev_gy, float *dev_dW, int64_t R,
                        int Fin, int Fout, int K, void *workspace, size_t workspace_bytes, void *stream);

/* ---- dense head -------------------------------------------------------- */
/* fc layers of cgcnn (lib/models.py:268-274: relu(x W + b)) and their gradients: a general fp32 GEMM on the
 * tensor cores, C[M x N] = op(A)[M x K] . op(B)[K x N] (+ bias[N]) (relu), all matrices row-major;
 * transA != 0: A is stored [K][lda >= M]; transB != 0: B is stored [N][ldb >= K].  Same bf16 hi+mid split
 * with fp32 accumulation as the filter kernels (fp32-level accuracy).  Scratch (split-K partial tiles) is
 * sized by cg_gemm_f32_workspace_bytes and may be NULL when that returns 0.                              */
size_t cg_gemm_f32_workspace_bytes(int M, int N, int K);
int cg_gemm_f32(const float *dev_A, const float *dev_B, float *dev_C, int M, int N, int K, int transA,
                int transB, int lda, int ldb, int ldc, const float *dev_bias, int relu, void *dev_workspace,
                size_t workspace_bytes, void *stream);

/* ---- batched small GEMM (spectral `fourier` filter) ---------------------- */
/* lib/filter.py:11-27 == lib/models.py:129-144: between the two dense graph-Fourier transforms the reference applies one
 * Fout x Fin matrix per graph frequency (tf.matmul(W, x), W [M, Fout, Fin], batched over the M frequencies).
 * C[b] (m x n) = op(A[b]) (m x k) . op(B[b]) (k x n) for b < batch, all row-major fp32 (FFMA, fp32 accumulate);
 * transA != 0: A[b] is stored [k][lda >= m]; transB != 0: B[b] is stored [n][ldb >= k]; stride_* in elements between
 * consecutive batch entries.  The two transforms themselves are cg_gemm_f32 calls against U / U^T.               */
int cg_bmm_f32(const float *dev_A, const float *dev_B, float *dev_C, int batch, int m, int n, int k, int transA,
               int transB, int lda, int ldb, int ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c,
               void *stream);

/* ---- coarsening.perm_data ---------------------------------------------- */
/* lib/coarsening.py:219-240: out[:, i] = x[:, perm[i]] if perm[i] < M else 0.
 * dev_x [N, M], dev_perm [Mnew] int32, dev_out [N, Mnew] (float32 on device;
 * the host API converts to the reference's float64).                         */
int cg_perm_data(const float *dev_x, const int32_t *dev_perm, float *dev_out, int64_t N, int M,
                 int Mnew, void *stream);

/* ---- graph-conv LSTM gates --------------------------------------------- */
/* lib/gconv_lstm.py:185-215 (variant 0, 'fork': z = tan, o = tanh) and
 * lib/gconvRNN.py:189-213 (variant 1, 'standard': z = tanh, o = sigmoid).
 * dev_pre [R, 4H] holds the four summed filter outputs, gate order z,i,f,o
 * (R = N*M rows); dev_bias [4H]; dev_c [R, H].  Writes new c and new h.
 * Backward: given g_h, g_c (either may be NULL) produces g_pre [R,4H],
 * g_cprev [R,H]; d_bias [4H] is overwritten when non-NULL.                   */
int cg_lstm_gates_fwd(const float *dev_pre, const float *dev_bias, const float *dev_c, float *dev_new_c,
                      float *dev_new_h, int64_t R, int H, int variant, void *stream);
int cg_lstm_gates_bwd(const float *dev_pre, const float *dev_bias, const float *dev_c,
                      const float *dev_new_c, const float *dev_g_h, const float *dev_g_c,
                      float *dev_g_pre, float *dev_g_cprev, float *dev_d_bias, int64_t R, int H,
                      int variant, void *stream);
/* The same with the pre-activations given as TWO addends (x-path and h-path filters, lib/gconv_lstm.py:185-207:
 * filter(x, W?x) + filter(h, W?h)): dev_pre2 [R][4H] or NULL; the sum is formed inside the gate kernels, the
 * gradient dev_g_pre belongs to both addends.                                                                   */
int cg_lstm_gates2_fwd(const float *dev_pre, const float *dev_pre2, const float *dev_bias, const float *dev_c, float *dev_new_c,
                       float *dev_new_h, int64_t R, int H, int variant, void *stream);
int cg_lstm_gates2_bwd(const float *dev_pre, const float *dev_pre2, const float *dev_bias, const float *dev_c, const float *dev_new_c,
                       const float *dev_g_h, const float *dev_g_c, float *dev_g_pre, float *dev_g_cprev, float *dev_d_bias,
                       int64_t R, int H, int variant, void *stream);


/* ---- launch accounting / per-kernel timing (measurement support) -------- */
/* cg_launch_count: kernels launched by this library since load (all threads).
 * cg_profile_enable(1): bracket every launch with CUDA events on its own stream;
 * cg_profile_query(i, ...) returns the number of distinct kernel names and, for a
 * valid i, that kernel's accumulated device time (ms) and launch count (it
 * synchronises the recorded events first); cg_profile_reset clears the totals. */
int64_t cg_launch_count(void);
int cg_profile_enable(int on);
int cg_profile_reset(void);
int cg_profile_query(int index, char *name, int name_cap, double *total_ms, int64_t *count);

/* ---- tcgen05 self-test ---------------------------------------------------- */
/* One CTA computes D[128][N] = A . B^T with bf16 operands (round-to-nearest from the fp32
 * inputs) and fp32 accumulation in TMEM, with the operands staged in shared memory in the
 * K-major (x_mn = 0: A [128][Kd], B [N][Kd]) or MN-major (x_mn = 1: A [Kd][128], B [Kd][N])
 * canonical layout.  Pins the descriptor encodings the fused kernels rely on.            */
int cg_debug_umma_gemm(const float *dev_A, const float *dev_B, float *dev_D, int N, int Kd, int a_mn,
                       int b_mn, void *stream);
/* Same with Mr = 64 or 128 rows of A; dev_D [128][N] receives all 128 TMEM lanes. */
int cg_debug_umma_gemm_m(const float *dev_A, const float *dev_B, float *dev_D, int Mr, int N, int Kd,
                         int a_mn, int b_mn, void *stream);

/* A operand in tensor memory (tcgen05.st + tcgen05.mma with a TMEM A operand): dev_A [128][Kd], dev_B [N][Kd]. */
int cg_debug_umma_gemm_ts(const float *dev_A, const float *dev_B, float *dev_D, int N, int Kd, void *stream);

/* ---- sparse input batches (lib/graph_model.py:145-158: scipy batches densified on the host, then fed) ------------- */
/* dev_out [out_rows][M] fp32 <- CSR batch (dev_indptr [csr_rows + 1], dev_indices, dev_values: int32 / fp32, on the
 * device); duplicates are summed like scipy's toarray(); rows csr_rows .. out_rows-1 are zero (padded last batch).   */
int cg_csr_densify(const int32_t *dev_indptr, const int32_t *dev_indices, const float *dev_values, float *dev_out, int csr_rows,
                   int out_rows, int M, void *stream);

/* ---- precision of the tensor-core products (process-wide) --------------------------------------------------------
 * CG_PRECISION_FP32 (default): every fp32 operand is split into bf16 hi + mid and the product is formed as hi*hi + mid*hi
 * + hi*mid with fp32 accumulation (error <= 2^-16 relative: BASELINE's fp32 tolerance, rtol 1e-4).
 * CG_PRECISION_BF16: hi*hi only -- one tensor-core pass instead of three (BASELINE's bf16 tolerance, 2e-2).  Storage, the
 * recurrence and all accumulation stay fp32 in both modes.                                                          */
enum { CG_PRECISION_FP32 = 0, CG_PRECISION_BF16 = 1 };
int cg_set_precision(int mode);
int cg_get_precision(void);

/* ---- loss + optimiser tail of cgcnn (lib/graph_model.py:246-310, upstream cgcnn.loss / training) ----------------- */
/* Softmax cross-entropy averaged over the batch AND its gradient in one launch: dev_logits [N][C] fp32, dev_labels [N]
 * int64 -> *dev_loss (scalar) and dev_dlogits [N][C] = (softmax - onehot) / N (tf.nn.sparse_softmax_cross_entropy_with_logits
 * + tf.reduce_mean, lib/graph_model.py:250-252).  Deterministic (fixed summation order).                               */
int cg_softmax_xent(const float *dev_logits, const long long *dev_labels, float *dev_loss, float *dev_dlogits, int N, int C,
                    void *stream);
/* Momentum SGD on every variable in one launch (tf.train.MomentumOptimizer, lib/graph_model.py:291-295):
 * host_table: ntensors records {float *param; const float *grad; float *momentum_buffer; int64 numel} (32 bytes each, HOST
 * memory, device pointers inside; they travel as kernel arguments, 64 per launch); buffer = momentum * buffer + grad;
 * param -= lr * buffer.  max_numel sizes the grid.                                                                     */
int cg_sgd_momentum(const void *host_table, int ntensors, long long max_numel, float lr, float momentum, void *stream);
/* The same with the learning rate read from device memory when dev_lr is not NULL (a replayed CUDA graph bakes kernel
 * arguments; the deferred data-parallel update of dist.DeferredGradAllReducer applies the PREVIOUS step's rate).      */
int cg_sgd_momentum_dev(const void *host_table, int ntensors, long long max_numel, float lr, const float *dev_lr, float momentum,
                        void *stream);

/* Adam step of every variable in one launch (lib/graph_model.py:293: the fork trains the gconv-LSTM models with
 * tf.train.AdamOptimizer; arithmetic of torch.optim.Adam: m += (1-b1)(g-m), v = b2 v + (1-b2) g^2,
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)).  host_table: ntensors records of five 64-bit words
 * {param, grad, exp_avg, exp_avg_sq (device pointers), element count}; dev_state: two int32 on the device, zero before
 * the first step: [0] = steps done (advanced by the kernel, so a replayed CUDA graph keeps counting), [1] = scratch.  */
int cg_adam(const void *host_table, int ntensors, long long max_numel, float lr, float beta1, float beta2, float eps,
            int *dev_state, void *stream);

/* Debug aids of the fused recurrence kernels.  cg_debug_fused_trace: device buffer [K][10] of int64 that receives
 * clock64 stamps of CTA 0's second group (NULL switches it off).  cg_debug_fused_plan_info: the plan of the most
 * recent fused forward ([0..3]) and Clenshaw ([4..7]) launch: {row-block gather used, samples per group, items per
 * thread, dynamic shared-memory bytes}.                                                                        */
int cg_debug_fused_trace(long long *dev_buf);
int cg_debug_clenshaw_trace(long long *dev_buf);     /* same for the row-block Clenshaw kernel */
int cg_debug_fused_plan_info(int *info8);
/* Switch of the streaming GEMM (cg_gemm_stream.cu; environment CG_GEMM_STREAM, default on): on = 1 / 0 sets it, on < 0 only
 * queries; returns the previous setting.  With it off cg_gemm_f32 and the filter contractions take the pipelined
 * kernel (cg_gemm_pipe.cu); the two produce bit-identical results (same split, same MMA order).                     */
int cg_debug_gemm_stream(int on);

/* ---- host-side native loops of the coarsening -------------------------- */
/* lib/coarsening.py:119-165 (metis_one_level): greedy matching, float32
 * arithmetic in the reference's order, including the reference's row-table
 * off-by-one.  rr sorted; arrays are host pointers; cluster_id has N = rr[nnz-1]+1
 * entries.  Returns the number of clusters through *nclusters.               */
int cg_host_metis_one_level(int64_t nnz, const int64_t *rr, const int64_t *cc, const float *vv,
                            const int64_t *rid, int64_t n_rid, const float *weights,
                            int32_t *cluster_id, int64_t *nclusters);
/* Same loop with the match score vv*(1/w[v] + 1/w[u]) in float64: float64 adjacencies, and float32
 * ones (widened by the caller) under numpy 1.x value-based promotion, where the Python scalar 1.0
 * makes the reference's score float64 (lib/coarsening.py:150).                              */
int cg_host_metis_one_level_f64(int64_t nnz, const int64_t *rr, const int64_t *cc, const double *vv,
                                const int64_t *rid, int64_t n_rid, const double *weights,
                                int32_t *cluster_id, int64_t *nclusters);
/* lib/coarsening.py:179-204 (one level of compute_perm): children of the
 * vertices listed in `order` (length n_order) under `parent` (length n_parent),
 * singletons padded with fake ids starting at n_parent.  out has 2*n_order.  */
int cg_host_perm_level(const int64_t *parent, int64_t n_parent, const int64_t *order, int64_t n_order,
                       int64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* CNN_GRAPH_B200_H */
