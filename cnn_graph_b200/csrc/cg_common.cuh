// Shared declarations of the cnn_graph_b200 native library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/cnn_graph_b200.h"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void cg_set_error(const char *fmt, ...);

#define CG_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            cg_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,              \
                         cudaGetErrorString(_e));                                        \
            return CG_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define CG_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            cg_set_error(__VA_ARGS__);                                                   \
            return CG_ERR_ARG;                                                           \
        }                                                                                \
    } while (0)

#define CG_LAUNCH_CHECK() CG_CHECK_CUDA(cudaGetLastError())

// Counts one kernel launch and, when profiling is enabled (cg_profile_enable), brackets it
// with CUDA events on its stream.  Put one in the scope of every kernel launch.
class CgProfScope {
public:
    CgProfScope(const char *name, cudaStream_t stream);
    ~CgProfScope();
    CgProfScope(const CgProfScope &) = delete;
    CgProfScope &operator=(const CgProfScope &) = delete;

private:
    const char *name_;
    cudaStream_t stream_;
    cudaEvent_t a_, b_;
    bool active_;
};

// ---------------------------------------------------------------------------
// packed operator
// ---------------------------------------------------------------------------
// One orientation of the rescaled Laplacian (L~ or L~^T), resident in HBM.
//   CSR  : rowptr[M+1], col[nnz], val[nnz]          -- streaming kernels
//   ELL  : ell[M_pad * width] of {val, col} pairs, column-major (slot-major:
//          entry j of row m at ell[j * M_pad + m]); padding entries are
//          {0.0f, 0} so they can be applied unconditionally.  Built only when
//          it is small enough to live in shared memory.
struct CgCsr {
    int *rowptr = nullptr;
    int *col = nullptr;
    float *val = nullptr;
    int *order = nullptr;    // rows sorted by descending length (stable): work dealing of the fused kernels
    float2 *ell = nullptr;   // .x = value, .y = __int_as_float(col)
    int width = 0;           // max row length (ELL width)
    int m_pad = 0;           // rows padded to a multiple of 32
    // Row-block form for the fused kernels (built with the ELL): block b = rows 4b .. 4b+3; its entries are the UNION of
    // the four rows' columns, each with the four rows' weights (0 where a row has no such entry).  Neighbouring rows of
    // a coarsened grid share most of their neighbours, so a thread that owns a block gathers every needed row of the
    // signal once instead of once per referencing row.
    int nblk = 0;            // ceil(M / 4)
    int blk_total = 0;       // sum of union sizes
    int *blk_ptr = nullptr;  // [nblk + 1]
    int *blk_col = nullptr;  // [blk_total]
    float4 *blk_w = nullptr; // [blk_total]  weights of rows 4b .. 4b+3 for that column
    int *blk_ps = nullptr;   // [2 * nblk + 1] interleaved {ptr[b], split[b]}: inside a block the entries whose column lies in
                             // the block's own 128-row window come first ("near": [ptr, split)), the others after ("far")
    int *blk_order = nullptr;            // blocks by descending union size (stable)
    std::vector<int> blk_len_sorted;     // host copy: union sizes in that order (kernel planning)
};

struct cg_graph {
    int M = 0;
    int64_t nnz = 0;
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;   // max dynamic shared memory per block
    CgCsr fwd;               // L~
    CgCsr adj;               // L~^T
    bool onchip = false;     // ELL of both orientations fits the SMEM kernels
};

static inline const CgCsr &cg_side(const cg_graph *g, int transpose) { return transpose ? g->adj : g->fwd; }

// ---------------------------------------------------------------------------
// internal entry points shared between translation units
// ---------------------------------------------------------------------------
// Internal layout of a signal slab: S[m][c], c in [0, C), C = N * F, column
// c = n * F + f (one vertex row holds all signals' features contiguously).
// A Chebyshev stack is K such slabs: stack[k][m][c].

// stack[0] must already hold the input slab; fills stack[1..K-1].
int cg_run_basis(const cg_graph *g, int transpose, float *stack, int64_t C, int K, cudaStream_t s, int flags);

// Sample-major variant of the on-chip basis: x [N][M][F] -> stack [K][N][M][F] directly (rows r = n*M + m).
bool cg_basis_samples_supported(const cg_graph *g, int transpose, int N, int F);
int cg_run_basis_samples(const cg_graph *g, int transpose, const float *x, float *stack, int N, int F, int K,
                         cudaStream_t s);

// in[A][B][F] -> out[B][A][F]
int cg_run_permute_abf(const float *in, float *out, int64_t A, int64_t B, int F, cudaStream_t s);

// y[(n*M+m)][fo] = sum_{k,f} stack[k][m*N+n][f] * W[f*K+k][fo]      (contract)
int cg_run_contract(const float *stack, const float *W, float *y, int N, int M, int F, int Fout, int K,
                    bool w_transposed, bool sample_major, cudaStream_t s);

// dW[(a*K+k)][b] (or [(b*K+k)][a] when swap) = sum_{m,n} stack[k][m*N+n][a] * T[(n*M+m)][b]
size_t cg_stack_t_plain_workspace(int N, int M, int Fa, int Fb, int K, int sm_count);
int cg_run_stack_t_plain(const float *stack, const float *T, float *dW, int N, int M, int Fa, int Fb, int K,
                         bool swap, bool sample_major, float *workspace, int sm_count, cudaStream_t s);

// dW[...] = sum over `splits` partial results part[split][k*Fa + a][b] (deterministic, no atomics)
int cg_reduce_partials(const float *part, float *dW, int splits, int Fa, int Fb, int K, bool swap, cudaStream_t s);

// Tensor-core form of stack_t_plain (cg_dw_umma.cu): same contract, bf16x3 split with fp32 accumulation in TMEM.
bool cg_dw_umma_supported(int N, int M, int Fa, int Fb, int K, int sm_count, size_t smem_limit);
size_t cg_dw_umma_workspace(int N, int M, int Fa, int Fb, int K, int sm_count, size_t smem_limit);
//   sample_major: stack rows are r = n*M + m (the layout of T itself) instead of r = m*N + n.
int cg_run_dw_umma(const float *stack, const float *T, float *dW, int N, int M, int Fa, int Fb, int K, bool swap,
                   bool sample_major, float *workspace, int sm_count, size_t smem_limit, cudaStream_t s);

// Fused recurrence + contraction (cg_fused.cu): y[n,m,:] = sum_k (T_k(L) x)[n,m,:] W_k with the operator side
// chosen by `transpose`; w_transposed selects the dx form (W_k^T).  workspace: cg_fused_workspace bytes.
// W [..] -> per-k bf16 hi|mid B operands (K-major canonical, n = output feature, q = reduction feature)
int cg_pack_w(const float *W, unsigned char *wp, int Q, int Nn, int K, bool transposed, cudaStream_t s);
bool cg_fused_supported(const cg_graph *g, int transpose, int N, int Fin, int Fout, int K);
size_t cg_fused_workspace(int Fin, int Fout, int K);
//   stack_out (optional): the basis X_k, [K][N][M][Fin] (sample-major), for a later weight gradient.
//   stack_planes: stack_out receives the staged bf16 hi | mid planes instead ([2][K][Fin/8][N*M][8], same byte size),
//   the input format of cg_run_dw_planes.
int cg_run_fused(const cg_graph *g, int transpose, const float *x, const float *W, float *y, float *stack_out, int N,
                 int Fin, int Fout, int K, bool w_transposed, void *workspace, cudaStream_t s, bool stack_planes = false);

// One adjoint (Clenshaw) step on sample-major tensors with per-tensor row strides, batched over the samples, and the
// regrouping of the weight rows by k that the product G = gy W^T needs (cg_spmm.cu).
bool cg_clenshaw_step_supported(int F);
int cg_run_clenshaw_step(const cg_graph *g, int transpose, const float *G, int64_t sg, const float *X1, int64_t s1,
                         const float *X0, int64_t s0, float *out, int64_t so, int N, int F, float alpha, cudaStream_t s);
int cg_run_regroup_w(const float *W, float *Wp, int Fin, int Fout, int K, cudaStream_t s);

// Weight gradient of a first layer (Fa = 1) on the FFMA pipe, HBM-bound streaming (cg_dw_thin.cu).
bool cg_dw_thin_supported(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit);
size_t cg_dw_thin_workspace(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit);
int cg_run_dw_thin(const float *stack, const float *T, float *dW, long long R, int Fb, int K, float *workspace, int sm_count,
                   size_t smem_limit, cudaStream_t s);

// ... and its variant for a first layer followed by bias + relu + max pooling of 4: takes the gradient of the POOLED output
bool cg_dw_thin_pooled_supported(long long R, int Fb, int K, int sm_count, size_t smem_limit);
size_t cg_dw_thin_pooled_workspace(long long R, int Fb, int K, int sm_count, size_t smem_limit);
int cg_run_dw_thin_pooled(const float *stack, const float *gp, const float *yp, const unsigned char *aux, float *dW, float *db,
                          long long R, int Fb, int K, float *workspace, int sm_count, size_t smem_limit, cudaStream_t s);

// Weight gradient straight from that plane image (cg_dw_planes.cu): no conversion on the stack side.
bool cg_dw_planes_supported(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit);
size_t cg_dw_planes_workspace(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit);
int cg_run_dw_planes(const void *planes, const float *T, float *dW, long long R, int Fa, int Fb, int K, float *workspace,
                     int sm_count, size_t smem_limit, cudaStream_t s);

// Tensor-core contraction of a sample-major basis that lives in HBM (cg_contract_umma.cu).
bool cg_contract_umma_supported(int N, int M, int Fa, int J, int K, size_t smem_limit);
//   yp != NULL: pooled epilogue (relu(y + bias) max-pooled over groups of 4 rows -> yp, aux [N*M/4][J]; y is not written)
int cg_run_contract_umma(const float *stack, const float *W, float *y, int N, int M, int Fa, int J, int K, int sm_count,
                         size_t smem_limit, cudaStream_t s, const float *bias = nullptr, float *yp = nullptr,
                         unsigned char *aux = nullptr);

// General fp32 GEMM on the tensor cores (cg_gemm_umma.cu) with optional K blocking of the operands:
//   a_kblk > 0: A element (m, q) at A[(q / a_kblk) * a_kbs + m * lda + q % a_kblk]   (a Chebyshev stack [K][R][F])
//   b_kblk > 0: B row of q is (q / b_kblk) * b_shi + (q % b_kblk) * b_slo            (W rows f*K + k for q = k*F + f)
size_t cg_gemm_workspace(int M, int N, int K);
int cg_run_gemm(const float *A, const float *B, float *C, int M, int N, int K, int transA, int transB, int lda, int ldb,
                int ldc, const float *bias, int relu, int a_kblk, long long a_kbs, int b_kblk, int b_shi, int b_slo,
                void *workspace, size_t workspace_bytes, cudaStream_t s);
//   a_mblk > 0 (transA only): row m of op(A) starts at element (m / a_mblk) * a_mbs + m % a_mblk -- all k slabs of a
//   Chebyshev stack as one transposed operand.  Only the pipelined kernel implements it: check cg_gemm_mblocked_ok.
bool cg_gemm_mblocked_ok(const float *A, const float *B, int M, int N, int K, int lda, int ldb, int a_mblk, long long a_mbs);
int cg_run_gemm_mblocked(const float *A, const float *B, float *C, int M, int N, int K, int lda, int ldb, int ldc, int a_mblk,
                         long long a_mbs, void *workspace, size_t workspace_bytes, cudaStream_t s);
// dW[f*K + k][:] = T[k*Fin + f][:]  (rows of a stacked gradient back into the order of the filter weights)
int cg_run_regroup_dw(const float *T, float *dW, int Fin, int Fout, int K, cudaStream_t s);

// Pipelined fast path of the same GEMM (cg_gemm_pipe.cu): 16-byte aligned operands, power-of-two K blocks
bool cg_gemm_pipe_eligible(const float *A, const float *B, int M, int N, int K, int lda, int ldb, int transA, int transB,
                           int a_kblk, long long a_kbs, int b_kblk, int a_mblk = 0, long long a_mbs = 0);
size_t cg_gemm_pipe_workspace(int M, int N, int K, int sm_count);
int cg_run_gemm_pipe(const float *A, const float *B, float *C, int M, int N, int K, int transA, int transB, int lda,
                     int ldb, int ldc, const float *bias, int relu, int a_kblk, long long a_kbs, int b_kblk, int b_shi,
                     int b_slo, void *workspace, size_t workspace_bytes, int sm_count, cudaStream_t s, int a_mblk = 0,
                     long long a_mbs = 0);
int cg_gemm_reduce(const float *part, const float *bias, float *C, int M, int N, int ldc, int split, int relu, cudaStream_t s);

// Short reductions (K * F <= 16) on the FFMA pipe, bound by the stream of y / T (cg_thin.cu); either stack layout.
// CG_THIN=0 disables them.
bool cg_thin_supported(int N, int M, int F, int J, int K);
size_t cg_thin_dw_workspace(int N, int M, int F, int J, int K, int sm_count);
int cg_run_thin_contract(const float *stack, const float *W, float *y, int N, int M, int F, int J, int K, bool sample_major,
                         int sm_count, cudaStream_t s);
int cg_run_thin_dw(const float *stack, const float *T, float *dW, int N, int M, int F, int J, int K, bool sample_major,
                   float *workspace, int sm_count, cudaStream_t s);

// Streaming form for a large left operand (cg_gemm_stream.cu): B split once per call into packed bf16 planes (workspace),
// the fp32 A tiles fetched by bulk copies of one producer warp; K a multiple of 32.  CG_GEMM_STREAM=0 disables it.
//   cg_run_gemm_stream returns CG_TRY_NEXT (nothing launched) when the tensor map of A cannot be encoded: take the next kernel
constexpr int CG_TRY_NEXT = -77;
size_t cg_gemm_stream_workspace(int M, int N, int K, int sm_count);
bool cg_gemm_stream_eligible(const float *A, int M, int N, int K, int lda, int transA, int transB, int a_kblk, long long a_kbs,
                             int b_kblk, int a_mblk, long long a_mbs, size_t workspace_bytes, int sm_count);
int cg_run_gemm_stream(const float *A, const float *B, float *C, int M, int N, int K, int transA, int transB, int lda, int ldb,
                       int ldc, const float *bias, int relu, int a_kblk, long long a_kbs, int b_kblk, int b_shi, int b_slo,
                       void *workspace, size_t workspace_bytes, int sm_count, cudaStream_t s, int a_mblk, long long a_mbs);

// Input gradient by the adjoint (Clenshaw) recurrence with gy resident in tensor memory (cg_clenshaw.cu).
bool cg_clenshaw_supported(const cg_graph *g, int N, int Fin, int Fout, int K);
int cg_run_clenshaw(const cg_graph *g, const float *gy, const float *W, float *dx, int N, int Fin, int Fout, int K,
                    void *workspace, cudaStream_t s);

// SMs the persistent kernels may occupy: all of them minus CG_SM_RESERVE (environment; default 0).  Data-parallel runs
// reserve a few SMs so that NCCL's all-reduce kernels can start while a persistent backward kernel is resident
// (one CTA per SM with ~200 KB of shared memory leaves no room for another CTA).
int cg_sm_budget(int device);

// MMA passes per tensor-core product: 3 = fp32-equivalent (bf16 hi*hi + mid*hi + hi*mid, the default), 1 = single-pass bf16
// (hi*hi only; BASELINE's 2e-2 tolerance).  Process-wide, set through cg_set_precision().
int cg_mma_passes();

static inline int64_t cg_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t cg_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
