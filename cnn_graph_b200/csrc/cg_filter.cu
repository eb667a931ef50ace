// Chebyshev filter forward / backward (chebyshev5, chebyshev2, cheby_conv):
// lib/models.py:161-224, lib/graph_conv.py:113-176, lib/filter.py:45-95.
//
// Forward   y  = contract(basis(L~, x), W)
// Backward  Z  = basis(L~^T, gy)                      (T_k(L~)^T = T_k(L~^T))
//           dx = contract(Z, W^T-view)                dx[n,m,fin] = sum_{k,fo} Z_k[n,m,fo] W[fin*K+k,fo]
//           dW = x^T Z_k   (stack_t_plain, swapped)   dW[fin*K+k,fo] = sum_{n,m} x[n,m,fin] Z_k[n,m,fo]
//   without dx (first layer, chebyshev2):  dW = X_k^T gy with X = basis(L~, x).
// The forward stack is never saved: backward needs only x, gy and W.
#include "cg_common.cuh"

static size_t stack_bytes(const cg_graph *g, int N, int F, int K) {
    return cg_align_up(sizeof(float) * (size_t)K * g->M * N * F, 256);
}

extern "C" size_t cg_cheb_filter_fwd_workspace_bytes(const cg_graph_t *g, int N, int Fin, int Fout, int K, int flags) {
    (void)Fout;
    (void)flags;
    if (!g || K <= 1) return 0;
    return stack_bytes(g, N, Fin, K);
}

extern "C" size_t cg_cheb_filter_bwd_workspace_bytes(const cg_graph_t *g, int N, int Fin, int Fout, int K,
                                                     int need_dx, int flags) {
    (void)flags;
    if (!g) return 0;
    if (need_dx)
        return stack_bytes(g, N, Fout, K) + cg_stack_t_plain_workspace(N, g->M, Fout, Fin, K, g->sm_count);
    return stack_bytes(g, N, Fin, K) + cg_stack_t_plain_workspace(N, g->M, Fin, Fout, K, g->sm_count);
}

static int check_dims(const char *who, const cg_graph *g, int N, int Fin, int Fout, int K) {
    CG_REQUIRE(g != nullptr, "%s: graph handle is NULL", who);
    CG_REQUIRE(N >= 0 && Fin > 0 && Fout > 0 && K >= 1, "%s: bad dims N=%d Fin=%d Fout=%d K=%d", who, N, Fin, Fout, K);
    CG_REQUIRE((int64_t)g->M * N * (int64_t)((Fin > Fout ? Fin : Fout)) < (int64_t)1 << 40, "%s: problem too large", who);
    return CG_OK;
}

extern "C" int cg_cheb_filter_fwd(const cg_graph_t *g, const float *x, const float *W, float *y, int N, int Fin,
                                  int Fout, int K, void *workspace, size_t workspace_bytes, int flags, void *stream) {
    int rc = check_dims("cg_cheb_filter_fwd", g, N, Fin, Fout, K);
    if (rc != CG_OK) return rc;
    if (N == 0) return CG_OK;
    CG_REQUIRE(x && W && y, "cg_cheb_filter_fwd: NULL tensor");
    cudaStream_t s = (cudaStream_t)stream;
    const int M = g->M;
    if (K == 1)   // y = x W: a per-vertex linear map (lib/models.py:205-206 with no SpMM)
        return cg_run_contract(x, W, y, 1, N * M, Fin, Fout, 1, false, s);
    const size_t need = cg_cheb_filter_fwd_workspace_bytes(g, N, Fin, Fout, K, flags);
    if (workspace == nullptr || workspace_bytes < need) {
        cg_set_error("cg_cheb_filter_fwd: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    float *stack = reinterpret_cast<float *>(workspace);
    rc = cg_run_permute_abf(x, stack, N, M, Fin, s);                          // [N][M][F] -> [M][N][F]
    if (rc == CG_OK) rc = cg_run_basis(g, 0, stack, (int64_t)N * Fin, K, s, flags);
    if (rc == CG_OK) rc = cg_run_contract(stack, W, y, N, M, Fin, Fout, K, false, s);
    return rc;
}

extern "C" int cg_cheb_filter_bwd(const cg_graph_t *g, const float *x, const float *W, const float *gy, float *dx,
                                  float *dW, int N, int Fin, int Fout, int K, void *workspace, size_t workspace_bytes,
                                  int flags, void *stream) {
    int rc = check_dims("cg_cheb_filter_bwd", g, N, Fin, Fout, K);
    if (rc != CG_OK) return rc;
    CG_REQUIRE(dW != nullptr, "cg_cheb_filter_bwd: dW is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const int M = g->M;
    if (N == 0) {
        CG_CHECK_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)Fin * K * Fout, s));
        return CG_OK;
    }
    CG_REQUIRE(x && W && gy, "cg_cheb_filter_bwd: NULL tensor");
    const int need_dx = dx != nullptr;
    const size_t need = cg_cheb_filter_bwd_workspace_bytes(g, N, Fin, Fout, K, need_dx, flags);
    if (workspace == nullptr || workspace_bytes < need) {
        cg_set_error("cg_cheb_filter_bwd: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    float *stack = reinterpret_cast<float *>(workspace);
    if (need_dx) {
        float *part = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + stack_bytes(g, N, Fout, K));
        rc = cg_run_permute_abf(gy, stack, N, M, Fout, s);
        if (rc == CG_OK) rc = cg_run_basis(g, 1, stack, (int64_t)N * Fout, K, s, flags);
        if (rc == CG_OK) rc = cg_run_contract(stack, W, dx, N, M, Fout, Fin, K, true, s);
        if (rc == CG_OK) rc = cg_run_stack_t_plain(stack, x, dW, N, M, Fout, Fin, K, true, part, g->sm_count, s);
    } else {
        float *part = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + stack_bytes(g, N, Fin, K));
        rc = cg_run_permute_abf(x, stack, N, M, Fin, s);
        if (rc == CG_OK) rc = cg_run_basis(g, 0, stack, (int64_t)N * Fin, K, s, flags);
        if (rc == CG_OK) rc = cg_run_stack_t_plain(stack, gy, dW, N, M, Fin, Fout, K, false, part, g->sm_count, s);
    }
    return rc;
}
