// Chebyshev filter forward / backward (chebyshev5, chebyshev2, cheby_conv):
// lib/models.py:161-224, lib/graph_conv.py:113-176, lib/filter.py:45-95.
//
// Forward   y  = contract(basis(L~, x), W)
// Backward  Z  = basis(L~^T, gy)                      (T_k(L~)^T = T_k(L~^T))
//           dx = contract(Z, W^T-view)                dx[n,m,fin] = sum_{k,fo} Z_k[n,m,fo] W[fin*K+k,fo]
//           dW = x^T Z_k   (stack_t_plain, swapped)   dW[fin*K+k,fo] = sum_{n,m} x[n,m,fin] Z_k[n,m,fo]
//   without dx (first layer, chebyshev2):  dW = X_k^T gy with X = basis(L~, x).
// The forward stack is never saved: backward needs only x, gy and W.
#include <algorithm>

#include "cg_common.cuh"

static size_t stack_bytes(const cg_graph *g, int N, int F, int K) {
    return cg_align_up(sizeof(float) * (size_t)K * g->M * N * F, 256);
}

extern "C" size_t cg_cheb_filter_fwd_workspace_bytes(const cg_graph_t *g, int N, int Fin, int Fout, int K, int flags) {
    (void)flags;
    if (!g) return 0;
    const int64_t R = (int64_t)N * g->M;
    const size_t gws = R > 0 && R < (int64_t)INT32_MAX ? cg_align_up(cg_gemm_workspace((int)R, Fout, K * Fin), 256) : 0;
    return (K > 1 ? stack_bytes(g, N, Fin, K) : 0) + gws + cg_fused_workspace(Fin, Fout, K);
}

// dW = stack^T T for all k in ONE tensor-core GEMM: the K slabs [R][Fa] form a transposed operand whose rows
// m = k*Fa + a are blocked by Fa (cg_run_gemm_mblocked); the [K*Fa][Fb] result is regrouped into W's row order.
// `scratch` holds K*Fa*Fb floats followed by the GEMM's split-K partials.
static size_t dw_allk_workspace(int64_t R, int Fa, int Fb, int K) {
    if (R <= 0 || R >= (int64_t)INT32_MAX) return 0;
    return cg_align_up(sizeof(float) * (size_t)K * Fa * Fb, 256) + cg_gemm_workspace(K * Fa, Fb, (int)R);
}
static bool dw_allk_ok(const float *stack, int64_t slab_stride, const float *T, int64_t R, int Fa, int Fb, int K) {
    if (R <= 0 || R >= (int64_t)INT32_MAX || (Fa & (Fa - 1)) != 0) return false;
    return cg_gemm_mblocked_ok(stack, T, K * Fa, Fb, (int)R, Fa, Fb, Fa, slab_stride);
}
static int run_dw_allk(const float *stack, int64_t slab_stride, const float *T, float *dW, int64_t R, int Fa, int Fb, int K,
                       void *scratch, cudaStream_t s) {
    float *tmp = reinterpret_cast<float *>(scratch);
    void *gws = reinterpret_cast<char *>(scratch) + cg_align_up(sizeof(float) * (size_t)K * Fa * Fb, 256);
    int rc = cg_run_gemm_mblocked(stack, T, tmp, K * Fa, Fb, (int)R, Fa, Fb, Fb, Fa, slab_stride, gws,
                                  cg_gemm_workspace(K * Fa, Fb, (int)R), s);
    if (rc == CG_OK) rc = cg_run_regroup_dw(tmp, dW, Fa, Fb, K, s);
    return rc;
}

// stack^T x plain: tensor-core kernel when the shape allows it, FFMA kernel otherwise
static size_t dw_workspace(const cg_graph *g, int N, int Fa, int Fb, int K) {
    const size_t a = cg_stack_t_plain_workspace(N, g->M, Fa, Fb, K, g->sm_count);
    const size_t b = cg_dw_umma_workspace(N, g->M, Fa, Fb, K, g->sm_count, g->smem_optin);
    const int64_t R = (int64_t)N * g->M;
    const size_t c = R < (int64_t)INT32_MAX ? cg_gemm_workspace(Fa, Fb, (int)R) : 0;       // per-k GEMM, split over R
    const size_t d = cg_dw_planes_workspace(R, Fa, Fb, K, g->sm_count, g->smem_optin);
    const size_t e = cg_dw_thin_workspace(R, Fa, Fb, K, g->sm_count, g->smem_optin);
    const size_t f = std::max(dw_allk_workspace(R, Fa, Fb, K), dw_allk_workspace(R, Fb, Fa, K));
    const size_t t = cg_thin_dw_workspace(N, g->M, Fa, Fb, K, g->sm_count);
    return std::max(std::max(std::max(a, t), std::max(d, std::max(e, f))), std::max(b, c));
}

// stack^T x plain for a scalar-signal stack (first layers): streaming FFMA kernel
static bool thin_ok(const cg_graph *g, const float *stack, const float *T, float *part, int N, int Fa, int Fb, int K, bool swap,
                    bool sample_major, int flags) {
    return !swap && sample_major && !(flags & (CG_FILTER_NO_FUSED | CG_FILTER_FORCE_STREAMING)) &&
           ((((uintptr_t)stack | (uintptr_t)T | (uintptr_t)part) & 15) == 0) &&
           cg_dw_thin_supported((long long)N * g->M, Fa, Fb, K, g->sm_count, g->smem_optin);
}

static int run_dw(const cg_graph *g, const float *stack, const float *T, float *dW, int N, int Fa, int Fb, int K,
                  bool swap, bool sample_major, float *part, int flags, cudaStream_t s) {
    if (thin_ok(g, stack, T, part, N, Fa, Fb, K, swap, sample_major, flags))
        return cg_run_dw_thin(stack, T, dW, (long long)N * g->M, Fb, K, part, g->sm_count, g->smem_optin, s);
    if (!swap && !(flags & CG_FILTER_NO_FUSED) && cg_thin_supported(N, g->M, Fa, Fb, K))      // K * Fa <= 16, either stack layout
        return cg_run_thin_dw(stack, T, dW, N, g->M, Fa, Fb, K, sample_major, part, g->sm_count, s);
    const bool tc = !(flags & (CG_FILTER_NO_FUSED | CG_FILTER_FORCE_STREAMING)) &&
                    cg_dw_umma_supported(N, g->M, Fa, Fb, K, g->sm_count, g->smem_optin) &&
                    ((((uintptr_t)stack | (uintptr_t)T | (uintptr_t)part) & 15) == 0);
    if (tc)
        return cg_run_dw_umma(stack, T, dW, N, g->M, Fa, Fb, K, swap, sample_major, part, g->sm_count, g->smem_optin, s);
    const int64_t R = (int64_t)N * g->M;
    if (sample_major && !(flags & (CG_FILTER_NO_FUSED | CG_FILTER_FORCE_STREAMING)) && R < (int64_t)INT32_MAX) {
        // shapes the TMEM-resident kernel does not take (Fb > 256, ...)
        if (!swap && dw_allk_ok(stack, R * Fa, T, R, Fa, Fb, K)) return run_dw_allk(stack, R * Fa, T, dW, R, Fa, Fb, K, part, s);
        // ... or one tensor-core GEMM per k,
        //   dW_k = stack_k^T T   (rows a*K + k)        or, swapped,   dW_k = T^T stack_k   (rows b*K + k)
        const size_t ws = cg_gemm_workspace(swap ? Fb : Fa, swap ? Fa : Fb, (int)R);
        for (int k = 0; k < K; ++k) {
            const float *sk = stack + (size_t)k * R * Fa;
            int rc;
            if (!swap)
                rc = cg_run_gemm(sk, T, dW + (size_t)k * Fb, Fa, Fb, (int)R, 1, 0, Fa, Fb, K * Fb, nullptr, 0, 0, 0, 0, 0, 0,
                                 part, ws, s);
            else
                rc = cg_run_gemm(T, sk, dW + (size_t)k * Fa, Fb, Fa, (int)R, 1, 0, Fb, Fa, K * Fa, nullptr, 0, 0, 0, 0, 0, 0,
                                 part, ws, s);
            if (rc != CG_OK) return rc;
        }
        return CG_OK;
    }
    return cg_run_stack_t_plain(stack, T, dW, N, g->M, Fa, Fb, K, swap, sample_major, part, g->sm_count, s);
}

// basis straight from / to the sample-major layout (no permute) when the on-chip kernel takes the operator
static bool samples_ok(const cg_graph *g, int transpose, int N, int F, int flags, const void *a, const void *b) {
    if (flags & CG_FILTER_FORCE_STREAMING) return false;
    return ((((uintptr_t)a | (uintptr_t)b) & 15) == 0) && cg_basis_samples_supported(g, transpose, N, F);
}

extern "C" size_t cg_cheb_filter_bwd_workspace_bytes(const cg_graph_t *g, int N, int Fin, int Fout, int K,
                                                     int need_dx, int flags) {
    (void)flags;
    if (!g) return 0;
    // either the Z-stack (width Fout) for dx and dW, or the X-stack (width Fin) for dW next to a fused dx
    const size_t wide = stack_bytes(g, N, Fout > Fin ? Fout : Fin, K);
    const size_t part_a = dw_workspace(g, N, Fout, Fin, K);
    const size_t part_b = dw_workspace(g, N, Fin, Fout, K);
    (void)need_dx;
    const int64_t R = (int64_t)N * g->M;
    size_t gdx = R > 0 && R < (int64_t)INT32_MAX ? cg_gemm_workspace((int)R, Fin, K * Fout) : 0;
    if (R > 0 && R < (int64_t)INT32_MAX) gdx = std::max(gdx, cg_gemm_workspace((int)R, K * Fin, Fout));      // G = gy W^T
    return wide + cg_align_up(std::max(gdx, std::max(part_a, part_b)), 256) + cg_fused_workspace(Fin, Fout, K);
}

// Unfused adjoint recurrence for dx (see cg_cheb_filter_bwd_ex): pays off when dx is narrower than gy; needs
// room for G, two step buffers and the regrouped weights inside the stack region of the workspace.
static bool unfused_clenshaw_ok(const cg_graph *g, int N, int Fin, int Fout, int K, int flags, const void *gy, const void *dx) {
    if (flags & (CG_FILTER_NO_FUSED | CG_FILTER_FORCE_STREAMING | CG_FILTER_NO_CLENSHAW)) return false;
    const int64_t R = (int64_t)N * g->M;
    if (K < 2 || Fin >= Fout || R >= (int64_t)INT32_MAX || N > 65535 || !cg_clenshaw_step_supported(Fin)) return false;
    if (((((uintptr_t)gy) | ((uintptr_t)dx)) & 15) != 0 || Fout % 4 != 0) return false;
    const size_t need = sizeof(float) * ((size_t)(K + 2) * R * Fin + (size_t)K * Fin * Fout);
    return need <= sizeof(float) * (size_t)K * R * Fout;      // the stack region holds K slabs of width max(Fin, Fout)
}

static bool aligned16(const void *a, const void *b, const void *c) {
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c)) & 15) == 0;
}

// fused kernel wanted and possible?  (CG_FILTER_FORCE_FUSED turns "not possible" into an error)
static int want_fused(const char *who, const cg_graph *g, int transpose, int N, int Fin, int Fout, int K, int flags,
                      bool *use) {
    *use = false;
    if (flags & (CG_FILTER_NO_FUSED | CG_FILTER_FORCE_STREAMING)) return CG_OK;
    *use = cg_fused_supported(g, transpose, N, Fin, Fout, K);
    if (!*use && (flags & CG_FILTER_FORCE_FUSED)) {
        cg_set_error("%s: fused kernel requested but the shape is not supported (M=%d Fin=%d Fout=%d K=%d)", who, g->M,
                     Fin, Fout, K);
        return CG_ERR_ARG;
    }
    return CG_OK;
}

static int check_dims(const char *who, const cg_graph *g, int N, int Fin, int Fout, int K) {
    CG_REQUIRE(g != nullptr, "%s: graph handle is NULL", who);
    CG_REQUIRE(N >= 0 && Fin > 0 && Fout > 0 && K >= 1, "%s: bad dims N=%d Fin=%d Fout=%d K=%d", who, N, Fin, Fout, K);
    CG_REQUIRE((int64_t)g->M * N * (int64_t)((Fin > Fout ? Fin : Fout)) < (int64_t)1 << 40, "%s: problem too large", who);
    return CG_OK;
}

// The forward pass can leave the basis X_k behind for the weight gradient ([K][N][M][Fin], sample-major) when
// both the fused kernel and the tensor-core dW kernel take the shape.
static bool can_save_stack(const cg_graph *g, int N, int Fin, int Fout, int K, int flags) {
    if (N <= 0 || K < 2 || (flags & CG_FILTER_FORCE_STREAMING)) return false;
    if (!(flags & CG_FILTER_NO_FUSED) && cg_fused_supported(g, 0, N, Fin, Fout, K)) return true;
    return cg_basis_samples_supported(g, 0, N, Fin);
}

// The saved basis can be the fused kernel's own bf16 operand planes (CG_FILTER_STACK_PLANES) when the fused
// forward kernel runs and the plane-streaming dW kernel takes the shape.
static bool planes_ok(const cg_graph *g, int N, int Fin, int Fout, int K, int flags) {
    if (N <= 0 || K < 2 || (flags & (CG_FILTER_FORCE_STREAMING | CG_FILTER_NO_FUSED))) return false;
    return Fin % 8 == 0 && cg_fused_supported(g, 0, N, Fin, Fout, K) &&
           cg_dw_planes_supported((long long)N * g->M, Fin, Fout, K, g->sm_count, g->smem_optin);
}

extern "C" int cg_cheb_filter_stack_planes(const cg_graph_t *g, int N, int Fin, int Fout, int K, int flags) {
    if (!g || Fin <= 0 || Fout <= 0 || K < 1) return 0;
    return planes_ok(g, N, Fin, Fout, K, flags) ? 1 : 0;
}

extern "C" size_t cg_cheb_filter_stack_bytes(const cg_graph_t *g, int N, int Fin, int Fout, int K, int flags) {
    if (!g || Fin <= 0 || Fout <= 0 || K < 1) return 0;
    if ((flags & CG_FILTER_STACK_PLANES) && planes_ok(g, N, Fin, Fout, K, flags))      // rows padded to whole chunks of 128
        return sizeof(float) * (size_t)K * Fin * (size_t)(cg_ceil_div((int64_t)N * g->M, 128) * 128);
    return can_save_stack(g, N, Fin, Fout, K, flags) ? sizeof(float) * (size_t)K * N * g->M * Fin : 0;
}

extern "C" int cg_cheb_filter_fwd(const cg_graph_t *g, const float *x, const float *W, float *y, int N, int Fin,
                                  int Fout, int K, void *workspace, size_t workspace_bytes, int flags, void *stream) {
    return cg_cheb_filter_fwd_ex(g, x, W, y, nullptr, N, Fin, Fout, K, workspace, workspace_bytes, flags, stream);
}

extern "C" int cg_cheb_filter_fwd_ex(const cg_graph_t *g, const float *x, const float *W, float *y, float *stack_out,
                                     int N, int Fin, int Fout, int K, void *workspace, size_t workspace_bytes,
                                     int flags, void *stream) {
    int rc = check_dims("cg_cheb_filter_fwd", g, N, Fin, Fout, K);
    if (rc != CG_OK) return rc;
    if (N == 0) return CG_OK;
    CG_REQUIRE(stack_out == nullptr || can_save_stack(g, N, Fin, Fout, K, flags),
               "cg_cheb_filter_fwd_ex: this shape cannot save the basis (cg_cheb_filter_stack_bytes returned 0)");
    CG_REQUIRE(x && W && y, "cg_cheb_filter_fwd: NULL tensor");
    cudaStream_t s = (cudaStream_t)stream;
    const int M = g->M;
    const size_t need = cg_cheb_filter_fwd_workspace_bytes(g, N, Fin, Fout, K, flags);
    if (workspace == nullptr || workspace_bytes < need) {
        cg_set_error("cg_cheb_filter_fwd: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    bool fused = false;
    rc = want_fused("cg_cheb_filter_fwd", g, 0, N, Fin, Fout, K, flags, &fused);
    if (rc != CG_OK) return rc;
    if (fused && !aligned16(x, y, stack_out)) {          // bulk copies and 128-bit accesses need 16-byte alignment
        CG_REQUIRE(!(flags & CG_FILTER_FORCE_FUSED), "cg_cheb_filter_fwd: fused kernel needs 16-byte aligned tensors");
        fused = false;
    }
    const bool planes = stack_out != nullptr && (flags & CG_FILTER_STACK_PLANES);
    CG_REQUIRE(!planes || (fused && planes_ok(g, N, Fin, Fout, K, flags)),
               "cg_cheb_filter_fwd_ex: CG_FILTER_STACK_PLANES needs cg_cheb_filter_stack_planes() == 1 and aligned tensors");
    if (fused) {
        void *wpack = reinterpret_cast<char *>(workspace) + (need - cg_fused_workspace(Fin, Fout, K));
        return cg_run_fused(g, 0, x, W, y, stack_out, N, Fin, Fout, K, false, wpack, s, planes);
    }
    if (K == 1)   // y = x W: a per-vertex linear map (lib/models.py:205-206 with no SpMM)
        return cg_run_contract(x, W, y, 1, N * M, Fin, Fout, 1, false, false, s);
    float *stack = reinterpret_cast<float *>(workspace);
    if (samples_ok(g, 0, N, Fin, flags, x, stack_out ? stack_out : stack)) {
        float *st = stack_out ? stack_out : stack;                            // [K][N][M][F], rows in y's order
        rc = cg_run_basis_samples(g, 0, x, st, N, Fin, K, s);
        if (rc != CG_OK) return rc;
        if (!(flags & CG_FILTER_NO_FUSED) && (((uintptr_t)y) & 15) == 0 &&
            cg_contract_umma_supported(N, M, Fin, Fout, K, g->smem_optin))      // also the arithmetic of the fused first layer
            return cg_run_contract_umma(st, W, y, N, M, Fin, Fout, K, g->sm_count, g->smem_optin, s);
        if (!(flags & CG_FILTER_NO_FUSED) && cg_thin_supported(N, M, Fin, Fout, K))     // K * Fin <= 16: bound by the stream of y
            return cg_run_thin_contract(st, W, y, N, M, Fin, Fout, K, true, g->sm_count, s);
        if (!(flags & CG_FILTER_NO_FUSED) && (int64_t)N * M < (int64_t)INT32_MAX) {
            // wide outputs (Fout > 128, e.g. the 4H gates of the gconv-LSTM): general tensor-core GEMM over the
            // K-blocked basis, q = k*Fin + f  <->  W row f*K + k
            const int64_t R = (int64_t)N * M;
            void *gws = reinterpret_cast<char *>(workspace) + stack_bytes(g, N, Fin, K);
            return cg_run_gemm(st, W, y, (int)R, Fout, K * Fin, 0, 0, Fin, Fout, Fout, nullptr, 0, Fin, (long long)R * Fin,
                               Fin, 1, K, gws, cg_gemm_workspace((int)R, Fout, K * Fin), s);
        }
        return cg_run_contract(st, W, y, N, M, Fin, Fout, K, false, true, s);
    }
    CG_REQUIRE(stack_out == nullptr, "cg_cheb_filter_fwd_ex: this call cannot save the basis (unaligned tensors)");
    rc = cg_run_permute_abf(x, stack, N, M, Fin, s);                          // [N][M][F] -> [M][N][F]
    if (rc == CG_OK) rc = cg_run_basis(g, 0, stack, (int64_t)N * Fin, K, s, flags);
    if (rc == CG_OK) rc = cg_run_contract(stack, W, y, N, M, Fin, Fout, K, false, false, s);
    return rc;
}

extern "C" int cg_cheb_filter_bwd(const cg_graph_t *g, const float *x, const float *W, const float *gy, float *dx,
                                  float *dW, int N, int Fin, int Fout, int K, void *workspace, size_t workspace_bytes,
                                  int flags, void *stream) {
    return cg_cheb_filter_bwd_ex(g, x, W, gy, nullptr, dx, dW, N, Fin, Fout, K, workspace, workspace_bytes, flags,
                                 stream);
}

extern "C" int cg_cheb_filter_bwd_ex(const cg_graph_t *g, const float *x, const float *W, const float *gy,
                                     const float *saved_stack, float *dx, float *dW, int N, int Fin, int Fout, int K,
                                     void *workspace, size_t workspace_bytes, int flags, void *stream) {
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
    bool have_dW = false;
    if (saved_stack != nullptr) {
        // dW[fin*K+k, fo] = sum_{n,m} X_k[n,m,fin] gy[n,m,fo] straight from the basis the forward pass left behind
        CG_REQUIRE(can_save_stack(g, N, Fin, Fout, K, flags), "cg_cheb_filter_bwd_ex: saved stack given for a shape that cannot save one");
        float *part = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + stack_bytes(g, N, Fin, K));
        if (flags & CG_FILTER_STACK_PLANES) {
            CG_REQUIRE(planes_ok(g, N, Fin, Fout, K, flags) && (((uintptr_t)gy) & 15) == 0,
                       "cg_cheb_filter_bwd_ex: CG_FILTER_STACK_PLANES given for a shape without plane support");
            rc = cg_run_dw_planes(saved_stack, gy, dW, (long long)N * M, Fin, Fout, K, part, g->sm_count, g->smem_optin, s);
        } else {
            rc = run_dw(g, saved_stack, gy, dW, N, Fin, Fout, K, false, true, part, flags, s);
        }
        if (rc != CG_OK) return rc;
        have_dW = true;
    }
    if (need_dx) {
        bool fused = false;
        const bool allow = !(flags & (CG_FILTER_NO_FUSED | CG_FILTER_FORCE_STREAMING));
        const bool al = aligned16(gy, dx, nullptr);
        const bool clenshaw = allow && al && !(flags & CG_FILTER_NO_CLENSHAW) && cg_clenshaw_supported(g, N, Fin, Fout, K);
        if (!clenshaw) {
            rc = want_fused("cg_cheb_filter_bwd", g, 1, N, Fout, Fin, K, flags, &fused);
            if (rc != CG_OK) return rc;
            if (fused && !al) {
                CG_REQUIRE(!(flags & CG_FILTER_FORCE_FUSED), "cg_cheb_filter_bwd: fused kernel needs 16-byte aligned tensors");
                fused = false;
            }
        }
        void *wpack = reinterpret_cast<char *>(workspace) + (need - cg_fused_workspace(Fin, Fout, K));
        if (clenshaw) {
            // adjoint recurrence at the width of dx, G_k = gy W_k^T from tensor memory
            rc = cg_run_clenshaw(g, gy, W, dx, N, Fin, Fout, K, wpack, s);
            if (rc != CG_OK) return rc;
        } else if (fused) {
            // dx by the fused kernel on L~^T (Z_k never materialised)
            rc = cg_run_fused(g, 1, gy, W, dx, nullptr, N, Fout, Fin, K, true, wpack, s);
            if (rc != CG_OK) return rc;
        } else if (have_dW && unfused_clenshaw_ok(g, N, Fin, Fout, K, flags, gy, dx)) {
            // adjoint recurrence without the fused kernel (wide gy, e.g. the 4H gates of the gconv-LSTM):
            //   G = gy [W_0^T .. W_{K-1}^T]  (one GEMM, [R, K*Fin]),   b_k = G_k + 2 L~^T b_{k+1} - b_{k+2},
            //   dx = G_0 + L~^T b_1 - b_2  -- K-1 sparse steps at the width of dx instead of the width of gy
            const int64_t R = (int64_t)N * M;
            float *G = stack, *buf0 = G + R * K * Fin, *buf1 = buf0 + R * Fin, *Wp = buf1 + R * Fin;
            float *part = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + stack_bytes(g, N, Fout > Fin ? Fout : Fin, K));
            rc = cg_run_regroup_w(W, Wp, Fin, Fout, K, s);
            if (rc == CG_OK)        // G[r, k*Fin + f] = sum_fo gy[r, fo] Wp[k*Fin + f, fo]
                rc = cg_run_gemm(gy, Wp, G, (int)R, K * Fin, Fout, 0, 1, Fout, Fout, K * Fin, nullptr, 0, 0, 0, 0, 0, 0, part,
                                 cg_gemm_workspace((int)R, K * Fin, Fout), s);
            const float *b1 = G + (size_t)(K - 1) * Fin, *b2 = nullptr;       // b_{K-1} = G_{K-1}, b_K = 0
            int64_t s1 = (int64_t)K * Fin, s2 = 0;
            float *bufs[2] = {buf0, buf1};
            int nb = 0;
            for (int k = K - 2; k >= 1 && rc == CG_OK; --k) {
                // out may alias b2 (read and written element-wise by the same thread); a G view is never overwritten
                float *out = (b2 == buf0 || b2 == buf1) ? const_cast<float *>(b2) : bufs[nb++ & 1];
                if (out == b1) out = bufs[nb++ & 1];
                rc = cg_run_clenshaw_step(g, 1, G + (size_t)k * Fin, (int64_t)K * Fin, b1, s1, b2, s2, out, Fin, N, Fin, 2.0f, s);
                b2 = b1;
                s2 = s1;
                b1 = out;
                s1 = Fin;
            }
            if (rc == CG_OK) rc = cg_run_clenshaw_step(g, 1, G, (int64_t)K * Fin, b1, s1, b2, s2, dx, Fin, N, Fin, 1.0f, s);
            if (rc != CG_OK) return rc;
        } else {
            // Z_k = T_k(L~^T) gy materialised: dx = Z W^T, and dW = x^T Z_k if still missing
            float *part = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + stack_bytes(g, N, Fout, K));
            const bool sm = samples_ok(g, 1, N, Fout, flags, gy, stack);
            if (sm) {
                rc = cg_run_basis_samples(g, 1, gy, stack, N, Fout, K, s);
            } else {
                rc = cg_run_permute_abf(gy, stack, N, M, Fout, s);
                if (rc == CG_OK) rc = cg_run_basis(g, 1, stack, (int64_t)N * Fout, K, s, flags);
            }
            if (rc == CG_OK) {
                const int64_t R = (int64_t)N * M;
                if (sm && !(flags & CG_FILTER_NO_FUSED) && R < (int64_t)INT32_MAX) {
                    // dx = [Z_0 .. Z_{K-1}] W^T: A = K-blocked Z stack, B = W read as [fin][k*Fout + fo] (transposed)
                    rc = cg_run_gemm(stack, W, dx, (int)R, Fin, K * Fout, 0, 1, Fout, K * Fout, Fin, nullptr, 0, Fout,
                                     (long long)R * Fout, 0, 0, 0, part, cg_gemm_workspace((int)R, Fin, K * Fout), s);
                } else {
                    rc = cg_run_contract(stack, W, dx, N, M, Fout, Fin, K, true, sm, s);
                }
            }
            if (rc == CG_OK && !have_dW) rc = run_dw(g, stack, x, dW, N, Fout, Fin, K, true, sm, part, flags, s);
            if (rc != CG_OK) return rc;
            have_dW = true;
        }
    }
    if (!have_dW) {
        // dW = X_k^T gy with the (narrower) X-stack recomputed
        float *part = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + stack_bytes(g, N, Fin, K));
        const bool sm = samples_ok(g, 0, N, Fin, flags, x, stack);
        if (sm) {
            rc = cg_run_basis_samples(g, 0, x, stack, N, Fin, K, s);
        } else {
            rc = cg_run_permute_abf(x, stack, N, M, Fin, s);
            if (rc == CG_OK) rc = cg_run_basis(g, 0, stack, (int64_t)N * Fin, K, s, flags);
        }
        if (rc == CG_OK) rc = run_dw(g, stack, gy, dW, N, Fin, Fout, K, false, sm, part, flags, s);
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Contractions over a caller-owned Chebyshev stack (row-partitioned filter of config C5: the stack is the
// [K][nloc + nhalo][F] buffer of cnn_graph_b200/partition.py, so the slab stride is not R * F).
// ---------------------------------------------------------------------------------------------------------
extern "C" size_t cg_cheb_contract_workspace_bytes(int64_t R, int Fin, int Fout, int K) {
    if (R <= 0 || R >= (int64_t)INT32_MAX || Fin <= 0 || Fout <= 0 || K < 1) return 0;
    const size_t a = cg_gemm_workspace((int)R, Fout, K * Fin), b = cg_gemm_workspace((int)R, Fin, K * Fout);
    const size_t c = std::max(cg_gemm_workspace(Fin, Fout, (int)R), dw_allk_workspace(R, Fin, Fout, K));
    return std::max(a, std::max(b, c));
}

extern "C" int cg_cheb_contract(const float *stack, int64_t slab_stride, const float *W, float *y, int64_t R, int Fin,
                                int Fout, int K, int transposed, void *workspace, size_t workspace_bytes, void *stream) {
    CG_REQUIRE(R >= 0 && R < (int64_t)INT32_MAX && Fin > 0 && Fout > 0 && K >= 1, "cg_cheb_contract: bad dims");
    if (R == 0) return CG_OK;
    CG_REQUIRE(stack && W && y, "cg_cheb_contract: NULL tensor");
    cudaStream_t s = (cudaStream_t)stream;
    if (!transposed) {
        // y[r, fo] = sum_{k,f} stack_k[r, f] W[f*K + k, fo]: q = k*Fin + f  <->  W row f*K + k
        CG_REQUIRE(slab_stride >= R * Fin, "cg_cheb_contract: slab stride smaller than a slab");
        return cg_run_gemm(stack, W, y, (int)R, Fout, K * Fin, 0, 0, Fin, Fout, Fout, nullptr, 0, Fin, slab_stride, Fin, 1, K,
                           workspace, workspace_bytes, s);
    }
    // y[r, f] = sum_{k,fo} stack_k[r, fo] W[f*K + k, fo]: B = W read as [f][k*Fout + fo]
    CG_REQUIRE(slab_stride >= R * Fout, "cg_cheb_contract: slab stride smaller than a slab");
    return cg_run_gemm(stack, W, y, (int)R, Fin, K * Fout, 0, 1, Fout, K * Fout, Fin, nullptr, 0, Fout, slab_stride, 0, 0, 0,
                       workspace, workspace_bytes, s);
}

extern "C" int cg_cheb_contract_dw(const float *stack, int64_t slab_stride, const float *gy, float *dW, int64_t R, int Fin,
                                   int Fout, int K, void *workspace, size_t workspace_bytes, void *stream) {
    CG_REQUIRE(R >= 0 && R < (int64_t)INT32_MAX && Fin > 0 && Fout > 0 && K >= 1, "cg_cheb_contract_dw: bad dims");
    CG_REQUIRE(dW != nullptr, "cg_cheb_contract_dw: dW is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (R == 0) {
        CG_CHECK_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)Fin * K * Fout, s));
        return CG_OK;
    }
    CG_REQUIRE(stack && gy && slab_stride >= R * Fin, "cg_cheb_contract_dw: bad stack");
    if (dw_allk_ok(stack, slab_stride, gy, R, Fin, Fout, K) && workspace_bytes >= dw_allk_workspace(R, Fin, Fout, K))
        return run_dw_allk(stack, slab_stride, gy, dW, R, Fin, Fout, K, workspace, s);
    // dW[f*K + k, fo] = sum_r stack_k[r, f] gy[r, fo]: one GEMM per k into rows k, K + k, ... of dW
    for (int k = 0; k < K; ++k) {
        const int rc = cg_run_gemm(stack + (size_t)k * slab_stride, gy, dW + (size_t)k * Fout, Fin, Fout, (int)R, 1, 0, Fin,
                                   Fout, K * Fout, nullptr, 0, 0, 0, 0, 0, 0, workspace, workspace_bytes, s);
        if (rc != CG_OK) return rc;
    }
    return CG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// First-layer fusion (Fin = 1): gradients of  pool_max4(relu(filter(x; W) + b))  from the gradient of the POOLED
// output (lib/models.py:226-257 behind lib/models.py:192-224; TF autodiff through lib/graph_model.py:296).
// ---------------------------------------------------------------------------------------------------------
extern "C" int cg_cheb_dw_pooled_supported(const cg_graph_t *g, int N, int Fout, int K, int p, int act, int kind,
                                           int bias_kind) {
    if (!g || N <= 0 || g->M % 4 != 0 || p != 4 || act != 1 || kind != 1 || bias_kind < 0 || bias_kind > 1) return 0;
    if (cg_cheb_filter_stack_bytes(g, N, 1, Fout, K, 0) == 0) return 0;          // needs the fp32 basis [K][N][M]
    return cg_dw_thin_pooled_supported((long long)N * g->M, Fout, K, g->sm_count, g->smem_optin) ? 1 : 0;
}

// forward of the same layer in two launches: on-chip recurrence (the basis stays for the backward), then the contraction
// with bias + relu + max pooling of 4 applied in its epilogue -- the [N, M, Fout] filter output is never written
extern "C" int cg_cheb_first_layer_fwd_supported(const cg_graph_t *g, int N, int Fout, int K, int bias_kind) {
    if (!cg_cheb_dw_pooled_supported(g, N, Fout, K, 4, 1, 1, bias_kind)) return 0;
    return cg_basis_samples_supported(g, 0, N, 1) && cg_contract_umma_supported(N, g->M, 1, Fout, K, g->smem_optin) ? 1 : 0;
}

extern "C" int cg_cheb_first_layer_fwd(const cg_graph_t *g, const float *x, const float *W, const float *bias, float *stack_out,
                                       float *y_pooled, uint8_t *aux, int N, int Fout, int K, void *stream) {
    CG_REQUIRE(g != nullptr && x && W && stack_out && y_pooled && aux, "cg_cheb_first_layer_fwd: NULL argument");
    CG_REQUIRE(cg_cheb_first_layer_fwd_supported(g, N, Fout, K, bias ? 1 : 0), "cg_cheb_first_layer_fwd: shape not supported");
    CG_REQUIRE(((((uintptr_t)x) | ((uintptr_t)stack_out)) & 15) == 0, "cg_cheb_first_layer_fwd: unaligned tensor");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = cg_run_basis_samples(g, 0, x, stack_out, N, 1, K, s);
    if (rc != CG_OK) return rc;
    return cg_run_contract_umma(stack_out, W, nullptr, N, g->M, 1, Fout, K, g->sm_count, g->smem_optin, s, bias, y_pooled, aux);
}

extern "C" size_t cg_cheb_dw_pooled_workspace_bytes(const cg_graph_t *g, int N, int Fout, int K) {
    if (!g || N <= 0) return 0;
    return cg_dw_thin_pooled_workspace((long long)N * g->M, Fout, K, g->sm_count, g->smem_optin);
}

extern "C" int cg_cheb_dw_pooled(const cg_graph_t *g, const float *stack, const float *g_pooled, const float *y_pooled,
                                 const uint8_t *aux, float *dW, float *db, int N, int Fout, int K, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    CG_REQUIRE(g != nullptr && N > 0 && Fout > 0 && K >= 1, "cg_cheb_dw_pooled: bad arguments");
    CG_REQUIRE(stack && g_pooled && y_pooled && aux && dW, "cg_cheb_dw_pooled: NULL tensor");
    const size_t need = cg_cheb_dw_pooled_workspace_bytes(g, N, Fout, K);
    CG_REQUIRE(need > 0, "cg_cheb_dw_pooled: shape not supported (cg_cheb_dw_pooled_supported returned 0)");
    if (workspace == nullptr || workspace_bytes < need) {
        cg_set_error("cg_cheb_dw_pooled: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    return cg_run_dw_thin_pooled(stack, g_pooled, y_pooled, aux, dW, db, (long long)N * g->M, Fout, K,
                                 reinterpret_cast<float *>(workspace), g->sm_count, g->smem_optin, (cudaStream_t)stream);
}
