// Batched small fp32 GEMM on the FFMA pipe: C[b] = op(A[b]) . op(B[b]) for many small independent products.
//
// Used by the spectral `fourier` filter (reference lib/filter.py:11-27, lib/models.py:129-144): between the two dense
// graph-Fourier transforms the reference applies one Fout x Fin matrix per graph frequency (`tf.matmul(W, x)` with W
// [M, Fout, Fin] batched over M).  Per frequency that is a (N x Fin) . (Fin x Fout) product -- far too small and too
// many for the tensor-core GEMM; the op is bound by streaming its operands once (AI = 2 Fin Fout N / 4 (Fin N + Fin Fout
// + Fout N) ~ 7 flop/B at 32/32/100), so a classic shared-memory tiled FFMA kernel with 128-bit global accesses is the
// right tool.  One CTA = one 64 x 64 output tile of one batch entry, 256 threads x (4 x 4) register tile, reduction in
// chunks of 16.
#include "cg_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_bmm(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C,
                                             int m, int n, int k, int lda, int ldb, int ldc, long long sa, long long sb,
                                             long long sc) {
    __shared__ float As[TK][TM + 4];   // As[q][i] = op(A)[i0 + i][q0 + q]
    __shared__ float Bs[TK][TN + 4];   // Bs[q][j] = op(B)[q0 + q][j0 + j]
    const int b = blockIdx.z;
    A += (long long)b * sa;
    B += (long long)b * sb;
    C += (long long)b * sc;
    const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int q0 = 0; q0 < k; q0 += TK) {
        // stage op(A): 64 x 16 elements, 4 per thread; consecutive threads walk the contiguous direction of the storage
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = t + e * 256;
            int i, q;
            if (TA) { i = idx & 63; q = idx >> 6; }      // A stored [k][lda >= m]: i contiguous
            else    { q = idx & 15; i = idx >> 4; }      // A stored [m][lda >= k]: q contiguous
            const int gi = i0 + i, gq = q0 + q;
            float v = 0.f;
            if (gi < m && gq < k) v = TA ? A[(long long)gq * lda + gi] : A[(long long)gi * lda + gq];
            As[q][i] = v;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = t + e * 256;
            int j, q;
            if (TB) { q = idx & 15; j = idx >> 4; }      // B stored [n][ldb >= k]: q contiguous
            else    { j = idx & 63; q = idx >> 6; }      // B stored [k][ldb >= n]: j contiguous
            const int gj = j0 + j, gq = q0 + q;
            float v = 0.f;
            if (gj < n && gq < k) v = TB ? B[(long long)gj * ldb + gq] : B[(long long)gq * ldb + gj];
            Bs[q][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TK; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[q][ty * 4]);
            const float4 bb = *reinterpret_cast<const float4 *>(&Bs[q][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = i0 + ty * 4 + i;
        if (gi >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = j0 + tx * 4 + j;
            if (gj < n) C[(long long)gi * ldc + gj] = acc[i][j];
        }
    }
}

}  // namespace

extern "C" int cg_bmm_f32(const float *dev_A, const float *dev_B, float *dev_C, int batch, int m, int n, int k, int transA,
                          int transB, int lda, int ldb, int ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c,
                          void *stream) {
    CG_REQUIRE(batch >= 0 && m >= 0 && n >= 0 && k >= 0, "cg_bmm_f32: negative size");
    if (batch == 0 || m == 0 || n == 0) return CG_OK;
    CG_REQUIRE(dev_A && dev_B && dev_C, "cg_bmm_f32: NULL operand");
    CG_REQUIRE(batch <= 65535, "cg_bmm_f32: batch %d exceeds 65535", batch);
    CG_REQUIRE(lda >= (transA ? m : k) && ldb >= (transB ? k : n) && ldc >= n, "cg_bmm_f32: leading dimension too small");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)cg_ceil_div(n, TN), (unsigned)cg_ceil_div(m, TM), (unsigned)batch);
    CgProfScope prof("bmm", s);
    if (transA) {
        if (transB) k_bmm<true, true><<<grid, 256, 0, s>>>(dev_A, dev_B, dev_C, m, n, k, lda, ldb, ldc, stride_a, stride_b, stride_c);
        else k_bmm<true, false><<<grid, 256, 0, s>>>(dev_A, dev_B, dev_C, m, n, k, lda, ldb, ldc, stride_a, stride_b, stride_c);
    } else {
        if (transB) k_bmm<false, true><<<grid, 256, 0, s>>>(dev_A, dev_B, dev_C, m, n, k, lda, ldb, ldc, stride_a, stride_b, stride_c);
        else k_bmm<false, false><<<grid, 256, 0, s>>>(dev_A, dev_B, dev_C, m, n, k, lda, ldb, ldc, stride_a, stride_b, stride_c);
    }
    CG_LAUNCH_CHECK();
    return CG_OK;
}
