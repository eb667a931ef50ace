// Unit-test kernel for the tcgen05 primitives in cg_umma.cuh: one CTA computes
//   D[128][N] = A * B^T   with bf16 operands, fp32 accumulation in TMEM,
// for every combination of K-major / MN-major operand storage.  Exposed through
// cg_debug_umma_gemm so the GPU parity tests can pin descriptor encodings before the fused
// kernels rely on them.
#include "cg_common.cuh"
#include "cg_umma.cuh"

extern "C" int cg_debug_umma_gemm_m(const float *A, const float *B, float *D, int Mr, int N, int Kd, int a_mn,
                                    int b_mn, void *stream);

// A_src: a_mn == 0 -> [128][Kd] (k contiguous);  a_mn == 1 -> [Kd][128] (m contiguous)
// B_src: b_mn == 0 -> [N][Kd];                   b_mn == 1 -> [Kd][N]
__global__ void __launch_bounds__(128, 1)
k_umma_test(const float *__restrict__ A_src, const float *__restrict__ B_src, float *__restrict__ D, int N, int Kd,
            int a_mn_flags, int b_mn, int Mr) {
    const int a_mn = a_mn_flags & 1;     // bit 1 of a_mn_flags: (M = 64 only) put the accumulator at TMEM lane offset 16
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __nv_bfloat16 *As = reinterpret_cast<__nv_bfloat16 *>(smem);
    __nv_bfloat16 *Bs = As + Mr * Kd;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // strides (bytes).  K-major: LBO = between K chunks, SBO = between 8-row groups.
    //                   MN-major: SBO = between 8-element MN chunks, LBO = between 8-row K groups.
    const uint32_t a_lbo = a_mn ? 128u * (Mr / 8) : 128u;          // MN: k-group stride = 16 mn-chunks * 128 B
    const uint32_t a_sbo = a_mn ? 128u : 128u * (Kd / 8);
    const uint32_t b_lbo = b_mn ? 128u * (N / 8) : 128u;
    const uint32_t b_sbo = b_mn ? 128u : 128u * (Kd / 8);

    for (int e = tid; e < Mr * Kd; e += 128) {
        int r, k;
        if (a_mn) { k = e / Mr; r = e % Mr; } else { r = e / Kd; k = e % Kd; }
        const float v = A_src[e];
        uint32_t off = a_mn ? (r / 8) * a_sbo + (k / 8) * a_lbo + (k % 8) * 16 + (r % 8) * 2
                            : (r / 8) * a_sbo + (k / 8) * a_lbo + (r % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<unsigned char *>(As) + off) = __float2bfloat16_rn(v);
    }
    for (int e = tid; e < N * Kd; e += 128) {
        int r, k;
        if (b_mn) { k = e / N; r = e % N; } else { r = e / Kd; k = e % Kd; }
        const float v = B_src[e];
        uint32_t off = b_mn ? (r / 8) * b_sbo + (k / 8) * b_lbo + (k % 8) * 16 + (r % 8) * 2
                            : (r / 8) * b_sbo + (k / 8) * b_lbo + (r % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<unsigned char *>(Bs) + off) = __float2bfloat16_rn(v);
    }
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::fence_mbar_init();
    }
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols *= 2;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, ncols);
    umma::fence_proxy_async();          // operand tiles written with st.shared -> async proxy
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_bf16(Mr, N, a_mn, b_mn);
        const uint32_t a0 = umma::smem_u32(As), b0 = umma::smem_u32(Bs);
        for (int k16 = 0; k16 < Kd / 16; ++k16) {
            const uint64_t ad = umma::make_desc(a0 + k16 * 2 * a_lbo, a_lbo, a_sbo);
            const uint64_t bd = umma::make_desc(b0 + k16 * 2 * b_lbo, b_lbo, b_sbo);
            umma::mma_bf16(tmem + ((uint32_t)(Mr == 64 ? (a_mn_flags >> 1) * 16 : 0) << 16), ad, bd, idesc, k16 > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();

    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 8) {
        float v[8];
        umma::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[row * N + c + i] = v[i];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

extern "C" int cg_debug_umma_gemm(const float *A, const float *B, float *D, int N, int Kd, int a_mn, int b_mn,
                                  void *stream) {
    return cg_debug_umma_gemm_m(A, B, D, 128, N, Kd, a_mn, b_mn, stream);
}

// Same with Mr = 64 or 128 rows of A; D always receives all 128 TMEM lanes x N columns.
extern "C" int cg_debug_umma_gemm_m(const float *A, const float *B, float *D, int Mr, int N, int Kd, int a_mn,
                                    int b_mn, void *stream) {
    CG_REQUIRE(Mr == 64 || Mr == 128, "cg_debug_umma_gemm_m: Mr must be 64 or 128");
    CG_REQUIRE(A && B && D, "cg_debug_umma_gemm: NULL tensor");
    CG_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256, "cg_debug_umma_gemm: N must be a multiple of 16 in [16, 256]");
    CG_REQUIRE(Kd % 16 == 0 && Kd >= 16 && Kd <= 256, "cg_debug_umma_gemm: Kd must be a multiple of 16 in [16, 256]");
    const size_t smem = sizeof(__nv_bfloat16) * (size_t)(128 + N) * Kd;
    CG_CHECK_CUDA(cudaFuncSetAttribute(k_umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CgProfScope prof("umma_test", (cudaStream_t)stream);
    k_umma_test<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, Kd, a_mn, b_mn, Mr);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------------------
// tcgen05.mma issue/throughput micro-benchmark (debug aid, not in the public header).
// One CTA, operands zero-filled in shared memory; `reps` rounds of `per` MMAs (M = 128, N, K = 16 each)
// on `nacc` accumulators, then one commit.  mode 0: K-major SWIZZLE_NONE (core matrices of a K chunk
// contiguous, LBO = 2048*..), 1: MN-major SWIZZLE_NONE, 2: K-major SWIZZLE_128B (rows of 64 bf16),
// 3: K-major SWIZZLE_NONE with the K chunks of an 8-row group adjacent (LBO = 128).
// out[0] = clocks from first issue to completion, out[1] = clocks spent issuing.
__global__ void __launch_bounds__(128, 1)
k_umma_bench(int mode, int N, int per, int reps, long long *out, int bg, int same_acc) {
    __shared__ volatile int stop_flag;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (64 * 1024) / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::fence_mbar_init();
        stop_flag = 0;
    }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t a0 = umma::smem_u32(smem), b0 = a0 + 32 * 1024;
        const int a_mn = mode == 1, b_mn = mode == 1;
        const uint32_t idesc = umma::make_idesc_bf16(128, N, a_mn, b_mn);
        uint64_t ad, bd;
        if (mode == 0) {            // K-major, chunk-major: LBO = 128 rows * 16 B, SBO = 128
            ad = umma::make_desc(a0, 2048u, 128u);
            bd = umma::make_desc(b0, (uint32_t)N * 16u, 128u);
        } else if (mode == 1) {     // MN-major: LBO = 128 (k groups), SBO = 256
            ad = umma::make_desc(a0, 128u, 256u);
            bd = umma::make_desc(b0, 128u, 256u);
        } else if (mode == 2) {     // K-major SWIZZLE_128B: SBO = 1024, layout type 2 (bits 61..63)
            ad = umma::make_desc(a0, 16u, 1024u) | ((uint64_t)2 << 61);
            bd = umma::make_desc(b0, 16u, 1024u) | ((uint64_t)2 << 61);
        } else {                    // K-major, K chunks adjacent: LBO = 128, SBO = 256
            ad = umma::make_desc(a0, 128u, 256u);
            bd = umma::make_desc(b0, 128u, 256u);
        }
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            uint32_t acc = tmem;
            for (int i = 0; i < per; ++i) {
                // distinct operand tiles per MMA (4 KB apart for A, N*32 bytes for B) when mode >= 4 is not set
                const uint64_t ao = (uint64_t)(((uint32_t)(i & 3) * 4096u) >> 4), bo = (uint64_t)(((uint32_t)(i & 3) * (uint32_t)N * 32u) >> 4);
                umma::mma_bf16(acc, ad + (same_acc >= 2 ? 0 : ao), bd + (same_acc >= 2 ? 0 : bo), idesc, true);
                if (!same_acc) {
                    acc += (uint32_t)N;
                    if (acc + (uint32_t)N > tmem + 512u) acc = tmem;
                }
            }
        }
        umma::commit(&bar);
        const long long t1 = clock64();
        umma::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t2 - t0;
        out[1] = t1 - t0;
        stop_flag = 1;
    } else if (bg && warp >= 1) {
        // background shared-memory traffic: conflict-free 128-bit loads (and stores when bg == 2)
        float4 accv = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 *base = reinterpret_cast<float4 *>(smem + 48 * 1024);
        int it = 0;
        while (!stop_flag) {
#pragma unroll 8
            for (int u = 0; u < 8; ++u) {
                const float4 v = base[((it + u) * 96 + (tid - 32)) & 1023];
                accv.x += v.x; accv.y += v.y; accv.z += v.z; accv.w += v.w;
                if (bg == 2) base[((it + u) * 96 + (tid - 32) + 512) & 1023] = accv;
            }
            it += 8;
        }
        if (accv.x == 123.456f) out[1] = 0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

extern "C" int cg_debug_umma_bench(int mode, int N, int per, int reps, long long *dev_out, int bg, int same_acc,
                                   void *stream) {
    CG_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256, "cg_debug_umma_bench: bad N");
    const size_t smem = 64 * 1024;
    CG_CHECK_CUDA(cudaFuncSetAttribute(k_umma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_bench<<<1, 128, smem, (cudaStream_t)stream>>>(mode, N, per, reps, dev_out, bg, same_acc);
    CG_LAUNCH_CHECK();
    return CG_OK;
}


// ---------------------------------------------------------------------------------------
// A operand in tensor memory (the Clenshaw dx kernel keeps gy there): D[128][N] = A * B^T with
// A [128][Kd] written by tcgen05.st (two bf16 per 32-bit column), B [N][Kd] K-major in shared memory.
__global__ void __launch_bounds__(128, 1)
k_umma_test_ts(const float *__restrict__ A_src, const float *__restrict__ B_src, float *__restrict__ D, int N, int Kd) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b_lbo = 128u, b_sbo = 128u * (Kd / 8);
    for (int e = tid; e < N * Kd; e += 128) {
        const int r = e / Kd, k = e % Kd;
        const uint32_t off = (r / 8) * b_sbo + (k / 8) * b_lbo + (r % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(smem + off) = __float2bfloat16_rn(B_src[e]);
    }
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t a_col = 256;                       // A lives at columns [256, 256 + Kd/2)
    {
        const int row = warp * 32 + lane;
        for (int c = 0; c < Kd / 2; c += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __nv_bfloat162 v = __floats2bfloat162_rn(A_src[row * Kd + 2 * (c + i)], A_src[row * Kd + 2 * (c + i) + 1]);
                r[i] = *reinterpret_cast<const uint32_t *>(&v);
            }
            umma::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + a_col + (uint32_t)c, r);
        }
        umma::tmem_st_wait();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_bf16(128, N, 0, 0);
        const uint32_t b0 = umma::smem_u32(smem);
        for (int k16 = 0; k16 < Kd / 16; ++k16) {
            const uint64_t bd = umma::make_desc(b0 + k16 * 2 * b_lbo, b_lbo, b_sbo);
            umma::mma_bf16_ts(tmem, tmem + a_col + (uint32_t)k16 * 8u, bd, idesc, k16 > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 8) {
        float v[8];
        umma::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[row * N + c + i] = v[i];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

extern "C" int cg_debug_umma_gemm_ts(const float *A, const float *B, float *D, int N, int Kd, void *stream) {
    CG_REQUIRE(A && B && D, "cg_debug_umma_gemm_ts: NULL tensor");
    CG_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256, "cg_debug_umma_gemm_ts: N must be a multiple of 16 in [16, 256]");
    CG_REQUIRE(Kd % 16 == 0 && Kd >= 16 && Kd <= 256, "cg_debug_umma_gemm_ts: Kd must be a multiple of 16 in [16, 256]");
    const size_t smem = sizeof(__nv_bfloat16) * (size_t)N * Kd;
    CG_CHECK_CUDA(cudaFuncSetAttribute(k_umma_test_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CgProfScope prof("umma_test_ts", (cudaStream_t)stream);
    k_umma_test_ts<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, Kd);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
