// General fp32 GEMM on the tensor cores:  C[M x N] = op(A)[M x K] . op(B)[K x N] (+ bias[N]) (relu)
//
// Used for the dense head of cgcnn (lib/models.py:268-274: fc = relu(x W + b), and its two gradients) and for
// contractions whose operands live in HBM with shapes the fused kernels do not take.  All matrices are fp32,
// row-major; the arithmetic is the same split as everywhere else in this library: every operand value becomes
// bf16 hi + mid, the product is  hi*hi + mid*hi + hi*mid  with fp32 accumulation in TMEM (error <= 2^-16 relative).
//
// One CTA owns a 128 x BN output tile (BN = 256 or 128) and a K range (split-K when the tile grid does not fill
// the SMs; partial tiles are summed by k_gemm_reduce).  Per K stage of 64:
//   16 compute warps read the fp32 operands straight from global memory in octets along their contiguous
//   dimension (two 128-bit loads), split them and store 16-byte bf16 octets into the canonical UMMA layouts
//   (K-major when K is the contiguous dimension, MN-major otherwise; (row, octet) pairs are dealt to lanes
//   diagonally so that the stores of a quarter-warp fill one 128-byte core-matrix line);
//   the issue warp (elect.sync) issues 3 x 4 MMAs of 128 x BN x 16 and commits to the stage's mbarrier.
// Two operand stages: the conversion of stage s+1 runs under the MMAs of stage s.
#include <stdlib.h>

#include <algorithm>

#include "cg_common.cuh"
#include "cg_umma.cuh"
#include "cg_fused_common.cuh"

namespace {

constexpr int GC = 512;          // compute threads
constexpr int GT = GC + 32;      // + issue warp
constexpr int BM = 128;
constexpr int BK = 64;

struct GemmParams {
    int npass;                   // MMA passes per product: 3 (fp32-equivalent hi/mid split) or 1 (single-pass bf16)
    const float *A, *B, *bias;
    float *C;                    // [M][ldc] (split == 1) or partials [split][M][N]
    int M, N, K, lda, ldb, ldc, transA, transB, relu, BN, split, k_per_split, vecA, vecB;
    // K-blocked operands (contractions over a Chebyshev stack): A element (m, q) at A[(q / a_kblk) * a_kbs + m * lda +
    // q % a_kblk] (transA == 0 only); B row of q (transB == 0 only) is (q / b_kblk) * b_shi + (q % b_kblk) * b_slo
    long long a_kbs;
    int a_kblk, b_kblk, b_shi, b_slo;
    uint32_t off_a, off_b, stage_bytes, a_plane, b_plane, off_bar;
};

// stage one operand tile [ROWS (m or n)] x [BK] from global memory into its canonical layout.
//   KC = true : K is the contiguous dimension of the source (element (r, k) at src[r * ld + k]); K-major layout
//               offset = (k/8) * (ROWS * 16) + (r/8) * 128 + (r%8) * 16
//   KC = false: the row index is contiguous (element (r, k) at src[k * ld + r]); MN-major layout
//               offset = (r/8) * (BK * 16) + (k/8) * 128 + (k%8) * 16
// r < r_lim, k < k_lim are the valid ranges (zero fill outside); `vec` = 128-bit loads allowed.
// K blocking: KC = true  -> element (r, q) at src[(q / kblk) * kbs + r * ld + q % kblk]   (kblk % 8 == 0 for `vec`)
//             KC = false -> source row of q is (q / kblk) * shi + (q % kblk) * slo
template <bool KC>
__device__ __forceinline__ void stage_operand(unsigned char *hi_plane, uint32_t plane_bytes, const float *src, int ld, int rows,
                                              int r0, int r_lim, int k0, int k_lim, int vec, int tid, int kblk, long long kbs,
                                              int shi, int slo) {
    const int n_r8 = rows / 8;                 // row octets (or rows / 8 blocks)
    const int total = n_r8 * (BK / 8) * 8;     // items: (8 x 8) blocks of (row-or-k, octet)
    for (int e = tid; e < total; e += GC) {
        const int blk = e >> 6, b = e & 63;
        const int i = b & 7, ph = b >> 3;
        float v[8];
        uint32_t off;
        if (KC) {
            // block = 8 rows x 8 k-octets; lane (i, ph): row i of the block, octet (i + ph) & 7  (BK / 8 == 8)
            const int r = blk * 8 + i, ko = (i + ph) & 7;
            const int gr = r0 + r, gk = k0 + ko * 8;
            if (gr < r_lim && gk + 7 < k_lim && vec) {
                const int kb = gk / kblk, ki = gk - kb * kblk;
                const float *p = src + (size_t)kb * kbs + (size_t)gr * ld + ki;
                const float4 a = *reinterpret_cast<const float4 *>(p), c = *reinterpret_cast<const float4 *>(p + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float x = 0.f;
                    if (gr < r_lim && gk + j < k_lim) {
                        const int kb = (gk + j) / kblk, ki = (gk + j) - kb * kblk;
                        x = src[(size_t)kb * kbs + (size_t)gr * ld + ki];
                    }
                    v[j] = x;
                }
            }
            off = (uint32_t)ko * (uint32_t)(rows * 16) + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
        } else {
            // block = 8 k x 8 row-octets; lane (i, ph): k = i of the block, row octet (i + ph) & 7
            const int nb_r = n_r8 / 8;                       // blocks along the row octets
            const int kb = blk / nb_r, rb = blk - kb * nb_r;
            const int k = kb * 8 + i, ro = rb * 8 + ((i + ph) & 7);
            const int gk = k0 + k, gr = r0 + ro * 8;
            const int qb = gk / kblk;
            const float *p = src + ((size_t)qb * shi + (size_t)(gk - qb * kblk) * slo) * ld + gr;
            if (gk < k_lim && gr + 7 < r_lim && vec) {
                const float4 a = *reinterpret_cast<const float4 *>(p), c = *reinterpret_cast<const float4 *>(p + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = (gk < k_lim && gr + j < r_lim) ? p[j] : 0.f;
            }
            off = (uint32_t)ro * (uint32_t)(BK * 16) + (uint32_t)(k >> 3) * 128u + (uint32_t)(k & 7) * 16u;
        }
        uint2 h0, m0, h1, m1;
        split4(make_float4(v[0], v[1], v[2], v[3]), h0, m0);
        split4(make_float4(v[4], v[5], v[6], v[7]), h1, m1);
        *reinterpret_cast<uint4 *>(hi_plane + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4 *>(hi_plane + plane_bytes + off) = make_uint4(m0.x, m0.y, m1.x, m1.y);
    }
}

__global__ void __launch_bounds__(GT, 1) k_gemm_umma(const GemmParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *mbar = bars;           // [2] MMAs of stage s completed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN;
    const int tiles_n = (p.N + BN - 1) / BN;
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x - tm * tiles_n;
    const int m0 = tm * BM, n0 = tn * BN;
    const int k_beg = blockIdx.y * p.k_per_split, k_end = min(p.K, k_beg + p.k_per_split);
    const int nst = k_end > k_beg ? (k_end - k_beg + BK - 1) / BK : 0;

    if (tid == 0) {
        umma::mbar_init(mbar, 1);
        umma::mbar_init(mbar + 1, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)BN);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t st0 = umma::smem_u32(smem);

    if (warp == GC / 32) {
        // =========================== MMA issue warp ======================================
        const uint32_t idesc = umma::make_idesc_bf16(BM, BN, p.transA ? 1 : 0, p.transB ? 0 : 1);
        // K-major: LBO = rows * 16 (between k octets), SBO = 128;  MN-major: LBO = 128 (k groups), SBO = BK * 16
        const uint32_t a_lbo = p.transA ? 128u : (uint32_t)BM * 16u, a_sbo = p.transA ? (uint32_t)BK * 16u : 128u;
        const uint32_t b_lbo = p.transB ? (uint32_t)BN * 16u : 128u, b_sbo = p.transB ? 128u : (uint32_t)BK * 16u;
        const uint32_t a_hi = umma::desc_hi(a_sbo), b_hi = umma::desc_hi(b_sbo);
        // one K = 16 step = two k octets: K-major -> 2 * LBO bytes, MN-major -> 2 * 128 bytes
        const uint32_t a_k = ((p.transA ? 256u : 2u * a_lbo)) >> 4, b_k = ((p.transB ? 2u * b_lbo : 256u)) >> 4;
        for (int s = 0; s < nst; ++s) {
            __syncthreads();                              // operands of stage s are staged
            if (umma::elect_one()) {
                umma::fence_after_sync();
                const uint32_t sb = st0 + (uint32_t)(s & 1) * p.stage_bytes;
                const uint32_t a_lo = umma::desc_lo(sb + p.off_a, a_lbo), b_lo = umma::desc_lo(sb + p.off_b, b_lbo);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    if (pass >= p.npass) break;
                    uint32_t al = a_lo + (pass == 1 ? (p.a_plane >> 4) : 0u), bl = b_lo + (pass == 2 ? (p.b_plane >> 4) : 0u);
#pragma unroll
                    for (int j = 0; j < BK / 16; ++j) {
                        umma::mma_bf16(tmem, umma::desc_join(al, a_hi), umma::desc_join(bl, b_hi), idesc, (s | pass | j) != 0);
                        al += a_k;
                        bl += b_k;
                    }
                }
                umma::commit(mbar + (s & 1));
            }
            __syncwarp();
        }
    } else {
        // =========================== compute warps ======================================
        for (int s = 0; s < nst; ++s) {
            if (s >= 2) umma::mbar_wait(mbar + (s & 1), (uint32_t)(((s - 2) >> 1) & 1));     // stage free again
            unsigned char *sb = smem + (size_t)(s & 1) * p.stage_bytes;
            const int k0 = k_beg + s * BK;
            if (p.transA)
                stage_operand<false>(sb + p.off_a, p.a_plane, p.A, p.lda, BM, m0, p.M, k0, k_end, p.vecA, tid, 1 << 30, 0, 0, 1);
            else
                stage_operand<true>(sb + p.off_a, p.a_plane, p.A, p.lda, BM, m0, p.M, k0, k_end, p.vecA, tid, p.a_kblk, p.a_kbs, 0, 1);
            if (p.transB)
                stage_operand<true>(sb + p.off_b, p.b_plane, p.B, p.ldb, BN, n0, p.N, k0, k_end, p.vecB, tid, 1 << 30, 0, 0, 1);
            else
                stage_operand<false>(sb + p.off_b, p.b_plane, p.B, p.ldb, BN, n0, p.N, k0, k_end, p.vecB, tid, p.b_kblk, 0, p.b_shi, p.b_slo);
            umma::fence_proxy_async();
            __syncthreads();
        }
        // ---- epilogue: TMEM -> (bias, relu) -> C  (thread = row)
        if (nst > 0) {
            umma::mbar_wait(mbar + ((nst - 1) & 1), (uint32_t)(((nst - 1) >> 1) & 1));
            umma::fence_after_sync();
        }
        const int qd = warp & 3, wq = warp >> 2;
        const int m = m0 + 32 * qd + lane;
        const bool final_out = p.split == 1;
        float *crow = final_out ? p.C + (size_t)m * p.ldc : p.C + ((size_t)blockIdx.y * p.M + m) * p.N;
        for (int c8 = wq; c8 < BN / 8; c8 += 4) {
            float v[8];
            if (nst > 0) {
                umma::tmem_ld8(tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(c8 * 8), v);
                umma::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
            const int n = n0 + c8 * 8;
            if (m < p.M) {
                if (final_out) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (p.bias && n + j < p.N) v[j] += p.bias[n + j];
                        if (p.relu) v[j] = fmaxf(v[j], 0.f);
                    }
                }
                const int ld = final_out ? p.ldc : p.N;
                if (n + 7 < p.N && (ld & 3) == 0 && ((((uintptr_t)crow) & 15) == 0)) {
                    *reinterpret_cast<float4 *>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4 *>(crow + n + 4) = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (n + j < p.N) crow[n + j] = v[j];
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)BN);
}

__global__ void __launch_bounds__(256)
k_gemm_reduce(const float *__restrict__ part, const float *__restrict__ bias, float *__restrict__ C, int M, int N, int ldc,
              int split, int relu) {
    const size_t total = (size_t)M * N;
    if ((N & 3) == 0 && (ldc & 3) == 0 && ((((uintptr_t)part) | ((uintptr_t)C)) & 15) == 0) {
        // four outputs per thread, two partial sums in flight per thread (same summation order for every output)
        const size_t total4 = total / 4;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
            int sp = 0;
            for (; sp + 1 < split; sp += 2) {
                const float4 a = reinterpret_cast<const float4 *>(part + (size_t)sp * total)[i];
                const float4 b = reinterpret_cast<const float4 *>(part + (size_t)(sp + 1) * total)[i];
                s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
                s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
            }
            if (sp < split) {
                const float4 a = reinterpret_cast<const float4 *>(part + (size_t)sp * total)[i];
                s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
            }
            float4 r = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
            const size_t e = i * 4;
            const int m = (int)(e / N), n = (int)(e - (size_t)m * N);
            if (bias) { r.x += bias[n]; r.y += bias[n + 1]; r.z += bias[n + 2]; r.w += bias[n + 3]; }
            if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            *reinterpret_cast<float4 *>(C + (size_t)m * ldc + n) = r;
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < split; ++sp) s += part[(size_t)sp * total + i];
        const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
        if (bias) s += bias[n];
        if (relu) s = fmaxf(s, 0.f);
        C[(size_t)m * ldc + n] = s;
    }
}

struct GPlan {
    int BN, split, k_per_split, tiles;
    size_t smem, ws;
    GemmParams gp;
};

static GPlan gemm_plan(int M, int N, int K, int sm_count) {
    GPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.gp.npass = cg_mma_passes();
    pl.BN = N > 128 ? 256 : 128;
    const int tiles = (int)(cg_ceil_div(M, BM) * cg_ceil_div(N, pl.BN));
    pl.tiles = tiles;
    int split = 1;
    if (tiles < sm_count) {
        split = std::max(1, sm_count / tiles);
        const int max_split = (int)cg_ceil_div(K, 4 * BK);       // at least four stages per CTA
        if (split > max_split) split = std::max(1, max_split);
    }
    int kps = (int)cg_ceil_div(K, split);
    kps = (int)cg_ceil_div(kps, BK) * BK;
    pl.split = (int)cg_ceil_div(K, kps);
    pl.k_per_split = kps;
    GemmParams &gp = pl.gp;
    gp.a_plane = (uint32_t)BM * BK * 2u;
    gp.b_plane = (uint32_t)pl.BN * BK * 2u;
    gp.off_a = 0;
    gp.off_b = 2 * gp.a_plane;
    gp.stage_bytes = 2 * gp.a_plane + 2 * gp.b_plane;
    gp.off_bar = 2 * gp.stage_bytes;
    pl.smem = gp.off_bar + 64;
    pl.ws = pl.split > 1 ? sizeof(float) * (size_t)pl.split * M * N : 0;
    return pl;
}

}  // namespace

extern "C" size_t cg_gemm_f32_workspace_bytes(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) sms = cg_sm_budget(dev);
    return std::max(std::max(gemm_plan(M, N, K, sms).ws, cg_gemm_pipe_workspace(M, N, K, sms)), cg_gemm_stream_workspace(M, N, K, sms));
}

int cg_gemm_reduce(const float *part, const float *bias, float *C, int M, int N, int ldc, int split, int relu, cudaStream_t s) {
    CgProfScope prof("gemm_reduce", s);
    const size_t total = (size_t)M * N;
    k_gemm_reduce<<<(unsigned)std::min<size_t>((total + 255) / 256, 148 * 8), 256, 0, s>>>(part, bias, C, M, N, ldc, split,
                                                                                            relu ? 1 : 0);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// C = op(A) op(B) with optional K blocking (see GemmParams); a_kblk <= 0 / b_kblk <= 0 mean "not blocked"
int cg_run_gemm(const float *A, const float *B, float *C, int M, int N, int K, int transA, int transB, int lda, int ldb,
                int ldc, const float *bias, int relu, int a_kblk, long long a_kbs, int b_kblk, int b_shi, int b_slo,
                void *workspace, size_t workspace_bytes, cudaStream_t s) {
    CG_REQUIRE(M >= 0 && N >= 0 && K >= 0, "cg_gemm_f32: negative dimension");
    if (M == 0 || N == 0) return CG_OK;
    CG_REQUIRE(A && B && C, "cg_gemm_f32: NULL matrix");
    CG_REQUIRE(K > 0, "cg_gemm_f32: K must be positive");
    CG_REQUIRE(!(a_kblk > 0 && transA) && !(b_kblk > 0 && transB), "cg_gemm_f32: K blocking needs the untransposed operand");
    int dev = 0, sms = 148;
    CG_CHECK_CUDA(cudaGetDevice(&dev));
    sms = cg_sm_budget(dev);
    if (workspace && cg_gemm_stream_eligible(A, M, N, K, lda, transA, transB, a_kblk, a_kbs, b_kblk, 0, 0, workspace_bytes, sms)) {
        const int rc = cg_run_gemm_stream(A, B, C, M, N, K, transA, transB, lda, ldb, ldc, bias, relu, a_kblk, a_kbs, b_kblk, b_shi,
                                          b_slo, workspace, workspace_bytes, sms, s, 0, 0);
        if (rc != CG_TRY_NEXT) return rc;
    }
    if (cg_gemm_pipe_eligible(A, B, M, N, K, lda, ldb, transA, transB, a_kblk, a_kbs, b_kblk))
        return cg_run_gemm_pipe(A, B, C, M, N, K, transA, transB, lda, ldb, ldc, bias, relu, a_kblk, a_kbs, b_kblk, b_shi,
                                b_slo, workspace, workspace_bytes, sms, s);
    GPlan pl = gemm_plan(M, N, K, sms);
    CG_REQUIRE(pl.ws == 0 || (workspace && workspace_bytes >= pl.ws), "cg_gemm_f32: workspace too small (%zu < %zu bytes)",
               workspace_bytes, pl.ws);
    GemmParams &gp = pl.gp;
    gp.A = A;
    gp.B = B;
    gp.bias = bias;
    gp.C = pl.split > 1 ? reinterpret_cast<float *>(workspace) : C;
    gp.M = M;
    gp.N = N;
    gp.K = K;
    gp.lda = lda;
    gp.ldb = ldb;
    gp.ldc = ldc;
    gp.transA = transA ? 1 : 0;
    gp.transB = transB ? 1 : 0;
    gp.relu = relu ? 1 : 0;
    gp.BN = pl.BN;
    gp.split = pl.split;
    gp.k_per_split = pl.k_per_split;
    gp.a_kblk = a_kblk > 0 ? a_kblk : (1 << 30);
    gp.a_kbs = a_kblk > 0 ? a_kbs : 0;
    gp.b_kblk = b_kblk > 0 ? b_kblk : (1 << 30);
    gp.b_shi = b_kblk > 0 ? b_shi : 0;
    gp.b_slo = b_kblk > 0 ? b_slo : 1;
    gp.vecA = (lda % 4 == 0) && ((((uintptr_t)A) & 15) == 0) && (a_kblk <= 0 || (a_kblk % 8 == 0 && a_kbs % 4 == 0));
    gp.vecB = (ldb % 4 == 0) && ((((uintptr_t)B) & 15) == 0);
    CG_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    {
        CgProfScope prof("gemm_umma", s);
        dim3 grid((unsigned)pl.tiles, (unsigned)pl.split);
        k_gemm_umma<<<grid, GT, pl.smem, s>>>(gp);
        CG_LAUNCH_CHECK();
    }
    if (pl.split > 1)
        return cg_gemm_reduce(reinterpret_cast<const float *>(workspace), bias, C, M, N, ldc, pl.split, relu, s);
    return CG_OK;
}

size_t cg_gemm_workspace(int M, int N, int K) { return cg_gemm_f32_workspace_bytes(M, N, K); }

extern "C" int cg_gemm_f32(const float *A, const float *B, float *C, int M, int N, int K, int transA, int transB, int lda,
                           int ldb, int ldc, const float *bias, int relu, void *workspace, size_t workspace_bytes,
                           void *stream) {
    return cg_run_gemm(A, B, C, M, N, K, transA, transB, lda, ldb, ldc, bias, relu, 0, 0, 0, 0, 0, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}

bool cg_gemm_mblocked_ok(const float *A, const float *B, int M, int N, int K, int lda, int ldb, int a_mblk, long long a_mbs) {
    return a_mblk > 0 && cg_gemm_pipe_eligible(A, B, M, N, K, lda, ldb, 1, 0, 0, 0, 0, a_mblk, a_mbs);
}

int cg_run_gemm_mblocked(const float *A, const float *B, float *C, int M, int N, int K, int lda, int ldb, int ldc, int a_mblk,
                         long long a_mbs, void *workspace, size_t workspace_bytes, cudaStream_t s) {
    CG_REQUIRE(cg_gemm_mblocked_ok(A, B, M, N, K, lda, ldb, a_mblk, a_mbs), "cg_run_gemm_mblocked: operands not eligible");
    int dev = 0, sms = 148;
    CG_CHECK_CUDA(cudaGetDevice(&dev));
    sms = cg_sm_budget(dev);
    if (workspace && cg_gemm_stream_eligible(A, M, N, K, lda, 1, 0, 0, 0, 0, a_mblk, a_mbs, workspace_bytes, sms)) {
        const int rc = cg_run_gemm_stream(A, B, C, M, N, K, 1, 0, lda, ldb, ldc, nullptr, 0, 0, 0, 0, 0, 0, workspace, workspace_bytes,
                                          sms, s, a_mblk, a_mbs);
        if (rc != CG_TRY_NEXT) return rc;
    }
    return cg_run_gemm_pipe(A, B, C, M, N, K, 1, 0, lda, ldb, ldc, nullptr, 0, 0, 0, 0, 0, 0, workspace, workspace_bytes, sms, s,
                            a_mblk, a_mbs);
}
