// Weight gradient of a first-layer Chebyshev filter (scalar input signal, Fin = 1; lib/models.py:222-223 through
// lib/graph_model.py:296):
//
//     P[k][b] = sum_r X_k[r] * gy[r][b],      r over all N*M vertex signals,  K <= 32,  Fb in {32, 64}
//
// The output is tiny (K x Fb) and the reduction very long: 2*K*Fb flops per 4*(K + Fb) bytes is under 30 flop/B, so
// the job is HBM-bound on the FFMA pipe and the tensor cores (128 accumulator lanes for K <= 32 rows) only add
// conversion work.  Every CTA streams a contiguous row range through a 3-stage shared-memory ring -- one producer
// warp issues cp.async.bulk copies: K pieces [RC] of the fp32 basis [K][R] and the [RC][Fb] block of gy -- and sixteen
// compute warps take groups of four rows: a lane owns output feature(s) b, reads X_k[r..r+3] with one broadcast
// 128-bit load per k and its gy values conflict-free, and keeps all K partial sums in registers.  Warps are
// reduced through shared memory at the end; k_reduce_partials sums the per-CTA results.
//
// POOLED variant (first layer followed by bias + relu + max pooling of 4, lib/models.py:226-257): the kernel takes
// the gradient of the POOLED output instead of gy.  Pool groups are 4 consecutive rows, so per group and feature
// gy has at most one non-zero, at the argmax row, equal to the pooled gradient where the pooled output is positive:
// each lane reads the basis value of its argmax row and does one FMA per k with gp * [yp > 0]; the bias gradient is the
// sum of those weights, and neither the pooling-backward kernel nor the 4x larger gy ever exist.
#include <algorithm>

#include "cg_common.cuh"
#include "cg_umma.cuh"
#include "cg_fused_common.cuh"

namespace {

constexpr int TW = 16;                 // compute warps
constexpr int TT = TW * 32 + 32;       // + producer warp
constexpr int RC = 256;                // rows per stage
constexpr int NST = 3;                 // at most; two when a stage is larger than a third of shared memory
constexpr int KMAX = 32;

struct ThinParams {
    const float *stack;     // [K][R]
    const float *T;         // gy [R][Fb];  POOLED: gradient of the pooled output [R/4][Fb]
    const float *yp;        // POOLED: pooled output [R/4][Fb] (relu mask)
    const unsigned char *aux;   // POOLED: argmax inside the group [R/4][Fb]
    float *part;            // [CTAs][K (+1: bias gradient)][Fb]
    long long R, rows_per_cta;
    int K, Fb, nst;
    uint32_t stage_bytes, off_g, off_y, off_a, off_bar;
};

// NF: output features per lane (Fb = 32 * NF); KT: K rounded up to a multiple of 4 -- the k loop is fully unrolled
// without predicates; the up to three extra rows of the X area are never loaded and their sums never written
template <int NF, int KT, bool POOLED>
__global__ void __launch_bounds__(TT, 1) k_dw_thin(const ThinParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *full = bars, *empty = bars + NST;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, Fb = p.Fb, nst = p.nst;
    const long long r_beg = (long long)blockIdx.x * p.rows_per_cta, r_end = std::min(p.R, r_beg + p.rows_per_cta);
    const int nchunks = r_end > r_beg ? (int)((r_end - r_beg + RC - 1) / RC) : 0;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) {
            umma::mbar_init(full + i, 1);
            umma::mbar_init(empty + i, TW);
        }
        umma::fence_mbar_init();
    }
    __syncthreads();
    const uint32_t st0 = umma::smem_u32(smem);
    float acc[KT][NF];
#pragma unroll
    for (int k = 0; k < KT; ++k)
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[k][f] = 0.f;
    float accb[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) accb[f] = 0.f;

    if (warp == TW) {
        // =========================== producer warp ======================================
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % nst;
            if (c >= nst) {
                if (lane == 0) umma::mbar_wait(empty + s, (uint32_t)((c / nst - 1) & 1));
                __syncwarp();
            }
            const long long rb = r_beg + (long long)c * RC;
            const uint32_t rows = (uint32_t)std::min<long long>(RC, r_end - rb);
            const uint32_t dst = st0 + (uint32_t)s * p.stage_bytes;
            if (lane == 0) {
                umma::fence_proxy_async();      // the slot was read through the generic proxy
                mbar_expect_tx(full + s, POOLED ? rows * 4u * (uint32_t)K + (rows / 4u) * (uint32_t)Fb * 9u
                                                : rows * 4u * (uint32_t)(K + Fb));
            }
            __syncwarp();
            for (int k = lane; k < K; k += 32) bulk_g2s(dst + (uint32_t)k * RC * 4u, p.stack + (size_t)k * p.R + rb, rows * 4u, full + s);
            if (!POOLED) {
                if (lane == 0) bulk_g2s(dst + p.off_g, p.T + (size_t)rb * Fb, rows * (uint32_t)Fb * 4u, full + s);
            } else {
                const size_t jb = (size_t)(rb / 4) * Fb;
                const uint32_t n = (rows / 4u) * (uint32_t)Fb;
                if (lane == 0) bulk_g2s(dst + p.off_g, p.T + jb, n * 4u, full + s);
                if (lane == 1) bulk_g2s(dst + p.off_y, p.yp + jb, n * 4u, full + s);
                if (lane == 2) bulk_g2s(dst + p.off_a, p.aux + jb, n, full + s);
            }
        }
    } else {
        // =========================== compute warps ======================================
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % nst;
            const long long rb = r_beg + (long long)c * RC;
            const int rows = (int)std::min<long long>(RC, r_end - rb);
            umma::mbar_wait(full + s, (uint32_t)((c / nst) & 1));
            // plain (non-volatile) shared loads: the compiler batches the broadcast reads of several k ahead of the FMAs
            const float *xs = reinterpret_cast<const float *>(smem + (size_t)s * p.stage_bytes);
            const float *gs = reinterpret_cast<const float *>(smem + (size_t)s * p.stage_bytes + p.off_g);
            for (int r = 4 * warp; r < rows; r += 4 * TW) {         // rows % 4 == 0 (R % 4 == 0 is required)
                if (POOLED) {
                    // one non-zero per group and feature: read the basis value of the argmax row directly (the lanes'
                    // addresses differ by at most 12 bytes: one wavefront) and do a single FMA per k
                    const float *ys = reinterpret_cast<const float *>(smem + (size_t)s * p.stage_bytes + p.off_y);
                    const unsigned char *as = smem + (size_t)s * p.stage_bytes + p.off_a;
                    float gv[NF];
                    const float *xa[NF];
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        const int i = (r >> 2) * Fb + lane + 32 * f;
                        gv[f] = ys[i] > 0.f ? gs[i] : 0.f;                  // relu'(pooled output) * pooled gradient
                        xa[f] = xs + r + as[i];
                        accb[f] += gv[f];
                    }
#pragma unroll
                    for (int k = 0; k < KT; ++k)
#pragma unroll
                        for (int f = 0; f < NF; ++f) acc[k][f] = fmaf(xa[f][k * RC], gv[f], acc[k][f]);
                    continue;
                }
                float g[4][NF];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int f = 0; f < NF; ++f) g[j][f] = gs[(r + j) * Fb + lane + 32 * f];
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    const float4 x = *reinterpret_cast<const float4 *>(xs + k * RC + r);
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        acc[k][f] = fmaf(x.x, g[0][f], acc[k][f]);
                        acc[k][f] = fmaf(x.y, g[1][f], acc[k][f]);
                        acc[k][f] = fmaf(x.z, g[2][f], acc[k][f]);
                        acc[k][f] = fmaf(x.w, g[3][f], acc[k][f]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    }
    // ---- reduce the compute warps through shared memory (every chunk has been consumed: the ring is idle), write
    // the CTA's partial result
    __syncthreads();
    const int KO = POOLED ? K + 1 : K;                    // output rows: K weight rows (+ the bias gradient)
    float *red = reinterpret_cast<float *>(smem);         // [TW][KO][Fb]
    if (warp < TW) {
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K)
#pragma unroll
                for (int f = 0; f < NF; ++f) red[((size_t)warp * KO + k) * Fb + lane + 32 * f] = acc[k][f];
        if (POOLED)
#pragma unroll
            for (int f = 0; f < NF; ++f) red[((size_t)warp * KO + K) * Fb + lane + 32 * f] = accb[f];
    }
    __syncthreads();
    for (int i = tid; i < KO * Fb; i += TT) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < TW; ++w) t += red[(size_t)w * KO * Fb + i];
        p.part[(size_t)blockIdx.x * KO * Fb + i] = t;
    }
}

struct ThinPlan {
    bool ok = false;
    int ctas = 0;
    size_t smem = 0;
    ThinParams tp;
};

static ThinPlan thin_plan(long long R, int Fb, int K, int sm_count, size_t smem_limit, bool pooled = false) {
    ThinPlan pl;
    if (K < 1 || K > KMAX || (Fb != 32 && Fb != 64) || R < 1 || R % 4 != 0) return pl;
    ThinParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.off_g = (uint32_t)((K + 3) / 4 * 4) * RC * 4u;       // X area: K rounded up to a multiple of 4 rows of [RC]
    if (!pooled) {
        tp.stage_bytes = tp.off_g + (uint32_t)RC * Fb * 4u;
    } else {                                                 // pooled gradient, pooled output (fp32), argmax (u8)
        tp.off_y = tp.off_g + (uint32_t)(RC / 4) * Fb * 4u;
        tp.off_a = tp.off_y + (uint32_t)(RC / 4) * Fb * 4u;
        tp.stage_bytes = (uint32_t)cg_align_up(tp.off_a + (uint32_t)(RC / 4) * Fb, 128);
    }
    tp.nst = (int)std::min<size_t>(NST, (smem_limit - 128) / tp.stage_bytes);
    if (tp.nst < 2 || (size_t)tp.nst * tp.stage_bytes < (size_t)TW * (K + 1) * Fb * 4) return pl;
    tp.off_bar = (uint32_t)tp.nst * tp.stage_bytes;
    pl.smem = tp.off_bar + 128;
    int ctas = (int)std::min<long long>(sm_count, cg_ceil_div(R, RC));
    long long rpc = cg_ceil_div(cg_ceil_div(R, ctas), RC) * RC;
    pl.ctas = (int)cg_ceil_div(R, rpc);
    tp.rows_per_cta = rpc;
    pl.tp = tp;
    pl.ok = true;
    return pl;
}

}  // namespace

bool cg_dw_thin_supported(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit) {
    return Fa == 1 && thin_plan(R, Fb, K, sm_count, smem_limit).ok;
}

size_t cg_dw_thin_workspace(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit) {
    if (Fa != 1) return 0;
    const ThinPlan pl = thin_plan(R, Fb, K, sm_count, smem_limit);
    return pl.ok ? sizeof(float) * (size_t)pl.ctas * K * Fb : 0;
}

// stack: fp32 basis [K][R] (Fa = 1), T = gy [R][Fb]; dW [K][Fb] (row a*K + k with a = 0)
int cg_run_dw_thin(const float *stack, const float *T, float *dW, long long R, int Fb, int K, float *workspace, int sm_count,
                   size_t smem_limit, cudaStream_t s) {
    ThinPlan pl = thin_plan(R, Fb, K, sm_count, smem_limit);
    CG_REQUIRE(pl.ok, "cg_run_dw_thin: shape not supported (Fb=%d K=%d)", Fb, K);
    CG_REQUIRE((((uintptr_t)stack | (uintptr_t)T | (uintptr_t)workspace) & 15) == 0, "cg_run_dw_thin: unaligned tensor");
    ThinParams &tp = pl.tp;
    tp.stack = stack;
    tp.T = T;
    tp.part = workspace;
    tp.R = R;
    tp.K = K;
    tp.Fb = Fb;
    {
        CgProfScope prof("dw_thin", s);
#define CG_THIN_LAUNCH(NF, KT)                                                                                   \
    do {                                                                                                          \
        CG_CHECK_CUDA(cudaFuncSetAttribute(k_dw_thin<NF, KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); \
        k_dw_thin<NF, KT, false><<<(unsigned)pl.ctas, TT, pl.smem, s>>>(tp);                                      \
    } while (0)
#define CG_THIN_K(NF)                                \
    switch ((K + 3) / 4) {                           \
        case 1: CG_THIN_LAUNCH(NF, 4); break;        \
        case 2: CG_THIN_LAUNCH(NF, 8); break;        \
        case 3: CG_THIN_LAUNCH(NF, 12); break;       \
        case 4: CG_THIN_LAUNCH(NF, 16); break;       \
        case 5: CG_THIN_LAUNCH(NF, 20); break;       \
        case 6: CG_THIN_LAUNCH(NF, 24); break;       \
        case 7: CG_THIN_LAUNCH(NF, 28); break;       \
        default: CG_THIN_LAUNCH(NF, 32); break;      \
    }
        if (Fb == 32) {
            CG_THIN_K(1)
        } else {
            CG_THIN_K(2)
        }
#undef CG_THIN_K
#undef CG_THIN_LAUNCH
        CG_LAUNCH_CHECK();
    }
    return cg_reduce_partials(workspace, dW, pl.ctas, 1, Fb, K, false, s);
}

// ---- pooled variant -----------------------------------------------------------------------------------------------
bool cg_dw_thin_pooled_supported(long long R, int Fb, int K, int sm_count, size_t smem_limit) {
    return thin_plan(R, Fb, K, sm_count, smem_limit, true).ok;
}

size_t cg_dw_thin_pooled_workspace(long long R, int Fb, int K, int sm_count, size_t smem_limit) {
    const ThinPlan pl = thin_plan(R, Fb, K, sm_count, smem_limit, true);
    // per-CTA partials [K + 1][Fb] and the reduced [K + 1][Fb] block (weight rows, then the bias gradient)
    return pl.ok ? cg_align_up(sizeof(float) * (size_t)pl.ctas * (K + 1) * Fb, 256) + sizeof(float) * (size_t)(K + 1) * Fb : 0;
}

// stack: fp32 basis [K][R]; gp / yp / aux: gradient of the pooled output, pooled output, argmax, each [R/4][Fb];
// dW [K][Fb]; db [Fb] or NULL
int cg_run_dw_thin_pooled(const float *stack, const float *gp, const float *yp, const unsigned char *aux, float *dW, float *db,
                          long long R, int Fb, int K, float *workspace, int sm_count, size_t smem_limit, cudaStream_t s) {
    ThinPlan pl = thin_plan(R, Fb, K, sm_count, smem_limit, true);
    CG_REQUIRE(pl.ok, "cg_run_dw_thin_pooled: shape not supported (Fb=%d K=%d)", Fb, K);
    CG_REQUIRE((((uintptr_t)stack | (uintptr_t)gp | (uintptr_t)yp | (uintptr_t)aux | (uintptr_t)workspace) & 15) == 0,
               "cg_run_dw_thin_pooled: unaligned tensor");
    ThinParams &tp = pl.tp;
    tp.stack = stack;
    tp.T = gp;
    tp.yp = yp;
    tp.aux = aux;
    tp.part = workspace;
    tp.R = R;
    tp.K = K;
    tp.Fb = Fb;
    {
        CgProfScope prof("dw_thin", s);
#define CG_THIN_LAUNCH(NF, KT)                                                                                   \
    do {                                                                                                          \
        CG_CHECK_CUDA(cudaFuncSetAttribute(k_dw_thin<NF, KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); \
        k_dw_thin<NF, KT, true><<<(unsigned)pl.ctas, TT, pl.smem, s>>>(tp);                                       \
    } while (0)
#define CG_THIN_K(NF)                                \
    switch ((K + 3) / 4) {                           \
        case 1: CG_THIN_LAUNCH(NF, 4); break;        \
        case 2: CG_THIN_LAUNCH(NF, 8); break;        \
        case 3: CG_THIN_LAUNCH(NF, 12); break;       \
        case 4: CG_THIN_LAUNCH(NF, 16); break;       \
        case 5: CG_THIN_LAUNCH(NF, 20); break;       \
        case 6: CG_THIN_LAUNCH(NF, 24); break;       \
        case 7: CG_THIN_LAUNCH(NF, 28); break;       \
        default: CG_THIN_LAUNCH(NF, 32); break;      \
    }
        if (Fb == 32) {
            CG_THIN_K(1)
        } else {
            CG_THIN_K(2)
        }
#undef CG_THIN_K
#undef CG_THIN_LAUNCH
        CG_LAUNCH_CHECK();
    }
    // sum the partials into [K + 1][Fb] (rows k of a one-feature gradient), then split weight rows and bias row
    float *red = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) +
                                           cg_align_up(sizeof(float) * (size_t)pl.ctas * (K + 1) * Fb, 256));
    int rc = cg_reduce_partials(workspace, red, pl.ctas, 1, Fb, K + 1, false, s);
    if (rc != CG_OK) return rc;
    CG_CHECK_CUDA(cudaMemcpyAsync(dW, red, sizeof(float) * (size_t)K * Fb, cudaMemcpyDeviceToDevice, s));
    if (db) CG_CHECK_CUDA(cudaMemcpyAsync(db, red + (size_t)K * Fb, sizeof(float) * (size_t)Fb, cudaMemcpyDeviceToDevice, s));
    return CG_OK;
}
