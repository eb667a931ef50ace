// Weight gradient of the Chebyshev filter from the bf16 plane image the fused forward kernel leaves behind
// (TF autodiff of the matmul at lib/models.py:222-223, reached through lib/graph_model.py:296):
//
//     P[q][b] = sum_r X_q[r] * gy[r][b],      q = k*Fa + a,   r = n*M + m  (all vertex signals of the batch)
//
// k_cheb_fused (cg_fused.cu) ships its staged operand planes as they are:  [hi | mid][K][Fa/8][R rows][8 features]
// bf16.  Per (k, feature octet) that is one contiguous run over the rows, already in the canonical MN-major
// core-matrix order (8 rows x 16 bytes) of a tcgen05 A operand whose M index is q and whose K index is r.  So this
// kernel converts nothing on the stack side: a producer warp streams CR-row pieces with cp.async.bulk straight
// into the operand stages, the issue warp runs  hi*hi + mid*hi + hi*mid  into a TMEM accumulator that holds the
// whole [K*Fa][Fb] gradient (q on the 128 lanes of up to 512/Fb row tiles), and the 16 compute warps only split
// the CR x Fb block of gy into the B operand planes once per row chunk (it is reused by every row tile).
// HBM sees the stack once and gy once; every CTA writes one partial, k_reduce_partials sums them.
#include <algorithm>

#include "cg_common.cuh"
#include "cg_umma.cuh"
#include "cg_fused_common.cuh"

namespace {

constexpr int PC = 512;            // compute threads
constexpr int PT = PC + 64;        // + issue warp + producer warp
constexpr int MAXST = 6;

struct DwpParams {
    const unsigned char *planes;   // [2][K][Fa/8][R][8] bf16
    const float *T;                // gy [R][Fb]
    float *part;                   // [CTAs][K*Fa][Fb]
    long long R, kf_stride, plane_stride;
    int Fa, Fb, K, CR, tiles, tmem_cols, nchunks, nstage;
    uint32_t a_plane, a_stage, b_plane, b_buf, off_b, off_bar;
};

__global__ void __launch_bounds__(PT, 1) k_dw_planes(const DwpParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *full = bars;                 // [MAXST] pieces of the stage landed
    uint64_t *empty = bars + MAXST;        // [MAXST] MMAs of the stage completed
    uint64_t *bfull = bars + 2 * MAXST;    // [2] gy planes of a chunk converted
    uint64_t *bfree = bfull + 2;           // [2] MMAs that read the gy planes completed
    uint64_t *done = bfree + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Fb = p.Fb, CR = p.CR, tiles = p.tiles, nstage = p.nstage;
    const int Q = p.K * p.Fa;
    // contiguous range of row chunks of this CTA
    const int c_beg = (int)((long long)blockIdx.x * p.nchunks / gridDim.x);
    const int c_end = (int)((long long)(blockIdx.x + 1) * p.nchunks / gridDim.x);

    if (tid == 0) {
        for (int i = 0; i < MAXST; ++i) {
            umma::mbar_init(full + i, 1);
            umma::mbar_init(empty + i, 1);
        }
        umma::mbar_init(bfull, PC / 32);
        umma::mbar_init(bfull + 1, PC / 32);
        umma::mbar_init(bfree, 1);
        umma::mbar_init(bfree + 1, 1);
        umma::mbar_init(done, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    {   // rows of a ragged last chunk and q octets beyond K*Fa are never loaded: they must hold finite values
        uint4 *z = reinterpret_cast<uint4 *>(smem);
        const int n16 = (int)((p.off_b + 2 * p.b_buf) / 16);
        for (int i = tid; i < n16; i += PT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t st0 = umma::smem_u32(smem);
    const uint32_t piece = (uint32_t)CR * 16u;          // one q (or feature) octet of a chunk: CR rows x 16 bytes

    if (warp == PC / 32 + 1) {
        // =========================== producer warp ======================================
        int it = 0;
        for (int ch = c_beg; ch < c_end; ++ch) {
            const long long r0 = (long long)ch * CR;
            const uint32_t bytes = (uint32_t)std::min<long long>(CR, p.R - r0) * 16u;
            for (int t = 0; t < tiles; ++t, ++it) {
                const int s = it % nstage;
                if (it >= nstage) {
                    if (lane == 0) umma::mbar_wait(empty + s, (uint32_t)((it / nstage - 1) & 1));
                    __syncwarp();
                }
                const int noct = min(16, (Q - t * 128) / 8);          // q octets of this row tile
                if (lane == 0) mbar_expect_tx(full + s, (uint32_t)(2 * noct) * bytes);
                __syncwarp();
                for (int j = lane; j < 2 * noct; j += 32) {
                    const int pl = j >= noct ? 1 : 0, jj = j - pl * noct;
                    const unsigned char *src = p.planes + (size_t)pl * p.plane_stride + (size_t)(t * 16 + jj) * p.kf_stride + (size_t)r0 * 16;
                    bulk_g2s(st0 + (uint32_t)s * p.a_stage + (uint32_t)pl * p.a_plane + (uint32_t)jj * piece, src, bytes, full + s);
                }
            }
        }
    } else if (warp == PC / 32) {
        // =========================== MMA issue warp ======================================
        const uint32_t idesc = umma::make_idesc_bf16(128, Fb, 1, 1);
        // both operands MN-major: LBO = 128 bytes between groups of 8 rows (K index), SBO = one piece between octets
        const uint32_t d_hi = umma::desc_hi(piece);
        int it = 0;
        for (int ch = c_beg; ch < c_end; ++ch) {
            const int cl = ch - c_beg, bb = cl & 1;
            if (lane == 0) umma::mbar_wait(bfull + bb, (uint32_t)((cl >> 1) & 1));
            __syncwarp();
            const uint32_t b_lo = umma::desc_lo(st0 + p.off_b + (uint32_t)bb * p.b_buf, 128u);
            for (int t = 0; t < tiles; ++t, ++it) {
                const int s = it % nstage;
                if (lane == 0) umma::mbar_wait(full + s, (uint32_t)((it / nstage) & 1));
                __syncwarp();
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint32_t a_lo = umma::desc_lo(st0 + (uint32_t)s * p.a_stage, 128u);
                    const uint32_t acc = tmem + (uint32_t)(t * Fb);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        uint32_t al = a_lo + (pass == 1 ? (p.a_plane >> 4) : 0u), bl = b_lo + (pass == 2 ? (p.b_plane >> 4) : 0u);
                        for (int j = 0; j < CR / 16; ++j) {
                            umma::mma_bf16(acc, umma::desc_join(al, d_hi), umma::desc_join(bl, d_hi), idesc, (cl | pass | j) != 0);
                            al += 16u;          // 16 rows = two groups of 128 bytes
                            bl += 16u;
                        }
                    }
                    umma::commit(empty + s);
                    if (t == tiles - 1) umma::commit(bfree + bb);
                    if (t == tiles - 1 && ch == c_end - 1) umma::commit(done);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== compute warps ======================================
        const int OB = Fb / 8;
        for (int ch = c_beg; ch < c_end; ++ch) {
            const int cl = ch - c_beg, bb = cl & 1;
            const long long r0 = (long long)ch * CR;
            const int rows = (int)std::min<long long>(CR, p.R - r0);
            if (cl >= 2) umma::mbar_wait(bfree + bb, (uint32_t)(((cl >> 1) - 1) & 1));
            unsigned char *bp = smem + p.off_b + (size_t)bb * p.b_buf;
            // item = (row, feature octet); a quarter-warp takes 8 consecutive rows of one octet: its 16-byte stores
            // fill one 128-byte core matrix
            for (int e = tid; e < CR * OB; e += PC) {
                const int i = e & 7, g = e >> 3;
                const int o = g % OB, rb = g / OB;
                const int row = rb * 8 + i;
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (row < rows) {
                    const float4 *src = reinterpret_cast<const float4 *>(p.T + (size_t)(r0 + row) * Fb + o * 8);
                    v0 = __ldg(src);
                    v1 = __ldg(src + 1);
                }
                uint2 h0, m0, h1, m1;
                split4(v0, h0, m0);
                split4(v1, h1, m1);
                const uint32_t off = (uint32_t)o * piece + (uint32_t)rb * 128u + (uint32_t)i * 16u;
                *reinterpret_cast<uint4 *>(bp + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<uint4 *>(bp + p.b_plane + off) = make_uint4(m0.x, m0.y, m1.x, m1.y);
            }
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bfull + bb);
        }
        // ---- epilogue: TMEM -> partial result
        umma::mbar_wait(done, 0u);
        umma::fence_after_sync();
        const int qd = warp & 3, wq = warp >> 2;
        const int nc8 = Fb / 8;
        float *dst0 = p.part + (size_t)blockIdx.x * Q * Fb;
        for (int idx = wq; idx < tiles * nc8; idx += 4) {
            const int t = idx / nc8, c8 = idx - t * nc8;
            const int q = t * 128 + 32 * qd + lane;
            float v[8];
            umma::tmem_ld8(tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(t * Fb + c8 * 8), v);
            umma::tmem_ld_wait();
            if (q < Q) {
                float *dst = dst0 + (size_t)q * Fb + c8 * 8;
                *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

struct DwpPlan {
    bool ok = false;
    int ctas = 0;
    size_t smem = 0;
    DwpParams dp;
};

static DwpPlan dwp_plan(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit) {
    DwpPlan pl;
    if (Fa % 8 != 0 || Fb % 16 != 0 || Fb < 16 || Fb > 256 || K < 1 || R < 1) return pl;
    const int tiles = (int)cg_ceil_div((int64_t)K * Fa, 128);
    if (tiles * Fb > 512) return pl;
    DwpParams dp;
    memset(&dp, 0, sizeof(dp));
    for (int CR = 64; CR >= 32 && !pl.ok; CR /= 2) {
        dp.CR = CR;
        dp.a_plane = 16u * (uint32_t)CR * 16u;          // 16 q octets x CR rows x 16 bytes
        dp.a_stage = 2 * dp.a_plane;
        dp.b_plane = (uint32_t)(Fb / 8) * (uint32_t)CR * 16u;
        dp.b_buf = 2 * dp.b_plane;
        const size_t fixed = 2 * (size_t)dp.b_buf + 256;
        if (fixed + 2 * (size_t)dp.a_stage > smem_limit) continue;
        dp.nstage = (int)std::min<size_t>(MAXST, (smem_limit - fixed) / dp.a_stage);
        dp.off_b = (uint32_t)dp.nstage * dp.a_stage;
        dp.off_bar = dp.off_b + 2 * dp.b_buf;
        pl.smem = dp.off_bar + 256;
        pl.ok = true;
    }
    if (!pl.ok) return pl;
    dp.tiles = tiles;
    int cols = 32;
    while (cols < tiles * Fb) cols *= 2;
    dp.tmem_cols = cols;
    dp.nchunks = (int)cg_ceil_div(R, dp.CR);
    pl.ctas = std::min(sm_count, dp.nchunks);
    pl.dp = dp;
    return pl;
}

}  // namespace

bool cg_dw_planes_supported(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit) {
    return R < (1LL << 31) * 32 && dwp_plan(R, Fa, Fb, K, sm_count, smem_limit).ok;
}

size_t cg_dw_planes_workspace(long long R, int Fa, int Fb, int K, int sm_count, size_t smem_limit) {
    const DwpPlan pl = dwp_plan(R, Fa, Fb, K, sm_count, smem_limit);
    return pl.ok ? sizeof(float) * (size_t)pl.ctas * K * Fa * Fb : 0;
}

// planes: the image written by cg_run_fused(..., stack_planes = true); T = gy [R][Fb]; dW [Fa*K][Fb] (row a*K + k)
int cg_run_dw_planes(const void *planes, const float *T, float *dW, long long R, int Fa, int Fb, int K, float *workspace,
                     int sm_count, size_t smem_limit, cudaStream_t s) {
    DwpPlan pl = dwp_plan(R, Fa, Fb, K, sm_count, smem_limit);
    CG_REQUIRE(pl.ok, "cg_run_dw_planes: shape not supported (Fa=%d Fb=%d K=%d)", Fa, Fb, K);
    CG_REQUIRE((((uintptr_t)planes | (uintptr_t)T | (uintptr_t)workspace) & 15) == 0, "cg_run_dw_planes: unaligned tensor");
    DwpParams &dp = pl.dp;
    dp.planes = reinterpret_cast<const unsigned char *>(planes);
    dp.T = T;
    dp.part = workspace;
    dp.R = R;
    dp.kf_stride = R * 16;
    dp.plane_stride = (long long)K * (Fa / 8) * dp.kf_stride;
    dp.Fa = Fa;
    dp.Fb = Fb;
    dp.K = K;
    {
        CgProfScope prof("dw_umma", s);
        CG_CHECK_CUDA(cudaFuncSetAttribute(k_dw_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        k_dw_planes<<<(unsigned)pl.ctas, PT, pl.smem, s>>>(dp);
        CG_LAUNCH_CHECK();
    }
    return cg_reduce_partials(workspace, dW, pl.ctas, Fa, Fb, K, false, s);
}
