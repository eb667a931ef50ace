// Weight gradient of the Chebyshev filter from the bf16 plane image the fused forward kernel leaves behind
// (TF autodiff of the matmul at lib/models.py:222-223, reached through lib/graph_model.py:296):
//
//     P[q][b] = sum_r X_q[r] * gy[r][b],      q = k*Fa + a,   r = n*M + m  (all vertex signals of the batch)
//
// k_cheb_fused (cg_fused.cu) ships its staged operand planes as they are:
//     [hi | mid][K][chunks of 128 rows][Fa/8][128 rows][8 features]   bf16
// Per (k, chunk, feature octet) that is 128 rows x 16 bytes, already in the canonical MN-major core-matrix order
// (8 rows x 16 bytes) of a tcgen05 A operand whose M index is q and whose K index is r; the octets of one k are
// adjacent, so a row tile of 128 q fetches (k, chunk) blocks of up to 32 KB with single bulk copies (the achieved
// HBM rate depends strongly on the piece size: 1 KB pieces 3.7 TB/s, 2 KB 4.9 TB/s -- measured).  This kernel converts
// nothing on the stack side: a producer warp streams the pieces with cp.async.bulk straight into the operand
// stages, the issue warp runs  hi*hi + mid*hi + hi*mid  into a TMEM accumulator that holds the whole [K*Fa][Fb]
// gradient (q on the 128 lanes of up to 512/Fb row tiles), and the 16 compute warps only split the CR x Fb block
// of gy into the B operand planes once per row chunk (prefetched into registers; reused by every row tile).
// HBM sees the stack once and gy once; every CTA writes one partial, k_reduce_partials sums them.
#include <stdlib.h>

#include <algorithm>

#include "cg_common.cuh"
#include "cg_umma.cuh"
#include "cg_fused_common.cuh"

namespace {

constexpr int PC = 512;            // compute threads
constexpr int PT = PC + 64;        // + issue warp + producer warp
constexpr int MAXST = 6;

struct DwpParams {
    int npass;                   // MMA passes per product: 3 (fp32-equivalent hi/mid split) or 1 (single-pass bf16)
    const unsigned char *planes;   // [2][K][ceil(R/128)][Fa/8][128][8] bf16
    const float *T;                // gy [R][Fb]
    float *part;                   // [CTAs][K*Fa][Fb]
    long long R, k_stride, plane_stride;
    int Fa, Fb, K, CR, tiles, tmem_cols, nchunks, nstage, nb;      // nb: gy plane buffers (1 or 2)
    uint32_t a_plane, a_stage, b_plane, b_buf, off_b, off_bar;
};

__global__ void __launch_bounds__(PT, 1) k_dw_planes(const DwpParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *full = bars;                 // [MAXST] pieces of the stage landed
    uint64_t *empty = bars + MAXST;        // [MAXST] MMAs of the stage completed
    uint64_t *bfull = bars + 2 * MAXST;    // [2] gy planes of a chunk converted (buffer chunk % nb)
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
        const int n16 = (int)((p.off_b + (uint32_t)p.nb * p.b_buf) / 16);
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
        const int OF = p.Fa / 8;                               // feature octets per k
        const uint32_t cs = (uint32_t)OF * 2048u;              // one (k, 128-row chunk) block
        int it = 0;
        for (int ch = c_beg; ch < c_end; ++ch) {
            const long long r0 = (long long)ch * CR;
            const int rows = (int)std::min<long long>(CR, p.R - r0);
            const uint32_t bytes = (uint32_t)rows * 16u;
            const size_t coff = (size_t)(r0 >> 7) * cs + (size_t)(r0 & 127) * 16;
            for (int t = 0; t < tiles; ++t, ++it) {
                const int s = it % nstage;
                if (it >= nstage) {
                    if (lane == 0) umma::mbar_wait(empty + s, (uint32_t)((it / nstage - 1) & 1));
                    __syncwarp();
                }
                const int q0 = t * 16, noct = min(16, Q / 8 - q0);      // q octets [q0, q0 + noct) of this row tile
                if (lane == 0) mbar_expect_tx(full + s, (uint32_t)(2 * noct) * bytes);
                __syncwarp();
                const uint32_t dst0 = st0 + (uint32_t)s * p.a_stage;
                if (rows == 128) {
                    // whole chunks: one copy per (plane, k) -- the octets of a k are adjacent in the image and in the stage
                    const int k_lo = q0 / OF, nk = (q0 + noct - 1) / OF - k_lo + 1;
                    for (int j = lane; j < 2 * nk; j += 32) {
                        const int pl = j >= nk ? 1 : 0, k = k_lo + (j - pl * nk);
                        const int a = max(q0, k * OF), b = min(q0 + noct, (k + 1) * OF);
                        const unsigned char *src = p.planes + (size_t)pl * p.plane_stride + (size_t)k * p.k_stride + coff + (size_t)(a - k * OF) * 2048;
                        bulk_g2s(dst0 + (uint32_t)pl * p.a_plane + (uint32_t)(a - q0) * piece, src, (uint32_t)(b - a) * 2048u, full + s);
                    }
                } else {
                    // half chunks (CR = 64) and the ragged last chunk: one copy per (plane, q octet), valid rows only
                    for (int j = lane; j < 2 * noct; j += 32) {
                        const int pl = j >= noct ? 1 : 0, jj = j - pl * noct;
                        const int k = (q0 + jj) / OF, fo = (q0 + jj) - k * OF;
                        const unsigned char *src = p.planes + (size_t)pl * p.plane_stride + (size_t)k * p.k_stride + coff + (size_t)fo * 2048;
                        bulk_g2s(dst0 + (uint32_t)pl * p.a_plane + (uint32_t)jj * piece, src, bytes, full + s);
                    }
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
            const int cl = ch - c_beg, bb = p.nb == 2 ? (cl & 1) : 0;
            if (lane == 0) umma::mbar_wait(bfull + bb, (uint32_t)((p.nb == 2 ? (cl >> 1) : cl) & 1));
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
                        if (pass >= p.npass) break;
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
        // item = (row, feature octet); a quarter-warp takes 8 consecutive rows of one octet: its 16-byte stores fill
        // one 128-byte core matrix.  The fp32 values of the NEXT chunk are loaded into registers before the wait for
        // the plane buffer, so a single buffer costs only the split + store time per chunk.
        const int OB = Fb / 8;
        constexpr int NI = 8;                                  // items per thread: CR * OB / 512 <= 128 * 32 / 512
        const int ni = (CR * OB + PC - 1) / PC;
        float4 v[NI][2];
        auto load_chunk = [&](int ch) {
            const long long r0 = (long long)ch * CR;
            const int rows = (int)std::min<long long>(CR, p.R - r0);
#pragma unroll
            for (int u = 0; u < NI; ++u) {
                const int e = tid + u * PC;
                if (u < ni && e < CR * OB) {
                    const int g = e >> 3, o = g % OB, row = (g / OB) * 8 + (e & 7);
                    v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < rows) {
                        const float4 *src = reinterpret_cast<const float4 *>(p.T + (size_t)(r0 + row) * Fb + o * 8);
                        v[u][0] = __ldg(src);
                        v[u][1] = __ldg(src + 1);
                    }
                }
            }
        };
        if (c_beg < c_end) load_chunk(c_beg);
        for (int ch = c_beg; ch < c_end; ++ch) {
            const int cl = ch - c_beg, bb = p.nb == 2 ? (cl & 1) : 0;
            if (cl >= p.nb) umma::mbar_wait(bfree + bb, (uint32_t)(((p.nb == 2 ? (cl >> 1) : cl) - 1) & 1));
            unsigned char *bp = smem + p.off_b + (size_t)bb * p.b_buf;
#pragma unroll
            for (int u = 0; u < NI; ++u) {
                const int e = tid + u * PC;
                if (u < ni && e < CR * OB) {
                    const int g = e >> 3, o = g % OB, rb = g / OB;
                    uint2 h0, m0, h1, m1;
                    split4(v[u][0], h0, m0);
                    split4(v[u][1], h1, m1);
                    const uint32_t off = (uint32_t)o * piece + (uint32_t)rb * 128u + (uint32_t)(e & 7) * 16u;
                    *reinterpret_cast<uint4 *>(bp + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                    *reinterpret_cast<uint4 *>(bp + p.b_plane + off) = make_uint4(m0.x, m0.y, m1.x, m1.y);
                }
            }
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bfull + bb);
            if (ch + 1 < c_end) load_chunk(ch + 1);
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
    dp.npass = cg_mma_passes();
    int cr0 = 128;
    if (const char *env = getenv("CG_DWP_CR")) cr0 = atoi(env);     // tuning aid: 128 or 64
    if (cr0 != 64) cr0 = 128;
    for (int CR = cr0; CR >= 64 && !pl.ok; CR /= 2) {
        dp.CR = CR;
        dp.a_plane = 16u * (uint32_t)CR * 16u;          // 16 q octets x CR rows x 16 bytes
        dp.a_stage = 2 * dp.a_plane;
        dp.b_plane = (uint32_t)(Fb / 8) * (uint32_t)CR * 16u;
        dp.b_buf = 2 * dp.b_plane;
        // prefer three operand stages with one gy buffer over two stages with two buffers
        for (int nb = 2; nb >= 1 && !pl.ok; --nb) {
            const size_t fixed = (size_t)nb * dp.b_buf + 256;
            if (fixed + 2 * (size_t)dp.a_stage > smem_limit) continue;
            const int nstage = (int)std::min<size_t>(MAXST, (smem_limit - fixed) / dp.a_stage);
            if (nb == 2 && nstage < 3 && (size_t)dp.b_buf + 256 + 3 * (size_t)dp.a_stage <= smem_limit) continue;
            dp.nb = nb;
            dp.nstage = nstage;
            dp.off_b = (uint32_t)nstage * dp.a_stage;
            dp.off_bar = dp.off_b + (uint32_t)nb * dp.b_buf;
            pl.smem = dp.off_bar + 256;
            pl.ok = true;
        }
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
    dp.k_stride = cg_ceil_div(R, 128) * (Fa / 8) * 2048;       // all 128-row chunks of one k
    dp.plane_stride = (long long)K * dp.k_stride;
    dp.Fa = Fa;
    dp.Fb = Fb;
    dp.K = K;
    {
        CgProfScope prof("dw_planes", s);
        CG_CHECK_CUDA(cudaFuncSetAttribute(k_dw_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        k_dw_planes<<<(unsigned)pl.ctas, PT, pl.smem, s>>>(dp);
        CG_LAUNCH_CHECK();
    }
    return cg_reduce_partials(workspace, dW, pl.ctas, Fa, Fb, K, false, s);
}
