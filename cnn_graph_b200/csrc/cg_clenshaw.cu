// Input gradient of the Chebyshev filter by the adjoint (Clenshaw) recurrence -- what TF's autodiff of
// lib/models.py:205-223 computes (lib/graph_model.py:296), fused into one persistent kernel:
//
//     G_k = gy W_k^T                                 [R][Fi]    tcgen05.mma, A = gy resident in TENSOR MEMORY
//     b_k = G_k + 2 L~^T b_{k+1} - b_{k+2}           k = K-1 .. 1   (b_K = b_{K+1} = 0)
//     dx  = G_0 +   L~^T b_1     - b_2
//
// so the recurrence runs at the width of dx (Fi = Fin of the filter) instead of the width of gy, and neither
// Z_k = T_k(L~^T) gy nor the G_k ever exist in HBM.  Per group of S samples (rows r = s*M + m):
//   * gy rows are split into bf16 hi + mid and written to TMEM with tcgen05.st (thread = row);
//   * the issue warp streams W_k^T (bulk copies) and issues  hi*Whi + mid*Whi + hi*Wmid  into a ring of
//     TMEM accumulators, NS steps ahead of the recurrence;
//   * the compute warps move G_k from TMEM (thread = row) to a swizzled shared-memory buffer, from which the
//     gather-mapped threads (8 lanes per row, as in cg_fused.cu) pick it up while they apply L~^T.
#include <stdlib.h>

#include <algorithm>

#include "cg_common.cuh"
#include "cg_umma.cuh"
#include "cg_fused_common.cuh"

// clock64 stamps for scripts/prof_fused.py: compiled in only with -DCG_TRACE_BUILD (CG_TRACE_BUILD=1 python -m
// cnn_graph_b200.build --force); even a predicated-off stamp costs issue slots in the hot loops
#ifdef CG_TRACE_BUILD
#define CG_STAMP(cond, idx)                    \
    do {                                       \
        if (cond) p.trace[idx] = clock64();    \
    } while (0)
#else
#define CG_STAMP(cond, idx) \
    do {                    \
        (void)(cond);       \
    } while (0)
#endif

namespace {

constexpr int CC = 512;        // compute threads
constexpr int CT = CC + 32;    // + issue warp
constexpr int MAX_NS = 6;

struct ClenshawParams {
    int npass;                   // MMA passes per product: 3 (fp32-equivalent hi/mid split) or 1 (single-pass bf16)
    const int *rowptr;
    const int *col;
    const float *val;
    const int *order;            // rows by descending length
    const float *gy;             // [N][M][Fo]
    const unsigned char *wp;     // packed W^T: [K][hi|mid][Fo*Fi] bf16, B operand (n = fi, q = fo), K-major
    float *dx;                   // [N][M][Fi]
    int N, M, Fi, Fo, K, S, tiles, tmem_cols, ns, estride;
    uint32_t off_ent, off_slab, slab_bytes, off_gbuf, off_w, wplane_bytes, off_bar;
    BlkTables bt;                // row-block form of L~^T (k_cheb_clenshaw_b); its tables live at off_ent
    uint32_t off_wsz;
    long long *trace;            // optional (debug): clock64 stamps of CTA 0's second group, [K][10] (k_cheb_clenshaw_b)
};

// LPR lanes per row (Fi = 4 * LPR), IPT items (row, 4-column chunk) per compute thread
template <int LPR, int IPT>
__global__ void __launch_bounds__(CT, 1) k_cheb_clenshaw(const ClenshawParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float2 *ent = reinterpret_cast<float2 *>(smem + p.off_ent);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *wbar = bars;              // [2] W_k landed
    uint64_t *gfull = bars + 2;         // [ns] MMAs of the G in this TMEM slot completed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 + MAX_NS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = p.M, Fi = p.Fi, Fo = p.Fo, K = p.K, S = p.S, NS = p.ns;
    const int R = S * M;
    const bool is_issuer = warp == CC / 32;
    constexpr uint32_t SWZ = LPR >= 8 ? 7u : (uint32_t)(LPR - 1);     // chunk swizzle mask of the G buffer

    // ---- one-time setup ------------------------------------------------------------
    for (int i = tid; i < M * p.estride; i += CT) {
        const int m = i / p.estride, j = i - m * p.estride;
        const int b = p.rowptr[m], n = p.rowptr[m + 1] - b;
        float2 v = make_float2(0.f, __int_as_float(0));
        if (j < n) {
            v.x = p.val[b + j];
            v.y = __int_as_float(p.col[b + j] * Fi * 4);
        }
        ent[i] = v;
    }
    if (tid == 0) {
        for (int i = 0; i < 2 + MAX_NS; ++i) umma::mbar_init(bars + i, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);

    // gather-mapped items: same dealing as cg_fused.cu (rows by descending length, pairs adjacent)
    uint32_t a_ent[IPT], a_g[IPT], a_soff[IPT], a_goff[IPT];
    int nlen[IPT];
    int wl[IPT / 2];
    const uint32_t ent0 = umma::smem_u32(ent);
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int qd = ((i >> 1) & 1) ? (CC / LPR - 1 - tid / LPR) : tid / LPR;
        const int o = ((i >> 1) * (CC / LPR) + qd) * 2 + (i & 1), l = tid % LPR;
        nlen[i] = 0;
        a_ent[i] = ent0;
        a_g[i] = 0;
        a_soff[i] = 0xFFFFFFF0u;     // "absent": compares above every group's limit
        a_goff[i] = 0;
        if (!is_issuer && o < R) {
            const int s = o % S, m = p.order[o / S];
            const int r = s * M + m;
            a_ent[i] = ent0 + 8u * (uint32_t)(m * p.estride);
            nlen[i] = p.rowptr[m + 1] - p.rowptr[m];
            a_g[i] = 4u * (uint32_t)(s * M * Fi + 4 * l);
            a_soff[i] = 4u * (uint32_t)(r * Fi + 4 * l);
            a_goff[i] = 4u * (uint32_t)(r * Fi) + 16u * ((uint32_t)l ^ ((uint32_t)r & SWZ));
        }
    }
#pragma unroll
    for (int pr = 0; pr < IPT / 2; ++pr)
        wl[pr] = __reduce_max_sync(0xffffffffu, max(nlen[2 * pr], nlen[2 * pr + 1]));
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int G = (p.N + S - 1) / S;
    const uint32_t slab0 = umma::smem_u32(smem + p.off_slab);
    const uint32_t gbuf0 = umma::smem_u32(smem + p.off_gbuf);
    const uint32_t wbytes = 2 * p.wplane_bytes;
    const uint32_t w0 = umma::smem_u32(smem + p.off_w);
    const uint32_t g_col0 = (uint32_t)(p.tiles * Fo);            // TMEM: A planes first, then the G ring
    const uint32_t g_slot = (uint32_t)(p.tiles * Fi);
    uint32_t gpar = 0;       // parity of the next phase of every gfull barrier (bit = slot)
    uint32_t wpar = 0;       // issuer: parity of the two W barriers

    for (int g = blockIdx.x; g < G; g += gridDim.x) {
        const int n0 = g * S;
        const int Sg = min(S, p.N - n0);
        const int Rg = Sg * M;
        const uint32_t limb = 4u * (uint32_t)(Rg * Fi);

        if (is_issuer) {
            // =========================== issue warp =====================================
            if (lane == 0) {
                mbar_expect_tx(wbar, wbytes);
                bulk_g2s(w0, p.wp + (size_t)(K - 1) * wbytes, wbytes, wbar);
                if (K > 1) {
                    mbar_expect_tx(wbar + 1, wbytes);
                    bulk_g2s(w0 + wbytes, p.wp + (size_t)(K - 2) * wbytes, wbytes, wbar + 1);
                }
            }
            const uint32_t idesc = umma::make_idesc_bf16(128, Fi, 0, 0);
            const uint32_t lbo_w = (uint32_t)Fi * 16u;
            const uint32_t d_hi = umma::desc_hi(128u);
            const int nk16 = Fo / 16;
            // G_k into ring slot (K-1-k) % NS; lane 0 only
            auto issue = [&](int k) {
                const int slot = (K - 1 - k) % NS, bw = (K - 1 - k) & 1;
                umma::mbar_wait(wbar + bw, (wpar >> bw) & 1u);
                wpar ^= 1u << bw;
                umma::fence_after_sync();
                const uint32_t b_lo = umma::desc_lo(w0 + (uint32_t)bw * wbytes, lbo_w);
                const uint32_t b_mid = p.wplane_bytes >> 4, b_k = (2u * lbo_w) >> 4;
                for (int t = 0; t < p.tiles; ++t) {
                    const uint32_t acc = tmem + g_col0 + (uint32_t)slot * g_slot + (uint32_t)(t * Fi);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        if (pass >= p.npass) break;
                        uint32_t a_col = tmem + (uint32_t)(t * Fo) + (pass == 1 ? (uint32_t)(Fo / 2) : 0u);
                        uint32_t bl = b_lo + (pass == 2 ? b_mid : 0u);
                        for (int j = 0; j < nk16; ++j) {
                            umma::mma_bf16_ts(acc, a_col, umma::desc_join(bl, d_hi), idesc, (pass | j) != 0);
                            a_col += 8u;
                            bl += b_k;
                        }
                    }
                }
                umma::commit(gfull + slot);
                // the MMAs of step k+1 have certainly read their W buffer: wait (keeps the parities in step) and
                // refill it with W_{k-1}
                if (k + 1 < K) {
                    const int sp = (K - 2 - k) % NS;
                    umma::mbar_wait(gfull + sp, (gpar >> sp) & 1u);
                    gpar ^= 1u << sp;
                    if (k >= 1) {
                        const int bn = (K - k) & 1;
                        mbar_expect_tx(wbar + bn, wbytes);
                        bulk_g2s(w0 + (uint32_t)bn * wbytes, p.wp + (size_t)(k - 1) * wbytes, wbytes, wbar + bn);
                    }
                }
                if (k == 0) {       // last of the group: wait for it as well
                    umma::mbar_wait(gfull + slot, (gpar >> slot) & 1u);
                    gpar ^= 1u << slot;
                }
            };
            __syncthreads();                         // sync A: gy of this group is in tensor memory
            if (lane == 0 && umma::elect_lane0())
                for (int k = K - 1; k >= 0 && k > K - 1 - NS; --k) issue(k);
            __syncwarp();
            for (int j = 0; j <= K; ++j) {
                __syncthreads();                     // sync j: G_{K-1-j} has left its slot
                const int kn = K - 1 - j - NS;
                if (kn >= 0 && lane == 0 && umma::elect_lane0()) issue(kn);
                __syncwarp();
            }
            // the compute warps waited for every slot once per use as well
            // (their parity mask advances identically; nothing to do here)
        } else {
            // =========================== compute warps ==================================
            const int q = warp & 3, sub = warp >> 2;
            const uint32_t lane_base = (uint32_t)(32 * q) << 16;
            // ---- gy rows -> bf16 hi | mid planes in tensor memory (thread = row, 16 features per slice)
            for (int t = 0; t < p.tiles; ++t) {
                const int r = t * 128 + 32 * q + lane;
                const float *src = p.gy + ((size_t)n0 * M + r) * Fo;
                for (int sl = sub; sl < Fo / 16; sl += 4) {
                    uint32_t hi[8], mid[8];
                    if (r < Rg) {
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float4 v = *reinterpret_cast<const float4 *>(src + sl * 16 + h * 4);
                            uint2 a, b;
                            split4(v, a, b);
                            hi[2 * h] = a.x;
                            hi[2 * h + 1] = a.y;
                            mid[2 * h] = b.x;
                            mid[2 * h + 1] = b.y;
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < 8; ++h) hi[h] = mid[h] = 0u;
                    }
                    umma::tmem_st8(tmem + lane_base + (uint32_t)(t * Fo + sl * 8), hi);
                    umma::tmem_st8(tmem + lane_base + (uint32_t)(t * Fo + Fo / 2 + sl * 8), mid);
                }
            }
            umma::tmem_st_wait();
            umma::fence_before_sync();
            __syncthreads();                         // sync A

            // G_k: TMEM slot -> swizzled shared-memory buffer (k & 1)
            auto dump = [&](int k) {
                const int slot = (K - 1 - k) % NS;
                umma::mbar_wait(gfull + slot, (gpar >> slot) & 1u);
                gpar ^= 1u << slot;
                umma::fence_after_sync();
                const uint32_t gb = gbuf0 + (uint32_t)(k & 1) * p.slab_bytes;
                const int nc8 = Fi / 8;
                for (int t = 0; t < p.tiles; ++t) {
                    const int r = t * 128 + 32 * q + lane;
                    for (int c = sub; c < nc8; c += 4) {
                        float v[8];
                        umma::tmem_ld8(tmem + lane_base + g_col0 + (uint32_t)slot * g_slot + (uint32_t)(t * Fi + c * 8), v);
                        umma::tmem_ld_wait();
                        if (r < Rg) {
                            const uint32_t row = gb + 4u * (uint32_t)(r * Fi);
                            sts128(row + 16u * ((uint32_t)(2 * c) ^ ((uint32_t)r & SWZ)), make_float4(v[0], v[1], v[2], v[3]));
                            sts128(row + 16u * ((uint32_t)(2 * c + 1) ^ ((uint32_t)r & SWZ)), make_float4(v[4], v[5], v[6], v[7]));
                        }
                    }
                }
            };

            dump(K - 1);
            umma::fence_before_sync();
            __syncthreads();                         // sync 0

            float4 res[IPT], old[IPT];               // the thread's own b_{k+1} (b_k after the step) and b_{k+2}
#pragma unroll
            for (int i = 0; i < IPT; ++i) res[i] = old[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = 0; s < K; ++s) {
                const int k = K - 1 - s;
                const uint32_t prev = slab0 + (uint32_t)((s + 1) & 1) * p.slab_bytes;     // b_{k+1}
                const uint32_t cur = slab0 + (uint32_t)(s & 1) * p.slab_bytes;            // receives b_k
                const uint32_t gb = gbuf0 + (uint32_t)(k & 1) * p.slab_bytes;
                const float c2 = k > 0 ? 2.f : 1.f;
                char *dxp = reinterpret_cast<char *>(p.dx + (size_t)n0 * M * Fi);
#pragma unroll
                for (int pr = 0; pr < IPT / 2; ++pr) {
                    const int i0 = 2 * pr, i1 = 2 * pr + 1;
                    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
                    if (s > 0) {
                        const uint32_t g0 = prev + a_g[i0], g1 = prev + a_g[i1];
                        const int trips = wl[pr];
                        const int jlast = p.estride - 2;
                        float4 e0 = lds128(a_ent[i0]), e1 = lds128(a_ent[i1]);     // two {weight, offset} entries each
                        for (int j = 0; j < trips; j += 2) {
                            const float4 x00 = lds128(g0 + (uint32_t)__float_as_int(e0.y));
                            const float4 x10 = lds128(g1 + (uint32_t)__float_as_int(e1.y));
                            const float4 x01 = lds128(g0 + (uint32_t)__float_as_int(e0.w));
                            const float4 x11 = lds128(g1 + (uint32_t)__float_as_int(e1.w));
                            const float w00 = e0.x, w01 = e0.z, w10 = e1.x, w11 = e1.z;
                            const uint32_t jn = 8u * (uint32_t)min(j + 2, jlast);
                            e0 = lds128(a_ent[i0] + jn);
                            e1 = lds128(a_ent[i1] + jn);
                            fma4(acc0, w00, x00);
                            fma4(acc1, w10, x10);
                            fma4(acc0, w01, x01);
                            fma4(acc1, w11, x11);
                        }
                    }
                    const bool v0 = a_soff[i0] < limb, v1 = a_soff[i1] < limb;
                    float4 G0 = make_float4(0.f, 0.f, 0.f, 0.f), G1 = G0;
                    if (v0) G0 = lds128(gb + a_goff[i0]);
                    if (v1) G1 = lds128(gb + a_goff[i1]);
                    const float4 o0 = old[i0], o1 = old[i1];
                    // b_k = G_k + c L^T b_{k+1} - b_{k+2}
                    acc0 = make_float4(fmaf(c2, acc0.x, G0.x) - o0.x, fmaf(c2, acc0.y, G0.y) - o0.y,
                                       fmaf(c2, acc0.z, G0.z) - o0.z, fmaf(c2, acc0.w, G0.w) - o0.w);
                    acc1 = make_float4(fmaf(c2, acc1.x, G1.x) - o1.x, fmaf(c2, acc1.y, G1.y) - o1.y,
                                       fmaf(c2, acc1.z, G1.z) - o1.z, fmaf(c2, acc1.w, G1.w) - o1.w);
                    old[i0] = res[i0];
                    old[i1] = res[i1];
                    res[i0] = acc0;
                    res[i1] = acc1;
                    if (k > 0) {
                        if (v0) sts128(cur + a_soff[i0], acc0);
                        if (v1) sts128(cur + a_soff[i1], acc1);
                    } else {
                        if (v0) *reinterpret_cast<float4 *>(dxp + a_soff[i0]) = acc0;
                        if (v1) *reinterpret_cast<float4 *>(dxp + a_soff[i1]) = acc1;
                    }
                }
                if (k > 0) dump(k - 1);
                umma::fence_before_sync();
                __syncthreads();                     // sync s+1
            }
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// The same kernel with the row-block gather of cg_fused_common.cuh (a compute thread owns IPB items of 4 consecutive
// rows x 4 features); issue warp, TMEM ring and the G dump are unchanged.
template <int LPR, int IPB>
__global__ void __launch_bounds__(CT, 1) k_cheb_clenshaw_b(const ClenshawParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *wbar = bars;              // [2] W_k landed
    uint64_t *gfull = bars + 2;         // [ns] MMAs of the G in this TMEM slot completed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 + MAX_NS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int Fi = 4 * LPR;                  // compile-time row width
    const int M = p.M, Fo = p.Fo, K = p.K, S = p.S, NS = p.ns;
    const bool is_issuer = warp == CC / 32;
    constexpr uint32_t SWZ = LPR >= 8 ? 7u : (uint32_t)(LPR - 1);     // chunk swizzle mask of the G buffer
    constexpr uint32_t rowb = (uint32_t)Fi * 4u;

    if (tid == 0) {
        for (int i = 0; i < 2 + MAX_NS; ++i) umma::mbar_init(bars + i, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);

    BlkItems<IPB> it;
    blk_setup<LPR, IPB, CC>(p.bt, S, rowb, umma::smem_u32(smem + p.off_ent), reinterpret_cast<int *>(smem + p.off_wsz), it);
    uint32_t a_g[IPB], a_soff[IPB], a_goff[IPB][4];
    int nrow[IPB];
    const uint32_t lc = (uint32_t)(tid % LPR);
#pragma unroll
    for (int i = 0; i < IPB; ++i) {
        const int r0 = it.samp[i] * M + it.row0[i];
        a_g[i] = 4u * (uint32_t)(it.samp[i] * M * Fi) + 16u * lc;
        a_soff[i] = 4u * (uint32_t)(r0 * Fi) + 16u * lc;
        nrow[i] = it.samp[i] < S ? min(4, M - it.row0[i]) : 0;
#pragma unroll
        for (int r = 0; r < 4; ++r)      // position of the thread's chunk of row r0 + r in the swizzled G buffer
            a_goff[i][r] = 4u * (uint32_t)((r0 + r) * Fi) + 16u * (lc ^ ((uint32_t)(r0 + r) & SWZ));
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int G = (p.N + S - 1) / S;
    const uint32_t slab0 = umma::smem_u32(smem + p.off_slab);
    const uint32_t gbuf0 = umma::smem_u32(smem + p.off_gbuf);
    const uint32_t wbytes = 2 * p.wplane_bytes;
    const uint32_t w0 = umma::smem_u32(smem + p.off_w);
    const uint32_t g_col0 = (uint32_t)(p.tiles * Fo);            // TMEM: A planes first, then the G ring
    const uint32_t g_slot = (uint32_t)(p.tiles * Fi);
    uint32_t gpar = 0;
    uint32_t wpar = 0;

    int gi = 0;
    for (int g = blockIdx.x; g < G; g += gridDim.x, ++gi) {
        const int n0 = g * S;
        const bool trg = p.trace != nullptr && blockIdx.x == 0 && gi == 1;
        const int Sg = min(S, p.N - n0);
        const int Rg = Sg * M;

        if (is_issuer) {
            // =========================== issue warp (as k_cheb_clenshaw) ================
            if (lane == 0) {
                mbar_expect_tx(wbar, wbytes);
                bulk_g2s(w0, p.wp + (size_t)(K - 1) * wbytes, wbytes, wbar);
                if (K > 1) {
                    mbar_expect_tx(wbar + 1, wbytes);
                    bulk_g2s(w0 + wbytes, p.wp + (size_t)(K - 2) * wbytes, wbytes, wbar + 1);
                }
            }
            const uint32_t idesc = umma::make_idesc_bf16(128, Fi, 0, 0);
            const uint32_t lbo_w = (uint32_t)Fi * 16u;
            const uint32_t d_hi = umma::desc_hi(128u);
            const int nk16 = Fo / 16;
            auto issue = [&](int k) {
                const int slot = (K - 1 - k) % NS, bw = (K - 1 - k) & 1;
                CG_STAMP(trg, (K - 1 - k) * 10 + 5);
                umma::mbar_wait(wbar + bw, (wpar >> bw) & 1u);
                wpar ^= 1u << bw;
                umma::fence_after_sync();
                const uint32_t b_lo = umma::desc_lo(w0 + (uint32_t)bw * wbytes, lbo_w);
                const uint32_t b_mid = p.wplane_bytes >> 4, b_k = (2u * lbo_w) >> 4;
                for (int t = 0; t < p.tiles; ++t) {
                    const uint32_t acc = tmem + g_col0 + (uint32_t)slot * g_slot + (uint32_t)(t * Fi);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        if (pass >= p.npass) break;
                        uint32_t a_col = tmem + (uint32_t)(t * Fo) + (pass == 1 ? (uint32_t)(Fo / 2) : 0u);
                        uint32_t bl = b_lo + (pass == 2 ? b_mid : 0u);
                        for (int j = 0; j < nk16; ++j) {
                            umma::mma_bf16_ts(acc, a_col, umma::desc_join(bl, d_hi), idesc, (pass | j) != 0);
                            a_col += 8u;
                            bl += b_k;
                        }
                    }
                }
                umma::commit(gfull + slot);
                CG_STAMP(trg, (K - 1 - k) * 10 + 6);
                if (k + 1 < K) {
                    const int sp = (K - 2 - k) % NS;
                    umma::mbar_wait(gfull + sp, (gpar >> sp) & 1u);
                    gpar ^= 1u << sp;
                    if (k >= 1) {
                        const int bn = (K - k) & 1;
                        mbar_expect_tx(wbar + bn, wbytes);
                        bulk_g2s(w0 + (uint32_t)bn * wbytes, p.wp + (size_t)(k - 1) * wbytes, wbytes, wbar + bn);
                    }
                }
                if (k == 0) {
                    umma::mbar_wait(gfull + slot, (gpar >> slot) & 1u);
                    gpar ^= 1u << slot;
                }
                CG_STAMP(trg, (K - 1 - k) * 10 + 7);
            };
            __syncthreads();                         // sync A: gy of this group is in tensor memory
            if (lane == 0 && umma::elect_lane0())
                for (int k = K - 1; k >= 0 && k > K - 1 - NS; --k) issue(k);
            __syncwarp();
            for (int j = 0; j <= K; ++j) {
                __syncthreads();                     // sync j: G_{K-1-j} has left its slot
                const int kn = K - 1 - j - NS;
                if (kn >= 0 && lane == 0 && umma::elect_lane0()) issue(kn);
                __syncwarp();
            }
        } else {
            // =========================== compute warps ==================================
            const int q = warp & 3, sub = warp >> 2;
            const uint32_t lane_base = (uint32_t)(32 * q) << 16;
            for (int t = 0; t < p.tiles; ++t) {
                const int r = t * 128 + 32 * q + lane;
                const float *src = p.gy + ((size_t)n0 * M + r) * Fo;
                for (int sl = sub; sl < Fo / 16; sl += 4) {
                    uint32_t hi[8], mid[8];
                    if (r < Rg) {
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float4 v = *reinterpret_cast<const float4 *>(src + sl * 16 + h * 4);
                            uint2 a, b;
                            split4(v, a, b);
                            hi[2 * h] = a.x;
                            hi[2 * h + 1] = a.y;
                            mid[2 * h] = b.x;
                            mid[2 * h + 1] = b.y;
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < 8; ++h) hi[h] = mid[h] = 0u;
                    }
                    umma::tmem_st8(tmem + lane_base + (uint32_t)(t * Fo + sl * 8), hi);
                    umma::tmem_st8(tmem + lane_base + (uint32_t)(t * Fo + Fo / 2 + sl * 8), mid);
                }
            }
            umma::tmem_st_wait();
            umma::fence_before_sync();
            __syncthreads();                         // sync A

            auto dump = [&](int k) {
                const int slot = (K - 1 - k) % NS;
                umma::mbar_wait(gfull + slot, (gpar >> slot) & 1u);
                gpar ^= 1u << slot;
                umma::fence_after_sync();
                const uint32_t gb = gbuf0 + (uint32_t)(k & 1) * p.slab_bytes;
                constexpr int nc8 = Fi / 8;
                for (int t = 0; t < p.tiles; ++t) {
                    const int r = t * 128 + 32 * q + lane;
                    for (int c = sub; c < nc8; c += 4) {
                        float v[8];
                        umma::tmem_ld8(tmem + lane_base + g_col0 + (uint32_t)slot * g_slot + (uint32_t)(t * Fi + c * 8), v);
                        umma::tmem_ld_wait();
                        if (r < Rg) {
                            const uint32_t row = gb + 4u * (uint32_t)(r * Fi);
                            sts128(row + 16u * ((uint32_t)(2 * c) ^ ((uint32_t)r & SWZ)), make_float4(v[0], v[1], v[2], v[3]));
                            sts128(row + 16u * ((uint32_t)(2 * c + 1) ^ ((uint32_t)r & SWZ)), make_float4(v[4], v[5], v[6], v[7]));
                        }
                    }
                }
            };

            dump(K - 1);
            umma::fence_before_sync();
            __syncthreads();                         // sync 0

            int nr[IPB];
            constexpr bool OLD_REGS = IPB == 1;      // two items per thread: b_{k+2} is read back from the slab it sits in
            float4 res[IPB][4], old[OLD_REGS ? IPB : 1][4];
#pragma unroll
            for (int i = 0; i < IPB; ++i) {
                nr[i] = it.samp[i] < Sg ? nrow[i] : 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) res[i][r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if constexpr (OLD_REGS) {
#pragma unroll
                for (int r = 0; r < 4; ++r) old[0][r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const bool trc = trg && tid == 0;
            for (int s = 0; s < K; ++s) {
                const int k = K - 1 - s;
                CG_STAMP(trc, s * 10 + 0);
                const uint32_t prev = slab0 + (uint32_t)((s + 1) & 1) * p.slab_bytes;     // b_{k+1}
                const uint32_t cur = slab0 + (uint32_t)(s & 1) * p.slab_bytes;            // receives b_k (holds b_{k+2})
                const uint32_t gb = gbuf0 + (uint32_t)(k & 1) * p.slab_bytes;
                const float c2 = k > 0 ? 2.f : 1.f;
                char *dxp = reinterpret_cast<char *>(p.dx + (size_t)n0 * M * Fi);
#pragma unroll
                for (int i = 0; i < IPB; ++i) {
                    float4 acc[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s > 0) blk_gather<LPR>(prev + a_g[i], it.tab[i], it.trips[i], acc);
                    CG_STAMP(trc && i == 0, s * 10 + 1);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const bool v = r < nr[i];
                        float4 Gv = make_float4(0.f, 0.f, 0.f, 0.f), o = Gv;
                        if (v) Gv = lds128(gb + a_goff[i][r]);
                        if constexpr (OLD_REGS) {
                            o = old[i][r];
                        } else {
                            if (v && s >= 2) o = lds128(cur + a_soff[i] + (uint32_t)r * rowb);
                        }
                        // b_k = G_k + c L^T b_{k+1} - b_{k+2}
                        acc[r] = make_float4(fmaf(c2, acc[r].x, Gv.x) - o.x, fmaf(c2, acc[r].y, Gv.y) - o.y,
                                             fmaf(c2, acc[r].z, Gv.z) - o.z, fmaf(c2, acc[r].w, Gv.w) - o.w);
                        if constexpr (OLD_REGS) old[i][r] = res[i][r];
                        res[i][r] = acc[r];
                        if (v) {
                            if (k > 0)
                                sts128(cur + a_soff[i] + (uint32_t)r * rowb, acc[r]);
                            else
                                *reinterpret_cast<float4 *>(dxp + a_soff[i] + (uint32_t)r * rowb) = acc[r];
                        }
                    }
                }
                CG_STAMP(trc, s * 10 + 2);
                if (k > 0) dump(k - 1);
                CG_STAMP(trc, s * 10 + 3);
                umma::fence_before_sync();
                __syncthreads();                     // sync s+1
                CG_STAMP(trc, s * 10 + 4);
            }
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

struct CPlan {
    bool ok = false;
    int S = 0, ipt = 0;
    bool blocked = false;
    ClenshawParams cp;
    size_t smem = 0;
};

static CPlan make_cplan(const cg_graph *g, int N, int Fi, int Fo, int K, bool blocked) {
    CPlan best;
    if (Fi % 16 != 0 || Fi > 128 || (Fi & (Fi - 1)) != 0) return best;     // LPR in {4, 8, 16, 32}
    if (Fo % 16 != 0 || Fo < 16 || Fo > 256) return best;
    if (N <= 0 || K < 1) return best;
    const CgCsr &side = g->adj;
    if (blocked && side.nblk == 0) return best;
    const int M = g->M, LPR = Fi / 4, width = side.width;
    const double avg = M > 0 ? (double)(blocked ? 4 * (int64_t)side.blk_total : g->nnz) / M : 0.0;
    const int estride = std::max(2, (width + 1) & ~1);
    double best_cost = 0.0;
    int s_lo = 1, s_hi = 64;
    if (const char *env = getenv("CG_FUSED_S")) {
        const int v = atoi(env);
        if (v > 0) s_lo = s_hi = v;
    }
    for (int S = s_lo; S <= N && S <= s_hi; ++S) {
        const int64_t R = (int64_t)S * M;
        const int tiles = (int)cg_ceil_div(R, 128);
        // TMEM: tiles * Fo columns of gy planes + at least two G slots of tiles * Fi columns
        const int ns = std::min(MAX_NS, (512 - tiles * Fo) / (tiles * Fi));
        if (tiles * Fo > 512 || ns < 2) break;
        int need, ipt;
        if (blocked) {
            need = (int)cg_ceil_div((int64_t)S * side.nblk * LPR, CC);
            if (need > 2) break;
            ipt = need;
        } else {
            need = (int)cg_ceil_div(R * LPR, CC);
            if (need > 8) break;
            ipt = need <= 2 ? 2 : need <= 4 ? 4 : 8;
        }
        const uint32_t slab = (uint32_t)cg_align_up((size_t)R * Fi * 4, 128);
        const uint32_t wplane = (uint32_t)Fi * Fo * 2u;
        const size_t ent_bytes = blocked ? cg_blk_table_bytes(side.blk_len_sorted, S, LPR, ipt, CC) : (size_t)M * estride * 8;
        ClenshawParams cp;
        memset(&cp, 0, sizeof(cp));
    cp.npass = cg_mma_passes();
        uint32_t off = 0;
        cp.off_bar = off;
        off += 128;
        cp.off_wsz = off;
        off += 256;
        cp.off_ent = off;
        off += (uint32_t)cg_align_up(ent_bytes, 128);
        cp.off_slab = off;
        off += 2 * slab;
        cp.off_gbuf = off;
        off += 2 * slab;
        cp.off_w = off;
        off += 4 * wplane;
        if (off > g->smem_optin) continue;
        const int64_t G = cg_ceil_div(N, S);
        const int64_t rounds = cg_ceil_div(G, g->sm_count);
        const double step = blocked ? (double)need * (avg * 5.5 + 140.0) + 300.0
                                    : (double)need * (avg * 7.0 + 60.0) * (need > 4 ? 1.6 : 1.0) + 300.0;
        const double cost = (double)rounds * ((double)K * step + 1500.0);
        if (!best.ok || cost < best_cost) {
            best.ok = true;
            best_cost = cost;
            best.S = S;
            best.ipt = ipt;
            best.blocked = blocked;
            cp.S = S;
            cp.tiles = tiles;
            cp.ns = ns;
            int cols = 32;
            while (cols < tiles * (Fo + ns * Fi)) cols *= 2;
            cp.tmem_cols = cols;
            cp.estride = estride;
            cp.slab_bytes = slab;
            cp.wplane_bytes = wplane;
            best.cp = cp;
            best.smem = off;
        }
    }
    return best;
}

static CPlan choose_cplan(const cg_graph *g, int N, int Fi, int Fo, int K) {
    bool blocked = g->adj.nblk > 0 && (int64_t)g->adj.blk_total * 10 <= g->nnz * 8;
    if (const char *env = getenv("CG_FUSED_BLOCK")) blocked = atoi(env) != 0;
    if (blocked) {
        CPlan pb = make_cplan(g, N, Fi, Fo, K, true);
        if (pb.ok) return pb;
    }
    return make_cplan(g, N, Fi, Fo, K, false);
}

template <int LPR>
static cudaError_t launch_c(const CPlan &pl, dim3 grid, cudaStream_t s) {
#define CG_CL_CASE(I)                                                                                              \
    case I: {                                                                                                      \
        cudaError_t e = cudaFuncSetAttribute(k_cheb_clenshaw<LPR, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)pl.smem);                                                        \
        if (e != cudaSuccess) return e;                                                                            \
        k_cheb_clenshaw<LPR, I><<<grid, CT, pl.smem, s>>>(pl.cp);                                                  \
        return cudaGetLastError();                                                                                 \
    }
    if (pl.blocked) {
#define CG_CL_BCASE(I)                                                                                               \
    case I: {                                                                                                        \
        cudaError_t e = cudaFuncSetAttribute(k_cheb_clenshaw_b<LPR, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)pl.smem);                                                          \
        if (e != cudaSuccess) return e;                                                                              \
        k_cheb_clenshaw_b<LPR, I><<<grid, CT, pl.smem, s>>>(pl.cp);                                                  \
        return cudaGetLastError();                                                                                   \
    }
        switch (pl.ipt) {
            CG_CL_BCASE(1)
            CG_CL_BCASE(2)
        }
#undef CG_CL_BCASE
        return cudaErrorInvalidValue;
    }
    switch (pl.ipt) {
        CG_CL_CASE(2)
        CG_CL_CASE(4)
        CG_CL_CASE(8)
    }
#undef CG_CL_CASE
    return cudaErrorInvalidValue;
}

}  // namespace

extern int g_fused_last_plan[8];
static long long *g_clenshaw_trace = nullptr;
// debug aid: clock64 stamps of CTA 0's second group in k_cheb_clenshaw_b, [K][10] int64 on the device
extern "C" int cg_debug_clenshaw_trace(long long *dev_buf) {
    g_clenshaw_trace = dev_buf;
    return CG_OK;
}

bool cg_clenshaw_supported(const cg_graph *g, int N, int Fin, int Fout, int K) {
    return choose_cplan(g, N, Fin, Fout, K).ok;
}

// dx[n,m,fin] = sum_{k,fo} (T_k(L~^T) gy)[n,m,fo] W[fin*K+k, fo];  workspace: cg_fused_workspace(Fin, Fout, K) bytes
int cg_run_clenshaw(const cg_graph *g, const float *gy, const float *W, float *dx, int N, int Fin, int Fout, int K,
                    void *workspace, cudaStream_t s) {
    CPlan pl = choose_cplan(g, N, Fin, Fout, K);
    CG_REQUIRE(pl.ok, "cg_run_clenshaw: shape not supported (M=%d Fin=%d Fout=%d)", g->M, Fin, Fout);
    CG_REQUIRE(workspace != nullptr, "cg_run_clenshaw: workspace is NULL");
    unsigned char *wp = reinterpret_cast<unsigned char *>(workspace);
    int rc = cg_pack_w(W, wp, Fout, Fin, K, true, s);       // B(n = fin, q = fout) = W[(n*K + k)*Fout + q]
    if (rc != CG_OK) return rc;
    ClenshawParams &cp = pl.cp;
    cp.rowptr = g->adj.rowptr;
    cp.col = g->adj.col;
    cp.val = g->adj.val;
    cp.order = g->adj.order;
    cp.bt.ptr = g->adj.blk_ptr;
    cp.bt.col = g->adj.blk_col;
    cp.bt.w = g->adj.blk_w;
    cp.bt.order = g->adj.blk_order;
    cp.bt.nblk = g->adj.nblk;
    cp.gy = gy;
    cp.wp = wp;
    cp.dx = dx;
    cp.N = N;
    cp.M = g->M;
    cp.Fi = Fin;
    cp.Fo = Fout;
    cp.K = K;
    cp.trace = g_clenshaw_trace;
    const int64_t G = cg_ceil_div(N, pl.S);
    dim3 grid((unsigned)std::min<int64_t>(G, g->sm_count));
    g_fused_last_plan[4] = pl.blocked ? 1 : 0;
    g_fused_last_plan[5] = pl.S;
    g_fused_last_plan[6] = pl.ipt;
    g_fused_last_plan[7] = (int)pl.smem;
    CgProfScope prof("clenshaw_dx", s);
    cudaError_t e;
    switch (Fin / 4) {
        case 4: e = launch_c<4>(pl, grid, s); break;
        case 8: e = launch_c<8>(pl, grid, s); break;
        case 16: e = launch_c<16>(pl, grid, s); break;
        default: e = launch_c<32>(pl, grid, s); break;
    }
    if (e != cudaSuccess) {
        cg_set_error("cg_run_clenshaw: launch failed: %s", cudaGetErrorString(e));
        return CG_ERR_CUDA;
    }
    return CG_OK;
}
