// Fused Chebyshev filter (lib/models.py:192-224, lib/filter.py:45-95) for operators that fit
// in shared memory: recurrence AND weight contraction in one persistent kernel; the
// Chebyshev stack never leaves the SM.
//
//   per CTA, per group of S samples (rows r = s*M + m, R = S*M):
//     X_0 = x                           cp.async.bulk of the contiguous [R][Fin] block (TMA engine)
//     X_k = 2 L~ X_{k-1} - X_{k-2}      fp32, slabs in SMEM, CSR of L~ in SMEM, 128-bit gathers
//     y  += X_k W_k                     tcgen05.mma, accumulators [R][Fout] in TMEM
//   The fp32 slab is the master copy; every step also writes X_k as two bf16 planes
//   (hi = rn(x), mid = rn(x - hi)) in the canonical K-major UMMA layout, and the product is
//   formed as  hi*Whi + mid*Whi + hi*Wmid  (three bf16 MMAs, fp32 accumulate; the dropped
//   terms are <= 2^-16 relative), which keeps the filter inside the reference's fp32
//   tolerance (rtol 1e-4) while running on the tensor cores.
//   Warp roles: 16 compute warps (gathers, staging, epilogue) + 1 issue warp (bulk copies of
//   x / W_k, tcgen05.mma issue, commits).  The MMAs of step k run under the gathers of k+1.
//
// The same kernel computes dx = sum_k T_k(L~^T) gy W_k^T  (operator side = transpose,
// W packed transposed, Fin <-> Fout).
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

constexpr int FC = 512;        // compute threads
constexpr int NSTORE = 3;                  // plane store warps (one 128-row chunk of the image each)
constexpr int FT = FC + 32 + 32 * NSTORE;  // + MMA issue warp + plane store warps = 20 warps: still 96 registers per thread

struct FusedParams {
    int npass;                   // MMA passes per product: 3 (fp32-equivalent hi/mid split) or 1 (single-pass bf16)
    const int *rowptr;
    const int *col;
    const float *val;
    const int *order;            // rows by descending length
    const float *x;              // [N][M][Fin]
    const unsigned char *wp;     // packed W: [K][hi|mid][Fin*Fout] bf16, canonical K-major B operand
    float *y;                    // [N][M][Fout]
    float *stack_out;            // optional: X_k for the weight gradient, [K][N][M][Fin] (sample-major)
    unsigned char *planes_out;   // optional, instead of stack_out: the staged bf16 hi | mid planes themselves,
                                 // [2][K][chunks of 128 rows][Fin/8][128 rows][8 features] over the rows n*M + m
                                 // (cg_dw_planes.cu reads them without conversion)
    long long planes_k, planes_pl;      // bytes per k (all chunks) / between the hi and mid halves
    long long *trace;            // optional (debug): clock64 stamps of CTA 0, second group: [K][10]
    int N, M, Fin, Fout, K, S, nnz, tiles, tmem_cols, nslab, nw, estride;
    uint32_t off_ent, off_slab, slab_bytes, off_stage, plane_bytes, lbo_a, off_w, wplane_bytes, off_bar;
    BlkTables bt;                // row-block form of the operator (k_cheb_fused_b); its tables live at off_ent
    uint32_t off_wsz;
};

// ---- the two service warps of both fused kernels --------------------------------------------------------------------
// A single warp runs its (uniform-datapath) issue code at one dependent instruction per ~5 cycles next to four busy
// compute warps on its scheduler: the ~25 bulk stores of a step's planes alone took 1200 - 2200 cycles in the MMA issue
// thread and delayed the MMAs behind them (ncu / clock64 trace, round 2).  So: warp FC/32 issues the loads of x and
// W_k and the MMAs of a step, warp FC/32 + 1 ships the staged planes; the staging buffer is released when both are
// done (sbar counts two arrivals when the plane side output is on).
struct FusedCtl {
    uint64_t *xbar, *wbar, *mbar, *sbar;
    uint32_t slab0, w0, stage0, tmem, wbytes;
};

// MMA issue warp, one group: x / W loads, K steps of MMAs.  NK16 = Fin / 16 MMAs along the reduction per pass.
template <int NK16>
__device__ __forceinline__ void fused_issue_group(const FusedParams &p, const FusedCtl &c, int g, int gi, int G, int sP, int sR,
                                                  uint32_t &wpar, uint32_t &mpar, int lane) {
    const int M = p.M, Fin = p.Fin, Fout = p.Fout, K = p.K, S = p.S;
    const int n0 = g * S;
    const int Sg = min(S, p.N - n0);
    const bool prefetch = p.nslab == 3;
    if (lane == 0) {
        // weights of steps 0 and 1 (all MMAs of the previous group have completed)
        mbar_expect_tx(c.wbar, c.wbytes);
        bulk_g2s(c.w0, p.wp, c.wbytes, c.wbar);
        if (K > 1 && p.nw > 1) {
            mbar_expect_tx(c.wbar + 1, c.wbytes);
            bulk_g2s(c.w0 + c.wbytes, p.wp + c.wbytes, c.wbytes, c.wbar + 1);
        }
        const int gn = g + gridDim.x;
        if (prefetch) {
            if (gn < G) {       // next group's x into the spare slab
                const int Sn = min(S, p.N - gn * S);
                const uint32_t bytes = (uint32_t)Sn * M * Fin * 4u;
                umma::fence_proxy_async();
                mbar_expect_tx(c.xbar + ((gi + 1) & 1), bytes);
                bulk_g2s(c.slab0 + (uint32_t)sR * p.slab_bytes, p.x + (size_t)gn * S * M * Fin, bytes, c.xbar + ((gi + 1) & 1));
            }
        } else if (gi > 0) {    // no spare slab: load this group's x now
            const uint32_t bytes = (uint32_t)Sg * M * Fin * 4u;
            umma::fence_proxy_async();
            mbar_expect_tx(c.xbar + (gi & 1), bytes);
            bulk_g2s(c.slab0 + (uint32_t)sP * p.slab_bytes, p.x + (size_t)n0 * M * Fin, bytes, c.xbar + (gi & 1));
        }
    }
    const uint32_t idesc = umma::make_idesc_bf16(128, Fout, 0, 0);
    const uint32_t lbo_w = (uint32_t)Fout * 16u;
    const uint32_t d_hi = umma::desc_hi(128u);      // SBO = 128 for both operands
    // descriptors advance by plain adds on the low word (units of 16 bytes)
    const uint32_t a_lo = umma::desc_lo(c.stage0, p.lbo_a);
    const uint32_t a_mid = p.plane_bytes >> 4, b_mid = p.wplane_bytes >> 4;
    const uint32_t a_k = (2u * p.lbo_a) >> 4, b_k = (2u * lbo_w) >> 4;
    for (int k = 0; k < K; ++k) {
        __syncthreads();                                  // staging of step k is complete
        if (lane == 0 && umma::elect_lane0()) {
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && gi == 1;
            CG_STAMP(tr, k * 10 + 4);
            const int b = k & (p.nw - 1);
            umma::mbar_wait(c.wbar + b, (wpar >> b) & 1u);
            wpar ^= 1u << b;
            umma::fence_after_sync();
            CG_STAMP(tr, k * 10 + 5);
            const uint32_t wb = c.w0 + (uint32_t)b * c.wbytes;
            const uint32_t b_lo = umma::desc_lo(wb, lbo_w);
            uint32_t at = a_lo, acc = c.tmem;
            for (int t = 0; t < p.tiles; ++t, at += 128u, acc += (uint32_t)Fout) {      // 2048 bytes per 128-row tile
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    if (pass >= p.npass) break;
#pragma unroll
                    for (int j = 0; j < NK16; ++j) {
                        const uint32_t al = at + (pass == 1 ? a_mid : 0u) + (uint32_t)j * a_k;
                        const uint32_t bl = b_lo + (pass == 2 ? b_mid : 0u) + (uint32_t)j * b_k;
                        umma::mma_bf16(acc, umma::desc_join(al, d_hi), umma::desc_join(bl, d_hi), idesc, (k | pass | j) != 0);
                    }
                }
            }
            umma::commit(c.mbar);
            CG_STAMP(tr, k * 10 + 6);
            // W_{k+nw} goes where W_k was, once the MMAs of step k have read it; waiting for every
            // step also guarantees that nothing is in flight when the group ends
            umma::mbar_wait(c.mbar, mpar);
            mbar_arrive(c.sbar);                    // compute warps may overwrite the planes (once the store warp agrees)
            CG_STAMP(tr, k * 10 + 7);
            if (k + p.nw < K) {
                mbar_expect_tx(c.wbar + b, c.wbytes);
                bulk_g2s(wb, p.wp + (size_t)(k + p.nw) * c.wbytes, c.wbytes, c.wbar + b);
            }
        }
        mpar ^= 1;
        __syncwarp();
    }
}

// plane store warp, one group: after every step's staging, ship the planes as they are.  The staged planes are, per
// feature octet, one contiguous run over the group's rows in the MN-major core-matrix order the weight-gradient kernel
// wants (global image: chunks of 128 rows, so that a chunk of one k is one contiguous block).
__device__ __forceinline__ void fused_store_group(const FusedParams &p, const FusedCtl &c, int g, int gi, int lane, int sw) {
    const int M = p.M, Fin = p.Fin, K = p.K, S = p.S;
    const int n0 = g * S;
    const int Rg = min(S, p.N - n0) * M;
    const long long g0 = (long long)n0 * M, g1 = g0 + Rg;
    const uint32_t cs = (uint32_t)(Fin / 8) * 2048u;
    for (int k = 0; k < K; ++k) {
        __syncthreads();                                  // staging of step k is complete
        if (p.planes_out != nullptr && lane == 0 && umma::elect_lane0()) {
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && gi == 1 && sw == 0;
            CG_STAMP(tr, k * 10 + 8);
            unsigned char *kbase = p.planes_out + (size_t)k * p.planes_k;
            for (long long ch = (g0 >> 7) + sw; ch <= (g1 - 1) >> 7; ch += NSTORE) {
                const long long a = max(g0, ch << 7), b = min(g1, (ch + 1) << 7);
                const uint32_t run = (uint32_t)(b - a) * 16u, so = (uint32_t)(a - g0) * 16u;
                unsigned char *dst = kbase + (size_t)ch * cs + (size_t)(a & 127) * 16;
                uint32_t src = c.stage0 + so;
                for (int fo = 0; fo < Fin / 8; ++fo) {
                    bulk_s2g(dst, src, run);
                    bulk_s2g(dst + p.planes_pl, src + p.plane_bytes, run);
                    dst += 2048;
                    src += p.lbo_a;
                }
            }
            bulk_commit();
            bulk_wait_read();
            mbar_arrive(c.sbar);
            CG_STAMP(tr, k * 10 + 9);
        }
        __syncwarp();
    }
}

// LPR lanes per row (Fin = 4 * LPR), IPT items (row, 4-column chunk) per compute thread
template <int LPR, int IPT>
__global__ void __launch_bounds__(FT, 1) k_cheb_fused(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float2 *ent = reinterpret_cast<float2 *>(smem + p.off_ent);
    unsigned char *stage = smem + p.off_stage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *xbar = bars;            // [2] x slab of a group landed
    uint64_t *wbar = bars + 2;        // [2] W_k landed
    uint64_t *mbar = bars + 4;        // MMAs of the last issued step completed
    uint64_t *sbar = bars + 5;        // staging planes free again (MMAs done and, with planes_out, bulk stores read)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = p.M, Fin = p.Fin, Fout = p.Fout, K = p.K, S = p.S;
    const int R = S * M;
    const bool is_issuer = warp >= FC / 32;      // MMA issue warp or plane store warp

    // ---- one-time setup ------------------------------------------------------------
    // operator as fixed-stride rows {weight, byte offset of the neighbour inside its sample's slab}; the tail of
    // a row is {0, 0}, so the gather loops need neither row-length predicates nor tail handling
    for (int i = tid; i < M * p.estride; i += FT) {
        const int m = i / p.estride, j = i - m * p.estride;
        const int b = p.rowptr[m], n = p.rowptr[m + 1] - b;
        float2 v = make_float2(0.f, __int_as_float(0));
        if (j < n) {
            v.x = p.val[b + j];
            v.y = __int_as_float(p.col[b + j] * Fin * 4);
        }
        ent[i] = v;
    }
    {   // pad rows of the A operand are never written by the steps: clear the staging planes once
        uint4 *z = reinterpret_cast<uint4 *>(stage);
        const int n16 = (int)(2 * p.plane_bytes / 16);
        for (int i = tid; i < n16; i += FT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) umma::mbar_init(bars + i, (i == 5 && p.planes_out != nullptr) ? 1 + NSTORE : 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);

    // per-item constants (identical for every step and group).  Rows are dealt to the items in order of
    // descending length, interleaving the samples of a group: the four rows of a warp-level item and the two
    // items of a pair have (nearly) equal lengths, and every warp gets the same mix of long and short rows.
    uint32_t a_ent[IPT], a_g[IPT], a_soff[IPT], a_stoff[IPT];
    int nlen[IPT];
    int wl[IPT / 2];
    const uint32_t ent0 = umma::smem_u32(ent);
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        // the two items of a pair are adjacent in the dealing order
        // (odd pair slots are dealt in reverse thread order: every warp gets the same mix of long and short rows)
        const int qd = ((i >> 1) & 1) ? (FC / LPR - 1 - tid / LPR) : tid / LPR;
        const int o = ((i >> 1) * (FC / LPR) + qd) * 2 + (i & 1), l = tid % LPR;
        nlen[i] = 0;
        a_ent[i] = ent0;
        a_g[i] = 0;
        a_soff[i] = 0xFFFFFFF0u;     // "absent": compares above every group's limit
        a_stoff[i] = 0;
        if (!is_issuer && o < R) {
            const int s = o % S, m = p.order[o / S];
            const int r = s * M + m;
            a_ent[i] = ent0 + 8u * (uint32_t)(m * p.estride);
            nlen[i] = p.rowptr[m + 1] - p.rowptr[m];
            a_g[i] = 4u * (uint32_t)(s * M * Fin + 4 * l);
            a_soff[i] = 4u * (uint32_t)(r * Fin + 4 * l);
            // the even lane of a pair stores the 16-byte hi octet (its 4 features and its partner's), the odd lane the
            // mid octet: with LBO = 32 and plane = 16 (mod 128) the 8 lanes of a row fill exactly one 128-byte wavefront
            a_stoff[i] = (uint32_t)((l >> 1) * (int)p.lbo_a + r * 16) + ((l & 1) ? p.plane_bytes : 0u);
        }
    }
#pragma unroll
    for (int pr = 0; pr < IPT / 2; ++pr)
        wl[pr] = __reduce_max_sync(0xffffffffu, max(nlen[2 * pr], nlen[2 * pr + 1]));
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int G = (p.N + S - 1) / S;
    const bool prefetch = p.nslab == 3;
    const uint32_t slab0 = umma::smem_u32(smem + p.off_slab);
    const uint32_t wbytes = 2 * p.wplane_bytes;
    const uint32_t w0 = umma::smem_u32(smem + p.off_w);
    const FusedCtl ctl = {xbar, wbar, mbar, sbar, slab0, w0, umma::smem_u32(stage), tmem, wbytes};
    int sP = 0, sQ = 1, sR = 2;           // slab roles: X_0 of this group, second slab, prefetch target
    uint32_t mpar = 0;                    // parity of the next MMA-complete phase to wait for
    uint32_t wpar = 0;                    // issuer: parities of the two W barriers (bit b)
    int gi = 0;                           // groups processed by this CTA

    if (warp == FC / 32 && lane == 0 && (int)blockIdx.x < G) {
        const int Sg = min(S, p.N - (int)blockIdx.x * S);
        const uint32_t bytes = (uint32_t)Sg * M * Fin * 4u;
        mbar_expect_tx(xbar, bytes);
        bulk_g2s(slab0, p.x + (size_t)blockIdx.x * S * M * Fin, bytes, xbar);
    }

    for (int g = blockIdx.x; g < G; g += gridDim.x, ++gi) {
        const int n0 = g * S;
        const int Sg = min(S, p.N - n0);
        const int Rg = Sg * M;
        const int lim = Rg * Fin;         // items with i_soff >= lim belong to absent samples

        if (warp == FC / 32) {
            fused_issue_group<LPR / 4>(p, ctl, g, gi, G, sP, sR, wpar, mpar, lane);
        } else if (warp > FC / 32) {
            fused_store_group(p, ctl, g, gi, lane, warp - FC / 32 - 1);
        } else {
            // =========================== compute warps ==================================
            umma::mbar_wait(xbar + (gi & 1), (uint32_t)((gi >> 1) & 1));
            const uint32_t limb = 4u * (uint32_t)lim;
            const uint32_t aP = slab0 + (uint32_t)sP * p.slab_bytes, aQ = slab0 + (uint32_t)sQ * p.slab_bytes;
            const uint32_t st0 = umma::smem_u32(stage);
            float4 res[IPT], old[IPT];     // the thread's own X_{k-1} (X_k after the step) and X_{k-2}
            // ---- step 0: X_0 = x
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                if (a_soff[i] < limb) res[i] = lds128(aP + a_soff[i]);
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && gi == 1 && tid == 0;
            for (int k = 0; k < K; ++k) {
                CG_STAMP(tr, k * 10 + 0);
                if (k > 0) {
                    const uint32_t prev = (k & 1) ? aP : aQ;     // X_{k-1}
                    const uint32_t cur = (k & 1) ? aQ : aP;      // X_{k-2} -> X_k
#pragma unroll
                    for (int pr = 0; pr < IPT / 2; ++pr) {
                        const int i0 = 2 * pr, i1 = 2 * pr + 1;
                        const uint32_t g0 = prev + a_g[i0], g1 = prev + a_g[i1];
                        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
                        const int trips = wl[pr];
                        const int jlast = p.estride - 2;
                        float4 e0 = lds128(a_ent[i0]), e1 = lds128(a_ent[i1]);     // two {weight, offset} entries each
                        for (int j = 0; j < trips; j += 2) {
                            const float4 x00 = lds128(g0 + (uint32_t)__float_as_int(e0.y));
                            const float4 x10 = lds128(g1 + (uint32_t)__float_as_int(e1.y));
                            const float4 x01 = lds128(g0 + (uint32_t)__float_as_int(e0.w));
                            const float4 x11 = lds128(g1 + (uint32_t)__float_as_int(e1.w));
                            const float w00 = e0.x, w01 = e0.z, w10 = e1.x, w11 = e1.z;
                            const uint32_t jn = 8u * (uint32_t)min(j + 2, jlast);     // next entries (clamped in-row)
                            e0 = lds128(a_ent[i0] + jn);
                            e1 = lds128(a_ent[i1] + jn);
                            fma4(acc0, w00, x00);
                            fma4(acc1, w10, x10);
                            fma4(acc0, w01, x01);
                            fma4(acc1, w11, x11);
                        }
                        // X_k = 2 L X_{k-1} - X_{k-2}; the thread's own X_{k-2} is still in registers
                        if (k > 1) {
                            const float4 o0 = old[i0], o1 = old[i1];
                            acc0 = make_float4(fmaf(2.f, acc0.x, -o0.x), fmaf(2.f, acc0.y, -o0.y), fmaf(2.f, acc0.z, -o0.z),
                                               fmaf(2.f, acc0.w, -o0.w));
                            acc1 = make_float4(fmaf(2.f, acc1.x, -o1.x), fmaf(2.f, acc1.y, -o1.y), fmaf(2.f, acc1.z, -o1.z),
                                               fmaf(2.f, acc1.w, -o1.w));
                        }
                        old[i0] = res[i0];
                        old[i1] = res[i1];
                        res[i0] = acc0;
                        res[i1] = acc1;
                        if (a_soff[i0] < limb) sts128(cur + a_soff[i0], acc0);
                        if (a_soff[i1] < limb) sts128(cur + a_soff[i1], acc1);
                    }
                    // the staging planes are free once the MMAs (and bulk stores) of step k-1 have read them
                    CG_STAMP(tr, k * 10 + 1);
                    umma::mbar_wait(sbar, mpar);
                    mpar ^= 1;
                }
                CG_STAMP(tr, k * 10 + 2);
                if (p.stack_out != nullptr) {       // side output for the backward pass (coalesced 128-bit stores)
                    char *dst = reinterpret_cast<char *>(p.stack_out + ((size_t)k * p.N + n0) * M * Fin);
#pragma unroll
                    for (int i = 0; i < IPT; ++i)
                        if (a_soff[i] < limb) *reinterpret_cast<float4 *>(dst + a_soff[i]) = res[i];
                }
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    uint2 hi, mid;
                    split4(res[i], hi, mid);
                    // lanes l and l ^ 1 hold the two halves of one feature octet of the same row (absent rows too:
                    // the exchange runs on every lane, only the store is predicated)
                    const bool odd = tid & 1;
                    const uint2 send = odd ? hi : mid;
                    uint2 recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                    const uint4 v = odd ? make_uint4(recv.x, recv.y, mid.x, mid.y) : make_uint4(hi.x, hi.y, recv.x, recv.y);
                    if (a_soff[i] < limb)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st0 + a_stoff[i]), "r"(v.x), "r"(v.y),
                                     "r"(v.z), "r"(v.w)
                                     : "memory");
                }
                umma::fence_proxy_async();
                if (k == 0) umma::fence_before_sync();     // orders the previous group's TMEM loads
                CG_STAMP(tr, k * 10 + 3);
                __syncthreads();
            }
            // ---- epilogue: TMEM -> registers -> y
            umma::mbar_wait(sbar, mpar);
            mpar ^= 1;
            umma::fence_after_sync();
            const int q = warp & 3, wq = warp >> 2;
            const int nc8 = Fout / 8;
            for (int idx = wq; idx < p.tiles * nc8; idx += 4) {
                const int t = idx / nc8, c = idx - t * nc8;
                const int r = t * 128 + 32 * q + lane;
                float v[8];
                umma::tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * Fout + c * 8), v);
                umma::tmem_ld_wait();
                if (r < Rg) {
                    float *dst = p.y + ((size_t)n0 * M + r) * Fout + c * 8;
                    *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
        // rotate the slabs: the prefetched block becomes X_0
        if (prefetch) {
            const int t = sP;
            sP = sR;
            sR = sQ;
            sQ = t;
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// The same kernel with the row-block gather of cg_fused_common.cuh: a compute thread owns IPB items of
// (4 consecutive rows) x (4 features); everything else (slabs, staging planes, issue warp, epilogue) is unchanged.
template <int LPR, int IPB>
__global__ void __launch_bounds__(FT, 1) k_cheb_fused_b(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *stage = smem + p.off_stage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *xbar = bars;            // [2] x slab of a group landed
    uint64_t *wbar = bars + 2;        // [2] W_k landed
    uint64_t *mbar = bars + 4;        // MMAs of the last issued step completed
    uint64_t *sbar = bars + 5;        // staging planes free again
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int Fin = 4 * LPR;                 // compile-time row width: no multiplications in the address arithmetic
    const int M = p.M, Fout = p.Fout, K = p.K, S = p.S;
    const bool is_issuer = warp >= FC / 32;      // MMA issue warp or plane store warp
    constexpr uint32_t rowb = (uint32_t)Fin * 4u;

    // ---- one-time setup ------------------------------------------------------------
    {   // pad rows of the A operand are never written by the steps: clear the staging planes once
        uint4 *z = reinterpret_cast<uint4 *>(stage);
        const int n16 = (int)(2 * p.plane_bytes / 16);
        for (int i = tid; i < n16; i += FT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) umma::mbar_init(bars + i, (i == 5 && p.planes_out != nullptr) ? 1 + NSTORE : 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);

    BlkItems<IPB> it;
    blk_setup<LPR, IPB, FC>(p.bt, S, rowb, umma::smem_u32(smem + p.off_ent), reinterpret_cast<int *>(smem + p.off_wsz), it);
    uint32_t a_g[IPB], a_soff[IPB], a_stoff[IPB];
    int nrow[IPB];
    {
        const int l = tid % LPR;
#pragma unroll
        for (int i = 0; i < IPB; ++i) {
            const int r0 = it.samp[i] * M + it.row0[i];
            a_g[i] = 4u * (uint32_t)(it.samp[i] * M * Fin + 4 * l);
            a_soff[i] = 4u * (uint32_t)(r0 * Fin + 4 * l);
            // even lanes store the 16-byte hi octet (own 4 features + partner's), odd lanes the mid octet (see k_cheb_fused)
            a_stoff[i] = (uint32_t)((l >> 1) * (int)p.lbo_a + r0 * 16) + ((l & 1) ? p.plane_bytes : 0u);
            nrow[i] = it.samp[i] < S ? min(4, M - it.row0[i]) : 0;
        }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int G = (p.N + S - 1) / S;
    const bool prefetch = p.nslab == 3;
    const uint32_t slab0 = umma::smem_u32(smem + p.off_slab);
    const uint32_t wbytes = 2 * p.wplane_bytes;
    const uint32_t w0 = umma::smem_u32(smem + p.off_w);
    const FusedCtl ctl = {xbar, wbar, mbar, sbar, slab0, w0, umma::smem_u32(stage), tmem, wbytes};
    int sP = 0, sQ = 1, sR = 2;
    uint32_t mpar = 0;
    uint32_t wpar = 0;
    int gi = 0;

    if (warp == FC / 32 && lane == 0 && (int)blockIdx.x < G) {
        const int Sg = min(S, p.N - (int)blockIdx.x * S);
        const uint32_t bytes = (uint32_t)Sg * M * Fin * 4u;
        mbar_expect_tx(xbar, bytes);
        bulk_g2s(slab0, p.x + (size_t)blockIdx.x * S * M * Fin, bytes, xbar);
    }

    for (int g = blockIdx.x; g < G; g += gridDim.x, ++gi) {
        const int n0 = g * S;
        const int Sg = min(S, p.N - n0);
        const int Rg = Sg * M;

        if (warp == FC / 32) {
            fused_issue_group<LPR / 4>(p, ctl, g, gi, G, sP, sR, wpar, mpar, lane);
        } else if (warp > FC / 32) {
            fused_store_group(p, ctl, g, gi, lane, warp - FC / 32 - 1);
        } else {
            // =========================== compute warps ==================================
            umma::mbar_wait(xbar + (gi & 1), (uint32_t)((gi >> 1) & 1));
            const uint32_t aP = slab0 + (uint32_t)sP * p.slab_bytes, aQ = slab0 + (uint32_t)sQ * p.slab_bytes;
            const uint32_t st0 = umma::smem_u32(stage);
            int nr[IPB];                      // rows of the item that exist in this group
            // the thread's own X_{k-1} (X_k after the step) and X_{k-2}; with two items per thread X_{k-2} is read back
            // from the slab position that X_k is about to overwrite instead (96-register cap at 544 threads)
            constexpr bool OLD_REGS = IPB == 1;
            float4 res[IPB][4], old[OLD_REGS ? IPB : 1][4];
#pragma unroll
            for (int i = 0; i < IPB; ++i) {
                nr[i] = it.samp[i] < Sg ? nrow[i] : 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    res[i][r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nr[i]) res[i][r] = lds128(aP + a_soff[i] + (uint32_t)r * rowb);
                }
            }
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && gi == 1 && tid == 0;
            for (int k = 0; k < K; ++k) {
                CG_STAMP(tr, k * 10 + 0);
                if (k > 0) {
                    const uint32_t prev = (k & 1) ? aP : aQ;     // X_{k-1}
                    const uint32_t cur = (k & 1) ? aQ : aP;      // X_{k-2} -> X_k
#pragma unroll
                    for (int i = 0; i < IPB; ++i) {
                        float4 acc[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                        blk_gather<LPR>(prev + a_g[i], it.tab[i], it.trips[i], acc);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (k > 1) {
                                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                                if constexpr (OLD_REGS) {
                                    o = old[i][r];
                                } else {
                                    if (r < nr[i]) o = lds128(cur + a_soff[i] + (uint32_t)r * rowb);
                                }
                                acc[r] = make_float4(fmaf(2.f, acc[r].x, -o.x), fmaf(2.f, acc[r].y, -o.y),
                                                     fmaf(2.f, acc[r].z, -o.z), fmaf(2.f, acc[r].w, -o.w));
                            }
                            if constexpr (OLD_REGS) old[i][r] = res[i][r];
                            res[i][r] = acc[r];
                            if (r < nr[i]) sts128(cur + a_soff[i] + (uint32_t)r * rowb, acc[r]);
                        }
                    }
                    CG_STAMP(tr, k * 10 + 1);
                    umma::mbar_wait(sbar, mpar);
                    mpar ^= 1;
                }
                CG_STAMP(tr, k * 10 + 2);
                if (p.stack_out != nullptr) {       // side output for the backward pass (coalesced 128-bit stores)
                    char *dst = reinterpret_cast<char *>(p.stack_out + ((size_t)k * p.N + n0) * M * Fin);
#pragma unroll
                    for (int i = 0; i < IPB; ++i)
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (r < nr[i]) *reinterpret_cast<float4 *>(dst + a_soff[i] + (uint32_t)r * rowb) = res[i][r];
                }
#pragma unroll
                for (int i = 0; i < IPB; ++i)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        uint2 hi, mid;
                        split4(res[i][r], hi, mid);
                        const bool odd = tid & 1;
                        const uint2 send = odd ? hi : mid;
                        uint2 recv;
                        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                        const uint4 v = odd ? make_uint4(recv.x, recv.y, mid.x, mid.y) : make_uint4(hi.x, hi.y, recv.x, recv.y);
                        if (r < nr[i])
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st0 + a_stoff[i] + 16u * (uint32_t)r),
                                         "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                                         : "memory");
                    }
                umma::fence_proxy_async();
                if (k == 0) umma::fence_before_sync();     // orders the previous group's TMEM loads
                CG_STAMP(tr, k * 10 + 3);
                __syncthreads();
            }
            // ---- epilogue: TMEM -> registers -> y
            umma::mbar_wait(sbar, mpar);
            mpar ^= 1;
            umma::fence_after_sync();
            const int q = warp & 3, wq = warp >> 2;
            const int nc8 = Fout / 8;
            for (int idx = wq; idx < p.tiles * nc8; idx += 4) {
                const int t = idx / nc8, c = idx - t * nc8;
                const int r = t * 128 + 32 * q + lane;
                float v[8];
                umma::tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * Fout + c * 8), v);
                umma::tmem_ld_wait();
                if (r < Rg) {
                    float *dst = p.y + ((size_t)n0 * M + r) * Fout + c * 8;
                    *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
        if (prefetch) {
            const int t = sP;
            sP = sR;
            sR = sQ;
            sQ = t;
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// W [Fin*K, Fout] (row = f*K + k)            -> wp[k][hi|mid] with B(n = fout, q = f)
// transposed: the operand is W^T per k: B(n = fin_ext, q = fout_ext), i.e. for the dx filter
// (stack feature = fout of W, output = fin of W):  B(n, q) = W[(n*K + k) * Q + q]
__global__ void __launch_bounds__(256)
k_pack_w(const float *__restrict__ W, unsigned char *__restrict__ wp, int Q, int Nn, int K, int transposed) {
    const int total = K * Q * Nn;
    const uint32_t plane = (uint32_t)Q * Nn * 2u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / (Q * Nn);
        const int rem = i - k * Q * Nn;
        const int q = rem / Nn, n = rem - q * Nn;
        const float w = transposed ? W[((size_t)n * K + k) * Q + q] : W[((size_t)q * K + k) * Nn + n];
        __nv_bfloat16 hi, mid;
        umma::split_bf16(w, hi, mid);
        const uint32_t off = (uint32_t)(q >> 3) * (uint32_t)Nn * 16u + (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u +
                             (uint32_t)(q & 7) * 2u;
        unsigned char *base = wp + (size_t)k * 2u * plane;
        *reinterpret_cast<__nv_bfloat16 *>(base + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(base + plane + off) = mid;
    }
}

struct Plan {
    bool ok = false;
    int S = 0, tiles = 0, ipt = 0, nslab = 0, tmem_cols = 0;
    bool blocked = false;
    double cost = 0.0;
    FusedParams fp;
    size_t smem = 0;
};

// blocked: the row-block gather kernel (k_cheb_fused_b); otherwise one row per item (k_cheb_fused)
static Plan make_plan(const cg_graph *g, const CgCsr &side, int64_t nnz, int N, int Fin, int Fout, int K, bool blocked) {
    Plan best;
    if (Fin % 16 != 0 || Fin > 128 || (Fin & (Fin - 1)) != 0) return best;     // LPR in {4, 8, 16, 32}
    if (Fout % 16 != 0 || Fout < 16 || Fout > 256) return best;
    if (N <= 0 || K < 1) return best;
    if (blocked && side.nblk == 0) return best;
    const int M = g->M, LPR = Fin / 4, width = side.width;
    const double avg = M > 0 ? (double)(blocked ? 4 * (int64_t)side.blk_total : nnz) / M : 0.0;   // gathers per 4 rows x 1/4
    double best_cost = 0.0;
    int s_lo = 1, s_hi = 64;
    if (const char *env = getenv("CG_FUSED_S")) {       // tuning / debugging aid: pin the samples per group
        const int v = atoi(env);
        if (v > 0) s_lo = s_hi = v;
    }
    for (int S = s_lo; S <= N && S <= s_hi; ++S) {
        const int64_t R = (int64_t)S * M;
        const int tiles = (int)cg_ceil_div(R, 128);
        if ((int64_t)tiles * Fout > 512) break;
        int need, ipt;
        if (blocked) {
            need = (int)cg_ceil_div((int64_t)S * side.nblk * LPR, FC);
            if (need > 2) break;        // res[] / old[] of 4 rows per item must stay in registers
            ipt = need;
        } else {
            need = (int)cg_ceil_div(R * LPR, FC);
            if (need > 8) break;        // res[] + item constants must stay in registers
            ipt = need <= 2 ? 2 : need <= 4 ? 4 : 8;
        }
        const uint32_t Rp = (uint32_t)tiles * 128u;
        // strides chosen for conflict-free staging stores (see a_stoff): LBO = 32, plane = 16 (mod 128)
        const uint32_t lbo_a = Rp * 16u + 32u;
        const uint32_t plane = (uint32_t)cg_align_up((size_t)(Fin / 8) * lbo_a, 128) + 16u;
        const uint32_t wplane = (uint32_t)Fin * Fout * 2u;
        const uint32_t slab = (uint32_t)cg_align_up((size_t)R * Fin * 4, 128);
        const int estride = std::max(2, (width + 1) & ~1);      // even: 16-byte aligned entry pairs
        const size_t ent_bytes = blocked ? cg_blk_table_bytes(side.blk_len_sorted, S, LPR, ipt, FC) : (size_t)M * estride * 8;
        for (int cfg = 0; cfg < 3; ++cfg) {
            const int nslab = cfg == 0 ? 3 : 2, nw = cfg == 2 ? 1 : 2;
            FusedParams fp;
            memset(&fp, 0, sizeof(fp));
    fp.npass = cg_mma_passes();
            uint32_t off = 0;
            fp.off_bar = off;
            off += 128;
            fp.off_wsz = off;
            off += 256;                  // blk_setup scratch: (FC / 32) * IPB ints
            fp.off_ent = off;
            off += (uint32_t)cg_align_up(ent_bytes, 128);
            fp.off_slab = off;
            off += (uint32_t)nslab * slab;
            fp.off_stage = off;
            off += (uint32_t)cg_align_up(2 * (size_t)plane, 128);
            fp.off_w = off;
            off += (uint32_t)nw * 2u * wplane;
            if (off > g->smem_optin) continue;
            const int64_t G = cg_ceil_div(N, S);
            const int64_t rounds = cg_ceil_div(G, g->sm_count);
            const double step = blocked ? (double)need * (avg * 5.5 + 120.0) + 300.0
                                        : (double)need * (avg * 7.0 + 40.0) * (need > 4 ? 1.6 : 1.0) + 300.0;   // > 4: register spills
            const double cost = (double)rounds * ((double)K * step + 600.0 + (double)R * Fout / 64.0) -
                                (nslab == 3 ? 1.0 : 0.0) - (nw == 2 ? 0.5 : 0.0);
            if (!best.ok || cost < best_cost) {
                best.ok = true;
                best_cost = cost;
                best.cost = cost;
                best.S = S;
                best.tiles = tiles;
                best.ipt = ipt;
                best.nslab = nslab;
                int cols = 32;
                while (cols < tiles * Fout) cols *= 2;
                best.tmem_cols = cols;
                fp.S = S;
                fp.tiles = tiles;
                fp.tmem_cols = cols;
                fp.nslab = nslab;
                fp.nw = nw;
                fp.estride = estride;
                fp.slab_bytes = slab;
                fp.plane_bytes = plane;
                fp.lbo_a = lbo_a;
                fp.wplane_bytes = wplane;
                best.fp = fp;
                best.smem = off;
                best.blocked = blocked;
            }
            break;      // 3 slabs fit: no need to look at 2
        }
    }
    return best;
}

// The row-block gather pays when neighbouring rows share neighbours (union well below the entry count); on operators
// without that locality it would only add multiplications by zero.  CG_FUSED_BLOCK=0/1 overrides (tests run both).
static bool want_blocked(const CgCsr &side, int64_t nnz) {
    if (const char *env = getenv("CG_FUSED_BLOCK")) return atoi(env) != 0;
    return side.nblk > 0 && (int64_t)side.blk_total * 10 <= nnz * 8;
}

static Plan choose_plan(const cg_graph *g, int transpose, int N, int Fin, int Fout, int K) {
    const CgCsr &side = cg_side(g, transpose);
    if (want_blocked(side, g->nnz)) {
        Plan pb = make_plan(g, side, g->nnz, N, Fin, Fout, K, true);
        if (pb.ok) return pb;
    }
    return make_plan(g, side, g->nnz, N, Fin, Fout, K, false);
}

template <int LPR>
static cudaError_t launch_lpr(const Plan &pl, dim3 grid, cudaStream_t s) {
#define CG_FUSED_CASE(I)                                                                                         \
    case I: {                                                                                                    \
        cudaError_t e = cudaFuncSetAttribute(k_cheb_fused<LPR, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)pl.smem);                                                      \
        if (e != cudaSuccess) return e;                                                                          \
        k_cheb_fused<LPR, I><<<grid, FT, pl.smem, s>>>(pl.fp);                                                   \
        return cudaGetLastError();                                                                               \
    }
    if (pl.blocked) {
#define CG_FUSED_BCASE(I)                                                                                          \
    case I: {                                                                                                      \
        cudaError_t e = cudaFuncSetAttribute(k_cheb_fused_b<LPR, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)pl.smem);                                                        \
        if (e != cudaSuccess) return e;                                                                            \
        k_cheb_fused_b<LPR, I><<<grid, FT, pl.smem, s>>>(pl.fp);                                                   \
        return cudaGetLastError();                                                                                 \
    }
        switch (pl.ipt) {
            CG_FUSED_BCASE(1)
            CG_FUSED_BCASE(2)
        }
#undef CG_FUSED_BCASE
        return cudaErrorInvalidValue;
    }
    switch (pl.ipt) {
        CG_FUSED_CASE(2)
        CG_FUSED_CASE(4)
        CG_FUSED_CASE(8)
    }
#undef CG_FUSED_CASE
    return cudaErrorInvalidValue;
}

}  // namespace

static long long *g_fused_trace = nullptr;
int g_fused_last_plan[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // [0..3] forward kernel, [4..7] Clenshaw kernel: blocked, S, items, smem
// debug aid: the plan of the most recent fused forward / Clenshaw launch of this process
extern "C" int cg_debug_fused_plan_info(int *info8) {
    if (!info8) return CG_ERR_ARG;
    for (int i = 0; i < 8; ++i) info8[i] = g_fused_last_plan[i];
    return CG_OK;
}
// debug aid (not in the public header): clock64 stamps of CTA 0's second group, [K][8] int64 on the device
extern "C" int cg_debug_fused_trace(long long *dev_buf) {
    g_fused_trace = dev_buf;
    return CG_OK;
}

int cg_pack_w(const float *W, unsigned char *wp, int Q, int Nn, int K, bool transposed, cudaStream_t s) {
    CgProfScope prof("pack_w", s);
    const int total = K * Q * Nn;
    k_pack_w<<<(unsigned)std::min<int64_t>(cg_ceil_div(total, 256), 1024), 256, 0, s>>>(W, wp, Q, Nn, K, transposed ? 1 : 0);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

bool cg_fused_supported(const cg_graph *g, int transpose, int N, int Fin, int Fout, int K) {
    return choose_plan(g, transpose, N, Fin, Fout, K).ok;
}

size_t cg_fused_workspace(int Fin, int Fout, int K) { return cg_align_up((size_t)K * Fin * Fout * 4, 256); }

int cg_run_fused(const cg_graph *g, int transpose, const float *x, const float *W, float *y, float *stack_out, int N,
                 int Fin, int Fout, int K, bool w_transposed, void *workspace, cudaStream_t s, bool stack_planes) {
    Plan pl = choose_plan(g, transpose, N, Fin, Fout, K);
    CG_REQUIRE(pl.ok, "cg_run_fused: shape not supported by the fused kernel (M=%d Fin=%d Fout=%d)", g->M, Fin, Fout);
    CG_REQUIRE(workspace != nullptr, "cg_run_fused: workspace is NULL");
    const CgCsr &L = cg_side(g, transpose);
    unsigned char *wp = reinterpret_cast<unsigned char *>(workspace);
    int rc = cg_pack_w(W, wp, Fin, Fout, K, w_transposed, s);
    if (rc != CG_OK) return rc;
    FusedParams &fp = pl.fp;
    fp.rowptr = L.rowptr;
    fp.col = L.col;
    fp.val = L.val;
    fp.order = L.order;
    fp.bt.ptr = L.blk_ptr;
    fp.bt.col = L.blk_col;
    fp.bt.w = L.blk_w;
    fp.bt.order = L.blk_order;
    fp.bt.nblk = L.nblk;
    fp.x = x;
    fp.wp = wp;
    fp.y = y;
    fp.stack_out = stack_planes ? nullptr : stack_out;
    fp.planes_out = stack_planes ? reinterpret_cast<unsigned char *>(stack_out) : nullptr;
    fp.planes_k = cg_ceil_div((long long)N * g->M, 128) * (Fin / 8) * 2048;     // all 128-row chunks of one k
    fp.planes_pl = (long long)K * fp.planes_k;                                  // hi half, then mid half
    CG_REQUIRE(!stack_planes || Fin % 8 == 0, "cg_run_fused: the plane side output needs Fin %% 8 == 0");
    fp.N = N;
    fp.M = g->M;
    fp.Fin = Fin;
    fp.Fout = Fout;
    fp.K = K;
    fp.nnz = (int)g->nnz;
    fp.trace = g_fused_trace;
    const int64_t G = cg_ceil_div(N, pl.S);
    dim3 grid((unsigned)std::min<int64_t>(G, g->sm_count));
    g_fused_last_plan[0] = pl.blocked ? 1 : 0;
    g_fused_last_plan[1] = pl.S;
    g_fused_last_plan[2] = pl.ipt;
    g_fused_last_plan[3] = (int)pl.smem;
    CgProfScope prof(transpose ? "fused_dx" : "fused_fwd", s);
    cudaError_t e;
    switch (Fin / 4) {
        case 4: e = launch_lpr<4>(pl, grid, s); break;
        case 8: e = launch_lpr<8>(pl, grid, s); break;
        case 16: e = launch_lpr<16>(pl, grid, s); break;
        default: e = launch_lpr<32>(pl, grid, s); break;
    }
    if (e != cudaSuccess) {
        cg_set_error("cg_run_fused: launch failed: %s", cudaGetErrorString(e));
        return CG_ERR_CUDA;
    }
    return CG_OK;
}
