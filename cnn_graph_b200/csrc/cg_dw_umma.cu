// Weight gradient of the Chebyshev filter on the tensor cores (TF autodiff of the matmul at
// lib/models.py:222-223, reached through lib/graph_model.py:296):
//
//     P[q][b] = sum_r stack_q[r] * T[trow(r)][b],      q = k*Fa + a,   r = m*N + n,   trow(r) = n*M + m
//
// A "TN" product with a very long reduction (r runs over all N*M vertex signals) and a small output
// (K*Fa x Fb): every CTA keeps its whole K*Fa x Fb accumulator in TMEM (q on the 128 lanes of up to
// 512/Fb row tiles), streams a contiguous range of rows through shared memory and writes one partial
// result; k_reduce_partials (cg_gemm.cu) sums the partials deterministically.
//
// Per chunk of KD rows:  cp.async.bulk (TMA engine) brings the fp32 pieces of the stack ([KD][Fa] per k)
// and of T into a 2-deep ring;  the compute warps split every value into bf16 hi + mid and store the
// MN-major canonical UMMA operands;  one thread issues  hi*hi + mid*hi + hi*mid  (fp32 accumulate).
#include <stdlib.h>

#include <algorithm>

#include "cg_common.cuh"
#include "cg_umma.cuh"

namespace {

constexpr int DC = 512;        // compute threads
constexpr int DT = DC + 64;    // + MMA issue warp + TMA producer warp

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_smem),
                 "l"(src), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void split8(const float *v, uint4 &hi, uint4 &mid) {
    uint32_t h[4], m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float2 f = __bfloat1622float2(hh);
        const __nv_bfloat162 mm = __floats2bfloat162_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
        h[i] = *reinterpret_cast<const uint32_t *>(&hh);
        m[i] = *reinterpret_cast<const uint32_t *>(&mm);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    mid = make_uint4(m[0], m[1], m[2], m[3]);
}

struct DwParams {
    int npass;                   // MMA passes per product: 3 (fp32-equivalent hi/mid split) or 1 (single-pass bf16)
    const float *stack;     // [K][R][Fa]
    const float *T;         // [R][Fb], row trow(r)
    float *part;            // [splits][K*Fa][Fb]
    long long *trace;       // optional (debug): clock64 stamps of CTA 0, chunks 8..23: [16][8]
    long long R, rows_per_cta;
    int N, M, Fa, Fb, K, KD, kgroup, tmem_cols, nstage, oa_shift, ob_shift, sample_major;
    uint32_t ring_bytes, ring_b_off, stage_bytes, stage_a_plane, stage_b_off, stage_b_plane, off_ring, off_stage, off_bar;
};

// FA8: Fa % 8 == 0 (two 128-bit reads per octet); otherwise scalar gathers across the k pieces
template <bool FA8>
__global__ void __launch_bounds__(DT, 1) k_dw_umma(const DwParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *full = bars;          // [2] ring slot landed
    uint64_t *mbar = bars + 2;      // [2] MMAs of chunk c completed (barrier c & 1)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_issuer = warp == DC / 32;
    const int Fa = p.Fa, Fb = p.Fb, KD = p.KD;
    const int k_lo = blockIdx.y * p.kgroup;
    const int nk = min(p.kgroup, p.K - k_lo);          // k-slabs of this CTA
    const int Ql = nk * Fa;                             // local q range
    const int tiles = (Ql + 127) / 128;
    const long long r_beg = (long long)blockIdx.x * p.rows_per_cta;
    const long long r_end = min(p.R, r_beg + p.rows_per_cta);
    const int nchunks = r_end > r_beg ? (int)((r_end - r_beg + KD - 1) / KD) : 0;

    const uint32_t ring0 = umma::smem_u32(smem + p.off_ring);
    const uint32_t st0 = umma::smem_u32(smem + p.off_stage);
    // stride between 8-element groups along M/N = the k groups of one octet.  Core matrices stay 128-byte
    // aligned (one shared-memory wavefront each for the tensor core); the conversion below deals (row, octet)
    // pairs to lanes diagonally so that both its reads and its 16-byte stores spread over the banks.
    const uint32_t sbo = (uint32_t)(KD / 8) * 128u;
    const uint32_t piece = (uint32_t)KD * Fa * 4u;      // one k piece of the ring: [KD][Fa] fp32

    if (tid == 0) {
        umma::mbar_init(full, 1);
        umma::mbar_init(full + 1, 1);
        umma::mbar_init(mbar, 1);
        umma::mbar_init(mbar + 1, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    {   // rows of the A operand beyond Ql are never written: clear both planes once
        uint4 *z = reinterpret_cast<uint4 *>(smem + p.off_stage);
        const int n16 = (int)(p.nstage * p.stage_bytes / 16);
        for (int i = tid; i < n16; i += DT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    // ring loads of chunk c, issued by the 32 lanes of the issue warp: nk stack pieces + KD rows of T
    auto issue_loads = [&](int c) {
        const long long rb = r_beg + (long long)c * KD;
        const int rows = (int)min((long long)KD, r_end - rb);
        const uint32_t dst = ring0 + (uint32_t)(c & 1) * p.ring_bytes;
        uint64_t *bar = full + (c & 1);
        if (lane == 0) {
            umma::fence_proxy_async();      // the slot was read by the conversion of chunk c-2 (generic proxy)
            mbar_expect_tx(bar, (uint32_t)rows * (uint32_t)(Ql + Fb) * 4u);
        }
        __syncwarp();
        for (int k = lane; k < nk; k += 32)
            bulk_g2s(dst + (uint32_t)k * piece, p.stack + ((long long)(k_lo + k) * p.R + rb) * Fa, (uint32_t)rows * Fa * 4u,
                     bar);
        if (p.sample_major) {       // T rows follow the stack rows: one contiguous block
            if (lane == 0) bulk_g2s(dst + p.ring_b_off, p.T + rb * Fb, (uint32_t)rows * Fb * 4u, bar);
        } else {
            const long long m0 = rb / p.N;
            const int n0 = (int)(rb - m0 * p.N);
            for (int kr = lane; kr < rows; kr += 32) {
                const int nn = n0 + kr;                     // row r = rb + kr = (m0 + nn / N) * N + nn % N
                const int dm = nn / p.N, n = nn - dm * p.N;
                const long long tr = (long long)n * p.M + (m0 + dm);
                bulk_g2s(dst + p.ring_b_off + (uint32_t)kr * Fb * 4u, p.T + tr * Fb, (uint32_t)Fb * 4u, bar);
            }
        }
    };

    if (warp == DC / 32 + 1) {
        // =========================== TMA producer warp ===================================
        if (nchunks > 0) issue_loads(0);
        if (nchunks > 1) issue_loads(1);
        for (int c = 0; c < nchunks; ++c) {
            __syncthreads();                              // ring slot c&1 has been converted
            if (c + 2 < nchunks) issue_loads(c + 2);
        }
    } else if (is_issuer) {
        // =========================== MMA issue warp ======================================
        const uint32_t idesc = umma::make_idesc_bf16(128, Fb, 1, 1);
        const uint32_t d_hi = umma::desc_hi(sbo);
        const uint32_t t_step = (16u * sbo) >> 4;           // 128 q of a row tile
        for (int c = 0; c < nchunks; ++c) {
            __syncthreads();                              // operands of chunk c are staged
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && c >= 8 && c < 24 && lane == 0;
            if (tr) p.trace[(c - 8) * 8 + 5] = clock64();
            if (umma::elect_one()) {
                umma::fence_after_sync();
                const uint32_t sb = st0 + (uint32_t)(c % p.nstage) * p.stage_bytes;
                const uint32_t a_lo = umma::desc_lo(sb, 128u), b_lo = umma::desc_lo(sb + p.stage_b_off, 128u);
                const uint32_t a_mid = p.stage_a_plane >> 4, b_mid = p.stage_b_plane >> 4;
                for (int t = 0; t < tiles; ++t) {
                    const uint32_t acc = tmem + (uint32_t)(t * Fb);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        if (pass >= p.npass) break;
                        uint32_t al = a_lo + (uint32_t)t * t_step + (pass == 1 ? a_mid : 0u);
                        uint32_t bl = b_lo + (pass == 2 ? b_mid : 0u);
                        for (int j = 0; j < KD / 16; ++j) {
                            if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && c == 30 && t * 3 + pass < 16)
                                p.trace[64 + t * 3 + pass] = clock64();
                            umma::mma_bf16(acc, umma::desc_join(al, d_hi), umma::desc_join(bl, d_hi), idesc,
                                           (c | pass | j) != 0);
                            al += 16u;                      // two k groups of 128 bytes
                            bl += 16u;
                        }
                    }
                }
                umma::commit(mbar + (c & 1));
            }
            __syncwarp();
            if (tr) p.trace[(c - 8) * 8 + 6] = clock64();
        }
        // nothing may be in flight when the CTA retires
        if (lane == 0) {
            if (nchunks > 1) umma::mbar_wait(mbar + ((nchunks - 2) & 1), (uint32_t)(((nchunks - 2) >> 1) & 1));
            if (nchunks > 0) umma::mbar_wait(mbar + ((nchunks - 1) & 1), (uint32_t)(((nchunks - 1) >> 1) & 1));
        }
        __syncwarp();
    } else {
        const int kd_shift = 31 - __clz(KD);                 // KD is a power of two
        for (int c = 0; c < nchunks; ++c) {
            const long long rb = r_beg + (long long)c * KD;
            const int rows = (int)min((long long)KD, r_end - rb);
            const unsigned char *ring = smem + p.off_ring + (size_t)(c & 1) * p.ring_bytes;
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && c >= 8 && c < 24 && tid == 0;
            if (tr) p.trace[(c - 8) * 8 + 0] = clock64();
            umma::mbar_wait(full + (c & 1), (uint32_t)((c >> 1) & 1));
            if (tr) p.trace[(c - 8) * 8 + 1] = clock64();
            if (c >= p.nstage) {                          // the operand stage is free once its previous MMAs completed
                const int cp = c - p.nstage;
                umma::mbar_wait(mbar + (cp & 1), (uint32_t)((cp >> 1) & 1));
            }
            unsigned char *stg = smem + p.off_stage + (size_t)(c % p.nstage) * p.stage_bytes;
            if (tr) p.trace[(c - 8) * 8 + 2] = clock64();
            // ---- A: stack pieces -> MN-major hi/mid planes
            if (FA8) {
                // the ring holds [k][kr][Fa] contiguously: octet e lives at byte 32 e;  e = (k*KD + kr)*OA + oa
                const int OA = Fa / 8;
                const int total = nk * KD * OA;
                for (int e = tid; e < total; e += DC) {
                    int kr, k, oa;
                    if (p.oa_shift >= 0) {
                        // blocks of 8 rows x OA octets; lane (i, ph) of a block takes row i, octet (i + ph) % OA:
                        // a quarter-warp stores 8 different rows (128 contiguous bytes of one or more core
                        // matrices) and reads 4 different 32-byte slots of the ring rows
                        const int blk = e >> (3 + p.oa_shift), b = e & (8 * OA - 1);
                        const int i = b & 7, ph = b >> 3;
                        const int kg = KD >> 3;
                        k = blk / kg;
                        kr = (blk - k * kg) * 8 + i;
                        oa = (i + ph) & (OA - 1);
                    } else {
                        const int t = e / OA;
                        oa = e - t * OA;
                        kr = t & (KD - 1);
                        k = t >> kd_shift;
                    }
                    const int es = (k * KD + kr) * OA + oa;          // source octet index in the ring
                    float v[8];
                    if (kr < rows) {
                        const float4 *src = reinterpret_cast<const float4 *>(ring) + (size_t)es * 2;
                        const float4 v0 = src[0], v1 = src[1];
                        v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w;
                        v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    }
                    uint4 hi, mid;
                    split8(v, hi, mid);
                    const uint32_t off = (uint32_t)(k * OA + oa) * sbo + (uint32_t)(kr >> 3) * 128u + (uint32_t)(kr & 7) * 16u;
                    *reinterpret_cast<uint4 *>(stg + off) = hi;
                    *reinterpret_cast<uint4 *>(stg + p.stage_a_plane + off) = mid;
                }
            } else {
                const int nocts = (Ql + 7) / 8;
                for (int e = tid; e < nocts * KD; e += DC) {
                    const int oct = e >> kd_shift, kr = e & (KD - 1);
                    float v[8];
                    int k = (oct * 8) / Fa, a = oct * 8 - k * Fa;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float x = 0.f;
                        if (oct * 8 + i < Ql && kr < rows) x = reinterpret_cast<const float *>(ring + (size_t)k * piece)[kr * Fa + a];
                        v[i] = x;
                        if (++a == Fa) { a = 0; ++k; }
                    }
                    uint4 hi, mid;
                    split8(v, hi, mid);
                    const uint32_t off = (uint32_t)oct * sbo + (uint32_t)(kr >> 3) * 128u + (uint32_t)(kr & 7) * 16u;
                    *reinterpret_cast<uint4 *>(stg + off) = hi;
                    *reinterpret_cast<uint4 *>(stg + p.stage_a_plane + off) = mid;
                }
            }
            // ---- B: T rows -> MN-major hi/mid planes
            {
                const int OB = Fb / 8;
                for (int e = tid; e < KD * OB; e += DC) {
                    int kr, ob;
                    if (p.ob_shift >= 0) {       // same diagonal dealing as for A
                        const int blk = e >> (3 + p.ob_shift), b = e & (8 * OB - 1);
                        const int i = b & 7, ph = b >> 3;
                        kr = blk * 8 + i;
                        ob = (i + ph) & (OB - 1);
                    } else {
                        kr = e / OB;
                        ob = e - kr * OB;
                    }
                    float v[8];
                    if (kr < rows) {
                        const float4 *src = reinterpret_cast<const float4 *>(ring + p.ring_b_off) + (size_t)(kr * OB + ob) * 2;
                        const float4 v0 = src[0], v1 = src[1];
                        v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w;
                        v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    }
                    uint4 hi, mid;
                    split8(v, hi, mid);
                    const uint32_t off = (uint32_t)ob * sbo + (uint32_t)(kr >> 3) * 128u + (uint32_t)(kr & 7) * 16u;
                    *reinterpret_cast<uint4 *>(stg + p.stage_b_off + off) = hi;
                    *reinterpret_cast<uint4 *>(stg + p.stage_b_off + p.stage_b_plane + off) = mid;
                }
            }
            if (tr) p.trace[(c - 8) * 8 + 3] = clock64();
            umma::fence_proxy_async();
            __syncthreads();
            if (tr) p.trace[(c - 8) * 8 + 4] = clock64();
        }
        // ---- epilogue: TMEM -> partial result
        if (nchunks > 0) {
            umma::mbar_wait(mbar + ((nchunks - 1) & 1), (uint32_t)(((nchunks - 1) >> 1) & 1));
            umma::fence_after_sync();
        }
        const int qd = warp & 3, wq = warp >> 2;
        const int nc8 = Fb / 8;
        float *dst0 = p.part + ((size_t)blockIdx.x * p.K * Fa + (size_t)k_lo * Fa) * Fb;
        for (int idx = wq; idx < tiles * nc8; idx += 4) {
            const int t = idx / nc8, c8 = idx - t * nc8;
            const int q = t * 128 + 32 * qd + lane;
            float v[8];
            if (nchunks > 0) {
                umma::tmem_ld8(tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(t * Fb + c8 * 8), v);
                umma::tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            if (q < Ql) {
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

static int pow2_shift(int v) {       // log2(v) when v is a power of two, else -1
    if (v <= 0 || (v & (v - 1)) != 0) return -1;
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

struct DwPlan {
    bool ok = false;
    int KD = 0, kgroup = 0, ngroups = 0, splits = 0, tmem_cols = 0;
    long long rows_per_cta = 0;
    size_t smem = 0;
    DwParams dp;
};

static DwPlan dw_plan(int N, int M, int Fa, int Fb, int K, int sm_count, size_t smem_limit) {
    DwPlan pl;
    const long long R = (long long)N * M;
    if (Fb % 16 != 0 || Fb < 16 || Fb > 256 || Fa < 1 || K < 1 || R < 1) return pl;
    if (((long long)R * Fa) % 4 != 0) return pl;             // 16-byte aligned k slabs for the bulk copies
    // k-slabs per CTA: the accumulator [ceil(kgroup*Fa/128)*128][Fb] must fit the 512 TMEM columns.  Prefer a
    // split that leaves room for two operand stages (conversion of chunk c+1 under the MMAs of chunk c).
    const int max_tiles = 512 / Fb;
    int kmax = K;
    while (kmax > 1 && cg_ceil_div((int64_t)kmax * Fa, 128) > max_tiles) --kmax;
    if (cg_ceil_div((int64_t)kmax * Fa, 128) > max_tiles) return pl;
    int a_lo = 0;
    if (const char *env = getenv("CG_DW_ATTEMPT")) a_lo = atoi(env);     // tuning aid
    for (int attempt = a_lo; attempt < 8 && !pl.ok; ++attempt) {
        // attempts 0..3: two stages with 1, 2, 3, 4 times the minimum number of k groups; then one stage
        const int nstage = attempt < 4 ? 2 : 1;
        const int mult = attempt < 4 ? attempt + 1 : attempt - 3;
        const int ngroups = (int)cg_ceil_div(K, kmax) * mult;
        if (ngroups > K) continue;
        const int kgroup = (int)cg_ceil_div(K, ngroups);
        const int Ql = kgroup * Fa;
        const int tiles = (int)cg_ceil_div(Ql, 128);
        const int Qp = tiles * 128;
        for (int KD = 64; KD >= 16 && !pl.ok; KD /= 2) {
            if (((long long)KD * Fa) % 4 != 0) continue;
            DwParams dp;
            memset(&dp, 0, sizeof(dp));
    dp.npass = cg_mma_passes();
            uint32_t off = 0;
            dp.off_bar = off;
            off += 128;
            dp.off_ring = off;
            dp.ring_b_off = (uint32_t)cg_align_up((size_t)Ql * KD * 4, 128);
            dp.ring_bytes = dp.ring_b_off + (uint32_t)cg_align_up((size_t)Fb * KD * 4, 128);
            off += 2 * dp.ring_bytes;
            dp.off_stage = off;
            const uint32_t sbo = (uint32_t)(KD / 8) * 128u;     // must match the kernel
            dp.stage_a_plane = (uint32_t)(Qp / 8) * sbo;
            dp.stage_b_off = 2 * dp.stage_a_plane;
            dp.stage_b_plane = (uint32_t)(Fb / 8) * sbo;
            dp.stage_bytes = 2 * dp.stage_a_plane + 2 * dp.stage_b_plane;
            off += (uint32_t)nstage * dp.stage_bytes;
            if (off > smem_limit) continue;
            pl.ok = true;
            pl.KD = KD;
            pl.kgroup = kgroup;
            pl.ngroups = (int)cg_ceil_div(K, kgroup);
            int cols = 32;
            while (cols < tiles * Fb) cols *= 2;
            pl.tmem_cols = cols;
            int splits = std::max(1, sm_count / pl.ngroups);
            const long long max_splits = cg_ceil_div(R, 4LL * KD);
            if (splits > max_splits) splits = (int)max_splits;
            long long rpc = cg_ceil_div(R, splits);
            rpc = cg_ceil_div(rpc, KD) * KD;
            pl.splits = (int)cg_ceil_div(R, rpc);
            pl.rows_per_cta = rpc;
            dp.KD = KD;
            dp.kgroup = kgroup;
            dp.tmem_cols = cols;
            dp.nstage = nstage;
            dp.oa_shift = pow2_shift(Fa / 8);
            dp.ob_shift = pow2_shift(Fb / 8);
            dp.rows_per_cta = rpc;
            pl.dp = dp;
            pl.smem = off;
        }
    }
    return pl;
}

}  // namespace

static long long *g_dw_trace = nullptr;
extern "C" int cg_debug_dw_trace(long long *dev_buf) {
    g_dw_trace = dev_buf;
    return CG_OK;
}

bool cg_dw_umma_supported(int N, int M, int Fa, int Fb, int K, int sm_count, size_t smem_limit) {
    return dw_plan(N, M, Fa, Fb, K, sm_count, smem_limit).ok;
}

size_t cg_dw_umma_workspace(int N, int M, int Fa, int Fb, int K, int sm_count, size_t smem_limit) {
    const DwPlan pl = dw_plan(N, M, Fa, Fb, K, sm_count, smem_limit);
    return pl.ok ? sizeof(float) * (size_t)pl.splits * K * Fa * Fb : 0;
}

int cg_run_dw_umma(const float *stack, const float *T, float *dW, int N, int M, int Fa, int Fb, int K, bool swap,
                   bool sample_major, float *workspace, int sm_count, size_t smem_limit, cudaStream_t s) {
    DwPlan pl = dw_plan(N, M, Fa, Fb, K, sm_count, smem_limit);
    CG_REQUIRE(pl.ok, "cg_run_dw_umma: shape not supported (Fa=%d Fb=%d K=%d)", Fa, Fb, K);
    CG_REQUIRE((((uintptr_t)stack | (uintptr_t)T | (uintptr_t)workspace) & 15) == 0, "cg_run_dw_umma: unaligned tensor");
    DwParams &dp = pl.dp;
    dp.stack = stack;
    dp.T = T;
    dp.part = workspace;
    dp.R = (long long)N * M;
    dp.N = N;
    dp.M = M;
    dp.Fa = Fa;
    dp.Fb = Fb;
    dp.K = K;
    dp.sample_major = sample_major ? 1 : 0;
    dp.trace = g_dw_trace;
    dim3 grid((unsigned)pl.splits, (unsigned)pl.ngroups);
    {
        CgProfScope prof("dw_umma", s);
        if (Fa % 8 == 0) {
            CG_CHECK_CUDA(cudaFuncSetAttribute(k_dw_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            k_dw_umma<true><<<grid, DT, pl.smem, s>>>(dp);
        } else {
            CG_CHECK_CUDA(cudaFuncSetAttribute(k_dw_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            k_dw_umma<false><<<grid, DT, pl.smem, s>>>(dp);
        }
        CG_LAUNCH_CHECK();
    }
    return cg_reduce_partials(workspace, dW, pl.splits, Fa, Fb, K, swap, s);
}
