// Streaming form of the fp32 tensor-core GEMM:  C[M x N] = op(A)[M x K] . op(B)[K x N] (+ bias[N]) (relu)
//
// cg_gemm_pipe.cu's converter warps spend most of their instruction issue on address generation for their own
// cp.async copies and on re-splitting the B tile for every output tile (ncu, C5 contraction shape: 4 600 warp
// instructions per K stage, 68 % issue utilisation at 0.33 of the HBM rate).  This kernel serves the shapes that
// stream a large A (the Chebyshev stack of lib/models.py:192-224 / lib/filter.py:89-95 as the left operand of the
// (Fin*K) x Fout contraction, of its input gradient and -- transposed -- of its weight gradient):
//   * B is split ONCE per call into bf16 hi | mid planes, stage by stage in the K-major canonical operand layout
//     (k_pack_b), and every stage of it arrives by ONE bulk copy straight into the MMA operand slot;
//   * the fp32 A tile of a stage (128 rows x 32 k) arrives by ONE tensor copy (cp.async.bulk.tensor.3d over a tensor
//     map of A that also expresses the [K][rows][F] blocking of a Chebyshev stack; tile tails are zero-filled by the
//     copy engine) into a raw ring, issued by one elected producer thread;
//   * the 16 converter warps only read their 2 x 16 bytes back, split them and store the half octets: same values,
//     same MMA order as cg_gemm_pipe.cu, so the results are bit-identical to that kernel's.
// Restrictions (the caller falls back to cg_gemm_pipe.cu otherwise): K blocks multiples of 32, row blocks powers of two
// >= 32, 16-byte aligned A with a leading dimension that is a multiple of 4.  A K tail needs nothing: the tensor copy
// zero-fills A beyond K and k_pack_b zero-fills B.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "cg_common.cuh"
#include "cg_umma.cuh"
#include "cg_fused_common.cuh"

namespace {

constexpr int SC = 512;                 // converter threads
constexpr int SE = 128;                 // epilogue threads
constexpr int ST = SC + SE + 64;        // + issue warp + producer warp
constexpr int BM = 128;
constexpr int BK = 32;
constexpr int MAX_STAGES = 6;
constexpr int MAX_RAW = 8;
constexpr uint32_t RAW_BYTES = BM * BK * 4;         // one stage of fp32 A
constexpr uint32_t MN_SBO = BK * 16 + 32;           // row-contiguous A: stride between 8-row groups (padded)
constexpr uint32_t KC_LBO = BM * 16 + 32;           // K-contiguous A: stride between k octets (padded)
constexpr uint32_t EP_ROW = 80;                     // epilogue staging: 16 floats + 16 bytes of padding per row
constexpr uint32_t EP_BYTES = 32 * EP_ROW;          // per epilogue warp

__host__ __device__ constexpr uint32_t a_plane_bytes(bool ta) { return ta ? (BM / 8) * MN_SBO : 4u * KC_LBO; }

struct StreamParams {
    int npass;
    const float *A, *bias;
    const unsigned char *Bp;            // packed B: [tiles_n][K / 32][hi plane | mid plane], plane = BN * 64 bytes
    float *C;                           // [M][ldc] (split == 1) or partials [split][M][N]
    int M, N, K, lda, ldc, relu, BN, split, k_per_split, tiles_n, n_work, nstage, nraw;
    // tensor-map coordinates of a tile: K-contiguous A (k & a_mask, m, k >> a_sh); row-contiguous A (m & a_mask, k, m >> a_sh)
    int a_sh, a_mask;
    int a_chunk;                        // row-contiguous A: contiguous rows per box line (min(128, row block)); raw tile = [128 / chunk][32 k][chunk]
    uint32_t a_plane, b_plane, off_b, stage_bytes, off_raw, off_ep, off_bar, tmem_cols;
};

template <int SLEEP>
__device__ __forceinline__ void wait_warp(uint64_t *bar, uint32_t parity, int lane) {
    if (lane == 0) {
        const uint32_t addr = umma::smem_u32(bar);
        uint32_t done = 0;
        while (true) {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}\n"
                : "=r"(done)
                : "r"(addr), "r"(parity)
                : "memory");
            if (done) break;
            if (SLEEP > 0) __nanosleep(SLEEP);
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void tmem_ld16s(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// one box of the tensor map -> shared memory, completion (full box bytes, out-of-range elements zero-filled) on the mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst_smem),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// B -> packed planes.  One thread = (n tile, K stage, k octet, n): eight k values of one column, 16 bytes of hi and of mid.
//   op(B)(k, n) = TB ? B[n * ldb + k] : B[row(k) * ldb + n],  row(k) = (k >> sh) * shi + (k & mask) * slo
__global__ void __launch_bounds__(256) k_pack_b(const float *__restrict__ B, unsigned char *__restrict__ out, int N, int K, int ldb, int tb,
                                                int BN, int nks, long long total, int b_sh, int b_mask, int b_shi, int b_slo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int nl = (int)(idx % BN);
    const int o = (int)((idx / BN) & 3);
    const long long st = idx / (4LL * BN);          // tn * nks + ks
    const int ks = (int)(st % nks), tn = (int)(st / nks);
    const int n = tn * BN + nl, k0 = ks * BK + o * 8;
    uint32_t hi[4], mid[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float v[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int k = k0 + 2 * i + j;
            v[j] = 0.f;
            if (n < N && k < K)
                v[j] = tb ? B[(size_t)n * ldb + k] : B[((size_t)(k >> b_sh) * b_shi + (size_t)(k & b_mask) * b_slo) * ldb + n];
        }
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
        const float2 f = __bfloat1622float2(h);
        const __nv_bfloat162 m = __floats2bfloat162_rn(v[0] - f.x, v[1] - f.y);
        hi[i] = *reinterpret_cast<const uint32_t *>(&h);
        mid[i] = *reinterpret_cast<const uint32_t *>(&m);
    }
    const size_t plane = (size_t)BN * 64;
    unsigned char *dst = out + (size_t)st * 2 * plane + (size_t)o * BN * 16 + (size_t)nl * 16;
    *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(dst + plane) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
}

// TA: op(A)(m, k) = A[k * lda + m] (row index contiguous, MN-major operand); otherwise A[m * lda + k] (K-major)
template <bool TA>
__global__ void __launch_bounds__(ST, 1) k_gemm_stream(const __grid_constant__ CUtensorMap map_a, const StreamParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *full = bars;                              // [MAX_STAGES] A planes converted + B planes landed
    uint64_t *empty = bars + MAX_STAGES;                // [MAX_STAGES] MMAs of the stage completed
    uint64_t *raw_full = bars + 2 * MAX_STAGES;         // [MAX_RAW] fp32 A of the stage landed
    uint64_t *raw_empty = raw_full + MAX_RAW;           // [MAX_RAW] converters are done with it
    uint64_t *acc_ready = raw_empty + MAX_RAW;          // [2]
    uint64_t *acc_free = acc_ready + 2;                 // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_free + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN, nstage = p.nstage, nraw = p.nraw;

    if (tid == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            umma::mbar_init(full + i, SC / 32 + 1);
            umma::mbar_init(empty + i, 1);
        }
        for (int i = 0; i < MAX_RAW; ++i) {
            umma::mbar_init(raw_full + i, 1);
            umma::mbar_init(raw_empty + i, SC / 32);
        }
        umma::mbar_init(acc_ready, 1);
        umma::mbar_init(acc_ready + 1, 1);
        umma::mbar_init(acc_free, SE / 32);
        umma::mbar_init(acc_free + 1, SE / 32);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, p.tmem_cols);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t st0 = umma::smem_u32(smem);

    auto item = [&](int w, int &m0, int &n0, int &tn, int &k_beg, int &k_end, int &sp) {
        sp = w % p.split;
        const int t = w / p.split;
        const int tm = t / p.tiles_n;
        tn = t - tm * p.tiles_n;
        m0 = tm * BM;
        n0 = tn * BN;
        k_beg = sp * p.k_per_split;
        k_end = min(p.K, k_beg + p.k_per_split);
    };

    if (warp < SC / 32) {
        // =========================== converter warps ====================================
        int total = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int m0, n0, tn, k_beg, k_end, sp;
            item(w, m0, n0, tn, k_beg, k_end, sp);
            total += (k_end - k_beg + BK - 1) / BK;
        }
        // this thread's two 16-byte pieces of the raw tile and the half octets they become
        uint32_t src, src_step, dst;
        {
            const uint32_t i = (uint32_t)(lane >> 3), o = (uint32_t)((lane & 7) >> 1), half = (uint32_t)(lane & 1) * 8u;
            if (!TA) {      // raw [128 rows][128 B]: rows 8 * warp + i and + 4, floats 4 * (lane & 7) ..
                src = ((uint32_t)(8 * warp) + i) * 128u + (uint32_t)(lane & 7) * 16u;
                src_step = 512u;
                dst = o * KC_LBO + (uint32_t)warp * 128u + i * 16u + half;
            } else {        // raw [128 / chunk][32 k][chunk rows]: k = 8 * (warp & 3) + i and + 4, rows 32 * (warp >> 2) + 4 * (lane & 7) ..
                const uint32_t ch = (uint32_t)p.a_chunk, m = (uint32_t)(32 * (warp >> 2) + 4 * (lane & 7));
                src = (m / ch) * (BK * ch * 4u) + ((uint32_t)(8 * (warp & 3)) + i) * ch * 4u + (m % ch) * 4u;
                src_step = 4u * ch * 4u;
                dst = ((uint32_t)(warp >> 2) * 4u + o) * MN_SBO + (uint32_t)(warp & 3) * 128u + i * 16u + half;
            }
        }
        const uint32_t raw0 = st0 + p.off_raw + src;
        int slot = 0, rs = 0;
        uint32_t use = 0, ruse = 0, sb = st0 + dst, rb = raw0;
        for (int gc = 0; gc < total; ++gc) {
            wait_warp<0>(raw_full + rs, ruse & 1, lane);
            const float4 a0 = lds128(rb), a1 = lds128(rb + src_step);
            if (use > 0) wait_warp<0>(empty + slot, (use - 1) & 1, lane);
            uint2 h0, m0, h1, m1;
            split4(a0, h0, m0);
            split4(a1, h1, m1);
            sts64(sb, h0);
            sts64(sb + p.a_plane, m0);
            sts64(sb + 64u, h1);
            sts64(sb + p.a_plane + 64u, m1);
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(full + slot);
                mbar_arrive(raw_empty + rs);
            }
            sb += p.stage_bytes;
            if (++slot == nstage) {
                slot = 0;
                sb = st0 + dst;
                ++use;
            }
            rb += RAW_BYTES;
            if (++rs == nraw) {
                rs = 0;
                rb = raw0;
                ++ruse;
            }
        }
    } else if (warp == (SC + SE) / 32 + 1) {
        // =========================== producer warp (one elected thread) ===================
        if (umma::elect_one()) {
            const uint32_t bstage = 2u * p.b_plane;
            const int nks = (p.K + BK - 1) / BK;
            int slot = 0, rs = 0;
            uint32_t use = 0, ruse = 0;
            for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
                int m0, n0, tn, k_beg, k_end, sp;
                item(w, m0, n0, tn, k_beg, k_end, sp);
                const int nst = (k_end - k_beg + BK - 1) / BK;
                const unsigned char *bsrc = p.Bp + ((size_t)tn * nks + (size_t)(k_beg / BK)) * bstage;
                for (int s = 0; s < nst; ++s) {
                    const int k0 = k_beg + s * BK;
                    if (ruse > 0) umma::mbar_wait(raw_empty + rs, (ruse - 1) & 1);
                    mbar_expect_tx(raw_full + rs, RAW_BYTES);
                    const uint32_t raw = st0 + p.off_raw + (uint32_t)rs * RAW_BYTES;
                    if (!TA) tma_load_3d(raw, &map_a, raw_full + rs, k0 & p.a_mask, m0, k0 >> p.a_sh);
                    else tma_load_3d(raw, &map_a, raw_full + rs, m0 & p.a_mask, k0, m0 >> p.a_sh);
                    if (use > 0) umma::mbar_wait(empty + slot, (use - 1) & 1);
                    mbar_expect_tx(full + slot, bstage);
                    bulk_g2s(st0 + (uint32_t)slot * p.stage_bytes + p.off_b, bsrc + (size_t)s * bstage, bstage, full + slot);
                    if (++slot == nstage) {
                        slot = 0;
                        ++use;
                    }
                    if (++rs == nraw) {
                        rs = 0;
                        ++ruse;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == (SC + SE) / 32) {
        // =========================== MMA issue warp ======================================
        const uint32_t idesc = umma::make_idesc_bf16(BM, BN, TA ? 1 : 0, 0);
        const uint32_t a_lbo = TA ? 128u : KC_LBO, a_sbo = TA ? MN_SBO : 128u;
        const uint32_t b_lbo = (uint32_t)BN * 16u, b_sbo = 128u;
        const uint32_t a_hi = umma::desc_hi(a_sbo), b_hi = umma::desc_hi(b_sbo);
        const uint32_t a_k = (TA ? 256u : 2u * a_lbo) >> 4, b_k = (2u * b_lbo) >> 4;
        int it = 0, slot = 0;
        uint32_t use = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x, ++it) {
            int m0, n0, tn, k_beg, k_end, sp;
            item(w, m0, n0, tn, k_beg, k_end, sp);
            const int nst = (k_end - k_beg + BK - 1) / BK;
            const int ab = it & 1;
            if (it >= 2) wait_warp<64>(acc_free + ab, (uint32_t)(((it >> 1) - 1) & 1), lane);
            umma::fence_after_sync();
            const uint32_t acc = tmem + (uint32_t)(ab * BN);
            for (int s = 0; s < nst; ++s) {
                wait_warp<0>(full + slot, use & 1, lane);
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint32_t sb = st0 + (uint32_t)slot * p.stage_bytes;
                    const uint32_t a_lo = umma::desc_lo(sb, a_lbo), b_lo = umma::desc_lo(sb + p.off_b, b_lbo);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        if (pass >= p.npass) break;
                        uint32_t al = a_lo + (pass == 1 ? (p.a_plane >> 4) : 0u), bl = b_lo + (pass == 2 ? (p.b_plane >> 4) : 0u);
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j) {
                            umma::mma_bf16(acc, umma::desc_join(al, a_hi), umma::desc_join(bl, b_hi), idesc, (s | pass | j) != 0);
                            al += a_k;
                            bl += b_k;
                        }
                    }
                    umma::commit(empty + slot);
                    if (s == nst - 1) umma::commit(acc_ready + ab);
                }
                __syncwarp();
                if (++slot == nstage) {
                    slot = 0;
                    ++use;
                }
            }
        }
    } else {
        // =========================== epilogue warps ======================================
        const int qd = warp & 3;
        const bool final_out = p.split == 1;
        const uint32_t ep = st0 + p.off_ep + (uint32_t)qd * EP_BYTES;
        const int ld = final_out ? p.ldc : p.N;
        const bool vec = (ld & 3) == 0 && ((((uintptr_t)p.C) & 15) == 0);
        const int cl = 4 * (lane & 3), rl = lane >> 2;
        int it = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x, ++it) {
            int m0, n0, tn, k_beg, k_end, sp;
            item(w, m0, n0, tn, k_beg, k_end, sp);
            const int ab = it & 1;
            wait_warp<256>(acc_ready + ab, (uint32_t)((it >> 1) & 1), lane);
            umma::fence_after_sync();
            float *cbase = final_out ? p.C : p.C + (size_t)sp * p.M * p.N;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                if (n0 + c0 >= p.N) break;
                float v[16];
                tmem_ld16s(tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(ab * BN + c0), v);
                umma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sts128(ep + (uint32_t)lane * EP_ROW + (uint32_t)j * 16u, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                __syncwarp();
                const int n = n0 + c0 + cl;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (final_out && p.bias) {
                    if (n < p.N) bv.x = p.bias[n];
                    if (n + 1 < p.N) bv.y = p.bias[n + 1];
                    if (n + 2 < p.N) bv.z = p.bias[n + 2];
                    if (n + 3 < p.N) bv.w = p.bias[n + 3];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int row = 8 * j + rl;
                    float4 x = lds128(ep + (uint32_t)row * EP_ROW + (uint32_t)cl * 4u);
                    const int m = m0 + 32 * qd + row;
                    if (final_out) {
                        x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
                        if (p.relu) {
                            x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
                        }
                    }
                    if (m < p.M) {
                        float *dstp = cbase + (size_t)m * ld + n;
                        if (vec && n + 3 < p.N) {
                            *reinterpret_cast<float4 *>(dstp) = x;
                        } else {
                            if (n < p.N) dstp[0] = x.x;
                            if (n + 1 < p.N) dstp[1] = x.y;
                            if (n + 2 < p.N) dstp[2] = x.z;
                            if (n + 3 < p.N) dstp[3] = x.w;
                        }
                    }
                }
                __syncwarp();
            }
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free + ab);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, p.tmem_cols);
}

static int pow2_shift(int v) {
    if (v <= 0 || (v & (v - 1)) != 0) return -1;
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

struct StreamPlan {
    int BN, split, k_per_split, tiles, nstage, nraw;
    size_t smem, ws_part, ws_pack;
    StreamParams pp;
};

static StreamPlan stream_plan(int M, int N, int K, bool ta, int sm_count) {
    StreamPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.pp.npass = cg_mma_passes();
    // as few column tiles as possible (every column tile converts the A tile again), of equal width (a multiple of 16)
    const int tiles_n = (int)cg_ceil_div(N, 256);
    pl.BN = std::max(32, (int)cg_ceil_div(cg_ceil_div(N, tiles_n), 16) * 16);
    pl.tiles = (int)cg_ceil_div(M, BM) * tiles_n;
    int split = 1;
    if (pl.tiles < sm_count) {
        split = std::max(1, sm_count / pl.tiles);
        const int max_split = (int)cg_ceil_div(K, 8 * BK);
        if (split > max_split) split = std::max(1, max_split);
    }
    int kps = (int)cg_ceil_div(K, split);
    kps = (int)cg_ceil_div(kps, BK) * BK;
    pl.split = (int)cg_ceil_div(K, kps);
    pl.k_per_split = kps;
    StreamParams &pp = pl.pp;
    pp.a_plane = a_plane_bytes(ta);
    pp.b_plane = (uint32_t)pl.BN * 64u;
    pp.off_b = 2 * pp.a_plane;
    pp.stage_bytes = (uint32_t)cg_align_up(2 * pp.a_plane + 2 * pp.b_plane, 128);
    const size_t ep = (size_t)(SE / 32) * EP_BYTES;
    const size_t budget = (size_t)227 * 1024 - ep - 512;
    // Raw stages hide the HBM latency of A, operand stages the conversion -> MMA -> release hand-over.  Measured at the C5
    // contraction shape (BN = 64), operand + raw stages: 6 + 4 1.13 ms, 5 + 5 1.10 ms, 4 + 6 1.27 ms, 3 + 8 1.56 ms -- the
    // hand-over needs the depth more than the copies do.  CG_GEMM_STREAM_RAW overrides the raw depth (experiments).
    static int raw_env = -1;
    if (raw_env < 0) {
        const char *e = getenv("CG_GEMM_STREAM_RAW");
        raw_env = e ? std::min(MAX_RAW, std::max(2, atoi(e))) : 0;
    }
    int raw_want = raw_env > 0 ? raw_env : ((budget - 5u * RAW_BYTES) / pp.stage_bytes >= 5 ? 5 : 4);
    if ((budget - (size_t)raw_want * RAW_BYTES) / pp.stage_bytes < 3) raw_want = 3;
    pl.nraw = raw_want;
    pl.nstage = (int)std::min<size_t>(MAX_STAGES, (budget - (size_t)pl.nraw * RAW_BYTES) / pp.stage_bytes);
    pp.off_raw = (uint32_t)pl.nstage * pp.stage_bytes;
    pp.off_ep = pp.off_raw + (uint32_t)pl.nraw * RAW_BYTES;
    pp.off_bar = pp.off_ep + (uint32_t)ep;
    pl.smem = pp.off_bar + 512;
    pp.tmem_cols = 32;
    while (pp.tmem_cols < 2u * (uint32_t)pl.BN) pp.tmem_cols *= 2;
    pp.tiles_n = tiles_n;
    pl.ws_part = pl.split > 1 ? cg_align_up(sizeof(float) * (size_t)pl.split * M * N, 256) : 0;
    pl.ws_pack = (size_t)tiles_n * (size_t)cg_ceil_div(K, BK) * 2 * pp.b_plane;
    return pl;
}

static int g_stream_on = -1;         // -1: not read yet (environment CG_GEMM_STREAM, default on)
static bool stream_enabled() {
    if (g_stream_on < 0) {
        const char *e = getenv("CG_GEMM_STREAM");
        g_stream_on = (e && e[0] == '0') ? 0 : 1;
    }
    return g_stream_on == 1;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else
            (void)cudaGetLastError();
        if (!fn) fprintf(stderr, "cnn_graph_b200: cuTensorMapEncodeTiled unavailable, streaming GEMM disabled\n");
    }
    return fn;
}
static bool encode_map(CUtensorMap *map, const float *A, const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box,
                       const cuuint32_t *estr) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(A), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static bool warned = false;
        if (!warned) {
            warned = true;
            fprintf(stderr, "cnn_graph_b200: cuTensorMapEncodeTiled failed (%d) for dims {%llu, %llu, %llu} strides {%llu, %llu} box {%u, %u, %u}\n",
                    (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                    (unsigned long long)strides[0], (unsigned long long)strides[1], box[0], box[1], box[2]);
        }
        return false;
    }
    return true;
}

}  // namespace

extern "C" int cg_debug_gemm_stream(int on) {
    const int before = stream_enabled() ? 1 : 0;
    if (on >= 0) g_stream_on = on ? 1 : 0;
    return before;
}

// shapes worth a packing launch: at least four row tiles re-use the packed B.  Measured (C4, weight gradient with M = 384
// and a 105 MB B): packing costs 0.11 ms per step and saves 0.07.  CG_GEMM_STREAM_MIN_M overrides, for experiments.
static bool stream_shape(int M, int N, int K) {
    static int min_m = -1;
    if (min_m < 0) {
        const char *e = getenv("CG_GEMM_STREAM_MIN_M");
        min_m = e ? std::max(1, atoi(e)) : 4 * BM;
    }
    return stream_enabled() && M >= min_m && N >= 1 && K >= BK;
}

size_t cg_gemm_stream_workspace(int M, int N, int K, int sm_count) {
    if (!stream_shape(M, N, K)) return 0;
    // the orientation of A does not change the split or the packed size
    const StreamPlan pl = stream_plan(M, N, K, false, sm_count);
    return pl.ws_part + pl.ws_pack;
}

bool cg_gemm_stream_eligible(const float *A, int M, int N, int K, int lda, int transA, int transB, int a_kblk, long long a_kbs,
                             int b_kblk, int a_mblk, long long a_mbs, size_t workspace_bytes, int sm_count) {
    if (!stream_shape(M, N, K)) return false;
    if ((((uintptr_t)A) & 15) != 0 || (lda & 3) != 0) return false;
    if (transA) {
        if (a_kblk > 0 || (M & 3) != 0) return false;
        if (a_mblk > 0 && (pow2_shift(a_mblk) < 5 || (a_mbs & 3) != 0)) return false;
    } else {
        if (a_mblk > 0) return false;
        if (a_kblk > 0 && (pow2_shift(a_kblk) < 5 || (a_kbs & 3) != 0)) return false;
    }
    if (b_kblk > 0 && (transB || pow2_shift(b_kblk) < 0)) return false;
    return workspace_bytes >= cg_gemm_stream_workspace(M, N, K, sm_count);
}

// same contract as cg_run_gemm_pipe; the caller has checked cg_gemm_stream_eligible
int cg_run_gemm_stream(const float *A, const float *B, float *C, int M, int N, int K, int transA, int transB, int lda, int ldb,
                       int ldc, const float *bias, int relu, int a_kblk, long long a_kbs, int b_kblk, int b_shi, int b_slo,
                       void *workspace, size_t workspace_bytes, int sm_count, cudaStream_t s, int a_mblk, long long a_mbs) {
    StreamPlan pl = stream_plan(M, N, K, transA != 0, sm_count);
    CG_REQUIRE(workspace && workspace_bytes >= pl.ws_part + pl.ws_pack, "cg_gemm_f32: workspace too small (%zu < %zu bytes)",
               workspace_bytes, pl.ws_part + pl.ws_pack);
    StreamParams &pp = pl.pp;
    pp.A = A;
    pp.bias = bias;
    pp.C = pl.split > 1 ? reinterpret_cast<float *>(workspace) : C;
    pp.M = M;
    pp.N = N;
    pp.K = K;
    pp.lda = lda;
    pp.ldc = ldc;
    pp.relu = relu ? 1 : 0;
    pp.BN = pl.BN;
    pp.split = pl.split;
    pp.k_per_split = pl.k_per_split;
    pp.n_work = pl.tiles * pl.split;
    pp.nstage = pl.nstage;
    pp.nraw = pl.nraw;
    // tensor map of A: [blocks][rows][contiguous run] with a box of one stage of one tile
    //   K-contiguous:   dims {k block (or K), M, k blocks},       box {32, 128, 1}
    //   row-contiguous: dims {row block (or M), K, row blocks},   box {chunk, 32, 128 / chunk}
    const int blk = transA ? a_mblk : a_kblk;
    const long long bstride = transA ? a_mbs : a_kbs;
    const int inner = transA ? M : K;
    pp.a_sh = blk > 0 ? pow2_shift(blk) : 31;
    pp.a_mask = blk > 0 ? blk - 1 : 0x7fffffff;
    pp.a_chunk = transA ? (blk > 0 ? std::min(BM, blk) : BM) : BK;
    CUtensorMap map;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)(blk > 0 ? blk : inner), (cuuint64_t)(transA ? K : M),
                                    (cuuint64_t)(blk > 0 ? inner / blk : 1)};
        const cuuint64_t strides[2] = {(cuuint64_t)lda * 4u, blk > 0 ? (cuuint64_t)bstride * 4u : (cuuint64_t)lda * 4u * dims[1]};
        const cuuint32_t box[3] = {(cuuint32_t)(transA ? pp.a_chunk : BK), (cuuint32_t)(transA ? BK : BM),
                                   (cuuint32_t)(transA ? BM / pp.a_chunk : 1)};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (!encode_map(&map, A, dims, strides, box, estr)) return CG_TRY_NEXT;
    }
    unsigned char *packed = reinterpret_cast<unsigned char *>(workspace) + pl.ws_part;
    const int nks = (int)cg_ceil_div(K, BK);
    {
        CgProfScope prof("gemm_pack_b", s);
        const long long total = (long long)pl.pp.tiles_n * nks * 4 * pl.BN;
        k_pack_b<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(B, packed, N, K, ldb, transB ? 1 : 0, pl.BN, nks, total,
                                                                 b_kblk > 0 ? pow2_shift(b_kblk) : 31, b_kblk > 0 ? b_kblk - 1 : 0x7fffffff,
                                                                 b_kblk > 0 ? b_shi : 0, b_kblk > 0 ? b_slo : 1);
        CG_LAUNCH_CHECK();
    }
    pp.Bp = packed;
    const unsigned grid = (unsigned)std::min(pp.n_work, sm_count);
    {
        CgProfScope prof("gemm_pipe", s);
        if (transA) {
            CG_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            k_gemm_stream<true><<<grid, ST, pl.smem, s>>>(map, pp);
        } else {
            CG_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            k_gemm_stream<false><<<grid, ST, pl.smem, s>>>(map, pp);
        }
        CG_LAUNCH_CHECK();
    }
    if (pl.split > 1) return cg_gemm_reduce(reinterpret_cast<const float *>(workspace), bias, C, M, N, ldc, pl.split, relu, s);
    return CG_OK;
}
