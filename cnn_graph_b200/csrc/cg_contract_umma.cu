// Weight contraction of the Chebyshev filter on the tensor cores for a basis that lives in HBM
// (lib/models.py:218-224: the [N*M, Fin*K] x [Fin*K, Fout] matmul, with the restack / transposes of
// :218-220 folded into the operand staging):
//
//     y[r][j] = sum_{k,f} stack[k][r][f] * W[f*K + k][j]          r = n*M + m  (sample-major basis)
//
// Used when the fused kernel (cg_fused.cu) does not take the shape -- narrow inputs (layer 1 of cgcnn has
// Fin = 1), or operators whose slabs do not fit shared memory.  Rows are the M dimension of the MMA, the
// flattened (k, f) index q is its K dimension:
//   * a TMA producer warp brings, per chunk of ROWS rows, the K contiguous pieces [ROWS][Fin] of the basis
//     into a 2-deep fp32 ring (cp.async.bulk);
//   * the compute warps split them into bf16 hi + mid and store the MN-major canonical A operand
//     (8 consecutive rows of one q = 16 bytes);  W is staged once per CTA as the MN-major B operand;
//   * the issue warp runs  hi*Whi + mid*Whi + hi*Wmid  into a double-buffered TMEM accumulator, and the
//     compute warps write chunk c-1 (thread = row, Fout contiguous floats) while chunk c is multiplied.
#include <stdlib.h>

#include <algorithm>

#include "cg_common.cuh"
#include "cg_umma.cuh"
#include "cg_fused_common.cuh"

namespace {

constexpr int XC = 512;          // compute threads
constexpr int XT = XC + 64;      // + MMA issue warp + TMA producer warp
constexpr int ROWS = 256;        // rows per chunk (two 128-row MMA tiles)

struct CtParams {
    int npass;                   // MMA passes per product: 3 (fp32-equivalent hi/mid split) or 1 (single-pass bf16)
    const float *stack;          // [K][R][Fa]
    const float *W;              // [Fa*K][J]  (row = f*K + k)
    float *y;                    // [R][J]
    // pooled epilogue (first layer, lib/models.py:226-257): relu(y + bias) max-pooled over groups of 4 consecutive rows
    const float *bias;           // [J] or NULL
    float *yp;                   // [R/4][J]; non-NULL selects the pooled epilogue (y is not written)
    unsigned char *aux;          // [R/4][J] first-max index inside the group (format of cg_bias_act_pool_fwd)
    long long R;
    int Fa, J, K, Q, Qp;         // Q = K*Fa, Qp = Q rounded up to 16
    uint32_t ring_bytes, piece, stage_bytes, a_plane, b_plane, off_ring, off_stage, off_b, off_ep, off_bar;
};

__global__ void __launch_bounds__(XT, 1) k_contract_umma(const CtParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *full = bars;           // [2] ring slot landed
    uint64_t *mbar = bars + 2;       // [2] MMAs of chunk c completed (barrier c & 1)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Fa = p.Fa, J = p.J, Q = p.Q, Qp = p.Qp;
    const long long nchunks_all = (p.R + ROWS - 1) / ROWS;
    // chunks c = blockIdx.x, blockIdx.x + gridDim.x, ...
    const int nloc = (int)((nchunks_all - blockIdx.x + gridDim.x - 1) / gridDim.x);

    const uint32_t ring0 = umma::smem_u32(smem + p.off_ring);
    const uint32_t st0 = umma::smem_u32(smem + p.off_stage);
    const uint32_t b0 = umma::smem_u32(smem + p.off_b);
    // MN-major canonical layouts: octet of 8 rows (A) / 8 output features (B) = 16 bytes; the 8 q of a k group
    // are 16 bytes apart (128-byte aligned core matrix), k groups 128 bytes apart, row octets SBO apart
    const uint32_t sbo = (uint32_t)(Qp / 8) * 128u;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) umma::mbar_init(bars + i, 1);
        umma::fence_mbar_init();
    }
    const int tmem_cols = 4 * J <= 32 ? 32 : (4 * J <= 64 ? 64 : (4 * J <= 128 ? 128 : (4 * J <= 256 ? 256 : 512)));
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
    // ---- W -> MN-major B operand, hi | mid planes, zero rows for q >= Q
    for (int e = tid; e < Qp * (J / 8); e += XT) {
        const int q = e / (J / 8), ob = e - q * (J / 8);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (q < Q) {
            const int k = q / Fa, f = q - k * Fa;
            const float *src = p.W + ((size_t)f * p.K + k) * J + ob * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = src[i];
        }
        uint2 h0, m0, h1, m1;
        split4(make_float4(v[0], v[1], v[2], v[3]), h0, m0);
        split4(make_float4(v[4], v[5], v[6], v[7]), h1, m1);
        const uint32_t off = (uint32_t)ob * sbo + (uint32_t)(q >> 3) * 128u + (uint32_t)(q & 7) * 16u;
        *reinterpret_cast<uint4 *>(smem + p.off_b + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4 *>(smem + p.off_b + p.b_plane + off) = make_uint4(m0.x, m0.y, m1.x, m1.y);
    }
    {   // q rows Q..Qp of the A planes are never written: clear both stages once
        uint4 *z = reinterpret_cast<uint4 *>(smem + p.off_stage);
        for (int i = tid; i < (int)(2 * p.stage_bytes / 16); i += XT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    auto chunk_rows = [&](int c, long long &rb) {
        rb = ((long long)blockIdx.x + (long long)c * gridDim.x) * ROWS;
        return (int)min((long long)ROWS, p.R - rb);
    };
    auto issue_loads = [&](int c) {             // all lanes of the producer warp
        long long rb;
        const int rows = chunk_rows(c, rb);
        const uint32_t dst = ring0 + (uint32_t)(c & 1) * p.ring_bytes;
        uint64_t *bar = full + (c & 1);
        if (lane == 0) {
            umma::fence_proxy_async();
            mbar_expect_tx(bar, (uint32_t)rows * (uint32_t)Q * 4u);
        }
        __syncwarp();
        for (int k = lane; k < p.K; k += 32)
            bulk_g2s(dst + (uint32_t)k * p.piece, p.stack + ((long long)k * p.R + rb) * Fa, (uint32_t)rows * Fa * 4u, bar);
    };

    if (warp == XC / 32 + 1) {
        // =========================== TMA producer warp ===================================
        if (nloc > 0) issue_loads(0);
        if (nloc > 1) issue_loads(1);
        for (int c = 0; c < nloc; ++c) {
            __syncthreads();                              // ring slot c&1 has been converted
            if (c + 2 < nloc) issue_loads(c + 2);
        }
    } else if (warp == XC / 32) {
        // =========================== MMA issue warp ======================================
        const uint32_t idesc = umma::make_idesc_bf16(128, J, 1, 1);
        const uint32_t d_hi = umma::desc_hi(sbo);
        for (int c = 0; c < nloc; ++c) {
            __syncthreads();                              // operands of chunk c are staged
            if (umma::elect_one()) {
                umma::fence_after_sync();
                const uint32_t sb = st0 + (uint32_t)(c & 1) * p.stage_bytes;
                const uint32_t a_lo = umma::desc_lo(sb, 128u), b_lo = umma::desc_lo(b0, 128u);
                const uint32_t a_mid = p.a_plane >> 4, b_mid = p.b_plane >> 4;
                const uint32_t t_step = (16u * sbo) >> 4;      // 128 rows = 16 row octets
                for (int t = 0; t < ROWS / 128; ++t) {
                    const uint32_t acc = tmem + (uint32_t)((c & 1) * (ROWS / 128) * J + t * J);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        if (pass >= p.npass) break;
                        uint32_t al = a_lo + (uint32_t)t * t_step + (pass == 1 ? a_mid : 0u);
                        uint32_t bl = b_lo + (pass == 2 ? b_mid : 0u);
                        for (int j = 0; j < Qp / 16; ++j) {
                            umma::mma_bf16(acc, umma::desc_join(al, d_hi), umma::desc_join(bl, d_hi), idesc, (pass | j) != 0);
                            al += 16u;                      // two k groups of 128 bytes
                            bl += 16u;
                        }
                    }
                }
                umma::commit(mbar + (c & 1));
            }
            __syncwarp();
        }
    } else {
        // =========================== compute warps ======================================
        const int qd = warp & 3, wq = warp >> 2;
        // chunk c: TMEM -> y (thread = row); warps of a lane quadrant split (tile, 8-column group) pairs
        auto epilogue = [&](int c) {
            long long rb;
            const int rows = chunk_rows(c, rb);
            umma::mbar_wait(mbar + (c & 1), (uint32_t)((c >> 1) & 1));
            umma::fence_after_sync();
            const int nc8 = J / 8;
            for (int idx = wq; idx < (ROWS / 128) * nc8; idx += 4) {
                const int t = idx / nc8, c8 = idx - t * nc8;
                const int r = t * 128 + 32 * qd + lane;
                float v[8];
                umma::tmem_ld8(tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)((c & 1) * (ROWS / 128) * J + t * J + c8 * 8), v);
                umma::tmem_ld_wait();
                if (p.yp != nullptr) {
                    // 32 rows x 8 columns of this warp -> [column][row] in a padded tile (36 floats per column: the
                    // 128-bit reads below hit 8 different bank groups), then one thread per (group of 4 rows, column)
                    float *tile = reinterpret_cast<float *>(smem + p.off_ep) + (size_t)warp * (8 * 36);
#pragma unroll
                    for (int j = 0; j < 8; ++j) tile[j * 36 + lane] = v[j];
                    __syncwarp();
                    const int colj = lane & 7;
                    const float b = p.bias != nullptr ? p.bias[c8 * 8 + colj] : 0.f;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int g = (lane >> 3) + 4 * h;
                        const float4 y4 = *reinterpret_cast<const float4 *>(tile + colj * 36 + 4 * g);
                        // same order as k_bias_act_pool_fwd: a = relu(y + b), the first maximum wins
                        const float a0 = fmaxf(y4.x + b, 0.f), a1 = fmaxf(y4.y + b, 0.f), a2 = fmaxf(y4.z + b, 0.f),
                                    a3 = fmaxf(y4.w + b, 0.f);
                        float best = a0;
                        int arg = 0;
                        if (a1 > best) { best = a1; arg = 1; }
                        if (a2 > best) { best = a2; arg = 2; }
                        if (a3 > best) { best = a3; arg = 3; }
                        const int row0 = t * 128 + 32 * qd + 4 * g;          // rows % 4 == 0: whole groups only
                        if (row0 < rows) {
                            const size_t o = (size_t)((rb + row0) >> 2) * J + c8 * 8 + colj;
                            p.yp[o] = best;
                            p.aux[o] = (unsigned char)arg;
                        }
                    }
                    __syncwarp();
                    continue;
                }
                if (r < rows) {
                    float *dst = p.y + (rb + r) * J + c8 * 8;
                    *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            umma::fence_before_sync();
        };
        for (int c = 0; c < nloc; ++c) {
            long long rb;
            const int rows = chunk_rows(c, rb);
            const unsigned char *ring = smem + p.off_ring + (size_t)(c & 1) * p.ring_bytes;
            unsigned char *stg = smem + p.off_stage + (size_t)(c & 1) * p.stage_bytes;
            umma::mbar_wait(full + (c & 1), (uint32_t)((c >> 1) & 1));
            // stage c&1 and accumulator c&1 were last used by chunk c-2, whose epilogue ran in iteration c-1
            // ---- A: for every q the rows of the chunk in octets of 8 -> one 16-byte store
            // blocks of 8 q x 8 row octets dealt diagonally: a quarter-warp stores 8 different q of the same k
            // group (the 128 contiguous bytes of a core matrix) and reads 8 different ring rows
            const int nq8 = (Q + 7) / 8;
            for (int e = tid; e < nq8 * 8 * (ROWS / 8); e += XC) {
                const int blk = e >> 6, b = e & 63;
                const int i = b & 7, ph = b >> 3;
                const int qg = blk % nq8, og = blk / nq8;
                const int q = qg * 8 + i, orow = og * 8 + ((i + ph) & 7);
                if (q >= Q) continue;
                const int k = q / Fa, f = q - k * Fa;
                const float *src = reinterpret_cast<const float *>(ring + (size_t)k * p.piece) + (size_t)orow * 8 * Fa + f;
                float v[8];
                if (Fa == 1) {
                    const float4 v0 = *reinterpret_cast<const float4 *>(src), v1 = *reinterpret_cast<const float4 *>(src + 4);
                    v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w;
                    v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = src[(size_t)i * Fa];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (orow * 8 + i >= rows) v[i] = 0.f;          // rows past the end of the tensor
                uint2 h0, m0, h1, m1;
                split4(make_float4(v[0], v[1], v[2], v[3]), h0, m0);
                split4(make_float4(v[4], v[5], v[6], v[7]), h1, m1);
                const uint32_t off = (uint32_t)orow * sbo + (uint32_t)(q >> 3) * 128u + (uint32_t)(q & 7) * 16u;
                *reinterpret_cast<uint4 *>(stg + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<uint4 *>(stg + p.a_plane + off) = make_uint4(m0.x, m0.y, m1.x, m1.y);
            }
            umma::fence_proxy_async();
            __syncthreads();
            if (c > 0) epilogue(c - 1);                   // under the MMAs of chunk c
        }
        if (nloc > 0) epilogue(nloc - 1);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

struct CtPlan {
    bool ok = false;
    CtParams cp;
    size_t smem = 0;
};

static CtPlan ct_plan(long long R, int Fa, int J, int K, size_t smem_limit) {
    CtPlan pl;
    if (J % 16 != 0 || J < 16 || J > 128 || Fa < 1 || K < 1 || R < 1) return pl;     // 2 * (ROWS/128) * J <= 512
    const int Q = K * Fa, Qp = (Q + 15) / 16 * 16;
    if ((R * Fa) % 4 != 0 || (ROWS * Fa) % 4 != 0) return pl;       // 16-byte aligned pieces
    CtParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.npass = cg_mma_passes();
    const uint32_t sbo = (uint32_t)(Qp / 8) * 128u;
    uint32_t off = 0;
    cp.off_bar = off;
    off += 128;
    cp.off_ring = off;
    cp.piece = (uint32_t)ROWS * Fa * 4u;
    cp.ring_bytes = (uint32_t)cg_align_up((size_t)K * cp.piece, 128);
    off += 2 * cp.ring_bytes;
    cp.off_stage = off;
    cp.a_plane = (uint32_t)(ROWS / 8) * sbo;
    cp.stage_bytes = 2 * cp.a_plane;
    off += 2 * cp.stage_bytes;
    cp.off_b = off;
    cp.b_plane = (uint32_t)(J / 8) * sbo;
    off += 2 * cp.b_plane;
    cp.off_ep = off;
    off += (uint32_t)(XC / 32) * 8 * 36 * 4;         // pooled epilogue: one [8][36] fp32 tile per compute warp
    if (off > smem_limit) return pl;
    cp.Q = Q;
    cp.Qp = Qp;
    pl.ok = true;
    pl.cp = cp;
    pl.smem = off;
    return pl;
}

}  // namespace

bool cg_contract_umma_supported(int N, int M, int Fa, int J, int K, size_t smem_limit) {
    return ct_plan((long long)N * M, Fa, J, K, smem_limit).ok;
}

// stack [K][N*M][Fa] sample-major, W [Fa*K][J] (row = f*K + k), y [N*M][J]
//   yp != NULL: pooled epilogue -- relu(y + bias) max-pooled over groups of 4 rows into yp / aux [N*M/4][J]; y unused
int cg_run_contract_umma(const float *stack, const float *W, float *y, int N, int M, int Fa, int J, int K, int sm_count,
                         size_t smem_limit, cudaStream_t s, const float *bias, float *yp, unsigned char *aux) {
    CtPlan pl = ct_plan((long long)N * M, Fa, J, K, smem_limit);
    CG_REQUIRE(pl.ok, "cg_run_contract_umma: shape not supported (Fa=%d J=%d K=%d)", Fa, J, K);
    CG_REQUIRE((((uintptr_t)stack | (uintptr_t)y) & 15) == 0, "cg_run_contract_umma: unaligned tensor");
    CG_REQUIRE(yp == nullptr || (aux != nullptr && M % 4 == 0), "cg_run_contract_umma: pooled epilogue needs aux and M %% 4 == 0");
    CtParams &cp = pl.cp;
    cp.stack = stack;
    cp.W = W;
    cp.y = y;
    cp.bias = bias;
    cp.yp = yp;
    cp.aux = aux;
    cp.R = (long long)N * M;
    cp.Fa = Fa;
    cp.J = J;
    cp.K = K;
    const long long nchunks = cg_ceil_div(cp.R, ROWS);
    CG_CHECK_CUDA(cudaFuncSetAttribute(k_contract_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    CgProfScope prof("contract_umma", s);
    k_contract_umma<<<(unsigned)std::min<long long>(nchunks, sm_count), XT, pl.smem, s>>>(cp);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
