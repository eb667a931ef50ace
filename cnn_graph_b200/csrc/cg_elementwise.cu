// Bandwidth-bound helpers around the filter: layout permutes, bias + activation
// (lib/models.py:226-247), permuted pooling (lib/models.py:249-266), perm_data
// (lib/coarsening.py:219-240) and the gconv-LSTM gate math (lib/gconv_lstm.py:185-215).
// All are single-pass, coalesced along the innermost (feature) axis, 128-bit where
// the feature count allows.
#include "cg_common.cuh"

static inline unsigned grid_for(int64_t work, int threads, int64_t cap = 148LL * 32) {
    int64_t b = cg_ceil_div(work, threads);
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (unsigned)b;
}

// ---------------------------------------------------------------------------
// in[A][B][F] -> out[B][A][F]
// ---------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
k_permute_direct(const float *__restrict__ in, float *__restrict__ out, int64_t A, int64_t B, int Fv) {
    // Fv = F / VEC vectors per chunk
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    const int64_t total = A * B * Fv;
    const V *src = reinterpret_cast<const V *>(in);
    V *dst = reinterpret_cast<V *>(out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i % Fv;
        const int64_t ba = i / Fv;
        const int64_t a = ba % A, b = ba / A;
        dst[i] = src[(a * B + b) * Fv + f];
    }
}

// small F: 32 x 32 tile of F-float elements through shared memory
__global__ void __launch_bounds__(256)
k_permute_tiled(const float *__restrict__ in, float *__restrict__ out, int64_t A, int64_t B, int F) {
    extern __shared__ float tile[];            // [32][32 * F + 1]
    const int pitch = 32 * F + 1;
    const int64_t a0 = (int64_t)blockIdx.y * 32, b0 = (int64_t)blockIdx.x * 32;
    const int nb = (int)min((int64_t)32, B - b0), na = (int)min((int64_t)32, A - a0);
    const int run_in = nb * F;                  // contiguous floats per a-row of the tile
    for (int r = threadIdx.y; r < na; r += blockDim.y)
        for (int i = threadIdx.x; i < run_in; i += blockDim.x)
            tile[r * pitch + i] = in[((a0 + r) * B + b0) * F + i];
    __syncthreads();
    const int run_out = na * F;
    for (int r = threadIdx.y; r < nb; r += blockDim.y)
        for (int i = threadIdx.x; i < run_out; i += blockDim.x) {
            const int a = i / F, f = i - a * F;
            out[((b0 + r) * A + a0) * F + i] = tile[a * pitch + r * F + f];
        }
}

int cg_run_permute_abf(const float *in, float *out, int64_t A, int64_t B, int F, cudaStream_t s) {
    if (A == 1 || B == 1) {
        CG_CHECK_CUDA(cudaMemcpyAsync(out, in, sizeof(float) * (size_t)(A * B * F), cudaMemcpyDeviceToDevice, s));
        return CG_OK;
    }
    CgProfScope prof("permute", s);
    if (F >= 8) {
        const bool v4 = F % 4 == 0 && ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
        if (v4)
            k_permute_direct<4><<<grid_for(A * B * (F / 4), 256), 256, 0, s>>>(in, out, A, B, F / 4);
        else
            k_permute_direct<1><<<grid_for(A * B * F, 256), 256, 0, s>>>(in, out, A, B, F);
    } else {
        dim3 grid((unsigned)cg_ceil_div(B, 32), (unsigned)cg_ceil_div(A, 32));
        CG_REQUIRE(grid.y <= 65535, "permute: A too large (%lld)", (long long)A);
        const size_t smem = sizeof(float) * 32 * (32 * F + 1);
        k_permute_tiled<<<grid, dim3(32, 8), smem, s>>>(in, out, A, B, F);
    }
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------
// bias + activation
// ---------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return tanhf(v);
    return v;
}
__device__ __forceinline__ float act_bwd_from_out(float y, float g, int act) {
    if (act == 1) return y > 0.f ? g : 0.f;
    if (act == 2) return g * (1.f - y * y);
    return g;
}

// period = number of bias entries (F for kind 1, M*F for kind 2, 0 for none)
__global__ void __launch_bounds__(256)
k_bias_act_fwd(const float *__restrict__ x, const float *__restrict__ bias, float *__restrict__ y,
               int64_t total, int64_t period, int act) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float v = x[i];
        if (period > 0) v += bias[i % period];
        y[i] = act_fwd(v, act);
    }
}

__global__ void __launch_bounds__(256)
k_bias_act_fwd_v4(const float4 *__restrict__ x, const float4 *__restrict__ bias, float4 *__restrict__ y,
                  int64_t total4, int64_t period4, int act) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = x[i];
        if (period4 > 0) {
            const float4 b = bias[i % period4];
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        y[i] = make_float4(act_fwd(v.x, act), act_fwd(v.y, act), act_fwd(v.z, act), act_fwd(v.w, act));
    }
}

extern "C" int cg_bias_act_fwd(const float *x, const float *bias, float *y, int N, int M, int F, int bias_kind,
                               int act, void *stream) {
    CG_REQUIRE(x && y, "cg_bias_act_fwd: NULL tensor");
    CG_REQUIRE(bias_kind >= 0 && bias_kind <= 2 && act >= 0 && act <= 2, "cg_bias_act_fwd: bad bias_kind/act");
    CG_REQUIRE(bias_kind == 0 || bias, "cg_bias_act_fwd: bias is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total = (int64_t)N * M * F;
    if (total == 0) return CG_OK;
    const int64_t period = bias_kind == 0 ? 0 : (bias_kind == 1 ? F : (int64_t)M * F);
    const bool v4 = F % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)bias)) & 15) == 0;
    CgProfScope prof("bias_act_fwd", s);
    if (v4)
        k_bias_act_fwd_v4<<<grid_for(total / 4, 256), 256, 0, s>>>((const float4 *)x, (const float4 *)bias, (float4 *)y,
                                                                   total / 4, period / 4, act);
    else
        k_bias_act_fwd<<<grid_for(total, 256), 256, 0, s>>>(x, bias, y, total, period, act);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// gx = gy * act'(y); dbias[c] += sum over rows of gx[r][c], c in [0, period).
// The tensor is viewed as [R][period]; block = TR x TC threads over a column tile.
__global__ void __launch_bounds__(256)
k_bias_act_bwd(const float *__restrict__ y, const float *__restrict__ gy, float *__restrict__ gx,
               float *__restrict__ dbias, int64_t R, int64_t period, int act, int tc, int64_t rows_per_block) {
    __shared__ float red[256];
    const int tx = threadIdx.x % tc, ty = threadIdx.x / tc, tr = 256 / tc;
    const int64_t c = (int64_t)blockIdx.x * tc + tx;
    const int64_t r_beg = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r_end = min(R, r_beg + rows_per_block);
    float acc = 0.f;
    if (c < period) {
        for (int64_t r = r_beg + ty; r < r_end; r += tr) {
            const int64_t i = r * period + c;
            const float g = act_bwd_from_out(y[i], gy[i], act);
            gx[i] = g;
            acc += g;
        }
    }
    if (dbias == nullptr) return;
    red[threadIdx.x] = acc;
    __syncthreads();
    if (ty == 0 && c < period) {
        float t = 0.f;
        for (int q = 0; q < tr; ++q) t += red[q * tc + tx];
        atomicAdd(dbias + c, t);
    }
}

extern "C" int cg_bias_act_bwd(const float *y, const float *gy, float *gx, float *dbias, int N, int M, int F,
                               int bias_kind, int act, void *stream) {
    CG_REQUIRE(y && gy && gx, "cg_bias_act_bwd: NULL tensor");
    CG_REQUIRE(bias_kind >= 0 && bias_kind <= 2 && act >= 0 && act <= 2, "cg_bias_act_bwd: bad bias_kind/act");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total = (int64_t)N * M * F;
    if (total == 0) return CG_OK;
    if (bias_kind == 0) dbias = nullptr;
    // view as [R][period]; without bias reduce nothing and use a wide period for coalescing
    int64_t period = bias_kind == 2 ? (int64_t)M * F : F;
    int64_t R = total / period;
    if (dbias) CG_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)period, s));
    int tc = 256;
    while (tc > 1 && tc / 2 >= period) tc /= 2;
    const int64_t col_blocks = cg_ceil_div(period, tc);
    int64_t row_blocks = cg_ceil_div(148LL * 16, col_blocks);
    const int tr = 256 / tc;
    if (row_blocks > cg_ceil_div(R, tr)) row_blocks = cg_ceil_div(R, tr);
    if (row_blocks < 1) row_blocks = 1;
    if (row_blocks > 65535) row_blocks = 65535;
    const int64_t rows_per_block = cg_ceil_div(R, row_blocks);
    dim3 grid((unsigned)col_blocks, (unsigned)cg_ceil_div(R, rows_per_block));
    CgProfScope prof("bias_act_bwd", s);
    k_bias_act_bwd<<<grid, 256, 0, s>>>(y, gy, gx, dbias, R, period, act, tc, rows_per_block);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------
// pooling over p consecutive vertices
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pool_fwd(const float *__restrict__ x, float *__restrict__ y, uint8_t *__restrict__ amax, int64_t NJ, int F, int p,
           int kind) {
    // one thread per output element (nj, f); x viewed as [NJ][p][F]
    const int64_t total = NJ * F;
    const float inv = 1.0f / (float)p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t nj = i / F;
        const int f = (int)(i - nj * F);
        const float *src = x + nj * p * F + f;
        if (kind == 1) {
            float best = src[0];
            int arg = 0;
            for (int q = 1; q < p; ++q) {
                const float v = src[(int64_t)q * F];
                if (v > best) { best = v; arg = q; }
            }
            y[i] = best;
            if (amax) amax[i] = (uint8_t)arg;
        } else {
            float sum = 0.f;
            for (int q = 0; q < p; ++q) sum += src[(int64_t)q * F];
            y[i] = sum * inv;
        }
    }
}

__global__ void __launch_bounds__(256)
k_pool_bwd(const float *__restrict__ gy, const uint8_t *__restrict__ amax, float *__restrict__ gx, int64_t NJ, int F,
           int p, int kind) {
    const int64_t total = NJ * F;
    const float inv = 1.0f / (float)p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t nj = i / F;
        const int f = (int)(i - nj * F);
        float *dst = gx + nj * p * F + f;
        const float g = gy[i];
        if (kind == 1) {
            const int arg = amax[i];
            for (int q = 0; q < p; ++q) dst[(int64_t)q * F] = q == arg ? g : 0.f;
        } else {
            const float v = g * inv;
            for (int q = 0; q < p; ++q) dst[(int64_t)q * F] = v;
        }
    }
}

extern "C" int cg_pool_fwd(const float *x, float *y, uint8_t *amax, int N, int M, int F, int p, int kind,
                           void *stream) {
    CG_REQUIRE(x && y, "cg_pool_fwd: NULL tensor");
    CG_REQUIRE(kind == 1 || kind == 2, "cg_pool_fwd: kind must be 1 (max) or 2 (avg)");
    CG_REQUIRE(p >= 1 && p <= 256 && M % p == 0, "cg_pool_fwd: p=%d must divide M=%d and be <= 256", p, M);
    const int64_t NJ = (int64_t)N * (M / p);
    if (NJ * F == 0) return CG_OK;
    CgProfScope prof("pool_fwd", (cudaStream_t)stream);
    k_pool_fwd<<<grid_for(NJ * F, 256), 256, 0, (cudaStream_t)stream>>>(x, y, amax, NJ, F, p, kind);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

extern "C" int cg_pool_bwd(const float *gy, const uint8_t *amax, float *gx, int N, int M, int F, int p, int kind,
                           void *stream) {
    CG_REQUIRE(gy && gx, "cg_pool_bwd: NULL tensor");
    CG_REQUIRE(kind == 1 || kind == 2, "cg_pool_bwd: kind must be 1 (max) or 2 (avg)");
    CG_REQUIRE(kind == 2 || amax, "cg_pool_bwd: max pooling needs the argmax tensor");
    CG_REQUIRE(p >= 1 && p <= 256 && M % p == 0, "cg_pool_bwd: p=%d must divide M=%d and be <= 256", p, M);
    const int64_t NJ = (int64_t)N * (M / p);
    if (NJ * F == 0) return CG_OK;
    CgProfScope prof("pool_bwd", (cudaStream_t)stream);
    k_pool_bwd<<<grid_for(NJ * F, 256), 256, 0, (cudaStream_t)stream>>>(gy, amax, gx, NJ, F, p, kind);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------
// fused bias + activation + pooling  (the brelu -> pool tail of every cgcnn layer,
// lib/models.py:226-266 as sequenced by _inference): one pass over the filter output
// instead of three, one pass over its gradient instead of two.
//   max pooling : aux = index of the FIRST maximal activated value (TF MaxPoolGrad routing)
//   avg pooling : aux = bit q set when activated element q is > 0 (relu) -- the relu mask
// VEC features per thread (4 when F % 4 == 0).
// ---------------------------------------------------------------------------
// P > 0: pool size known at compile time (the q loops unroll and all loads of a thread are issued together)
template <int VEC, int P>
__global__ void __launch_bounds__(256)
k_bias_act_pool_fwd(const float *__restrict__ x, const float *__restrict__ bias, float *__restrict__ y,
                    uint8_t *__restrict__ aux, int64_t NJ, int Mp, int F, int p_rt, int bias_kind, int act, int kind) {
    const int p = P > 0 ? P : p_rt;
    const int FV = F / VEC;
    const int64_t total = NJ * FV;
    const float inv = 1.0f / (float)p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t nj = i / FV;
        const int f = (int)(i - nj * FV) * VEC;
        const int j = (int)(nj % Mp);                       // pooled vertex: bias kind 2 is indexed by (vertex, f)
        const float *src = x + nj * p * F + f;
        float best[VEC], sum[VEC];
        int arg[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) { best[e] = 0.f; sum[e] = 0.f; arg[e] = 0; }
#pragma unroll
        for (int q = 0; q < p; ++q) {
            float v[VEC];
            if (VEC == 4) {
                const float4 t = __ldcs(reinterpret_cast<const float4 *>(src + (int64_t)q * F));
                v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
            } else {
                v[0] = src[(int64_t)q * F];
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                float a = v[e];
                if (bias_kind == 1) a += bias[f + e];
                else if (bias_kind == 2) a += bias[((int64_t)j * p + q) * F + f + e];
                a = act_fwd(a, act);
                if (kind == 1) {
                    if (q == 0 || a > best[e]) { best[e] = a; arg[e] = q; }
                } else {
                    sum[e] += a;
                    if (a > 0.f) arg[e] |= 1 << q;
                }
            }
        }
        if (VEC == 4) {      // F % 4 == 0: 16-byte aligned pooled row piece, 4-byte aligned aux piece
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = kind == 1 ? best[e % VEC] : sum[e % VEC] * inv;
            *reinterpret_cast<float4 *>(y + nj * F + f) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uchar4 *>(aux + nj * F + f) =
                make_uchar4((uint8_t)arg[0], (uint8_t)arg[1 % VEC], (uint8_t)arg[2 % VEC], (uint8_t)arg[3 % VEC]);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                y[nj * F + f + e] = kind == 1 ? best[e] : sum[e] * inv;
                aux[nj * F + f + e] = (uint8_t)arg[e];
            }
        }
    }
}

// gx[n, j*p+q, f] from the pooled gradient; db (optional) accumulates the bias gradient per thread and is
// flushed with one atomicAdd per (thread, feature[, vertex]).
template <int VEC, int P>
__global__ void __launch_bounds__(256)
k_bias_act_pool_bwd(const float *__restrict__ gy, const float *__restrict__ yp, const uint8_t *__restrict__ aux,
                    float *__restrict__ gx, float *__restrict__ dbias, int64_t NJ, int Mp, int F, int p_rt, int bias_kind,
                    int act, int kind) {
    const int p = P > 0 ? P : p_rt;
    const int FV = F / VEC;
    const int64_t total = NJ * FV;
    const float inv = 1.0f / (float)p;
    // with a per-filter bias every thread keeps its feature columns: the grid stride is a multiple of FV
    float dacc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) dacc[e] = 0.f;
    int f_keep = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t nj = i / FV;
        const int f = (int)(i - nj * FV) * VEC;
        const int j = (int)(nj % Mp);
        f_keep = f;
        float *dst = gx + nj * p * F + f;
        float g[VEC];
        int a[VEC];
        float go[VEC], yo[VEC];
        if (VEC == 4) {
            const float4 g4 = __ldcs(reinterpret_cast<const float4 *>(gy + nj * F + f));
            const float4 y4 = __ldcs(reinterpret_cast<const float4 *>(yp + nj * F + f));
            const uchar4 a4 = *reinterpret_cast<const uchar4 *>(aux + nj * F + f);
            go[0] = g4.x; go[1 % VEC] = g4.y; go[2 % VEC] = g4.z; go[3 % VEC] = g4.w;
            yo[0] = y4.x; yo[1 % VEC] = y4.y; yo[2 % VEC] = y4.z; yo[3 % VEC] = y4.w;
            a[0] = a4.x; a[1 % VEC] = a4.y; a[2 % VEC] = a4.z; a[3 % VEC] = a4.w;
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                go[e] = gy[nj * F + f + e];
                yo[e] = yp[nj * F + f + e];
                a[e] = aux[nj * F + f + e];
            }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e)   // max: the routed element's activation derivative comes from the pooled value itself
            g[e] = kind == 1 ? act_bwd_from_out(yo[e], go[e], act) : go[e] * inv;
#pragma unroll
        for (int q = 0; q < p; ++q) {
            float o[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (kind == 1) o[e] = q == a[e] ? g[e] : 0.f;
                else o[e] = (act == 1) ? (((a[e] >> q) & 1) ? g[e] : 0.f) : g[e];
                if (bias_kind == 1) dacc[e] += o[e];
                else if (bias_kind == 2 && dbias != nullptr && o[e] != 0.f)
                    atomicAdd(dbias + ((int64_t)j * p + q) * F + f + e, o[e]);
            }
            if (VEC == 4) {
                __stcs(reinterpret_cast<float4 *>(dst + (int64_t)q * F), make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]));
            } else {
                dst[(int64_t)q * F] = o[0];
            }
        }
    }
    if (bias_kind == 1 && dbias != nullptr) {
        // block reduction per feature column (threads t, t + FV, ... of the block share their columns when
        // 256 % FV == 0; otherwise every thread flushes on its own), then one atomicAdd per block and column
        __shared__ float red[256 * VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) red[threadIdx.x * VEC + e] = f_keep >= 0 ? dacc[e] : 0.f;
        __syncthreads();
        if (256 % FV == 0) {
            if ((int)threadIdx.x < FV) {
                const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
                const int f = (int)(i0 % FV) * VEC;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    float t = 0.f;
                    for (int q = threadIdx.x; q < 256; q += FV) t += red[q * VEC + e];
                    atomicAdd(dbias + f + e, t);
                }
            }
        } else if (f_keep >= 0) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) atomicAdd(dbias + f_keep + e, dacc[e]);
        }
    }
}

static int check_fused_pool(const char *who, int M, int F, int p, int bias_kind, int act, int kind) {
    CG_REQUIRE(kind == 1 || kind == 2, "%s: kind must be 1 (max) or 2 (avg)", who);
    CG_REQUIRE(bias_kind >= 0 && bias_kind <= 2 && act >= 0 && act <= 2, "%s: bad bias_kind/act", who);
    CG_REQUIRE(p >= 1 && p <= 8 && M % p == 0, "%s: p=%d must divide M=%d and be <= 8", who, p, M);
    CG_REQUIRE(!(kind == 2 && act == 2), "%s: avg pooling after tanh is not fused (use cg_bias_act + cg_pool)", who);
    CG_REQUIRE(F > 0, "%s: F must be positive", who);
    return CG_OK;
}

// grid whose stride is a multiple of the per-row thread count FV (so a thread keeps its feature columns)
static unsigned fused_pool_grid(int64_t total, int FV) {
    int64_t blocks = cg_ceil_div(total, 256);
    const int64_t cap = 148LL * 16;
    if (blocks > cap) blocks = cap;
    // gridDim.x * 256 must be a multiple of FV: round the block count up to a multiple of FV / gcd(FV, 256)
    int64_t a = FV, b = 256;
    while (b) { const int64_t t = a % b; a = b; b = t; }
    const int64_t unit = FV / a;
    blocks = cg_ceil_div(blocks, unit) * unit;
    return (unsigned)blocks;
}

extern "C" int cg_bias_act_pool_fwd(const float *x, const float *bias, float *y, uint8_t *aux, int N, int M, int F,
                                    int p, int bias_kind, int act, int kind, void *stream) {
    int rc = check_fused_pool("cg_bias_act_pool_fwd", M, F, p, bias_kind, act, kind);
    if (rc != CG_OK) return rc;
    const int64_t NJ = (int64_t)N * (M / p);
    if (NJ == 0) return CG_OK;
    CG_REQUIRE(x && y && aux && (bias_kind == 0 || bias), "cg_bias_act_pool_fwd: NULL tensor");
    cudaStream_t s = (cudaStream_t)stream;
    const bool v4 = F % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0 && (((uintptr_t)aux) & 3) == 0;
    CgProfScope prof("bias_act_pool_fwd", s);
    if (v4 && p == 4)
        k_bias_act_pool_fwd<4, 4><<<grid_for(NJ * (F / 4), 256), 256, 0, s>>>(x, bias, y, aux, NJ, M / p, F, p, bias_kind, act, kind);
    else if (v4 && p == 2)
        k_bias_act_pool_fwd<4, 2><<<grid_for(NJ * (F / 4), 256), 256, 0, s>>>(x, bias, y, aux, NJ, M / p, F, p, bias_kind, act, kind);
    else if (v4)
        k_bias_act_pool_fwd<4, 0><<<grid_for(NJ * (F / 4), 256), 256, 0, s>>>(x, bias, y, aux, NJ, M / p, F, p, bias_kind, act, kind);
    else
        k_bias_act_pool_fwd<1, 0><<<grid_for(NJ * F, 256), 256, 0, s>>>(x, bias, y, aux, NJ, M / p, F, p, bias_kind, act, kind);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

extern "C" int cg_bias_act_pool_bwd(const float *gy, const float *y, const uint8_t *aux, float *gx, float *dbias,
                                    int N, int M, int F, int p, int bias_kind, int act, int kind, void *stream) {
    int rc = check_fused_pool("cg_bias_act_pool_bwd", M, F, p, bias_kind, act, kind);
    if (rc != CG_OK) return rc;
    const int64_t NJ = (int64_t)N * (M / p);
    cudaStream_t s = (cudaStream_t)stream;
    if (bias_kind == 0) dbias = nullptr;
    if (dbias) CG_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)(bias_kind == 1 ? F : (int64_t)M * F), s));
    if (NJ == 0) return CG_OK;
    CG_REQUIRE(gy && y && aux && gx, "cg_bias_act_pool_bwd: NULL tensor");
    const bool v4 = F % 4 == 0 && ((((uintptr_t)gx) | ((uintptr_t)gy) | ((uintptr_t)y)) & 15) == 0 && (((uintptr_t)aux) & 3) == 0;
    CgProfScope prof("bias_act_pool_bwd", s);
    if (v4 && p == 4)
        k_bias_act_pool_bwd<4, 4><<<fused_pool_grid(NJ * (F / 4), F / 4), 256, 0, s>>>(gy, y, aux, gx, dbias, NJ, M / p, F, p,
                                                                                       bias_kind, act, kind);
    else if (v4 && p == 2)
        k_bias_act_pool_bwd<4, 2><<<fused_pool_grid(NJ * (F / 4), F / 4), 256, 0, s>>>(gy, y, aux, gx, dbias, NJ, M / p, F, p,
                                                                                       bias_kind, act, kind);
    else if (v4)
        k_bias_act_pool_bwd<4, 0><<<fused_pool_grid(NJ * (F / 4), F / 4), 256, 0, s>>>(gy, y, aux, gx, dbias, NJ, M / p, F, p,
                                                                                       bias_kind, act, kind);
    else
        k_bias_act_pool_bwd<1, 0><<<fused_pool_grid(NJ * F, F), 256, 0, s>>>(gy, y, aux, gx, dbias, NJ, M / p, F, p,
                                                                             bias_kind, act, kind);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------
// perm_data
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_perm_data(const float *__restrict__ x, const int *__restrict__ perm, float *__restrict__ out, int64_t N, int M,
            int Mnew) {
    const int64_t total = N * Mnew;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / Mnew;
        const int j = perm[i - n * Mnew];
        out[i] = (j >= 0 && j < M) ? x[n * M + j] : 0.f;
    }
}

extern "C" int cg_perm_data(const float *x, const int32_t *perm, float *out, int64_t N, int M, int Mnew,
                            void *stream) {
    CG_REQUIRE(Mnew >= M && M > 0, "cg_perm_data: need Mnew >= M > 0 (M=%d Mnew=%d)", M, Mnew);
    if (N == 0) return CG_OK;
    CG_REQUIRE(x && perm && out, "cg_perm_data: NULL tensor");
    CgProfScope prof("perm_data", (cudaStream_t)stream);
    k_perm_data<<<grid_for(N * Mnew, 256), 256, 0, (cudaStream_t)stream>>>(x, perm, out, N, M, Mnew);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------
// gconv-LSTM gates
// ---------------------------------------------------------------------------
// Gate nonlinearities on the special-function unit (ex2 + fast reciprocal: absolute error ~1e-7, two orders below the
// parity tolerance); the precise expf / tanhf / division sequences made the gate kernels instruction bound (~200
// instructions per element).  tan (the fork's z gate) stays the precise tanf: it is ill-conditioned near its poles.
__device__ __forceinline__ float sigmoidf_(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float tanhf_(float v) { return 1.0f - __fdividef(2.0f, __expf(2.0f * v) + 1.0f); }

__global__ void __launch_bounds__(256)
k_lstm_gates_fwd(const float *__restrict__ pre, const float *__restrict__ pre2, const float *__restrict__ bias,
                 const float *__restrict__ c, float *__restrict__ new_c, float *__restrict__ new_h, int64_t R, int H, int variant) {
    const int64_t total = R * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / H;
        const int h = (int)(i - r * H);
        const float *p = pre + r * 4 * H;
        float az = p[h] + bias[h], ai = p[H + h] + bias[H + h];
        float af = p[2 * H + h] + bias[2 * H + h], ao = p[3 * H + h] + bias[3 * H + h];
        if (pre2 != nullptr) {      // x-path and h-path pre-activations summed here (lib/gconv_lstm.py:185-207), not by a separate pass
            const float *q = pre2 + r * 4 * H;
            az += q[h];
            ai += q[H + h];
            af += q[2 * H + h];
            ao += q[3 * H + h];
        }
        const float z = variant == 0 ? tanf(az) : tanhf_(az);
        const float o = variant == 0 ? tanhf_(ao) : sigmoidf_(ao);
        const float ig = sigmoidf_(ai), fg = sigmoidf_(af);
        const float cn = fg * c[i] + ig * z;
        new_c[i] = cn;
        new_h[i] = o * tanhf_(cn);
    }
}

__global__ void __launch_bounds__(256)
k_lstm_gates_bwd(const float *__restrict__ pre, const float *__restrict__ pre2, const float *__restrict__ bias,
                 const float *__restrict__ c, const float *__restrict__ new_c, const float *__restrict__ g_h, const float *__restrict__ g_c,
                 float *__restrict__ g_pre, float *__restrict__ g_cprev, float *__restrict__ d_bias, int64_t R, int H,
                 int variant, int64_t rows_per_block) {
    // block covers rows [blockIdx.y * rows_per_block, ...) x columns h = blockIdx.x * 256 + tid
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    const int64_t r_beg = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r_end = min(R, r_beg + rows_per_block);
    float sz = 0.f, si = 0.f, sf = 0.f, so = 0.f;
    const float bz = bias[h], bi = bias[H + h], bf = bias[2 * H + h], bo = bias[3 * H + h];
    for (int64_t r = r_beg; r < r_end; ++r) {
        const int64_t i = r * H + h;
        const float *p = pre + r * 4 * H;
        float az = p[h] + bz, ai = p[H + h] + bi, af = p[2 * H + h] + bf, ao = p[3 * H + h] + bo;
        if (pre2 != nullptr) {
            const float *q = pre2 + r * 4 * H;
            az += q[h];
            ai += q[H + h];
            af += q[2 * H + h];
            ao += q[3 * H + h];
        }
        const float z = variant == 0 ? tanf(az) : tanhf_(az);
        const float o = variant == 0 ? tanhf_(ao) : sigmoidf_(ao);
        const float ig = sigmoidf_(ai), fg = sigmoidf_(af);
        const float tc = tanhf_(new_c[i]);
        const float gh = g_h ? g_h[i] : 0.f;
        const float dcn = (g_c ? g_c[i] : 0.f) + gh * o * (1.f - tc * tc);
        const float d_o = gh * tc;
        const float dz = dcn * ig, di = dcn * z, df = dcn * c[i];
        const float gz = dz * (variant == 0 ? (1.f + z * z) : (1.f - z * z));
        const float go = d_o * (variant == 0 ? (1.f - o * o) : o * (1.f - o));
        const float gi = di * ig * (1.f - ig), gf = df * fg * (1.f - fg);
        float *q = g_pre + r * 4 * H;
        q[h] = gz; q[H + h] = gi; q[2 * H + h] = gf; q[3 * H + h] = go;
        g_cprev[i] = dcn * fg;
        sz += gz; si += gi; sf += gf; so += go;
    }
    if (d_bias) {
        atomicAdd(d_bias + h, sz);
        atomicAdd(d_bias + H + h, si);
        atomicAdd(d_bias + 2 * H + h, sf);
        atomicAdd(d_bias + 3 * H + h, so);
    }
}

extern "C" int cg_lstm_gates2_fwd(const float *pre, const float *pre2, const float *bias, const float *c, float *new_c, float *new_h,
                                  int64_t R, int H, int variant, void *stream);
extern "C" int cg_lstm_gates_fwd(const float *pre, const float *bias, const float *c, float *new_c, float *new_h,
                                 int64_t R, int H, int variant, void *stream) {
    return cg_lstm_gates2_fwd(pre, nullptr, bias, c, new_c, new_h, R, H, variant, stream);
}

extern "C" int cg_lstm_gates2_fwd(const float *pre, const float *pre2, const float *bias, const float *c, float *new_c, float *new_h,
                                  int64_t R, int H, int variant, void *stream) {
    CG_REQUIRE(pre && bias && c && new_c && new_h, "cg_lstm_gates_fwd: NULL tensor");
    CG_REQUIRE(variant == 0 || variant == 1, "cg_lstm_gates_fwd: variant must be 0 (fork) or 1 (standard)");
    if (R * H == 0) return CG_OK;
    CgProfScope prof("lstm_gates_fwd", (cudaStream_t)stream);
    k_lstm_gates_fwd<<<grid_for(R * H, 256), 256, 0, (cudaStream_t)stream>>>(pre, pre2, bias, c, new_c, new_h, R, H, variant);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

extern "C" int cg_lstm_gates2_bwd(const float *pre, const float *pre2, const float *bias, const float *c, const float *new_c,
                                  const float *g_h, const float *g_c, float *g_pre, float *g_cprev, float *d_bias, int64_t R, int H,
                                  int variant, void *stream);
extern "C" int cg_lstm_gates_bwd(const float *pre, const float *bias, const float *c, const float *new_c,
                                 const float *g_h, const float *g_c, float *g_pre, float *g_cprev, float *d_bias,
                                 int64_t R, int H, int variant, void *stream) {
    return cg_lstm_gates2_bwd(pre, nullptr, bias, c, new_c, g_h, g_c, g_pre, g_cprev, d_bias, R, H, variant, stream);
}

extern "C" int cg_lstm_gates2_bwd(const float *pre, const float *pre2, const float *bias, const float *c, const float *new_c,
                                  const float *g_h, const float *g_c, float *g_pre, float *g_cprev, float *d_bias, int64_t R, int H,
                                  int variant, void *stream) {
    CG_REQUIRE(pre && bias && c && new_c && g_pre && g_cprev, "cg_lstm_gates_bwd: NULL tensor");
    CG_REQUIRE(variant == 0 || variant == 1, "cg_lstm_gates_bwd: variant must be 0 (fork) or 1 (standard)");
    if (R * H == 0) return CG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (d_bias) CG_CHECK_CUDA(cudaMemsetAsync(d_bias, 0, sizeof(float) * 4 * (size_t)H, s));
    const int threads = H >= 256 ? 256 : (H + 31) / 32 * 32;
    const int64_t col_blocks = cg_ceil_div(H, threads);
    int64_t row_blocks = cg_ceil_div(148LL * 16, col_blocks);
    if (row_blocks > R) row_blocks = R;
    if (row_blocks > 65535) row_blocks = 65535;
    const int64_t rows_per_block = cg_ceil_div(R, row_blocks);
    dim3 grid((unsigned)col_blocks, (unsigned)cg_ceil_div(R, rows_per_block));
    CgProfScope prof("lstm_gates_bwd", s);
    k_lstm_gates_bwd<<<grid, threads, 0, s>>>(pre, pre2, bias, c, new_c, g_h, g_c, g_pre, g_cprev, d_bias, R, H, variant,
                                          rows_per_block);
    CG_LAUNCH_CHECK();
    return CG_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Loss + optimiser tail of cgcnn (lib/graph_model.py:246-310): softmax cross-entropy with its
// gradient in one launch, momentum SGD over every variable in one launch.
// ---------------------------------------------------------------------------------------------------------------
// loss = mean_n ( logsumexp(z_n) - z_n[y_n] ),  dz[n, c] = (softmax(z_n)[c] - [c == y_n]) / N.  One block: a thread owns
// whole rows (C is the class count: 3 .. 20 here), partial losses are summed in a fixed order (deterministic).
__global__ void __launch_bounds__(1024) k_softmax_xent(const float *__restrict__ z, const long long *__restrict__ y,
                                                       float *__restrict__ loss, float *__restrict__ dz, int N, int C) {
    __shared__ float red[32];
    float part = 0.f;
    const float invN = 1.f / (float)N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float *row = z + (size_t)n * C;
        float mx = row[0];
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, row[c]);
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
        const int label = (int)y[n];
        const float lse = logf(sum) + mx;
        if (label >= 0 && label < C) part += lse - row[label];
        const float inv = 1.f / sum;
        float *drow = dz + (size_t)n * C;
        for (int c = 0; c < C; ++c) drow[c] = (expf(row[c] - mx) * inv - (c == label ? 1.f : 0.f)) * invN;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) *loss = v * invN;
    }
}

extern "C" int cg_softmax_xent(const float *dev_logits, const long long *dev_labels, float *dev_loss, float *dev_dlogits, int N,
                               int C, void *stream) {
    CG_REQUIRE(dev_logits && dev_labels && dev_loss && dev_dlogits, "cg_softmax_xent: NULL tensor");
    CG_REQUIRE(N > 0 && C > 0 && C <= 4096, "cg_softmax_xent: bad shape N=%d C=%d", N, C);
    CgProfScope prof("softmax_xent", (cudaStream_t)stream);
    k_softmax_xent<<<1, 1024, 0, (cudaStream_t)stream>>>(dev_logits, dev_labels, dev_loss, dev_dlogits, N, C);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// table[t] = {param, grad, momentum buffer, element count}; buf = momentum * buf + grad; param -= lr * buf
// (torch.optim.SGD / tf.train.MomentumOptimizer without Nesterov; a zero buffer makes the first step buf = grad).
struct CgSgdEntry {
    float *p;
    const float *g;
    float *buf;
    long long n;
};

constexpr int CG_SGD_MAX = 64;      // records per launch: they travel as kernel arguments (2 KB), so a captured CUDA graph
struct CgSgdTable {                 // replays them without any device-side table to keep up to date
    CgSgdEntry e[CG_SGD_MAX];
};

__global__ void __launch_bounds__(256) k_sgd_momentum(const __grid_constant__ CgSgdTable table, float lr, float momentum,
                                                       const float *__restrict__ lr_dev) {
    if (lr_dev != nullptr) lr = *lr_dev;      // learning rate read at run time (a replayed CUDA graph bakes kernel arguments)
    const CgSgdEntry e = table.e[blockIdx.y];
    const long long n4 = ((((uintptr_t)e.p | (uintptr_t)e.g | (uintptr_t)e.buf) & 15) == 0) ? e.n / 4 : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 g = reinterpret_cast<const float4 *>(e.g)[i];
        float4 b = reinterpret_cast<float4 *>(e.buf)[i], w = reinterpret_cast<float4 *>(e.p)[i];
        b.x = fmaf(momentum, b.x, g.x);
        b.y = fmaf(momentum, b.y, g.y);
        b.z = fmaf(momentum, b.z, g.z);
        b.w = fmaf(momentum, b.w, g.w);
        w.x = fmaf(-lr, b.x, w.x);
        w.y = fmaf(-lr, b.y, w.y);
        w.z = fmaf(-lr, b.z, w.z);
        w.w = fmaf(-lr, b.w, w.w);
        reinterpret_cast<float4 *>(e.buf)[i] = b;
        reinterpret_cast<float4 *>(e.p)[i] = w;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += stride) {
        const float b = fmaf(momentum, e.buf[i], e.g[i]);
        e.buf[i] = b;
        e.p[i] = fmaf(-lr, b, e.p[i]);
    }
}

extern "C" int cg_sgd_momentum_dev(const void *host_table, int ntensors, long long max_numel, float lr, const float *dev_lr,
                                   float momentum, void *stream);
extern "C" int cg_sgd_momentum(const void *host_table, int ntensors, long long max_numel, float lr, float momentum, void *stream) {
    return cg_sgd_momentum_dev(host_table, ntensors, max_numel, lr, nullptr, momentum, stream);
}

extern "C" int cg_sgd_momentum_dev(const void *host_table, int ntensors, long long max_numel, float lr, const float *dev_lr,
                                   float momentum, void *stream) {
    CG_REQUIRE(host_table != nullptr && ntensors > 0, "cg_sgd_momentum: bad table (%d tensors)", ntensors);
    long long blocks = cg_ceil_div(cg_ceil_div(max_numel, 4), 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 592) blocks = 592;       // 4 x 148: the large fc weight streams at full rate, small tensors take one block
    const CgSgdEntry *src = reinterpret_cast<const CgSgdEntry *>(host_table);
    for (int t0 = 0; t0 < ntensors; t0 += CG_SGD_MAX) {
        const int n = ntensors - t0 < CG_SGD_MAX ? ntensors - t0 : CG_SGD_MAX;
        CgSgdTable tab;
        memset(&tab, 0, sizeof(tab));
        for (int i = 0; i < n; ++i) {
            tab.e[i] = src[t0 + i];
            CG_REQUIRE(tab.e[i].p && tab.e[i].g && tab.e[i].buf && tab.e[i].n >= 0, "cg_sgd_momentum: NULL tensor in record %d", t0 + i);
        }
        CgProfScope prof("sgd_momentum", (cudaStream_t)stream);
        k_sgd_momentum<<<dim3((unsigned)blocks, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(tab, lr, momentum, dev_lr);
        CG_LAUNCH_CHECK();
    }
    return CG_OK;
}


// Adam (lib/graph_model.py:293 tf.train.AdamOptimizer; here with torch.optim.Adam's arithmetic, which the host model used
// before): every variable in ONE launch.  table[t] = {param, grad, exp_avg, exp_avg_sq, element count}.  The step count
// lives on the device (state[0]; a replayed CUDA graph bakes kernel arguments): every block reads it when it starts, the
// block that finishes last (ticket in state[1]) advances it -- no block can still be waiting to read by then.
struct CgAdamEntry {
    float *p;
    const float *g;
    float *m, *v;
    long long n;
};
constexpr int CG_ADAM_MAX = 48;
struct CgAdamTable {
    CgAdamEntry e[CG_ADAM_MAX];
};

__global__ void __launch_bounds__(256) k_adam(const __grid_constant__ CgAdamTable table, float lr, float beta1, float beta2, float eps,
                                               int *__restrict__ state, int bump) {
    __shared__ float coef[2];
    if (threadIdx.x == 0) {
        const int t = *reinterpret_cast<volatile int *>(state) + 1;
        coef[0] = (float)((double)lr / (1.0 - pow((double)beta1, (double)t)));      // step size
        coef[1] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)t)));         // 1 / sqrt(bias correction 2)
    }
    __syncthreads();
    const float step_size = coef[0], inv_bc2 = coef[1];
    const CgAdamEntry e = table.e[blockIdx.y];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += stride) {
        const float g = e.g[i];
        float m = e.m[i], v = e.v[i];
        m = fmaf(1.f - beta1, g - m, m);                        // exp_avg.lerp_(grad, 1 - beta1)
        v = fmaf(1.f - beta2, g * g, beta2 * v);                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        e.m[i] = m;
        e.v[i] = v;
        e.p[i] -= step_size * (m / (sqrtf(v) * inv_bc2 + eps));
    }
    if (bump) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned total = gridDim.x * gridDim.y;
            const unsigned ticket = atomicAdd(reinterpret_cast<unsigned *>(state + 1), 1u);
            if (ticket == total - 1) {
                state[0] = state[0] + 1;
                state[1] = 0;
            }
        }
    }
}

extern "C" int cg_adam(const void *host_table, int ntensors, long long max_numel, float lr, float beta1, float beta2, float eps,
                       int *dev_state, void *stream) {
    CG_REQUIRE(host_table != nullptr && ntensors > 0 && dev_state != nullptr, "cg_adam: bad table (%d tensors) or NULL state", ntensors);
    long long blocks = cg_ceil_div(max_numel, 1024);
    if (blocks < 1) blocks = 1;
    if (blocks > 592) blocks = 592;
    const CgAdamEntry *src = reinterpret_cast<const CgAdamEntry *>(host_table);
    for (int t0 = 0; t0 < ntensors; t0 += CG_ADAM_MAX) {
        const int n = ntensors - t0 < CG_ADAM_MAX ? ntensors - t0 : CG_ADAM_MAX;
        CgAdamTable tab;
        memset(&tab, 0, sizeof(tab));
        for (int i = 0; i < n; ++i) {
            tab.e[i] = src[t0 + i];
            CG_REQUIRE(tab.e[i].p && tab.e[i].g && tab.e[i].m && tab.e[i].v && tab.e[i].n >= 0, "cg_adam: NULL tensor in record %d", t0 + i);
        }
        CgProfScope prof("adam", (cudaStream_t)stream);
        k_adam<<<dim3((unsigned)blocks, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(tab, lr, beta1, beta2, eps, dev_state,
                                                                                      t0 + n >= ntensors ? 1 : 0);
        CG_LAUNCH_CHECK();
    }
    return CG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Sparse input batches (lib/graph_model.py:150-151 densifies scipy batches on the host and feeds
// the dense array): the CSR batch travels to the device as it is (indptr, indices, values: ~1 % of the dense bytes
// for bag-of-words rows) and is expanded here.  One block per row: zero the row, then scatter its entries; rows of
// the batch beyond the CSR (zero padding of a last batch) are cleared.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_csr_densify(const int *__restrict__ indptr, const int *__restrict__ indices,
                                                      const float *__restrict__ values, float *__restrict__ out, int rows, int M) {
    const int n = blockIdx.x;
    float *row = out + (size_t)n * M;
    const int m4 = ((((uintptr_t)row) & 15) == 0) ? M / 4 : 0;
    for (int i = threadIdx.x; i < m4; i += blockDim.x) reinterpret_cast<float4 *>(row)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 4 * m4 + threadIdx.x; i < M; i += blockDim.x) row[i] = 0.f;
    if (n >= rows) return;
    __syncthreads();
    const int beg = indptr[n], end = indptr[n + 1];
    for (int e = beg + threadIdx.x; e < end; e += blockDim.x) {
        const int c = indices[e];
        if (c >= 0 && c < M) atomicAdd(row + c, values[e]);       // duplicates are summed, as scipy's toarray() does
    }
}

extern "C" int cg_csr_densify(const int32_t *dev_indptr, const int32_t *dev_indices, const float *dev_values, float *dev_out,
                              int csr_rows, int out_rows, int M, void *stream) {
    CG_REQUIRE(dev_indptr && dev_out && out_rows >= csr_rows && csr_rows >= 0 && M > 0, "cg_csr_densify: bad arguments");
    if (out_rows == 0) return CG_OK;
    CgProfScope prof("csr_densify", (cudaStream_t)stream);
    k_csr_densify<<<(unsigned)out_rows, 256, 0, (cudaStream_t)stream>>>(dev_indptr, dev_indices, dev_values, dev_out, csr_rows, M);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
