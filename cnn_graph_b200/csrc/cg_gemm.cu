// K2 (fp32 CUDA-core form): the dense contractions around the Chebyshev stack.
//
//   contract      y[(n*M+m)][j]   = sum_{k,f} stack[k][m*N+n][f] * Wq(k,f,j)
//                 (lib/models.py:218-224: the [N*M, Fin*K] x [Fin*K, Fout] matmul, with the
//                 restack / transposes of :218-220 folded into the addressing)
//   stack_t_plain P(k,a,b)        = sum_{m,n} stack[k][m*N+n][a] * T[(n*M+m)][b]
//                 (the dW = stack^T . gy half of TF's MatMul gradient)
//
// Both flatten (k, f) into one reduction / output index q = k*F + f so that narrow
// feature counts (Fin = 1, 2) still fill the tiles.  fp32 FFMA with register tiles;
// operands staged through shared memory with 128-bit loads where F % 4 == 0.
#include "cg_common.cuh"

// ---------------------------------------------------------------------------
// contract
// ---------------------------------------------------------------------------
constexpr int CT_BM = 128;      // rows per CTA
constexpr int CT_QC = 16;       // reduction chunk
constexpr int CT_AP = 20;       // As pitch (floats): rows rg + 16 i land in distinct bank groups
constexpr int CT_THREADS = 128;

// weight element for reduction index (k, f) and output column j
//   normal     : W[(f*K + k) * J + j]          (reference layout, row = fin*K + k)
//   transposed : W[(j*K + k) * F + f]          (dx = Z . W^T: stack feature is fout, output is fin)
__device__ __forceinline__ float w_elem(const float *__restrict__ W, int k, int f, int j, int K, int F, int J,
                                        bool transposed) {
    return transposed ? W[((int64_t)j * K + k) * F + f] : W[((int64_t)f * K + k) * J + j];
}

template <int TN>   // columns per thread (4 or 8); BN = 8 * TN
__global__ void __launch_bounds__(CT_THREADS)
k_contract(const float *__restrict__ stack, const float *__restrict__ W, float *__restrict__ y, int64_t R, int N,
           int M, int F, int J, int K, int w_transposed, int vecA, int sample_major) {
    constexpr int BN = 8 * TN;
    constexpr int NJ = TN / 4;
    __shared__ __align__(16) float As[CT_BM * CT_AP];
    __shared__ __align__(16) float Bs[CT_QC * BN];

    const int t = threadIdx.x;
    const int rg = t >> 3, cg = t & 7;
    const int64_t r0 = (int64_t)blockIdx.x * CT_BM;
    const int j0 = blockIdx.y * BN;
    const int Q = K * F;
    const int64_t slab = R * (int64_t)F;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[4];                 // staged A (vector path) -- reused as 16 scalars on the scalar path
    float rb[CT_QC * BN / CT_THREADS];

    auto load_chunk = [&](int q0) {
        if (vecA) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = t + CT_THREADS * i;
                const int row = e >> 2, q = q0 + (e & 3) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + row < R && q < Q) {
                    const int k = q / F, f = q - k * F;
                    v = *reinterpret_cast<const float4 *>(stack + (int64_t)k * slab + (r0 + row) * F + f);
                }
                ra[i] = v;
            }
        } else {
            float *rs = reinterpret_cast<float *>(ra);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int e = t + CT_THREADS * i;
                int row, qq;
                if (F >= 8) { row = e >> 4; qq = e & 15; } else { row = e & 127; qq = e >> 7; }
                const int q = q0 + qq;
                float v = 0.f;
                if (r0 + row < R && q < Q) {
                    const int k = q / F, f = q - k * F;
                    v = stack[(int64_t)k * slab + (r0 + row) * F + f];
                }
                rs[i] = v;
            }
        }
#pragma unroll
        for (int i = 0; i < CT_QC * BN / CT_THREADS; ++i) {
            const int e = t + CT_THREADS * i;
            const int qq = e / BN, j = j0 + (e - qq * BN);
            const int q = q0 + qq;
            float v = 0.f;
            if (q < Q && j < J) {
                const int k = q / F, f = q - k * F;
                v = w_elem(W, k, f, j, K, F, J, w_transposed);
            }
            rb[i] = v;
        }
    };
    auto store_chunk = [&]() {
        if (vecA) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = t + CT_THREADS * i;
                *reinterpret_cast<float4 *>(As + (e >> 2) * CT_AP + (e & 3) * 4) = ra[i];
            }
        } else {
            const float *rs = reinterpret_cast<const float *>(ra);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int e = t + CT_THREADS * i;
                int row, qq;
                if (F >= 8) { row = e >> 4; qq = e & 15; } else { row = e & 127; qq = e >> 7; }
                As[row * CT_AP + qq] = rs[i];
            }
        }
#pragma unroll
        for (int i = 0; i < CT_QC * BN / CT_THREADS; ++i) Bs[t + CT_THREADS * i] = rb[i];
    };

    const int nchunks = (Q + CT_QC - 1) / CT_QC;
    load_chunk(0);
    for (int ch = 0; ch < nchunks; ++ch) {
        store_chunk();
        __syncthreads();
        if (ch + 1 < nchunks) load_chunk((ch + 1) * CT_QC);
#pragma unroll
        for (int q4 = 0; q4 < CT_QC / 4; ++q4) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(As + (rg + 16 * i) * CT_AP + q4 * 4);
#pragma unroll
            for (int qi = 0; qi < 4; ++qi) {
                float4 b[NJ];
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj)
                    b[jj] = *reinterpret_cast<const float4 *>(Bs + (q4 * 4 + qi) * BN + cg * 4 + 32 * jj);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av = qi == 0 ? a[i].x : (qi == 1 ? a[i].y : (qi == 2 ? a[i].z : a[i].w));
#pragma unroll
                    for (int jj = 0; jj < NJ; ++jj) {
                        acc[i][jj * 4 + 0] = fmaf(av, b[jj].x, acc[i][jj * 4 + 0]);
                        acc[i][jj * 4 + 1] = fmaf(av, b[jj].y, acc[i][jj * 4 + 1]);
                        acc[i][jj * 4 + 2] = fmaf(av, b[jj].z, acc[i][jj * 4 + 2]);
                        acc[i][jj * 4 + 3] = fmaf(av, b[jj].w, acc[i][jj * 4 + 3]);
                    }
                }
            }
        }
        __syncthreads();
    }

    const bool vecY = (J % 4 == 0) && ((((uintptr_t)y) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = r0 + rg + 16 * i;
        if (r >= R) continue;
        const int64_t n = r % N, m = r / N;
        float *dst = y + (sample_major ? r : n * M + m) * (int64_t)J;      // stack rows already in y's order?
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const int j = j0 + cg * 4 + 32 * jj;
            if (vecY && j + 3 < J) {
                *reinterpret_cast<float4 *>(dst + j) =
                    make_float4(acc[i][jj * 4], acc[i][jj * 4 + 1], acc[i][jj * 4 + 2], acc[i][jj * 4 + 3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (j + e < J) dst[j + e] = acc[i][jj * 4 + e];
            }
        }
    }
}

int cg_run_contract(const float *stack, const float *W, float *y, int N, int M, int F, int J, int K,
                    bool w_transposed, bool sample_major, cudaStream_t s) {
    const int64_t R = (int64_t)N * M;
    if (R == 0 || J == 0) return CG_OK;
    if (!w_transposed && cg_thin_supported(N, M, F, J, K)) {
        int dev = 0;
        CG_CHECK_CUDA(cudaGetDevice(&dev));
        return cg_run_thin_contract(stack, W, y, N, M, F, J, K, sample_major, cg_sm_budget(dev), s);
    }
    const int vecA = (F % 4 == 0) && ((((uintptr_t)stack) & 15) == 0);
    const unsigned gx = (unsigned)cg_ceil_div(R, CT_BM);
    CgProfScope prof("contract", s);
    if (J > 32) {
        dim3 grid(gx, (unsigned)cg_ceil_div(J, 64));
        k_contract<8><<<grid, CT_THREADS, 0, s>>>(stack, W, y, R, N, M, F, J, K, w_transposed ? 1 : 0, vecA,
                                                  sample_major ? 1 : 0);
    } else {
        dim3 grid(gx, 1);
        k_contract<4><<<grid, CT_THREADS, 0, s>>>(stack, W, y, R, N, M, F, J, K, w_transposed ? 1 : 0, vecA,
                                                  sample_major ? 1 : 0);
    }
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------
// stack_t_plain: P[q][b] = sum_r S_q[r] * T[trow(r)][b],  q = k*Fa + a
// ---------------------------------------------------------------------------
constexpr int SP_THREADS = 128;
constexpr int SP_RC = 32;       // rows per stage

template <int TA, int TB>       // output tile TA (q) x TB (b); micro tile 4 x 4; (TA/4)*(TB/4) == 128
__global__ void __launch_bounds__(SP_THREADS)
k_stack_t_plain(const float *__restrict__ stack, const float *__restrict__ T, float *__restrict__ part, int64_t R,
                int N, int M, int Fa, int Fb, int K, int64_t rows_per_split, int vecS, int vecT, int sample_major) {
    static_assert((TA / 4) * (TB / 4) == SP_THREADS, "tile/thread mismatch");
    __shared__ __align__(16) float Ss[SP_RC * TA];
    __shared__ __align__(16) float Ts[SP_RC * TB];
    const int t = threadIdx.x;
    const int Q = K * Fa;
    const int q0 = blockIdx.y * TA;
    const int b0 = blockIdx.z * TB;
    const int64_t r_beg = (int64_t)blockIdx.x * rows_per_split;
    const int64_t r_end = min(R, r_beg + rows_per_split);
    const int64_t slab = R * (int64_t)Fa;
    constexpr int BG = TB / 4;
    const int ag = t / BG, bg = t % BG;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t rb = r_beg; rb < r_end; rb += SP_RC) {
        // ---- stage S: SP_RC x TA
        if (vecS) {
            for (int e = t; e < SP_RC * TA / 4; e += SP_THREADS) {
                const int row = e / (TA / 4), qq = (e % (TA / 4)) * 4;
                const int q = q0 + qq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rb + row < r_end && q < Q) {
                    const int k = q / Fa, a = q - k * Fa;
                    v = *reinterpret_cast<const float4 *>(stack + (int64_t)k * slab + (rb + row) * Fa + a);
                }
                *reinterpret_cast<float4 *>(Ss + row * TA + qq) = v;
            }
        } else {
            for (int e = t; e < SP_RC * TA; e += SP_THREADS) {
                int row, qq;
                if (Fa >= 8) { row = e / TA; qq = e % TA; } else { row = e % SP_RC; qq = e / SP_RC; }
                const int q = q0 + qq;
                float v = 0.f;
                if (rb + row < r_end && q < Q) {
                    const int k = q / Fa, a = q - k * Fa;
                    v = stack[(int64_t)k * slab + (rb + row) * Fa + a];
                }
                Ss[row * TA + qq] = v;
            }
        }
        // ---- stage T: SP_RC x TB
        if (vecT) {
            for (int e = t; e < SP_RC * TB / 4; e += SP_THREADS) {
                const int row = e / (TB / 4), bb = (e % (TB / 4)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int64_t r = rb + row;
                if (r < r_end && b0 + bb < Fb) {
                    const int64_t n = r % N, m = r / N;
                    v = *reinterpret_cast<const float4 *>(T + (sample_major ? r : n * M + m) * (int64_t)Fb + b0 + bb);
                }
                *reinterpret_cast<float4 *>(Ts + row * TB + bb) = v;
            }
        } else {
            for (int e = t; e < SP_RC * TB; e += SP_THREADS) {
                const int row = e / TB, bb = e % TB;
                float v = 0.f;
                const int64_t r = rb + row;
                if (r < r_end && b0 + bb < Fb) {
                    const int64_t n = r % N, m = r / N;
                    v = T[(sample_major ? r : n * M + m) * (int64_t)Fb + b0 + bb];
                }
                Ts[row * TB + bb] = v;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < SP_RC; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(Ss + r * TA + ag * 4);
            const float4 b = *reinterpret_cast<const float4 *>(Ts + r * TB + bg * 4);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    // partial tile -> part[split][q][b]
    float *dst = part + (int64_t)blockIdx.x * Q * Fb;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ag * 4 + i;
        if (q >= Q) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = b0 + bg * 4 + j;
            if (b < Fb) dst[(int64_t)q * Fb + b] = acc[i][j];
        }
    }
}

// dW[...] = sum_split part[split][q][b];  q = k*Fa + a
//   direct: dW[(a*K + k) * Fb + b]      swap: dW[(b*K + k) * Fa + a]
// A block owns 32 consecutive outputs; its 8 warps each sum every 8th partial (one coalesced 128-byte read per
// partial) and the 8 group sums are added in a fixed order: deterministic, and enough loads in flight for the
// 30 MB of partials the plane-streaming dW kernel leaves behind (a thread per output summing 148 values one after
// the other took 25 us for them).
__global__ void __launch_bounds__(256)
k_reduce_partials(const float *__restrict__ part, float *__restrict__ dW, int splits, int Fa, int Fb, int K, int swap) {
    __shared__ float red[8][32];
    const int64_t total = (int64_t)K * Fa * Fb;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
        const int64_t i = base + lane;
        float s0 = 0.f, s1 = 0.f;
        if (i < total) {
            int sp = grp;
            for (; sp + 8 < splits; sp += 16) {
                s0 += part[(int64_t)sp * total + i];
                s1 += part[(int64_t)(sp + 8) * total + i];
            }
            if (sp < splits) s0 += part[(int64_t)sp * total + i];
        }
        red[grp][lane] = s0 + s1;
        __syncthreads();
        if (grp == 0 && i < total) {
            float s = red[0][lane];
#pragma unroll
            for (int g = 1; g < 8; ++g) s += red[g][lane];
            const int b = (int)(i % Fb);
            const int q = (int)(i / Fb);
            const int k = q / Fa, a = q - k * Fa;
            if (swap)
                dW[((int64_t)b * K + k) * Fa + a] = s;
            else
                dW[((int64_t)a * K + k) * Fb + b] = s;
        }
        __syncthreads();
    }
}

int cg_reduce_partials(const float *part, float *dW, int splits, int Fa, int Fb, int K, bool swap, cudaStream_t s) {
    CgProfScope prof("reduce_partials", s);
    const int64_t total = (int64_t)K * Fa * Fb;
    int64_t blocks = cg_ceil_div(total, 32);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_reduce_partials<<<(unsigned)blocks, 256, 0, s>>>(part, dW, splits, Fa, Fb, K, swap ? 1 : 0);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

static int sp_splits(int N, int M, int Fa, int Fb, int K, int sm_count) {
    const int64_t R = (int64_t)N * M;
    const bool wide_a = Fb <= 32;
    const int TA = wide_a ? 64 : 32, TB = wide_a ? 32 : 64;
    const int64_t tiles = cg_ceil_div((int64_t)K * Fa, TA) * cg_ceil_div(Fb, TB);
    int64_t splits = cg_ceil_div(4LL * sm_count, tiles);
    const int64_t max_splits = cg_ceil_div(R, 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    return (int)splits;
}

size_t cg_stack_t_plain_workspace(int N, int M, int Fa, int Fb, int K, int sm_count) {
    return sizeof(float) * (size_t)sp_splits(N, M, Fa, Fb, K, sm_count) * K * Fa * Fb;
}

int cg_run_stack_t_plain(const float *stack, const float *T, float *dW, int N, int M, int Fa, int Fb, int K,
                         bool swap, bool sample_major, float *workspace, int sm_count, cudaStream_t s) {
    const int64_t R = (int64_t)N * M;
    const int splits = sp_splits(N, M, Fa, Fb, K, sm_count);
    int64_t rows_per_split = cg_ceil_div(R, splits);
    rows_per_split = cg_ceil_div(rows_per_split, SP_RC) * SP_RC;
    const int used = (int)cg_ceil_div(R, rows_per_split);
    const int vecS = (Fa % 4 == 0) && ((((uintptr_t)stack) & 15) == 0);
    const int vecT = (Fb % 4 == 0) && ((((uintptr_t)T) & 15) == 0);
    const bool wide_a = Fb <= 32;
    {
    CgProfScope prof("stack_t_plain", s);
    if (wide_a) {
        dim3 grid((unsigned)used, (unsigned)cg_ceil_div((int64_t)K * Fa, 64), (unsigned)cg_ceil_div(Fb, 32));
        CG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "stack_t_plain: problem too large");
        k_stack_t_plain<64, 32><<<grid, SP_THREADS, 0, s>>>(stack, T, workspace, R, N, M, Fa, Fb, K, rows_per_split,
                                                           vecS, vecT, sample_major ? 1 : 0);
    } else {
        dim3 grid((unsigned)used, (unsigned)cg_ceil_div((int64_t)K * Fa, 32), (unsigned)cg_ceil_div(Fb, 64));
        CG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "stack_t_plain: problem too large");
        k_stack_t_plain<32, 64><<<grid, SP_THREADS, 0, s>>>(stack, T, workspace, R, N, M, Fa, Fb, K, rows_per_split,
                                                           vecS, vecT, sample_major ? 1 : 0);
    }
    CG_LAUNCH_CHECK();
    }
    return cg_reduce_partials(workspace, dW, used, Fa, Fb, K, swap, s);
}
