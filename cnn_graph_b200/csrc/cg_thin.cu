// Contractions with a SHORT reduction: Q = K * F <= 16 terms per vertex signal.
//
//   forward          y[r][j]      = sum_q  X_q[row(r)] * W[q][j]            (lib/models.py:218-220 with Fin * K <= 16)
//   weight gradient  dW[q][j]     = sum_r  X_q[row(r)] * T[r][j]            (its adjoint, lib/graph_model.py:296)
//
// q = k * F + f walks the slabs of a Chebyshev stack [K][rows][F]; r = n * M + m are the rows of y / T (sample-major)
// and row(r) = r for a sample-major stack or m * N + n for the vertex-major stack of the streaming recurrence.
// Callers: the 20NEWS-shaped first layer (Fin = 1, K = 5, 32 output features, vertex-major: notebooks/20news.ipynb cell 12)
// and the input filter of the gconv-LSTM gates (Fin = 2 zero-padded to 4, K = 3, 4H = 512 outputs: lib/gconv_lstm.py:185-207).
//
// 2 Q J flops per 4 J output bytes (forward) or 4 (J + Q) input bytes (gradient) is at most 8 flop/B: both are bound by
// the HBM stream of y / T, and a tensor-core tile (128 x J x 16 with Q <= 16 real terms, operands converted to bf16
// hi + mid) only adds work.  Here a warp takes tiles of 32 rows; a lane owns CPL consecutive output columns (the
// warp covers 32 * CPL), holds the Q stack values of ONE row of the tile and the row loop broadcasts them with
// shuffles: Q shuffles + Q * CPL fp32 FMAs + one load / store per row.  (First version, measured: a warp per row run
// with Q broadcast loads per row -- 270 instructions per row at Q = 12 and every batch of rows exposed to the
// memory latency: 88 us per launch against 50 us for the tensor-core GEMM it was meant to undercut.)
#include <algorithm>
#include <cstdlib>

#include "cg_common.cuh"

namespace {

constexpr int TW = 8;                   // warps per CTA
constexpr int QMAX = 16;

// A warp works on TILES of 32 rows: lane l holds the Q stack values of row l of the tile (loaded once per tile, all
// loads of the tile in flight together) and the row loop broadcasts them with shuffles -- no memory latency inside it.
//   sample-major stack: tile t = rows 32 t .. 32 t + 31 (lane = row: coalesced loads of the slabs)
//   vertex-major stack: tile (tm, tn) = vertices 4 tm .. + 3  x  samples 8 tn .. + 7, lane = 8 * (vertex) + sample: every
//                       load instruction reads four full 32-byte sectors of X_q[m][8 tn ..]; a warp walks tm, so the rows
//                       n * M + m of y / T it touches are runs of consecutive vertices of eight samples
struct ThinGeom {
    int64_t R;
    int N, M, sample_major, tiles_per_warp, tiles_m;     // tiles_m: vertex-major tiles along the vertices (per sample block)
    int64_t tiles;
};

struct ThinTile {
    int64_t r0;     // sample-major: first row
    int n0, m0;     // vertex-major: first sample / vertex
};
__device__ __forceinline__ ThinTile thin_tile(const ThinGeom &g, int64_t t) {
    ThinTile x;
    x.r0 = t * 32;
    x.n0 = (int)(t / g.tiles_m) * 8;
    x.m0 = (int)(t % g.tiles_m) * 4;
    return x;
}
// row i (0..31) of the tile: its index in y / T, or -1 beyond the tensor
__device__ __forceinline__ int64_t thin_row(const ThinGeom &g, const ThinTile &x, int i) {
    if (g.sample_major) {
        const int64_t r = x.r0 + i;
        return r < g.R ? r : -1;
    }
    const int n = x.n0 + (i & 7), m = x.m0 + (i >> 3);
    return (n < g.N && m < g.M) ? (int64_t)n * g.M + m : -1;
}
__device__ __forceinline__ int64_t thin_srow(const ThinGeom &g, const ThinTile &x, int i) {
    if (g.sample_major) return x.r0 + i;
    return (int64_t)(x.m0 + (i >> 3)) * g.N + x.n0 + (i & 7);
}

template <int CPL>
struct Cols {
    float v[CPL];
};
template <int CPL>
__device__ __forceinline__ Cols<CPL> ld_cols(const float *p, bool full, int j, int J) {
    Cols<CPL> c;
    if (CPL == 4 && full) {
        const float4 x = *reinterpret_cast<const float4 *>(p);
        c.v[0] = x.x; c.v[1 % CPL] = x.y; c.v[2 % CPL] = x.z; c.v[3 % CPL] = x.w;
    } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) c.v[i] = j + i < J ? p[i] : 0.f;
    }
    return c;
}
template <int CPL>
__device__ __forceinline__ void st_cols(float *p, bool full, int j, int J, const Cols<CPL> &c) {
    if (CPL == 4 && full) {
        *reinterpret_cast<float4 *>(p) = make_float4(c.v[0], c.v[1 % CPL], c.v[2 % CPL], c.v[3 % CPL]);
    } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i)
            if (j + i < J) p[i] = c.v[i];
    }
}

// W element of reduction term q and output column j: rows of W are f * K + k (lib/models.py:214-216)
__device__ __forceinline__ float w_at(const float *W, int q, int j, int K, int F, int J) {
    const int k = q / F, f = q - k * F;
    return W[((int64_t)f * K + k) * J + j];
}

// this lane's Q stack values of its row of the tile (zero beyond the tensor and for the padding terms q >= Q).  The loads are
// unconditional on clamped addresses and the zeros come from selects: a branch per load would serialise their latencies.
template <int QP>
__device__ __forceinline__ void thin_load_a(float (&a)[QP], const float *__restrict__ stack, const ThinGeom &g, const ThinTile &x,
                                            int lane, int Q, int F, int64_t slab) {
    const bool valid = thin_row(g, x, lane) >= 0;
    const float *sp = stack + (valid ? thin_srow(g, x, lane) : 0) * F;
    float v[QP];
#pragma unroll
    for (int q = 0; q < QP; ++q) {
        const int qq = q < Q ? q : 0;
        v[q] = __ldg(sp + (int64_t)(qq / F) * slab + (qq % F));
    }
#pragma unroll
    for (int q = 0; q < QP; ++q) a[q] = (valid && q < Q) ? v[q] : 0.f;
}

template <int QP, int CPL>
__global__ void __launch_bounds__(TW * 32, 2) k_thin_contract(const float *__restrict__ stack, const float *__restrict__ W,
                                                               float *__restrict__ y, const ThinGeom g, int F, int J, int K) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Q = K * F;
    const int j = (blockIdx.y * 32 + lane) * CPL;
    const bool full = j + CPL <= J;
    const int64_t slab = g.R * (int64_t)F;
    float w[QP][CPL];
#pragma unroll
    for (int q = 0; q < QP; ++q)
#pragma unroll
        for (int c = 0; c < CPL; ++c) w[q][c] = (q < Q && j + c < J) ? w_at(W, q, j + c, K, F, J) : 0.f;
    const int64_t t_beg = ((int64_t)blockIdx.x * TW + warp) * g.tiles_per_warp;
    const int64_t t_end = min(g.tiles, t_beg + g.tiles_per_warp);
    for (int64_t t = t_beg; t < t_end; ++t) {
        const ThinTile x = thin_tile(g, t);
        float a[QP];
        thin_load_a<QP>(a, stack, g, x, lane, Q, F, slab);
        // element offset of this lane's row in y (R * J < 2^31), broadcast with the stack values: no per-row index arithmetic
        const int64_t myr = thin_row(g, x, lane);
        const uint32_t myoff = myr >= 0 ? (uint32_t)(myr * J) : 0xffffffffu;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            Cols<CPL> acc;
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc.v[c] = 0.f;
#pragma unroll
            for (int q = 0; q < QP; ++q) {
                const float ai = __shfl_sync(0xffffffffu, a[q], i);
#pragma unroll
                for (int c = 0; c < CPL; ++c) acc.v[c] = fmaf(ai, w[q][c], acc.v[c]);
            }
            const uint32_t off = __shfl_sync(0xffffffffu, myoff, i);
            if (off != 0xffffffffu && j < J) st_cols<CPL>(y + off + j, full, j, J, acc);
        }
    }
}

template <int QP, int CPL>
__global__ void __launch_bounds__(TW * 32, 2) k_thin_dw(const float *__restrict__ stack, const float *__restrict__ T,
                                                         float *__restrict__ part, const ThinGeom g, int F, int J, int K) {
    __shared__ float red[QP][32 * CPL];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Q = K * F;
    const int j = (blockIdx.y * 32 + lane) * CPL;
    const bool active = j + CPL <= J;       // CPL = 4 needs J % 4 == 0: a lane is either fully inside or outside
    const int64_t slab = g.R * (int64_t)F;
    float acc[QP][CPL];
#pragma unroll
    for (int q = 0; q < QP; ++q)
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[q][c] = 0.f;
    const int64_t t_beg = ((int64_t)blockIdx.x * TW + warp) * g.tiles_per_warp;
    const int64_t t_end = min(g.tiles, t_beg + g.tiles_per_warp);
    for (int64_t t = t_beg; t < t_end; ++t) {
        const ThinTile x = thin_tile(g, t);
        float a[QP];
        thin_load_a<QP>(a, stack, g, x, lane, Q, F, slab);
        // rows in batches of RB: all T loads of a batch first (unconditional: a row beyond the tensor reads row 0 and has
        // a = 0 in its lane, a column beyond J reads column 0 and is never written), then the shuffles and FMAs
        constexpr int RB = CPL == 4 ? 4 : 8;
        const int64_t myr = thin_row(g, x, lane);
        const uint32_t myoff = myr >= 0 ? (uint32_t)(myr * J) : 0u;
        const float *tj = T + (active ? j : 0);
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += RB) {
            Cols<CPL> gv[RB];
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const float *tp = tj + __shfl_sync(0xffffffffu, myoff, i0 + b);
                if (CPL == 4) {
                    const float4 t4 = __ldg(reinterpret_cast<const float4 *>(tp));
                    gv[b].v[0] = t4.x; gv[b].v[1 % CPL] = t4.y; gv[b].v[2 % CPL] = t4.z; gv[b].v[3 % CPL] = t4.w;
                } else {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) gv[b].v[c] = __ldg(tp + c);
                }
            }
#pragma unroll
            for (int b = 0; b < RB; ++b) {
#pragma unroll
                for (int q = 0; q < QP; ++q) {
                    const float ai = __shfl_sync(0xffffffffu, a[q], i0 + b);
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[q][c] = fmaf(ai, gv[b].v[c], acc[q][c]);
                }
            }
        }
    }
    // warps add their sums one after the other (fixed order: deterministic), then the CTA writes its partial
    for (int w = 0; w < TW; ++w) {
        if (warp == w) {
#pragma unroll
            for (int q = 0; q < QP; ++q)
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    float *slot = &red[q][lane * CPL + c];
                    *slot = w == 0 ? acc[q][c] : *slot + acc[q][c];
                }
        }
        __syncthreads();
    }
    float *pp = part + (size_t)blockIdx.x * Q * J;
    for (int e = threadIdx.x; e < Q * 32 * CPL; e += TW * 32) {
        const int q = e / (32 * CPL), c = e - q * (32 * CPL);
        const int jj = blockIdx.y * 32 * CPL + c;
        if (jj < J) pp[(size_t)q * J + jj] = red[q][c];
    }
}

struct ThinPlan {
    int cpl, qp, gx, gy;
    ThinGeom g;
};

// warps_per_sm: how many warps per SM the grid aims for over all column groups (the gradient takes fewer, longer runs:
// every CTA leaves a partial sum)
static ThinPlan thin_plan(int N, int M, int F, int J, int K, bool sample_major, bool vec, int sm_count, int warps_per_sm) {
    ThinPlan p;
    const int Q = K * F;
    p.qp = Q <= 6 ? Q : Q <= 8 ? 8 : Q <= 12 ? 12 : 16;
    p.cpl = (vec && J % 4 == 0 && J >= 128) ? 4 : 1;
    p.gy = (int)cg_ceil_div(J, 32 * p.cpl);
    ThinGeom &g = p.g;
    g.R = (int64_t)N * M;
    g.N = N;
    g.M = M;
    g.sample_major = sample_major ? 1 : 0;
    g.tiles_m = (int)cg_ceil_div(M, 4);
    g.tiles = sample_major ? cg_ceil_div(g.R, 32) : (int64_t)g.tiles_m * cg_ceil_div(N, 8);
    const int64_t want_warps = std::max<int64_t>(TW, (int64_t)sm_count * warps_per_sm / p.gy);
    g.tiles_per_warp = (int)std::max<int64_t>(1, cg_ceil_div(g.tiles, want_warps));
    p.gx = (int)cg_ceil_div(g.tiles, (int64_t)g.tiles_per_warp * TW);
    return p;
}

}  // namespace

static bool thin_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("CG_THIN");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

bool cg_thin_supported(int N, int M, int F, int J, int K) {
    const int64_t R = (int64_t)N * M;
    return thin_enabled() && K * F >= 1 && K * F <= QMAX && J >= 1 && R > 0 && R * (int64_t)J < (1LL << 31) && R * (int64_t)K * F < (1LL << 31);
}

constexpr int THIN_FWD_WARPS = 64, THIN_DW_WARPS = 64;

size_t cg_thin_dw_workspace(int N, int M, int F, int J, int K, int sm_count) {
    if (!cg_thin_supported(N, M, F, J, K)) return 0;
    int gx = 0;
    for (int mode = 0; mode < 4; ++mode)
        gx = std::max(gx, thin_plan(N, M, F, J, K, (mode & 1) != 0, (mode & 2) != 0, sm_count, THIN_DW_WARPS).gx);
    return sizeof(float) * (size_t)gx * K * F * J;
}

#define CG_THIN_Q(KERNEL, CPL, ...)                                                                              \
    switch (p.qp) {                                                                                               \
        case 1: KERNEL<1, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                      \
        case 2: KERNEL<2, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                      \
        case 3: KERNEL<3, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                      \
        case 4: KERNEL<4, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                      \
        case 5: KERNEL<5, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                      \
        case 6: KERNEL<6, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                      \
        case 8: KERNEL<8, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                      \
        case 12: KERNEL<12, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                    \
        default: KERNEL<16, CPL><<<grid, TW * 32, 0, s>>>(__VA_ARGS__); break;                                    \
    }
#define CG_THIN_DISPATCH(KERNEL, ...)                                                                            \
    do {                                                                                                          \
        dim3 grid((unsigned)p.gx, (unsigned)p.gy);                                                                \
        if (p.cpl == 4) {                                                                                         \
            CG_THIN_Q(KERNEL, 4, __VA_ARGS__)                                                                     \
        } else {                                                                                                  \
            CG_THIN_Q(KERNEL, 1, __VA_ARGS__)                                                                     \
        }                                                                                                         \
    } while (0)

// y [N*M][J] = contraction of stack [K][N*M][F] with W [F*K][J]
int cg_run_thin_contract(const float *stack, const float *W, float *y, int N, int M, int F, int J, int K, bool sample_major,
                         int sm_count, cudaStream_t s) {
    CG_REQUIRE(cg_thin_supported(N, M, F, J, K), "cg_run_thin_contract: unsupported shape");
    if (N == 1) sample_major = true;        // one sample: both layouts coincide
    const ThinPlan p = thin_plan(N, M, F, J, K, sample_major, (((uintptr_t)y) & 15) == 0, sm_count, THIN_FWD_WARPS);
    CgProfScope prof("thin_contract", s);
    CG_THIN_DISPATCH(k_thin_contract, stack, W, y, p.g, F, J, K);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// dW [F*K][J] (rows f*K + k) = stack^T T;  workspace: cg_thin_dw_workspace bytes
int cg_run_thin_dw(const float *stack, const float *T, float *dW, int N, int M, int F, int J, int K, bool sample_major,
                   float *workspace, int sm_count, cudaStream_t s) {
    CG_REQUIRE(cg_thin_supported(N, M, F, J, K) && workspace, "cg_run_thin_dw: unsupported shape or no workspace");
    if (N == 1) sample_major = true;
    const ThinPlan p = thin_plan(N, M, F, J, K, sample_major, (((uintptr_t)T) & 15) == 0, sm_count, THIN_DW_WARPS);
    {
        CgProfScope prof("thin_dw", s);
        CG_THIN_DISPATCH(k_thin_dw, stack, T, workspace, p.g, F, J, K);
        CG_LAUNCH_CHECK();
    }
    return cg_reduce_partials(workspace, dW, p.gx, F, J, K, false, s);
}
