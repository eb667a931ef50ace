// K1: Chebyshev recurrence  X_k = 2 L~ X_{k-1} - X_{k-2}  (lib/graph.py:241-258,
// lib/models.py:205-217, lib/filter.py:80-87) on a slab layout S[m][c].
//
// Two kernels:
//   k_basis_onchip  -- all K steps in ONE launch.  A CTA owns CW columns; the whole
//                      operator (ELL) and two M x CW signal slabs live in shared
//                      memory, neighbours are gathered with 128-bit LDS, every
//                      step's slab is streamed out to HBM with 128-bit stores.
//                      HBM traffic = read X once + write the K-1 new slabs.
//   k_spmm_step     -- one step per launch for operators too large for SMEM
//                      (CSR from HBM/L2, 128-bit gathers along the column axis).
#include <stdlib.h>

#include <algorithm>

#include "cg_common.cuh"

// ---------------------------------------------------------------------------
// streaming step:  out = alpha * L X1 - X0   (X0 may be null -> out = alpha * L X1)
// ---------------------------------------------------------------------------
template <int VEC>
struct VecT;
template <>
struct VecT<4> {
    using type = float4;
};
template <>
struct VecT<1> {
    using type = float;
};

__device__ __forceinline__ void fma_acc(float4 &a, float s, const float4 &x) {
    a.x = fmaf(s, x.x, a.x);
    a.y = fmaf(s, x.y, a.y);
    a.z = fmaf(s, x.z, a.z);
    a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void fma_acc(float &a, float s, const float &x) { a = fmaf(s, x, a); }
__device__ __forceinline__ float4 axmb(float alpha, const float4 &a, const float4 &b) {
    return make_float4(fmaf(alpha, a.x, -b.x), fmaf(alpha, a.y, -b.y), fmaf(alpha, a.z, -b.z), fmaf(alpha, a.w, -b.w));
}
__device__ __forceinline__ float axmb(float alpha, const float &a, const float &b) { return fmaf(alpha, a, -b); }
__device__ __forceinline__ float4 scale(float alpha, const float4 &a) {
    return make_float4(alpha * a.x, alpha * a.y, alpha * a.z, alpha * a.w);
}
__device__ __forceinline__ float scale(float alpha, const float &a) { return alpha * a; }
template <typename T>
__device__ __forceinline__ T zero_v();
template <>
__device__ __forceinline__ float4 zero_v<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <>
__device__ __forceinline__ float zero_v<float>() { return 0.f; }

// block = 256 threads = (256 / lpr) rows x lpr lanes; lane handles VEC columns.
// U gathers in flight per thread, at least B resident blocks per SM (the step is bound by load latency: ncu shows
// long-scoreboard stalls and 65 % occupancy at U = 4 without a register cap)
template <int VEC, int U, int B>
__global__ void __launch_bounds__(256, B)
k_spmm_step(const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
            const float *__restrict__ X1, const float *__restrict__ X0, float *__restrict__ out, int M,
            int64_t C, float alpha, int lpr) {
    using V = typename VecT<VEC>::type;
    const int lane = threadIdx.x % lpr;
    const int rows_per_block = 256 / lpr;
    const int64_t m = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / lpr;
    const int64_t c = ((int64_t)blockIdx.y * lpr + lane) * VEC;
    if (m >= M || c >= C) return;
    const int beg = rowptr[m], end = rowptr[m + 1];
    V acc = zero_v<V>();
    int e = beg;
    for (; e + U - 1 < end; e += U) {
        int cc[U];
        float vv[U];
        V xx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cc[u] = col[e + u];
            vv[u] = val[e + u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xx[u] = *reinterpret_cast<const V *>(X1 + (int64_t)cc[u] * C + c);
#pragma unroll
        for (int u = 0; u < U; ++u) fma_acc(acc, vv[u], xx[u]);
    }
    for (; e < end; ++e) {
        const V x = *reinterpret_cast<const V *>(X1 + (int64_t)col[e] * C + c);
        fma_acc(acc, val[e], x);
    }
    V r;
    if (X0 != nullptr) {
        const V old = *reinterpret_cast<const V *>(X0 + m * C + c);
        r = axmb(alpha, acc, old);
    } else {
        r = scale(alpha, acc);
    }
    *reinterpret_cast<V *>(out + m * C + c) = r;
}

// The same step on the row-block form of the operator (CgCsr::blk_*): a thread owns 4 consecutive rows x 4 columns and
// walks the UNION of the four rows' entries -- one 128-bit gather of X1 feeds 16 FMAs, and a neighbour that several of
// the four rows reference is fetched once.  On a locality-ordered kNN graph (C5: Morton order, 18 entries per row) the
// union is 0.44 of the entry count, which is what the L1 gather path -- the limiter of k_spmm_step, not HBM -- sees.
// block = 256 threads = (256 / lpr) row blocks x lpr lanes.
template <int U, int B>
__global__ void __launch_bounds__(256, B)
k_spmm_step_b(const int *__restrict__ bptr, const int *__restrict__ bcol, const float4 *__restrict__ bw,
              const float *__restrict__ X1, const float *__restrict__ X0, float *__restrict__ out, int M, int64_t C,
              float alpha, int lpr) {
    const int lane = threadIdx.x % lpr;
    const int64_t blk = (int64_t)blockIdx.x * (256 / lpr) + threadIdx.x / lpr;
    const int64_t m0 = 4 * blk;
    const int64_t c = ((int64_t)blockIdx.y * lpr + lane) * 4;
    if (m0 >= M || c >= C) return;
    const int beg = bptr[blk], end = bptr[blk + 1];
    float4 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *x1 = X1 + c;
    int e = beg;
    for (; e + U - 1 < end; e += U) {
        int cc[U];
        float4 ww[U], xx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cc[u] = bcol[e + u];
            ww[u] = bw[e + u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xx[u] = *reinterpret_cast<const float4 *>(x1 + (int64_t)cc[u] * C);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            fma_acc(acc[0], ww[u].x, xx[u]);
            fma_acc(acc[1], ww[u].y, xx[u]);
            fma_acc(acc[2], ww[u].z, xx[u]);
            fma_acc(acc[3], ww[u].w, xx[u]);
        }
    }
    for (; e < end; ++e) {
        const float4 w = bw[e];
        const float4 x = *reinterpret_cast<const float4 *>(x1 + (int64_t)bcol[e] * C);
        fma_acc(acc[0], w.x, x);
        fma_acc(acc[1], w.y, x);
        fma_acc(acc[2], w.z, x);
        fma_acc(acc[3], w.w, x);
    }
    const int nrow = (int)min((int64_t)4, (int64_t)M - m0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r >= nrow) break;
        float4 v;
        if (X0 != nullptr) {
            const float4 old = *reinterpret_cast<const float4 *>(X0 + (m0 + r) * C + c);
            v = axmb(alpha, acc[r], old);
        } else {
            v = scale(alpha, acc[r]);
        }
        *reinterpret_cast<float4 *>(out + (m0 + r) * C + c) = v;
    }
}

// Tiled form of the row-block step for slabs of 16 .. 128 columns: a tile is TR = 4 * (512 / lpr) consecutive rows
// (128 at C = 64).  The tile's rows of X1 and X0, its block pointers and its entries (columns + weights, one contiguous
// run of the block arrays) come in by five bulk copies (TMA engine, one mbarrier): every streaming operand of the step
// is read by large sequential copies with nothing in registers (the register path is latency bound: ncu shows 71 - 78 %
// long-scoreboard stalls at 33 % of the DRAM rate).  Entries whose column falls inside the tile -- 3/4 of them on a
// Morton-ordered kNN graph -- are gathered from the staged X1 tile in shared memory; the others ("far", stored last in
// every block) are fetched from L1 / L2 eight at a time before the near ones are applied.
constexpr int ST_THREADS = 512;
constexpr int ST_TILE_BYTES = ST_THREADS * 64;      // bytes of one X tile: every thread owns 4 rows x 16 bytes

struct TileParams {
    const int *bps, *bcol;       // CgCsr::blk_ps, blk_col
    const float4 *bw;
    const float *X1, *X0;
    float *out;
    const int *tiles;            // optional: the tiles to process (row-partitioned runs do interior and boundary tiles apart)
    int flags;                   // bit 0: L2 prefetch two tiles ahead
    int M, Mx, C, lpr, TR;       // M rows to compute; Mx >= M rows of X1 exist (row partition: halo rows follow the local ones)
    float alpha;
};

__device__ __forceinline__ uint32_t st_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Persistent and double-buffered: one CTA per SM walks tiles blockIdx.x, + gridDim.x, ...; a producer warp keeps the
// NEXT tile's copies in flight while the 16 compute warps work on the current one.
constexpr int SP_EMAX = 1536;              // entries staged per tile and stage: 24 KB of weights + 6 KB of columns
constexpr int SP_PTRS = 264;               // {ptr, split} pairs of a tile (<= 2 * 128 + 1 + alignment slack)
constexpr uint32_t SP_OFF_X0 = ST_TILE_BYTES, SP_OFF_W = 2 * ST_TILE_BYTES, SP_OFF_C = SP_OFF_W + SP_EMAX * 16,
                   SP_OFF_P = SP_OFF_C + (SP_EMAX + 8) * 4, SP_OFF_H = SP_OFF_P + SP_PTRS * 4,
                   SP_STAGE = (SP_OFF_H + 32 + 127) / 128 * 128;
static_assert(SP_STAGE % 128 == 0, "stage size keeps the tiles 128-byte aligned");

__device__ __forceinline__ void sp_bulk(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void sp_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
}

__device__ __forceinline__ void sp_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 sp_lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int sp_lds32(uint32_t a) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// one tile's copies (or, with bar == 0, their L2 prefetch two tiles ahead: one stage pair holds 2 x 95 KB, which is
// not enough bytes in flight per SM to cover the HBM latency; the prefetch makes the real copies L2 hits)
__device__ __forceinline__ void sp_issue_tile(const TileParams &p, int tile, unsigned char *st, uint32_t bar) {
    const int C = p.C;
    const int row0 = tile * p.TR;
    const int rows = min(p.TR, p.M - row0);
    const int b0 = row0 >> 2, nb = (rows + 3) >> 2;
    const int e0 = p.bps[2 * b0], e1 = p.bps[2 * (b0 + nb)];
    const bool staged = e1 - e0 <= SP_EMAX;
    const uint32_t xb = (uint32_t)rows * (uint32_t)C * 4u;
    const uint32_t x1b = (uint32_t)min(p.TR, p.Mx - row0) * (uint32_t)C * 4u;      // the whole window of X1: near entries may name rows >= M
    const int a0 = e0 & ~3, pa0 = (2 * b0) & ~3;      // 16-byte aligned starts of the column and pointer runs
    const uint32_t cb = staged ? (uint32_t)(((e1 - a0) + 3) & ~3) * 4u : 0u;
    const uint32_t wb = staged ? (uint32_t)(e1 - e0) * 16u : 0u;
    const uint32_t pb = (uint32_t)(((2 * (b0 + nb) + 1 - pa0) + 3) & ~3) * 4u;
    if (bar == 0u) {
        sp_prefetch_l2(p.X1 + (size_t)row0 * C, x1b);
        if (p.X0 != nullptr) sp_prefetch_l2(p.X0 + (size_t)row0 * C, xb);
        if (wb != 0u) sp_prefetch_l2(p.bw + e0, wb);
        if (cb != 0u) sp_prefetch_l2(p.bcol + a0, cb);
        return;
    }
    int *hdr = reinterpret_cast<int *>(st + SP_OFF_H);
    hdr[0] = e0;
    hdr[1] = e1 - e0;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                 "r"(x1b + (p.X0 != nullptr ? xb : 0u) + cb + wb + pb)
                 : "memory");
    sp_bulk(st_smem(st), p.X1 + (size_t)row0 * C, x1b, bar);
    if (p.X0 != nullptr) sp_bulk(st_smem(st + SP_OFF_X0), p.X0 + (size_t)row0 * C, xb, bar);
    sp_bulk(st_smem(st + SP_OFF_P), p.bps + pa0, pb, bar);
    if (wb != 0u) sp_bulk(st_smem(st + SP_OFF_W), p.bw + e0, wb, bar);
    if (cb != 0u) sp_bulk(st_smem(st + SP_OFF_C), p.bcol + a0, cb, bar);
}

// STAGED: the tile's entries are in shared memory (else, more than SP_EMAX of them: read from global memory).
// `st` points into the kernel's shared-memory array (plain loads: after inlining the compiler emits LDS and batches them).
// One walk over the block's entries in stored order, four gathers in flight; an entry whose row lies in the tile is read
// from the staged X1 rows, any other from L1 / L2.
template <bool STAGED>
__device__ __forceinline__ void sp_block(const TileParams &p, const unsigned char *st, int row0, int rows1, int e0, int beg,
                                         int end, int c, float4 (&acc)[4]) {
    const int *cols = STAGED ? reinterpret_cast<const int *>(st + SP_OFF_C) + (e0 & 3) : p.bcol + e0;
    const float4 *ws = STAGED ? reinterpret_cast<const float4 *>(st + SP_OFF_W) : p.bw + e0;
    // ONE generic load per entry and no branch: the BASE is selected (a warp holds two row blocks whose entries fall
    // inside / outside the tile independently), the index is the global row for both -- the shared-memory base is
    // pre-shifted by row0 rows.  Plain 64-bit integer arithmetic: 5 instructions per address.
    const uint32_t rowb = 4u * (uint32_t)p.C;
    const uint64_t gb = (uint64_t)(uintptr_t)(p.X1 + c);
    const uint64_t sb = (uint64_t)(uintptr_t)(reinterpret_cast<const float *>(st) + c) - (uint64_t)(uint32_t)row0 * rowb;
    auto gather = [&](int cc) {
        const uint64_t base = (unsigned)(cc - row0) < (unsigned)rows1 ? sb : gb;
        return *reinterpret_cast<const float4 *>((uintptr_t)(base + (uint64_t)(uint32_t)cc * rowb));
    };
    constexpr int U = 4;
    int e = beg;
    for (; e + U - 1 < end; e += U) {
        int cc[U];
        float4 ww[U], xx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cc[u] = cols[e + u];
            ww[u] = ws[e + u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xx[u] = gather(cc[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            fma_acc(acc[0], ww[u].x, xx[u]);
            fma_acc(acc[1], ww[u].y, xx[u]);
            fma_acc(acc[2], ww[u].z, xx[u]);
            fma_acc(acc[3], ww[u].w, xx[u]);
        }
    }
    for (; e < end; ++e) {
        const float4 w = ws[e];
        const float4 x = gather(cols[e]);
        fma_acc(acc[0], w.x, x);
        fma_acc(acc[1], w.y, x);
        fma_acc(acc[2], w.z, x);
        fma_acc(acc[3], w.w, x);
    }
}

__global__ void __launch_bounds__(ST_THREADS + 32, 1) k_spmm_tile_p(const TileParams p, int ntiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * SP_STAGE);      // full[2], empty[2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lpr = p.lpr, C = p.C;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem(bars + i)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem(bars + 2 + i)), "r"(ST_THREADS / 32) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t full0 = st_smem(bars), empty0 = st_smem(bars + 2);

    if (warp == ST_THREADS / 32) {
        // =========================== producer warp =====================================
        if (lane == 0) {
            int it = 0;
            for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x, ++it) {
                const int tile = p.tiles != nullptr ? p.tiles[ti] : ti;
                const int s = it & 1;
                if ((p.flags & 1) && p.tiles == nullptr) {
                    if (it == 0 && tile + (int)gridDim.x < ntiles) sp_issue_tile(p, tile + gridDim.x, nullptr, 0u);
                    if (tile + 2 * (int)gridDim.x < ntiles) sp_issue_tile(p, tile + 2 * gridDim.x, nullptr, 0u);      // L2 prefetch
                }
                if (it >= 2) sp_wait(empty0 + 8u * s, (uint32_t)(((it >> 1) - 1) & 1));
                sp_issue_tile(p, tile, smem + (size_t)s * SP_STAGE, full0 + 8u * s);
            }
        }
        return;
    }
    // =========================== compute warps =========================================
    const int ln = tid % lpr, bl = tid / lpr;
    const int c = ln * 4;
    int it = 0;
    for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x, ++it) {
        const int tile = p.tiles != nullptr ? p.tiles[ti] : ti;
        const int s = it & 1;
        unsigned char *st = smem + (size_t)s * SP_STAGE;
        const int row0 = tile * p.TR;
        const int rows = min(p.TR, p.M - row0);
        const int b0 = row0 >> 2, nb = (rows + 3) >> 2;
        if (p.flags & 2) {
            if (lane == 0) sp_wait(full0 + 8u * s, (uint32_t)((it >> 1) & 1));      // one polling lane per warp
            __syncwarp();
        } else {
            sp_wait(full0 + 8u * s, (uint32_t)((it >> 1) & 1));
        }
        const int *hdr = reinterpret_cast<const int *>(st + SP_OFF_H);
        const int e0 = hdr[0], ne = hdr[1];
        if (bl < nb && c < C) {
            const int *pt = reinterpret_cast<const int *>(st + SP_OFF_P) + ((2 * b0) & 3) + 2 * bl;
            const int beg = pt[0] - e0, end = pt[2] - e0;
            const int rows1 = min(p.TR, p.Mx - row0);       // rows of X1 staged (>= rows)
            float4 acc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ne <= SP_EMAX)
                sp_block<true>(p, st, row0, rows1, e0, beg, end, c, acc);
            else
                sp_block<false>(p, st, row0, rows1, e0, beg, end, c, acc);
            const int nrow = min(4, rows - 4 * bl);
            const float *x0t = reinterpret_cast<const float *>(st + SP_OFF_X0);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r >= nrow) break;
                const int lr = 4 * bl + r;
                float4 v;
                if (p.X0 != nullptr) {
                    const float4 old = *reinterpret_cast<const float4 *>(x0t + (size_t)lr * C + c);
                    v = axmb(p.alpha, acc[r], old);
                } else {
                    v = scale(p.alpha, acc[r]);
                }
                *reinterpret_cast<float4 *>(p.out + (size_t)(row0 + lr) * C + c) = v;
            }
        }
        // this warp has finished reading the stage
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8u * s) : "memory");
    }
}

// row-block form wanted?  It pays when neighbouring rows share neighbours; CG_SPMM_BLOCK=0/1 overrides (tests).
static bool step_blocked(const CgCsr &L, int64_t nnz_hint) {
    if (L.nblk == 0) return false;
    if (const char *env = getenv("CG_SPMM_BLOCK")) return atoi(env) != 0;
    return (int64_t)L.blk_total * 10 <= nnz_hint * 7;
}

// rows of one tile of the tiled step for slabs of C columns on this operator side; 0 when the tiled form does not apply
static int step_tile_rows(const CgCsr &L, int64_t nnz, int64_t C) {
    if (C % 4 != 0 || C / 4 < 4 || C / 4 > 32 || !step_blocked(L, nnz)) return 0;
    if (const char *env = getenv("CG_SPMM_TILE"))
        if (atoi(env) == 0) return 0;
    int lpr = 32;
    while (lpr > 1 && lpr / 2 >= C / 4) lpr /= 2;
    return 4 * (ST_THREADS / lpr);
}

static int launch_step(const CgCsr &L, int64_t nnz, int Mx, int M, const float *X1, const float *X0, float *out, int64_t C,
                       float alpha, cudaStream_t s, const int *tiles = nullptr, int ntiles_list = 0) {
    const bool vec4 = (C % 4 == 0) && ((((uintptr_t)X1 | (uintptr_t)out | (uintptr_t)X0) & 15) == 0);
    const int vec = vec4 ? 4 : 1;
    int64_t lanes_needed = cg_ceil_div(C, vec);
    int lpr = 32;
    while (lpr > 1 && lpr / 2 >= lanes_needed) lpr /= 2;
    CgProfScope prof("spmm_step", s);
    if (vec4 && step_blocked(L, nnz) && lanes_needed >= 4 && lanes_needed <= 32 && (int64_t)M * C < ((int64_t)1 << 31)) {
        // (16 <= C <= 128: at most 128 row blocks per tile, whole rows in one contiguous X tile)
        // tiled form (k_spmm_tile_p); CG_SPMM_TILE=0 keeps the register-path block kernel
        const char *env = getenv("CG_SPMM_TILE");
        if (env == nullptr || atoi(env) != 0) {
            TileParams tp;
            tp.bps = L.blk_ps;
            tp.bcol = L.blk_col;
            tp.bw = L.blk_w;
            tp.X1 = X1;
            tp.X0 = X0;
            tp.out = out;
            {
                const char *pf = getenv("CG_SPMM_PREFETCH");
                tp.flags = (pf != nullptr && atoi(pf) != 0) ? 1 : 0;      // L2 prefetch two tiles ahead: measured slower (11.6 vs 11.0 ms per 38 steps), off by default
                const char *pl = getenv("CG_SPMM_POLL1");
                if (pl != nullptr && atoi(pl) != 0) tp.flags |= 2;
            }
            tp.M = M;
            tp.Mx = Mx;
            tp.C = (int)C;
            tp.lpr = lpr;
            tp.TR = 4 * (ST_THREADS / lpr);
            tp.alpha = alpha;
            tp.tiles = tiles;
            const int ntiles = tiles != nullptr ? ntiles_list : (int)cg_ceil_div(M, tp.TR);
            if (ntiles == 0) return CG_OK;
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            sms = cg_sm_budget(dev);
            const size_t smem_p = 2 * (size_t)SP_STAGE + 64;
            static bool attr_p = false;
            if (!attr_p) {
                CG_CHECK_CUDA(cudaFuncSetAttribute(k_spmm_tile_p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
                attr_p = true;
            }
            k_spmm_tile_p<<<(unsigned)std::min(ntiles, sms), ST_THREADS + 32, smem_p, s>>>(tp, ntiles);
            CG_LAUNCH_CHECK();
            return CG_OK;
        }
    }
    CG_REQUIRE(tiles == nullptr, "spmm_step: a tile list needs the tiled step (16 <= C <= 128, C %% 4 == 0, aligned slabs, a local operator)");
    if (vec4 && step_blocked(L, nnz)) {
        const int blocks_per_cta = 256 / lpr;
        dim3 grid((unsigned)cg_ceil_div(cg_ceil_div(M, 4), blocks_per_cta), (unsigned)cg_ceil_div(lanes_needed, lpr));
        CG_REQUIRE(grid.y <= 65535, "spmm_step: too many columns (C=%lld)", (long long)C);
        k_spmm_step_b<4, 3><<<grid, 256, 0, s>>>(L.blk_ptr, L.blk_col, L.blk_w, X1, X0, out, M, C, alpha, lpr);
        CG_LAUNCH_CHECK();
        return CG_OK;
    }
    const int rows_per_block = 256 / lpr;
    dim3 grid((unsigned)cg_ceil_div(M, rows_per_block), (unsigned)cg_ceil_div(lanes_needed, lpr));
    CG_REQUIRE(grid.y <= 65535, "spmm_step: too many columns (C=%lld)", (long long)C);
    // measured at C5 (2^20 vertices, 18.6 M entries, C = 64): U = 4 gathers in flight at full occupancy (register cap 32)
    // 0.339 ms; without the cap (39 registers, 6 blocks) 0.378; U = 8 0.374 (6 blocks) / 0.465 (4); U = 2 at 8 blocks 0.351
    if (vec4)
        k_spmm_step<4, 4, 8><<<grid, 256, 0, s>>>(L.rowptr, L.col, L.val, X1, X0, out, M, C, alpha, lpr);
    else
        k_spmm_step<1, 4, 6><<<grid, 256, 0, s>>>(L.rowptr, L.col, L.val, X1, X0, out, M, C, alpha, lpr);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------
// adjoint (Clenshaw) step on sample-major data, batched over the samples:
//   out[n, m, :] = G[n, m, :] + alpha * sum_j L[m, j] X1[n, j, :] - X0[n, m, :]        (X0 may be null)
// Every tensor has its own row stride (elements between consecutive (n, m) rows), so that G_k can be a column block
// of the [N*M, K*F] product gy W^T and X1 the previous block.  block = 256 threads = (256 / lpr) rows x lpr lanes,
// a lane owns 4 columns; blockIdx.y = sample.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 8)
k_clenshaw_step(const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
                const float *__restrict__ G, int64_t sg, const float *__restrict__ X1, int64_t s1,
                const float *__restrict__ X0, int64_t s0, float *__restrict__ out, int64_t so, int M, int F, float alpha,
                int lpr) {
    const int lane = threadIdx.x % lpr;
    const int m = blockIdx.x * (256 / lpr) + threadIdx.x / lpr;
    const int c = lane * 4;
    if (m >= M || c >= F) return;
    const int64_t row0 = (int64_t)blockIdx.y * M;          // first row of the sample
    const int beg = rowptr[m], end = rowptr[m + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *x1 = X1 + row0 * s1 + c;
    int e = beg;
    for (; e + 3 < end; e += 4) {
        int cc[4];
        float vv[4];
        float4 xx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            cc[u] = col[e + u];
            vv[u] = val[e + u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xx[u] = *reinterpret_cast<const float4 *>(x1 + (int64_t)cc[u] * s1);
#pragma unroll
        for (int u = 0; u < 4; ++u) fma_acc(acc, vv[u], xx[u]);
    }
    for (; e < end; ++e) fma_acc(acc, val[e], *reinterpret_cast<const float4 *>(x1 + (int64_t)col[e] * s1));
    const int64_t r = row0 + m;
    const float4 g = *reinterpret_cast<const float4 *>(G + r * sg + c);
    float4 o = make_float4(fmaf(alpha, acc.x, g.x), fmaf(alpha, acc.y, g.y), fmaf(alpha, acc.z, g.z), fmaf(alpha, acc.w, g.w));
    if (X0 != nullptr) {
        const float4 b = *reinterpret_cast<const float4 *>(X0 + r * s0 + c);
        o.x -= b.x; o.y -= b.y; o.z -= b.z; o.w -= b.w;
    }
    *reinterpret_cast<float4 *>(out + r * so + c) = o;
}

// F % 4 == 0, F <= 128 * ... (one lane per 4 columns, at most 32 lanes per row: F <= 128); 16-byte aligned tensors and
// strides that are multiples of 4
bool cg_clenshaw_step_supported(int F) { return F % 4 == 0 && F >= 4 && F <= 128; }

int cg_run_clenshaw_step(const cg_graph *g, int transpose, const float *G, int64_t sg, const float *X1, int64_t s1,
                         const float *X0, int64_t s0, float *out, int64_t so, int N, int F, float alpha, cudaStream_t s) {
    CG_REQUIRE(cg_clenshaw_step_supported(F) && sg % 4 == 0 && s1 % 4 == 0 && so % 4 == 0 && (X0 == nullptr || s0 % 4 == 0),
               "clenshaw_step: unsupported width / strides (F=%d)", F);
    CG_REQUIRE(((((uintptr_t)G) | ((uintptr_t)X1) | ((uintptr_t)X0) | ((uintptr_t)out)) & 15) == 0, "clenshaw_step: unaligned tensor");
    CG_REQUIRE(N <= 65535, "clenshaw_step: too many samples for one launch (N=%d)", N);
    const CgCsr &L = cg_side(g, transpose);
    int lpr = 32;
    while (lpr > 1 && lpr / 2 >= F / 4) lpr /= 2;
    dim3 grid((unsigned)cg_ceil_div(g->M, 256 / lpr), (unsigned)N);
    CgProfScope prof("clenshaw_step", s);
    k_clenshaw_step<<<grid, 256, 0, s>>>(L.rowptr, L.col, L.val, G, sg, X1, s1, X0, s0, out, so, g->M, F, alpha, lpr);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// Wp[k*Fin + f][:] = W[f*K + k][:]   (rows of the filter weights regrouped by k)
__global__ void __launch_bounds__(256) k_regroup_w(const float *__restrict__ W, float *__restrict__ Wp, int Fin, int Fout, int K) {
    const int64_t total = (int64_t)Fin * K * Fout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / Fout;
        const int fo = (int)(i - row * Fout);
        const int k = (int)(row / Fin), f = (int)(row - (int64_t)k * Fin);
        Wp[i] = W[((int64_t)f * K + k) * Fout + fo];
    }
}

// dW[f*K + k][:] = T[k*Fin + f][:]
__global__ void __launch_bounds__(256) k_regroup_dw(const float *__restrict__ T, float *__restrict__ dW, int Fin, int Fout, int K) {
    const int64_t total = (int64_t)Fin * K * Fout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / Fout;
        const int fo = (int)(i - row * Fout);
        const int k = (int)(row / Fin), f = (int)(row - (int64_t)k * Fin);
        dW[((int64_t)f * K + k) * Fout + fo] = T[i];
    }
}

int cg_run_regroup_dw(const float *T, float *dW, int Fin, int Fout, int K, cudaStream_t s) {
    const int64_t total = (int64_t)Fin * K * Fout;
    k_regroup_dw<<<(unsigned)std::min<int64_t>(cg_ceil_div(total, 256), 1184), 256, 0, s>>>(T, dW, Fin, Fout, K);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

int cg_run_regroup_w(const float *W, float *Wp, int Fin, int Fout, int K, cudaStream_t s) {
    const int64_t total = (int64_t)Fin * K * Fout;
    k_regroup_w<<<(unsigned)std::min<int64_t>(cg_ceil_div(total, 256), 1184), 256, 0, s>>>(W, Wp, Fin, Fout, K);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ---------------------------------------------------------------------------
// on-chip fused recurrence
// ---------------------------------------------------------------------------
// shared memory:  ell[width][m_pad] float2 | len[m_pad] int | S0[M][CW] | S1[M][CW]
constexpr int kOnchipThreads = 512;

// Column c of the slab is signal n = c / F, feature f = c % F.  Two global layouts:
//   SM == 0  "vertex-major"  element (m, c) at m * C + c                  (slab layout S[m][c], lib/graph.py:241)
//   SM == 1  "sample-major"  element (m, c) at (n * M + m) * F + f        (the layout of x [N][M][F] itself)
template <int CW, int SM>
__global__ void __launch_bounds__(kOnchipThreads, 1)
k_basis_onchip(const float2 *__restrict__ ell_g, const int *__restrict__ rowptr, int width, int m_pad, int M,
               const float *__restrict__ in, float *__restrict__ stack, int64_t C, int K, int write_slab0, int F) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LPR = CW / 4;                 // lanes per row
    constexpr int RPP = kOnchipThreads / LPR;   // rows per pass
    float2 *ell = reinterpret_cast<float2 *>(smem_raw);
    int *len = reinterpret_cast<int *>(ell + (size_t)width * m_pad);
    float *S0 = reinterpret_cast<float *>(len + m_pad);
    float *S1 = S0 + (size_t)M * CW;

    const int tid = threadIdx.x;
    const int lane = tid % LPR;
    const int row0 = tid / LPR;
    const int64_t slab = (int64_t)M * C;

    // the operator is staged once per CTA; the CTA then walks its column groups (persistent when there are more groups
    // than SMs: the gate filters of the gconv-LSTM have N * H / CW = 800 groups, and 800 CTAs each staging the 78 KB
    // operator for a 32 KB tile spent most of their time on it -- 70 us per launch)
    for (int i = tid; i < width * m_pad; i += kOnchipThreads) ell[i] = ell_g[i];
    for (int m = tid; m < m_pad; m += kOnchipThreads) len[m] = m < M ? rowptr[m + 1] - rowptr[m] : 0;

    const int64_t ngroups = (C + CW - 1) / CW;
    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int64_t cbase = grp * CW + lane * 4;
        const bool cvalid = cbase < C;              // C % 4 == 0 -> the whole float4 is in range

        // global element offsets of the thread's four columns at m = 0, and the stride between vertices
        int64_t o[4];
        int64_t rs;
        bool vec;                                   // the four columns are contiguous and 16-byte aligned
        if (SM) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t c = cbase + i, n = c / F, f = c - n * F;
                o[i] = n * (int64_t)M * F + f;
            }
            rs = F;
            vec = (F % 4) == 0;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = cbase + i;
            rs = C;
            vec = true;
        }
        auto gload = [&](const float *base, int m) {
            if (vec) return *reinterpret_cast<const float4 *>(base + o[0] + (int64_t)m * rs);
            return make_float4(base[o[0] + (int64_t)m * rs], base[o[1] + (int64_t)m * rs], base[o[2] + (int64_t)m * rs],
                               base[o[3] + (int64_t)m * rs]);
        };
        auto gstore = [&](float *base, int m, const float4 v) {
            if (vec) {
                *reinterpret_cast<float4 *>(base + o[0] + (int64_t)m * rs) = v;
            } else {
                base[o[0] + (int64_t)m * rs] = v.x;
                base[o[1] + (int64_t)m * rs] = v.y;
                base[o[2] + (int64_t)m * rs] = v.z;
                base[o[3] + (int64_t)m * rs] = v.w;
            }
        };

        // (the previous group's last step ended with a barrier: nobody reads S0 / S1 any more)
        for (int m = row0; m < M; m += RPP) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cvalid) v = gload(in, m);
            *reinterpret_cast<float4 *>(S0 + m * CW + lane * 4) = v;
            if (write_slab0 && cvalid) gstore(stack, m, v);
        }
        __syncthreads();

        float *prev = S0, *cur = S1;                // cur holds X_{k-2} and receives X_k
        for (int k = 1; k < K; ++k) {
            float *dst = stack + (int64_t)k * slab;
            for (int m = row0; m < M; m += RPP) {
                const int n = len[m];
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int j = 0;
                for (; j + 1 < n; j += 2) {
                    const float2 e0 = ell[j * m_pad + m];
                    const float2 e1 = ell[(j + 1) * m_pad + m];
                    const float4 x0 = *reinterpret_cast<const float4 *>(prev + __float_as_int(e0.y) * CW + lane * 4);
                    const float4 x1 = *reinterpret_cast<const float4 *>(prev + __float_as_int(e1.y) * CW + lane * 4);
                    fma_acc(acc, e0.x, x0);
                    fma_acc(acc, e1.x, x1);
                }
                if (j < n) {
                    const float2 e0 = ell[j * m_pad + m];
                    const float4 x0 = *reinterpret_cast<const float4 *>(prev + __float_as_int(e0.y) * CW + lane * 4);
                    fma_acc(acc, e0.x, x0);
                }
                float4 *slot = reinterpret_cast<float4 *>(cur + m * CW + lane * 4);
                float4 r = acc;
                if (k > 1) r = axmb(2.0f, acc, *slot);
                *slot = r;
                if (cvalid) gstore(dst, m, r);
            }
            __syncthreads();
            float *t = prev;
            prev = cur;
            cur = t;
        }
    }
}

// widest column group (multiple of 8, power of two, <= 128) whose two slabs fit next to the ELL
static int onchip_cw(const cg_graph *g, const CgCsr &L, int64_t C) {
    if (!g->onchip || L.ell == nullptr || C % 4 != 0) return 0;
    const size_t fixed = (size_t)L.width * L.m_pad * sizeof(float2) + (size_t)L.m_pad * sizeof(int);
    const size_t limit = g->smem_optin;
    int best = 0;
    for (int cw = 8; cw <= 128; cw *= 2) {
        if (fixed + 2 * (size_t)g->M * cw * sizeof(float) <= limit) best = cw;
    }
    // do not pick a group much wider than the problem
    while (best > 8 && best / 2 >= C) best /= 2;
    // one CTA owns one column group for all K steps: prefer narrower groups while the grid does not fill the SMs
    while (best > 8 && cg_ceil_div(C, best) < g->sm_count) best /= 2;
    return best;
}

template <int CW, int SM>
static int launch_onchip(const cg_graph *g, const CgCsr &L, const float *in, float *stack, int64_t C, int K,
                         int write_slab0, int F, cudaStream_t s) {
    const size_t smem = (size_t)L.width * L.m_pad * sizeof(float2) + (size_t)L.m_pad * sizeof(int) +
                        2 * (size_t)g->M * CW * sizeof(float);
    CG_CHECK_CUDA(cudaFuncSetAttribute(k_basis_onchip<CW, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<int64_t>(cg_ceil_div(C, CW), g->sm_count);       // persistent over the column groups
    CgProfScope prof("basis_onchip", s);
    k_basis_onchip<CW, SM><<<grid, kOnchipThreads, smem, s>>>(L.ell, L.rowptr, L.width, L.m_pad, g->M, in, stack, C, K,
                                                              write_slab0, F);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

template <int SM>
static int dispatch_onchip(int cw, const cg_graph *g, const CgCsr &L, const float *in, float *stack, int64_t C, int K,
                           int w0, int F, cudaStream_t s) {
    switch (cw) {
        case 8: return launch_onchip<8, SM>(g, L, in, stack, C, K, w0, F, s);
        case 16: return launch_onchip<16, SM>(g, L, in, stack, C, K, w0, F, s);
        case 32: return launch_onchip<32, SM>(g, L, in, stack, C, K, w0, F, s);
        case 64: return launch_onchip<64, SM>(g, L, in, stack, C, K, w0, F, s);
        default: return launch_onchip<128, SM>(g, L, in, stack, C, K, w0, F, s);
    }
}

// Sample-major basis straight from x [N][M][F] into stack [K][N][M][F] (no permute): possible whenever the
// operator fits the on-chip kernel.
bool cg_basis_samples_supported(const cg_graph *g, int transpose, int N, int F) {
    const CgCsr &L = cg_side(g, transpose);
    return N > 0 && L.width > 0 && onchip_cw(g, L, (int64_t)N * F) > 0;
}

int cg_run_basis_samples(const cg_graph *g, int transpose, const float *x, float *stack, int N, int F, int K,
                         cudaStream_t s) {
    const CgCsr &L = cg_side(g, transpose);
    const int64_t C = (int64_t)N * F;
    const int cw = L.width > 0 ? onchip_cw(g, L, C) : 0;
    CG_REQUIRE(cw > 0, "cg_run_basis_samples: operator / columns do not fit the on-chip kernel");
    CG_REQUIRE(((((uintptr_t)x) | ((uintptr_t)stack)) & 15) == 0, "cg_run_basis_samples: unaligned tensor");
    if (K <= 1) {
        if (x != stack)
            CG_CHECK_CUDA(cudaMemcpyAsync(stack, x, sizeof(float) * (size_t)C * g->M, cudaMemcpyDeviceToDevice, s));
        return CG_OK;
    }
    return dispatch_onchip<1>(cw, g, L, x, stack, C, K, x != stack, F, s);
}

// `in` is slab 0 (may alias stack); fills stack[1..K-1] (and stack[0] when in != stack).
static int run_basis_from(const cg_graph *g, int transpose, const float *in, float *stack, int64_t C, int K,
                          cudaStream_t s, int flags) {
    const CgCsr &L = cg_side(g, transpose);
    const int M = g->M;
    const int64_t slab = (int64_t)M * C;
    const bool aligned = ((((uintptr_t)in) | ((uintptr_t)stack)) & 15) == 0;
    int cw = (flags & CG_FILTER_FORCE_STREAMING) || !aligned ? 0 : onchip_cw(g, L, C);
    if (L.width == 0) cw = 0;   // empty operator: streaming path handles it (all zeros)
    if ((flags & CG_FILTER_FORCE_ONCHIP) && cw == 0 && L.width > 0) {
        cg_set_error("cheb basis: on-chip kernel requested but operator/columns do not fit (M=%d width=%d C=%lld)",
                     M, L.width, (long long)C);
        return CG_ERR_ARG;
    }
    if (cw > 0 && K > 1) {
        return dispatch_onchip<0>(cw, g, L, in, stack, C, K, in != stack, 1, s);
    }
    if (in != stack)
        CG_CHECK_CUDA(cudaMemcpyAsync(stack, in, sizeof(float) * (size_t)slab, cudaMemcpyDeviceToDevice, s));
    for (int k = 1; k < K; ++k) {
        const float *x1 = stack + (int64_t)(k - 1) * slab;
        const float *x0 = k > 1 ? stack + (int64_t)(k - 2) * slab : nullptr;
        int rc = launch_step(L, g->nnz, g->M, M, x1, x0, stack + (int64_t)k * slab, C, k > 1 ? 2.0f : 1.0f, s);
        if (rc != CG_OK) return rc;
    }
    return CG_OK;
}

int cg_run_basis(const cg_graph *g, int transpose, float *stack, int64_t C, int K, cudaStream_t s, int flags) {
    return run_basis_from(g, transpose, stack, stack, C, K, s, flags);
}

extern "C" int cg_cheb_basis(const cg_graph_t *g, int transpose, const float *dev_X, float *dev_Xt, int64_t C,
                             int K, int flags, void *stream) {
    CG_REQUIRE(g && dev_X && dev_Xt, "cg_cheb_basis: NULL argument");
    CG_REQUIRE(C > 0 && K >= 1, "cg_cheb_basis: C and K must be positive (C=%lld K=%d)", (long long)C, K);
    return run_basis_from(g, transpose, dev_X, dev_Xt, C, K, (cudaStream_t)stream, flags);
}

// One recurrence step on device-resident slabs: out = alpha * L X1 - X0 (X0 may be NULL).  The building block of
// the row-partitioned recurrence (config C5): between two steps the caller exchanges the halo rows of X1.
// Only the first `rows` rows are computed (rows <= M; the remaining rows of the padded operator are halo slots).
extern "C" int cg_cheb_step(const cg_graph_t *g, int transpose, const float *dev_X1, const float *dev_X0,
                            float *dev_out, int rows, int64_t C, float alpha, void *stream) {
    CG_REQUIRE(g && dev_X1 && dev_out, "cg_cheb_step: NULL argument");
    CG_REQUIRE(C > 0 && rows >= 0 && rows <= g->M, "cg_cheb_step: bad C / rows (C=%lld rows=%d M=%d)", (long long)C, rows, g->M);
    if (rows == 0) return CG_OK;
    return launch_step(cg_side(g, transpose), g->nnz, g->M, rows, dev_X1, dev_X0, dev_out, C, alpha, (cudaStream_t)stream);
}

extern "C" int cg_cheb_step_tile_rows(const cg_graph_t *g, int transpose, int64_t C) {
    if (g == nullptr || C <= 0) return 0;
    return step_tile_rows(cg_side(g, transpose), g->nnz, C);
}

extern "C" int cg_cheb_step_tiles(const cg_graph_t *g, int transpose, const float *dev_X1, const float *dev_X0, float *dev_out,
                                  int rows, int64_t C, float alpha, const int32_t *dev_tiles, int ntiles, void *stream) {
    CG_REQUIRE(g && dev_X1 && dev_out && dev_tiles, "cg_cheb_step_tiles: NULL argument");
    CG_REQUIRE(C > 0 && rows >= 0 && rows <= g->M && ntiles >= 0, "cg_cheb_step_tiles: bad C / rows / ntiles");
    CG_REQUIRE(step_tile_rows(cg_side(g, transpose), g->nnz, C) > 0, "cg_cheb_step_tiles: the tiled step does not take C=%lld on this operator",
               (long long)C);
    CG_REQUIRE(((((uintptr_t)dev_X1) | ((uintptr_t)dev_X0) | ((uintptr_t)dev_out)) & 15) == 0, "cg_cheb_step_tiles: unaligned slab");
    if (ntiles == 0) return CG_OK;
    return launch_step(cg_side(g, transpose), g->nnz, g->M, rows, dev_X1, dev_X0, dev_out, C, alpha, (cudaStream_t)stream, dev_tiles,
                       ntiles);
}


// Halo rows of a row-partitioned slab fetched straight from the owners' buffers (peer memory over NVLink / NVSwitch):
// dst[i][:] = peer[src_rank[i]][slab_off + src_row[i] * C ...].  8 lanes x 16 bytes cover a 128-byte line per access.
__global__ void __launch_bounds__(256) k_halo_pull(const float *const *__restrict__ peers, const int *__restrict__ src_rank,
                                                   const int *__restrict__ src_row, long long slab_off, float *__restrict__ dst,
                                                   long long nhalo, int C) {
    const int c4 = C / 4;
    const long long total = nhalo * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c4;
        const int q = (int)(i - r * c4);
        const float4 *src = reinterpret_cast<const float4 *>(peers[src_rank[r]] + slab_off + (long long)src_row[r] * C) + q;
        reinterpret_cast<float4 *>(dst + r * C)[q] = *src;
    }
}

extern "C" int cg_halo_pull(const void *dev_peer_ptrs, const int32_t *dev_src_rank, const int32_t *dev_src_row, int64_t slab_offset,
                            float *dev_dst, int64_t nhalo, int C, void *stream) {
    CG_REQUIRE(dev_peer_ptrs && dev_dst, "cg_halo_pull: NULL argument");
    CG_REQUIRE(C > 0 && C % 4 == 0 && nhalo >= 0 && slab_offset % 4 == 0, "cg_halo_pull: C must be a multiple of 4 (C=%d)", C);
    if (nhalo == 0) return CG_OK;
    CG_REQUIRE(dev_src_rank && dev_src_row, "cg_halo_pull: NULL index arrays");
    const long long total = nhalo * (C / 4);
    const long long blocks = std::min<long long>(cg_ceil_div(total, 256), 148 * 8);
    CgProfScope prof("halo_pull", (cudaStream_t)stream);
    k_halo_pull<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float *const *>(dev_peer_ptrs), dev_src_rank,
                                                                   dev_src_row, slab_offset, dev_dst, nhalo, C);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
