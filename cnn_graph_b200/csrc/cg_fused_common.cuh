// Device helpers shared by the fused recurrence kernels (cg_fused.cu, cg_clenshaw.cu): mbarrier / bulk-copy
// wrappers, shared-memory accesses with 32-bit addresses, the bf16 hi/mid split.
#pragma once
#include <algorithm>
#include <vector>

#include "cg_common.cuh"
#include "cg_umma.cuh"

namespace {

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// 1-D bulk copy global -> shared (TMA engine), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_smem),
                 "l"(src), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}

// 1-D bulk copy shared -> global (TMA engine), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void split4(const float4 v, uint2 &hi, uint2 &mid) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    const __nv_bfloat162 m01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
    const __nv_bfloat162 m23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
    hi.x = *reinterpret_cast<const uint32_t *>(&h01);
    hi.y = *reinterpret_cast<const uint32_t *>(&h23);
    mid.x = *reinterpret_cast<const uint32_t *>(&m01);
    mid.y = *reinterpret_cast<const uint32_t *>(&m23);
}

// shared-memory accesses with 32-bit addresses (one IADD of addressing)
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, const uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void fma4(float4 &a, float s, const float4 x) {
    a.x = fmaf(s, x.x, a.x);
    a.y = fmaf(s, x.y, a.y);
    a.z = fmaf(s, x.z, a.z);
    a.w = fmaf(s, x.w, a.w);
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Row-block gather shared by k_cheb_fused_b and k_cheb_clenshaw_b.
//
// A thread-level item is (block of 4 consecutive rows, sample, 4-feature chunk l): LPR = F / 4 lanes share a block and
// each gathers its own 16 bytes of every row the block's UNION of columns names (CgCsr::blk_*), applying the four
// rows' weights to it: one 128-bit shared-memory load feeds 16 FMAs, and a neighbour row that several of the four
// rows reference is read once (coarsened grids: 0.6 of the loads of a row-per-item gather).
//
// A warp holds QPW = 32 / LPR items side by side.  Their entry lists are interleaved in one table per (warp, item
// slot): record j = QPW slots of 32 bytes {w0, w1, w2, w3, byte offset of the gathered row, pad}, so the weights of
// the warp's items come with one conflict-free LDS.128 (slot stride 32 B) and the offsets with one LDS.32.  Lists
// are padded to the warp's longest (even) length with {0, 0}: blocks are dealt in order of descending length, sample
// fastest, so the items of a warp have (nearly) equal lengths and groups of several samples share identical lists.
// ---------------------------------------------------------------------------------------------------------------
struct BlkTables {
    const int *ptr, *col, *order;
    const float4 *w;
    int nblk;
};

template <int LPR>
struct BlkGeom {
    static constexpr int QPW = 32 / LPR;
    static constexpr int REC = QPW * 32;
};

// dealing position of item slot i of thread group tg: even slots in reverse order (the longest lists go to the last
// warps -- warp 0's scheduler also hosts the MMA issue warp), odd slots forward (every warp gets the same mix)
__host__ __device__ __forceinline__ int blk_deal(int i, int tg, int ngroups) {
    return i * ngroups + ((i & 1) ? tg : ngroups - 1 - tg);
}

template <int IPB>
struct BlkItems {
    uint32_t tab[IPB];      // shared-memory address of this thread's slot in record 0 of its table
    int trips[IPB];         // records to apply (even, >= 2, warp-uniform)
    int samp[IPB];          // sample of the item inside the group (>= S: absent)
    int row0[IPB];          // first row of the block
};

// Builds the per-warp tables in shared memory at `tab_base` and fills `it`.  Called by EVERY thread of the CTA
// (contains a __syncthreads); threads with tid >= NC (issue warp) get absent items.  `wsz`: NC / 32 * IPB ints of
// shared scratch.  The caller synchronises the CTA once more before the first gather.
template <int LPR, int IPB, int NC>
__device__ __forceinline__ void blk_setup(const BlkTables bt, int S, uint32_t rowbytes, uint32_t tab_base, int *wsz,
                                          BlkItems<IPB> &it) {
    constexpr int QPW = BlkGeom<LPR>::QPW, REC = BlkGeom<LPR>::REC, NG = NC / LPR;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool compute = tid < NC;
    const int tg = tid / LPR, sg = lane / LPR, l = lane % LPR;
    int blk[IPB], len[IPB];
#pragma unroll
    for (int i = 0; i < IPB; ++i) {
        const int o = blk_deal(i, tg, NG);
        blk[i] = -1;
        len[i] = 0;
        it.samp[i] = S;
        it.row0[i] = 0;
        if (compute && o < bt.nblk * S) {
            blk[i] = bt.order[o / S];
            it.samp[i] = o % S;
            it.row0[i] = 4 * blk[i];
            len[i] = bt.ptr[blk[i] + 1] - bt.ptr[blk[i]];
        }
        const int mx = __reduce_max_sync(0xffffffffu, len[i]);
        it.trips[i] = max(2, (mx + 1) & ~1);
        if (compute && lane == 0) wsz[warp * IPB + i] = it.trips[i];
    }
    __syncthreads();
    if (compute) {
        uint32_t off = 0;
        for (int j = 0; j < warp * IPB; ++j) off += (uint32_t)wsz[j] * REC;
#pragma unroll
        for (int i = 0; i < IPB; ++i) {
            const uint32_t t0 = tab_base + off + (uint32_t)sg * 32u;
            it.tab[i] = t0;
            const int b0 = blk[i] >= 0 ? bt.ptr[blk[i]] : 0;
            for (int j = l; j < it.trips[i]; j += LPR) {
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t o = 0;
                if (j < len[i]) {
                    w = bt.w[b0 + j];
                    o = (uint32_t)bt.col[b0 + j] * rowbytes;
                }
                sts128(t0 + (uint32_t)j * REC, w);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(t0 + (uint32_t)j * REC + 16u), "r"(o) : "memory");
            }
            off += (uint32_t)it.trips[i] * REC;
        }
    } else {
#pragma unroll
        for (int i = 0; i < IPB; ++i) it.tab[i] = tab_base;
    }
}

// acc[r] += sum_j w_j[r] * X[col_j]  for the four rows of one item; gbase = slab + sample offset + 16 * l.
// Reads two records past the list (the table region carries that much slack).
template <int LPR>
__device__ __forceinline__ void blk_gather(uint32_t gbase, uint32_t tab, int trips, float4 (&acc)[4]) {
    constexpr uint32_t REC = BlkGeom<LPR>::REC;
    float4 wa = lds128(tab), wb = lds128(tab + REC);
    uint32_t oa = lds32(tab + 16u), ob = lds32(tab + REC + 16u);
#pragma unroll 2
    for (int j = 0; j < trips; j += 2) {
        const float4 xa = lds128(gbase + oa), xb = lds128(gbase + ob);
        const float4 ca = wa, cb = wb;
        tab += 2u * REC;
        wa = lds128(tab);
        oa = lds32(tab + 16u);
        wb = lds128(tab + REC);
        ob = lds32(tab + REC + 16u);
        fma4(acc[0], ca.x, xa);
        fma4(acc[1], ca.y, xa);
        fma4(acc[2], ca.z, xa);
        fma4(acc[3], ca.w, xa);
        fma4(acc[0], cb.x, xb);
        fma4(acc[1], cb.y, xb);
        fma4(acc[2], cb.z, xb);
        fma4(acc[3], cb.w, xb);
    }
}

}  // namespace

// Shared-memory bytes of the block tables for a dealing of `nblk * S` items over NC / LPR thread groups (host side of
// blk_setup; includes the two records of read-ahead slack).
static inline size_t cg_blk_table_bytes(const std::vector<int> &len_sorted, int S, int LPR, int IPB, int NC) {
    const int QPW = 32 / LPR, REC = QPW * 32, NG = NC / LPR;
    const long long total = (long long)len_sorted.size() * S;
    size_t bytes = 0;
    for (int w = 0; w < NC / 32; ++w)
        for (int i = 0; i < IPB; ++i) {
            int mx = 0;
            for (int sg = 0; sg < QPW; ++sg) {
                const int o = blk_deal(i, w * QPW + sg, NG);
                if (o < total) mx = std::max(mx, len_sorted[(size_t)(o / S)]);
            }
            bytes += (size_t)std::max(2, (mx + 1) & ~1) * REC;
        }
    return bytes + 2 * (size_t)REC;
}
