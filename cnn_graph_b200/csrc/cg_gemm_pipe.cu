// Pipelined fp32 GEMM on the tensor cores:  C[M x N] = op(A)[M x K] . op(B)[K x N] (+ bias[N]) (relu)
//
// The fast path behind cg_gemm_f32 / cg_run_gemm (cg_gemm_umma.cu keeps the generic kernel for operands that are
// not 16-byte aligned).  Same arithmetic as everywhere in this library: every fp32 operand value becomes
// bf16 hi + mid and the product is  hi*hi + mid*hi + hi*mid  with fp32 accumulation in tensor memory.
// Callers: the gate filters of the gconv-LSTM (lib/gconv_lstm.py:185-207 through lib/filter.py:89-95: the
// (Fin*K) x Fout contraction with Fout = 4H), their two gradients, and the dense head (lib/models.py:268-274).
//
// Persistent CTAs (one per SM) walk over (output tile 128 x BN, K split) work items.  Roles:
//   16 converter warps  per K stage (BK = 32) every warp loads 8 rows x 128 bytes (K-contiguous operand) or
//                       8 k x 128 bytes (row-contiguous operand) with two 16-byte copies per lane, each warp-wide copy covering
//                       four FULL 128-byte lines (a lane-per-row pattern costs one L1 wavefront per 32-byte sector
//                       and saturates the L1 data pipe -- measured).  Every lane splits its own 4-element pieces and
//                       stores the 8-byte bf16 half octets into the hi and mid planes of the canonical UMMA layout;
//                       the strides between core-matrix groups are padded by 32 bytes, which makes those 64-bit
//                       stores bank-conflict free (a lane-pair exchange to full 16-byte octets cost 16 more
//                       instructions per slot on the pipe that bounds the kernel: instruction issue).  The loads are
//                       cp.async copies into raw fp32 rings (A two stages ahead, B one) laid out line by line; a
//                       thread reads back only the 16-byte pieces it copied itself, so the rings need no barriers.
//   1 issue warp        waits for "full", issues 3 x 2 MMAs of 128 x BN x 16 (elect.sync), commits to "empty";
//                       after the last stage of a work item commits to the accumulator's "ready" barrier.
//   4 epilogue warps    drain the other of two TMEM accumulators (tcgen05.ld, 32 columns at a time), transpose
//                       through a padded shared-memory tile so that every store instruction writes four full
//                       128-byte lines, apply bias / relu, write C or the split-K partial, release the accumulator.
#include <algorithm>

#include "cg_common.cuh"
#include "cg_umma.cuh"
#include "cg_fused_common.cuh"

namespace {

constexpr int PC = 512;                 // converter threads
constexpr int PE = 128;                 // epilogue threads
constexpr int PT = PC + PE + 32;        // + issue warp
constexpr int BM = 128;
constexpr int BK = 32;
constexpr int MAX_STAGES = 6;
constexpr int RAW_A = 3, RAW_B = 2;        // raw fp32 rings: A (streamed from HBM) is fetched two stages ahead of the
                                        // conversion, B (L2-resident for most callers) one stage ahead
constexpr uint32_t MN_SBO = BK * 16 + 32;           // row-contiguous operand: stride between 8-row groups (padded)
constexpr uint32_t EP_ROW = 80;                     // epilogue staging: 16 floats + 16 bytes of padding per row
constexpr uint32_t EP_BYTES = 32 * EP_ROW;          // per epilogue warp

__host__ __device__ constexpr uint32_t kc_lbo(int rows) { return (uint32_t)rows * 16u + 32u; }   // between k octets (padded)
__host__ __device__ constexpr uint32_t plane_bytes(int rows) {
    return (uint32_t)(rows / 8) * MN_SBO > 4u * kc_lbo(rows) ? (uint32_t)(rows / 8) * MN_SBO : 4u * kc_lbo(rows);
}

struct PipeParams {
    int npass;                   // MMA passes per product: 3 (fp32-equivalent hi/mid split) or 1 (single-pass bf16)
    const float *A, *B, *bias;
    float *C;                           // [M][ldc] (split == 1) or partials [split][M][N]
    long long *trace;                   // optional (debug): clock64 stamps of CTA 0, stages 8..39: [32][8]
    int M, N, K, lda, ldb, ldc, relu, BN, split, k_per_split, tiles_n, n_work, nstage;
    // K blocking (see cg_common.cuh): block index = k >> sh, index inside = k & mask (sh = 31: not blocked)
    int a_sh, a_mask, b_sh, b_mask, b_shi, b_slo;
    long long a_kbs;
    // M blocking of a row-contiguous A (TA): element offset of row m = (m >> a_msh) * a_mbs + (m & a_mmask)
    int a_msh, a_mmask;
    long long a_mbs;
    uint32_t a_plane, b_plane, off_b, stage_bytes, off_raw_a, off_raw_b, raw_b_bytes, off_ep, off_bar, tmem_cols;
};

// 16-byte asynchronous copy global -> shared; bytes beyond `bytes` (0..16) are zero-filled
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// one lane polls the barrier (a spinning warp costs an LSU wavefront per probe), the others wait at the warp barrier;
// SLEEP > 0: back off between probes (long waits: accumulator hand-over)
template <int SLEEP>
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity, int lane) {
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// A "slot" is the work of one thread on one operand tile per stage: item warp `we` (= item index / 32) of the tile
// and the lane.  Loads (each lane 16 bytes; a quarter-warp = one 128-byte line):
//   KC (K contiguous source)    lines = rows 8*we + (lane>>3) [load 0] and + 4 [load 1], floats 4*(lane&7) .. + 3 of
//                               the stage's 32 k
//   MN (row contiguous source)  lines = k rows 8*(we&3) + (lane>>3) [load 0] and + 4 [load 1], row elements
//                               32*(we>>2) + 4*(lane&7) .. + 3
// Each 4-element piece is half an octet of the operand layout: octet (lane&7)>>1 (along k for KC, along the rows
// for MN), half lane&1, of row (KC) / k (MN) 8*we' + (lane>>3) (+ 4 for load 1).
struct Slot {
    uint32_t off;       // element offset from the operand base: KC row (load 0) * ld; MN first row element
    int nv;             // KC: bit 0 / bit 1 = row of load 0 / load 1 inside the matrix; MN: valid row elements (0..4)
};

template <bool KC>
__device__ __forceinline__ void slot_bind(Slot &s, int we, int lane, int ld, int r0, int r_lim, int msh = 31, int mmask = 0x7fffffff,
                                          long long mbs = 0) {
    if (KC) {
        const int r = r0 + 8 * we + (lane >> 3);
        s.nv = (r < r_lim ? 1 : 0) | (r + 4 < r_lim ? 2 : 0);
        s.off = (uint32_t)(r < r_lim ? r : 0) * (uint32_t)ld;
    } else {
        const int c = r0 + 32 * (we >> 2) + 4 * (lane & 7);
        s.nv = max(0, min(4, r_lim - c));
        s.off = s.nv > 0 ? (uint32_t)((long long)(c >> msh) * mbs + (c & mmask)) : 0u;
    }
}
// KC: blocked element address  base + (k >> sh) * kbs + (k & mask);   MN: source row (k >> sh) * shi + (k & mask) * slo
// The two 16-byte pieces go to the thread's own places in the raw ring (raw, raw + 512).
template <bool KC>
__device__ __forceinline__ void slot_fetch(const Slot &s, const float *src, int we, int lane, uint32_t raw, int k0, int k_lim,
                                           int ld, int sh, int mask, long long kbs, int shi, int slo) {
    if (KC) {
        const int gk = k0 + 4 * (lane & 7);
        const uint32_t nb = (uint32_t)max(0, min(4, k_lim - gk)) * 4u;
        const float *p = src + ((size_t)s.off + (size_t)(gk >> sh) * kbs + (gk & mask));
        cp_async16(raw, nb ? p : src, (s.nv & 1) ? nb : 0u);
        cp_async16(raw + 512u, ((s.nv & 2) && nb) ? p + 4 * (size_t)ld : src, (s.nv & 2) ? nb : 0u);
    } else {
        const int gk = k0 + 8 * (we & 3) + (lane >> 3), g1 = gk + 4;
        const float *p0 = src + ((size_t)s.off + ((size_t)(gk >> sh) * shi + (size_t)(gk & mask) * slo) * ld);
        const float *p1 = src + ((size_t)s.off + ((size_t)(g1 >> sh) * shi + (size_t)(g1 & mask) * slo) * ld);
        const uint32_t nb = (uint32_t)s.nv * 4u;
        cp_async16(raw, gk < k_lim ? p0 : src, gk < k_lim ? nb : 0u);
        cp_async16(raw + 512u, g1 < k_lim ? p1 : src, g1 < k_lim ? nb : 0u);
    }
}
// byte offset (hi plane) of the 8-byte half octet that load 0 of the thread produces; load 1 lands 64 bytes further
// (4 rows resp. 4 k down inside the same core matrix).  With the strides padded to 32 (mod 128) the 16 lanes of a
// half-warp hit 16 different 8-byte bank pairs: the 64-bit stores are conflict-free.
template <bool KC>
__device__ __forceinline__ uint32_t slot_dst(int we, int lane, int rows) {
    const uint32_t i = (uint32_t)(lane >> 3), o = (uint32_t)((lane & 7) >> 1), half = (uint32_t)(lane & 1) * 8u;
    if (KC)             // K-major: (k/8) * LBO + (r/8) * 128 + (r%8) * 16 + (k%8) * 2
        return o * kc_lbo(rows) + (uint32_t)we * 128u + i * 16u + half;
    // MN-major: (r/8) * SBO + (k/8) * 128 + (k%8) * 16 + (r%8) * 2
    return ((uint32_t)(we >> 2) * 4u + o) * MN_SBO + (uint32_t)(we & 3) * 128u + i * 16u + half;
}
// split the two 4-float pieces of the thread and store their bf16 hi / mid halves (8 bytes each)
__device__ __forceinline__ void slot_store(uint32_t hi_addr, uint32_t plane, const float4 v0, const float4 v1) {
    uint2 h0, m0, h1, m1;
    split4(v0, h0, m0);
    split4(v1, h1, m1);
    sts64(hi_addr, h0);
    sts64(hi_addr + plane, m0);
    sts64(hi_addr + 64u, h1);
    sts64(hi_addr + plane + 64u, m1);
}

// TA / TB: the operand is stored with its row index (m resp. n) contiguous -> MN-major; otherwise K contiguous
//   op(A)(m, k) = TA ? A[k * lda + m] : A[m * lda + k]        op(B)(k, n) = TB ? B[n * ldb + k] : B[k * ldb + n]
template <bool TA, bool TB>
__global__ void __launch_bounds__(PT, 1) k_gemm_pipe(const PipeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *full = bars;                          // [MAX_STAGES] converted operands of the stage are in place
    uint64_t *empty = bars + MAX_STAGES;            // [MAX_STAGES] MMAs of the stage completed
    uint64_t *acc_ready = bars + 2 * MAX_STAGES;    // [2] accumulator holds a finished work item
    uint64_t *acc_free = acc_ready + 2;             // [2] accumulator drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_free + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN, nstage = p.nstage;

    if (tid == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            umma::mbar_init(full + i, PC / 32);
            umma::mbar_init(empty + i, 1);
        }
        umma::mbar_init(acc_ready, 1);
        umma::mbar_init(acc_ready + 1, 1);
        umma::mbar_init(acc_free, PE / 32);
        umma::mbar_init(acc_free + 1, PE / 32);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, p.tmem_cols);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t st0 = umma::smem_u32(smem);

    // work item w -> (tile row tm, tile column tn, K split sp)
    auto item = [&](int w, int &m0, int &n0, int &k_beg, int &k_end, int &sp) {
        sp = w % p.split;
        const int t = w / p.split;
        const int tm = t / p.tiles_n, tn = t - tm * p.tiles_n;
        m0 = tm * BM;
        n0 = tn * BN;
        k_beg = sp * p.k_per_split;
        k_end = min(p.K, k_beg + p.k_per_split);
    };

    if (warp < PC / 32) {
        // =========================== converter warps ====================================
        // slots: A item warp = warp; B item warps = warp and warp + 16 (tile of BN rows = BN / 8 item warps)
        const bool b0_on = warp < BN / 8, b1_on = warp + PC / 32 < BN / 8;
        Slot sa, sb0, sb1;
        // two fetch cursors over the CTA's stage sequence (A runs RAW_A - 1 stages ahead, B RAW_B - 1)
        struct Cursor {
            int wi, si, nst, kbeg, kend, g;     // g = byte offset of the cursor's stage in its raw ring
        } ca = {0, 0, 0, 0, 0, 0}, cb = {0, 0, 0, 0, 0, 0};
        auto bind_a = [&]() {
            int m0, n0, sp;
            item(ca.wi, m0, n0, ca.kbeg, ca.kend, sp);
            ca.nst = (ca.kend - ca.kbeg + BK - 1) / BK;
            slot_bind<!TA>(sa, warp, lane, p.lda, m0, p.M, p.a_msh, p.a_mmask, p.a_mbs);
        };
        auto bind_b = [&]() {
            int m0, n0, sp;
            item(cb.wi, m0, n0, cb.kbeg, cb.kend, sp);
            cb.nst = (cb.kend - cb.kbeg + BK - 1) / BK;
            slot_bind<TB>(sb0, warp, lane, p.ldb, n0, p.N);
            slot_bind<TB>(sb1, warp + PC / 32, lane, p.ldb, n0, p.N);
        };
        ca.wi = cb.wi = blockIdx.x;
        ca.si = cb.si = ca.g = cb.g = 0;
        if (ca.wi < p.n_work) {
            bind_a();
            bind_b();
        }
        // raw rings: per stage one KB per item warp (two loads x 4 lines); the thread's own 16 bytes at lane * 16
        const uint32_t raw_a = st0 + p.off_raw_a + (uint32_t)warp * 1024u + (uint32_t)lane * 16u;
        const uint32_t raw_b0 = st0 + p.off_raw_b + (uint32_t)warp * 1024u + (uint32_t)lane * 16u;
        const uint32_t raw_b1 = raw_b0 + (uint32_t)(PC / 32) * 1024u;
        auto fetch_a = [&]() {
            if (ca.wi >= p.n_work) return;
            const uint32_t ro = (uint32_t)ca.g;
            slot_fetch<!TA>(sa, p.A, warp, lane, raw_a + ro, ca.kbeg + ca.si * BK, ca.kend, p.lda, p.a_sh, p.a_mask, p.a_kbs, 0, 1);
            ca.g = ca.g == (RAW_A - 1) * (BM / 8) * 1024 ? 0 : ca.g + (BM / 8) * 1024;
            if (++ca.si == ca.nst) {
                ca.si = 0;
                ca.wi += gridDim.x;
                if (ca.wi < p.n_work) bind_a();
            }
        };
        auto fetch_b = [&]() {
            if (cb.wi >= p.n_work) return;
            const uint32_t ro = (uint32_t)cb.g;
            const int k0 = cb.kbeg + cb.si * BK;
            if (b0_on) slot_fetch<TB>(sb0, p.B, warp, lane, raw_b0 + ro, k0, cb.kend, p.ldb, p.b_sh, p.b_mask, 0, p.b_shi, p.b_slo);
            if (b1_on)
                slot_fetch<TB>(sb1, p.B, warp + PC / 32, lane, raw_b1 + ro, k0, cb.kend, p.ldb, p.b_sh, p.b_mask, 0, p.b_shi, p.b_slo);
            cb.g = cb.g == (RAW_B - 1) * (int)p.raw_b_bytes ? 0 : cb.g + (int)p.raw_b_bytes;
            if (++cb.si == cb.nst) {
                cb.si = 0;
                cb.wi += gridDim.x;
                if (cb.wi < p.n_work) bind_b();
            }
        };
        int total = 0;      // stages of this CTA
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            int m0, n0, k_beg, k_end, sp;
            item(w, m0, n0, k_beg, k_end, sp);
            total += (k_end - k_beg + BK - 1) / BK;
        }
        const uint32_t da = slot_dst<!TA>(warp, lane, BM);
        const uint32_t db0 = p.off_b + slot_dst<TB>(warp, lane, BN), db1 = p.off_b + slot_dst<TB>(warp + PC / 32, lane, BN);
        // copy groups: one per iteration, holding A of stage gc + 2 and B of stage gc + 1 (possibly empty: uniform
        // group arithmetic); "all but the newest group complete" = A and B of stage gc have landed
        fetch_a();
        cp_async_commit();
        fetch_a();
        fetch_b();
        cp_async_commit();
        int slot = 0;
        uint32_t use = 0, roa = 0, rob = 0, sb = st0;       // use = gc / nstage; ring offsets of stage gc
        for (int gc = 0; gc < total; ++gc) {
            fetch_a();
            fetch_b();
            cp_async_commit();
            cp_async_wait<1>();
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && tid == 0 && gc >= 8 && gc < 40;
            if (tr) p.trace[(gc - 8) * 8 + 0] = clock64();
            if (use > 0) mbar_wait_warp<0>(empty + slot, (use - 1) & 1, lane);
            if (tr) p.trace[(gc - 8) * 8 + 1] = clock64();
            // all read-backs first (their latencies overlap), then split / store slot by slot
            const float4 a0 = lds128(raw_a + roa), a1 = lds128(raw_a + roa + 512u);
            float4 b00, b01, b10, b11;
            if (b0_on) {
                b00 = lds128(raw_b0 + rob);
                b01 = lds128(raw_b0 + rob + 512u);
            }
            if (b1_on) {
                b10 = lds128(raw_b1 + rob);
                b11 = lds128(raw_b1 + rob + 512u);
            }
            slot_store(sb + da, p.a_plane, a0, a1);
            if (b0_on) slot_store(sb + db0, p.b_plane, b00, b01);
            if (b1_on) slot_store(sb + db1, p.b_plane, b10, b11);
            if (tr) p.trace[(gc - 8) * 8 + 2] = clock64();
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + slot);
            if (tr) p.trace[(gc - 8) * 8 + 3] = clock64();
            sb += p.stage_bytes;
            if (++slot == nstage) {
                slot = 0;
                sb = st0;
                ++use;
            }
            roa = roa == (uint32_t)((RAW_A - 1) * (BM / 8) * 1024) ? 0u : roa + (uint32_t)(BM / 8) * 1024u;
            rob = rob == (uint32_t)(RAW_B - 1) * p.raw_b_bytes ? 0u : rob + p.raw_b_bytes;
        }
        cp_async_wait<0>();
    } else if (warp == (PC + PE) / 32) {
        // =========================== MMA issue warp ======================================
        const uint32_t idesc = umma::make_idesc_bf16(BM, BN, TA ? 1 : 0, TB ? 0 : 1);
        // K-major: LBO = padded stride between k octets, SBO = 128;  MN-major: LBO = 128 (k groups), SBO = padded
        const uint32_t a_lbo = TA ? 128u : kc_lbo(BM), a_sbo = TA ? MN_SBO : 128u;
        const uint32_t b_lbo = TB ? kc_lbo(BN) : 128u, b_sbo = TB ? 128u : MN_SBO;
        const uint32_t a_hi = umma::desc_hi(a_sbo), b_hi = umma::desc_hi(b_sbo);
        // one K = 16 step = two k octets: K-major -> 2 * LBO bytes, MN-major -> 2 * 128 bytes
        const uint32_t a_k = (TA ? 256u : 2u * a_lbo) >> 4, b_k = (TB ? 2u * b_lbo : 256u) >> 4;
        int g = 0, it = 0, slot = 0;
        uint32_t use = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x, ++it) {
            int m0, n0, k_beg, k_end, sp;
            item(w, m0, n0, k_beg, k_end, sp);
            const int nst = (k_end - k_beg + BK - 1) / BK;
            const int ab = it & 1;
            if (it >= 2) mbar_wait_warp<64>(acc_free + ab, (uint32_t)(((it >> 1) - 1) & 1), lane);
            umma::fence_after_sync();
            const uint32_t acc = tmem + (uint32_t)(ab * BN);
            for (int s = 0; s < nst; ++s, ++g) {
                const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && g >= 8 && g < 40;
                if (tr) p.trace[(g - 8) * 8 + 4] = clock64();
                mbar_wait_warp<0>(full + slot, use & 1, lane);
                if (tr) p.trace[(g - 8) * 8 + 5] = clock64();
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
                if (tr) p.trace[(g - 8) * 8 + 6] = clock64();
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
        const uint32_t ep = st0 + p.off_ep + (uint32_t)qd * EP_BYTES;       // this warp's staging tile [32][EP_ROW]
        const int ld = final_out ? p.ldc : p.N;
        const bool vec = (ld & 3) == 0 && ((((uintptr_t)p.C) & 15) == 0);
        const int cl = 4 * (lane & 3), rl = lane >> 2;                      // store side: 4 columns of row 8*j + rl
        int it = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x, ++it) {
            int m0, n0, k_beg, k_end, sp;
            item(w, m0, n0, k_beg, k_end, sp);
            const int ab = it & 1;
            mbar_wait_warp<256>(acc_ready + ab, (uint32_t)((it >> 1) & 1), lane);
            umma::fence_after_sync();
            float *cbase = final_out ? p.C : p.C + (size_t)sp * p.M * p.N;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                if (n0 + c0 >= p.N) break;          // warp-uniform
                float v[16];
                tmem_ld16(tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(ab * BN + c0), v);
                umma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)         // thread = row: its 16 columns into the staging row
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
                for (int j = 0; j < 4; ++j) {       // four lanes = 64 contiguous bytes (two full sectors) of one row
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
                        float *dst = cbase + (size_t)m * ld + n;
                        if (vec && n + 3 < p.N) {
                            *reinterpret_cast<float4 *>(dst) = x;
                        } else {
                            if (n < p.N) dst[0] = x.x;
                            if (n + 1 < p.N) dst[1] = x.y;
                            if (n + 2 < p.N) dst[2] = x.z;
                            if (n + 3 < p.N) dst[3] = x.w;
                        }
                    }
                }
                __syncwarp();                       // staging tile is reused by the next 16 columns
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

static int pow2_shift(int v) {       // log2(v) when v is a power of two, else -1
    if (v <= 0 || (v & (v - 1)) != 0) return -1;
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

struct PipePlan {
    int BN, split, k_per_split, tiles, nstage;
    size_t smem, ws;
    PipeParams pp;
};

static PipePlan pipe_plan(int M, int N, int K, int sm_count) {
    PipePlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.pp.npass = cg_mma_passes();
    pl.BN = N > 128 ? 256 : N > 64 ? 128 : N > 32 ? 64 : 32;
    const int tiles_n = (int)cg_ceil_div(N, pl.BN);
    pl.tiles = (int)cg_ceil_div(M, BM) * tiles_n;
    int split = 1;
    if (pl.tiles < sm_count) {
        split = std::max(1, sm_count / pl.tiles);
        const int max_split = (int)cg_ceil_div(K, 8 * BK);       // at least eight stages per work item
        if (split > max_split) split = std::max(1, max_split);
    }
    int kps = (int)cg_ceil_div(K, split);
    kps = (int)cg_ceil_div(kps, BK) * BK;
    pl.split = (int)cg_ceil_div(K, kps);
    pl.k_per_split = kps;
    PipeParams &pp = pl.pp;
    pp.a_plane = plane_bytes(BM);
    pp.b_plane = plane_bytes(pl.BN);
    pp.off_b = 2 * pp.a_plane;
    pp.stage_bytes = (uint32_t)cg_align_up(2 * pp.a_plane + 2 * pp.b_plane, 128);
    pp.raw_b_bytes = (uint32_t)(pl.BN / 8) * 1024u;
    const size_t raw = (size_t)RAW_A * (BM / 8) * 1024 + (size_t)RAW_B * pp.raw_b_bytes;
    const size_t fixed = (size_t)(PE / 32) * EP_BYTES + 256 + raw;
    pl.nstage = (int)std::min<size_t>(MAX_STAGES, ((size_t)(227 * 1024) - fixed) / pp.stage_bytes);
    pp.off_raw_a = (uint32_t)pl.nstage * pp.stage_bytes;
    pp.off_raw_b = pp.off_raw_a + (uint32_t)(RAW_A * (BM / 8) * 1024);
    pp.off_ep = pp.off_raw_a + (uint32_t)raw;
    pp.off_bar = pp.off_ep + (uint32_t)(PE / 32) * EP_BYTES;
    pl.smem = pp.off_bar + 256;
    pp.tmem_cols = (uint32_t)std::max(32, 2 * pl.BN);
    pp.tiles_n = tiles_n;
    pl.ws = pl.split > 1 ? sizeof(float) * (size_t)pl.split * M * N : 0;
    return pl;
}

}  // namespace

static long long *g_gemm_trace = nullptr;
extern "C" int cg_debug_gemm_trace(long long *dev_buf) {
    g_gemm_trace = dev_buf;
    return CG_OK;
}

size_t cg_gemm_pipe_workspace(int M, int N, int K, int sm_count) { return pipe_plan(M, N, K, sm_count).ws; }

bool cg_gemm_pipe_eligible(const float *A, const float *B, int M, int N, int K, int lda, int ldb, int transA, int transB,
                           int a_kblk, long long a_kbs, int b_kblk, int a_mblk, long long a_mbs) {
    if (a_mblk > 0 && (!transA || pow2_shift(a_mblk) < 2 || (a_mbs & 3) != 0 ||
                       (long long)cg_ceil_div(M, a_mblk) * a_mbs >= (1LL << 32)))
        return false;
    // row offsets are kept as 32-bit element counts
    if ((long long)(transA ? 1 : M) * lda >= (1LL << 32) || (long long)(transB ? N : 1) * ldb >= (1LL << 32)) return false;
    (void)K;
    if (((((uintptr_t)A) | ((uintptr_t)B)) & 15) != 0 || (lda & 3) != 0 || (ldb & 3) != 0) return false;
    if (a_kblk > 0 && (pow2_shift(a_kblk) < 2 || (a_kbs & 3) != 0)) return false;      // 16-byte pieces never straddle a block
    if (b_kblk > 0 && pow2_shift(b_kblk) < 0) return false;
    return true;
}

// same contract as cg_run_gemm (cg_gemm_umma.cu); the caller has checked cg_gemm_pipe_eligible
int cg_run_gemm_pipe(const float *A, const float *B, float *C, int M, int N, int K, int transA, int transB, int lda,
                     int ldb, int ldc, const float *bias, int relu, int a_kblk, long long a_kbs, int b_kblk, int b_shi,
                     int b_slo, void *workspace, size_t workspace_bytes, int sm_count, cudaStream_t s, int a_mblk,
                     long long a_mbs) {
    PipePlan pl = pipe_plan(M, N, K, sm_count);
    CG_REQUIRE(pl.ws == 0 || (workspace && workspace_bytes >= pl.ws), "cg_gemm_f32: workspace too small (%zu < %zu bytes)",
               workspace_bytes, pl.ws);
    PipeParams &pp = pl.pp;
    pp.A = A;
    pp.B = B;
    pp.bias = bias;
    pp.trace = g_gemm_trace;
    pp.C = pl.split > 1 ? reinterpret_cast<float *>(workspace) : C;
    pp.M = M;
    pp.N = N;
    pp.K = K;
    pp.lda = lda;
    pp.ldb = ldb;
    pp.ldc = ldc;
    pp.relu = relu ? 1 : 0;
    pp.BN = pl.BN;
    pp.split = pl.split;
    pp.k_per_split = pl.k_per_split;
    pp.n_work = pl.tiles * pl.split;
    pp.nstage = pl.nstage;
    pp.a_sh = a_kblk > 0 ? pow2_shift(a_kblk) : 31;
    pp.a_mask = a_kblk > 0 ? a_kblk - 1 : 0x7fffffff;
    pp.a_kbs = a_kblk > 0 ? a_kbs : 0;
    pp.b_sh = b_kblk > 0 ? pow2_shift(b_kblk) : 31;
    pp.b_mask = b_kblk > 0 ? b_kblk - 1 : 0x7fffffff;
    pp.b_shi = b_kblk > 0 ? b_shi : 0;
    pp.b_slo = b_kblk > 0 ? b_slo : 1;
    pp.a_msh = a_mblk > 0 ? pow2_shift(a_mblk) : 31;
    pp.a_mmask = a_mblk > 0 ? a_mblk - 1 : 0x7fffffff;
    pp.a_mbs = a_mblk > 0 ? a_mbs : 0;
    const unsigned grid = (unsigned)std::min(pp.n_work, sm_count);
    {
        CgProfScope prof("gemm_pipe", s);
#define CG_PIPE_LAUNCH(TA, TB)                                                                                       \
    do {                                                                                                              \
        CG_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_pipe<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); \
        k_gemm_pipe<TA, TB><<<grid, PT, pl.smem, s>>>(pp);                                                            \
    } while (0)
        if (transA && transB) CG_PIPE_LAUNCH(true, true);
        else if (transA) CG_PIPE_LAUNCH(true, false);
        else if (transB) CG_PIPE_LAUNCH(false, true);
        else CG_PIPE_LAUNCH(false, false);
#undef CG_PIPE_LAUNCH
        CG_LAUNCH_CHECK();
    }
    if (pl.split > 1) return cg_gemm_reduce(reinterpret_cast<const float *>(workspace), bias, C, M, N, ldc, pl.split, relu, s);
    return CG_OK;
}
