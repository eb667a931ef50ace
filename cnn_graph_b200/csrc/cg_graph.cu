// Packed rescaled-Laplacian handle: host CSR -> device CSR + ELL, both orientations.
// Replaces the per-call COO -> tf.SparseTensor -> tf.sparse_reorder staging of the
// reference (lib/models.py:198-201, lib/filter.py:66-70).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "cg_common.cuh"

static thread_local char g_err[512] = "";

void cg_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *cg_last_error(void) { return g_err; }

static int g_mma_passes = 3;
int cg_mma_passes() { return g_mma_passes; }
extern "C" int cg_set_precision(int mode) {
    CG_REQUIRE(mode == CG_PRECISION_FP32 || mode == CG_PRECISION_BF16, "cg_set_precision: mode must be CG_PRECISION_FP32 (0) or CG_PRECISION_BF16 (1)");
    g_mma_passes = mode == CG_PRECISION_BF16 ? 1 : 3;
    return CG_OK;
}
extern "C" int cg_get_precision(void) { return g_mma_passes == 1 ? CG_PRECISION_BF16 : CG_PRECISION_FP32; }

int cg_sm_budget(int device) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    if (v <= 0) v = 148;
    if (const char *env = getenv("CG_SM_RESERVE")) {
        const int r = atoi(env);
        if (r > 0 && r < v) v -= r;
    }
    return v;
}
extern "C" int cg_abi_version(void) { return CG_ABI_VERSION; }

// ELL budget: the on-chip kernels keep one orientation's ELL plus two signal
// slabs in shared memory; the operator may take at most this many bytes.
static const size_t kEllSmemBudget = 160 * 1024;

static int upload_side(CgCsr &dst, int M, int64_t nnz, const std::vector<int> &rowptr,
                       const std::vector<int> &col, const std::vector<float> &val, bool want_ell) {
    CG_CHECK_CUDA(cudaMalloc(&dst.rowptr, sizeof(int) * (size_t)(M + 1)));
    CG_CHECK_CUDA(cudaMalloc(&dst.col, sizeof(int) * (size_t)std::max<int64_t>(nnz, 1)));
    CG_CHECK_CUDA(cudaMalloc(&dst.val, sizeof(float) * (size_t)std::max<int64_t>(nnz, 1)));
    CG_CHECK_CUDA(cudaMemcpy(dst.rowptr, rowptr.data(), sizeof(int) * (size_t)(M + 1), cudaMemcpyHostToDevice));
    if (nnz > 0) {
        CG_CHECK_CUDA(cudaMemcpy(dst.col, col.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
        CG_CHECK_CUDA(cudaMemcpy(dst.val, val.data(), sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice));
    }
    {
        std::vector<int> order(M);
        for (int m = 0; m < M; ++m) order[m] = m;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return rowptr[a + 1] - rowptr[a] > rowptr[b + 1] - rowptr[b];
        });
        CG_CHECK_CUDA(cudaMalloc(&dst.order, sizeof(int) * (size_t)M));
        CG_CHECK_CUDA(cudaMemcpy(dst.order, order.data(), sizeof(int) * (size_t)M, cudaMemcpyHostToDevice));
    }
    int width = 0;
    for (int m = 0; m < M; ++m) width = std::max(width, rowptr[m + 1] - rowptr[m]);
    dst.width = width;
    dst.m_pad = (M + 31) / 32 * 32;
    if (want_ell && width > 0) {
        std::vector<float2> ell((size_t)dst.m_pad * width, make_float2(0.f, 0.f));  // {0.0f, col 0}
        for (int m = 0; m < M; ++m) {
            int j = 0;
            for (int e = rowptr[m]; e < rowptr[m + 1]; ++e, ++j) {
                float2 v;
                v.x = val[e];
                int c = col[e];
                memcpy(&v.y, &c, sizeof(int));
                ell[(size_t)j * dst.m_pad + m] = v;
            }
        }
        CG_CHECK_CUDA(cudaMalloc(&dst.ell, sizeof(float2) * ell.size()));
        CG_CHECK_CUDA(cudaMemcpy(dst.ell, ell.data(), sizeof(float2) * ell.size(), cudaMemcpyHostToDevice));
    }
    {
        // row-block (4 rows) union form, see CgCsr (built for every operator: the streaming step uses it too)
        const int nblk = (M + 3) / 4;
        // experiment switch: near entries (the block's own 128-row window) first; default off (column-sorted run: the tiled
        // step with a far batch was slower than the plain sorted walk)
        const bool split_near_far = getenv("CG_BLK_SPLIT") != nullptr && atoi(getenv("CG_BLK_SPLIT")) != 0;
        std::vector<int> bptr(nblk + 1, 0), bcol, bps(2 * (size_t)nblk + 1 + 8, 0);
        std::vector<float4> bw;
        for (int b = 0; b < nblk; ++b) {
            int cur[4], end[4];
            for (int i = 0; i < 4; ++i) {
                const int m = 4 * b + i;
                cur[i] = m < M ? rowptr[m] : 0;
                end[i] = m < M ? rowptr[m + 1] : 0;
            }
            for (;;) {      // 4-way merge of the sorted rows
                int c = INT32_MAX;
                for (int i = 0; i < 4; ++i)
                    if (cur[i] < end[i]) c = std::min(c, col[cur[i]]);
                if (c == INT32_MAX) break;
                float w[4] = {0.f, 0.f, 0.f, 0.f};
                for (int i = 0; i < 4; ++i)
                    if (cur[i] < end[i] && col[cur[i]] == c) w[i] = val[cur[i]++];
                bcol.push_back(c);
                bw.push_back(make_float4(w[0], w[1], w[2], w[3]));
            }
            // near entries (column inside the block's 128-row window) first: the tiled streaming step gathers them from
            // shared memory and batches the far ones (cg_spmm.cu); a stable partition keeps both parts sorted by column
            if (split_near_far) {
                const int lo = bptr[b], hi = (int)bcol.size(), win = (4 * b) >> 7;
                std::vector<int> tc;
                std::vector<float4> tw;
                for (int pass = 0; pass < 2; ++pass)
                    for (int e = lo; e < hi; ++e)
                        if (((bcol[e] >> 7) == win) == (pass == 0)) {
                            tc.push_back(bcol[e]);
                            tw.push_back(bw[e]);
                        }
                int nnear = 0;
                for (int e = lo; e < hi; ++e) nnear += (bcol[e] >> 7) == win ? 1 : 0;
                std::copy(tc.begin(), tc.end(), bcol.begin() + lo);
                std::copy(tw.begin(), tw.end(), bw.begin() + lo);
                bps[2 * b] = lo;
                bps[2 * b + 1] = lo + nnear;
            } else {
                bps[2 * b] = bptr[b];
                bps[2 * b + 1] = (int)bcol.size();      // everything "near": one column-sorted run
            }
            bptr[b + 1] = (int)bcol.size();
            bps[2 * b + 2] = (int)bcol.size();
        }
        std::vector<int> border(nblk);
        for (int b = 0; b < nblk; ++b) border[b] = b;
        std::stable_sort(border.begin(), border.end(),
                         [&](int a, int b) { return bptr[a + 1] - bptr[a] > bptr[b + 1] - bptr[b]; });
        dst.nblk = nblk;
        dst.blk_total = (int)bcol.size();
        dst.blk_len_sorted.resize(nblk);
        for (int i = 0; i < nblk; ++i) dst.blk_len_sorted[i] = bptr[border[i] + 1] - bptr[border[i]];
        const size_t nt = std::max<size_t>(bcol.size(), 1);
        CG_CHECK_CUDA(cudaMalloc(&dst.blk_ptr, sizeof(int) * (size_t)(nblk + 1 + 8)));     // + slack, as blk_col
        CG_CHECK_CUDA(cudaMemset(dst.blk_ptr, 0, sizeof(int) * (size_t)(nblk + 1 + 8)));
        CG_CHECK_CUDA(cudaMalloc(&dst.blk_order, sizeof(int) * (size_t)nblk));
        CG_CHECK_CUDA(cudaMalloc(&dst.blk_ps, sizeof(int) * bps.size()));
        CG_CHECK_CUDA(cudaMemcpy(dst.blk_ps, bps.data(), sizeof(int) * bps.size(), cudaMemcpyHostToDevice));
        CG_CHECK_CUDA(cudaMalloc(&dst.blk_col, sizeof(int) * (nt + 8)));      // + slack: the tiled step copies 16-byte aligned runs
        CG_CHECK_CUDA(cudaMemset(dst.blk_col, 0, sizeof(int) * (nt + 8)));
        CG_CHECK_CUDA(cudaMalloc(&dst.blk_w, sizeof(float4) * nt));
        CG_CHECK_CUDA(cudaMemcpy(dst.blk_ptr, bptr.data(), sizeof(int) * (size_t)(nblk + 1), cudaMemcpyHostToDevice));
        CG_CHECK_CUDA(cudaMemcpy(dst.blk_order, border.data(), sizeof(int) * (size_t)nblk, cudaMemcpyHostToDevice));
        if (!bcol.empty()) {
            CG_CHECK_CUDA(cudaMemcpy(dst.blk_col, bcol.data(), sizeof(int) * bcol.size(), cudaMemcpyHostToDevice));
            CG_CHECK_CUDA(cudaMemcpy(dst.blk_w, bw.data(), sizeof(float4) * bw.size(), cudaMemcpyHostToDevice));
        }
    }
    return CG_OK;
}

static void free_side(CgCsr &s) {
    cudaFree(s.rowptr);
    cudaFree(s.col);
    cudaFree(s.val);
    cudaFree(s.order);
    cudaFree(s.ell);
    cudaFree(s.blk_ptr);
    cudaFree(s.blk_col);
    cudaFree(s.blk_w);
    cudaFree(s.blk_order);
    cudaFree(s.blk_ps);
    s = CgCsr();
}

extern "C" int cg_graph_create(cg_graph_t **out, int M, int64_t nnz, const int32_t *indptr,
                               const int32_t *indices, const float *values) {
    CG_REQUIRE(out != nullptr, "cg_graph_create: out is NULL");
    *out = nullptr;
    CG_REQUIRE(M > 0, "cg_graph_create: M must be positive (got %d)", M);
    CG_REQUIRE(nnz >= 0 && nnz < (int64_t)INT32_MAX, "cg_graph_create: nnz out of range");
    CG_REQUIRE(indptr && (nnz == 0 || (indices && values)), "cg_graph_create: NULL CSR arrays");
    CG_REQUIRE(indptr[0] == 0 && indptr[M] == nnz, "cg_graph_create: indptr[0] must be 0 and indptr[M] == nnz");

    std::vector<int> rowptr(indptr, indptr + M + 1), col(indices, indices + nnz);
    std::vector<float> val(values, values + nnz);
    for (int m = 0; m < M; ++m) {
        CG_REQUIRE(rowptr[m] <= rowptr[m + 1], "cg_graph_create: indptr not monotone at row %d", m);
        for (int e = rowptr[m]; e < rowptr[m + 1]; ++e) {
            CG_REQUIRE(col[e] >= 0 && col[e] < M, "cg_graph_create: column index %d out of range in row %d", col[e], m);
            CG_REQUIRE(e == rowptr[m] || col[e - 1] < col[e],
                       "cg_graph_create: row %d is not sorted / has duplicates (sum duplicates and sort first)", m);
        }
    }
    // transpose (stable counting sort keeps rows sorted inside every column)
    std::vector<int> t_rowptr(M + 1, 0), t_col(nnz);
    std::vector<float> t_val(nnz);
    for (int64_t e = 0; e < nnz; ++e) t_rowptr[col[e] + 1]++;
    for (int m = 0; m < M; ++m) t_rowptr[m + 1] += t_rowptr[m];
    {
        std::vector<int> fill(t_rowptr.begin(), t_rowptr.end() - 1);
        for (int m = 0; m < M; ++m)
            for (int e = rowptr[m]; e < rowptr[m + 1]; ++e) {
                int p = fill[col[e]]++;
                t_col[p] = m;
                t_val[p] = val[e];
            }
    }
    int w_f = 0, w_t = 0;
    for (int m = 0; m < M; ++m) {
        w_f = std::max(w_f, rowptr[m + 1] - rowptr[m]);
        w_t = std::max(w_t, t_rowptr[m + 1] - t_rowptr[m]);
    }
    size_t m_pad = (size_t)(M + 31) / 32 * 32;
    bool onchip = M <= 65535 && m_pad * (size_t)std::max(w_f, w_t) * sizeof(float2) <= kEllSmemBudget;

    cg_graph *g = new cg_graph();
    g->M = M;
    g->nnz = nnz;
    g->onchip = onchip;
    cudaError_t e = cudaGetDevice(&g->device);
    if (e != cudaSuccess) {
        delete g;
        cg_set_error("cg_graph_create: no CUDA device: %s", cudaGetErrorString(e));
        return CG_ERR_CUDA;
    }
    int v = 0;
    g->sm_count = cg_sm_budget(g->device);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, g->device);
    g->smem_optin = (size_t)v;
    int rc = upload_side(g->fwd, M, nnz, rowptr, col, val, onchip);
    if (rc == CG_OK) rc = upload_side(g->adj, M, nnz, t_rowptr, t_col, t_val, onchip);
    if (rc != CG_OK) {
        free_side(g->fwd);
        free_side(g->adj);
        delete g;
        return rc;
    }
    *out = g;
    return CG_OK;
}

extern "C" int cg_graph_destroy(cg_graph_t *g) {
    if (!g) return CG_OK;
    free_side(g->fwd);
    free_side(g->adj);
    delete g;
    return CG_OK;
}

extern "C" int cg_graph_info(const cg_graph_t *g, int64_t info[5]) {
    CG_REQUIRE(g && info, "cg_graph_info: NULL argument");
    info[0] = g->M;
    info[1] = g->nnz;
    info[2] = g->fwd.width;
    info[3] = g->adj.width;
    info[4] = g->onchip ? 1 : 0;
    return CG_OK;
}
