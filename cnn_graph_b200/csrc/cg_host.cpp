// Host-side native loops of the coarsening: the two pure-Python
// loops of the reference that make large graphs impractical.  The match score
// vv * (1/w[v] + 1/w[u]) is evaluated in the reference's order, no fast-math, in the dtype numpy
// gives it: float32 for a float32 adjacency under numpy >= 2 (weak Python scalars; what the pinned
// fixtures of tests/golden were generated with), float64 for a float64 adjacency, and -- the
// `_f64` entry fed with widened float32 inputs -- float64 under the legacy value-based promotion of
// the numpy 1.x the reference was written for.  Parity is against the oracle port and the
// reference-generated fixtures (tests/test_host_lib.py).
#include <stdint.h>

#include <vector>

#include "../../include/cnn_graph_b200.h"

void cg_set_error(const char *fmt, ...);

// lib/coarsening.py:119-165
template <typename T>
static int metis_one_level(int64_t nnz, const int64_t *rr, const int64_t *cc, const T *vv, const int64_t *rid,
                           int64_t n_rid, const T *weights, int32_t *cluster_id, int64_t *nclusters) {
    if (nnz <= 0 || !rr || !cc || !vv || !rid || !weights || !cluster_id || !nclusters) {
        cg_set_error("cg_host_metis_one_level: bad arguments");
        return CG_ERR_ARG;
    }
    const int64_t N = rr[nnz - 1] + 1;
    if (n_rid < N) {
        cg_set_error("cg_host_metis_one_level: rid has %lld entries, need %lld", (long long)n_rid, (long long)N);
        return CG_ERR_ARG;
    }
    std::vector<char> marked((size_t)N, 0);
    std::vector<int64_t> rowstart((size_t)N, 0), rowlength((size_t)N, 0);
    // Row table exactly as the reference builds it: the entry that opens row r+1 is still
    // counted for row r, and rows are numbered by order of appearance.
    int64_t seen = rr[0], r = 0;
    for (int64_t e = 0; e < nnz; ++e) {
        rowlength[(size_t)r] += 1;
        if (rr[e] > seen) {
            seen = rr[e];
            if (r + 1 >= N) {
                cg_set_error("cg_host_metis_one_level: rr is not sorted");
                return CG_ERR_ARG;
            }
            rowstart[(size_t)(r + 1)] = e;
            r += 1;
        }
    }
    for (int64_t i = 0; i < N; ++i) cluster_id[i] = 0;
    int64_t count = 0;
    for (int64_t t = 0; t < N; ++t) {
        const int64_t v = rid[t];
        if (v < 0 || v >= N) {
            cg_set_error("cg_host_metis_one_level: rid[%lld]=%lld out of range", (long long)t, (long long)v);
            return CG_ERR_ARG;
        }
        if (marked[(size_t)v]) continue;
        marked[(size_t)v] = 1;
        T best_w = (T)0;
        int64_t best = -1;
        const int64_t base = rowstart[(size_t)v];
        for (int64_t j = 0; j < rowlength[(size_t)v]; ++j) {
            if (base + j >= nnz) {
                cg_set_error("cg_host_metis_one_level: row table overruns the edge list");
                return CG_ERR_ARG;
            }
            const int64_t u = cc[base + j];
            if (u < 0 || u >= N) {
                cg_set_error("cg_host_metis_one_level: column %lld out of range", (long long)u);
                return CG_ERR_ARG;
            }
            T w;
            if (marked[(size_t)u]) {
                w = (T)0;
            } else {
                const T inv_v = (T)1 / weights[v];
                const T inv_u = (T)1 / weights[u];
                const T s = inv_v + inv_u;
                w = vv[base + j] * s;
            }
            if (w > best_w) {
                best_w = w;
                best = u;
            }
        }
        cluster_id[v] = (int32_t)count;
        if (best > -1) {
            cluster_id[best] = (int32_t)count;
            marked[(size_t)best] = 1;
        }
        count += 1;
    }
    *nclusters = count;
    return CG_OK;
}

extern "C" int cg_host_metis_one_level(int64_t nnz, const int64_t *rr, const int64_t *cc, const float *vv,
                                       const int64_t *rid, int64_t n_rid, const float *weights,
                                       int32_t *cluster_id, int64_t *nclusters) {
    return metis_one_level<float>(nnz, rr, cc, vv, rid, n_rid, weights, cluster_id, nclusters);
}

extern "C" int cg_host_metis_one_level_f64(int64_t nnz, const int64_t *rr, const int64_t *cc, const double *vv,
                                           const int64_t *rid, int64_t n_rid, const double *weights,
                                           int32_t *cluster_id, int64_t *nclusters) {
    return metis_one_level<double>(nnz, rr, cc, vv, rid, n_rid, weights, cluster_id, nclusters);
}

// lib/coarsening.py:179-204, one level
extern "C" int cg_host_perm_level(const int64_t *parent, int64_t n_parent, const int64_t *order, int64_t n_order,
                                  int64_t *out) {
    if (!parent || !order || !out || n_parent < 0 || n_order < 0) {
        cg_set_error("cg_host_perm_level: bad arguments");
        return CG_ERR_ARG;
    }
    int64_t top = -1;
    for (int64_t i = 0; i < n_parent; ++i) {
        if (parent[i] < 0) {
            cg_set_error("cg_host_perm_level: negative parent id");
            return CG_ERR_ARG;
        }
        if (parent[i] > top) top = parent[i];
    }
    std::vector<int64_t> kid((size_t)(2 * (top + 1)), -1);
    for (int64_t i = 0; i < n_parent; ++i) {          // ascending i == np.where order
        int64_t *slot = &kid[(size_t)(2 * parent[i])];
        if (slot[0] < 0) {
            slot[0] = i;
        } else if (slot[1] < 0) {
            slot[1] = i;
        } else {
            cg_set_error("cg_host_perm_level: cluster %lld has more than two children", (long long)parent[i]);
            return CG_ERR_ARG;
        }
    }
    int64_t fake = n_parent;
    for (int64_t i = 0; i < n_order; ++i) {
        const int64_t node = order[i];
        int64_t a = -1, b = -1;
        if (node >= 0 && node <= top) {
            a = kid[(size_t)(2 * node)];
            b = kid[(size_t)(2 * node + 1)];
        }
        if (a < 0) a = fake++;
        if (b < 0) b = fake++;
        out[2 * i] = a;
        out[2 * i + 1] = b;
    }
    return CG_OK;
}
