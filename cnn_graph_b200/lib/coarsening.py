"""Graph coarsening and the pooling permutation, with the call surface of the
reference's ``lib/coarsening.py``.

The graph algebra between levels uses the same numpy / scipy calls as the reference
(so ties in its unstable argsorts and the order of duplicate summation resolve
identically on the same host), while the two Python loops that dominate its run time
-- the greedy matching and the O(M^2) child lookup of ``compute_perm`` -- run in the
native library (``cg_host_metis_one_level``, ``cg_host_perm_level``).
Results are bit-identical to the reference (tests/test_coarsening.py).
"""
import ctypes

import numpy as np
import scipy.sparse

from .. import _native

__all__ = ['coarsen', 'metis', 'metis_one_level', 'compute_perm', 'perm_data', 'perm_adjacency']


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def metis_one_level(rr, cc, vv, rid, weights):
    """Greedy heavy-edge matching of one level (reference lib/coarsening.py:119-165).

    Native loop; float32 arithmetic and the reference's row-table convention are kept so
    the cluster ids are identical.
    """
    rr, cc, rid = _i64(rr), _i64(cc), _i64(rid)
    vv = np.ascontiguousarray(vv, dtype=np.float32)
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    nnz = rr.shape[0]
    if nnz == 0:
        raise ValueError('metis_one_level: empty graph')
    N = int(rr[-1]) + 1
    if weights.shape[0] < N:
        raise ValueError('metis_one_level: weights shorter than the vertex count')
    cluster_id = np.zeros(N, np.int32)
    count = ctypes.c_int64(0)
    _native.check(_native.lib().cg_host_metis_one_level(
        nnz, rr.ctypes.data, cc.ctypes.data, vv.ctypes.data, rid.ctypes.data, rid.shape[0],
        weights.ctypes.data, cluster_id.ctypes.data, ctypes.byref(count)), 'cg_host_metis_one_level')
    return cluster_id


def metis(W, levels, rid=None):
    """Multilevel Graclus-style coarsening (reference lib/coarsening.py:34-115).

    Returns (graphs, parents): ``graphs[i]`` is the weight matrix of level i and
    ``parents[i][v]`` the cluster of vertex v in level i+1.
    """
    N = W.shape[0]
    if rid is None:
        rid = np.random.permutation(range(N))    # global RNG, as the reference (:55-56)
    graphs, parents = [W], []
    degree = W.sum(axis=0) - W.diagonal()
    for _ in range(levels):
        weights = np.array(degree).squeeze()
        # edge list sorted by row -- the same find + argsort pair as the reference (:76-81)
        row, col, val = scipy.sparse.find(W)
        by_row = np.argsort(row)
        rr, cc, vv = row[by_row], col[by_row], val[by_row]
        cluster_id = metis_one_level(rr, cc, vv, rid, weights)
        parents.append(cluster_id)
        n_coarse = int(cluster_id.max()) + 1
        W = scipy.sparse.csr_matrix((vv, (cluster_id[rr], cluster_id[cc])), shape=(n_coarse, n_coarse))
        W.eliminate_zeros()
        graphs.append(W)
        degree = W.sum(axis=0)
        rid = np.argsort(np.array(W.sum(axis=0)).squeeze())   # visit light vertices first (:112-113)
    return graphs, parents


def compute_perm(parents):
    """Vertex orderings that turn the cluster hierarchy into a balanced binary tree
    (reference lib/coarsening.py:167-214): ``result[i]`` lists, for level i, real vertex ids
    and fake ids (>= level size) such that positions 2j, 2j+1 are the children of position j
    of level i+1."""
    orders = []
    if len(parents) > 0:
        orders.append(np.arange(int(np.max(parents[-1])) + 1, dtype=np.int64))
    lib = _native.lib()
    for parent in parents[::-1]:
        parent = _i64(parent)
        above = orders[-1]
        layer = np.empty(2 * above.shape[0], np.int64)
        _native.check(lib.cg_host_perm_level(parent.ctypes.data, parent.shape[0], above.ctypes.data,
                                             above.shape[0], layer.ctypes.data), 'cg_host_perm_level')
        orders.append(layer)
    n_top = orders[0].shape[0] if orders else 0
    for i, layer in enumerate(orders):
        # every id exactly once: a true permutation of range(M_last * 2^i)
        if not np.array_equal(np.sort(layer), np.arange(n_top * 2 ** i)):
            raise AssertionError('compute_perm: level %d is not a permutation' % i)
    return [layer.tolist() for layer in orders[::-1]]


def perm_data(x, indices):
    """Reorder the vertex axis of a data matrix and zero-fill fake vertices
    (reference lib/coarsening.py:219-240).  Host numpy in, float64 out like the reference;
    the device-side equivalent is ``cnn_graph_b200.ops.perm_data_device``."""
    if indices is None:
        return x
    N, M = x.shape
    idx = np.asarray(indices, dtype=np.int64)
    Mnew = idx.shape[0]
    assert Mnew >= M
    real = idx < M
    out = np.zeros((N, Mnew))
    out[:, real] = x[:, idx[real]]
    return out


def perm_adjacency(A, indices):
    """Pad an adjacency with isolated (fake) vertices and relabel it with the ordering
    (reference lib/coarsening.py:242-269)."""
    if indices is None:
        return A
    M = A.shape[0]
    Mnew = len(indices)
    assert Mnew >= M
    A = A.tocoo()
    if Mnew > M:
        A = scipy.sparse.vstack([A, scipy.sparse.coo_matrix((Mnew - M, M), dtype=np.float32)])
        A = scipy.sparse.hstack([A, scipy.sparse.coo_matrix((Mnew, Mnew - M), dtype=np.float32)])
    position = np.argsort(indices)
    A.row = np.array(position)[A.row]
    A.col = np.array(position)[A.col]
    assert scipy.sparse.isspmatrix_coo(A)
    return A


def coarsen(A, levels, self_connections=False, verbose=True):
    """Coarsen ``A`` ``levels`` times; returns (graphs, perm) with every graph already
    permuted/padded and ``perm`` the ordering of the finest level
    (reference lib/coarsening.py:5-31)."""
    graphs, parents = metis(A, levels)
    perms = compute_perm(parents)
    for i, G in enumerate(graphs):
        M = G.shape[0]
        if not self_connections:
            G = G.tocoo()
            G.setdiag(0)
        if i < levels:
            G = perm_adjacency(G, perms[i])
        G = G.tocsr()
        G.eliminate_zeros()
        graphs[i] = G
        if verbose:
            Mnew = G.shape[0]
            print('Layer {0}: M_{0} = |V| = {1} nodes ({2} added),|E| = {3} edges'.format(i, Mnew, Mnew - M, G.nnz // 2))
    return graphs, perms[0] if levels > 0 else None
