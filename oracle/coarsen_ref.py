"""Oracle restatement of the reference's Graclus-style coarsening and of the
pooling permutation.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Plain Python loops on purpose: this is the slow, obviously-right checker for
``cnn_graph_b200.lib.coarsening`` (whose loops are native).  Bit-exactness with
the reference holds for the same numpy/scipy on the same host: the reference
draws ``rid`` from numpy's global RNG and uses unstable argsorts on tied keys
(lib/coarsening.py:55-56,78,113), and its float32 scalar arithmetic follows
numpy's promotion rules.  Pinned by ``tests/golden/coarsen_*.npz`` and by the
reference's own known-answer test (lib/coarsening.py:216-217).
"""
import numpy as np
import scipy.sparse


def metis_one_level(rr, cc, vv, rid, weights):
    """One greedy matching pass.  lib/coarsening.py:119-165.

    ``rr`` must be sorted.  The row table is built exactly like the reference
    builds it, INCLUDING its off-by-one: the first entry of row r+1 is counted
    in ``rowlength`` of row r (so the first row scans one foreign entry and the
    last row scans one entry fewer), and rows are numbered by order of
    appearance, not by id.
    """
    nnz = rr.shape[0]
    N = rr[nnz - 1] + 1
    marked = np.zeros(N, bool)
    rowstart = np.zeros(N, np.int32)
    rowlength = np.zeros(N, np.int32)
    cluster_id = np.zeros(N, np.int32)

    seen = rr[0]
    r = 0
    for e in range(nnz):
        rowlength[r] += 1
        if rr[e] > seen:
            seen = rr[e]
            rowstart[r + 1] = e
            r += 1

    nclusters = 0
    for t in range(N):
        v = rid[t]
        if marked[v]:
            continue
        marked[v] = True
        best, best_w = -1, 0.0
        base = rowstart[v]
        for j in range(rowlength[v]):
            u = cc[base + j]
            if marked[u]:
                w = 0.0
            else:
                w = vv[base + j] * (1.0 / weights[v] + 1.0 / weights[u])
            if w > best_w:
                best_w, best = w, u
        cluster_id[v] = nclusters
        if best > -1:
            cluster_id[best] = nclusters
            marked[best] = True
        nclusters += 1
    return cluster_id


def metis(W, levels, rid=None):
    """Multilevel coarsening.  lib/coarsening.py:34-115."""
    N = W.shape[0]
    if rid is None:
        rid = np.random.permutation(range(N))
    parents, graphs = [], [W]
    degree = W.sum(axis=0) - W.diagonal()
    for _ in range(levels):
        weights = np.array(degree).squeeze()
        row, col, val = scipy.sparse.find(W)
        order = np.argsort(row)
        rr, cc, vv = row[order], col[order], val[order]
        cluster_id = metis_one_level(rr, cc, vv, rid, weights)
        parents.append(cluster_id)
        Nnew = cluster_id.max() + 1
        W = scipy.sparse.csr_matrix((vv, (cluster_id[rr], cluster_id[cc])), shape=(Nnew, Nnew))
        W.eliminate_zeros()
        graphs.append(W)
        degree = W.sum(axis=0)
        rid = np.argsort(np.array(W.sum(axis=0)).squeeze())
    return graphs, parents


def compute_perm(parents):
    """Binary-tree vertex ordering with fake nodes.  lib/coarsening.py:167-214."""
    orders = []
    if len(parents) > 0:
        orders.append(list(range(max(parents[-1]) + 1)))
    for parent in reversed(parents):
        next_fake = len(parent)
        layer = []
        for node in orders[-1]:
            kids = list(np.where(parent == node)[0])
            assert len(kids) <= 2
            while len(kids) < 2:          # singleton -> 1 fake; fake parent -> 2 fakes
                kids.append(next_fake)
                next_fake += 1
            layer.extend(kids)
        orders.append(layer)
    for i, layer in enumerate(orders):
        assert sorted(layer) == list(range(len(orders[0]) * 2 ** i))
    return orders[::-1]


def perm_data(x, indices):
    """Gather along the vertex axis, zero-fill fakes; float64 out.  lib/coarsening.py:219-240."""
    if indices is None:
        return x
    N, M = x.shape
    assert len(indices) >= M
    out = np.empty((N, len(indices)))
    for i, j in enumerate(indices):
        out[:, i] = x[:, j] if j < M else 0.0
    return out


def perm_adjacency(A, indices):
    """Pad with isolated vertices and relabel.  lib/coarsening.py:242-269."""
    if indices is None:
        return A
    M = A.shape[0]
    Mnew = len(indices)
    assert Mnew >= M
    A = A.tocoo()
    if Mnew > M:
        A = scipy.sparse.vstack([A, scipy.sparse.coo_matrix((Mnew - M, M), dtype=np.float32)])
        A = scipy.sparse.hstack([A, scipy.sparse.coo_matrix((Mnew, Mnew - M), dtype=np.float32)])
    where = np.argsort(indices)
    A.row = np.array(where)[A.row]
    A.col = np.array(where)[A.col]
    return A


def coarsen(A, levels, self_connections=False):
    """graphs, perm = coarsen(A, levels).  lib/coarsening.py:5-31 (prints dropped)."""
    graphs, parents = metis(A, levels)
    perms = compute_perm(parents)
    for i, G in enumerate(graphs):
        if not self_connections:
            G = G.tocoo()
            G.setdiag(0)
        if i < levels:
            G = perm_adjacency(G, perms[i])
        G = G.tocsr()
        G.eliminate_zeros()
        graphs[i] = G
    return graphs, perms[0] if levels > 0 else None
