"""Oracle restatement of the numpy graph numerics of the reference.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites
the reference lines it restates (paths relative to /root/reference).  The
restatement keeps the reference's arithmetic order and dtypes so results are
bit-identical to the reference run on the same numpy/scipy; it is pinned by
``tests/golden/graph_*.npz``.
"""
import numpy as np
import scipy.sparse
import scipy.spatial.distance
import sklearn.metrics


def grid(m, dtype=np.float32):
    """2-D embedding of an m x m grid.  lib/graph.py:10-19."""
    ticks = np.linspace(0, 1, m, dtype=dtype)
    xx, yy = np.meshgrid(ticks, ticks)
    out = np.empty((m * m, 2), dtype)
    out[:, 0] = xx.reshape(-1)
    out[:, 1] = yy.reshape(-1)
    return out


def _knn_from_dense(d, k):
    # lib/graph.py:27-30 / 38-41: unstable argsort of every row, drop self.
    idx = np.argsort(d)[:, 1:k + 1]
    d.sort()
    return d[:, 1:k + 1], idx


def distance_scipy_spatial(z, k=4, metric='euclidean'):
    """Exact kNN through pdist.  lib/graph.py:22-30."""
    d = scipy.spatial.distance.squareform(scipy.spatial.distance.pdist(z, metric))
    return _knn_from_dense(d, k)


def distance_sklearn_metrics(z, k=4, metric='euclidean'):
    """Exact kNN through sklearn pairwise distances.  lib/graph.py:33-41."""
    d = sklearn.metrics.pairwise.pairwise_distances(z, metric=metric, n_jobs=2)
    return _knn_from_dense(d, k)


def adjacency(dist, idx):
    """Gaussian-weighted symmetric kNN adjacency.  lib/graph.py:57-83."""
    M, k = dist.shape
    assert dist.min() >= 0
    sigma2 = np.mean(dist[:, -1]) ** 2
    w = np.exp(-dist ** 2 / sigma2)
    rows = np.arange(0, M).repeat(k)
    W = scipy.sparse.coo_matrix((w.reshape(M * k), (rows, idx.reshape(M * k))), shape=(M, M))
    W.setdiag(0)
    # keep the larger of W[i, j], W[j, i] on both sides
    bigger = W.T > W
    W = W - W.multiply(bigger) + W.T.multiply(bigger)
    assert W.nnz % 2 == 0
    assert np.abs(W - W.T).mean() < 1e-10
    return W


def laplacian(W, normalized=True):
    """Graph Laplacian; degrees are COLUMN sums.  lib/graph.py:117-136."""
    d = W.sum(axis=0)
    if not normalized:
        return scipy.sparse.diags(d.A.squeeze(), 0) - W
    d += np.spacing(np.array(0, W.dtype))
    d = 1 / np.sqrt(d)
    D = scipy.sparse.diags(d.A.squeeze(), 0)
    I = scipy.sparse.identity(d.size, dtype=W.dtype)
    return I - D * W * D


def lmax(L, normalized=True):
    """Spectrum upper bound.  lib/graph.py:139-145."""
    if normalized:
        return 2
    return scipy.sparse.linalg.eigsh(L, k=1, which='LM', return_eigenvectors=False)[0]


def rescale_L(L, lmax=2):
    """L~ = L / (lmax/2) - I, spectrum in [-1, 1].  lib/graph.py:232-238.

    Like the reference this divides IN PLACE (``L /= ...``) and then rebinds on
    the subtraction, so callers that care pass a copy.
    """
    M = L.shape[0]
    I = scipy.sparse.identity(M, format='csr', dtype=L.dtype)
    L /= lmax / 2
    L -= I
    return L


def chebyshev(L, X, K):
    """Chebyshev basis T_k(L) X, k < K, as [K, M, N].  lib/graph.py:241-258."""
    M, N = X.shape
    assert L.dtype == X.dtype
    Xt = np.empty((K, M, N), L.dtype)
    Xt[0] = X
    if K > 1:
        Xt[1] = L.dot(X)
    for k in range(2, K):
        Xt[k] = 2 * L.dot(Xt[k - 1]) - Xt[k - 2]
    return Xt


def fourier(L):
    """Dense eigendecomposition (algo='eigh').  lib/graph.py:148-166."""
    return np.linalg.eigh(L.toarray())
