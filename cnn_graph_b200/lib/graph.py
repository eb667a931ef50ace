"""Host-side graph numerics with the call surface of the reference's ``lib/graph.py``.

Input producers (``grid``, ``distance_*``, ``adjacency``, ``laplacian``) are numpy/scipy
like the reference and return bit-identical matrices; the hot-path function
``chebyshev`` runs the native CUDA recurrence (no CPU fallback).  ``knn_kdtree`` is an
addition for graphs too large for a dense distance matrix (config C5).
"""
import numpy as np
import scipy.sparse
import scipy.sparse.linalg
import scipy.spatial
import scipy.spatial.distance

__all__ = ['grid', 'distance_scipy_spatial', 'distance_sklearn_metrics', 'knn_kdtree', 'adjacency',
           'replace_random_edges', 'laplacian', 'lmax', 'fourier', 'rescale_L', 'chebyshev']


def grid(m, dtype=np.float32):
    """Coordinates of the m*m vertices of a regular grid in [0,1]^2 (reference lib/graph.py:10-19)."""
    t = np.linspace(0, 1, m, dtype=dtype)
    gx, gy = np.meshgrid(t, t)
    return np.stack([gx.reshape(m * m), gy.reshape(m * m)], axis=1).astype(dtype, copy=False)


def _k_nearest(d, k):
    # same calls as the reference (lib/graph.py:27-30): unstable argsort, self at rank 0 dropped
    idx = np.argsort(d)[:, 1:k + 1]
    d.sort()
    return d[:, 1:k + 1], idx


def distance_scipy_spatial(z, k=4, metric='euclidean'):
    """Exact kNN via scipy pdist (reference lib/graph.py:22-30)."""
    d = scipy.spatial.distance.squareform(scipy.spatial.distance.pdist(z, metric))
    return _k_nearest(d, k)


def distance_sklearn_metrics(z, k=4, metric='euclidean'):
    """Exact kNN via sklearn pairwise distances (reference lib/graph.py:33-41)."""
    import sklearn.metrics
    d = sklearn.metrics.pairwise.pairwise_distances(z, metric=metric, n_jobs=2)
    return _k_nearest(d, k)


def knn_kdtree(z, k=4):
    """Exact euclidean kNN through a k-d tree: O(M log M) memory-light producer for graphs
    with 10^5..10^6 vertices (no reference equivalent; dense pdist is O(M^2))."""
    tree = scipy.spatial.cKDTree(z)
    d, idx = tree.query(z, k=k + 1, workers=-1)
    return d[:, 1:].astype(z.dtype, copy=False), idx[:, 1:]


def adjacency(dist, idx):
    """Symmetric Gaussian-kernel kNN adjacency (reference lib/graph.py:57-83)."""
    M, k = dist.shape
    assert idx.shape == (M, k)
    assert dist.min() >= 0
    sigma2 = np.mean(dist[:, -1]) ** 2
    weight = np.exp(-dist ** 2 / sigma2)
    W = scipy.sparse.coo_matrix((weight.reshape(-1), (np.arange(0, M).repeat(k), idx.reshape(-1))), shape=(M, M))
    W.setdiag(0)
    transposed_wins = W.T > W
    W = W - W.multiply(transposed_wins) + W.T.multiply(transposed_wins)
    assert W.nnz % 2 == 0
    assert np.abs(W - W.T).mean() < 1e-10
    assert scipy.sparse.isspmatrix_csr(W)
    return W


def replace_random_edges(A, noise_level):
    """Swap a fraction of the edges for uniformly random ones (reference lib/graph.py:86-114)."""
    M = A.shape[0]
    n = int(noise_level * A.nnz // 2)
    victims = np.random.permutation(A.nnz // 2)[:n]
    rows = np.random.randint(0, M, n)
    cols = np.random.randint(0, M, n)
    np.random.uniform(0, 1, n)   # the reference draws (unused) weights here; keep the RNG stream aligned
    upper = scipy.sparse.triu(A, format='coo')
    assert upper.nnz == A.nnz // 2 >= n
    A = A.tolil()
    for e, r, c in zip(victims, rows, cols):
        i, j = upper.row[e], upper.col[e]
        A[i, j] = 0
        A[j, i] = 0
        A[r, c] = 1
        A[c, r] = 1
    A.setdiag(0)
    A = A.tocsr()
    A.eliminate_zeros()
    return A


def laplacian(W, normalized=True):
    """Combinatorial or symmetric-normalised Laplacian; degrees are column sums plus the
    smallest subnormal, as in the reference (lib/graph.py:117-136)."""
    deg = W.sum(axis=0)
    if normalized:
        deg += np.spacing(np.array(0, W.dtype))
        deg = 1 / np.sqrt(deg)
        D = scipy.sparse.diags(deg.A.squeeze(), 0)
        L = scipy.sparse.identity(deg.size, dtype=W.dtype) - D * W * D
    else:
        L = scipy.sparse.diags(deg.A.squeeze(), 0) - W
    assert scipy.sparse.isspmatrix_csr(L)
    return L


def lmax(L, normalized=True):
    """Upper bound of the spectrum (reference lib/graph.py:139-145)."""
    if normalized:
        return 2
    return scipy.sparse.linalg.eigsh(L, k=1, which='LM', return_eigenvectors=False)[0]


def fourier(L, algo='eigh', k=1):
    """Graph Fourier basis (reference lib/graph.py:148-166).  Not on the hot path."""
    if algo == 'eigh':
        return np.linalg.eigh(L.toarray())
    if algo == 'eig':
        lamb, U = np.linalg.eig(L.toarray())
    elif algo == 'eigs':
        lamb, U = scipy.sparse.linalg.eigs(L, k=k, which='SM')
    elif algo == 'eigsh':
        return scipy.sparse.linalg.eigsh(L, k=k, which='SM')
    else:
        raise ValueError(algo)
    order = lamb.argsort()
    return lamb[order], U[:, order]


def rescale_L(L, lmax=2):
    """L~ = L / (lmax / 2) - I: spectrum mapped into [-1, 1] (reference lib/graph.py:232-238).

    Same semantics as the reference, including the in-place division of ``L`` (callers that
    share L pass a copy, like lib/models.py:196).
    """
    M = L.shape[0]
    L /= lmax / 2
    L -= scipy.sparse.identity(M, format='csr', dtype=L.dtype)
    return L


def chebyshev(L, X, K):
    """Chebyshev basis ``Xt[k] = T_k(L) X`` as [K, M, N] (reference lib/graph.py:241-258).

    ``L`` is an already rescaled operator; the recurrence runs on the GPU (fused K-step
    SpMM kernel).  numpy in -> numpy out; torch CUDA tensor in -> torch tensor out.
    """
    import torch
    from .. import ops
    M, N = X.shape
    assert L.dtype == X.dtype
    handle = ops.GraphHandle(L)
    if isinstance(X, np.ndarray):
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()
        return ops.cheb_basis(handle, Xd, K).cpu().numpy()
    return ops.cheb_basis(handle, X, K)
