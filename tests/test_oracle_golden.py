"""Pin the oracle (oracle/graph_ref.py, oracle/coarsen_ref.py) against the fixtures produced
by the UNMODIFIED reference modules (tests/golden/make_golden.py) and against the reference's
only known-answer test (lib/coarsening.py:216-217).  CPU only."""
import contextlib
import io

import numpy as np
import scipy.sparse

from conftest import csr_from, same_csr
from oracle import coarsen_ref, graph_ref

KAT_PARENTS = [np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]), np.array([2, 1, 0, 1, 0])]
KAT_PERMS = [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]


def test_compute_perm_known_answer():
    assert [list(map(int, p)) for p in coarsen_ref.compute_perm(KAT_PARENTS)] == KAT_PERMS


def test_grid_knn_adjacency_c2(c2):
    z = graph_ref.grid(28)
    dist, idx = graph_ref.distance_sklearn_metrics(z, k=8, metric='euclidean')
    assert np.array_equal(dist, c2['knn_dist'])
    assert np.array_equal(idx, c2['knn_idx'])
    assert same_csr(graph_ref.adjacency(dist, idx), csr_from(c2, 'A'))


def test_coarsen_c2_bit_exact(c2):
    A = csr_from(c2, 'A')
    np.random.seed(0)
    graphs, parents = coarsen_ref.metis(A, 4)
    for i, par in enumerate(parents):
        assert np.array_equal(par, c2['parent%d' % i])
    np.random.seed(0)
    graphs, perm = coarsen_ref.coarsen(A, levels=4, self_connections=False)
    assert np.array_equal(np.asarray(perm), c2['perm'])
    for i, G in enumerate(graphs):
        assert same_csr(G, csr_from(c2, 'G%d' % i))
        L = graph_ref.laplacian(G, normalized=True)
        assert same_csr(L, csr_from(c2, 'L%d' % i))
        Lr = graph_ref.rescale_L(scipy.sparse.csr_matrix(L, copy=True), lmax=2)
        assert same_csr(Lr, csr_from(c2, 'Lr%d' % i))


def test_coarsen_c1_bit_exact(c1):
    dist, idx = graph_ref.distance_scipy_spatial(c1['Xd'].T, k=10, metric='euclidean')
    assert np.array_equal(dist, c1['knn_dist']) and np.array_equal(idx, c1['knn_idx'])
    A = graph_ref.adjacency(dist, idx).astype(np.float32)
    assert same_csr(A, csr_from(c1, 'A'))
    np.random.seed(3)
    graphs, perm = coarsen_ref.coarsen(A, levels=3, self_connections=False)
    assert np.array_equal(np.asarray(perm), c1['perm'])
    for i, G in enumerate(graphs):
        assert same_csr(G, csr_from(c1, 'G%d' % i))
    assert np.array_equal(coarsen_ref.perm_data(c1['Xd'][:4], perm), c1['pd_y'])


def test_chebyshev_basis_matches_reference(c2, directed):
    Lr2 = csr_from(c2, 'Lr2')
    for K in (1, 2, 7):
        got = graph_ref.chebyshev(Lr2, c2['basis_X'], K)
        assert got.dtype == np.float32 and np.array_equal(got, c2['basis_K%d' % K])
    assert np.array_equal(graph_ref.chebyshev(csr_from(c2, 'Lr0'), c2['basis0_X'], 25), c2['basis0_K25'])
    Lr = graph_ref.rescale_L(scipy.sparse.csr_matrix(csr_from(directed, 'L'), copy=True), lmax=3.5)
    assert same_csr(Lr, csr_from(directed, 'Lr'))
    assert np.array_equal(graph_ref.chebyshev(scipy.sparse.csr_matrix(Lr), directed['X'], 6), directed['basis_K6'])


def test_perm_data_matches_reference(c2):
    out = coarsen_ref.perm_data(c2['pd_x'], c2['perm'])
    assert out.dtype == np.float64 and np.array_equal(out, c2['pd_y'])
