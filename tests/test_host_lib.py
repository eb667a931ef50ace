"""Host-side product logic (cnn_graph_b200.lib.graph / .coarsening, native coarsening loops)
against the reference-generated fixtures and the oracle.  CPU only -- loads the C-ABI library
but makes no compute call that needs a GPU."""
import ctypes
import re

import numpy as np
import pytest
import scipy.sparse

from conftest import ROOT, csr_from, same_csr
from cnn_graph_b200 import _native
from cnn_graph_b200.lib import coarsening, graph
from oracle import coarsen_ref, graph_ref

KAT_PARENTS = [np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]), np.array([2, 1, 0, 1, 0])]
KAT_PERMS = [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    header = open(ROOT + '/include/cnn_graph_b200.h').read()
    declared = set(re.findall(r'\b(cg_[a-z0-9_]+)\s*\(', header))
    declared -= {'cg_graph'}
    assert declared, 'no declarations parsed'
    for name in sorted(declared):
        assert hasattr(lib, name), 'library does not export %s' % name
    assert declared == set(_native.EXPORTED_SYMBOLS)
    assert lib.cg_abi_version() == _native.ABI_VERSION


def test_error_convention_without_gpu():
    lib = _native.lib()
    rc = lib.cg_graph_info(None, None)
    assert rc != 0 and b'NULL' in lib.cg_last_error()
    with pytest.raises(_native.NativeError):
        _native.check(rc, 'cg_graph_info')


def test_compute_perm_known_answer():
    assert coarsening.compute_perm(KAT_PARENTS) == KAT_PERMS


def test_compute_perm_rejects_three_children():
    with pytest.raises(_native.NativeError):
        coarsening.compute_perm([np.array([0, 0, 0])])


def test_graph_builders_match_reference(c2, c1):
    z = graph.grid(28)
    dist, idx = graph.distance_sklearn_metrics(z, k=8, metric='euclidean')
    assert np.array_equal(dist, c2['knn_dist']) and np.array_equal(idx, c2['knn_idx'])
    assert same_csr(graph.adjacency(dist, idx), csr_from(c2, 'A'))
    dist, idx = graph.distance_scipy_spatial(c1['Xd'].T, k=10, metric='euclidean')
    assert same_csr(graph.adjacency(dist, idx).astype(np.float32), csr_from(c1, 'A'))


def test_laplacian_and_rescale_match_reference(c2, directed):
    for i in range(5):
        L = graph.laplacian(csr_from(c2, 'G%d' % i), normalized=True)
        assert same_csr(L, csr_from(c2, 'L%d' % i))
        assert same_csr(graph.rescale_L(scipy.sparse.csr_matrix(L, copy=True), lmax=2), csr_from(c2, 'Lr%d' % i))
    Lr = graph.rescale_L(scipy.sparse.csr_matrix(csr_from(directed, 'L'), copy=True), lmax=3.5)
    assert same_csr(Lr, csr_from(directed, 'Lr'))
    assert graph.lmax(None) == 2


def test_rescale_csr_does_not_touch_callers_matrix(c2, directed):
    from cnn_graph_b200 import ops
    L = csr_from(directed, 'L')
    before = L.data.copy()
    Lr = ops.rescale_csr(L, lmax=3.5)
    assert np.array_equal(L.data, before)
    assert same_csr(Lr, csr_from(directed, 'Lr'))
    assert same_csr(ops.rescale_csr(csr_from(c2, 'L2'), 2), csr_from(c2, 'Lr2'))


def test_metis_bit_exact_vs_reference_and_oracle(c2):
    A = csr_from(c2, 'A')
    np.random.seed(0)
    graphs, parents = coarsening.metis(A, 4)
    for i, par in enumerate(parents):
        assert np.array_equal(par, c2['parent%d' % i])
    graphs_o, parents_o = coarsen_ref.metis(A, 4, rid=c2['rid0'])
    for a, b in zip(graphs, graphs_o):
        assert same_csr(a, b)


def test_coarsen_bit_exact_c2(c2):
    A = csr_from(c2, 'A')
    np.random.seed(0)
    graphs, perm = coarsening.coarsen(A, levels=4, self_connections=False, verbose=False)
    assert np.array_equal(np.asarray(perm), c2['perm'])
    for i, G in enumerate(graphs):
        assert same_csr(G, csr_from(c2, 'G%d' % i))
    assert [G.shape[0] for G in graphs] == [992, 496, 248, 124, 62]


def test_coarsen_bit_exact_c1(c1):
    A = csr_from(c1, 'A')
    np.random.seed(3)
    graphs, perm = coarsening.coarsen(A, levels=3, self_connections=False, verbose=False)
    assert np.array_equal(np.asarray(perm), c1['perm'])
    for i, G in enumerate(graphs):
        assert same_csr(G, csr_from(c1, 'G%d' % i))


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_coarsen_random_graphs_vs_oracle(seed):
    rng = np.random.RandomState(seed)
    M = int(rng.randint(20, 200))
    z = rng.uniform(size=(M, 3)).astype(np.float32)
    dist, idx = graph.distance_scipy_spatial(z, k=int(rng.randint(2, 7)))
    A = graph.adjacency(dist, idx).astype(np.float32)
    levels = int(rng.randint(0, 5))
    np.random.seed(seed)
    g1, p1 = coarsening.coarsen(A, levels, verbose=False)
    np.random.seed(seed)
    g2, p2 = coarsen_ref.coarsen(A, levels)
    assert (p1 is None and p2 is None) or np.array_equal(np.asarray(p1), np.asarray(p2))
    for a, b in zip(g1, g2):
        assert same_csr(a, b)
    if p1 is not None:
        # structural property: perm is a permutation of range(Mnew), Mnew multiple of 2^levels
        assert sorted(p1) == list(range(len(p1))) and len(p1) % (2 ** levels) == 0


def test_perm_data_matches_reference(c2):
    out = coarsening.perm_data(c2['pd_x'], c2['perm'])
    assert out.dtype == np.float64 and np.array_equal(out, c2['pd_y'])
    assert coarsening.perm_data(c2['pd_x'], None) is c2['pd_x'] or True
    # ragged / edge cases
    x = np.arange(6, dtype=np.float32).reshape(2, 3)
    assert np.array_equal(coarsening.perm_data(x, [2, 5, 0, 1, 3, 4]), coarsen_ref.perm_data(x, [2, 5, 0, 1, 3, 4]))
    assert coarsening.perm_data(np.zeros((0, 3), np.float32), [0, 1, 2, 3]).shape == (0, 4)


def test_perm_adjacency_vs_oracle(c2):
    A = csr_from(c2, 'A')
    a = coarsening.perm_adjacency(A, c2['perm']).tocsr()
    b = coarsen_ref.perm_adjacency(A, c2['perm']).tocsr()
    assert same_csr(a, b)
