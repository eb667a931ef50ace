#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules
``/root/reference/lib/graph.py`` and ``/root/reference/lib/coarsening.py``.

Run in the build container (the reference tree does not exist on the GPU
box): ``python tests/golden/make_golden.py``.  The reference files are loaded
by path (they are not a valid package without TensorFlow) with a stub
``matplotlib`` in ``sys.modules`` -- ``graph.py`` only touches ``plt`` inside
``plot_spectrum``.  Nothing is copied from the reference; only its outputs on
seeded inputs are stored.

Recorded with numpy/scipy/sklearn versions printed into ``versions.json``.
"""
import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np
import scipy
import scipy.sparse
import sklearn

REF = '/root/reference/lib'
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    warnings.filterwarnings('ignore')
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', plt)
    mods = {}
    for name in ('graph', 'coarsening'):
        spec = importlib.util.spec_from_file_location('_reference_' + name, os.path.join(REF, name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods['graph'], mods['coarsening']


def csr_parts(prefix, A):
    A = scipy.sparse.csr_matrix(A)
    return {prefix + '_indptr': A.indptr.astype(np.int64), prefix + '_indices': A.indices.astype(np.int64),
            prefix + '_data': A.data, prefix + '_shape': np.array(A.shape, np.int64)}


def silent(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    graph, coarsening = load_reference()

    # ---- known-answer test of the reference itself (lib/coarsening.py:216-217)
    kat = coarsening.compute_perm([np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]), np.array([2, 1, 0, 1, 0])])
    assert kat == [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]

    # ---- C2: MNIST-shaped 28x28 8-NN grid, 4 coarsening levels (nips2016/mnist.ipynb cells 1,3)
    out = {}
    z = graph.grid(28)
    dist, idx = graph.distance_sklearn_metrics(z, k=8, metric='euclidean')
    A = graph.adjacency(dist, idx)
    out.update(csr_parts('A', A))
    out['knn_dist'] = dist
    out['knn_idx'] = idx.astype(np.int64)
    np.random.seed(0)
    rid0 = np.random.permutation(range(A.shape[0]))            # what metis will draw
    np.random.seed(0)
    graphs_m, parents = coarsening.metis(A, 4)
    for i, par in enumerate(parents):
        out['parent%d' % i] = np.asarray(par, np.int64)
    out['rid0'] = rid0.astype(np.int64)
    np.random.seed(0)
    graphs, perm = silent(coarsening.coarsen, A, levels=4, self_connections=False)
    out['perm'] = np.asarray(perm, np.int64)
    perms_all = coarsening.compute_perm(parents)
    for i, pm in enumerate(perms_all):
        out['perm_level%d' % i] = np.asarray(pm, np.int64)
    for i, G in enumerate(graphs):
        out.update(csr_parts('G%d' % i, G))
        L = graph.laplacian(G, normalized=True)
        out.update(csr_parts('L%d' % i, L))
        Lr = graph.rescale_L(scipy.sparse.csr_matrix(L, copy=True), lmax=2)
        out.update(csr_parts('Lr%d' % i, Lr))
    # basis on level 2 (M=248), and perm_data of a small batch
    rng = np.random.RandomState(11)
    Lr2 = scipy.sparse.csr_matrix((out['Lr2_data'], out['Lr2_indices'], out['Lr2_indptr']), shape=tuple(out['Lr2_shape']))
    X = rng.standard_normal((248, 12)).astype(np.float32)
    out['basis_X'] = X
    out['basis_K7'] = graph.chebyshev(Lr2, X, 7)
    out['basis_K1'] = graph.chebyshev(Lr2, X, 1)
    out['basis_K2'] = graph.chebyshev(Lr2, X, 2)
    Lr0 = scipy.sparse.csr_matrix((out['Lr0_data'], out['Lr0_indices'], out['Lr0_indptr']), shape=tuple(out['Lr0_shape']))
    X0 = rng.standard_normal((992, 5)).astype(np.float32)
    out['basis0_X'] = X0
    out['basis0_K25'] = graph.chebyshev(Lr0, X0, 25)
    imgs = rng.uniform(0, 1, (6, 784)).astype(np.float32)
    out['pd_x'] = imgs
    out['pd_y'] = coarsening.perm_data(imgs, perm)
    np.savez_compressed(os.path.join(OUT, 'c2_grid28.npz'), **out)

    # ---- C1: usage.ipynb-shaped feature graph (cells 3,7,9), shrunk to 300 samples
    out = {}
    np.random.seed(1)
    d, n, c = 100, 300, 5
    Xd = np.random.normal(0, 1, (n, d)).astype(np.float32)
    Xd += np.linspace(0, 1, c).repeat(d // c)
    dist, idx = graph.distance_scipy_spatial(Xd.T, k=10, metric='euclidean')
    A = graph.adjacency(dist, idx).astype(np.float32)
    out['Xd'] = Xd
    out['knn_dist'] = dist
    out['knn_idx'] = idx.astype(np.int64)
    out.update(csr_parts('A', A))
    np.random.seed(3)
    graphs, perm = silent(coarsening.coarsen, A, levels=3, self_connections=False)
    out['perm'] = np.asarray(perm, np.int64)
    for i, G in enumerate(graphs):
        out.update(csr_parts('G%d' % i, G))
        L = graph.laplacian(G, normalized=True)
        out.update(csr_parts('L%d' % i, L))
        out.update(csr_parts('Lr%d' % i, graph.rescale_L(scipy.sparse.csr_matrix(L, copy=True), lmax=2)))
    out['pd_y'] = coarsening.perm_data(Xd[:4], perm)
    np.savez_compressed(os.path.join(OUT, 'c1_usage.npz'), **out)

    # ---- directed / non-symmetric operator and lmax != 2 (humantraffic.py:40-44 style edge matrix)
    out = {}
    rng = np.random.RandomState(5)
    M = 57
    dense = (rng.uniform(size=(M, M)) < 0.08) * rng.standard_normal((M, M))
    Ld = scipy.sparse.csr_matrix(dense.astype(np.float32))
    out.update(csr_parts('L', Ld))
    Lr = graph.rescale_L(scipy.sparse.csr_matrix(Ld, copy=True), lmax=3.5)
    out.update(csr_parts('Lr', Lr))
    X = rng.standard_normal((M, 9)).astype(np.float32)
    out['X'] = X
    out['basis_K6'] = graph.chebyshev(scipy.sparse.csr_matrix(Lr), X, 6)
    np.savez_compressed(os.path.join(OUT, 'directed57.npz'), **out)

    with open(os.path.join(OUT, 'versions.json'), 'w') as f:
        json.dump({'numpy': np.__version__, 'scipy': scipy.__version__, 'sklearn': sklearn.__version__,
                   'python': sys.version.split()[0]}, f, indent=1)
    print('golden fixtures written to', OUT)


if __name__ == '__main__':
    main()
