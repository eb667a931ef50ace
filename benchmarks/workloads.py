"""Synthetic workloads of the five BASELINE.json configs (SURVEY.md 8(d) "Synthetic inputs").

Graph construction takes the host library as an argument: the GPU arm passes the product's reference-named modules
(``cnn_graph_b200.lib.graph`` / ``coarsening`` -- native coarsening loops), the CPU / reference arm passes the oracle's
(``oracle.graph_ref`` / ``coarsen_ref`` -- pure numpy / Python), so the reference arm never loads the product library.
Both produce bit-identical graphs (tests/test_host_lib.py, tests/test_oracle_golden.py)."""
import types

import numpy as np
import scipy.sparse


def host_lib(kind):
    if kind == 'product':
        from cnn_graph_b200.lib import coarsening, graph
        return types.SimpleNamespace(graph=graph, coarsen=lambda A, levels: coarsening.coarsen(A, levels=levels, self_connections=False, verbose=False))
    from oracle import coarsen_ref, graph_ref
    return types.SimpleNamespace(graph=graph_ref, coarsen=lambda A, levels: coarsen_ref.coarsen(A, levels, False))


CGCNN = {
    # usage.ipynb cells 3-13: 100-feature kNN graph, 3 coarsening levels, apool1
    'c1': dict(F=[32, 64], K=[20, 20], p=[4, 2], M=[512, 3], pool='apool1', batch=100, classes=3,
               workload='C1 usage.ipynb-shaped: 100-feature 10-NN graph, 3-level coarsening, cgcnn F=[32,64] K=[20,20] '
                        'p=[4,2] M=[512,3] apool1, full training step'),
    # nips2016/mnist.ipynb cells 1,3,14: 28x28 8-NN grid, 4 levels
    'c2': dict(F=[32, 64], K=[25, 25], p=[4, 4], M=[512, 10], pool='mpool1', batch=1024, classes=10,
               workload='C2 MNIST-shaped synthetic: 28x28 8-NN grid graph, 4-level coarsening (M=992), '
                        'cgcnn F=[32,64] K=[25,25] p=[4,4] M=[512,10], full training step'),
    # nips2016/20news.ipynb cells 1,22 at the paper's 10k words
    'c3': dict(F=[32], K=[5], p=[1], M=[20], pool='mpool1', batch=100, classes=20,
               workload='C3 20NEWS-shaped synthetic: 10k-word 16-NN cosine feature graph, cgcnn F=[32] K=[5] p=[1] M=[20], '
                        'sparse bag-of-words input, full training step'),
}
HYPER = dict(regularization=5e-4, dropout=0.5, learning_rate=0.02, decay_rate=0.95, momentum=0.9)


def cgcnn_graphs(config, lib, seed=0):
    """(list of Laplacians per coarsening level, perm) for c1 / c2 / c3."""
    g = lib.graph
    np.random.seed(seed)
    if config == 'c2':
        A = g.adjacency(*g.distance_sklearn_metrics(g.grid(28), k=8, metric='euclidean'))
        levels = 4
    elif config == 'c1':
        d, n, c = 100, 1000, 5
        X = np.random.normal(0, 1, (n, d)).astype(np.float32)
        X += np.linspace(0, 1, c).repeat(d // c)
        A = g.adjacency(*g.distance_scipy_spatial(X.T, k=10, metric='euclidean')).astype(np.float32)
        levels = 3
    elif config == 'c3':
        emb = np.random.normal(0, 1, (10000, 100)).astype(np.float32)
        A = g.adjacency(*g.distance_sklearn_metrics(emb, k=16, metric='cosine')).astype(np.float32)
        levels = 0
    else:
        raise KeyError(config)
    graphs, perm = lib.coarsen(A, levels)
    L = [g.laplacian(G, normalized=True).astype(np.float32) for G in graphs]
    return L, perm


def model_laplacians(L, p):
    """The Laplacian each graph-conv layer runs on (lib/models.py:79-85)."""
    out, j = [], 0
    for pp in p:
        out.append(L[j])
        j += int(np.log2(pp)) if pp > 1 else 0
    return out


def cgcnn_batch(config, L, perm, batch, seed):
    """Raw host batch (what the user feeds) for a cgcnn config: dense [batch, M_raw] float32 signals, int64 labels.
    c3 rows are l1-normalised sparse counts (~0.7 % density, Zipf column popularity), densified like the reference's
    fit() does per batch (lib/graph_model.py:150-151)."""
    rng = np.random.RandomState(seed)
    M_raw = {'c1': 100, 'c2': 784, 'c3': L[0].shape[0]}[config]
    cfg = CGCNN[config]
    if config == 'c3':
        pop = 1.0 / np.arange(1, M_raw + 1) ** 0.8
        pop /= pop.sum()
        x = np.zeros((batch, M_raw), np.float32)
        for i in range(batch):
            cols = rng.choice(M_raw, size=70, replace=False, p=pop)
            x[i, cols] = rng.randint(1, 6, size=70)
        x /= x.sum(axis=1, keepdims=True)
    else:
        x = rng.uniform(0, 1, (batch, M_raw)).astype(np.float32)
    labels = rng.randint(0, cfg['classes'], batch).astype(np.int64)
    return x, labels


def grid32_laplacian(lib):
    g = lib.graph
    A = g.adjacency(*g.distance_sklearn_metrics(g.grid(32), k=8, metric='euclidean'))
    return g.laplacian(A, normalized=True).astype(np.float32)


def morton_order(z, bits=16):
    q = np.minimum((z * (1 << bits)).astype(np.uint64), (1 << bits) - 1)

    def spread(v):
        v = v & 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1), kind='stable')


def knn_graph_laplacian(log2m, k=16, order='morton', seed=2017):
    """C5: 2^log2m points U[0,1)^2, exact k-NN (k-d tree), Gaussian weights as graph.adjacency, symmetrised,
    normalised Laplacian; vertices in Morton order (or a random order, to expose the locality dependence)."""
    import scipy.spatial
    M = 1 << log2m
    rng = np.random.RandomState(seed)
    z = rng.uniform(0, 1, (M, 2)).astype(np.float32)
    z = z[morton_order(z)] if order == 'morton' else z[rng.permutation(M)]
    tree = scipy.spatial.cKDTree(z)
    d, idx = tree.query(z, k=k + 1, workers=-1)
    d, idx = d[:, 1:].astype(np.float32), idx[:, 1:]
    sigma2 = np.mean(d[:, -1]) ** 2
    w = np.exp(-d ** 2 / sigma2)
    W = scipy.sparse.coo_matrix((w.reshape(-1), (np.arange(M).repeat(k), idx.reshape(-1))), shape=(M, M)).tocsr()
    W.setdiag(0)
    bigger = W.T > W
    W = W - W.multiply(bigger) + W.T.multiply(bigger)
    deg = np.asarray(W.sum(axis=0)).ravel() + np.spacing(np.array(0, W.dtype))
    dinv = scipy.sparse.diags((1 / np.sqrt(deg)).astype(np.float32), 0)
    L = scipy.sparse.identity(M, dtype=np.float32) - dinv @ W @ dinv
    return scipy.sparse.csr_matrix(L, dtype=np.float32)
