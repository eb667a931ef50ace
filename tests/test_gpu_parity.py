"""Parity of the CUDA hot path (through the C ABI) against the oracle on identical seeded
inputs, against the reference-generated golden fixtures, and -- at full BASELINE sizes --
through size-independent properties (linearity, adjoint identities).

Tolerances (BASELINE.json north_star): bit-exact for permutations, perm_data and pooling
indices; fp32 filter outputs and gradients within rtol 1e-4 with atol = 1e-4 * max|ref|.
"""
import numpy as np
import pytest
import scipy.sparse
import torch

from conftest import csr_from

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def close(got, ref, rtol=RTOL):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-30) if ref.size else 1.0
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) if ref.size else 0.0
    assert err <= rtol * scale, 'max abs err %.3e > %.1e * %.3e' % (err, rtol, scale)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope='module')
def ops():
    from cnn_graph_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


@pytest.fixture(scope='module')
def tf_ref():
    from oracle import tf_ref as t
    return t


@pytest.fixture(params=['rows', 'blocks'])
def gather(request):
    """Both gather forms of the fused recurrence kernels: one row per item (k_cheb_fused / k_cheb_clenshaw) and the
    4-row union blocks (k_cheb_fused_b / k_cheb_clenshaw_b); the library picks by the operator's locality unless
    CG_FUSED_BLOCK says otherwise."""
    import os
    os.environ['CG_FUSED_BLOCK'] = '1' if request.param == 'blocks' else '0'
    yield request.param
    os.environ.pop('CG_FUSED_BLOCK', None)


@pytest.fixture(params=['csr', 'blocks'])
def spmm_form(request):
    """Both forms of the streaming recurrence step: one row per thread group (k_spmm_step) and 4-row union blocks
    (k_spmm_step_b); the library picks by the operator's locality unless CG_SPMM_BLOCK says otherwise."""
    import os
    os.environ['CG_SPMM_BLOCK'] = '1' if request.param == 'blocks' else '0'
    yield request.param
    os.environ.pop('CG_SPMM_BLOCK', None)


FLAG_SETS = [0, 1, 4]   # default (fused / on-chip when it fits), forced streaming, on-chip without the fused kernel


# --------------------------------------------------------------------------- basis
@pytest.mark.parametrize('flags', FLAG_SETS)
def test_basis_matches_reference_fixtures(ops, c2, directed, flags):
    h2 = ops.GraphHandle(csr_from(c2, 'Lr2'))
    assert h2.info()['onchip']
    for K in (1, 2, 7):
        got = ops.cheb_basis(h2, dev(c2['basis_X']), K, flags=flags)
        close(got, c2['basis_K%d' % K], 1e-5)
    h0 = ops.GraphHandle(csr_from(c2, 'Lr0'))
    close(ops.cheb_basis(h0, dev(c2['basis0_X'][:, :4]), 25, flags=flags), c2['basis0_K25'][:, :, :4], 1e-5)
    close(ops.cheb_basis(h0, dev(c2['basis0_X']), 25, flags=flags), c2['basis0_K25'], 1e-5)    # C = 5: scalar path
    hd = ops.GraphHandle(csr_from(directed, 'Lr'))
    close(ops.cheb_basis(hd, dev(directed['X']), 6, flags=flags), directed['basis_K6'], 1e-5)


def test_streaming_step_forms_vs_reference(ops, c2, c1, directed, spmm_form):
    """Forced streaming path (one launch per recurrence step) in both step forms: reference fixtures of the basis, a
    directed operator on both sides, a vertex count that is not a multiple of 4, wide and narrow slabs."""
    from oracle import graph_ref
    h0 = ops.GraphHandle(csr_from(c2, 'Lr0'))
    close(ops.cheb_basis(h0, dev(c2['basis0_X'][:, :4]), 25, flags=1), c2['basis0_K25'][:, :, :4], 1e-5)
    h2 = ops.GraphHandle(csr_from(c2, 'Lr2'))
    close(ops.cheb_basis(h2, dev(c2['basis_X']), 7, flags=1), c2['basis_K7'], 1e-5)
    Lr = csr_from(directed, 'Lr')           # M = 57
    hd = ops.GraphHandle(Lr)
    rng = np.random.RandomState(3)
    for C in (4, 64, 520):
        X = rng.standard_normal((Lr.shape[0], C)).astype(np.float32)
        close(ops.cheb_basis(hd, dev(X), 6, flags=1), graph_ref.chebyshev(Lr, X, 6), 1e-5)
        close(ops.cheb_basis(hd, dev(X), 6, transpose=True, flags=1), graph_ref.chebyshev(scipy.sparse.csr_matrix(Lr.T), X, 6), 1e-5)
    L1 = csr_from(c1, 'Lr1')
    X = rng.standard_normal((L1.shape[0], 96)).astype(np.float32)
    close(ops.cheb_basis(ops.GraphHandle(L1), dev(X), 9, flags=1), graph_ref.chebyshev(L1, X, 9), 1e-5)


def test_basis_transpose_and_wide(ops, directed):
    from oracle import graph_ref
    Lr = csr_from(directed, 'Lr')
    h = ops.GraphHandle(Lr)
    rng = np.random.RandomState(0)
    X = rng.standard_normal((Lr.shape[0], 520)).astype(np.float32)
    ref = graph_ref.chebyshev(scipy.sparse.csr_matrix(Lr.T), X, 9)
    for flags in FLAG_SETS:
        close(ops.cheb_basis(h, dev(X), 9, transpose=True, flags=flags), ref, 1e-5)


def test_lib_graph_chebyshev_host_api(c2):
    from cnn_graph_b200.lib import graph
    got = graph.chebyshev(csr_from(c2, 'Lr2'), c2['basis_X'], 7)
    assert isinstance(got, np.ndarray) and got.dtype == np.float32
    close(got, c2['basis_K7'], 1e-5)


# --------------------------------------------------------------------------- filter
SHAPES = [  # (level, N, Fin, Fout, K)
    (2, 5, 32, 64, 25), (0, 3, 1, 32, 25), (3, 7, 1, 3, 1), (3, 2, 3, 5, 2), (4, 9, 6, 10, 3), (3, 4, 33, 130, 4),
    (2, 1, 64, 2, 5), (4, 16, 2, 512, 3), (3, 1, 8, 8, 20), (4, 33, 5, 7, 6),
]


@pytest.mark.parametrize('flags', FLAG_SETS)
@pytest.mark.parametrize('level,N,Fin,Fout,K', SHAPES)
def test_filter_forward_backward_vs_oracle(ops, tf_ref, c2, level, N, Fin, Fout, K, flags):
    L = csr_from(c2, 'L%d' % level)
    M = L.shape[0]
    rng = np.random.RandomState(100 * level + K)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    xt = dev(x).requires_grad_(True)
    Wt = dev(W).requires_grad_(True)
    y = ops.cheb_filter(xt, Wt, L, K, lmax=2, flags=flags)
    close(y, tf_ref.chebyshev5(x, L, W, K))
    y.backward(dev(gy))
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
    close(xt.grad, dx)
    close(Wt.grad, dW)


# short reductions (cg_thin.cu, K * Fin <= 16): every width of the register tile (Q = 2, 3, 4, 7 -> 8, 9 -> 12, 13 -> 16, 16),
# scalar and 4-column lanes, ragged column groups, K = 1 (one "sample" of N * M rows), both stack layouts
THIN_SHAPES = [(3, 9, 2, 40, 1), (3, 5, 3, 128, 1), (4, 12, 4, 20, 1), (3, 6, 1, 64, 7), (3, 4, 3, 132, 3), (4, 7, 13, 24, 1),
               (3, 3, 4, 256, 4), (2, 70, 1, 32, 16), (1, 11, 2, 36, 5)]


@pytest.mark.parametrize('flags', [0, 1])        # sample-major stack of the on-chip basis / vertex-major stack of the streaming steps
@pytest.mark.parametrize('level,N,Fin,Fout,K', THIN_SHAPES)
def test_short_reduction_filter_vs_oracle(ops, tf_ref, c2, level, N, Fin, Fout, K, flags):
    L = csr_from(c2, 'L%d' % level)
    M = L.shape[0]
    rng = np.random.RandomState(7 * level + K + Fin)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    for grad_x in (True, False):                 # without dx the weight gradient comes from the X stack (first layers)
        xt = dev(x).requires_grad_(grad_x)
        Wt = dev(W).requires_grad_(True)
        y = ops.cheb_filter(xt, Wt, L, K, lmax=2, flags=flags, grad_x=grad_x)
        close(y, tf_ref.chebyshev5(x, L, W, K))
        y.backward(dev(gy))
        dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
        close(Wt.grad, dW)
        if grad_x:
            close(xt.grad, dx)


# shapes the fused recurrence+contraction (tcgen05) kernel must accept: (level, N, Fin, Fout, K)
FUSED_SHAPES = [
    (2, 5, 32, 64, 25),      # C2 layer 2
    (2, 301, 32, 64, 5),     # several groups per CTA (x prefetch, slab rotation), odd tail
    (4, 37, 16, 48, 4),      # several samples per group, ragged last group
    (3, 10, 64, 32, 3),
    (4, 5, 128, 16, 2),
    (3, 3, 32, 256, 1),      # K = 1: no recurrence
    (1, 3, 16, 64, 4),       # M = 496: four MMA row tiles
    (3, 150, 16, 16, 7),
]


@pytest.mark.parametrize('level,N,Fin,Fout,K', FUSED_SHAPES)
def test_fused_filter_forward_vs_oracle(ops, tf_ref, c2, gather, level, N, Fin, Fout, K):
    L = csr_from(c2, 'L%d' % level)
    M = L.shape[0]
    rng = np.random.RandomState(10 * level + K + N)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    y = ops.cheb_filter(dev(x), dev(W), L, K, lmax=2, flags=ops.FILTER_FORCE_FUSED)
    close(y, tf_ref.chebyshev5(x, L, W, K))
    # and the unfused CUDA path agrees too (same inputs, independent kernels)
    y2 = ops.cheb_filter(dev(x), dev(W), L, K, lmax=2, flags=ops.FILTER_NO_FUSED)
    close(y, y2.cpu().numpy())


@pytest.mark.parametrize('level,N,Fin,Fout,K', [(2, 5, 32, 64, 25), (4, 37, 64, 16, 4), (3, 10, 32, 64, 3),
                                                (3, 9, 16, 16, 1), (3, 5, 64, 64, 25), (4, 130, 16, 32, 3), (2, 301, 32, 64, 5), (4, 700, 16, 16, 4)])
@pytest.mark.parametrize('dx_kernel', ['clenshaw', 'forward_form'])
def test_fused_filter_backward_vs_oracle(ops, tf_ref, c2, gather, level, N, Fin, Fout, K, dx_kernel):
    """dx through the adjoint (Clenshaw) kernel or the forward-form fused kernel on L~^T; dW from the basis."""
    L = csr_from(c2, 'L%d' % level)
    M = L.shape[0]
    rng = np.random.RandomState(20 * level + K + N)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
    flags = ops.FILTER_FORCE_FUSED | (ops.FILTER_NO_CLENSHAW if dx_kernel == 'forward_form' else 0)
    y = ops.cheb_filter(xt, Wt, L, K, lmax=2, flags=flags)
    y.backward(dev(gy))
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
    close(xt.grad, dx)
    close(Wt.grad, dW)


def test_fused_filter_directed(ops, tf_ref, directed, gather):
    """Non-symmetric operator: the fused dx must use the true transpose."""
    L = csr_from(directed, 'L')
    M = L.shape[0]
    rng = np.random.RandomState(11)
    x = rng.standard_normal((6, M, 16)).astype(np.float32)
    W = (0.1 * rng.standard_normal((16 * 5, 32))).astype(np.float32)
    gy = rng.standard_normal((6, M, 32)).astype(np.float32)
    for flags in (ops.FILTER_FORCE_FUSED, ops.FILTER_FORCE_FUSED | ops.FILTER_NO_CLENSHAW):
        xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
        y = ops.cheb_filter(xt, Wt, L, 5, lmax=3.5, flags=flags)
        close(y, tf_ref.chebyshev5(x, L, W, 5, lmax=3.5))
        y.backward(dev(gy))
        dx, dW = tf_ref.chebyshev5_backward(x, L, W, 5, gy, lmax=3.5)
        close(xt.grad, dx)
        close(Wt.grad, dW)


def test_fused_unsupported_shape_raises(ops, c2):
    from cnn_graph_b200 import _native
    L = csr_from(c2, 'L3')
    x = torch.zeros((2, L.shape[0], 3), device='cuda')
    W = torch.zeros((3 * 4, 8), device='cuda')
    with pytest.raises(_native.NativeError):
        ops.cheb_filter(x, W, L, 4, flags=ops.FILTER_FORCE_FUSED)


def test_filter_directed_lmax(ops, tf_ref, directed):
    L = csr_from(directed, 'L')
    M = L.shape[0]
    rng = np.random.RandomState(7)
    x = rng.standard_normal((6, M, 4)).astype(np.float32)
    W = (0.1 * rng.standard_normal((4 * 5, 12))).astype(np.float32)
    gy = rng.standard_normal((6, M, 12)).astype(np.float32)
    before = L.data.copy()
    for flags in FLAG_SETS:
        xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
        y = ops.cheb_filter(xt, Wt, L, 5, lmax=3.5, flags=flags)
        close(y, tf_ref.chebyshev5(x, L, W, 5, lmax=3.5))
        y.backward(dev(gy))
        dx, dW = tf_ref.chebyshev5_backward(x, L, W, 5, gy, lmax=3.5)
        close(xt.grad, dx)
        close(Wt.grad, dW)
    assert np.array_equal(L.data, before)            # caller's Laplacian untouched


def test_chebyshev2_has_no_input_gradient(ops, tf_ref, c2):
    L = csr_from(c2, 'L3')
    rng = np.random.RandomState(3)
    x = rng.standard_normal((4, L.shape[0], 1)).astype(np.float32)
    W = (0.1 * rng.standard_normal((6, 8))).astype(np.float32)
    gy = rng.standard_normal((4, L.shape[0], 8)).astype(np.float32)
    xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
    y = ops.cheb_filter(xt, Wt, L, 6, grad_x=False)
    close(y, tf_ref.chebyshev2(x, L, W, 6))
    y.backward(dev(gy))
    assert xt.grad is None
    close(Wt.grad, tf_ref.chebyshev5_backward(x, L, W, 6, gy)[1])


def test_empty_batch_and_errors(ops, c2):
    from cnn_graph_b200 import _native
    L = csr_from(c2, 'L4')
    W = dev(np.zeros((2 * 3, 4), np.float32))
    y = ops.cheb_filter(dev(np.zeros((0, L.shape[0], 2), np.float32)), W, L, 3)
    assert tuple(y.shape) == (0, L.shape[0], 4)
    with pytest.raises(ValueError):
        ops.cheb_filter(dev(np.zeros((1, L.shape[0] + 1, 2), np.float32)), W, L, 3)
    with pytest.raises(ValueError):
        ops.cheb_filter(dev(np.zeros((1, L.shape[0], 3), np.float32)), W, L, 3)
    with pytest.raises(_native.NativeError):
        ops.cheb_filter(torch.zeros((1, L.shape[0], 2)), W.cpu(), L, 3)      # CPU tensors: no fallback
    with pytest.raises(_native.NativeError):        # unsorted / duplicate column indices are rejected by the C ABI
        import ctypes
        h = ctypes.c_void_p()
        ip = np.array([0, 2], np.int32); ix = np.array([0, 0], np.int32); v = np.ones(2, np.float32)
        _native.check(_native.lib().cg_graph_create(ctypes.byref(h), 1, 2, ip.ctypes.data, ix.ctypes.data, v.ctypes.data), 'create')


# --------------------------------------------------------------------------- full-size properties
def test_full_size_c2_linearity_and_adjoint(ops, c2):
    """BASELINE config C2 layer shapes at batch 100: linear in x, bilinear pairing identities
    <F(x), g> = <x, dx(g)> = <W, dW(x, g)> (no oracle needed at this size)."""
    torch.manual_seed(0)
    for level, Fin, Fout, K in ((0, 1, 32, 25), (2, 32, 64, 25)):
        L = csr_from(c2, 'L%d' % level)
        M = L.shape[0]
        N = 100
        x1 = torch.randn(N, M, Fin, device='cuda')
        x2 = torch.randn(N, M, Fin, device='cuda')
        W = (0.1 * torch.randn(Fin * K, Fout, device='cuda')).requires_grad_(True)
        y1 = ops.cheb_filter(x1, W, L, K)
        y2 = ops.cheb_filter(x2, W, L, K)
        y12 = ops.cheb_filter(2.0 * x1 - 0.5 * x2, W, L, K)
        close(y12, (2.0 * y1 - 0.5 * y2).detach().cpu().numpy())
        xa = x1.clone().requires_grad_(True)
        ya = ops.cheb_filter(xa, W, L, K)
        # g correlated with y: <y, g> ~ |y|^2 / 2 is a well-conditioned sum (an independent g gives a sum of
        # 3e6 cancelling terms whose value is pure rounding noise at any fp32 precision)
        g = 0.5 * ya.detach() + 0.5 * ya.detach().std() * torch.randn(N, M, Fout, device='cuda')
        ya.backward(g)
        lhs = float((ya.detach().double() * g.double()).sum())
        assert abs(lhs - float((xa.detach().double() * xa.grad.double()).sum())) <= 1e-4 * abs(lhs)
        assert abs(lhs - float((W.detach().double() * W.grad.double()).sum())) <= 1e-4 * abs(lhs)
        # K = 1 is the per-vertex linear map y = x W
        W1 = 0.1 * torch.randn(Fin, Fout, device='cuda')
        close(ops.cheb_filter(x1, W1, L, 1), (x1 @ W1).cpu().numpy())


# --------------------------------------------------------------------------- bias/act, pool, perm_data
@pytest.mark.parametrize('act', ['relu', 'tanh', 'none'])
@pytest.mark.parametrize('kind,F', [(0, 5), (1, 32), (1, 3), (2, 8), (2, 7)])
def test_bias_act_vs_torch(ops, act, kind, F):
    torch.manual_seed(1)
    N, M = 6, 20
    x = torch.randn(N, M, F, device='cuda', requires_grad=True)
    b = None if kind == 0 else torch.randn((1, 1, F) if kind == 1 else (1, M, F), device='cuda', requires_grad=True)
    g = torch.randn(N, M, F, device='cuda')
    y = ops.bias_act(x, b, act)
    fn = {'relu': torch.relu, 'tanh': torch.tanh, 'none': lambda t: t}[act]
    xr = x.detach().clone().requires_grad_(True)
    br = None if b is None else b.detach().clone().requires_grad_(True)
    yr = fn(xr if br is None else xr + br)
    close(y, yr.detach().cpu().numpy(), 1e-6)
    y.backward(g)
    yr.backward(g)
    close(x.grad, xr.grad.cpu().numpy(), 1e-5)
    if b is not None:
        close(b.grad, br.grad.cpu().numpy(), 1e-5)


@pytest.mark.parametrize('p', [1, 2, 4, 8])
@pytest.mark.parametrize('F', [1, 32, 5])
def test_pool_bit_exact_vs_oracle(ops, tf_ref, p, F):
    rng = np.random.RandomState(p * 10 + F)
    x = rng.standard_normal((5, 16 * p, F)).astype(np.float32)
    x[x < -0.3] = 0.0                        # ties between zeros as after a ReLU with fake vertices
    g = rng.standard_normal((5, 16, F)).astype(np.float32)
    xt = dev(x).requires_grad_(True)
    y = ops.pool(xt, p, 'max')
    assert np.array_equal(y.detach().cpu().numpy(), tf_ref.mpool1(x, p))            # bit-exact values
    if p > 1:
        _, am = ops.pool_argmax(dev(x), p)
        assert np.array_equal(am.cpu().numpy(), tf_ref.mpool1_argmax(x, p))          # bit-exact first-max indices
        y.backward(dev(g))
        assert np.array_equal(xt.grad.cpu().numpy(), tf_ref.mpool1_backward(x, p, g))
    xa = dev(x).requires_grad_(True)
    ya = ops.pool(xa, p, 'avg')
    close(ya, tf_ref.apool1(x, p), 1e-6)
    if p > 1:
        ya.backward(dev(g))
        close(xa.grad, tf_ref.apool1_backward(x, p, g), 1e-6)


def test_perm_data_bit_exact(ops, c2, c1):
    got = ops.perm_data_device(dev(c2['pd_x']), c2['perm'])
    assert np.array_equal(got.cpu().numpy().astype(np.float64), c2['pd_y'])          # reference fixture
    got = ops.perm_data_device(dev(c1['Xd'][:4]), c1['perm'])
    assert np.array_equal(got.cpu().numpy().astype(np.float64), c1['pd_y'])
    assert ops.perm_data_device(dev(np.zeros((0, 784), np.float32)), c2['perm']).shape == (0, 992)


# --------------------------------------------------------------------------- LSTM
@pytest.mark.parametrize('variant', ['fork', 'standard'])
def test_lstm_cell_step_vs_oracle(ops, tf_ref, c2, variant):
    from cnn_graph_b200.lib import gconv_lstm, variables
    L = csr_from(c2, 'L4')
    M, N, Fin, H, K = L.shape[0], 3, 2, 8, 3
    rng = np.random.RandomState(5)
    x = (0.5 * rng.standard_normal((N, M, Fin))).astype(np.float32)
    h = (0.3 * rng.standard_normal((N, M, H))).astype(np.float32)
    c = (0.3 * rng.standard_normal((N, M, H))).astype(np.float32)
    store = variables.VariableStore(device='cuda', seed=1)
    cell = gconv_lstm.GConvLSTMCell(num_units=H, laplacian=L, lmax=2, K=K, feat_in=Fin, nNode=M, gate_variant=variant)
    with variables.use_store(store):
        new_h, state = cell(dev(x), (dev(c), dev(h)))
    v = {k.split('/')[-1]: p.detach().cpu().numpy() for k, p in store.vars.items()}
    Wx = {g: v['W%sxt' % g] for g in 'zifo'}
    Wh = {g: v['W%sht' % g] for g in 'zifo'}
    b = {g: v['b%st' % g] for g in 'zifo'}
    ref_h, ref_c = tf_ref.gconv_lstm_step(x, c, h, L, 2, K, Wx, Wh, b, variant)
    close(state.c, ref_c)
    close(new_h, ref_h)
    assert state.h is new_h and cell.output_size == H and cell.state_size.c == (M, H)


def test_c4_humanflow_shaped_cell_forward_backward(ops, tf_ref):
    """Config C4: 32x32 8-NN grid (M = 1024, gconvTest.py:79), Fin = 2, H = 128, K = 3 -- the gate filters have
    Fout = 4H = 512 and Fin in {2, 128}: the wide-output tensor-core contraction and its two backward GEMMs."""
    from cnn_graph_b200.lib import graph
    from oracle import graph_ref
    A = graph_ref.adjacency(*graph_ref.distance_sklearn_metrics(graph_ref.grid(32), k=8, metric='euclidean'))
    L = graph_ref.laplacian(A, normalized=True)
    M, N, Fin, H, K = 1024, 3, 2, 128, 3
    assert L.shape == (M, M)
    rng = np.random.RandomState(11)
    x = rng.uniform(0, 1, (N, M, Fin)).astype(np.float32)
    h = (0.3 * rng.standard_normal((N, M, H))).astype(np.float32)
    c = (0.3 * rng.standard_normal((N, M, H))).astype(np.float32)
    Wx = rng.uniform(-0.1, 0.1, (Fin * K, 4 * H)).astype(np.float32)
    Wh = rng.uniform(-0.1, 0.1, (H * K, 4 * H)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, 4 * H).astype(np.float32)
    gh = rng.standard_normal((N, M, H)).astype(np.float32)
    xt, ht, Wxt, Wht = (dev(a).requires_grad_(True) for a in (x, h, Wx, Wh))
    px = ops.cheb_filter(xt, Wxt, L, K)
    ph = ops.cheb_filter(ht, Wht, L, K)
    close(px, tf_ref.chebyshev5(x, L, Wx, K))
    close(ph, tf_ref.chebyshev5(h, L, Wh, K))
    new_h, new_c = ops.lstm_gates(px + ph, dev(b), dev(c), 'standard')
    sl = {g: slice(i * H, (i + 1) * H) for i, g in enumerate('zifo')}
    ref_h, ref_c = tf_ref.gconv_lstm_step(x, c, h, L, 2, K, {g: Wx[:, s] for g, s in sl.items()},
                                          {g: Wh[:, s] for g, s in sl.items()}, {g: b[s] for g, s in sl.items()},
                                          'standard')
    close(new_c, ref_c)
    close(new_h, ref_h)
    # backward of the two filters against the oracle's adjoint, seeded with the gate gradient the GPU produced
    new_h.backward(dev(gh))
    pre = (px + ph).detach().requires_grad_(True)
    ops.lstm_gates(pre, dev(b), dev(c), 'standard')[0].backward(dev(gh))
    g = pre.grad.cpu().numpy()
    dx, dWx = tf_ref.chebyshev5_backward(x, L, Wx, K, g)
    dh, dWh = tf_ref.chebyshev5_backward(h, L, Wh, K, g)
    close(xt.grad, dx)
    close(ht.grad, dh)
    close(Wxt.grad, dWx)
    close(Wht.grad, dWh)


@pytest.mark.parametrize('variant', ['fork', 'standard'])
def test_lstm_gates_gradients_vs_torch(ops, variant):
    torch.manual_seed(2)
    R, H = 37, 5
    pre = (0.5 * torch.randn(1, R, 4 * H, device='cuda')).requires_grad_(True)
    bias = (0.1 * torch.randn(4 * H, device='cuda')).requires_grad_(True)
    c = torch.randn(1, R, H, device='cuda', requires_grad=True)
    gh, gc = torch.randn(1, R, H, device='cuda'), torch.randn(1, R, H, device='cuda')
    nh, nc = ops.lstm_gates(pre, bias, c, variant)
    (nh * gh).sum().add((nc * gc).sum()).backward()
    p2, b2, c2_ = (t.detach().double().clone().requires_grad_(True) for t in (pre, bias, c))
    a = p2 + b2
    z, i, f, o = a[..., :H], a[..., H:2 * H], a[..., 2 * H:3 * H], a[..., 3 * H:]
    zz = torch.tan(z) if variant == 'fork' else torch.tanh(z)
    oo = torch.tanh(o) if variant == 'fork' else torch.sigmoid(o)
    rc = torch.sigmoid(f) * c2_ + torch.sigmoid(i) * zz
    rh = oo * torch.tanh(rc)
    (rh * gh.double()).sum().add((rc * gc.double()).sum()).backward()
    close(nc, rc.detach().cpu().numpy(), 1e-5)
    close(nh, rh.detach().cpu().numpy(), 1e-5)
    close(pre.grad, p2.grad.cpu().numpy())
    close(c.grad, c2_.grad.cpu().numpy())
    close(bias.grad, b2.grad.cpu().numpy())


# --------------------------------------------------------------------------- models
def test_cgcnn_forward_backward_vs_oracle_c2(ops, tf_ref, c2):
    """The MNIST-shaped model of BASELINE config C2 (GC32-P4-GC64-P4-FC512-FC10) end to end."""
    from cnn_graph_b200.lib import models
    L = [csr_from(c2, 'L%d' % i) for i in range(5)]
    N = 6
    model = models.cgcnn(L, F=[32, 64], K=[25, 25], p=[4, 4], M=[512, 10], batch_size=N, dropout=1)
    model.fuse_brelu_pool = False          # keep nets['conv*/bias_relu'] (the fused tail is checked below)
    rng = np.random.RandomState(9)
    x = rng.uniform(0, 1, (N, 992)).astype(np.float32)
    labels = rng.randint(0, 10, N)
    logits = model.inference(dev(x), 1)
    v = {k: p.detach().cpu().numpy() for k, p in model.store.vars.items()}
    a1 = tf_ref.chebyshev5(x[:, :, None], L[0], v['conv1/weights'], 25)
    r1 = tf_ref.b1relu(a1, v['conv1/bias'])
    p1 = tf_ref.mpool1(r1, 4)
    a2 = tf_ref.chebyshev5(p1, L[2], v['conv2/weights'], 25)
    r2 = tf_ref.b1relu(a2, v['conv2/bias'])
    p2 = tf_ref.mpool1(r2, 4)
    f1 = tf_ref.fc(p2.reshape(N, -1), v['fc1/weights'], v['fc1/bias'])
    ref_logits = tf_ref.fc(f1, v['logits/weights'], v['logits/bias'], relu=False)
    close(logits, ref_logits)
    # backward: seed with d(sum of logits * g) and push it through the oracle chain by hand.
    # ReLU masks and max-pool routing are discontinuous in the activations, so the oracle
    # chain is evaluated AT the product's (already parity-checked) activations: a value that
    # is +1e-8 on one side and 0 on the other would otherwise flip a whole gradient entry.
    nets = {k: t.detach().cpu().numpy() for k, t in model.nets.items()}
    close(nets['conv1/bias_relu'], r1)
    close(nets['conv2/bias_relu'], r2)
    close(nets['conv1/pooling'], p1)
    close(nets['fc1'], f1)
    r1, r2, p1, f1 = nets['conv1/bias_relu'], nets['conv2/bias_relu'], nets['conv1/pooling'], nets['fc1']
    g = rng.standard_normal(ref_logits.shape).astype(np.float32)
    (logits * dev(g)).sum().backward()
    gf1 = (g @ v['logits/weights'].T) * (f1 > 0)
    gp2 = (gf1 @ v['fc1/weights'].T).reshape(p2.shape)
    gr2 = tf_ref.mpool1_backward(r2, 4, gp2)
    ga2 = gr2 * (r2 > 0)
    gp1, dW2 = tf_ref.chebyshev5_backward(p1, L[2], v['conv2/weights'], 25, ga2)
    gr1 = tf_ref.mpool1_backward(r1, 4, gp1)
    ga1 = gr1 * (r1 > 0)
    _, dW1 = tf_ref.chebyshev5_backward(x[:, :, None], L[0], v['conv1/weights'], 25, ga1)
    close(model.store.vars['conv2/weights'].grad, dW2)
    close(model.store.vars['conv1/weights'].grad, dW1)
    close(model.store.vars['conv2/bias'].grad.reshape(-1), ga2.sum(axis=(0, 1)))
    close(model.store.vars['conv1/bias'].grad.reshape(-1), ga1.sum(axis=(0, 1)))
    # one optimisation step runs and changes the weights
    before = v['conv1/weights'].copy()
    loss = model.train_step(dev(x), torch.from_numpy(labels).cuda())
    assert np.isfinite(float(loss))
    assert not np.array_equal(before, model.get_var('conv1/weights'))


def test_cgcnn_fused_brelu_pool_matches_unfused_model(ops, c2):
    """The default model fuses brelu -> pool into one kernel: same logits and parameter gradients as the
    model that runs them separately (which the test above checks against the oracle)."""
    from cnn_graph_b200.lib import models
    L = [csr_from(c2, 'L%d' % i) for i in range(5)]
    N = 8
    rng = np.random.RandomState(4)
    x = dev(rng.uniform(0, 1, (N, 992)).astype(np.float32))
    g = dev(rng.standard_normal((N, 10)).astype(np.float32))
    results = []
    for fuse in (True, False):
        torch.manual_seed(77)
        model = models.cgcnn(L, F=[32, 64], K=[25, 25], p=[4, 4], M=[512, 10], batch_size=N, dropout=1)
        model.fuse_brelu_pool = fuse
        logits = model.inference(x, 1)
        assert ('conv1/bias_relu' in model.nets) == (not fuse)
        (logits * g).sum().backward()
        results.append((logits.detach().cpu().numpy(),
                        {k: p.grad.detach().cpu().numpy() for k, p in model.store.vars.items()}))
    (lf, gf), (lu, gu) = results
    close(lf, lu, 1e-6)
    assert set(gf) == set(gu)
    for k in gu:
        close(gf[k], gu[k], 1e-5)


def test_graphconv_and_glstm_models_run(ops, tf_ref, c2):
    from cnn_graph_b200.lib import gconv_lstm, graph_conv
    L = csr_from(c2, 'L3')
    M = L.shape[0]
    rng = np.random.RandomState(4)
    # fork residual net: conv_init -> 1 residual layer -> convN, plain-ReLU b1relu
    net = graph_conv.GraphConv([L], F=[8], K=[3], p=[1], M=[2], _nfilter=8, _nres_layer_count=1, C_0=[6], batch_size=3)
    x = rng.uniform(0, 1, (3, M, 6)).astype(np.float32)
    out = net.inference(dev(x), 0)
    v = {k: p.detach().cpu().numpy() for k, p in net.store.vars.items()}
    a = np.maximum(tf_ref.chebyshev5(x, L, v['conv_init/weights'], 3), 0)
    b = np.maximum(tf_ref.chebyshev5(a, L, v['residual_layer_0/sublayer0/weights'], 3), 0)
    c = np.maximum(tf_ref.chebyshev5(b, L, v['residual_layer_0/sublayer1/weights'], 3) + a, 0)
    close(out, tf_ref.chebyshev5(c, L, v['convN/weights'], 3))
    y = rng.uniform(0, 1, (3, M, 2)).astype(np.float32)
    assert np.isfinite(float(net.train_step(dev(x), dev(y))))
    # gLSTM model: 3 frames of 2 features, injected all-ones dropout masks -> deterministic
    gm = gconv_lstm.GconvModel(L, 3, 0, 0, filter_num=8, kernel_num=3, feature_num=6, in_feature_num=2,
                               infer_func='inference_glstm', batch_size=3, gate_variant='standard', output_keep_prob=1.0)
    out = gm.inference(dev(x), 0)
    v = {k.split('/')[-1] if 'Cell' in k else k: p.detach().cpu().numpy() for k, p in gm.store.vars.items()}
    Wx = {g: v['W%sxt' % g] for g in 'zifo'}
    Wh = {g: v['W%sht' % g] for g in 'zifo'}
    bb = {g: v['b%st' % g] for g in 'zifo'}
    h = np.zeros((3, M, 8), np.float32)
    cc = np.zeros((3, M, 8), np.float32)
    frames = x.reshape(3, M, 2, 3)
    for t in range(3):
        h, cc = tf_ref.gconv_lstm_step(frames[..., t], cc, h, L, 2, 3, Wx, Wh, bb, 'standard')
    close(out, tf_ref.chebyshev5(h, L, v['conv_init/weights'], 3))
    assert np.isfinite(float(gm.train_step(dev(x), dev(y))))


def test_saved_stack_backward_matches_recompute(ops, tf_ref, c2, gather):
    """The forward pass leaves the basis behind (sample-major) and the backward pass uses it for dW:
    same gradients as the recomputing path and as the oracle."""
    L = csr_from(c2, 'L2')
    M = L.shape[0]
    rng = np.random.RandomState(5)
    N, Fin, Fout, K = 37, 32, 64, 9
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    grads = []
    for save in (True, False):
        ops.set_save_stack(save)
        try:
            xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
            y = ops.cheb_filter(xt, Wt, L, K)
            assert (y.grad_fn.saved_tensors[2] is not None) == save
            y.backward(dev(gy))
            grads.append((xt.grad.cpu().numpy(), Wt.grad.cpu().numpy()))
        finally:
            ops.set_save_stack(True)
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
    for gx, gw in grads:
        close(gx, dx)
        close(gw, dW)
    # default format of the saved basis for this shape (the loop above used it): the fused kernel's bf16 operand
    # planes [2][K][128-row chunks][Fin/8][128][8]; hi + mid reproduce the reference's graph.chebyshev of every sample
    # to 2^-16
    from oracle import graph_ref
    Lr = ops.rescale_csr(L)
    ref = graph_ref.chebyshev(Lr, np.ascontiguousarray(x.transpose(1, 0, 2).reshape(M, N * Fin)), K)
    y = ops.cheb_filter(dev(x).requires_grad_(True), dev(W), L, K)
    assert y.grad_fn.stack_planes
    nch = (N * M + 127) // 128
    pl = y.grad_fn.saved_tensors[2].reshape(2, K, nch, Fin // 8, 128, 8).float().sum(0)          # [K, chunk, octet, row, 8]
    basis = pl.permute(0, 1, 3, 2, 4).reshape(K, nch * 128, Fin)[:, :N * M].reshape(K, N, M, Fin).cpu().numpy()
    close(basis.transpose(0, 2, 1, 3).reshape(K, M, N * Fin), ref, 3e-5)
    # the fp32 format [K, N, M, Fin] on request: same basis, same gradients
    ops.set_stack_planes(False)
    try:
        xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
        y = ops.cheb_filter(xt, Wt, L, K)
        assert not y.grad_fn.stack_planes
        stack = y.grad_fn.saved_tensors[2].cpu().numpy()
        y.backward(dev(gy))
    finally:
        ops.set_stack_planes(True)
    close(stack.transpose(0, 2, 1, 3).reshape(K, M, N * Fin), ref, 1e-5)
    close(xt.grad, dx)
    close(Wt.grad, dW)


@pytest.mark.parametrize('kind', ['max', 'avg'])
@pytest.mark.parametrize('act,bias_kind', [('relu', 0), ('relu', 1), ('relu', 2), ('tanh', 1), ('none', 1)])
@pytest.mark.parametrize('N,M,F,p', [(7, 24, 32, 4), (3, 16, 5, 2), (2, 64, 64, 8), (5, 8, 12, 4)])
def test_fused_bias_act_pool_matches_separate_ops(ops, kind, act, bias_kind, N, M, F, p):
    """pool(bias_act(x)) in one kernel: pooled values bit-equal to the two-kernel path, same gradients."""
    if not ops.bias_act_pool_supported(act, p, kind):
        pytest.skip('combination is not fused')
    torch.manual_seed(N * 100 + M + F + p)
    x = torch.randn(N, M, F, device='cuda')
    x[0, :p] = 0.25                                   # ties inside one pooling window: first index must win
    bias = None if bias_kind == 0 else (0.3 * torch.randn(F if bias_kind == 1 else M * F, device='cuda'))
    gy = torch.randn(N, M // p, F, device='cuda')
    outs = []
    for fused in (True, False):
        xt = x.clone().requires_grad_(True)
        bt = None if bias is None else bias.clone().requires_grad_(True)
        if fused:
            y = ops.bias_act_pool(xt, bt, act, p, kind)
        else:
            y = ops.pool(ops.bias_act(xt, bt, act), p, kind)
        y.backward(gy)
        outs.append((y.detach(), xt.grad, None if bt is None else bt.grad))
    (y1, gx1, gb1), (y0, gx0, gb0) = outs
    assert torch.equal(y1, y0)
    assert torch.equal(gx1, gx0) if kind == 'max' and act != 'tanh' else torch.allclose(gx1, gx0, rtol=1e-6, atol=1e-7)
    if gb0 is not None:
        assert torch.allclose(gb1, gb0, rtol=1e-4, atol=1e-5 * float(gb0.abs().max()))


@pytest.mark.parametrize('level,N,Fin,Fout,K', [(0, 4, 1, 32, 25), (3, 12, 1, 16, 5), (4, 64, 2, 32, 7), (2, 9, 4, 64, 3),
                                                (4, 66, 2, 128, 2), (3, 8, 8, 48, 20), (4, 4, 1, 16, 1)])
def test_tensor_core_contraction_of_hbm_basis(ops, tf_ref, c2, level, N, Fin, Fout, K):
    """Narrow inputs (Fin < 16) cannot use the fused kernel: sample-major basis in HBM, then the tcgen05
    contraction kernel (ragged last chunk, Q = K*Fin not a multiple of 16, several chunks per CTA)."""
    L = csr_from(c2, 'L%d' % level)
    M = L.shape[0]
    rng = np.random.RandomState(31 * level + K + N)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
    y = ops.cheb_filter(xt, Wt, L, K)
    close(y, tf_ref.chebyshev5(x, L, W, K))
    close(y, ops.cheb_filter(dev(x), dev(W), L, K, flags=ops.FILTER_NO_FUSED).cpu().numpy())
    y.backward(dev(gy))
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
    close(xt.grad, dx)
    close(Wt.grad, dW)


# --------------------------------------------------------------------------- BASELINE config C1 (usage.ipynb)
def test_c1_usage_config_filters_and_model(ops, tf_ref, c1):
    """Config C1: 100-feature kNN graph (irregular rows, up to 31 entries), 3 coarsening levels
    [104, 52, 26], cgcnn F=[32,64] K=[20,20] p=[4,2] M=[512,3] with average pooling, batch 100.
    Layer 2 (M = 26) packs several samples into one MMA row tile of the fused kernel."""
    from cnn_graph_b200.lib import models
    L = [csr_from(c1, 'L%d' % i) for i in range(3)]
    N = 100
    rng = np.random.RandomState(42)
    # the two graph-conv layers on their own, forward and backward
    for Lk, Fin, Fout in ((L[0], 1, 32), (L[2], 32, 64)):
        M = Lk.shape[0]
        x = rng.standard_normal((N, M, Fin)).astype(np.float32)
        W = (0.1 * rng.standard_normal((Fin * 20, Fout))).astype(np.float32)
        gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
        xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
        y = ops.cheb_filter(xt, Wt, Lk, 20)
        close(y, tf_ref.chebyshev5(x, Lk, W, 20))
        y.backward(dev(gy))
        dx, dW = tf_ref.chebyshev5_backward(x, Lk, W, 20, gy)
        close(xt.grad, dx)
        close(Wt.grad, dW)
    # the model: filter -> b1relu -> apool1 twice, then the dense head
    # the constructor wants all 1 + log2(4) + log2(2) levels; the last one (13 vertices) is never filtered on
    L4 = L + [scipy.sparse.identity(13, format='csr', dtype=np.float32)]
    model = models.cgcnn(L4, F=[32, 64], K=[20, 20], p=[4, 2], M=[512, 3], pool='apool1', batch_size=N, dropout=1)
    x = rng.standard_normal((N, 104)).astype(np.float32)
    logits = model.inference(dev(x), 1)
    v = {k: p_.detach().cpu().numpy() for k, p_ in model.store.vars.items()}
    a1 = tf_ref.chebyshev5(x[:, :, None], L[0], v['conv1/weights'], 20)
    p1 = tf_ref.apool1(tf_ref.b1relu(a1, v['conv1/bias']), 4)
    a2 = tf_ref.chebyshev5(p1, L[2], v['conv2/weights'], 20)
    p2 = tf_ref.apool1(tf_ref.b1relu(a2, v['conv2/bias']), 2)
    f1 = tf_ref.fc(p2.reshape(N, -1), v['fc1/weights'], v['fc1/bias'])
    close(logits, tf_ref.fc(f1, v['logits/weights'], v['logits/bias'], relu=False))
    close(model.nets['conv1/pooling'], p1)
    close(model.nets['conv2/pooling'], p2)


# --------------------------------------------------------------------------- BASELINE config C3 (20NEWS-shaped)
def test_c3_20news_shaped_config(ops, tf_ref):
    """Config C3: 10 000-word 16-NN cosine feature graph (no coarsening), cgcnn F=[32] K=[5] p=[1] M=[20],
    l1-normalised sparse bag-of-words rows densified per batch (lib/graph_model.py:150-151), batch 100.
    The operator does not fit shared memory: streaming recurrence (one CSR step per launch)."""
    import scipy.sparse
    from cnn_graph_b200.lib import graph, models
    rng = np.random.RandomState(2017)
    Mw, N = 10000, 100
    emb = rng.standard_normal((Mw, 100)).astype(np.float32)
    dist_, idx = graph.distance_sklearn_metrics(emb, k=16, metric='cosine')
    L = graph.laplacian(graph.adjacency(dist_, idx), normalized=True)
    assert not ops.get_handle(L).info()['onchip']
    counts = scipy.sparse.random(N, Mw, density=0.008, random_state=rng, format='csr', dtype=np.float32)
    counts.data = np.ceil(counts.data * 5)
    x = np.asarray((counts.multiply(1.0 / counts.sum(axis=1))).todense(), dtype=np.float32)      # rows sum to 1
    W = (0.1 * rng.standard_normal((5, 32))).astype(np.float32)
    gy = rng.standard_normal((N, Mw, 32)).astype(np.float32)
    xt, Wt = dev(x[:, :, None]).requires_grad_(True), dev(W).requires_grad_(True)
    y = ops.cheb_filter(xt, Wt, L, 5)
    close(y, tf_ref.chebyshev5(x[:, :, None], L, W, 5))
    y.backward(dev(gy))
    dx, dW = tf_ref.chebyshev5_backward(x[:, :, None], L, W, 5, gy)
    close(xt.grad, dx)
    close(Wt.grad, dW)
    model = models.cgcnn([L], F=[32], K=[5], p=[1], M=[20], batch_size=N, dropout=1)
    logits = model.inference(dev(x), 1)
    v = {k: p_.detach().cpu().numpy() for k, p_ in model.store.vars.items()}
    r1 = tf_ref.b1relu(tf_ref.chebyshev5(x[:, :, None], L, v['conv1/weights'], 5), v['conv1/bias'])
    close(logits, tf_ref.fc(r1.reshape(N, -1), v['logits/weights'], v['logits/bias'], relu=False))


# --------------------------------------------------------------------------- dense head GEMM
@pytest.mark.parametrize('M,N,K', [(1024, 512, 3968), (100, 10, 512), (7, 300, 65), (129, 257, 64), (256, 256, 1000),
                                   # 16-byte aligned leading dimensions: the pipelined kernel (ragged M/N/K tails, every BN,
                                   # several work items per CTA, split-K)
                                   (132, 260, 68), (20000, 512, 384), (4, 12, 4), (384, 512, 5120), (1000, 36, 200),
                                   (300, 100, 36), (2560, 128, 1536)])
@pytest.mark.parametrize('transA,transB', [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_f32_tensor_core(ops, M, N, K, transA, transB):
    """cg_gemm_f32 against float64 matmul: fp32-level accuracy (bf16 hi+mid split, fp32 accumulation), all four
    operand orientations, ragged tiles, split-K."""
    torch.manual_seed(M + N + K)
    A = torch.randn((K, M) if transA else (M, K), device='cuda')
    B = torch.randn((N, K) if transB else (K, N), device='cuda')
    bias = torch.randn(N, device='cuda')
    ref = (A.double().t() if transA else A.double()) @ (B.double().t() if transB else B.double())
    close(ops.gemm(A, B, transA=transA, transB=transB), ref.cpu().numpy())
    got = ops.gemm(A, B, transA=transA, transB=transB, bias=bias, relu=True)
    close(got, torch.relu(ref + bias.double()).cpu().numpy())


@pytest.mark.parametrize('M,N,K', [(20000, 512, 384), (1024, 512, 3968), (2560, 128, 1536), (5000, 384, 512), (777, 40, 96),
                                   (1000, 36, 200), (640, 20, 100)])       # K tails: zero-filled by the tensor copy
@pytest.mark.parametrize('transA,transB', [(False, False), (False, True), (True, False), (True, True)])
def test_streaming_gemm_bit_equal_to_pipelined(ops, M, N, K, transA, transB):
    """cg_gemm_stream.cu (tensor-map A tiles, B packed once) against cg_gemm_pipe.cu on the same operands: the same
    bf16 hi/mid split and the same MMA order, so every bit of the result agrees (ragged M / N tiles, split-K, column
    tiles of 192)."""
    from cnn_graph_b200 import _native
    lib = _native.lib()
    torch.manual_seed(M + N + K)
    A = torch.randn((K, M) if transA else (M, K), device='cuda')
    B = torch.randn((N, K) if transB else (K, N), device='cuda')
    bias = torch.randn(N, device='cuda')
    before = lib.cg_debug_gemm_stream(1)
    try:
        got = ops.gemm(A, B, transA=transA, transB=transB, bias=bias, relu=True)
        lib.cg_debug_gemm_stream(0)
        want = ops.gemm(A, B, transA=transA, transB=transB, bias=bias, relu=True)
    finally:
        lib.cg_debug_gemm_stream(before)
    assert torch.equal(got, want)
    ref = (A.double().t() if transA else A.double()) @ (B.double().t() if transB else B.double())
    close(got, torch.relu(ref + bias.double()).cpu().numpy())


def test_native_adam_matches_torch(ops):
    """ops.NativeAdam (cg_adam: every variable in one launch, step count on the device) against torch.optim.Adam over
    several steps and tensor sizes (one not a multiple of 4, one larger than a block's stride)."""
    torch.manual_seed(5)
    shapes = [(7,), (33, 5), (128, 64), (1, 1), (700, 1001)]
    ps = [torch.randn(s, device='cuda').requires_grad_(True) for s in shapes]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    mine, ref = ops.NativeAdam(ps, lr=3e-3), torch.optim.Adam(qs, lr=3e-3)
    for it in range(7):
        gs = [torch.randn(s, device='cuda') * (0.1 + it) for s in shapes]
        for p, q, g in zip(ps, qs, gs):
            p.grad, q.grad = g.clone(), g.clone()
        mine.step()
        ref.step()
    for p, q in zip(ps, qs):
        np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().cpu().numpy(), rtol=2e-5, atol=2e-6)
    assert int(mine.state[ps[0]]['step_state'][0]) == 7 and int(mine.state[ps[0]]['step_state'][1]) == 0


def test_linear_layer_gradients(ops):
    torch.manual_seed(3)
    x = torch.randn(64, 200, device='cuda', requires_grad=True)
    W = (0.1 * torch.randn(200, 48, device='cuda')).requires_grad_(True)
    b = torch.randn(48, device='cuda', requires_grad=True)
    g = torch.randn(64, 48, device='cuda')
    y = ops.linear(x, W, b, relu=True)
    y.backward(g)
    xd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (x, W, b))
    yd = torch.relu(xd @ Wd + bd)
    yd.backward(g.double())
    close(y, yd.detach().cpu().numpy())
    close(x.grad, xd.grad.cpu().numpy())
    close(W.grad, Wd.grad.cpu().numpy())
    close(b.grad, bd.grad.cpu().numpy())


# --------------------------------------------------------------------------- captured step / pipelined feeder
def _small_cgcnn(c2, seed):
    from cnn_graph_b200.lib import models
    L = [csr_from(c2, 'L%d' % i) for i in range(5)]
    torch.manual_seed(seed)
    return models.cgcnn(L, F=[8, 16], K=[5, 4], p=[4, 4], M=[32, 10], batch_size=16, dropout=1, learning_rate=0.05,
                        decay_rate=0.9, decay_steps=3, momentum=0.9, regularization=1e-3)


def test_graphed_step_matches_eager_steps(ops, c2):
    """train_step_graphed (CUDA-graph replay; re-captured when the staircase learning rate changes) walks the same
    trajectory as eager train_step on the same batches."""
    rng = np.random.RandomState(3)
    M = csr_from(c2, 'L0').shape[0]
    batches = [(dev(rng.rand(16, M).astype(np.float32)), torch.from_numpy(rng.randint(0, 10, 16)).cuda())
               for _ in range(6)]
    a, b = _small_cgcnn(c2, 7), _small_cgcnn(c2, 7)
    for pa, pb in zip(a.store.parameters(), b.store.parameters()):
        pb.data.copy_(pa.data)
    for x, y in batches:                    # decay_steps = 3: the fourth step re-captures with the decayed rate
        la = a.train_step(x, y)
        lb = b.train_step_graphed(x, y).clone()
        close(lb, la.detach().cpu().numpy(), 2e-4)
    assert a.global_step == b.global_step == len(batches)
    assert b.graphed_native_launches() > 0
    for pa, pb in zip(a.store.parameters(), b.store.parameters()):
        close(pb, pa.detach().cpu().numpy(), 1e-3)


@pytest.mark.parametrize('use_graph', [False, True])
def test_pipelined_trainer_matches_direct_steps(ops, c2, use_graph):
    """PipelinedTrainer (H2D on a copy stream, lagged loss read-back) returns the losses of the same steps driven
    one by one from device tensors, in order."""
    from cnn_graph_b200.lib import coarsening
    rng = np.random.RandomState(5)
    perm = [int(v) for v in c2['perm']]
    raw = [torch.from_numpy(rng.rand(16, 784).astype(np.float32)).pin_memory() for _ in range(5)]
    lab = [torch.from_numpy(rng.randint(0, 10, 16)).pin_memory() for _ in range(5)]
    a, b = _small_cgcnn(c2, 11), _small_cgcnn(c2, 11)
    for pa, pb in zip(a.store.parameters(), b.store.parameters()):
        pb.data.copy_(pa.data)
    step = a.train_step_graphed if use_graph else a.train_step
    want = []
    for x, y in zip(raw, lab):
        xd = dev(coarsening.perm_data(x.numpy(), perm).astype(np.float32))
        want.append(float(step(xd, y.cuda())))
    tr = b.pipelined_trainer(perm=perm, depth=2, use_graph=use_graph)
    for x, y in zip(raw, lab):
        tr.submit(x, y)
    got = tr.drain()
    assert len(got) == len(want)
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-6)
    assert tr.h2d_bytes_per_step == 16 * 784 * 4 + 16 * 8


# --------------------------------------------------------------------------- row-partitioned filter (config C5)
@pytest.mark.parametrize('Fin,Fout,K', [(64, 64, 5), (8, 16, 4), (3, 5, 3)])
def test_partitioned_filter_single_rank_vs_oracle(ops, tf_ref, directed, Fin, Fout, K):
    """PartitionedFilter at world 1 (no halo): cg_cheb_step recurrences on L~ and L~^T, cg_cheb_contract and
    cg_cheb_contract_dw over the strided stack -- against the oracle on a directed operator (N = 1)."""
    from cnn_graph_b200 import partition
    L = csr_from(directed, 'L')
    M = L.shape[0]
    Lr = ops.rescale_csr(L, 3.5)
    rng = np.random.RandomState(17)
    x = rng.standard_normal((M, Fin)).astype(np.float32)
    W = (0.2 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((M, Fout)).astype(np.float32)
    pf = partition.PartitionedFilter(Lr, K, rank=0, world=1)
    y = pf.forward(dev(x), dev(W))
    dx, dW = pf.backward(dev(gy))
    close(y, tf_ref.chebyshev5(x[None], L, W, K, lmax=3.5)[0])
    rdx, rdW = tf_ref.chebyshev5_backward(x[None], L, W, K, gy[None], lmax=3.5)
    close(dx, rdx[0])
    close(dW, rdW)


# --------------------------------------------------------------------------- fused first layer
@pytest.mark.parametrize('level,N,Fout,K,with_bias', [(0, 8, 32, 25, True), (1, 16, 64, 5, True), (2, 8, 32, 7, False),
                                                     (0, 64, 32, 20, True)])
def test_first_layer_fused_node_vs_separate_ops_and_oracle(ops, tf_ref, c2, level, N, Fout, K, with_bias):
    """ops.first_layer (filter + bias/relu/max-pool-4 as one node; backward from the POOLED gradient straight to dW
    and db) against the three separate ops (pooled values bit-equal) and against the oracle's gradients."""
    L = csr_from(c2, 'L%d' % level)
    M = L.shape[0]
    rng = np.random.RandomState(31 + level + K)
    x = rng.standard_normal((N, M, 1)).astype(np.float32)
    W = (0.3 * rng.standard_normal((K, Fout))).astype(np.float32)
    b = (0.2 * rng.standard_normal(Fout)).astype(np.float32) if with_bias else None
    g = rng.standard_normal((N, M // 4, Fout)).astype(np.float32)
    Wt = dev(W).requires_grad_(True)
    bt = dev(b).requires_grad_(True) if with_bias else None
    assert ops.first_layer_supported(dev(x), Wt, bt, L, K, 'relu', 4, 'max')
    yp = ops.first_layer(dev(x), Wt, bt, L, K)
    ypn, aux = (t.detach().cpu().numpy() for t in yp.grad_fn.saved_tensors[:2])       # pooled output, argmax bytes
    yp.backward(dev(g))
    W2 = dev(W).requires_grad_(True)
    b2 = dev(b).requires_grad_(True) if with_bias else None
    ref = ops.bias_act_pool(ops.cheb_filter(dev(x), W2, L, K, grad_x=False), b2, 'relu', 4, 'max')
    ref.backward(dev(g))
    assert torch.equal(yp, ref)
    close(Wt.grad, W2.grad.cpu().numpy())
    # pooling inside the contraction's epilogue (default) and as a separate kernel: same bits
    ops.set_first_layer_epilogue(False)
    try:
        yp_sep = ops.first_layer(dev(x), dev(W), None if b is None else dev(b), L, K)
    finally:
        ops.set_first_layer_epilogue(True)
    assert torch.equal(yp, yp_sep)
    # oracle.  The gradient is routed by the argmax of every pool group; among ~10^5..10^6 groups a near-tie or two
    # resolves differently in the oracle's forward pass, so the routing uses the decisions the GPU made (its argmax
    # bytes, checked to point at a maximal element of the oracle's own activations)
    a = tf_ref.chebyshev5(x, L, W, K)
    r = tf_ref.b1relu(a, b) if with_bias else np.maximum(a, 0)
    close(yp, tf_ref.mpool1(r, 4))
    rg = r.reshape(N, M // 4, 4, Fout)
    picked = np.take_along_axis(rg, aux[:, :, None, :].astype(np.int64), axis=2)[:, :, 0, :]
    assert np.abs(picked - rg.max(axis=2)).max() <= 1e-4 * np.abs(r).max()
    ga = np.zeros((N, M // 4, 4, Fout), np.float32)
    np.put_along_axis(ga, aux[:, :, None, :].astype(np.int64), (g * (ypn > 0))[:, :, None, :], axis=2)
    ga = ga.reshape(N, M, Fout)
    _, dW = tf_ref.chebyshev5_backward(x, L, W, K, ga)
    close(Wt.grad, dW)
    if with_bias:
        close(bt.grad, ga.sum(axis=(0, 1)))
        close(bt.grad, b2.grad.cpu().numpy())
    # an input that needs a gradient, other pool sizes: not this node
    assert not ops.first_layer_supported(dev(x).requires_grad_(True), Wt, bt, L, K, 'relu', 4, 'max')
    assert not ops.first_layer_supported(dev(x), Wt, bt, L, K, 'relu', 2, 'max')
    assert not ops.first_layer_supported(dev(x), Wt, bt, L, K, 'tanh', 4, 'max')


# --------------------------------------------------------------------------- unfused adjoint recurrence (wide gy)
@pytest.mark.parametrize('N,Fin,Fout,K', [(2, 32, 128, 2), (3, 32, 128, 4), (2, 32, 160, 5), (2, 64, 256, 3)])
def test_unfused_clenshaw_input_gradient(ops, tf_ref, N, Fin, Fout, K):
    """dx when neither fused kernel takes the shape (32x32 grid, M = 1024: the slabs do not fit shared memory) and gy
    is wider than dx: G = gy W^T by one GEMM, then K-1 batched adjoint steps at the width of dx -- against the oracle,
    and against the forward-form path (CG_FILTER_NO_CLENSHAW) on the same inputs."""
    from oracle import graph_ref
    A = graph_ref.adjacency(*graph_ref.distance_sklearn_metrics(graph_ref.grid(32), k=8, metric='euclidean'))
    L = graph_ref.laplacian(A, normalized=True)
    M = L.shape[0]
    rng = np.random.RandomState(7 * K + Fin)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
    got = []
    for flags in (ops.FILTER_DEFAULT, ops.FILTER_NO_CLENSHAW):
        xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
        y = ops.cheb_filter(xt, Wt, L, K, flags=flags)
        y.backward(dev(gy))
        close(xt.grad, dx)
        close(Wt.grad, dW)
        got.append(xt.grad.cpu().numpy())
    close(got[0], got[1], 2e-5)


def test_second_backward_through_retained_graph(ops, tf_ref, c2):
    """retain_graph=True: the saved basis lives in autograd's saved tensors, so a second backward through the same
    graph gives the same gradients (filter node and fused first-layer node)."""
    L = csr_from(c2, 'L2')
    M = L.shape[0]
    rng = np.random.RandomState(11)
    N, Fin, Fout, K = 9, 32, 64, 6
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
    y = ops.cheb_filter(xt, Wt, L, K)
    g1 = torch.autograd.grad(y, (xt, Wt), dev(gy), retain_graph=True)
    g2 = torch.autograd.grad(y, (xt, Wt), dev(gy))
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)
    close(g2[0], dx)
    close(g2[1], dW)
    L0 = csr_from(c2, 'L0')
    M0 = L0.shape[0]
    x0 = rng.standard_normal((16, M0, 1)).astype(np.float32)
    W0 = dev((0.1 * rng.standard_normal((25, 32))).astype(np.float32)).requires_grad_(True)
    b0 = dev((0.1 * rng.standard_normal(32)).astype(np.float32)).requires_grad_(True)
    if ops.first_layer_supported(dev(x0), W0, b0, L0, 25, 'relu', 4, 'max'):
        yp = ops.first_layer(dev(x0), W0, b0, L0, 25)
        g = dev(rng.standard_normal(tuple(yp.shape)).astype(np.float32))
        h1 = torch.autograd.grad(yp, (W0, b0), g, retain_graph=True)
        h2 = torch.autograd.grad(yp, (W0, b0), g)
        for a, b in zip(h1, h2):
            assert torch.equal(a, b)


def _plan_info():
    import ctypes
    from cnn_graph_b200 import _native
    buf = (ctypes.c_int * 8)()
    _native.check(_native.lib().cg_debug_fused_plan_info(buf), 'cg_debug_fused_plan_info')
    return list(buf)


# (fixture, operator, N, Fin, Fout, K): one and two block items per thread, every lanes-per-row variant, ragged last
# groups, a vertex count that is not a multiple of 4 (directed57), an operator without locality (c1)
BLOCK_SHAPES = [('c2', 'L2', 7, 32, 64, 25), ('c2', 'L2', 5, 16, 32, 6), ('c2', 'L2', 3, 64, 64, 5), ('c2', 'L3', 2, 128, 32, 3),
                ('c2', 'L3', 11, 32, 64, 7), ('c2', 'L3', 9, 16, 16, 4), ('c2', 'L4', 50, 32, 32, 5), ('c2', 'L4', 33, 64, 16, 3),
                ('directed', 'L', 13, 16, 32, 6), ('directed', 'L', 5, 32, 16, 4), ('c1', 'L2', 100, 32, 64, 20), ('c1', 'L1', 37, 16, 32, 5)]


@pytest.mark.parametrize('fix,name,N,Fin,Fout,K', BLOCK_SHAPES)
def test_row_block_gather_kernels_vs_oracle(ops, tf_ref, request, fix, name, N, Fin, Fout, K):
    """k_cheb_fused_b / k_cheb_clenshaw_b (4-row union blocks) forced on: y, dx, dW and the saved basis path against the
    oracle, and bit-equal y / dx between the two gather forms is NOT expected (different summation order) -- both sit
    inside the fp32 tolerance."""
    import os
    npz = request.getfixturevalue(fix)
    L = csr_from(npz, name)
    M = L.shape[0]
    lmax = 3.5 if fix == 'directed' else 2
    rng = np.random.RandomState(N + Fin + K)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    os.environ['CG_FUSED_BLOCK'] = '1'
    try:
        xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
        y = ops.cheb_filter(xt, Wt, L, K, lmax=lmax, flags=ops.FILTER_FORCE_FUSED)
        info_f = _plan_info()
        y.backward(dev(gy))
        info_b = _plan_info()
    finally:
        os.environ.pop('CG_FUSED_BLOCK', None)
    # wide layers whose tables + slabs exceed shared memory fall back to the row-per-item kernels (M = 248 with Fin >= 64)
    if Fin <= 32 or M <= 128:
        assert info_f[0] == 1, 'forward did not take the row-block kernel: %r' % info_f
        assert info_b[4] == 1, 'dx did not take the row-block Clenshaw kernel: %r' % info_b
    close(y, tf_ref.chebyshev5(x, L, W, K, lmax=lmax))
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy, lmax=lmax)
    close(xt.grad, dx)
    close(Wt.grad, dW)


def test_native_softmax_xent_vs_torch(ops):
    """cg_softmax_xent (loss + gradient in one launch) against torch's cross-entropy in float64: value 1e-6, gradient
    1e-6 of its largest entry; class counts of the three cgcnn configs and a ragged batch."""
    rng = np.random.RandomState(0)
    for N, C in ((100, 3), (1024, 10), (37, 20), (4096, 10), (5, 257)):
        z = (3.0 * rng.standard_normal((N, C))).astype(np.float32)
        y = rng.randint(0, C, size=N).astype(np.int64)
        zt = dev(z).requires_grad_(True)
        loss = ops.softmax_xent(zt, dev(y))
        (2.5 * loss).backward()
        zr = torch.from_numpy(z).double().requires_grad_(True)
        ref = torch.nn.functional.cross_entropy(zr, torch.from_numpy(y))
        (2.5 * ref).backward()
        assert abs(float(loss) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
        close(zt.grad, zr.grad.numpy(), 2e-6)


def test_native_momentum_sgd_matches_torch(ops):
    """cg_sgd_momentum (every variable in one launch) follows torch.optim.SGD(momentum) step for step -- bitwise up to
    the fused multiply-adds (1e-6) -- on aligned, odd-sized and tiny tensors, with and without momentum."""
    rng = np.random.RandomState(1)
    shapes = [(3968, 512), (512,), (25, 32), (7,), (1,), (800, 64), (13, 5)]
    for momentum in (0.9, 0.0):
        init = [rng.standard_normal(sh).astype(np.float32) for sh in shapes]
        pa = [dev(a).requires_grad_(True) for a in init]
        pb = [dev(a).requires_grad_(True) for a in init]
        oa = ops.NativeMomentumSGD(pa, lr=0.05, momentum=momentum)
        ob = torch.optim.SGD(pb, lr=0.05, momentum=momentum)
        for step in range(4):
            grads = [rng.standard_normal(sh).astype(np.float32) for sh in shapes]
            for p, q, g in zip(pa, pb, grads):
                p.grad = dev(g)
                q.grad = dev(g)
            if step == 2:
                for grp in oa.param_groups + ob.param_groups:
                    grp['lr'] = 0.02
            oa.step()
            ob.step()
            for p, q in zip(pa, pb):
                close(p, q.detach().cpu().numpy(), 1e-6)


def test_step_tiles_and_halo_pull_single_process(ops):
    """The pieces of the peer-memory halo exchange that one process can check: cg_cheb_step_tiles over two disjoint tile
    lists equals cg_cheb_step bit for bit, and cg_halo_pull gathers rows through a table of buffer pointers (here: two
    local buffers standing in for two ranks)."""
    import ctypes
    from cnn_graph_b200 import _native
    lib = _native.lib()
    rng = np.random.RandomState(4)
    n, C = 1500, 64
    pts = rng.rand(n, 2)
    order = np.lexsort((pts[:, 1], (pts[:, 0] * 12).astype(int)))          # strips: neighbours stay close in index
    pts = pts[order]
    import scipy.spatial
    _, idx = scipy.spatial.cKDTree(pts).query(pts, k=9)
    rows = np.repeat(np.arange(n), 8)
    A = scipy.sparse.csr_matrix((np.full(n * 8, -0.1, np.float32), (rows, idx[:, 1:].ravel())), shape=(n, n))
    A = ((A + A.T) * 0.5).tocsr().astype(np.float32)
    A.sum_duplicates()
    A.sort_indices()
    h = ops.GraphHandle(A)
    tr = lib.cg_cheb_step_tile_rows(h.handle, 0, C)
    assert tr > 0, 'tiled step expected for a locality-ordered operator'
    x1 = dev(rng.standard_normal((n, C)).astype(np.float32))
    x0 = dev(rng.standard_normal((n, C)).astype(np.float32))
    ref, got = torch.empty_like(x1), torch.zeros_like(x1)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _native.check(lib.cg_cheb_step(h.handle, 0, x1.data_ptr(), x0.data_ptr(), ref.data_ptr(), n, C, ctypes.c_float(2.0), stream), 'step')
    ntiles = -(-n // tr)
    perm = rng.permutation(ntiles).astype(np.int32)
    for part in (perm[:ntiles // 2], perm[ntiles // 2:]):
        t = dev(np.ascontiguousarray(part))
        _native.check(lib.cg_cheb_step_tiles(h.handle, 0, x1.data_ptr(), x0.data_ptr(), got.data_ptr(), n, C, ctypes.c_float(2.0),
                                             t.data_ptr(), t.numel(), stream), 'step_tiles')
    assert torch.equal(ref, got)
    # oracle: 2 A x1 - x0
    close(ref, 2.0 * (A @ x1.cpu().numpy()) - x0.cpu().numpy(), 1e-5)
    # halo pull: rows of two "ranks"
    bufs = [dev(rng.standard_normal((3, 40, C)).astype(np.float32)) for _ in range(2)]
    table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device='cuda')
    src_rank = rng.randint(0, 2, size=25).astype(np.int32)
    src_row = rng.randint(0, 40, size=25).astype(np.int32)
    dst = torch.zeros((25, C), device='cuda')
    k = 2
    d_rank, d_row = dev(src_rank), dev(src_row)        # kept alive: the launch is asynchronous
    _native.check(lib.cg_halo_pull(table.data_ptr(), d_rank.data_ptr(), d_row.data_ptr(), k * 40 * C, dst.data_ptr(), 25, C,
                                   stream), 'halo_pull')
    torch.cuda.synchronize()
    want = np.stack([bufs[r][k, j].cpu().numpy() for r, j in zip(src_rank, src_row)])
    assert np.array_equal(dst.cpu().numpy(), want)


def test_lstm_gates_two_addends_and_zero_state_skip(ops, tf_ref, c2):
    """The gate kernels sum the x-path and h-path pre-activations themselves (same values and gradients as the sum formed
    outside), and a cell step on the marked all-zero initial state (its h-path filter is skipped: the filter is linear)
    equals the step on an ordinary zero tensor."""
    rng = np.random.RandomState(2)
    N, M, H = 3, 40, 16
    px, ph = (rng.standard_normal((N, M, 4 * H)).astype(np.float32) * 0.5 for _ in range(2))
    b = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
    c = rng.standard_normal((N, M, H)).astype(np.float32)
    gh, gc = (rng.standard_normal((N, M, H)).astype(np.float32) for _ in range(2))
    for variant in ('fork', 'standard'):
        outs = []
        for two in (True, False):
            tx, th, tb, tc = (dev(a).requires_grad_(True) for a in (px, ph, b, c))
            if two:
                nh, nc = ops.lstm_gates(tx, tb, tc, variant, pre2=th)
            else:
                nh, nc = ops.lstm_gates(tx + th, tb, tc, variant)
            torch.autograd.backward([nh, nc], [dev(gh), dev(gc)])
            outs.append([t.detach().cpu().numpy() for t in (nh, nc, tx.grad, th.grad, tb.grad, tc.grad)])
        for a, r in zip(*outs):
            close(a, r, 2e-6)
    from cnn_graph_b200.lib import gconv_lstm, variables
    L = csr_from(c2, 'L3')
    Mg = L.shape[0]
    x = dev(rng.standard_normal((2, Mg, 2)).astype(np.float32))
    res = []
    for marked in (True, False):
        torch.manual_seed(0)
        store = variables.VariableStore(torch.device('cuda'))
        with variables.use_store(store):
            cell = gconv_lstm.GConvLSTMCell(num_units=16, laplacian=L, lmax=2, K=3, feat_in=2, nNode=Mg, gate_variant='standard')
            state = cell.zero_state(2) if marked else (torch.zeros(2, Mg, 16, device='cuda'), torch.zeros(2, Mg, 16, device='cuda'))
            assert ops.is_marked_zero(state[1]) == marked
            h1, st1 = cell(x, state)
            h2, _ = cell(x, (st1.c, st1.h))
        res.append((h1.detach().cpu().numpy(), h2.detach().cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0])
    close(res[0][1], res[1][1], 1e-6)


BF16_RTOL = 2e-2       # BASELINE.json north_star: "within rtol 1e-4 fp32 / 2e-2 bf16 for filter outputs and gradients"


@pytest.mark.parametrize('level,N,Fin,Fout,K,flags', [(2, 19, 32, 64, 25, 0), (3, 7, 16, 32, 5, 0), (2, 6, 32, 64, 9, 4), (0, 5, 1, 32, 25, 0),
                                                     (4, 33, 64, 512, 3, 0), (2, 4, 8, 16, 6, 1)])
def test_single_pass_bf16_mode_within_2e_2(ops, tf_ref, c2, level, N, Fin, Fout, K, flags):
    """Opt-in single-pass bf16 products (cg_set_precision): one tensor-core pass instead of three in every MMA kernel --
    fused forward / Clenshaw / dW kernels, the HBM-basis contraction, the pipelined GEMMs.  Outputs and gradients stay
    inside the bf16 tolerance, and the mode is really on (its error exceeds the fp32 mode's wherever tensor cores run)."""
    L = csr_from(c2, 'L%d' % level)
    M = L.shape[0]
    rng = np.random.RandomState(7 * level + K)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    ref_y = tf_ref.chebyshev5(x, L, W, K)
    ref_dx, ref_dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)

    def run():
        xt, Wt = dev(x).requires_grad_(True), dev(W).requires_grad_(True)
        y = ops.cheb_filter(xt, Wt, L, K, flags=flags)
        y.backward(dev(gy))
        return [t.detach().cpu().numpy() for t in (y, xt.grad, Wt.grad)]

    def rel(a, r):
        return float(np.abs(a.astype(np.float64) - r).max() / max(np.abs(r).max(), 1e-30))

    full = run()
    assert ops.get_precision() == 'fp32'
    ops.set_precision('bf16')
    try:
        assert ops.get_precision() == 'bf16'
        half = run()
    finally:
        ops.set_precision('fp32')
    for a, r in zip(half, (ref_y, ref_dx, ref_dW)):
        close(a, r, BF16_RTOL)
    for a, r in zip(full, (ref_y, ref_dx, ref_dW)):
        close(a, r, RTOL)
    if flags == 0 and Fin >= 16:        # fused kernel: the forward contraction certainly runs on the tensor cores
        assert rel(half[0], ref_y) > 4 * rel(full[0], ref_y), 'bf16 mode did not change the forward products'


def test_csr_densify_matches_scipy_toarray(ops):
    """cg_csr_densify: the CSR batch expanded on the device equals scipy's toarray() bit for bit (bag-of-words rows,
    duplicates summed, an empty row, odd widths, zero rows appended for a padded last batch)."""
    rng = np.random.RandomState(5)
    for rows, M, out_rows in ((100, 10000, 100), (7, 1001, 7), (3, 18, 5), (1, 4, 1)):
        nnz = max(1, min(M, 70))
        ri = np.repeat(np.arange(rows), nnz)
        ci = rng.randint(0, M, size=rows * nnz)
        va = rng.randint(1, 6, size=rows * nnz).astype(np.float32)
        if rows > 2:
            keep = ri != 1                      # an empty row
            ri, ci, va = ri[keep], ci[keep], va[keep]
        A = scipy.sparse.csr_matrix((va, (ri, ci)), shape=(rows, M))          # sums duplicates
        B = scipy.sparse.coo_matrix((va, (ri, ci)), shape=(rows, M)).tocsr()   # canonical as well
        assert (A != B).nnz == 0
        got = ops.sparse_batch_to_device(A, torch.device('cuda'), out_rows=out_rows).cpu().numpy()
        want = np.zeros((out_rows, M), np.float32)
        want[:rows] = A.toarray()
        assert np.array_equal(got, want)
    # un-canonical CSR with duplicate entries: summed like toarray()
    ip = dev(np.array([0, 3, 4], np.int32))
    ix = dev(np.array([2, 2, 0, 1], np.int32))
    va = dev(np.array([1.0, 2.0, 4.0, 8.0], np.float32))
    got = ops.csr_densify(ip, ix, va, 4).cpu().numpy()
    assert np.array_equal(got, np.array([[4, 0, 3, 0], [0, 8, 0, 0]], np.float32))


def test_c5_shaped_streaming_properties(ops):
    """BASELINE config C5's operator family at 2^17 vertices (Morton-ordered 16-NN graph, ~18 entries per row, C = 64;
    the bench runs 2^20): the three streaming step forms agree bit for bit (zero weights of the row blocks add exact
    zeros and the summation order is the stored column order in all of them), the basis is linear in X, and
    <T_k(L~) x, g> = <x, T_k(L~^T) g> for every k (no oracle needed at this size)."""
    import os
    from benchmarks import workloads
    L = workloads.knn_graph_laplacian(17, 16, 'morton')
    Lr = ops.rescale_csr(L, 2)
    h = ops.GraphHandle(Lr)
    M, C, K = Lr.shape[0], 64, 6
    gen = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(M, C, device='cuda', generator=gen)
    g = torch.randn(M, C, device='cuda', generator=gen)
    outs = {}
    for name, env in (('tiled', {}), ('blocks', {'CG_SPMM_TILE': '0'}), ('csr', {'CG_SPMM_BLOCK': '0'})):
        os.environ.update(env)
        try:
            outs[name] = ops.cheb_basis(h, x, K, flags=1)
        finally:
            for k in env:
                os.environ.pop(k, None)
    assert torch.equal(outs['tiled'], outs['blocks']) and torch.equal(outs['tiled'], outs['csr'])
    B = outs['tiled']
    x2 = torch.randn(M, C, device='cuda', generator=gen)
    B2 = ops.cheb_basis(h, x2, K, flags=1)
    B12 = ops.cheb_basis(h, 1.5 * x - 0.25 * x2, K, flags=1)
    close(B12, (1.5 * B - 0.25 * B2).cpu().numpy(), 2e-5)
    Bt = ops.cheb_basis(h, g, K, transpose=True, flags=1)
    for k in range(K):
        lhs = float((B[k].double() * g.double()).sum())
        rhs = float((x.double() * Bt[k].double()).sum())
        scale = float(B[k].double().norm() * g.double().norm())
        assert abs(lhs - rhs) <= 1e-6 * scale, (k, lhs, rhs)


def test_deferred_reducer_graphed_trajectory_matches_plain_steps(c2):
    """dist.DeferredGradAllReducer(force=True) on one process: the large gradients' update moved to the start of the next
    step (side stream, device-side learning rate, captured into the step graph) gives the same weights as plain eager
    steps -- across a learning-rate boundary of the staircase schedule and after flush()."""
    from cnn_graph_b200 import dist as cgdist
    rng = np.random.RandomState(3)
    xs = [dev(rng.uniform(0, 1, (16, 992)).astype(np.float32)) for _ in range(7)]
    ys = [dev(rng.randint(0, 10, 16).astype(np.int64)) for _ in range(7)]
    ref = _small_cgcnn(c2, 5)
    for x, y in zip(xs, ys):
        ref.train_step(x, y)
    for graphed in (False, True):
        m = _small_cgcnn(c2, 5)
        m.grad_hook = cgdist.DeferredGradAllReducer(m, min_numel=2000, force=True)
        assert m.grad_hook.active() and len(m.grad_hook.big) >= 1 and len(m.grad_hook.small) >= 1
        for x, y in zip(xs, ys):
            (m.train_step_graphed if graphed else m.train_step)(x, y)
        assert m.grad_hook.valid            # the last step's large gradient is still pending ...
        for name in ref.store.vars:         # ... get_var flushes it
            close(m.get_var(name), ref.get_var(name), 2e-5)
