"""Pins the TF half of the oracle (oracle/tf_ref.py) to outputs of the reference's OWN source files.

tests/golden/tf_*.npz were produced in the build container by executing the unmodified
/root/reference/lib/{filter,models,graph_conv,gconv_lstm,gconvRNN}.py under a torch-backed `tensorflow`
stand-in (tests/golden/tf_shim.py, generator tests/golden/make_golden_tf.py); gradients are torch.autograd
through the reference's own op graph.  Tolerance: rtol 1e-4 with atol = 1e-4 * max|ref| (north star, fp32).
"""
import numpy as np
import pytest

from conftest import csr_from, load_golden
from oracle import tf_ref

RTOL = 1e-4
FILTER_CASES = ['c2l2', 'c2l0', 'c1l0', 'c1l2', 'c4', 'c4h', 'k1', 'k2', 'directed']


def close(got, ref, rtol=RTOL):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-30)
    err = float(np.abs(got - ref).max())
    assert err <= rtol * scale, 'max abs err %.3e > %.1e * %.3e' % (err, rtol, scale)


@pytest.fixture(scope='module')
def tff():
    return load_golden('tf_filter.npz')


@pytest.fixture(scope='module')
def tfo():
    return load_golden('tf_model_ops.npz')


@pytest.fixture(scope='module')
def tfl():
    return load_golden('tf_lstm.npz')


@pytest.mark.parametrize('name', FILTER_CASES)
def test_cheby_conv_matches_reference_source(tff, name):
    """lib/filter.py:45-95 executed as is: layouts (column fin*N+n, the (3,1,2,0) restack, W row fin*K+k),
    K = 1 / K = 2 edge cases, lmax != 2 on a directed operator, and the autodiff gradients."""
    L = csr_from(tff, name + '_L')
    lmax, N, Fin, Fout, K = tff[name + '_meta']
    N, Fin, Fout, K = int(N), int(Fin), int(Fout), int(K)
    x, W, gy = tff[name + '_x'], tff[name + '_W'], tff[name + '_gy']
    before = L.copy()
    y = tf_ref.cheby_conv(x, L, lmax, Fout, K, W)
    assert (L != before).nnz == 0                      # the oracle never rescales the caller's L
    close(y, tff[name + '_y'])
    close(tf_ref.chebyshev5(x, L, W, K, lmax), tff[name + '_y'])
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy, lmax)
    close(dx, tff[name + '_dx'])
    close(dW, tff[name + '_dW'])
    if name + '_y_f32' in tff.files:                   # the reference graph evaluated in float32 agrees with its float64 run
        close(tff[name + '_y_f32'], tff[name + '_y'])
        close(tff[name + '_dW_f32'], tff[name + '_dW'])


def test_cheby_conv_creates_weights_in_ambient_scope(tff):
    assert list(tff['defaultW_names']) == ['layer/weights']          # lib/filter.py:62-64
    assert tuple(tff['defaultW_shape']) == (4 * 3, 5)                 # [K * feat_in, feat_out]


def test_fourier_conv_matches_reference_source(tff):
    L = csr_from(tff, 'fourier_L')
    y = tf_ref.fourier_conv(tff['fourier_x'], L, tff['fourier_W'], U=tff['fourier_U'])
    close(y, tff['fourier_y'])
    close(tf_ref.fourier_conv(tff['fourier_x'], L, tff['fourier_W']), tff['fourier_y'], 1e-3)   # this host's eigh


def test_cgcnn_filters_match_reference_source(tfo):
    """cgcnn.chebyshev5 / chebyshev2 (lib/models.py:161-224)."""
    L = csr_from(tfo, 'L')
    x, gy = tfo['x'], tfo['gy']
    for name in ('chebyshev5', 'chebyshev2'):
        W = tfo[name + '_W']
        K = W.shape[0] // x.shape[2]
        fn = tf_ref.chebyshev5 if name == 'chebyshev5' else tf_ref.chebyshev2
        close(fn(x, L, W, K), tfo[name + '_y'])
        dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
        close(dW, tfo[name + '_dW'])
        if name == 'chebyshev5':
            close(dx, tfo[name + '_dx'])
    Lk = csr_from(tfo, 'Lk')
    close(tf_ref.fourier_conv(x, Lk, tfo['fourier_W'], U=tfo['fourier_U']), tfo['fourier_y'])


def test_bias_activation_pool_fc_match_reference_source(tfo):
    a, gy = tfo['act_x'], tfo['gy']
    close(tf_ref.b1relu(a), tfo['b1relu_y'])                                   # fork: no bias
    close(gy * (a > 0), tfo['b1relu_dx'])
    close(tf_ref.b1tanh(a, tfo['b1tanh_b']), tfo['b1tanh_y'])
    close(gy * (1 - tfo['b1tanh_y'] ** 2), tfo['b1tanh_dx'])
    close((gy * (1 - tfo['b1tanh_y'] ** 2)).sum(axis=(0, 1)).reshape(1, 1, -1), tfo['b1tanh_db'])
    close(tf_ref.b2relu(a, tfo['b2relu_b']), tfo['b2relu_y'])
    close((gy * (tfo['b2relu_y'] > 0)).sum(axis=0, keepdims=True), tfo['b2relu_db'])
    for p in (1, 2, 4):
        y = tf_ref.mpool1(a, p)
        assert np.array_equal(y, tfo['mpool1_p%d_y' % p])                      # exact: max of float32 values
        assert np.array_equal(tf_ref.mpool1_backward(a, p, tfo['mpool1_p%d_gy' % p]), tfo['mpool1_p%d_dx' % p])
        close(tf_ref.apool1(a, p), tfo['apool1_p%d_y' % p], 1e-6)
        close(tf_ref.apool1_backward(a, p, tfo['apool1_p%d_gy' % p]), tfo['apool1_p%d_dx' % p], 1e-6)
    for tag, relu in (('fc_relu', True), ('fc_lin', False)):
        close(tf_ref.fc(tfo['fc_x'], tfo[tag + '_W'], tfo[tag + '_b'], relu=relu), tfo[tag + '_y'])


@pytest.mark.parametrize('variant', ['fork', 'standard'])
def test_lstm_step_matches_reference_source(tfl, variant):
    """GConvLSTMCell.__call__ (lib/gconv_lstm.py:77-221: z = tan, o = tanh) and gconvRNN's gconvLSTMCell
    (lib/gconvRNN.py:123-219: z = tanh, o = sigmoid), executed as is on a non-zero state."""
    L = csr_from(tfl, 'L')
    N, Fin, H, K = (int(v) for v in tfl['cell_meta'])
    pre = variant + '_'
    Wx = {g: tfl[pre + 'W%sxt' % g] for g in 'zifo'}
    Wh = {g: tfl[pre + 'W%sht' % g] for g in 'zifo'}
    b = {g: tfl[pre + 'b%st' % g] for g in 'zifo'}
    new_h, new_c = tf_ref.gconv_lstm_step(tfl['cell_x'], tfl['cell_c'], tfl['cell_h'], L, 2, K, Wx, Wh, b, variant)
    close(new_h, tfl[pre + 'new_h'])
    close(new_c, tfl[pre + 'new_c'])


def test_reference_inference_variants_recorded(tfl):
    """Which GconvModel.inference_* run in the reference as shipped (lib/gconv_lstm.py:264-607)."""
    status = dict(s.split(' => ') for s in tfl['infer_status'])
    assert len(status) == 12
    broken = {k for k, v in status.items() if v != 'ok'}
    assert broken == {'inference_glstm_period_no_expand'}            # float reshape at lib/gconv_lstm.py:288


def test_torch_mirror_matches_reference_source(tff, tfl):
    """oracle/torch_ref.py (the CPU arm of bench.py --config c4|c5) against the same fixtures."""
    import torch
    from oracle import torch_ref
    for name in ('c4', 'directed', 'k1'):
        L = csr_from(tff, name + '_L')
        lmax, N, Fin, Fout, K = tff[name + '_meta']
        x = torch.tensor(tff[name + '_x'], requires_grad=True)
        W = torch.tensor(tff[name + '_W'], requires_grad=True)
        y = torch_ref.cheby_conv(x, torch_ref.sparse_operator(L, lmax), int(K), W)
        y.backward(torch.tensor(tff[name + '_gy']))
        close(y.detach().numpy(), tff[name + '_y'])
        close(x.grad.numpy(), tff[name + '_dx'])
        close(W.grad.numpy(), tff[name + '_dW'])
    L = csr_from(tfl, 'L')
    N, Fin, H, K = (int(v) for v in tfl['cell_meta'])
    Ls = torch_ref.sparse_operator(L, 2)
    for variant in ('fork', 'standard'):
        pre = variant + '_'
        Wx = {g: torch.tensor(tfl[pre + 'W%sxt' % g]) for g in 'zifo'}
        Wh = {g: torch.tensor(tfl[pre + 'W%sht' % g]) for g in 'zifo'}
        b = {g: torch.tensor(tfl[pre + 'b%st' % g]) for g in 'zifo'}
        x = torch.tensor(tfl['cell_x'], requires_grad=True)
        h, c = torch_ref.lstm_cell(x, torch.tensor(tfl['cell_c']), torch.tensor(tfl['cell_h']), Ls, K, Wx, Wh, b, variant)
        torch.autograd.backward([h, c], [torch.tensor(tfl['cell_gh']), torch.tensor(tfl['cell_gc'])])
        close(h.detach().numpy(), tfl[pre + 'new_h'])
        close(c.detach().numpy(), tfl[pre + 'new_c'])
        close(x.grad.numpy(), tfl[pre + 'dx'])
