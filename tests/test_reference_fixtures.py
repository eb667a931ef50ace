"""The drop-in modules (cnn_graph_b200.lib.*) against outputs of the reference's OWN source files.

tests/golden/tf_*.npz come from executing the unmodified /root/reference/lib/{filter,models,graph_conv,gconv_lstm,
gconvRNN}.py under the torch-backed `tensorflow` stand-in (tests/golden/make_golden_tf.py).  CPU part: every model
variant declares exactly the reference's variables (names and shapes, i.e. TF's scoping rules).  GPU part (-m gpu):
forward values and all gradients of the CUDA path, called through the reference-named Python surface -> ctypes ->
C ABI, within rtol 1e-4 (atol = 1e-4 * max|ref|) of the reference run.
"""
import contextlib
import io

import numpy as np
import pytest
import torch

from conftest import csr_from, load_golden

RTOL = 1e-4
FILTER_CASES = ['c2l2', 'c2l0', 'c1l0', 'c1l2', 'c4', 'c4h', 'k1', 'k2', 'directed']
GRAPH_CONV_CASES = {
    'resgnn': dict(C_0=[6], model_name='ResGNN'),
    'plain': dict(C_0=[6], model_name='GNN'),
    'tanh': dict(C_0=[4], model_name='ResGNN', brelu='b1tanh'),
    'stack2': dict(C_0=[16], _STACK_NUM=2, model_name='ResGNN'),
}
INFER_FUNCS = ['inference_glstm', 'inference_gconv', 'inference_gconv_period_no_expand', 'inference_gconv_period_expand',
               'inference_glstm_gconv', 'inference_glstm_gconv_no_expand', 'inference_glstm_gconv_split',
               'inference_glstm_period_expand', 'inference_glstm_period_expand_gconv1',
               'inference_glstm_period_expand_gconv2', 'inference_glstm_period_expand_gconv3']


def close(got, ref, rtol=RTOL, what=''):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-30) if ref.size else 1.0
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) if ref.size else 0.0
    assert err <= rtol * scale, '%s: max abs err %.3e > %.1e * %.3e' % (what, err, rtol, scale)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.fixture(scope='module')
def tff():
    return load_golden('tf_filter.npz')


@pytest.fixture(scope='module')
def tfo():
    return load_golden('tf_model_ops.npz')


@pytest.fixture(scope='module')
def tfg():
    return load_golden('tf_graph_conv.npz')


@pytest.fixture(scope='module')
def tfl():
    return load_golden('tf_lstm.npz')


def fixture_vars(z, case):
    names = [str(n) for n in z[case + '/var_names']]
    return {n: (z['%s/var%d' % (case, i)], z['%s/grad%d' % (case, i)], bool(z['%s/hasgrad%d' % (case, i)]))
            for i, n in enumerate(names)}


def build_graph_conv(z, case):
    from cnn_graph_b200.lib import graph_conv
    L = csr_from(z, 'L').astype(np.float32)
    kw = dict(F=[8], K=[3], p=[1], M=[2], _nfilter=8, _nres_layer_count=2, filter='chebyshev5', brelu='b1relu',
              pool='mpool1', batch_size=2)
    kw.update(GRAPH_CONV_CASES[case])
    return quiet(graph_conv.GraphConv, [L], **kw)


def build_gconv_model(z, fn):
    from cnn_graph_b200.lib import gconv_lstm
    L = csr_from(z, 'L').astype(np.float32)
    feats = int(z[fn + '/x'].shape[2])
    return gconv_lstm.GconvModel(L, seq_num_closeness=2, seq_num_period=2, seq_num_trend=2, filter_num=4, conv_layer_num=1,
                                 filter='cheby_conv', batch_size=2, kernel_num=2, in_feature_num=2, out_feature_num=2,
                                 lstm_layer_count=2, num_hidden_conv=6, feature_num=feats, infer_func=fn)


def assert_same_variables(model, ref_vars):
    ours = {n: tuple(p.shape) for n, p in model.store.vars.items()}
    theirs = {n: tuple(v[0].shape) for n, v in ref_vars.items()}
    assert ours == theirs, (sorted(set(ours) ^ set(theirs)), [(n, ours[n], theirs[n]) for n in ours if n in theirs and ours[n] != theirs[n]])


# ---------------------------------------------------------------------------------------------- CPU: declared variables
@pytest.mark.parametrize('case', list(GRAPH_CONV_CASES))
def test_graph_conv_declares_reference_variables(tfg, case):
    assert_same_variables(build_graph_conv(tfg, case), fixture_vars(tfg, case))


@pytest.mark.parametrize('fn', INFER_FUNCS)
def test_gconv_model_declares_reference_variables(tfl, fn):
    assert_same_variables(build_gconv_model(tfl, fn), fixture_vars(tfl, fn))


def test_cgcnn_upstream_variable_names():
    """Upstream cgcnn scopes its layers with tf.variable_scope('conv{i}') and tf.name_scope for the blocks inside, so
    the variables are conv1/weights, conv1/bias, ... (usage.ipynb reads them with get_var)."""
    import scipy.sparse
    from cnn_graph_b200.lib import models
    L = [scipy.sparse.identity(16, format='csr', dtype=np.float32) for _ in range(3)]
    m = models.cgcnn(L, F=[4, 6], K=[3, 3], p=[2, 2], M=[5, 3], batch_size=2)
    assert list(m.store.vars) == ['conv1/weights', 'conv1/bias', 'conv2/weights', 'conv2/bias', 'fc1/weights', 'fc1/bias',
                                  'logits/weights', 'logits/bias']
    assert m.get_var('conv1/weights').shape == (3, 4)


# ---------------------------------------------------------------------------------------------- GPU: values + gradients
gpu = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@gpu
@pytest.mark.parametrize('name', FILTER_CASES)
def test_cheby_conv_cuda_vs_reference_source(tff, name):
    from cnn_graph_b200.lib import filter as flt
    L = csr_from(tff, name + '_L')
    lmax, N, Fin, Fout, K = tff[name + '_meta']
    Fout, K = int(Fout), int(K)
    x = dev(tff[name + '_x']).requires_grad_(True)
    W = dev(tff[name + '_W']).requires_grad_(True)
    before = L.copy()
    y = flt.cheby_conv(x, L, float(lmax), Fout, K, W)
    y.backward(dev(tff[name + '_gy']))
    assert (L != before).nnz == 0
    close(y, tff[name + '_y'], what='y')
    close(x.grad, tff[name + '_dx'], what='dx')
    close(W.grad, tff[name + '_dW'], what='dW')


@gpu
def test_fourier_conv_cuda_vs_reference_source(tff):
    from cnn_graph_b200 import ops
    from cnn_graph_b200.lib import filter as flt
    L = csr_from(tff, 'fourier_L')
    x = dev(tff['fourier_x']).requires_grad_(True)
    W = dev(tff['fourier_W']).requires_grad_(True)
    y = ops.fourier_filter(x, W, L, U=dev(tff['fourier_U'].astype(np.float32)))
    y.backward(dev(tff['fourier_gy']))
    close(y, tff['fourier_y'], what='y')
    close(x.grad, tff['fourier_dx'], what='dx')
    close(W.grad, tff['fourier_dW'], what='dW')
    # the public entry point computes the eigenbasis itself (numpy eigh on this host)
    y2 = flt.fourier_conv(dev(tff['fourier_x']), L, 2, W.shape[1], L.shape[0], W.detach())
    close(y2, tff['fourier_y'], 1e-3, what='y (own eigenbasis)')


@gpu
def test_cgcnn_ops_cuda_vs_reference_source(tfo):
    """cgcnn.chebyshev5 / chebyshev2 / fourier / b1tanh / b2relu / mpool1 / apool1 / fc bound by name, as the reference
    binds them (lib/models.py:120-122), with the reference's variable values."""
    from cnn_graph_b200 import ops
    from cnn_graph_b200.lib import models, variables
    L, Lk = csr_from(tfo, 'L'), csr_from(tfo, 'Lk')
    m = models.cgcnn([L], F=[4], K=[3], p=[1], M=[2], batch_size=3)
    m.b1relu_has_bias = False                                    # the fork's b1relu (lib/models.py:229-235)
    gy = dev(tfo['gy'])

    def run(method, args, values, x_np, g):
        store = variables.VariableStore(device=torch.device('cuda'))
        x = dev(x_np).requires_grad_(True)
        with variables.use_store(store), variables.variable_scope('op'):
            for nm, val in values.items():
                p = variables.get_variable(nm, val.shape, variables.constant_initializer(0.0))
                p.data.copy_(dev(val))
            y = getattr(m, method)(x, *args)
        y.backward(g)
        return x, y, store.vars

    for name in ('chebyshev5', 'chebyshev2'):
        W = tfo[name + '_W']
        x, y, v = run(name, (L, W.shape[1], W.shape[0] // tfo['x'].shape[2]), {'weights': W}, tfo['x'], gy)
        close(y, tfo[name + '_y'], what=name)
        close(v['op/weights'].grad, tfo[name + '_dW'], what=name + ' dW')
        if name == 'chebyshev5':
            close(x.grad, tfo[name + '_dx'], what=name + ' dx')
        else:
            assert x.grad is None                                # tf.py_func: no gradient to x (lib/models.py:183)
    ops._fourier_cache.clear()
    Wf = tfo['fourier_W']
    x, y, v = run('fourier', (Lk, Wf.shape[1], Lk.shape[0]), {'weights': Wf}, tfo['x'], gy)
    close(y, tfo['fourier_y'], 1e-3, what='fourier')             # eigenbasis recomputed on this host
    a = tfo['act_x']
    x, y, v = run('b1relu', (), {}, a, gy)
    close(y, tfo['b1relu_y']); close(x.grad, tfo['b1relu_dx'])
    for name in ('b1tanh', 'b2relu'):
        x, y, v = run(name, (), {'bias': tfo[name + '_b']}, a, gy)
        close(y, tfo[name + '_y'], what=name)
        close(x.grad, tfo[name + '_dx'], what=name + ' dx')
        close(v['op/bias'].grad, tfo[name + '_db'], what=name + ' db')
    for name in ('mpool1', 'apool1'):
        for p in (1, 2, 4):
            x, y, v = run(name, (p,), {}, a, dev(tfo['%s_p%d_gy' % (name, p)]))
            if name == 'mpool1':
                assert np.array_equal(y.detach().cpu().numpy(), tfo['mpool1_p%d_y' % p])      # bit-exact
                assert np.array_equal(x.grad.cpu().numpy(), tfo['mpool1_p%d_dx' % p])
            else:
                close(y, tfo['apool1_p%d_y' % p], 1e-6)
                close(x.grad, tfo['apool1_p%d_dx' % p], 1e-6)
    for tag, relu in (('fc_relu', True), ('fc_lin', False)):
        store = variables.VariableStore(device=torch.device('cuda'))
        x = dev(tfo['fc_x']).requires_grad_(True)
        with variables.use_store(store), variables.variable_scope('fc'):
            for nm in ('weights', 'bias'):
                val = tfo[tag + ('_W' if nm == 'weights' else '_b')]
                variables.get_variable(nm, val.shape, variables.constant_initializer(0.0)).data.copy_(dev(val))
            y = m.fc(x, 7, relu=relu)
        y.backward(dev(tfo['fc_gy']))
        close(y, tfo[tag + '_y'], what=tag)
        close(x.grad, tfo[tag + '_dx'], what=tag + ' dx')
        close(store.vars['fc/weights'].grad, tfo[tag + '_dW'], what=tag + ' dW')
        close(store.vars['fc/bias'].grad, tfo[tag + '_db'], what=tag + ' db')


def run_model_against_fixture(model, z, case):
    ref = fixture_vars(z, case)
    assert_same_variables(model, ref)
    with torch.no_grad():
        for n, p in model.store.vars.items():
            p.copy_(dev(ref[n][0]))
    n_masks = int(z[case + '/n_masks']) if case + '/n_masks' in z.files else 0
    if n_masks:
        model.dropout_masks = [dev(z['%s/mask%d' % (case, i)]) for i in range(n_masks)]
    x = dev(z[case + '/x']).requires_grad_(True)
    y = model.inference(x, 1.0)
    if n_masks:
        assert model._mask_cursor == n_masks                      # same number of DropoutWrapper calls as the reference
    y.backward(dev(z[case + '/gy']))
    close(y, z[case + '/y'], what=case + ' y')
    close(x.grad, z[case + '/dx'], what=case + ' dx')
    for n, p in model.store.vars.items():
        val, grad, has = ref[n]
        if has:
            close(p.grad, grad, what='%s d(%s)' % (case, n))
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n


@gpu
@pytest.mark.parametrize('case', list(GRAPH_CONV_CASES))
def test_graph_conv_cuda_vs_reference_source(tfg, case):
    """GraphConv._inference: ResGNN / plain / b1tanh residual networks and the _STACK_NUM = 2 merge
    (lib/graph_conv.py:234-330), values and every gradient."""
    run_model_against_fixture(build_graph_conv(tfg, case), tfg, case)


@gpu
@pytest.mark.parametrize('fn', INFER_FUNCS)
def test_gconv_model_cuda_vs_reference_source(tfl, fn):
    """Every GconvModel.inference_* the reference can run (lib/gconv_lstm.py:264-607), two stacked GConvLSTMCells under
    the reference's dropout masks, values and every gradient."""
    run_model_against_fixture(build_gconv_model(tfl, fn), tfl, fn)


@gpu
@pytest.mark.parametrize('variant', ['fork', 'standard'])
def test_lstm_cell_cuda_vs_reference_source(tfl, variant):
    """GConvLSTMCell.__call__ on a non-zero state (fork gates lib/gconv_lstm.py:185-215; standard gates
    lib/gconvRNN.py:189-213): new_h, new_c and the gradients w.r.t. x, c, h and all twelve variables."""
    from cnn_graph_b200.lib import gconv_lstm, variables
    L = csr_from(tfl, 'L').astype(np.float32)
    N, Fin, H, K = (int(v) for v in tfl['cell_meta'])
    pre = variant + '_'
    store = variables.VariableStore(device=torch.device('cuda'))
    cell = gconv_lstm.GConvLSTMCell(num_units=H, laplacian=L, lmax=2, K=K, feat_in=Fin, nNode=L.shape[0],
                                    filter_type='cheby_conv', gate_variant=variant)
    x = dev(tfl['cell_x']).requires_grad_(True)
    c = dev(tfl['cell_c']).requires_grad_(True)
    h = dev(tfl['cell_h']).requires_grad_(True)
    names = ['W%sxt' % g for g in 'zifo'] + ['W%sht' % g for g in 'zifo'] + ['b%st' % g for g in 'zifo']
    with variables.use_store(store):
        with variables.variable_scope('GConvLSTMCell'):
            for nm in names:
                val = tfl[pre + nm]
                variables.get_variable(nm, val.shape, variables.constant_initializer(0.0)).data.copy_(dev(val))
        new_h, state = cell(x, (c, h))
    assert state.h is new_h
    torch.autograd.backward([new_h, state.c], [dev(tfl['cell_gh']), dev(tfl['cell_gc'])])
    close(new_h, tfl[pre + 'new_h'], what='new_h')
    close(state.c, tfl[pre + 'new_c'], what='new_c')
    close(x.grad, tfl[pre + 'dx'], what='dx')
    close(c.grad, tfl[pre + 'dc'], what='dc')
    close(h.grad, tfl[pre + 'dh'], what='dh')
    for nm in names:
        close(store.vars['GConvLSTMCell/' + nm].grad, tfl[pre + 'd' + nm], what='d' + nm)
