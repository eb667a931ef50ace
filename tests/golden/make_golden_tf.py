#!/usr/bin/env python3
"""Generate tests/golden/tf_*.npz by EXECUTING the reference's own TensorFlow-side source files,
UNMODIFIED, under the torch-backed ``tensorflow`` stand-in of ``tf_shim.py``:

  /root/reference/lib/filter.py       cheby_conv, fourier_conv                     (:11-95)
  /root/reference/lib/models.py       cgcnn.chebyshev5 / chebyshev2 / fourier / b1relu / b1tanh / b2relu /
                                      mpool1 / apool1 / fc (:129-274); ``base_model`` (undefined in the fork,
                                      models.py:20) is supplied as the reference's own GraphModel
  /root/reference/lib/graph_conv.py   GraphConv._inference, residual_network, residual_layer, _STACK_NUM = 2
  /root/reference/lib/gconv_lstm.py   GConvLSTMCell.__call__, GconvModel.inference_* (all thirteen), glstm_layer
  /root/reference/lib/gconvRNN.py     gconvLSTMCell.__call__ (the standard gate variant, :123-219)

Run in the build container (``python tests/golden/make_golden_tf.py``); the reference tree does not exist
on the GPU box.  The files are loaded by path; nothing is copied from them, only their outputs on seeded
inputs are stored.  Gradients are ``torch.autograd`` through the reference's own op graph (stand-in for TF
autodiff).  Every case is run in float64 ("truth", stored rounded to float32: 6e-8 relative, far below the
1e-4 tolerance) on float32-representable inputs.  ``build_graph`` (placeholders, session, Adam) is not part
of the hot path and is overridden by a no-op in a subclass; constructors and every op method are the
reference's.
"""
import builtins
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import types
import warnings

import numpy as np
import scipy.sparse
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim as tf  # noqa: E402

REF = '/root/reference/lib'


def load_reference():
    warnings.filterwarnings('ignore')
    tf.install()
    tf.VERSION = '1.4.0'
    mpl = types.ModuleType('matplotlib')
    mpl.pyplot = types.ModuleType('matplotlib.pyplot')
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', mpl.pyplot)
    ipy = types.ModuleType('IPython')
    ipy.embed = lambda *a, **k: None
    sys.modules.setdefault('IPython', ipy)
    utils = types.ModuleType('utils')                      # gconvRNN.py:9 imports a helper that does not exist
    utils.show_all_variables = lambda *a, **k: None
    sys.modules['utils'] = utils
    sys.path.insert(0, REF)                                 # `import graph`, `import filter` (absolute, as the fork does)
    pkg = types.ModuleType('_reference_lib')
    pkg.__path__ = [REF]
    sys.modules['_reference_lib'] = pkg
    mods = {}
    with contextlib.redirect_stdout(io.StringIO()):
        mods['graph'] = importlib.import_module('graph')
        mods['filter'] = importlib.import_module('filter')
        mods['graph_model'] = importlib.import_module('_reference_lib.graph_model')
        builtins.base_model = mods['graph_model'].GraphModel          # lib/models.py:20 names an undefined base
        mods['models'] = importlib.import_module('_reference_lib.models')
        mods['graph_conv'] = importlib.import_module('_reference_lib.graph_conv')
        mods['gconv_lstm'] = importlib.import_module('_reference_lib.gconv_lstm')
        mods['gconvRNN'] = importlib.import_module('gconvRNN')
    return mods


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def csr_parts(prefix, A):
    A = scipy.sparse.csr_matrix(A)
    A.sort_indices()
    return {prefix + '_indptr': A.indptr.astype(np.int64), prefix + '_indices': A.indices.astype(np.int64),
            prefix + '_data': A.data.astype(np.float32), prefix + '_shape': np.array(A.shape, np.int64)}


def f32(t):
    return t.detach().to(torch.float32).cpu().numpy()


def T(a, grad=False):
    t = tf.convert(np.asarray(a, np.float32))
    return t.requires_grad_(True) if grad else t


def leaf(a, dtype=torch.float64):
    """Differentiable leaf of the shim's tensor type holding the float32 values of ``a``."""
    return tf.convert(torch.as_tensor(np.asarray(a, np.float32)).to(dtype)).requires_grad_(True)


def grid_laplacian(graph, side, k=8):
    z = graph.grid(side)
    dist, idx = graph.distance_sklearn_metrics(z, k=k, metric='euclidean')
    A = graph.adjacency(dist, idx)
    return graph.laplacian(A, normalized=True).astype(np.float32)


def knn_laplacian(graph, rng, M, k=6):
    """Random geometric kNN graph: a generic spectrum (no repeated eigenvalues, unlike grids)."""
    z = rng.uniform(size=(M, 3)).astype(np.float32)
    dist, idx = graph.distance_scipy_spatial(z, k=k, metric='euclidean')
    A = graph.adjacency(dist, idx)
    return graph.laplacian(A, normalized=True).astype(np.float32)


def copyL(L):
    return scipy.sparse.csr_matrix(L, copy=True)


# ------------------------------------------------------------------------------------------------
def gen_filter(mods, out_path):
    """lib/filter.py cheby_conv / fourier_conv: y and (dx, dW) of <y, gy>."""
    graph, flt = mods['graph'], mods['filter']
    rng = np.random.RandomState(2017)
    c2 = np.load(os.path.join(HERE, 'c2_grid28.npz'))
    c1 = np.load(os.path.join(HERE, 'c1_usage.npz'))
    d57 = np.load(os.path.join(HERE, 'directed57.npz'))

    def csr(z, p):
        return scipy.sparse.csr_matrix((z[p + '_data'], z[p + '_indices'], z[p + '_indptr']), shape=tuple(z[p + '_shape']))

    cases = {
        # name: (L, lmax, N, Fin, Fout, K)
        'c2l2': (csr(c2, 'L2').astype(np.float32), 2, 3, 32, 16, 25),      # MNIST level-2 operator, K = 25
        'c2l0': (csr(c2, 'L0').astype(np.float32), 2, 2, 1, 32, 25),       # MNIST level-0, scalar input
        'c1l0': (csr(c1, 'L0').astype(np.float32), 2, 4, 1, 32, 20),       # usage.ipynb level-0 (31 entries / row)
        'c1l2': (csr(c1, 'L2').astype(np.float32), 2, 4, 32, 8, 20),
        'c4': (grid_laplacian(graph, 8), 2, 3, 2, 16, 3),                    # humanflow-shaped grid, K = 3
        'c4h': (grid_laplacian(graph, 8), 2, 3, 8, 16, 2),
        'k1': (grid_laplacian(graph, 6), 2, 2, 3, 5, 1),
        'k2': (grid_laplacian(graph, 6), 2, 2, 3, 5, 2),
        'directed': (csr(d57, 'L').astype(np.float32), 3.5, 3, 4, 6, 6),     # non-symmetric operator, lmax != 2
    }
    out = {}
    for name, (L, lmax, N, Fin, Fout, K) in cases.items():
        M = L.shape[0]
        x = rng.standard_normal((N, M, Fin)).astype(np.float32)
        W = (0.1 * rng.standard_normal((K * Fin, Fout))).astype(np.float32)
        gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
        res = {}
        for tag, dt in (('', torch.float64), ('_f32', torch.float32))[:2 if name in ('c2l2', 'c4') else 1]:
            tf.set_dtype(dt)
            tf.reset_default_graph()
            xt, Wt = leaf(x, dt), leaf(W, dt)
            y = flt.cheby_conv(xt, copyL(L), lmax, Fout, K, Wt)        # cheby_conv rescales its L argument in place
            (y * torch.as_tensor(gy).to(dt)).sum().backward()
            res['y' + tag], res['dx' + tag], res['dW' + tag] = f32(y), f32(xt.grad), f32(Wt.grad)
        out.update(csr_parts(name + '_L', L))
        out[name + '_meta'] = np.array([lmax, N, Fin, Fout, K], np.float64)
        out[name + '_x'], out[name + '_W'], out[name + '_gy'] = x, W, gy
        for k_, v in res.items():
            out[name + '_' + k_] = v
    # default-W creation in the ambient scope (lib/filter.py:62-64)
    tf.set_dtype(torch.float64)
    g = tf.reset_default_graph()
    with tf.variable_scope('layer'):
        L = grid_laplacian(graph, 6)
        flt.cheby_conv(T(np.zeros((2, 36, 3))), copyL(L), 2, 5, 4)
    out['defaultW_names'] = np.array(sorted(g.variables.keys()))
    out['defaultW_shape'] = np.array(g.variables['layer/weights'].shape, np.int64)

    # fourier_conv (lib/filter.py:11-42) on a generic-spectrum graph; U stored so that a different LAPACK
    # build cannot change the comparison
    L = knn_laplacian(graph, rng, 40)
    N, Fin, Fout = 3, 4, 6
    x = rng.standard_normal((N, 40, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((40, Fout, Fin))).astype(np.float32)
    gy = rng.standard_normal((N, 40, Fout)).astype(np.float32)
    tf.reset_default_graph()
    xt, Wt = leaf(x), leaf(W)
    y = flt.fourier_conv(xt, L, 2, Fout, 40, Wt)
    (y * torch.as_tensor(gy).double()).sum().backward()
    out.update(csr_parts('fourier_L', L))
    out['fourier_U'] = graph.fourier(L)[1]
    out['fourier_x'], out['fourier_W'], out['fourier_gy'] = x, W, gy
    out['fourier_y'], out['fourier_dx'], out['fourier_dW'] = f32(y), f32(xt.grad), f32(Wt.grad)
    np.savez_compressed(out_path, **out)
    return out


# ------------------------------------------------------------------------------------------------
def bare(cls, **attrs):
    """Instance of a reference model class with the graph build switched off (see module docstring)."""
    sub = type(cls.__name__, (cls,), {'build_graph': lambda self, *a, **k: None})
    return sub, attrs


def gen_model_ops(mods, out_path):
    """cgcnn's op methods (lib/models.py:129-274), each on seeded inputs."""
    graph, models = mods['graph'], mods['models']
    rng = np.random.RandomState(7)
    tf.set_dtype(torch.float64)
    L = [grid_laplacian(graph, 6), knn_laplacian(graph, rng, 36)]
    Cg, _ = bare(models.cgcnn)
    m = quiet(Cg, [L[0]], [4], [3], [1], [2], filter='chebyshev5', brelu='b1relu', pool='mpool1', C_0=[3])
    out = {}
    out.update(csr_parts('L', L[0]))
    out.update(csr_parts('Lk', L[1]))
    N, M, Fin, Fout, K = 3, 36, 3, 5, 4
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    out['x'], out['gy'] = x, gy

    def run(fn, *args, wname='weights', bname='bias'):
        g = tf.reset_default_graph()
        xt = leaf(x)
        with tf.variable_scope('op'):
            y = fn(xt, *args)
        return g, xt, y

    for name in ('chebyshev5', 'chebyshev2'):
        g, xt, y = run(getattr(m, name), L[0], Fout, K)
        (y * torch.as_tensor(gy).double()).sum().backward()
        W = g.variables['op/weights']
        out[name + '_W'], out[name + '_y'], out[name + '_dW'] = f32(W), f32(y), f32(W.grad)
        if name == 'chebyshev5':
            out[name + '_dx'] = f32(xt.grad)
        else:
            assert xt.grad is None                 # tf.py_func has no gradient (lib/models.py:183)
    g, xt, y = run(m.fourier, L[1], Fout, M)
    (y * torch.as_tensor(gy).double()).sum().backward()
    W = g.variables['op/weights']
    out['fourier_U'] = graph.fourier(L[1])[1]
    out['fourier_W'], out['fourier_y'], out['fourier_dW'], out['fourier_dx'] = f32(W), f32(y), f32(W.grad), f32(xt.grad)
    # bias / activation: give the bias a non-trivial value first
    a = rng.standard_normal((N, M, Fout)).astype(np.float32)
    out['act_x'] = a
    for name in ('b1relu', 'b1tanh', 'b2relu'):
        g = tf.reset_default_graph()
        at = leaf(a)
        with tf.variable_scope('op'):
            if name != 'b1relu':
                shape = [1, 1, Fout] if name == 'b1tanh' else [1, M, Fout]
                b = tf.get_variable('bias', shape, tf.float32, initializer=tf.constant_initializer(0.1))
                with torch.no_grad():
                    b.copy_(torch.as_tensor(0.3 * rng.standard_normal(shape).astype(np.float32)))
                tf.get_variable_scope().reuse_variables()
            y = getattr(m, name)(at)
        (y * torch.as_tensor(gy).double()).sum().backward()
        out[name + '_y'], out[name + '_dx'] = f32(y), f32(at.grad)
        if name != 'b1relu':
            out[name + '_b'], out[name + '_db'] = f32(b), f32(b.grad)
        else:
            assert not g.variables               # the fork's b1relu has no bias (lib/models.py:229-235)
    for name in ('mpool1', 'apool1'):
        for p in (1, 2, 4):
            at = leaf(a)
            y = getattr(m, name)(at, p)
            gp = rng.standard_normal(tuple(y.shape)).astype(np.float32)
            (y * torch.as_tensor(gp).double()).sum().backward()
            out['%s_p%d_y' % (name, p)], out['%s_p%d_gy' % (name, p)], out['%s_p%d_dx' % (name, p)] = f32(y), gp, f32(at.grad)
    # fc
    xf = rng.standard_normal((N, 20)).astype(np.float32)
    gf = rng.standard_normal((N, 7)).astype(np.float32)
    for relu in (True, False):
        g = tf.reset_default_graph()
        xt = leaf(xf)
        with tf.variable_scope('fc'):
            y = m.fc(xt, 7, relu=relu)
        (y * torch.as_tensor(gf).double()).sum().backward()
        tag = 'fc_relu' if relu else 'fc_lin'
        W, b = g.variables['fc/weights'], g.variables['fc/bias']
        out[tag + '_W'], out[tag + '_b'], out[tag + '_y'] = f32(W), f32(b), f32(y)
        out[tag + '_dx'], out[tag + '_dW'], out[tag + '_db'] = f32(xt.grad), f32(W.grad), f32(b.grad)
    out['fc_x'], out['fc_gy'] = xf, gf
    np.savez_compressed(out_path, **out)


# ------------------------------------------------------------------------------------------------
def run_model(build, x, dtype=torch.float64, seed=2017):
    """Build a reference model in a fresh shim graph, run _inference on x, return outputs, all variables
    (creation order), gradients of <out, gy> and the dropout masks in call order."""
    tf.set_dtype(dtype)
    g = tf.reset_default_graph(seed)
    model = quiet(build)
    xt = leaf(x, dtype)
    y = quiet(model._inference, xt, 1.0)
    rng = np.random.RandomState(seed)
    gy = rng.standard_normal(tuple(y.shape)).astype(np.float32)
    (y * torch.as_tensor(gy).to(dtype)).sum().backward()
    res = {'x': np.asarray(x, np.float32), 'y': f32(y), 'gy': gy, 'dx': f32(xt.grad),
           'var_names': np.array(list(g.variables.keys()))}
    for i, (name, v) in enumerate(g.variables.items()):
        res['var%d' % i] = f32(v)
        res['grad%d' % i] = f32(v.grad) if v.grad is not None else np.zeros(tuple(v.shape), np.float32)
        res['hasgrad%d' % i] = np.array(v.grad is not None)
    for i, mk in enumerate(g.dropout_masks):
        res['mask%d' % i] = f32(mk)
    res['n_masks'] = np.array(len(g.dropout_masks))
    return res


def gen_graph_conv(mods, out_path):
    """GraphConv topologies (lib/graph_conv.py:234-330): ResGNN, plain, and the two-branch _STACK_NUM = 2 merge."""
    graph, gc = mods['graph'], mods['graph_conv']
    rng = np.random.RandomState(3)
    L = grid_laplacian(graph, 5)
    M = L.shape[0]
    Sub, _ = bare(gc.GraphConv)
    out = {}
    out.update(csr_parts('L', L))
    common = dict(F=[8], K=[3], p=[1], M=[2], _nfilter=8, _nres_layer_count=2, filter='chebyshev5', brelu='b1relu',
                  pool='mpool1', batch_size=2)
    cases = {
        'resgnn': dict(common, C_0=[6], model_name='ResGNN'),
        'plain': dict(common, C_0=[6], model_name='GNN'),
        'tanh': dict(common, C_0=[4], model_name='ResGNN', brelu='b1tanh'),
        'stack2': dict(common, C_0=[16], _STACK_NUM=2, model_name='ResGNN'),
    }
    for name, kw in cases.items():
        C = int(np.sum(kw['C_0']))
        x = rng.standard_normal((2, M, C)).astype(np.float32)
        res = run_model(lambda: Sub([L], **kw), x)
        for k_, v in res.items():
            out[name + '/' + k_] = v
    np.savez_compressed(out_path, **out)


INFER_FUNCS = ['inference_glstm', 'inference_glstm_period_no_expand', 'inference_gconv',
               'inference_gconv_period_no_expand', 'inference_gconv_period_expand', 'inference_glstm_gconv',
               'inference_glstm_gconv_no_expand', 'inference_glstm_gconv_split', 'inference_glstm_period_expand',
               'inference_glstm_period_expand_gconv1', 'inference_glstm_period_expand_gconv2',
               'inference_glstm_period_expand_gconv3']


def gen_lstm(mods, out_path):
    """GConvLSTMCell steps (fork gates, lib/gconv_lstm.py:77-221; standard gates, lib/gconvRNN.py:123-219) and
    every GconvModel.inference_* variant (lib/gconv_lstm.py:264-607)."""
    graph, gl, grnn = mods['graph'], mods['gconv_lstm'], mods['gconvRNN']
    rng = np.random.RandomState(11)
    L = grid_laplacian(graph, 5)
    M = L.shape[0]
    out = {}
    out.update(csr_parts('L', L))
    # ---- single cell steps with a non-zero state
    N, Fin, H, K = 3, 2, 6, 3
    x = rng.uniform(0, 1, (N, M, Fin)).astype(np.float32)
    c0 = (0.5 * rng.standard_normal((N, M, H))).astype(np.float32)
    h0 = (0.5 * rng.standard_normal((N, M, H))).astype(np.float32)
    gh = rng.standard_normal((N, M, H)).astype(np.float32)
    gc_ = rng.standard_normal((N, M, H)).astype(np.float32)
    out['cell_x'], out['cell_c'], out['cell_h'], out['cell_gh'], out['cell_gc'] = x, c0, h0, gh, gc_
    out['cell_meta'] = np.array([N, Fin, H, K], np.int64)
    for variant in ('fork', 'standard'):
        tf.set_dtype(torch.float64)
        g = tf.reset_default_graph(5)
        if variant == 'fork':
            cell = gl.GConvLSTMCell(num_units=H, laplacian=L, lmax=2, K=K, feat_in=Fin, nNode=M, filter_type='cheby_conv')
        else:
            cell = grnn.gconvLSTMCell(num_units=H, laplacian=L, lmax=2, K=K, feat_in=Fin, nNode=M)
        xt = leaf(x)
        ct = leaf(c0)
        ht = leaf(h0)
        new_h, state = cell(xt, (ct, ht))
        assert state.h is new_h
        (new_h * torch.as_tensor(gh).double()).sum().backward(retain_graph=True)
        (state.c * torch.as_tensor(gc_).double()).sum().backward()
        pre = variant + '_'
        out[pre + 'new_h'], out[pre + 'new_c'] = f32(new_h), f32(state.c)
        out[pre + 'dx'], out[pre + 'dc'], out[pre + 'dh'] = f32(xt.grad), f32(ct.grad), f32(ht.grad)
        out[pre + 'var_names'] = np.array(list(g.variables.keys()))
        for name, v in g.variables.items():
            short = name.split('/')[-1]
            out[pre + short], out[pre + 'd' + short] = f32(v), f32(v.grad)
    # ---- model variants
    Sub, _ = bare(gl.GconvModel)
    Tn = 2
    kw = dict(seq_num_closeness=Tn, seq_num_period=Tn, seq_num_trend=Tn, filter_num=4, conv_layer_num=1,
              filter='cheby_conv', batch_size=2, kernel_num=2, in_feature_num=2, out_feature_num=2,
              lstm_layer_count=2, num_hidden_conv=6)
    status = {}
    for fn in INFER_FUNCS:
        feats = 2 * Tn * 3 if ('period_expand' in fn or fn == 'inference_gconv_period_expand') else \
            (2 * Tn * 2 if fn == 'inference_glstm_gconv_split' else 2 * Tn)
        x = rng.uniform(0, 1, (2, M, feats)).astype(np.float32)
        try:
            res = run_model(lambda: Sub(L, feature_num=feats, infer_func=fn, **kw), x)
        except Exception as exc:      # noqa: BLE001 -- some variants are broken as shipped; record how
            status[fn] = '%s: %s' % (type(exc).__name__, exc)
            continue
        status[fn] = 'ok'
        for k_, v in res.items():
            out[fn + '/' + k_] = v
    out['infer_status'] = np.array(['%s => %s' % kv for kv in status.items()])
    for kv in status.items():
        print('  %-44s %s' % kv)
    np.savez_compressed(out_path, **out)


def main():
    mods = load_reference()
    gen_filter(mods, os.path.join(HERE, 'tf_filter.npz'))
    gen_model_ops(mods, os.path.join(HERE, 'tf_model_ops.npz'))
    gen_graph_conv(mods, os.path.join(HERE, 'tf_graph_conv.npz'))
    gen_lstm(mods, os.path.join(HERE, 'tf_lstm.npz'))
    for f in ('tf_filter', 'tf_model_ops', 'tf_graph_conv', 'tf_lstm'):
        print(f, os.path.getsize(os.path.join(HERE, f + '.npz')) // 1024, 'KB')


if __name__ == '__main__':
    main()
