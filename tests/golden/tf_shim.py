"""A torch-backed stand-in for the slice of TensorFlow 1.x that the reference's hot path uses.

TEST INFRASTRUCTURE ONLY (used by ``make_golden_tf.py`` in the build container).  TensorFlow
is not installed and cannot be (no network), so nothing could execute the reference's own
``lib/filter.py``, ``lib/models.py``, ``lib/graph_conv.py`` and ``lib/gconv_lstm.py``.  This
module is registered as ``tensorflow`` in ``sys.modules`` so that those files run UNMODIFIED,
eagerly, on torch tensors: every ``tf.*`` call below is the plain mathematical definition of
the TF op of the same name (layout ops, dense / sparse matmul, elementwise functions, 1-D
pooling, variable scopes with TF's create / reuse rules, the static RNN plumbing).  Because
the tensors are torch tensors, ``torch.autograd`` differentiates the reference's own op graph,
which stands in for TF autodiff (``optimizer.compute_gradients``, lib/graph_model.py:296).

``set_dtype(torch.float64)`` makes every ``tf.float32`` request compute in float64 (the
"truth" fixtures); float32 reproduces TF's arithmetic type (not its summation order).

Nothing here is copied from TensorFlow or from the reference; it only implements the
documented semantics of the API names the reference calls.
"""
import contextlib
import sys
import types

import numpy as np
import torch

_DTYPE = [torch.float32]


def set_dtype(dtype):
    _DTYPE[0] = dtype


def _dt(requested=None):
    if requested in (None, 'float32', float32, torch.float32, np.float32):
        return _DTYPE[0]
    if requested is bool_:
        return torch.bool
    return requested


class _DType(object):
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return 'tf.' + self.name


float32 = _DType('float32')
bool_ = _DType('bool')

_counter = [0]


class TensorShape(tuple):
    def as_list(self):
        return [int(d) for d in self]


class Tensor(torch.Tensor):
    """torch tensor with the two TF tensor methods the reference uses (get_shape, name)."""

    def get_shape(self):
        return TensorShape(int(d) for d in self.shape)

    @property
    def name(self):
        n = self.__dict__.get('_tf_name')
        if n is None:
            _counter[0] += 1
            n = 'tensor_%d:0' % _counter[0]
            self.__dict__['_tf_name'] = n
        return n

    @property
    def op(self):
        return types.SimpleNamespace(name=self.name.split(':')[0])


def _wrap(t):
    return t if isinstance(t, Tensor) else t.as_subclass(Tensor)


def convert(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return _wrap(x)
    a = np.asarray(x)
    if a.dtype.kind == 'f':
        return _wrap(torch.as_tensor(a).to(_dt(dtype)))
    return _wrap(torch.as_tensor(a))


# ---------------------------------------------------------------------------- layout ops
def transpose(x, perm=None):
    x = convert(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return _wrap(x.permute(*[int(p) for p in perm]))


def reshape(x, shape):
    shape = [d if isinstance(d, (int, np.integer)) else _strict_int(d) for d in shape]
    return _wrap(convert(x).reshape(*[int(d) for d in shape]))


def _strict_int(d):
    # TF refuses non-integer shapes; the reference has one such call (gconv_lstm.py:288)
    if isinstance(d, float):
        raise TypeError('tf.reshape: shape must be integers, got %r' % (d,))
    return int(d)


def expand_dims(x, axis):
    return _wrap(convert(x).unsqueeze(int(axis)))


def squeeze(x, axis=None):
    x = convert(x)
    if axis is None:
        return _wrap(x.squeeze())
    for a in sorted([int(a) for a in axis], reverse=True):
        x = x.squeeze(a)
    return _wrap(x)


def concat(values, axis):
    return _wrap(torch.cat([convert(v) for v in values], dim=int(axis)))


def unstack(x, num=None, axis=0):
    x = convert(x)
    if num is not None:
        assert int(x.shape[axis]) == int(num), (tuple(x.shape), num, axis)
    return [_wrap(t) for t in torch.unbind(x, dim=int(axis))]


def split(value, num_or_size_splits, axis=0):
    value = convert(value)
    n = int(num_or_size_splits)
    return [_wrap(t) for t in torch.split(value, value.shape[axis] // n, dim=int(axis))]


def zeros(shape, dtype=None, name=None):
    return _wrap(torch.zeros([int(d) for d in shape], dtype=_dt(dtype)))


def constant(value, dtype=None):
    return convert(np.asarray(value), dtype)


def identity(x, name=None):
    return convert(x)


# ---------------------------------------------------------------------------- arithmetic
def matmul(a, b):
    return _wrap(torch.matmul(convert(a), convert(b)))


class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        self.indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
        self.values = np.asarray(values)
        self.dense_shape = tuple(int(d) for d in dense_shape)


def sparse_reorder(sp):
    order = np.lexsort((sp.indices[:, 1], sp.indices[:, 0]))      # canonical row-major ordering
    return SparseTensor(sp.indices[order], sp.values[order], sp.dense_shape)


def sparse_tensor_dense_matmul(sp, dense):
    dense = convert(dense)
    idx = torch.as_tensor(sp.indices.T.copy())
    vals = torch.as_tensor(sp.values).to(dense.dtype)
    A = torch.sparse_coo_tensor(idx, vals, sp.dense_shape)
    return _wrap(torch.sparse.mm(A, dense.as_subclass(torch.Tensor)))


def tan(x):
    return _wrap(torch.tan(convert(x)))


def tanh(x):
    return _wrap(torch.tanh(convert(x)))


def sigmoid(x):
    return _wrap(torch.sigmoid(convert(x)))


def square(x):
    return _wrap(convert(x) ** 2)


def subtract(a, b):
    return _wrap(convert(a) - convert(b))


def reduce_mean(x):
    return _wrap(convert(x).mean())


def py_func(func, inp, Tout):
    """Host round trip: numpy in, numpy out, no gradient (as tf.py_func)."""
    args = [convert(a).detach().cpu().numpy() for a in inp]
    args = [a.astype(np.float32) if a.dtype.kind == 'f' else a for a in args]     # TF tensors here are tf.float32
    out = func(*args)
    if not isinstance(out, (list, tuple)):
        out = [out]
    return [convert(np.asarray(o), float32) for o in out]


# ---------------------------------------------------------------------------- variables
class _Scope(object):
    def __init__(self, name, reuse=None):
        self.name = name          # full path
        self.reuse = reuse

    def reuse_variables(self):
        self.reuse = True


class _Graph(object):
    def __init__(self, seed=2017):
        self.variables = {}       # insertion ordered name -> Tensor (leaf, requires_grad)
        self.scopes = [_Scope('')]
        self.generator = torch.Generator().manual_seed(seed)
        self.dropout_generator = torch.Generator().manual_seed(seed + 1)
        self.dropout_masks = []   # in call order


_graph = [_Graph()]


def reset_default_graph(seed=2017):
    _graph[0] = _Graph(seed)
    _counter[0] = 0
    return _graph[0]


def default_graph():
    return _graph[0]


def get_variable_scope():
    return _graph[0].scopes[-1]


@contextlib.contextmanager
def variable_scope(name_or_scope, reuse=None):
    g = _graph[0]
    cur = g.scopes[-1]
    if isinstance(name_or_scope, _Scope):
        # re-entering a captured scope object: same path, shared reuse flag (TF semantics)
        scope = name_or_scope
        if reuse:
            scope.reuse = True
    else:
        path = (cur.name + '/' if cur.name else '') + str(name_or_scope)
        scope = _Scope(path, True if (reuse or cur.reuse) else None)
    g.scopes.append(scope)
    try:
        yield scope
    finally:
        g.scopes.pop()


@contextlib.contextmanager
def name_scope(name):
    yield name


def get_variable(name, shape=None, dtype=None, initializer=None):
    g = _graph[0]
    scope = g.scopes[-1]
    full = (scope.name + '/' if scope.name else '') + name
    if full in g.variables:
        if not scope.reuse:
            raise ValueError('Variable %s already exists, disallowed. Did you mean to set reuse=True?' % full)
        v = g.variables[full]
        if shape is not None and tuple(int(d) for d in shape) != tuple(v.shape):
            raise ValueError('shape mismatch for %s' % full)
        return v
    if scope.reuse:
        raise ValueError('Variable %s does not exist, or was not created with tf.get_variable().' % full)
    shape = tuple(int(d) for d in shape)
    if initializer is None:
        initializer = glorot_uniform_initializer()
    value = initializer(shape, g.generator).to(_dt(dtype))
    v = _wrap(value).requires_grad_(True)
    v.__dict__['_tf_name'] = full + ':0'
    g.variables[full] = v
    return v


def truncated_normal_initializer(mean=0.0, stddev=1.0):
    def init(shape, gen):
        t = torch.empty(shape, dtype=torch.float64)
        torch.nn.init.trunc_normal_(t, mean=mean, std=stddev, a=mean - 2 * stddev, b=mean + 2 * stddev, generator=gen)
        return t.to(torch.float32)      # variable VALUES are fp32-representable in every dtype mode
    return init


def random_uniform_initializer(minval=0.0, maxval=1.0):
    def init(shape, gen):
        return (torch.rand(shape, dtype=torch.float64, generator=gen) * (maxval - minval) + minval).to(torch.float32)
    return init


def constant_initializer(value):
    def init(shape, gen):
        return torch.full(shape, float(value), dtype=torch.float32)
    return init


def glorot_uniform_initializer():
    def init(shape, gen):
        fan_in = shape[0] if len(shape) > 1 else shape[0]
        fan_out = shape[-1]
        limit = (6.0 / (fan_in + fan_out)) ** 0.5
        return ((torch.rand(shape, dtype=torch.float64, generator=gen) * 2 - 1) * limit).to(torch.float32)
    return init


# ---------------------------------------------------------------------------- tf.nn
def _pool1(x, ksize, strides, padding, kind):
    x = convert(x)                       # N x M x F x 1, window over axis 1
    assert list(ksize) == list(strides) and ksize[0] == 1 and ksize[2] == 1 and ksize[3] == 1, (ksize, strides)
    p = int(ksize[1])
    N, M, F, one = x.shape
    assert M % p == 0, 'SAME padding never pads here: M is a multiple of p by construction (coarsening.py:208-212)'
    w = x.reshape(N, M // p, p, F, one)
    return _wrap(w.max(dim=2).values if kind == 'max' else w.mean(dim=2))


nn = types.ModuleType('tensorflow.nn')
nn.relu = lambda x: _wrap(torch.relu(convert(x)))
nn.tanh = tanh
nn.sigmoid = sigmoid
nn.softmax = lambda x: _wrap(torch.softmax(convert(x), dim=-1))
nn.l2_loss = lambda x: _wrap(0.5 * (convert(x) ** 2).sum())
nn.max_pool = lambda x, ksize, strides, padding: _pool1(x, ksize, strides, padding, 'max')
nn.avg_pool = lambda x, ksize, strides, padding: _pool1(x, ksize, strides, padding, 'avg')


def _dropout(x, keep_prob):
    g = _graph[0]
    x = convert(x)
    mask = (torch.rand(x.shape, generator=g.dropout_generator, dtype=torch.float64) < keep_prob).to(x.dtype) / keep_prob
    g.dropout_masks.append(mask)
    return _wrap(x * mask)


nn.dropout = _dropout


class RNNCell(object):
    def __init__(self, _reuse=None, **kwargs):
        self._reuse = _reuse


class DropoutWrapper(RNNCell):
    """output = dropout(cell output, output_keep_prob); state passes through."""

    def __init__(self, cell, input_keep_prob=1.0, output_keep_prob=1.0):
        self._cell, self._keep = cell, output_keep_prob

    @property
    def state_size(self):
        return self._cell.state_size

    @property
    def output_size(self):
        return self._cell.output_size

    def zero_state(self, batch_size, dtype):
        return self._cell.zero_state(batch_size, dtype)

    def __call__(self, inputs, state, scope=None):
        out, new_state = self._cell(inputs, state, scope)
        if self._keep < 1:
            out = _dropout(out, self._keep)
        return out, new_state


class MultiRNNCell(RNNCell):
    def __init__(self, cells, state_is_tuple=True):
        self._cells = cells

    def zero_state(self, batch_size, dtype):
        return tuple(c.zero_state(batch_size, dtype) for c in self._cells)

    def __call__(self, inputs, state, scope=None):
        cur, new_states = inputs, []
        with variable_scope(scope or 'multi_rnn_cell'):
            for i, cell in enumerate(self._cells):
                with variable_scope('cell_%d' % i):
                    cur, ns = cell(cur, state[i])
                    new_states.append(ns)
        return cur, tuple(new_states)


def static_rnn(cell, inputs, initial_state=None, dtype=None, scope=None):
    outputs = []
    with variable_scope(scope or 'rnn') as varscope:
        batch = int(inputs[0].shape[0])
        state = initial_state if initial_state is not None else cell.zero_state(batch, dtype)
        for time, inp in enumerate(inputs):
            if time > 0:
                varscope.reuse_variables()
            out, state = cell(inp, state)
            outputs.append(out)
    return outputs, state


nn.rnn_cell = types.ModuleType('tensorflow.nn.rnn_cell')
nn.rnn_cell.DropoutWrapper = DropoutWrapper
nn.rnn_cell.MultiRNNCell = MultiRNNCell
nn.rnn_cell.RNNCell = RNNCell
nn.static_rnn = static_rnn


def install():
    """Register this module as ``tensorflow`` (+ the one private path the reference imports)."""
    me = sys.modules[__name__]
    me.bool = bool_
    sys.modules['tensorflow'] = me
    python = types.ModuleType('tensorflow.python')
    ops = types.ModuleType('tensorflow.python.ops')
    impl = types.ModuleType('tensorflow.python.ops.rnn_cell_impl')
    impl.RNNCell = RNNCell
    python.ops = ops
    ops.rnn_cell_impl = impl
    me.python = python
    sys.modules['tensorflow.python'] = python
    sys.modules['tensorflow.python.ops'] = ops
    sys.modules['tensorflow.python.ops.rnn_cell_impl'] = impl
    sys.modules['tensorflow.nn'] = nn
    return me
