"""A minimal stand-in for TF-1 variable scopes, so that the reference's call surface
(``self.filter(x, L, Fout, K)`` / ``filter.cheby_conv(x, L, lmax, Fout, K)`` creating their
own ``weights`` in the ambient scope -- lib/graph_model.py:326-342, lib/filter.py:62-64) can
be kept on PyTorch.  Variables are ``torch.nn.Parameter`` objects stored by scoped name;
asking for an existing name returns the same parameter (TF ``reuse`` semantics).
"""
import contextlib
import math

import torch

_stack = []   # active VariableStore objects (innermost last)


class VariableStore:
    def __init__(self, device=None, seed=2017):
        self.vars = {}                      # insertion-ordered: name -> Parameter
        self.scope = []
        self.device = device
        self.generator = torch.Generator(device='cpu')
        self.generator.manual_seed(seed)    # tf.set_random_seed(2017), lib/graph_model.py:41

    def full_name(self, name):
        return '/'.join(self.scope + [name])

    def get(self, name, shape, initializer):
        key = self.full_name(name)
        if key in self.vars:
            p = self.vars[key]
            if tuple(p.shape) != tuple(shape):
                raise ValueError('variable %s exists with shape %r, asked for %r' % (key, tuple(p.shape), tuple(shape)))
            return p
        value = initializer(tuple(int(s) for s in shape), self.generator)
        p = torch.nn.Parameter(value.to(self.device) if self.device is not None else value)
        self.vars[key] = p
        return p

    def parameters(self):
        return list(self.vars.values())


def truncated_normal_initializer(mean=0.0, stddev=0.1):
    """tf.truncated_normal_initializer: normal re-drawn outside two standard deviations."""
    def init(shape, gen):
        t = torch.empty(shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, mean=mean, std=stddev, a=mean - 2 * stddev, b=mean + 2 * stddev, generator=gen)
        return t
    return init


def constant_initializer(value):
    def init(shape, gen):
        return torch.full(shape, float(value), dtype=torch.float32)
    return init


def random_uniform_initializer(minval=-0.1, maxval=0.1):
    def init(shape, gen):
        return torch.rand(shape, dtype=torch.float32, generator=gen) * (maxval - minval) + minval
    return init


def glorot_uniform_initializer():
    """tf.get_variable's default initializer (used for the LSTM biases, lib/gconv_lstm.py:177-180)."""
    def init(shape, gen):
        fan_in = shape[0] if len(shape) else 1
        fan_out = shape[-1] if len(shape) else 1
        if len(shape) == 1:
            fan_in = fan_out = shape[0]
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, dtype=torch.float32, generator=gen) * 2 - 1) * limit
    return init


def current_store():
    if not _stack:
        raise RuntimeError('no active variable store: call inside a model (or `with variables.use_store(store):`)')
    return _stack[-1]


@contextlib.contextmanager
def use_store(store):
    _stack.append(store)
    try:
        yield store
    finally:
        _stack.pop()


@contextlib.contextmanager
def variable_scope(name):
    store = current_store()
    store.scope.append(name)
    try:
        yield
    finally:
        store.scope.pop()


def get_variable(name, shape, initializer=None):
    if initializer is None:
        initializer = glorot_uniform_initializer()
    return current_store().get(name, shape, initializer)
