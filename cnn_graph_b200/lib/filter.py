"""Stand-alone graph filters with the call surface of the reference's ``lib/filter.py``.

``cheby_conv(x, L, lmax, feat_out, K, W=None)`` (reference lib/filter.py:45-95; copies at
lib/models.py:416-460 and lib/gconvRNN.py:27-71) keeps its positional order, its weight
layout ``[K*feat_in, feat_out]`` with row index ``fin*K + k``, and the creation of a
``weights`` variable in the ambient scope when ``W`` is None.  The op chain behind it
(SparseTensor staging, K-1 SpMM launches, growing concat, stack transpose, matmul) is
replaced by the native kernels in ``cnn_graph_b200.ops``.
"""
from .. import ops
from . import variables

__all__ = ['cheby_conv']


def cheby_conv(x, L, lmax, feat_out, K, W=None):
    """x [nSample, nNode, feat_in] -> [nSample, nNode, feat_out]; no bias (lib/filter.py:93)."""
    nSample, nNode, feat_in = (int(d) for d in x.shape)
    if W is None:
        W = variables.get_variable('weights', [K * feat_in, feat_out], variables.truncated_normal_initializer(0, 0.1))
    if tuple(W.shape) != (K * feat_in, feat_out):
        raise ValueError('cheby_conv: W must be [K*feat_in, feat_out] = [%d, %d], got %r'
                         % (K * feat_in, feat_out, tuple(W.shape)))
    return ops.cheb_filter(x, W, L, K, lmax=lmax)


def fourier_conv(x, L, lmax, Fout, K, W=None):
    """The dense-EVD spectral filter of the reference (lib/filter.py:29-42) is outside the
    Chebyshev hot path and is not provided."""
    raise NotImplementedError('fourier_conv is out of scope of the B200 Chebyshev hot path')
