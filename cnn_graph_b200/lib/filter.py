"""Stand-alone graph filters with the call surface of the reference's ``lib/filter.py``.

``cheby_conv(x, L, lmax, feat_out, K, W=None)`` (reference lib/filter.py:45-95; copies at
lib/models.py:416-460 and lib/gconvRNN.py:27-71) keeps its positional order, its weight
layout ``[K*feat_in, feat_out]`` with row index ``fin*K + k``, and the creation of a
``weights`` variable in the ambient scope when ``W`` is None; ``fourier_conv`` (:11-42) is the
dense-eigenbasis filter the fork's recorded runs used.  The op chain behind cheby_conv
(SparseTensor staging, K-1 SpMM launches, growing concat, stack transpose, matmul) is
replaced by the native kernels in ``cnn_graph_b200.ops``.
"""
from .. import ops
from . import variables

__all__ = ['cheby_conv', 'fourier_conv']


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
    """Dense spectral filter (lib/filter.py:29-42): x [N, M, Fin] -> [N, M, Fout] with one Fout x Fin matrix per graph
    frequency, W [M, Fout, Fin] (created as ``weights`` in the ambient scope when None).  ``lmax`` and ``K`` are
    accepted and ignored, as in the reference (K is overwritten by M there, :31)."""
    N, M, Fin = (int(d) for d in x.shape)
    if W is None:
        W = variables.get_variable('weights', [M, Fout, Fin], variables.truncated_normal_initializer(0, 0.1))
    if tuple(W.shape) != (M, Fout, Fin):
        raise ValueError('fourier_conv: W must be [M, Fout, Fin] = [%d, %d, %d], got %r' % (M, Fout, Fin, tuple(W.shape)))
    return ops.fourier_filter(x, W, L)
