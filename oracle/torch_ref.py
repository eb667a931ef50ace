"""Op-for-op PyTorch-CPU mirror of the reference's TF graph for the Chebyshev filter and the gconv-LSTM cell
(SURVEY.md 8(d) "CPU baseline (B2)"): ``torch.sparse.mm`` for ``tf.sparse_tensor_dense_matmul``, ``cat`` for the
growing ``tf.concat``, ``permute`` / ``reshape`` for the restack, ``matmul``, and torch autograd for TF autodiff --
including the extra traffic of the TF graph (K-1 concats, stack transpose, eight separate filters per LSTM step).

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for multi-step LSTM gradients and the
CPU arm of ``bench.py --config c4|c5``.  PINNED by tests/test_oracle_tf_pinned.py against tests/golden/tf_*.npz
(outputs of the reference's own sources).  Follows lib/filter.py:45-95 and lib/gconv_lstm.py:185-215 /
lib/gconvRNN.py:189-213.
"""
import numpy as np
import scipy.sparse
import torch

from . import graph_ref


def sparse_operator(L, lmax=2, dtype=torch.float32):
    """graph.rescale_L + tocoo + tf.SparseTensor + tf.sparse_reorder (lib/filter.py:65-70) as a torch sparse tensor."""
    Lr = scipy.sparse.csr_matrix(graph_ref.rescale_L(scipy.sparse.csr_matrix(L, copy=True), lmax))
    Lr.sort_indices()
    Lr = Lr.tocoo()
    idx = torch.as_tensor(np.vstack([Lr.row, Lr.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.as_tensor(Lr.data).to(dtype), Lr.shape).coalesce().to_sparse_csr()


def cheby_conv(x, Ls, K, W):
    """lib/filter.py:72-95 with the rescaled sparse operator ``Ls`` given (hoisted: the reference rebuilds it per call
    at graph-construction time, not per step)."""
    nSample, nNode, feat_in = x.shape
    x0 = x.permute(1, 2, 0).reshape(nNode, feat_in * nSample)
    xs = x0.unsqueeze(0)
    if K > 1:
        x1 = torch.sparse.mm(Ls, x0)
        xs = torch.cat([xs, x1.unsqueeze(0)], dim=0)
    for _ in range(2, K):
        x2 = 2 * torch.sparse.mm(Ls, x1) - x0
        xs = torch.cat([xs, x2.unsqueeze(0)], dim=0)
        x0, x1 = x1, x2
    xs = xs.reshape(K, nNode, feat_in, nSample).permute(3, 1, 2, 0).reshape(nSample * nNode, feat_in * K)
    return torch.matmul(xs, W).reshape(nSample, nNode, W.shape[1])


def lstm_cell(x, c, h, Ls, K, Wx, Wh, b, variant='fork'):
    """One GConvLSTMCell step as the reference computes it: eight separate filters (lib/gconv_lstm.py:185-215).
    Wx / Wh / b: dicts keyed 'z', 'i', 'f', 'o'."""
    pre = {g: cheby_conv(x, Ls, K, Wx[g]) + cheby_conv(h, Ls, K, Wh[g]) + b[g] for g in 'zifo'}
    if variant == 'fork':
        z, o = torch.tan(pre['z']), torch.tanh(pre['o'])
    else:
        z, o = torch.tanh(pre['z']), torch.sigmoid(pre['o'])
    i, f = torch.sigmoid(pre['i']), torch.sigmoid(pre['f'])
    new_c = f * c + i * z
    return o * torch.tanh(new_c), new_c
