"""Oracle restatement of the TensorFlow half of the hot path, in numpy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: TensorFlow 1.x
(``requirements.txt:7``, no version pin) cannot be installed here and the
reference holds no golden vector for these ops, so the reference's own source
files (``lib/filter.py``, ``lib/models.py``, ``lib/gconv_lstm.py``,
``lib/gconvRNN.py``) were EXECUTED UNMODIFIED in the build container under a
torch-backed ``tensorflow`` stand-in (``tests/golden/tf_shim.py``; generator
``tests/golden/make_golden_tf.py``), outputs and autograd gradients stored in
``tests/golden/tf_*.npz``; ``tests/test_oracle_tf_pinned.py`` checks every
function below against them.  What remains outside the pin is TensorFlow's own
kernel arithmetic (summation order inside its SpMM / GEMM), which the 1e-4
tolerance covers.  The functions follow the reference's op sequence literally
(same layouts, same transposes, same order of the recurrence); the arithmetic
of each TF op is restated from its documented meaning:

* ``tf.sparse_tensor_dense_matmul(L, x)`` after ``tf.sparse_reorder``  ->
  row-major CSR x dense, float32 (scipy ``csr_matvecs``; the same kernel the
  reference's numpy twin ``graph.chebyshev`` uses).
* ``tf.matmul``  -> float32 GEMM (numpy BLAS).
* ``tf.nn.max_pool`` / ``avg_pool`` with ksize = strides = [1, p, 1, 1] -> max /
  mean over p consecutive vertices.

Backward passes have no reference code (TF autodiff); they are derived in
SURVEY.md appendix A.3 and are checked in ``tests/test_oracle.py`` against
float64 torch autograd of the restated forward.
"""
import numpy as np
import scipy.sparse

from . import graph_ref


# --------------------------------------------------------------------------
# Chebyshev filter
# --------------------------------------------------------------------------

def _rescaled_csr(L, lmax=2):
    # lib/models.py:196-197 (copy, then graph.rescale_L) / lib/filter.py:65
    L = scipy.sparse.csr_matrix(L, copy=True)
    L = graph_ref.rescale_L(L, lmax)
    L = scipy.sparse.csr_matrix(L)
    L.sort_indices()                      # tf.sparse_reorder, lib/models.py:201
    return L


def cheb_basis_tf(x, L, K, lmax=2):
    """x [N, M, Fin] -> stack [K, M, Fin*N], column c = fin*N + n.

    lib/models.py:203-217 == lib/filter.py:72-87 (the concat only grows the
    stack, so it is written in place here).
    """
    N, M, Fin = x.shape
    Lr = _rescaled_csr(L, lmax)
    x0 = np.ascontiguousarray(np.transpose(x, (1, 2, 0))).reshape(M, Fin * N)
    stack = np.empty((K, M, Fin * N), np.float32)
    stack[0] = x0
    if K > 1:
        x1 = Lr.dot(x0)
        stack[1] = x1
    for k in range(2, K):
        x2 = 2 * Lr.dot(x1) - x0
        stack[k] = x2
        x0, x1 = x1, x2
    return stack


def chebyshev5(x, L, W, K, lmax=2, return_stack=False):
    """y [N, M, Fout] = chebyshev5(x [N, M, Fin]); W [Fin*K, Fout], row = fin*K + k.

    lib/models.py:192-224, lib/graph_conv.py:144-176, lib/filter.py:45-95.  ``return_stack``: also return the
    restacked [N*M, Fin*K] matmul operand (lib/models.py:218-220), the tensor TF autodiff keeps for the backward.
    """
    N, M, Fin = x.shape
    Fout = W.shape[1]
    assert W.shape[0] == Fin * K
    stack = cheb_basis_tf(x.astype(np.float32, copy=False), L, K, lmax)
    a = stack.reshape(K, M, Fin, N)
    a = np.ascontiguousarray(np.transpose(a, (3, 1, 2, 0))).reshape(N * M, Fin * K)
    y = np.matmul(a, W.astype(np.float32, copy=False)).reshape(N, M, Fout)
    return (y, a) if return_stack else y


def cheby_conv(x, L, lmax, feat_out, K, W):
    """lib/filter.py:45-95 with the weight supplied by the caller."""
    assert W.shape[1] == feat_out
    return chebyshev5(x, L, W, K, lmax)


def fourier_conv(x, L, W, U=None):
    """Dense spectral filter, lib/filter.py:11-42 == lib/models.py:129-157.

    x [N, M, Fin], W [M, Fout, Fin] (one Fout x Fin matrix per graph frequency) ->
    y [N, M, Fout] = U^T-side transform as the reference writes it: with Ut = U.T (U = eigenvectors of L,
    ``graph.fourier``), xh = Ut x, yh[m] = W[m] xh[m], y = (yh^T Ut) re-laid out as N x M x Fout.
    """
    N, M, Fin = x.shape
    if U is None:
        U = graph_ref.fourier(L)[1]
    Ut = np.asarray(U.T, np.float32)                       # tf.constant(U.T, dtype=tf.float32)
    xh = np.transpose(x, (1, 2, 0)).reshape(M, Fin * N)     # M x Fin*N
    xh = np.matmul(Ut, xh).reshape(M, Fin, N)
    yh = np.matmul(W, xh)                                   # M x Fout x N (batched over the frequency index)
    yh = np.transpose(yh).reshape(N * W.shape[1], M)        # tf.transpose without perm reverses: N x Fout x M
    y = np.matmul(yh, Ut).reshape(N, W.shape[1], M)
    return np.ascontiguousarray(np.transpose(y, (0, 2, 1)))


def chebyshev2(x, L, W, K):
    """Same forward as chebyshev5, basis from graph.chebyshev.  lib/models.py:161-190."""
    N, M, Fin = x.shape
    Lr = scipy.sparse.csr_matrix(L, copy=True)
    Lr = graph_ref.rescale_L(Lr, lmax=2)
    x0 = np.ascontiguousarray(np.transpose(x, (1, 2, 0))).reshape(M, Fin * N)
    stack = graph_ref.chebyshev(scipy.sparse.csr_matrix(Lr), x0, K)
    a = np.transpose(stack.reshape(K, M, Fin, N), (3, 1, 2, 0)).reshape(N * M, Fin * K)
    return np.matmul(a, W).reshape(N, M, W.shape[1])


def chebyshev5_backward(x, L, W, K, gy, lmax=2, a=None, need_dx=True):
    """(dx, dW) for y = chebyshev5(x).  SURVEY.md appendix A.3.  ``a``: the restacked operand kept by the forward
    (``chebyshev5(..., return_stack=True)``) -- what TF autodiff does; None recomputes the basis.

    dW[fin*K+k, fo] = sum_{n,m} Xk[m, fin*N+n] gy[n,m,fo]
    Gk = gy W_k^T;  for k = K-1..2: G_{k-1} += 2 L~^T G_k, G_{k-2} -= G_k;
    G_0 += L~^T G_1;  dx = G_0.
    """
    N, M, Fin = x.shape
    Fout = W.shape[1]
    if a is None:
        stack = cheb_basis_tf(x.astype(np.float32, copy=False), L, K, lmax)
        a = np.transpose(stack.reshape(K, M, Fin, N), (3, 1, 2, 0)).reshape(N * M, Fin * K)
    g2 = gy.reshape(N * M, Fout).astype(np.float32, copy=False)
    dW = np.matmul(a.T, g2)
    if not need_dx:
        return None, dW
    ga = np.matmul(g2, W.T)                               # [N*M, Fin*K]
    G = np.transpose(ga.reshape(N, M, Fin, K), (3, 1, 2, 0)).reshape(K, M, Fin * N).copy()
    Lt = _rescaled_csr(L, lmax).T.tocsr()
    for k in range(K - 1, 1, -1):
        G[k - 1] += 2 * Lt.dot(G[k])
        G[k - 2] -= G[k]
    if K > 1:
        G[0] += Lt.dot(G[1])
    dx = np.transpose(G[0].reshape(M, Fin, N), (2, 0, 1))
    return np.ascontiguousarray(dx), dW


# --------------------------------------------------------------------------
# bias / activation / pooling
# --------------------------------------------------------------------------

def b1relu(x, b=None):
    """Fork: relu(x) (lib/models.py:226-235); upstream: relu(x + b[1,1,F])."""
    return np.maximum(x if b is None else x + b.reshape(1, 1, -1), 0)


def b1tanh(x, b):
    """tanh(x + b[1,1,F]).  lib/models.py:237-241."""
    return np.tanh(x + b.reshape(1, 1, -1))


def b2relu(x, b):
    """relu(x + b[1,M,F]).  lib/models.py:243-247."""
    return np.maximum(x + b.reshape(1, x.shape[1], x.shape[2]), 0)


def mpool1(x, p):
    """Max over p consecutive vertices.  lib/models.py:249-257."""
    if p <= 1:
        return x
    N, M, F = x.shape
    return x.reshape(N, M // p, p, F).max(axis=2)


def mpool1_argmax(x, p):
    """Index (0..p-1) of the first maximal element per window (TF max-pool grad routing)."""
    N, M, F = x.shape
    return x.reshape(N, M // p, p, F).argmax(axis=2)


def apool1(x, p):
    """Mean over p consecutive vertices, fake zeros included.  lib/models.py:259-266."""
    if p <= 1:
        return x
    N, M, F = x.shape
    return x.reshape(N, M // p, p, F).mean(axis=2, dtype=np.float32)


def mpool1_backward(x, p, g):
    if p <= 1:
        return g
    N, M, F = x.shape
    am = mpool1_argmax(x, p)
    out = np.zeros((N, M // p, p, F), g.dtype)
    n, j, f = np.meshgrid(np.arange(N), np.arange(M // p), np.arange(F), indexing='ij')
    out[n, j, am, f] = g
    return out.reshape(N, M, F)


def apool1_backward(x, p, g):
    if p <= 1:
        return g
    return np.repeat(g / np.float32(p), p, axis=1)


def fc(x, W, b, relu=True):
    """lib/models.py:268-274."""
    y = np.matmul(x, W) + b
    return np.maximum(y, 0) if relu else y


# --------------------------------------------------------------------------
# graph-conv LSTM step
# --------------------------------------------------------------------------

def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def gconv_lstm_step(x, c, h, L, lmax, K, Wx, Wh, b, variant='fork'):
    """One GConvLSTMCell step.  lib/gconv_lstm.py:185-215 ('fork': z = tan(.),
    o = tanh(.)) and lib/gconvRNN.py:189-213 ('standard': z = tanh, o = sigmoid).

    Wx / Wh / b: dicts keyed 'z','i','f','o' of [K*Fin, H] / [K*H, H] / [H].
    Returns (new_h, new_c).
    """
    H = b['z'].shape[0]
    pre = {}
    for g in 'zifo':
        pre[g] = (cheby_conv(x, L, lmax, H, K, Wx[g]) + cheby_conv(h, L, lmax, H, K, Wh[g])
                  + b[g].reshape(1, 1, H))
    if variant == 'fork':
        z, o = np.tan(pre['z']), np.tanh(pre['o'])
    else:
        z, o = np.tanh(pre['z']), _sigmoid(pre['o'])
    i, f = _sigmoid(pre['i']), _sigmoid(pre['f'])
    new_c = f * c + i * z
    new_h = o * np.tanh(new_c)
    return new_h.astype(np.float32), new_c.astype(np.float32)


def mse_loss(pred, labels):
    """lib/graph_model.py:255."""
    return np.mean(np.square(labels - pred))
