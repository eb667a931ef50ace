"""The TF half of the oracle (oracle/tf_ref.py) has no reference-run golden (TensorFlow is not
installable; SURVEY.md 8(c): parity unpinned at the TF boundary).  These CPU tests tie it to
what IS available: the reference's own numpy basis (pinned in test_oracle_golden.py), the
float64 identities of trials/1_learning_filters.ipynb cells 39/41/43, and float64 torch
autograd of an independent dense formulation for every backward formula."""
import numpy as np
import pytest
import scipy.sparse
import torch

from conftest import csr_from
from oracle import graph_ref, tf_ref


def dense_filter_torch(x, Lr_dense, W, K):
    """Independent float64 formulation: y = sum_k T_k(L~) x W_k with explicit dense T_k."""
    N, M, Fin = x.shape
    Wk = W.reshape(Fin, K, -1)                         # row fin*K + k
    T0 = torch.eye(M, dtype=x.dtype)
    y = torch.einsum('ij,njf,fo->nio', T0, x, Wk[:, 0])
    if K > 1:
        T1 = Lr_dense
        y = y + torch.einsum('ij,njf,fo->nio', T1, x, Wk[:, 1])
    for k in range(2, K):
        T2 = 2 * Lr_dense @ T1 - T0
        y = y + torch.einsum('ij,njf,fo->nio', T2, x, Wk[:, k])
        T0, T1 = T1, T2
    return y


def rescaled(L, lmax=2):
    return graph_ref.rescale_L(scipy.sparse.csr_matrix(L, copy=True), lmax)


def test_trial_identities_float64(c2):
    # cells 39 / 43: basis[0] == X and recurrence == U T_k(lambda~) U^T X
    L = csr_from(c2, 'L3').astype(np.float64)
    Lr = rescaled(L)
    Lr = scipy.sparse.csr_matrix((Lr + Lr.T) / 2)       # fp32 D.W.D is symmetric only to ~1e-8
    lam, U = np.linalg.eigh(Lr.toarray())
    rng = np.random.RandomState(0)
    X = rng.standard_normal((L.shape[0], 4))
    K = 9
    Xt = graph_ref.chebyshev(Lr, X, K)
    assert np.array_equal(Xt[0], X)
    Tk = np.empty((K, lam.size))                        # T_k(lambda) by the scalar recurrence
    Tk[0], Tk[1] = 1.0, lam
    for k in range(2, K):
        Tk[k] = 2 * lam * Tk[k - 1] - Tk[k - 2]
    for k in range(K):
        ref = U @ (Tk[k][:, None] * (U.T @ X))
        assert np.abs(Xt[k] - ref).max() < 1e-10
    # cell 41: Clenshaw evaluation of sum_k c_k T_k(L~) X equals the explicit basis combination
    c = rng.standard_normal(K)
    direct = np.tensordot(c, Xt, axes=(0, 0))
    b1 = b2 = np.zeros_like(X)
    for k in range(K - 1, 0, -1):
        b1, b2 = c[k] * X + 2 * Lr.dot(b1) - b2, b1
    clenshaw = c[0] * X + Lr.dot(b1) - b2
    assert np.abs(direct - clenshaw).max() < 1e-10


@pytest.mark.parametrize('N,Fin,Fout,K', [(3, 1, 4, 1), (2, 1, 5, 2), (4, 3, 6, 5), (2, 5, 2, 7)])
def test_chebyshev5_forward_backward_vs_autograd(c2, N, Fin, Fout, K):
    L = csr_from(c2, 'L3')          # M = 124
    M = L.shape[0]
    rng = np.random.RandomState(K)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    W = (0.1 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((N, M, Fout)).astype(np.float32)
    y = tf_ref.chebyshev5(x, L, W, K)
    assert y.shape == (N, M, Fout) and y.dtype == np.float32
    # layout identity: chebyshev5 == chebyshev2 == einsum over the reference's own basis
    assert np.allclose(y, tf_ref.chebyshev2(x, L, W, K), rtol=1e-6, atol=1e-6)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    Wt = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    Lr = torch.tensor(rescaled(L.astype(np.float64)).toarray())
    yt = dense_filter_torch(xt, Lr, Wt, K)
    assert np.abs(yt.detach().numpy() - y).max() < 1e-4 * max(1.0, np.abs(y).max())
    yt.backward(torch.tensor(gy, dtype=torch.float64))
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy)
    assert np.abs(dx - xt.grad.numpy()).max() < 1e-4 * np.abs(xt.grad.numpy()).max()
    assert np.abs(dW - Wt.grad.numpy()).max() < 1e-4 * np.abs(Wt.grad.numpy()).max()


def test_directed_operator_uses_true_transpose(directed):
    L = csr_from(directed, 'L')
    M = L.shape[0]
    rng = np.random.RandomState(1)
    x = rng.standard_normal((2, M, 3)).astype(np.float32)
    W = (0.1 * rng.standard_normal((3 * 4, 5))).astype(np.float32)
    gy = rng.standard_normal((2, M, 5)).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    Wt = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    Lr = torch.tensor(rescaled(L.astype(np.float64), 3.5).toarray())
    dense_filter_torch(xt, Lr, Wt, 4).backward(torch.tensor(gy, dtype=torch.float64))
    dx, dW = tf_ref.chebyshev5_backward(x, L, W, 4, gy, lmax=3.5)
    assert np.abs(dx - xt.grad.numpy()).max() < 1e-4 * np.abs(xt.grad.numpy()).max()
    assert np.abs(dW - Wt.grad.numpy()).max() < 1e-4 * np.abs(Wt.grad.numpy()).max()


def test_pool_and_activation_vs_torch():
    rng = np.random.RandomState(2)
    x = rng.standard_normal((3, 16, 5)).astype(np.float32)
    x[0, :4, 0] = 1.5                                   # a tie: first index must win
    g = rng.standard_normal((3, 4, 5)).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True)
    mp = torch.nn.functional.max_pool1d(xt.permute(0, 2, 1), 4).permute(0, 2, 1)
    assert np.array_equal(tf_ref.mpool1(x, 4), mp.detach().numpy())
    assert tf_ref.mpool1_argmax(x, 4)[0, 0, 0] == 0
    gm = tf_ref.mpool1_backward(x, 4, g)
    assert gm[0, 0, 0] == g[0, 0, 0] and np.all(gm[0, 1:4, 0] == 0)
    ap = torch.nn.functional.avg_pool1d(xt.permute(0, 2, 1), 4).permute(0, 2, 1)
    assert np.allclose(tf_ref.apool1(x, 4), ap.detach().numpy(), atol=1e-6)
    ap.backward(torch.tensor(g))
    assert np.allclose(tf_ref.apool1_backward(x, 4, g), xt.grad.numpy(), atol=1e-7)
    assert tf_ref.mpool1(x, 1) is x and tf_ref.apool1(x, 1) is x
    b = rng.standard_normal(5).astype(np.float32)
    assert np.array_equal(tf_ref.b1relu(x), np.maximum(x, 0))
    assert np.allclose(tf_ref.b1relu(x, b), torch.relu(torch.tensor(x) + torch.tensor(b)).numpy())
    b2 = rng.standard_normal((16, 5)).astype(np.float32)
    assert np.allclose(tf_ref.b2relu(x, b2), torch.relu(torch.tensor(x) + torch.tensor(b2)).numpy())


@pytest.mark.parametrize('variant', ['fork', 'standard'])
def test_lstm_step_matches_eight_separate_filters(c2, variant):
    L = csr_from(c2, 'L4')          # M = 62
    M, N, Fin, H, K = L.shape[0], 2, 2, 3, 3
    rng = np.random.RandomState(3)
    x = rng.standard_normal((N, M, Fin)).astype(np.float32)
    h = (0.5 * rng.standard_normal((N, M, H))).astype(np.float32)
    c = (0.5 * rng.standard_normal((N, M, H))).astype(np.float32)
    Wx = {g: rng.uniform(-0.1, 0.1, (K * Fin, H)).astype(np.float32) for g in 'zifo'}
    Wh = {g: rng.uniform(-0.1, 0.1, (K * H, H)).astype(np.float32) for g in 'zifo'}
    b = {g: rng.uniform(-0.1, 0.1, H).astype(np.float32) for g in 'zifo'}
    new_h, new_c = tf_ref.gconv_lstm_step(x, c, h, L, 2, K, Wx, Wh, b, variant)
    # fused formulation used by the product: one filter on [x|h] with stacked weights
    Wcat = np.concatenate([np.concatenate([Wx[g] for g in 'zifo'], 1), np.concatenate([Wh[g] for g in 'zifo'], 1)], 0)
    pre = tf_ref.chebyshev5(np.concatenate([x, h], 2), L, Wcat, K) + np.concatenate([b[g] for g in 'zifo'])
    z, i, f, o = (pre[..., j * H:(j + 1) * H] for j in range(4))
    sig = lambda v: 1 / (1 + np.exp(-v))
    zz = np.tan(z) if variant == 'fork' else np.tanh(z)
    oo = np.tanh(o) if variant == 'fork' else sig(o)
    cc = sig(f) * c + sig(i) * zz
    assert np.allclose(new_c, cc, rtol=1e-5, atol=1e-6)
    assert np.allclose(new_h, oo * np.tanh(cc), rtol=1e-5, atol=1e-6)
