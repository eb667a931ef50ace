"""oracle/model_ref.py (the CPU baseline / model-level checker) against float64 torch autograd of
an independent dense formulation of the same cgcnn step.  CPU only, small shapes."""
import numpy as np
import scipy.sparse
import torch

from conftest import csr_from
from oracle import graph_ref, model_ref


def dense_T(L, K):
    Lr = graph_ref.rescale_L(scipy.sparse.csr_matrix(L.astype(np.float64), copy=True), 2).toarray()
    Ts = [np.eye(L.shape[0]), Lr]
    for k in range(2, K):
        Ts.append(2 * Lr @ Ts[-1] - Ts[-2])
    return torch.tensor(np.stack(Ts[:K]))


def test_model_ref_gradients_vs_autograd(c2):
    L = [csr_from(c2, 'L3'), csr_from(c2, 'L4')]          # 124 -> pool 2 -> 62 -> pool 2 -> 31
    F, K, p, M = [3, 4], [4, 3], [2, 2], [6, 5]
    N = 5
    rng = np.random.RandomState(0)
    x = rng.uniform(0, 1, (N, 124)).astype(np.float32)
    labels = rng.randint(0, 5, N)
    params = model_ref.init_params(L, F, K, p, M, seed=1)
    masks = [(rng.uniform(size=(N, 6)) < 0.5).astype(np.float32) / 0.5]
    loss, grads = model_ref.forward_backward(params, L, F, K, p, M, x, labels, 5e-4, 'mpool1', masks)

    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    h = torch.tensor(x, dtype=torch.float64)[:, :, None]
    for i in range(2):
        T = dense_T(L[i], K[i])
        Fin = h.shape[2]
        W = tp['conv%d/weights' % (i + 1)].reshape(Fin, K[i], F[i])
        a = torch.einsum('kij,njf,fko->nio', T, h, W)
        r = torch.relu(a + tp['conv%d/bias' % (i + 1)])
        h = torch.nn.functional.max_pool1d(r.permute(0, 2, 1), p[i]).permute(0, 2, 1)
    z = h.reshape(N, -1)
    z = torch.relu(z @ tp['fc1/weights'] + tp['fc1/bias']) * torch.tensor(masks[0], dtype=torch.float64)
    logits = z @ tp['logits/weights'] + tp['logits/bias']
    ref = torch.nn.functional.cross_entropy(logits, torch.tensor(labels))
    ref = ref + 5e-4 * sum(0.5 * (tp[n] ** 2).sum() for n in ('fc1/weights', 'fc1/bias', 'logits/weights', 'logits/bias'))
    ref.backward()
    assert abs(loss - float(ref)) < 1e-5
    for name, g in grads.items():
        want = tp[name].grad.numpy()
        assert np.abs(g - want).max() <= 1e-4 * max(np.abs(want).max(), 1e-12), name
    vel = {}
    before = params['fc1/weights'].copy()
    model_ref.sgd_momentum_step(params, grads, vel, 0.02, 0.9)
    assert np.allclose(params['fc1/weights'], before - 0.02 * grads['fc1/weights'])
