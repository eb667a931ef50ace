"""Oracle restatement of one ``cgcnn`` training step on the CPU (numpy / scipy), used as the
checker for model-level parity tests and as the CPU baseline of ``bench.py``.

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  It chains the restated
TF ops of ``oracle/tf_ref.py`` in the order of the upstream ``cgcnn._inference``
(lib/models.py:25-42, 79-111, 268-274): per layer filter -> b1relu -> pool,
flatten, fc stack, softmax cross-entropy + L2, momentum-SGD update.  The Chebyshev basis is
computed the way the reference's own numpy path does (scipy CSR x dense, single thread);
the dense products use numpy BLAS with all host threads.
"""
import numpy as np

from . import tf_ref


def init_params(L, F, K, p, M, seed=0, bias=True):
    rng = np.random.RandomState(seed)
    params = {}
    Fin = 1
    for i, (Fo, Kk) in enumerate(zip(F, K)):
        params['conv%d/weights' % (i + 1)] = np.clip(0.1 * rng.standard_normal((Fin * Kk, Fo)), -0.2, 0.2).astype(np.float32)
        if bias:
            params['conv%d/bias' % (i + 1)] = np.full((1, 1, Fo), 0.1, np.float32)
        Fin = Fo
    width = L[len(F) - 1].shape[0] * F[-1] // p[-1] if len(F) else L[0].shape[0]
    names = ['fc%d' % (i + 1) for i in range(len(M) - 1)] + ['logits']
    for name, out in zip(names, M):
        params[name + '/weights'] = np.clip(0.1 * rng.standard_normal((width, out)), -0.2, 0.2).astype(np.float32)
        params[name + '/bias'] = np.full((out,), 0.1, np.float32)
        width = out
    return params


def forward_backward(params, L, F, K, p, M, x, labels, regularization=0.0, pool='mpool1', dropout_masks=None,
                     keep_stack=True):
    """Returns (loss, grads).  x [N, M0] float32, labels [N] int.  ``keep_stack``: the forward keeps each layer's
    restacked Chebyshev operand for the backward (what TF autodiff does, lib/models.py:207-220); False recomputes it."""
    N = x.shape[0]
    nconv = len(F)
    acts = []
    h = x[:, :, None].astype(np.float32)
    for i in range(nconv):
        W = params['conv%d/weights' % (i + 1)]
        b = params.get('conv%d/bias' % (i + 1))
        a, kept = tf_ref.chebyshev5(h, L[i], W, K[i], return_stack=True)
        r = tf_ref.b1relu(a, b)
        q = tf_ref.mpool1(r, p[i]) if pool == 'mpool1' else tf_ref.apool1(r, p[i])
        acts.append((h, r, kept if keep_stack else None))
        h = q
    flat = h.reshape(N, -1)
    names = ['fc%d' % (i + 1) for i in range(len(M) - 1)] + ['logits']
    fcs = []
    z = flat
    for j, name in enumerate(names):
        last = j == len(names) - 1
        out = tf_ref.fc(z, params[name + '/weights'], params[name + '/bias'], relu=not last)
        mask = None
        fcs.append((z, out, dropout_masks[j] if (not last and dropout_masks is not None) else None))
        if not last and dropout_masks is not None:
            out = out * dropout_masks[j]      # masks already carry the 1/keep scaling
        z = out
    logits = z
    shifted = logits - logits.max(axis=1, keepdims=True)
    logp = shifted - np.log(np.exp(shifted).sum(axis=1, keepdims=True))
    loss = float(-logp[np.arange(N), labels].mean())
    reg_names = [n + s for n in names for s in ('/weights', '/bias')]
    if regularization:
        loss += regularization * sum(0.5 * float((params[n] ** 2).sum()) for n in reg_names)

    grads = {}
    g = np.exp(logp)
    g[np.arange(N), labels] -= 1.0
    g = (g / N).astype(np.float32)
    for j in range(len(names) - 1, -1, -1):
        name = names[j]
        zin, out, mask = fcs[j]
        if j != len(names) - 1:          # `out` is the relu output before the dropout mask
            if mask is not None:
                g = g * mask
            g = g * (out > 0)
        grads[name + '/weights'] = zin.T @ g
        grads[name + '/bias'] = g.sum(axis=0)
        g = g @ params[name + '/weights'].T
    if regularization:
        for n in reg_names:
            grads[n] = grads[n] + regularization * params[n]
    g = g.reshape(h.shape)
    for i in range(nconv - 1, -1, -1):
        hin, r, kept = acts[i]
        gr = tf_ref.mpool1_backward(r, p[i], g) if pool == 'mpool1' else tf_ref.apool1_backward(r, p[i], g)
        ga = (gr * (r > 0)).astype(np.float32)
        if 'conv%d/bias' % (i + 1) in params:
            grads['conv%d/bias' % (i + 1)] = ga.sum(axis=(0, 1)).reshape(1, 1, -1)
        W = params['conv%d/weights' % (i + 1)]
        g, dW = tf_ref.chebyshev5_backward(hin, L[i], W, K[i], ga, a=kept, need_dx=i > 0)
        grads['conv%d/weights' % (i + 1)] = dW
    return loss, grads


def sgd_momentum_step(params, grads, velocity, lr, momentum):
    for n, gval in grads.items():
        v = velocity.setdefault(n, np.zeros_like(params[n]))
        v *= momentum
        v += gval
        params[n] -= lr * v
