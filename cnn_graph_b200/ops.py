"""PyTorch-facing operators of the hot path: thin ``autograd.Function`` wrappers that hand
raw device pointers and the current CUDA stream to the C ABI (``include/cnn_graph_b200.h``).
PyTorch is plumbing only (device memory, streams, autograd tape); every kernel is native.
There is no CPU path: tensors must live on a CUDA device.
"""
import ctypes
import weakref

import numpy as np
import scipy.sparse
import torch

from . import _native
from ._native import check, ptr

FILTER_DEFAULT = 0
FILTER_FORCE_STREAMING = 1
FILTER_FORCE_ONCHIP = 2
FILTER_NO_FUSED = 4
FILTER_FORCE_FUSED = 8
FILTER_NO_CLENSHAW = 16


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _native.NativeError('cnn_graph_b200 ops need CUDA tensors (no CPU fallback); got device %s' % t.device)


def _f32c(t):
    """float32, contiguous and 16-byte aligned (bulk copies and 128-bit accesses of the native kernels)."""
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


# ---------------------------------------------------------------------------------------
# rescaled Laplacian handle
# ---------------------------------------------------------------------------------------

def rescale_csr(L, lmax=2):
    """L~ = L / (lmax/2) - I as a fresh sorted float32 CSR (lib/graph.py:232-238).

    Unlike the reference (lib/models.py:196 aliases the caller's data, lib/filter.py:65 does
    not copy at all) the caller's matrix is never modified; identical for lmax = 2.
    """
    L = scipy.sparse.csr_matrix(L, dtype=np.float32, copy=True)
    M = L.shape[0]
    L /= lmax / 2
    L -= scipy.sparse.identity(M, format='csr', dtype=L.dtype)
    L = scipy.sparse.csr_matrix(L)
    L.sum_duplicates()
    L.sort_indices()
    return L


class GraphHandle:
    """Packed L~ / L~^T resident on the current CUDA device (``cg_graph_t``)."""

    def __init__(self, L_rescaled):
        L = scipy.sparse.csr_matrix(L_rescaled, dtype=np.float32)
        L.sum_duplicates()
        L.sort_indices()
        if L.shape[0] != L.shape[1]:
            raise ValueError('Laplacian must be square, got %r' % (L.shape,))
        if not torch.cuda.is_available():
            raise _native.NativeError('cnn_graph_b200 needs a CUDA device (no CPU fallback)')
        self.M = int(L.shape[0])
        self.nnz = int(L.nnz)
        indptr = np.ascontiguousarray(L.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(L.indices, dtype=np.int32)
        data = np.ascontiguousarray(L.data, dtype=np.float32)
        self.device = torch.cuda.current_device()
        h = ctypes.c_void_p()
        check(_native.lib().cg_graph_create(ctypes.byref(h), self.M, self.nnz, indptr.ctypes.data,
                                            indices.ctypes.data, data.ctypes.data), 'cg_graph_create')
        self._h = h

    @classmethod
    def from_laplacian(cls, L, lmax=2):
        return cls(rescale_csr(L, lmax))

    @property
    def handle(self):
        if self._h is None:
            raise _native.NativeError('graph handle already destroyed')
        return self._h

    def info(self):
        out = (ctypes.c_int64 * 5)()
        check(_native.lib().cg_graph_info(self.handle, out), 'cg_graph_info')
        return {'M': out[0], 'nnz': out[1], 'width': out[2], 'width_t': out[3], 'onchip': bool(out[4])}

    def close(self):
        if getattr(self, '_h', None) is not None:
            _native.lib().cg_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handle_cache = {}


def get_handle(L, lmax=2):
    """Handle for a scipy Laplacian, cached per (object, lmax, device) -- the reference
    re-stages L on every filter call site (lib/models.py:196-201)."""
    if isinstance(L, GraphHandle):
        return L
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    key = (id(L), float(lmax), dev)
    hit = _handle_cache.get(key)
    if hit is not None:
        ref, nnz, h = hit
        if ref() is L and nnz == L.nnz:
            return h
    h = GraphHandle.from_laplacian(L, lmax)
    try:
        ref = weakref.ref(L, lambda _r, k=key: _handle_cache.pop(k, None))
    except TypeError:
        ref = (lambda obj: (lambda: obj))(L)
    _handle_cache[key] = (ref, L.nnz, h)
    return h


# ---------------------------------------------------------------------------------------
# Chebyshev basis (graph.chebyshev)
# ---------------------------------------------------------------------------------------

def cheb_basis(handle, X, K, transpose=False, flags=FILTER_DEFAULT):
    """Xt [K, M, C] = T_k(L~) X for X [M, C] on the device (lib/graph.py:241-258)."""
    _require_cuda(X)
    X = _f32c(X)
    M, C = X.shape
    if M != handle.M:
        raise ValueError('X has %d rows, graph has %d vertices' % (M, handle.M))
    Xt = torch.empty((K, M, C), dtype=torch.float32, device=X.device)
    check(_native.lib().cg_cheb_basis(handle.handle, int(bool(transpose)), ptr(X), ptr(Xt), C, K, int(flags),
                                      _stream()),
          'cg_cheb_basis')
    return Xt


# ---------------------------------------------------------------------------------------
# Chebyshev filter
# ---------------------------------------------------------------------------------------

_STACK_SAVE_LIMIT = 8 << 30      # bytes: larger bases are recomputed in the backward pass instead of saved
_save_stack = True


def set_precision(mode):
    """'fp32' (default: bf16 hi+mid split, three tensor-core passes, inside rtol 1e-4) or 'bf16' (single pass, BASELINE's
    2e-2 tolerance) for every tensor-core product of the library; storage and the recurrence stay fp32."""
    check(_native.lib().cg_set_precision({'fp32': 0, 'bf16': 1}[mode]), 'cg_set_precision')


def get_precision():
    return 'bf16' if _native.lib().cg_get_precision() == 1 else 'fp32'


_stack_planes = True
FILTER_STACK_PLANES = 32


def save_stack_enabled():
    return _save_stack


def set_stack_planes(flag):
    """Saved basis as the fused kernel's bf16 operand planes (default, where the library supports the shape) or
    always as the fp32 [K, N, M, Fin] stack (the layout of the reference's graph.chebyshev)."""
    global _stack_planes
    _stack_planes = bool(flag)


def set_save_stack(flag):
    """Let the forward pass keep the Chebyshev basis for the backward pass (default) or always recompute it."""
    global _save_stack
    _save_stack = bool(flag)


class ChebFilterFn(torch.autograd.Function):
    """y = chebyshev5(x; L~, W)  (lib/models.py:192-224).  x [N,M,Fin], W [Fin*K, Fout]."""

    @staticmethod
    def forward(ctx, x, W, handle, K, grad_x, flags):
        _require_cuda(x, W)
        x = _f32c(x)
        W = _f32c(W)
        N, M, Fin = x.shape
        if M != handle.M:
            raise ValueError('x has %d vertices, graph has %d' % (M, handle.M))
        if W.shape[0] != Fin * K:
            raise ValueError('W must be [Fin*K, Fout] = [%d, *], got %r' % (Fin * K, tuple(W.shape)))
        Fout = W.shape[1]
        lib = _native.lib()
        y = torch.empty((N, M, Fout), dtype=torch.float32, device=x.device)
        nbytes = lib.cg_cheb_filter_fwd_workspace_bytes(handle.handle, N, Fin, Fout, K, flags)
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=x.device)
        # when a gradient will be asked for, let the forward leave the basis X_k behind (shapes that allow it)
        stack = None
        if any(ctx.needs_input_grad[:2]) and save_stack_enabled():
            sbytes = lib.cg_cheb_filter_stack_bytes(handle.handle, N, Fin, Fout, K, flags)
            if 0 < sbytes <= _STACK_SAVE_LIMIT:
                if _stack_planes and lib.cg_cheb_filter_stack_planes(handle.handle, N, Fin, Fout, K, flags):
                    # bf16 hi | mid operand planes for the dW kernel: [2, K, ceil(N*M/128), Fin/8, 128, 8]
                    flags |= FILTER_STACK_PLANES
                    sbytes = lib.cg_cheb_filter_stack_bytes(handle.handle, N, Fin, Fout, K, flags)
                    stack = torch.empty((sbytes // 2,), dtype=torch.bfloat16, device=x.device)
                else:
                    stack = torch.empty((K, N, M, Fin), dtype=torch.float32, device=x.device)
        check(lib.cg_cheb_filter_fwd_ex(handle.handle, ptr(x), ptr(W), ptr(y), ptr(stack), N, Fin, Fout, K, ptr(ws),
                                        nbytes, flags, _stream()), 'cg_cheb_filter_fwd_ex')
        ctx.stack_planes = bool(flags & FILTER_STACK_PLANES)
        # the saved basis goes through save_for_backward: autograd frees it with the graph and keeps it under
        # retain_graph=True (a second backward sees the same basis, not a NULL pointer)
        ctx.save_for_backward(x, W, stack)
        ctx.handle, ctx.K, ctx.grad_x, ctx.flags = handle, K, grad_x, flags
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, stack = ctx.saved_tensors
        handle, K, flags = ctx.handle, ctx.K, ctx.flags
        gy = _f32c(gy)
        N, M, Fin = x.shape
        Fout = W.shape[1]
        lib = _native.lib()
        need_dx = bool(ctx.needs_input_grad[0] and ctx.grad_x)
        dx = torch.empty_like(x) if need_dx else None
        dW = torch.empty_like(W)
        nbytes = lib.cg_cheb_filter_bwd_workspace_bytes(handle.handle, N, Fin, Fout, K, int(need_dx), flags)
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=x.device)
        check(lib.cg_cheb_filter_bwd_ex(handle.handle, ptr(x), ptr(W), ptr(gy), ptr(stack), ptr(dx), ptr(dW), N,
                                        Fin, Fout, K, ptr(ws), nbytes, flags, _stream()), 'cg_cheb_filter_bwd_ex')
        return dx, dW, None, None, None, None


def cheb_filter(x, W, L, K, lmax=2, grad_x=True, flags=FILTER_DEFAULT):
    """Functional Chebyshev filter; ``L`` is a scipy Laplacian (rescaled here) or a GraphHandle."""
    if x.is_meta:     # shape tracing while a model declares its variables (no device work)
        return x.new_empty((x.shape[0], x.shape[1], W.shape[1]))
    Fin = int(x.shape[2])
    if Fin > 1 and Fin % 4 != 0 and W.shape[1] > 128 and W.shape[0] == Fin * K:
        # narrow inputs under wide outputs (the x path of the gconv-LSTM: Fin = 2, Fout = 4H): zero features up to a
        # multiple of 4 keep every operand 16-byte aligned for the pipelined tensor-core GEMM.  W rows are f*K + k, so
        # the rows of the extra features are appended; autograd slices the gradients back.
        pad = 4 - Fin % 4
        x = torch.cat([x, x.new_zeros((x.shape[0], x.shape[1], pad))], dim=2)
        W = torch.cat([W, W.new_zeros((pad * K, W.shape[1]))], dim=0)
    return ChebFilterFn.apply(x, W, get_handle(L, lmax), int(K), bool(grad_x), int(flags))


# ---------------------------------------------------------------------------------------
# bias + activation, pooling
# ---------------------------------------------------------------------------------------

ACT = {'none': 0, 'relu': 1, 'tanh': 2}


class BiasActFn(torch.autograd.Function):
    """act(x + bias) with bias None | [F] | [M, F]  (lib/models.py:226-247)."""

    @staticmethod
    def forward(ctx, x, bias, act):
        _require_cuda(x, bias)
        x = _f32c(x)
        N, M, F = x.shape
        kind = 0
        if bias is not None:
            bias = _f32c(bias)
            if bias.numel() == F:
                kind = 1
            elif bias.numel() == M * F:
                kind = 2
            else:
                raise ValueError('bias must have F=%d or M*F=%d entries, got %d' % (F, M * F, bias.numel()))
        y = torch.empty_like(x)
        check(_native.lib().cg_bias_act_fwd(ptr(x), ptr(bias), ptr(y), N, M, F, kind, act, _stream()),
              'cg_bias_act_fwd')
        ctx.save_for_backward(y)
        ctx.kind, ctx.act = kind, act
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        gy = _f32c(gy)
        N, M, F = y.shape
        gx = torch.empty_like(y)
        db = None
        if ctx.kind != 0 and ctx.needs_input_grad[1]:
            db = torch.empty(ctx.bias_shape, dtype=torch.float32, device=y.device)
        check(_native.lib().cg_bias_act_bwd(ptr(y), ptr(gy), ptr(gx), ptr(db), N, M, F, ctx.kind, ctx.act,
                                            _stream()), 'cg_bias_act_bwd')
        return gx, db, None


def bias_act(x, bias, act):
    if x.is_meta:
        return x.new_empty(x.shape)
    return BiasActFn.apply(x, bias, ACT[act] if isinstance(act, str) else int(act))


class PoolFn(torch.autograd.Function):
    """mpool1 (kind 1) / apool1 (kind 2) over p consecutive vertices (lib/models.py:249-266)."""

    @staticmethod
    def forward(ctx, x, p, kind):
        _require_cuda(x)
        x = _f32c(x)
        N, M, F = x.shape
        if M % p != 0:
            raise ValueError('pool size %d does not divide M=%d' % (p, M))
        y = torch.empty((N, M // p, F), dtype=torch.float32, device=x.device)
        amax = torch.empty((N, M // p, F), dtype=torch.uint8, device=x.device) if kind == 1 else None
        check(_native.lib().cg_pool_fwd(ptr(x), ptr(y), ptr(amax), N, M, F, p, kind, _stream()), 'cg_pool_fwd')
        ctx.amax, ctx.p, ctx.kind, ctx.M = amax, p, kind, M
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = _f32c(gy)
        N, Mp, F = gy.shape
        gx = torch.empty((N, ctx.M, F), dtype=torch.float32, device=gy.device)
        check(_native.lib().cg_pool_bwd(ptr(gy), ptr(ctx.amax), ptr(gx), N, ctx.M, F, ctx.p, ctx.kind, _stream()),
              'cg_pool_bwd')
        return gx, None, None


def pool(x, p, kind):
    if p <= 1:
        return x
    if x.is_meta:
        return x.new_empty((x.shape[0], x.shape[1] // p, x.shape[2]))
    return PoolFn.apply(x, int(p), 1 if kind in (1, 'max') else 2)


class BiasActPoolFn(torch.autograd.Function):
    """pool_p(act(x + bias)) in one pass (the brelu -> pool tail of a cgcnn layer, lib/models.py:226-266)."""

    @staticmethod
    def forward(ctx, x, bias, act, p, kind):
        _require_cuda(x, bias)
        x = _f32c(x)
        N, M, F = x.shape
        bkind = 0
        if bias is not None:
            bias = _f32c(bias)
            if bias.numel() == F:
                bkind = 1
            elif bias.numel() == M * F:
                bkind = 2
            else:
                raise ValueError('bias must have F=%d or M*F=%d entries, got %d' % (F, M * F, bias.numel()))
        if M % p != 0:
            raise ValueError('pool size %d does not divide M=%d' % (p, M))
        y = torch.empty((N, M // p, F), dtype=torch.float32, device=x.device)
        aux = torch.empty((N, M // p, F), dtype=torch.uint8, device=x.device)
        check(_native.lib().cg_bias_act_pool_fwd(ptr(x), ptr(bias), ptr(y), ptr(aux), N, M, F, p, bkind, act, kind,
                                                 _stream()), 'cg_bias_act_pool_fwd')
        ctx.save_for_backward(y, aux)
        ctx.cfg = (N, M, F, p, bkind, act, kind)
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        y, aux = ctx.saved_tensors
        N, M, F, p, bkind, act, kind = ctx.cfg
        gy = _f32c(gy)
        gx = torch.empty((N, M, F), dtype=torch.float32, device=y.device)
        db = None
        if bkind != 0 and ctx.needs_input_grad[1]:
            db = torch.empty(ctx.bias_shape, dtype=torch.float32, device=y.device)
        check(_native.lib().cg_bias_act_pool_bwd(ptr(gy), ptr(y), ptr(aux), ptr(gx), ptr(db), N, M, F, p, bkind, act,
                                                 kind, _stream()), 'cg_bias_act_pool_bwd')
        return gx, db, None, None, None


class FirstLayerFn(torch.autograd.Function):
    """pool_max4(relu(chebyshev5(x; W) + b)) for a scalar-input first layer (x [N, M, 1], no gradient to x) as ONE
    autograd node: the forward is the usual filter + fused bias/relu/pool pair, the backward goes from the gradient
    of the pooled output straight to dW and db (cg_cheb_dw_pooled) -- no pooling-backward pass, no [N, M, Fout]
    gradient tensor.  Same values as bias_act_pool(cheb_filter(x, W), b) and its gradients."""

    @staticmethod
    def forward(ctx, x, W, bias, handle, K):
        _require_cuda(x, W, bias)
        x, W = _f32c(x), _f32c(W)
        N, M, Fin = x.shape
        if Fin != 1 or M % 4 != 0 or M != handle.M or W.shape[0] != K:
            raise ValueError('FirstLayerFn: needs x [N, M, 1] with M %% 4 == 0 on the graph (M = %d) and W [K, Fout]; got x %r, '
                             'W %r, K = %d' % (handle.M, tuple(x.shape), tuple(W.shape), K))
        Fout = W.shape[1]
        lib = _native.lib()
        bkind = 0
        if bias is not None:
            bias = _f32c(bias)
            if bias.numel() != Fout:
                raise ValueError('bias must have Fout=%d entries, got %d' % (Fout, bias.numel()))
            bkind = 1
        stack = torch.empty((K, N, M, 1), dtype=torch.float32, device=x.device)
        yp = torch.empty((N, M // 4, Fout), dtype=torch.float32, device=x.device)
        aux = torch.empty((N, M // 4, Fout), dtype=torch.uint8, device=x.device)
        if _first_layer_epilogue and lib.cg_cheb_first_layer_fwd_supported(handle.handle, N, Fout, K, bkind):
            # recurrence + contraction with bias / relu / pool in its epilogue: the [N, M, Fout] output is never written
            check(lib.cg_cheb_first_layer_fwd(handle.handle, ptr(x), ptr(W), ptr(bias), ptr(stack), ptr(yp), ptr(aux), N, Fout, K,
                                              _stream()), 'cg_cheb_first_layer_fwd')
        else:
            y = torch.empty((N, M, Fout), dtype=torch.float32, device=x.device)
            nbytes = lib.cg_cheb_filter_fwd_workspace_bytes(handle.handle, N, 1, Fout, K, 0)
            ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=x.device)
            check(lib.cg_cheb_filter_fwd_ex(handle.handle, ptr(x), ptr(W), ptr(y), ptr(stack), N, 1, Fout, K, ptr(ws), nbytes, 0,
                                            _stream()), 'cg_cheb_filter_fwd_ex')
            check(lib.cg_bias_act_pool_fwd(ptr(y), ptr(bias), ptr(yp), ptr(aux), N, M, Fout, 4, bkind, ACT['relu'], 1, _stream()),
                  'cg_bias_act_pool_fwd')
        ctx.save_for_backward(yp, aux, stack)
        ctx.handle, ctx.K = handle, K
        ctx.w_shape, ctx.bias_shape = tuple(W.shape), None if bias is None else tuple(bias.shape)
        return yp

    @staticmethod
    def backward(ctx, g):
        yp, aux, stack = ctx.saved_tensors
        g = _f32c(g)
        N, Mp, Fout = yp.shape
        K, handle = ctx.K, ctx.handle
        lib = _native.lib()
        dW = torch.empty(ctx.w_shape, dtype=torch.float32, device=yp.device)
        db = None
        if ctx.bias_shape is not None and ctx.needs_input_grad[2]:
            db = torch.empty(ctx.bias_shape, dtype=torch.float32, device=yp.device)
        nbytes = lib.cg_cheb_dw_pooled_workspace_bytes(handle.handle, N, Fout, K)
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=yp.device)
        check(lib.cg_cheb_dw_pooled(handle.handle, ptr(stack), ptr(g), ptr(yp), ptr(aux), ptr(dW), ptr(db), N, Fout, K,
                                    ptr(ws), nbytes, _stream()), 'cg_cheb_dw_pooled')
        return None, dW, db, None, None


_first_layer_fusion = True
_first_layer_epilogue = True      # bias / relu / pool inside the contraction's epilogue (False: separate pooling kernel)


def set_first_layer_fusion(flag):
    """Allow (default) or forbid the fused first-layer node (tests compare both paths)."""
    global _first_layer_fusion
    _first_layer_fusion = bool(flag)


def set_first_layer_epilogue(flag):
    global _first_layer_epilogue
    _first_layer_epilogue = bool(flag)


def first_layer_supported(x, W, bias, L, K, act, p, kind, lmax=2):
    """True when pool(brelu(cheb_filter(x, W))) can run as the fused first-layer node: scalar input signal that needs
    no gradient, relu, max pooling of 4, one bias per filter (or none), a shape the streaming dW kernel takes."""
    if not _first_layer_fusion or x.is_meta or x.dim() != 3 or x.shape[2] != 1 or x.requires_grad or not save_stack_enabled():
        return False
    if bias is not None and bias.numel() != W.shape[1]:
        return False
    act = ACT[act] if isinstance(act, str) else int(act)
    kind = 1 if kind in (1, 'max') else 2
    handle = get_handle(L, lmax)
    return bool(_native.lib().cg_cheb_dw_pooled_supported(handle.handle, int(x.shape[0]), int(W.shape[1]), int(K), int(p), act,
                                                          kind, 0 if bias is None else 1))


def first_layer(x, W, bias, L, K, lmax=2):
    """pool_max4(relu(cheb_filter(x, W) + bias)); check first_layer_supported first."""
    return FirstLayerFn.apply(x, W, bias, get_handle(L, lmax), int(K))


def bias_act_pool_supported(act, p, kind):
    act = ACT[act] if isinstance(act, str) else int(act)
    kind = 1 if kind in (1, 'max') else 2
    return 1 < p <= 8 and not (kind == 2 and act == 2)


def bias_act_pool(x, bias, act, p, kind):
    """Fused bias + activation + pooling; same values as pool(bias_act(x, bias, act), p, kind)."""
    act = ACT[act] if isinstance(act, str) else int(act)
    kind = 1 if kind in (1, 'max') else 2
    if x.is_meta:
        return x.new_empty((x.shape[0], x.shape[1] // p, x.shape[2]))
    return BiasActPoolFn.apply(x, bias, act, int(p), kind)


def pool_argmax(x, p):
    """Pooled values and first-max indices (uint8) -- used by the bit-exact parity tests."""
    _require_cuda(x)
    x = _f32c(x)
    N, M, F = x.shape
    y = torch.empty((N, M // p, F), dtype=torch.float32, device=x.device)
    amax = torch.empty((N, M // p, F), dtype=torch.uint8, device=x.device)
    check(_native.lib().cg_pool_fwd(ptr(x), ptr(y), ptr(amax), N, M, F, p, 1, _stream()), 'cg_pool_fwd')
    return y, amax


# ---------------------------------------------------------------------------------------
# dense head (fc layers) on the tensor cores
# ---------------------------------------------------------------------------------------

def gemm(A, B, transA=False, transB=False, bias=None, relu=False):
    """C = op(A) @ op(B) (+ bias) (relu) through cg_gemm_f32; A, B 2-D float32 CUDA tensors (row-major)."""
    _require_cuda(A, B, bias)
    A, B = _f32c(A), _f32c(B)
    M, K = (A.shape[1], A.shape[0]) if transA else (A.shape[0], A.shape[1])
    Kb, N = (B.shape[1], B.shape[0]) if transB else (B.shape[0], B.shape[1])
    if K != Kb:
        raise ValueError('inner dimensions differ: %d vs %d' % (K, Kb))
    if bias is not None:
        bias = _f32c(bias).reshape(-1)
        if bias.numel() != N:
            raise ValueError('bias must have N=%d entries' % N)
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    lib = _native.lib()
    nbytes = lib.cg_gemm_f32_workspace_bytes(M, N, K)
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=A.device)
    check(lib.cg_gemm_f32(ptr(A), ptr(B), ptr(C), M, N, K, int(transA), int(transB), A.shape[1], B.shape[1], N,
                          ptr(bias), int(relu), ptr(ws), nbytes, _stream()), 'cg_gemm_f32')
    return C


class LinearFn(torch.autograd.Function):
    """y = relu?(x W + b)  (lib/models.py:268-274) with both gradients on the tensor cores."""

    @staticmethod
    def forward(ctx, x, W, b, relu):
        y = gemm(x, W, bias=b, relu=relu)
        ctx.save_for_backward(x, W, y if relu else None)
        ctx.relu = relu
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, y = ctx.saved_tensors
        gy = _f32c(gy)
        if ctx.relu:
            gy = gy * (y > 0)
        dx = gemm(gy, W, transB=True) if ctx.needs_input_grad[0] else None          # [N, out] . [in, out]^T
        dW = gemm(x, gy, transA=True) if ctx.needs_input_grad[1] else None          # [N, in]^T . [N, out]
        db = gy.sum(dim=0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dW, db, None


def linear(x, W, b=None, relu=False):
    if x.is_meta:
        return x.new_empty((x.shape[0], W.shape[1]))
    return LinearFn.apply(x, W, b, bool(relu))


# ---------------------------------------------------------------------------------------
# spectral (Fourier) filter
# ---------------------------------------------------------------------------------------

def bmm(A, B, transA=False, transB=False):
    """C[b] = op(A[b]) @ op(B[b]) through cg_bmm_f32; A, B 3-D float32 CUDA tensors."""
    _require_cuda(A, B)
    A, B = _f32c(A), _f32c(B)
    batch = A.shape[0]
    m, k = (A.shape[2], A.shape[1]) if transA else (A.shape[1], A.shape[2])
    kb, n = (B.shape[2], B.shape[1]) if transB else (B.shape[1], B.shape[2])
    if k != kb or B.shape[0] != batch:
        raise ValueError('bmm: incompatible operands %r, %r' % (tuple(A.shape), tuple(B.shape)))
    C = torch.empty((batch, m, n), dtype=torch.float32, device=A.device)
    check(_native.lib().cg_bmm_f32(ptr(A), ptr(B), ptr(C), batch, m, n, k, int(transA), int(transB), A.shape[2], B.shape[2], n,
                                   A.shape[1] * A.shape[2], B.shape[1] * B.shape[2], m * n, _stream()), 'cg_bmm_f32')
    return C


class FourierFilterFn(torch.autograd.Function):
    """y = U (W_m (U^T x)_m)_m  (lib/filter.py:11-27, lib/models.py:129-144): x [N, M, Fin], W [M, Fout, Fin], U [M, M]
    (columns = eigenvectors of L) -> y [N, M, Fout].  Two dense transforms on the tensor-core GEMM (cg_gemm_f32) around
    one batched per-frequency product (cg_bmm_f32); the backward is the same three products transposed."""

    @staticmethod
    def forward(ctx, x, W, U):
        _require_cuda(x, W, U)
        x, W, U = _f32c(x), _f32c(W), _f32c(U)
        N, M, Fin = x.shape
        Fout = W.shape[1]
        if tuple(W.shape) != (M, Fout, Fin) or tuple(U.shape) != (M, M):
            raise ValueError('fourier filter: W must be [M, Fout, Fin] and U [M, M]; got %r, %r' % (tuple(W.shape), tuple(U.shape)))
        xm = x.permute(1, 0, 2).reshape(M, N * Fin)                 # vertex-major view of the batch
        xh = gemm(U, xm, transA=True).view(M, N, Fin)               # U^T x: graph Fourier transform
        yh = bmm(xh, W, transB=True)                                # per frequency m: [N, Fin] . W[m]^T -> [N, Fout]
        ym = gemm(U, yh.view(M, N * Fout))                          # back to the vertex domain
        ctx.save_for_backward(xh, W, U)
        ctx.dims = (N, M, Fin, Fout)
        return ym.view(M, N, Fout).permute(1, 0, 2).contiguous()

    @staticmethod
    def backward(ctx, gy):
        xh, W, U = ctx.saved_tensors
        N, M, Fin, Fout = ctx.dims
        gm = _f32c(gy).permute(1, 0, 2).reshape(M, N * Fout)
        gyh = gemm(U, gm, transA=True).view(M, N, Fout)             # U^T gy
        dx = dW = None
        if ctx.needs_input_grad[1]:
            dW = bmm(gyh, xh, transA=True)                          # [Fout, N] . [N, Fin] per frequency
        if ctx.needs_input_grad[0]:
            gxh = bmm(gyh, W)                                       # [N, Fout] . [Fout, Fin]
            dx = gemm(U, gxh.view(M, N * Fin)).view(M, N, Fin).permute(1, 0, 2).contiguous()
        return dx, dW, None


_fourier_cache = {}


def fourier_basis(L):
    """Device copy of U = eigenvectors of L (graph.fourier, lib/graph.py:148-166: numpy eigh on the host, once per
    Laplacian object), cached like the packed operators."""
    dev = torch.cuda.current_device()
    key = (id(L), dev)
    hit = _fourier_cache.get(key)
    if hit is not None and hit[0]() is L:
        return hit[1]
    _, U = np.linalg.eigh(L.toarray())
    Ut = torch.as_tensor(np.ascontiguousarray(U, dtype=np.float32), device=torch.device('cuda', dev))
    try:
        ref = weakref.ref(L, lambda _r, k=key: _fourier_cache.pop(k, None))
    except TypeError:
        ref = (lambda obj: (lambda: obj))(L)
    _fourier_cache[key] = (ref, Ut)
    return Ut


def fourier_filter(x, W, L, U=None):
    """Spectral filter with one Fout x Fin weight matrix per graph frequency; ``U`` overrides the cached eigenbasis."""
    if x.is_meta:
        return x.new_empty((x.shape[0], x.shape[1], W.shape[1]))
    if U is None:
        U = fourier_basis(L)
    return FourierFilterFn.apply(x, W, U)


# ---------------------------------------------------------------------------------------
# perm_data on the device
# ---------------------------------------------------------------------------------------

def perm_data_device(x, perm, out=None, perm_t=None):
    """out[:, i] = x[:, perm[i]] if perm[i] < M else 0  (lib/coarsening.py:219-240), float32.
    ``perm_t`` (int32 device tensor) skips the upload of ``perm``; ``out`` is written in place when given."""
    _require_cuda(x)
    x = _f32c(x)
    N, M = x.shape
    if perm_t is None:
        perm_t = torch.as_tensor(np.asarray(perm, dtype=np.int32), device=x.device)
    if out is None:
        out = torch.empty((N, perm_t.numel()), dtype=torch.float32, device=x.device)
    elif out.shape != (N, perm_t.numel()) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError('perm_data_device: out must be a contiguous float32 [N, len(perm)] tensor')
    check(_native.lib().cg_perm_data(ptr(x), ptr(perm_t), ptr(out), N, M, perm_t.numel(), _stream()), 'cg_perm_data')
    return out


# ---------------------------------------------------------------------------------------
# gconv-LSTM gates
# ---------------------------------------------------------------------------------------

class LstmGatesFn(torch.autograd.Function):
    """Gate nonlinearities + state update of GConvLSTMCell (lib/gconv_lstm.py:185-215).

    pre [N, M, 4H] (z|i|f|o) (+ pre2, the second addend: x-path and h-path filters are summed inside the gate kernels),
    bias [4H], c [N, M, H]  ->  new_h, new_c.
    """

    @staticmethod
    def forward(ctx, pre, pre2, bias, c, variant):
        _require_cuda(pre, bias, c)
        pre, bias, c = _f32c(pre), _f32c(bias), _f32c(c)
        pre2 = _f32c(pre2) if pre2 is not None else None
        H = c.shape[-1]
        R = c.numel() // H
        new_c = torch.empty_like(c)
        new_h = torch.empty_like(c)
        check(_native.lib().cg_lstm_gates2_fwd(ptr(pre), ptr(pre2), ptr(bias), ptr(c), ptr(new_c), ptr(new_h), R, H, variant,
                                               _stream()), 'cg_lstm_gates2_fwd')
        ctx.save_for_backward(pre, pre2, bias, c, new_c)
        ctx.variant = variant
        return new_h, new_c

    @staticmethod
    def backward(ctx, g_h, g_c):
        pre, pre2, bias, c, new_c = ctx.saved_tensors
        H = c.shape[-1]
        R = c.numel() // H
        g_h = _f32c(g_h) if g_h is not None else None
        g_c = _f32c(g_c) if g_c is not None else None
        g_pre = torch.empty_like(pre)
        g_cprev = torch.empty_like(c)
        d_bias = torch.empty_like(bias)
        check(_native.lib().cg_lstm_gates2_bwd(ptr(pre), ptr(pre2), ptr(bias), ptr(c), ptr(new_c), ptr(g_h), ptr(g_c),
                                               ptr(g_pre), ptr(g_cprev), ptr(d_bias), R, H, ctx.variant, _stream()),
              'cg_lstm_gates2_bwd')
        return g_pre, (g_pre if pre2 is not None else None), d_bias, g_cprev, None


def lstm_gates(pre, bias, c, variant='fork', pre2=None):
    """pre (+ pre2) are the gate pre-activations without bias; returns (new_h, new_c)."""
    v = {'fork': 0, 'standard': 1}[variant] if isinstance(variant, str) else int(variant)
    if pre.is_meta:
        return c.new_empty(c.shape), c.new_empty(c.shape)
    return LstmGatesFn.apply(pre, pre2, bias, c, v)


def mark_zero(t):
    """Tag a tensor as known all-zero (initial LSTM state): a linear filter of it is exactly zero and can be skipped."""
    t._cg_zero = True
    return t


def is_marked_zero(t):
    return bool(getattr(t, '_cg_zero', False))


# ---------------------------------------------------------------------------------------
# loss + optimiser tail (native: one launch each)
# ---------------------------------------------------------------------------------------

class SoftmaxXentFn(torch.autograd.Function):
    """mean softmax cross-entropy of logits [N, C] against int64 labels [N] (lib/graph_model.py:250-252); the forward
    launch also leaves the gradient (softmax - onehot) / N behind."""

    @staticmethod
    def forward(ctx, logits, labels):
        _require_cuda(logits, labels)
        logits = _f32c(logits)
        labels = labels.to(torch.int64).contiguous()
        N, C = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dz = torch.empty_like(logits)
        check(_native.lib().cg_softmax_xent(ptr(logits), ptr(labels), ptr(loss), ptr(dz), N, C, _stream()), 'cg_softmax_xent')
        ctx.save_for_backward(dz)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return dz * g, None


def softmax_xent(logits, labels):
    return SoftmaxXentFn.apply(logits, labels)


class NativeMomentumSGD(torch.optim.Optimizer):
    """torch.optim.SGD(momentum) semantics with the whole update in ONE native launch (cg_sgd_momentum): the
    {param, grad, momentum buffer, numel} records travel as kernel arguments, so a captured CUDA graph replays the
    addresses it was captured with (its private pool keeps them alive).  Gradients are read where autograd left them."""

    def __init__(self, params, lr, momentum=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum))

    @torch.no_grad()
    def apply(self, params, grads, lr, momentum, lr_dev=None):
        """buffer = momentum * buffer + grad; param -= lr * buffer for the given (param, grad) pairs in one launch; with
        `lr_dev` (a one-element float32 device tensor) the rate is read on the device when the kernel runs."""
        rec = []
        for p, g in zip(params, grads):
            st = self.state[p]
            if 'momentum_buffer' not in st or st['momentum_buffer'] is None:
                st['momentum_buffer'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous() or not g.is_contiguous():
                raise TypeError('NativeMomentumSGD: fp32 contiguous variables and gradients only')
            rec += [p.data_ptr(), g.data_ptr(), st['momentum_buffer'].data_ptr(), p.numel()]
        if not rec:
            return
        table = np.array(rec, dtype=np.int64)
        check(_native.lib().cg_sgd_momentum_dev(table.ctypes.data, len(params), max(p.numel() for p in params), ctypes.c_float(lr),
                                                ptr(lr_dev), ctypes.c_float(momentum), _stream()), 'cg_sgd_momentum_dev')

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.grad is not None]
            for p in ps:
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
            self.apply(ps, [p.grad for p in ps], group['lr'], group['momentum'])
        return None


class NativeAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad) with the whole step in ONE native launch (cg_adam).
    The step count lives on the device next to the first variable's moments, so a captured CUDA graph keeps counting
    when it is replayed (what torch needs ``capturable=True`` and ~25 small launches per step for)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            rec = []
            for p in ps:
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError('NativeAdam: fp32 contiguous variables and gradients only')
                st = self.state[p]
                if 'exp_avg' not in st:
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                rec += [p.data_ptr(), p.grad.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), p.numel()]
            # {steps done, scratch}: kept in the state of the group's first variable, so whatever snapshots / restores the
            # optimiser state (GraphModel._capture_step) carries the step count along
            head = self.state[group['params'][0]]
            if 'step_state' not in head:
                head['step_state'] = torch.zeros(2, dtype=torch.int32, device=ps[0].device)
            table = np.array(rec, dtype=np.int64)
            b1, b2 = group['betas']
            check(_native.lib().cg_adam(table.ctypes.data, len(ps), max(p.numel() for p in ps), ctypes.c_float(group['lr']),
                                        ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(group['eps']),
                                        ptr(head['step_state']), _stream()), 'cg_adam')
        return None


# ---------------------------------------------------------------------------------------
# sparse input batches
# ---------------------------------------------------------------------------------------

def csr_densify(indptr, indices, values, M, out_rows=None, out=None):
    """Dense [out_rows, M] float32 device tensor of a CSR batch whose arrays are already on the device (int32 indptr /
    indices, float32 values); rows beyond the CSR stay zero."""
    _require_cuda(indptr, indices, values)
    rows = indptr.numel() - 1
    out_rows = rows if out_rows is None else int(out_rows)
    if out is None:
        out = torch.empty((out_rows, M), dtype=torch.float32, device=indptr.device)
    elif tuple(out.shape) != (out_rows, M) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError('csr_densify: out must be a contiguous float32 [%d, %d] tensor' % (out_rows, M))
    check(_native.lib().cg_csr_densify(ptr(indptr), ptr(indices), ptr(values), ptr(out), rows, out_rows, M, _stream()),
          'cg_csr_densify')
    return out


def sparse_batch_to_device(a, device, out_rows=None):
    """scipy sparse batch -> dense float32 device tensor without a host-side toarray(): uploads the CSR arrays (about
    1 % of the dense bytes for bag-of-words rows) and expands them on the device (cg_csr_densify)."""
    a = scipy.sparse.csr_matrix(a, dtype=np.float32)
    ip = torch.from_numpy(a.indptr.astype(np.int32)).to(device, non_blocking=True)
    ix = torch.from_numpy(a.indices.astype(np.int32)).to(device, non_blocking=True)
    va = torch.from_numpy(np.ascontiguousarray(a.data, dtype=np.float32)).to(device, non_blocking=True)
    return csr_densify(ip, ix, va, a.shape[1], out_rows=out_rows)
