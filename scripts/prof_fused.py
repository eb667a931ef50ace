"""Time (CUDA events) one Chebyshev filter forward/backward at a C2 layer shape -- the command ncu wraps."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import scipy.sparse
from cnn_graph_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument('--level', type=int, default=2)
ap.add_argument('--N', type=int, default=1024)
ap.add_argument('--Fin', type=int, default=32)
ap.add_argument('--Fout', type=int, default=64)
ap.add_argument('--K', type=int, default=25)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--flags', type=int, default=0)
ap.add_argument('--bwd', type=int, default=1)
ap.add_argument('--gradx', type=int, default=1)
ap.add_argument('--kernels', type=int, default=0)
a = ap.parse_args()
c2 = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'c2_grid28.npz'))
pre = 'L%d' % a.level
L = scipy.sparse.csr_matrix((c2[pre + '_data'], c2[pre + '_indices'], c2[pre + '_indptr']),
                            shape=tuple(int(v) for v in c2[pre + '_shape']))
M = L.shape[0]
torch.manual_seed(0)
x = torch.randn(a.N, M, a.Fin, device='cuda', requires_grad=bool(a.gradx))
W = (0.1 * torch.randn(a.Fin * a.K, a.Fout, device='cuda')).requires_grad_(True)
gy = torch.randn(a.N, M, a.Fout, device='cuda')
import ctypes
from cnn_graph_b200 import _native
_lib = _native.lib()
def kernel_times():
    out = {}
    name = ctypes.create_string_buffer(64); ms = ctypes.c_double(); cnt = ctypes.c_int64()
    n = _lib.cg_profile_query(-1, name, 64, ctypes.byref(ms), ctypes.byref(cnt))
    for i in range(n):
        _lib.cg_profile_query(i, name, 64, ctypes.byref(ms), ctypes.byref(cnt))
        out[name.value.decode()] = (ms.value, cnt.value)
    return out
for it in range(a.iters):
    if a.kernels and it == a.iters - 1:
        _lib.cg_profile_enable(1); _lib.cg_profile_reset()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    y = ops.cheb_filter(x, W, L, a.K, flags=a.flags)
    e[1].record()
    if a.bwd:
        y.backward(gy)
    e[2].record()
    torch.cuda.synchronize()
    print('iter %d fwd %.3f ms bwd %.3f ms' % (it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])), flush=True)
if a.kernels:
    for k, (ms, c) in sorted(kernel_times().items(), key=lambda kv: -kv[1][0]):
        print('  %-18s %8.3f ms  x%d' % (k, ms, c))
    _lib.cg_profile_enable(0)
if os.environ.get('CG_TRACE'):
    import ctypes
    from cnn_graph_b200 import _native
    lib = _native.lib()
    buf = torch.zeros(a.K * 10, dtype=torch.int64, device='cuda')
    lib._handle if False else None
    fn = ctypes.CDLL(_native.LIB_PATH).cg_debug_fused_trace
    fn.argtypes = [ctypes.c_void_p]
    fn(buf.data_ptr())
    if os.environ.get('CG_TRACE_NOGRAD'):      # inference form: no saved basis (no plane side output)
        with torch.no_grad():
            y = ops.cheb_filter(x, W, L, a.K, flags=a.flags)
    else:
        y = ops.cheb_filter(x, W, L, a.K, flags=a.flags)
    torch.cuda.synchronize()
    fn(None)
    t = buf.cpu().numpy().reshape(a.K, 10)
    t0 = t[0, 0]
    names = ['A_start', 'A_end', 'wait_end', 'B_end', 'I_sync', 'I_wbar', 'I_issued', 'I_done', 'S_start', 'S_done']
    print(' k ' + ' '.join('%9s' % n for n in names))
    for k in range(a.K):
        print('%2d ' % k + ' '.join('%9d' % (v - t0 if v else -1) for v in t[k]))

if os.environ.get('CG_TRACE_CL'):
    fn = _lib.cg_debug_clenshaw_trace
    buf = torch.zeros(a.K * 10, dtype=torch.int64, device='cuda')
    fn(buf.data_ptr())
    y = ops.cheb_filter(x, W, L, a.K, flags=a.flags)
    y.backward(gy)
    torch.cuda.synchronize()
    fn(None)
    t = buf.cpu().numpy().reshape(a.K, 10)
    t0 = t[0, 0]
    names = ['start', 'gathered', 'stored', 'dumped', 'synced', 'I_start', 'I_issued', 'I_done', '-', '-']
    print(' s ' + ' '.join('%9s' % n for n in names))
    for k in range(a.K):
        print('%2d ' % k + ' '.join('%9d' % (v - t0 if v else -1) for v in t[k]))

if os.environ.get('CG_TRACE_DW'):
    fn = ctypes.CDLL(_native.LIB_PATH).cg_debug_dw_trace
    fn.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(16 * 8, dtype=torch.int64, device='cuda')
    fn(buf.data_ptr())
    y = ops.cheb_filter(x, W, L, a.K, flags=a.flags)
    y.backward(gy)
    torch.cuda.synchronize()
    fn(None)
    t = buf.cpu().numpy().reshape(16, 8)
    t0 = t[0, 0]
    names = ['start', 'full', 'mmafree', 'conv_end', 'sync', 'I_sync', 'I_issued', 'I_loads']
    print(' c ' + ' '.join('%9s' % n for n in names))
    for c in range(16):
        print('%2d ' % c + ' '.join('%9d' % (v - t0 if v else -1) for v in t[c]))
    tt = buf.cpu().numpy()[64:80]
    print('per-MMA issue clocks (chunk 30):', [int(v - tt[0]) for v in tt if v])
