"""Config C4: gconv-LSTM cell on a 32x32 8-NN grid (M = 1024), batch 50, Fin = 2, H = 128, K = 3:
forward + backward of T unrolled cell steps (lib/gconv_lstm.py:77-221)."""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cnn_graph_b200 import _native, ops
from cnn_graph_b200.lib import graph, gconv_lstm, variables
ap = argparse.ArgumentParser()
ap.add_argument('--H', type=int, default=128); ap.add_argument('--K', type=int, default=3)
ap.add_argument('--T', type=int, default=3); ap.add_argument('--N', type=int, default=50)
ap.add_argument('--iters', type=int, default=4)
a = ap.parse_args()
A = graph.adjacency(*graph.distance_sklearn_metrics(graph.grid(32), k=8, metric='euclidean'))
L = graph.laplacian(A, normalized=True)
cell = gconv_lstm.GConvLSTMCell(a.H, 2, L, 2, nNode=1024, K=a.K) if False else None
lib = _native.lib()
torch.manual_seed(0)
N, M, H, K = a.N, 1024, a.H, a.K
Fin = 2
Wx = (0.1 * torch.randn(Fin * K, 4 * H, device='cuda')).requires_grad_(True)
Wh = (0.1 * torch.randn(H * K, 4 * H, device='cuda')).requires_grad_(True)
bias = torch.zeros(4 * H, device='cuda', requires_grad=True)
xs = [torch.rand(N, M, Fin, device='cuda') for _ in range(a.T)]
def run():
    c = torch.zeros(N, M, H, device='cuda'); h = torch.zeros(N, M, H, device='cuda')
    for t in range(a.T):
        pre = ops.cheb_filter(xs[t], Wx, L, K) + ops.cheb_filter(h, Wh, L, K)
        h, c = ops.lstm_gates(pre, bias, c, 'standard')
    return h
for it in range(a.iters):
    if it == a.iters - 1:
        lib.cg_profile_enable(1); lib.cg_profile_reset()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); h = run(); e[1].record(); h.sum().backward(); e[2].record(); torch.cuda.synchronize()
    print('iter %d: %d cell steps fwd %.3f ms, bwd %.3f ms' % (it, a.T, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
name = ctypes.create_string_buffer(64); ms = ctypes.c_double(); cnt = ctypes.c_int64()
n = lib.cg_profile_query(-1, name, 64, ctypes.byref(ms), ctypes.byref(cnt))
rows = []
for i in range(n):
    lib.cg_profile_query(i, name, 64, ctypes.byref(ms), ctypes.byref(cnt)); rows.append((ms.value, name.value.decode(), cnt.value))
for ms_, nm, c_ in sorted(rows, reverse=True): print('  %-18s %8.3f ms  x%d' % (nm, ms_, c_))
