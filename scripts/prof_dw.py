"""Time the layer-2 filter of config C2 (forward with plane side output, backward) kernel by kernel."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cnn_graph_b200 import _native, ops
L, perm = bench.build_graphs()
lib = _native.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = L[2].shape[0]
torch.manual_seed(0)
x = torch.randn(N, M, 32, device='cuda', requires_grad=True)
W = (0.1 * torch.randn(32 * 25, 64, device='cuda')).requires_grad_(True)
gy = torch.randn(N, M, 64, device='cuda')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
for it in range(4):
    flush.fill_(0.)
    if it == 3:
        lib.cg_profile_enable(1); lib.cg_profile_reset()
    y = ops.cheb_filter(x, W, L[2], 25)
    y.backward(gy)
torch.cuda.synchronize()
name = ctypes.create_string_buffer(64); ms = ctypes.c_double(); cnt = ctypes.c_int64()
n = lib.cg_profile_query(-1, name, 64, ctypes.byref(ms), ctypes.byref(cnt))
for i in range(n):
    lib.cg_profile_query(i, name, 64, ctypes.byref(ms), ctypes.byref(cnt))
    print('  %-18s %8.3f ms  x%d' % (name.value.decode(), ms.value, cnt.value))
