"""tcgen05.mma issue/completion latency versus batch size (one CTA, one commit per batch)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cnn_graph_b200 import _native
lib = ctypes.CDLL(_native.LIB_PATH)
fn = lib.cg_debug_umma_bench
fn.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device='cuda')
for N in (64,):
    for mode in (0, 1):
        for per in (1, 2, 4, 8, 12, 16, 24, 32, 64, 128):
            fn(mode, N, per, 1, out.data_ptr(), 0, 0, None)
            torch.cuda.synchronize()
            tot, iss = out.tolist()
            print('N=%3d mode=%d batch=%3d: total %6d clk, issue %6d clk' % (N, mode, per, tot, iss))
