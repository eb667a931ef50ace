"""Probe: where do the 64 rows of an M=64 tcgen05.mma accumulator land in TMEM?"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cnn_graph_b200 import _native
lib = _native.lib()
for a_mn, b_mn, off in ((0, 0, 0), (0, 0, 2), (1, 1, 2)):
    N, Kd = 32, 64
    torch.manual_seed(0)
    A = torch.randn(64, Kd, device='cuda'); B = torch.randn(N, Kd, device='cuda')
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    A_src = A.t().contiguous() if a_mn else A
    B_src = B.t().contiguous() if b_mn else B
    D = torch.full((128, N), float('nan'), device='cuda')
    _native.check(lib.cg_debug_umma_gemm_m(A_src.data_ptr(), B_src.data_ptr(), D.data_ptr(), 64, N, Kd, a_mn | off, b_mn, None), 'x')
    torch.cuda.synchronize()
    mapping = []
    for i in range(64):
        d = (D - ref[i]).abs().amax(dim=1)
        j = int(torch.argmin(torch.nan_to_num(d, nan=1e9)))
        mapping.append((i, j, float(d[j])))
    print('a_mn', a_mn, 'off', off, 'row->lane', [(i, j) for i, j, e in mapping if e < 1e-3][::5])
    print('unmatched rows', [i for i, j, e in mapping if e >= 1e-3])
