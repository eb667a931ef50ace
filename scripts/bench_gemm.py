"""Time cg_gemm_f32 on the dense-head shapes of config C2 against torch (cuBLAS fp32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cnn_graph_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x = torch.randn(1024, 3968, device='cuda'); W = torch.randn(3968, 512, device='cuda'); g = torch.randn(1024, 512, device='cuda')
for name, ours, ref, fl in (
        ('fwd  x W      ', lambda: ops.gemm(x, W), lambda: x @ W, 2 * 1024 * 3968 * 512),
        ('dx   g W^T    ', lambda: ops.gemm(g, W, transB=True), lambda: g @ W.t(), 2 * 1024 * 3968 * 512),
        ('dW   x^T g    ', lambda: ops.gemm(x, g, transA=True), lambda: x.t() @ g, 2 * 1024 * 3968 * 512)):
    a, b = t(ours), t(ref)
    print('%s ours %.3f ms (%.0f TFLOP/s fp32-equivalent)   torch fp32 %.3f ms' % (name, a, fl / a / 1e9, b))
