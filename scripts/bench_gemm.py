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
# the gate-filter shapes of config C4 (R = 50 * 1024 rows, K*Fin = 384, 4H = 512)
R = 51200
a = torch.randn(R, 384, device='cuda'); w = torch.randn(384, 512, device='cuda'); gy = torch.randn(R, 512, device='cuda')
w2 = torch.randn(128, 1536, device='cuda'); z = torch.randn(R, 1536, device='cuda'); a1 = torch.randn(R, 128, device='cuda')
for name, ours, ref, fl in (
        ('C4 fwd  [R,384] [384,512]    ', lambda: ops.gemm(a, w), lambda: a @ w, 2 * R * 384 * 512),
        ('C4 dx   [R,1536] [128,1536]^T', lambda: ops.gemm(z, w2, transB=True), lambda: z @ w2.t(), 2 * R * 1536 * 128),
        ('C4 dW_k [R,128]^T [R,512]    ', lambda: ops.gemm(a1, gy, transA=True), lambda: a1.t() @ gy, 2 * R * 128 * 512)):
    t0, t1 = t(ours), t(ref)
    print('%s ours %.3f ms (%.0f TFLOP/s fp32-equivalent)   torch fp32 %.3f ms' % (name, t0, fl / t0 / 1e9, t1))
