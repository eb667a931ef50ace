"""One shape of cg_gemm_f32 a few times (for ncu): python scripts/prof_gemm.py [fwd|dx|dw|fc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cnn_graph_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
R = 51200
torch.manual_seed(0)
if which == 'fwd':
    a, b, kw = torch.randn(R, 384, device='cuda'), torch.randn(384, 512, device='cuda'), {}
elif which == 'dx':
    a, b, kw = torch.randn(R, 1536, device='cuda'), torch.randn(128, 1536, device='cuda'), {'transB': True}
elif which == 'dw':
    a, b, kw = torch.randn(R, 128, device='cuda'), torch.randn(R, 512, device='cuda'), {'transA': True}
else:
    a, b, kw = torch.randn(1024, 3968, device='cuda'), torch.randn(3968, 512, device='cuda'), {}
for _ in range(3):
    c = ops.gemm(a, b, **kw)
torch.cuda.synchronize()
print(which, float(c.abs().max()))
