"""Per-stage clock64 stamps of k_gemm_pipe (CTA 0): converter warp 0 and the issue warp."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cnn_graph_b200 import _native, ops
which = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
R = 51200
if which == 'fwd':
    a, b, kw = torch.randn(R, 384, device='cuda'), torch.randn(384, 512, device='cuda'), {}
elif which == 'dx':
    a, b, kw = torch.randn(R, 1536, device='cuda'), torch.randn(128, 1536, device='cuda'), {'transB': True}
else:
    a, b, kw = torch.randn(R, 128, device='cuda'), torch.randn(R, 512, device='cuda'), {'transA': True}
fn = ctypes.CDLL(_native.LIB_PATH).cg_debug_gemm_trace
fn.argtypes = [ctypes.c_void_p]
for _ in range(2): ops.gemm(a, b, **kw)
buf = torch.zeros(32 * 8, dtype=torch.int64, device='cuda')
fn(buf.data_ptr())
ops.gemm(a, b, **kw)
torch.cuda.synchronize()
fn(None)
t = buf.cpu().view(32, 8)
t0 = int(t[0, 0])
print('stage: conv_start  empty_ok  stored  arrived | mma_wait  full_ok  issued   (cycles from first stamp)')
for i in range(32):
    print('%3d: ' % (i + 8) + ' '.join('%8d' % (int(v) - t0) for v in t[i, :7]))
