"""tcgen05 primitive self-test (cg_debug_umma_gemm): descriptor encodings for K-major and
MN-major shared-memory operands, TMEM accumulation and read-back, against a torch fp32
reference of the same bf16-rounded product."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize('N,Kd', [(64, 32), (32, 64), (16, 16), (256, 128), (64, 256)])
def test_umma_gemm(a_mn, b_mn, N, Kd):
    from cnn_graph_b200 import _native
    lib = _native.lib()
    torch.manual_seed(N * 1000 + Kd)
    A = torch.randn(128, Kd, device='cuda')
    B = torch.randn(N, Kd, device='cuda')
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    A_src = A.t().contiguous() if a_mn else A.contiguous()
    B_src = B.t().contiguous() if b_mn else B.contiguous()
    D = torch.full((128, N), float('nan'), device='cuda')
    _native.check(lib.cg_debug_umma_gemm(A_src.data_ptr(), B_src.data_ptr(), D.data_ptr(), N, Kd, a_mn, b_mn,
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'cg_debug_umma_gemm')
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize('N,Kd', [(32, 64), (64, 16), (16, 128), (256, 32)])
def test_umma_gemm_a_in_tmem(N, Kd):
    """A operand written to tensor memory with tcgen05.st, B K-major in shared memory."""
    from cnn_graph_b200 import _native
    lib = _native.lib()
    torch.manual_seed(N + Kd)
    A = torch.randn(128, Kd, device='cuda')
    B = torch.randn(N, Kd, device='cuda')
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    D = torch.full((128, N), float('nan'), device='cuda')
    _native.check(lib.cg_debug_umma_gemm_ts(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, Kd,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  'cg_debug_umma_gemm_ts')
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item(), (err, ref.abs().max().item())
