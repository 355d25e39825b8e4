"""tcgen05 building block: the TMEM-A / smem-B UMMA with the operand layouts of the fused kernel."""
import pytest
import torch

from pesto_b200 import _lib

pytestmark = pytest.mark.gpu


def probe(A, B, split, lbo=-1, sbo=-1, idesc=0):
    lib = _lib.load()
    K, N = A.shape[1], B.shape[0]
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(lib.pesto_debug_umma_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, int(split), lbo, sbo, idesc, None), "probe")
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("K,N", [(32, 128), (32, 32), (64, 64), (32, 16), (128, 128), (96, 48)])
def test_umma_matches_matmul(K, N):
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = (torch.randn(128, K, generator=g) * 3).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    ref64 = (A.double() @ B.double().T)
    # single pass: operands rounded to bf16, fp32 accumulation
    D1 = probe(A, B, 0)
    ref1 = (A.bfloat16().double() @ B.bfloat16().double().T)
    assert (D1.double() - ref1).abs().max().item() < 1e-3 * ref1.abs().max().item()
    # 3-term split: close to the fp32 product
    D3 = probe(A, B, 1)
    assert (D3.double() - ref64).abs().max().item() < 2e-5 * ref64.abs().max().item()
