"""GPU: the tcgen05/TMA bf16 GEMM against an fp64 product of the SAME bf16 operands (so only the
fp32 accumulation order differs: tolerance 2e-5 of the output's max)."""
import pytest
import torch

from cassnat_asr_public_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 512), (300, 512, 512), (1000, 5000, 512),
                                   (77, 120, 128), (4096, 2048, 512), (513, 1536, 2048), (32000, 512, 512),
                                   (1, 8, 8), (250, 264, 72)])
def test_gemm_bf16_tc(M, N, K):
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K).bfloat16()
    w = (torch.randn(N, K) / K ** 0.5).bfloat16()
    b = torch.randn(N)
    r = torch.randn(M, N)
    ref = x.double() @ w.double().t()
    y = ops.linear_bf16(x.to(DEV), w.to(DEV))
    assert rel(y, ref) < 2e-5
    y = ops.linear_bf16(x.to(DEV), w.to(DEV), b.to(DEV), r.to(DEV), 0.5, "relu")
    assert rel(y, 0.5 * torch.relu(ref + b.double()) + r.double()) < 2e-5
    y32, y16 = ops.linear_bf16(x.to(DEV), w.to(DEV), b.to(DEV), None, 1.0, "swish", also_bf16=True)
    z = ref + b.double()
    assert rel(y32, z * torch.sigmoid(z)) < 1e-4
    assert y16.dtype == torch.bfloat16 and rel(y16.float(), y32) < 1e-2


def test_gemm_bf16_strided_views_and_res_rows():
    torch.manual_seed(0)
    M, N, K = 384, 512, 256
    big = torch.randn(M, 3 * K).bfloat16().to(DEV)
    x = big[:, K:2 * K]                               # row stride 3K, 16-byte aligned offset
    w = (torch.randn(N, K) / 16).bfloat16().to(DEV)
    pe = torch.randn(96, N).to(DEV)
    y = ops.linear_bf16(x, w, None, pe, 1.0, None, res_rows=96)
    ref = x.double().cpu() @ w.double().cpu().t() + pe.double().cpu().repeat(4, 1)
    assert rel(y, ref) < 2e-5


def test_cast_bf16():
    x = torch.randn(1000, 512, device=DEV)
    assert torch.equal(ops.cast_bf16(x), x.bfloat16())
