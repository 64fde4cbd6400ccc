"""GPU: the fast-mode encoder front-end (direct conv1 + implicit-GEMM conv2 on tcgen05 with 4-D strided TMA
boxes, channels-last bf16) against torch's fp32 convolutions (src/models/modules/embedding.py:103-119).
Stated bf16 tolerance: 2e-2 of the output's max."""

import pytest
import torch

from cassnat_asr_public_b200 import modules as MD
from cassnat_asr_public_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()


@pytest.mark.parametrize("B,Tin,Fin,C", [(2, 64, 16, 256), (3, 203, 80, 256), (1, 7, 5, 256), (2, 401, 83, 512)])
def test_conv_frontend_vs_torch(B, Tin, Fin, C):
    torch.manual_seed(Tin + Fin)
    x = torch.randn(B, Tin, Fin, device=DEV)
    x[-1, Tin // 2:] = 0                                   # padded tail of a shorter utterance
    conv = torch.nn.Sequential(torch.nn.Conv2d(1, C, 3, 2, 1), torch.nn.ReLU(), torch.nn.Conv2d(C, C, 3, 2, 1),
                               torch.nn.ReLU()).to(DEV)
    with torch.no_grad():
        conv[2].weight.mul_(3.0)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            want = conv(x.unsqueeze(1))                        # [B, C, T2, F2]
        got = ops.conv_frontend(x, conv[0].weight, conv[0].bias, conv[2].weight, conv[2].bias)   # [B, T2, F2, C]
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == (B, want.shape[2], want.shape[3], C)
    assert rel(got.float().permute(0, 3, 1, 2), want) < 2e-2


def test_conv_embedding_fast_mode_matches_parity_mode():
    """Whole ConvEmbedding (convs, output projection with the permuted weight, sqrt(d) scale, positional
    encoding, mask subsampling) in fast mode vs the fp32 parity mode."""
    torch.manual_seed(5)
    d, Fin = 256, 80
    emb = MD.ConvEmbedding(Fin, d, 0.0, MD.PositionalEncoding(d, 0.0)).to(DEV).eval()
    x = torch.randn(3, 301, Fin, device=DEV)
    mask = torch.ones(3, 1, 301, dtype=torch.bool, device=DEV)
    mask[2, :, 200:] = False
    with torch.no_grad():
        ops.set_precision("fp32")
        want, m0 = emb(x, mask)
        ops.set_precision("bf16")
        try:
            got, m1 = emb(x, mask)
        finally:
            ops.set_precision("fp32")
    assert torch.equal(m0, m1) and got.shape == want.shape
    assert rel(got.float(), want) < 2e-2
