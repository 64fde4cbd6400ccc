"""GPU: tcgen05 attention (bf16 in/out) against the fp32 SIMT kernel / oracle on the same bf16-rounded
inputs.  Stated tolerance: 2e-2 of the output's max (P and V are rounded to bf16 before the PV MMA)."""
import pytest
import torch

from cassnat_asr_public_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()


def ref_attention(q, k, v, mask, h):
    B, Tq, d = q.shape
    dk = d // h
    qh, kh, vh = (t.double().view(B, -1, h, dk).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) / dk ** 0.5
    if mask is not None:
        m = mask if mask.dim() == 3 else mask.unsqueeze(1)
        s = s.masked_fill(m.unsqueeze(1) == 0, float(torch.finfo(torch.float32).min))
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Tq, d)


@pytest.mark.parametrize("B,Tq,Tk,h", [(2, 128, 128, 1), (3, 100, 77, 2), (2, 390, 500, 8), (1, 7, 512, 4), (2, 257, 129, 2)])
def test_attention_tc(B, Tq, Tk, h):
    d = 64 * h
    torch.manual_seed(Tq * 7 + Tk)
    q, k, v = (torch.randn(B, n, d).bfloat16() for n in (Tq, Tk, Tk))
    lens = torch.randint(Tk // 2, Tk + 1, (B,))
    key_mask = (torch.arange(Tk)[None] < lens[:, None]).unsqueeze(1)
    full = torch.rand(B, Tq, Tk) < 0.3
    full[:, Tq // 2:] = False                                  # fully masked rows -> uniform over all keys
    for mask in (None, key_mask, full):
        got = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), mask.to(DEV) if mask is not None else None, h)
        assert got.dtype == torch.bfloat16
        want = ref_attention(q.float(), k.float(), v.float(), mask, h)
        assert rel(got.float(), want) < 2e-2, (None if mask is None else tuple(mask.shape))


def test_attention_tc_on_fused_buffers():
    """q/k/v as strided views of fused projection outputs, like the model path."""
    B, T, h = 2, 200, 4
    d = 64 * h
    torch.manual_seed(1)
    qkv = torch.randn(B, T, 3 * d).bfloat16().to(DEV)
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    got = ops.attention(q, k, v, None, h)
    want = ref_attention(q.float().cpu(), k.float().cpu(), v.float().cpu(), None, h)
    assert rel(got.float(), want) < 2e-2
