"""Conformer convolution-module core (GLU -> depthwise conv -> GroupNorm(1,C) -> activation,
/root/reference/src/models/modules/conformer_related.py:33-38) against the same sequence of torch ops in fp64:
the fp32 path, and the fast mode's bf16-in / bf16-out path that evaluates the convolution twice instead of storing it."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def reference(y, w, b, gamma, beta, eps, act):
    z = F.glu(y.double(), dim=-1).transpose(1, 2)                                   # [B,C,L]
    C, k = w.shape[0], w.shape[-1]
    z = F.conv1d(z, w.double(), b.double(), padding=(k - 1) // 2, groups=C)
    z = F.group_norm(z, 1, gamma.double(), beta.double(), eps)
    z = z * torch.sigmoid(z) if act == "swish" else torch.relu(z)
    return z.transpose(1, 2)


@pytest.mark.parametrize("act", ["swish", "relu"])
@pytest.mark.parametrize("B,L,C,k", [(3, 77, 128, 3), (2, 500, 512, 3), (2, 130, 256, 7), (2, 200, 128, 15),
                                     (2, 300, 256, 31), (1, 5, 64, 3), (2, 90, 96, 5)])
@pytest.mark.parametrize("mode", ["f32", "bf16"])
def test_conv_module_core(B, L, C, k, act, mode):
    from cassnat_asr_public_b200 import ops
    torch.manual_seed(L + k)
    wide = torch.randn(B, L, 2 * C + 8)
    w, b = torch.randn(C, 1, k) * 0.5, torch.randn(C) * 0.1
    gamma, beta = 1 + 0.2 * torch.randn(C), 0.1 * torch.randn(C)
    if mode == "bf16":
        wide = wide.bfloat16()
    y = wide[..., :2 * C]                                # a strided view, like a slice of a fused buffer
    want = reference(y.float(), w, b, gamma, beta, 1e-5, act)
    got = ops.glu_dwconv_groupnorm_act(wide.to(DEV)[..., :2 * C], w.to(DEV), b.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-5, act,
                                       out_bf16=(mode == "bf16"))
    assert got.dtype == (torch.bfloat16 if mode == "bf16" else torch.float32)
    err = (got.double().cpu() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= (1e-2 if mode == "bf16" else 2e-5) * scale, (err, scale)
