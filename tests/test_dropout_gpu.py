"""GPU: training-mode dropout (the reference trains with dropout 0.1: egs/librispeech/conf/cassnat_train.yaml:11).
The kernels' keep mask is a stateless function of (seed, site, element index); `ops.dropout_mask` returns the very mask
a site used, so every dropout site is compared with plain torch arithmetic under the SAME mask, forward and backward:
  residual branch     modules/utils.py:32          FFN hidden     modules/positionff.py:16
  attention probs     modules/attention.py:22-23   relative attn  modules/attention.py:140-141
  positional enc.     modules/embedding.py:31,60"""
import math

import pytest
import torch

from cassnat_asr_public_b200 import modules as MD
from cassnat_asr_public_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()


def test_mask_statistics_and_determinism():
    ops.dropout_begin_step(seed=123)
    m1 = ops.dropout_mask((4096, 512), 0.1, 7, DEV)
    m2 = ops.dropout_mask((4096, 512), 0.1, 7, DEV)
    m3 = ops.dropout_mask((4096, 512), 0.1, 8, DEV)
    assert torch.equal(m1, m2)
    assert not torch.equal(m1, m3)
    keep = (m1 != 0).float().mean().item()
    assert abs(keep - 0.9) < 2e-3
    assert torch.all((m1 == 0) | ((m1 - 1 / 0.9).abs() < 1e-6))
    # independent of the neighbouring element / row (no visible structure): column and row means are all near 0.9
    assert ((m1 != 0).float().mean(0) - 0.9).abs().max().item() < 0.03
    ops.dropout_begin_step(seed=124)
    assert not torch.equal(m1, ops.dropout_mask((4096, 512), 0.1, 7, DEV))
    # device-side seed overrides the host seed (graph-replayed steps advance it in-graph)
    ops.DROPOUT_SEED_DEV = torch.tensor([123], dtype=torch.int64, device=DEV)
    try:
        assert torch.equal(m1, ops.dropout_mask((4096, 512), 0.1, 7, DEV))
    finally:
        ops.DROPOUT_SEED_DEV = None
        ops.dropout_begin_step(seed=0)


def test_feed_forward_and_residual_with_dropout():
    torch.manual_seed(0)
    d, dff, p = 128, 256, 0.1
    ff = MD.PositionwiseFeedForward(d, dff, p).to(DEV).train()
    sub = MD.SublayerConnection(d, p).to(DEV).train()
    x = torch.randn(3, 17, d, device=DEV, requires_grad=True)
    ops.dropout_begin_step(seed=5)
    y = sub(x, ff, scale=0.5)                    # site 1: FFN hidden, site 2: residual branch
    g = torch.randn_like(y)
    y.backward(g)
    m_hid = ops.dropout_mask((3, 17, dff), p, 1, DEV)
    m_res = ops.dropout_mask((3, 17, d), p, 2, DEV)
    xr = x.detach().clone().requires_grad_(True)
    mu, sd = xr.mean(-1, keepdim=True), xr.std(-1, keepdim=True)
    n = sub.norm.a_2 * (xr - mu) / (sd + sub.norm.eps) + sub.norm.b_2
    h = torch.relu(torch.nn.functional.linear(n, ff.w_1.weight, ff.w_1.bias)) * m_hid
    want = xr + 0.5 * torch.nn.functional.linear(h, ff.w_2.weight, ff.w_2.bias) * m_res
    want.backward(g)
    assert rel(y, want) < 1e-4 and rel(x.grad, xr.grad) < 1e-3


@pytest.mark.parametrize("relative", [False, True])
def test_attention_probability_dropout(relative):
    torch.manual_seed(1)
    B, L, h, d, p = 2, 21, 2, 128, 0.2
    dk = d // h
    q, k, v = (torch.randn(B, L, d, device=DEV, requires_grad=True) for _ in range(3))
    mask = torch.ones(B, 1, L, dtype=torch.uint8, device=DEV)
    mask[1, :, 15:] = 0
    relp = None
    if relative:
        m = 4
        table = torch.randn(2 * m + 1, d, device=DEV) * 0.3
        bu, bv = torch.randn(h, dk, device=DEV) * 0.3, torch.randn(h, dk, device=DEV) * 0.3
        relp = (table, bu, bv, m)
    ops.dropout_begin_step(seed=9)
    out = ops.attention(q, k, v, mask, h, relp, p_drop=p)        # site 1
    g = torch.randn_like(out)
    out.backward(g)
    pm = ops.dropout_mask((B, h, L, L), p, 1, DEV)
    qr, kr, vr = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    hv = lambda t: t.view(B, L, h, dk).transpose(1, 2)
    if relative:
        qu, qv = hv(qr) + bu[None, :, None, :], hv(qr) + bv[None, :, None, :]
        ac = qu @ hv(kr).transpose(-2, -1)
        idx = (torch.arange(L, device=DEV)[None, :] - torch.arange(L, device=DEV)[:, None]).clamp(-m, m) + m     # [i, j] -> j - i
        P = table.view(2 * m + 1, h, dk)                          # [r, h, dk]
        bd = torch.einsum("bhid,rhd->bhir", qv, P).gather(-1, idx[None, None].expand(B, h, L, L))
        s = (ac + bd) / math.sqrt(dk)
    else:
        s = hv(qr) @ hv(kr).transpose(-2, -1) / math.sqrt(dk)
    masked = (mask.unsqueeze(1) == 0)
    s = s.masked_fill(masked, torch.finfo(torch.float32).min)
    pr = torch.softmax(s, -1)
    if relative:
        pr = pr.masked_fill(masked, 0.0)
    want = ((pr * pm) @ hv(vr)).transpose(1, 2).reshape(B, L, d)
    want.backward(g)
    assert rel(out, want) < 1e-4
    for a, b in ((q, qr), (k, kr), (v, vr)):
        assert rel(a.grad, b.grad) < 1e-3


def test_positional_encodings_with_dropout():
    p = 0.1
    x = torch.randn(2, 30, 128, device=DEV)
    pe = MD.PositionalEncoding(128, p).to(DEV).train()
    ops.dropout_begin_step(seed=3)
    y = pe(x)
    assert rel(y, (x + pe.pe[:, :30]) * ops.dropout_mask(x.shape, p, 1, DEV)) < 1e-6
    rp = MD.RelativePositionalEncoding(128, p, 8).to(DEV).train()
    ops.dropout_begin_step(seed=3)
    yx, pos = rp(x)
    assert pos.shape == (59, 128) and getattr(pos, "_cn_table", None) is None
    assert rel(yx, x * ops.dropout_mask(x.shape, p, 1, DEV)) < 1e-6
    idx = torch.arange(-29, 30, device=DEV).clamp(-8, 8) + 8
    assert rel(pos, rp.embedding.weight[idx] * ops.dropout_mask((59, 128), p, 2, DEV)) < 1e-6
    assert rp.eval()(x)[0] is x                                  # eval mode: identity, table shortcut back


@pytest.mark.parametrize("family", ["trf", "con"])
def test_training_step_runs_with_reference_dropout(family):
    """The reference's training configuration (dropout 0.1) end to end: finite losses, every parameter gets a gradient,
    the step is reproducible under the same seed and differs under another one and from the dropout-free step."""
    from types import SimpleNamespace

    import bench
    from cassnat_asr_public_b200.cassnat import make_model
    w = dict(d_model=128, n_head=2, d_ff=256, N_enc=2, N_mix_dec=2)
    over = dict(vocab_size=60, dropout=0.1)
    if family == "con":
        over.update(use_conv_enc=True, use_conv_dec=True, pos_type="relative")
    args = bench.model_args(w, **over)
    torch.manual_seed(4)
    model = make_model(16, args).to(DEV).train()
    B, Tin = 3, 64
    src = torch.randn(B, Tin, 16, device=DEV)
    mask = torch.ones(B, 1, Tin, dtype=torch.bool, device=DEV)
    labels = torch.randint(4, 60, (B, 7), device=DEV)
    labels[:, -1] = 2
    sizes = torch.full((B,), 6, dtype=torch.int64, device=DEV)
    ratio = torch.ones(B, device=DEV)

    def step(seed):
        model.zero_grad(set_to_none=True)
        ops.dropout_begin_step(seed=seed)
        out = model(src, mask, ratio, labels, sizes, args)
        out[2].backward()
        return float(out[2]), torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])

    l1, g1 = step(11)
    l2, g2 = step(11)
    l3, _ = step(12)
    assert math.isfinite(l1) and torch.isfinite(g1).all()
    assert l1 == l2 and rel(g1, g2) < 1e-5          # (the cuDNN front-end backward is not bitwise reproducible)
    assert l1 != l3
    missing = [n for n, p in model.named_parameters() if p.grad is None and p.requires_grad]      # (frozen: the
    assert not missing, missing                                   # relative-position embedding tables, from_pretrained)
    model.zero_dropout()
    l0, _ = step(11)
    assert l0 != l1
