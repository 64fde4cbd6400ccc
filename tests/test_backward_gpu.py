"""Backward pass of the training call on the GPU kernels (grad.py + csrc/backward.cu).

* each autograd Function against torch autograd of the same fp32 expression (reference formulas:
  modules/norm.py:15-18, modules/attention.py:13-24, modules/positionff.py:15-16, utils/loss.py:21-35);
* the whole `loss.backward()` of CassNAT.forward against the parameter gradients of the unmodified
  reference (tests/golden/train_grads_*.npz): rel 1e-3 of each tensor's max in the fp32 mode,
  rel 5e-2 in the bf16 tensor-core mode (stated bf16 tolerance for gradients)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from cassnat_asr_public_b200 import grad, ops
from cassnat_asr_public_b200.cassnat import make_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()


@pytest.fixture(autouse=True)
def _fp32_mode():
    ops.set_precision("fp32")
    yield
    ops.set_precision("fp32")


def leaf(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).requires_grad_(True)


def ref_layer_norm(x, a, b, eps=1e-6):
    return a * (x - x.mean(-1, keepdim=True)) / (x.std(-1, keepdim=True) + eps) + b


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("M,N,K,act,use_res,alpha", [(70, 48, 40, None, False, 1.0), (300, 64, 96, "relu", False, 1.0),
                                                     (129, 40, 72, None, True, 0.5), (1000, 136, 64, None, True, 1.0)])
def test_linear_backward(precision, tol, M, N, K, act, use_res, alpha):
    ops.set_precision(precision)

    x, w, b = leaf(M, K, seed=1), leaf(N, K, seed=2, scale=0.2), leaf(N, seed=3)
    res = leaf(M, N, seed=4) if use_res else None
    y = ops.linear(x, w, b, residual=res, alpha=alpha, act=act)
    gy = torch.randn(M, N, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    y.backward(gy)
    got = [t.grad.clone() for t in (x, w, b)] + ([res.grad.clone()] if use_res else [])
    for t in (x, w, b, res):
        if t is not None:
            t.grad = None
    # bf16 mode: the reference sees the same bf16-rounded operands, so the ReLU mask is the same
    rnd = (lambda t: t + (t.detach().bfloat16().float() - t.detach())) if precision == "bf16" else (lambda t: t)
    z = torch.nn.functional.linear(rnd(x), rnd(w), b)
    if act == "relu":
        z = torch.relu(z)
    z = alpha * z + (res if use_res else 0)
    assert rel(y, z) < tol
    z.backward(gy)
    want = [t.grad for t in (x, w, b)] + ([res.grad] if use_res else [])
    for g, wnt, nm in zip(got, want, "x w b res".split()):
        assert rel(g, wnt) < tol, nm


@pytest.mark.parametrize("rows,d", [(7, 16), (100, 64), (257, 512), (33, 130)])
def test_layernorm_backward(rows, d):
    x, a, b = leaf(rows, d, seed=1), leaf(d, seed=2), leaf(d, seed=3)
    y = ops.layer_norm(x, a, b)
    gy = torch.randn(rows, d, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    y.backward(gy)
    got = [t.grad.clone() for t in (x, a, b)]
    for t in (x, a, b):
        t.grad = None
    z = ref_layer_norm(x, a, b)
    assert rel(y, z) < 1e-5
    z.backward(gy)
    for g, t, nm in zip(got, (x, a, b), "x a b".split()):
        assert rel(g, t.grad) < 2e-5, nm


def ref_attention(q, k, v, mask, h):
    B, Tq, d = q.shape
    dk = d // h
    sp = lambda t: t.reshape(B, -1, h, dk).transpose(1, 2)
    s = sp(q) @ sp(k).transpose(-1, -2) / dk ** 0.5
    if mask is not None:
        s = s.masked_fill(mask.unsqueeze(1) == 0, torch.finfo(torch.float32).min)
    return (torch.softmax(s, -1) @ sp(v)).transpose(1, 2).reshape(B, Tq, d)


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("B,h,Tq,Tk,dk,kind", [(2, 2, 5, 9, 64, "key"), (3, 4, 33, 33, 64, "full"), (2, 8, 70, 130, 64, "key"),
                                               (2, 8, 64, 64, 64, "none"), (2, 2, 41, 250, 64, "full"),
                                               (2, 2, 128, 200, 64, "full"), (2, 2, 150, 90, 64, "key")])
def test_attention_backward(precision, tol, B, h, Tq, Tk, dk, kind):
    """Includes fully masked rows (uniform probabilities, no score gradient) in the 'full' case.  Up to 128 query rows
    the backward is the fused single launch (cn_attention_bwd_small), beyond that the composed strided-GEMM path."""
    ops.set_precision(precision)
    d = h * dk
    dt = torch.bfloat16 if precision == "bf16" and dk == 64 else torch.float32
    qkv = leaf(B, max(Tq, Tk), 3 * d, seed=1)
    gen = torch.Generator(device=DEV).manual_seed(9)
    if kind == "key":
        lens = torch.randint(1, Tk + 1, (B,), device=DEV, generator=gen)
        mask = (torch.arange(Tk, device=DEV)[None] < lens[:, None]).unsqueeze(1)
    elif kind == "full":
        mask = torch.rand(B, Tq, Tk, device=DEV, generator=gen) > 0.4
        mask[:, 3] = False
    else:
        mask = None

    def run(fn, cast):
        t = qkv.to(dt) if cast else qkv
        q, k, v = t[:, :Tq, :d], t[:, :Tk, d:2 * d], t[:, :Tk, 2 * d:]
        return fn(q, k, v)
    out = run(lambda q, k, v: ops.attention(q, k, v, mask, h), True)
    go = torch.randn(B, Tq, d, device=DEV, generator=gen)
    out.backward(go.to(out.dtype))
    got = qkv.grad.clone()
    qkv.grad = None
    ref = run(lambda q, k, v: ref_attention(q, k, v, mask, h), False)
    assert rel(out.float(), ref) < tol
    ref.backward(go)
    assert rel(got, qkv.grad) < tol


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_attention_backward_fused_equals_composed(dt):
    """The fused small-Tq backward against the composed path (strided GEMMs + cn_attn_softmax_bwd) on the same
    inputs, with dropout on the probabilities: same masks by construction (seed, site)."""
    from cassnat_asr_public_b200 import grad
    ops.set_precision("fp32" if dt == torch.float32 else "bf16")
    B, h, Tq, Tk = 3, 4, 41, 250
    d = 64 * h
    gen = torch.Generator(device=DEV).manual_seed(5)
    mask = torch.rand(B, Tq, Tk, device=DEV, generator=gen) > 0.3
    mask[:, 7] = False
    go = torch.randn(B, Tq, d, device=DEV, generator=gen).to(dt)
    res = {}
    for fused in (True, False):
        grad.FUSED_SMALL_ATTENTION = fused
        try:
            q, k, v = (torch.randn(B, n, d, device=DEV, generator=torch.Generator(device=DEV).manual_seed(s)).to(dt).requires_grad_(True)
                       for n, s in ((Tq, 1), (Tk, 2), (Tk, 3)))
            ops.dropout_begin_step(seed=77)
            out = grad.Attention.apply(q, k, v, mask, h, 0.1, 3)
            out.backward(go)
            res[fused] = (out.detach().float(), q.grad.float(), k.grad.float(), v.grad.float())
        finally:
            grad.FUSED_SMALL_ATTENTION = True
    tol = 2e-5 if dt == torch.float32 else 2e-2
    for a, b_, nm in zip(res[True], res[False], ("out", "dq", "dk", "dv")):
        assert rel(a, b_) < tol, (nm, rel(a, b_))


def ref_rel_attention(q, k, v, P, u, vb, mask, h, m):
    B, T, d = q.shape
    dk = d // h
    sp = lambda t: t.reshape(B, -1, h, dk).transpose(1, 2)
    qu, qv = sp(q) + u[None, :, None, :], sp(q) + vb[None, :, None, :]
    ac = qu @ sp(k).transpose(-1, -2)
    ar = torch.arange(T, device=q.device)
    idx = (ar[None, :] - ar[:, None]).clamp(-m, m) + m
    Ph = P.reshape(-1, h, dk).permute(1, 0, 2)
    bd = (qv @ Ph.transpose(-1, -2)[None]).gather(-1, idx[None, None].expand(B, h, T, T))
    s = (ac + bd) / dk ** 0.5
    if mask is not None:
        s = s.masked_fill(mask.unsqueeze(1) == 0, torch.finfo(torch.float32).min)
    p = torch.softmax(s, -1)
    if mask is not None:
        p = p.masked_fill(mask.unsqueeze(1) == 0, 0.0)
    return (p @ sp(v)).transpose(1, 2).reshape(B, T, d)


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("B,h,T,m,kind", [(2, 2, 9, 3, "key"), (2, 4, 40, 8, "full"), (3, 8, 70, 8, "key"), (2, 2, 5, 8, "none")])
def test_rel_attention_backward(precision, tol, B, h, T, m, kind):
    ops.set_precision(precision)
    dk = 64
    d = h * dk
    mm = min(m, T - 1)
    qkv, P = leaf(B, T, 3 * d, seed=1), leaf(2 * mm + 1, d, seed=2)
    u, vb = leaf(h, dk, seed=3, scale=0.3), leaf(h, dk, seed=4, scale=0.3)
    gen = torch.Generator(device=DEV).manual_seed(9)
    if kind == "key":
        lens = torch.randint(1, T + 1, (B,), device=DEV, generator=gen)
        mask = (torch.arange(T, device=DEV)[None] < lens[:, None]).unsqueeze(1)
    elif kind == "full":
        mask = torch.rand(B, T, T, device=DEV, generator=gen) > 0.4
        mask[:, 3] = False
    else:
        mask = None
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    leaves = (qkv, P, u, vb)
    t = qkv.to(dt)
    out = ops.attention(t[..., :d], t[..., d:2 * d], t[..., 2 * d:], mask, h, (P, u, vb, mm))
    go = torch.randn(B, T, d, device=DEV, generator=gen)
    out.backward(go.to(out.dtype))
    got = [x.grad.clone() for x in leaves]
    for x in leaves:
        x.grad = None
    ref = ref_rel_attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], P, u, vb, mask, h, mm)
    assert rel(out.float(), ref) < tol
    ref.backward(go)
    for g, x, nm in zip(got, leaves, "qkv P u v".split()):
        assert rel(g, x.grad) < tol, nm


@pytest.mark.parametrize("B,L,C,ksz,act", [(2, 7, 16, 3, "swish"), (3, 70, 128, 7, "swish"), (2, 33, 300, 15, "relu")])
def test_conv_core_backward(B, L, C, ksz, act):
    F = torch.nn.functional
    y, w, b = leaf(B, L, 2 * C, seed=1), leaf(C, 1, ksz, seed=2, scale=0.4), leaf(C, seed=3, scale=0.2)
    gamma, beta = leaf(C, seed=4), leaf(C, seed=5, scale=0.3)
    leaves = (y, w, b, gamma, beta)
    out = ops.glu_dwconv_groupnorm_act(y, w, b, gamma, beta, 1e-5, act)
    go = torch.randn(B, L, C, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6))
    out.backward(go)
    got = [x.grad.clone() for x in leaves]
    for x in leaves:
        x.grad = None
    x = F.glu(y.transpose(1, 2), dim=1)
    x = F.group_norm(F.conv1d(x, w, b, padding=(ksz - 1) // 2, groups=C), 1, gamma, beta, 1e-5)
    x = (x * torch.sigmoid(x) if act == "swish" else torch.relu(x)).transpose(1, 2)
    assert rel(out, x) < 2e-5
    x.backward(go)
    for g, t, nm in zip(got, leaves, "y w b gamma beta".split()):
        assert rel(g, t.grad) < 1e-4, nm


def test_swish_linear_backward():
    x, w, b = leaf(200, 64, seed=1), leaf(96, 64, seed=2, scale=0.3), leaf(96, seed=3)
    y = ops.linear(x, w, b, act="swish")
    go = torch.randn(200, 96, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6))
    y.backward(go)
    got = [t.grad.clone() for t in (x, w, b)]
    for t in (x, w, b):
        t.grad = None
    z = torch.nn.functional.linear(x, w, b)
    z = z * torch.sigmoid(z)
    assert rel(y, z) < 2e-5
    z.backward(go)
    for g, t, nm in zip(got, (x, w, b), "x w b".split()):
        assert rel(g, t.grad) < 3e-5, nm


def test_label_smoothing_and_ctc_heads():
    R, V, B, T, L = 24, 50, 4, 30, 5
    gen = torch.Generator(device=DEV).manual_seed(3)
    logits = leaf(R, V, seed=1)
    target = torch.randint(1, V, (R,), device=DEV, generator=gen)
    target[::5] = 0
    logp, loss = grad.LabelSmoothingFromLogits.apply(logits, target, 0, 0.1)
    (loss * 0.7).backward()
    got = logits.grad.clone()
    logits.grad = None
    lp = torch.log_softmax(logits, -1)
    true = torch.full_like(lp, 0.1 / (V - 1))
    keep = target != 0
    true.scatter_(1, target.unsqueeze(1), 0.9)
    want = torch.nn.functional.kl_div(lp, true, reduction="none").masked_fill(~keep.unsqueeze(1), 0).sum() / keep.sum()
    assert rel(logp, lp) < 1e-5 and abs(float(loss.detach()) - float(want.detach())) < 1e-5 * abs(float(want.detach()))
    (want * 0.7).backward()
    assert rel(got, logits.grad) < 1e-4

    z = leaf(B, T, V, seed=2)
    labels = torch.randint(1, V, (B, L + 1), device=DEV, generator=gen)
    il = torch.tensor([30, 22, 17, 30], device=DEV)
    tl = torch.tensor([5, 3, 4, 2], device=DEV)
    logp, loss = grad.CtcFromLogits.apply(z, labels, il, tl, 0)
    (loss * 0.5).backward()
    got = z.grad.clone()
    z.grad = None
    want = torch.nn.functional.ctc_loss(torch.log_softmax(z, -1).transpose(0, 1), labels, il, tl, blank=0, reduction="mean",
                                        zero_infinity=True)
    assert abs(float(loss) - float(want)) < 1e-4 * abs(float(want))
    (want * 0.5).backward()
    assert rel(got, z.grad) < 1e-3


def _load(name):
    import json
    z = np.load(os.path.join(GOLDEN, f"blocks_{name}.npz"))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    cfg = SimpleNamespace(**json.loads(bytes(z["cfg_json"]).decode()))
    return z, sd, cfg, np.load(os.path.join(GOLDEN, f"train_grads_{name}.npz"))


@pytest.mark.parametrize("precision,tol,loss_tol", [("fp32", 1e-3, 1e-4), ("bf16", 0.25, 2e-2)])
@pytest.mark.parametrize("name", ["trf", "con"])
def test_whole_model_backward_vs_reference(name, precision, tol, loss_tol):
    """CassNAT.forward(...)[2].backward() == the reference's parameter gradients, encoder included."""
    z, sd, cfg, g = _load(name)
    args = SimpleNamespace(**vars(cfg))
    model = make_model(16, args)
    model.load_state_dict(sd, strict=False)
    model = model.to(DEV).train().zero_dropout()
    model.set_precision(precision)
    src, mask, size = (torch.from_numpy(z[k]).to(DEV) for k in ("src", "src_mask", "src_size"))
    labels, sizes = torch.from_numpy(z["train::labels"]).to(DEV), torch.from_numpy(z["train::label_sizes"]).to(DEV)
    ctc_out, att_out, loss, ctc_loss, att_loss = model(src, mask, size, labels, sizes, args)
    assert loss.requires_grad and not ctc_out.requires_grad
    for got, key in ((loss, "loss"), (ctc_loss, "ctc_loss"), (att_loss, "att_loss")):
        assert abs(float(got) - float(g[key])) < loss_tol * max(1.0, abs(float(g[key]))), key
    loss.backward()
    params = dict(model.named_parameters())
    worst = {}
    for key in g.files:
        if not key.startswith("grad::"):
            continue
        p = params[key[6:]]
        assert p.grad is not None, key
        if np.abs(g[key]).max() < 1e-7:          # analytically zero (key biases: softmax is shift invariant)
            worst[key] = float(p.grad.abs().max()) / 0.1 * tol
        else:
            worst[key] = rel(p.grad, g[key])
    bad = {k: v for k, v in worst.items() if not v < tol}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:40]
    assert len(worst) > 100
    # bf16 mode: ReLU units near zero flip under operand rounding, which is visible per tensor on this
    # 24-token fixture; the direction of the whole gradient must still agree
    mine = torch.cat([params[k[6:]].grad.flatten().double().cpu() for k in g.files if k.startswith("grad::")])
    ref = torch.cat([torch.from_numpy(g[k]).flatten().double() for k in g.files if k.startswith("grad::")])
    cos = float(torch.dot(mine, ref) / (mine.norm() * ref.norm()))
    assert cos > (0.999999 if precision == "fp32" else 0.998), cos
    model.set_precision("fp32")


INTER_CASES = [("trf", "interctc_mixce"), ("trf", "unimask"), ("con", "interctc_selfce"), ("con", "mixce")]


@pytest.mark.parametrize("fam,case", INTER_CASES)
def test_training_call_intermediate_losses_and_unimask(fam, case):
    """CassNAT.forward with interctc / interce (self- and mix-decoder taps) and use_unimask against the unmodified
    reference (tests/golden/inter.npz): log-probabilities and losses rel 1e-4, gradients of the generator and
    LayerNorm parameters rel 1e-3 (fp32 mode)."""
    import json
    z = np.load(os.path.join(GOLDEN, "inter.npz"))
    pre = f"{fam}::{case}::"
    cfg = json.loads(bytes(z[pre + "cfg_json"]).decode())
    args = SimpleNamespace(**cfg)
    build_args = SimpleNamespace(**{**cfg, "interctc_alpha": 0.3, "interce_alpha": 0.2, "interce_layer": 3})
    model = make_model(16, build_args)                       # built with both intermediate generators, like the fixture
    sd = {k[len(fam) + 6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{fam}::sd::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("pos_enc.pe") for k in missing), (missing, unexpected)
    model = model.to(DEV).train().zero_dropout()
    src, mask, size = (torch.from_numpy(z[k]).to(DEV) for k in ("src", "src_mask", "src_size"))
    labels, sizes = torch.from_numpy(z["labels"]).to(DEV), torch.from_numpy(z["label_sizes"]).to(DEV)
    ctc_out, att_out, loss, ctc_loss, att_loss = model(src, mask, size, labels, sizes, args)
    assert rel(ctc_out, z[pre + "ctc_out"]) < 1e-4 and rel(att_out, z[pre + "att_out"]) < 1e-4
    want = z[pre + "losses"]
    for got, w in zip((loss, ctc_loss, att_loss), want):
        assert abs(float(got.detach()) - float(w)) < 1e-4 * max(1.0, abs(float(w)))
    loss.backward()
    params = dict(model.named_parameters())
    n = 0
    for key in z.files:
        if key.startswith(pre + "grad::"):
            g = z[key]
            if np.abs(g).max() < 1e-7:
                continue
            assert rel(params[key[len(pre) + 6:]].grad, g) < 1e-3, key
            n += 1
    assert n > 20
    # and the same values without a tape
    with torch.no_grad():
        out = model(src, mask, size, labels, sizes, args)
    assert abs(float(out[2]) - float(want[0])) < 1e-4 * max(1.0, abs(float(want[0])))


def test_fused_linear_cross_entropy():
    """grad.LinearCE (head + label-smoothing loss without logits) against the unfused pair
    ops.linear -> LabelSmoothingFromLogits on the same bf16 GEMM arithmetic, and against torch in fp64."""
    from cassnat_asr_public_b200 import grad
    ops.set_precision("bf16")
    R, K, V, pad, ls = 700, 256, 1000, 0, 0.1
    gen = torch.Generator(device=DEV).manual_seed(11)
    x0 = torch.randn(R, K, device=DEV, generator=gen)
    w0 = torch.randn(V, K, device=DEV, generator=gen) * 0.05
    b0 = torch.randn(V, device=DEV, generator=gen) * 0.1
    tgt = torch.randint(1, V, (R,), device=DEV, generator=gen)
    tgt[::7] = pad
    res = {}
    for name in ("fused", "plain"):
        x, w, b = (t.clone().requires_grad_(True) for t in (x0, w0, b0))
        if name == "fused":
            loss, ids = grad.LinearCE.apply(x, w, b, tgt, pad, ls)
        else:
            z = ops.linear(x, w, b)
            _, loss = grad.LabelSmoothingFromLogits.apply(z, tgt, pad, ls)
            ids = z.argmax(-1)
        (loss * 3.0).backward()
        res[name] = (loss.detach(), x.grad, w.grad, b.grad, ids)
    assert abs(float(res["fused"][0]) - float(res["plain"][0])) < 2e-5 * abs(float(res["plain"][0]))
    for a, b_, nm in zip(res["fused"][1:4], res["plain"][1:4], ("dx", "dw", "db")):
        assert rel(a, b_) < 1.5e-2, (nm, rel(a, b_))                 # dz is rounded to bf16 before the two GEMMs
    assert (res["fused"][4] == res["plain"][4]).float().mean() > 0.995
    # torch fp64 on the bf16-rounded operands
    xd, wd, bd = x0.bfloat16().double().requires_grad_(True), w0.bfloat16().double().requires_grad_(True), b0.double().requires_grad_(True)
    logp = torch.log_softmax(xd @ wd.T + bd, -1)
    keep = tgt != pad
    true = torch.full_like(logp, ls / (V - 1)).scatter_(1, tgt.unsqueeze(1), 1.0 - ls)
    want = (torch.nn.functional.kl_div(logp, true, reduction="none").sum(1) * keep).sum() / keep.sum()
    (want * 3.0).backward()
    assert abs(float(res["fused"][0]) - float(want)) < 1e-4 * abs(float(want))
    assert rel(res["fused"][1], xd.grad) < 1.5e-2 and rel(res["fused"][2], wd.grad) < 1.5e-2 and rel(res["fused"][3], bd.grad) < 1.5e-2


def test_training_call_with_fused_cross_entropy():
    """args.fused_ce (opt-in): the training call without [B,U,V] log-probs returns the same loss and parameter
    gradients as the default path (bf16 GEMM arithmetic in both), att_out is None, argmax ids in last_att_pred."""
    import bench
    from cassnat_asr_public_b200.cassnat import make_model
    w = dict(d_model=128, n_head=2, d_ff=256, N_enc=2, N_mix_dec=2)
    args = bench.model_args(w, vocab_size=64, dropout=0.0, precision="bf16")
    torch.manual_seed(4)
    model = make_model(16, args).to(DEV).train()
    model.set_precision("bf16")
    model.zero_dropout()
    B, Tin = 3, 64
    gen = torch.Generator(device=DEV).manual_seed(2)
    src = torch.randn(B, Tin, 16, device=DEV, generator=gen)
    mask = torch.ones(B, 1, Tin, dtype=torch.bool, device=DEV)
    labels = torch.randint(4, 64, (B, 7), device=DEV, generator=gen)
    labels[:, -1] = 2
    sizes = torch.full((B,), 6, dtype=torch.int64, device=DEV)
    ratio = torch.ones(B, device=DEV)
    res = {}
    for fused in (False, True):
        args.fused_ce = fused
        model.zero_grad(set_to_none=True)
        out = model(src, mask, ratio, labels, sizes, args)
        out[2].backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        res[fused] = (float(out[2]), float(out[4]), grads, out[1])
    args.fused_ce = False
    ops.set_precision("fp32")
    assert res[True][3] is None and res[False][3] is not None
    assert abs(res[True][1] - res[False][1]) < 1e-4 * abs(res[False][1])
    assert abs(res[True][0] - res[False][0]) < 1e-4 * abs(res[False][0])
    assert res[True][2].keys() == res[False][2].keys()
    for n, g in res[False][2].items():
        assert rel(res[True][2][n], g) < 3e-2, n
    pred = model.last_att_pred
    assert pred.shape == labels.shape and pred.dtype == torch.int64
    assert (pred == res[False][3].argmax(-1)).float().mean() > 0.9
