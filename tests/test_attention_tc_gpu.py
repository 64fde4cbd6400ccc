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


@pytest.mark.parametrize("B,Tq,Tk,h", [(2, 128, 128, 1), (3, 100, 77, 2), (2, 390, 500, 8), (1, 7, 512, 4), (2, 257, 129, 2),
                                       (2, 300, 875, 2), (1, 520, 2048, 1), (3, 40, 1300, 4)])
def test_attention_tc(B, Tq, Tk, h):
    d = 64 * h
    torch.manual_seed(Tq * 7 + Tk)
    q, k, v = (torch.randn(B, n, d).bfloat16() for n in (Tq, Tk, Tk))
    lens = torch.randint(Tk // 2, Tk + 1, (B,))
    key_mask = (torch.arange(Tk)[None] < lens[:, None]).unsqueeze(1)
    full = torch.rand(B, Tq, Tk) < 0.3
    full[:, Tq // 2:] = False                                  # fully masked rows -> uniform over all keys
    for mask in (None, key_mask) + ((full,) if Tk <= 512 else ()):      # per-row masks: bit-packed, Tk <= 512
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


@pytest.mark.parametrize("table", ["pre-kernel", "f32", "bf16"])
@pytest.mark.parametrize("B,U,h,m", [(2, 90, 2, 8), (3, 300, 4, 20), (2, 500, 8, 8), (2, 130, 2, 22), (1, 5, 1, 4),
                                     (2, 700, 2, 8), (1, 1100, 1, 20)])       # (> 512 keys: K/V tiles stream through the ring)
def test_rel_attention_tc(B, U, h, m, table):
    """tcgen05 relative-position attention vs the oracle's rel_mha core on bf16-rounded q/k/v; the relative term's
    table either from the pre-kernel or handed in (what the fused projection's extra columns hold), f32 or bf16."""
    d = 64 * h
    torch.manual_seed(U + m)
    q, k, v = (torch.randn(B, U, d).bfloat16() for _ in range(3))
    u_b, v_b = torch.randn(h, 64) * 0.3, torch.randn(h, 64) * 0.3
    relp = torch.randn(2 * m + 1, d)
    lens = torch.randint(U // 2, U + 1, (B,))
    mask = (torch.arange(U)[None] < lens[:, None]).unsqueeze(1)
    # reference (fp64): scores = ((q+u).k + (q+v).P[clamp(j-i)]) / 8, masked -> finfo.min, softmax, re-zero, @ v
    qh, kh, vh = (t.double().view(B, U, h, 64).transpose(1, 2) for t in (q, k, v))
    ac = (qh + u_b.double()[None, :, None, :]) @ kh.transpose(-1, -2)
    P = relp.double().view(2 * m + 1, h, 64).permute(1, 0, 2)                       # [h, 2m+1, 64]
    qp = (qh + v_b.double()[None, :, None, :]) @ P.transpose(-1, -2)[None]          # [B,h,U,2m+1]
    idx = (torch.arange(U)[None, :] - torch.arange(U)[:, None]).clamp(-m, m) + m    # [i,j]
    bd = torch.gather(qp, -1, idx[None, None].expand(B, h, U, U))
    s = (ac + bd) / 8.0
    mm = (mask.unsqueeze(1) == 0)
    s = s.masked_fill(mm, float(torch.finfo(torch.float32).min))
    want = (torch.softmax(s, -1).masked_fill(mm, 0.0) @ vh).transpose(1, 2).reshape(B, U, d)
    # kernel: Q operand = q + u (what project_qkv_rel produces)
    q_u = (q.float().view(B, U, h, 64) + u_b[None, None]).reshape(B, U, d).bfloat16()
    rel_args = (relp.to(DEV), u_b.to(DEV), v_b.to(DEV), m)
    if table != "pre-kernel":
        # (q + v).P for the bf16-rounded q, laid out [B,U,h*(2m+1)] inside a wider buffer (row pitch != width)
        tab = qp.float().permute(0, 2, 1, 3).reshape(B, U, h * (2 * m + 1))
        wide = torch.zeros(B, U, h * (2 * m + 1) + 24, dtype=torch.bfloat16 if table == "bf16" else torch.float32)
        wide[..., 8:8 + tab.size(-1)] = tab
        rel_args = (None, u_b.to(DEV), v_b.to(DEV), m, wide.to(DEV)[..., 8:8 + tab.size(-1)])
    got = ops.attention(q_u.to(DEV), k.to(DEV), v.to(DEV), mask.to(DEV), h, rel_args)
    assert got.dtype == torch.bfloat16
    # (q+u) is rounded to bf16 once more than in the reference: tolerance 3e-2 of the output's max
    assert rel(got.float(), want) < 3e-2


@pytest.mark.parametrize("Tq,Tk,klen,qlen", [(400, 500, [500, 130, 257, 1], [400, 0, 129, 300]),
                                             (600, 1500, [1500, 640, 513, 1100], [600, 10, 300, 513])])
def test_attention_tc_length_hints(Tq, Tk, klen, qlen):
    """kv_len / q_len hints: skipped key tiles carry exactly zero weight, unneeded rows are zero-filled,
    needed rows are unchanged (also when the K/V tiles stream through the ring, Tk > 512)."""
    B, h = 4, 2
    d = 64 * h
    torch.manual_seed(3)
    q, k, v = (torch.randn(B, n, d).bfloat16().to(DEV) for n in (Tq, Tk, Tk))
    klen = torch.tensor(klen)
    qlen = torch.tensor(qlen)
    mask = (torch.arange(Tk)[None] < klen[:, None]).unsqueeze(1).to(DEV)
    plain = ops.attention(q, k, v, mask, h).float()
    ops.attach_length_hints(mask, kv_len=klen.to(DEV), q_len=qlen.to(DEV))
    hinted = ops.attention(q, k, v, mask, h).float()
    for b in range(B):
        n = int(qlen[b])
        assert torch.equal(hinted[b, :n], plain[b, :n]), b
        # rows >= ceil(q_len/128)*128 are zero-filled; rows between q_len and the tile edge are computed
        edge = min(-(-n // 128) * 128, Tq)
        assert float(hinted[b, edge:].abs().max()) == 0.0 if edge < Tq else True


@pytest.mark.parametrize("left,right", [(0, 0), (1, 1), (1, 0)])
def test_segment_attention_tc_matches_fp32_segment_kernel(left, right):
    """Extractor attention on tcgen05 with the trigger mask generated bit-packed from segment ids
    (cn_trigger_mask_bits + cn_attention_bf16_bits, shared query block) == the fp32 segment-sparse kernel,
    and the bit mask itself == the dense byte mask of cn_align_to_mask (bit-exact)."""
    import numpy as np

    import oracle.align_np as A
    B, T, h, V = 4, 300, 2, 30
    d = 64 * h
    rng = np.random.default_rng(5)
    lp = rng.standard_normal((B, T, V)).astype(np.float32)
    lp[:, :, 0] += 2.0
    xl = np.array([T, 190, 161, 240])
    mask = (np.arange(T)[None] < xl[:, None])[:, None]
    al, yl, ym = A.best_path_align(lp, mask, 0)
    c = lambda t: t.to(DEV)
    al_t, yl_t, mk_t, xl_t = (c(torch.from_numpy(x)) for x in (al, yl, mask, xl))
    dense, _, ym1, seg = ops.align_to_mask(al_t, yl_t, ym, mk_t, xl_t, 0, left, right, True, want_mask=True, want_seg=True)
    U = ym1 + 3                                                   # extra fully masked rows
    torch.manual_seed(3)
    q_rows = torch.randn(U, d).bfloat16().to(DEV)
    kv = torch.randn(B, T, 2 * d).bfloat16().to(DEV)
    k, v = kv[..., :d], kv[..., d:]
    got = ops.segment_attention(q_rows, k, v, seg, yl_t, mk_t, xl_t, left, right, h)
    assert got.dtype == torch.bfloat16
    want = ops.segment_attention(q_rows.float(), k.float(), v.float(), seg, yl_t, mk_t, xl_t, left, right, h)
    assert rel(got.float(), want) < 2e-2
    # the packed bits against the dense mask
    from cassnat_asr_public_b200._lib import call, ptr, stream
    bits = torch.empty(B, U, 16, dtype=torch.int32, device=DEV)
    call("cn_trigger_mask_bits", ptr(seg), ptr(yl_t), ptr(mk_t.reshape(B, T).to(torch.uint8).contiguous()), ptr(xl_t), B, T, U,
         int(left > 0), int(right > 0), ptr(bits), stream())
    b = bits.cpu().numpy().view(np.uint32)
    unpacked = ((b[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(B, U, 512)[..., :T].astype(bool)
    want_mask = np.zeros((B, U, T), bool)
    want_mask[:, :ym1] = dense.cpu().numpy() != 0
    assert np.array_equal(unpacked, want_mask)
    # q_len hint: rows below the hint unchanged
    q_len = (yl_t + 2).clamp(max=U)
    hinted = ops.segment_attention(q_rows, k, v, seg, yl_t, mk_t, xl_t, left, right, h, q_len=q_len)
    for bi in range(B):
        n = int(q_len[bi])
        assert torch.equal(hinted[bi, :n], got[bi, :n])


def test_attention_tc_causal_hint():
    """Causal per-row mask (LM ranker): with the causal hint the key tiles above the diagonal are skipped and the
    result is unchanged."""
    B, U, h = 3, 500, 2
    d = 64 * h
    torch.manual_seed(8)
    q, k, v = (torch.randn(B, U, d).bfloat16().to(DEV) for _ in range(3))
    lens = torch.tensor([500, 257, 130])
    key = (torch.arange(U)[None] < lens[:, None]).unsqueeze(1)
    mask = (key & torch.tril(torch.ones(U, U, dtype=torch.bool)).unsqueeze(0)).to(DEV)
    plain = ops.attention(q, k, v, mask, h).float()
    hinted = ops.attention(q, k, v, ops.mark_causal(mask.clone()), h).float()
    valid = mask.any(-1)                                                   # rows with at least one key
    assert torch.equal(plain[valid], hinted[valid])
    want = ref_attention(q.float().cpu(), k.float().cpu(), v.float().cpu(), mask.cpu(), h)
    assert rel(plain.cpu()[valid.cpu()], want[valid.cpu()]) < 2e-2


def _split(t):
    hi = t.bfloat16()
    lo = (t - hi.float()).bfloat16()
    full = torch.cat([hi, lo], -1).to(DEV)
    n = t.shape[-1]
    out = full[..., :n]
    out._cn_lo = full[..., n:]
    return out


@pytest.mark.parametrize("B,Tq,Tk,h", [(2, 128, 128, 1), (3, 100, 77, 2), (2, 390, 500, 8), (1, 7, 512, 4), (2, 257, 129, 2),
                                       (2, 300, 875, 2), (1, 520, 2048, 1), (3, 40, 1300, 4)])
def test_attention_split_bf16(B, Tq, Tk, h):
    """cn_attention_bf16x3: (hi, lo) bf16 Q/K/V, three tensor-core passes for the scores and for the context, fp32 out.
    Error vs fp64 on the fp32 inputs is that of the split (2^-17 per operand), far below plain bf16's 2^-9."""
    d = 64 * h
    torch.manual_seed(Tq * 7 + Tk)
    q, k, v = (torch.randn(B, n, d) for n in (Tq, Tk, Tk))
    lens = torch.randint(Tk // 2, Tk + 1, (B,))
    key_mask = (torch.arange(Tk)[None] < lens[:, None]).unsqueeze(1)
    for mask in (None, key_mask):
        got = ops.attention(_split(q), _split(k), _split(v), mask.to(DEV) if mask is not None else None, h)
        assert got.dtype == torch.bfloat16 and ops.split_of(got) is not None      # hi half of a split operand
        got = ops.unsplit(got)
        want = ref_attention(q, k, v, mask, h)
        assert rel(got, want) < 5e-5, (None if mask is None else tuple(mask.shape), rel(got, want))


def test_attention_split_bf16_hints_and_fallback():
    """Length hints skip work without changing the needed rows; per-row masks fall back to the fp32 kernel."""
    B, Tq, Tk, h = 4, 400, 700, 2
    d = 64 * h
    torch.manual_seed(5)
    q, k, v = (torch.randn(B, n, d) for n in (Tq, Tk, Tk))
    klen = torch.tensor([700, 130, 257, 1])
    qlen = torch.tensor([400, 0, 129, 300])
    mask = (torch.arange(Tk)[None] < klen[:, None]).unsqueeze(1).to(DEV)
    plain = ops.unsplit(ops.attention(_split(q), _split(k), _split(v), mask, h))
    ops.attach_length_hints(mask, kv_len=klen.to(DEV), q_len=qlen.to(DEV))
    hinted = ops.unsplit(ops.attention(_split(q), _split(k), _split(v), mask, h))
    for b in range(B):
        n = int(qlen[b])
        assert torch.equal(hinted[b, :n], plain[b, :n]), b
        edge = min(-(-n // 128) * 128, Tq)
        assert edge == Tq or float(hinted[b, edge:].abs().max()) == 0.0
    full = (torch.rand(B, Tq, Tk) < 0.3)
    got = ops.attention(_split(q), _split(k), _split(v), full.to(DEV), h)
    assert rel(got, ref_attention(q, k, v, full, h)) < 1e-5


@pytest.mark.parametrize("h,m,L", [(2, 8, 90), (4, 20, 150), (8, 3, 40)])
def test_rel_table_from_fused_projection(h, m, L):
    """modules.RelMultiHeadedAttention.project_qkv_rel(x, table, m): q (+ pos_bias_u), k, v and the relative term's
    table (q + pos_bias_v) . linear_pos(table)[r] come out of ONE GEMM whose weight folds linear_q, linear_pos and the
    biases -- checked against the unfused fp64 arithmetic of modules/attention.py:104-117 (bf16 operands: 2e-2 of the
    output scale), also after a weight update (the folded weight is rebuilt)."""
    from cassnat_asr_public_b200 import modules
    d = 64 * h
    torch.manual_seed(h + m)
    attn = modules.RelMultiHeadedAttention(h, d, dropout=0.0).to(DEV).eval()
    table = torch.randn(2 * m + 1, d, device=DEV)
    x = torch.randn(2, L, d, device=DEV)
    ops.set_precision("bf16")
    try:
        for _ in range(2):
            with torch.no_grad():
                q, k, v, tab = attn.project_qkv_rel(x, table, m)
            xd = x.double()
            lin = lambda i: xd @ attn.linears[i].weight.double().t() + attn.linears[i].bias.double()
            qd = lin(0).view(2, L, h, 64)
            P = (table.double() @ attn.linear_pos.weight.double().t()).view(2 * m + 1, h, 64)
            want_tab = torch.einsum("blhk,rhk->blhr", qd + attn.pos_bias_v.double(), P).reshape(2, L, h * (2 * m + 1))
            want_q = (qd + attn.pos_bias_u.double()).reshape(2, L, d)
            for got, want in ((tab, want_tab), (q, want_q), (k, lin(1)), (v, lin(2))):
                assert got.dtype == torch.bfloat16 and got.shape == want.shape
                assert rel(got.float(), want) < 2e-2
            with torch.no_grad():                      # an optimizer step: the cached folded weight must follow
                attn.linears[0].weight.mul_(1.5)
                attn.pos_bias_v.add_(0.25)
    finally:
        ops.set_precision("fp32")
