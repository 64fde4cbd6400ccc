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


@pytest.mark.parametrize("M,N,K", [(300, 5000, 512), (129, 120, 128), (1000, 4990, 256), (32, 257, 64)])
def test_fused_head_rowstats(M, N, K):
    """K1/K9 fused epilogue vs row_stats on the materialised logits of the same bf16 GEMM."""
    torch.manual_seed(M + N)
    x = torch.randn(M, K).bfloat16().to(DEV)
    w = (torch.randn(N, K) / K ** 0.5).bfloat16().to(DEV)
    w[N // 3] = w[0]                                   # exact duplicate column -> argmax / top-2 ties
    b = torch.randn(N).to(DEV)
    b[N // 3] = b[0]
    logits = ops.linear_bf16(x, w, b)
    want = ops.row_stats(logits, top2=True, peak=True, lse=True)
    from cassnat_asr_public_b200 import _lib
    from cassnat_asr_public_b200._lib import call, ptr, stream
    got = {k: torch.empty(M, dtype=(torch.int32 if k.startswith("top") else torch.float32), device=DEV)
           for k in ("top1", "top2", "peak", "lse")}
    nbytes = _lib.load().cn_gemm_rowstats_workspace_bytes(M, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    call("cn_gemm_bf16_rowstats", ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(b), 1.0, M, N, K, ptr(got["top1"]),
         ptr(got["top2"]), ptr(got["peak"]), ptr(got["lse"]), None, None, ptr(ws), nbytes, stream())
    assert torch.equal(got["top1"], want["top1"]) and torch.equal(got["top2"], want["top2"])
    assert torch.equal(got["peak"], want["peak"])
    assert (got["lse"] - want["lse"]).abs().max().item() < 1e-4
    # and without the runner-up (the greedy decode configuration)
    call("cn_gemm_bf16_rowstats", ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(b), 1.0, M, N, K, ptr(got["top1"]),
         None, ptr(got["peak"]), ptr(got["lse"]), None, None, ptr(ws), nbytes, stream())
    assert torch.equal(got["top1"], want["top1"]) and (got["lse"] - want["lse"]).abs().max().item() < 1e-4
    # the logit of a given column per row (rescoring: log p(id) = target_logit - lse), ids outside [0,N) -> 0
    tgt = torch.randint(0, N, (M,), device=DEV, dtype=torch.int32)
    tgt[::7] = -1
    st = ops.head_rowstats(x, w, b, target=tgt)
    ref = logits.gather(1, tgt.clamp(min=0).long().unsqueeze(1)).squeeze(1).masked_fill(tgt < 0, 0.0)
    assert torch.equal(st["target_logit"], ref)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 136), (1312, 512, 2048), (1000, 72, 520), (5000, 512, 1312)])
def test_gemm_mn_major_operands(M, N, K):
    """cn_gemm_bf16_mn: the backward GEMMs with operands in contraction-major storage (no transposed copies):
    C = A W with W stored [K,N] (dX = dY W), and C = A^T W with A stored [K,M], W stored [K,N] (dW = dY^T X)."""
    from cassnat_asr_public_b200._lib import call, ptr, stream
    torch.manual_seed(M + N + K)
    pad = lambda n: (n + 7) // 8 * 8
    a = torch.randn(M, pad(K)).bfloat16().to(DEV)[:, :K]              # [M,K] K-major, padded pitch
    at = torch.randn(K, pad(M)).bfloat16().to(DEV)[:, :M]             # [K,M] contraction-major
    w = (torch.randn(K, pad(N)) / K ** 0.5).bfloat16().to(DEV)[:, :N]  # [K,N] contraction-major
    for a_mn, A in ((0, a), (1, at)):
        out = torch.empty(M, N, dtype=torch.float32, device=DEV)
        call("cn_gemm_bf16_mn", ptr(A), A.stride(0), a_mn, ptr(w), w.stride(0), 1, ptr(out), None, N, M, N, K, stream())
        ref = (A.double().t() if a_mn else A.double()) @ w.double()
        assert ((out.double() - ref).abs().max() / ref.abs().max()).item() < 2e-5, (a_mn, M, N, K)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 512, 512), (1000, 5000, 512), (513, 1536, 2048), (77, 120, 128)])
def test_gemm_bf16x3(M, N, K):
    """bf16x3 (split) GEMM: fp32 operands carried as (hi, lo) bf16 pairs, three tensor-core passes.  Against the
    fp64 product of the fp32 operands the error is ~2^-17 of |a||w| (plain bf16: ~2^-9)."""
    torch.manual_seed(M * 3 + N + K)
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    b = torch.randn(N, device=DEV)
    r = torch.randn(M, N, device=DEV)
    ref = x.double() @ w.double().t()
    ops.set_precision("bf16x3")
    try:
        y = ops.linear(x, w)
        assert y.dtype == torch.float32 and rel(y, ref) < 3e-5
        y = ops.linear(x, w, b, r, 0.5, "relu")
        assert rel(y, 0.5 * torch.relu(ref + b.double()) + r.double()) < 3e-5
        # split output feeding the next GEMM: hi + lo reproduces the fp32 result to 2^-16
        h = ops.linear(x, w, b, out_bf16=True, out_split=True)
        full = ops.split_of(h)
        assert full is not None and full.shape == (M, 2 * N)
        assert rel(full[:, :N].float() + full[:, N:].float(), ref + b.double()) < 3e-5
        if N % 64 == 0:                                          # (the split contraction needs K % 64 == 0)
            w2 = torch.randn(64, N, device=DEV) / N ** 0.5
            y2 = ops.linear(h, w2)
            assert rel(y2, (ref + b.double()) @ w2.double().t()) < 5e-5
        # plain bf16 A (an attention context): only the weight is split (two passes)
        xb = x.bfloat16()
        y3 = ops.linear(xb, w)
        assert rel(y3, xb.double() @ w.double().t()) < 3e-5
        # LayerNorm emitting the split operand
        a2, b2 = torch.rand(K, device=DEV) + 0.5, torch.randn(K, device=DEV)
        ln = ops.layer_norm(x, a2, b2, bf16=True)
        lnf = ops.split_of(ln)
        ops.set_precision("fp32")
        want = ops.layer_norm(x, a2, b2)
        assert rel(lnf[:, :K].float() + lnf[:, K:].float(), want) < 2e-5
    finally:
        ops.set_precision("fp32")
    plain = ops.linear_bf16(x.bfloat16(), w.bfloat16())
    assert rel(plain, ref) > 3e-4                                # plain bf16 operands are visibly coarser


@pytest.mark.parametrize("M,N,K", [(300, 5000, 512), (129, 120, 128), (1000, 4990, 256)])
def test_decision_safe_head(M, N, K):
    """K1/K9 as bf16x3: argmax and runner-up equal those of the fp32 logits wherever the fp32 top-2 (resp. 2-3)
    margin exceeds 1e-4; peak / lse within 1e-4.  (cassnat.py:367-389 argmax -> alignment, :611 token pick.)"""
    torch.manual_seed(M + N + 1)
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    b = torch.randn(N, device=DEV) * 0.1
    z = (x.double() @ w.double().t() + b.double())
    top = z.topk(3, -1)
    ops.set_precision("bf16")
    try:
        st = ops.head_rowstats(x, w, b, top2=True)              # safe by default
        assert st["top1"].dtype == torch.int32
        clear1 = (top.values[:, 0] - top.values[:, 1]) > 1e-4
        clear2 = clear1 & ((top.values[:, 1] - top.values[:, 2]) > 1e-4)
        assert torch.equal(st["top1"].long()[clear1], top.indices[:, 0][clear1])
        assert torch.equal(st["top2"].long()[clear2], top.indices[:, 1][clear2])
        assert (st["peak"].double() - top.values[:, 0]).abs().max().item() < 1e-4
        assert (st["lse"].double() - torch.logsumexp(z, -1)).abs().max().item() < 1e-4
        plain = ops.head_rowstats(x, w, b, safe=False)
        assert (plain["peak"].double() - top.values[:, 0]).abs().max().item() > 1e-4      # the bf16 head is coarser
    finally:
        ops.set_precision("fp32")


@pytest.mark.parametrize("M,N,K,res", [(300, 512, 512, True), (32000, 512, 512, True), (1000, 512, 2048, True),
                                       (500, 512, 512, "rows"), (129, 256, 256, True), (4000, 256, 2048, True),
                                       (77, 128, 128, True), (640, 512, 512, False), (1, 512, 64, True)])
def test_gemm_with_fused_layernorm(M, N, K, res):
    """cn_gemm_bf16_ln: residual GEMM + the LayerNorm that consumes it (modules/utils.py:25-32 + norm.py:15-18) in one
    kernel.  The fp32 stream must equal the unfused GEMM bit for bit; the normalised bf16 copy must equal the
    stand-alone LayerNorm of that stream up to one bf16 rounding (reciprocal instead of division)."""
    from types import SimpleNamespace
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=DEV)
    if res == "rows":
        r, rr = torch.randn(100, N, device=DEV) * 2 + 0.3, 100
    else:
        r, rr = ((torch.randn(M, N, device=DEV) * 2 + 0.3) if res else None), 0
    ln = SimpleNamespace(a_2=torch.rand(N, device=DEV) + 0.5, b_2=torch.randn(N, device=DEV), eps=1e-6)
    want32 = ops.linear_bf16(x, w, b, r, 0.5, None, rr)
    got32 = ops.linear_bf16_ln(x, w, b, r, 0.5, rr, ln)
    assert torch.equal(got32, want32)
    tag, got16 = got32._cn_ln
    assert tag is ln and got16.dtype == torch.bfloat16 and got16.shape == (M, N)
    want = ops.layer_norm_forward(want32, ln.a_2, ln.b_2, ln.eps)                       # fp32 LayerNorm of the same rows
    err = (got16.float() - want).abs()
    tol = want.abs() * 2.0 ** -8 + 1e-6                                                  # one bf16 rounding
    assert bool((err <= tol).all()), float((err - tol).max())


@pytest.mark.parametrize("M,d,N2,K1,act", [(300, 512, 1536, 512, None), (32000, 512, 2048, 2048, "relu"), (129, 256, 512, 256, "swish"),
                                           (1000, 128, 384, 128, None), (77, 512, 512, 512, None)])
def test_layernorm_folded_into_gemms(M, d, N2, K1, act):
    """LayerNorm folding: the producer residual GEMM emits fp32 rows + raw bf16 rows + partial row sums, the consumer GEMM
    applies LN(x) W^T = rinv (x W'^T - mean c) + bias' in its epilogue (cn_gemm_bf16_rowpart / cn_gemm_bf16_folded).
    Against the unfolded chain (GEMM -> LayerNorm kernel -> GEMM) on the same operands: the fp32 stream bit-identical,
    the consumer output within bf16 operand rounding (the two forms round different quantities to bf16)."""
    from types import SimpleNamespace
    torch.manual_seed(M + d + N2)
    x = torch.randn(M, K1, device=DEV).bfloat16()
    w1 = (torch.randn(d, K1, device=DEV) / K1 ** 0.5).bfloat16()
    b1 = torch.randn(d, device=DEV)
    res = torch.randn(M, d, device=DEV) * 2 + 0.2
    ln = SimpleNamespace(a_2=torch.rand(d, device=DEV) + 0.5, b_2=torch.randn(d, device=DEV) * 0.3, eps=1e-6)
    w2 = torch.randn(N2, d, device=DEV) / d ** 0.5
    b2 = torch.randn(N2, device=DEV)
    want32 = ops.linear_bf16(x, w1, b1, res, 0.5, None, 0)
    got32 = ops.linear_bf16_rowpart(x, w1, b1, res, 0.5, 0, ln)
    assert torch.equal(got32, want32)
    tag, x16, part, n_parts = got32._cn_fold
    assert tag is ln and torch.equal(x16, want32.bfloat16())
    tot = part.sum(1)
    assert rel(tot[:, 0], want32.sum(1)) < 1e-5 and rel(tot[:, 1], want32.square().sum(1)) < 1e-5
    y_fold = ops.linear_folded(ops.FoldedLN(x16, part, n_parts, ln), w2, b2, 1.0, act, out_bf16=False)
    xn = ops.layer_norm_forward(want32, ln.a_2, ln.b_2, ln.eps)                         # fp32 LayerNorm
    z = xn.double() @ w2.double().t() + b2.double()
    ref = torch.relu(z) if act == "relu" else (z * torch.sigmoid(z) if act == "swish" else z)
    y_chain = ops.linear_bf16(xn.bfloat16(), w2.bfloat16(), b2, None, 1.0, act)           # the unfolded fast-mode chain
    e_fold, e_chain = rel(y_fold, ref), rel(y_chain, ref)
    assert e_fold < 2e-2 and e_fold < 2.5 * e_chain + 1e-3, (e_fold, e_chain)
