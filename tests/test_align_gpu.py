"""GPU parity: K2/K3/K4/K5 through the C ABI against the oracle and the reference-generated
golden vectors.  Everything here is integer work -> bit-exact."""
import numpy as np
import pytest
import torch

import oracle.align_np as A
from cassnat_asr_public_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def unpack(bits, T):
    return np.unpackbits(bits, axis=-1)[..., :T].astype(bool)


def test_known_answer_vector(align_golden):
    g = align_golden
    a, yl, ym = ops.best_path_align(cu(g["kat_lp"]), cu(g["kat_mask"]), cu(g["kat_src_size"]), 0)
    assert a.cpu().tolist() == [[0, 0, 3, 0, 0, 2, 0, 0], [0, 1, 0, 0, 1, 0, 2, 0]]
    assert yl.cpu().tolist() == [2, 3] and ym == 3 and a.dtype == torch.int64
    tm, yl2, ym2 = ops.align_to_mask(a, yl, ym, cu(g["kat_mask"]), cu(g["kat_src_size"]), 0)
    assert np.array_equal(tm.cpu().numpy(), g["kat_trigger"]) and yl2.cpu().tolist() == [3, 4] and ym2 == 4
    ex = ops.expand_trigger_mask(tm, 1, 1) & cu(g["kat_mask"])
    assert np.array_equal(ex.cpu().numpy(), g["kat_expanded"])
    va, _, _ = ops.viterbi_align(cu(g["kat_lp"]), cu(g["kat_mask"]), cu(g["kat_src_size"]), cu(g["kat_ys"]),
                                 cu(np.array([3, 3])), 0)
    assert np.array_equal(va.cpu().numpy(), g["kat_vit_aligned"])


def test_golden_cases(align_golden):
    g = align_golden
    for c in range(int(g["n_case"])):
        p = f"c{c}_"
        lp, mask, xl, ys, yls = (cu(g[p + k]) for k in ("ctc_out", "mask", "src_size", "ys", "ylens"))
        T = lp.shape[1]
        a, yl, ym = ops.best_path_align(lp, mask, xl, 0)
        assert np.array_equal(a.cpu().numpy(), g[p + "bp_aligned"]), p
        assert np.array_equal(yl.cpu().numpy(), g[p + "bp_ylen"]) and ym == int(g[p + "bp_ymax"])
        if p + "bps_select" in g:
            S = 3
            torch.manual_seed(1000 + c + 1)            # the seed make_golden.py used for this draw
            sa, syl, sym = ops.best_path_align(lp.repeat_interleave(S, 0), mask.repeat_interleave(S, 0),
                                               xl.repeat_interleave(S, 0), 0, S, 0.9)
            assert np.array_equal(sa.cpu().numpy(), g[p + "bps_aligned"]), p
            assert np.array_equal(syl.cpu().numpy(), g[p + "bps_ylen"]) and sym == int(g[p + "bps_ymax"])
        va, _, vym = ops.viterbi_align(lp, mask, xl, ys, yls, 0)
        assert np.array_equal(va.cpu().numpy(), g[p + "vit_aligned"]), p
        torch.manual_seed(2000 + c + 1)
        vs, _, _ = ops.viterbi_align(lp, mask, xl, ys, yls, 0, 4)
        assert np.array_equal(vs.cpu().numpy(), g[p + "vits_aligned"]), p
        # the same alignments from the compact emissions [B,T,U+1] (gathered once, coalesced DP reads)
        emis = ops.gather_emissions(lp, ys, 0)
        want_e = torch.cat([lp[:, :, :1], torch.gather(lp, 2, ys.clamp(0, lp.shape[2] - 1).unsqueeze(1).expand(-1, T, -1))], 2)
        assert torch.equal(emis, want_e), p
        ca, _, _ = ops.viterbi_align_compact(emis, lp.shape[2], mask, xl, ys, yls, 0)
        assert np.array_equal(ca.cpu().numpy(), g[p + "vit_aligned"]), p
        torch.manual_seed(2000 + c + 1)
        cs, _, _ = ops.viterbi_align_compact(emis, lp.shape[2], mask, xl, ys, yls, 0, 4)
        assert np.array_equal(cs.cpu().numpy(), g[p + "vits_aligned"]), p
        for tag, (al, yin, ymx) in {"bp": (a, yl, ym), "vit": (va, yls, vym)}.items():
            tm, tyl, tym = ops.align_to_mask(al, yin, ymx, mask, xl, 0)
            assert np.array_equal(tm.cpu().numpy(), unpack(g[p + tag + "_trigger"], T)), (p, tag)
            assert np.array_equal(tyl.cpu().numpy(), g[p + tag + "_trigger_ylen"])
            for (l, r) in [(0, 1), (1, 0), (1, 1)]:
                want = unpack(g[p + tag + f"_trigger_l{l}r{r}"], T)
                ex = ops.expand_trigger_mask(tm, l, r) & mask
                assert np.array_equal(ex.cpu().numpy(), want), (p, tag, l, r)
                fused, _, _ = ops.align_to_mask(al, yin, ymx, mask, xl, 0, l, r, True)
                assert np.array_equal(fused.cpu().numpy(), want), (p, tag, l, r, "fused")
            tg = ops.tgt_length_mask(tyl, tym)
            assert tg.dtype == torch.int64 and np.array_equal(tg.cpu().numpy()[:, 0], g[p + tag + "_tgt_mask"][:, 0])


def planted(seed, B, T, V, U, peak=8.0, ragged=True):
    """SURVEY §8d cfg2 generator: noise + a planted monotone alignment (numpy, seeded)."""
    rng = np.random.default_rng(seed)
    ys = rng.integers(4, V, size=(B, U))
    ylens = rng.integers(U // 2, U + 1, size=B) if ragged else np.full(B, U)
    xlens = rng.integers(T // 2, T + 1, size=B) if ragged else np.full(B, T)
    xlens = np.minimum(np.maximum(xlens, 2 * ylens + 2), T)
    xlens[0] = T
    logits = rng.standard_normal((B, T, V), dtype=np.float32)
    for b in range(B):
        n, L = int(ylens[b]), int(xlens[b])
        cuts = np.sort(rng.choice(np.arange(1, L), size=2 * n, replace=False))
        for u in range(n):
            logits[b, cuts[2 * u]:cuts[2 * u + 1], ys[b, u]] += peak
    lp = torch.log_softmax(torch.from_numpy(logits), -1).numpy()
    mask = (np.arange(T)[None, :] < xlens[:, None])[:, None, :]
    return lp, mask, xlens.astype(np.int64), ys.astype(np.int64), ylens.astype(np.int64)


@pytest.mark.parametrize("B,T,V,U,peak", [(6, 200, 300, 40, 8.0), (4, 333, 97, 100, 2.0), (3, 150, 5000, 7, 0.0),
                                          (2, 520, 64, 255, 4.0), (2, 700, 50, 300, 3.0), (5, 64, 40, 1, 1.0),
                                          (2, 1000, 40, 450, 3.0), (3, 257, 80, 31, 2.0), (3, 300, 80, 63, 2.0)])
def test_random_vs_oracle(B, T, V, U, peak):
    lp, mask, xl, ys, yls = planted(B * 1000 + T, B, T, V, U, peak)
    a, yl, ym = ops.best_path_align(cu(lp), cu(mask), cu(xl), 0)
    oa, oyl, oym = A.best_path_align(lp, mask, 0)
    assert np.array_equal(a.cpu().numpy(), oa) and np.array_equal(yl.cpu().numpy(), oyl) and ym == oym
    va, _, _ = ops.viterbi_align(cu(lp), cu(mask), cu(xl), cu(ys), cu(yls), 0)
    ova, _, _ = A.viterbi_align(lp, mask, xl, ys, yls, 0)
    assert np.array_equal(va.cpu().numpy(), ova)
    ca, _, _ = ops.viterbi_align_compact(ops.gather_emissions(cu(lp), cu(ys), 0), V, cu(mask), cu(xl), cu(ys), cu(yls), 0)
    assert np.array_equal(ca.cpu().numpy(), ova)
    for (al, yin, ymx) in ((oa, oyl, oym), (ova, yls, U)):
        for (l, r) in [(0, 0), (1, 1)]:
            tm, tyl, tym, seg = ops.align_to_mask(cu(al), cu(yin), ymx, cu(mask), cu(xl), 0, l, r, True, want_seg=True)
            otm, otyl, otym = A.align_to_mask(al, yin, ymx, mask, xl, 0)
            otm = A.expand_trigger_mask(otm, l, r) & (mask != 0)
            assert np.array_equal(tm.cpu().numpy(), otm) and np.array_equal(tyl.cpu().numpy(), otyl) and tym == otym
            assert np.array_equal(seg.cpu().numpy(), np.cumsum(al != 0, 1))


def test_row_stats_vs_torch():
    torch.manual_seed(0)
    for (rows, V) in [(37, 5000), (200, 4990), (8, 7), (1, 1), (64, 1031)]:
        x = torch.randn(rows, V, device=DEV) * 3
        x[:, V // 2] = x[:, 0]                       # exact tie inside every row
        out = torch.empty_like(x)
        st = ops.row_stats(x, top2=True, peak=True, lse=True, logp_out=out)
        assert torch.equal(st["top1"].long(), x.argmax(-1))
        assert torch.equal(st["peak"], x.max(-1)[0])
        if V > 1:
            order = np.argsort(-x.double().cpu().numpy(), axis=-1, kind="stable")
            assert np.array_equal(st["top2"].cpu().numpy(), order[:, 1])
        ref = torch.log_softmax(x.double(), -1)
        assert (out.double() - ref).abs().max().item() < 2e-5
        assert (st["lse"].double() - torch.logsumexp(x.double(), -1)).abs().max().item() < 2e-5


def test_cfg2_full_size_properties():
    """BASELINE configs[1]: B=256, T=1000, V=5000, U=200 (5.12 GB of posteriors), checked through
    size-independent properties + the oracle on a slice of utterances."""
    B, T, V, U = 256, 1000, 5000, 200
    g = torch.Generator(device=DEV).manual_seed(1234)
    ys = torch.randint(4, V, (B, U), device=DEV, generator=g)
    logits = torch.randn(B, T, V, device=DEV, generator=g)
    # planted alignment: token u occupies frames [5u+1, 5u+3]
    t_idx = (torch.arange(U, device=DEV) * 5)[None, :, None] + torch.arange(1, 4, device=DEV)[None, None, :]
    b_idx = torch.arange(B, device=DEV)[:, None, None].expand(B, U, 3)
    logits[b_idx, t_idx.expand(B, U, 3), ys[:, :, None].expand(B, U, 3)] += 8.0
    lp = torch.log_softmax(logits, -1)
    del logits
    mask = torch.ones(B, 1, T, dtype=torch.bool, device=DEV)
    xl = torch.full((B,), T, dtype=torch.int64, device=DEV)
    yls = torch.full((B,), U, dtype=torch.int64, device=DEV)
    a, yl, ym = ops.best_path_align(lp, mask, xl, 0)
    va, _, _ = ops.viterbi_align(lp, mask, xl, ys, yls, 0)
    # property 1: the collapsed Viterbi alignment spells exactly ys, in order
    for b in range(0, B, 17):
        toks = va[b][va[b] != 0]
        assert torch.equal(toks, ys[b]), b
    # property 2: with a strongly planted path best-path == Viterbi on the planted tokens' frames
    assert (yl >= U).all()
    # property 3: trigger mask rows partition the frames (each frame in exactly one row)
    tm, tyl, tym = ops.align_to_mask(va, yls, U, mask, xl, 0)
    assert tm.shape == (B, U + 1, T) and bool((tm.sum(1) == 1).all())
    assert bool((tm.any(-1).sum(-1) == U + 1).all())
    # oracle on a slice
    sl = slice(0, 3)
    ova, _, _ = A.viterbi_align(lp[sl].cpu().numpy(), mask[sl].cpu().numpy(), xl[sl].cpu().numpy(),
                                ys[sl].cpu().numpy(), yls[sl].cpu().numpy(), 0)
    assert np.array_equal(va[sl].cpu().numpy(), ova)
    oa, oyl, _ = A.best_path_align(lp[sl].cpu().numpy(), mask[sl].cpu().numpy(), 0)
    assert np.array_equal(a[sl].cpu().numpy(), oa) and np.array_equal(yl[sl].cpu().numpy(), oyl)


def test_decode_hints_kernel():
    """cn_decode_hints: rows needed, valid tokens, last valid frame and the heaviest-first (stable) orders, against
    the same quantities computed with torch."""
    g = torch.Generator().manual_seed(3)
    for B, T, ymax, use_q in ((1, 5, 3, True), (64, 500, 501, True), (37, 130, 77, False), (300, 33, 20, True)):
        ylen = torch.randint(1, ymax + 3, (B,), generator=g)
        lens = torch.randint(0, T + 1, (B,), generator=g)
        mask = torch.arange(T)[None] < lens[:, None]
        if B > 2:
            mask[1, :] = False                           # an utterance without any valid frame
            mask[2, T // 2] = False                      # a hole inside the valid range
        h = ops.decode_hints(ylen.to("cuda:0"), ymax, mask.to("cuda:0"), use_q=use_q)
        need = (ylen + 1).clamp(max=ymax)
        kv = ylen.clamp(max=ymax)
        last = (mask * torch.arange(1, T + 1)).amax(1)
        assert torch.equal(h["need_rows"].cpu().long(), need) and torch.equal(h["kv_tgt"].cpu().long(), kv)
        assert torch.equal(h["last_key"].cpu().long(), last)
        qt = (need + 127) // 128 if use_q else torch.ones_like(need)
        for key, got in ((((kv + 127) // 128) * qt, h["order_tgt"]), (((last + 127) // 128) * qt, h["order_src"])):
            want = torch.argsort(key, descending=True, stable=True)
            assert torch.equal(got.cpu().long(), want)


@pytest.mark.parametrize("U", [5, 40, 120, 200])
def test_viterbi_with_masked_frames_inside_the_utterance(U):
    """Frames masked out INSIDE the valid length.  The reference's prefix masks never produce them, and its own result
    is then ill-defined (every alpha drops below LOGZERO, so the forbidden / out-of-lattice candidates, which hold
    LOGZERO, win and the back-pointers of states 0 and 1 index state -1): what is pinned here is that the kernel stays
    memory-safe and deterministic, and that the posterior path and the compact path (cp.async ring + patching of the
    masked rows) agree bit for bit."""
    B, T, V = 4, 2 * U + 150, 60
    lp, mask, xl, ys, yls = planted(77 + U, B, T, V, U, 3.0)
    rng = np.random.default_rng(U)
    mask = mask.copy()
    for b in range(B):
        holes = rng.choice(np.arange(1, int(xl[b]) - 1), size=7, replace=False)
        mask[b, 0, holes] = False
    va, _, _ = ops.viterbi_align(cu(lp), cu(mask), cu(xl), cu(ys), cu(yls), 0)
    vb, _, _ = ops.viterbi_align(cu(lp), cu(mask), cu(xl), cu(ys), cu(yls), 0)
    ca, _, _ = ops.viterbi_align_compact(ops.gather_emissions(cu(lp), cu(ys), 0), V, cu(mask), cu(xl), cu(ys), cu(yls), 0)
    assert torch.equal(va, vb) and torch.equal(va, ca)
    labels = set(np.unique(ys).tolist()) | {0}
    assert set(np.unique(va.cpu().numpy()).tolist()) <= labels
