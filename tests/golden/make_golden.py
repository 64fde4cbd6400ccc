"""Generates tests/golden/*.npz by calling the UNMODIFIED reference (/root/reference).

Run in the build container only:   python tests/golden/make_golden.py
The reference ships no tests or golden vectors, so these fixtures -- the
outputs of the reference's own functions on seeded inputs -- are what pins the oracle
(oracle/*.py) and, through it, the CUDA path.  Inputs are stored next to outputs so the
tests never need the reference at run time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_harness import VOCAB, load_reference, make_args  # noqa: E402

M = load_reference()


def _bare_model():
    """A CassNAT instance whose helper methods we can call (weights irrelevant)."""
    args = make_args(d_model=16, n_head=1, d_encff=16, d_decff=16, d_ff=16, vocab_size=8,
                     N_enc=1, N_mix_dec=1)
    return M.make_model(8, args)


def planted_posteriors(g, B, T, V, U, peak=6.0, ragged=True, repeat_p=0.3):
    """log-softmax of noise + a planted monotone alignment; returns tensors."""
    ys = torch.randint(3, V, (B, U), generator=g)
    rep = torch.rand(B, U, generator=g) < repeat_p
    for u in range(1, U):
        ys[:, u] = torch.where(rep[:, u], ys[:, u - 1], ys[:, u])
    ylens = torch.randint(max(U // 2, 0), U + 1, (B,), generator=g) if ragged else torch.full((B,), U)
    xlens = torch.randint(T // 2, T + 1, (B,), generator=g) if ragged else torch.full((B,), T)
    xlens = torch.maximum(xlens, 2 * ylens + 1).clamp(max=T)
    xlens[0] = T
    logits = torch.randn(B, T, V, generator=g)
    for b in range(B):
        n, L = int(ylens[b]), int(xlens[b])
        if n == 0:
            logits[b, :L, 0] += peak
            continue
        cuts = torch.sort(torch.randperm(L - 1, generator=g)[: 2 * n] + 1)[0].tolist() if L - 1 >= 2 * n else None
        if cuts is None:
            continue
        frame = torch.zeros(L, dtype=torch.long)
        for u in range(n):
            frame[cuts[2 * u]:cuts[2 * u + 1]] = ys[b, u]
        logits[b, torch.arange(L), frame] += peak
    ctc_out = torch.log_softmax(logits, -1)
    mask = (torch.arange(T)[None, :] < xlens[:, None]).unsqueeze(1)
    return ctc_out, mask, xlens, ys, ylens


def gen_align():
    import oracle.align_np as O
    model = _bare_model()
    out = {}
    # ---- SURVEY App. C known-answer vector -------------------------------------------
    fr = [[0, 3, 3, 0, 2, 2, 2, 1], [1, 1, 0, 1, 0, 2, 0, 0]]
    lp = torch.full((2, 8, 4), -5.0)
    for b in range(2):
        for t in range(8):
            lp[b, t, fr[b][t]] = -0.1
    mask = torch.tensor([[1] * 8, [1] * 6 + [0] * 2]).bool().unsqueeze(1)
    ssz = torch.tensor([8, 6])
    a, yl, ym = model.best_path_align(lp, mask, ssz, 0)
    tm, yl2, ym2 = model.align_to_mask(a, yl, ym, mask, ssz, 0)
    ex = model.expand_trigger_mask(tm, 1, 1) & mask
    ys = torch.tensor([[3, 2, 1], [1, 1, 2]])
    va, vyl, vym = model.viterbi_align(lp, mask, ssz, ys, torch.tensor([3, 3]), 0, 0)
    out.update(kat_lp=lp.numpy(), kat_mask=mask.numpy(), kat_src_size=ssz.numpy(),
               kat_bp_aligned=a.numpy(), kat_bp_ylen=yl.numpy(), kat_bp_ymax=np.int64(ym),
               kat_trigger=tm.numpy(), kat_trigger_ylen=yl2.numpy(), kat_trigger_ymax=np.int64(ym2),
               kat_expanded=ex.numpy(), kat_ys=ys.numpy(), kat_vit_aligned=va.numpy())
    assert a.tolist() == [[0, 0, 3, 0, 0, 2, 0, 0], [0, 1, 0, 0, 1, 0, 2, 0]] and yl.tolist() == [2, 3]

    # ---- randomised differential cases ------------------------------------------------
    g = torch.Generator().manual_seed(20251018)
    n_case = 0
    shapes = [(4, 24, 12, 5), (3, 40, 9, 8), (5, 17, 6, 3), (2, 64, 30, 20), (6, 33, 7, 1), (3, 21, 5, 0)]
    for (B, T, V, U) in shapes:
        for variant in range(3):
            peak = [6.0, 1.5, 0.0][variant]
            if U == 0:
                # ymax == 0 makes the reference's gather fail; use U=1 with ylens of 0/1
                ctc_out, mask, xl, ys, yls = planted_posteriors(g, B, T, V, 1, peak)
                yls = torch.randint(0, 2, (B,), generator=g)
            else:
                ctc_out, mask, xl, ys, yls = planted_posteriors(g, B, T, V, U, peak)
            if variant == 2:
                # quantise to create exact ties everywhere
                ctc_out = torch.round(ctc_out * 2) / 2
            pre = f"c{n_case}_"
            n_case += 1
            out[pre + "ctc_out"] = ctc_out.numpy()
            out[pre + "mask"] = mask.numpy()
            out[pre + "src_size"] = xl.numpy()
            out[pre + "ys"] = ys.numpy()
            out[pre + "ylens"] = yls.numpy()
            # best path, greedy
            a, yl, ym = model.best_path_align(ctc_out, mask, xl, 0)
            oa, oyl, oym = O.best_path_align(ctc_out.numpy(), mask.numpy(), 0)
            assert np.array_equal(oa, a.numpy()) and np.array_equal(oyl, yl.numpy()) and oym == ym
            out[pre + "bp_aligned"], out[pre + "bp_ylen"], out[pre + "bp_ymax"] = a.numpy(), yl.numpy(), np.int64(ym)
            # best path, sampled (skipped on the exact-tie variant: torch.topk's order among
            # EQUAL values over the vocab axis is implementation-defined -- it differs between
            # torch CPU and CUDA -- so only tie-free posteriors pin this mode)
            if variant != 2:
                # best path, sampled: replicate x S like cassnat.py:442-445 and draw like :373
                S = 3
                rep = lambda x: x.unsqueeze(1).repeat(1, S, *([1] * (x.dim() - 1))).reshape(-1, *x.shape[1:])
                ctc_r, mask_r, xl_r = rep(ctc_out), rep(mask), rep(xl)
                torch.manual_seed(1000 + n_case)
                sa, syl, sym = model.best_path_align(ctc_r, mask_r, xl_r, 0, S, 0.9)
                torch.manual_seed(1000 + n_case)
                select = torch.randint(0, 2, (B * S, T, 1))
                osa, osyl, osym = O.best_path_align(ctc_r.numpy(), mask_r.numpy(), 0, S, 0.9, True,
                                                    select.numpy()[..., 0])
                assert np.array_equal(osa, sa.numpy()) and np.array_equal(osyl, syl.numpy()) and osym == sym, pre
                out[pre + "bps_select"] = select.numpy()[..., 0].astype(np.uint8)
                out[pre + "bps_aligned"], out[pre + "bps_ylen"], out[pre + "bps_ymax"] = sa.numpy(), syl.numpy(), np.int64(sym)
            # viterbi exact
            va, _, vym = model.viterbi_align(ctc_out, mask, xl, ys, yls, 0, 0)
            ova, _, ovym = O.viterbi_align(ctc_out.numpy(), mask.numpy(), xl.numpy(), ys.numpy(), yls.numpy(), 0)
            assert np.array_equal(ova, va.numpy()) and ovym == vym, pre
            out[pre + "vit_aligned"] = va.numpy()
            # viterbi with 2nd-best transitions at sampled frames (sample_topk=4)
            torch.manual_seed(2000 + n_case)
            vsa, _, _ = model.viterbi_align(ctc_out, mask, xl, ys, yls, 0, 4)
            torch.manual_seed(2000 + n_case)
            t_sample = torch.randint(1, T, (1, 4)).numpy().tolist()[0]
            ovsa, _, _ = O.viterbi_align(ctc_out.numpy(), mask.numpy(), xl.numpy(), ys.numpy(), yls.numpy(), 0, t_sample)
            assert np.array_equal(ovsa, vsa.numpy()), pre
            out[pre + "vits_t_sample"] = np.asarray(t_sample, dtype=np.int64)
            out[pre + "vits_aligned"] = vsa.numpy()
            # trigger masks from both alignments, all dilation settings
            for tag, (al, yl_in, ym_in) in {"bp": (a, yl, ym), "vit": (va, yls, vym)}.items():
                tm, tyl, tym = model.align_to_mask(al, yl_in, ym_in, mask, xl, 0)
                otm, otyl, otym = O.align_to_mask(al.numpy(), yl_in.numpy(), ym_in, mask.numpy(), xl.numpy(), 0)
                assert np.array_equal(otm, tm.numpy()) and np.array_equal(otyl, tyl.numpy()) and otym == tym
                out[pre + tag + "_trigger"] = np.packbits(tm.numpy(), axis=-1)
                out[pre + tag + "_trigger_ylen"] = tyl.numpy()
                for (l, r) in [(0, 1), (1, 0), (1, 1)]:
                    ex = model.expand_trigger_mask(tm, l, r) & mask
                    oex = O.expand_trigger_mask(otm, l, r) & (mask.numpy() != 0)
                    assert np.array_equal(oex, ex.numpy())
                    out[pre + tag + f"_trigger_l{l}r{r}"] = np.packbits(ex.numpy(), axis=-1)
                tgt = torch.full((B, tym), 1).type_as(mask)
                tgt = tgt.scatter(1, tyl.unsqueeze(1) - 1, 0).cumprod(1)
                tgt = tgt.scatter(1, tyl.unsqueeze(1) - 1, 1).unsqueeze(1)
                assert np.array_equal(O.target_length_mask(tyl.numpy(), tym), tgt.numpy())
                out[pre + tag + "_tgt_mask"] = tgt.numpy().astype(np.uint8)
    out["n_case"] = np.int64(n_case)
    np.savez_compressed(os.path.join(HERE, "align.npz"), **out)
    print("align.npz:", n_case, "cases")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    which = sys.argv[1:] or ["align", "blocks"]
    if "align" in which:
        gen_align()
    if "blocks" in which:
        from make_golden_blocks import gen_blocks
        gen_blocks()
