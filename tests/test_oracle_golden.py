"""CPU: the oracle (oracle/*.py) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden*.py).  Bit-exact for integer outputs."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import oracle.align_np as A
import oracle.nat_torch as O


def test_known_answer_vector(align_golden):
    g = align_golden
    a, yl, ym = A.best_path_align(g["kat_lp"], g["kat_mask"], 0)
    assert a.tolist() == [[0, 0, 3, 0, 0, 2, 0, 0], [0, 1, 0, 0, 1, 0, 2, 0]]
    assert np.array_equal(a, g["kat_bp_aligned"]) and yl.tolist() == [2, 3] and ym == 3
    tm, yl2, ym2 = A.align_to_mask(a, yl, ym, g["kat_mask"], g["kat_src_size"], 0)
    assert np.array_equal(tm, g["kat_trigger"]) and yl2.tolist() == [3, 4] and ym2 == 4
    ex = A.expand_trigger_mask(tm, 1, 1) & (g["kat_mask"] != 0)
    assert np.array_equal(ex, g["kat_expanded"])
    va, _, _ = A.viterbi_align(g["kat_lp"], g["kat_mask"], g["kat_src_size"], g["kat_ys"], np.array([3, 3]), 0)
    assert np.array_equal(va, g["kat_vit_aligned"])


def test_alignment_cases(align_golden):
    g = align_golden
    for c in range(int(g["n_case"])):
        p = f"c{c}_"
        lp, mask, xl, ys, yls = g[p + "ctc_out"], g[p + "mask"], g[p + "src_size"], g[p + "ys"], g[p + "ylens"]
        T = lp.shape[1]
        a, yl, ym = A.best_path_align(lp, mask, 0)
        assert np.array_equal(a, g[p + "bp_aligned"]) and np.array_equal(yl, g[p + "bp_ylen"]) and ym == int(g[p + "bp_ymax"])
        if p + "bps_select" in g:
            S = 3
            rep = lambda x: np.repeat(x, S, axis=0)
            sa, syl, sym = A.best_path_align(rep(lp), rep(mask), 0, S, 0.9, True, g[p + "bps_select"])
            assert np.array_equal(sa, g[p + "bps_aligned"]) and np.array_equal(syl, g[p + "bps_ylen"])
        va, _, vym = A.viterbi_align(lp, mask, xl, ys, yls, 0)
        assert np.array_equal(va, g[p + "vit_aligned"])
        vs, _, _ = A.viterbi_align(lp, mask, xl, ys, yls, 0, g[p + "vits_t_sample"].tolist())
        assert np.array_equal(vs, g[p + "vits_aligned"])
        for tag, (al, yin, ymx) in {"bp": (a, yl, ym), "vit": (va, yls, vym)}.items():
            tm, tyl, tym = A.align_to_mask(al, yin, ymx, mask, xl, 0)
            assert np.array_equal(tm, np.unpackbits(g[p + tag + "_trigger"], axis=-1)[..., :T].astype(bool))
            assert np.array_equal(tyl, g[p + tag + "_trigger_ylen"])
            for (l, r) in [(0, 1), (1, 0), (1, 1)]:
                ex = A.expand_trigger_mask(tm, l, r) & (mask != 0)
                assert np.array_equal(ex, np.unpackbits(g[p + tag + f"_trigger_l{l}r{r}"], axis=-1)[..., :T].astype(bool))
            assert np.array_equal(A.target_length_mask(tyl, tym)[:, 0], g[p + tag + "_tgt_mask"][:, 0])


def _rel(a, b):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-6)).item()


def _hyps(z, tag):
    flat, lens = z[tag + "::hyp_flat"], z[tag + "::hyp_len"]
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n].tolist())
        o += n
    return out


def test_decode_port(blocks_golden):
    name, z, sd, cfg = blocks_golden
    src, mask, size = (torch.from_numpy(z[k]) for k in ("src", "src_mask", "src_size"))
    for tag, (l, r) in {"greedy": (0, 0), "greedy_l1r1": (1, 1)}.items():
        c = SimpleNamespace(**{**vars(cfg), "left_trigger": l, "right_trigger": r})
        with torch.no_grad():
            o = O.beam_decode(src, mask, size, sd, c)
        for key in ("enc_h", "ctc_out", "ac_embed", "pred_embed", "dec_h", "att_out"):
            assert _rel(o[key], z[f"{tag}::{key}"]) < 2e-5, (name, tag, key)
        assert np.array_equal(o["aligned"], z[tag + "::aligned"])
        T = o["trigger"].shape[-1]
        assert np.array_equal(o["trigger"], np.unpackbits(z[tag + "::trigger"], axis=-1)[..., :T].astype(bool))
        assert o["hyps"] == _hyps(z, tag)
        assert np.allclose(o["scores"], z[tag + "::score"], rtol=1e-5)


def test_train_port(blocks_golden):
    name, z, sd, cfg = blocks_golden
    enc_h = torch.from_numpy(z["train::enc_h"])
    x_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2]
    o = O.train_forward_from_encoder(enc_h, x_mask, torch.from_numpy(z["train::frames"]),
                                     torch.from_numpy(z["train::labels"]), torch.from_numpy(z["train::label_sizes"]),
                                     sd, cfg)
    assert np.array_equal(o["aligned"], z["train::aligned"])
    assert _rel(o["ctc_out"], z["train::ctc_out"]) < 2e-5 and _rel(o["att_out"], z["train::att_out"]) < 2e-5
    for k in ("loss", "ctc_loss", "att_loss"):
        assert abs(float(o[k]) - float(z["train::" + k])) < 1e-4 * max(1.0, abs(float(z["train::" + k])))


def test_train_port_gradients(blocks_golden):
    """The oracle port differentiated by torch autograd reproduces the parameter gradients of the
    unmodified reference's loss.backward() (whole model, encoder included): the pin for the
    backward-pass parity tests."""
    name, z, sd, cfg = blocks_golden
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"train_grads_{name}.npz"))
    leaf = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    src, mask, size = (torch.from_numpy(z[k]) for k in ("src", "src_mask", "src_size"))
    enc_h, x_mask = O.encode(src, mask, leaf, cfg)
    o = O.train_forward_from_encoder(enc_h, x_mask, torch.from_numpy(z["train::frames"]),
                                     torch.from_numpy(z["train::labels"]), torch.from_numpy(z["train::label_sizes"]),
                                     leaf, cfg)
    o["loss"].backward()
    checked = 0
    for key in g.files:
        if not key.startswith("grad::"):
            continue
        want = g[key]
        got = leaf[key[6:]].grad
        assert got is not None, key
        scale = max(np.abs(want).max(), 1e-6)
        assert np.abs(got.numpy() - want).max() / scale < 2e-4, (name, key)
        checked += 1
    assert checked > 100


def test_lm_port():
    """Oracle port of the LM ranker == the unmodified reference's TransformerLM (tests/golden/lm.npz)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lm.npz"))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    with torch.no_grad():
        got = O.lm_forward(torch.from_numpy(z["ids"]), torch.from_numpy(z["mask"]), sd, int(z["args"][0]))
    assert _rel(got, z["lm_out"]) < 2e-5


def test_train_port_intermediate_losses():
    """Oracle port of the training call with interctc / interce / use_unimask == the unmodified reference
    (tests/golden/inter.npz)."""
    import json
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "inter.npz"))
    src, mask, size = (torch.from_numpy(z[k]) for k in ("src", "src_mask", "src_size"))
    labels, sizes = torch.from_numpy(z["labels"]), torch.from_numpy(z["label_sizes"])
    for fam, case in (("trf", "interctc_mixce"), ("trf", "unimask"), ("con", "interctc_selfce"), ("con", "mixce")):
        pre = f"{fam}::{case}::"
        cfg = SimpleNamespace(**json.loads(bytes(z[pre + "cfg_json"]).decode()))
        sd = {k[len(fam) + 6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{fam}::sd::")}
        tap = [cfg.interctc_layer, None] if cfg.interctc_alpha > 0 else None
        with torch.no_grad():
            enc_h, x_mask = O.encode(src, mask, sd, cfg, tap=tap)
            o = O.train_forward_from_encoder(enc_h, x_mask, (size * enc_h.size(1)).long(), labels, sizes, sd, cfg,
                                             inter_h=tap[1] if tap else None)
        assert _rel(o["ctc_out"], z[pre + "ctc_out"]) < 2e-5 and _rel(o["att_out"], z[pre + "att_out"]) < 2e-5, (fam, case)
        assert abs(float(o["loss"]) - float(z[pre + "losses"][0])) < 1e-4 * abs(float(z[pre + "losses"][0])), (fam, case)


def test_beam_assembly_port():
    """Oracle restatement of the beam / LM-fusion hypothesis assembly (cassnat.py:574-637) == the unmodified
    reference (tests/golden/beam.npz on the blocks_trf fixture)."""
    from conftest import load_blocks
    z, sd, cfg = load_blocks("trf")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "beam.npz"))

    class TinyLM(torch.nn.Module):
        def __init__(self, V, seed=7):
            super().__init__()
            gen = torch.Generator().manual_seed(seed)
            self.emb, self.out = torch.randn(V, 16, generator=gen), torch.randn(16, V, generator=gen)

        def forward(self, ids, mask):
            m = (mask != 0).float()
            return torch.log_softmax(torch.matmul(m / m.sum(-1, keepdim=True).clamp(min=1), self.emb[ids]) @ self.out, -1)

    src, mask, size = (torch.from_numpy(z[k]) for k in ("src", "src_mask", "src_size"))
    for case in ("beam3", "beam2_lm"):
        bw, lw, lp = g[case + "::flags"]
        c = SimpleNamespace(**{**vars(cfg), "beam_width": int(bw), "lm_weight": float(lw), "length_penalty": float(lp)})
        with torch.no_grad():
            o = O.beam_decode(src, mask, size, sd, c, None, TinyLM(c.vocab_size))
        flat = [t for beams in o["beam_hyps"] for h in beams for t in h]
        assert flat == g[case + "::hyp_flat"].tolist(), case
        assert np.allclose(np.asarray(sum(o["beam_scores"], [])), g[case + "::score"], rtol=1e-5, atol=1e-5), case


@pytest.mark.parametrize("kind", ["n-gram", "at_baseline"])
def test_other_rankers_port(kind):
    """rank_model 'n-gram' / 'at_baseline' of the oracle against the hypotheses the unmodified reference produced with
    the same stand-in rankers (tests/golden/rankers.npz, made by tests/golden/make_golden_rankers.py)."""
    import os
    import sys
    from types import SimpleNamespace

    import numpy as np
    import torch

    import oracle.nat_torch as O
    from conftest import load_blocks
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from ranker_standins import TinyAT, TinyNgram, index2word
    z, sd, cfg = load_blocks("trf")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rankers.npz"))
    cfg = SimpleNamespace(**{**vars(cfg), "sample_num": 3, "threshold": 0.9, "rank_model": kind})
    src, mask, size = (torch.from_numpy(z[k]) for k in ("src", "src_mask", "src_size"))
    ranker = TinyNgram() if kind == "n-gram" else TinyAT(src.size(-1), cfg.vocab_size)
    vocab = SimpleNamespace(index2word=index2word(cfg.vocab_size))
    select = z["sampled::select"]                       # the reference's CPU-generator draw under manual_seed(555)
    with torch.no_grad():
        o = O.beam_decode(src, mask, size, sd, cfg, select, ranker, vocab)
    tag = kind.replace("-", "")
    flat, lens = g[tag + "::hyp_flat"], g[tag + "::hyp_len"]
    want, at = [], 0
    for n in lens:
        want.append(flat[at:at + n].tolist())
        at += n
    assert o["hyps"] == want
    assert np.allclose(o["scores"], g[tag + "::score"], rtol=1e-4, atol=1e-4)
