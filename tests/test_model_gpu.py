"""GPU parity of the CassNAT mirror (make_model / beam_decode / decode_from_encoder) against the
fixtures produced by the unmodified reference (tests/golden/blocks_*.npz): alignments and token ids
bit-exact, embeddings within rel 1e-4 (north_star: 1e-3 in fp32)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from cassnat_asr_public_b200.cassnat import make_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
VOCAB = SimpleNamespace(word2index={"sos": 1, "eos": 2, "blank": 0})
TOL = 1e-4


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-9)).item()


def build(sd, cfg, **over):
    args = SimpleNamespace(**{**vars(cfg), **over})
    model = make_model(16, args)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("pos_enc.pe") for k in missing), missing
    return model.to(DEV).eval(), args


def hyps_of(z, tag):
    flat, lens = z[tag + "::hyp_flat"], z[tag + "::hyp_len"]
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n].tolist())
        o += n
    return out


def test_state_dict_names_match_reference(blocks_golden):
    name, z, sd, cfg = blocks_golden
    model = make_model(16, SimpleNamespace(**vars(cfg)))
    mine = {k for k in model.state_dict() if not k.endswith("pos_enc.pe")}
    assert mine == set(sd), (sorted(mine - set(sd))[:5], sorted(set(sd) - mine)[:5])
    for k, v in model.state_dict().items():
        if k in sd:
            assert tuple(v.shape) == tuple(sd[k].shape), k


@pytest.mark.parametrize("tag,left,right", [("greedy", 0, 0), ("greedy_l1r1", 1, 1)])
def test_decode_from_encoder(blocks_golden, tag, left, right):
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg, left_trigger=left, right_trigger=right)
    enc_h = torch.from_numpy(z[tag + "::enc_h"]).to(DEV)
    src_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2].to(DEV)
    size = torch.from_numpy(z["src_size"]).to(DEV)
    with torch.no_grad():
        out, _, keep = model.decode_from_encoder(enc_h, src_mask, size, VOCAB, args, return_intermediates=True)
    assert np.array_equal(keep["aligned"].cpu().numpy(), z[tag + "::aligned"])
    assert np.array_equal(keep["ylen"].cpu().numpy(), z[tag + "::ylen"])
    for key in ("ac_embed", "pred_embed", "dec_h"):
        assert rel(keep[key], z[f"{tag}::{key}"]) < TOL, (name, tag, key)
    att = torch.from_numpy(z[tag + "::att_out"])
    assert torch.equal(keep["ids"].cpu().long(), att.argmax(-1))
    assert (keep["logp"].cpu() - att.max(-1)[0]).abs().max().item() < 1e-4
    assert [o[0]["hyp"] for o in out] == hyps_of(z, tag)
    assert np.allclose([o[0]["score"] for o in out], z[tag + "::score"], rtol=1e-5, atol=1e-4)


def test_beam_decode_from_features(blocks_golden):
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    src, mask, size = (torch.from_numpy(z[k]).to(DEV) for k in ("src", "src_mask", "src_size"))
    with torch.no_grad():
        enc_h, m2 = model.encode(src, mask)
    assert rel(enc_h, z["greedy::enc_h"]) < TOL
    out, args2 = model.beam_decode(src, mask, size, VOCAB, args)
    assert args2 is args
    assert [o[0]["hyp"] for o in out] == hyps_of(z, "greedy")
    assert out[0][0]["ys"].tolist() == [[1]]


def test_generator_log_softmax(blocks_golden):
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    enc_h = torch.from_numpy(z["greedy::enc_h"]).to(DEV)
    ctc_out = model.ctc_generator(enc_h)
    assert (ctc_out.cpu() - torch.from_numpy(z["greedy::ctc_out"])).abs().max().item() < 1e-4


def test_sampled_decode_with_ranker(blocks_golden):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg, sample_num=3, threshold=0.9, rank_model="lm")

    class TinyLM(torch.nn.Module):          # same stand-in ranker as tests/golden/make_golden_blocks.py
        def __init__(self, V, seed=7):
            super().__init__()
            g = torch.Generator().manual_seed(seed)
            self.emb = torch.nn.Parameter(torch.randn(V, 16, generator=g), requires_grad=False)
            self.out = torch.nn.Parameter(torch.randn(16, V, generator=g), requires_grad=False)

        def forward(self, ids, mask):
            x = self.emb[ids]
            m = (mask != 0).float()
            ctx = torch.matmul(m / m.sum(-1, keepdim=True).clamp(min=1), x)
            return torch.log_softmax(ctx @ self.out, -1)

    lm = TinyLM(cfg.vocab_size).to(DEV)
    src, mask, size = (torch.from_numpy(z[k]).to(DEV) for k in ("src", "src_mask", "src_size"))
    torch.manual_seed(555)                   # make_golden_blocks.py seeds the reference the same way
    with torch.no_grad():
        enc_h, m2 = model.encode(src, mask)
        out, _, keep = model.decode_from_encoder(enc_h, m2, size, VOCAB, args, lm_model=lm, return_intermediates=True)
    assert np.array_equal(keep["aligned"].cpu().numpy(), z["sampled::aligned"])
    assert np.array_equal(keep["best_sample"].cpu().numpy(), z["sampled::best_sample"])
    assert [o[0]["hyp"] for o in out] == hyps_of(z, "sampled")


def test_deferred_hypotheses(blocks_golden):
    """decode_from_encoder(defer=True): the handle's result() equals the synchronous call's return value, also with
    two batches in flight (each keeps its own read-back buffer) and with CUDA graphs on."""
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    src, mask, size = (torch.from_numpy(z[k]).to(DEV) for k in ("src", "src_mask", "src_size"))
    with torch.no_grad():
        enc_h, m2 = model.encode(src, mask)
        enc_b, m2_b, size_b = enc_h[[1, 0, 3]].contiguous(), m2[[1, 0, 3]].contiguous(), size[[1, 0, 3]].contiguous()
        for graphs in (False, True):
            model.use_cuda_graphs = graphs
            want_a, _ = model.decode_from_encoder(enc_h, m2, size, VOCAB, args)
            want_b, _ = model.decode_from_encoder(enc_b, m2_b, size_b, VOCAB, args)
            pa, _ = model.decode_from_encoder(enc_h, m2, size, VOCAB, args, defer=True)
            pb, _ = model.decode_from_encoder(enc_b, m2_b, size_b, VOCAB, args, defer=True)
            for pend, want in ((pa, want_a), (pb, want_b)):
                got = pend.result()
                assert got is pend.result()
                assert [o[0]["hyp"] for o in got] == [o[0]["hyp"] for o in want]
                assert [o[0]["score"] for o in got] == [o[0]["score"] for o in want]
    model.use_cuda_graphs = False


@pytest.mark.parametrize("kind", ["n-gram", "at_baseline"])
def test_sampled_decode_other_rankers(kind):
    """rank_model 'n-gram' / 'at_baseline' (cassnat.py:513-536): the hypotheses the UNMODIFIED reference produced with
    the stand-in rankers of tests/golden/ranker_standins.py (tests/golden/make_golden_rankers.py), same CPU-generator
    draw for the sampled alignments."""
    import os, sys
    from conftest import load_blocks
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from ranker_standins import TinyAT, TinyNgram, index2word
    z, sd, cfg = load_blocks("trf")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rankers.npz"))
    model, args = build(sd, cfg, sample_num=3, threshold=0.9, rank_model=kind)
    src, mask, size = (torch.from_numpy(z[k]).to(DEV) for k in ("src", "src_mask", "src_size"))
    vocab = SimpleNamespace(word2index=VOCAB.word2index, index2word=index2word(cfg.vocab_size))
    ranker = TinyNgram() if kind == "n-gram" else TinyAT(src.size(-1), cfg.vocab_size).to(DEV)
    torch.manual_seed(555)
    out, _ = model.beam_decode(src, mask, size, vocab, args, lm_model=ranker)
    tag = kind.replace("-", "")
    assert [o[0]["hyp"] for o in out] == hyps_of(g, tag)
    assert np.allclose([o[0]["score"] for o in out], g[tag + "::score"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("mode", ["bf16", "bf16x3"])
def test_bf16_tensor_core_mode(blocks_golden, mode):
    """Fast mode (bf16 operands on tcgen05, fp32 accumulation) and the precise tensor-core mode (bf16x3 GEMMs).
    The CTC head runs as bf16x3 in both (decision-safe heads), so the best-path alignment is the reference's, bit for
    bit, in BOTH block families.  Stated tolerance after that: embeddings within rel 3e-2 (bf16) / 1e-2 (bf16x3: the
    attention operands stay bf16) of the fp32 reference, log-probs within 0.05, token ids equal wherever the
    reference's top-2 margin exceeds 0.1 nat."""
    from cassnat_asr_public_b200 import ops
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    enc_h = torch.from_numpy(z["greedy::enc_h"]).to(DEV)
    src_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2].to(DEV)
    size = torch.from_numpy(z["src_size"]).to(DEV)
    try:
        model.set_precision(mode)
        with torch.no_grad():
            out, _, keep = model.decode_from_encoder(enc_h, src_mask, size, VOCAB, args, return_intermediates=True)
    finally:
        model.set_precision("fp32")
    assert np.array_equal(keep["aligned"].cpu().numpy(), z["greedy::aligned"]), (name, mode)
    ylen = torch.from_numpy(z["greedy::ylen"])
    U = keep["dec_h"].shape[1]
    needed = torch.arange(U)[None, :] <= ylen[:, None]          # rows 0..ylen[b]: what the pick loop reads
    for key in ("ac_embed", "pred_embed", "dec_h"):
        got, want = keep[key].float().cpu()[needed], torch.from_numpy(z[f"greedy::{key}"])[needed]
        assert rel(got, want) < (3e-2 if mode == "bf16" else 1e-2), (name, key, mode)
    att = torch.from_numpy(z["greedy::att_out"])
    top2 = att.topk(2, -1)[0]
    clear = ((top2[..., 0] - top2[..., 1]) > 0.1) & needed
    assert torch.equal(keep["ids"].cpu().long()[clear], att.argmax(-1)[clear])
    assert (keep["logp"].cpu() - att.max(-1)[0])[clear].abs().max().item() < 0.05


def test_training_forward(blocks_golden):
    """CassNAT.forward values (cassnat.py:136-257): Viterbi alignment bit-exact, log-probs and the three
    losses within rel 1e-4, fused CTC gradient (ctc_alpha * d ctc_loss / d logits) within rel 1e-3."""
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    enc_h = torch.from_numpy(z["train::enc_h"]).to(DEV)
    x_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2].to(DEV)
    size = torch.from_numpy(z["src_size"]).to(DEV)
    labels, sizes = torch.from_numpy(z["train::labels"]), torch.from_numpy(z["train::label_sizes"])
    with torch.no_grad():
        ctc_out, att_out, loss, ctc_loss, att_loss = model.train_forward_from_encoder(enc_h, x_mask, size, labels, sizes, args)
    assert rel(ctc_out, z["train::ctc_out"]) < TOL and rel(att_out, z["train::att_out"]) < TOL
    for got, key in ((loss, "loss"), (ctc_loss, "ctc_loss"), (att_loss, "att_loss")):
        want = float(z["train::" + key])
        assert abs(float(got) - want) < 1e-4 * max(1.0, abs(want)), key
    want_grad = torch.from_numpy(z["train::dctc_dlogits"]) * args.ctc_alpha
    assert rel(model.last_ctc_grad, want_grad) < 1e-3
    # whole call from features
    src, mask = torch.from_numpy(z["src"]).to(DEV), torch.from_numpy(z["src_mask"]).to(DEV)
    with torch.no_grad():
        out = model(src, mask, size, labels, sizes, args)
    assert abs(float(out[2]) - float(z["train::loss"])) < 1e-3 * abs(float(z["train::loss"]))


def test_cuda_graph_replay_matches_eager(blocks_golden):
    """The sync-free part of the decode replayed as a CUDA graph gives the same hypotheses as eager
    launches, also when the data (not the shapes) change between replays."""
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    enc_h = torch.from_numpy(z["greedy::enc_h"]).to(DEV)
    src_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2].to(DEV)
    size = torch.from_numpy(z["src_size"]).to(DEV)
    want = hyps_of(z, "greedy")
    with torch.no_grad():
        model.use_cuda_graphs = True
        for it in range(3):                                   # eager, eager + capture (the shape came back), replay
            out, _ = model.decode_from_encoder(enc_h, src_mask, size, VOCAB, args)
            assert [o[0]["hyp"] for o in out] == want, it
            assert len(model._graphs) == (0 if it == 0 else 1)
        # different content, same shapes: compare with an eager run of the same input
        g = torch.Generator(device=DEV).manual_seed(7)
        enc2 = enc_h + 0.05 * torch.randn(enc_h.shape, device=DEV, generator=g)
        model.use_cuda_graphs = False
        ref, _ = model.decode_from_encoder(enc2, src_mask, size, VOCAB, args)
        model.use_cuda_graphs = True
        got, _ = model.decode_from_encoder(enc2, src_mask, size, VOCAB, args)
        model.use_cuda_graphs = False
    if max(len(o[0]["hyp"]) for o in ref) == max(len(o[0]["hyp"]) for o in out):     # same ymax -> same graph
        assert [o[0]["hyp"] for o in got] == [o[0]["hyp"] for o in ref]


def test_hitrate_diagnostic(blocks_golden):
    """args.test_hitrate (cassnat.py:438-462): trigger hit-rate counters of the best-path alignment against the
    forced alignment of the labels, one utterance per call, compared with the same counts from the oracle."""
    import oracle.align_np as A
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg, test_hitrate=True)
    args.total = args.num_correct = args.length_total = args.length_correct = 0
    enc_h = torch.from_numpy(z["greedy::enc_h"])
    src_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2]
    size = torch.from_numpy(z["src_size"])
    labels = torch.from_numpy(z["train::labels"])
    labels = torch.cat([torch.ones(labels.size(0), 1, dtype=torch.long), labels], 1)      # [sos, tokens..., eos]
    sizes = torch.from_numpy(z["train::label_sizes"])
    want = dict(total=0, num_correct=0, length_total=0, length_correct=0)
    T = enc_h.size(1)
    for b in range(enc_h.size(0)):
        sl = slice(b, b + 1)
        with torch.no_grad():
            model.decode_from_encoder(enc_h[sl].to(DEV), src_mask[sl].to(DEV), size[sl].to(DEV), VOCAB, args,
                                      labels=labels[sl], label_sizes=sizes[sl])
        lp = z["greedy::ctc_out"][sl]
        frames = (size[sl] * T).long().numpy()
        al, _, _ = A.best_path_align(lp, src_mask[sl].numpy(), 0)
        n = int(sizes[b])
        oa, _, _ = A.viterbi_align(lp, src_mask[sl].numpy(), frames, labels[sl, 1:1 + n].numpy(), sizes[sl].numpy(), 0, None)
        a, o = al != 0, oa != 0
        want["total"] += int(o.sum())
        want["num_correct"] += int(((a == o) & (a | o)).sum())
        want["length_total"] += 1
        want["length_correct"] += int(a.sum() == o.sum())
    got = {k: getattr(args, k) for k in want}
    assert got == want, (got, want)


class _TinyLM(torch.nn.Module):
    """Same stand-in ranker as tests/golden/make_golden_blocks.py (CPU generator, seed 7)."""

    def __init__(self, V, seed=7):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = torch.nn.Parameter(torch.randn(V, 16, generator=g), requires_grad=False)
        self.out = torch.nn.Parameter(torch.randn(16, V, generator=g), requires_grad=False)

    def forward(self, ids, mask):
        x = self.emb[ids]
        m = (mask != 0).float()
        ctx = torch.matmul(m / m.sum(-1, keepdim=True).clamp(min=1), x)
        return torch.log_softmax(ctx @ self.out, -1)


@pytest.mark.parametrize("case", ["beam3", "beam2_lm"])
def test_beam_width_and_lm_fusion_vs_reference(case):
    """Hypothesis assembly with beam_width > 1 and token-level LM fusion (cassnat.py:574-637) against the unmodified
    reference (tests/golden/beam.npz, generated on the blocks_trf fixture): every beam's hypothesis identical,
    scores within 1e-4."""
    import os
    from conftest import load_blocks
    z, sd, cfg = load_blocks("trf")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "beam.npz"))
    bw, lw, lp = g[case + "::flags"]
    model, args = build(sd, cfg, beam_width=int(bw), lm_weight=float(lw), length_penalty=float(lp))
    src, mask, size = (torch.from_numpy(z[k]).to(DEV) for k in ("src", "src_mask", "src_size"))
    lm = _TinyLM(args.vocab_size).to(DEV)
    out, _ = model.beam_decode(src, mask, size, VOCAB, args, lm_model=lm)
    flat, lens, nb, score = (g[case + "::" + k] for k in ("hyp_flat", "hyp_len", "n_beams", "score"))
    want, o, i = [], 0, 0
    for n in nb:
        beams = []
        for _ in range(int(n)):
            beams.append((flat[o:o + lens[i]].tolist(), float(score[i])))
            o += int(lens[i])
            i += 1
        want.append(beams)
    assert [len(b) for b in out] == [len(b) for b in want]
    for got_b, want_b in zip(out, want):
        for seq, (hyp, sc) in zip(got_b, want_b):
            assert seq["hyp"] == hyp
            assert abs(seq["score"] - sc) < 1e-4 * max(1.0, abs(sc))


def test_cuda_graph_length_buckets(blocks_golden):
    """model.graph_bucket: batches whose token-axis lengths fall into the same bucket replay ONE graph (transformer
    family) and return the same hypotheses as eager launches; the conformer family keeps exact lengths."""
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    enc_h = torch.from_numpy(z["greedy::enc_h"]).to(DEV)
    src_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2].to(DEV)
    size = torch.from_numpy(z["src_size"]).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(11)
    variants = [enc_h, enc_h + 0.3 * torch.randn(enc_h.shape, device=DEV, generator=g),
                enc_h + 0.6 * torch.randn(enc_h.shape, device=DEV, generator=g)]
    with torch.no_grad():
        model.use_cuda_graphs = False
        eager = [[o[0]["hyp"] for o in model.decode_from_encoder(v, src_mask, size, VOCAB, args)[0]] for v in variants]
        model.use_cuda_graphs, model.graph_bucket = True, 32
        for rep in range(2):
            got = [[o[0]["hyp"] for o in model.decode_from_encoder(v, src_mask, size, VOCAB, args)[0]] for v in variants]
            assert got == eager, (name, rep)
    lens = {max(len(h) for h in hyps) for hyps in eager}
    if name == "trf":
        assert len(model._graphs) == 1 or len(lens) == 1 or max(lens) - min(lens) >= 32
    else:
        assert len(model._graphs) == len(lens)


def test_cuda_graph_dropped_when_weights_change(blocks_golden):
    """A captured graph holds raw pointers to derived weight copies (bf16 casts, fused QKV concatenations): after an
    in-place parameter update or load_state_dict the cached graphs must be dropped, not replayed on stale weights."""
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    enc_h = torch.from_numpy(z["greedy::enc_h"]).to(DEV)
    src_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2].to(DEV)
    size = torch.from_numpy(z["src_size"]).to(DEV)
    try:
        model.set_precision("bf16")
        with torch.no_grad():
            model.use_cuda_graphs = True
            for _ in range(2):
                model.decode_from_encoder(enc_h, src_mask, size, VOCAB, args)
            assert len(model._graphs) == 1
            # new weights for the output head and one FFN (what optimizer.step() / load_state_dict do)
            g = torch.Generator(device=DEV).manual_seed(11)
            model.att_generator.proj.weight.copy_(torch.randn(model.att_generator.proj.weight.shape, device=DEV, generator=g) * 0.2)
            sd2 = {k: v.clone() for k, v in model.state_dict().items()}
            k_ff = next(k for k in sd2 if k.endswith(".w_1.weight") and k.startswith("decoder"))
            sd2[k_ff] = sd2[k_ff] * -1.5
            model.load_state_dict(sd2)
            got, _ = model.decode_from_encoder(enc_h, src_mask, size, VOCAB, args)          # must re-capture
            model.use_cuda_graphs = False
            want, _ = model.decode_from_encoder(enc_h, src_mask, size, VOCAB, args)
        assert [o[0]["hyp"] for o in got] == [o[0]["hyp"] for o in want]
        assert [o[0]["score"] for o in got] == [o[0]["score"] for o in want]
    finally:
        model.use_cuda_graphs = False
        model.set_precision("fp32")


@pytest.mark.parametrize("variant", ["fused_epilogue", "folded"])
def test_fused_layernorm_does_not_change_the_decode(blocks_golden, variant):
    """The two ways the fast mode gets rid of LayerNorm kernels -- ops.FOLD_LN (default: the norm is applied
    algebraically in the consumer GEMM's epilogue) and ops.FUSE_LN (opt-in: normalised copy written by the producer
    GEMM) -- against the same mode with stand-alone LayerNorm kernels: identical alignment, embeddings within bf16
    operand rounding, same tokens wherever the head margin is clear."""
    from cassnat_asr_public_b200 import ops
    name, z, sd, cfg = blocks_golden
    model, args = build(sd, cfg)
    enc_h = torch.from_numpy(z["greedy::enc_h"]).to(DEV)
    src_mask = torch.from_numpy(z["src_mask"])[:, :, ::2][:, :, ::2].to(DEV)
    size = torch.from_numpy(z["src_size"]).to(DEV)
    runs = {}
    try:
        model.set_precision("bf16")
        for fuse in (True, False):
            ops.FUSE_LN = fuse and variant == "fused_epilogue"
            ops.FOLD_LN = fuse and variant == "folded"
            with torch.no_grad():
                _, _, keep = model.decode_from_encoder(enc_h, src_mask, size, VOCAB, args, return_intermediates=True)
            runs[fuse] = keep
    finally:
        ops.FUSE_LN, ops.FOLD_LN = False, False
        model.set_precision("fp32")
    assert torch.equal(runs[True]["aligned"], runs[False]["aligned"])
    for key in ("ac_embed", "pred_embed", "dec_h"):
        assert rel(runs[True][key], runs[False][key]) < 1e-2, (name, key)
    att = torch.from_numpy(z["greedy::att_out"])
    top2 = att.topk(2, -1)[0]
    clear = (top2[..., 0] - top2[..., 1]) > 0.1
    assert torch.equal(runs[True]["ids"].cpu()[clear], runs[False]["ids"].cpu()[clear])
