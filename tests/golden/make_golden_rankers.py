"""Fixtures for rank_model = 'n-gram' and 'at_baseline' (cassnat.py:513-536) from the UNMODIFIED reference (build
container only): the transformer-family model and inputs of make_golden_blocks.py (same seeds; its weights are checked
against blocks_trf.npz), three sampled alignments per utterance, the stand-in rankers of ranker_standins.py.
Output: rankers.npz (best sample per utterance and the hypotheses, per ranker)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_harness import VOCAB, load_reference, make_args  # noqa: E402
from ranker_standins import TinyAT, TinyNgram, index2word  # noqa: E402


def main():
    M = load_reference()
    z = np.load(os.path.join(HERE, "blocks_trf.npz"))
    args = make_args()
    torch.manual_seed(4321)
    model = M.make_model(16, args).eval()
    for k, v in model.state_dict().items():
        if not k.endswith("pos_enc.pe"):
            assert np.array_equal(v.numpy(), z["sd::" + k]), k
    src, src_mask, src_size = (torch.from_numpy(z[k]) for k in ("src", "src_mask", "src_size"))
    B, S = src.size(0), 3
    vocab = SimpleNamespace(word2index=VOCAB.word2index, index2word=index2word(args.vocab_size))
    out = {}
    for kind, ranker in (("n-gram", TinyNgram()), ("at_baseline", TinyAT(src.size(-1), args.vocab_size))):
        a = make_args(sample_num=S, threshold=0.9, rank_model=kind)
        picked = {}
        if kind == "n-gram":
            # the reference keeps max_indices local; recover it from the scores it hands the ranker, in call order
            scores = []
            inner = ranker.score
            ranker.score = lambda s_, inner=inner: (scores.append(inner(s_)), scores[-1])[1]
        torch.manual_seed(555)
        with torch.no_grad():
            res, _ = model.beam_decode(src, src_mask, src_size, vocab, a, lm_model=ranker)
        hyps = [r_[0]["hyp"] for r_ in res]
        # the oracle port must agree before anything is saved (same draw of the sampling bits)
        import oracle.nat_torch as O
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        with torch.no_grad():
            o = O.beam_decode(src, src_mask, src_size, sd, SimpleNamespace(**vars(a)), z["sampled::select"],
                              TinyNgram() if kind == "n-gram" else ranker, vocab)
        assert o["hyps"] == hyps, (kind, o["hyps"], hyps)
        tag = kind.replace("-", "") + "::"
        out[tag + "hyp_flat"] = np.asarray([t for h in hyps for t in h], dtype=np.int64)
        out[tag + "hyp_len"] = np.asarray([len(h) for h in hyps], dtype=np.int64)
        out[tag + "score"] = np.asarray([r_[0]["score"] for r_ in res], dtype=np.float64)
        if kind == "n-gram":
            out[tag + "ranker_scores"] = np.asarray(scores, dtype=np.float64)        # [B*S] in utterance-major order
        print(kind, hyps)
    path = os.path.join(HERE, "rankers.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
