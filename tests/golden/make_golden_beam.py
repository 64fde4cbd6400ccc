"""Golden fixture for the hypothesis assembly with beam_width > 1 and token-level LM fusion (cassnat.py:574-637):
the unmodified reference's beam_decode on the blocks_trf fixture (same seed -> same weights and inputs) with
(beam_width, lm_weight, length_penalty) in {(3, 0, 0), (2, 0.3, 0.1)} -> tests/golden/beam.npz (hyps and scores of
every beam).  Asserts oracle == reference before saving."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_harness import VOCAB, load_reference, make_args  # noqa: E402
from make_golden_blocks import TinyLM  # noqa: E402

CASES = {"beam3": dict(beam_width=3, lm_weight=0, length_penalty=0),
         "beam2_lm": dict(beam_width=2, lm_weight=0.3, length_penalty=0.1)}


def main():
    import oracle.nat_torch as O
    M = load_reference()
    z = np.load(os.path.join(HERE, "blocks_trf.npz"))
    src, src_mask, src_size = (torch.from_numpy(z[k]) for k in ("src", "src_mask", "src_size"))
    out = {}
    for name, flags in CASES.items():
        args = make_args(**flags)
        torch.manual_seed(4321)
        model = M.make_model(16, args).eval()
        sd = model.state_dict()
        assert all(np.array_equal(sd[k[4:]].numpy(), z[k]) for k in z.files if k.startswith("sd::")), "weights differ from blocks_trf"
        lm = TinyLM(args.vocab_size)
        with torch.no_grad():
            res, _ = model.beam_decode(src, src_mask, src_size, VOCAB, args, lm_model=lm)
            o = O.beam_decode(src, src_mask, src_size, {k: v.detach() for k, v in sd.items()}, SimpleNamespace(**vars(args)),
                              None, lm)
        hyps = [[s["hyp"] for s in beams] for beams in res]
        scores = [[s["score"] for s in beams] for beams in res]
        assert o["beam_hyps"] == hyps, (name, o["beam_hyps"], hyps)
        assert np.allclose(np.asarray(sum(o["beam_scores"], [])), np.asarray(sum(scores, [])), rtol=1e-5, atol=1e-5)
        flat = [t for beams in hyps for h in beams for t in h]
        out[name + "::hyp_flat"] = np.asarray(flat, dtype=np.int64)
        out[name + "::hyp_len"] = np.asarray([len(h) for beams in hyps for h in beams], dtype=np.int64)
        out[name + "::n_beams"] = np.asarray([len(beams) for beams in hyps], dtype=np.int64)
        out[name + "::score"] = np.asarray(sum(scores, []), dtype=np.float64)
        out[name + "::flags"] = np.asarray([flags["beam_width"], flags["lm_weight"], flags["length_penalty"]], dtype=np.float64)
        print(name, [len(b) for b in hyps], scores[0])
    path = os.path.join(HERE, "beam.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
