"""Loads the UNMODIFIED reference (/root/reference) for golden-vector generation.

Only usable in the build container (the GPU box has no /root/reference).  Nothing
under tests/ imports this at test time; it is used by make_golden.py only.
Shims: stub `editdistance`, skip models/__init__ (fairseq),
identity Tensor.cuda on CPU (cassnat.py:361 hard-codes .cuda()).
"""
import sys
import types
from types import SimpleNamespace

import torch

REF_SRC = "/root/reference/src"


def load_reference():
    if "models.cassnat" in sys.modules:
        return sys.modules["models.cassnat"]
    ed = types.ModuleType("editdistance")
    ed.eval = lambda a, b: 0
    sys.modules["editdistance"] = ed
    sys.path.insert(0, REF_SRC)
    pkg = types.ModuleType("models")
    pkg.__path__ = [REF_SRC + "/models"]
    sys.modules["models"] = pkg
    torch.Tensor.cuda = lambda self, *a, **k: self
    import models.cassnat as M  # noqa
    return M


def make_args(**kw):
    a = dict(
        use_conv_enc=False, use_conv_dec=False, model_type="transformer", pos_type="absolute",
        n_head=2, d_model=128, d_encff=256, d_decff=256, d_ff=256, dropout=0.0,
        N_enc=1, N_extra=1, N_self_dec=1, N_mix_dec=2, vocab_size=120,
        interctc_alpha=0, interce_alpha=0, interctc_layer=0, interce_layer=0,
        padding_idx=0, label_smooth=0.1, share_ff=False,
        enc_max_relative_len=20, enc_kernel_size=31, dec_max_relative_len=8, dec_kernel_size=3,
        use_trigger=True, use_best_path=False, sample_topk=0, ctc_alpha=0.5, att_alpha=1.0,
        save_embedding=False, use_unimask=False, src_trigger=False, left_trigger=0, right_trigger=0,
        test_hitrate=False, sample_num=0, decode_type="att_only", threshold=0.9, rank_model="lm",
        lm_weight=0, beam_width=1, length_penalty=0, use_gpu=False,
    )
    a.update(kw)
    return SimpleNamespace(**a)


VOCAB = SimpleNamespace(word2index={"sos": 1, "eos": 2, "blank": 0})
