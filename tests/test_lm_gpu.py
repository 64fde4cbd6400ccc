"""GPU: the LM ranker mirror (cassnat_asr_public_b200/lm.py) against the unmodified reference's TransformerLM
(tests/golden/lm.npz): same state_dict keys, log-probabilities within rel 1e-4 (fp32 mode) / 0.05 nat (bf16
mode), and the fused `score_tokens` == gather of the full log-probabilities."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from cassnat_asr_public_b200 import lm as LM
from cassnat_asr_public_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _load():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lm.npz"))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    h, d, dff, V, N = (int(v) for v in z["args"])
    args = SimpleNamespace(n_head=h, d_model=d, d_ff=dff, dropout=0.0, vocab_size=V, N=N)
    model = LM.make_model(args)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(".pe") for k in missing), (missing, unexpected)
    return z, model.to(DEV).eval()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", None)])
def test_lm_matches_reference(precision, tol):
    z, model = _load()
    ids, mask = torch.from_numpy(z["ids"]).to(DEV), torch.from_numpy(z["mask"]).to(DEV)
    want = torch.from_numpy(z["lm_out"])
    ops.set_precision(precision)
    try:
        with torch.no_grad():
            got = model(ids, mask).float().cpu()
            tgt = torch.roll(ids, -1, 1)
            sc = model.score_tokens(ids, mask, tgt).float().cpu()
            sc_causal = model.score_tokens(ids, mask.clone(), tgt, causal=True).float().cpu()   # tile-skipping hint
    finally:
        ops.set_precision("fp32")
    valid = torch.from_numpy(z["mask"]).diagonal(dim1=1, dim2=2)            # positions that can see themselves
    if tol is not None:
        assert ((got - want).abs().max() / want.abs().max()).item() < tol
    else:
        assert (got - want)[valid].abs().max().item() < 0.05
    assert torch.equal(sc[valid], sc_causal[valid])
    ref_sc = want.gather(-1, tgt.cpu().unsqueeze(-1)).squeeze(-1)
    assert (sc - ref_sc)[valid].abs().max().item() < (1e-3 if tol is not None else 0.05)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampled_decode_with_lm_mirror_matches_generic_ranker_path(precision):
    """Sampled-alignment decode (cassnat.py:499-561) with the LM mirror as ranker: the fused `score_tokens` branch of
    `_rescore` picks the same samples and returns the same hypotheses as the generic branch that materialises the
    [B*S,U,V] log-probabilities of the same LM."""
    from cassnat_asr_public_b200.cassnat import make_model
    a = dict(use_conv_enc=False, use_conv_dec=False, model_type="transformer", pos_type="absolute", n_head=2, d_model=128,
             d_encff=256, d_decff=256, d_ff=256, dropout=0.0, N_enc=1, N_extra=1, N_self_dec=1, N_mix_dec=2, vocab_size=60,
             interctc_alpha=0, interce_alpha=0, interctc_layer=0, interce_layer=0, padding_idx=0, label_smooth=0.1,
             share_ff=False, enc_max_relative_len=20, enc_kernel_size=7, dec_max_relative_len=8, dec_kernel_size=3,
             use_trigger=True, use_best_path=False, sample_topk=0, ctc_alpha=0.5, att_alpha=1.0, save_embedding=False,
             use_unimask=False, src_trigger=False, left_trigger=0, right_trigger=0, test_hitrate=False, sample_num=4,
             decode_type="att_only", threshold=0.9, rank_model="lm", lm_weight=0, beam_width=1, length_penalty=0, use_gpu=True)
    args = SimpleNamespace(**a)
    torch.manual_seed(3)
    model = make_model(16, args).to(DEV).eval()
    _, lm = _load()

    class Plain(torch.nn.Module):                      # same LM without the fused entry point
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, ids, mask):
            return self.inner(ids, mask)

    g = torch.Generator().manual_seed(9)
    B, T = 3, 40
    enc_h = torch.randn(B, T, 128, generator=g).to(DEV)
    lens = torch.tensor([40, 25, 33])
    mask = (torch.arange(T)[None] < lens[:, None]).unsqueeze(1).to(DEV)
    ratio = ((lens.float() + 0.5) / T).to(DEV)
    vocab = SimpleNamespace(word2index={"sos": 1, "eos": 2, "blank": 0})
    model.set_precision(precision)
    try:
        outs = []
        for ranker in (lm, Plain(lm)):
            torch.manual_seed(555)                         # same host-drawn select bits (cassnat.py:373)
            with torch.no_grad():
                out, _, keep = model.decode_from_encoder(enc_h, mask, ratio, vocab, args, lm_model=ranker, return_intermediates=True)
            outs.append(([o[0]["hyp"] for o in out], keep["best_sample"].cpu().tolist()))
    finally:
        model.set_precision("fp32")
    if precision == "fp32":
        assert outs[0] == outs[1]
    else:               # bf16: the two heads round differently (fused epilogue vs materialised logits); near-ties may flip
        assert len(outs[0][0]) == B and len(outs[1][0]) == B
