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
