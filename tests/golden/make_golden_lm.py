"""Golden fixture for the LM ranker: the unmodified reference `models.lm.make_model` on a small
configuration -> weights, token ids, mask and log-probabilities (tests/golden/lm.npz)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_harness import load_reference  # noqa: E402


def main():
    load_reference()
    import models.lm as LM
    args = SimpleNamespace(n_head=2, d_model=128, d_ff=256, dropout=0.0, vocab_size=60, N=2)
    torch.manual_seed(77)
    model = LM.make_model(args).eval()
    g = torch.Generator().manual_seed(5)
    N, U = 5, 23
    ids = torch.randint(1, 60, (N, U), generator=g)
    lens = torch.tensor([23, 9, 17, 1, 20])
    key = (torch.arange(U)[None] < lens[:, None]).unsqueeze(1)
    mask = key & torch.tril(torch.ones(U, U, dtype=torch.uint8)).unsqueeze(0).bool()
    with torch.no_grad():
        out = model(ids, mask)
    z = {"sd::" + k: v.numpy() for k, v in model.state_dict().items() if not k.endswith(".pe")}
    z.update(ids=ids.numpy(), mask=mask.numpy(), lm_out=out.numpy(), args=np.asarray([2, 128, 256, 60, 2]))
    path = os.path.join(HERE, "lm.npz")
    np.savez_compressed(path, **z)
    print(path, os.path.getsize(path) // 1024, "KiB", sorted(model.state_dict())[:4])


if __name__ == "__main__":
    main()
