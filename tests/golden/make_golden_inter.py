"""Golden fixtures for the intermediate losses and use_unimask of the training call (cassnat.py:148-257): the
unmodified reference's CassNAT.forward with interctc_alpha / interce_alpha > 0 (self- and mix-decoder taps) and
with use_unimask, both block families where the reference supports them -> tests/golden/inter.npz
(weights, inputs, the five returned values per case).  Asserts oracle == reference before saving."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_harness import load_reference, make_args  # noqa: E402

FAMILIES = {
    "trf": dict(),
    "con": dict(use_conv_dec=True, use_conv_enc=True, pos_type="relative", enc_kernel_size=7),
}
# One model per family, built WITH the intermediate generators (interctc_alpha / interce_alpha > 0 at construction);
# the cases below only change forward-time flags.  make_model splits interce_layer into the self- / mix-decoder tap
# (cassnat.py:73-83: layer <= N_self_dec -> selfce, else mixce with layer - N_self_dec); the same rule is applied
# here by hand for the cases that move the tap.  N_self_dec = 1, N_mix_dec = 2 in these fixtures.
BUILD = dict(interctc_alpha=0.3, interctc_layer=1, interce_alpha=0.2, interce_layer=3)
CASES = {
    "trf": {"interctc_mixce": dict(), "unimask": dict(use_unimask=True, interctc_alpha=0, interce_alpha=0, mixce_alpha=0)},
    "con": {"interctc_selfce": dict(selfce_alpha=0.2, mixce_alpha=0, interce_layer=1), "mixce": dict(interctc_alpha=0)},
}


def main():
    import copy
    import json

    import oracle.nat_torch as O
    M = load_reference()
    out = {}
    g = torch.Generator().manual_seed(99)
    B, Tin, Fdim, L = 4, 96, 16, 5
    src = torch.randn(B, Tin, Fdim, generator=g)
    lens = [96, 60, 80, 44]
    for b, n in enumerate(lens):
        src[b, n:] = 0
    src_mask = (src[:, :, 0] != 0).unsqueeze(1)
    src_size = torch.tensor([n / Tin for n in lens])
    labels = torch.zeros(B, L + 1, dtype=torch.long)
    label_sizes = torch.tensor([5, 3, 4, 2])
    for b in range(B):
        n = int(label_sizes[b])
        labels[b, :n] = torch.randint(3, 120, (n,), generator=g)
        labels[b, n] = 2
    out.update(src=src.numpy(), src_mask=src_mask.numpy(), src_size=src_size.numpy(), labels=labels.numpy(),
               label_sizes=label_sizes.numpy())
    for fam, over in FAMILIES.items():
        args_build = make_args(**over, **BUILD)
        torch.manual_seed(4321)
        model = M.make_model(16, args_build).eval()    # mutates args_build: selfce_alpha / mixce_alpha / interce_layer
        for m_ in model.modules():
            if isinstance(m_, torch.nn.Dropout):
                m_.p = 0.0
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        out.update({f"{fam}::sd::" + k: v.numpy() for k, v in sd.items() if not k.endswith("pos_enc.pe")})
        model.train()
        for case, flags in CASES[fam].items():
            args = copy.copy(args_build)
            for k, v in flags.items():
                setattr(args, k, v)
            model.zero_grad()
            ctc_out, att_out, loss, ctc_loss, att_loss = model(src, src_mask, src_size, labels, label_sizes, args)
            cfg = SimpleNamespace(**vars(args))
            tap = [args.interctc_layer, None] if args.interctc_alpha > 0 else None
            with torch.no_grad():
                enc_h, x_mask = O.encode(src, src_mask, sd, cfg, tap=tap)
                frames = (src_size * enc_h.size(1)).long()
                o = O.train_forward_from_encoder(enc_h, x_mask, frames, labels, label_sizes, sd, cfg,
                                                 inter_h=tap[1] if tap else None)
            for a, b_, what in ((o["ctc_out"], ctc_out, "ctc_out"), (o["att_out"], att_out, "att_out"), (o["loss"], loss, "loss")):
                err = (a - b_.detach()).abs().max().item() / max(b_.detach().abs().max().item(), 1e-6)
                assert err < 2e-5, (fam, case, what, err)
            pre = f"{fam}::{case}::"
            out[pre + "cfg_json"] = np.frombuffer(json.dumps(vars(args)).encode(), dtype=np.uint8)
            out[pre + "ctc_out"] = ctc_out.detach().numpy()
            out[pre + "att_out"] = att_out.detach().numpy()
            out[pre + "losses"] = np.asarray([float(loss.detach()), float(ctc_loss.detach()), float(att_loss.detach())])
            loss.backward()
            probe = [k for k, p in model.named_parameters() if p.grad is not None and ("generator" in k or "norm" in k)]
            for k in probe:
                out[pre + "grad::" + k] = dict(model.named_parameters())[k].grad.numpy().copy()
            print(fam, case, "loss", float(loss.detach()), "selfce/mixce/layer", args.selfce_alpha, args.mixce_alpha,
                  args.interce_layer, "grads", len(probe))
    path = os.path.join(HERE, "inter.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
