"""Block-level / end-to-end fixtures from the UNMODIFIED reference (build container only).

For each block family (transformer, conformer) builds a tiny reference CassNAT
(d=128, h=2 -> d_k=64, V=120), runs beam_decode and forward on seeded input and stores
the weights (reference parameter names), the inputs and every intermediate tensor of the
alignment-conditioned path, captured with forward hooks on the reference's own modules.
Each tensor is also compared with the oracle port (oracle/nat_torch.py) before saving.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_harness import VOCAB, load_reference, make_args  # noqa: E402


def _capture(model, names):
    store, hooks = {}, []
    for n in names:
        mod = getattr(model, n)
        hooks.append(mod.register_forward_hook(lambda m, i, o, n=n: store.__setitem__(n, o)))
    return store, hooks


def _first(x):
    return x[0] if isinstance(x, tuple) else x


def _close(a, b, tol=2e-5, what=""):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)
    assert err < tol, (what, err)


class TinyLM(torch.nn.Module):
    """Stand-in ranker with the TransformerLM call contract (ids, mask) -> log-probs.
    Deterministic and cheap; the real LM is a 'next' row."""

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


def gen_family(name, **over):
    import oracle.nat_torch as O
    M = load_reference()
    args = make_args(**over)
    torch.manual_seed(4321)
    model = M.make_model(16, args).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    out = {"sd::" + k: v.numpy() for k, v in sd.items() if not k.endswith("pos_enc.pe")}
    cfg = SimpleNamespace(**vars(args))

    g = torch.Generator().manual_seed(99)
    B, Tin, Fdim = 4, 96, 16
    src = torch.randn(B, Tin, Fdim, generator=g)
    lens = [96, 60, 80, 44]
    for b, L in enumerate(lens):
        src[b, L:] = 0
    src_mask = (src[:, :, 0] != 0).unsqueeze(1)
    src_size = torch.tensor([L / Tin for L in lens])
    out.update(src=src.numpy(), src_mask=src_mask.numpy(), src_size=src_size.numpy())

    names = ["encoder", "ctc_generator", "acembed_extractor", "embed_mapper", "decoder", "att_generator"]
    # ------------------------------------------------------------- greedy decode
    for (l, r, tag) in [(0, 0, "greedy"), (1, 1, "greedy_l1r1")]:
        args_d = make_args(**over, left_trigger=l, right_trigger=r)
        store, hooks = _capture(model, names)
        with torch.no_grad():
            res, _ = model.beam_decode(src, src_mask, src_size, VOCAB, args_d)
        for hk in hooks:
            hk.remove()
        cfg_d = SimpleNamespace(**vars(args_d))
        with torch.no_grad():
            o = O.beam_decode(src, src_mask, src_size, sd, cfg_d)
        hyps = [r_[0]["hyp"] for r_ in res]
        scores = [r_[0]["score"] for r_ in res]
        _close(o["enc_h"], store["encoder"], what="enc_h")
        _close(o["ctc_out"], store["ctc_generator"], what="ctc_out")
        _close(o["ac_embed"], _first(store["acembed_extractor"]), what="ac_embed")
        _close(o["pred_embed"], _first(store["embed_mapper"]), what="pred_embed")
        _close(o["dec_h"], store["decoder"], what="dec_h")
        _close(o["att_out"], store["att_generator"], what="att_out")
        assert o["hyps"] == hyps, (o["hyps"], hyps)
        _close(o["scores"], scores, 1e-5, "scores")
        pre = tag + "::"
        out[pre + "enc_h"] = store["encoder"].numpy()
        out[pre + "ctc_out"] = store["ctc_generator"].numpy()
        out[pre + "aligned"] = o["aligned"]
        out[pre + "trigger"] = np.packbits(o["trigger"], axis=-1)
        out[pre + "ylen"] = o["ylen"]
        out[pre + "ac_embed"] = _first(store["acembed_extractor"]).numpy()
        out[pre + "pred_embed"] = _first(store["embed_mapper"]).numpy()
        out[pre + "dec_h"] = store["decoder"].numpy()
        out[pre + "att_out"] = store["att_generator"].numpy()
        out[pre + "hyp_flat"] = np.asarray([t for h in hyps for t in h], dtype=np.int64)
        out[pre + "hyp_len"] = np.asarray([len(h) for h in hyps], dtype=np.int64)
        out[pre + "score"] = np.asarray(scores, dtype=np.float64)

    # ------------------------------------------------------------- sampled decode + LM ranker
    S = 3
    args_s = make_args(**over, sample_num=S, threshold=0.9, rank_model="lm")
    lm = TinyLM(args.vocab_size)
    torch.manual_seed(555)
    with torch.no_grad():
        res, _ = model.beam_decode(src, src_mask, src_size, VOCAB, args_s, lm_model=lm)
    torch.manual_seed(555)
    T = out["greedy::enc_h"].shape[1]
    select = torch.randint(0, 2, (B * S, T, 1))[..., 0]
    with torch.no_grad():
        o = O.beam_decode(src, src_mask, src_size, sd, SimpleNamespace(**vars(args_s)), select.numpy(), lm)
    hyps = [r_[0]["hyp"] for r_ in res]
    assert o["hyps"] == hyps, (o["hyps"], hyps)
    out["sampled::select"] = select.numpy().astype(np.uint8)
    out["sampled::aligned"] = o["aligned"]
    out["sampled::best_sample"] = o["best_sample"].numpy()
    out["sampled::hyp_flat"] = np.asarray([t for h in hyps for t in h], dtype=np.int64)
    out["sampled::hyp_len"] = np.asarray([len(h) for h in hyps], dtype=np.int64)

    # ------------------------------------------------------------- training forward/backward
    L = 5
    labels = torch.zeros(B, L + 1, dtype=torch.long)
    label_sizes = torch.tensor([5, 3, 4, 2])
    for b in range(B):
        n = int(label_sizes[b])
        labels[b, :n] = torch.randint(3, args.vocab_size, (n,), generator=g)
        labels[b, n] = 2
    model.train()   # dropout = 0 in these configs
    store, hooks = _capture(model, names)

    def _keep(key):
        def hook(m, i, o_):
            o_.retain_grad()
            store[key] = o_
        return hook
    hooks.append(model.ctc_generator.proj.register_forward_hook(_keep("ctc_logits")))
    hooks.append(model.att_generator.proj.register_forward_hook(_keep("att_logits")))
    ctc_out, att_out, loss, ctc_loss, att_loss = model(src, src_mask, src_size, labels, label_sizes, args)
    loss.backward()
    for hk in hooks:
        hk.remove()
    # Parameter gradients of the unmodified reference (whole model, encoder included) for the
    # backward-pass parity tests; kept in their own file.  The transformer-family make_model builds
    # the encoder's MultiHeadedAttention with its default dropout=0.1 regardless of args.dropout
    # (cassnat.py make_model), so a .train() forward is random there; this run zeroes every Dropout.p
    # (module attributes only, no code change) to make loss.backward() reproducible.
    for m_ in model.modules():
        if isinstance(m_, torch.nn.Dropout):
            m_.p = 0.0
    model.zero_grad()
    _, _, loss_d, ctc_d, att_d = model(src, src_mask, src_size, labels, label_sizes, args)
    loss_d.backward()
    grads = {"grad::" + k: p.grad.numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
    grads["no_grad"] = np.asarray([k for k, p in model.named_parameters() if p.grad is None])
    grads["loss"] = loss_d.detach().numpy()
    grads["ctc_loss"] = ctc_d.detach().numpy()
    grads["att_loss"] = att_d.detach().numpy()
    gpath = os.path.join(HERE, f"train_grads_{name}.npz")
    np.savez_compressed(gpath, **grads)
    print(gpath, os.path.getsize(gpath) // 1024, "KiB", len(grads) - 4, "tensors")
    enc_h = store["encoder"].detach()
    x_mask = src_mask[:, :, ::2][:, :, ::2]
    frames = (src_size * enc_h.size(1)).long()
    o = O.train_forward_from_encoder(enc_h, x_mask, frames, labels, label_sizes, sd, cfg)
    _close(o["ctc_out"], ctc_out, what="train ctc_out")
    _close(o["att_out"], att_out, what="train att_out")
    _close(o["loss"], loss, 1e-5, "loss")
    out.update({"train::labels": labels.numpy(), "train::label_sizes": label_sizes.numpy(),
                "train::enc_h": enc_h.numpy(), "train::frames": frames.numpy(),
                "train::ctc_out": ctc_out.detach().numpy(), "train::att_out": att_out.detach().numpy(),
                "train::aligned": o["aligned"], "train::loss": loss.detach().numpy(),
                "train::ctc_loss": ctc_loss.detach().numpy(), "train::att_loss": att_loss.detach().numpy(),
                "train::dloss_dctc_logits": store["ctc_logits"].grad.numpy(),
                "train::dloss_datt_logits": store["att_logits"].grad.numpy()})
    # ctc-only gradient wrt logits (what the fused CTC kernel must reproduce)
    lg = store["ctc_logits"].detach().clone().requires_grad_(True)
    cl = torch.nn.functional.ctc_loss(torch.log_softmax(lg, -1).transpose(0, 1), labels, frames, label_sizes,
                                      blank=0, reduction="mean", zero_infinity=True)
    cl.backward()
    out["train::ctc_logits"] = lg.detach().numpy()
    out["train::dctc_dlogits"] = lg.grad.numpy()
    out["cfg_json"] = np.frombuffer(__import__("json").dumps(vars(args)).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f"blocks_{name}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


def gen_blocks():
    gen_family("trf")
    gen_family("con", use_conv_dec=True, use_conv_enc=True, pos_type="relative", enc_kernel_size=7)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    gen_blocks()
