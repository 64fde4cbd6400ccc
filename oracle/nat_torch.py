"""ORACLE (test infrastructure, not product code): CPU/torch-fp32 restatement of the
floating-point part of the CASS-NAT alignment-conditioned path, written functionally
over a flat state dict that uses the reference's parameter names.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  The product path (cassnat_asr_public_b200) never does.

Parity pin: outputs are compared with fixtures produced by the unmodified reference
modules (tests/golden/make_golden_blocks.py -> tests/golden/blocks_*.npz); on CPU the
agreement is to float32 round-off of the same torch ops.

Reference code restated (paths relative to /root/reference/src/models):
  LayerNorm (unbiased std, eps on std)        modules/norm.py:6-18
  scaled dot-product attention + MHA           modules/attention.py:13-66
  relative-position MHA (Transformer-XL style) modules/attention.py:68-147
  position-wise FFN / Swish                    modules/positionff.py:6-16, modules/conformer_related.py:10-12
  conformer conv module                        modules/conformer_related.py:15-44
  pre-norm residual                            modules/utils.py:13-32
  conv front-end, abs/rel positional encoding  modules/embedding.py:10-124
  FaNAT transformer blocks                     blocks/fanat_blocks.py:9-159
  FaNAT conformer blocks                       blocks/fanat_conformer_blocks.py:9-243
  Generator, create_pe, CassNAT.forward/beam_decode   cassnat.py:91-113, 136-257, 420-637
  LabelSmoothing                               ../utils/loss.py:6-35
CTC loss is torch.nn.functional.ctc_loss (ATen; the reference calls torch.nn.CTCLoss,
cassnat.py:126,232 — third-party arithmetic, torch pinned 1.2.0 in requirements_original.txt).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import align_np as A

F32_MIN = float(np.finfo(np.float32).min)


# ----------------------------------------------------------------------------- modules
def layer_norm(x, sd, pfx, eps=1e-6):
    mu = x.mean(-1, keepdim=True)
    sigma = x.std(-1, keepdim=True)           # unbiased (N-1), eps added to std
    return sd[pfx + ".a_2"] * (x - mu) / (sigma + eps) + sd[pfx + ".b_2"]


def linear(x, sd, pfx, bias=True):
    return F.linear(x, sd[pfx + ".weight"], sd[pfx + ".bias"] if bias else None)


def swish(x):
    return x * torch.sigmoid(x)


def ffn(x, sd, pfx, act):
    return linear(act(linear(x, sd, pfx + ".w_1")), sd, pfx + ".w_2")


def _heads(x, h):
    b, n, d = x.shape
    return x.view(b, n, h, d // h).transpose(1, 2)


def mha(q_in, kv_in, mask, sd, pfx, h):
    """modules/attention.py:44-66; mask broadcastable to [B,1|U,T] (0 = masked).
    Fully masked rows become uniform over all keys (finfo.min fill, :19-21)."""
    q = _heads(linear(q_in, sd, pfx + ".linears.0"), h)
    k = _heads(linear(kv_in, sd, pfx + ".linears.1"), h)
    v = _heads(linear(kv_in, sd, pfx + ".linears.2"), h)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.size(-1))
    if mask is not None:
        # fp32: np.finfo(float32).min like the reference; under bf16 autocast (bench.py's GPU-eager comparator only)
        # the fill must be representable in the score dtype
        s = s.masked_fill(mask.unsqueeze(1) == 0, F32_MIN if s.dtype == torch.float32 else torch.finfo(s.dtype).min)
    p = F.softmax(s, dim=-1)
    ctx = torch.matmul(p, v).transpose(1, 2).contiguous().view(q_in.size(0), -1, q.size(-1) * h)
    return linear(ctx, sd, pfx + ".linears.3")


def rel_shift(bd, t_q):
    """modules/attention.py:129-132: pad one zero column, view-transpose, drop a row."""
    b, h, n, m = bd.shape
    padded = torch.cat([bd.new_zeros(b, h, n, 1), bd], dim=-1).view(b, h, m + 1, n)
    return padded[:, :, 1:].reshape(b, h, n, m)[:, :, :, :t_q]


def rel_mha(x, mask, pos_embed, sd, pfx, h):
    """modules/attention.py:91-147 (self-attention only: q=k=v=x)."""
    B, n, d = x.shape
    dk = d // h
    q = linear(x, sd, pfx + ".linears.0").view(B, n, h, dk)
    k = linear(x, sd, pfx + ".linears.1").view(B, n, h, dk)
    v = linear(x, sd, pfx + ".linears.2").view(B, n, h, dk)
    p = F.linear(pos_embed, sd[pfx + ".linear_pos.weight"]).view(-1, h, dk).unsqueeze(0)
    qu = (q + sd[pfx + ".pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[pfx + ".pos_bias_v"]).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(1, 2).transpose(-2, -1))
    bd = rel_shift(torch.matmul(qv, p.transpose(1, 2).transpose(-2, -1)), n)
    s = (ac + bd) / math.sqrt(dk)
    m = mask.unsqueeze(1) == 0
    s = s.masked_fill(m, F32_MIN)
    pr = F.softmax(s, dim=-1).masked_fill(m, 0.0)
    ctx = torch.matmul(pr, v.transpose(1, 2)).transpose(1, 2).contiguous().view(B, n, d)
    return linear(ctx, sd, pfx + ".linears.3")


def conv_module(x, sd, pfx):
    """modules/conformer_related.py:29-44 with Swish activation."""
    y = x.transpose(1, 2)
    y = F.glu(F.conv1d(y, sd[pfx + ".pointwise_conv1.weight"], sd[pfx + ".pointwise_conv1.bias"]), dim=1)
    ksz = sd[pfx + ".depthwise_conv.weight"].size(-1)
    y = F.conv1d(y, sd[pfx + ".depthwise_conv.weight"], sd[pfx + ".depthwise_conv.bias"],
                 padding=(ksz - 1) // 2, groups=y.size(1))
    y = swish(F.group_norm(y, 1, sd[pfx + ".norm.weight"], sd[pfx + ".norm.bias"]))
    y = F.conv1d(y, sd[pfx + ".pointwise_conv2.weight"], sd[pfx + ".pointwise_conv2.bias"])
    return y.transpose(1, 2)


def sinusoid_table(d_model, max_len=5000):
    """cassnat.py:91-99 / embedding.py:19-25."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0., max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def rel_pos_embed(n, table, max_rel):
    """embedding.py:52-60: rows of the (2m+1)-row table for offsets -(n-1)..(n-1), clamped."""
    r = torch.arange(-(n - 1), n, device=table.device).clamp(-max_rel, max_rel) + max_rel
    return table[r]


def generator(x, sd, pfx):
    return F.log_softmax(linear(x, sd, pfx + ".proj"), dim=-1)


# ------------------------------------------------------------------- transformer family
def _res(x, sd, pfx, fn, scale=1.0):
    return x + scale * fn(layer_norm(x, sd, pfx + ".norm"))


def trf_self_layer(x, mask, sd, p, h):
    x = _res(x, sd, p + ".sublayer.0", lambda y: mha(y, y, mask, sd, p + ".self_attn", h))
    return _res(x, sd, p + ".sublayer.1", lambda y: ffn(y, sd, p + ".feed_forward", F.relu))


def trf_src_layer(x, mem, mask, sd, p, h):
    x = _res(x, sd, p + ".sublayer.0", lambda y: mha(y, mem, mask, sd, p + ".src_attn", h))
    return _res(x, sd, p + ".sublayer.1", lambda y: ffn(y, sd, p + ".feed_forward", F.relu))


def trf_mix_layer(x, mem, src_mask, self_mask, sd, p, h):
    x = _res(x, sd, p + ".sublayer.0", lambda y: mha(y, y, self_mask, sd, p + ".self_attn", h))
    x = _res(x, sd, p + ".sublayer.1", lambda y: mha(y, mem, src_mask, sd, p + ".src_attn", h))
    return _res(x, sd, p + ".sublayer.2", lambda y: ffn(y, sd, p + ".feed_forward", F.relu))


def _n_layers(sd, pfx):
    n = 0
    while any(k.startswith(f"{pfx}.layers.{n}.") for k in sd):
        n += 1
    return n


def trf_extractor(pe_rows, mem, trigger, sd, h, pfx="acembed_extractor"):
    x = pe_rows
    for i in range(_n_layers(sd, pfx)):
        x = trf_src_layer(x, mem, trigger, sd, f"{pfx}.layers.{i}", h)
    return x


def trf_sad(x, tgt_mask, sd, h, pfx="embed_mapper", tap=None):
    for i in range(_n_layers(sd, pfx)):
        x = trf_self_layer(x, tgt_mask, sd, f"{pfx}.layers.{i}", h)
        if tap is not None and i == tap[0] - 1:
            tap[1] = x
    return x


def trf_mad(x, mem, src_mask, tgt_mask, sd, h, pfx="decoder", tap=None):
    for i in range(_n_layers(sd, pfx)):
        x = trf_mix_layer(x, mem, src_mask, tgt_mask, sd, f"{pfx}.layers.{i}", h)
        if tap is not None and i == tap[0] - 1:
            tap[1] = x
    return layer_norm(x, sd, pfx + ".norm")


def trf_encoder(x, mask, sd, h, pfx="encoder", tap=None):
    """tap = [layer, None]: tap[1] receives the (un-normalised) output of that layer (interctc, fanat_blocks.py:92-96)."""
    for i in range(_n_layers(sd, pfx)):
        x = trf_self_layer(x, mask, sd, f"{pfx}.layers.{i}", h)
        if tap is not None and i == tap[0] - 1:
            tap[1] = x
    return layer_norm(x, sd, pfx + ".norm")


# --------------------------------------------------------------------- conformer family
def con_self_layer(x, mask, pos, sd, p, h, with_src=None):
    """fanat_conformer_blocks.py:26-37 / :84-97 for pos_type == 'relative':
    1/2 FFN1 -> rel self-attn -> conv -> [src-attn] -> 1/2 FFN2."""
    x = _res(x, sd, p + ".sublayer.0", lambda y: ffn(y, sd, p + ".feed_forward1", swish), 0.5)
    x = _res(x, sd, p + ".sublayer.2", lambda y: rel_mha(y, mask, pos, sd, p + ".self_attn", h))
    x = _res(x, sd, p + ".sublayer.1", lambda y: conv_module(y, sd, p + ".conv_module"))
    if with_src is not None:
        mem, src_mask = with_src
        x = _res(x, sd, p + ".sublayer.3", lambda y: mha(y, mem, src_mask, sd, p + ".src_attn", h))
        return _res(x, sd, p + ".sublayer.4", lambda y: ffn(y, sd, p + ".feed_forward2", swish), 0.5)
    return _res(x, sd, p + ".sublayer.3", lambda y: ffn(y, sd, p + ".feed_forward2", swish), 0.5)


def con_extractor(pe_rows, mem, trigger, sd, h, max_rel, pfx="acembed_extractor"):
    """fanat_conformer_blocks.py:52-61: no pre-norm / residual around the attention."""
    p = pfx + ".layers.0"
    x = mha(pe_rows, mem, trigger, sd, p + ".src_attn", h) * math.sqrt(pe_rows.size(-1))
    pos = rel_pos_embed(x.size(1), sd[p + ".pos_enc.embedding.weight"], max_rel)
    x = _res(x, sd, p + ".sublayer", lambda y: ffn(y, sd, p + ".feed_forward", swish))
    return x, pos


def con_sad(x, pos, tgt_mask, sd, h, pfx="embed_mapper", tap=None):
    for i in range(_n_layers(sd, pfx)):
        x = con_self_layer(x, tgt_mask, pos, sd, f"{pfx}.layers.{i}", h)
        if tap is not None and i == tap[0] - 1:
            tap[1] = x
    return x, pos


def con_mad(x, pos, mem, src_mask, tgt_mask, sd, h, pfx="decoder", tap=None):
    for i in range(_n_layers(sd, pfx)):
        x = con_self_layer(x, tgt_mask, pos, sd, f"{pfx}.layers.{i}", h, with_src=(mem, src_mask))
        if tap is not None and i == tap[0] - 1:
            tap[1] = x
    return layer_norm(x, sd, pfx + ".norm")


def con_encoder(x, pos, mask, sd, h, pfx="encoder", tap=None):
    for i in range(_n_layers(sd, pfx)):
        x = con_self_layer(x, mask, pos, sd, f"{pfx}.layers.{i}", h)
        if tap is not None and i == tap[0] - 1:
            tap[1] = x
    return layer_norm(x, sd, pfx + ".norm")


# --------------------------------------------------------------------------- front-end
def conv_embedding(src, src_mask, sd, cfg, pfx="src_embed"):
    """embedding.py:112-124 (non-causal): two 3x3 stride-2 convs + ReLU, linear, *sqrt(d), PE."""
    d = cfg.d_model
    x = src.unsqueeze(1)
    x = F.relu(F.conv2d(x, sd[pfx + ".conv.0.weight"], sd[pfx + ".conv.0.bias"], stride=2, padding=1))
    x = F.relu(F.conv2d(x, sd[pfx + ".conv.2.weight"], sd[pfx + ".conv.2.bias"], stride=2, padding=1))
    b, c, t, f = x.shape
    x = linear(x.transpose(1, 2).contiguous().view(b, t, c * f), sd, pfx + ".linear_out") * math.sqrt(d)
    mask = src_mask[:, :, ::2][:, :, ::2]
    if cfg.use_conv_enc:
        pos = rel_pos_embed(t, sd[pfx + ".pos_enc.embedding.weight"], cfg.enc_max_relative_len)
        return (x, pos), mask
    return x + sinusoid_table(d)[:t].unsqueeze(0), mask


def encode(src, src_mask, sd, cfg, tap=None):
    x, mask = conv_embedding(src, src_mask, sd, cfg)
    if cfg.use_conv_enc:
        return con_encoder(x[0], x[1], mask, sd, cfg.n_head, tap=tap), mask
    return trf_encoder(x, mask, sd, cfg.n_head, tap=tap), mask


# ----------------------------------------------------------------- the conditioned path
def decoder_stack(enc_h, trigger, src_mask, tgt_mask, ymax, sd, cfg, taps=None):
    """cassnat.py:481-497 (== :191-226): PE queries -> extractor -> SAD -> MAD -> log-probs.
    Returns (att_out, ac_embed, pred_embed, dec_h).  taps = {"sad": [layer, None], "mad": [layer, None]} collect
    the intermediate hidden states of the interce losses (cassnat.py:194-197,216-221)."""
    taps = taps or {}
    bs, _, d = enc_h.shape
    h = cfg.n_head
    pe = sinusoid_table(d).to(enc_h.device).unsqueeze(0).repeat(bs, 1, 1)[:, :ymax, :]
    mem_mask = trigger if cfg.src_trigger else src_mask

    def unimask(pred):
        """cassnat.py:486-489: shift the token embeddings right by one (zeros first), causal target mask."""
        if not getattr(cfg, "use_unimask", False):
            return pred, tgt_mask
        if cfg.use_conv_dec:
            raise AttributeError("use_unimask with the conformer decoder fails in the reference "
                                 "(cassnat.py:487 calls .size on the (x, pos_embed) tuple)")
        shifted = torch.cat([torch.zeros_like(pred[:, :1]), pred[:, :-1]], dim=1)
        causal = torch.tril(torch.ones(ymax, ymax, dtype=torch.uint8, device=pred.device)).unsqueeze(0)
        return shifted, tgt_mask & causal.type_as(tgt_mask)

    if cfg.use_conv_dec:
        ac, pos = con_extractor(pe, enc_h, trigger, sd, h, cfg.dec_max_relative_len)
        pred, pos = con_sad(ac, pos, tgt_mask, sd, h, tap=taps.get("sad"))
        dec_in, dec_mask = unimask(pred)
        dec_h = con_mad(dec_in, pos, enc_h, mem_mask, dec_mask, sd, h, tap=taps.get("mad"))
    else:
        ac = trf_extractor(pe, enc_h, trigger, sd, h)
        pred = trf_sad(ac, tgt_mask, sd, h, tap=taps.get("sad"))
        dec_in, dec_mask = unimask(pred)
        dec_h = trf_mad(dec_in, enc_h, mem_mask, dec_mask, sd, h, tap=taps.get("mad"))
    return generator(dec_h, sd, "att_generator"), ac, pred, dec_h


def pick_hypotheses(att_out, ylen, ymax, sos=1):
    """cassnat.py:574-637 for beam_width == 1, lm_weight == 0: token i of utterance b is
    kept while i <= ylen[b] (note: one row past the eos row) and i < ymax."""
    ids = att_out.argmax(-1)
    score = att_out.max(-1)[0]
    hyps, scores = [], []
    for b in range(att_out.size(0)):
        n = min(int(ylen[b]) + 1, ymax)
        hyps.append([sos] + ids[b, :n].tolist())
        s = 0.0
        for i in range(n):
            s = s + score[b, i].item()
        scores.append(s)
    return hyps, scores


def beam_assemble(att_out, ylen, ymax, cfg, lm_model=None, sos=1):
    """cassnat.py:574-637 in full: per-position top-`beam_width` extension of every kept hypothesis, optional
    token-level LM fusion (the LM scores the hypothesis prefix, :600-606), sort by score + length penalty (:633-635).
    Positions i > ylen[b] leave the hypotheses of utterance b untouched.  Returns (hyps, scores) per utterance/beam."""
    bs = att_out.size(0)
    ys0 = torch.ones(1, 1).fill_(sos).long()
    tops = [[{"ys": ys0, "score": 0.0, "hyp": [sos]}] for _ in range(bs)]
    for i in range(ymax):
        all_seqs, ys, att_prob = [], [], []
        for b in range(bs):
            all_seqs.append([])
            for seq in tops[b]:
                if i > int(ylen[b]):
                    all_seqs[b].append(seq)
                    continue
                att_prob.append(att_out[b, i:i + 1, :])
                if cfg.lm_weight > 0:
                    ys.append(seq["ys"])
        if not att_prob:
            break
        local = torch.cat(att_prob, 0)
        if cfg.lm_weight > 0:
            ysb = torch.cat(ys, 0)
            n = ysb.size(-1)
            m = (ysb != cfg.padding_idx).unsqueeze(1) & torch.tril(torch.ones(n, n, dtype=torch.uint8)).unsqueeze(0).bool()
            local = local + cfg.lm_weight * lm_model(ysb, m)[:, -1, :]
        sc, idx = torch.topk(local, cfg.beam_width, dim=-1)
        s_idx = -1
        for b in range(bs):
            for seq in tops[b]:
                if i > int(ylen[b]):
                    continue
                s_idx += 1
                for j in range(cfg.beam_width):
                    tok = idx[s_idx][j]
                    nys = torch.cat([seq["ys"], tok.view(-1, 1)], -1) if cfg.lm_weight > 0 else seq["ys"]
                    all_seqs[b].append({"ys": nys, "score": seq["score"] + sc[s_idx][j].item(), "hyp": seq["hyp"] + [tok.item()]})
            key = (lambda x: x["score"] + (len(x["hyp"]) - 1) * cfg.length_penalty) if cfg.length_penalty is not None \
                else (lambda x: x["score"])
            tops[b] = sorted(all_seqs[b], key=key, reverse=True)[:cfg.beam_width]
    return [[s["hyp"] for s in t] for t in tops], [[s["score"] for s in t] for t in tops]


def decode_from_encoder(enc_h, src_mask, src_size_frames, sd, cfg, select=None, lm_model=None, vocab=None, features=None):
    """Post-encoder part of CassNAT.beam_decode (cassnat.py:433-637), greedy pick.
    src_size_frames: int64 [B] == (ratio * T).long() of :436.
    select: host-drawn bits for sampled alignments (cfg.sample_num > 1).
    Returns dict with every intermediate the parity tests compare."""
    ctc_out = generator(enc_h, sd, "ctc_generator")
    S = cfg.sample_num
    if S > 1:                                                     # :441-445
        rep = lambda x: x.unsqueeze(1).repeat(1, S, *([1] * (x.dim() - 1))).reshape(-1, *x.shape[1:])
        ctc_r, enc_h, src_mask, src_size_frames = rep(ctc_out), rep(enc_h), rep(src_mask), rep(src_size_frames)
    else:
        ctc_r = ctc_out
    aligned, ylen, ymax = A.best_path_align(ctc_r.numpy(), src_mask.numpy(), 0, S, cfg.threshold, True, select)
    trig, ylen1, ymax1 = A.align_to_mask(aligned, ylen, ymax, src_mask.numpy(), src_size_frames.numpy(), 0)
    trig = A.expand_trigger_mask(trig, cfg.left_trigger, cfg.right_trigger) & (src_mask.numpy() != 0)
    trigger = torch.from_numpy(trig)
    tgt_mask = torch.from_numpy(A.target_length_mask(ylen1, ymax1))
    att_out, ac, pred, dec_h = decoder_stack(enc_h, trigger, src_mask, tgt_mask, ymax1, sd, cfg)
    out = dict(ctc_out=ctc_out, aligned=aligned, trigger=trig, ylen=ylen1, ymax=ymax1,
               ac_embed=ac, pred_embed=pred, dec_h=dec_h, att_out=att_out)
    ylen_t = torch.from_numpy(ylen1)
    if S > 1:                                                     # :499-561
        seql, dim = att_out.shape[1:]
        att_pred = att_out.argmax(-1)
        kind = getattr(cfg, "rank_model", "lm")
        if kind == "n-gram":                                      # :525-536: text of every sample scored on the host
            tgt_len = tgt_mask.sum(-1)
            mean = torch.zeros(att_pred.size(0))
            for i in range(att_pred.size(0)):
                pieces = [vocab.index2word[int(t)] for t in att_pred[i, :int(tgt_len[i])] if int(t) != 2]
                mean[i] = lm_model.score("".join(pieces).replace("\u2581", " ").strip()) / tgt_len[i]
            mean = mean.reshape(-1, S)
        else:
            lm_in = torch.cat([torch.full((att_out.size(0), 1), 1, dtype=torch.long), att_pred[:, :-1]], 1)
            causal = torch.tril(torch.ones(ymax1, ymax1, dtype=torch.uint8)).unsqueeze(0)
            lm_mask = tgt_mask & causal.type_as(tgt_mask)
            if kind == "at_baseline":                             # :513-519: the AT model re-encodes the features
                src, x_mask = features
                x, at_mask = lm_model.src_embed(src, x_mask)
                at_h = lm_model.encoder(x, at_mask)
                at_h = at_h.unsqueeze(1).repeat(1, S, 1, 1).reshape(-1, at_h.size(1), at_h.size(2))
                at_mask = at_mask.unsqueeze(1).repeat(1, S, 1, 1).reshape(-1, at_mask.size(1), at_mask.size(2))
                lm_out = lm_model.forward_decoder(at_h, lm_in, at_mask, lm_mask)
            else:
                lm_out = lm_model(lm_in, lm_mask)
            sc = torch.gather(lm_out, -1, att_pred.unsqueeze(-1)).squeeze(-1)
            sc = sc.reshape(-1, S, seql).masked_fill(tgt_mask.reshape(-1, S, seql) == 0, 0)
            mean = sc.sum(-1) / (sc != 0).sum(-1).float()
        best = mean.max(-1, keepdim=True)[1]
        att_out = att_out.reshape(-1, S, seql, dim).masked_fill(
            tgt_mask.reshape(-1, S, 1, seql).transpose(2, 3) == 0, 0)
        att_out = torch.gather(att_out, 1, best.unsqueeze(2).unsqueeze(3).repeat(1, 1, seql, dim)).squeeze(1)
        ylen_t = ylen_t.reshape(-1, S).gather(1, best).squeeze(1)
        ymax1 = int(ylen_t.max())
        out.update(best_sample=best.squeeze(1), lm_score=mean, att_out_best=att_out)
    hyps, scores = pick_hypotheses(att_out, ylen_t, ymax1)
    out.update(hyps=hyps, scores=scores)
    if getattr(cfg, "beam_width", 1) > 1 or getattr(cfg, "lm_weight", 0) > 0:
        out["beam_hyps"], out["beam_scores"] = beam_assemble(att_out, ylen_t, ymax1, cfg, lm_model)
    return out


def beam_decode(src, src_mask, src_size, sd, cfg, select=None, lm_model=None, vocab=None):
    """Whole CassNAT.beam_decode (cassnat.py:420-637): encoder + the conditioned path."""
    enc_h, mask = encode(src, src_mask, sd, cfg)
    frames = (src_size * enc_h.size(1)).long()                   # :436
    out = decode_from_encoder(enc_h, mask, frames, sd, cfg, select, lm_model, vocab, (src, src_mask))
    out["enc_h"] = enc_h
    return out


def label_smoothing_loss(logp, target, V, padding_idx, smoothing):
    """utils/loss.py:21-35: KL(smoothed one-hot || p) summed over non-pad tokens / #tokens."""
    true = torch.full_like(logp, smoothing / (V - 1))
    keep = target != padding_idx
    true.scatter_(1, target.masked_fill(~keep, 0).unsqueeze(1), 1.0 - smoothing)
    kl = F.kl_div(logp, true, reduction="none")
    return kl.masked_fill(~keep.unsqueeze(1), 0).sum() / keep.sum().item()


def generator_norm(x, sd, pfx):
    """Generator(add_norm=True) (cassnat.py:101-113): LayerNorm, projection, log-softmax."""
    return F.log_softmax(linear(layer_norm(x, sd, pfx + ".norm"), sd, pfx + ".proj"), dim=-1)


def train_forward_from_encoder(enc_h, x_mask, src_size_frames, tgt_label, label_sizes, sd, cfg, t_sample=None, inter_h=None):
    """Post-encoder part of CassNAT.forward (cassnat.py:148-257) with use_trigger=True, use_best_path=False.
    Intermediate losses: interctc_alpha > 0 takes `inter_h` (the encoder tap, :151-153,236-239); interce_alpha > 0
    taps the SAD (selfce, :194-197) or the MAD (mixce, :216-221) after make_model's split of interce_layer (:73-83)."""
    ctc_out = generator(enc_h, sd, "ctc_generator")
    aligned, ylen, ymax = A.viterbi_align(ctc_out.detach().numpy(), x_mask.numpy(), src_size_frames.numpy(),
                                          tgt_label[:, :-1].numpy(), label_sizes.numpy(), cfg.padding_idx, t_sample)
    trig, ylen1, ymax1 = A.align_to_mask(aligned, ylen, ymax, x_mask.numpy(), src_size_frames.numpy(), cfg.padding_idx)
    trig = A.expand_trigger_mask(trig, cfg.left_trigger, cfg.right_trigger) & (x_mask.numpy() != 0)
    tgt_mask = torch.from_numpy(A.target_length_mask(ylen1, ymax1))
    taps = {}
    selfce, mixce = getattr(cfg, "selfce_alpha", 0), getattr(cfg, "mixce_alpha", 0)
    if selfce > 0:
        taps["sad"] = [cfg.interce_layer, None]
    if mixce > 0:
        taps["mad"] = [cfg.interce_layer, None]
    att_out, ac, pred, dec_h = decoder_stack(enc_h, torch.from_numpy(trig), x_mask, tgt_mask, ymax1, sd, cfg, taps)
    ctc_loss = F.ctc_loss(ctc_out.transpose(0, 1), tgt_label, src_size_frames, label_sizes,
                          blank=0, reduction="mean", zero_infinity=True)
    att_loss = label_smoothing_loss(att_out.reshape(-1, att_out.size(-1)), tgt_label.reshape(-1),
                                    cfg.vocab_size, cfg.padding_idx, cfg.label_smooth)
    loss = cfg.ctc_alpha * ctc_loss + cfg.att_alpha * att_loss
    extra = {}
    if getattr(cfg, "interctc_alpha", 0) > 0:
        ic = generator_norm(inter_h, sd, "interctc_generator")
        extra["interctc_loss"] = F.ctc_loss(ic.transpose(0, 1), tgt_label, src_size_frames, label_sizes,
                                            blank=0, reduction="mean", zero_infinity=True)
        loss = loss + cfg.interctc_alpha * extra["interctc_loss"]
    if getattr(cfg, "interce_alpha", 0) > 0:
        tap = taps["sad"][1] if selfce > 0 else taps["mad"][1]
        ie = generator_norm(tap, sd, "interce_generator")
        extra["interce_loss"] = label_smoothing_loss(ie.reshape(-1, ie.size(-1)), tgt_label.reshape(-1), cfg.vocab_size,
                                                     cfg.padding_idx, cfg.label_smooth)
        loss = loss + cfg.interce_alpha * extra["interce_loss"]
    return dict(ctc_out=ctc_out, att_out=att_out, loss=loss, ctc_loss=ctc_loss, att_loss=att_loss,
                aligned=aligned, trigger=trig, ac_embed=ac, pred_embed=pred, dec_h=dec_h, **extra)


# --------------------------------------------------------------------- LM ranker
def lm_forward(ids, mask, sd, n_head):
    """TransformerLM.forward (src/models/lm.py:52-56): TextEmbedding (embedding.py:71-78) + absolute positional
    encoding (embedding.py:29-31) -> transformer_blocks.Encoder (:51-71, same layer as the FaNAT encoder) ->
    Generator (lm.py:35-43).  ids int64 [N,U], mask [N,U,U] -> log-probabilities [N,U,V]."""
    w = sd["text_embed.0.lut.weight"]
    d = w.shape[1]
    x = w[ids] * math.sqrt(d) + sinusoid_table(d)[: ids.shape[1]].unsqueeze(0)
    return generator(trf_encoder(x, mask, sd, n_head), sd, "out_generator")
