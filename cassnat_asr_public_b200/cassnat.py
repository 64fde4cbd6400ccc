"""`CassNAT` and `make_model` with the reference's surface (src/models/cassnat.py:21-89, 115-637):
same sub-module attribute names (src_embed, encoder, acembed_extractor, embed_mapper, decoder,
ctc_generator, att_generator, pe), same state_dict keys, same helper-method signatures and the same
`beam_decode(...) -> (batch_top_seqs, args)` contract — executing the alignment-conditioned path
(everything after `enc_h`) on the sm_100a kernels of libcassnat_b200.so.

What is different in HOW: the CTC posteriors are never materialised in decoding (the head GEMM is
followed by one row-statistics pass), the trigger mask is never materialised (segment ids drive
the extractor), the MAD memory K/V projections of all layers are one GEMM, and the B x U Python
`.item()` pick loop (cassnat.py:574-637) is one device pass + one device->host copy.
"""
import copy
import math

import torch
import torch.nn as nn

from . import blocks as BK
from . import modules as MD
from . import ops


def create_pe(d_model, max_len=5000):
    """cassnat.py:91-99"""
    pe = torch.zeros(max_len, d_model, requires_grad=False)
    position = torch.arange(0., max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


class Generator(nn.Module):
    """cassnat.py:101-113: linear + log-softmax (full log-probabilities on request only)."""

    def __init__(self, d_model, vocab, add_norm=False):
        super().__init__()
        self.proj = nn.Linear(d_model, vocab)
        self.add_norm = add_norm
        if add_norm:
            self.norm = MD.LayerNorm(d_model)

    def logits(self, x, T=1.0):
        if self.add_norm:
            x = self.norm(x)
        return ops.linear(x, self.proj.weight, self.proj.bias, alpha=1.0 / T)

    def forward(self, x, T=1.0):
        z = self.logits(x, T)
        ops.row_stats(z, lse=True, logp_out=z)      # in-place log-softmax
        return z

    def greedy(self, x, want_top2=False):
        """argmax id, its log-probability (and runner-up id) per row without writing log-probs."""
        if ops.is_bf16() and not self.add_norm and x.shape[-1] % 8 == 0:
            st = ops.head_rowstats(x, self.proj.weight, self.proj.bias, top2=want_top2)    # logits stay on chip
        else:
            st = ops.row_stats(self.logits(x), top2=want_top2, peak=True, lse=True)
        st["logp"] = st["peak"] - st["lse"]
        return st


def make_model(input_size, args):
    """cassnat.py:21-89 (same args attributes read, same args mutations, same initialisation)."""
    c = copy.deepcopy
    if args.use_conv_enc:
        assert args.pos_type == "relative", "conformer must use relative positional encoding"
        enc_position = MD.RelativePositionalEncoding(args.d_model, args.dropout, args.enc_max_relative_len)
        enc_attn = MD.RelMultiHeadedAttention(args.n_head, args.d_model, args.dropout)
        enc_conv = MD.ConvModule(args.d_model, args.enc_kernel_size, activation=MD.Swish())
        enc_ff = MD.PositionwiseFeedForward(args.d_model, args.d_encff, args.dropout, activation=MD.Swish())
        encoder = BK.ConEncoder(args.d_model, c(enc_ff), enc_attn, enc_conv, c(enc_ff), args.dropout, args.N_enc,
                                args.pos_type, args.share_ff)
    else:
        assert args.model_type == "transformer"
        attn = MD.MultiHeadedAttention(args.n_head, args.d_model)
        ff = MD.PositionwiseFeedForward(args.d_model, args.d_encff, args.dropout)
        enc_position = MD.PositionalEncoding(args.d_model, args.dropout)
        encoder = BK.TrfEncoder(args.d_model, c(attn), c(ff), args.dropout, args.N_enc)

    if args.use_conv_dec:
        self_attn = MD.RelMultiHeadedAttention(args.n_head, args.d_model, args.dropout)
        src_attn = MD.MultiHeadedAttention(args.n_head, args.d_model, args.dropout)
        conv = MD.ConvModule(args.d_model, args.dec_kernel_size, activation=MD.Swish())
        position = MD.RelativePositionalEncoding(args.d_model, args.dropout, args.dec_max_relative_len)
        ff = MD.PositionwiseFeedForward(args.d_model, args.d_decff, args.dropout, activation=MD.Swish())
        ff_extra = MD.PositionwiseFeedForward(args.d_model, args.d_ff, args.dropout, activation=MD.Swish())
        extra = BK.ConAcExtra(args.d_model, c(src_attn), ff_extra, position, args.pos_type, args.dropout, args.N_extra)
        sad = BK.ConSAD(args.d_model, c(ff), c(self_attn), c(conv), c(ff), args.dropout, args.N_self_dec,
                        args.pos_type, args.share_ff)
        mad = BK.ConMAD(args.d_model, c(ff), c(self_attn), c(conv), c(src_attn), c(ff), args.dropout, args.N_mix_dec,
                        args.pos_type, args.share_ff)
    else:
        ff = MD.PositionwiseFeedForward(args.d_model, args.d_decff, args.dropout)
        attn = MD.MultiHeadedAttention(args.n_head, args.d_model, args.dropout)
        extra = BK.TrfAcExtra(args.d_model, c(attn), c(ff), args.dropout, args.N_extra)
        sad = BK.TrfSAD(args.d_model, c(attn), c(ff), args.dropout, args.N_self_dec)
        mad = BK.TrfMAD(args.d_model, c(attn), c(attn), c(ff), args.dropout, args.N_mix_dec)

    gen = Generator(args.d_model, args.vocab_size)
    interctc_gen = Generator(args.d_model, args.vocab_size, add_norm=True) if args.interctc_alpha > 0 else None
    interce_gen = Generator(args.d_model, args.vocab_size, add_norm=True) if args.interce_alpha > 0 else None
    model = CassNAT(MD.ConvEmbedding(input_size, args.d_model, args.dropout, enc_position), encoder, extra, sad, mad,
                    c(gen), c(gen), create_pe(args.d_model), interctc_gen, interce_gen, args)

    if args.interce_alpha > 0:
        if args.interce_layer <= args.N_self_dec:
            args.selfce_alpha, args.mixce_alpha = args.interce_alpha, 0
        else:
            args.selfce_alpha, args.mixce_alpha = 0, args.interce_alpha
            args.interce_layer = args.interce_layer - args.N_self_dec
    else:
        args.selfce_alpha = args.mixce_alpha = 0
    for p in model.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)
    return model


class PendingHypotheses:
    """Handle returned by decode_from_encoder(..., defer=True): the batch's device work is queued; result() waits for
    its read-back and assembles the reference's `batch_top_seqs` (once)."""

    def __init__(self, assemble, done):
        self._assemble, self._done, self._out = assemble, done, None

    def result(self):
        if self._assemble is not None:
            if self._done is not None:
                self._done.synchronize()
            self._out, self._assemble = self._assemble(), None
        return self._out


class CassNAT(nn.Module):
    def __init__(self, src_embed, encoder, acembed_extractor, embed_mapper, decoder, ctc_gen, att_gen, pe,
                 interctc_gen, interce_gen, args):
        super().__init__()
        self.src_embed = src_embed
        self.encoder = encoder
        self.acembed_extractor = acembed_extractor
        self.embed_mapper = embed_mapper
        self.decoder = decoder
        self.ctc_generator = ctc_gen
        self.att_generator = att_gen
        self.pe = pe                       # plain tensor attribute like the reference (not in state_dict)
        if interctc_gen is not None:
            self.interctc_generator = interctc_gen
        if interce_gen is not None:
            self.interce_generator = interce_gen
        self._conformer_dec = isinstance(acembed_extractor, BK.ConAcEmbedExtractor)

    # ------------------------------------------------------------------ helper methods (public API)
    def expand_trigger_mask(self, trigger_mask, left_trigger, right_trigger):
        return ops.expand_trigger_mask(trigger_mask, left_trigger, right_trigger)

    def viterbi_align(self, ctc_out, src_mask, src_size, ys, ylens, blank, sample_topk):
        return ops.viterbi_align(ctc_out, src_mask, src_size, ys, ylens, blank, sample_topk)

    def align_to_mask(self, aligned_seq_shift, ylens, ymax, src_mask, src_size, blank):
        return ops.align_to_mask(aligned_seq_shift, ylens, ymax, src_mask, src_size, blank)

    def best_path_align(self, ctc_out, src_mask, src_size, blank, sample_num=0, threshold=0.9, include_best=True):
        return ops.best_path_align(ctc_out, src_mask, src_size, blank, sample_num, threshold, include_best)

    def subsequent_mask(self, size):
        ret = torch.ones(size, size, dtype=torch.uint8)
        return torch.tril(ret, out=ret).unsqueeze(0)

    def _pe_rows(self, like, n):
        if self.pe.device != like.device:
            self.pe = self.pe.to(device=like.device)
        return self.pe[:n]

    def set_precision(self, mode):
        """'fp32' (parity mode, FFMA) or 'bf16' (tcgen05 tensor cores, fp32 accumulate); process-wide."""
        ops.set_precision(mode)
        return self

    # ------------------------------------------------------------------ the accelerated path
    def encode(self, src, x_mask):
        """src_embed + encoder (cassnat.py:431-432) -> enc_h [B,T,d], src_mask [B,1,T]."""
        x, src_mask = self.src_embed(src, x_mask)
        return self.encoder(x, src_mask), src_mask

    def decode_from_encoder(self, enc_h, src_mask, src_size, vocab, args, lm_model=None, labels=None,
                            label_sizes=None, return_intermediates=False, features=None, defer=False):
        """Everything CassNAT.beam_decode does after the encoder (cassnat.py:433-637).
        src_size is the length RATIO tensor the reference takes (frames = (ratio*T).long(), :436)."""
        if defer and return_intermediates:
            raise ValueError("defer=True returns a handle to the hypotheses only; intermediates need the synchronous call")
        beam = args.lm_weight > 0 or args.beam_width != 1        # cassnat.py:600-636: per-position top-k (+ token-level LM)
        hitrate = bool(getattr(args, "test_hitrate", False)) and args.use_trigger
        if hitrate and enc_h.size(0) != 1:
            raise ValueError("test_hitrate works on one utterance at a time (cassnat.py:459 calls .item() on a [B] tensor)")
        sos, blank = vocab.word2index["sos"], vocab.word2index["blank"]
        B, T, d = enc_h.shape
        dev = enc_h.device
        S = args.sample_num if (args.use_trigger and args.sample_num > 1) else 1
        mask2 = src_mask.reshape(B, T)
        frames = (src_size.to(dev) * T).long()                                       # :436 / :471
        keep = {}
        graphable = (self.use_cuda_graphs and args.use_trigger and S == 1 and not args.use_unimask and not beam and
                     not args.src_trigger and not getattr(args, "save_embedding", False) and not return_intermediates)
        # host work the graph replay needs but that does not depend on the token count is done here, while the device
        # is still busy with whatever was queued before this call (the replay itself has to wait for `ymax`)
        stamp = self._weight_stamp() if graphable else None
        if ops.is_bf16():
            # one bf16 copy feeds the CTC head and every K/V projection; with decision-safe heads (or in bf16x3 mode)
            # it is the hi half of a split operand whose lo half the CTC head also reads.  With graphs on it is written
            # straight into the buffer the captured graph reads (no 65 MB copy into a static input per call).
            split = ops.SAFE_HEADS or ops.is_x3()
            buf = None
            if graphable and enc_h.dtype == torch.float32 and d % 4 == 0:
                bufs = self.__dict__.setdefault("_enc_static", {})
                shape = (B, T, 2 * d if split else d)
                buf = bufs.get(shape)
                if buf is None:
                    if len(bufs) >= self._GRAPH_CACHE_SIZE:
                        bufs.pop(next(iter(bufs)))
                    buf = bufs[shape] = torch.empty(shape, dtype=torch.bfloat16, device=dev)
            enc_h = ops.cast_bf16(enc_h, split=split, out=buf)

        # ---- K1 + K2: CTC head -> per-frame argmax/top-2 -> collapsed, shifted alignment
        if args.use_trigger and args.decode_type == "oracle_att":
            ctc_out = self.ctc_generator(enc_h)
            aligned, ylen, ymax = ops.viterbi_align(ctc_out, src_mask, frames, labels[:, 1:-1].to(dev),
                                                    label_sizes.to(dev), blank, 0)
        elif args.use_trigger and args.decode_type == "ctc_att":
            raise NotImplementedError("decode_type='ctc_att' raises TypeError in the reference (cassnat.py:413)")
        else:
            st = self.ctc_generator.greedy(enc_h, want_top2=S > 1)
            aligned, ylen, ymax = self._collapse(st, mask2, B, T, S, blank, args.threshold if S > 1 else 0.9)
        if hitrate:                                                                    # :438-439, :453-462
            # trigger hit rate of the best-path alignment against the forced alignment of the reference labels
            oracle_al, _, _ = ops.viterbi_align(self.ctc_generator(enc_h), src_mask, frames, labels[:, 1:-1].to(dev),
                                                label_sizes.to(dev), blank, 0)
            if S < 2:
                a, o = aligned != 0, oracle_al != 0
                args.total += int(o.sum().item())
                args.num_correct += int(((a == o) & (a | o)).sum().item())
                args.length_total += 1
                args.length_correct += int(a.sum().item() == o.sum().item())
        if S > 1:                                                                      # :441-445
            rep = lambda x: x.repeat_interleave(S, 0)
            enc_h, src_mask, mask2, frames = rep(enc_h), rep(src_mask), rep(mask2), rep(frames)
        BS = B * S

        # ---- K4..K9: segments -> extractor -> SAD -> MAD -> head (optionally replayed as a CUDA graph)
        if beam and S > 1:
            raise NotImplementedError("beam_width > 1 / lm_weight > 0 together with sampled alignments")
        if graphable:
            ids, logp, ylen, ymax = self._stage_b_graphed(enc_h, src_mask, frames, aligned, ylen, ymax, args, blank, stamp)
            ac_embed = pred_embed = dec_h = seg = None
        else:
            ids, logp, ylen, ymax, ac_embed, pred_embed, dec_h, seg, tgt_mask = self._stage_b(
                enc_h, src_mask, frames, aligned, ylen, ymax, args, blank)
        if return_intermediates:
            first = lambda x: x[0] if isinstance(x, tuple) else x
            keep.update(aligned=aligned, seg=seg, ylen=ylen, ymax=ymax,
                        ac_embed=first(ac_embed), pred_embed=first(pred_embed), dec_h=dec_h, ids=ids, logp=logp)
        if S > 1:
            ids, logp, ylen, ymax, best = self._rescore(ids, logp, ylen, tgt_mask, lm_model, args, sos, B, S, ymax,
                                                        vocab=vocab, features=features)
            keep["best_sample"] = best
        if beam:
            out = self._assemble_beam(self.att_generator(dec_h).float(), ylen, ymax, args, lm_model, sos)
            if defer:
                return PendingHypotheses(lambda: out, None), args
            return (out, args, keep) if return_intermediates else (out, args)
        # ---- hypothesis assembly (cassnat.py:574-637 for beam_width 1): one D2H copy
        # one device->host transfer: [ids | logp bits | ylen] packed into a pinned int32 buffer
        U_ = ids.shape[1]
        packed = torch.cat([ids.reshape(B, U_).to(torch.int32), logp.reshape(B, U_).view(torch.int32),
                            ylen.reshape(B, 1).to(torch.int32)], 1)
        host = self._pinned(packed.shape, slot=self._next_slot() if defer else 0)
        host.copy_(packed, non_blocking=True)
        if defer:
            # serving loops: the device work of this batch is queued, the hypotheses are assembled when the caller
            # asks for them -- typically after queueing the next batch, so that the host-side assembly overlaps it
            done = torch.cuda.Event()
            done.record()
            return PendingHypotheses(lambda: self._assemble_greedy(host, U_, B, ymax, sos, args, dev), done), args
        torch.cuda.current_stream().synchronize()
        out = self._assemble_greedy(host, U_, B, ymax, sos, args, dev)
        if return_intermediates:
            return out, args, keep
        return out, args

    @staticmethod
    def _assemble_greedy(host, U_, B, ymax, sos, args, dev):
        """cassnat.py:574-637 for beam_width 1 on the packed [ids | logp bits | ylen] rows read back from the device."""
        arr = host.numpy()
        ids_h, ylen_h = arr[:, :U_], arr[:, 2 * U_].tolist()
        # sequential float64 running sum == the reference's `seq['score'] + token_score` chain
        csum = arr[:, U_:2 * U_].view("float32").astype("float64").cumsum(axis=1)
        ys0 = torch.ones(1, 1).fill_(sos).long()
        if getattr(args, "use_gpu", False):
            ys0 = ys0.to(dev)
        out = []
        for b in range(B):
            n = min(int(ylen_h[b]) + 1, ymax)          # the reference keeps i <= ylen[b] (one past eos)
            out.append([{"ys": ys0, "score": float(csum[b, n - 1]) if n else 0.0,
                         "hyp": [sos] + ids_h[b, :n].tolist()}])
        return out

    def _next_slot(self):
        self.__dict__["_pin_slot"] = 1 + (self.__dict__.get("_pin_slot", 0) % 2)       # slots 1, 2 (0: synchronous calls)
        return self.__dict__["_pin_slot"]

    def _assemble_beam(self, att_out, ylen, ymax, args, lm_model, sos):
        """cassnat.py:574-637 for beam_width > 1 and / or lm_weight > 0: every kept hypothesis is extended by the
        top-`beam_width` tokens of its position (plus lm_weight times the LM's next-token log-probabilities of the
        hypothesis prefix), then the best `beam_width` by score + length penalty survive.  Positions i > ylen[b] leave
        utterance b untouched.  Without LM fusion the top-k of all positions is taken in one pass and read back once."""
        B, U, _ = att_out.shape
        dev = att_out.device
        k, lw, pad = int(args.beam_width), float(args.lm_weight), args.padding_idx
        lp = getattr(args, "length_penalty", None)
        key = (lambda x: x["score"] + (len(x["hyp"]) - 1) * lp) if lp is not None else (lambda x: x["score"])
        ylen_h = ylen.tolist()
        ys0 = torch.ones(1, 1).fill_(sos).long()
        if getattr(args, "use_gpu", False):
            ys0 = ys0.to(dev)
        tops = [[{"ys": ys0, "score": 0.0, "hyp": [sos]}] for _ in range(B)]
        if lw <= 0:
            sc_all, idx_all = att_out.topk(k, -1)
            sc_all, idx_all = sc_all.cpu().numpy(), idx_all.cpu().numpy()
        for i in range(ymax):
            active = [(b, seq) for b in range(B) for seq in tops[b] if i <= ylen_h[b]]
            if not active:
                break
            if lw > 0:
                rows = torch.cat([att_out[b, i:i + 1] for b, _ in active], 0)
                ysb = torch.cat([seq["ys"].to(dev) for _, seq in active], 0)
                n = ysb.size(-1)
                m = (ysb != pad).unsqueeze(1) & torch.tril(torch.ones(n, n, dtype=torch.bool, device=dev)).unsqueeze(0)
                rows = rows + lw * lm_model(ysb, m)[:, -1, :].float()
                sc, idx = rows.topk(k, -1)
                sc, idx = sc.tolist(), idx.tolist()
            s_idx = -1
            for b in range(B):
                if i > ylen_h[b]:
                    continue
                cand = []
                for seq in tops[b]:
                    s_idx += 1
                    for j in range(k):
                        tok, ts = (idx[s_idx][j], sc[s_idx][j]) if lw > 0 else (int(idx_all[b, i, j]), float(sc_all[b, i, j]))
                        ys = torch.cat([seq["ys"], torch.tensor([[tok]], dtype=torch.long, device=seq["ys"].device)], -1) \
                            if lw > 0 else seq["ys"]
                        cand.append({"ys": ys, "score": seq["score"] + ts, "hyp": seq["hyp"] + [tok]})
                tops[b] = sorted(cand, key=key, reverse=True)[:k]
        return tops

    def _stage_b(self, enc_h, src_mask, frames, aligned, ylen, ymax, args, blank):
        """cassnat.py:465-497 after the alignment: no host synchronisation in here (CUDA-graph safe)."""
        BS, T, _ = enc_h.shape
        dev = enc_h.device
        mask2 = src_mask.reshape(BS, T)
        # ---- K4/K5: segment ids (trigger mask stays implicit), ylen+1, target-length mask
        seg = None
        if args.use_trigger:
            _, ylen, ymax, seg = ops.align_to_mask(aligned, ylen, ymax, mask2, frames, blank, want_mask=False,
                                                   want_seg=True)
            seg_ctx = (seg, ylen - 1, mask2, frames, args.left_trigger, args.right_trigger)
            dense_trigger = lambda: ops.align_to_mask(aligned, ylen - 1, ymax - 1, mask2, frames, blank,
                                                      args.left_trigger, args.right_trigger, True)[0]
        else:                                                                          # :470-473
            seg_ctx = None
        tgt_mask = ops.tgt_length_mask(ylen, ymax, as_u8=True).unsqueeze(1)            # [BS,1,U]
        if ops.is_bf16() and not args.use_unimask and self.use_length_hints:
            # Length hints for the tcgen05 attention (exact): key tiles past the last valid key carry
            # zero weight; in the transformer family token rows past ylen[b] (the pick loop reads rows
            # 0..ylen[b], cassnat.py:587) never influence a row that is read, so they are zero-filled.
            # The conformer family mixes rows (depthwise conv, GroupNorm over the padded slab): no q hint.
            if BS <= 1024:                  # one launch for every hint and both processing orders
                h = ops.decode_hints(ylen, ymax, mask2, use_q=not self._conformer_dec)
                need_rows = None if self._conformer_dec else h["need_rows"]
                tgt_mask._cn_kv_len, tgt_mask._cn_order = h["kv_tgt"], h["order_tgt"]
                if need_rows is not None:
                    tgt_mask._cn_q_len = need_rows
                if not args.src_trigger:
                    src_mask = src_mask.reshape(BS, 1, T)
                    src_mask._cn_kv_len, src_mask._cn_order = h["last_key"], h["order_src"]
                    if need_rows is not None:
                        src_mask._cn_q_len = need_rows
            else:
                need_rows = None if self._conformer_dec else (ylen + 1).clamp(max=ymax)
                ops.attach_length_hints(tgt_mask, kv_len=ylen.clamp(max=ymax), q_len=need_rows)
                if not args.src_trigger:
                    valid = mask2 != 0
                    last_key = (valid * torch.arange(1, T + 1, device=dev)).amax(1)
                    src_mask = ops.attach_length_hints(src_mask.reshape(BS, 1, T), kv_len=last_key, q_len=need_rows)
            if seg_ctx is not None and need_rows is not None:
                seg_ctx = seg_ctx + (need_rows,)                       # extractor: token rows the pick loop can read
        # ---- K6: token acoustic extractor
        pe_rows = self._pe_rows(enc_h, ymax)
        # (fast mode: every residual GEMM also emits the LayerNorm its consumer applies first -- the stacks hand each
        # other the norm that comes next; use_unimask shifts the SAD output, so no hand-over there)
        sad_norm = BK._first_norm(self.embed_mapper.layers[0])
        mad_norm = None if args.use_unimask else BK._first_norm(self.decoder.layers[0])
        if seg_ctx is not None:
            ac_embed = self.acembed_extractor.forward_segments(pe_rows, enc_h, seg_ctx, dense_trigger, next_norm=sad_norm)
        else:
            ac_embed = self.acembed_extractor(pe_rows.unsqueeze(0).repeat(BS, 1, 1), enc_h, src_mask, next_norm=sad_norm)
        # ---- K7: SAD
        pred_embed = self.embed_mapper(ac_embed, tgt_mask, next_norm=mad_norm)
        if getattr(args, "save_embedding", False) and getattr(args, "sample_num", 0) <= 1:
            first = lambda x: x[0] if isinstance(x, tuple) else x
            args.ac_embed, args.pred_embed = first(ac_embed)[0].cpu(), first(pred_embed)[0].cpu()
        # ---- K8: MAD
        dec_in = pred_embed
        if args.use_unimask:                                                           # :486-489
            if isinstance(pred_embed, tuple):
                raise NotImplementedError("use_unimask with the conformer decoder raises in the reference "
                                          "(cassnat.py:487 calls .size on a tuple)")
            dec_in = torch.cat([torch.zeros_like(pred_embed[:, :1]), pred_embed[:, :-1]], 1)
            causal = torch.tril(torch.ones(ymax, ymax, dtype=torch.uint8, device=dev)).unsqueeze(0)
            dec_mask = tgt_mask & causal
        else:
            dec_mask = tgt_mask
        mem_mask = dense_trigger() if (args.src_trigger and seg_ctx is not None) else src_mask
        dec_h = self.decoder(dec_in, enc_h, mem_mask, dec_mask)
        # ---- K9: output head + greedy pick
        st = self.att_generator.greedy(dec_h)
        return st["top1"], st["logp"], ylen, ymax, ac_embed, pred_embed, dec_h, seg, tgt_mask

    use_cuda_graphs = False          # opt-in: replay the sync-free part of the decode as a CUDA graph
    use_length_hints = True          # fast mode: skip attention tiles that cannot carry weight / are never read (exact)
    graph_bucket = 0                 # opt-in: token-axis bucket size for graph reuse across batches (0 = exact length)
    _GRAPH_CACHE_SIZE = 4

    def _stage_b_graphed(self, enc_h, src_mask, frames, aligned, ylen, ymax, args, blank, stamp=None):
        """Same as _stage_b, captured once per (batch, frames, tokens, precision) shape and replayed:
        ~150 kernel launches become one graph launch.  The first call of a shape runs eagerly (it also
        warms the weight caches); the `graph_capture_after`-th call of the same shape (default: the second) runs
        eagerly once more and captures, later calls replay."""
        # The graph is specialised on the token-axis length.  With `model.graph_bucket = n` (> 1, transformer family:
        # token rows are independent of the padded length) the length is rounded up to a multiple of n — extra rows are
        # masked and never read — so that batches with similar lengths share one graph; the conformer family
        # normalises over the padded slab (conformer_related.py:36-40) and always keeps the exact length.
        ymax_true = int(ymax)
        bucket = int(getattr(self, "graph_bucket", 0))
        if bucket > 1 and not self._conformer_dec:
            padded = -(-(ymax_true + 1) // bucket) * bucket
            if ymax_true + 1 <= 512:
                padded = min(padded, 512)                      # stay inside the resident-K/V attention variant
            ymax = padded - 1
        # A captured graph holds raw pointers to derived weight copies (bf16 casts, fused QKV / KV concatenations):
        # every graph is dropped when any parameter was updated, replaced or moved since it was captured.
        if stamp is None:
            stamp = self._weight_stamp()
        cache = self.__dict__.setdefault("_graphs", {})
        if self.__dict__.get("_graphs_stamp") != stamp:
            cache.clear()
            self.__dict__["_graphs_stamp"] = stamp
        enc_full = ops.split_of(enc_h)                          # bf16x3 / safe heads: enc_h is the hi half of [hi | lo]
        key = (tuple(enc_h.shape), str(enc_h.dtype), enc_full is not None, int(ymax), ops.get_precision(), ops.SAFE_HEADS,
               int(args.left_trigger), int(args.right_trigger), int(blank), str(src_mask.dtype))
        hit = cache.get(key)
        fix = lambda r: (r[0], r[1], r[2], ymax_true + 1)      # the pick loop needs the true (max ylens) + 1
        live = (enc_full if enc_full is not None else enc_h, src_mask, frames, aligned, ylen)
        if hit is None:
            out = self._stage_b(enc_h, src_mask, frames, aligned, ylen, ymax, args, blank)
            # A capture costs a synchronisation and a second pass over the stage: only shapes that come back are worth
            # it (with real, variable-length batches most (frames, tokens) shapes are seen once).
            seen = self.__dict__.setdefault("_graph_seen", {})
            if len(seen) > 256:
                seen.clear()
            seen[key] = seen.get(key, 0) + 1
            if seen[key] < int(getattr(self, "graph_capture_after", 2)):
                return fix(out[:4])
            # (the encoder operand usually IS a persistent buffer already, see decode_from_encoder: captured in place)
            persistent = any(live[0] is b for b in self.__dict__.get("_enc_static", {}).values())
            statics = [t if (i == 0 and persistent) else t.clone() for i, t in enumerate(live)]
            s_enc = ops._hi_view(statics[0]) if enc_full is not None else statics[0]
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            from . import _lib
            before = _lib.launch_count
            with torch.cuda.graph(graph):
                res = self._stage_b(s_enc, statics[1], statics[2], statics[3], statics[4], ymax, args, blank)
            n_kernels = _lib.launch_count - before          # our kernels recorded into the graph
            _lib.launch_count = before                       # capture launched nothing
            if len(cache) >= self._GRAPH_CACHE_SIZE:
                cache.pop(next(iter(cache)))
            cache[key] = (graph, statics, res[:4], n_kernels)
            return fix(out[:4])
        graph, statics, res, n_kernels = hit
        for dst, src in zip(statics, live):
            if dst is not src:
                dst.copy_(src, non_blocking=True)
        graph.replay()
        from . import _lib
        _lib.launch_count += n_kernels
        return fix(res)

    def _weight_stamp(self):
        """(storage address, version counter) of every parameter: changes on optimizer.step(), load_state_dict(),
        .to(), .half() ... anything after which cached derived weights (and graphs that point at them) are stale."""
        ps = self.__dict__.get("_stamp_params")
        if ps is None:                  # the parameters the captured stage reads (everything after the encoder)
            ps = self.__dict__["_stamp_params"] = [p for n, p in self.named_parameters()
                                                   if not n.startswith(("src_embed.", "encoder.", "ctc_generator."))]
        return hash(tuple([p._version for p in ps] + [p.data_ptr() for p in ps]))

    def _pinned(self, shape, slot=0):
        """Reusable pinned int32 staging buffer for the hypothesis read-back (deferred calls alternate between two
        buffers of their own: at most two batches are in flight)."""
        bufs = self.__dict__.setdefault("_pin_bufs", {})
        buf = bufs.get(slot)
        n = int(shape[0]) * int(shape[1])
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 1 << 16), dtype=torch.int32, pin_memory=True)
            bufs[slot] = buf
        return buf[:n].view(*shape)

    def _collapse(self, st, mask2, B, T, S, blank, threshold):
        dev = mask2.device
        top1, top2, peak, select = st["top1"], st.get("top2"), None, None
        if S > 1:
            # same CPU-generator draw as cassnat.py:373, on the [B*S, T] replicated batch
            select = torch.randint(0, 2, (B * S, T, 1)).to(torch.uint8).reshape(B * S, T).to(dev, non_blocking=True)
            rep = lambda x: x.reshape(B, T).repeat_interleave(S, 0).contiguous()
            top1, top2, peak = rep(top1), rep(top2), rep(st["logp"])
            mask2 = mask2.repeat_interleave(S, 0)
        mask8 = (mask2 != 0).to(torch.uint8).contiguous()
        BS = B * S
        aligned = torch.empty(BS, T, dtype=torch.int64, device=dev)
        ylen = torch.empty(BS, dtype=torch.int64, device=dev)
        ymax = torch.empty(1, dtype=torch.int32, device=dev)
        from ._lib import call, ptr, stream
        call("cn_bestpath_collapse", ptr(top1.contiguous()), ptr(top2), ptr(peak), ptr(select), ptr(mask8), BS, T,
             int(blank), int(S if S > 1 else 0), float(threshold), 1, ptr(aligned), ptr(ylen), ptr(ymax), stream())
        return aligned, ylen, int(ymax.item())          # the one host sync the API needs (cassnat.py:387)

    def _rescore(self, ids, logp, ylen, tgt_mask, lm_model, args, sos, B, S, ymax, vocab=None, features=None):
        """cassnat.py:499-561: the S sampled hypotheses of an utterance are ranked by `lm_model`, the caller's stock
        module, and the best one is kept.  rank_model 'lm' (a language model on the predicted ids,
        :510-512), 'at_baseline' (an autoregressive ASR model re-encoding the utterance's features and scoring the
        ids with its decoder, :513-519; `features` = the (src, x_mask) beam_decode was called with) or 'n-gram'
        (`lm_model.score(sentence)` on the detokenised text, host side like the reference, :525-536).
        K11 = gather / mean / first arg-max."""
        if args.rank_model == "n-gram":
            return self._rescore_ngram(ids, logp, ylen, tgt_mask, lm_model, vocab, B, S)
        if args.rank_model not in ("lm", "at_baseline"):
            raise NotImplementedError(f"rank_model={args.rank_model!r}")
        if args.rank_model == "at_baseline" and features is None:
            raise ValueError("rank_model='at_baseline' re-encodes the features: call beam_decode, or pass features=(src, x_mask)")
        ids64 = ids.long()
        lm_in = torch.cat([torch.full((B * S, 1), sos, dtype=torch.long, device=ids.device), ids64[:, :-1]], 1)
        causal = torch.tril(torch.ones(ymax, ymax, dtype=torch.uint8, device=ids.device)).unsqueeze(0)
        lm_mask = tgt_mask & causal
        if args.rank_model == "at_baseline":
            src, x_mask = features
            x, at_mask = lm_model.src_embed(src, x_mask)                                   # :515-519
            at_h = lm_model.encoder(x, at_mask)
            at_h = at_h.unsqueeze(1).repeat(1, S, 1, 1).reshape(-1, at_h.size(1), at_h.size(2))
            at_mask = at_mask.unsqueeze(1).repeat(1, S, 1, 1).reshape(-1, at_mask.size(1), at_mask.size(2))
            lm_out = lm_model.forward_decoder(at_h, lm_in, at_mask, lm_mask.to(at_mask.dtype))
        elif hasattr(lm_model, "score_tokens") and S <= 32:
            # this package's LM mirror: log p(id) per position straight from the fused head epilogue (fast mode),
            # the [B*S,U,V] tensor is never materialised; K11 then runs on the [B*S,U] scores (V = 1, id 0)
            sc = lm_model.score_tokens(lm_in, lm_mask, ids, causal=True).float().contiguous()
            from ._lib import call, ptr, stream
            mean = torch.empty(B * S, dtype=torch.float32, device=ids.device)
            best = torch.empty(B, dtype=torch.int64, device=ids.device)
            zero_ids = torch.zeros(B * S, ymax, dtype=torch.int32, device=ids.device)
            call("cn_rescore_select", ptr(sc), sc.stride(0), 1, ptr(zero_ids), ptr(tgt_mask.reshape(B * S, ymax).contiguous()),
                 B, S, ymax, ptr(mean), ptr(best), stream())
            best = best.unsqueeze(1)
            lm_out = None
        else:
            lm_out = lm_model(lm_in, lm_mask.long())
        if lm_out is None:
            pass
        elif lm_out.is_cuda and lm_out.dtype == torch.float32 and lm_out.stride(-1) == 1 and S <= 32:
            from ._lib import call, ptr, stream                                            # K11
            mean = torch.empty(B * S, dtype=torch.float32, device=ids.device)
            best = torch.empty(B, dtype=torch.int64, device=ids.device)
            ids_c, tm_c = ids.contiguous(), tgt_mask.reshape(B * S, ymax).contiguous()
            call("cn_rescore_select", ptr(lm_out), lm_out.stride(0), lm_out.stride(1), ptr(ids_c),
                 ptr(tm_c), B, S, ymax, ptr(mean), ptr(best), stream())
            best = best.unsqueeze(1)
        else:
            valid = tgt_mask.reshape(B, S, ymax) != 0
            sc = torch.gather(lm_out, -1, ids64.unsqueeze(-1)).squeeze(-1).reshape(B, S, ymax).masked_fill(~valid, 0)
            mean = sc.sum(-1) / (sc != 0).sum(-1).float()
            best = mean.max(-1, keepdim=True)[1]                                        # [B,1]
        pick = lambda x: x.reshape(B, S, -1).gather(1, best.unsqueeze(-1).expand(B, 1, x.shape[-1])).squeeze(1)
        # the reference zeroes att_out beyond ylen before picking (:536) -> id 0 / score 0 there
        ids = pick(ids.masked_fill(tgt_mask.squeeze(1) == 0, 0))
        logp = pick(logp.masked_fill(tgt_mask.squeeze(1) == 0, 0.0))
        ylen = ylen.reshape(B, S).gather(1, best).squeeze(1)
        return ids, logp, ylen, int(ylen.max().item()), best.squeeze(1)

    def _rescore_ngram(self, ids, logp, ylen, tgt_mask, lm_model, vocab, B, S):
        """cassnat.py:525-536: every sampled hypothesis is detokenised on the host (eos pieces dropped, sentencepiece
        word marks turned into blanks) and scored by the caller's n-gram model; score / length, first maximum."""
        tgt_len = tgt_mask.reshape(B * S, -1).sum(-1).cpu()
        ids_h = ids.cpu().tolist()
        prob_sum = torch.zeros(B * S)
        for i in range(B * S):
            n = int(tgt_len[i])
            pieces = [vocab.index2word[t] for t in ids_h[i][:n] if t != 2]
            prob_sum[i] = lm_model.score("".join(pieces).replace("\u2581", " ").strip()) / tgt_len[i]
        best = prob_sum.reshape(B, S).max(-1, keepdim=True)[1].to(ids.device)                # [B,1]
        pick = lambda x: x.reshape(B, S, -1).gather(1, best.unsqueeze(-1).expand(B, 1, x.shape[-1])).squeeze(1)
        ids = pick(ids.masked_fill(tgt_mask.squeeze(1) == 0, 0))
        logp = pick(logp.masked_fill(tgt_mask.squeeze(1) == 0, 0.0))
        ylen = ylen.reshape(B, S).gather(1, best).squeeze(1)
        return ids, logp, ylen, int(ylen.max().item()), best.squeeze(1)

    def beam_decode(self, src, x_mask, src_size, vocab, args, lm_model=None, ctc_top_seqs=None, labels=None,
                    label_sizes=None):
        """cassnat.py:420-637 -> (batch_top_seqs, args)."""
        with torch.no_grad():
            enc_h, src_mask = self.encode(src, x_mask)
            return self.decode_from_encoder(enc_h, src_mask, src_size, vocab, args, lm_model, labels, label_sizes,
                                            features=(src, x_mask))

    def forward(self, src, src_mask, src_size, tgt_label, label_sizes, args):
        """cassnat.py:136-257 -> (ctc_out, att_out, loss, ctc_loss, att_loss).  With grad mode on and
        parameters requiring grad, `loss` is connected to the autograd tape and `loss.backward()`
        (cassnat_task.py:291) runs on this package's backward kernels (grad.py); `ctc_out` / `att_out`
        are the log-probabilities the caller decodes for WER and carry no gradient of their own.
        Under torch.no_grad() the same values are produced without a tape, and
        `self.last_ctc_grad` = ctc_alpha * d ctc_loss / d ctc logits (K10) is kept for inspection."""
        ops.dropout_begin_step()
        x, x_mask = self.src_embed(src, src_mask)
        enc = self.encoder(x, x_mask, args.interctc_alpha, args.interctc_layer)         # :139-143
        enc_h, inter_h = enc if args.interctc_alpha > 0 else (enc, None)                  # :151
        return self.train_forward_from_encoder(enc_h, x_mask, src_size, tgt_label, label_sizes, args, inter_h=inter_h,
                                               _begin_step=False)

    def zero_dropout(self):
        """Sets every dropout probability to 0 (deterministic training step, e.g. for gradient parity tests: the
        reference's transformer-family make_model hard-codes 0.1 for the encoder's attention dropout whatever
        args.dropout says, so even its dropout=0 configuration is random there).  With p > 0 the modules apply the
        stateless-mask dropout of ops.dropout wherever the reference applies nn.Dropout."""
        for m in self.modules():
            if hasattr(m, "p_drop"):
                m.p_drop = 0.0
        return self

    def train_forward_from_encoder(self, enc_h, x_mask, src_size, tgt_label, label_sizes, args, inter_h=None, _begin_step=True):
        """Post-encoder part of CassNAT.forward (cassnat.py:148-257): CTC posteriors, Viterbi forced
        alignment, trigger segments, extractor -> SAD -> MAD -> log-probs, CTC + label-smoothing loss, plus the
        optional intermediate losses: interctc on `inter_h` (the encoder tap, :151-153,236-239) and interce on the
        SAD (selfce) or MAD (mixce) tap (:194-197,216-221,253-255)."""
        if not args.use_trigger or args.use_best_path:
            raise NotImplementedError("training needs use_trigger with the Viterbi alignment (use_best_path is broken "
                                      "upstream: cassnat.py:168,246)")
        if _begin_step:
            ops.dropout_begin_step()                 # (CassNAT.forward starts the step before the encoder runs)
        assert args.ctc_alpha > 0
        selfce, mixce = getattr(args, "selfce_alpha", 0), getattr(args, "mixce_alpha", 0)
        if args.interctc_alpha > 0 and inter_h is None:
            raise ValueError("interctc_alpha > 0 needs the encoder's intermediate output (inter_h)")
        if selfce > 0 and not self._conformer_dec:
            raise NotImplementedError("selfce with the transformer decoder: the reference hands the MAD a 1-tuple there "
                                      "(cassnat.py:196) and fails")
        B, T, d = enc_h.shape
        dev = enc_h.device
        blank = args.padding_idx
        tgt_label, label_sizes = tgt_label.to(dev), label_sizes.to(dev)
        frames = (src_size.to(dev) * T).long()                                          # :165
        mask2 = x_mask.reshape(B, T)
        taped = torch.is_grad_enabled() and (enc_h.requires_grad or any(p.requires_grad for p in self.parameters()))
        mem = ops.cast_bf16(enc_h) if ops.is_bf16() else enc_h
        if taped:
            from . import grad
            ctc_out, ctc_loss = grad.CtcFromLogits.apply(self.ctc_generator.logits(mem), tgt_label, frames,
                                                         label_sizes, blank)                # :148 + :232 (K1, K10)
        else:
            ctc_out = self.ctc_generator(mem)                                           # :148 (K1, materialised)
        aligned, ylen, ymax = ops.viterbi_align(ctc_out, x_mask, frames, tgt_label[:, :-1], label_sizes, blank,
                                                args.sample_topk)                       # :170 (K3)
        _, ylen, ymax, seg = ops.align_to_mask(aligned, ylen, ymax, mask2, frames, blank, want_mask=False,
                                               want_seg=True)                           # :172-175 (K4, implicit mask)
        seg_ctx = (seg, ylen - 1, mask2, frames, args.left_trigger, args.right_trigger)
        dense_trigger = lambda: ops.align_to_mask(aligned, ylen - 1, ymax - 1, mask2, frames, blank,
                                                  args.left_trigger, args.right_trigger, True)[0]
        tgt_mask = ops.tgt_length_mask(ylen, ymax, as_u8=True).unsqueeze(1)             # :185-187 (K5)
        ac_embed = self.acembed_extractor.forward_segments(self._pe_rows(enc_h, ymax), mem, seg_ctx, dense_trigger)
        inter_tap = None
        if selfce > 0:                                                                  # :193-197 (conformer: (x, pos, tap))
            sad = self.embed_mapper(ac_embed, tgt_mask, selfce, args.interce_layer)
            pred_embed, inter_tap = sad[:-1], sad[-1]
        else:
            pred_embed = self.embed_mapper(ac_embed, tgt_mask)
        dec_in, dec_mask = pred_embed, tgt_mask
        if args.use_unimask:                                                            # :205-208
            if isinstance(pred_embed, tuple):
                raise NotImplementedError("use_unimask with the conformer decoder raises in the reference (cassnat.py:206)")
            dec_in = torch.cat([torch.zeros_like(pred_embed[:, :1]), pred_embed[:, :-1]], 1)
            dec_mask = tgt_mask & torch.tril(torch.ones(ymax, ymax, dtype=torch.uint8, device=dev)).unsqueeze(0)
        mem_mask = dense_trigger() if args.src_trigger else x_mask
        if mixce > 0:                                                                   # :216-221
            dec_h, inter_tap = self.decoder(dec_in, mem, mem_mask, dec_mask, mixce, args.interce_layer)
        else:
            dec_h = self.decoder(dec_in, mem, mem_mask, dec_mask)
        if taped and getattr(args, "fused_ce", False) and ops.is_bf16() and not ops.is_x3():
            # opt-in: head + label-smoothing loss without the [B,U,V] logits / log-probs; the
            # call then returns att_out = None and leaves the argmax ids in self.last_att_pred (token accuracy)
            proj = self.att_generator.proj
            hd = self.att_generator.norm(dec_h) if self.att_generator.add_norm else dec_h
            att_loss, self.last_att_pred = grad.LinearCE.apply(hd.reshape(-1, hd.size(-1)), proj.weight, proj.bias,
                                                               tgt_label.reshape(-1), args.padding_idx, args.label_smooth)
            self.last_att_pred = self.last_att_pred.view(tgt_label.shape)
            att_out = None
        elif taped:
            z = self.att_generator.logits(dec_h)                                        # :226 + :248
            att_out, att_loss = grad.LabelSmoothingFromLogits.apply(z.reshape(-1, z.size(-1)), tgt_label.reshape(-1),
                                                                    args.padding_idx, args.label_smooth)
            att_out = att_out.view(z.shape)
        else:
            att_out = self.att_generator(dec_h)                                         # :226
            ctc_loss, _, g = ops.ctc_loss(ctc_out, tgt_label, frames, label_sizes, blank,
                                          want_grad=True, grad_out=float(args.ctc_alpha))  # :232 (K10)
            self.last_ctc_grad = g
            att_loss = label_smoothing_loss(att_out.view(-1, att_out.size(-1)), tgt_label.reshape(-1),
                                            args.vocab_size, args.padding_idx, args.label_smooth)   # :248
        loss = args.ctc_alpha * ctc_loss + args.att_alpha * att_loss
        if args.interctc_alpha > 0:                                                     # :236-239
            if taped:
                _, ic_loss = grad.CtcFromLogits.apply(self.interctc_generator.logits(inter_h), tgt_label, frames,
                                                      label_sizes, blank)
            else:
                ic_loss = ops.ctc_loss(self.interctc_generator(inter_h), tgt_label, frames, label_sizes, blank,
                                       want_grad=False)[0]
            loss = loss + args.interctc_alpha * ic_loss
        if args.interce_alpha > 0:                                                      # :253-255
            if taped:
                zi = self.interce_generator.logits(inter_tap)
                _, ie_loss = grad.LabelSmoothingFromLogits.apply(zi.reshape(-1, zi.size(-1)), tgt_label.reshape(-1),
                                                                 args.padding_idx, args.label_smooth)
            else:
                io = self.interce_generator(inter_tap)
                ie_loss = label_smoothing_loss(io.view(-1, io.size(-1)), tgt_label.reshape(-1), args.vocab_size,
                                               args.padding_idx, args.label_smooth)
            loss = loss + args.interce_alpha * ie_loss
        return ctc_out, att_out, loss, ctc_loss, att_loss


def label_smoothing_loss(logp, target, size, padding_idx, smoothing):
    """LabelSmoothing.forward of src/utils/loss.py:21-35 in closed form (no [N,V] true_dist clone):
    KL(smoothed one-hot || p) summed over non-pad tokens / #tokens.  Adjacent to the accelerated
    path: a handful of torch reductions over the already materialised log-probs."""
    keep = target != padding_idx
    tokens = keep.sum()
    tgt = target.masked_fill(~keep, 0)
    lp_t = logp.gather(1, tgt.unsqueeze(1)).squeeze(1)
    conf, eps = 1.0 - smoothing, smoothing / (size - 1)
    const = (conf * math.log(conf) if conf > 0 else 0.0) + ((size - 1) * eps * math.log(eps) if eps > 0 else 0.0)
    per_tok = const - conf * lp_t - eps * (logp.sum(1) - lp_t)
    return per_tok.masked_fill(~keep, 0).sum() / tokens
