"""Transformer language model used as the ranker of the sampled-alignment decode: mirror of the
reference's `src/models/lm.py` (`make_model`, `TransformerLM`, identical state_dict keys) on this package's
kernels.  `forward(tgt, tgt_mask)` returns the [N,U,V] log-probabilities like the reference; `score_tokens`
returns only log p(ids) per position — in the fast mode straight out of the fused head epilogue
(`cn_gemm_bf16_rowstats` with a target column), so the [N,U,V] tensor the reference gathers from
(`cassnat.py:508,517`) is never materialised."""
import copy
import math

import torch
import torch.nn as nn

from . import blocks as BK
from . import modules as MD
from . import ops
from .cassnat import Generator


class TextEmbedding(nn.Module):
    """embedding.py:71-78"""

    def __init__(self, d_model, vocab):
        super().__init__()
        self.lut = nn.Embedding(vocab, d_model)
        self.d_model = d_model

    def forward(self, x):
        return self.lut(x) * math.sqrt(self.d_model)


class TransformerLM(nn.Module):
    """lm.py:45-56"""

    def __init__(self, hidden_size, text_embed, encoder, out_gen):
        super().__init__()
        self.dim = hidden_size
        self.text_embed = text_embed
        self.encoder = encoder
        self.out_generator = out_gen

    def set_precision(self, mode):
        ops.set_precision(mode)
        return self

    def _hidden(self, tgt, tgt_mask, causal=False):
        if causal and ops.is_bf16():
            ops.mark_causal(tgt_mask)
        return self.encoder(self.text_embed(tgt), tgt_mask)

    def forward(self, tgt, tgt_mask):
        return self.out_generator(self._hidden(tgt, tgt_mask))

    def score_tokens(self, tgt, tgt_mask, ids, causal=False):
        """log p(ids[n,u] | tgt[n,:u+1]) for every position: [N,U] f32 (what cassnat.py:517 gathers).
        causal=True promises tgt_mask is zero above the diagonal (mask & tril), which lets the fast-mode attention
        skip those key tiles."""
        h = self._hidden(tgt, tgt_mask, causal)
        gen = self.out_generator
        if ops.is_bf16() and not gen.add_norm and h.shape[-1] % 8 == 0:
            st = ops.head_rowstats(h, gen.proj.weight, gen.proj.bias, target=ids, safe=False)   # a score, not a decision
            return (st["target_logit"].reshape(-1) - st["lse"].reshape(-1)).reshape(ids.shape)
        return gen(h).gather(-1, ids.long().unsqueeze(-1)).squeeze(-1)

    def _subsequent_mask(self, size):
        return torch.tril(torch.ones(size, size, dtype=torch.uint8)).unsqueeze(0)


def make_model(args):
    """lm.py:16-32: reads args.{n_head,d_model,d_ff,dropout,vocab_size,N}."""
    c = copy.deepcopy
    attn = MD.MultiHeadedAttention(args.n_head, args.d_model)
    ff = MD.PositionwiseFeedForward(args.d_model, args.d_ff, args.dropout)
    position = MD.PositionalEncoding(args.d_model, args.dropout)
    model = TransformerLM(args.d_model, nn.Sequential(TextEmbedding(args.d_model, args.vocab_size), c(position)),
                          BK.Encoder(args.d_model, c(attn), c(ff), args.dropout, args.N),
                          Generator(args.d_model, args.vocab_size))
    for p in model.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)
    return model
