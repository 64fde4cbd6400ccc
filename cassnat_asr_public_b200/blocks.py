"""FaNAT blocks (token acoustic extractor, self-attention decoder "SAD", mixed-attention decoder
"MAD", encoder stack), mirroring the class names, constructor signatures, parameter names and
forward contracts of the reference's `src/models/blocks/fanat_blocks.py` (Trf*) and
`fanat_conformer_blocks.py` (Con*), executing on the sm_100a kernels.

Differences in HOW (not what): residual adds, the conformer 1/2 FFN scale and bias+activation are
fused into GEMM epilogues; Q/K/V (and K/V of the memory) are single fused GEMMs; the extractor has
an extra `forward_segments` entry that consumes segment ids instead of a dense trigger mask.
"""
import copy
import math

import torch
import torch.nn as nn

from . import ops
from .modules import LayerNorm, SublayerConnection, _p


def clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


def _residual(sub, x, branch, scale=1.0, next_norm=None, **kw):
    """x + scale * dropout(branch(...)) of SublayerConnection (modules/utils.py:32).  Without active dropout the add (and
    the next LayerNorm) is fused into the branch's last GEMM; with dropout the branch output is dropped first."""
    p = _p(sub)
    if p > 0:
        return ops.dropout_add(branch(residual=None, scale=1.0, **kw), x, p, scale)     # one launch (was dropout, mul, add)
    return branch(residual=x, scale=scale, next_ln=next_norm, **kw)


def _self_attention_block(x, sub, attn, mask, pos_embed=None, next_norm=None):
    """x + self_attn(norm(x)) with the residual (and, in fast mode, the LayerNorm `next_norm` that consumes the
    result next) fused into the output projection.  sub: the SublayerConnection (norm + dropout probability)."""
    norm = sub.norm
    y = norm.for_gemm(x)
    if pos_embed is None:
        q, k, v = attn.project_qkv(y, **ops.attn_kw())            # bf16 (or split bf16) -> tcgen05 attention
        ctx = ops.attention(q, k, v, mask, attn.h, p_drop=_p(attn))
    else:
        from .modules import distinct_rel_rows
        table, m = distinct_rel_rows(pos_embed, x.size(1))
        key_mask = mask is None or mask.dim() == 2 or mask.shape[1] == 1        # (per-row masks: the fp32 kernel)
        fast = ops.attn_bf16() and x.size(1) <= 2048 and attn.h <= 8 and m <= ops.REL_TC_MAX and key_mask
        if fast and not (torch.is_grad_enabled() and (y.requires_grad or attn.pos_bias_u.requires_grad)) and _p(attn) == 0:
            # fast: q + pos_bias_u in bf16, and the relative term's table out of the same GEMM
            q, k, v, tab = attn.project_qkv_rel(y, table, m)
            ctx = ops.attention(q, k, v, mask, attn.h, (None, attn.pos_bias_u, attn.pos_bias_v, m, tab))
        else:
            rel_pos = ops.linear(table, attn.linear_pos.weight)
            if torch.is_grad_enabled() and (y.requires_grad or attn.pos_bias_u.requires_grad):
                q, k, v = attn.project_qkv(y, out_bf16=fast)           # training: raw q, grad.RelAttention adds u
            else:
                q, k, v = attn.project_qkv_rel(y) if fast else attn.project_qkv(y)     # fast: q + pos_bias_u, bf16
            ctx = ops.attention(q, k, v, mask, attn.h, (rel_pos, attn.pos_bias_u, attn.pos_bias_v, m), p_drop=_p(attn))
    return _residual(sub, x, lambda **kw: attn.output(ctx, **kw), next_norm=next_norm)


def _src_attention_block(x, sub, attn, memory, mask, kv=None, next_norm=None):
    y = sub.norm.for_gemm(x)
    q = attn.project_q(y, **ops.attn_kw())
    k, v = kv if kv is not None else attn.project_kv(memory, **ops.attn_kw())
    ctx = ops.attention(q, k, v, mask, attn.h, p_drop=_p(attn))
    return _residual(sub, x, lambda **kw: attn.output(ctx, **kw), next_norm=next_norm)


def _ffn_block(x, sub, ff, scale=1.0, next_norm=None):
    y = sub.norm.for_gemm(x)
    return _residual(sub, x, lambda **kw: ff(y, **kw), scale=scale, next_norm=next_norm)


def _first_norm(layer):
    """The LayerNorm a layer applies to its input first (sublayer[0].norm / sublayer.norm), or None."""
    sub = getattr(layer, "sublayer", None)
    if sub is None:
        return None
    return sub[0].norm if isinstance(sub, nn.ModuleList) else sub.norm


def _next_norms(layers, after=None):
    """For every layer of a stack: the norm that consumes its output (the next layer's first norm; `after` for
    the last layer: the stack's final norm or the first norm of whatever follows)."""
    return [(_first_norm(layers[n + 1]) if n + 1 < len(layers) else after) for n in range(len(layers))]


# =============================================================================== transformer family
class SelfAttLayer(nn.Module):
    """fanat_blocks.py:9-22"""

    def __init__(self, size, self_attn, feed_forward, dropout):
        super().__init__()
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.sublayer = clones(SublayerConnection(size, dropout), 2)
        self.size = size

    def forward(self, x, self_mask, next_norm=None):
        x = _self_attention_block(x, self.sublayer[0], self.self_attn, self_mask, next_norm=self.sublayer[1].norm)
        return _ffn_block(x, self.sublayer[1], self.feed_forward, next_norm=next_norm)


class SrcAttLayer(nn.Module):
    """fanat_blocks.py:24-37"""

    def __init__(self, size, src_attn, feed_forward, dropout):
        super().__init__()
        self.src_attn = src_attn
        self.feed_forward = feed_forward
        self.sublayer = clones(SublayerConnection(size, dropout), 2)
        self.size = size

    def forward(self, x, m, src_mask, next_norm=None):
        x = _src_attention_block(x, self.sublayer[0], self.src_attn, m, src_mask, next_norm=self.sublayer[1].norm)
        return _ffn_block(x, self.sublayer[1], self.feed_forward, next_norm=next_norm)

    def forward_segments(self, pe_rows, m, seg_ctx, kv=None, next_norm=None):
        """Same result as forward(pe_rows repeated over the batch, m, trigger_mask) with the
        trigger mask given implicitly by `seg_ctx` = (seg, ylens, src_mask, src_size, left, right)."""
        attn = self.src_attn
        fast = ops.attn_bf16() and attn.d_k == 64 and m.size(1) <= 512              # tcgen05 attention, mask fused
        q_rows = attn.project_q(self.sublayer[0].norm.for_gemm(pe_rows), out_bf16=fast)   # batch independent [U,d]
        k, v = kv if kv is not None else attn.project_kv(m, out_bf16=fast)
        ctx = ops.segment_attention(q_rows, k, v, *seg_ctx[:6], attn.h, q_len=seg_ctx[6] if len(seg_ctx) > 6 else None)
        x = attn.output(ctx, residual=pe_rows, res_rows=pe_rows.size(0), next_ln=self.sublayer[1].norm)
        return _ffn_block(x, self.sublayer[1], self.feed_forward, next_norm=next_norm)


class MixAttLayer(nn.Module):
    """fanat_blocks.py:39-55"""

    def __init__(self, size, self_attn, src_attn, feed_forward, dropout):
        super().__init__()
        self.size = size
        self.self_attn = self_attn
        self.src_attn = src_attn
        self.feed_forward = feed_forward
        self.sublayer = clones(SublayerConnection(size, dropout), 3)

    def forward(self, x, memory, src_mask, self_mask, kv=None, next_norm=None):
        x = _self_attention_block(x, self.sublayer[0], self.self_attn, self_mask, next_norm=self.sublayer[1].norm)
        x = _src_attention_block(x, self.sublayer[1], self.src_attn, memory, src_mask, kv, next_norm=self.sublayer[2].norm)
        return _ffn_block(x, self.sublayer[2], self.feed_forward, next_norm=next_norm)


class Encoder(nn.Module):
    """fanat_blocks.py:78-99 (upstream of the accelerated path; same kernels as the SAD)."""

    def __init__(self, size, self_attn, feed_forward, dropout, N):
        super().__init__()
        layer = SelfAttLayer(size, self_attn, feed_forward, dropout)
        self.layers = clones(layer, N)
        self.norm = LayerNorm(layer.size)

    def forward(self, x, mask, interctc_alpha=0, interctc_layer=6):
        inter = None
        nxt = _next_norms(self.layers)
        for n, layer in enumerate(self.layers):
            x = layer(x, mask, next_norm=nxt[n])
            if interctc_alpha > 0 and n == interctc_layer - 1:
                inter = x
        return (self.norm(x), inter) if interctc_alpha > 0 else self.norm(x)


class AcEmbedExtractor(nn.Module):
    """fanat_blocks.py:101-113"""

    def __init__(self, size, src_attn, feed_forward, dropout, N):
        super().__init__()
        layer = SrcAttLayer(size, src_attn, feed_forward, dropout)
        self.layers = clones(layer, N)

    def forward(self, x, memory, trigger_mask, next_norm=None):
        nxt = _next_norms(self.layers, next_norm)
        for n, layer in enumerate(self.layers):
            x = layer(x, memory, trigger_mask, next_norm=nxt[n])
        return x

    def forward_segments(self, pe_rows, memory, seg_ctx, dense_mask_fn=None, next_norm=None):
        """next_norm: the LayerNorm of whatever consumes the embeddings next (the SAD's first pre-norm)."""
        if torch.is_grad_enabled() and (memory.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training: dense trigger mask through the attention that has a backward
            return self.forward(pe_rows.unsqueeze(0).expand(memory.size(0), -1, -1), memory, dense_mask_fn())
        nxt = _next_norms(self.layers, next_norm)
        x = self.layers[0].forward_segments(pe_rows, memory, seg_ctx, next_norm=nxt[0])
        if len(self.layers) > 1:
            mask = dense_mask_fn()
            for n, layer in enumerate(self.layers[1:], 1):
                x = layer(x, memory, mask, next_norm=nxt[n])
        return x


class SelfAttDecoder(nn.Module):
    """fanat_blocks.py:115-136"""

    def __init__(self, size, self_attn, feed_forward, dropout, N):
        super().__init__()
        layer = SelfAttLayer(size, self_attn, feed_forward, dropout)
        self.layers = clones(layer, N)

    def forward(self, x, mask, interce_alpha=0, interce_layer=4, next_norm=None):
        inter = None
        nxt = _next_norms(self.layers, next_norm)
        for n, layer in enumerate(self.layers):
            x = layer(x, mask, next_norm=nxt[n])
            if interce_alpha > 0 and n == interce_layer - 1:
                inter = x
        return (x, inter) if interce_alpha > 0 else x


class MixAttDecoder(nn.Module):
    """fanat_blocks.py:138-159"""

    def __init__(self, size, self_attn, src_attn, feed_forward, dropout, N):
        super().__init__()
        layer = MixAttLayer(size, self_attn, src_attn, feed_forward, dropout)
        self.layers = clones(layer, N)
        self.norm = LayerNorm(layer.size)

    def project_memory(self, memory):
        """K/V of `memory` for every layer in ONE GEMM ([B,T,d] x [d, 2d*N]); the per-layer K/V
        projections are ~60% of the decoder-side FLOPs in the reference."""
        return _project_memory_all(self, memory)

    def forward(self, x, memory, src_mask, tgt_mask, interce_alpha=0, interce_layer=4):
        kvs = self.project_memory(memory)
        inter = None
        nxt = _next_norms(self.layers)           # (the final norm stays a separate pass: it also feeds the fp32 API)
        for n, layer in enumerate(self.layers):
            x = layer(x, memory, src_mask, tgt_mask, kvs[n], next_norm=nxt[n])
            if interce_alpha > 0 and n == interce_layer - 1:
                inter = x
        return (self.norm.for_head(x), inter) if interce_alpha > 0 else self.norm.for_head(x)


def _project_memory_all(dec, memory):
    from .modules import _cat_cached
    attns = [layer.src_attn for layer in dec.layers]
    w = _cat_cached(dec, "kv_all_w", [t for a in attns for t in (a.linears[1].weight, a.linears[2].weight)])
    b = _cat_cached(dec, "kv_all_b", [t for a in attns for t in (a.linears[1].bias, a.linears[2].bias)])
    kv = ops.linear(memory, w, b, **ops.attn_kw())                            # [B,T,2d*N]
    d = memory.size(-1)
    parts = ops.split_cols_multi(kv, tuple(i * d for i in range(2 * len(attns) + 1)))
    return [(parts[2 * i], parts[2 * i + 1]) for i in range(len(attns))]


# ================================================================================= conformer family
class ConSelfAttLayer(nn.Module):
    """fanat_conformer_blocks.py:9-37: 1/2 FFN -> (rel) self-attn -> conv module -> 1/2 FFN for
    pos_type 'relative'; conv before attention for 'absolute'."""

    def __init__(self, size, feed_forward1, self_attn, conv_module, feed_forward2, dropout, pos_type, share_ff=False,
                 ff_scale=0.5):
        super().__init__()
        self.self_attn = self_attn
        self.feed_forward1 = feed_forward1
        self.conv_module = conv_module
        self.feed_forward2 = feed_forward1 if share_ff else feed_forward2
        self.sublayer = clones(SublayerConnection(size, dropout), 4)
        self.size = size
        self.pos_type = pos_type
        self.ff_scale = ff_scale

    def _conv(self, x, next_norm=None):
        y = self.sublayer[1].norm.for_gemm(x)
        return _residual(self.sublayer[1], x, lambda scale=1.0, **kw: self.conv_module(y, **kw), next_norm=next_norm)

    def forward(self, x, mask, pos_embed, next_norm=None):
        sl = self.sublayer
        if self.pos_type == "absolute":
            x = _ffn_block(x, sl[0], self.feed_forward1, self.ff_scale, next_norm=sl[1].norm)
            x = self._conv(x, sl[2].norm)
            x = _self_attention_block(x, sl[2], self.self_attn, mask, None, next_norm=sl[3].norm)
        else:
            x = _ffn_block(x, sl[0], self.feed_forward1, self.ff_scale, next_norm=sl[2].norm)
            x = _self_attention_block(x, sl[2], self.self_attn, mask, pos_embed, next_norm=sl[1].norm)
            x = self._conv(x, sl[3].norm)
        return _ffn_block(x, sl[3], self.feed_forward2, self.ff_scale, next_norm=next_norm)


class ConSrcAttLayer(nn.Module):
    """fanat_conformer_blocks.py:40-61: attention WITHOUT pre-norm or residual, times sqrt(d),
    positional encoding, then a pre-norm FFN."""

    def __init__(self, size, src_attn, feed_forward, pos_enc, pos_type, dropout):
        super().__init__()
        self.src_attn = src_attn
        self.feed_forward = feed_forward
        self.sublayer = SublayerConnection(size, dropout)
        self.pos_enc = pos_enc
        self.pos_type = pos_type
        self.size = size

    def _finish(self, ctx, next_norm=None):
        if self.pos_type == "relative":
            # relative positions: pos_enc leaves x untouched, so the pre-norm of the FFN can ride on the projection
            x, pos_embed = self.pos_enc(self.src_attn.output(ctx, scale=math.sqrt(self.size), next_ln=self.sublayer.norm))
            return _ffn_block(x, self.sublayer, self.feed_forward, next_norm=next_norm), pos_embed
        x = self.pos_enc(self.src_attn.output(ctx, scale=math.sqrt(self.size)))
        return _ffn_block(x, self.sublayer, self.feed_forward, next_norm=next_norm)

    def forward(self, x, m, src_mask, next_norm=None):
        attn = self.src_attn
        q = attn.project_q(x)
        k, v = attn.project_kv(m)
        return self._finish(ops.attention(q, k, v, src_mask, attn.h), next_norm)

    def forward_segments(self, pe_rows, m, seg_ctx, kv=None, next_norm=None):
        attn = self.src_attn
        fast = ops.attn_bf16() and attn.d_k == 64 and m.size(1) <= 512
        q_rows = attn.project_q(pe_rows, out_bf16=fast)
        k, v = kv if kv is not None else attn.project_kv(m, out_bf16=fast)
        return self._finish(ops.segment_attention(q_rows, k, v, *seg_ctx[:6], attn.h), next_norm)


class ConMixAttLayer(nn.Module):
    """fanat_conformer_blocks.py:63-97"""

    def __init__(self, size, feed_forward1, self_attn, conv_module, src_attn, feed_forward2, dropout, pos_type,
                 share_ff=False, ff_scale=0.5):
        super().__init__()
        self.size = size
        self.src_attn = src_attn
        self.self_attn = self_attn
        self.feed_forward1 = feed_forward1
        self.conv_module = conv_module
        self.feed_forward2 = feed_forward1 if share_ff else feed_forward2
        self.sublayer = clones(SublayerConnection(size, dropout), 5)
        self.pos_type = pos_type
        self.ff_scale = ff_scale

    _conv = ConSelfAttLayer._conv

    def forward(self, x, memory, src_mask, self_mask, pos_embed, kv=None, next_norm=None):
        sl = self.sublayer
        if self.pos_type == "absolute":
            x = _ffn_block(x, sl[0], self.feed_forward1, self.ff_scale, next_norm=sl[1].norm)
            x = self._conv(x, sl[2].norm)
            x = _self_attention_block(x, sl[2], self.self_attn, self_mask, None, next_norm=sl[3].norm)
        else:
            x = _ffn_block(x, sl[0], self.feed_forward1, self.ff_scale, next_norm=sl[2].norm)
            x = _self_attention_block(x, sl[2], self.self_attn, self_mask, pos_embed, next_norm=sl[1].norm)
            x = self._conv(x, sl[3].norm)
        x = _src_attention_block(x, sl[3], self.src_attn, memory, src_mask, kv, next_norm=sl[4].norm)
        return _ffn_block(x, sl[4], self.feed_forward2, self.ff_scale, next_norm=next_norm)


def _split_pos(x, pos_type):
    if pos_type == "relative":
        return x[0], x[1]
    return x, None


class ConEncoder(nn.Module):
    """fanat_conformer_blocks.py:141-169"""

    def __init__(self, size, feed_forward1, self_attn, conv_module, feed_forward2, dropout, N, pos_type,
                 share_ff=False, ff_scale=0.5):
        super().__init__()
        layer = ConSelfAttLayer(size, feed_forward1, self_attn, conv_module, feed_forward2, dropout, pos_type,
                                share_ff, ff_scale)
        self.layers = clones(layer, N)
        self.pos_type = pos_type
        self.norm = LayerNorm(size)

    def forward(self, x, mask, interctc_alpha=0, interctc_layer=0):
        x, pos_embed = _split_pos(x, self.pos_type)
        inter = None
        nxt = _next_norms(self.layers)
        for n, layer in enumerate(self.layers):
            x = layer(x, mask, pos_embed, next_norm=nxt[n])
            if interctc_alpha > 0 and n == interctc_layer - 1:
                inter = x
        return (self.norm(x), inter) if interctc_alpha > 0 else self.norm(x)


class ConAcEmbedExtractor(nn.Module):
    """fanat_conformer_blocks.py:171-183"""

    def __init__(self, size, src_attn, feed_forward, pos_enc, pos_type, dropout, N):
        super().__init__()
        layer = ConSrcAttLayer(size, src_attn, feed_forward, pos_enc, pos_type, dropout)
        self.layers = clones(layer, N)
        assert N == 1, "if more than 1 layer use, code needs to be modified"

    def forward(self, x, memory, trigger_mask, next_norm=None):
        for layer in self.layers:
            x = layer(x, memory, trigger_mask, next_norm=next_norm)
        return x

    def forward_segments(self, pe_rows, memory, seg_ctx, dense_mask_fn=None, next_norm=None):
        if torch.is_grad_enabled() and (memory.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self.forward(pe_rows.unsqueeze(0).expand(memory.size(0), -1, -1), memory, dense_mask_fn())
        return self.layers[0].forward_segments(pe_rows, memory, seg_ctx, next_norm=next_norm)


class ConSelfAttDecoder(nn.Module):
    """fanat_conformer_blocks.py:185-211"""

    def __init__(self, size, feed_forward1, self_attn, conv_module, feed_forward2, dropout, N, pos_type,
                 share_ff=False, ff_scale=0.5):
        super().__init__()
        layer = ConSelfAttLayer(size, feed_forward1, self_attn, conv_module, feed_forward2, dropout, pos_type,
                                share_ff, ff_scale)
        self.layers = clones(layer, N)
        self.pos_type = pos_type

    def forward(self, x, mask, interce_alpha=0, interce_layer=0, next_norm=None):
        x, pos_embed = _split_pos(x, self.pos_type)
        inter = None
        nxt = _next_norms(self.layers, next_norm)
        for n, layer in enumerate(self.layers):
            x = layer(x, mask, pos_embed, next_norm=nxt[n])
            if interce_alpha > 0 and n == interce_layer - 1:
                inter = x
        return (x, pos_embed, inter) if interce_alpha > 0 else (x, pos_embed)


class ConMixAttDecoder(nn.Module):
    """fanat_conformer_blocks.py:213-243"""

    def __init__(self, size, feed_forward1, self_attn, conv_module, src_attn, feed_forward2, dropout, N, pos_type,
                 share_ff=False, ff_scale=0.5):
        super().__init__()
        layer = ConMixAttLayer(size, feed_forward1, self_attn, conv_module, src_attn, feed_forward2, dropout, pos_type,
                               share_ff, ff_scale)
        self.layers = clones(layer, N)
        self.pos_type = pos_type
        self.norm = LayerNorm(layer.size)

    def project_memory(self, memory):
        return _project_memory_all(self, memory)

    def forward(self, x, memory, src_mask, tgt_mask, interce_alpha=0, interce_layer=0, is_last=True):
        x, pos_embed = _split_pos(x, self.pos_type)
        kvs = self.project_memory(memory)
        inter = None
        nxt = _next_norms(self.layers)
        for n, layer in enumerate(self.layers):
            x = layer(x, memory, src_mask, tgt_mask, pos_embed, kvs[n], next_norm=nxt[n])
            if interce_alpha > 0 and n == interce_layer - 1:
                inter = x
        out = self.norm.for_head(x)
        if interce_alpha > 0:
            return (out, inter) if is_last else (out, pos_embed, inter)
        return out if is_last else (out, pos_embed)


# names the reference exports from models/blocks/__init__.py:2-13
TrfEncoder, TrfAcExtra, TrfSAD, TrfMAD = Encoder, AcEmbedExtractor, SelfAttDecoder, MixAttDecoder
ConAcExtra, ConSAD, ConMAD = ConAcEmbedExtractor, ConSelfAttDecoder, ConMixAttDecoder
