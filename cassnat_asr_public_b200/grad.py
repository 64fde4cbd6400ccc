"""Autograd bindings of the CUDA kernels: what makes `loss.backward()` of the training call
(src/tasks/cassnat_task.py:271,291) run on this package's kernels.

Each op of `ops.py` that carries a gradient gets a `torch.autograd.Function` whose backward is built
from the backward kernels of csrc/backward.cu (strided fp32 GEMM, LayerNorm / activation / softmax
backward, label smoothing) plus, in the bf16 mode, the tcgen05 GEMM on transposed bf16 operands.
`ops.linear / layer_norm / attention / cast_bf16` dispatch here when grad mode is on and an input
requires grad; PyTorch autograd only provides the tape (slicing, concatenation of the fused QKV
weights, residual adds)."""
import math

import torch

from . import _lib
from . import ops
from ._lib import call, ptr, stream

_ACT = ops.ACT


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _f32_2d(t, cols):
    t = t.reshape(-1, cols)
    if t.dtype != torch.float32:
        t = t.float()
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def gemm_strided(A, a_strides, B, b_strides, C, c_strides, nz1, nz2, M, N, K, alpha=1.0, beta=0.0):
    """C[z](m,n) = alpha * sum_k A[z](m,k) B[z](n,k) + beta C; strides are (z1, z2, row, k) in elements."""
    call("cn_gemm_f32_strided", ptr(A), *a_strides, ptr(B), *b_strides, ptr(C), *c_strides, nz1, nz2, M, N, K,
         float(alpha), float(beta), stream(), work=dict(flops=2.0 * nz1 * nz2 * M * N * K))
    return C


def act_bwd_colsum(dy, saved, act, alpha, want_dpre, want_colsum):
    """-> (dpre or None, colsum or None) of cn_act_bwd_colsum on 2-D fp32 dy."""
    M, N = dy.shape
    dpre = torch.empty(M, N, dtype=torch.float32, device=dy.device) if want_dpre else None
    col = torch.empty(N, dtype=torch.float32, device=dy.device) if want_colsum else None
    ws_bytes = _lib.load().cn_colsum_workspace_bytes(M, N) if want_colsum else 0
    ws = _ws(ws_bytes, dy.device)
    call("cn_act_bwd_colsum", ptr(dy), dy.stride(0), ptr(saved), saved.stride(0) if saved is not None else 0,
         _ACT[act], float(alpha), ptr(dpre), N, M, N, ptr(col), 0, ptr(ws), ws_bytes, stream(),
         work=dict(bytes=4.0 * M * N * (2 + (saved is not None))))
    return dpre, col


def transpose_bf16(x2, pad_to=8):
    """[R,C] f32|bf16 -> bf16 [C, Rpad] (row pitch padded to a multiple of 8 elements for TMA);
    returns the padded buffer, whose first R columns are x^T."""
    R, C = x2.shape
    Rp = (R + pad_to - 1) // pad_to * pad_to
    out = torch.empty(C, Rp, dtype=torch.bfloat16, device=x2.device)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    call("cn_transpose_bf16", ptr(x2), int(x2.dtype == torch.bfloat16), x2.stride(0), R, C, ptr(out), Rp, stream(),
         work=dict(bytes=float(R * C) * (x2.element_size() + 2)))
    return out


def _gemm_bf16_raw(A, lda, W, ldw, M, N, K, out_dtype, device):
    """out[M,N] = A[M,K] W[N,K]^T on tcgen05 (bf16 operands, row pitches lda/ldw)."""
    o32 = torch.empty(M, N, dtype=torch.float32, device=device) if out_dtype == torch.float32 else None
    o16 = torch.empty(M, N, dtype=torch.bfloat16, device=device) if out_dtype == torch.bfloat16 else None
    call("cn_gemm_bf16", ptr(A), lda, ptr(W), ldw, None, None, 0, 0, 1.0, 0, ptr(o32), ptr(o16), N, M, N, K, stream(),
         work=dict(flops=2.0 * M * N * K, bytes=2.0 * (M * K + N * K) + M * N * (4.0 if o32 is not None else 2.0)),
         tag=(f"cn_gemm_bf16[{M}x{N}x{K}]" if _lib.PROFILE_SHAPES else None))
    return o32 if o32 is not None else o16


def act_forward(x, act, out_bf16=False):
    """act(x) elementwise (f32 in; f32 or bf16 out)."""
    xc = x.contiguous()
    y = torch.empty(xc.shape, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
    call("cn_act_fwd", ptr(xc), _ACT[act], None if out_bf16 else ptr(y), ptr(y) if out_bf16 else None, xc.numel(), stream(),
         work=dict(bytes=float(xc.numel()) * (4 + y.element_size())))
    return y


USE_MN_GEMM = True      # backward GEMMs read dY / X / W in place (MN-major tcgen05 operands) instead of transposed copies


def gemm_bf16_mn(A, a_mn, W, M, N, K, out_dtype):
    """out[M,N] = op(A) @ W with W stored [K,N]; a_mn: A stored [K,M] (contraction-major) instead of [M,K]."""
    o32 = torch.empty(M, N, dtype=torch.float32, device=A.device) if out_dtype == torch.float32 else None
    o16 = torch.empty(M, N, dtype=torch.bfloat16, device=A.device) if out_dtype == torch.bfloat16 else None
    call("cn_gemm_bf16_mn", ptr(A), A.stride(0), int(a_mn), ptr(W), W.stride(0), 1, ptr(o32), ptr(o16), N, M, N, K, stream(),
         work=dict(flops=2.0 * M * N * K, bytes=2.0 * (M * K + N * K) + M * N * (4.0 if o32 is not None else 2.0)),
         tag=(f"cn_gemm_bf16_mn[{M}x{N}x{K}]" if _lib.PROFILE_SHAPES else None))
    return o32 if o32 is not None else o16


# =============================================================================== linear
def act_bwd_colsum_bf16(dy, saved, act, alpha, want_colsum, want_dpre=True):
    """-> (dpre bf16 [M,N] or None, colsum or None) of cn_act_bwd_colsum_bf16 on 2-D dy / saved in f32 or bf16 (no dtype
    copies)."""
    M, N = dy.shape
    dpre = torch.empty(M, N, dtype=torch.bfloat16, device=dy.device) if want_dpre else None
    col = torch.empty(N, dtype=torch.float32, device=dy.device) if want_colsum else None
    ws_bytes = _lib.load().cn_colsum_workspace_bytes(M, N) if want_colsum else 0
    ws = _ws(ws_bytes, dy.device)
    call("cn_act_bwd_colsum_bf16", ptr(dy), int(dy.dtype == torch.bfloat16), dy.stride(0), ptr(saved),
         int(saved is not None and saved.dtype == torch.bfloat16), saved.stride(0) if saved is not None else 0, _ACT[act],
         float(alpha), ptr(dpre), N, M, N, ptr(col), 0, ptr(ws), ws_bytes, stream(),
         work=dict(bytes=float(M * N) * (dy.element_size() + 2 + (saved.element_size() if saved is not None else 0))))
    return dpre, col


def _rows_2d(t, cols):
    """2-D view [-1, cols] with unit column stride in f32 or bf16 (other types -> f32)."""
    t = t.reshape(-1, cols)
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float()
    return t if t.stride(-1) == 1 else t.contiguous()


class Linear(torch.autograd.Function):
    """residual + alpha * act(x W^T + b)  (ops.linear); dX = dpre W, dW = dpre^T X, db = colsum(dpre),
    dpre = alpha * dY * act'."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, alpha, act, res_rows, out_bf16):
        swish = act == "swish"            # the backward needs the pre-activation: GEMM, then a separate activation
        assert not (swish and residual is not None)
        assert not (act == "relu" and (residual is not None or alpha != 1.0)), "relu mask is taken from the output"
        ctx.tc = ops.is_bf16() and x.shape[-1] % 8 == 0
        ctx.alpha, ctx.act, ctx.res_rows = float(alpha), act, int(res_rows)
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        ctx.wb = None
        with torch.no_grad():
            if ctx.tc:
                xs = ops.cast_bf16(x.detach())                # the operands the forward GEMM consumes
                ctx.wb = ops.bf16_weight(weight)
                y = ops.linear_bf16(xs, ctx.wb, bias, residual, 1.0 if swish else alpha, None if swish else act, res_rows,
                                    torch.bfloat16 if (out_bf16 and not swish) else torch.float32)
            else:
                xs = x
                y = ops.linear_forward(x, weight, bias, residual, 1.0 if swish else alpha, None if swish else act,
                                       res_rows, None, out_bf16 and not swish)
            saved = y if act == "relu" else None
            if swish:
                saved = y
                y = act_forward(saved, "swish", out_bf16)
                if alpha != 1.0:
                    y = y * alpha
        ctx.save_for_backward(xs, weight, saved)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y_act = ctx.saved_tensors
        N, K = weight.shape
        needs = ctx.needs_input_grad
        fused16 = ctx.tc and N % 8 == 0 and USE_MN_GEMM       # tensor-core backward: dY goes to the GEMMs as bf16
        gy2 = _rows_2d(gy, N) if fused16 else _f32_2d(gy, N)
        M = gy2.shape[0]
        dev = gy2.device
        g_res = None
        if ctx.has_res and needs[3]:
            g_res = gy if gy.dtype == torch.float32 else gy.float()
            if ctx.res_rows:
                g_res = g_res.reshape(-1, ctx.res_rows, N).sum(0)
        want_dpre = ctx.act is not None or ctx.alpha != 1.0
        want_db = ctx.has_bias and needs[2]
        d16 = None
        if fused16:
            saved = _rows_2d(y_act, N) if y_act is not None else None
            if want_dpre or want_db or gy2.dtype != torch.bfloat16:
                d16, db = act_bwd_colsum_bf16(gy2, saved, ctx.act, ctx.alpha, want_db)      # one pass: act', bf16, colsum
            else:
                d16, db = gy2, None
                if d16.stride(0) % 8 or d16.data_ptr() % 16:
                    d16 = d16.contiguous()
            dpre = None
        else:
            saved = _f32_2d(y_act, N) if y_act is not None else None
            if want_dpre or want_db:
                dpre, db = act_bwd_colsum(gy2, saved, ctx.act, ctx.alpha, want_dpre, want_db)
                if dpre is None:
                    dpre = gy2
            else:
                dpre, db = gy2, None
        x2 = x.reshape(-1, K)
        dx = dw = None
        if fused16:
            if needs[0]:
                dx = gemm_bf16_mn(d16, False, ctx.wb, M, K, N, ctx.x_dtype)        # dY [M,N] x W [N,K] (contraction-major)
            if needs[1]:
                xb = x2 if x2.is_contiguous() else x2.contiguous()
                dw = gemm_bf16_mn(d16, True, xb, N, K, M, torch.float32)           # dY^T x X, both contraction-major
        elif ctx.tc and N % 8 == 0:
            d16 = ops.cast_bf16(dpre)
            if needs[0]:
                wt = transpose_bf16(ctx.wb)                                        # [K, Npad]
                dx = _gemm_bf16_raw(d16, N, wt, wt.stride(0), M, K, N, ctx.x_dtype, dev)
            if needs[1]:
                dt = transpose_bf16(dpre)                                          # [N, Mpad]
                xt = transpose_bf16(x2)                                            # [K, Mpad]
                dw = _gemm_bf16_raw(dt, dt.stride(0), xt, xt.stride(0), N, K, M, torch.float32, dev)
        else:
            x2 = _f32_2d(x2, K)
            w = weight if weight.stride(-1) == 1 else weight.contiguous()
            if needs[0]:
                dx = torch.empty(M, K, dtype=torch.float32, device=dev)
                gemm_strided(dpre, (0, 0, dpre.stride(0), 1), w, (0, 0, 1, w.stride(0)), dx, (0, 0, K), 1, 1, M, K, N)
                if ctx.x_dtype != torch.float32:
                    dx = dx.to(ctx.x_dtype)
            if needs[1]:
                dw = torch.empty(N, K, dtype=torch.float32, device=dev)
                gemm_strided(dpre, (0, 0, 1, dpre.stride(0)), x2, (0, 0, 1, x2.stride(0)), dw, (0, 0, K), 1, 1, N, K, M)
        if dx is not None:
            dx = dx.reshape(ctx.x_shape)
        return dx, dw, db, g_res, None, None, None, None


# =============================================================================== layer norm
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a_2, b_2, eps, bf16):
        with torch.no_grad():
            y = ops.layer_norm_forward(x, a_2, b_2, eps, bf16, split_ok=False)     # training runs plain bf16 GEMMs
        ctx.eps = float(eps)
        ctx.save_for_backward(x, a_2)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, a = ctx.saved_tensors
        d = x.shape[-1]
        x2, g2 = _f32_2d(x, d), _f32_2d(gy, d)
        rows = x2.shape[0]
        dx = torch.empty(rows, d, dtype=torch.float32, device=x.device)
        da = torch.empty(d, dtype=torch.float32, device=x.device)
        db = torch.empty(d, dtype=torch.float32, device=x.device)
        ws_bytes = _lib.load().cn_layernorm_bwd_workspace_bytes(rows, d)
        ws = _ws(ws_bytes, x.device)
        call("cn_layernorm_bwd", ptr(x2), x2.stride(0), ptr(g2), g2.stride(0), ptr(a), ctx.eps, rows, d, ptr(dx), d,
             ptr(da), ptr(db), ptr(ws), ws_bytes, stream(), work=dict(bytes=float(rows * d) * 20))
        dx = dx.reshape(x.shape)
        if x.dtype != torch.float32:
            dx = dx.to(x.dtype)
        return dx, da, db, None, None


# =============================================================================== dropout
class Dropout(torch.autograd.Function):
    """y = x * mask / (1 - p); the mask is regenerated from (seed, site) in backward, never stored."""

    @staticmethod
    def forward(ctx, x, p, site):
        ctx.p, ctx.site = p, site
        return ops.dropout_forward(x, p, site)

    @staticmethod
    def backward(ctx, gy):
        return ops.dropout_forward(gy, ctx.p, ctx.site), None, None


class DropoutAdd(torch.autograd.Function):
    """y = residual + scale * x * mask / (1 - p) in one pass; backward: d x = scale * mask / (1 - p) * g (mask regenerated
    from (seed, site)), d residual = g."""

    @staticmethod
    def forward(ctx, x, residual, p, scale, site):
        ctx.p, ctx.scale, ctx.site, ctx.x_dtype = p, scale, site, x.dtype
        return ops.dropout_add_forward(x, residual, p, scale, site)

    @staticmethod
    def backward(ctx, gy):
        gx = ops.dropout_add_forward(gy, None, ctx.p, ctx.scale, ctx.site) if ctx.needs_input_grad[0] else None
        if gx is not None and ctx.x_dtype != torch.float32:
            gx = gx.to(ctx.x_dtype)
        return gx, (gy if ctx.needs_input_grad[1] else None), None, None, None


def _attn_probs_dropped(S, mask, B, h, Tq, Tk, zero_masked, p, site):
    """In place on the scaled scores S [B,h,Tq,Tk]: softmax with the reference's masking, then dropout."""
    m, mb, mi = ops._mask_strides(mask, B, Tq, Tk)
    call("cn_attn_softmax_bwd", ptr(S), None, ptr(m), mb, mi, B, h, Tq, Tk, int(zero_masked), stream(),
         work=dict(bytes=8.0 * B * h * Tq * Tk))
    return ops.dropout_forward(S, p, site, inplace=True)


# =============================================================================== attention
class LinearCE(torch.autograd.Function):
    """Fused output head + label-smoothing loss (src/models/cassnat.py:226,248 with
    src/utils/loss.py:21-35): loss = sum over non-pad rows of KL(smoothed one-hot || softmax(x W^T + b)) / #tokens,
    without the [rows, V] logits / log-probs / gradient tensors of the unfused path.  Forward: the tcgen05 GEMM's
    row-statistics epilogue gives log-sum-exp, target logit and argmax per row; the row sum of the logits, which the
    smoothing term needs, is x . colsum(W) + sum(b).  Backward: the logits are recomputed tile by tile and leave the
    GEMM as dz in bf16 (cn_gemm_bf16_ce_grad), feeding the dX / dW tensor-core GEMMs.  Returns (loss, argmax ids)."""

    @staticmethod
    def forward(ctx, x, weight, bias, target, padding_idx, smoothing):
        with torch.no_grad():
            K, V = x.shape[-1], weight.shape[0]
            xs = ops.cast_bf16(x.detach()).reshape(-1, K)
            if xs.stride(-1) != 1 or xs.stride(0) % 8 or xs.data_ptr() % 16:
                xs = xs.contiguous()
            wb = ops.bf16_weight(weight)
            tgt = target.reshape(-1)
            st = ops.head_rowstats(xs, weight, bias, target=tgt, safe=False)
            keep = tgt != padding_idx
            tokens = keep.sum().clamp(min=1).float()
            wsum = wb.float().sum(0)                                              # [K]: column sums of the rounded weight
            sumz = xs.float() @ wsum + (bias.float().sum() if bias is not None else 0.0)
            conf, eps = 1.0 - smoothing, smoothing / (V - 1)
            const = (conf * math.log(conf) if conf > 0 else 0.0) + ((V - 1) * eps * math.log(eps) if eps > 0 else 0.0)
            lp_t = st["target_logit"] - st["lse"]
            per_tok = const - conf * lp_t - eps * ((sumz - V * st["lse"]) - lp_t)
            loss = per_tok.masked_fill(~keep, 0).sum() / tokens
        ctx.save_for_backward(xs, weight, bias, st["lse"], tgt.to(torch.int32), (keep.float() / tokens))
        ctx.wb, ctx.conf, ctx.eps, ctx.x_shape, ctx.x_dtype = wb, float(conf), float(eps), x.shape, x.dtype
        ctx.mark_non_differentiable(st["top1"])
        return loss, st["top1"].to(torch.int64)

    @staticmethod
    def backward(ctx, g_loss, _g_ids):
        xs, weight, bias, lse, tgt, rowscale = ctx.saved_tensors
        V, K = weight.shape
        R = xs.shape[0]
        rs = (rowscale * g_loss).contiguous()
        dz = torch.empty(R, V, dtype=torch.bfloat16, device=xs.device)
        call("cn_gemm_bf16_ce_grad", ptr(xs), xs.stride(0), ptr(ctx.wb), ctx.wb.stride(0), ptr(bias), ptr(lse), ptr(tgt), ptr(rs),
             ctx.conf, ctx.eps, ptr(dz), V, R, V, K, stream(),
             work=dict(flops=2.0 * R * V * K, bytes=2.0 * (R * K + V * K + R * V)))
        needs = ctx.needs_input_grad
        dx = dw = db = None
        if needs[0]:
            dx = gemm_bf16_mn(dz, False, ctx.wb, R, K, V, ctx.x_dtype).reshape(ctx.x_shape)
        if needs[1]:
            dw = gemm_bf16_mn(dz, True, xs, V, K, R, torch.float32)
        if bias is not None and needs[2]:
            _, db = act_bwd_colsum_bf16(dz, None, None, 1.0, True, want_dpre=False)
        return dx, dw, db, None, None, None


class SplitCols(torch.autograd.Function):
    """Column blocks of a fused projection output (q | k | v, or the K/V of every decoder layer) as views.  Backward
    is ONE concatenation of the block gradients -- autograd's own slicing would materialise a zero-padded full-width
    gradient per block and add them up (3 + 2 kernels per attention instead of 1)."""

    @staticmethod
    def forward(ctx, t, *bounds):
        ctx.bounds = bounds
        ctx.meta = (t.shape, t.dtype, t.device)
        return tuple(t[..., lo:hi] for lo, hi in zip(bounds[:-1], bounds[1:]))

    @staticmethod
    def backward(ctx, *gs):
        shape, dtype, dev = ctx.meta
        b = ctx.bounds
        parts = []
        if b[0] > 0:
            parts.append(torch.zeros(*shape[:-1], b[0], dtype=dtype, device=dev))
        for g, lo, hi in zip(gs, b[:-1], b[1:]):
            parts.append(g.to(dtype) if g is not None else torch.zeros(*shape[:-1], hi - lo, dtype=dtype, device=dev))
        if b[-1] < shape[-1]:
            parts.append(torch.zeros(*shape[:-1], shape[-1] - b[-1], dtype=dtype, device=dev))
        return (torch.cat(parts, -1),) + (None,) * len(b)


FUSED_SMALL_ATTENTION = True      # attention of up to 128 query rows: backward (and the forward with dropout on the
#                                   probabilities) as one launch each; tests compare with the composed path


class Attention(torch.autograd.Function):
    """Masked multi-head attention (ops.attention, plain variant).  Backward recomputes the
    probabilities: S = QK^T/sqrt(dk), dP = dO V^T, softmax backward in place, then
    dV = P^T dO, dQ = dS K/sqrt(dk), dK = dS^T Q/sqrt(dk) as batched strided GEMMs."""

    @staticmethod
    def forward(ctx, q, k, v, mask, n_head, p=0.0, site=0):
        with torch.no_grad():
            if p > 0.0 and FUSED_SMALL_ATTENTION and q.shape[1] <= 128 and q.shape[-1] // n_head == 64 and \
                    q.dtype in (torch.float32, torch.bfloat16) and k.dtype == q.dtype and v.dtype == q.dtype:
                # few query rows: dropout on the probabilities inside one fused launch (cn_attention_fwd_small)
                B, Tq, d = q.shape
                Tk = k.shape[1]
                for t in (q, k, v):
                    assert t.stride(-1) == 1
                out = torch.empty(B, Tq, d, dtype=q.dtype, device=q.device)
                m, mb, mi = ops._mask_strides(mask, B, Tq, Tk)
                call("cn_attention_fwd_small", ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0), ptr(v),
                     v.stride(1), v.stride(0), int(q.dtype == torch.bfloat16), ptr(m), mb, mi, B, n_head, Tq, Tk, 64, float(p),
                     ptr(ops.DROPOUT_SEED_DEV), int(ops.DROPOUT_SEED), int(site), ptr(out), out.stride(1), out.stride(0),
                     stream(), work=dict(flops=4.0 * B * Tq * Tk * d))
            elif p > 0.0:
                # dropout on the probabilities (attention.py:22-23): materialised scores, P, dropout, P V
                B, Tq, d = q.shape
                Tk, dk = k.shape[1], d // n_head
                f = lambda t: t if t.dtype == torch.float32 else t.float()
                qf, kf, vf = f(q), f(k), f(v)
                S = torch.empty(B, n_head, Tq, Tk, dtype=torch.float32, device=q.device)
                st = lambda t: (t.stride(0), dk, t.stride(1), 1)
                gemm_strided(qf, st(qf), kf, st(kf), S, (n_head * Tq * Tk, Tq * Tk, Tk), B, n_head, Tq, Tk, dk, alpha=1.0 / dk ** 0.5)
                _attn_probs_dropped(S, mask, B, n_head, Tq, Tk, 0, p, site)
                out = torch.empty(B, Tq, d, dtype=torch.float32, device=q.device)
                gemm_strided(S, (n_head * Tq * Tk, Tq * Tk, Tk, 1), vf, (vf.stride(0), dk, 1, vf.stride(1)), out, (Tq * d, dk, d),
                             B, n_head, Tq, dk, Tk)
                if q.dtype != torch.float32:
                    out = out.to(q.dtype)
            else:
                out = ops.attention_forward(q, k, v, mask, n_head)
        ctx.n_head, ctx.p, ctx.site = n_head, p, site
        ctx.mask = mask
        ctx.save_for_backward(q, k, v, out)
        return out

    @staticmethod
    def backward(ctx, go):
        q, k, v, out = ctx.saved_tensors
        h = ctx.n_head
        B, Tq, d = q.shape
        Tk = k.shape[1]
        dk = d // h
        dev = q.device
        if FUSED_SMALL_ATTENTION and Tq <= 128 and dk == 64 and q.dtype in (torch.float32, torch.bfloat16) and \
                k.dtype == q.dtype and v.dtype == q.dtype and go.dtype == q.dtype and out.dtype == q.dtype:
            # few query rows (the label-token side of training): one fused launch (cn_attention_bwd_small)
            g = go if go.stride(-1) == 1 else go.contiguous()
            for t in (q, k, v, out):
                assert t.stride(-1) == 1
            dq = torch.empty(B, Tq, d, dtype=q.dtype, device=dev)
            dk_ = torch.empty(B, Tk, d, dtype=q.dtype, device=dev)
            dv = torch.empty(B, Tk, d, dtype=q.dtype, device=dev)
            m, mb, mi = ops._mask_strides(ctx.mask, B, Tq, Tk)
            call("cn_attention_bwd_small", ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0), ptr(v), v.stride(1),
                 v.stride(0), ptr(out), out.stride(1), out.stride(0), ptr(g), g.stride(1), g.stride(0),
                 int(q.dtype == torch.bfloat16), ptr(m), mb, mi, B, h, Tq, Tk, dk, float(ctx.p), ptr(ops.DROPOUT_SEED_DEV),
                 int(ops.DROPOUT_SEED), int(ctx.site), ptr(dq), dq.stride(1), dq.stride(0), ptr(dk_), dk_.stride(1),
                 dk_.stride(0), ptr(dv), dv.stride(1), dv.stride(0), stream(),
                 work=dict(flops=12.0 * B * Tq * Tk * d))
            return dq, dk_, dv, None, None, None, None
        f = lambda t: t if t.dtype == torch.float32 else t.float()
        qf, kf, vf, gf = f(q), f(k), f(v), f(go)
        if gf.stride(-1) != 1:
            gf = gf.contiguous()
        scale = 1.0 / (dk ** 0.5)
        S = torch.empty(B, h, Tq, Tk, dtype=torch.float32, device=dev)
        dP = torch.empty(B, h, Tq, Tk, dtype=torch.float32, device=dev)
        sz = (h * Tq * Tk, Tq * Tk, Tk)
        st = lambda t: (t.stride(0), dk, t.stride(1), 1)                 # (batch, head, row, feature)
        gemm_strided(qf, st(qf), kf, st(kf), S, sz, B, h, Tq, Tk, dk, alpha=scale)
        gemm_strided(gf, st(gf), vf, st(vf), dP, sz, B, h, Tq, Tk, dk)
        if ctx.p > 0.0:
            ops.dropout_forward(dP, ctx.p, ctx.site, inplace=True)          # d(dropped P) -> dP through the same mask
        m, mb, mi = ops._mask_strides(ctx.mask, B, Tq, Tk)
        call("cn_attn_softmax_bwd", ptr(S), ptr(dP), ptr(m), mb, mi, B, h, Tq, Tk, 0, stream(),
             work=dict(bytes=16.0 * B * h * Tq * Tk))
        if ctx.p > 0.0:
            ops.dropout_forward(S, ctx.p, ctx.site, inplace=True)           # dV needs the dropped probabilities
        dq = torch.empty(B, Tq, d, dtype=torch.float32, device=dev)
        dk_ = torch.empty(B, Tk, d, dtype=torch.float32, device=dev)
        dv = torch.empty(B, Tk, d, dtype=torch.float32, device=dev)
        tr = lambda t: (t.stride(0), dk, 1, t.stride(1))                 # operand read as [feature, row]
        P_t = (h * Tq * Tk, Tq * Tk, 1, Tk)                              # P^T / dS^T: rows = keys, k = queries
        P_n = (h * Tq * Tk, Tq * Tk, Tk, 1)
        gemm_strided(S, P_t, gf, tr(gf), dv, (Tk * d, dk, d), B, h, Tk, dk, Tq)
        gemm_strided(dP, P_n, kf, tr(kf), dq, (Tq * d, dk, d), B, h, Tq, dk, Tk, alpha=scale)
        gemm_strided(dP, P_t, qf, tr(qf), dk_, (Tk * d, dk, d), B, h, Tk, dk, Tq, alpha=scale)
        c = lambda g, ref: g if ref.dtype == torch.float32 else g.to(ref.dtype)
        return c(dq, q), c(dk_, k), c(dv, v), None, None, None, None


class RelAttention(torch.autograd.Function):
    """Relative-position self-attention (attention.py:104-143) on raw projected q (without pos_bias_u):
    score(i,j) = ((q_i+u).k_j + (q_i+v).P[clamp(j-i,-m,m)+m]) / sqrt(dk), masked probabilities re-zeroed.
    Backward = plain attention backward on q+u, plus the tabulated-bias path: bucket dS by relative offset,
    then d(q+v) = G P and dP = G^T (q+v) as strided GEMMs."""

    @staticmethod
    def forward(ctx, q, k, v, rel_pos, bias_u, bias_v, mask, n_head, m, p=0.0, site=0):
        with torch.no_grad():
            key_mask = mask is None or mask.dim() == 2 or mask.shape[1] == 1
            fast = (q.dtype == torch.bfloat16 and key_mask and q.shape[-1] // n_head == 64 and k.shape[1] <= 512
                    and n_head <= 8 and m <= 64)
            if p > 0.0:
                # dropout on the probabilities (attention.py:140-141): materialised scores incl. the relative term
                B, Tq, d = q.shape
                Tk, dk, R = k.shape[1], d // n_head, 2 * int(m) + 1
                f = lambda t: t if t.dtype == torch.float32 else t.float()
                qf, kf, vf = f(q), f(k), f(v)
                Pm = f(rel_pos).contiguous()
                qu, qv = qf + bias_u.reshape(1, 1, -1), qf + bias_v.reshape(1, 1, -1)
                sc = 1.0 / dk ** 0.5
                S = torch.empty(B, n_head, Tq, Tk, dtype=torch.float32, device=q.device)
                Tb = torch.empty(B, n_head, Tq, R, dtype=torch.float32, device=q.device)
                st = lambda t: (t.stride(0), dk, t.stride(1), 1)
                gemm_strided(qu, st(qu), kf, st(kf), S, (n_head * Tq * Tk, Tq * Tk, Tk), B, n_head, Tq, Tk, dk, alpha=sc)
                gemm_strided(qv, st(qv), Pm, (0, dk, d, 1), Tb, (n_head * Tq * R, Tq * R, R), B, n_head, Tq, R, dk)
                call("cn_rel_bias_add", ptr(S), ptr(Tb), B * n_head, Tq, Tk, int(m), sc, stream(), work=dict(bytes=8.0 * B * n_head * Tq * Tk))
                _attn_probs_dropped(S, mask, B, n_head, Tq, Tk, 1, p, site)
                out = torch.empty(B, Tq, d, dtype=torch.float32, device=q.device)
                gemm_strided(S, (n_head * Tq * Tk, Tq * Tk, Tk, 1), vf, (vf.stride(0), dk, 1, vf.stride(1)), out, (Tq * d, dk, d),
                             B, n_head, Tq, dk, Tk)
                if q.dtype != torch.float32:
                    out = out.to(q.dtype)
            elif fast:                         # tcgen05 kernel takes q + u as its Q operand
                qu = (q.float() + bias_u.reshape(1, 1, -1)).to(torch.bfloat16)
                out = ops.attention_forward(qu, k, v, mask, n_head, (rel_pos, bias_u, bias_v, m))
            else:
                out = ops.attention_forward(q.float(), k.float(), v.float(), mask, n_head, (rel_pos, bias_u, bias_v, m))
                out = out.to(q.dtype)
        ctx.n_head, ctx.m, ctx.mask, ctx.p, ctx.site = n_head, int(m), mask, p, site
        ctx.save_for_backward(q, k, v, rel_pos, bias_u, bias_v)
        return out

    @staticmethod
    def backward(ctx, go):
        q, k, v, rel_pos, bias_u, bias_v = ctx.saved_tensors
        h, m = ctx.n_head, ctx.m
        B, Tq, d = q.shape
        Tk = k.shape[1]
        dk = d // h
        R = 2 * m + 1
        dev = q.device
        f = lambda t: t if t.dtype == torch.float32 else t.float()
        qf, kf, vf, gf = f(q), f(k), f(v), f(go)
        if gf.stride(-1) != 1:
            gf = gf.contiguous()
        P = f(rel_pos).contiguous()                                           # [R, d]
        qu = qf + bias_u.reshape(1, 1, -1)
        qv = qf + bias_v.reshape(1, 1, -1)
        scale = 1.0 / (dk ** 0.5)
        Z = B * h
        S = torch.empty(B, h, Tq, Tk, dtype=torch.float32, device=dev)
        dP = torch.empty(B, h, Tq, Tk, dtype=torch.float32, device=dev)
        Tb = torch.empty(B, h, Tq, R, dtype=torch.float32, device=dev)
        sz = (h * Tq * Tk, Tq * Tk, Tk)
        st = lambda t: (t.stride(0), dk, t.stride(1), 1)
        tr = lambda t: (t.stride(0), dk, 1, t.stride(1))
        gemm_strided(qu, st(qu), kf, st(kf), S, sz, B, h, Tq, Tk, dk, alpha=scale)
        gemm_strided(qv, st(qv), P, (0, dk, d, 1), Tb, (h * Tq * R, Tq * R, R), B, h, Tq, R, dk)
        call("cn_rel_bias_add", ptr(S), ptr(Tb), Z, Tq, Tk, m, scale, stream(), work=dict(bytes=8.0 * Z * Tq * Tk))
        gemm_strided(gf, st(gf), vf, st(vf), dP, sz, B, h, Tq, Tk, dk)
        if ctx.p > 0.0:
            ops.dropout_forward(dP, ctx.p, ctx.site, inplace=True)
        mk, mb, mi = ops._mask_strides(ctx.mask, B, Tq, Tk)
        call("cn_attn_softmax_bwd", ptr(S), ptr(dP), ptr(mk), mb, mi, B, h, Tq, Tk, 1, stream(),
             work=dict(bytes=16.0 * Z * Tq * Tk))
        if ctx.p > 0.0:
            ops.dropout_forward(S, ctx.p, ctx.site, inplace=True)
        dqu = torch.empty(B, Tq, d, dtype=torch.float32, device=dev)
        dqv = torch.empty(B, Tq, d, dtype=torch.float32, device=dev)
        dk_ = torch.empty(B, Tk, d, dtype=torch.float32, device=dev)
        dv = torch.empty(B, Tk, d, dtype=torch.float32, device=dev)
        P_t = (h * Tq * Tk, Tq * Tk, 1, Tk)
        P_n = (h * Tq * Tk, Tq * Tk, Tk, 1)
        gemm_strided(S, P_t, gf, tr(gf), dv, (Tk * d, dk, d), B, h, Tk, dk, Tq)
        gemm_strided(dP, P_n, kf, tr(kf), dqu, (Tq * d, dk, d), B, h, Tq, dk, Tk, alpha=scale)
        gemm_strided(dP, P_t, qu, tr(qu), dk_, (Tk * d, dk, d), B, h, Tk, dk, Tq, alpha=scale)
        G = Tb                                                                 # reuse: [B,h,Tq,R]
        call("cn_rel_bucket", ptr(dP), ptr(G), Z, Tq, Tk, m, stream(), work=dict(bytes=4.0 * Z * Tq * (Tk + R)))
        gemm_strided(G, (h * Tq * R, Tq * R, R, 1), P, (0, dk, 1, d), dqv, (Tq * d, dk, d), B, h, Tq, dk, R, alpha=scale)
        dPb = torch.empty(B, R, d, dtype=torch.float32, device=dev)
        gemm_strided(G, (h * Tq * R, Tq * R, 1, R), qv, tr(qv), dPb, (R * d, dk, d), B, h, R, dk, Tq, alpha=scale)
        _, du = act_bwd_colsum(dqu.view(-1, d), None, None, 1.0, False, True)
        _, dvb = act_bwd_colsum(dqv.view(-1, d), None, None, 1.0, False, True)
        dq = dqu + dqv
        c = lambda g, ref: g if ref.dtype == torch.float32 else g.to(ref.dtype)
        return (c(dq, q), c(dk_, k), c(dv, v), c(dPb.sum(0), rel_pos), du.view_as(bias_u), dvb.view_as(bias_v),
                None, None, None, None, None)


# =============================================================================== conformer conv module core
class ConvCore(torch.autograd.Function):
    """GLU -> depthwise conv -> GroupNorm(1,C) -> activation (conformer_related.py:33-38)."""

    @staticmethod
    def forward(ctx, y, dw_weight, dw_bias, gamma, beta, eps, act, out_bf16):
        with torch.no_grad():
            out = ops.glu_dwconv_groupnorm_act_forward(y, dw_weight, dw_bias, gamma, beta, eps, act, out_bf16)
        ctx.eps, ctx.act = float(eps), act
        ctx.save_for_backward(y, dw_weight, dw_bias, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, go):
        y, dw_weight, dw_bias, gamma, beta = ctx.saved_tensors
        B, L, C2 = y.shape
        C = C2 // 2
        ksz = dw_weight.shape[-1]
        dev = y.device
        gf = go if go.dtype == torch.float32 else go.float()
        gf = gf.contiguous()
        yc = y if y.stride(-1) == 1 else y.contiguous()
        dy = torch.empty(B, L, C2, dtype=torch.float32, device=dev)
        dw = torch.empty(C, ksz, dtype=torch.float32, device=dev)
        db, dg, dbt = (torch.empty(C, dtype=torch.float32, device=dev) for _ in range(3))
        ws_bytes = _lib.load().cn_conv_module_bwd_workspace_bytes(B, L, C, ksz)
        ws = _ws(ws_bytes, dev)
        call("cn_glu_dwconv_groupnorm_act_bwd", ptr(yc), int(yc.dtype == torch.bfloat16), yc.stride(1), yc.stride(0),
             ptr(dw_weight.reshape(C, ksz).contiguous()), ptr(dw_bias), ptr(gamma), ptr(beta), ctx.eps, _ACT[ctx.act],
             B, L, C, ksz, ptr(gf), ptr(dy), ptr(dw), ptr(db), ptr(dg), ptr(dbt), ptr(ws), ws_bytes, stream(),
             work=dict(bytes=float(B * L * C) * 40))
        if y.dtype != torch.float32:
            dy = dy.to(y.dtype)
        return dy, dw.view_as(dw_weight), db if dw_bias is not None else None, dg, dbt, None, None, None


# =============================================================================== bf16 cast
class CastBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        with torch.no_grad():
            return ops.cast_bf16(x.detach())

    @staticmethod
    def backward(ctx, g):
        return g.float()


# =============================================================================== heads + losses
class CtcFromLogits(torch.autograd.Function):
    """logits [B,T,V] -> (log-probs [B,T,V] (no gradient of its own: the caller only decodes it for
    WER, cassnat_task.py:276-279), CTC loss).  The gradient at the logits comes straight out of the
    fused alpha-beta kernel (K10): softmax - occupancy, scaled by the incoming loss gradient."""

    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, blank):
        with torch.no_grad():
            logp = torch.empty_like(logits, dtype=torch.float32)
            ops.row_stats(logits, lse=True, logp_out=logp)
            loss, _, grad = ops.ctc_loss(logp, targets, input_lengths, target_lengths, blank, want_grad=True, grad_out=1.0)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(logp)
        return logp, loss

    @staticmethod
    def backward(ctx, _g_logp, g_loss):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None, None


class LabelSmoothingFromLogits(torch.autograd.Function):
    """logits [R,V] -> (log-probs, label-smoothing KL loss) with dlogits = (p - w) / tokens."""

    @staticmethod
    def forward(ctx, logits, target, pad, smoothing):
        with torch.no_grad():
            R, V = logits.shape
            logp = torch.empty_like(logits, dtype=torch.float32)
            ops.row_stats(logits, lse=True, logp_out=logp)
            loss = torch.empty(1, dtype=torch.float32, device=logits.device)
            grad = torch.empty(R, V, dtype=torch.float32, device=logits.device)
            tg = target.to(torch.int64).contiguous()
            ws = _ws((R + 4) * 4, logits.device)
            call("cn_label_smoothing", ptr(logp), logp.stride(0), R, V, ptr(tg), int(pad), float(smoothing), 1.0,
                 ptr(loss), ptr(grad), V, ptr(ws), ws.numel(), stream(), work=dict(bytes=12.0 * R * V))
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(logp)
        return logp, loss[0]

    @staticmethod
    def backward(ctx, _g_logp, g_loss):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)
