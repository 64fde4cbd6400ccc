"""Tensor-level operators of the CASS-NAT alignment-conditioned path, each a thin wrapper that
allocates outputs with torch and calls one entry point of the C ABI (include/cassnat_b200.h).
Signatures and return conventions follow the reference methods they stand in for
(src/models/cassnat.py:259-418); every operator raises if its input is not on a CUDA device.
"""

import torch

from . import _lib
from ._lib import call, ptr, require_cuda, stream

ACT = {None: 0, "none": 0, "relu": 1, "swish": 2}


def _u8(mask, B, T):
    """[B,1,T] / [B,T] mask of any integer/bool dtype -> contiguous uint8 [B,T]."""
    m = mask.reshape(B, T)
    if m.dtype != torch.uint8:
        m = (m != 0).to(torch.uint8)
    return m.contiguous()


def _i64(x):
    return x.to(torch.int64).contiguous()


# ----------------------------------------------------------------------------- K1/K2/K9
def row_stats(x, top2=False, peak=False, lse=False, logp_out=None):
    """argmax (+runner-up, max, log-sum-exp, log-softmax) of every row of a [...,V] f32 tensor.
    Returns dict with int32 'top1' (+ 'top2', 'peak', 'lse'); logp_out (same shape) is filled
    with x - lse when given."""
    require_cuda(x)
    V = x.shape[-1]
    x2 = x.reshape(-1, V)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    rows = x2.shape[0]
    dev = x.device
    out = {"top1": torch.empty(rows, dtype=torch.int32, device=dev)}
    if top2:
        out["top2"] = torch.empty(rows, dtype=torch.int32, device=dev)
    if peak:
        out["peak"] = torch.empty(rows, dtype=torch.float32, device=dev)
    if lse or logp_out is not None:
        out["lse"] = torch.empty(rows, dtype=torch.float32, device=dev)
    lp = logp_out.reshape(-1, V) if logp_out is not None else None
    call("cn_row_stats", ptr(x2), rows, V, x2.stride(0), ptr(out["top1"]), ptr(out.get("top2")),
         ptr(out.get("peak")), ptr(out.get("lse")), ptr(lp), lp.stride(0) if lp is not None else 0, stream(),
         work=dict(bytes=4.0 * rows * V * (3 if lp is not None else 1)))
    _lib.launch_count += 1 if lp is not None else 0
    lead = x.shape[:-1]
    return {k: v.reshape(lead) for k, v in out.items()}


def best_path_align(ctc_out, src_mask, src_size, blank, sample_num=0, threshold=0.9, include_best=True):
    """CassNAT.best_path_align (cassnat.py:367-389) -> (aligned_seq_shift i64 [B,T], ylen i64 [B], ymax int).
    With sample_num > 1 the per-frame coin flips are drawn here with torch.randint on the CPU
    default generator, the call the reference makes at cassnat.py:373."""
    require_cuda(ctc_out)
    B, T, V = ctc_out.shape
    lp = ctc_out if ctc_out.stride(-1) == 1 and ctc_out.stride(0) == T * ctc_out.stride(1) else ctc_out.contiguous()
    dev = ctc_out.device
    select = None
    if sample_num > 1:
        select = torch.randint(0, 2, (B, T, 1)).to(torch.uint8).reshape(B, T).to(dev, non_blocking=True)
    mask = _u8(src_mask, B, T)
    aligned = torch.empty(B, T, dtype=torch.int64, device=dev)
    ylen = torch.empty(B, dtype=torch.int64, device=dev)
    ymax = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = _lib.load().cn_bestpath_workspace_bytes(B, T)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("cn_bestpath_from_logp", ptr(lp), B, T, V, lp.stride(0), lp.stride(1), ptr(mask), ptr(select),
         int(blank), int(sample_num), float(threshold), int(bool(include_best)),
         ptr(aligned), ptr(ylen), ptr(ymax), ptr(ws), ws_bytes, stream())
    return aligned, ylen, int(ymax.item())      # .item(): the reference's host sync at cassnat.py:387


def viterbi_align(ctc_out, src_mask, src_size, ys, ylens, blank, sample_topk=0):
    """CassNAT.viterbi_align (cassnat.py:272-353) -> (aligned_seq_shift i64 [B,T], ylens, ymax=ys.size(1))."""
    require_cuda(ctc_out)
    B, T, V = ctc_out.shape
    U = ys.size(1)
    dev = ctc_out.device
    lp = ctc_out if ctc_out.stride(-1) == 1 else ctc_out.contiguous()
    t_sample = None
    n_sample = 0
    if sample_topk > 1:
        # same draw as cassnat.py:312 (CPU default generator)
        t_sample = torch.randint(1, T, (1, sample_topk)).to(torch.int32).reshape(-1).to(dev)
        n_sample = int(sample_topk)
    mask = _u8(src_mask, B, T)
    ys64, yl64, sz64 = _i64(ys), _i64(ylens), _i64(src_size)
    aligned = torch.empty(B, T, dtype=torch.int64, device=dev)
    ws_bytes = _lib.load().cn_viterbi_workspace_bytes(B, T, U)
    if ws_bytes < 0:
        raise RuntimeError(f"viterbi_align: unsupported label length U={U} (1..511)")
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    call("cn_viterbi_align", ptr(lp), B, T, V, lp.stride(0), lp.stride(1), ptr(mask), ptr(sz64), ptr(ys64), U,
         ys64.stride(0), ptr(yl64), int(blank), ptr(t_sample), n_sample, ptr(aligned), ptr(ws), ws_bytes, stream())
    return aligned, ylens, U


def gather_emissions(ctc_out, ys, blank):
    """The U+1 posteriors per frame the forced alignment (and the CTC loss) look at: emis[b,t,0] = ctc_out[b,t,blank],
    emis[b,t,u] = ctc_out[b,t,ys[b,u-1]] -> f32 [B,T,U+1] (cn_gather_emissions)."""
    require_cuda(ctc_out)
    B, T, V = ctc_out.shape
    U = ys.size(1)
    lp = ctc_out if ctc_out.stride(-1) == 1 else ctc_out.contiguous()
    ys64 = _i64(ys)
    emis = torch.empty(B, T, U + 1, dtype=torch.float32, device=ctc_out.device)
    call("cn_gather_emissions", ptr(lp), B, T, V, lp.stride(0), lp.stride(1), ptr(ys64), U, ys64.stride(0), int(blank), ptr(emis),
         stream(), work=dict(bytes=4.0 * B * T * (U + 1) * 2))
    return emis


def viterbi_align_compact(emis, vocab_size, src_mask, src_size, ys, ylens, blank, sample_topk=0):
    """CassNAT.viterbi_align on compact emissions [B,T,U+1] (see gather_emissions): the same dynamic program, the same
    alignments bit for bit, without touching the [B,T,V] posteriors.  -> (aligned_seq_shift i64 [B,T], ylens, U)."""
    require_cuda(emis)
    B, T, U1 = emis.shape
    U = ys.size(1)
    assert U1 == U + 1 and emis.stride(-1) == 1
    dev = emis.device
    t_sample, n_sample = None, 0
    if sample_topk > 1:
        t_sample = torch.randint(1, T, (1, sample_topk)).to(torch.int32).reshape(-1).to(dev)     # cassnat.py:312
        n_sample = int(sample_topk)
    mask = _u8(src_mask, B, T)
    ys64, yl64, sz64 = _i64(ys), _i64(ylens), _i64(src_size)
    aligned = torch.empty(B, T, dtype=torch.int64, device=dev)
    ws_bytes = _lib.load().cn_viterbi_workspace_bytes(B, T, U)
    if ws_bytes < 0:
        raise RuntimeError(f"viterbi_align: unsupported label length U={U} (1..511)")
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    call("cn_viterbi_align_compact", ptr(emis), B, T, int(vocab_size), emis.stride(0), emis.stride(1), ptr(mask), ptr(sz64), ptr(ys64),
         U, ys64.stride(0), ptr(yl64), int(blank), ptr(t_sample), n_sample, ptr(aligned), ptr(ws), ws_bytes, stream())
    return aligned, ylens, U


def align_to_mask(aligned_seq_shift, ylens, ymax, src_mask, src_size, blank, left=0, right=0, and_src_mask=False,
                  want_mask=True, want_seg=False):
    """CassNAT.align_to_mask (cassnat.py:355-365) -> (bool [B,ymax+1,T], ylens+1, ymax+1).
    left/right/and_src_mask fold expand_trigger_mask and the `& src_mask` of cassnat.py:467-468 into
    the same kernel; want_seg additionally returns the int32 segment ids the fused extractor uses."""
    require_cuda(aligned_seq_shift)
    B, T = aligned_seq_shift.shape
    dev = aligned_seq_shift.device
    al, yl, sz = _i64(aligned_seq_shift), _i64(ylens), _i64(src_size)
    mask_in = _u8(src_mask, B, T)
    out = torch.empty(B, ymax + 1, T, dtype=torch.uint8, device=dev) if want_mask else None
    ylen_out = torch.empty(B, dtype=torch.int64, device=dev)
    seg = torch.empty(B, T, dtype=torch.int32, device=dev) if want_seg else None
    call("cn_align_to_mask", ptr(al), ptr(yl), ptr(mask_in), ptr(sz), B, T, int(ymax), int(blank),
         int(left > 0), int(right > 0), int(bool(and_src_mask)), ptr(out), ptr(ylen_out), ptr(seg), stream())
    res = (out.view(torch.bool) if out is not None else None, ylen_out, ymax + 1)
    return res + (seg,) if want_seg else res


def expand_trigger_mask(trigger_mask, left_trigger, right_trigger):
    """CassNAT.expand_trigger_mask (cassnat.py:259-270)."""
    require_cuda(trigger_mask)
    if left_trigger <= 0 and right_trigger <= 0:
        return trigger_mask
    B, U, T = trigger_mask.shape
    m = trigger_mask.contiguous()
    m8 = m.view(torch.uint8) if m.dtype == torch.bool else (m != 0).to(torch.uint8)
    out = torch.empty_like(m8)
    call("cn_expand_trigger_mask", ptr(m8), B, U, T, int(left_trigger), int(right_trigger), ptr(out), stream())
    return out.view(torch.bool)


def tgt_length_mask(ylen, ymax, as_u8=False):
    """The scatter/cumprod/scatter of cassnat.py:476-478 -> int64 [B,1,ymax] (1 where u < ylen[b])."""
    require_cuda(ylen)
    B = ylen.numel()
    yl = _i64(ylen)
    o64 = None if as_u8 else torch.empty(B, 1, ymax, dtype=torch.int64, device=ylen.device)
    o8 = torch.empty(B, ymax, dtype=torch.uint8, device=ylen.device) if as_u8 else None
    call("cn_tgt_length_mask", ptr(yl), B, int(ymax), ptr(o64), ptr(o8), stream())
    return o8 if as_u8 else o64


# ----------------------------------------------------------------------------- fp32 blocks
_PRECISION = "fp32"          # "fp32": FFMA GEMMs (parity mode) | "bf16" / "bf16x3": tcgen05 GEMMs, fp32 accumulate
SAFE_HEADS = True            # bf16 mode: the two decision heads (K1 CTC argmax, K9 token pick) run as bf16x3
_BF16_WEIGHTS = {}           # id(tensor) -> (version stamp, bf16 copy)
WEIGHT_EPOCH = 0             # bumped whenever a derived weight copy is rebuilt (captured CUDA graphs key on it)


def set_precision(mode):
    """Selects the arithmetic of every linear layer on the path.
    'fp32'   parity mode: FFMA GEMMs, fp32 attention (bit-exact hypotheses vs the reference);
    'bf16'   fast mode: bf16 operands on the 5th-gen tensor cores, fp32 accumulation in TMEM, under the stated bf16
             tolerance -- except the two decision heads, which run as bf16x3 (SAFE_HEADS) so that a rounding flip of
             an argmax cannot change the alignment or a token given the same head input;
    'bf16x3' precise tensor-core mode: every GEMM operand is carried as a (hi, lo) bf16 pair and multiplied in three
             passes (error ~2^-17); attention runs on the fp32 kernel (X3_FP32_ATTENTION) or, optionally, in bf16."""
    global _PRECISION
    if mode not in ("fp32", "bf16", "bf16x3"):
        raise ValueError(mode)
    _PRECISION = mode


def get_precision():
    return _PRECISION


def is_bf16():
    """True in both tensor-core modes (bf16 activations between kernels)."""
    return _PRECISION != "fp32"


def is_x3():
    return _PRECISION == "bf16x3"


X3_FP32_ATTENTION = True     # bf16x3 mode: Q/K/V stay fp32 and the attention runs on the fp32 kernel, so that no bf16
#                              rounding is left anywhere on the path (the GEMMs are the 3-pass tensor-core ones)


X3_SPLIT_ATTENTION = True    # ... unless the split-bf16 tcgen05 attention applies (key masks, d_k = 64, no relative term):
#                              Q/K/V then stay (hi, lo) bf16 pairs and scores / context take three tensor-core passes each


def attn_bf16():
    """True when Q/K/V are produced in plain bf16 for the tcgen05 attention kernel."""
    return _PRECISION == "bf16" or (_PRECISION == "bf16x3" and not X3_FP32_ATTENTION and not attn_split())


def attn_split():
    """True when Q/K/V are produced as split (hi, lo) bf16 operands for cn_attention_bf16x3."""
    return _PRECISION == "bf16x3" and X3_SPLIT_ATTENTION and not torch.is_grad_enabled()     # (forward-only kernel)


def attn_kw():
    """Keyword arguments of the Q/K/V projections for the attention kernel of the current mode."""
    return dict(out_bf16=attn_bf16() or attn_split(), out_split=attn_split())


def split_cols(t, lo, hi):
    """Columns [lo, hi) of a projection output.  For the hi half of a split operand the slice carries its lo twin
    as `_cn_lo` (same shape and pitches), which is how Q / K / V reach the split attention kernel."""
    full = split_of(t)
    out = t[..., lo:hi]
    if full is not None:
        n = t.shape[-1]
        out._cn_lo = full[..., n + lo:n + hi]
    return out


def split_cols_multi(t, bounds):
    """Consecutive column blocks t[..., bounds[i]:bounds[i+1]] of a fused projection output (see split_cols).  In
    training the blocks come from one autograd node whose backward is a single concatenation (grad.SplitCols)."""
    if torch.is_grad_enabled() and t.requires_grad and split_of(t) is None:
        from . import grad
        return grad.SplitCols.apply(t, *bounds)
    return tuple(split_cols(t, lo, hi) for lo, hi in zip(bounds[:-1], bounds[1:]))


def unsplit(t):
    """fp32 value of a (hi, lo) pair (or of a plain tensor)."""
    lo = getattr(t, "_cn_lo", None)
    if lo is None and split_of(t) is not None:
        lo = split_of(t)[..., t.shape[-1]:]
    return t.float() + lo.float() if lo is not None else t.float()


def split_of(x):
    """The [.., 2K] split twin ([hi | lo]) of a bf16 tensor produced as the hi half of a bf16x3 operand, or None."""
    sp = getattr(x, "_cn_split", None)
    if sp is not None and sp.data_ptr() == x.data_ptr() and sp.shape[-1] == 2 * x.shape[-1] and sp.shape[:-1] == x.shape[:-1]:
        return sp
    return None


def _hi_view(full):
    """[.., 2K] split operand -> its hi half [.., K] (a strided view) tagged with the full tensor."""
    K = full.shape[-1] // 2
    hi = full[..., :K]
    hi._cn_split = full
    return hi


def bf16_weight(w):
    """bf16 copy of a weight, converted once and refreshed when the parameter changes."""
    if w.grad_fn is not None and w._base is None:          # rebuilt every training step (torch.cat of parameters)
        src = w.detach()
        return cast_bf16(src if src.is_contiguous() else src.contiguous())
    base = w._base if w._base is not None else w          # views of a parameter share its cache entry
    key = (id(base), w.storage_offset(), tuple(w.shape), tuple(w.stride()))
    stamp = (base.data_ptr(), base._version)
    hit = _BF16_WEIGHTS.get(key)
    if hit is None or hit[0] != stamp:
        global WEIGHT_EPOCH
        WEIGHT_EPOCH += 1
        with torch.no_grad():
            src = w.detach()
            hit = (stamp, cast_bf16(src if src.is_contiguous() else src.contiguous()), base)
        _BF16_WEIGHTS[key] = hit
    return hit[1]


def bf16_weight_split(w):
    """Split ([hi | lo], [N, 2K]) copy of a weight for the bf16x3 GEMM, cached like bf16_weight."""
    base = w._base if w._base is not None else w
    key = ("x3", id(base), w.storage_offset(), tuple(w.shape), tuple(w.stride()))
    stamp = (base.data_ptr(), base._version)
    hit = _BF16_WEIGHTS.get(key)
    if hit is None or hit[0] != stamp:
        global WEIGHT_EPOCH
        WEIGHT_EPOCH += 1
        with torch.no_grad():
            hit = (stamp, cast_bf16_split(w.detach().float(), tag=False), base)
        _BF16_WEIGHTS[key] = hit
    return hit[1]


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def linear(x, weight, bias=None, residual=None, alpha=1.0, act=None, res_rows=0, out=None, out_bf16=False, out_split=False,
           next_ln=None):
    """residual + alpha * act(x @ weight.T + bias) over the last dim of x.
    fp32 mode: FFMA kernel on f32 operands.  bf16 mode: x (f32 or bf16) and the weight go through the
    tcgen05 GEMM as bf16 with fp32 accumulation; out_bf16 returns a bf16 tensor (GEMM->GEMM chains).
    bf16x3 mode: three-pass split GEMM; out_split (with out_bf16) returns the hi half of a split operand for the
    next GEMM.  With grad mode on and an input requiring grad the call is recorded for backward (grad.Linear).
    x may be a FoldedLN (fast mode): the LayerNorm it stands for is applied in this GEMM's epilogue."""
    if isinstance(x, FoldedLN):
        return linear_forward(x, weight, bias, residual, alpha, act, res_rows, out, out_bf16, out_split, next_ln)
    if _needs_grad(x, weight, bias, residual):
        from . import grad
        assert out is None
        return grad.Linear.apply(x, weight, bias, residual, alpha, act, res_rows, out_bf16)
    return linear_forward(x, weight, bias, residual, alpha, act, res_rows, out, out_bf16, out_split, next_ln)


FOLD_LN = False      # opt-in: LayerNorm folded into its producer / consumer GEMMs (no LayerNorm kernel, see FoldedLN).
#                      Measured on B200 (cfg3): the 22 LayerNorm launches (0.50 ms) disappear, but the extra bf16 copy +
#                      partial sums in the producers' epilogues (+0.30 ms) and the rank-1 correction in the consumers'
#                      (+0.25 ms) cost as much -- the GEMM epilogues are the contended resource of these shapes
FUSE_LN_MAX_ROWS = 0      # (automatic use below this many rows: measured slower at cfg1 too, 0.967 vs 0.939 ms/step -> off)
FUSE_LN = False      # opt-in: the LayerNorm that consumes a residual GEMM's output runs in that GEMM's epilogue
#                      (cn_gemm_bf16_ln).  Measured on B200 at [32000x512x512]: 69 us fused vs 35 + 23 us for the GEMM and
#                      the stand-alone LayerNorm -- the epilogue warps are latency-bound, the second pass over the
#                      L2-hot rows costs more there than a separate bandwidth-bound kernel; kept for small batches


class FoldedLN:
    """What LayerNorm.for_gemm returns in fast mode when the GEMM that produced x already emitted the raw bf16 copy of
    its rows and their partial sums (linear_bf16_rowpart): not a tensor but a promise -- the next ops.linear applies
    the LayerNorm algebraically in its own epilogue (cn_gemm_bf16_folded), with the gain folded into its weight."""
    __slots__ = ("x16", "part", "n_parts", "ln")
    requires_grad = False
    dtype = torch.bfloat16

    def __init__(self, x16, part, n_parts, ln):
        self.x16, self.part, self.n_parts, self.ln = x16, part, n_parts, ln

    @property
    def shape(self):
        return self.x16.shape

    @property
    def device(self):
        return self.x16.device

    def size(self, i=None):
        return self.x16.size() if i is None else self.x16.size(i)

    def dim(self):
        return self.x16.dim()


_FOLDED = {}         # (weight id, ln id) -> (stamp, W' bf16, column sums f32, bias' f32)


def folded_weight(weight, bias, ln):
    """W' = bf16(W diag(a)), c = row sums of the ROUNDED W' (fp32), bias' = bias + W b: LN(x) W^T + bias =
    rinv * (x W'^T - mean * c) + bias'.  Cached; rebuilt when the weight, the bias or the norm's parameters change."""
    key = (id(weight), id(ln), None if bias is None else id(bias))
    stamp = (weight.data_ptr(), weight._version, ln.a_2.data_ptr(), ln.a_2._version, ln.b_2._version,
             None if bias is None else (bias.data_ptr(), bias._version))
    hit = _FOLDED.get(key)
    if hit is None or hit[0] != stamp:
        global WEIGHT_EPOCH
        WEIGHT_EPOCH += 1
        with torch.no_grad():
            w = weight.detach().float()
            wf = cast_bf16((w * ln.a_2.detach().float()[None, :]).contiguous())
            c = wf.float().sum(1).contiguous()
            bf = torch.mv(w, ln.b_2.detach().float())
            if bias is not None:
                bf = bf + bias.detach().float()
            hit = (stamp, wf, c, bf.contiguous(), weight, ln, bias)
        _FOLDED[key] = hit
    return hit[1], hit[2], hit[3]


def linear_folded(f, weight, bias, alpha, act, out_bf16):
    """act(LN(x) W^T + bias) * alpha with x given as a FoldedLN (raw bf16 rows + partial row sums)."""
    x16 = f.x16
    K = x16.shape[-1]
    N = weight.shape[0]
    wf, c, bf = folded_weight(weight, bias, f.ln)
    x2 = x16.reshape(-1, K)
    M = x2.shape[0]
    lead = x16.shape[:-1]
    o32 = None if out_bf16 else torch.empty(*lead, N, dtype=torch.float32, device=x16.device)
    o16 = torch.empty(*lead, N, dtype=torch.bfloat16, device=x16.device) if out_bf16 else None
    call("cn_gemm_bf16_folded", ptr(x2), x2.stride(0), ptr(wf), wf.stride(0), ptr(bf), ptr(c), ptr(f.part), int(f.n_parts), K,
         float(f.ln.eps), float(alpha), ACT[act], ptr(o32), ptr(o16), N, M, N, K, stream(),
         work=dict(flops=2.0 * M * N * K, bytes=2.0 * (M * K + N * K) + M * N * (2.0 if out_bf16 else 4.0)),
         tag=(f"cn_gemm_bf16[{M}x{N}x{K}]" if _lib.PROFILE_SHAPES else "cn_gemm_bf16"))
    return o16 if out_bf16 else o32


def linear_bf16_rowpart(x, weight_bf16, bias, residual, alpha, res_rows, ln):
    """The residual GEMM as the producer of a folded LayerNorm: fp32 rows + their raw bf16 copy + per-row partial sums
    (cn_gemm_bf16_rowpart); the pair rides on the fp32 tensor as `_cn_fold` where `ln.for_gemm` finds it."""
    K = x.shape[-1]
    N = weight_bf16.shape[0]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    M = x2.shape[0]
    lead = x.shape[:-1]
    dev = x.device
    o32 = torch.empty(*lead, N, dtype=torch.float32, device=dev)
    o16 = torch.empty(*lead, N, dtype=torch.bfloat16, device=dev)
    n_parts = int(_lib.load().cn_gemm_rowpart_count(N, K))
    part = torch.empty(M, n_parts, 2, dtype=torch.float32, device=dev)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, N)
        if r2.stride(-1) != 1 or r2.stride(0) % 4 or r2.data_ptr() % 16:
            r2 = r2.contiguous()
    call("cn_gemm_bf16_rowpart", ptr(x2), x2.stride(0), ptr(weight_bf16), weight_bf16.stride(0), ptr(bias), ptr(r2),
         r2.stride(0) if r2 is not None else 0, int(res_rows), float(alpha), ptr(o32), ptr(o16), N, ptr(part), M, N, K, stream(),
         work=dict(flops=2.0 * M * N * K, bytes=2.0 * (M * K + N * K) + M * N * (6.0 + (4.0 if r2 is not None else 0.0))),
         tag=(f"cn_gemm_bf16[{M}x{N}x{K}]" if _lib.PROFILE_SHAPES else "cn_gemm_bf16"))
    o32._cn_fold = (ln, o16, part, n_parts)
    return o32


def linear_forward(x, weight, bias=None, residual=None, alpha=1.0, act=None, res_rows=0, out=None, out_bf16=False,
                   out_split=False, next_ln=None):
    """next_ln (fast mode): the LayerNorm module that will normalise this output next (the following sublayer's
    pre-norm).  Its bf16 result is produced by the same kernel and attached to the returned tensor, where
    LayerNorm.for_gemm finds it (cn_gemm_bf16_ln)."""
    if isinstance(x, FoldedLN):
        assert residual is None and out is None
        return linear_folded(x, weight, bias, alpha, act, out_bf16)
    require_cuda(x, weight)
    if _PRECISION == "bf16x3" and x.shape[-1] % 64 == 0:
        return linear_x3(x, weight, bias, residual, alpha, act, res_rows, out_bf16, out_split)
    if _PRECISION != "fp32" and x.shape[-1] % 8 == 0:
        N = weight.shape[0]
        rows = x.numel() // x.shape[-1]
        if (FOLD_LN and next_ln is not None and _PRECISION == "bf16" and not out_bf16 and act is None and N % 8 == 0 and
                N == next_ln.a_2.numel()):
            return linear_bf16_rowpart(cast_bf16(x), bf16_weight(weight), bias, residual, alpha, res_rows, next_ln)
        if ((FUSE_LN or rows <= FUSE_LN_MAX_ROWS) and next_ln is not None and _PRECISION == "bf16" and not out_bf16 and
                act is None and N <= 512 and N % 4 == 0 and N >= 8):
            return linear_bf16_ln(cast_bf16(x), bf16_weight(weight), bias, residual, alpha, res_rows, next_ln)
        return linear_bf16(cast_bf16(x), bf16_weight(weight), bias, residual, alpha, act, res_rows,
                           torch.bfloat16 if out_bf16 else torch.float32)
    if x.dtype != torch.float32:
        x = x.float()
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    if out is None:
        out = torch.empty(*x.shape[:-1], N, dtype=torch.float32, device=x.device)
    o2 = out.reshape(-1, N)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, N)
        if r2.stride(-1) != 1:
            r2 = r2.contiguous()
    w = weight if weight.stride(-1) == 1 else weight.contiguous()
    call("cn_gemm_f32", ptr(x2), x2.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(r2),
         r2.stride(0) if r2 is not None else 0, int(res_rows), float(alpha), ACT[act], ptr(o2), o2.stride(0),
         M, N, K, stream(), work=dict(flops=2.0 * M * N * K, bytes=4.0 * (M * K + N * K + M * N)))
    return out


def layer_norm(x, a_2, b_2, eps=1e-6, bf16=False):
    """LayerNorm of modules/norm.py:15-18 (unbiased std, eps added to std)."""
    if _needs_grad(x, a_2, b_2):
        from . import grad
        return grad.LayerNormFn.apply(x, a_2, b_2, eps, bf16)
    return layer_norm_forward(x, a_2, b_2, eps, bf16)


def layer_norm_forward(x, a_2, b_2, eps=1e-6, bf16=False, split_ok=True):
    require_cuda(x)
    d = x.shape[-1]
    x2 = x.reshape(-1, d)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    if bf16 and split_ok and _PRECISION == "bf16x3" and d % 4 == 0:
        full = torch.empty(*x.shape[:-1], 2 * d, dtype=torch.bfloat16, device=x.device)
        call("cn_layernorm_bf16x2", ptr(x2), x2.shape[0], d, x2.stride(0), ptr(a_2), ptr(b_2), float(eps), None, 0, ptr(full),
             2 * d, stream(), work=dict(bytes=float(x2.shape[0] * d * 8)))
        return _hi_view(full)
    out = torch.empty(x.shape, dtype=torch.bfloat16 if bf16 else torch.float32, device=x.device)
    o2 = out.reshape(-1, d)
    call("cn_layernorm", ptr(x2), x2.shape[0], d, x2.stride(0), ptr(a_2), ptr(b_2), float(eps),
         None if bf16 else ptr(o2), ptr(o2) if bf16 else None, o2.stride(0), stream(),
         work=dict(bytes=float(x2.shape[0] * d * (4 + (2 if bf16 else 4)))))
    return out


def layer_norm_for_head(x, a_2, b_2, eps=1e-6):
    """LayerNorm in fp32 (what the API returns) plus, in one pass, its split bf16 twin for a decision-safe head:
    the f32 result carries `_cn_split16` = [.., 2d] bf16 [hi | lo], which head_rowstats consumes."""
    require_cuda(x)
    d = x.shape[-1]
    x2 = x.reshape(-1, d)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    full = torch.empty(*x.shape[:-1], 2 * d, dtype=torch.bfloat16, device=x.device)
    call("cn_layernorm_bf16x2", ptr(x2), x2.shape[0], d, x2.stride(0), ptr(a_2), ptr(b_2), float(eps), ptr(out), d, ptr(full),
         2 * d, stream(), work=dict(bytes=float(x2.shape[0] * d * 12)))
    out._cn_split16 = full
    return out


def _mask_strides(mask, B, Tq, Tk):
    if mask is None:
        return None, 0, 0
    m = mask
    if m.dtype != torch.uint8:
        m = m.view(torch.uint8) if m.dtype == torch.bool else (m != 0).to(torch.uint8)
    if m.dim() == 2:
        m = m.unsqueeze(1)
    assert m.shape[0] == B and m.shape[-1] == Tk and m.shape[1] in (1, Tq), tuple(m.shape)
    m = m.contiguous()
    return m, m.stride(0), (0 if m.shape[1] == 1 else m.stride(1))


# ----------------------------------------------------------------------------- dropout (training)
DROPOUT_SEED = 0             # host seed, used when DROPOUT_SEED_DEV is None
DROPOUT_SEED_DEV = None      # optional int64 [1] device tensor: a CUDA-graph-replayed training step advances it in-graph
_dropout_site = 0            # running id of the dropout call sites of one step (every site draws an independent mask)


def dropout_begin_step(seed=None):
    """Call at the start of every training forward: resets the call-site counter (so that forward and its recorded
    backward agree on the site ids) and optionally sets the host seed of the step."""
    global _dropout_site, DROPOUT_SEED
    _dropout_site = 0
    if seed is not None:
        DROPOUT_SEED = int(seed)


def _next_site():
    global _dropout_site
    _dropout_site += 1
    return _dropout_site


def dropout_forward(x, p, site, inplace=False):
    """x * keep / (1 - p) with the stateless keep mask of (seed, site, element index) (cn_dropout)."""
    require_cuda(x)
    xc = x if x.is_contiguous() else x.contiguous()
    if xc.dtype not in (torch.float32, torch.bfloat16):
        xc = xc.float()
    out = xc if (inplace and xc is x) else torch.empty_like(xc)
    call("cn_dropout", ptr(xc), int(xc.dtype == torch.bfloat16), ptr(out), int(out.dtype == torch.bfloat16), xc.numel(),
         float(p), ptr(DROPOUT_SEED_DEV), int(DROPOUT_SEED), int(site), stream(),
         work=dict(bytes=2.0 * xc.numel() * xc.element_size()))
    return out


def dropout_add_forward(x, residual, p, scale, site):
    """residual + scale * dropout(x) (residual None: scale * dropout(x)) in one launch, f32 out; the mask of `site`."""
    require_cuda(x)
    xc = x if x.is_contiguous() else x.contiguous()
    if xc.dtype not in (torch.float32, torch.bfloat16):
        xc = xc.float()
    rc = None
    if residual is not None:
        rc = residual if (residual.is_contiguous() and residual.dtype == torch.float32) else residual.float().contiguous()
        assert rc.shape == xc.shape
    out = torch.empty(xc.shape, dtype=torch.float32, device=xc.device)
    call("cn_dropout_add", ptr(xc), int(xc.dtype == torch.bfloat16), ptr(rc), ptr(out), xc.numel(), float(p), float(scale),
         ptr(DROPOUT_SEED_DEV), int(DROPOUT_SEED), int(site), stream(),
         work=dict(bytes=float(xc.numel()) * (xc.element_size() + 4 + (4 if rc is not None else 0))))
    return out


def dropout_add(x, residual, p, scale=1.0):
    """residual + scale * nn.Dropout(p)(x) of SublayerConnection (modules/utils.py:32), training mode, one launch (and
    one in backward: the branch gradient is scale * mask * g, the residual's is g itself)."""
    site = _next_site()
    if _needs_grad(x, residual):
        from . import grad
        return grad.DropoutAdd.apply(x, residual, float(p), float(scale), site)
    return dropout_add_forward(x, residual, p, scale, site)


def dropout_mask(shape, p, site, device):
    """The multiplicative mask (0 or 1/(1-p)) the kernels use for `site` under the current seed (tests)."""
    return dropout_forward(torch.ones(shape, dtype=torch.float32, device=device), p, site)


def dropout(x, p, training=True):
    """nn.Dropout of the reference's modules; identity in eval mode or at p = 0."""
    if not training or p <= 0.0:
        return x
    site = _next_site()
    if _needs_grad(x):
        from . import grad
        return grad.Dropout.apply(x, float(p), site)
    return dropout_forward(x, p, site)


REL_TC_MAX = 22           # largest max_rel of the tcgen05 relative attention (table rows staged in shared memory)


def attention(q, k, v, mask, n_head, rel=None, out=None, p_drop=0.0):
    """Masked multi-head attention on projected q [B,Tq,d], k/v [B,Tk,d] (views with a row stride
    are fine, e.g. slices of a fused QKV buffer).  mask: [B,1,Tk] or [B,Tq,Tk], 0 = masked.
    rel = (rel_pos [2m+1,d], pos_bias_u [h,dk], pos_bias_v [h,dk], m) selects the relative-position
    variant (modules/attention.py:104-143).  p_drop > 0 (training): dropout on the attention probabilities
    (attention.py:22-23,140-141) through the materialised path of grad.py."""
    site = _next_site() if p_drop > 0.0 else 0
    if p_drop > 0.0 or _needs_grad(q, k, v) or (rel is not None and _needs_grad(rel[0], rel[1], rel[2])):
        from . import grad
        assert out is None
        if rel is not None:
            return grad.RelAttention.apply(q, k, v, rel[0], rel[1], rel[2], mask, n_head, rel[3], float(p_drop), site)
        return grad.Attention.apply(q, k, v, mask, n_head, float(p_drop), site)
    return attention_forward(q, k, v, mask, n_head, rel, out)


def attention_forward(q, k, v, mask, n_head, rel=None, out=None):
    require_cuda(q, k, v)
    B, Tq, d = q.shape
    Tk = k.shape[1]
    dk = d // n_head
    for t in (q, k, v):
        assert t.stride(-1) == 1
    if getattr(q, "_cn_lo", None) is not None:
        key_mask = mask is None or mask.dim() == 2 or mask.shape[1] == 1
        if rel is None and key_mask and dk == 64 and Tk <= 2048 and getattr(k, "_cn_lo", None) is not None and \
                getattr(v, "_cn_lo", None) is not None and out is None:
            return _attention_x3(q, k, v, mask, n_head)
        q, k, v = unsplit(q), unsplit(k), unsplit(v)          # (relative term / per-row masks: fp32 kernel)
    if q.dtype == torch.bfloat16:
        key_mask = mask is None or mask.dim() == 2 or mask.shape[1] == 1
        if dk == 64 and Tk <= (2048 if key_mask else 512) and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16 and \
                (rel is None or (n_head <= 8 and rel[3] <= REL_TC_MAX and key_mask)):
            return _attention_bf16(q, k, v, mask, n_head, rel)      # rel: q must already hold q + pos_bias_u
        if rel is not None and key_mask:
            raise RuntimeError("bf16 relative attention outside the tcgen05 kernel's range (h<=8, m<=22)")
        q, k, v = q.float(), k.float(), v.float()        # (rel + per-row mask: fp32 kernel, q without pos_bias_u)
    if out is None:
        out = torch.empty(B, Tq, d, dtype=torch.float32, device=q.device)
    m, mb, mi = _mask_strides(mask, B, Tq, Tk)
    rp = ru = rv = None
    mr = 0
    if rel is not None:
        rp, ru, rv, mr = rel[:4]
        rp, ru, rv = rp.contiguous(), ru.contiguous(), rv.contiguous()
    call("cn_attention_f32", ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0),
         ptr(v), v.stride(1), v.stride(0), ptr(m), mb, mi, ptr(rp), ptr(ru), ptr(rv), int(mr),
         B, n_head, Tq, Tk, dk, ptr(out), out.stride(1), out.stride(0), stream(),
         work=dict(flops=4.0 * B * Tq * Tk * d, bytes=4.0 * B * d * (2 * Tq + 2 * Tk)))
    return out


def segment_attention(q_rows, k, v, seg, ylens, src_mask, src_size, left, right, n_head, out=None, q_len=None):
    """Extractor attention with the trigger mask evaluated on the fly from segment ids
    (cassnat.py:465-468 + modules/attention.py:13-24).  q_rows [U,d] is batch independent.
    bf16 operands (fast mode, d_k = 64, T <= 512): the mask is generated bit-packed from the segment ids
    and fused into the tcgen05 attention; otherwise the fp32 segment-sparse kernel."""
    require_cuda(q_rows, k, v, seg)
    U, d = q_rows.shape
    B, T, _ = k.shape
    if any(getattr(t, "_cn_lo", None) is not None for t in (q_rows, k, v)):
        q_rows, k, v = unsplit(q_rows), unsplit(k), unsplit(v)
    if q_rows.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16 and \
            d // n_head == 64 and T <= 512 and out is None:
        return _segment_attention_bf16(q_rows, k, v, seg, ylens, src_mask, src_size, left, right, n_head, q_len)
    if q_rows.dtype != torch.float32:
        q_rows, k, v = q_rows.float(), k.float(), v.float()
    if out is None:
        out = torch.empty(B, U, d, dtype=torch.float32, device=k.device)
    ws_bytes = _lib.load().cn_segment_attention_workspace_bytes(B, U, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=k.device)
    mask = _u8(src_mask, B, T)
    yl64, sz64 = _i64(ylens), _i64(src_size)           # converted copies must outlive the call (no inline temporaries)
    call("cn_segment_attention_f32", ptr(q_rows), q_rows.stride(0), ptr(k), k.stride(1), k.stride(0),
         ptr(v), v.stride(1), v.stride(0), ptr(seg), ptr(yl64), ptr(mask), ptr(sz64),
         int(left > 0), int(right > 0), B, n_head, U, T, d // n_head, ptr(out), out.stride(1), out.stride(0),
         ptr(ws), ws_bytes, stream(), work=dict(flops=4.0 * B * T * d, bytes=4.0 * B * d * (3 * T + U)))
    return out


def glu_dwconv_groupnorm_act(y, dw_weight, dw_bias, gamma, beta, eps, act, out_bf16=False):
    """Core of the conformer ConvModule (conformer_related.py:33-38) on y [B,L,2C] (f32 or bf16)
    -> [B,L,C] (f32, or bf16 for the next tensor-core GEMM)."""
    if _needs_grad(y, dw_weight, dw_bias, gamma, beta):
        from . import grad
        return grad.ConvCore.apply(y, dw_weight, dw_bias, gamma, beta, eps, act, out_bf16)
    return glu_dwconv_groupnorm_act_forward(y, dw_weight, dw_bias, gamma, beta, eps, act, out_bf16)


def glu_dwconv_groupnorm_act_forward(y, dw_weight, dw_bias, gamma, beta, eps, act, out_bf16=False):
    require_cuda(y)
    B, L, C2 = y.shape
    C = C2 // 2
    assert y.stride(-1) == 1
    ksz = dw_weight.shape[-1]
    out = torch.empty(B, L, C, dtype=torch.float32, device=y.device)
    o16 = torch.empty(B, L, C, dtype=torch.bfloat16, device=y.device) if out_bf16 else None
    ws_bytes = _lib.load().cn_conv_module_workspace_bytes(B, L)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=y.device)
    call("cn_glu_dwconv_groupnorm_act", ptr(y), int(y.dtype == torch.bfloat16), y.stride(1), y.stride(0),
         ptr(dw_weight.reshape(C, ksz).contiguous()), ptr(dw_bias), ptr(gamma), ptr(beta), float(eps), ACT[act],
         B, L, C, ksz, ptr(out), ptr(o16), ptr(ws), ws_bytes, stream(),
         work=dict(bytes=float(B * L * C) * (2 * y.element_size() + 8 + (2 if out_bf16 else 4))))
    return o16 if out_bf16 else out


# ----------------------------------------------------------------------------- bf16 tensor-core path
def cast_bf16_split(x, tag=True, out=None):
    """f32 [.., K] -> split bf16 operand [.., 2K] = [hi | lo] (see cn_gemm_bf16x3).  Returns the hi half (a strided
    [.., K] view, usable wherever a plain bf16 tensor is) tagged with the full tensor; tag=False returns the full one."""
    require_cuda(x)
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 4 or x2.data_ptr() % 16 or x2.dtype != torch.float32:
        x2 = x2.float().contiguous()
    full = out if out is not None else torch.empty(*x.shape[:-1], 2 * K, dtype=torch.bfloat16, device=x.device)
    assert full.shape == (*x.shape[:-1], 2 * K) and full.dtype == torch.bfloat16 and full.is_contiguous()
    call("cn_cast_bf16_split", ptr(x2), x2.stride(0), ptr(full), 2 * K, x2.shape[0], K, stream(),
         work=dict(bytes=8.0 * x2.shape[0] * K))
    return _hi_view(full) if tag else full


def linear_x3(x, weight, bias=None, residual=None, alpha=1.0, act=None, res_rows=0, out_bf16=False, out_split=False):
    """bf16x3 GEMM: x is the hi half of a split operand (three passes), an f32 tensor (split here) or a plain bf16
    tensor such as an attention context (two passes: only the weight is split)."""
    K = x.shape[-1]
    N = weight.shape[0]
    w = bf16_weight_split(weight)
    a_split = 1
    if x.dtype == torch.bfloat16:
        full = split_of(x)
        if full is None:
            a_split, full = 0, x
    else:
        full = cast_bf16_split(x, tag=False)
    a2 = full.reshape(-1, full.shape[-1])
    if a2.stride(-1) != 1 or a2.stride(0) % 8 or a2.data_ptr() % 16:
        a2 = a2.contiguous()
    M = a2.shape[0]
    lead = x.shape[:-1]
    dev = x.device
    o32 = o16 = None
    if out_bf16:
        o16 = torch.empty(*lead, 2 * N if out_split else N, dtype=torch.bfloat16, device=dev)
    else:
        o32 = torch.empty(*lead, N, dtype=torch.float32, device=dev)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, N)
        if r2.stride(-1) != 1:
            r2 = r2.contiguous()
    split_out = bool(out_bf16 and out_split)
    call("cn_gemm_bf16x3", ptr(a2), a2.stride(0), a_split, ptr(w), w.stride(0), ptr(bias), ptr(r2),
         r2.stride(0) if r2 is not None else 0, int(res_rows), float(alpha), ACT[act], ptr(o32), ptr(o16), int(split_out),
         2 * N if split_out else N, M, N, K, stream(),
         work=dict(flops=2.0 * M * N * K, bytes=2.0 * (M * K + N * K) * 2 + M * N * (4.0 if o32 is not None else 2.0)),
         tag=(f"cn_gemm_bf16x3[{M}x{N}x{K}]" if _lib.PROFILE_SHAPES else None))
    if o32 is not None:
        return o32
    return _hi_view(o16) if split_out else o16


def cast_bf16(x, split=False, out=None):
    """f32 -> bf16 copy (operands of the tensor-core GEMMs that arrive in fp32).  split=True also produces the lo
    half (the returned tensor is the hi half of a split operand, see cast_bf16_split).  out: a caller-owned buffer of
    the result's full shape ([.., 2K] for a split operand) -- used when it can be written directly."""
    require_cuda(x)
    if x.dtype == torch.bfloat16:
        return x
    if split and not _needs_grad(x) and x.shape[-1] % 4 == 0:
        return cast_bf16_split(x, out=out)
    if _needs_grad(x):
        from . import grad
        return grad.CastBf16.apply(x)
    xc = x.contiguous()
    if out is None or out.shape != xc.shape or out.dtype != torch.bfloat16 or not out.is_contiguous():
        out = torch.empty(xc.shape, dtype=torch.bfloat16, device=x.device)
    n = xc.numel()
    if n % 4 == 0:
        call("cn_cast_bf16", ptr(xc), ptr(out), n, stream(), work=dict(bytes=6.0 * n))
    else:
        out.copy_(xc)
    return out


def linear_bf16(x, weight_bf16, bias=None, residual=None, alpha=1.0, act=None, res_rows=0, out_dtype=torch.float32,
                also_bf16=False):
    """Tensor-core GEMM: residual + alpha * act(x @ W.T + bias), bf16 operands, fp32 accumulation.
    x bf16 [..., K] (row stride multiple of 8), weight_bf16 [N, K].  Returns f32 or bf16 (or both
    with also_bf16: the fp32 residual stream plus its bf16 copy for the next GEMM)."""
    require_cuda(x, weight_bf16)
    assert x.dtype == torch.bfloat16 and weight_bf16.dtype == torch.bfloat16
    K = x.shape[-1]
    N = weight_bf16.shape[0]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    M = x2.shape[0]
    lead = x.shape[:-1]
    o32 = torch.empty(*lead, N, dtype=torch.float32, device=x.device) if (out_dtype == torch.float32) else None
    o16 = torch.empty(*lead, N, dtype=torch.bfloat16, device=x.device) if (out_dtype == torch.bfloat16 or also_bf16) else None
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, N)
        if r2.stride(-1) != 1:
            r2 = r2.contiguous()
    call("cn_gemm_bf16", ptr(x2), x2.stride(0), ptr(weight_bf16), weight_bf16.stride(0), ptr(bias), ptr(r2),
         r2.stride(0) if r2 is not None else 0, int(res_rows), float(alpha), ACT[act], ptr(o32), ptr(o16), N,
         M, N, K, stream(),
         work=dict(flops=2.0 * M * N * K, bytes=2.0 * (M * K + N * K) + M * N * ((4.0 if o32 is not None else 0.0) +
                                                                              (2.0 if o16 is not None else 0.0) +
                                                                              (4.0 if r2 is not None else 0.0))),
         tag=(f"cn_gemm_bf16[{M}x{N}x{K}]" if _lib.PROFILE_SHAPES else None))
    if o32 is not None and o16 is not None:
        return o32, o16
    return o32 if o32 is not None else o16


def linear_bf16_ln(x, weight_bf16, bias, residual, alpha, res_rows, ln):
    """residual + alpha * (x W^T + bias) as fp32, plus ln(that) as bf16 from the same kernel; the bf16 tensor rides on
    the fp32 one as `_cn_ln = (ln module, tensor)`."""
    K = x.shape[-1]
    N = weight_bf16.shape[0]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    M = x2.shape[0]
    lead = x.shape[:-1]
    o32 = torch.empty(*lead, N, dtype=torch.float32, device=x.device)
    o16 = torch.empty(*lead, N, dtype=torch.bfloat16, device=x.device)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, N)
        if r2.stride(-1) != 1 or r2.stride(0) % 4 or r2.data_ptr() % 16:
            r2 = r2.contiguous()
    call("cn_gemm_bf16_ln", ptr(x2), x2.stride(0), ptr(weight_bf16), weight_bf16.stride(0), ptr(bias), ptr(r2),
         r2.stride(0) if r2 is not None else 0, int(res_rows), float(alpha), ptr(o32), N, ptr(ln.a_2), ptr(ln.b_2),
         float(ln.eps), ptr(o16), N, M, N, K, stream(),
         work=dict(flops=2.0 * M * N * K, bytes=2.0 * (M * K + N * K) + M * N * (4.0 + 2.0 + (4.0 if r2 is not None else 0.0))),
         tag=(f"cn_gemm_bf16_ln[{M}x{N}x{K}]" if _lib.PROFILE_SHAPES else None))
    o32._cn_ln = (ln, o16)
    return o32


def decode_hints(ylen, ymax, src_mask2, use_q=True):
    """All length hints of a decode step in one launch -> dict(need_rows, kv_tgt, last_key, order_tgt, order_src)
    of int32 [B] tensors (see cn_decode_hints)."""
    B, T = src_mask2.shape
    dev = src_mask2.device
    buf = torch.empty(5, B, dtype=torch.int32, device=dev)
    yl64, m8 = _i64(ylen), _u8(src_mask2, B, T)
    call("cn_decode_hints", ptr(yl64), int(ymax), ptr(m8), B, T, int(bool(use_q)),
         ptr(buf[0]), ptr(buf[1]), ptr(buf[2]), ptr(buf[3]), ptr(buf[4]), stream())
    return dict(need_rows=buf[0], kv_tgt=buf[1], last_key=buf[2], order_tgt=buf[3], order_src=buf[4])


def mark_causal(mask):
    """Tags a [B,U,U] mask whose entries above the diagonal are all zero (mask & tril): the tcgen05 attention then
    skips the key tiles above the diagonal (the mask bits are still applied inside the diagonal tiles)."""
    mask._cn_causal = True
    return mask


def attach_length_hints(mask, kv_len=None, q_len=None):
    """Tags a KEY mask ([B,1,Tk], prefix-shaped or not) with per-utterance hints for the tcgen05
    attention: kv_len[b] = (index of the last unmasked key) + 1, q_len[b] = number of leading query
    rows the caller needs.  Both int32 [B] device tensors; the kernel skips fully masked key tiles
    (exactly zero weight) and zero-fills unneeded query rows."""
    if kv_len is not None:
        mask._cn_kv_len = kv_len.to(torch.int32).contiguous()
    if q_len is not None:
        mask._cn_q_len = q_len.to(torch.int32).contiguous()
    if kv_len is not None or q_len is not None:
        # heaviest utterances first: (key tiles) x (query tiles) of 128, so that the static round-robin over the
        # persistent CTAs balances the ragged batch
        B = mask.shape[0]
        one = torch.ones(B, dtype=torch.int32, device=mask.device)
        kt = (mask._cn_kv_len + 127) // 128 if kv_len is not None else one
        qt = (mask._cn_q_len + 127) // 128 if q_len is not None else one
        mask._cn_order = torch.argsort(kt * qt, descending=True, stable=True).to(torch.int32).contiguous()
    return mask


def _segment_attention_bf16(q_rows, k, v, seg, ylens, src_mask, src_size, left, right, n_head, q_len=None):
    U, d = q_rows.shape
    B, T, _ = k.shape
    dev = k.device
    bits = torch.empty(B, U, 16, dtype=torch.int32, device=dev)
    yl64, m8, sz64 = _i64(ylens), _u8(src_mask, B, T), _i64(src_size)
    call("cn_trigger_mask_bits", ptr(seg), ptr(yl64), ptr(m8), ptr(sz64), B, T, U,
         int(left > 0), int(right > 0), ptr(bits), stream(), work=dict(bytes=4.0 * B * T + 64.0 * B * U))
    out = torch.empty(B, U, d, dtype=torch.bfloat16, device=dev)
    # no kv_len hint: the first row past eos is fully masked and must stay uniform over ALL T frames
    # (modules/attention.py:19-21); q_len only drops whole query tiles the caller never reads
    kv_len = None
    if q_len is not None:
        q_len = q_len.to(torch.int32).contiguous()
    if q_rows.stride(-1) != 1 or q_rows.stride(0) % 8:
        q_rows = q_rows.contiguous()
    call("cn_attention_bf16_bits", ptr(q_rows), q_rows.stride(0), 0, ptr(k), k.stride(1), k.stride(0),
         ptr(v), v.stride(1), v.stride(0), ptr(bits), B, n_head, U, T, d // n_head, ptr(out), out.stride(1), out.stride(0),
         ptr(kv_len), ptr(q_len), 0, stream(), work=dict(flops=4.0 * B * U * T * d, bytes=2.0 * B * d * (2 * T + U)))
    return out


def _attention_bf16(q, k, v, mask, n_head, rel=None):
    """tcgen05 attention: bf16 q/k/v (strided views of fused projection buffers are fine) -> bf16 ctx."""
    B, Tq, d = q.shape
    Tk = k.shape[1]
    out = torch.empty(B, Tq, d, dtype=torch.bfloat16, device=q.device)
    kv_len = getattr(mask, "_cn_kv_len", None) if mask is not None else None
    q_len = getattr(mask, "_cn_q_len", None) if mask is not None else None
    order = getattr(mask, "_cn_order", None) if mask is not None else None
    m, mb, mi = _mask_strides(mask, B, Tq, Tk)
    if m is not None and mi != 0:                    # per-row mask: bit-pack it for the tensor-core kernel
        if rel is not None:
            raise RuntimeError("bf16 relative attention takes key masks only")
        bits = getattr(mask, "_cn_bits", None)           # packed once per mask tensor (every layer reuses it)
        if bits is None or bits.shape[:2] != (B, Tq):
            bits = torch.empty(B, Tq, 16, dtype=torch.int32, device=q.device)
            call("cn_pack_mask_bits", ptr(m), mb, mi, B, Tq, Tk, ptr(bits), stream(), work=dict(bytes=float(B * Tq * (Tk + 64))))
            try:
                mask._cn_bits = bits
            except AttributeError:
                pass
        call("cn_attention_bf16_bits", ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0),
             ptr(v), v.stride(1), v.stride(0), ptr(bits), B, n_head, Tq, Tk, d // n_head, ptr(out), out.stride(1),
             out.stride(0), None, None, int(bool(getattr(mask, "_cn_causal", False)) and Tq == Tk), stream(),
             work=dict(flops=4.0 * B * Tq * Tk * d, bytes=2.0 * B * d * (2 * Tq + 2 * Tk)))
        return out
    if rel is not None and len(rel) > 4 and rel[4] is not None:
        tab, mr = rel[4], rel[3]                      # [B,Tq,h*(2m+1)] view of the fused projection's extra columns
        assert tab.dtype in (torch.bfloat16, torch.float32) and tab.stride(2) == 1
        call("cn_rel_attention_bf16_table", ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0),
             ptr(v), v.stride(1), v.stride(0), ptr(m), mb, mi, ptr(tab), int(tab.dtype == torch.bfloat16), tab.stride(1),
             tab.stride(0), 2 * int(mr) + 1, int(mr), B, n_head, Tq, Tk, d // n_head, ptr(out), out.stride(1),
             out.stride(0), ptr(kv_len), stream(),
             work=dict(flops=4.0 * B * Tq * Tk * d, bytes=2.0 * B * d * (2 * Tq + 2 * Tk)))
        return out
    if rel is not None:
        rp, ru, rv, mr = rel[:4]
        ws_bytes = _lib.load().cn_rel_attention_workspace_bytes(B, n_head, Tq, int(mr))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
        rp, ru, rv = rp.float().contiguous(), ru.contiguous(), rv.contiguous()
        call("cn_rel_attention_bf16", ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0),
             ptr(v), v.stride(1), v.stride(0), ptr(m), mb, mi, ptr(rp), ptr(ru),
             ptr(rv), int(mr), B, n_head, Tq, Tk, d // n_head, ptr(out), out.stride(1), out.stride(0),
             ptr(ws), ws_bytes, ptr(kv_len), stream(),
             work=dict(flops=4.0 * B * Tq * Tk * d, bytes=2.0 * B * d * (2 * Tq + 2 * Tk)))
        return out
    call("cn_attention_bf16", ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0),
         ptr(v), v.stride(1), v.stride(0), ptr(m), mb, mi, B, n_head, Tq, Tk, d // n_head,
         ptr(out), out.stride(1), out.stride(0), ptr(kv_len), ptr(q_len), ptr(order), stream(),
         work=dict(flops=4.0 * B * Tq * Tk * d, bytes=2.0 * B * d * (2 * Tq + 2 * Tk)))
    return out


def _attention_x3(q, k, v, mask, n_head):
    """Split-bf16 tcgen05 attention (precise mode): (hi, lo) bf16 Q/K/V -> context as the hi half of a split operand
    (what the three-pass output projection reads; unsplit() gives the fp32 value)."""
    B, Tq, d = q.shape
    Tk = k.shape[1]
    full = torch.empty(B, Tq, 2 * d, dtype=torch.bfloat16, device=q.device)
    out = _hi_view(full)
    kv_len = getattr(mask, "_cn_kv_len", None) if mask is not None else None
    q_len = getattr(mask, "_cn_q_len", None) if mask is not None else None
    order = getattr(mask, "_cn_order", None) if mask is not None else None
    m, mb, mi = _mask_strides(mask, B, Tq, Tk)
    ql, kl, vl = q._cn_lo, k._cn_lo, v._cn_lo
    for a, b_ in ((q, ql), (k, kl), (v, vl)):
        assert a.stride() == b_.stride() and a.shape == b_.shape
    call("cn_attention_bf16x3", ptr(q), ptr(ql), q.stride(1), q.stride(0), ptr(k), ptr(kl), k.stride(1), k.stride(0),
         ptr(v), ptr(vl), v.stride(1), v.stride(0), ptr(m), mb, B, n_head, Tq, Tk, d // n_head, None, ptr(full), d,
         full.stride(1), full.stride(0), ptr(kv_len), ptr(q_len), ptr(order), stream(),
         work=dict(flops=4.0 * B * Tq * Tk * d, bytes=4.0 * B * d * (2 * Tq + 2 * Tk)))
    return out


def conv_frontend(x, w1, b1, w2, b2):
    """relu(conv2(relu(conv1(x)))) of ConvEmbedding (embedding.py:103-104): x f32 [B,Tin,F] ->
    bf16 [B, T2, F2, C] channels-last (fast mode; conv2 = implicit GEMM on tcgen05)."""
    require_cuda(x, w1, w2)
    B, Tin, Fin = x.shape
    C = w1.shape[0]
    T1, F1 = (Tin - 1) // 2 + 1, (Fin - 1) // 2 + 1
    T2, F2 = (T1 - 1) // 2 + 1, (F1 - 1) // 2 + 1
    xc = x if (x.stride(-1) == 1 and x.dtype == torch.float32) else x.float().contiguous()
    key = ("conv2_taps", id(w2))
    stamp = (w2.data_ptr(), w2._version)
    hit = _BF16_WEIGHTS.get(key)
    if hit is None or hit[0] != stamp:                 # [C_out, C_in, 3, 3] -> [C_out, (dt, df, c_in)] bf16
        with torch.no_grad():
            hit = (stamp, cast_bf16(w2.detach().permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()), w2)
        _BF16_WEIGHTS[key] = hit
    mid = torch.empty(B, T1, F1, C, dtype=torch.bfloat16, device=x.device)
    out = torch.empty(B, T2, F2, C, dtype=torch.bfloat16, device=x.device)
    call("cn_conv_frontend_bf16", ptr(xc), xc.stride(0), xc.stride(1), B, Tin, Fin, ptr(w1.detach().reshape(C, 9).contiguous()),
         ptr(b1), ptr(hit[1]), ptr(b2), C, ptr(mid), ptr(out), stream(),
         work=dict(flops=2.0 * B * T2 * F2 * C * 9 * C + 18.0 * B * T1 * F1 * C, bytes=2.0 * B * C * (2 * T1 * F1 + T2 * F2)))
    return out


# ----------------------------------------------------------------------------- K10 CTC loss
def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, want_grad=True, grad_out=1.0):
    """CTC loss of cassnat.py:232 — torch.nn.CTCLoss(reduction='mean', zero_infinity=True) — on
    log_probs [B,T,V] (batch-first; the reference passes the transposed [T,B,V] view of the same
    tensor).  Returns (loss scalar, nll [B], d loss / d logits [B,T,V] or None); the gradient is with
    respect to the logits that log_probs = log_softmax(logits) came from."""
    require_cuda(log_probs)
    B, T, V = log_probs.shape
    lp = log_probs if log_probs.stride(-1) == 1 else log_probs.contiguous()
    tg = _i64(targets.to(lp.device))
    L = tg.shape[1]
    il, tl = _i64(input_lengths.to(lp.device)), _i64(target_lengths.to(lp.device))
    loss = torch.empty(1, dtype=torch.float32, device=lp.device)
    nll = torch.empty(B, dtype=torch.float32, device=lp.device)
    grad = torch.empty(B, T, V, dtype=torch.float32, device=lp.device) if want_grad else None
    ws_bytes = _lib.load().cn_ctc_loss_workspace_bytes(B, T, L)
    if ws_bytes < 0:
        raise RuntimeError(f"ctc_loss: unsupported target length L={L} (<= 511)")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=lp.device)
    call("cn_ctc_loss", ptr(lp), B, T, V, lp.stride(0), lp.stride(1), ptr(tg), L, tg.stride(0), ptr(il), ptr(tl),
         int(blank), float(grad_out), ptr(loss), ptr(nll), ptr(grad), grad.stride(0) if want_grad else 0,
         grad.stride(1) if want_grad else 0, ptr(ws), ws_bytes, stream(),
         work=dict(bytes=4.0 * B * T * V * (2 if want_grad else 0) + 8.0 * B * T * (2 * L + 1)))
    return loss[0], nll, grad


def head_rowstats(x, weight, bias=None, alpha=1.0, top2=False, target=None, safe=None):
    """K1/K9 fused (fast mode): argmax / runner-up / max / log-sum-exp of every row of
    alpha * (x @ weight.T + bias) straight out of the tcgen05 GEMM's epilogue; the logits are never
    written.  Returns the same dict as row_stats."""
    require_cuda(x, weight)
    K, N = x.shape[-1], weight.shape[0]
    # safe (default: SAFE_HEADS): run the head as bf16x3 so that the argmax / runner-up are those of the fp32 logits
    safe = (SAFE_HEADS if safe is None else safe) or _PRECISION == "bf16x3"
    x3 = safe and K % 64 == 0 and (x.dtype != torch.bfloat16 or split_of(x) is not None)
    if x3:                        # decision-safe head: split operands, three passes
        twin = getattr(x, "_cn_split16", None)                  # produced together with x by layer_norm_for_head
        if x.dtype == torch.bfloat16:
            full = split_of(x)
        elif twin is not None and twin.shape[:-1] == x.shape[:-1] and twin.shape[-1] == 2 * K:
            full = twin
        else:
            full = cast_bf16_split(x, tag=False)
        wb = bf16_weight_split(weight)
        x2 = full.reshape(-1, 2 * K)
    else:
        xb, wb = cast_bf16(x), bf16_weight(weight)
        x2 = xb.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    M = x2.shape[0]
    dev = x.device
    out = {"top1": torch.empty(M, dtype=torch.int32, device=dev), "peak": torch.empty(M, dtype=torch.float32, device=dev),
           "lse": torch.empty(M, dtype=torch.float32, device=dev)}
    if top2:
        out["top2"] = torch.empty(M, dtype=torch.int32, device=dev)
    tgt = None
    if target is not None:                 # also return z[row, target[row]] (0 where the id is outside [0, N))
        tgt = target.reshape(-1).to(torch.int32).contiguous()
        out["target_logit"] = torch.zeros(M, dtype=torch.float32, device=dev)
    ws_bytes = _lib.load().cn_gemm_rowstats_workspace_bytes(M, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    name = "cn_gemm_bf16x3_rowstats" if x3 else "cn_gemm_bf16_rowstats"
    call(name, ptr(x2), x2.stride(0), ptr(wb), wb.stride(0), ptr(bias), float(alpha), M, N, K,
         ptr(out["top1"]), ptr(out.get("top2")), ptr(out["peak"]), ptr(out["lse"]), ptr(tgt), ptr(out.get("target_logit")),
         ptr(ws), ws_bytes, stream(),
         work=dict(flops=2.0 * M * N * K, bytes=2.0 * (M * K + N * K) * (2 if x3 else 1)),
         tag=(f"{name}[{M}x{N}x{K}]" if _lib.PROFILE_SHAPES else None))
    lead = x.shape[:-1]
    return {k: v.reshape(lead) for k, v in out.items()}
