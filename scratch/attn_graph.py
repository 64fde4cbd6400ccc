"""Attention kernel time without host launch overhead: 20 calls captured in a CUDA graph and replayed
(scratch/diagnostic).  CN_ATTN_DBG selects what-if variants compiled into the kernel during tuning."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cassnat_asr_public_b200 import ops
DEV = 'cuda:0'
B, Tq, Tk, h = 64, int(os.environ.get("TQ", 500)), int(os.environ.get("TK", 500)), 8
d = 64 * h
q = torch.randn(B, Tq, d, device=DEV).bfloat16(); k = torch.randn(B, Tk, d, device=DEV).bfloat16(); v = torch.randn(B, Tk, d, device=DEV).bfloat16()
for name, mask in (("no-mask", None),):
    for _ in range(3):
        y = ops.attention(q, k, v, mask, h)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            y = ops.attention(q, k, v, mask, h)
    g.replay(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); g.replay(); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 40
    print(f"dbg={os.environ.get('CN_ATTN_DBG', '0')} {name} {Tq}x{Tk}: {ms * 1e3:.1f} us/call, {4.0 * B * Tq * Tk * d / ms / 1e9:.0f} TF/s")
