"""GEMM shapes of the decode step, graph-replayed (no host launch overhead).  CN_GEMM_DBG: what-if switches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cassnat_asr_public_b200 import ops
DEV = 'cuda:0'
M = 32000
shapes = [("ffn1 bf16out relu", 2048, 512, False, True, "relu"), ("qkv bf16out", 1536, 512, False, True, None),
          ("dxd res f32out", 512, 512, True, False, None), ("ffn2 res f32out", 512, 2048, True, False, None),
          ("kvall bf16out", 6144, 512, False, True, None)]
for name, N, K, res, bf, act in shapes:
    x = torch.randn(M, K, device=DEV).bfloat16(); w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16(); b = torch.randn(N, device=DEV)
    r = torch.randn(M, N, device=DEV) if res else None
    f = lambda: ops.linear_bf16(x, w, b, r, 1.0, act, 0, torch.bfloat16 if bf else torch.float32)
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): f()
    g.replay(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); g.replay(); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f"dbg={os.environ.get('CN_GEMM_DBG','0')} {name:20s} [{M}x{N}x{K}] {ms*1e3:7.1f} us  {2.0*M*N*K/ms/1e9:7.0f} TF/s")
