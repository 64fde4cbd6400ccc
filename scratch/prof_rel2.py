import sys, torch
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
dev = 'cuda:0'
B, U, h, m = 64, 500, 8, 8
d = 64 * h
qkv = torch.randn(B, U, 3 * d, device=dev).bfloat16()
q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
P = torch.randn(2 * m + 1, d, device=dev); u = torch.randn(h, 64, device=dev) * 0.1; vb = torch.randn(h, 64, device=dev) * 0.1
lens = torch.randint(300, 501, (B,), device=dev); mask = (torch.arange(U, device=dev)[None] < lens[:, None]).unsqueeze(1)
ops.attach_length_hints(mask, kv_len=lens)
def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("plain us", t(lambda: ops.attention(q, k, v, mask, h)))
print("rel   us (incl. table kernel)", t(lambda: ops.attention(q, k, v, mask, h, (P, u, vb, m))))
