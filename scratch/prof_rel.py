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
for _ in range(3): o = ops.attention(q, k, v, mask, h, (P, u, vb, m))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    for _ in range(5): o = ops.attention(q, k, v, mask, h, (P, u, vb, m))
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=80))
