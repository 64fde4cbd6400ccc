import sys, torch
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
dev = 'cuda:0'
B, L, C, k = 64, 500, 512, (int(sys.argv[1]) if len(sys.argv) > 1 else 3)
y = torch.randn(B, L, 2 * C, device=dev).bfloat16()
w = torch.randn(C, 1, k, device=dev); b = torch.randn(C, device=dev); g = torch.ones(C, device=dev); be = torch.zeros(C, device=dev)
for _ in range(3): o = ops.glu_dwconv_groupnorm_act(y, w, b, g, be, 1e-5, "swish", out_bf16=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    for _ in range(5): o = ops.glu_dwconv_groupnorm_act(y, w, b, g, be, 1e-5, "swish", out_bf16=True)
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=80))
