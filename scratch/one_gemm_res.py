import sys, torch
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
DEV = 'cuda:0'
M, N, K = 32000, 512, 512
x = torch.randn(M, K, device=DEV).bfloat16(); w = torch.randn(N, K, device=DEV).bfloat16(); b = torch.randn(N, device=DEV)
r = torch.randn(M, N, device=DEV)
for _ in range(3): y = ops.linear_bf16(x, w, b, residual=r)
torch.cuda.synchronize()
print("ok", float(y[0, 0]))
