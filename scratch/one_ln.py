import torch, sys
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
dev = 'cuda:0'
ops.set_precision("bf16")
x = torch.randn(32000, 512, device=dev); a = torch.ones(512, device=dev); b = torch.zeros(512, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): y = ops.layer_norm_forward(x, a, b, 1e-6, True)
torch.cuda.synchronize()
ms = 0
for _ in range(10):
    flush.fill_(0)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); y = ops.layer_norm_forward(x, a, b, 1e-6, True); e.record(); torch.cuda.synchronize(); ms += s.elapsed_time(e)
print("LN cold ms", ms / 10, y.dtype)
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): y = ops.layer_norm_forward(x, a, b, 1e-6, True)
e.record(); torch.cuda.synchronize(); print("LN warm (L2) ms", s.elapsed_time(e) / 10)
