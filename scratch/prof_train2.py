import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from cassnat_asr_public_b200 import _lib, ops
from cassnat_asr_public_b200.cassnat import make_model
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["cfg3"]
args = bench.model_args(w, precision="bf16")
drop = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
args.dropout = drop
torch.manual_seed(1234)
model = make_model(80, args).to(dev)
model.set_precision("bf16")
model.train()
if drop == 0: model.zero_dropout()
B, T, L = 32, 250, 40
g = torch.Generator().manual_seed(4321)
enc_h = torch.randn(B, T, 512, generator=g).to(dev).requires_grad_(True)
lens = torch.randint(200, T + 1, (B,), generator=g); lens[0] = T
x_mask = (torch.arange(T)[None] < lens[:, None]).unsqueeze(1).to(dev)
ratio = ((lens.float() + 0.5) / T).to(dev)
labels = torch.randint(4, 5000, (B, L + 1), generator=g); labels[:, L] = 2
labels, sizes = labels.to(dev), torch.full((B,), L, dtype=torch.int64, device=dev)
params = [p for n, p in model.named_parameters() if not n.startswith(("src_embed.", "encoder."))]
def one():
    for p in params: p.grad = None
    enc_h.grad = None
    out = model.train_forward_from_encoder(enc_h, x_mask, ratio, labels, sizes, args)
    out[2].backward()
for _ in range(3): one()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    one(); torch.cuda.synchronize()
rows = [(e.key, e.count, e.self_device_time_total) for e in p.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print("total CUDA us", tot, "kernel launches", sum(r[1] for r in rows))
for k, c, t in rows[:45]:
    print(f"{t:9.1f} us x{c:4d}  {k[:110]}")
