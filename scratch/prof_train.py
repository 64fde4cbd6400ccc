import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from cassnat_asr_public_b200 import _lib, ops
from cassnat_asr_public_b200.cassnat import make_model
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["cfg3"]
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
args = bench.model_args(w, precision=prec)
torch.manual_seed(1234)
model = make_model(80, args).to(dev)
model.set_precision(prec)
model.train(); model.zero_dropout()
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
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): one()
torch.cuda.synchronize(); print("wall ms/step", (time.perf_counter() - t0) * 100)
# forward only
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    out = model.train_forward_from_encoder(enc_h, x_mask, ratio, labels, sizes, args)
torch.cuda.synchronize(); print("fwd-only wall ms/step", (time.perf_counter() - t0) * 100)
_lib.PROFILE_SHAPES = True
_lib.start_profile(); one(); prof = _lib.stop_profile()
tot = sum(v["ms"] for v in prof.values()); n = sum(v["calls"] for v in prof.values())
print("our kernels: total ms", tot, "calls", n)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:22]:
    print(f"{k:44s} {v['calls']:4d} {v['ms']:8.3f}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    one(); torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
