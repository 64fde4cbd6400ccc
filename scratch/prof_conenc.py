import sys, torch
sys.path.insert(0, '.')
import bench
from cassnat_asr_public_b200 import _lib
from cassnat_asr_public_b200.cassnat import make_model
w = bench.WORKLOADS['cfg3']
args = bench.model_args(w, use_conv_dec=True, use_conv_enc=True, pos_type="relative", precision="bf16")
torch.manual_seed(1234); dev = 'cuda:0'
model = make_model(80, args).to(dev).eval(); model.set_precision('bf16')
B, Tin = 64, w["frames_in"]
src = torch.randn(B, Tin, 80, device=dev); mask = torch.ones(B, 1, Tin, dtype=torch.bool, device=dev)
with torch.no_grad():
    for _ in range(2):
        x, xm = model.src_embed(src, mask); enc_h = model.encoder(x, xm)
    _lib.PROFILE_SHAPES = True
    _lib.start_profile()
    enc_h = model.encoder(x, xm)
    prof = _lib.stop_profile()
tot = sum(v['ms'] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:16]:
    print(f"{v['ms']:8.3f} ms x{v['calls']:3d}  {k}")
print("total our kernels", tot)
from torch.profiler import profile, ProfilerActivity
with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as p:
    enc_h = model.encoder(x, xm); torch.cuda.synchronize()
rows = [(e.key, e.count, e.self_device_time_total) for e in p.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[2])
print("total CUDA us", sum(r[2] for r in rows))
for k, c, t in rows[:14]:
    print(f"{t:9.1f} us x{c:4d}  {k[:100]}")
