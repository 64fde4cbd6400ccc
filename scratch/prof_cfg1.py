import sys, torch
sys.path.insert(0, '.')
import bench
from cassnat_asr_public_b200 import _lib
from cassnat_asr_public_b200.cassnat import make_model
w = bench.WORKLOADS['cfg1']; args = bench.model_args(w)
torch.manual_seed(1234); dev = 'cuda:0'
model = make_model(80, args).to(dev).eval(); model.set_precision('bf16')
enc_h, mask, ratio, lens = bench.synth_encoder_output(w, w['batch'], 1234)
inp = [t.to(dev) for t in (enc_h, mask, ratio)]
for _ in range(3):
    with torch.no_grad(): model.decode_from_encoder(*inp, bench.VOCAB, args)
_lib.PROFILE_SHAPES = True
_lib.start_profile()
with torch.no_grad(): model.decode_from_encoder(*inp, bench.VOCAB, args)
prof = _lib.stop_profile()
tot = sum(v['ms'] for v in prof.values()); n = sum(v['calls'] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:30]:
    print(f"{v['ms']*1e3:8.1f} us x{v['calls']:3d}  {v['ms']*1e3/v['calls']:6.1f} us/call  {k}")
print("total ms", tot, "calls", n)
model.use_cuda_graphs = True
with torch.no_grad():
    for _ in range(5): model.decode_from_encoder(*inp, bench.VOCAB, args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): model.decode_from_encoder(*inp, bench.VOCAB, args)
    e1.record(); torch.cuda.synchronize()
print("graph ms/step", e0.elapsed_time(e1) / 50)
