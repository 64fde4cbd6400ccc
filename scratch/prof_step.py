import sys, torch
sys.path.insert(0,'.')
import bench
from cassnat_asr_public_b200 import _lib, ops
from cassnat_asr_public_b200.cassnat import make_model
w = bench.WORKLOADS['cfg3']; args = bench.model_args(w)
torch.manual_seed(1234); dev='cuda:0'
model = make_model(80,args).to(dev).eval(); model.set_precision('bf16')
enc_h, mask, ratio, lens = bench.synth_encoder_output(w, 64, 1234)
inp = [t.to(dev) for t in (enc_h,mask,ratio)]
for _ in range(3):
    with torch.no_grad(): model.decode_from_encoder(*inp, bench.VOCAB, args)
_lib.PROFILE_SHAPES=True
_lib.start_profile()
with torch.no_grad(): model.decode_from_encoder(*inp, bench.VOCAB, args)
prof=_lib.stop_profile()
tot=sum(v['ms'] for v in prof.values())
for k,v in sorted(prof.items(), key=lambda kv:-kv[1]['ms']):
    tf = v['flops']/v['ms']/1e9 if v['flops'] else 0; gb = v['bytes']/v['ms']/1e6
    print(f"{v['ms']:8.3f} ms x{v['calls']:3d}  {tf:7.0f} TF/s {gb:7.0f} GB/s  {k}")
print("total", tot)
