import sys, torch
sys.path.insert(0,'.')
import bench
from cassnat_asr_public_b200 import _lib
from cassnat_asr_public_b200.cassnat import make_model
w = bench.WORKLOADS['cfg3']; args = bench.model_args(w)
args.use_conv_dec=True; args.pos_type='relative'; args.use_conv_enc=True; args.d_decff=1024; args.d_encff=1024
torch.manual_seed(1234); dev='cuda:0'
model = make_model(80,args).to(dev).eval(); model.set_precision('bf16')
enc_h, mask, ratio, lens = bench.synth_encoder_output(w, 64, 1234)
inp = [t.to(dev) for t in (enc_h,mask,ratio)]
for _ in range(2):
    with torch.no_grad(): model.decode_from_encoder(*inp, bench.VOCAB, args)
_lib.PROFILE_SHAPES=True
_lib.start_profile()
with torch.no_grad(): model.decode_from_encoder(*inp, bench.VOCAB, args)
prof=_lib.stop_profile()
tot=sum(v['ms'] for v in prof.values())
for k,v in sorted(prof.items(), key=lambda kv:-kv[1]['ms'])[:14]:
    print(f"{v['ms']:8.3f} ms x{v['calls']:3d}  {k}")
print("total", tot)
import cProfile, pstats, time
torch.cuda.synchronize(); t0=time.perf_counter()
pr = cProfile.Profile(); pr.enable()
with torch.no_grad(): model.decode_from_encoder(*inp, bench.VOCAB, args)
torch.cuda.synchronize(); pr.disable(); print("wall ms", (time.perf_counter()-t0)*1e3)
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
