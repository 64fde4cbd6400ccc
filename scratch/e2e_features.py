"""Whole beam_decode(src, ...) from fbank features: encoder + front-end included."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from cassnat_asr_public_b200 import _lib
from cassnat_asr_public_b200.cassnat import make_model
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["cfg3"]
args = bench.model_args(w)
torch.manual_seed(1234)
model = make_model(80, args).to(dev).eval()
model.set_precision("bf16")
model.use_cuda_graphs = True
B, Tin = 64, 2000
g = torch.Generator().manual_seed(1)
src = torch.randn(B, Tin, 80, generator=g)
lens = torch.randint(1200, Tin + 1, (B,), generator=g); lens[0] = Tin
for b in range(B): src[b, lens[b]:] = 0
mask = (torch.arange(Tin)[None] < lens[:, None]).unsqueeze(1)
src, mask = src.to(dev), mask.to(dev)
ratio = (lens.float() / Tin).to(dev)
def ev():
    return torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for _ in range(2):
        out, _ = model.beam_decode(src, mask, ratio, bench.VOCAB, args)
    torch.cuda.synchronize()
    e = [ev() for _ in range(4)]
    e[0].record(); x, xm = model.src_embed(src, mask); e[1].record()
    enc_h = model.encoder(x, xm); e[2].record()
    out, _ = model.decode_from_encoder(enc_h, xm, ratio, bench.VOCAB, args); e[3].record()
    torch.cuda.synchronize()
    print("front-end ms", e[0].elapsed_time(e[1]), "encoder ms", e[1].elapsed_time(e[2]), "post-encoder ms", e[2].elapsed_time(e[3]))
    _lib.PROFILE_SHAPES = True
    _lib.start_profile(); enc_h = model.encoder(x, xm); prof = _lib.stop_profile()
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]: print(f"  {k:40s} {v['calls']:3d} {v['ms']:.3f}")
    t0 = time.perf_counter()
    for _ in range(5): out, _ = model.beam_decode(src, mask, ratio, bench.VOCAB, args)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("beam_decode from features: ms", dt * 1e3, "utt/s", B / dt, "U mean", sum(len(o[0]['hyp']) - 1 for o in out) / B)
with torch.no_grad():
    _lib.PROFILE_SHAPES = True
    _lib.start_profile(); x, xm = model.src_embed(src, mask); prof = _lib.stop_profile()
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]: print(f"  {k:40s} {v['calls']:3d} {v['ms']:.3f}")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        x, xm = model.src_embed(src, mask); torch.cuda.synchronize()
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=70))
