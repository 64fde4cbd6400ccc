"""torchrun scratch: where does the N>1 step time go (decode vs hypothesis gather)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from cassnat_asr_public_b200 import sharding
from cassnat_asr_public_b200.cassnat import make_model

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
w = bench.WORKLOADS["cfg3"]
args = bench.model_args(w, precision="bf16")
torch.manual_seed(1234)
model = make_model(80, args).to(dev).eval()
model.set_precision("bf16")
enc_h, mask, ratio, lens = bench.synth_encoder_output(w, 64, 1234 + rank)
dev_in = [t.to(dev) for t in (enc_h, mask, ratio)]
T = enc_h.size(1)

def run(gather, n=30):
    ts = [0.0, 0.0]
    for i in range(n + 3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            out, _ = model.decode_from_encoder(*dev_in, bench.VOCAB, args)
        hyps = [o[0]["hyp"] for o in out]
        t1 = time.perf_counter()
        if gather and world > 1:
            hyps = sharding.gather_hypotheses(hyps, 64 * world, T + 2, dev)
        t2 = time.perf_counter()
        if i >= 3:
            ts[0] += t1 - t0; ts[1] += t2 - t1
    return [x / n * 1e3 for x in ts]

for g in (False, True):
    r = run(g)
    print(f"rank {rank} gather={g}: decode {r[0]:.2f} ms  gather {r[1]:.2f} ms  OMP={os.environ.get('OMP_NUM_THREADS')} threads={torch.get_num_threads()}", flush=True)
if world > 1:
    dist.destroy_process_group()
