"""torchrun --nproc-per-node N scratch/train_reserve.py : cfg5 training step at several GEMM CTA caps (cn_set_sm_limit)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
import bench
from cassnat_asr_public_b200 import _lib
from cassnat_asr_public_b200.cassnat import make_model
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
w = bench.WORKLOADS['cfg3']; args = bench.model_args(w, precision='bf16')
torch.manual_seed(1234)
model = make_model(80, args).to(dev).eval(); model.set_precision('bf16')
def sync_all():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for lim in [int(x) for x in (sys.argv[1:] or ["0", "140", "132"])]:
    _lib.load().cn_set_sm_limit(lim)
    r = bench.training_step(model, args, w, dev, rank, world, sync_all)
    if rank == 0:
        print(f"limit {lim}: {r['ms_per_step']:.3f} ms/step with dropout, {r['no_dropout']['ms_per_step']:.3f} without", flush=True)
_lib.load().cn_set_sm_limit(0)
if world > 1: dist.destroy_process_group()
