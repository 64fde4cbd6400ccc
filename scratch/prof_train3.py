import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(__file__), "prof_train2.py")).read().split("from torch.profiler")[0])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as p:
    one(); torch.cuda.synchronize()
want = ("aten::copy_", "aten::add_", "aten::mul", "aten::zeros", "aten::fill_", "aten::cat", "aten::add", "aten::zero_", "aten::sum", "aten::to")
acc = collections.defaultdict(lambda: [0, 0.0])
for e in p.events():
    if e.name in want and e.device_time_total > 0:
        st = [s for s in (e.stack or []) if "cassnat_asr_public_b200" in s or "bench.py" in s]
        key = (e.name, " <- ".join(s.split("/")[-1] for s in st[:3]))
        acc[key][0] += 1
        acc[key][1] += e.device_time_total
rows = sorted(acc.items(), key=lambda kv: -kv[1][1])
for (name, st), (n, t) in rows[:40]:
    print(f"{t:8.1f} us x{n:3d} {name:14s} {st}")
