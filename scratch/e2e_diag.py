"""Why is the end-to-end loop slower than the device-resident loop?  Wall clock per step of the pipelined decode loop with
(1) inputs resident, (2) pinned H2D into freshly allocated tensors on a side stream (bench.py's e2e), (3) H2D into two
preallocated device buffers."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cassnat_asr_public_b200.cassnat import make_model

dev = torch.device("cuda:0")
w = bench.WORKLOADS["cfg3"]
args = bench.model_args(w, precision="bf16")
torch.manual_seed(1234)
model = make_model(80, args).to(dev).eval()
model.set_precision("bf16")
model.use_cuda_graphs = True
enc_h, mask, ratio, _ = bench.synth_encoder_output(w, 64, 1234)
host = [t.pin_memory() for t in (enc_h, mask, ratio)]
res = [t.to(dev) for t in host]
N = 40
n_valid = mask.reshape(64, -1).sum(1).tolist()
copy_stream = torch.cuda.Stream(device=dev)


def run(kind):
    bufs2 = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]

    def stage(i):
        with torch.cuda.stream(copy_stream):
            if kind == "fresh":
                b = [t.to(dev, non_blocking=True) for t in host]
            elif kind == "ragged":
                b = bufs2[i % 2]
                for u in range(64):
                    b[0][u, :n_valid[u]].copy_(host[0][u, :n_valid[u]], non_blocking=True)
                b[1].copy_(host[1], non_blocking=True); b[2].copy_(host[2], non_blocking=True)
            else:
                b = bufs2[i % 2]
                for d, s in zip(b, host):
                    d.copy_(s, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        return b, ev
    with torch.no_grad():
        for _ in range(3):
            model.decode_from_encoder(*res, bench.VOCAB, args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pend = None
        nxt = stage(0) if kind != "resident" else None
        for i in range(N):
            if kind == "resident":
                inp = res
            else:
                inp, ev = nxt
                torch.cuda.current_stream().wait_event(ev)
                if kind == "fresh":
                    for t in inp:
                        t.record_stream(torch.cuda.current_stream())
                if i + 1 < N:
                    if kind in ("prealloc", "ragged"):          # the buffer about to be overwritten was read by step i-1's stage A: done
                        copy_stream.wait_stream(torch.cuda.current_stream())
                    nxt = stage(i + 1)
            q, _ = model.decode_from_encoder(*inp, bench.VOCAB, args, defer=True)
            if pend is not None:
                pend.result()
            pend = q
        pend.result()
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / N * 1e3
        t0 = time.perf_counter()
        if kind != "resident":
            for i in range(10):
                stage(i)
        t_host = (time.perf_counter() - t0) / 10 * 1e3
        torch.cuda.synchronize()
        return t_all, t_host


for kind in ("resident", "fresh", "prealloc", "ragged", "prealloc", "ragged"):
    print(kind, "%.3f ms/step, stage() host time %.3f ms" % run(kind), flush=True)
