"""Where the non-kernel time of a decode step goes: wall clock of the three host phases with and without a device
synchronisation between them (scratch/diagnostic; run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cassnat_asr_public_b200 import ops
from cassnat_asr_public_b200.cassnat import make_model

dev = torch.device("cuda:0")
w = bench.WORKLOADS["cfg3"]
args = bench.model_args(w)
torch.manual_seed(1234)
model = make_model(80, args).to(dev).eval()
model.set_precision("bf16")
model.use_cuda_graphs = True
enc_h, mask, ratio, _ = bench.synth_encoder_output(w, 64, 1234)
inp = [t.to(dev) for t in (enc_h, mask, ratio)]
with torch.no_grad():
    for _ in range(5):
        model.decode_from_encoder(*inp, bench.VOCAB, args)
    torch.cuda.synchronize()
    N = 30
    t0 = time.perf_counter()
    for _ in range(N):
        model.decode_from_encoder(*inp, bench.VOCAB, args)
    torch.cuda.synchronize()
    print("whole call, back to back: %.3f ms" % ((time.perf_counter() - t0) / N * 1e3))
    # phase A: cast + K1 + collapse + .item()
    B, T, d = inp[0].shape
    mask2 = inp[1].reshape(B, T)
    ta = tb = tc = 0.0
    for _ in range(N):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = ops.cast_bf16(inp[0], split=True)
        st = model.ctc_generator.greedy(e)
        t1 = time.perf_counter()
        aligned, ylen, ymax = model._collapse(st, mask2, B, T, 1, 0, 0.9)
        t2 = time.perf_counter()
        frames = (inp[2] * T).long()
        ids, logp, ylen2, ymax2 = model._stage_b_graphed(e, inp[1], frames, aligned, ylen, ymax, args, 0)
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        ta += t1 - t0; tb += t2 - t1; tc += t3 - t2
        td = t4 - t3
    print("host: K1 launches %.3f ms | collapse+item (waits for K1) %.3f ms | graph stage launch %.3f ms | graph drain %.3f ms"
          % (ta / N * 1e3, tb / N * 1e3, tc / N * 1e3, td * 1e3))
    # phase C: hypothesis assembly alone (device idle)
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(N):
        model.decode_from_encoder(*inp, bench.VOCAB, args)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
