#!/usr/bin/env python
"""Benchmark of the CASS-NAT alignment-conditioned decode path (BASELINE.json metric:
utterances/sec of CASS-NAT decode; alignment-lattice GB/s for the cfg2 microbench).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg1|cfg2]

A "step" is one pass of the path (everything CassNAT.beam_decode does after the encoder output
`enc_h` exists: CTC head -> best-path alignment -> segment ids -> extractor -> SAD -> MAD -> output
head -> hypotheses) over one batch of synthetic encoder output of the named shape, random-init
weights.  N > 1: one process per GPU under torchrun, utterance-sharded, weak scaling (the batch per
GPU is fixed); no data-path collective, one end-of-job count all_reduce.  One JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOCAB = SimpleNamespace(word2index={"sos": 1, "eos": 2, "blank": 0})

WORKLOADS = {
    # BASELINE.json configs[2]: LibriSpeech-shaped base model, 64 x 20 s per GPU
    "cfg3": dict(d_model=512, n_head=8, d_ff=2048, N_enc=12, N_mix_dec=6, batch=64, frames_in=2000, audio_s=20.0,
                 name="cfg3: CASS-NAT base (512-d, 12 enc / 1+1+6 dec, V=5000) greedy decode, 64 x 20 s per GPU, "
                      "post-encoder path on synthetic enc_h"),
    # BASELINE.json configs[0]: small model, 16 x 10 s
    "cfg1": dict(d_model=256, n_head=4, d_ff=2048, N_enc=12, N_mix_dec=3, batch=16, frames_in=1000, audio_s=10.0,
                 name="cfg1: CASS-NAT small (256-d, 12 enc / 1+1+3 dec, V=5000) greedy decode, 16 x 10 s"),
}


def model_args(w, **over):
    a = dict(use_conv_enc=False, use_conv_dec=False, model_type="transformer", pos_type="absolute",
             n_head=w["n_head"], d_model=w["d_model"], d_encff=w["d_ff"], d_decff=w["d_ff"], d_ff=w["d_ff"],
             dropout=0.0, N_enc=w["N_enc"], N_extra=1, N_self_dec=1, N_mix_dec=w["N_mix_dec"], vocab_size=5000,
             interctc_alpha=0, interce_alpha=0, interctc_layer=0, interce_layer=0, padding_idx=0, label_smooth=0.1,
             share_ff=False, enc_max_relative_len=20, enc_kernel_size=31, dec_max_relative_len=8, dec_kernel_size=3,
             use_trigger=True, use_best_path=False, sample_topk=0, ctc_alpha=0.5, att_alpha=1.0, save_embedding=False,
             use_unimask=False, src_trigger=False, left_trigger=0, right_trigger=0, test_hitrate=False, sample_num=0,
             decode_type="att_only", threshold=0.9, rank_model="lm", lm_weight=0, beam_width=1, length_penalty=0,
             use_gpu=True)
    a.update(over)
    return SimpleNamespace(**a)


def synth_encoder_output(w, batch, seed):
    """Synthetic encoder output of the named shape: unit-variance rows (what the encoder's final
    LayerNorm emits at init), ragged lengths U{60%..100%} with utterance 0 full length."""
    g = torch.Generator().manual_seed(seed)
    T = (w["frames_in"] + 3) // 4
    enc_h = torch.randn(batch, T, w["d_model"], generator=g)
    lens = torch.randint(int(0.6 * T), T + 1, (batch,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None, :] < lens[:, None]).unsqueeze(1)
    # exact ratios (cassnat.py:436 truncates): nudge so that (ratio*T).long() == lens
    ratio = (lens.float() + 0.5) / T
    return enc_h, mask, ratio, lens


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tensor=j.get("bf16_tflops_sustained", j["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tensor=1400.0, source="fallback")


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        self.tmp.flush()
        rows = [r.strip().split(",") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ reference arm
def run_reference(a, w):
    """The reference's CPU implementation of the path = the oracle port (the reference is Python and
    does not travel to the GPU box), all host threads, on a bounded sample of the same workload."""
    import oracle.nat_torch as O
    from cassnat_asr_public_b200.cassnat import make_model
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    args = model_args(w, use_gpu=False)
    torch.manual_seed(1234)
    model = make_model(80, args)                    # parameter container only (never runs on this arm)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    # utterances per step: the workload's full per-GPU batch when the whole run then stays within ~2.5 minutes of CPU
    # work (about 14 utt/s on 16 cores), else a bounded sample of it
    full = a.batch or w["batch"]
    sample = a.ref_sample or max(8, min(full, int(150 * 14 / (a.steps + max(a.warmup, 1))) // 8 * 8))
    enc_h, mask, ratio, lens = synth_encoder_output(w, sample, 1234)
    frames = (ratio * enc_h.size(1)).long()
    cfg = SimpleNamespace(**vars(args))

    def step():
        with torch.no_grad():
            return O.decode_from_encoder(enc_h, mask, frames, sd, cfg)

    for _ in range(max(a.warmup, 1)):
        out = step()
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        out = step()
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    value = sample * len(ts) / total
    T = enc_h.size(1)
    line = {"impl": "reference", "metric": "utterances/sec", "value": value, "unit": "utt/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": max(a.warmup, 1), "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same workload as the other arm; each step decodes a bounded sample of it (cpu_baseline.sample)
            "config": {"workload": w["name"], "global_batch": (a.batch or w["batch"]) * a.gpus, "per_gpu_batch": a.batch or w["batch"],
                       "frames_T": T, "tokens_U_max": int(out["ymax"]) - 1, "precision": "fp32 (reference arithmetic, host cores)",
                       "parallelism": "host threads", "sample_per_step": sample},
            "rtf": (total / len(ts)) / (sample * w["audio_s"]),
            "cpu_baseline": {"value": value, "unit": "utt/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{sample} utterances of the workload per step (oracle/nat_torch.py on CPU)"},
            "e2e": {"value": value, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm
def run_ours(a, w):
    import torch.distributed as dist

    from cassnat_asr_public_b200 import _lib, sharding
    from cassnat_asr_public_b200.cassnat import make_model
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if _lib.load().cn_device_check() != 0:
        raise RuntimeError("bench.py needs an sm_100 device (no fallback)")
    B = a.batch or w["batch"]
    args = model_args(w, precision=a.precision)
    torch.manual_seed(1234)
    model = make_model(80, args).to(dev).eval()
    model.set_precision(a.precision)
    model.use_cuda_graphs = not a.no_graphs
    if a.scaling == "strong":
        # the workload's batch is fixed; every GPU takes a contiguous block of B/N utterances
        assert B % world == 0, "strong scaling needs the batch to divide by the number of GPUs"
        g_enc, g_mask, g_ratio, _ = synth_encoder_output(w, B, 1234)
        B = B // world
        cut = slice(rank * B, (rank + 1) * B)
        enc_h, mask, ratio = g_enc[cut].contiguous(), g_mask[cut].contiguous(), g_ratio[cut].contiguous()
    else:
        enc_h, mask, ratio, lens = synth_encoder_output(w, B, 1234 + rank)        # a different shard per rank
    T = enc_h.size(1)
    host = [t.pin_memory() for t in (enc_h, mask, ratio)]
    dev_in = [t.to(dev) for t in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(inputs):
        # Each rank decodes its own shard and keeps its own hypotheses, exactly like the reference's
        # decode job (one process per device writing its own result file, run_hubert.sh:94-116):
        # no data-path collective.  The job-level exchange is the (errors, tokens) sum below.
        with torch.no_grad():
            out, _ = model.decode_from_encoder(inputs[0], inputs[1], inputs[2], VOCAB, args)
        return [o[0]["hyp"] for o in out]

    def step_async(inputs):
        # serving-loop form of the same call: the batch's device work is queued and a handle comes back; the host-side
        # hypothesis assembly (collect) is done after the NEXT batch has been queued, so it overlaps that batch
        with torch.no_grad():
            pend, _ = model.decode_from_encoder(inputs[0], inputs[1], inputs[2], VOCAB, args, defer=True)
        return pend

    def collect(pend):
        return [o[0]["hyp"] for o in pend.result()]

    def job_totals(hyps, n_steps):
        # end-of-job count exchange (the WER numerator/denominator all_reduce, cassnat_task.py:186-189)
        return sharding.reduce_error_counts(0, n_steps * sum(len(h) - 1 for h in hyps), dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None     # samples during warm-up + timed steps (all under load)
    for _ in range(a.warmup):
        hyps = step(dev_in)
    # ---- device-resident throughput: per-step CUDA events, L2 flushed between steps
    sync_all()
    _lib.launch_count = 0
    evs = []
    pend = None
    for _ in range(a.steps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nxt = step_async(dev_in)
        if pend is not None:
            hyps = collect(pend)
        pend = nxt
        e1.record()
        evs.append((e0, e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hyps = collect(pend)                     # the last batch's hypotheses: inside the clock as well
    job_totals(hyps, a.steps)
    e1.record()
    evs.append((e0, e1))
    sync_all()
    launches = _lib.launch_count
    clocks = sampler.stop() if sampler else None
    # the synchronous form of the call (hypotheses returned before the next batch is queued), for the record
    sync_evs = []
    for _ in range(min(a.steps, 10)):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(dev_in)
        e1.record()
        sync_evs.append((e0, e1))
    sync_all()
    sync_call_ms = sum(e0.elapsed_time(e1) for e0, e1 in sync_evs) / len(sync_evs)
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    # ---- end to end: pinned host buffers in, hypotheses (python lists) out, copies inside the clock
    # (double-buffered: the copy of step i+1 runs on a side stream while step i computes)
    # two preallocated device batches (no allocation inside the loop).  Copying only the frames an utterance has (one copy
    # per utterance, 19 % fewer bytes) was measured and is slower on one GPU: 5.90 vs 5.48-5.54 ms/step
    # (scratch/e2e_diag.py) -- 64 small copies per step cost more than the padding they skip.
    copy_stream = torch.cuda.Stream(device=dev)
    stage_bufs = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
    reader_done = [None, None]               # the step that last read a staging batch has been queued up to this event
    h2d = sum(t.numel() * t.element_size() for t in host)

    def stage(k):
        bufs = stage_bufs[k % 2]
        with torch.cuda.stream(copy_stream):
            if reader_done[k % 2] is not None:
                copy_stream.wait_event(reader_done[k % 2])
            for d_, s_ in zip(bufs, host):
                d_.copy_(s_, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs, ev

    sync_all()
    t0 = time.perf_counter()
    for j in range(4):                       # (copy alone, for the record: what the host link of this box delivers)
        stage(j)[1].synchronize()
    h2d_alone_s = (time.perf_counter() - t0) / 4
    sync_all()
    t0 = time.perf_counter()
    nxt = stage(0)
    pend = None
    for i in range(a.steps):
        staged, ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        if i + 1 < a.steps:
            nxt = stage(i + 1)
        queued = step_async(staged)
        reader_done[i % 2] = torch.cuda.Event()
        reader_done[i % 2].record()
        if pend is not None:
            hyps = collect(pend)
        pend = queued
    hyps = collect(pend)
    job_totals(hyps, a.steps)
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allt = torch.empty(world, 2, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, t.unsqueeze(0))
        per_rank = [round(v / a.steps, 4) for v in allt[:, 0].tolist()]      # device ms/step of every rank (diagnostic)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = t.tolist()
    strong = None
    if world > 1 and a.scaling == "weak" and w["batch"] % world == 0:
        strong = strong_scaling_leg(model, args, w, dev, rank, world, sync_all, flush, a.steps)
    training = None
    if not a.no_training:
        training = training_step(model, args, w, dev, rank, world, sync_all, graphed=not a.no_graphs)
    # ---- per-kernel durations for the roofline (separate profiled steps, CUDA events per C-ABI call)
    prof = None
    if rank == 0:
        model.use_cuda_graphs = False           # per-call CUDA events need eager launches
        _lib.start_profile()
        for _ in range(2):
            step(dev_in)                         # rank 0 only: no collective in here
        prof = _lib.stop_profile()
        _lib.PROFILE_SHAPES = True
        _lib.start_profile()
        step(dev_in)
        prof_shapes = _lib.stop_profile()
        _lib.PROFILE_SHAPES = False
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_tok = sum(len(h) - 1 for h in hyps[:B]) / B
    units = B * world * a.steps
    value = units / (total_ms / 1e3)
    pk = peaks()
    top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    name, d = top
    if d["flops"] > 0:
        ach = d["flops"] / (d["ms"] / 1e3) / 1e12
        roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor"], "traffic": None}
    else:
        ach = d["bytes"] / (d["ms"] / 1e3) / 1e9
        roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": None}
    roof["peak_source"] = pk["source"]
    roof["avg_launch_ms"] = d["ms"] / d["calls"]
    roof["share_of_step"] = d["ms"] / sum(v["ms"] for v in prof.values())
    # the whole step: algorithmic flops over the step time, and the time-weighted roofline fraction of its kernels
    # (each kernel against whichever of the two bounds it is closer to)
    k_ms = sum(v["ms"] for v in prof.values()) / 2              # (the profile covers two eager steps)
    step_ms = total_ms / a.steps
    fr = lambda v: max(v["flops"] / (v["ms"] / 1e3) / 1e12 / pk["tensor"] if v["flops"] else 0.0,
                       v["bytes"] / (v["ms"] / 1e3) / 1e9 / pk["hbm"] if v["bytes"] else 0.0)
    roof["whole_step"] = {"algorithmic_tflop": sum(v["flops"] for v in prof.values()) / 2 / 1e12,
                          "frac_tensor": sum(v["flops"] for v in prof.values()) / 2 / (step_ms / 1e3) / 1e12 / pk["tensor"],
                          "kernel_time_weighted_frac": sum(v["ms"] * fr(v) for v in prof.values()) / 2 / k_ms,
                          "kernel_ms": k_ms, "step_ms": step_ms}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and a.workload == "cfg3" and a.precision == "bf16" and B == w["batch"]:
        roof["traffic"] = json.load(open(tpath)).get(name)      # ncu capture of exactly this workload (profiles/)
    ymax = max(len(h) - 1 for h in hyps[:B])
    d2h = B * ymax * (4 + 8) + B * 8 + 4          # ids i32 + logp f64 + ylen i64 + the ymax scalar
    line = {"metric": "utterances/sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f32" if a.precision == "fp32" else a.precision, "data": "synthetic",
            "config": {"workload": w["name"], "global_batch": B * world, "per_gpu_batch": B, "frames_T": T,
                       "tokens_U_mean": n_tok, "tokens_U_max": ymax, "precision": a.precision, "cuda_graph": not a.no_graphs,
                       "parallelism": f"utterance-sharded x{world}", "l2": "flushed (256 MiB write) between steps",
                       "call": "decode_from_encoder(defer=True): hypotheses of batch i are assembled on the host "
                               "after batch i+1 has been queued; every batch's hypotheses are collected inside the clock"},
            "synchronous_call": {"ms_per_step": sync_call_ms, "utt_per_s": B * world / (sync_call_ms / 1e3),
                                 "note": "the same call returning the hypotheses before the next batch is queued "
                                         "(rank 0's clock)"},
            "rtf": (total_ms / 1e3 / a.steps) / (B * w["audio_s"]),
            "e2e": {"value": units / e2e_s, "unit": "utt/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / a.steps * 1e3, "h2d_alone_ms": h2d_alone_s * 1e3,
                    "h2d_alone_GBps": h2d / h2d_alone_s / 1e9,
                    # the copy of batch i+1 overlaps batch i: when the host link needs longer for one batch than the device
                    # does, the end-to-end number measures the link of this box, not the path
                    "bound": "host link" if h2d_alone_s * 1e3 > total_ms / a.steps else "device"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            # per kernel: algorithmic flops and operand + result bytes over its CUDA-event time, each against the
            # measured peak -- a kernel is at its roofline when EITHER fraction is near 1 (the d x d projections with
            # their fp32 residual stream are bound by HBM, not by the tensor pipe)
            "kernels": {k: {"ms_per_step": round(v["ms"], 4), "calls_per_step": v["calls"],
                            **({"TFLOPs": round(v["flops"] / v["ms"] / 1e9, 1),
                                "frac_tensor": round(v["flops"] / v["ms"] / 1e9 / pk["tensor"], 3)} if v["flops"] else {}),
                            **({"GBps": round(v["bytes"] / v["ms"] / 1e6, 1),
                                "frac_hbm": round(v["bytes"] / v["ms"] / 1e6 / pk["hbm"], 3)} if v["bytes"] else {})}
                        for k, v in sorted(prof_shapes.items(), key=lambda kv: -kv[1]["ms"])}}
    if per_rank is not None:
        line["ms_per_step_by_rank"] = per_rank
    if strong is not None:
        line["strong_scaling"] = strong
    if training is not None:
        line["training_step"] = training
    def leg(key, fn, *fa, **kw):
        """Auxiliary measurement: a failure is recorded in the line instead of losing the headline numbers."""
        try:
            line[key] = fn(*fa, **kw)
        except Exception as e:                                  # noqa: BLE001
            line[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.synchronize()

    if world == 1 and not a.no_parity and a.precision != "fp32":
        model.use_cuda_graphs = False
        leg("parity", parity_report, model, args, w, dev_in, a.precision)
        if "fp32_parity_mode" in line["parity"]:
            line["fp32_parity_mode"] = line["parity"].pop("fp32_parity_mode")
    if world == 1 and not a.no_parity and a.precision == "bf16":
        leg("bit_exact_mode", bit_exact_mode, model, args, dev_in, flush, a.steps, a.warmup, not a.no_graphs, a.precision)
    if world == 1 and not a.no_eager:
        leg("gpu_eager_baseline", gpu_eager_baseline, model, args, dev_in, value, units / e2e_s)
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, w, args)
    if world == 1 and a.workload == "cfg3" and not a.no_sampled:
        leg("cfg1_decode", small_decode, dev, a.precision, not a.no_graphs)
    if world == 1 and not a.no_sampled and a.workload == "cfg3":
        model.use_cuda_graphs = not a.no_graphs
        leg("speechlike_decode", speechlike_decode, model, args, w, dev_in, B, dev)
        leg("features_decode", features_decode, model, args, w, dev, B)
        leg("conformer_decode", conformer_decode, w, dev_in, B, dev, a.precision, not a.no_graphs)
        leg("features_decode_conformer", features_decode_conformer, w, dev, B, a.precision, not a.no_graphs)
        leg("sampled_decode", sampled_decode, model, w, dev)
        leg("sampled_decode_lm", sampled_decode, model, w, dev, iters=3, real_lm=True)
    if world == 1 and not a.no_lattice:
        del model, dev_in, flush
        torch.cuda.empty_cache()
        leg("alignment_lattice", alignment_lattice, dev, pk)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def strong_scaling_leg(model, args, w, dev, rank, world, sync_all, flush, steps):
    """BASELINE configs[2] read literally: the 64-utterance batch is split into contiguous blocks of 64/N
    utterances, one per GPU.  Device time of `steps` decodes, max over ranks.  Runs on every rank."""
    import torch.distributed as dist
    B = w["batch"] // world
    g_enc, g_mask, g_ratio, _ = synth_encoder_output(w, w["batch"], 1234)
    cut = slice(rank * B, (rank + 1) * B)
    inp = [t[cut].contiguous().to(dev) for t in (g_enc, g_mask, g_ratio)]
    with torch.no_grad():
        for _ in range(3):
            model.decode_from_encoder(inp[0], inp[1], inp[2], VOCAB, args)
        sync_all()
        total = 0.0
        evs = []
        for _ in range(steps):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.decode_from_encoder(inp[0], inp[1], inp[2], VOCAB, args)
            e1.record()
            evs.append((e0, e1))
        sync_all()
    t = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    return {"workload": f"cfg3 batch of {w['batch']} utterances split into blocks of {B} per GPU", "scaling": "strong",
            "ms_per_step": ms, "utt_per_s": w["batch"] / (ms / 1e3)}


def parity_report(model, args, w, dev_in, precision):
    """Agreement of the benchmarked tensor-core mode with the fp32 parity mode on the full-size workload.
    The fp32 parity mode (FFMA kernels) is the mode the golden tests pin to the reference (identical hypotheses on
    every fixture); a few utterances are also decoded by the oracle port here to anchor it at this size.
    Margins are top-1 minus top-2 of the fp32 logits (torch fp32 matmul on the same tensors, TF32 off)."""
    import numpy as np

    import oracle.nat_torch as O
    from cassnat_asr_public_b200 import ops
    enc_h, mask, ratio = dev_in
    B, T, _ = enc_h.shape
    valid = mask.reshape(B, T)
    torch.backends.cuda.matmul.allow_tf32 = False

    def run(mode, safe=True):
        ops.SAFE_HEADS = safe
        model.set_precision(mode)
        with torch.no_grad():
            out, _, keep = model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args, return_intermediates=True)
        return [o[0]["hyp"] for o in out], keep

    try:
        ref_h, ref = run("fp32")
        # timing of the parity mode (the mode with proven parity), same inputs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            for _ in range(3):
                model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args)
        e1.record()
        torch.cuda.synchronize()
        fp32_ms = e0.elapsed_time(e1) / 3
        with torch.no_grad():
            z1 = torch.nn.functional.linear(enc_h, model.ctc_generator.proj.weight, model.ctc_generator.proj.bias)
            m1 = z1.topk(2, -1)[0]
            m1 = (m1[..., 0] - m1[..., 1])[valid]
            del z1
            z9 = torch.nn.functional.linear(ref["dec_h"], model.att_generator.proj.weight, model.att_generator.proj.bias)
            t9 = z9.topk(2, -1)[0]
            m9 = t9[..., 0] - t9[..., 1]
            del z9, t9
        U = ref["ids"].shape[1]
        ylen = ref["ylen"]
        read = torch.arange(U, device=enc_h.device)[None, :] <= ylen[:, None]        # rows the pick loop reads
        n_oracle = 4
        frames = (ratio * T).long()
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        with torch.no_grad():
            want = O.decode_from_encoder(enc_h[:n_oracle].cpu(), mask[:n_oracle].cpu(), frames[:n_oracle].cpu(), sd,
                                         SimpleNamespace(**{**vars(args), "use_gpu": False}))
        # the oracle pads its sub-batch to its own ymax: compare the hypotheses (what the caller gets), token by token
        # with the oracle's own top-2 margin (fp32 summation order differs between FFMA tiles and the CPU BLAS, so
        # tokens whose margin is inside fp32 round-off are undecidable)
        omarg = want["att_out"].topk(2, -1)[0]
        omarg = (omarg[..., 0] - omarg[..., 1])
        n_tok = n_diff = n_diff_clear = 0
        same_len = True
        for b in range(n_oracle):
            hw, hg = want["hyps"][b], ref_h[b]
            same_len &= abs(len(hw) - len(hg)) <= 1     # min(ylen + 1, ymax): ymax is the sub-batch's own maximum
            for i, (x, y) in enumerate(zip(hw[1:], hg[1:])):
                n_tok += 1
                if x != y:
                    n_diff += 1
                    n_diff_clear += int(float(omarg[b, i]) > 1e-4)
        frac = lambda num, den: {"equal": int(num), "of": int(den), "rate": (float(num) / float(den)) if int(den) else None}
        rep = {"reference_mode": "this package's fp32 parity mode on the same inputs (cfg3 full size)",
               "fp32_mode_vs_oracle_port": {"utterances": n_oracle, "same_lengths": bool(same_len), "tokens": n_tok,
                                            "tokens_differing": n_diff, "differing_with_oracle_margin_gt_1e-4": n_diff_clear,
                                            "alignment_identical": bool((ref["aligned"][:n_oracle].cpu().numpy() == want["aligned"]).all())},
               "frames": int(valid.sum()), "token_rows_read": int(read.sum()), "utterances": B}
        variants = [(precision, True, precision + " (decision-safe heads: K1/K9 as bf16x3)")]
        if precision == "bf16":
            variants += [("bf16", False, "bf16 with plain bf16 heads"),
                         ("bf16x3", True, "bf16x3 (all GEMMs and the attention three-pass on tcgen05)")]
        for mode, safe, label in variants:
            hyp, got = run(mode, safe)
            same_al = got["aligned"] == ref["aligned"]
            utt_al = same_al.all(1)
            r = {"alignment_frames": frac(same_al[valid].sum(), valid.sum()), "alignments_identical": f"{int(utt_al.sum())}/{B}",
                 "alignment_frames_where_fp32_margin_gt_1e-4": frac(same_al[valid][m1 > 1e-4].sum(), (m1 > 1e-4).sum())}
            # tokens: only utterances whose alignment (hence token grid) is identical are comparable row by row
            if got["ids"].shape == ref["ids"].shape:
                rows = read & utt_al[:, None]
                eq = (got["ids"] == ref["ids"])
                r["tokens"] = frac(eq[rows].sum(), rows.sum())
                for thr in (1e-4, 1e-3, 1e-2, 5e-2):
                    sel = rows & (m9 > thr)
                    r[f"tokens_where_fp32_margin_gt_{thr:g}"] = frac(eq[sel].sum(), sel.sum())
                r["fraction_of_rows_with_margin_lt_1e-2"] = float((m9[rows] < 1e-2).double().mean()) if int(rows.sum()) else None
                for key in ("ac_embed", "pred_embed", "dec_h"):
                    x, y = got[key].float()[rows], ref[key].float()[rows]
                    r[f"rel_err_{key}"] = float((x - y).norm() / y.norm())
            r["hypotheses_identical"] = f"{sum(int(a == b) for a, b in zip(hyp, ref_h))}/{B}"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.no_grad():
                for _ in range(3):
                    model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args)
            e1.record()
            torch.cuda.synchronize()
            r["ms_per_step_eager"] = e0.elapsed_time(e1) / 3
            r["utt_per_s_eager"] = B / (r["ms_per_step_eager"] / 1e3)
            rep[label] = r
        rep["fp32_parity_mode"] = {"ms_per_step": fp32_ms, "utt_per_s": B / (fp32_ms / 1e3),
                                   "note": "FFMA GEMMs + fp32 attention; hypotheses identical to the reference on every golden fixture"}
        return rep
    finally:
        ops.SAFE_HEADS = True
        model.set_precision(precision)


def bit_exact_mode(model, args, dev_in, flush, steps, warmup, graphs, restore):
    """The same workload in the bf16x3 mode (every GEMM and the attention as three tcgen05 passes on split-bf16
    operands): the mode whose alignments AND hypotheses equal the fp32 parity mode's, timed exactly like the headline
    (device events per step, L2 flushed between steps, CUDA-graph replay)."""
    B = dev_in[0].size(0)

    def run():
        with torch.no_grad():
            out, _ = model.decode_from_encoder(dev_in[0], dev_in[1], dev_in[2], VOCAB, args)
        return [o[0]["hyp"] for o in out]
    try:
        model.use_cuda_graphs = False
        model.set_precision("fp32")
        want = run()
        model.set_precision("bf16x3")
        model.use_cuda_graphs = graphs
        for _ in range(max(warmup, 3)):
            got = run()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(steps):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            got = run()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= steps
        return {"precision": "bf16x3", "value": B / (ms / 1e3), "unit": "utt/s", "ms_per_step": ms, "cuda_graph": graphs,
                "hypotheses_identical_to_fp32_mode": f"{sum(int(x == y) for x, y in zip(got, want))}/{B}",
                "note": "three tensor-core passes per GEMM / attention product: 3x the headline's flops, fp32-grade results"}
    finally:
        model.set_precision(restore)
        model.use_cuda_graphs = graphs


def gpu_eager_baseline(model, args, dev_in, ours_value, ours_e2e, iters=5):
    """The "library kernels" comparator (SURVEY §8d / BASELINE.md §3): the same path written with stock PyTorch ops on
    the same GPU -- cuBLAS GEMMs, ATen softmax / masked_fill / LayerNorm arithmetic, materialised posteriors and
    trigger masks exactly like src/models/cassnat.py:433-501 -- in fp32 (TF32 off) and under bf16 autocast.  The block
    arithmetic is the oracle port run on CUDA tensors; the hypothesis pick uses one device->host copy instead of the
    reference's B x U `.item()` loop (cassnat.py:574-637), which favours the baseline."""
    import oracle.nat_torch as O
    enc_h, mask, ratio = dev_in
    B, T, _ = enc_h.shape
    dev = enc_h.device
    frames = (ratio * T).long()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = SimpleNamespace(**vars(args))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def decode():
        ctc_out = O.generator(enc_h, sd, "ctc_generator").float()
        best = ctc_out.argmax(-1).masked_fill(mask.squeeze(1) == 0, 0)                 # cassnat.py:378-387
        shift = best.new_zeros(best.size())
        shift[:, 1:] = best[:, :-1]
        best = best.masked_fill(best == shift, 0)
        shift[:, 1:] = best[:, :-1]
        ylen = (shift != 0).sum(1)
        ymax = int(ylen.max().item())
        trig = (shift != 0).cumsum(1).unsqueeze(1).repeat(1, ymax + 1, 1)              # :358-361
        trig = trig == torch.arange(ymax + 1, device=dev).type_as(trig)[None, :, None]
        trig.masked_fill_(mask == 0, 0)
        trig[torch.arange(B, device=dev), ylen, frames - 1] = 1
        ylen, ymax = ylen + 1, ymax + 1
        trig = trig & mask
        tgt = (torch.arange(ymax, device=dev)[None, :] < ylen[:, None]).unsqueeze(1)
        att_out = O.decoder_stack(enc_h, trig, mask, tgt, ymax, sd, cfg)[0].float()
        ids = att_out.argmax(-1).cpu()
        yl = ylen.cpu().tolist()
        return [[1] + ids[b, :min(yl[b] + 1, ymax)].tolist() for b in range(B)]

    out = {}
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            for _ in range(2):
                hyps = decode()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                hyps = decode()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / iters
        out[name] = {"ms_per_step": dt * 1e3, "utt_per_s": B / dt, "ours_over_eager": ours_value / (B / dt),
                     "ours_e2e_over_eager": ours_e2e / (B / dt)}
        torch.cuda.empty_cache()
    out["note"] = ("stock PyTorch eager on the same B200, same weights and inputs resident in HBM, wall clock incl. the "
                   "hypothesis read-back; TF32 off for fp32; reported comparator, not a target")
    return out


def small_decode(dev, precision, graphs, iters=50):
    """BASELINE configs[0] (cfg1): the small model (256-d, 4 heads, 1+1+3 decoder layers), 16 x 10 s -> T=250.
    4 000 rows per GEMM cannot fill 148 SMs with 128-row tiles: this is the launch-latency-bound end of the path."""
    from cassnat_asr_public_b200.cassnat import make_model
    w = WORKLOADS["cfg1"]
    args = model_args(w)
    torch.manual_seed(1234)
    model = make_model(80, args).to(dev).eval()
    model.set_precision(precision)
    model.use_cuda_graphs = graphs
    enc_h, mask, ratio, _ = synth_encoder_output(w, w["batch"], 1234)
    inp = [t.to(dev) for t in (enc_h, mask, ratio)]
    with torch.no_grad():
        for _ in range(5):
            out, _ = model.decode_from_encoder(*inp, VOCAB, args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out, _ = model.decode_from_encoder(*inp, VOCAB, args)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    toks = [len(o[0]["hyp"]) - 1 for o in out]
    return {"workload": w["name"] + ", post-encoder path", "ms_per_step": ms, "utt_per_s": w["batch"] / (ms / 1e3),
            "rtf": (ms / 1e3) / (w["batch"] * w["audio_s"]), "tokens_U_mean": sum(toks) / len(toks), "tokens_U_max": max(toks),
            "note": "no L2 flush between steps (the working set is smaller than L2 anyway)"}


def training_step(model, args, w, dev, rank, world, sync_all, graphed=True, B=32, frames_in=1000, L=40, iters=10, dropout=0.1):
    """SURVEY §8 cfg5 on the post-encoder path: forward (CTC head, Viterbi forced alignment, trigger segments,
    extractor -> SAD -> MAD -> head, CTC + label-smoothing loss) and loss.backward() on the package's backward kernels
    (gradient returned at enc_h for the encoder), with the reference's training dropout (conf/cassnat_train.yaml:11:
    0.1; stateless masks, the step's seed lives on the device and is advanced inside the graph).  N > 1: the gradients
    are views of one flat buffer whose bf16 buckets are all_reduced from post-accumulate hooks, i.e. overlapped with
    the rest of backward; zero + forward + backward + exchange are captured as ONE CUDA graph.  Runs on every rank;
    time = max over ranks of the CUDA-event time of `iters` steps.  Also reports the dropout-free step."""
    import torch.distributed as dist

    from cassnat_asr_public_b200 import ops, sharding
    T = (frames_in + 3) // 4
    g = torch.Generator().manual_seed(4321 + rank)
    enc_h = torch.randn(B, T, w["d_model"], generator=g).to(dev).requires_grad_(True)
    lens = torch.randint(int(0.8 * T), T + 1, (B,), generator=g)
    lens[0] = T
    x_mask = (torch.arange(T)[None, :] < lens[:, None]).unsqueeze(1).to(dev)
    ratio = ((lens.float() + 0.5) / T).to(dev)
    labels = torch.randint(4, 5000, (B, L + 1), generator=g)
    labels[:, L] = 2
    labels, sizes = labels.to(dev), torch.full((B,), L, dtype=torch.int64, device=dev)
    model.train()
    params = [p for n, p in model.named_parameters() if not n.startswith(("src_embed.", "encoder."))]
    n_par = sum(p.numel() for p in params)
    seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)

    def set_dropout(p):
        for m in model.modules():
            if hasattr(m, "p_drop"):
                m.p_drop = p

    def measure(p_drop):
        set_dropout(p_drop)
        ops.DROPOUT_SEED_DEV = seed_dev
        flat = sharding.OverlappedFlatGradients(params, comm_dtype=torch.bfloat16)

        def fwd_bwd():
            seed_dev.add_(1)                               # a fresh mask every step, also under graph replay
            flat.zero()
            out = model.train_forward_from_encoder(enc_h, x_mask, ratio, labels, sizes, args)
            out[2].backward()
            flat.finish()                                  # waits for the bucket all_reduces launched during backward
            return out[2]

        # warm-up on a side stream (PyTorch's whole-step capture recipe: the gradient accumulators of the leaves
        # must not be tied to the legacy default stream when backward is captured)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                enc_h.grad = None
                loss = fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        graph = None
        if graphed:
            # the training forward has no host synchronisation (ymax comes from the label tensor's shape), so the
            # whole step is captured once and replayed: gradients live in the flat buffer (static memory)
            enc_h.grad = None
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = fwd_bwd()

        def one():
            if graph is not None:
                graph.replay()
                return static_loss
            enc_h.grad = None
            return fwd_bwd()

        for _ in range(2):
            loss = one()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            loss = one()
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        flat.remove()
        ops.DROPOUT_SEED_DEV = None
        for p in params:
            p.grad = None
        return float(t[0]) / iters, float(loss.detach())

    fused = None
    try:
        ms, loss = measure(dropout)
        ms0, loss0 = measure(0.0)
        if ops.get_precision() == "bf16":
            args.fused_ce = True                   # head + label-smoothing loss without [B,U,V] log-probs
            msf, lossf = measure(dropout)
            fused = {"ms_per_step": msf, "utt_per_s": B * world / (msf / 1e3), "loss": lossf,
                     "note": "args.fused_ce: fused linear + label-smoothed cross entropy (no logits / log-probs tensor)"}
    finally:
        args.fused_ce = False
        set_dropout(0.0)
        model.eval()
        ops.DROPOUT_SEED_DEV = None
    return {"workload": f"cfg5 (post-encoder path): base model, {B} x {frames_in / 100:.0f} s per GPU, T={T}, {L} labels/utt, "
                        f"fwd + bwd, dropout {dropout}" + (", bf16 bucketed gradient all_reduce overlapped with backward (in-graph)"
                                                         if world > 1 else ""),
            "ms_per_step": ms, "utt_per_s": B * world / (ms / 1e3), "loss": loss,
            "no_dropout": {"ms_per_step": ms0, "utt_per_s": B * world / (ms0 / 1e3), "loss": loss0},
            "fused_cross_entropy": fused, "params_in_step": n_par, "precision": ops.get_precision(), "cuda_graph": graphed}


def speechlike_decode(model, args, w, dev_in, B, dev, iters=20, target=0.15):
    """The headline workload has U ~ T because random-init CTC posteriors start a new token on almost every
    frame; real speech has U ~ 0.1-0.2 T.  Same model, same inputs, but the CTC head's blank
    bias is raised (bisection on the measured token rate) until tokens/frames ~ 0.15: what the path costs
    at a speech-like token rate.  Restores the bias afterwards."""
    bias = model.ctc_generator.proj.bias
    saved = bias.detach().clone()
    enc_h, mask, ratio = dev_in
    frames = float(mask.sum())

    def rate(b):
        with torch.no_grad():
            bias.copy_(saved)
            bias[0] += b
            out, _ = model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args)
        return sum(len(o[0]["hyp"]) - 2 for o in out) / frames, out

    lo, hi = 0.0, 8.0
    for _ in range(10):
        mid = 0.5 * (lo + hi)
        r, _ = rate(mid)
        lo, hi = (mid, hi) if r > target else (lo, mid)
    r, out = rate(hi)
    for _ in range(3):
        model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for _ in range(iters):
            out, _ = model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    with torch.no_grad():
        bias.copy_(saved)
    toks = [len(o[0]["hyp"]) - 1 for o in out]
    return {"workload": f"cfg3 model and inputs, CTC blank bias +{hi:.2f} so that tokens/frames = {r:.3f}",
            "ms_per_step": ms, "utt_per_s": B / (ms / 1e3), "tokens_U_mean": sum(toks) / len(toks), "tokens_U_max": max(toks),
            "note": "no L2 flush between steps; everything else as the headline run"}


def features_decode(model, args, w, dev, B, iters=7):
    """SURVEY §8f-1 (first 'next' row): the whole `beam_decode(src, ...)` call from 80-dim fbank features —
    conv front-end (direct conv1 + implicit-GEMM conv2 on tcgen05), 12-layer encoder on the same kernels as the
    decoder blocks, then the accelerated path.  CUDA events around each part, wall clock around the call."""
    Tin = w["frames_in"]
    g = torch.Generator().manual_seed(99)
    src = torch.randn(B, Tin, 80, generator=g)
    lens = torch.randint(int(0.6 * Tin), Tin + 1, (B,), generator=g)
    lens[0] = Tin
    for b in range(B):
        src[b, lens[b]:] = 0
    mask = (torch.arange(Tin)[None] < lens[:, None]).unsqueeze(1)
    src, mask, ratio = src.to(dev), mask.to(dev), (lens.float() / Tin).to(dev)
    with torch.no_grad():
        for _ in range(2):
            out, _ = model.beam_decode(src, mask, ratio, VOCAB, args)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        x, xm = model.src_embed(src, mask)
        e[1].record()
        enc_h = model.encoder(x, xm)
        e[2].record()
        model.decode_from_encoder(enc_h, xm, ratio, VOCAB, args)
        e[3].record()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter()
            out, _ = model.beam_decode(src, mask, ratio, VOCAB, args)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        dt = sorted(ts)[len(ts) // 2]              # median call (one run saw a single 150 ms stall among five calls)
    return {"workload": f"cfg3 from features: {B} x {w['audio_s']:.0f} s of 80-dim fbank -> hypotheses (front-end + 12-layer "
                        f"encoder + post-encoder path)", "ms_per_step": dt * 1e3, "utt_per_s": B / dt,
            "rtf": dt / (B * w["audio_s"]), "front_end_ms": e[0].elapsed_time(e[1]), "encoder_ms": e[1].elapsed_time(e[2]),
            "post_encoder_ms": e[2].elapsed_time(e[3]), "tokens_U_mean": sum(len(o[0]["hyp"]) - 1 for o in out) / B,
            "timing": f"wall clock of the call, median of {iters}"}


def conformer_decode(w, dev_in, B, dev, precision, graphs, iters=10):
    """The same workload on the conformer block family (use_conv_dec / relative positions: the reference's shipped
    YAML default): relative-position attention on tcgen05, conv module, macaron half-step FFNs."""
    from cassnat_asr_public_b200.cassnat import make_model
    args = model_args(w, use_conv_dec=True, use_conv_enc=True, pos_type="relative", precision=precision)
    torch.manual_seed(1234)
    model = make_model(80, args).to(dev).eval()
    model.set_precision(precision)
    model.use_cuda_graphs = graphs
    enc_h, mask, ratio = dev_in
    with torch.no_grad():
        for _ in range(3):
            out, _ = model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out, _ = model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    toks = [len(o[0]["hyp"]) - 1 for o in out]
    return {"workload": "cfg3 shapes, conformer decoder family (1+1+6 layers, kernel 3, max relative length 8)",
            "ms_per_step": ms, "utt_per_s": B / (ms / 1e3), "tokens_U_mean": sum(toks) / len(toks), "tokens_U_max": max(toks)}


def features_decode_conformer(w, dev, B, precision, graphs):
    """features_decode on the conformer family end to end: conv front-end, 12 CONFORMER encoder layers (relative-position
    attention, conv module, macaron FFNs: use_conv_enc) and the conformer decoder blocks -- the reference's shipped
    YAML default for both stacks."""
    from cassnat_asr_public_b200.cassnat import make_model
    args = model_args(w, use_conv_dec=True, use_conv_enc=True, pos_type="relative", precision=precision)
    torch.manual_seed(1234)
    model = make_model(80, args).to(dev).eval()
    model.set_precision(precision)
    model.use_cuda_graphs = graphs
    r = features_decode(model, args, w, dev, B)
    r["workload"] = r["workload"].replace("12-layer encoder", "12-layer conformer encoder").replace("post-encoder path", "conformer decoder path")
    return r


class StandInRanker(torch.nn.Module):
    """Ranker with the TransformerLM call contract (ids [N,U], mask [N,U,U]) -> log-probs [N,U,V]
    (src/models/lm.py:52-56).  The real LM is the caller's stock PyTorch module (not part of
    the path); this stand-in is a causal running-mean bag of embeddings so that the rescoring input has
    the real shape and dtype."""

    def __init__(self, V, d=64, seed=7):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = torch.nn.Parameter(torch.randn(V, d, generator=g) * 0.3, requires_grad=False)
        self.out = torch.nn.Parameter(torch.randn(d, V, generator=g) * 0.3, requires_grad=False)

    def forward(self, ids, mask):
        x = self.emb[ids].cumsum(1) / torch.arange(1, ids.size(1) + 1, device=ids.device)[None, :, None]
        return torch.log_softmax(x @ self.out, -1)


def sampled_decode(model, w, dev, B=32, S=10, iters=5, real_lm=False):
    """BASELINE configs[3] / SURVEY cfg4: decode with S=10 sampled alignments per utterance (threshold 0.9,
    host-drawn select bits as cassnat.py:373) + ranker rescoring (K11), batch 32 -> effective batch 320."""
    args = model_args(w, sample_num=S, threshold=0.9, rank_model="lm")
    enc_h, mask, ratio, lens = synth_encoder_output(w, B, 777)
    enc_h, mask, ratio = enc_h.to(dev), mask.to(dev), ratio.to(dev)
    ranker = StandInRanker(5000).to(dev)
    if real_lm:
        # the reference's ranker: TransformerLM of conf/lm.yaml (16 x 512, d_ff 2048, 8 heads; vocabulary widened to
        # the ASR model's 5000 so that every id is in range), random init, on this package's kernels
        from cassnat_asr_public_b200 import lm as LM
        torch.manual_seed(4321)
        ranker = LM.make_model(SimpleNamespace(n_head=8, d_model=512, d_ff=2048, dropout=0.0, vocab_size=5000, N=16)).to(dev).eval()

    def once():
        with torch.no_grad():
            return model.decode_from_encoder(enc_h, mask, ratio, VOCAB, args, lm_model=ranker)

    for _ in range(2):
        out, _ = once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        out, _ = once()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / iters * 1e3
    U = max(len(o[0]["hyp"]) - 1 for o in out)
    ids = torch.randint(1, 5000, (B * S, U), device=dev)
    lm_mask = torch.tril(torch.ones(U, U, dtype=torch.bool, device=dev)).unsqueeze(0).expand(B * S, U, U).contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(iters):
            if real_lm:
                ranker.score_tokens(ids, lm_mask, ids, causal=True)
            else:
                ranker(ids, None)
    torch.cuda.synchronize()
    rk = (time.perf_counter() - t0) / iters * 1e3
    what = ("TransformerLM 16 x 512 (conf/lm.yaml shape) on this package's kernels, fused target log-prob head" if real_lm
            else "stand-in ranker (stock torch, not part of the path)")
    return {"workload": f"cfg4: base model, {B} utterances x {S} sampled alignments (threshold 0.9) + rescoring, T={enc_h.size(1)}, "
                        f"tokens_U_max={U}", "ms_per_step": ms, "utt_per_s": B / (ms / 1e3),
            "alignments_per_s": B * S / (ms / 1e3), "ranker": what, "ranker_ms": rk,
            "note": "wall clock incl. the host RNG draw, the ranker and the hypothesis read-back"}


def alignment_lattice(dev, pk, B=256, T=1000, V=5000, U=200, iters=5):
    """BASELINE configs[1]: CTC best-path + Viterbi forced alignment + trigger masks on synthetic
    posteriors (T=1000, V=5000, U=200, batch 256; 5.12 GB).  Kernel times from CUDA events around the
    C-ABI calls.  "lattice bytes" follow SURVEY §8d: B*T*S*5 + B*T*9 (S = 2U+1)."""
    from cassnat_asr_public_b200 import _lib, ops
    g = torch.Generator(device=dev).manual_seed(1234)
    ys = torch.randint(4, V, (B, U), device=dev, generator=g)
    logits = torch.randn(B, T, V, device=dev, generator=g)
    t_idx = (torch.arange(U, device=dev) * 5)[None, :, None] + torch.arange(1, 4, device=dev)[None, None, :]
    b_idx = torch.arange(B, device=dev)[:, None, None].expand(B, U, 3)
    logits[b_idx, t_idx.expand(B, U, 3), ys[:, :, None].expand(B, U, 3)] += 8.0        # planted alignment
    lp = torch.log_softmax(logits, -1)
    del logits
    mask = torch.ones(B, 1, T, dtype=torch.bool, device=dev)
    xl = torch.full((B,), T, dtype=torch.int64, device=dev)
    yls = torch.full((B,), U, dtype=torch.int64, device=dev)

    def once():
        a, yl, ym = ops.best_path_align(lp, mask, xl, 0)
        ops.align_to_mask(a, yl, ym, mask, xl, 0)
        va, _, _ = ops.viterbi_align(lp, mask, xl, ys, yls, 0)
        ops.align_to_mask(va, yls, U, mask, xl, 0)
        emis = ops.gather_emissions(lp, ys, 0)
        ca, _, _ = ops.viterbi_align_compact(emis, V, mask, xl, ys, yls, 0)
        return va, ca

    for _ in range(2):
        va, ca = once()
    same = bool(torch.equal(va, ca))
    _lib.start_profile()
    for _ in range(iters):
        once()
    prof = _lib.stop_profile()
    ms = {k: v["ms"] / v["calls"] for k, v in prof.items()}
    S = 2 * U + 1
    lattice = B * T * S * 5 + B * T * 9
    vit, bp = ms["cn_viterbi_align"], ms["cn_bestpath_from_logp"]
    vc, ge = ms["cn_viterbi_align_compact"], ms["cn_gather_emissions"]
    return {"workload": f"cfg2: B={B} T={T} V={V} U={U} planted posteriors ({4 * B * T * V / 1e9:.2f} GB f32)",
            "viterbi_ms": vit, "viterbi_utt_per_s": B / (vit / 1e3),
            "viterbi_lattice_GBps": lattice / (vit / 1e3) / 1e9, "viterbi_lattice_frac_of_hbm": lattice / (vit / 1e3) / 1e9 / pk["hbm"],
            "viterbi_gather_traffic_GBps": B * T * (U + 1) * 64.0 / (vit / 1e3) / 1e9,
            "viterbi_gather_traffic_frac_of_hbm": B * T * (U + 1) * 64.0 / (vit / 1e3) / 1e9 / pk["hbm"],
            "viterbi_note": "from [B,T,V] posteriors the kernel is bound by the emission gather: one 64-B DRAM atom per "
                            "(frame, label) = B*T*(U+1)*64 bytes of real traffic (ncu: profiles/r02_summary.md); the "
                            "lattice figure above divides SURVEY §8d's algorithmic bytes by the same time",
            "viterbi_compact_ms": vc, "viterbi_compact_lattice_GBps": lattice / (vc / 1e3) / 1e9,
            "viterbi_compact_lattice_frac_of_hbm": lattice / (vc / 1e3) / 1e9 / pk["hbm"],
            "gather_emissions_ms": ge, "compact_alignments_identical": same,
            "viterbi_compact_note": "the same DP on compact [B,T,U+1] emissions (gathered once by cn_gather_emissions; shareable "
                                    "with the CTC loss): 4(U+1) coalesced bytes per frame instead of a 64-B atom per value",
            "bestpath_ms": bp, "bestpath_GBps": 4.0 * B * T * V / (bp / 1e3) / 1e9,
            "bestpath_frac_of_hbm": 4.0 * B * T * V / (bp / 1e3) / 1e9 / pk["hbm"],
            "align_to_mask_ms": ms["cn_align_to_mask"],
            "align_to_mask_GBps": (8.0 * B * T + B * (U + 1) * T) / (ms["cn_align_to_mask"] / 1e3) / 1e9}


def cpu_baseline(a, w, args):
    """Oracle port on the host cores, bounded sample (rank 0, N=1 only).  Reported, not a target."""
    import oracle.nat_torch as O
    from cassnat_asr_public_b200.cassnat import make_model
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1234)
    cargs = SimpleNamespace(**{**vars(args), "use_gpu": False})
    sd = {k: v.detach() for k, v in make_model(80, cargs).state_dict().items()}
    sample, chunk = a.cpu_sample, 16
    enc_h, mask, ratio, lens = synth_encoder_output(w, sample, 1234)
    frames = (ratio * enc_h.size(1)).long()
    with torch.no_grad():
        O.decode_from_encoder(enc_h[:1], mask[:1], frames[:1], sd, cargs)            # warm-up
        t0 = time.perf_counter()
        for i in range(0, sample, chunk):                # bounded sample (~10-20 s of CPU work), 16 utterances per call
            O.decode_from_encoder(enc_h[i:i + chunk], mask[i:i + chunk], frames[i:i + chunk], sd, cargs)
        dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "utt/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample} utterances of the workload in calls of {chunk} ({dt:.1f} s; oracle/nat_torch.py, torch CPU f32)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--precision", default=os.environ.get("CASSNAT_PRECISION", "bf16"), choices=["fp32", "bf16", "bf16x3"])
    ap.add_argument("--cpu-sample", type=int, default=192, help="utterances of the cpu_baseline leg (~10-20 s of CPU work)")
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="utterances per step of the --impl reference arm (0: the full batch if the run stays within minutes)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the per-GPU batch is fixed; strong: the workload's batch is split over the GPUs")
    ap.add_argument("--no-parity", action="store_true", help="skip the fast-mode vs parity-mode agreement report")
    ap.add_argument("--no-eager", action="store_true", help="skip the stock-PyTorch-eager-on-GPU comparator")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lattice", action="store_true", help="skip the cfg2 alignment-lattice microbench")
    ap.add_argument("--no-sampled", action="store_true", help="skip the cfg4 sampled-alignment decode measurement")
    ap.add_argument("--no-training", action="store_true", help="skip the cfg5 training-step measurement")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    w = WORKLOADS[a.workload]
    if a.impl == "reference":
        run_reference(a, w)
    else:
        run_ours(a, w)


if __name__ == "__main__":
    main()
