"""Turns the ncu outputs of scratch/profile_round.sh (gpurun_out/) into profiles/: launch-list table, GEMM DRAM
traffic per launch (profiles/ncu_traffic.json, read by bench.py's roofline object) and the metric tables of the
full captures.  Run in the build container after the gpurun call."""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launch_table():
    rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", re.sub(r"^void ", "", r[ix["Kernel Name"]])).replace("<unnamed>::", "")
        name = name.replace("(anonymous namespace)::", "")
        v, unit = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
        v = v / 1000.0 if unit in ("nsecond", "ns") else (v * 1000.0 if unit in ("msecond", "ms") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [f"| `{k[:70]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |" for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    return sum(a[0] for a in agg.values()), tot, lines[:24]


def gemm_traffic():
    rows = [r for r in csv.reader(open(os.path.join(G, "gemm_traffic.csv"))) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    per = {}
    for r in rows[1:]:
        v, u = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
        v *= {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9}.get(u, 1.0)
        per.setdefault(r[ix["ID"]], {})[r[ix["Metric Name"]]] = v
    n = len(per)
    return n, sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in per.values()) / n


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def full_table(rep):
    txt = subprocess.run(["ncu", "-i", os.path.join(G, rep + ".ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    return "\n".join(["| metric | value | unit |", "|---|---:|---|"] + [f"| `{w}` | {d[w][0]} | {d[w][1]} |" for w in WANT if w in d])


if __name__ == "__main__":
    n, tot, lines = launch_table()
    ng, tr = gemm_traffic()
    shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, "r02_bench_launches.csv"))
    shutil.copy(os.path.join(G, "gemm_traffic.csv"), os.path.join(P, "r02_gemm_traffic.csv"))
    json.dump({"cn_gemm_bf16": tr, "_note": f"average dram__bytes_read.sum + dram__bytes_write.sum per launch over {ng} consecutive "
               "gemm_bf16_tc_kernel launches of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lattice --no-training "
               "--no-sampled --no-parity --no-eager`; ncu --metrics, --clock-control none (profiles/r02_gemm_traffic.csv)"},
              open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    out = {"launches": n, "total_us": tot, "launch_table": lines, "gemm": full_table("one_gemm_full"), "attn": full_table("one_attn_full"),
           "gemm_traffic": tr}
    json.dump(out, open("/tmp/profile_summary.json", "w"), indent=1)
    print(n, tot, tr)
    print("\n".join(lines[:12]))
