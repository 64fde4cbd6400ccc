#!/bin/bash
# Round profile capture: each ncu pass runs only after the same command exited 0 without ncu.
# Usage (on the GPU box): bash scratch/profile_round.sh
set -u
ARGS="--steps 3 --warmup 3 --no-cpu-baseline --no-lattice --no-training --no-sampled --no-parity --no-eager"
timeout 300 python bench.py $ARGS > gpurun_out/prof_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
echo "plain bench ok"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/prof_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:gemm_bf16_tc -s 120 -c 60 --csv --log-file gpurun_out/gemm_traffic.csv python bench.py $ARGS > gpurun_out/prof_ncu2.log 2>&1
echo "gemm traffic rc=$?"
for s in one_gemm one_attn; do
  timeout 200 python scratch/$s.py > gpurun_out/$s.log 2>&1 || { echo "$s failed"; continue; }
  k=gemm_bf16_tc; [ $s = one_attn ] && k=attention_tc
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${s}_full \
    python scratch/$s.py > gpurun_out/${s}_ncu.log 2>&1
  echo "$s full rc=$?"
done
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv gpurun_out/gemm_traffic.csv
