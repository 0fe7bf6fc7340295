#!/bin/bash
# Run under gpurun (one GPU).  Produces gpurun_out/prof_int4_dec8.ncu-rep: `--set full` capture (with source) of two
# launches of the INT4 decode kernel (duo_attn_int4_dec8_kernel) inside a 1M-token decode step of bench.py.
# Read here with: ncu -i gpurun_out/prof_int4_dec8.ncu-rep --page raw --csv / --page source --csv
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
COMMON="--kv-format int4 --steps 1 --warmup 3 --no-cpu-baseline --no-fa2 --no-graph --no-prefill"
timeout 600 $NCU --nvtx --nvtx-include "timed_decode/" --metrics gpu__time_duration.sum --csv \
    --log-file gpurun_out/launches_int4.csv python bench.py $COMMON > gpurun_out/ncu_int4_launches_stdout.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:duo_attn_int4_dec8_kernel -s 100 -c 2 -f \
    -o gpurun_out/prof_int4_dec8 python bench.py $COMMON > gpurun_out/ncu_int4_dec8_stdout.log 2>&1
ls -la gpurun_out/*.ncu-rep
