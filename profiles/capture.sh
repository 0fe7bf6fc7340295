#!/bin/bash
# Run under gpurun (one GPU).  Produces under gpurun_out/:
#   launches_decode.csv    every kernel launch of ONE timed decode step (eager driver, nvtx range timed_decode)
#   launches_prefill.csv   every kernel launch of ONE timed 128K prefill (nvtx range timed_prefill)
#   prof_decode.ncu-rep    `--set full` capture of the split-KV decode kernel (2 launches of a decode step)
#   prof_prefill.ncu-rep   `--set full` capture of the tcgen05 prefill kernel on one layer's last 32K chunk
#                          (profiles/bench_tc.py: 32K queries over 96K past, n_full = 4)
# Numbers printed by a run under ncu are never bench values.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
COMMON="--steps 1 --warmup 3 --no-cpu-baseline --no-fa2 --no-graph"
$NCU --nvtx --nvtx-include "timed_decode/" --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_decode.csv \
    python bench.py $COMMON --no-prefill > gpurun_out/ncu_decode_stdout.log 2>&1
$NCU --set full --import-source on -k regex:duo_attn_mma_kernel -s 100 -c 2 -o gpurun_out/prof_decode \
    python bench.py $COMMON --no-prefill > gpurun_out/ncu_decode_full_stdout.log 2>&1
$NCU --nvtx --nvtx-include "timed_prefill/" --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_prefill.csv \
    python bench.py $COMMON --prefill-reps 1 > gpurun_out/ncu_prefill_stdout.log 2>&1
$NCU --set full --import-source on -k regex:duo_attn_tc_kernel -s 2 -c 1 -o gpurun_out/prof_prefill \
    python profiles/bench_tc.py > gpurun_out/ncu_prefill_full_stdout.log 2>&1
ls -la gpurun_out/
