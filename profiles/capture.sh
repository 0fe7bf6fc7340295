#!/bin/bash
# Run under gpurun (one GPU).  Produces under gpurun_out/:
#   launches_decode.csv   every kernel launch of the timed decode steps with its device time (cold, serialised)
#   launches_prefill.csv  same for the timed 128K prefill
#   prof_decode.ncu-rep   `--set full` capture of the split-KV decode kernel (3 launches)
#   prof_prefill.ncu-rep  `--set full` capture of the tcgen05 prefill kernel (2 launches)
# Numbers printed by a run under ncu are never bench values.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --nvtx --nvtx-include "timed_decode/" --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_decode.csv \
    python bench.py --steps 1 --warmup 3 --no-prefill --no-cpu-baseline --no-fa2 > gpurun_out/ncu_decode_stdout.log 2>&1
$NCU --set full --import-source on -k regex:duo_attn_mma_kernel -s 200 -c 3 -o gpurun_out/prof_decode \
    python bench.py --steps 1 --warmup 3 --no-prefill --no-cpu-baseline --no-fa2 > gpurun_out/ncu_decode_full_stdout.log 2>&1
$NCU --nvtx --nvtx-include "timed_prefill/" --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_prefill.csv \
    python bench.py --steps 1 --warmup 3 --prefill-reps 1 --no-cpu-baseline --no-fa2 > gpurun_out/ncu_prefill_stdout.log 2>&1
$NCU --set full --import-source on -k regex:duo_attn_tc_kernel -s 110 -c 2 -o gpurun_out/prof_prefill \
    python bench.py --steps 1 --warmup 3 --prefill-reps 1 --no-cpu-baseline --no-fa2 > gpurun_out/ncu_prefill_full_stdout.log 2>&1
ls -la gpurun_out/
