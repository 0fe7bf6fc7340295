#!/bin/bash
# Run under gpurun (one GPU) AFTER profiles/validate_experimental.sh has passed.  Produces under gpurun_out/:
#   prof_int4_base.ncu-rep     `--set full` capture of the shipped INT4 decode kernel (duo_attn_int4_kernel<4>)
#   prof_int4_swapab.ncu-rep   same for the swapped-operand kernel (duo_attn_int4_dec8_kernel, DUO_INT4_SWAPAB=1)
# Read here with: ncu -i gpurun_out/prof_int4_swapab.ncu-rep --page raw --csv | grep -E
#   "gpu__time_duration|dram__throughput|sm__inst_executed_pipe_tensor|smsp__issue_active|achieved_occupancy|registers"
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
COMMON="--kv-format int4 --steps 1 --warmup 3 --no-cpu-baseline --no-fa2 --no-graph --no-prefill"
timeout 900 $NCU --set full --import-source on -k regex:duo_attn_int4_kernel -s 100 -c 2 -o gpurun_out/prof_int4_base \
    python bench.py $COMMON > gpurun_out/ncu_int4_base_stdout.log 2>&1
DUO_INT4_SWAPAB=1 timeout 900 $NCU --set full --import-source on -k regex:duo_attn_int4_dec8_kernel -s 100 -c 2 \
    -o gpurun_out/prof_int4_swapab python bench.py $COMMON > gpurun_out/ncu_int4_swapab_stdout.log 2>&1
ls -la gpurun_out/
