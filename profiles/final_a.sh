#!/bin/bash
# Final single-GPU run, part A (ncu captures + the two headline bench lines).  Run under gpurun (one GPU).
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
timeout 900 bash profiles/capture.sh > gpurun_out/capture.log 2>&1
timeout 600 bash profiles/capture_int4.sh > gpurun_out/capture_int4.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err
timeout 400 python bench.py --kv-format int4 --prefill-reps 1 > gpurun_out/r2_final_int4.json 2> gpurun_out/r2_final_int4.err
tail -c 600 gpurun_out/r2_final_n1.json; tail -c 600 gpurun_out/r2_final_int4.json
