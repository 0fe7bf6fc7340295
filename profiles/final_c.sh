#!/bin/bash
# Final single-GPU run of round 2 with the last kernels (INT4 one-launch decode, packed fp32 pairs, overlapped issuer waits):
# the whole GPU suite, the two headline bench lines, and the ncu captures of the two kernels that changed.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_final.log
timeout 600 python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err
timeout 400 python bench.py --kv-format int4 --prefill-reps 1 > gpurun_out/r2_final_int4.json 2> gpurun_out/r2_final_int4.err
NCU="ncu --clock-control none"
timeout 300 $NCU --set full --import-source on -k regex:duo_attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_prefill \
    python profiles/bench_tc.py > gpurun_out/ncu_prefill_full_stdout.log 2>&1
timeout 600 bash profiles/capture_int4.sh > gpurun_out/capture_int4.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -c 300 gpurun_out/r2_final_n1.json; echo; tail -c 300 gpurun_out/r2_final_int4.json
