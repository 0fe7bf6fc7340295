#!/bin/bash
# A/B of tuning builds of the tcgen05 prefill kernel (duo_attention_b200/csrc/_var/libduo_<name>.so) against the default
# build on one GPU:  bash profiles/tc_ab.sh name1 name2 ...   (isolated 32K x 128K layer, then the tc parity tests)
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; rm -f gpurun_out/tc_base_out.pt
V=$PWD/duo_attention_b200/csrc/_var
export DUO_TC_REF=gpurun_out/tc_base_out.pt DUO_TC_REPS=6
{
  timeout 200 python profiles/bench_tc.py
  for n in "$@"; do DUO_B200_LIB=$V/libduo_$n.so timeout 200 python profiles/bench_tc.py; done
  timeout 200 python profiles/bench_tc.py
} 2>&1 | grep -v Warning | tee gpurun_out/tc_ab.log
rm -f gpurun_out/tc_base_out.pt
unset DUO_TC_REF
for n in "$@"; do
  DUO_B200_LIB=$V/libduo_$n.so timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_bench_shapes.py tests/test_gpu_oracle_pin.py -q -x -k "tc or prefill or accuracy or sharp or benchmarked or batch2 or fp16" > gpurun_out/tc_parity_$n.log 2>&1
  echo "parity under $n:"; tail -4 gpurun_out/tc_parity_$n.log
done
