#!/bin/bash
# One-GPU validation + A/B timing of the opt-in experimental paths (run through gpurun, ~10 GPU-minutes):
#   gpurun --timeout 1500 -- 'bash profiles/validate_experimental.sh > gpurun_out/validate_experimental.log 2>&1'
# Every step runs under its own `timeout`; results land in gpurun_out/exp_*.{log,json}.
#   DUO_INT4_SWAPAB=1   INT4 decode kernel with keys as the MMA M dimension (attn_int4.cu: duo_attn_int4_dec8_kernel)
#   DUO_INT4_FAST=1     shipped INT4 kernels with the single-LOP3 conversion + interior-tile loader path (template FAST)
#   DUO_TC_WARP_ARRIVE=1  tcgen05 prefill kernel: one p_full mbarrier arrival per softmax warp instead of per thread
#   DUO_WIDE_MERGE=1    split-KV last-CTA merge with 16 loads in flight (duo_common.cuh: split_merge_rows4)
#   DUO_INT4_PREFILL_SCRATCH=1  INT4 chunks >= 128 tokens: dequantise to an fp16 scratch + tcgen05 kernel (kv_cache.py)
# (DUO_FUSED_ALLREDUCE=1 needs 2 GPUs: tests/multi_gpu/fused_allreduce_check.py, then tests/multi_gpu/tp_check.py)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@"; echo "=== rc=$?"; }

# 1. parity under each switch (the regular suites exercise the switched kernels)
DUO_EXPERIMENTAL=1 DUO_INT4_SWAPAB=1 run python -m pytest tests/test_gpu_int4_attention.py tests/test_gpu_experimental.py -x -q \
  > gpurun_out/exp_int4_swapab_tests.log 2>&1
DUO_EXPERIMENTAL=1 DUO_INT4_FAST=1 run python -m pytest tests/test_gpu_int4_attention.py tests/test_gpu_experimental.py -x -q \
  > gpurun_out/exp_int4_fast_tests.log 2>&1
DUO_WIDE_MERGE=1 run python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py tests/test_gpu_int4_attention.py -x -q \
  > gpurun_out/exp_wide_merge_tests.log 2>&1
DUO_EXPERIMENTAL=1 DUO_INT4_PREFILL_SCRATCH=1 run python -m pytest tests/test_gpu_int4_attention.py tests/test_gpu_experimental.py -x -q \
  > gpurun_out/exp_int4_scratch_tests.log 2>&1
tail -3 gpurun_out/exp_int4_swapab_tests.log gpurun_out/exp_int4_fast_tests.log gpurun_out/exp_wide_merge_tests.log gpurun_out/exp_int4_scratch_tests.log

DUO_TC_WARP_ARRIVE=1 run python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py -x -q \
  > gpurun_out/exp_tc_warp_arrive_tests.log 2>&1
tail -3 gpurun_out/exp_tc_warp_arrive_tests.log
run python profiles/bench_tc.py > gpurun_out/exp_tc_base.log 2>&1
DUO_TC_WARP_ARRIVE=1 run python profiles/bench_tc.py > gpurun_out/exp_tc_warp_arrive.log 2>&1
tail -2 gpurun_out/exp_tc_base.log gpurun_out/exp_tc_warp_arrive.log

# 2. A/B timing: default bench line (bf16 decode @1M) with and without the wide merge
run python bench.py --steps 8 --warmup 3 > gpurun_out/exp_bf16_base.json 2> gpurun_out/exp_bf16_base.err
DUO_WIDE_MERGE=1 run python bench.py --steps 8 --warmup 3 > gpurun_out/exp_bf16_wide.json 2> gpurun_out/exp_bf16_wide.err

# 3. A/B timing: INT4 decode @1M, current kernel vs swapped-operand kernel (both with the wide merge off, then on)
run python bench.py --kv-format int4 --no-prefill --steps 8 --warmup 3 > gpurun_out/exp_int4_base.json 2> gpurun_out/exp_int4_base.err
DUO_INT4_FAST=1 run python bench.py --kv-format int4 --no-prefill --steps 8 --warmup 3 > gpurun_out/exp_int4_fast.json 2> gpurun_out/exp_int4_fast.err
DUO_INT4_SWAPAB=1 run python bench.py --kv-format int4 --no-prefill --steps 8 --warmup 3 > gpurun_out/exp_int4_swapab.json 2> gpurun_out/exp_int4_swapab.err
DUO_EXPERIMENTAL=1 DUO_INT4_SWAPAB=1 DUO_WIDE_MERGE=1 run python bench.py --kv-format int4 --no-prefill --steps 8 --warmup 3 \
  > gpurun_out/exp_int4_swapab_graph.json 2> gpurun_out/exp_int4_swapab_graph.err   # + CUDA-graph decode driver
# 4. INT4 prefill @128K through the scratch + tcgen05 path (the default INT4 chunk kernel is mma.sync: much slower; skip it)
DUO_INT4_SWAPAB=1 DUO_INT4_PREFILL_SCRATCH=1 run python bench.py --kv-format int4 --prefill-reps 1 --steps 4 --warmup 3 \
  > gpurun_out/exp_int4_scratch_prefill.json 2> gpurun_out/exp_int4_scratch_prefill.err
# 5. reference-protocol harness (eval/efficiency/benchmark_static.py), Llama-3-8B-1048k arch, 100K context
run python eval/efficiency/benchmark_static.py --random_init llama3-8b-1048k --sparsity 0.5 --max_length 100000 \
  --prefilling_chunk_size 32000 --ctx_steps 2 --gen_steps 50 --cuda_graph --output_dir gpurun_out/exp_harness \
  --attn_load_dir attn_patterns/Llama-3-8B-Instruct-Gradient-1048k/lr=0.02-reg=0.05-ctx=1000_32000-multi_passkey10 \
  > gpurun_out/exp_harness.log 2>&1
cat gpurun_out/exp_harness/benchmark_result.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/exp_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value", round(d["value"], 2), d["unit"], "attn_ms", round(d["roofline"]["attn_ms_per_step"], 3),
              "frac", round(d["roofline"]["frac"], 3), "prefill", (d.get("prefill") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
