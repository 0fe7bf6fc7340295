cd "$(dirname "$0")/.."; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int4_attention.py tests/test_gpu_model.py tests/test_gpu_bench_shapes.py tests/test_gpu_harness.py tests/test_gpu_kv_ops.py -q -x -k "int4 or INT4 or kv4 or harness or quant or rope" > gpurun_out/int4_fused_tests.log 2>&1; tail -5 gpurun_out/int4_fused_tests.log
timeout 400 python bench.py --kv-format int4 --no-prefill --no-cpu-baseline --no-fa2 > gpurun_out/int4_fused_bench.json 2> gpurun_out/int4_fused_bench.err; tail -c 1500 gpurun_out/int4_fused_bench.json; tail -3 gpurun_out/int4_fused_bench.err
