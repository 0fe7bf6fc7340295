#!/bin/bash
# Final single-GPU run, part B (the whole GPU suite with the parity log, the other configurations, the harness).
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1; tail -4 gpurun_out/r2_gpu_tests_final.log
timeout 300 python bench.py --arch mistral-7b-v0.3 --ctx 65536 --prefill-ctx 65536 --no-fa2 > gpurun_out/r2_final_mistral64k.json 2> gpurun_out/r2_final_mistral64k.err
timeout 300 python bench.py --arch llama2-7b-32k --ctx 32768 --prefill-ctx 32768 --chunk 8192 --sparsity 0.75 --no-fa2 > gpurun_out/r2_final_llama2_32k.json 2> gpurun_out/r2_final_llama2_32k.err
PAT=attn_patterns/Llama-3-8B-Instruct-Gradient-1048k/lr=0.02-reg=0.05-ctx=1000_32000-multi_passkey10
for V in "" "--cuda_graph" "--kv_format int4 --cuda_graph"; do
  timeout 300 python eval/efficiency/benchmark_static.py --random_init llama3-8b-1048k --sparsity 0.5 --max_length 100000 \
    --prefilling_chunk_size 32000 --ctx_steps 2 --gen_steps 50 $V --output_dir "gpurun_out/harness_$(echo $V | tr -d ' -')" \
    --attn_load_dir $PAT > "gpurun_out/harness_$(echo $V | tr -d ' -').log" 2>&1
done
for d in gpurun_out/harness_*/; do echo $d; cat $d/benchmark_result.txt | head -4; done
tail -c 400 gpurun_out/r2_final_mistral64k.json; tail -c 400 gpurun_out/r2_final_llama2_32k.json
