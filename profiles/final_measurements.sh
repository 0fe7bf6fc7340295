#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_final.log 2>&1; tail -4 gpurun_out/r2_gpu_tests_final.log
timeout 900 bash profiles/capture.sh > gpurun_out/capture.log 2>&1
timeout 600 bash profiles/capture_int4.sh > gpurun_out/capture_int4.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err
timeout 400 python bench.py --kv-format int4 --prefill-reps 1 > gpurun_out/r2_final_int4.json 2> gpurun_out/r2_final_int4.err
timeout 300 python bench.py --arch mistral-7b-v0.3 --ctx 65536 --prefill-ctx 65536 --no-fa2 > gpurun_out/r2_final_mistral64k.json 2> gpurun_out/r2_final_mistral64k.err
timeout 300 python bench.py --arch llama2-7b-32k --ctx 32768 --prefill-ctx 32768 --chunk 8192 --sparsity 0.75 --no-fa2 > gpurun_out/r2_final_llama2_32k.json 2> gpurun_out/r2_final_llama2_32k.err
PAT=attn_patterns/Llama-3-8B-Instruct-Gradient-1048k/lr=0.02-reg=0.05-ctx=1000_32000-multi_passkey10
for V in "" "--cuda_graph" "--kv_format int4 --cuda_graph"; do
  timeout 300 python eval/efficiency/benchmark_static.py --random_init llama3-8b-1048k --sparsity 0.5 --max_length 100000 \
    --prefilling_chunk_size 32000 --ctx_steps 2 --gen_steps 50 $V --output_dir "gpurun_out/harness_$(echo $V | tr -d ' -')" \
    --attn_load_dir $PAT > "gpurun_out/harness_$(echo $V | tr -d ' -').log" 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_final_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        p = d.get("prefill") or {}
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), "attn_ms",
              round(d["roofline"]["attn_ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"],
              "prefill", p.get("value"), (p.get("roofline") or {}).get("frac"))
        if "reference_gpu_same_box" in d: print("   ref:", json.dumps(d["reference_gpu_same_box"])[:600])
    except Exception as e:
        print(f, "unreadable:", e)
PY
for d in gpurun_out/harness_*/; do echo $d; cat $d/benchmark_result.txt | head -4; done
