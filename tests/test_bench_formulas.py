"""The algorithmic work bench.py divides by (roofline numerators) equals BASELINE.md §4, and the reference arm /
JSON contract keys exist.  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_decode_bytes_and_prefill_flops_match_baseline_md():
    mask, sp = bench.head_pattern("Llama-3-8B-Instruct-Gradient-1048k", 0.5)
    assert mask.shape == (32, 8) and mask.sum() == 128 and abs(sp - 0.5) < 1e-12
    gb = lambda n, rb=256: bench.decode_bytes_per_token(mask, n, rb) / 1e9
    assert abs(gb(131072) - 8.61) < 0.01
    assert abs(gb(1048576) - 68.74) < 0.01
    assert abs(gb(2097152) - 137.5) < 0.1
    assert abs(gb(1048576, 68) - 18.26) < 0.01
    assert abs(bench.prefill_flops(mask, 131072, 32768) / 1e15 - 2.823) < 0.001
    assert abs(bench.prefill_flops(mask, 131072, 131072) / 1e15 - 4.504) < 0.001
    assert abs(bench.prefill_flops(mask, 65536, 32768) / 1e15 - 0.847) < 0.001


def test_every_architecture_has_a_pattern():
    for arch, (_, pat) in bench.ARCHS.items():
        m, _ = bench.head_pattern(pat, 0.5)
        assert m.shape[0] == 32 and m.shape[1] in (8, 32), arch


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--ctx", "4096"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["cores"] >= 1
