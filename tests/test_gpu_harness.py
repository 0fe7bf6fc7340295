"""The ported efficiency harness (eval/efficiency/benchmark_static.py, reference protocol) end to end on the GPU."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("extra", [[], ["--cuda_graph"], ["--kv_format", "int4"]])
def test_benchmark_static_harness_smoke(tmp_path, extra):
    """eval/efficiency/benchmark_static.py (reference protocol) end to end on a 2-layer random-init model."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("benchmark_static",
                                                  os.path.join(root, "eval", "efficiency", "benchmark_static.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pat = os.path.join(root, "attn_patterns", "Llama-3-8B-Instruct-Gradient-1048k",
                       "lr=0.02-reg=0.05-ctx=1000_32000-multi_passkey10")
    mod.main(["--random_init", "llama3-8b-1048k", "--num_layers", "2", "--attn_load_dir", pat, "--sparsity", "0.5",
              "--max_length", "3000", "--prefilling_chunk_size", "1024", "--ctx_steps", "1", "--gen_steps", "3",
              "--output_dir", str(tmp_path)] + extra)
    lines = open(tmp_path / "benchmark_result.txt").read().splitlines()
    assert len(lines) == 9 and lines[0].startswith("Average generation time: ") and lines[5] == "Context length: 3000"
