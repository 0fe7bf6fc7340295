"""Host-side checks of the efficiency harness port (eval/efficiency/benchmark_static.py): CLI surface and the exact
result-file format of the reference (eval/efficiency/benchmark_static.py:107-119 there)."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("benchmark_static", os.path.join(ROOT, "eval", "efficiency",
                                                                                    "benchmark_static.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_result_file_has_the_reference_lines():
    m = _load()
    txt = m.format_result(13.91234, 40000.5, 4100.25, 52000.125, "llama3-8b-1048k", 100000, 0.5, 32000, 1234.5)
    assert txt.splitlines() == [
        "Average generation time: 13.9123 ms",
        "Peak generation memory usage: 40000.5000 MB",
        "Average context time: 4100.2500 ms",
        "Peak context memory usage: 52000.1250 MB",
        "Model name: llama3-8b-1048k",
        "Context length: 100000",
        "Sparsity: 0.5",
        "Prefilling chunk size: 32000",
        "KV cache memory usage: 1234.5000 MB",
    ]


def test_cli_matches_reference_names_and_defaults():
    m = _load()
    a = m.parse_args(["--random_init", "llama3-8b-1048k", "--attn_load_dir", "x"])
    assert (a.max_length, a.prefilling_chunk_size, a.device, a.seed, a.output_dir, a.sparsity, a.threshold) == \
        (4096, 4096, "0", 42, "outputs", None, 0.5)
    with pytest.raises(SystemExit):
        m.parse_args([])  # needs a model
    with pytest.raises(SystemExit):
        m.parse_args(["--random_init", "llama3-8b-1048k", "--model_name", "/x"])
    assert set(m.ARCHS) == {"llama3-8b-1048k", "llama3-8b-4194k", "llama2-7b-32k", "mistral-7b-v0.3"}


def test_needs_cuda_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = _load()
    with pytest.raises(RuntimeError, match="CUDA"):
        m.main(["--random_init", "llama2-7b-32k", "--attn_load_dir", "x"])
