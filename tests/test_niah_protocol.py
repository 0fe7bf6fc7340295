"""Call sequence of the NIAH generation protocol port (eval/needle/niah_protocol.py) with a recording stand-in model:
chunked pre-fill, one-token simulation calls, greedy loop with EOS stop — needle_in_haystack.py:262-314."""
import importlib.util
import os
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _protocol():
    spec = importlib.util.spec_from_file_location("niah_protocol", os.path.join(ROOT, "eval", "needle", "niah_protocol.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Recorder:
    def __init__(self, vocab=16, eos_at=None):
        self.calls, self.vocab, self.eos_at, self.n = [], vocab, eos_at, 0

    def __call__(self, input_ids, past_key_values, use_cache):
        assert use_cache is True
        self.calls.append((tuple(input_ids.shape), past_key_values))
        self.n += 1
        logits = torch.zeros(1, input_ids.shape[1], self.vocab)
        nxt = 3 if self.eos_at is None or self.n < self.eos_at else 7
        logits[0, -1, nxt] = 1.0
        return types.SimpleNamespace(logits=logits, past_key_values=self.n)


def test_call_sequence_matches_the_reference_harness():
    m = Recorder()
    gen, past = _protocol().niah_generate(m, torch.zeros(1, 100, dtype=torch.long), simulation_length=10,
                                          prefilling_chunk_size=32, max_new_tokens=5)
    shapes = [c[0] for c in m.calls]
    assert shapes == [(1, 32), (1, 32), (1, 26)] + [(1, 1)] * 10 + [(1, 1)] * 5
    assert [c[1] for c in m.calls] == [None] + list(range(1, len(m.calls)))   # past fed back verbatim
    assert gen == [3] * 6 and past == len(m.calls)


def test_single_shot_prefill_and_eos_stop():
    m = Recorder(eos_at=14)
    gen, _ = _protocol().niah_generate(m, torch.zeros(1, 60, dtype=torch.long), simulation_length=10,
                                       prefilling_chunk_size=None, max_new_tokens=50, eos_token_ids=[7])
    assert [c[0] for c in m.calls][:1] == [(1, 50)] and gen[-1] == 7 and len(m.calls) == 14
