"""The generation protocol of the reference's Needle-in-a-Haystack harness (eval/needle/needle_in_haystack.py:262-314) on
token ids, for any model that follows the drop-in call protocol (``model(input_ids=..., past_key_values=..., use_cache=True)``
-> ``.logits[:, -1]``, ``.past_key_values``):

  1. the prompt minus its last ``simulation_length`` tokens is pre-filled, in chunks of ``prefilling_chunk_size`` tokens
     when that is set (:274-291);
  2. the remaining prompt tokens are fed ONE AT A TIME ("simulate multi-round conversation", :293-299);
  3. greedy decoding of up to ``max_new_tokens`` tokens, stopping at an EOS id (:301-314).

Tokenizer, haystack text, needle placement and the ROUGE scorer are the harness's own business (out of scope: no
tokenizers or datasets offline); this is the part that drives the hot path, and what tests/test_gpu_niah.py runs on a
tiny random-init model against the CPU oracle."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch


@torch.no_grad()
def niah_generate(model, prompt_input_ids: torch.Tensor, simulation_length: int = 50,
                  prefilling_chunk_size: Optional[int] = None, max_new_tokens: int = 50,
                  eos_token_ids: Iterable[int] = (), past_key_values=None, collect_logits: bool = False):
    """-> ``(generated token ids, past_key_values[, list of last-position logits of every call])``."""
    eos = set(int(e) for e in eos_token_ids)
    n = prompt_input_ids.size(1)
    start = n - simulation_length
    question, context = prompt_input_ids[:, start:], prompt_input_ids[:, :start]
    logits: List[torch.Tensor] = []

    def call(ids, past):
        out = model(input_ids=ids, past_key_values=past, use_cache=True)
        if collect_logits:
            logits.append(out.logits[:, -1, :].float().cpu())
        return out

    past = past_key_values
    step = prefilling_chunk_size if prefilling_chunk_size is not None else max(context.size(1), 1)
    for i in range(0, context.size(1), step):
        out = call(context[:, i : i + step], past)
        past = out.past_key_values
    for tok in question[0]:
        out = call(tok.view(1, 1), past)
        past = out.past_key_values
    pred = out.logits[:, -1, :].argmax(dim=-1).unsqueeze(1)
    generated = [int(pred.item())]
    for _ in range(max_new_tokens):
        if generated[-1] in eos:
            break
        out = call(pred, past)
        past = out.past_key_values
        pred = out.logits[:, -1, :].argmax(dim=-1).unsqueeze(1)
        generated.append(int(pred.item()))
    return (generated, past, logits) if collect_logits else (generated, past)
