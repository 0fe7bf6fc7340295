"""Sequence-sharding index arithmetic (duo_attention_b200/seqshard.py) and the slice-merge algebra the device kernels
implement (duo_attention_partial / duo_merge_partials): attention over position slices + log-sum-exp merge == attention
over the whole cache."""
import numpy as np
import pytest
import torch

from duo_attention_b200.seqshard import SeqShardPlan


@pytest.mark.parametrize("world,block", [(1, 4), (2, 4), (3, 5), (8, 1024), (4, 1)])
def test_plan_is_a_balanced_ordered_partition(world, block):
    plan = SeqShardPlan(world, block)
    for n in [0, 1, block - 1, block, block + 1, block * world, block * world + 3, 7 * block * world + block // 2 + 1, 5000]:
        if n < 0:
            continue
        parts = plan.split(n)
        allpos = torch.cat(parts)
        assert sorted(allpos.tolist()) == list(range(n))                      # partition
        lens = [len(p) for p in parts]
        assert lens == [plan.local_len(r, n) for r in range(world)]
        assert max(lens) - min(lens) <= block                                  # balanced at any length
        for r, p in enumerate(parts):
            assert torch.all(p[1:] > p[:-1])                                   # slice order == position order
            for i in (0, len(p) // 2, len(p) - 1):
                if len(p):
                    pos = int(p[i])
                    assert plan.owner(pos) == r and plan.local_index(pos) == i
        # appending position n goes to the end of its owner's slice
        assert plan.local_index(n) == plan.local_len(plan.owner(n), n)
        assert plan.capacity(n) == max(lens + [0])


def _attn(q, k, v, scale):
    s = (q @ k.T) * scale
    m = s.max(dim=1, keepdim=True).values
    p = torch.exp2((s - m) * 1.4426950408889634)
    l = p.sum(dim=1, keepdim=True)
    return (p @ v) / l, (m * 1.4426950408889634 + torch.log2(l)).squeeze(1)  # normalised O, log2-domain log-sum-exp


def test_slice_attention_plus_lse_merge_equals_full_attention():
    g = torch.Generator().manual_seed(0)
    n, d, rows, world = 700, 128, 4, 3
    q = torch.randn(rows, d, generator=g, dtype=torch.float64)
    k = torch.randn(n, d, generator=g, dtype=torch.float64)
    v = torch.randn(n, d, generator=g, dtype=torch.float64)
    scale = d ** -0.5
    ref, lse_ref = _attn(q, k, v, scale)
    plan = SeqShardPlan(world, 64)
    os_, ls_ = [], []
    for idx in plan.split(n) + [torch.zeros(0, dtype=torch.long)]:  # plus an empty slice
        if len(idx) == 0:
            os_.append(torch.zeros(rows, d, dtype=torch.float64))
            ls_.append(torch.full((rows,), -np.inf, dtype=torch.float64))
            continue
        o, l = _attn(q, k[idx], v[idx], scale)
        os_.append(o)
        ls_.append(l)
    O, L = torch.stack(os_), torch.stack(ls_)
    mx = L.max(dim=0).values
    w = torch.exp2(L - mx)                                  # duo_merge_partials: 2^(lse_p - max)
    merged = (w[:, :, None] * O).sum(0) / w.sum(0)[:, None]
    torch.testing.assert_close(merged, ref, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(mx + torch.log2(w.sum(0)), lse_ref, rtol=1e-12, atol=1e-12)
