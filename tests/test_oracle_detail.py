"""CPU checks of the full-size parity checker itself (oracle.two_scale_eval_detail / compare_ranking): the batched
evaluator must equal the plain per-branch formulas, and the ranking comparison must accept exactly what it documents."""
import numpy as np
import torch

from oracle import oracle as O
from tests import synth


def _setup(Nv=9, L=64, D=64, M=23, T=32):
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=5)
    frames2 = synth.encoded_corpus(Nv, L, D, seed=6)[0] * mask[:, :, None]
    g = torch.Generator().manual_seed(7)
    params = [(0.05 * torch.randn(D, D, generator=g), 0.01 * torch.randn(D, generator=g),
               0.05 * torch.randn(D, D, generator=g), torch.zeros(D)) for _ in range(2)]
    qs = [synth.encoded_queries(M, D, seed=8), synth.encoded_queries(M, D, seed=9)]
    return [frames, frames2], mask, params, qs


def test_eval_detail_equals_plain_formulas():
    frames, mask, params, qs = _setup()
    props, keys, vals = O.two_scale_corpus(frames, mask, params)
    det = O.two_scale_eval_detail(qs, props, keys, vals, mask, bsz=7)
    br = [O.two_scale_branch(q, f, mask, *p) for q, f, p in zip(qs, frames, params)]
    fused = O.fuse_branches(br[0]["branch"].numpy(), br[1]["branch"].numpy())
    assert np.abs(det["fused"] - fused).max() <= 1e-6
    for b in range(2):
        assert np.array_equal(det["key_clip"][b], br[b]["key_clip"].numpy().astype(np.int32))
        assert np.abs(det["clip"][b] - br[b]["clip"].numpy()).max() <= 1e-6
    # the batched CPU-baseline loop computes the same matrix
    fused2, order = O.cpu_eval_two_scale(qs, props, keys, vals, mask, bsz=5, K=5)
    assert np.abs(fused2 - fused).max() <= 1e-6 and order.shape == (qs[0].shape[0], 5)
    assert det["tie"].shape == fused.shape and det["tie"].dtype == bool


def test_compare_ranking_rules():
    rng = np.random.default_rng(0)
    fused = rng.standard_normal((6, 40)).astype(np.float32)
    tie = np.zeros_like(fused, dtype=bool)
    top = O.topk_ids(fused, 10)
    assert O.compare_ranking(fused, tie, top, 10) == dict(queries=6, queries_identical=6, swaps=0, tie_pairs_in_lists=0,
                                                          mismatches=0)
    # a swap of two scores further apart than the tolerance is a mismatch ...
    bad = top.copy()
    bad[2, [3, 4]] = bad[2, [4, 3]]
    assert O.compare_ranking(fused, tie, bad, 10)["mismatches"] == 2
    # ... unless the two scores are within fp32 noise of each other
    f2 = fused.copy()
    f2[2, top[2, 4]] = f2[2, top[2, 3]] - 1e-6
    r = O.compare_ranking(f2, tie, bad, 10)
    assert r["mismatches"] == 0 and r["swaps"] == 2
    # a tie pair may sit anywhere in the device list: it is dropped from both sides
    t2 = tie.copy()
    v = top[4, 0]
    t2[4, v] = True
    moved = np.concatenate([top[4, 1:], [v]])[None]
    r = O.compare_ranking(fused[4:5], t2[4:5], moved, 10)
    assert r["mismatches"] == 0 and r["tie_pairs_in_lists"] == 1
    assert O.recall_counts([1, 3, 7, 200]) == [1, 2, 3, 3]
