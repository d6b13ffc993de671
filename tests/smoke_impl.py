"""__graft_entry__.smoke(): one small end-to-end pass of the hot path on cuda:0, checked against the oracle."""
import numpy as np
import torch

from oracle import oracle as O
from tests import synth


def run():
    assert torch.cuda.is_available(), "smoke() needs cuda:0"
    from dkd_b200 import engine, ops, _lib
    _lib.load()
    dev = torch.device("cuda:0")
    Nv, L, D, M, K = 64, 128, 384, 200, 100
    g = torch.Generator().manual_seed(0)
    frames, mask, lengths = synth.encoded_corpus(Nv, L, D, seed=1)
    frames2, _, _ = synth.encoded_corpus(Nv, L, D, seed=2)
    frames2 = frames2 * mask[:, :, None]
    qs = [synth.encoded_queries(M, D, seed=3), synth.encoded_queries(M, D, seed=4)]
    params = []
    for _ in range(2):
        params.append((0.05 * torch.randn(D, D, generator=g), torch.zeros(D), 0.05 * torch.randn(D, D, generator=g), torch.zeros(D)))
    # oracle (CPU, fp32)
    br = [O.two_scale_branch(q, f, mask, *p) for q, f, p in zip(qs, (frames, frames2), params)]
    fused_ref = O.fuse_branches(br[0]["branch"].numpy(), br[1]["branch"].numpy())
    top_ref = O.topk_ids(fused_ref, K)[:, :Nv]
    # device
    pc = engine.prepare_corpus([frames.to(dev), frames2.to(dev)], mask.to(dev),
                               [tuple(t.to(dev) for t in p) for p in params])
    pq = engine.prepare_queries([q.to(dev) for q in qs])
    s_ex, i_ex = engine.rank(pc, pq, K=Nv, head="two_scale", precision="exact")
    s_bf, i_bf = engine.rank(pc, pq, K=Nv, head="two_scale", precision="bf16", Kc=Nv)
    torch.cuda.synchronize()
    # Dense fused scores of the exact path vs the oracle.  The key clip is an argmax over 528 cosine
    # scores: where the oracle's own top-2 proposals are closer than fp32 summation noise (~1e-6) either
    # choice is "the" key clip and the frame-scale term differs — those pairs (a few per 10^4) are the
    # documented ties; everywhere else the fused score must agree to 5e-6.
    fused_ex, per = engine.score_two_scale_head(pc, pq, "exact")
    d_ex = np.abs(fused_ex.cpu().numpy() - fused_ref)
    tie = np.zeros_like(d_ex, dtype=bool)
    for b in range(2):
        allp = O.clip_scale_scores(qs[b], br[b]["proposals"])[1]            # (M, P, Nv)
        top2 = torch.topk(allp, 2, dim=1).values
        tie |= ((top2[:, 0] - top2[:, 1]) <= 2e-6).numpy()
        assert (per[b]["clip"].cpu() - br[b]["clip"]).abs().max() <= 2e-6, "exact path: clip-scale scores off"
        kk = per[b]["key_clip"].cpu().long()
        assert bool(((kk == br[b]["key_clip"]) | torch.from_numpy(tie)).all()), "exact path: key clip differs beyond ties"
    assert d_ex[~tie].max() <= 5e-6, "exact path: fused scores off"
    assert tie.mean() < 0.10
    # Ranking vs the oracle: a tie pair may sit anywhere in the device ranking (its frame-scale term follows
    # the other, equally valid, key clip), so tie pairs are dropped from both lists; what remains must be the
    # same sequence up to swaps of scores closer than the fp32 noise floor.
    got = i_ex.cpu().numpy()
    n_swapped = 0
    for m in range(M):
        a = [v for v in got[m] if not tie[m, v]]
        b = [v for v in top_ref[m] if not tie[m, v]]
        assert sorted(a) == sorted(b)
        for x, y in zip(a, b):
            if x != y:
                n_swapped += 1
                assert abs(fused_ref[m, x] - fused_ref[m, y]) <= 1e-5, "exact path: ranking differs from the oracle"
    assert torch.equal(i_bf, i_ex) and torch.equal(s_bf, s_ex), "bf16+rescoring differs from the exact path"
    # dense bf16-path scores: within the north_star tolerance of the oracle outside the tie pairs (there the
    # oracle's key clip is a coin flip), and within it of the device's own exact path everywhere (the bound the
    # ranking certificate relies on)
    fused_bf, _ = engine.score_two_scale_head(pc, pq, "bf16")
    d_bf = np.abs(fused_bf.cpu().numpy() - fused_ref)
    d_dev = float((fused_bf - fused_ex).abs().max())
    assert d_bf[~tie].max() <= 1e-3, "bf16 fused scores beyond 1e-3 of the oracle"
    assert d_dev <= engine.CERT_EPS, "bf16 fused scores beyond the certificate bound of the exact path"
    print("smoke ok: two-scale rank on cuda:0 matches the oracle "
          f"(max |d| exact {d_ex[~tie].max():.2e} outside {int(tie.sum())} fp32 key-clip ties, "
          f"{n_swapped} near-equal swaps in the ranking, bf16 vs oracle max {d_bf[~tie].max():.2e}, "
          f"bf16 vs exact path max {d_dev:.2e})")
