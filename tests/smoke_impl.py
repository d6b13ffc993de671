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
    ref_sorted = np.take_along_axis(fused_ref, top_ref, 1)
    assert np.abs(s_ex.cpu().numpy() - ref_sorted).max() <= 5e-6, "exact path: fused scores off"
    same = (i_ex.cpu().numpy() == top_ref)
    assert same.mean() > 0.995, f"exact path: ranked ids differ from the oracle ({same.mean():.4f})"
    assert torch.equal(i_bf, i_ex) and torch.equal(s_bf, s_ex), "bf16+rescoring differs from the exact path"
    fused_bf, _ = engine.score_two_scale_head(pc, pq, "bf16")
    assert np.abs(fused_bf.cpu().numpy() - fused_ref).max() <= 1e-3, "bf16 fused scores beyond 1e-3"
    print("smoke ok: two-scale rank on cuda:0 matches the oracle "
          f"(max |d| exact {np.abs(s_ex.cpu().numpy() - ref_sorted).max():.2e}, "
          f"bf16 {np.abs(fused_bf.cpu().numpy() - fused_ref).max():.2e})")
