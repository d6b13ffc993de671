import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g; g.load_package()
from dkd_b200 import engine, ops
from dkd_b200.model import DLDKD
import bench
from tests import synth
dev = torch.device("cuda")

def stats(pc, pq, tag, K=100):
    f_ex, per_ex = engine.score_two_scale_head(pc, pq, "exact")
    for tau in (0.0, 2e-4, 5e-4, 1e-3, 2e-3):
        f_bf, per_bf = engine.score_two_scale_head(pc, pq, "bf16", tau=tau)
        d = (f_ex - f_bf).abs()
        fl = [(per_ex[b]["key_clip"] != per_bf[b]["key_clip"]).float().mean().item() for b in range(2)]
        _, _, gap = ops.score_max_bf16(pq.qb[0], pq.M, pc.branches[0].prop_b, pc.Nv, pc.P, want_gap=True)
        print(tag, f"tau={tau:g}: flagged {(gap < tau).float().mean().item():.4f}  fused |d| max {d.max().item():.2e} p99.99 {torch.quantile(d.flatten()[:10_000_000], 0.9999).item():.2e}  flips {fl[0]:.5f} {fl[1]:.5f}")
    f_bf, per_bf = engine.score_two_scale_head(pc, pq, "bf16")
    d = (f_ex - f_bf).abs()
    print(tag, "fused |d| max %.2e mean %.2e" % (d.max().item(), d.mean().item()))
    for b in range(len(per_ex)):
        dc = (per_ex[b]["clip"] - per_bf[b]["clip"]).abs()
        flips = (per_ex[b]["key_clip"] != per_bf[b]["key_clip"]).float().mean().item()
        print(tag, " branch", b, "clip |d| max %.2e  key-clip flips %.3f" % (dc.max().item(), flips))
    # containment: rank (in approx order) of each exact top-K element
    s_ex, i_ex = ops.topk(f_ex, K)
    M, Nv = f_ex.shape
    # approx rank of item = number of items with approx score greater
    appr = f_bf
    got = torch.gather(appr, 1, i_ex.long())
    rk = (appr.unsqueeze(1) > got.unsqueeze(2)).sum(-1) if M * K * Nv < 2e8 else None
    if rk is not None:
        print(tag, " max approx-rank of an exact top-%d item: %d ; 99.9pct %d" % (K, rk.max().item(), int(torch.quantile(rk.float().flatten(), 0.999).item())))
    gaps = (s_ex[:, :-1] - s_ex[:, 1:])
    print(tag, " exact top-K adjacent gap median %.2e ; score range %.3f..%.3f; gap(100th - 128th) median %.2e" % (
        gaps.median().item(), f_ex.min().item(), f_ex.max().item(),
        (ops.topk(f_ex, 128)[0][:, 99] - ops.topk(f_ex, 128)[0][:, 127]).median().item()))
    for Kc in (128, 160, 256):
        s_bf, i_bf = engine.rank(pc, pq, K=K, head="two_scale", precision="bf16", Kc=Kc)
        same_q = (i_bf == i_ex).all(1).float().mean().item()
        print(tag, f" Kc={Kc}: queries with identical top-{K}: {same_q:.4f}")

# (a) realistic: random-init encoders, reduced TVR
shape = dict(bench.TVR); shape["Nv"] = 700; shape["Nq"] = 1000
model, frames, mask, qs = bench.synth_encoded(shape, dev, 0, DLDKD)
pc = engine.prepare_corpus(frames, mask, [tuple(t.detach() for t in p) for p in model.attention_params()])
pq = engine.prepare_queries(qs)
stats(pc, pq, "[random-init encoders]")
# (b) synthetic test data
Nv, L, D, M = 700, 128, 384, 300
fr, mk, _ = synth.encoded_corpus(Nv, L, D, seed=7, shared=1.5)
fr2, _, _ = synth.encoded_corpus(Nv, L, D, seed=8, shared=1.5); fr2 = fr2 * mk[:, :, None]
gen = torch.Generator().manual_seed(9)
params = [(0.05 * torch.randn(D, D, generator=gen), torch.zeros(D), 0.05 * torch.randn(D, D, generator=gen), torch.zeros(D)) for _ in range(2)]
qq = [synth.encoded_queries(M, D, seed=10), synth.encoded_queries(M, D, seed=11)]
pc = engine.prepare_corpus([fr.to(dev), fr2.to(dev)], mk.to(dev), [tuple(t.to(dev) for t in p) for p in params])
pq = engine.prepare_queries([q.to(dev) for q in qq])
stats(pc, pq, "[synthetic]")
