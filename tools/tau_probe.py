"""Developer probe: how large can the bf16 best-vs-runner-up gap be when the bf16 key clip is WRONG (TVR / ANet shape)?
The ambiguity threshold tau must exceed that."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.load_package()
import bench
from dkd_b200 import engine, ops
from dkd_b200.model import DLDKD
dev = torch.device("cuda")
SH = {"tvr": dict(bench.TVR), "anet": dict(Nv=4885, L=128, Dv=1024, Nq=17031, Lq=30, Dq=1024, H=384, T=32),
      "charades": dict(Nv=1334, L=128, Dv=1024, Nq=3720, Lq=30, Dq=768, H=384, T=32)}
for name in sys.argv[1:] or ["tvr"]:
    shape = SH[name]
    model, frames, mask, qs = bench.synth_encoded(shape, dev, 0, DLDKD)
    pc = engine.prepare_corpus(frames, mask, [tuple(t.detach() for t in p) for p in model.attention_params()],
                               T=32, heads=("two_scale",), precisions=("exact", "bf16", "fp16"))
    pq = engine.prepare_queries([q.contiguous() for q in qs])
    for b, bd in enumerate(pc.branches):
        s_ex, k_ex = ops.clip_score_f32(pq.qn[b], bd.clip_planes, bd.prop_scale)
        for prec, q, prop in (("bf16", pq.qb[b], bd.prop_b), ("fp16", pq.qh[b], bd.prop_h)):
            s, k, gap = ops.score_max_bf16(q, pq.M, prop, pc.Nv, pc.P, want_gap=True)
            wrong = k != k_ex
            gw = gap[wrong]
            err = (s - s_ex).abs()
            qs_ = torch.quantile(gw.float()[:10_000_000], torch.tensor([0.5, 0.99, 0.9999], device=dev)) if gw.numel() else None
            print(f"{name} branch {b} {prec}: wrong key {int(wrong.sum())} of {wrong.numel()} ({100 * wrong.float().mean():.2f} %), "
                  f"gap of wrong pairs max {float(gw.max()) if gw.numel() else 0:.2e} q50/q99/q9999 {[f'{float(x):.1e}' for x in qs_] if qs_ is not None else None}; "
                  f"score err max {float(err.max()):.2e}; pairs with gap < 1e-3: {100 * (gap < 1e-3).float().mean():.2f} %, "
                  f"< 5e-4: {100 * (gap < 5e-4).float().mean():.2f} %, < 1.25e-4: {100 * (gap < 1.25e-4).float().mean():.2f} %")
    del pc, pq, frames, qs
    torch.cuda.empty_cache()
