"""Developer probe: how concentrated are the key clips of one video across queries (TVR shape)?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.load_package()
import bench
from dkd_b200 import engine
from dkd_b200.model import DLDKD
dev = torch.device("cuda")
shape = dict(bench.TVR)
model, frames, mask, qs = bench.synth_encoded(shape, dev, 0, DLDKD)
pc = engine.prepare_corpus(frames, mask, [tuple(t.detach() for t in p) for p in model.attention_params()],
                           T=32, heads=("two_scale",), precisions=("exact",))
pq = engine.prepare_queries([q.contiguous() for q in qs])
_, per = engine.score_two_scale_head(pc, pq, "exact")
for b, d in enumerate(per):
    k = d["key_clip"].long()                      # (M, Nv)
    M, Nv = k.shape
    hist = torch.zeros(Nv, 528, device=dev)
    hist.scatter_add_(1, k.t().contiguous(), torch.ones(Nv, M, device=dev))
    srt = hist.sort(dim=1, descending=True).values
    cum = srt.cumsum(1) / M
    print(f"branch {b}: distinct key clips per video: mean {float((hist > 0).sum(1).float().mean()):.1f}; "
          f"coverage of the top-1/4/8/16/32 proposals: " + ", ".join(f"{float(cum[:, j - 1].mean()):.3f}" for j in (1, 4, 8, 16, 32)))
    # consecutive-query locality: fraction of (m, n) whose key clip equals that of query m-1
    print("  same key clip as the previous query:", float((k[1:] == k[:-1]).float().mean()))
