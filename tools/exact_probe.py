"""Developer probe: the exact clip-scale kernel (dense, TVR shape) alone, for ncu."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.load_package()
from dkd_b200 import ops
Nq, Nv, L, D = int(os.environ.get("NQ", 10895)), int(os.environ.get("NV", 2179)), 128, 384
dev = "cuda"
torch.manual_seed(0)
frames = torch.randn(Nv, L, D, device=dev) + 0.6 * torch.randn(Nv, 1, D, device=dev)
lengths = torch.full((Nv,), L, dtype=torch.int32, device=dev)
q = torch.randn(Nq, D, device=dev)
qn, _ = ops.normalize_rows(q)
clips = ops.downsample_clips(frames, lengths)
_, ps, _ = ops.build_proposals(clips, want_bf16=False)
planes = ops.pack_clips(clips)
for _ in range(2):
    ops.clip_score_f32(qn, planes, ps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.clip_score_f32(qn, planes, ps)
e1.record(); torch.cuda.synchronize()
print(f"clip_score_f32 dense {e0.elapsed_time(e1) / 3:.3f} ms for {Nq} x {Nv} pairs")
