"""Developer probe: the key-clip attention table kernel alone at TVR shape (for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.load_package()
from dkd_b200 import ops
Nv, L, D = int(os.environ.get("NV", 2179)), 128, 384
dev = "cuda"
torch.manual_seed(0)
frames = torch.randn(Nv, L, D, device=dev)
lengths = torch.full((Nv,), L, dtype=torch.int32, device=dev)
clips = ops.downsample_clips(frames, lengths)
key = torch.randn(Nv, L, D, device=dev) * 0.3
val = torch.randn(Nv, L, D, device=dev)
for _ in range(3):
    tf, tb = ops.frame_attn_table(key, val, clips, lengths)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.frame_attn_table(key, val, clips, lengths)
e1.record(); torch.cuda.synchronize()
print(f"frame_attn_table (dots + table) {e0.elapsed_time(e1) / 5:.3f} ms")
