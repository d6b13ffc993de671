"""Developer probe: time each kernel of the scoring path at TVR shape (not the bench contract)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.load_package()
from dkd_b200 import ops

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

Nq, Nv, L, D, T = 10895, 2179, 128, 384, 32
P = 528
dev = "cuda"
torch.manual_seed(0)
frames = torch.randn(Nv, L, D, device=dev) + 0.6 * torch.randn(Nv, 1, D, device=dev)
lengths = torch.full((Nv,), L, dtype=torch.int32, device=dev)
q = torch.randn(Nq, D, device=dev)
Mpad = ops.round_up(Nq, 128)
qn, qb, qh = ops.normalize_rows(q, True, True, rows_pad=Mpad, want_f16=True)
clips = ops.downsample_clips(frames, lengths)
t = timeit(lambda: ops.downsample_clips(frames, lengths)); print(f"downsample_clips {t:.3f} ms")
pb, ps, _ = ops.build_proposals(clips)
t = timeit(lambda: ops.build_proposals(clips)); print(f"build_proposals {t:.3f} ms  write {Nv*P*D*2/t/1e6:.1f} GB/s")
key = torch.randn(Nv, L, D, device=dev) * 0.3; val = torch.randn(Nv, L, D, device=dev)
tf, tb = ops.frame_attn_table(key, val, clips, lengths)
t = timeit(lambda: ops.frame_attn_table(key, val, clips, lengths), iters=2, warm=1); print(f"frame_attn_table {t:.3f} ms")
om = torch.empty(Nq, Nv, device=dev); oa = torch.empty(Nq, Nv, dtype=torch.int32, device=dev)
fn = lambda: ops.score_max_bf16(qb, Nq, pb.view(-1, D), Nv, P, None, om, oa)
t = timeit(fn, iters=10, warm=3)
fl = 2.0 * Mpad * Nv * P * D
print(f"score_max_bf16 P=528: {t:.3f} ms  {fl/t/1e9:.1f} TFLOP/s (padded M) ; algorithmic {2.0*Nq*Nv*P*D/t/1e9:.1f}")
_, fb = ops.normalize_rows(frames, False, True)
fn2 = lambda: ops.score_max_bf16(qb, Nq, fb.view(-1, D), Nv, L, None, om, oa)
t2 = timeit(fn2, iters=10, warm=3)
print(f"score_max_bf16 R=128: {t2:.3f} ms  {2.0*Nq*Nv*L*D/t2/1e9:.1f} TFLOP/s")
fn(); 
fused = torch.empty(Nq, Nv, device=dev)
t = timeit(lambda: ops.frame_fuse(qh[:Nq], tb, om, oa, 0.7, 0.3, 0.7, fused=fused, accumulate=False)); print(f"frame_fuse fp16 {t:.3f} ms  gather {Nq*Nv*D*2/t/1e6:.1f} GB/s")
t = timeit(lambda: ops.frame_fuse(qn[:Nq], tf, om, oa, 0.7, 0.3, 0.7, fused=fused, accumulate=False)); print(f"frame_fuse f32 {t:.3f} ms")
t = timeit(lambda: ops.topk(fused, 128)); print(f"topk128 {t:.3f} ms")
t = timeit(lambda: ops.topk(fused, 100)); print(f"topk100 {t:.3f} ms")
qsp = qn[:Nq]; csp = ops.pack_clips(clips)
t = timeit(lambda: ops.clip_score_f32(qsp, csp, ps), iters=2, warm=1); print(f"clip_score_f32 dense {t:.3f} ms")
fn32, _ = ops.normalize_rows(frames)
t = timeit(lambda: ops.score_max_f32(qn[:Nq], fn32.view(Nv, L, D), None), iters=2, warm=1); print(f"score_max_f32 dense R=128 {t:.3f} ms")
ts, ti = ops.topk(fused, 128)
def resc():
    csr = ops.candidates_to_csr(ti, Nv)
    cs, ck = ops.clip_score_f32(qsp, csp, ps, csr=csr[:2])
    cand = torch.empty(Nq, 128, device=dev)
    ops.frame_fuse_csr(qn[:Nq], tf, cs, ck, csr, 0.7, 0.3, 0.7, cand, False)
    return ops.sort_candidates(cand, ti, 100)
t = timeit(resc, iters=3, warm=1); print(f"rescore one branch {t:.3f} ms")
