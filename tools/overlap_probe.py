"""Developer probe: does the frame-scale gather co-run with the persistent scoring GEMM (2 streams)?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g; g.load_package()
from dkd_b200 import ops
Nq, Nv, D, P = 10895, 2179, 384, 528
dev = "cuda"; torch.manual_seed(0)
q = torch.randn(Nq, D, device=dev); x = torch.randn(Nv * P, D, device=dev)
Mpad = ops.round_up(Nq, 256)
_, qb = ops.normalize_rows(q, False, True, rows_pad=Mpad)
_, xb = ops.normalize_rows(x, False, True)
tb = xb.view(Nv, P, D)
om = torch.empty(Nq, Nv, device=dev); oa = torch.empty(Nq, Nv, dtype=torch.int32, device=dev); og = torch.empty(Nq, Nv, device=dev)
om2 = torch.rand(Nq, Nv, device=dev); oa2 = torch.randint(0, P, (Nq, Nv), dtype=torch.int32, device=dev)
fused = torch.empty(Nq, Nv, device=dev)
sA = torch.cuda.Stream(priority=-1); sB = torch.cuda.Stream(priority=0)
def gemm(): ops.score_max_bf16(qb, Nq, xb, Nv, P, None, om, oa)
def fuse(): ops.frame_fuse(qb[:Nq], tb, om2, oa2, 0.7, 0.3, 0.7, fused=fused, accumulate=False)
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
print(f"gemm alone   {timeit(gemm):.3f} ms")
print(f"fuse alone   {timeit(fuse):.3f} ms")
def both(order):
    cur = torch.cuda.current_stream()
    sA.wait_stream(cur); sB.wait_stream(cur)
    if order == "gemm_first":
        with torch.cuda.stream(sA): gemm()
        with torch.cuda.stream(sB): fuse()
    else:
        with torch.cuda.stream(sB): fuse()
        with torch.cuda.stream(sA): gemm()
    cur.wait_stream(sA); cur.wait_stream(sB)
for order in ("gemm_first", "fuse_first"):
    print(f"both ({order}) {timeit(lambda: both(order)):.3f} ms")
