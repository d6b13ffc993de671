import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g; g.load_package()
from dkd_b200 import ops
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
Nq, Nv, D, P = 10895, 2179, 384, 528
dev="cuda"; torch.manual_seed(0)
q = torch.randn(Nq, D, device=dev); x = torch.randn(Nv*P, D, device=dev)
Mpad = ops.round_up(Nq,128)
_, qb = ops.normalize_rows(q, False, True, rows_pad=Mpad)
_, xb = ops.normalize_rows(x, False, True)
om = torch.empty(Nq, Nv, device=dev); oa = torch.empty(Nq, Nv, dtype=torch.int32, device=dev)
t = timeit(lambda: ops.score_max_bf16(qb, Nq, xb, Nv, P, None, om, oa))
print(f"DKD_GEMM_DEBUG={os.environ.get('DKD_GEMM_DEBUG','0')}: {t:.3f} ms  {2.0*Nq*Nv*P*D/t/1e9:.1f} TFLOP/s")
