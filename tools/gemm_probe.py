"""Developer probe: A/B the scoring GEMM variants inside ONE process (same box, same clocks)."""
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
Nq, Nv, D = 10895, 2179, 384
dev="cuda"; torch.manual_seed(0)
configs = sys.argv[1:] or ["cta=2,kbs=1", "cta=1,kbs=1"]
for R in (528, 128):
    q = torch.randn(Nq, D, device=dev); x = torch.randn(Nv*R, D, device=dev)
    Mpad = ops.round_up(Nq,256)
    _, qb = ops.normalize_rows(q, False, True, rows_pad=Mpad)
    _, xb = ops.normalize_rows(x, False, True)
    om = torch.empty(Nq, Nv, device=dev); oa = torch.empty(Nq, Nv, dtype=torch.int32, device=dev); og = torch.empty(Nq, Nv, device=dev)
    for rep in range(2):
        for cfg in configs:
            kv = dict(x.split("=") for x in cfg.split(","))
            os.environ["DKD_GEMM_CTA"] = kv.get("cta", "2"); os.environ["DKD_GEMM_KBS"] = kv.get("kbs", "1"); os.environ["DKD_GEMM_DEBUG"] = kv.get("dbg", "0")
            try:
                t = timeit(lambda: ops.score_max_bf16(qb, Nq, xb, Nv, R, None, om, oa))
                print(f"R={R} {cfg:24s}: {t:.3f} ms  {2.0*Nq*Nv*R*D/t/1e9:.1f} TFLOP/s", flush=True)
            except Exception as e:
                print(f"R={R} {cfg}: {e}")
