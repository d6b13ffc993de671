import os, sys, time, threading, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g; g.load_package()
from dkd_b200 import ops
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
class S(threading.Thread):
    def __init__(s): super().__init__(daemon=True); s.stop=False; s.clk=[]; s.pw=[]
    def run(s):
        while not s.stop:
            s.clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); s.pw.append(pynvml.nvmlDeviceGetPowerUsage(h)/1000.0); time.sleep(0.01)
Nq, Nv, D, R = 10895, 2179, 384, 528
dev="cuda"; torch.manual_seed(0)
q = torch.randn(Nq, D, device=dev); x = torch.randn(Nv*R, D, device=dev)
Mpad = ops.round_up(Nq,256)
_, qb = ops.normalize_rows(q, False, True, rows_pad=Mpad); _, xb = ops.normalize_rows(x, False, True)
om = torch.empty(Nq, Nv, device=dev); oa = torch.empty(Nq, Nv, dtype=torch.int32, device=dev)
for cfg in sys.argv[1:]:
    kv = dict(x.split("=") for x in cfg.split(","))
    os.environ["DKD_GEMM_CTA"] = kv.get("cta", "2"); os.environ["DKD_GEMM_KBS"] = kv.get("kbs", "1"); os.environ["DKD_GEMM_DEBUG"] = kv.get("dbg", "0")
    for _ in range(3): ops.score_max_bf16(qb, Nq, xb, Nv, R, None, om, oa)
    torch.cuda.synchronize(); s = S(); s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 150
    for _ in range(n): ops.score_max_bf16(qb, Nq, xb, Nv, R, None, om, oa)
    e1.record(); torch.cuda.synchronize(); s.stop=True; s.join()
    t = e0.elapsed_time(e1)/n
    print(f"{cfg:22s}: {t:.3f} ms {2.0*Nq*Nv*R*D/t/1e9:.0f} TF  clk median {np.median(s.clk):.0f} min {min(s.clk)} MHz  power median {np.median(s.pw):.0f} max {max(s.pw):.0f} W  (n={len(s.clk)})", flush=True)
    time.sleep(1.0)
