"""Time DLDKD.encode_context on the TVR-shaped corpus: PyTorch/cuBLAS mirror vs the fused encoder kernels."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dkd_b200.model import DLDKD  # noqa: E402

shape = dict(bench.TVR)
dev = torch.device("cuda")
cfg, opt = bench.model_config(shape)
torch.manual_seed(0)
model = DLDKD(cfg, opt).to(dev).eval()
g = torch.Generator(device=dev).manual_seed(1)
B = 200
x = torch.randn(B, shape["L"], shape["Dv"], device=dev, generator=g)
x = x / (x.norm(dim=-1, keepdim=True) + 1e-5)
mask = torch.ones(B, shape["L"], device=dev)
nb = (shape["Nv"] + B - 1) // B


def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        model.encode_context(x, mask)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            a, b = model.encode_context(x, mask)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, a, b


t_ref, ra, rb = run(nb)
model.enable_fused_encoder()
t_fused, fa, fb = run(nb)
flops = 2.0 * B * shape["L"] * (shape["Dv"] * 384 + 5 * 384 * 384) * 2 + 2.0 * 2 * 2 * B * 128 * 128 * 384
print(json.dumps({"batch_videos": B, "ms_per_batch_pytorch": t_ref, "ms_per_batch_fused": t_fused,
                  "tvr_corpus_ms_pytorch": t_ref * nb, "tvr_corpus_ms_fused": t_fused * nb,
                  "fp32_grade_tflops_fused": flops / (t_fused * 1e-3) / 1e12,
                  "max_abs_diff": [float((fa - ra).abs().max()), float((fb - rb).abs().max())]}))
