"""In-process, interleaved A/B of builds of the scoring GEMM (tools/build_variant.sh): every library is loaded into
ONE process and the launches alternate, so box-to-box and run-to-run clock drift cancels.

    python tools/ab_gemm_inproc.py [--shape two_scale|frame] [--rounds 12] name=path.so ...

Prints the median / min launch time per build (TVR shape: 10,895 queries x 2,179 videos, D = 384; R = 528 proposals with
the ambiguous-pair lists, or R = 128 frames with a mask)."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dkd_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="two_scale")
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--no-mask", action="store_true", help="frame shape: pass no mask (the unmasked kernel)")
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    M, Nv, D = 10895, 2179, 384
    R = 528 if a.shape == "two_scale" else 128
    Mpad = (M + 255) // 256 * 256
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn(Mpad, D, device="cuda", generator=g), dim=-1).bfloat16()
    x = torch.nn.functional.normalize(torch.randn(Nv * R, D, device="cuda", generator=g), dim=-1).bfloat16()
    om = torch.empty((M, Nv), dtype=torch.float32, device="cuda")
    oa = torch.empty((M, Nv), dtype=torch.int32, device="cuda")
    cnt = torch.zeros((Nv,), dtype=torch.int32, device="cuda")
    lst = torch.empty((Nv, M), dtype=torch.int32, device="cuda")
    mask = torch.ones((Nv, R), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    P = ctypes.c_void_p
    handles = {}
    for spec in a.libs:
        name, path = spec.split("=", 1)
        lib = ctypes.CDLL(os.path.abspath(path))
        for fn in ("dkd_score_max_bf16_lists", "dkd_score_max_bf16"):
            getattr(lib, fn).argtypes = _lib.PROTOTYPES[fn]
            getattr(lib, fn).restype = ctypes.c_int32
        handles[name] = lib

    def launch(lib):
        if a.shape == "two_scale":
            cnt.zero_()
            rc = lib.dkd_score_max_bf16_lists(P(q.data_ptr()), M, Mpad, P(x.data_ptr()), Nv, R, D, None, P(om.data_ptr()),
                                              P(oa.data_ptr()), Nv, 1e-3, P(cnt.data_ptr()), P(lst.data_ptr()), M, 0, P(st))
        else:
            rc = lib.dkd_score_max_bf16(P(q.data_ptr()), M, Mpad, P(x.data_ptr()), Nv, R, D,
                                        None if a.no_mask else P(mask.data_ptr()),
                                        P(om.data_ptr()), P(oa.data_ptr()), None, Nv, None, 0.0, P(st))
        assert rc == 0, rc

    ref = None
    times = {n: [] for n in handles}
    for n, lib in handles.items():            # warm-up + equality of the results across builds
        for _ in range(3):
            launch(lib)
        torch.cuda.synchronize()
        cur = (om.clone(), oa.clone())
        if ref is None:
            ref = cur
        else:
            assert torch.equal(ref[0], cur[0]) and torch.equal(ref[1], cur[1]), f"{n}: results differ"
    for _ in range(a.rounds):
        for n, lib in handles.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                launch(lib)
            e1.record()
            torch.cuda.synchronize()
            times[n].append(e0.elapsed_time(e1) / 4)
    flops = 2.0 * M * Nv * R * D
    for n, t in times.items():
        med = float(np.median(t))
        print(f"{n:10s} median {med:7.3f} ms  min {min(t):7.3f}  max {max(t):7.3f}  -> {flops / med / 1e9:7.1f} TFLOP/s (median)")


if __name__ == "__main__":
    main()
