#!/bin/bash
# A/B of library builds on the frame-head workload: bash tools/ab_frame.sh <variant names...>  (default build runs last)
mkdir -p gpurun_out
run() {
  timeout 200 $2 --workload tvr_frame --steps 20 --warmup 3 --no-reference-leg --no-cpu-baseline --no-encoder --no-variants --no-eval-epoch --no-strong --no-c4 > gpurun_out/abf_$1.json 2> gpurun_out/abf_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/abf_$1.json').read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), d['parity']['top100_ids_identical_to_exact_fp32'], round(d['roofline']['avg_launch_ms'],3), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], {k:v['ms_per_step'] for k,v in d['kernels_ms'].items()})"
}
for v in "$@"; do run $v "python tools/ab_bench.py dl-dkd_b200/variants/libdkd_b200_$v.so"; done
run default "python bench.py"
