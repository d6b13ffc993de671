#!/bin/bash
# A/B of library builds inside the whole step: bash tools/ab_run.sh <variant names...>  (the default build runs last as "default")
mkdir -p gpurun_out
run() {  # name, launcher
  timeout 300 $2 --steps 10 --warmup 3 --no-strong --no-c4 --no-reference-leg --no-cpu-baseline --no-encoder --no-variants --no-eval-epoch > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$1.json').read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), d['parity']['top100_ids_identical_to_exact_fp32'], {k:(v['calls_per_step'], v['ms_per_step']) for k,v in d['kernels_ms'].items()})"
}
for v in "$@"; do run $v "python tools/ab_bench.py dl-dkd_b200/variants/libdkd_b200_$v.so"; done
run default "python bench.py"
