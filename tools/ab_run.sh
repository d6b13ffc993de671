mkdir -p gpurun_out
for v in a b d e; do
  timeout 300 python tools/ab_bench.py dl-dkd_b200/variants/libdkd_b200_$v.so --steps 10 --warmup 3 --no-strong --no-c4 --no-reference-leg --no-cpu-baseline --no-encoder --no-variants > gpurun_out/r2f_ab_$v.json 2> gpurun_out/r2f_ab_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2f_ab_$v.json').read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],3), {k:(v['calls_per_step'], v['ms_per_step']) for k,v in d['kernels_ms'].items()})"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-strong --no-c4 --no-reference-leg --no-cpu-baseline --no-encoder --no-variants > gpurun_out/r2f_ab_c.json 2> gpurun_out/r2f_ab_c.err
python -c "
import json; d=json.loads(open('gpurun_out/r2f_ab_c.json').read().strip().splitlines()[-1]); print('c', round(d['ms_per_step'],3), {k:(v['calls_per_step'], v['ms_per_step']) for k,v in d['kernels_ms'].items()})"
