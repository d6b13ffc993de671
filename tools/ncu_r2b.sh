#!/bin/bash
# Round-2 (second session) ncu evidence for the frame head (pair-tile GEMM): launch list + one --set full capture.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_frame.csv \
    python bench.py --workload tvr_frame --profile --steps 2 --warmup 1 > gpurun_out/r2b_launches_frame.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'score_max_bf16_kernel|exact_umma_kernel' \
    -c 4 -o gpurun_out/r2b_frame python bench.py --workload tvr_frame --profile --steps 1 --warmup 0 > gpurun_out/r2b_frame.log 2>&1
ncu -i gpurun_out/r2b_frame.ncu-rep --page raw --csv > gpurun_out/r2b_frame_raw.csv 2>/dev/null
rm -f gpurun_out/r2b_frame.ncu-rep
tail -3 gpurun_out/r2b_frame.log
