#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU): launch list of the step + --set full captures.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'score_max_bf16_kernel|exact_umma_kernel|frame_fuse_h_kernel|frame_fuse_csr_kernel|select_topk_kernel|build_proposals_kernel|frame_attn_table_v2_kernel' \
    -c 17 -o gpurun_out/r2_prof python bench.py --profile --steps 1 --warmup 0 > gpurun_out/r2_prof.log 2>&1
ncu -i gpurun_out/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'exact_umma_kernel|mha_small_kernel|layernorm_rows_kernel|row_stats_kernel' \
    -c 14 -o gpurun_out/r2_enc python tools/encoder_probe.py > gpurun_out/r2_enc.log 2>&1
ncu -i gpurun_out/r2_enc.ncu-rep --page raw --csv > gpurun_out/r2_enc_raw.csv 2>/dev/null
rm -f gpurun_out/r2_enc.ncu-rep
ls -la gpurun_out | tail -12
