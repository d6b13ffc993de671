"""Summarise an ncu report (--page raw --csv export) into a markdown table for profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv "title" > profiles/rN_ncu_xxx.md
"""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__grid_size",
    "launch__block_size",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:48] for r in data]
    print(f"# {title}\n")
    print("| metric | unit | " + " | ".join(f"{i}: {n}" for i, n in enumerate(names)) + " |")
    print("|---|---|" + "---|" * len(names))
    for m in METRICS:
        if m not in idx:
            continue
        vals = []
        for r in data:
            v = r[idx[m]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {m} | {units[idx[m]]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
