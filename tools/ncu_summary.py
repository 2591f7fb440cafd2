#!/usr/bin/env python3
"""Turn raw Nsight Compute output into the small CSV summaries kept under profiles/.

  launches : ncu --metrics gpu__time_duration.sum --csv --log-file L.csv ...   ->  per-kernel time shares
  kernel   : ncu -i X.ncu-rep --page raw --csv > R.csv                         ->  selected metrics of launch N

usage: ncu_summary.py launches L.csv "header comment" > profiles/rNN_launches_summary.csv
       ncu_summary.py kernel R.csv "header comment" [launch-id] > profiles/rNN_<kernel>_ncu.csv
"""
import csv, sys, collections, re

KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__inst_executed.sum",
    "launch__local_size_per_thread" if False else "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]


def rows_after_header(path):
    with open(path, newline="") as f:
        lines = f.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.reader(lines[start:]))


def launches(path, comment):
    rows = rows_after_header(path)
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    t, n = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        name = re.sub(r"^void ", "", r[ki])
        name = re.sub(r"\(.*$", "", name)[:90]
        t[name] += v
        n[name] += 1
    tot = sum(t.values())
    OURS = ("fq3", "pf::", "fe::", "to_channels_last_kernel", "conv_out_kernel", "conv_gemm", "set_state_kernel",
            "get_hidden_kernel", "sample_kernel", "pack_kernel", "pack_mma_kernel", "mtp_table_kernel", "cast_strided")
    ours = sum(v for k, v in t.items() if any(s in k for s in OURS))
    print(f"# {comment}")
    print(f"# {sum(n.values())} launches, {tot / 1e3:.1f} ms total device time; hand-written kernels {100 * ours / tot:.1f}% of device time.")
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
    print("time_us,share_pct,launches,kernel")
    for k, v in t.most_common(24):
        print(f"{v:.1f},{100 * v / tot:.2f},{n[k]},{k}")


def kernel(path, comment, which=0):
    rows = rows_after_header(path)
    hdr, units = rows[0], rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    r = data[which]
    print(f"# {comment}")
    print("launch,metric,unit,value")
    for want in KEEP:
        for i, h in enumerate(hdr):
            if h == want or h.endswith("." + want):
                print(f"{which},{want},{units[i]},{r[i].replace(',', '') if h != 'Kernel Name' else r[i]}")
                break


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
