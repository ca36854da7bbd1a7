"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
    python tools/summarize_ncu.py full     gpurun_out/prof_conv.ncu-rep     > profiles/rNN_conv_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "launch__cluster_dim_x", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time (ncu, serialised, cold cache)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} n={v[0]:5d} total={v[1] / 1e6:9.3f} ms share={v[1] / tot * 100:5.1f}% avg={v[1] / v[0] / 1e3:8.1f} us")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEEP]
    for r in rows[2:]:
        print("----")
        for i in idx:
            print(f"{hdr[i]:70s} {r[i]:>24s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
