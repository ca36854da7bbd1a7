"""Target of the ncu pass over the streaming metric kernel: k_metrics_pairs on the BASELINE configs[4] chunk
([8192, 50, 3, 32, 32] fp32 x 2 = 10.07 GB, the shape bench.py times) and on a bench-workload-shaped pair ([4736, 51, 256] x 2).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \\
        --clock-control none -k regex:k_metrics_pairs --csv --log-file gpurun_out/ncu_metrics.csv python tools/ncu_metrics.py
    python tools/ncu_metrics.py table gpurun_out/ncu_metrics.csv >> profiles/rNN_metrics_ncu.txt

(the same tensors as bench.py's `roofline_other` legs; one warm-up call and two profiled calls per shape)"""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run():
    import torch
    from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm
    dev = torch.device("cuda", 0)
    for N, L, D in ((8192, 50, 3072), (4736, 51, 256)):
        gen = torch.Generator(device=dev).manual_seed(1234)
        t = torch.randn(N, 1, D, device=dev, generator=gen) + 0.1 * torch.cumsum(torch.randn(N, L, D, device=dev, generator=gen), dim=1)
        s = t + 0.05 * torch.randn(N, L, D, device=dev, generator=gen)
        for _ in range(3):
            tm.pair_reductions(t, s)
        torch.cuda.synchronize()
        del t, s
        torch.cuda.empty_cache()
    print("done")


def table(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    iid, ik, im, iv, ig = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    by = {}
    for r in rows[1:]:
        by.setdefault(r[iid], {"kernel": r[ik], "grid": r[ig]})[r[im]] = float(r[iv].replace(",", ""))
    print(f"{'launch':>6s} {'kernel':28s} {'grid':>14s} {'time_us':>9s} {'dram_read_MB':>13s} {'dram_write_MB':>14s} {'dram_pct_peak':>14s} {'GB/s (read+write)':>18s}")
    for k, v in by.items():
        us = v["gpu__time_duration.sum"] / 1e3
        rd, wr = v["dram__bytes_read.sum"] / 1e6, v["dram__bytes_write.sum"] / 1e6
        print(f"{k:>6s} {v['kernel'][:28]:28s} {v['grid']:>14s} {us:9.1f} {rd:13.1f} {wr:14.2f} "
              f"{v.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):14.1f} {(rd + wr) / us * 1e3:18.0f}")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "table":
        table(sys.argv[2])
    else:
        run()
