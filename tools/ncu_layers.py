"""Target + table for the per-layer ncu pass of one forward + step of the teacher and of the sf = 0.5 student at the bench batch.

    python tools/ncu_layers.py run [seeds]              # the workload: a 3-step S2 loop of each model, one stream;
                                                        # without ncu it also writes the launch names of a step to gpurun_out/ncu_layers_names.txt
    DTRAJ_UNDER_NCU=1 ncu --metrics <M> --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/ncu_layers.csv python tools/ncu_layers.py run
    python tools/ncu_layers.py table gpurun_out/ncu_layers.csv gpurun_out/ncu_layers_names.txt   > profiles/rNN_forward_layers_ncu.txt

M = gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum
A step's launches are the same for every timestep, so a 3-step loop (config.timesteps = 3) shows the same layers as the 50-step one;
the table takes the SECOND step of each loop (warm weights, real data)."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(seeds):
    os.environ.setdefault("DTRAJ_OVERLAP_MODELS", "0")     # one stream: a deterministic launch order
    import torch
    import bench
    from distillation_trajectories_b200 import grid
    from distillation_trajectories_b200.engine import UNetEngine

    class Cfg3(bench.Cfg):
        timesteps = 3

    dev = torch.device("cuda", 0)
    teacher, student = bench.make_model(Cfg3, 1.0, 0, dev), bench.make_model(Cfg3, 0.5, 1050, dev)
    ck = grid.stage_chunk(list(range(seeds)), Cfg3, bench.GUIDANCE, dev)
    grid.run_chunk(teacher, [student], ck, dev, "f16")     # (captured loops: ncu profiles the graph's kernel nodes one by one)
    torch.cuda.synchronize()
    if not os.environ.get("DTRAJ_UNDER_NCU"):
        with open(os.path.join(ROOT, "gpurun_out", "ncu_layers_names.txt"), "w") as fh:
            for m, tag in ((teacher, "teacher"), (student, "student")):
                s = next(reversed(UNetEngine.for_model(m, 16, Cfg3.timesteps, "f16", dev)._samplers.values()))
                for name, g, us, fl in s.profile_layers(1):
                    fh.write(f"{tag}\t{name}\t{us:.1f}\n")
    print("done")


def table(csv_path, names_path):
    rows = list(csv.reader(l for l in open(csv_path) if not l.startswith("==")))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iid = hdr.index("ID")
    launches = {}
    order = []
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        k = int(r[iid])
        if k not in launches:
            launches[k] = {"name": r[ik]}
            order.append(k)
        launches[k][r[im]] = float(r[iv].replace(",", ""))
    seq = [launches[k] for k in order]
    names = {"teacher": [], "student": []}
    for ln in open(names_path):
        tag, name, us = ln.rstrip("\n").split("\t")
        names[tag].append((name, float(us)))
    # steps end with k_step; the teacher's loop comes first
    steps, cur = [], []
    for L in seq:
        cur.append(L)
        if L["name"].startswith("k_step"):
            steps.append(cur)
            cur = []
    per_model = len(steps) // 2
    print("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,l1tex__m_xbar2l1tex_read_bytes.sum")
    print("#     --clock-control none -k regex:^k_ : python tools/ncu_layers.py run   (3-step S2 loops of the teacher and the sf = 0.5 student, 592 seeds x 8 scales")
    print("# = 8880 forward rows, fp16 mode, captured loops on one stream); the SECOND step of each loop.  Times are ncu's (serialised, cold L2):")
    print("# compare SHARES with the CUDA-event numbers of bench.py / tools/profile_layers.py (last column), not absolutes.  xbar = L2 -> SM bytes.")
    for mi, tag in enumerate(("teacher", "student")):
        st = steps[mi * per_model + 1]
        nm = names[tag]
        print(f"\n== {tag}")
        print(f"{'layer':40s} {'time_us':>9s} {'dram_MB':>8s} {'xbar_GB':>8s} {'tensor_pipe_active_%':>21s} {'events_us':>10s}")
        tot = 0.0
        for i, L in enumerate(st):
            label, ev = nm[i] if i < len(nm) else (L["name"][:38], float("nan"))
            us = L.get("gpu__time_duration.sum", 0.0) / 1e3
            tot += us
            dram = (L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)) / 1e6
            xbar = L.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0.0) / 1e9
            tp = L.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
            print(f"{label:40s} {us:9.1f} {dram:8.1f} {xbar:8.2f} {tp:21.1f} {ev:10.1f}")
        print(f"{'sum (this forward + step)':40s} {tot:9.1f}")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 592)
    else:
        table(sys.argv[2], sys.argv[3])
