#!/usr/bin/env python
"""Benchmark of the trajectory hot path (BASELINE.json metric: denoising trajectories/sec, 50 steps, CFG).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPUs

Workload (BASELINE.json configs[1]): teacher U-Net (size_factor 1.0) vs student (0.5), 1x16x16, 50
timesteps, classifier-free-guidance sweep w in {1, 2, 3, 5, 7.5, 10, 15, 20}; one STEP = `--seeds` seeds
x 8 guidance scales on every GPU: teacher trajectories + student trajectories (sampler S2,
analysis/trajectory_engine.py:24-115) + the pair metrics of compute_trajectory_metrics.  A trajectory is
one (model, seed, w) run of 51 frames.  Seeds are sharded over ranks (weak scaling: `--seeds` per GPU),
the only collective is the final all-reduce of the metric sums.

value   device-resident: inputs already in HBM, K x (2 captured sampling loops + metric kernels), CUDA events.
e2e     through the public API (grid.sweep, the batched compare_trajectories): host RNG draws, pinned
        host -> device copies, device work, device -> host copy of the reductions, f64 scalar formulas.
"""
import argparse
import contextlib
import functools
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GUIDANCE = [1.0, 2.0, 3.0, 5.0, 7.5, 10.0, 15.0, 20.0]
METRIC = "denoising trajectories/sec (50 steps, CFG)"
UNIT = "trajectories/s"


class Cfg:
    channels, image_size, timesteps, dropout = 1, 16, 50, 0.3
    sample_steps = teacher_steps = student_steps = 50
    beta_start, beta_end = 1e-4, 0.02


def workload_config(seeds, world):
    return {"workload": "configs[1]: teacher(sf=1.0) vs student(sf=0.5), 1x16x16, 50 steps, CFG sweep "
                        "w in {1,2,3,5,7.5,10,15,20}, sampler S2 + pair metrics",
            "seeds_per_gpu_per_step": seeds, "guidance_scales": GUIDANCE,
            "trajectories_per_step": 2 * seeds * len(GUIDANCE) * world,
            "parallelism": f"seed-sharded x{world}, one all-reduce of metric sums",
            "dedup": "teacher trajectories computed once per (seed, w); at the first step the 8 scales of a seed still hold "
                     "the same x_T, so its forward rows are evaluated once per (seed, conditioning variant) -- bit-identical "
                     "frames (tests/test_gpu_samplers.py::test_first_step_row_sharing_is_exact)",
            "l2": f"per-step working set (2 x {seeds * 8 * 51 * 256 * 4 / 1e6:.0f} MB trajectory buffers + GBs of activations) "
                  "exceeds the 126 MB L2; no explicit flush"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# --------------------------------------------------------------------------------------- CPU arm
def cpu_step(n_seeds, first_seed, state):
    """The reference algorithm (oracle port of compare_trajectories) on the host CPUs:
    n_seeds x 8 guidance scales x {teacher, student} trajectories + their metrics."""
    import numpy as np
    import torch
    from oracle import metrics as om
    from oracle import samplers as osmp
    ft, fs = state
    n = 0
    for s in range(first_seed, first_seed + n_seeds):
        seed = 42 + s
        torch.manual_seed(seed)
        np.random.seed(seed)
        noise = torch.randn(1, Cfg.channels, Cfg.image_size, Cfg.image_size)
        for gs in GUIDANCE:
            a = osmp.s2_generate_trajectory(ft, noise, Cfg.timesteps, seed=seed, guidance_scale=gs)
            b = osmp.s2_generate_trajectory(fs, noise, Cfg.timesteps, seed=seed, guidance_scale=gs)
            om.trajectory_metrics(a, b)
            n += 2
    return n


def cpu_state():
    import torch
    from oracle import unet as ounet
    from distillation_trajectories_b200.models import DiffusionUNet
    fns = []
    for sf, seed in ((1.0, 0), (0.5, 1050)):
        torch.manual_seed(seed)
        with quiet():
            m = DiffusionUNet(Cfg, sf).eval()
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        fns.append(functools.partial(ounet.unet_forward, sd))
    return fns


def run_reference(args, emit=print):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the Python reference
    cannot travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    state = cpu_state()
    seeds = 1
    for i in range(args.warmup):
        cpu_step(seeds, i, state)
    t0 = time.perf_counter()
    n = 0
    for i in range(args.steps):
        n += cpu_step(seeds, args.warmup + i, state)
    dt = time.perf_counter() - t0
    v = n / dt
    sample = f"{seeds} seed x {len(GUIDANCE)} guidance scales x (teacher, student) = {2 * seeds * len(GUIDANCE)} trajectories + {seeds * len(GUIDANCE)} metric pairs per step"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (random-init weights, seeded N(0,1) noise)",
        "config": dict(workload_config(seeds, 1), note="CPU arm: rank 0 only, bounded sample per step"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                         "torch": torch.__version__},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# --------------------------------------------------------------------------------------- GPU arm
def run_ours(args, emit=print):
    import numpy as np
    import torch
    import torch.distributed as dist

    from distillation_trajectories_b200 import grid
    from distillation_trajectories_b200.engine import UNetEngine
    from distillation_trajectories_b200.models import DiffusionUNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("DTRAJ_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU on one host: do not let every rank's tiny host-side torch ops spawn a full-size thread pool
        torch.set_num_threads(max(1, (os.cpu_count() or world) // world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    models = []
    for sf, seed in ((1.0, 0), (0.5, 1050)):
        torch.manual_seed(seed)
        with quiet():
            models.append(DiffusionUNet(Cfg, sf).eval().to(dev))
    teacher, student = models
    S, K, W, G = args.seeds, args.steps, args.warmup, len(GUIDANCE)

    # ---- device-resident leg: stage W + K different chunks, then time K x run_chunk with CUDA events
    chunks = [grid.stage_chunk([(i * world + rank) * S + j for j in range(S)], Cfg, GUIDANCE, dev) for i in range(W + K)]
    torch.cuda.synchronize()
    keep = []
    for i in range(W):
        keep = grid.run_chunk(teacher, [student], chunks[i], dev, args.precision)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(K):
            keep = grid.run_chunk(teacher, [student], chunks[W + i], dev, args.precision)
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    traj_per_step = 2 * S * G * world
    value = traj_per_step * K / (ms / 1e3)
    samplers = [next(iter(UNetEngine.for_model(m, Cfg.image_size, Cfg.timesteps, args.precision, dev)._samplers.values()))
                for m in models]
    launches = K * (sum(s.launches for s in samplers) + 2)

    # ---- end-to-end leg: the public sweep API with host-side inputs, copies inside the timed region
    stats = {}
    e2e_pairs = max(G, (S // max(1, args.e2e_chunks)) * G)
    for i in range(max(1, W // 2)):
        grid.sweep(teacher, {"student": student}, Cfg, GUIDANCE, S * world, dev, rank, world, max_pairs=e2e_pairs,
                   precision=args.precision)
    barrier()
    # ONE sweep call over K steps' worth of seeds, chunked at a step (S seeds x G scales per rank): the API's own
    # software pipeline (stage chunk i+1 / run chunk i / finish chunk i-1) is what a sweep larger than one batch gets.
    # Every chunk's host draws, pinned H2D copies, D2H read-back and f64 host formulas are inside the timed region.
    t0 = time.perf_counter()
    res = grid.sweep(teacher, {"student": student}, Cfg, GUIDANCE, K * S * world, dev, rank, world, max_pairs=e2e_pairs,
                     precision=args.precision, stats=stats)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if os.environ.get("DTRAJ_SWEEP_TIMING"):
        print(f"[rank {rank}] e2e {e2e_s * 1e3:.0f} ms; host ms per phase: " +
              ", ".join(f"{k[7:]} {v * 1e3:.0f}" for k, v in stats.items() if k.startswith("host_s_")), file=sys.stderr)
    e2e = {"value": traj_per_step * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": stats["h2d_bytes"] // K,
           "d2h_bytes_per_step": stats["d2h_bytes"] // K, "ms_per_step": e2e_s / K * 1e3,
           "api": "distillation_trajectories_b200.grid.sweep (batched compare_trajectories): one call over steps x seeds, "
                  "one chunk per step, chunks software-pipelined by the API",
           "chunks_per_step": args.e2e_chunks,
           "check": {"trajectory_mse@w=7.5": res["student"][7.5]["trajectory_mse"],
                     "distribution_similarity@w=7.5": res["student"][7.5]["distribution_similarity"]}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: per-launch CUDA events over one un-captured pass of each loop
    hbm, tc_burst, tc_sust, src = peaks()
    prof = [s.profile() for s in samplers]
    conv_ms = sum(p["ms"][0] for p in prof)
    conv_fl = sum(p["conv_flops"] for p in prof)
    conv_n = sum(p["launches"][0] for p in prof)
    all_ms = sum(sum(p["ms"]) for p in prof)
    ach = conv_fl / (conv_ms / 1e3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("k_conv_umma_dram_bytes_per_launch")
    kname = {"fp32": "k_conv_simt", "f16": "k_conv_umma (tcgen05 kind::f16 implicit-GEMM conv, fp32 accumulate)"}.get(
        args.precision, "k_conv_umma (tcgen05 kind::tf32 implicit-GEMM conv)")
    pnote = ": dense bf16 sustained" + ("; kind::f16 issues at the bf16 rate" if args.precision == "f16"
                                        else "; kind::tf32 issues at half the bf16 rate")
    roofline = {"kernel": kname,
                "bound": "tensor", "achieved": ach, "peak": tc_sust, "unit": "TFLOP/s", "frac": ach / tc_sust,
                "traffic": traffic, "peak_source": src + pnote,
                "launches_timed": conv_n, "avg_launch_us": conv_ms / max(conv_n, 1) * 1e3,
                "flops_per_launch": conv_fl / max(conv_n, 1), "share_of_loop_time": conv_ms / all_ms,
                "class_ms": {"conv": conv_ms, "first_conv": sum(p["ms"][1] for p in prof),
                             "pool_upsample_final": sum(p["ms"][2] for p in prof), "step": sum(p["ms"][3] for p in prof),
                             "enc1_fused": sum(p["ms"][4] for p in prof)}}
    e1_ms, e1_fl = sum(p["ms"][4] for p in prof), sum(p["enc1_flops"] for p in prof)
    if e1_ms > 0:
        roofline["enc1_kernel"] = {"kernel": "k_enc1_umma (conv1 on CUDA cores into smem + conv2 by tcgen05 tap views + pool)",
                                   "achieved": e1_fl / (e1_ms / 1e3) / 1e12, "unit": "TFLOP/s (conv2 flops only)",
                                   "frac": e1_fl / (e1_ms / 1e3) / 1e12 / tc_sust}
    # HBM-bound side kernels (algorithmic bytes: SURVEY.md 8d)
    D, L, N = Cfg.channels * Cfg.image_size ** 2, Cfg.timesteps + 1, S * G
    step_ms = sum(p["ms"][3] for p in prof)
    step_bytes = sum(((16 if w <= 1.0 else 20) * D) for w in GUIDANCE) * S * (Cfg.timesteps - 1) * 2 + 2 * N * D * 8
    from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm

    def time_pairs(ta, sa, reps=10):
        for _ in range(3):
            tm.pair_reductions(ta, sa)
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(reps):
            tm.pair_reductions(ta, sa)
        m1.record()
        torch.cuda.synchronize()
        return m0.elapsed_time(m1) / reps

    met_ms = time_pairs(samplers[0].traj.reshape(N, L, D), samplers[1].traj.reshape(N, L, D))
    met_bytes = 2 * N * L * D * 4
    # BASELINE configs[4] ("metric-kernel bandwidth test"): a chunk of the synthetic [262144, 50, 3, 32, 32] pair
    # (the full tensors are 2 x 150 GiB and are processed in chunks of N on one GPU)
    N5, L5, D5 = 8192, 50, 3 * 32 * 32
    gen = torch.Generator(device=dev).manual_seed(1234)
    t5 = torch.randn(N5, 1, D5, device=dev, generator=gen) + 0.1 * torch.cumsum(torch.randn(N5, L5, D5, device=dev, generator=gen), dim=1)
    s5 = t5 + 0.05 * torch.randn(N5, L5, D5, device=dev, generator=gen)
    m5_ms = time_pairs(t5, s5)
    m5_bytes = 2 * N5 * L5 * D5 * 4
    del t5, s5
    side = [{"kernel": "k_step (+ k_copy_frame)", "bound": "hbm", "achieved": step_bytes / (step_ms / 1e3) / 1e9, "peak": hbm,
             "unit": "GB/s", "frac": step_bytes / (step_ms / 1e3) / 1e9 / hbm, "traffic": None,
             "note": f"{step_bytes / (2 * (Cfg.timesteps - 1)) / 1e6:.0f} MB per launch: launch-latency bound at this batch, "
                     f"{step_ms / all_ms * 100:.1f} % of the loop time"},
            {"kernel": f"k_metrics_pairs (this workload: [{N}, {L}, {D}] x 2)", "bound": "hbm",
             "achieved": met_bytes / (met_ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
             "frac": met_bytes / (met_ms / 1e3) / 1e9 / hbm, "traffic": None,
             "note": f"{met_bytes / 1e6:.0f} MB per launch"},
            {"kernel": "k_metrics_pairs (BASELINE configs[4] chunk: [8192, 50, 3, 32, 32] x 2, path length + directional "
                       "consistency + MSE reductions)", "bound": "hbm", "achieved": m5_bytes / (m5_ms / 1e3) / 1e9, "peak": hbm,
             "unit": "GB/s", "frac": m5_bytes / (m5_ms / 1e3) / 1e9 / hbm, "traffic": None,
             "note": f"{m5_bytes / 1e9:.1f} GB per launch, every element read once; metric-kernel GB/s of BASELINE.json's metric"}]

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": {"tf32": "tf32", "tf32x3": "tf32x3", "fp32": "f32", "f16": "f16 (fp32 accumulate)"}[args.precision],
           "data": "synthetic (random-init weights, seeded N(0,1) noise)",
           "config": workload_config(S, world), "e2e": e2e, "gpu_launches": launches,
           "clocks": clk.summary(), "roofline": roofline, "roofline_other": side}

    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        state = cpu_state()
        cpu_step(1, 0, state)                       # warm-up (thread pools, oneDNN primitive caches)
        n_seeds = 2
        t0 = time.perf_counter()
        n = cpu_step(n_seeds, 1, state)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{n_seeds} seeds x {G} guidance scales x (teacher, student) = {n} trajectories "
                                         f"+ {n_seeds * G} metric pairs, {dt:.1f} s",
                               "torch": torch.__version__}
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seeds", type=int, default=592, help="seeds per GPU per step (x 8 guidance scales x 2 models); "
                    "a multiple of 148 keeps every conv grid a whole number of waves")
    ap.add_argument("--precision", default="f16", choices=["tf32", "tf32x3", "fp32", "f16"])
    ap.add_argument("--e2e-chunks", type=int, default=1,
                    help="chunks a sweep is cut into in the end-to-end leg (host staging of chunk i+1 overlaps chunk i on the GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL's version banner,
    # model constructors) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    emit = lines.append
    try:
        if args.impl == "reference":
            run_reference(args, emit)
        else:
            run_ours(args, emit)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
