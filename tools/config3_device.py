"""GPU: device-resident throughput of the BASELINE configs[2] sweep (teacher + 11 students, 3x32x32, w = 7.5) for a given chunk
size and number of student streams:  python tools/config3_device.py <seeds per chunk> [DTRAJ_MODEL_STREAMS]"""
import os, sys
if len(sys.argv) > 2:
    os.environ["DTRAJ_MODEL_STREAMS"] = sys.argv[2]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distillation_trajectories_b200 import grid
from distillation_trajectories_b200.engine import UNetEngine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
t32 = bench.make_model(bench.Cfg32, 1.0, 0, dev)
studs = [bench.make_model(bench.Cfg32, sf, 1000 + int(sf * 100), dev) for sf in bench.SF11]
chunks = [grid.stage_chunk(list(range(i * S, (i + 1) * S)), bench.Cfg32, [7.5], dev) for i in range(3)]
grid.run_chunk(t32, studs, chunks[0], dev, "f16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in (1, 2):
    grid.run_chunk(t32, studs, chunks[i], dev, "f16")
e1.record()
torch.cuda.synchronize()
sec = e0.elapsed_time(e1) / 1e3
fl = sum(sum(next(reversed(UNetEngine.for_model(m, 32, 50, "f16", dev)._samplers.values())).flops()) for m in [t32] + studs)
print(f"seeds/chunk {S} streams {os.environ.get('DTRAJ_MODEL_STREAMS', '3')}: {2 * 12 * S / sec:.0f} trajectories/s, {2 * fl / sec / 1e12:.0f} TFLOP/s algorithmic, "
      f"{sec / 2 * 1e3:.0f} ms per chunk")
