"""ncu target: one chunk of the bench workload (teacher as both models, 592 seeds x 8 scales = 8880 forward rows, fp16) -- used with
    ncu --set full --import-source on -k regex:k_conv_umma_t --launch-skip 15 --launch-count 1 python tools/_ncu_layer.py
to capture enc2.conv2 +res +pool of the second forward (profiles/r02m_enc2conv2_full.txt, r02i_epilogue_stalls.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from distillation_trajectories_b200 import grid
dev = torch.device("cuda", 0)
ck = grid.stage_chunk(list(range(592)), bench.Cfg, bench.GUIDANCE, dev)
m = bench.make_model(bench.Cfg, 1.0, 0, dev)
grid.run_chunk(m, [m], ck, dev, "f16")
torch.cuda.synchronize()
