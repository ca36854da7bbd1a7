"""Model containers with the reference's constructor signatures and state_dict layout
(drop-in for /root/reference/models.py:85-242).

The modules below only HOLD parameters -- created in the reference's construction order so that
``torch.manual_seed(k); DiffusionUNet(config, sf)`` yields the same random-init weights and the
same 146 state_dict keys -- and ``forward`` hands the work to the packed CUDA engine
(``engine.UNetEngine``): tcgen05 implicit-GEMM convolutions with folded eval-mode BatchNorm,
ReLU, time-embedding add and residual add in their epilogues.  Training-mode forward/backward is
out of scope and raises.
"""
import torch
import torch.nn as nn

from ._lib import DtrajError, VAR_COND0, VAR_COND1, VAR_NONE
from .engine import UNetEngine, check_device_errors, get_precision


class SinusoidalPositionEmbeddings(nn.Module):
    """Parameter-free placeholder at index 0 of ``time_mlp`` (keeps the key ``time_mlp.1.*``);
    the embedding itself (models.py:15-39) is evaluated by the time-table kernel."""

    def __init__(self, dim):
        super().__init__()
        self.dim = max(dim, 2)


class Block(nn.Module):
    """Parameters of one conv-BN-ReLU x2 block with time MLP and optional 1x1 residual
    (models.py:45-57); registration order matters for random-init parity."""

    def __init__(self, in_ch, out_ch, time_emb_dim=None):
        super().__init__()
        self.time_mlp = nn.Linear(time_emb_dim, out_ch) if time_emb_dim else None
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.norm1 = nn.BatchNorm2d(out_ch)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
        self.norm2 = nn.BatchNorm2d(out_ch)
        self.relu = nn.ReLU()
        self.residual_conv = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else nn.Identity()


class DiffusionUNet(nn.Module):
    """4-level encoder / bottleneck / 3-level decoder U-Net whose widths scale with ``size_factor``."""

    def __init__(self, config, size_factor=1.0):
        super().__init__()
        self.channels = config.channels
        self.size_factor = size_factor
        self.time_emb_dim = max(int(256 * size_factor), 16)
        self.base_channels = max(int(128 * size_factor), 16)
        self.channel_multipliers = [1, 2, 2, 2]
        self.dims = [max(16, int(self.base_channels * m)) for m in self.channel_multipliers]
        print(f"Model size factor: {size_factor}")
        print(f"Model dimensions: {self.dims}")
        d, te = self.dims, self.time_emb_dim
        self.dropout = nn.Dropout(config.dropout)
        self.time_mlp = nn.Sequential(SinusoidalPositionEmbeddings(te), nn.Linear(te, te), nn.ReLU())
        self.cond_emb = nn.Sequential(nn.Linear(1, te), nn.ReLU(), nn.Linear(te, te))
        self.pool = nn.MaxPool2d(2)
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.enc1 = Block(self.channels, d[0], te)
        self.enc2 = Block(d[0], d[1], te)
        self.enc3 = Block(d[1], d[2], te)
        self.enc4 = Block(d[2], d[3], te)
        self.bottleneck = Block(d[3], d[3], te)
        self.dec3 = Block(d[3] + d[3], d[2], te)
        self.dec2 = Block(d[2] + d[2], d[1], te)
        self.dec1 = Block(d[1] + d[1], d[0], te)
        self.final = nn.Conv2d(d[0], self.channels, 1)

    @torch.no_grad()
    def forward(self, x, t, cond=None):
        """eps = U-Net(x, t, cond) (models.py:159-224).  ``t`` [B] (or [B,1]): one shared value (the sampling loops) or
        a timestep per row (p_losses, the distillation step: scripts/train_students.py:131-141); ``cond`` is None or
        a [B,1] tensor of 0/1 flags (the only values the reference feeds)."""
        if self.training:
            raise DtrajError("DiffusionUNet.forward: training mode is out of scope; call model.eval()")
        tv = t.reshape(t.shape[0], -1)[:, 0]
        t0 = int(tv[0].item())
        shared = bool((tv == t0).all())
        if int(tv.min()) < 0:
            raise DtrajError("DiffusionUNet.forward: negative timestep")
        eng = UNetEngine.for_model(self, x.shape[2], (t0 if shared else int(tv.max())) + 1, get_precision("forward"))
        variants = None
        if cond is not None:
            c = cond.reshape(cond.shape[0], -1)[:, 0].to(torch.float32)
            if not bool(((c == 0) | (c == 1)).all()):
                raise DtrajError("DiffusionUNet.forward: cond must be 0/1 flags")
            variants = torch.where(c > 0.5, VAR_COND1, VAR_COND0).to(torch.int32)
        eps = eng.forward(x, t0 if shared else tv, variants)
        check_device_errors(eng.device)          # an fp16 overflow / pipeline time-out must not leave as a plain tensor
        return eps


class SimpleUNet(DiffusionUNet):
    """Teacher alias: size_factor = 1.0 (models.py:227-232)."""

    def __init__(self, config):
        super().__init__(config, size_factor=1.0)


class StudentUNet(DiffusionUNet):
    """Student alias; ``architecture_type`` is accepted and ignored, as in the reference (models.py:234-242)."""

    def __init__(self, config, size_factor=1.0, architecture_type=None):
        if architecture_type is not None:
            print(f"Warning: architecture_type '{architecture_type}' is ignored in the new unified model architecture")
        super().__init__(config, size_factor=size_factor)
