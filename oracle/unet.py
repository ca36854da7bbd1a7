"""Oracle: functional fp32 CPU restatement of the reference U-Net forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Operates on a bare ``state_dict``
(the reference's checkpoint format, scripts/train_teacher.py:86) instead of an
``nn.Module`` so that it shares no code with the product package.

Follows /root/reference/models.py:
  * sinusoidal embedding ............ models.py:15-39
  * Block (conv-bn-relu, +temb, conv-bn-relu, +residual) ... models.py:59-83
  * U-Net wiring ..................... models.py:159-224
BatchNorm is evaluated in eval mode (running statistics, eps=1e-5) and dropout is
the identity, which is what every caller of the hot path uses.
"""
import math

import torch
import torch.nn.functional as F

BLOCKS = ("enc1", "enc2", "enc3", "enc4", "bottleneck", "dec3", "dec2", "dec1")


def sinusoidal_embedding(t, dim):
    """models.py:15-39.  ``t`` is a 1-d tensor (any dtype); returns [B, max(dim,2)]."""
    dim = max(dim, 2)
    half = max(dim // 2, 1)
    scale = math.log(10000) / (half - 1 + 1e-8)
    freqs = torch.exp(torch.arange(half) * -scale)
    arg = t[:, None] * freqs[None, :]
    emb = torch.cat((arg.sin(), arg.cos()), dim=-1)
    if emb.shape[-1] < dim:
        emb = torch.cat((emb, torch.zeros(emb.shape[0], dim - emb.shape[-1])), dim=-1)
    elif emb.shape[-1] > dim:
        emb = emb[:, :dim]
    return emb


def time_embedding(sd, t, cond):
    """models.py:175-185: time MLP, plus cond MLP iff ``cond`` is given."""
    temb_dim = sd["time_mlp.1.weight"].shape[0]
    if t.dim() > 1:
        t = t.reshape(t.shape[0], -1)[:, 0]
    emb = sinusoidal_embedding(t, temb_dim)
    emb = F.relu(F.linear(emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    if cond is not None:
        c = F.relu(F.linear(cond, sd["cond_emb.0.weight"], sd["cond_emb.0.bias"]))
        emb = emb + F.linear(c, sd["cond_emb.2.weight"], sd["cond_emb.2.bias"])
    return emb


def _bn_eval(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=1e-5)


def block_forward(sd, name, x, temb):
    """models.py:59-83."""
    key = name + ".residual_conv.weight"
    if key in sd:
        res = F.conv2d(x, sd[key], sd[name + ".residual_conv.bias"])
    else:
        res = x
    h = F.relu(_bn_eval(F.conv2d(x, sd[name + ".conv1.weight"], sd[name + ".conv1.bias"], padding=1),
                        sd, name + ".norm1"))
    tb = F.relu(F.linear(temb, sd[name + ".time_mlp.weight"], sd[name + ".time_mlp.bias"]))
    h = h + tb[:, :, None, None]
    h = F.relu(_bn_eval(F.conv2d(h, sd[name + ".conv2.weight"], sd[name + ".conv2.bias"], padding=1),
                        sd, name + ".norm2"))
    return h + res


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


@torch.no_grad()
def unet_forward(sd, x, t, cond=None):
    """models.py:159-224.  ``x`` [B,C,H,W] fp32, ``t`` [B] int64, ``cond`` None or [B,1]."""
    temb = time_embedding(sd, t, cond)
    x1 = block_forward(sd, "enc1", x, temb)
    x2 = block_forward(sd, "enc2", F.max_pool2d(x1, 2), temb)
    x3 = block_forward(sd, "enc3", F.max_pool2d(x2, 2), temb)
    x4 = block_forward(sd, "enc4", F.max_pool2d(x3, 2), temb)
    h = block_forward(sd, "bottleneck", F.max_pool2d(x4, 2), temb)
    h = block_forward(sd, "dec3", torch.cat([_up2(h), x4], dim=1), temb)
    h = block_forward(sd, "dec2", torch.cat([_up2(h), x3], dim=1), temb)
    h = block_forward(sd, "dec1", torch.cat([_up2(h), x2], dim=1), temb)
    return F.conv2d(_up2(h), sd["final.weight"], sd["final.bias"])


def block_time_bias(sd, name, temb):
    """relu(time_mlp(temb)) of one block, models.py:66-67 (used to check the
    product's per-(t, variant) table)."""
    return F.relu(F.linear(temb, sd[name + ".time_mlp.weight"], sd[name + ".time_mlp.bias"]))
