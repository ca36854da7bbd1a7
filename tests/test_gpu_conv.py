"""GPU: one convolution layer through each implementation (CUDA-core fp32, tcgen05 TF32, tcgen05
3xTF32) against an f64 torch convolution, via the C-ABI test hook dtraj_test_conv."""
import ctypes as C

import numpy as np
import pytest
import torch

from distillation_trajectories_b200 import _lib, umma_error_flag

pytestmark = pytest.mark.gpu


def tf32_trunc(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_rna(a):
    u = a.view(np.uint32).astype(np.uint64) + 0x1000
    return (u & 0xFFFFE000).astype(np.uint32).view(np.float32)


def rup(v, m=32):
    return (v + m - 1) // m * m


def run_conv(prec, x0, x1, w, b, ksize, relu, resid):
    """x0/x1 [n,c,H,W] np fp32 (x1 may be None) -> out [n,cout,H,W] through libdtraj."""
    lib = _lib.load()
    dev = torch.device("cuda")
    n, c0, H, W = x0.shape
    c1 = 0 if x1 is None else x1.shape[1]
    cout = w.shape[0]

    f16 = prec == _lib.PREC_F16
    pad = 32                             # channel padding unit of every mode (fp16: a 128-byte K block carries 64 channels, so a
                                         # source padded to an odd multiple of 32 ends in a half-filled, TMA-zero-filled block)
    dt = torch.float16 if f16 else torch.float32

    def nhwc(a, c):
        t = torch.zeros(n, H, W, rup(c, pad), dtype=dt)
        t[..., :c] = torch.from_numpy(a).permute(0, 2, 3, 1).to(dt)
        return t.to(dev).contiguous()

    d0 = nhwc(x0, c0)
    d1 = None if x1 is None else nhwc(x1, c1)
    dr = None if resid is None else nhwc(resid, cout)
    out = torch.full((n, H, W, rup(cout, pad)), float("nan"), dtype=dt, device=dev)
    wc = np.ascontiguousarray(w, np.float32)
    bc = np.ascontiguousarray(b, np.float32)
    rc = lib.dtraj_test_conv(prec, _lib.ptr(d0), c0, _lib.ptr(d1), c1, n, H, W, wc.ctypes.data_as(C.c_void_p),
                             bc.ctypes.data_as(C.c_void_p), cout, ksize, 1 if relu else 0, _lib.ptr(dr), _lib.ptr(out),
                             _lib.stream_ptr())
    _lib.check(rc)
    torch.cuda.synchronize()
    assert umma_error_flag() == 0, "a tcgen05 role timed out on an mbarrier"
    o = out.float().cpu()
    assert torch.isfinite(o).all()
    assert (o[..., cout:] == 0).all() or cout == rup(cout, pad)      # pad channels carry only zeros
    return o[..., :cout].permute(0, 3, 1, 2).numpy()


def reference(x0, x1, w, b, ksize, relu, resid, round_w):
    x = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
    ww = tf32_rna(w) if round_w else w
    y = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(ww).double(),
                                   torch.from_numpy(b).double(), padding=ksize // 2)
    if relu:
        y = y.relu()
    if resid is not None:
        y = y + torch.from_numpy(resid).double()
    return y.numpy()


SHAPES = [
    # n, c0, c1, cout, H, ksize
    (3, 32, 0, 32, 16, 3),
    (2, 16, 0, 25, 16, 3),       # ragged channels (student widths)
    (5, 64, 64, 64, 8, 3),       # decoder: two sources
    (3, 50, 50, 50, 4, 3),       # ragged two-source
    (9, 32, 0, 64, 2, 3),
    (130, 32, 0, 32, 1, 3),      # bottleneck at 1x1: centre tap only, > one tile of images
    (2, 128, 0, 256, 8, 3),      # teacher widths
    (1, 256, 256, 128, 16, 3),
    (2, 32, 0, 32, 32, 3),       # 32x32 level 0: 4-row boxes
    (592, 128, 0, 256, 8, 3),    # 296 tiles: CTA pairs (tcgen05.mma.cta_group::2), persistent loop with 2 tiles per pair
    (300, 64, 64, 128, 8, 3),    # 150 tiles of 128 rows: single-CTA persistent path
    (160, 128, 0, 128, 16, 3),   # 320 tiles of half a 16x16 image: CTA pairs on the im2col path
    (3, 64, 0, 32, 16, 1),       # 1x1 residual
    (4, 38, 38, 76, 4, 1),
    (3, 76, 0, 152, 16, 3),      # sf 0.6 widths: 96 -> 160 padded channels (fp16: half-filled last K block, N = 160)
    (3, 152, 152, 76, 8, 3),     # ... two half-filled sources (160 + 160), N = 96, halo mode in fp16
    (300, 204, 0, 204, 8, 3),    # sf 0.8: 224 channels, CTA pairs with 112 weight rows per CTA
    (2, 16, 0, 32, 32, 3),       # tiny students: one half-filled K block per tap, N = 32
    (2400, 64, 64, 64, 4, 3),    # fp16: position-major tiles (one output position of 128 images, taps outside the map skipped), CTA pairs
    (9500, 32, 0, 64, 2, 3),     # ... 2x2 maps: 4 of 9 taps per tile, 75 image blocks per position padded to 76 for the pairs
]


def f16_round(a):
    return a.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("prec", [_lib.PREC_FP32, _lib.PREC_TF32, _lib.PREC_TF32X3, _lib.PREC_F16],
                         ids=["fp32", "tf32", "tf32x3", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "n%d_c%d+%d_o%d_h%d_k%d" % s)
def test_conv_layer(prec, shape):
    n, c0, c1, cout, H, ksize = shape
    rng = np.random.RandomState(hash(shape) % (2 ** 31))
    x0 = rng.randn(n, c0, H, H).astype(np.float32)
    x1 = rng.randn(n, c1, H, H).astype(np.float32) if c1 else None
    w = (rng.randn(cout, c0 + c1, ksize, ksize) / np.sqrt((c0 + c1) * ksize * ksize)).astype(np.float32)
    b = rng.randn(cout).astype(np.float32) * 0.1
    resid = rng.randn(n, cout, H, H).astype(np.float32)
    if prec == _lib.PREC_TF32:
        # single-pass TF32 sees trunc_tf32(activations) x rna_tf32(weights): feed exact-tf32 activations
        x0 = tf32_trunc(x0)
        x1 = None if x1 is None else tf32_trunc(x1)
    if prec == _lib.PREC_F16:
        # fp16 operands (the same 11-bit significand as tf32), fp32 accumulation, ONE fp16 rounding of the output
        x0 = f16_round(x0)
        x1 = None if x1 is None else f16_round(x1)
        resid = f16_round(resid)
        w = f16_round(w)
    for relu, rs in ((True, resid), (False, None)):
        got = run_conv(prec, x0, x1, w, b, ksize, relu, rs)
        want = reference(x0, x1, w, b, ksize, relu, rs, round_w=(prec == _lib.PREC_TF32))
        scale = np.abs(want).max()
        if prec == _lib.PREC_F16:
            K = (c0 + c1) * ksize * ksize
            bound = (4e-6 + 4e-9 * K) * scale + 2.0 ** -11 * np.abs(want) + 6e-8     # accumulate + output rounding (+ subnormal step)
            assert (np.abs(got - want) <= bound).all(), f"max excess {np.max(np.abs(got - want) - bound):.3e}"
            continue
        # CUDA-core path: fp32 FMA chain.  tcgen05 paths: the tensor core adds into its fp32 accumulator
        # with round-toward-zero, measured ~2e-9 * K of max|ref| (tools/probe_conv_accuracy.py); 3xTF32
        # additionally drops the lo*lo term (2^-22 relative).
        K = (c0 + c1) * ksize * ksize
        tol = 3e-6 if prec == _lib.PREC_FP32 else 4e-6 + 4e-9 * K
        err = np.abs(got - want).max() / scale
        assert err < tol, f"max err / max|ref| = {err:.3e}"
