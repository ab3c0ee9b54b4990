#!/usr/bin/env python
"""Diagnostics for the tcgen05 implicit-GEMM kernels (developer tool, run on the GPU box).

Feeds the single-layer C-ABI entry points structured inputs (identity weights on one tap, ramps) so that a wrong
descriptor / swizzle / tap mapping shows up as a recognisable pattern instead of a bare mismatch."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cartseg import _lib                                    # noqa: E402
from gpu_util import bf16_round, from_nhwc, layer_scratch, rel_l2, stream, to_nhwc_bf16   # noqa: E402

L = _lib.lib()


def conv_fprop(x, w, stats=False):
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    y = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    buf, scratch = layer_scratch(Cin, Cout)
    xg, wg = to_nhwc_bf16(x), w.cuda()                  # keep alive until the synchronize
    _lib.check(L.cs_conv3x3_fprop(xg.data_ptr(), B, H, W, Cin, wg.data_ptr(), Cout, y.data_ptr(),
                                  None, None, scratch, stream()), "fprop")
    torch.cuda.synchronize()
    return from_nhwc(y)


def main():
    torch.manual_seed(0)
    print("device:", torch.cuda.get_device_name(0))
    # 1. one tile, identity on each tap
    B, H, W, C = 1, 16, 8, 64
    x = bf16_round(torch.randn(B, C, H, W))
    for kh in range(3):
        for kw in range(3):
            w = torch.zeros(C, C, 3, 3)
            w[:, :, kh, kw] = torch.eye(C)
            ref = F.conv2d(x, w, padding=1)
            got = conv_fprop(x, w)
            e = (got - ref).abs().max().item()
            msg = f"tap(kh={kh},kw={kw}) max|err|={e:.4f} nan={int(torch.isnan(got).sum())}"
            if e > 1e-2:
                # which tap does the output actually look like?
                best = None
                for a in range(3):
                    for b in range(3):
                        w2 = torch.zeros(C, C, 3, 3); w2[:, :, a, b] = torch.eye(C)
                        r2 = F.conv2d(x, w2, padding=1)
                        d = (torch.nan_to_num(got) - r2).abs().max().item()
                        if best is None or d < best[0]:
                            best = (d, a, b)
                msg += f"  closest tap=({best[1]},{best[2]}) err={best[0]:.4f}"
            print(msg)
    # 2. channel permutation check: centre tap, w = permutation matrix (co <- ci = (co*7+3) % 64)
    w = torch.zeros(C, C, 3, 3)
    perm = [(co * 7 + 3) % C for co in range(C)]
    for co, ci in enumerate(perm):
        w[co, ci, 1, 1] = 1.0
    got = conv_fprop(x, w)
    ref = F.conv2d(x, w, padding=1)
    print("channel permutation: max|err| =", (got - ref).abs().max().item())
    # 3. random, growing sizes
    for (B, H, W, Ci, Co) in [(1, 16, 8, 64, 64), (1, 16, 16, 64, 64), (2, 32, 32, 64, 64), (1, 16, 8, 128, 64),
                              (1, 16, 8, 64, 128), (1, 16, 8, 64, 256), (1, 16, 8, 256, 256), (1, 14, 14, 64, 64),
                              (4, 56, 56, 128, 128)]:
        x = bf16_round(torch.randn(B, Ci, H, W))
        w = bf16_round(torch.randn(Co, Ci, 3, 3) * (2.0 / (9 * Ci)) ** 0.5)
        got = conv_fprop(x, w)
        ref = F.conv2d(x, w, padding=1)
        print(f"random B{B} {H}x{W} {Ci}->{Co}: rel-L2 {rel_l2(torch.nan_to_num(got), ref):.3e} "
              f"nan={int(torch.isnan(got).sum())}")


if __name__ == "__main__":
    main()
