#!/usr/bin/env python
"""Diagnostics for stem_gemm_kernel (developer tool, run on the GPU box): runs a training forward with structured
first-convolution weights (a single tap of a single input channel copied to every output channel) and reports, per
(kh, kw, c), how the raw output of conv1.0 compares with the shifted input — a wrong patch offset / tap order / channel
order shows up as "closest = another tap" instead of a bare mismatch."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cart-segmentation-unet_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cartseg                                              # noqa: E402
from cartseg import _lib, ops                               # noqa: E402


def read(plan, kind, index):
    L = _lib.lib()
    dims = (C.c_int * 4)()
    _lib.check(L.cs_unet_debug_read(plan.handle, kind, index, dims, None, None), "cs_unet_debug_read")
    out = torch.empty(tuple(int(d) for d in dims), dtype=torch.float32, device="cuda")
    _lib.check(L.cs_unet_debug_read(plan.handle, kind, index, dims, out.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "cs_unet_debug_read")
    torch.cuda.synchronize()
    return out.cpu()


def shifted(x, c, dh, dw):
    """x[:, c] sampled at (h + dh, w + dw) with zero padding."""
    B, _, H, W = x.shape
    pad = torch.zeros(B, H + 8, W + 8)
    pad[:, 4:4 + H, 4:4 + W] = x[:, c]
    return pad[:, 4 + dh:4 + dh + H, 4 + dw:4 + dw + W]


def main():
    B, H, W = int(os.environ.get("DIAG_B", 2)), int(os.environ.get("DIAG_H", 32)), int(os.environ.get("DIAG_W", 32))
    torch.manual_seed(0)
    x = torch.randn(B, 3, H, W).bfloat16().float()
    m = cartseg.UNet().cuda().train()
    for kh in range(3):
        for kw in range(3):
            for c in range(3):
                with torch.no_grad():
                    w = torch.zeros(64, 3, 3, 3)
                    w[:, c, kh, kw] = (torch.arange(64) + 1.0) / 64.0     # distinct per output channel: exposes permutations
                    m.conv1.conv[0].weight.copy_(w.cuda())
                z = m(x.cuda())
                plan = ops.get_plan(B, 3, H, W, torch.device("cuda"), inference_only=False)
                yall = read(plan, 0, 0)
                gain = ((torch.arange(64) + 1.0) / 64.0).view(1, 64, 1, 1)
                ref_all = (shifted(x, c, kh - 1, kw - 1).unsqueeze(1) * gain).bfloat16().float()
                chan_err = (yall - ref_all).abs().amax((0, 2, 3))
                y = yall[:, 63]                                 # gain 1.0
                ref = shifted(x, c, kh - 1, kw - 1)
                err = chan_err.max().item()
                if err > 1e-2:
                    print("   wrong output channels:", (chan_err > 1e-2).nonzero().flatten().tolist())
                msg = f"tap(kh={kh},kw={kw},c={c}) max|err|={err:.4f}"
                if err > 1e-2:
                    best = min(((y - shifted(x, cc, dh, dw)).abs().max().item(), cc, dh, dw)
                               for cc in range(3) for dh in range(-3, 4) for dw in range(-4, 5))
                    bad = (y - ref).abs() > 1e-2
                    hw = bad.any(0)
                    msg += (f"  closest: channel {best[1]} shift ({best[2]},{best[3]}) err {best[0]:.4f};"
                            f" wrong pixels {int(bad.sum())}/{bad.numel()}; wrong columns {sorted(set(hw.nonzero()[:, 1].tolist()))[:20]}"
                            f" wrong rows {sorted(set(hw.nonzero()[:, 0].tolist()))[:20]}")
                print(msg)
                del z


if __name__ == "__main__":
    main()
