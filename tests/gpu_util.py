"""Helpers shared by the -m gpu parity tests (they call the product only through its public ops / C ABI)."""
import ctypes as C
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def unpack_bits(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(bool)


def to_nhwc_bf16(x_nchw: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW (CPU) -> contiguous NHWC bf16 on the GPU."""
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def from_nhwc(x_nhwc: torch.Tensor) -> torch.Tensor:
    """NHWC (GPU, any dtype) -> fp32 NCHW on the CPU."""
    return x_nhwc.float().cpu().permute(0, 3, 1, 2).contiguous()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()


def rel_l2(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.double().flatten(), ref.double().flatten()
    return float((got - ref).norm() / max(ref.norm().item(), 1e-30))


def layer_scratch(cin, cout):
    from cartseg import _lib
    n = int(_lib.lib().cs_layer_scratch_bytes(cin, cout))
    buf = torch.empty(n + 1024, dtype=torch.uint8, device="cuda")
    base = (buf.data_ptr() + 1023) & ~1023
    return buf, base


def stream():
    return torch.cuda.current_stream().cuda_stream
