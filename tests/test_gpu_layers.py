"""GPU parity of the tcgen05 implicit-GEMM kernels, one layer at a time, through the C ABI
(cs_conv3x3_* / cs_convT2x2_*, include/cartseg.h).  Reference: torch CPU fp32 on the SAME
bf16-rounded operands (these are floating-point kernels: bf16 inputs, fp32 accumulate, bf16 / fp32
outputs).  Tolerance: 1e-2 of the output scale (one bf16 output rounding is 2^-9 relative)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16_round, from_nhwc, layer_scratch, rel_l2, stream, to_nhwc_bf16

pytestmark = pytest.mark.gpu

# (B, H, W, Cin, Cout): covers BLOCK_N 64 / 128 / 256, ragged tiles (14x14, 28x28), multi-chunk K
CONV_SHAPES = [
    (2, 16, 16, 64, 64),
    (1, 32, 24, 64, 128),
    (2, 14, 14, 128, 256),
    (1, 28, 28, 256, 128),
    (3, 16, 8, 128, 64),
    (1, 48, 40, 64, 64),
    (2, 14, 14, 512, 512),
    (8, 112, 112, 64, 64),        # 1568 tiles: several per persistent CTA, split-K wgrad
    (4, 56, 56, 128, 256),
]


def _lib():
    from cartseg import _lib
    return _lib.lib(), _lib.check


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return bf16_round(torch.randn(*shape, generator=g) * scale)


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_fprop_and_stats(B, H, W, Cin, Cout):
    L, check = _lib()
    x = _rand((B, Cin, H, W), 1)
    w = _rand((Cout, Cin, 3, 3), 2, scale=(2.0 / (9 * Cin)) ** 0.5)
    ref = F.conv2d(x, w, padding=1)
    xg, wg = to_nhwc_bf16(x), w.cuda()
    y = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    ssum = torch.zeros(Cout, dtype=torch.float64, device="cuda")
    ssq = torch.zeros(Cout, dtype=torch.float64, device="cuda")
    buf, scratch = layer_scratch(Cin, Cout)
    check(L.cs_conv3x3_fprop(xg.data_ptr(), B, H, W, Cin, wg.data_ptr(), Cout, y.data_ptr(), ssum.data_ptr(),
                             ssq.data_ptr(), scratch, stream()), "cs_conv3x3_fprop")
    torch.cuda.synchronize()
    got = from_nhwc(y)
    assert torch.isfinite(got).all()
    err = rel_l2(got, ref)
    assert err < 1e-2, err
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-2 * ref.abs().max().item(), rtol=2e-2)
    # statistics are taken over the stored (bf16-rounded) outputs
    np.testing.assert_allclose(ssum.cpu().numpy(), got.double().sum((0, 2, 3)).numpy(), rtol=1e-4,
                               atol=1e-3 * B * H * W ** 0.5)
    np.testing.assert_allclose(ssq.cpu().numpy(), (got.double() ** 2).sum((0, 2, 3)).numpy(), rtol=1e-4)


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_dgrad(B, H, W, Cin, Cout):
    L, check = _lib()
    dy = _rand((B, Cout, H, W), 3)
    w = _rand((Cout, Cin, 3, 3), 4, scale=(2.0 / (9 * Cout)) ** 0.5)
    ref = F.conv_transpose2d(dy, w, padding=1)            # gradient of conv2d(pad=1) w.r.t. its input
    dx = torch.full((B, H, W, Cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    buf, scratch = layer_scratch(Cin, Cout)
    dyg, wg = to_nhwc_bf16(dy), w.cuda()                  # kept alive until the synchronize below
    check(L.cs_conv3x3_dgrad(dyg.data_ptr(), B, H, W, Cin, wg.data_ptr(), Cout, dx.data_ptr(),
                             scratch, stream()), "cs_conv3x3_dgrad")
    torch.cuda.synchronize()
    got = from_nhwc(dx)
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < 1e-2


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_wgrad(B, H, W, Cin, Cout):
    L, check = _lib()
    x = _rand((B, Cin, H, W), 5)
    dy = _rand((B, Cout, H, W), 6)
    xr = x.clone().requires_grad_(False)
    wr = torch.zeros(Cout, Cin, 3, 3, requires_grad=True)
    F.conv2d(xr, wr, padding=1).backward(dy)
    ref = wr.grad
    dw = torch.full((Cout, Cin, 3, 3), float("nan"), dtype=torch.float32, device="cuda")
    buf, scratch = layer_scratch(Cin, Cout)
    xg, dyg = to_nhwc_bf16(x), to_nhwc_bf16(dy)
    check(L.cs_conv3x3_wgrad(xg.data_ptr(), dyg.data_ptr(), B, H, W, Cin, Cout, dw.data_ptr(), scratch, stream()),
          "cs_conv3x3_wgrad")
    torch.cuda.synchronize()
    got = dw.cpu()
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < 2e-3                        # fp32 accumulate and fp32 output: no bf16 rounding


def _bnrelu_operand(B, Cin, H, W, seed):
    """Raw operand, per-channel affine (both signs of scale, shifts that leave ~half of the values positive) and the
    activation the kernels must see: relu(fma(x, sc, sh)) rounded to bf16, as bn_relu_kernel stores it."""
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(B, Cin, H, W, generator=g))
    sc = (torch.rand(Cin, generator=g) + 0.5) * torch.where(torch.rand(Cin, generator=g) < 0.25, -1.0, 1.0)
    sh = torch.randn(Cin, generator=g) * 0.5
    act = bf16_round(torch.relu(torch.addcmul(sh.view(1, -1, 1, 1).double(), x.double(), sc.view(1, -1, 1, 1).double()).float()))
    return x, sc, sh, act


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_fprop_with_fused_bn_relu_operand(B, H, W, Cin, Cout):
    """convX.3 of a DoubleConv reading the RAW output of convX.0 (create_testset.py:40-52): BN + ReLU applied to the
    operand patch in shared memory, zero padding AFTER the activation (a shift > 0 must not leak into the border)."""
    L, check = _lib()
    x, sc, sh, act = _bnrelu_operand(B, Cin, H, W, 21)
    w = _rand((Cout, Cin, 3, 3), 22, scale=(2.0 / (9 * Cin)) ** 0.5)
    ref = F.conv2d(act, w, padding=1)
    xg, wg, scg, shg = to_nhwc_bf16(x), w.cuda(), sc.cuda(), sh.cuda()
    y = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    ssum = torch.zeros(Cout, dtype=torch.float64, device="cuda")
    ssq = torch.zeros(Cout, dtype=torch.float64, device="cuda")
    buf, scratch = layer_scratch(Cin, Cout)
    check(L.cs_conv3x3_fprop_bnrelu(xg.data_ptr(), scg.data_ptr(), shg.data_ptr(), B, H, W, Cin, wg.data_ptr(), Cout,
                                    y.data_ptr(), ssum.data_ptr(), ssq.data_ptr(), scratch, stream()), "cs_conv3x3_fprop_bnrelu")
    torch.cuda.synchronize()
    got = from_nhwc(y)
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < 1e-2
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-2 * ref.abs().max().item(), rtol=2e-2)
    np.testing.assert_allclose(ssq.cpu().numpy(), (got.double() ** 2).sum((0, 2, 3)).numpy(), rtol=1e-4)
    # the raw operand is left untouched in HBM
    assert torch.equal(from_nhwc(xg), x)


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_wgrad_with_fused_bn_relu_operand(B, H, W, Cin, Cout):
    L, check = _lib()
    x, sc, sh, act = _bnrelu_operand(B, Cin, H, W, 23)
    dy = _rand((B, Cout, H, W), 24)
    wr = torch.zeros(Cout, Cin, 3, 3, requires_grad=True)
    F.conv2d(act, wr, padding=1).backward(dy)
    dw = torch.full((Cout, Cin, 3, 3), float("nan"), dtype=torch.float32, device="cuda")
    buf, scratch = layer_scratch(Cin, Cout)
    xg, dyg, scg, shg = to_nhwc_bf16(x), to_nhwc_bf16(dy), sc.cuda(), sh.cuda()
    check(L.cs_conv3x3_wgrad_bnrelu(xg.data_ptr(), scg.data_ptr(), shg.data_ptr(), dyg.data_ptr(), B, H, W, Cin, Cout,
                                    dw.data_ptr(), scratch, stream()), "cs_conv3x3_wgrad_bnrelu")
    torch.cuda.synchronize()
    got = dw.cpu()
    assert torch.isfinite(got).all()
    assert rel_l2(got, wr.grad) < 2e-3


CONVT_SHAPES = [(2, 8, 8, 128, 64), (1, 14, 14, 256, 128), (2, 7, 7, 1024, 512), (1, 16, 24, 512, 256)]


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONVT_SHAPES)
def test_convT2x2_fprop_into_concat_slot(B, H, W, Cin, Cout):
    L, check = _lib()
    x = _rand((B, Cin, H, W), 7)
    w = _rand((Cin, Cout, 2, 2), 8, scale=(1.0 / Cin) ** 0.5)
    bias = torch.randn(Cout, generator=torch.Generator().manual_seed(9)) * 0.1
    ref = F.conv_transpose2d(x, w, bias, stride=2)
    pitch = 2 * Cout                                        # written into the first half of a concat buffer
    y = torch.full((B, 2 * H, 2 * W, pitch), 7.0, dtype=torch.bfloat16, device="cuda")
    buf, scratch = layer_scratch(Cin, Cout)
    xg, wg, bg = to_nhwc_bf16(x), w.cuda(), bias.cuda()
    check(L.cs_convT2x2_fprop(xg.data_ptr(), B, H, W, Cin, wg.data_ptr(), bg.data_ptr(),
                              Cout, y.data_ptr(), pitch, scratch, stream()), "cs_convT2x2_fprop")
    torch.cuda.synchronize()
    got = from_nhwc(y[..., :Cout])
    assert rel_l2(got, ref) < 1e-2
    assert (y[..., Cout:].float() == 7.0).all()             # the skip half is left untouched


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONVT_SHAPES)
def test_convT2x2_dgrad_and_wgrad(B, H, W, Cin, Cout):
    L, check = _lib()
    x = _rand((B, Cin, H, W), 10)
    w = _rand((Cin, Cout, 2, 2), 11, scale=(1.0 / Cin) ** 0.5)
    dy = _rand((B, Cout, 2 * H, 2 * W), 12)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, None, stride=2).backward(dy)
    pitch = 2 * Cout
    dyg = torch.zeros((B, 2 * H, 2 * W, pitch), dtype=torch.bfloat16, device="cuda")
    dyg[..., :Cout] = to_nhwc_bf16(dy)
    dyg[..., Cout:] = 3.0                                   # must be ignored
    dx = torch.full((B, H, W, Cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    buf, scratch = layer_scratch(Cin, Cout)
    wg, xg = w.cuda(), to_nhwc_bf16(x)
    check(L.cs_convT2x2_dgrad(dyg.data_ptr(), pitch, B, H, W, Cin, wg.data_ptr(), Cout, dx.data_ptr(), scratch,
                              stream()), "cs_convT2x2_dgrad")
    dw = torch.full((Cin, Cout, 2, 2), float("nan"), dtype=torch.float32, device="cuda")
    check(L.cs_convT2x2_wgrad(xg.data_ptr(), dyg.data_ptr(), pitch, B, H, W, Cin, Cout, dw.data_ptr(),
                              scratch, stream()), "cs_convT2x2_wgrad")
    torch.cuda.synchronize()
    assert rel_l2(from_nhwc(dx), xr.grad) < 1e-2
    assert rel_l2(dw.cpu(), wr.grad) < 2e-3
