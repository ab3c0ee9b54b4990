"""Stage-by-stage ("teacher-forced") GPU parity of the U-Net op.

End-to-end gradient comparisons of a ReLU network are ill-conditioned: a 0.5 % forward deviation flips ~0.5 % of the
ReLU masks, which is a ~7 % L2 change of the back-propagated signal per layer (tests/test_host_cpu.py demonstrates
this with the CPU oracle alone).  So every kernel is checked here IN the network, on the tensors the GPU path really
produced: the internal NHWC bf16 buffers are read back through the test hook cs_unet_debug_read
(include/cartseg.h), and each stage's output is compared with plain torch fp32 applied to that stage's actual
inputs (reference semantics: src/create_testset.py:40-83 — Conv3x3 / train-mode BatchNorm2d / ReLU / MaxPool2d /
ConvTranspose2d / cat / 1x1 conv, and their autograd).  Tolerances: one bf16 store (2^-9 relative per element,
~1.5e-3 rel-L2) for bf16 outputs, 3e-3 for fp32 parameter gradients; the north-star bar is 3e-2."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from gpu_util import rel_l2

pytestmark = pytest.mark.gpu

CONV_NAMES = ["conv1.conv.0", "conv1.conv.3", "conv2.conv.0", "conv2.conv.3", "conv3.conv.0", "conv3.conv.3",
              "conv4.conv.0", "conv4.conv.3", "conv5.conv.0", "conv5.conv.3", "dconv4.conv.0", "dconv4.conv.3",
              "dconv3.conv.0", "dconv3.conv.3", "dconv2.conv.0", "dconv2.conv.3", "dconv1.conv.0", "dconv1.conv.3"]
UP_NAMES = ["upconv4", "upconv3", "upconv2", "upconv1"]
Y, OUT, DY, G_OUT, POOLED, G_POOL, UP_OUT, G_UP = range(8)

TOL_BF16 = 4e-3      # tensors stored in bf16
TOL_F32 = 3e-3       # fp32 parameter gradients (bf16 operands, fp32 accumulation)


def _read(plan, kind, index):
    from cartseg import _lib
    L = _lib.lib()
    dims = (C.c_int * 4)()
    _lib.check(L.cs_unet_debug_read(plan.handle, kind, index, dims, None, None), "cs_unet_debug_read")
    out = torch.empty(tuple(int(d) for d in dims), dtype=torch.float32, device="cuda")
    _lib.check(L.cs_unet_debug_read(plan.handle, kind, index, dims, out.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "cs_unet_debug_read")
    torch.cuda.synchronize()
    return out.cpu()


def _bf16(t):
    return t.to(torch.bfloat16).float()


def _snapshot(plan):
    snap = {}
    for i in range(18):
        for kind in (Y, OUT, DY, G_OUT):
            snap[(kind, i)] = _read(plan, kind, i)
        if i < 8 and i % 2 == 1:
            snap[(POOLED, i)] = _read(plan, POOLED, i)
            snap[(G_POOL, i)] = _read(plan, G_POOL, i)
    for k in range(4):
        snap[(UP_OUT, k)] = _read(plan, UP_OUT, k)
        snap[(G_UP, k)] = _read(plan, G_UP, k)
    return snap


class _LazySnapshot(dict):
    """Reads internal tensors on demand (full-size runs: one level-1 tensor of K2 is 0.8 GB as fp32 on the host)."""

    def __init__(self, plan):
        super().__init__()
        self.plan = plan

    def __missing__(self, key):
        return _read(self.plan, key[0], key[1])          # not cached: the caller keeps what it needs

    def __contains__(self, key):
        kind, i = key
        if kind in (POOLED, G_POOL):
            return i < 8 and i % 2 == 1
        return True


def _conv_input(snap, i, x):
    """(input activation of conv i as the GPU saw it, the snapshot keys its gradient is written to)."""
    if i == 0:
        return _bf16(x), None
    if i < 10 and i % 2 == 0:
        return snap[(POOLED, i - 1)], [(G_POOL, i - 1)]
    if i >= 10 and i % 2 == 0:
        k = (i - 10) // 2                       # dconv level 4-k: cat([up_k.out, skip of conv (2*(4-k)-1)])
        skip = 2 * (4 - k) - 1
        return torch.cat([snap[(UP_OUT, k)], snap[(OUT, skip)]], 1), [(G_UP, k), (G_OUT, skip)]
    return snap[(OUT, i - 1)], [(G_OUT, i - 1)]


def _run(model, crit, x, tgt):
    model.zero_grad(set_to_none=True)
    z = model(x.cuda())
    z.retain_grad()
    loss = crit(z, tgt.cuda())
    loss.backward()
    torch.cuda.synchronize()
    return z, loss


def check_stages(snap, sd, x, z, dlogits, grads, conv_ids, up_ids, head):
    """Teacher-forced checks of the listed stages; returns [(name, rel-L2, tolerance)]."""
    rows = []

    def cmp(name, got, ref, tol):
        rows.append((name, rel_l2(got, ref), tol))

    for i in conv_ids:
        n = CONV_NAMES[i]
        inp, gin_keys = _conv_input(snap, i, x)
        w = _bf16(sd[n + ".weight"])
        gamma, beta = sd[n[:-1] + str(int(n[-1]) + 1) + ".weight"], sd[n[:-1] + str(int(n[-1]) + 1) + ".bias"]
        bn = n[:-1] + str(int(n[-1]) + 1)
        # ---- forward: conv (no bias: it cancels in train-mode BN and the kernels never add it)
        y_gpu = snap[(Y, i)]
        cmp(f"fwd {n}: conv", y_gpu, F.conv2d(inp, w, padding=1), TOL_BF16)
        # ---- forward: BN (batch statistics of the stored y) + ReLU (+ pool), autograd graph for the backward check
        yv = y_gpu.clone().requires_grad_(True)
        gv, bv = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        a = F.relu(F.batch_norm(yv, None, None, gv, bv, training=True, eps=1e-5))
        a_q = a + (_bf16(a) - a).detach()                    # stored activation (bf16), identity gradient
        out_gpu = snap[(OUT, i)]
        cmp(f"fwd {n}: bn+relu", out_gpu, a_q.detach(), TOL_BF16)
        outs, gs = [a_q], [snap[(G_OUT, i)]]
        if (POOLED, i) in snap:
            pooled = F.max_pool2d(a_q, 2, 2)
            assert torch.equal(snap[(POOLED, i)], F.max_pool2d(out_gpu, 2, 2)), f"{n}: pool"   # exact
            outs.append(pooled)
            gs.append(snap[(G_POOL, i)])
        del out_gpu
        # ---- backward: BN + ReLU (+ pool routing + skip add)
        torch.autograd.backward(outs, gs)
        dy = snap[(DY, i)]
        cmp(f"bwd {n}: bn+relu dy", dy, yv.grad, TOL_BF16)
        cmp(f"bwd {bn}.weight", grads[bn + ".weight"], gv.grad, TOL_F32)
        cmp(f"bwd {bn}.bias", grads[bn + ".bias"], bv.grad, TOL_F32)
        assert grads[n + ".bias"].abs().max().item() == 0.0          # exactly zero: BN removes the conv bias
        del yv, a, a_q, outs, gs
        # ---- backward: wgrad / dgrad on the GPU's own dy
        iv = inp.clone().requires_grad_(i > 0)
        wv = w.clone().requires_grad_(True)
        F.conv2d(iv, wv, padding=1).backward(dy)
        cmp(f"bwd {n}.weight (wgrad)", grads[n + ".weight"], wv.grad, TOL_F32)
        if gin_keys:
            got = torch.cat([snap[k] for k in gin_keys], 1)
            cmp(f"bwd {n}: dgrad", got, iv.grad, TOL_BF16)
    for k in up_ids:
        n = UP_NAMES[k]
        src = 9 if k == 0 else 11 + 2 * (k - 1)
        inp = snap[(OUT, src)]
        w = _bf16(sd[n + ".weight"])
        iv, wv, bv = inp.clone().requires_grad_(True), w.clone().requires_grad_(True), sd[n + ".bias"].clone().requires_grad_(True)
        out = F.conv_transpose2d(iv, wv, bv, stride=2)
        cmp(f"fwd {n}", snap[(UP_OUT, k)], out.detach(), TOL_BF16)
        out.backward(snap[(G_UP, k)])
        cmp(f"bwd {n}.weight", grads[n + ".weight"], wv.grad, TOL_F32)
        cmp(f"bwd {n}.bias", grads[n + ".bias"], bv.grad, TOL_F32)
        cmp(f"bwd {n}: dgrad", snap[(G_OUT, src)], iv.grad, TOL_BF16)
    if not head:
        return rows
    # ---- head
    a17 = snap[(OUT, 17)].clone().requires_grad_(True)
    hw, hb = sd["final_conv.weight"].clone().requires_grad_(True), sd["final_conv.bias"].clone().requires_grad_(True)
    zr = F.conv2d(a17, hw, hb)
    cmp("fwd final_conv", z, zr.detach(), 1e-5)
    zr.backward(dlogits)
    cmp("bwd final_conv.weight", grads["final_conv.weight"], hw.grad, 1e-4)
    cmp("bwd final_conv.bias", grads["final_conv.bias"], hb.grad, 1e-4)
    cmp("bwd final_conv: dgrad", snap[(G_OUT, 17)], a17.grad, TOL_BF16)

    return rows


@pytest.mark.parametrize("B,H,W,init", [(2, 64, 64, "synth"), (3, 96, 80, "torch"), (2, 224, 224, "torch"), (4, 512, 512, "torch")])
def test_every_stage_teacher_forced(B, H, W, init):
    import cartseg
    from cartseg import ops
    from oracle import unet_oracle as O
    x, tgt = O.synth_batch(B, H, W, seed=5)
    if init == "synth":
        sd = O.synth_state_dict(seed=1)
    else:                                    # the reference's own default initialisation (UNet() under a seed)
        torch.manual_seed(0)
        sd = {k: v.detach().clone() for k, v in cartseg.UNet().state_dict().items()}
    m = cartseg.UNet()
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    m = m.cuda().train()
    z, loss = _run(m, cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7), x, tgt)
    plan = ops.get_plan(B, 3, H, W, torch.device("cuda"), inference_only=False)
    snap = _snapshot(plan)
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    dlogits = z.grad.detach().cpu()
    rows = check_stages(snap, sd, x, z.detach().cpu(), dlogits, grads, range(18), range(4), head=True)
    worst = sorted(rows, key=lambda r: -r[1] / r[2])[:8]
    print(f"\n[{init} B{B} {H}x{W}] {len(rows)} stage checks; worst (rel-L2 / tolerance):")
    for n, e, tol in worst:
        print(f"   {n:40s} {e:.3e} / {tol:g}")
    bad = [(n, f"{e:.3e}", tol) for n, e, tol in rows if not e < tol]
    assert not bad, bad


def test_forward_and_backward_are_run_to_run_deterministic():
    """Same model, same batch, twice: every internal tensor must repeat bit for bit in the forward pass and up to
    fp32 atomic-add ordering in the backward pass; the first tensor that does not is named."""
    import cartseg
    from cartseg import ops
    from oracle import unet_oracle as O
    B, H, W = 2, 64, 64
    x, tgt = O.synth_batch(B, H, W, seed=7)
    torch.manual_seed(1)
    m = cartseg.UNet().cuda().train()              # the reference's default initialisation
    crit = cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)
    plan = None
    snaps, zs, gr = [], [], []
    for _ in range(2):
        z, _ = _run(m, crit, x, tgt)
        plan = plan or ops.get_plan(B, 3, H, W, torch.device("cuda"), inference_only=False)
        snaps.append(_snapshot(plan))
        zs.append(z.detach().cpu())
        gr.append({k: p.grad.detach().cpu().clone() for k, p in m.named_parameters()})
    order = []
    for i in range(18):                                           # forward execution order
        if i >= 10 and i % 2 == 0:
            order.append((UP_OUT, (i - 10) // 2))
        order += [(Y, i), (OUT, i)] + ([(POOLED, i)] if (POOLED, i) in snaps[0] else [])
    fwd_diff = [(k, rel_l2(snaps[1][k], snaps[0][k])) for k in order if not torch.equal(snaps[0][k], snaps[1][k])]
    print("forward tensors that differ run to run (kind, index, rel-L2):", fwd_diff[:6])
    assert not fwd_diff and torch.equal(zs[0], zs[1])
    border = []
    for i in reversed(range(18)):                                 # backward execution order
        border += [(G_OUT, i)] + ([(G_POOL, i)] if (G_POOL, i) in snaps[0] else []) + [(DY, i)]
        if i >= 10 and i % 2 == 0:
            border.append((G_UP, (i - 10) // 2))
    bwd_diff = [(k, rel_l2(snaps[1][k], snaps[0][k])) for k in border]
    print("run-to-run backward-tensor differences in backward order:",
          [(f"{'gout,dy,gpool,gup'.split(',')[{G_OUT: 0, DY: 1, G_POOL: 2, G_UP: 3}[k[0]]]}{k[1]}", f"{e:.1e}") for k, e in bwd_diff])
    bwd_diff = [(k, e) for k, e in bwd_diff if e > 0]
    worst = max((rel_l2(gr[1][k], gr[0][k]), k) for k in gr[0] if gr[0][k].abs().max() > 0)
    print("worst run-to-run parameter-gradient difference:", worst)
    # the activation-gradient chain has no atomics (fixed-order BN reductions): bit-identical run to run; weight
    # gradients are reduced with fp32 atomic adds across split-K CTAs: identical up to summation order
    assert not bwd_diff and worst[0] < 1e-4


def test_stages_with_the_operand_transform_option():
    """CARTSEG_XFORM=1 (convX.3 applies the BatchNorm + ReLU of convX.0 to its operand patch in shared memory, forward and
    weight gradient; the activation between the two convolutions is never stored): the same teacher-forced stage checks
    in a fresh process, because the option is read once per process.  Default is off (measured break-even)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, CARTSEG_XFORM="1")
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-x", "-m", "gpu", "--tb=short", "-k",
                        "teacher_forced and (3-96-80 or 2-224-224)"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "2 passed" in r.stdout, r.stdout[-2000:]


def test_stages_with_the_alternative_kernels_selected():
    """Every same-box A/B knob at its non-default value at once (round-1 pixel GEMM for the 3x3 convolutions, im2col +
    pointwise stem, wgrad_gemm for Cout = 64, stand-alone head kernels, scalar BN-backward kernels, channel_sum for the
    conv-transpose bias, ascending BN traversal): the alternatives the measurements in DESIGN.md were taken against stay
    correct — same teacher-forced stage checks, fresh process (the knobs are read once per process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, CARTSEG_CONV3="0", CARTSEG_STEM="0", CARTSEG_WGRAD9="0", CARTSEG_FUSE_HEAD="0",
               CARTSEG_POOL_PACKED="0", CARTSEG_UPBIAS_FUSED="0", CARTSEG_BN_REVERSE="0", CARTSEG_BN_REVERSE_BWD="0",
               CARTSEG_WGRAD_AFTER_UP="0")
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-x", "-m", "gpu", "--tb=short", "-k",
                        "teacher_forced and 3-96-80"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "1 passed" in r.stdout, r.stdout[-2000:]
