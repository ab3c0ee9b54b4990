"""GPU parity of the whole U-Net op (cartseg.UNet -> cartseg::unet_forward/backward -> cs_unet_*) against
the CPU oracle (a restatement of src/create_testset.py:40-83 pinned to the reference's own outputs).
Loss 1e-2 relative and inference-mask Dice >= 0.999 vs the fp32 oracle (north star); gradients are held to the
3e-2 bar kernel by kernel in tests/test_gpu_unet_stages.py and compared end to end here (see
test_train_step_vs_oracle for why the end-to-end comparison of a ReLU network cannot be per-tensor tight)."""
import numpy as np
import pytest
import torch

from gpu_util import load_golden, rel_l2

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-2
GRAD_TOL = 3e-2


def _model(sd, **kw):
    import cartseg
    m = cartseg.UNet(**kw)
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return m.cuda()


def _oracle_train(O, x, tgt, sd, loss_fn, emulate_bf16=False):
    sd = {k: v.clone() for k, v in sd.items()}
    keys = O.param_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    z = O.unet_logits(x, sd, training=True, emulate_bf16=emulate_bf16)
    loss = loss_fn(z, tgt)
    loss.backward()
    return z.detach(), loss.item(), {k: sd[k].grad for k in keys}, sd


def test_state_dict_keys_and_groups():
    import cartseg
    from oracle import unet_oracle as O
    m = cartseg.UNet()
    spec = O.state_dict_spec()
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    assert all(tuple(sd[k].shape) == s for k, s in spec)
    n_all = sum(p.numel() for p in m.parameters())
    assert n_all == 31043521
    n_grp = sum(p.numel() for g in (m.encoder, m.decoder, m.segmentation_head) for p in g.parameters())
    assert n_grp == n_all


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (1, 48, 16), (4, 64, 64), (2, 224, 224)])
def test_eval_forward_vs_oracle(B, H, W):
    from oracle import unet_oracle as O
    x, _ = O.synth_batch(B, H, W, seed=5)
    sd = O.synth_state_dict(seed=1)
    with torch.no_grad():
        ref = O.unet_logits(x, {k: v.clone() for k, v in sd.items()}, training=False)
    m = _model(sd).eval()
    with torch.no_grad():
        got = m(x.cuda()).cpu()
    assert got.shape == ref.shape and got.dtype == torch.float32
    err = rel_l2(got, ref)
    print(f"eval logits rel-L2 {err:.3e}")
    assert err < 2e-2
    # inference masks: Dice(new, ref) >= 0.999 over pixels that are not within bf16 noise of the threshold
    a, b = (torch.sigmoid(got) >= 0.5), (torch.sigmoid(ref) >= 0.5)
    if H * W * B >= 4096 and b.any():
        dice = 2.0 * (a & b).sum().item() / (a.sum().item() + b.sum().item())
        print(f"mask dice {dice:.5f}")
        assert dice >= 0.999


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_eval_forward_vs_reference_golden(tag):
    from oracle import unet_oracle as O
    g = load_golden("model.npz")
    B, C, H, W = (int(v) for v in g[f"{tag}_shape"])
    x, _ = O.synth_batch(B, H, W, seed=5)
    m = _model(O.synth_state_dict(seed=1)).eval()
    with torch.no_grad():
        got = m(x.cuda()).cpu()
    assert rel_l2(got, torch.from_numpy(g[f"{tag}_eval_logits"])) < 2e-2


def _torch_init_state_dict(seed):
    import cartseg
    torch.manual_seed(seed)                 # the reference's default initialisation: UNet() under a seed
    return {k: v.detach().clone() for k, v in cartseg.UNet().state_dict().items()}


def _global_grad_agreement(g_gpu, g_ref):
    keys = [k for k in g_ref if not (k.endswith(".conv.0.bias") or k.endswith(".conv.3.bias"))]
    a = torch.cat([g_gpu[k].double().flatten() for k in keys])
    b = torch.cat([g_ref[k].double().flatten() for k in keys])
    return float((a - b).norm() / b.norm()), float(torch.dot(a, b) / (a.norm() * b.norm()))


@pytest.mark.parametrize("B,H,W,loss", [(4, 64, 64, "bce_dice"), (2, 224, 224, "focal_dice"), (2, 96, 160, "composite")])
def test_train_step_vs_oracle(B, H, W, loss):
    """One training step against the oracle.  Loss: <= 1e-2 relative vs the fp32 oracle (north star).  Gradients:
    every kernel is held to <= 4e-3 on its actual inputs by tests/test_gpu_unet_stages.py (north-star bar 3e-2);
    END-TO-END a ReLU network turns any 16-bit activation storage into mask flips (the oracle's own bf16 emulation
    differs from its fp32 run by 15-45 % per tensor, tests/test_host_cpu.py), so here the whole gradient is compared
    with the bf16-emulating oracle and with the fp32 oracle by direction and overall size."""
    import cartseg
    from oracle import unet_oracle as O
    x, tgt = O.synth_batch(B, H, W, seed=5)
    sd = _torch_init_state_dict(0)
    ref_fn, crit = {
        "bce_dice": (lambda z, t: O.bce_dice_loss(z, t), cartseg.BCEDiceLoss()),
        "focal_dice": (lambda z, t: O.focal_dice_loss(z, t, 0.5, 2.0, 1.0, 0.7), cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)),
        "composite": (lambda z, t: O.composite_seg_loss(z, t, 0.5, 0.3), cartseg.CompositeSegLoss(0.5, 0.3)),
    }[loss]
    z_ref, loss_ref, g_ref, sd_after = _oracle_train(O, x, tgt, sd, ref_fn)
    z_emu, loss_emu, g_emu, _ = _oracle_train(O, x, tgt, sd, ref_fn, emulate_bf16=True)

    m = _model(sd).train()
    z = m(x.cuda())
    out = crit(z, tgt.cuda())
    out.backward()
    torch.cuda.synchronize()

    e_logits = rel_l2(z.detach().cpu(), z_ref)
    e_loss = abs(out.item() - loss_ref) / abs(loss_ref)
    print(f"train logits rel-L2 vs fp32 {e_logits:.3e} (vs bf16-emulating {rel_l2(z.detach().cpu(), z_emu):.3e}); "
          f"loss rel {e_loss:.3e}")
    assert e_logits < 3e-2
    assert e_loss < LOSS_TOL
    assert abs(out.item() - loss_emu) / abs(loss_emu) < LOSS_TOL
    g_gpu = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    for k, g in g_gpu.items():
        assert torch.isfinite(g).all(), k
    err_f, cos_f = _global_grad_agreement(g_gpu, g_ref)
    err_e, cos_e = _global_grad_agreement(g_gpu, g_emu)
    err_oo, cos_oo = _global_grad_agreement(g_emu, g_ref)
    print(f"whole-gradient rel-L2 / cosine: GPU vs fp32 oracle {err_f:.3f} / {cos_f:.4f}; GPU vs bf16-emulating oracle "
          f"{err_e:.3f} / {cos_e:.4f}; bf16-emulating vs fp32 oracle (CPU only) {err_oo:.3f} / {cos_oo:.4f}")
    # The bars DESIGN.md §1 states.  Measured on B200 at these worst-conditioned sizes (B <= 4, noise images, default
    # init): cosine 0.9715 / 0.9906 / 0.9937 vs fp32 and 0.9902 / 0.9966 / 0.9979 vs the bf16-emulating oracle; the
    # bars sit half a percent under the worst case because every change of summation order re-rolls the mask flips.
    assert cos_f >= 0.965 and cos_e >= 0.985, (cos_f, cos_e)
    assert err_f <= 1.1 * err_oo + 0.01, (err_f, err_oo)   # no worse than what bf16 storage alone does to the oracle
    # per tensor, so that one wrong tensor cannot hide in the global norm: the head and the last decoder block meet the
    # north-star 3e-2 against the bf16-emulating oracle; every other tensor stays within what the oracle's own bf16
    # emulation loses against its fp32 run on that tensor (x2 + 5e-2: two roundings per stored tensor on the GPU path
    # — y and the activation — where the emulation has the same two; mask flips differ per realisation)
    worst = []
    for k in g_ref:
        if k.endswith(".conv.0.bias") or k.endswith(".conv.3.bias"):
            assert g_gpu[k].abs().max().item() == 0.0, k          # exactly zero: BN removes the conv bias
            continue
        e_emu, e_oo = rel_l2(g_gpu[k], g_emu[k]), rel_l2(g_emu[k], g_ref[k])
        c = float(torch.dot(g_gpu[k].double().flatten(), g_emu[k].double().flatten())
                  / (g_gpu[k].double().norm() * g_emu[k].double().norm()).clamp_min(1e-300))
        worst.append((e_emu, e_oo, c, k))
        if k.startswith("final_conv."):
            assert e_emu < GRAD_TOL, (k, e_emu)
        if k.startswith("upconv") and k.endswith(".bias"):
            # pixel sum of an activation gradient that cancels to ~1e-3 of its terms: the bf16 storage of GRADIENTS,
            # which the emulation (fp32 gradients) does not have, shows here; measured 2e-2 ... 1.3e-1
            assert e_emu <= max(1.0 * e_oo + 3e-2, 0.3), (k, e_emu, e_oo)
            continue
        assert e_emu <= 1.0 * e_oo + 3e-2, (k, e_emu, e_oo)     # measured: e_emu ~ 0.63 e_oo on the worst tensors
        assert c >= 0.93, (k, c)                                 # measured worst 0.951 (conv5.conv.1.bias, B = 2)
    worst.sort(reverse=True)
    print("per-tensor worst (rel-L2 GPU vs emu, emu vs fp32, cosine):",
          [(k, f"{a:.3f}", f"{b:.3f}", f"{c:.4f}") for a, b, c, k in worst[:5]])
    # BN running statistics were updated in place (momentum 0.1, unbiased variance)
    bufs = dict(m.named_buffers())
    for k, v in sd_after.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert rel_l2(bufs[k].cpu(), v) < 2e-2, k
        if k.endswith("num_batches_tracked"):
            assert int(bufs[k].item()) == int(v.item()) == 1


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_train_loss_vs_reference_golden(tag):
    """The reference's own UNet + BCEDiceLoss on these inputs (oracle/make_golden.py): train-mode logits and loss.
    (The cases are tiny — the bottleneck BN sees 3..8 values per channel — so gradients are checked elsewhere.)"""
    import cartseg
    from oracle import unet_oracle as O
    g = load_golden("model.npz")
    B, C, H, W = (int(v) for v in g[f"{tag}_shape"])
    x, tgt = O.synth_batch(B, H, W, seed=5)
    m = _model(O.synth_state_dict(seed=1)).train()
    z = m(x.cuda())
    loss = cartseg.BCEDiceLoss()(z, tgt.cuda())
    loss.backward()
    assert abs(loss.item() - float(g[f"{tag}_train_loss"])) / float(g[f"{tag}_train_loss"]) < LOSS_TOL
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    if tag == "c":
        # 3 x 64 x 96: the head's gradients against the REFERENCE's own (fp32) gradients, north-star bar 3e-2
        named = dict(m.named_parameters())
        for k in ("final_conv.weight", "final_conv.bias"):
            assert rel_l2(named[k].grad.cpu(), torch.from_numpy(g[f"{tag}_gfull/{k}"])) < GRAD_TOL, k


def test_frozen_encoder_and_param_groups():
    """src/train_with_focalDice.py:384-391: encoder frozen -> only decoder + head receive gradients, and they
    equal the gradients of the unfrozen run."""
    import cartseg
    from oracle import unet_oracle as O
    x, tgt = O.synth_batch(2, 96, 96, seed=7)
    sd = _torch_init_state_dict(2)
    crit = cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)
    from cartseg import ops
    full = _model(sd).train()
    crit(full(x.cuda()), tgt.cuda()).backward()
    torch.cuda.synchronize()
    # Fresh plan with a POISONED workspace: a gradient buffer the frozen run forgets to write must not be rescued by
    # what the unfrozen run left behind (round 1 skipped dconv4.0's dgrad when the whole encoder was frozen — the only
    # writer of the buffer upconv4's weight gradient reads).
    ops.release_plans()
    plan = ops.get_plan(2, 3, 96, 96, torch.device("cuda"), inference_only=False)
    plan.workspace.view(torch.int16).fill_(0x7FC0)            # bf16 NaN everywhere (fp32: NaN too)
    frozen = _model(sd).train()
    for p in frozen.encoder.parameters():
        p.requires_grad = False
    crit(frozen(x.cuda()), tgt.cuda()).backward()
    enc = {id(p) for p in frozen.encoder.parameters()}
    fp = dict(full.named_parameters())
    for k, p in frozen.named_parameters():
        if id(p) in enc:
            assert p.grad is None
        else:
            assert rel_l2(p.grad, fp[k].grad) < 1e-4, k
    # the same frozen step against the oracle (frozen from step 0, nothing warm): decoder + head gradients
    z_ref, loss_ref, g_ref, _ = _oracle_train(O, x, tgt, sd, lambda z, t: O.focal_dice_loss(z, t, 0.5, 2.0, 1.0, 0.7),
                                              emulate_bf16=True)
    for k, p in frozen.named_parameters():
        if id(p) not in enc and not (k.endswith(".conv.0.bias") or k.endswith(".conv.3.bias")):
            assert torch.isfinite(p.grad).all(), k
            c = float(torch.dot(p.grad.cpu().double().flatten(), g_ref[k].double().flatten())
                      / (p.grad.double().norm().cpu() * g_ref[k].double().norm()))
            assert c > 0.95, (k, c)
    # half-frozen: conv1..conv3 frozen, gradients of everything above unchanged
    ops.release_plans()
    plan = ops.get_plan(2, 3, 96, 96, torch.device("cuda"), inference_only=False)
    plan.workspace.view(torch.int16).fill_(0x7FC0)
    half = _model(sd).train()
    for mod in (half.conv1, half.conv2, half.conv3):
        for p in mod.parameters():
            p.requires_grad = False
    crit(half(x.cuda()), tgt.cuda()).backward()
    for k, p in half.named_parameters():
        if k.startswith(("conv1.", "conv2.", "conv3.")):
            assert p.grad is None
        else:
            assert rel_l2(p.grad, fp[k].grad) < 1e-4, k
    # non-prefix freezing (decoder frozen, encoder + head trainable): the frozen tensors get no gradient and no
    # weight-gradient GEMM, everything else is unchanged
    ops.release_plans()
    plan = ops.get_plan(2, 3, 96, 96, torch.device("cuda"), inference_only=False)
    plan.workspace.view(torch.int16).fill_(0x7FC0)
    dec = _model(sd).train()
    for p in dec.decoder.parameters():
        p.requires_grad = False
    n0 = cartseg.lib().cs_kernel_launch_count()
    crit(dec(x.cuda()), tgt.cuda()).backward()
    torch.cuda.synchronize()
    n_dec = cartseg.lib().cs_kernel_launch_count() - n0
    frozen_ids = {id(p) for p in dec.decoder.parameters()}
    for k, p in dec.named_parameters():
        if id(p) in frozen_ids:
            assert p.grad is None
        else:
            assert rel_l2(p.grad, fp[k].grad) < 1e-4, k
    n0 = cartseg.lib().cs_kernel_launch_count()
    full.zero_grad(set_to_none=True)
    crit(full(x.cuda()), tgt.cuda()).backward()
    torch.cuda.synchronize()
    assert n_dec < cartseg.lib().cs_kernel_launch_count() - n0       # 12 wgrad GEMMs (+ unpacks) were not launched


def test_frozen_head_only():
    """segmentation_head frozen (final_conv.weight / bias without gradients): everything else unchanged."""
    import cartseg
    from oracle import unet_oracle as O
    x, tgt = O.synth_batch(2, 64, 64, seed=8)
    sd = _torch_init_state_dict(4)
    crit = cartseg.BCEDiceLoss()
    full = _model(sd).train()
    crit(full(x.cuda()), tgt.cuda()).backward()
    fp = dict(full.named_parameters())
    m = _model(sd).train()
    m.segmentation_head.requires_grad_(False)
    crit(m(x.cuda()), tgt.cuda()).backward()
    for k, p in m.named_parameters():
        if k.startswith("final_conv."):
            assert p.grad is None
        else:
            assert rel_l2(p.grad, fp[k].grad) < 1e-4, k


def test_weight_tied_multi_step_drift():
    """Equivalence-by-training in the style of the reference's only numerical check
    (src/training/losses/label_smooth.py:216-259): tie weights, run the same SGD steps on identical batches in
    both implementations, compare the loss trajectory and the accumulated parameter update."""
    import cartseg
    from oracle import unet_oracle as O
    sd0 = _torch_init_state_dict(3)
    m = _model(sd0).train()
    opt = torch.optim.SGD(m.parameters(), lr=1e-2)
    crit = cartseg.BCEDiceLoss()
    sd = {k: v.clone() for k, v in sd0.items()}
    keys = O.param_keys(sd)
    losses = []
    for step in range(4):
        x, tgt = O.synth_batch(4, 64, 64, seed=100 + step)
        for k in keys:                                           # oracle step
            sd[k] = sd[k].detach().requires_grad_(True)
        lo = O.bce_dice_loss(O.unet_logits(x, sd, training=True), tgt)
        lo.backward()
        with torch.no_grad():
            for k in keys:
                sd[k] = sd[k] - 1e-2 * sd[k].grad
        opt.zero_grad()                                          # GPU step
        lg = crit(m(x.cuda()), tgt.cuda())
        lg.backward()
        opt.step()
        losses.append((lo.item(), lg.item()))
    print("loss trajectory (oracle, GPU):", losses)
    for lo, lg in losses:
        assert abs(lo - lg) / abs(lo) < LOSS_TOL, losses
    named = dict(m.named_parameters())
    upd_gpu = {k: named[k].detach().cpu() - sd0[k] for k in keys}
    upd_ref = {k: sd[k].detach() - sd0[k] for k in keys}
    err, cos = _global_grad_agreement(upd_gpu, upd_ref)
    print(f"accumulated update after 4 steps: rel-L2 {err:.3f}, cosine {cos:.4f}")
    assert cos >= 0.99                                    # measured 0.998


def test_fused_optimizer_updates_reach_the_kernels():
    """torch's fused AdamW updates parameters WITHOUT bumping their version counters: the bf16 weight packs must
    still follow (training forwards always re-pack; eval plans re-pack after a training step)."""
    import cartseg
    from oracle import unet_oracle as O
    x, tgt = O.synth_batch(2, 64, 64, seed=9)
    m = _model(_torch_init_state_dict(5)).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2, fused=True)
    crit = cartseg.BCEDiceLoss()
    with torch.no_grad():
        m.eval()
        before = m(x.cuda()).clone()
        m.train()
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss = crit(m(x.cuda()), tgt.cuda())
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.02, losses                  # the same batch six times: the loss must fall
    m.eval()
    with torch.no_grad():
        after = m(x.cuda())
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        ref = O.unet_logits(x, sd, training=False)
    assert rel_l2(after.cpu(), ref) < 2e-2                        # eval uses the UPDATED weights
    assert rel_l2(after, before) > 0.05


def test_eval_packs_follow_data_updates_and_model_recreation():
    """(1) in-place updates through ``p.data`` (EMA / SWA / checkpoint averaging) do not bump version
    counters, and such a model never runs a training forward; (2) a model rebuilt in a loop can land on the freed
    model's addresses with identical versions.  Eval-mode forwards must see the current weights in both cases."""
    import cartseg
    from oracle import unet_oracle as O
    x, _ = O.synth_batch(1, 32, 32, seed=3)
    xg = x.cuda()
    m = _model(O.synth_state_dict(seed=1)).eval()
    with torch.no_grad():
        a = m(xg).clone()
        for p in m.parameters():
            p.data.mul_(0.5)                              # version counters unchanged
        b = m(xg).clone()
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        ref = O.unet_logits(x, sd, training=False)
    assert rel_l2(b.cpu(), ref) < 2e-2 and rel_l2(b, a) > 0.05
    # opt-in: packs frozen -> an edit of the PACKED tensors (3x3 conv / conv-transpose weights; BN parameters, biases and
    # the 1x1 head are read live) is by contract NOT seen until freeze_packed is called again
    m.freeze_packed()
    packed = [p for k, p in m.named_parameters() if p.dim() == 4 and not k.startswith("final_conv")]
    with torch.no_grad():
        c0 = m(xg).clone()
        for p in packed:
            p.data.mul_(2.0)
        c1 = m(xg).clone()
        assert torch.equal(c0, c1)
        m.freeze_packed()
        c2 = m(xg).clone()
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        ref = O.unet_logits(x, sd, training=False)
    assert rel_l2(c2.cpu(), ref) < 2e-2 and not torch.equal(c2, c0)
    # checkpoint loop: same addresses, same versions, different weights
    outs = []
    for seed in (1, 2, 3):
        mm = _model(O.synth_state_dict(seed=seed)).eval()
        with torch.no_grad():
            outs.append(mm(xg).clone())
        ref = O.unet_logits(x, O.synth_state_dict(seed=seed), training=False)
        assert rel_l2(outs[-1].cpu(), ref) < 2e-2, seed
        del mm
    assert rel_l2(outs[1], outs[0]) > 1e-3 and rel_l2(outs[2], outs[1]) > 1e-3


def test_chunked_eval_forward_equals_one_piece():
    """Eval batches above UNet.EVAL_CHUNK run chunk by chunk (ragged last chunk included): bit-identical logits."""
    import cartseg
    from oracle import unet_oracle as O
    x, _ = O.synth_batch(10, 32, 48, seed=4)
    m = _model(O.synth_state_dict(seed=1)).eval()
    with torch.no_grad():
        whole = m(x.cuda()).clone()
        m.EVAL_CHUNK = 4                                   # 4 + 4 + 2
        chunked = m(x.cuda())
    assert chunked.shape == whole.shape and torch.equal(chunked, whole)


def test_per_group_bn_mode_and_eval_backward_raise_clearly():
    import cartseg
    from cartseg import CartsegError
    from oracle import unet_oracle as O
    x, _ = O.synth_batch(1, 32, 32, seed=1)
    m = _model(O.synth_state_dict(seed=1)).train()
    m.encoder.eval()                                       # the usual way to freeze BN statistics: unsupported, loud
    with pytest.raises(CartsegError, match="single BatchNorm mode"):
        m(x.cuda())
    m.train()
    m(x.cuda())
    m.eval()
    z = m(x.cuda())                                        # grad mode on, eval mode
    with pytest.raises(CartsegError, match="eval-mode forward"):
        z.sum().backward()
    # dtype / device conversions after a forward are re-validated on every call
    m.half()
    with pytest.raises(CartsegError, match="float32"):
        with torch.no_grad():
            m(x.cuda())


def test_backward_after_overwritten_forward_raises():
    import cartseg
    from oracle import unet_oracle as O
    x, tgt = O.synth_batch(1, 32, 32, seed=1)
    m = _model(O.synth_state_dict(seed=1)).train()
    z1 = m(x.cuda())
    z2 = m(x.cuda())
    with pytest.raises(Exception):
        z1.sum().backward()
    z2.sum().backward()


def test_rejects_cpu_input_and_bad_shapes():
    import cartseg
    m = cartseg.UNet()
    with pytest.raises(Exception):
        m(torch.zeros(1, 3, 32, 32))
    m = m.cuda()
    with pytest.raises(Exception):
        m(torch.zeros(1, 3, 30, 32, device="cuda"))
    with pytest.raises(Exception):
        m(torch.zeros(1, 4, 32, 32, device="cuda"))


@pytest.mark.parametrize("cin", [1, 4, 7])
def test_other_input_channel_counts(cin):
    """UNet(in_channels=...) (src/create_testset.py:54): the first convolution's generic path (any 1..7 channels; the
    benchmarked 3-channel path is a compile-time specialisation) — eval logits and one training step against the oracle."""
    import cartseg
    from oracle import unet_oracle as O
    torch.manual_seed(40 + cin)
    sd = {k: v.detach().clone() for k, v in cartseg.UNet(in_channels=cin).state_dict().items()}
    g = torch.Generator().manual_seed(cin)
    x = torch.randn(2, cin, 48, 32, generator=g)
    tgt = (torch.rand(2, 1, 48, 32, generator=g) > 0.6).float()
    with torch.no_grad():
        ref = O.unet_logits(x, {k: v.clone() for k, v in sd.items()}, training=False)
    m = _model(sd, in_channels=cin)
    with torch.no_grad():
        got = m.eval()(x.cuda()).cpu()
    assert rel_l2(got, ref) < 2e-2
    m.train()
    crit = cartseg.BCEDiceLoss()
    loss = crit(m(x.cuda()), tgt.cuda())
    loss.backward()
    z_ref = O.unet_logits(x, {k: v.clone() for k, v in sd.items()}, training=True)
    assert abs(loss.item() - float(O.bce_dice_loss(z_ref, tgt))) < 1e-2
    gw = m.conv1.conv[0].weight.grad
    assert gw is not None and torch.isfinite(gw).all() and gw.abs().max() > 0 and tuple(gw.shape) == (64, cin, 3, 3)
    # the first convolution's weight gradient (im2col of the cin-channel image on the side stream): direction against the
    # fp32 oracle (end-to-end gradients of this ReLU network are ill-conditioned, DESIGN.md section 1: cosine, not rel-L2)
    g_ref = _oracle_train(O, x, tgt, sd, O.bce_dice_loss)[2]["conv1.conv.0.weight"]
    cos = float(torch.dot(gw.cpu().double().flatten(), g_ref.double().flatten()) / (gw.cpu().double().norm() * g_ref.double().norm()))
    print(f"cin={cin}: conv1.conv.0.weight gradient cosine vs fp32 oracle {cos:.4f}")
    assert cos > 0.9
