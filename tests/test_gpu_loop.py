"""The reference's training-loop body, unmodified, on the cartseg objects:
train_bce_dice.py:328-338 — zero_grad / autocast forward / criterion / GradScaler backward / step / update / .item() —
plus the encoder-freeze + three-LR-group policy of src/train_with_focalDice.py:383-420 and
src/train_with_focalDice_unfrozen.py:388-392, and a validation pass in eval mode (train_bce_dice.py:343-352)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _loader(n, B, S, seed):
    from oracle import unet_oracle as O
    return [O.synth_batch(B, S, S, seed=seed + i) for i in range(n)]


@pytest.mark.parametrize("crit_name", ["bce_dice", "focal_dice", "composite"])
def test_reference_loop_body_runs_unmodified_and_learns(crit_name):
    import cartseg
    DEVICE = torch.device("cuda")
    torch.manual_seed(0)
    model = cartseg.UNet(in_channels=3, out_channels=1).to(DEVICE)
    criterion = {"bce_dice": cartseg.BCEDiceLoss(bce_weight=0.5),
                 "focal_dice": cartseg.FocalDiceLoss(alpha=0.5, gamma=2.0, w_focal=0.7),
                 "composite": cartseg.CompositeSegLoss(bce_weight=0.5, boundary_weight=0.3)}[crit_name]
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=4)
    scaler = torch.cuda.amp.GradScaler()
    train_loader = _loader(3, 4, 64, seed=0)
    val_loader = _loader(1, 4, 64, seed=50)
    epoch_losses = []
    for epoch in range(4):
        model.train()
        train_loss = 0.0
        for data, target in train_loader:                      # ---- train_bce_dice.py:328-338, verbatim
            data, target = data.to(DEVICE), target.to(DEVICE)
            optimizer.zero_grad()
            with torch.cuda.amp.autocast():
                logits = model(data)
                loss = criterion(logits, target)
            scaler.scale(loss).backward()
            scaler.step(optimizer)
            scaler.update()
            train_loss += loss.item()
        epoch_losses.append(train_loss / len(train_loader))
        model.eval()                                           # ---- :340-352
        val_dice = 0.0
        with torch.no_grad():
            for data, target in val_loader:
                data, target = data.to(DEVICE), target.to(DEVICE)
                logits = model(data)
                val_dice += cartseg.dice_metric(logits, target)
                cartseg.iou_metric(logits, target)
        scheduler.step()
    assert scaler.get_scale() >= 65536.0                       # no inf/nan step was ever skipped
    assert epoch_losses[-1] < epoch_losses[0] - 0.03, epoch_losses
    assert 0.0 < val_dice <= 1.0
    best_t, best_d = cartseg.find_best_threshold(model, val_loader, DEVICE)
    assert 0.2 <= best_t <= 0.8 and 0.0 <= best_d <= 1.0


def test_freeze_unfreeze_and_three_lr_groups():
    """src/train_with_focalDice.py:384-391 (freeze the encoder, train decoder + head), :413-419 (unfreeze at an epoch)
    and src/train_with_focalDice_unfrozen.py:388-392 (encoder LR x0.1, decoder LR, head LR x3)."""
    import cartseg
    DEVICE = torch.device("cuda")
    torch.manual_seed(1)
    model = cartseg.UNet().to(DEVICE)
    criterion = cartseg.FocalDiceLoss(alpha=0.5, gamma=2.0, w_focal=0.7)
    for p in model.encoder.parameters():
        p.requires_grad = False
    base_lr = 1e-3
    optimizer = torch.optim.AdamW([
        {"params": model.decoder.parameters(), "lr": base_lr},
        {"params": model.segmentation_head.parameters(), "lr": base_lr * 3.0},
    ], weight_decay=1e-4)
    enc_before = [p.detach().clone() for p in model.encoder.parameters()]
    dec_before = [p.detach().clone() for p in model.decoder.parameters()]
    data, target = _loader(1, 4, 64, seed=3)[0]
    data, target = data.to(DEVICE), target.to(DEVICE)
    model.train()
    for _ in range(2):
        optimizer.zero_grad()
        criterion(model(data), target).backward()
        optimizer.step()
    assert all(torch.equal(a, b) for a, b in zip(enc_before, model.encoder.parameters()))
    assert any(not torch.equal(a, b) for a, b in zip(dec_before, model.decoder.parameters()))
    assert all(p.grad is None for p in model.encoder.parameters())
    # unfreeze: add the encoder as a third group
    for p in model.encoder.parameters():
        p.requires_grad = True
    optimizer.add_param_group({"params": list(model.encoder.parameters()), "lr": base_lr * 0.1})
    optimizer.zero_grad()
    criterion(model(data), target).backward()
    optimizer.step()
    assert all(p.grad is not None for p in model.encoder.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(enc_before, model.encoder.parameters()))
    assert len(optimizer.param_groups) == 3


def test_checkpoint_roundtrip_with_reference_keys(tmp_path):
    """train_bce_dice.py:368-374 saves {'model_state_dict': ...}; create_testset.py:87-89 loads it with strict=True."""
    import cartseg
    from oracle import unet_oracle as O
    torch.manual_seed(2)
    model = cartseg.UNet().cuda()
    path = tmp_path / "dice_model_1.pth"
    torch.save({"epoch": 1, "model_state_dict": model.state_dict()}, path)
    ckpt = torch.load(path, map_location="cuda")
    state = ckpt.get("model_state_dict", ckpt)
    assert list(state.keys()) == [k for k, _ in O.state_dict_spec()]
    other = cartseg.UNet(final_sigmoid=True).cuda()
    other.load_state_dict(state, strict=True)
    other.eval()
    model.eval()
    x, _ = O.synth_batch(2, 64, 64, seed=4)
    with torch.no_grad():
        assert torch.allclose(other(x.cuda()), torch.sigmoid(model(x.cuda())))


def test_torch_compile_of_the_model_matches_eager():
    """create_pseudo_labels_gpu.py:164 wraps the model in torch.compile before inference."""
    import cartseg
    from oracle import unet_oracle as O
    torch.manual_seed(3)
    model = cartseg.UNet().cuda().eval()
    x, _ = O.synth_batch(2, 64, 64, seed=6)
    with torch.inference_mode():
        ref = model(x.cuda())
        compiled = torch.compile(model)
        got = compiled(x.cuda())
    assert torch.equal(got, ref)


def test_cuda_graph_capture_of_a_training_step_replays_identically():
    """nothing in the path synchronises or allocates behind the caller's back, so forward + loss +
    backward (incl. the two internal backward streams, forked / joined with events made at plan-bind time) capture into
    ONE CUDA graph.  The replay must reproduce the eager step: logits / loss / activation-gradient chain bit for bit,
    weight gradients up to the fp32 atomic ordering of the split-K accumulation; and it must follow new inputs and
    new weights (the bf16 re-pack is part of the graph)."""
    import cartseg
    from oracle import unet_oracle as O
    torch.manual_seed(4)
    model = cartseg.UNet().cuda().train()
    crit = cartseg.CompositeSegLoss(bce_weight=0.5, boundary_weight=0.3)      # exercises the EDT + fused loss too
    (x0, t0), (x1, t1) = _loader(2, 4, 64, seed=20)
    x0, t0, x1, t1 = x0.cuda(), t0.cuda(), x1.cuda(), t1.cuda()

    def eager(x, t):
        model.zero_grad(set_to_none=True)
        z = model(x)
        loss = crit(z, t)
        loss.backward()
        torch.cuda.synchronize()
        return z.detach().clone(), loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    step = cartseg.GraphedTrainStep(model, crit, x0, t0)
    n0 = cartseg.lib().cs_kernel_launch_count()
    for (x, t) in ((x0, t0), (x1, t1), (x0, t0)):
        loss_g = step(x, t).clone()
        torch.cuda.synchronize()
        g_graph = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        z_e, loss_e, g_e = eager(x, t)
        assert torch.equal(loss_g, loss_e), (float(loss_g), float(loss_e))
        for k in g_e:
            d = (g_graph[k] - g_e[k]).norm() / g_e[k].norm().clamp_min(1e-30)
            assert float(d) < 1e-4, (k, float(d))
        step.restore_grads()           # eager() replaced the parameters' .grad tensors: point them at the graph's again
    # weights change between replays: the graph re-packs them
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(0.9)
    loss_g = step(x1, t1).clone()
    _, loss_e, _ = eager(x1, t1)
    assert torch.equal(loss_g, loss_e)
    assert cartseg.lib().cs_kernel_launch_count() > n0


def test_cuda_graph_inference_matches_eager_and_is_one_launch():
    import cartseg
    from oracle import unet_oracle as O
    torch.manual_seed(5)
    model = cartseg.UNet().cuda().eval()
    x, _ = O.synth_batch(1, 224, 224, seed=2)
    xg = x.cuda()
    with torch.no_grad():
        ref_mask = cartseg.pseudo_label_mask(model(xg), 0.5)
        ref_logits = model(xg).clone()
    g_mask = cartseg.GraphedInference(model, xg, threshold=0.5)
    g_logits = cartseg.GraphedInference(model, xg)
    assert torch.equal(g_mask(xg).reshape(ref_mask.shape), ref_mask)
    assert torch.equal(g_logits(xg), ref_logits)
    x2, _ = O.synth_batch(1, 224, 224, seed=3)
    with torch.no_grad():
        assert torch.equal(g_logits(x2.cuda()), model(x2.cuda()))
    # parameter edits between replays are seen (the re-pack is inside the graph) ...
    with torch.no_grad():
        for p in model.parameters():
            p.data.mul_(0.5)
        assert torch.equal(g_logits(xg), model(xg))
    # ... unless the packs were frozen before capture (lowest latency)
    model.freeze_packed()
    g_frozen = cartseg.GraphedInference(model, xg)
    with torch.no_grad():
        a = g_frozen(xg).clone()
        assert torch.equal(a, model(xg))


def test_results_do_not_depend_on_the_persistent_grid_size():
    """Race / pipeline-phase check that needs no external tool (compute-sanitizer is closed on this pool): the persistent
    GEMM kernels take their tiles in a static round-robin, so capping the grid changes the tile -> CTA assignment, the
    number of tiles per CTA, and with them every ring phase / accumulator parity / staging-buffer rotation of every CTA.
    (1) single layers through the C ABI (CARTSEG_LAYER_SMS): conv fprop and dgrad outputs bit-identical for every grid
    size; (2) the whole eval-mode forward (no batch statistics: nothing may depend on the partition) bit-identical;
    (3) a training step: the per-CTA fp32 partial sums of the BN statistics do depend on the partition (last-bit changes
    of scale / shift), so loss to 1e-5 and gradients to cosine 0.999."""
    import os
    import cartseg
    from cartseg import _lib
    from gpu_util import layer_scratch, stream, to_nhwc_bf16
    L = _lib.lib()
    g = torch.Generator().manual_seed(11)
    for (B, H, W, Cin, Cout) in [(4, 56, 56, 64, 64), (2, 28, 28, 256, 128), (2, 14, 14, 512, 512), (3, 40, 24, 128, 64)]:
        x = to_nhwc_bf16(torch.randn(B, Cin, H, W, generator=g))
        dy = to_nhwc_bf16(torch.randn(B, Cout, H, W, generator=g))
        w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).cuda()
        buf, scratch = layer_scratch(Cin, Cout)
        outs = []
        for sms in ("148", "146", "96", "36", "2"):
            os.environ["CARTSEG_LAYER_SMS"] = sms
            y = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
            dx = torch.full((B, H, W, Cin), float("nan"), dtype=torch.bfloat16, device="cuda")
            ssum = torch.zeros(Cout, dtype=torch.float64, device="cuda")
            ssq = torch.zeros(Cout, dtype=torch.float64, device="cuda")
            _lib.check(L.cs_conv3x3_fprop(x.data_ptr(), B, H, W, Cin, w.data_ptr(), Cout, y.data_ptr(), ssum.data_ptr(),
                                          ssq.data_ptr(), scratch, stream()), "cs_conv3x3_fprop")
            _lib.check(L.cs_conv3x3_dgrad(dy.data_ptr(), B, H, W, Cin, w.data_ptr(), Cout, dx.data_ptr(), scratch, stream()),
                       "cs_conv3x3_dgrad")
            torch.cuda.synchronize()
            outs.append((y, dx, ssum))
        os.environ.pop("CARTSEG_LAYER_SMS")
        for y, dx, ssum in outs[1:]:
            assert torch.equal(y.view(torch.int16), outs[0][0].view(torch.int16)), (B, H, W, Cin, Cout)
            assert torch.equal(dx.view(torch.int16), outs[0][1].view(torch.int16)), (B, H, W, Cin, Cout)
            assert torch.allclose(ssum, outs[0][2], rtol=1e-5, atol=1e-3)
    torch.manual_seed(6)
    model = cartseg.UNet().cuda()
    crit = cartseg.FocalDiceLoss(0.5, 2.0, 1.0, 0.7)
    (x, t), = _loader(1, 3, 96, seed=30)
    xg, tg = x.cuda(), t.cuda()
    ref_eval = ref_train = None
    model.eval()
    for reserve in (0, 2, 52, 112, 146):                 # eval first: training steps move the running statistics
        model._reserve_sms = reserve
        with torch.no_grad():
            ze = model(xg).clone()
        if ref_eval is None:
            ref_eval = ze
        assert torch.equal(ze, ref_eval), reserve
    model.train()
    for reserve in (0, 2, 52, 112, 146):
        model._reserve_sms = reserve
        model.zero_grad(set_to_none=True)
        loss = crit(model(xg), tg)
        loss.backward()
        torch.cuda.synchronize()
        flat = torch.cat([p.grad.flatten() for p in model.parameters()]).double()
        if ref_train is None:
            ref_train = (float(loss), flat)
            continue
        assert abs(float(loss) - ref_train[0]) <= 1e-5 * abs(ref_train[0]), reserve
        cos = float(torch.dot(flat, ref_train[1]) / (flat.norm() * ref_train[1].norm()))
        assert cos > 0.999, (reserve, cos)
    model._reserve_sms = 0
