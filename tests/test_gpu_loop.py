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
