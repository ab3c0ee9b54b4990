"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares, the
host mirror keeps the reference's state-dict layout, thresholds map to exact logit bounds, and the
data-parallel gradient bucketing / averaging works across two gloo ranks."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "cartseg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from cartseg import _lib
    names = _header_functions()
    assert len(names) >= 20
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f"libcartseg.so does not export {n}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names        # the ctypes table covers the header, nothing else
    assert L.cs_version() >= 1


def test_plan_argument_validation_without_gpu():
    import ctypes as C
    from cartseg import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.cs_unet_plan_create(C.byref(h), 2, 3, 30, 32, 0) != 0          # H not a multiple of 16
    assert b"multiples of 16" in L.cs_last_error()
    assert L.cs_unet_plan_create(C.byref(h), 0, 3, 32, 32, 0) != 0
    assert L.cs_unet_plan_create(C.byref(h), 2, 3, 224, 224, 0) == 0
    train_bytes = L.cs_unet_plan_workspace_bytes(h)
    L.cs_unet_plan_destroy(h)
    assert L.cs_unet_plan_create(C.byref(h), 2, 3, 224, 224, 1) == 0
    infer_bytes = L.cs_unet_plan_workspace_bytes(h)
    L.cs_unet_plan_destroy(h)
    assert 0 < infer_bytes < train_bytes / 2
    # 23 backward stages cover the 82 parameters exactly once
    seen = []
    buf = (C.c_int * 8)()
    for s in range(_lib.NUM_BWD_STAGES):
        n = L.cs_unet_stage_params(s, buf, 8)
        seen += [buf[i] for i in range(n)]
    assert sorted(seen) == list(range(82))
    assert seen[:2] == [80, 81]                                             # the head finishes first


def test_module_state_dict_matches_reference_layout():
    import cartseg
    from oracle import unet_oracle as O
    m = cartseg.UNet()
    spec = O.state_dict_spec()
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    assert all(tuple(sd[k].shape) == s for k, s in spec)
    flat = m._flat_params()
    named = {id(p): k for k, p in m.named_parameters()}
    assert [named[id(p)] for p in flat] == O.param_keys(sd)                 # C-ABI order == state-dict order
    assert len(m._flat_buffers()) == 54
    # loading a reference-format checkpoint works with strict=True (create_testset.py:89)
    m.load_state_dict(O.synth_state_dict(seed=4), strict=True)
    # smp-style groups used by the training scripts (train_with_focalDice.py:384-391)
    for p in m.encoder.parameters():
        p.requires_grad = False
    assert m._frozen_encoder_convs(m._flat_params()) == 10
    assert all(p.requires_grad for p in m.decoder.parameters())
    assert sum(1 for _ in m.segmentation_head.parameters()) == 2


def test_no_cpu_fallback():
    import cartseg
    m = cartseg.UNet()
    with pytest.raises(cartseg.CartsegError):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(cartseg.CartsegError):
        cartseg.BCEDiceLoss()(torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))
    with pytest.raises(cartseg.CartsegError):
        cartseg.batch_sdf_from_masks(torch.zeros(1, 1, 16, 16))
    with pytest.raises(cartseg.CartsegError):
        cartseg.dice_metric(torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))


def test_no_cpu_fallback_next_rows():
    import cartseg
    z, t = torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16)
    for fn in (lambda: cartseg.ABL()(z, t), lambda: cartseg.BCEDiceABL()(z, t),
               lambda: cartseg.pseudo_label_qc(torch.zeros(1, 16, 16)),
               lambda: cartseg.clean_mask(torch.zeros(16, 16, dtype=torch.uint8)),
               lambda: cartseg.clean_mask_largest_component(torch.zeros(16, 16, dtype=torch.uint8)),
               lambda: cartseg.letterbox_resize_normalize([torch.zeros(8, 8, 3, dtype=torch.uint8)], 16),
               lambda: cartseg.resize_masks([torch.zeros(8, 8, dtype=torch.uint8)], 16),
               lambda: cartseg.ensemble_forward([lambda x: z], [1.0], torch.zeros(1, 3, 16, 16))):
        with pytest.raises(cartseg.CartsegError):
            fn()


def test_host_helpers_of_the_next_rows_match_the_oracle():
    """Host-side pieces that need no GPU: the C letterbox geometry (Python round() semantics), the ABL threshold ladder,
    the ensemble weight normalisation, should_accept."""
    import cartseg
    from cartseg import ops
    from oracle import abl_oracle as A
    from oracle import postproc_oracle as P
    from oracle import preproc_oracle as R
    for h, w in [(10, 25), (10, 35), (120, 50), (60, 96), (1080, 1920), (75, 101), (5, 5), (480, 645)]:
        for ratio in (0.1, 0.0, 0.25):
            assert cartseg.preproc.letterbox_geometry(h, w, ratio) == R.letterbox_geometry(h, w, ratio), (h, w, ratio)
    import numpy as np
    lad = np.asarray(ops.abl_eps_ladder(), dtype=np.float64).astype(np.float32)
    assert np.array_equal(lad, A.eps_ladder(len(lad)))
    w = cartseg.postproc.normalize_weights([0.7, 0.3, 2.0])
    assert w == (np.array([0.7, 0.3, 2.0], np.float32) / np.array([0.7, 0.3, 2.0], np.float32).sum()).tolist()
    for args in [(0.2, 0.9, 0.1), (0.004, 0.9, 0.1), (0.61, 0.9, 0.1), (0.2, 0.64, 0.1), (0.2, 0.9, 0.36)]:
        assert cartseg.should_accept(*args) == P.should_accept(*args)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cart-segmentation-unet_b200", "cartseg")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read().replace("the oracle", ""), f


@pytest.mark.parametrize("t", [0.05, 0.2, 0.35, 0.5, 0.65, 0.8, 0.95])
@pytest.mark.parametrize("ge", [False, True])
def test_logit_bound_is_the_exact_float32_threshold(t, ge):
    from cartseg import ops
    xb = ops.logit_bound(t, ge)
    x = torch.tensor([xb], dtype=torch.float32)
    below = torch.nextafter(x, torch.tensor([-float("inf")]))
    tt = torch.tensor(t, dtype=torch.float32)
    f = (lambda v: torch.sigmoid(v.repeat(16))[0] >= tt) if ge else (lambda v: torch.sigmoid(v.repeat(16))[0] > tt)
    assert bool(f(x)) and not bool(f(below))
    # and it reproduces sigmoid-then-compare on a dense sample around the boundary
    xs = x + torch.linspace(-1e-4, 1e-4, 20001)
    ref = (torch.sigmoid(xs) >= tt) if ge else (torch.sigmoid(xs) > tt)
    assert torch.equal(xs >= x, ref)


def test_plan_buckets_cover_all_stages():
    from cartseg.parallel import plan_buckets
    off = [0, 65, 37000, 110000, 118000, 200000, 500000, 510000, 900000, 5000000, 5000100]
    for be in (1, 50000, 10 ** 6, 10 ** 9):
        b = plan_buckets(off, be)
        assert b[0][0] == 0 and b[-1][1] == len(off) - 1
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert all(off[e] - off[s] >= be for s, e in b[:-1])


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], "cart-segmentation-unet_b200"))
from cartseg.parallel import GradSync, shard_batch
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
sync = GradSync(None, bucket_mb=0.001)
# a fake "backward": 6 stages, each fills its slice of the flat buffer with rank-dependent values
stage_off = [0, 10, 700, 705, 2000, 2600, 4000]
flat = torch.empty(stage_off[-1])
buckets = sync.stage_buckets(stage_off)
assert buckets[0][0] == 0 and buckets[-1][1] == 6
for s0, s1 in buckets:
    for s in range(s0, s1):
        flat[stage_off[s]:stage_off[s + 1]] = torch.arange(stage_off[s], stage_off[s + 1]).float() * (rank + 1) + s
    sync.reduce_async(flat[stage_off[s0]:stage_off[s1]])
sync.finish()
expect = torch.empty(stage_off[-1])
for s in range(6):
    idx = torch.arange(stage_off[s], stage_off[s + 1]).float()
    expect[stage_off[s]:stage_off[s + 1]] = sum(idx * (r + 1) + s for r in range(world)) / world
assert torch.allclose(flat, expect), (flat - expect).abs().max()
x = torch.arange(8 * 3).reshape(8, 3)
sh = shard_batch(x, rank, world)
assert sh.shape[0] == 8 // world and int(sh[0, 0]) == rank * (8 // world) * 3
# mean of equal-shard means == global mean (SURVEY 8e): the loss-averaging identity the DP path relies on
g = torch.Generator().manual_seed(0); v = torch.randn(8, 5, generator=g)
local = shard_batch(v, rank, world).mean().reshape(1)
dist.all_reduce(local); local /= world
assert torch.allclose(local, v.mean().reshape(1), atol=1e-6)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_data_parallel_gradient_average_two_gloo_ranks(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_bf16_storage_alone_moves_relu_network_gradients_by_tens_of_percent():
    """Why tests/test_gpu_unet_stages.py checks gradients stage by stage: with the CPU oracle ALONE, rounding the
    stored activations / weights to bfloat16 (fp32 arithmetic otherwise) leaves the loss unchanged to ~1e-4 but moves
    per-tensor gradients by >10 % (ReLU masks flip for pre-activations within the rounding noise of zero)."""
    from oracle import unet_oracle as O
    import cartseg
    x, t = O.synth_batch(2, 64, 64, seed=5)
    torch.manual_seed(0)
    sd0 = {k: v.detach().clone() for k, v in cartseg.UNet().state_dict().items()}
    res = []
    for emu in (False, True):
        sd = {k: v.clone() for k, v in sd0.items()}
        keys = O.param_keys(sd)
        for k in keys:
            sd[k].requires_grad_(True)
        loss = O.focal_dice_loss(O.unet_logits(x, sd, training=True, emulate_bf16=emu), t, 0.5, 2.0, 1.0, 0.7)
        loss.backward()
        res.append((loss.item(), sd["conv3.conv.0.weight"].grad.clone()))
    (l32, g32), (l16, g16) = res
    assert abs(l32 - l16) / l32 < 1e-3
    assert float((g16 - g32).norm() / g32.norm()) > 0.1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract's
    keys; without a GPU the product arm refuses to run instead of falling back."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "train_images_per_sec" and line["unit"] == "img/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    from oracle import ref_lift
    cb = line["cpu_baseline"]
    # the reference's own classes when its tree is present (build container), the pinned restatement elsewhere (GPU box)
    assert cb["kind"] == ("reference" if ref_lift.available() else "port") and cb["cores"] >= 1
    assert line["steps"] >= 10 and line["warmup"] >= 3                  # BASELINE.md §4 protocol
    assert cb["best_img_per_s"] >= cb["median_img_per_s"] > 0 and cb["cpu_model"]
    assert line["e2e"] == {"value": line["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                           text=True, timeout=600)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_default_bucket_policy_depends_on_world_size():
    """Overlapped 16 MB buckets up to 2 ranks, one all-reduce after the backward pass from 4 ranks on (measured on
    8 x B200: cartseg/parallel.py, profiles/r2_bench_k2_8gpu_bucket*mb.json); either way every stage is covered once."""
    from cartseg import parallel
    assert parallel.default_bucket_mb(1) == parallel.DEFAULT_BUCKET_MB == parallel.default_bucket_mb(2)
    assert parallel.default_bucket_mb(4) == parallel.default_bucket_mb(8) == parallel.LARGE_WORLD_BUCKET_MB
    stage_off = [0, 65, 65 + 36_928, 200_000, 5_000_000, 31_000_000]
    one = parallel.plan_buckets(stage_off, int(parallel.LARGE_WORLD_BUCKET_MB * (1 << 20) / 4))
    assert one == [(0, 5)]
    many = parallel.plan_buckets(stage_off, int(parallel.DEFAULT_BUCKET_MB * (1 << 20) / 4))
    assert many[0][0] == 0 and many[-1][1] == 5 and all(a[1] == b[0] for a, b in zip(many, many[1:]))
