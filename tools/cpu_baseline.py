#!/usr/bin/env python
"""The CPU-baseline plan of BASELINE.md §4 in one go, on the host cores of the box it runs on (run it in the SAME gpurun
invocation as the GPU numbers it stands beside).  Prints one JSON document.

  K1   fwd+bwd, B=4, 3x224x224, BCE+Dice, fp32 — >= 3 warm-ups, >= 10 timed steps, best + median; and with AdamW
  focal-Dice and Composite(boundary, scipy EDT) steps at B=4
  signed_distance_map_np (two scipy EDTs) per mask at 224^2 and 512^2, one thread
  eval-mode forward at B in {1, 8, 64}

The reference's own classes are used when /root/reference is present (kind "reference"), else the restatement in
oracle/ (kind "port") — which tests/test_oracle_golden.py pins to the reference's outputs.
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import cpu_model_string, cpu_reference_step_rate  # noqa: E402
from oracle import edt_oracle, ref_lift  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402


def timed(fn, warmup, reps):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return {"best_ms": 1e3 * min(ts), "median_ms": 1e3 * statistics.median(ts), "reps": reps, "warmup": warmup}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--quick", action="store_true", help="fewer repetitions (smoke run)")
    args = ap.parse_args()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    steps = 3 if args.quick else max(args.steps, 10)
    out = {"cpu_model": cpu_model_string(), "cores": cores, "torch_threads": torch.get_num_threads(),
           "torch": torch.__version__, "kind": "reference" if ref_lift.available() else "port"}

    def strip(r):
        return {k: r[k] for k in ("value", "best_img_per_s", "median_img_per_s", "best_ms", "median_ms", "kind", "sample")}

    out["k1_bce_dice_fwd_bwd"] = strip(cpu_reference_step_rate("bce_dice", 224, steps, 3))
    out["k1_bce_dice_fwd_bwd_adamw"] = strip(cpu_reference_step_rate("bce_dice", 224, steps, 3, optimizer=True))
    out["focal_dice_fwd_bwd"] = strip(cpu_reference_step_rate("focal_dice", 224, steps, 3))
    out["composite_boundary_fwd_bwd"] = strip(cpu_reference_step_rate("composite", 224, steps, 3))

    # ---- scipy SDF per mask (src/train_with_boundary_loss.py:191-202), single thread as the reference calls it
    if ref_lift.available():
        sdf = ref_lift.lift("src/train_with_boundary_loss.py", ["signed_distance_map_np"])["signed_distance_map_np"]
    else:
        from scipy.ndimage import distance_transform_edt

        def sdf(mask):                      # restatement of :191-202 (scipy does the arithmetic either way)
            mask = mask.astype(bool)
            if mask.all() or not mask.any():
                return np.zeros(mask.shape, np.float32)
            return (distance_transform_edt(~mask) - distance_transform_edt(mask)).astype(np.float32)
    out["sdf_scipy_per_mask"] = {}
    for S in (224, 512):
        _, m = O.synth_batch(1, S, S, seed=3)
        mask = m[0, 0].numpy() > 0.5
        r = timed(lambda: sdf(mask), 2, 5 if args.quick else 20)
        out["sdf_scipy_per_mask"][str(S)] = r
    # ---- eval-mode forward
    out["eval_forward"] = {}
    if ref_lift.available():
        torch.manual_seed(0)
        net, logits_of, _ = ref_lift.reference_model_and_loss("bce_dice")
        net.eval()
        fwd = lambda x: logits_of(x)                                      # noqa: E731
    else:
        sd = O.synth_state_dict(seed=0)
        fwd = lambda x: O.unet_logits(x, sd, training=False)              # noqa: E731
    for B in (1, 8) if args.quick else (1, 8, 64):
        x, _ = O.synth_batch(B, 224, 224, seed=1)
        with torch.no_grad():
            r = timed(lambda: fwd(x), 1 if B == 64 else 3, 3 if (B == 64 or args.quick) else 10)
        r["img_per_s_best"] = B / (r["best_ms"] / 1e3)
        r["img_per_s_median"] = B / (r["median_ms"] / 1e3)
        out["eval_forward"][str(B)] = r
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
