"""Lift the reference's OWN hot-path classes out of /root/reference by ``ast`` (TEST / BASELINE INFRASTRUCTURE).

The reference scripts cannot be imported (private paths and missing packages at import time), so class / function
definitions are taken from the parsed source and ``exec``-ed unchanged in a namespace that supplies torch / numpy /
scipy.  Nothing is copied into this repository.  Used by ``oracle/make_golden.py`` (golden vectors) and by the CPU legs
of ``bench.py`` when the reference tree is present (``cpu_baseline.kind == "reference"``); on a box without
/root/reference those legs fall back to the restatement in ``oracle/unet_oracle.py`` (``kind == "port"``).
"""
from __future__ import annotations

import ast
import os

REF = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "create_testset.py"))


def lift(relpath: str, names, extra=None):
    import numpy as np
    import torch
    import torch.nn as nn
    src = open(os.path.join(REF, relpath)).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "np": np, "F": torch.nn.functional}
    ns.update(extra or {})
    from scipy.ndimage import distance_transform_edt
    ns["distance_transform_edt"] = distance_transform_edt
    got = []
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), relpath, "exec")
            exec(code, ns)
            got.append(node.name)
    missing = set(names) - set(got)
    assert not missing, f"{relpath}: missing {missing}"
    return ns


def reference_model_and_loss(loss_name: str):
    """(UNet instance of src/create_testset.py:53-83, callable net -> logits, criterion) built from the reference's own
    source.  The class's forward ends in a sigmoid (:83) that no training script uses; logits are read at final_conv."""
    ct = lift("src/create_testset.py", ["DoubleConv", "UNet"])
    net = ct["UNet"](in_channels=3, out_channels=1)

    def logits_of(x):
        acts = {}
        hnd = net.final_conv.register_forward_hook(lambda m, i, o: acts.__setitem__("z", o))
        net(x)
        hnd.remove()
        return acts["z"]

    if loss_name == "bce_dice":
        crit = lift("train_bce_dice.py", ["BCEDiceLoss"])["BCEDiceLoss"](bce_weight=0.5, smooth=1.0)
    elif loss_name == "focal_dice":
        crit = lift("src/train_with_focalDice.py", ["FocalLoss", "FocalDiceLoss"])["FocalDiceLoss"](
            alpha=0.5, gamma=2.0, smooth=1.0, w_focal=0.7)
    elif loss_name == "composite":
        crit = lift("src/train_with_boundary_loss.py",
                    ["signed_distance_map_np", "batch_sdf_from_masks", "SymmetricBoundaryLoss", "BCEDiceLoss",
                     "CompositeSegLoss"])["CompositeSegLoss"](bce_weight=0.5, boundary_weight=0.3)
    else:
        raise ValueError(loss_name)
    return net, logits_of, crit
