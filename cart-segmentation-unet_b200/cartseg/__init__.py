"""cartseg — B200-native U-Net training / inference hot path of endressa/cart-segmentation-unet.

Importing the package loads libcartseg.so (hand-written sm_100a CUDA behind a C ABI) and registers
the ``cartseg::`` torch.library ops.  There is no CPU or PyTorch fallback: a missing library raises.
"""
from ._lib import CartsegError, LIB_PATH, lib
from . import ops
from .modules import (ABL, BCEDiceABL, BCEDiceLoss, BCEDiceLossPerSample, CompositeSegLoss, DoubleConv, FocalDiceLoss, FocalLoss,
                      SymmetricBoundaryLoss, UNet, batch_sdf_from_masks)
from .metrics import (dice_iou_at_t, dice_metric, find_best_threshold, hard_dice_metric, hard_iou_metric, iou_metric,
                      precision_recall_f1, pseudo_label_mask, sweep_thresholds, threshold_sums)
from . import parallel
from . import graphs
from .graphs import GraphedInference, GraphedTrainStep
from . import postproc
from . import preproc
from .preproc import letterbox_resize_normalize, resize_masks
from .postproc import (clean_mask, clean_mask_largest_component, ensemble_forward, pseudo_label_qc, should_accept)

lib()   # fail loudly at import time if the extension has not been built

__all__ = [
    "CartsegError", "LIB_PATH", "lib", "ops", "parallel", "graphs", "GraphedInference", "GraphedTrainStep",
    "UNet", "DoubleConv", "BCEDiceLoss", "BCEDiceLossPerSample", "FocalLoss", "FocalDiceLoss",
    "SymmetricBoundaryLoss", "CompositeSegLoss", "batch_sdf_from_masks", "ABL", "BCEDiceABL",
    "dice_metric", "iou_metric", "precision_recall_f1", "dice_iou_at_t", "hard_dice_metric", "hard_iou_metric",
    "sweep_thresholds", "threshold_sums", "find_best_threshold", "pseudo_label_mask",
    "preproc", "letterbox_resize_normalize", "resize_masks",
    "postproc", "ensemble_forward", "pseudo_label_qc", "should_accept", "clean_mask", "clean_mask_largest_component",
]
