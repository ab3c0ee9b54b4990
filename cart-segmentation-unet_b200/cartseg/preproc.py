"""Input side on the device: letterbox -> bilinear resize -> normalise -> [B,3,S,S] float32, and
the nearest-neighbour mask resize, for a batch of variable-size uint8 images in one launch each.

  letterbox_image_with_side_padding      train_bce_dice.py:42-85
  cv2.resize INTER_LINEAR / INTER_NEAREST :147-148        A.Resize / A.Normalize / ToTensorV2  :171-176
  the pseudo-label transform              src/data_preprocessing/create_pseudo_labels_gpu.py:113-117

At >= 6 k images/s per GPU the reference's 2-4 CPU DataLoader workers cannot feed the model; decoded uint8 images go to
the device as they are (3 B/px instead of 12 B/px of float32) and everything else happens there.  The resize is
OpenCV's 8-bit bilinear kernel bit for bit.  Augmentations (flips, rotations, colour jitter: train_bce_dice.py:161-170)
stay host-side and are out of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import CartsegError, ImageDesc, check, ptr

IMAGENET_MEAN = (0.485, 0.456, 0.406)        # create_pseudo_labels_gpu.py:54-55
IMAGENET_STD = (0.229, 0.224, 0.225)


def letterbox_geometry(height: int, width: int, side_padding_ratio: float = 0.1) -> Tuple[int, int, int]:
    """(canvas side, x offset, y offset) of train_bce_dice.py:56-80."""
    side, x0, y0 = C.c_int(), C.c_int(), C.c_int()
    check(_lib.lib().cs_letterbox_geometry(int(height), int(width), float(side_padding_ratio), C.byref(side),
                                           C.byref(x0), C.byref(y0)), "cs_letterbox_geometry")
    return side.value, x0.value, y0.value


def _desc_table(images: Sequence[Tensor], channels: int, geometry) -> Tensor:
    if not images:
        raise CartsegError("empty batch")
    dev = images[0].device
    table = (ImageDesc * len(images))()
    for i, im in enumerate(images):
        if not im.is_cuda or im.device != dev:
            raise CartsegError("preprocessing takes CUDA tensors on one device only (no CPU fallback)")
        if im.dtype != torch.uint8 or not im.is_contiguous():
            raise CartsegError("images must be contiguous uint8")
        if channels == 3 and (im.dim() != 3 or im.shape[2] != 3):
            raise CartsegError("images must be [H,W,3] uint8")
        if channels == 1 and im.dim() != 2:
            raise CartsegError("masks must be [H,W] uint8")
        H, W = int(im.shape[0]), int(im.shape[1])
        ch, cw, x0, y0 = geometry(H, W)
        table[i] = ImageDesc(im.data_ptr(), H, W, W * channels, ch, cw, x0, y0, 0)
    host = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).pin_memory()
    return host.to(dev, non_blocking=True)


@torch.no_grad()
def letterbox_resize_normalize(images: Sequence[Tensor], size: int, mean: Sequence[float] = IMAGENET_MEAN,
                               std: Sequence[float] = IMAGENET_STD, side_padding_ratio: float = 0.1,
                               bgr: bool = False, letterbox: bool = True) -> Tensor:
    """``images``: uint8 CUDA tensors [H_i, W_i, 3] (decoded frames, any sizes).  Returns float32 [B,3,size,size]:
    letterbox (black side padding of ``round(W * ratio)`` columns, then square), OpenCV-exact bilinear resize,
    ``(x - mean*255) / (std*255)``, CHW.  ``bgr=True`` takes cv2.imread's channel order.  ``mean=(0,0,0), std=(1,1,1)``
    is the training transform of train_bce_dice.py:174."""
    if letterbox:
        def geometry(H, W):
            side, x0, y0 = letterbox_geometry(H, W, side_padding_ratio)
            return side, side, x0, y0
    else:
        def geometry(H, W):
            return H, W, 0, 0
    table = _desc_table(images, 3, geometry)
    dev = images[0].device
    out = torch.empty((len(images), 3, size, size), dtype=torch.float32, device=dev)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(dev):
        check(_lib.lib().cs_preproc_images(ptr(table), len(images), int(size), m, s, int(bgr), ptr(out),
                                           _lib.current_stream()), "cs_preproc_images")
    return out


@torch.no_grad()
def resize_masks(masks: Sequence[Tensor], size: int) -> Tensor:
    """``masks``: uint8 CUDA tensors [H_i, W_i] in {0,255}.  Returns float32 [B,1,size,size] = nearest-neighbour resize
    / 255 (train_bce_dice.py:148,154)."""
    table = _desc_table(masks, 1, lambda H, W: (H, W, 0, 0))
    dev = masks[0].device
    out = torch.empty((len(masks), 1, size, size), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cs_preproc_masks(ptr(table), len(masks), int(size), ptr(out), _lib.current_stream()),
              "cs_preproc_masks")
    return out
