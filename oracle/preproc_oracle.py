"""CPU oracle for the input side ("next" row N4 of SURVEY.md §8f): letterbox -> resize -> normalise -> CHW float32.

TEST INFRASTRUCTURE — never imported by the product package.

Restates
  * ``letterbox_image_with_side_padding``   train_bce_dice.py:42-85 (same function in create_pseudo_labels_gpu.py:68-111)
  * ``cv2.resize(img, IMG_SIZE, interpolation=cv2.INTER_LINEAR)``  train_bce_dice.py:147 / ``A.Resize`` :173,
    create_pseudo_labels_gpu.py:113-114 — image path
  * ``cv2.resize(mask, IMG_SIZE, interpolation=cv2.INTER_NEAREST)`` train_bce_dice.py:148 and ``/ 255.0`` :154 — mask path
  * ``A.Normalize(mean, std)`` + ``ToTensorV2``   train_bce_dice.py:174-175, create_pseudo_labels_gpu.py:115-116

Third-party arithmetic.  (i) OpenCV (unpinned in the reference; 4.13.0 in this image): ``resize_linear_u8`` restates its
8-bit bilinear kernel — 11-bit fixed-point coefficients from float32 fractions, horizontal pass in int32, vertical pass
``(((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2``, x coefficients clamped at the borders, y coefficients not (rows
clipped instead), and the exact 2x down-scale routed to the 2x2 area average — and is pinned bit-for-bit against
``cv2.resize`` by tests/golden/preproc.npz.  (ii) albumentations (absent from this image, unpinned): ``A.Normalize`` is
restated from its published formula (albumentations 1.x ``functional.normalize``): float32 image, ``img -= mean*255``,
``img *= 1/(std*255)`` with both constants rounded to float32 — **parity unpinned** for this one step (no way to run
the library here); the tolerance against any float32 evaluation order of that formula is 1 ulp.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def letterbox_geometry(height: int, width: int, side_padding_ratio: float = 0.1):
    """train_bce_dice.py:56-80 -> (square side L, x offset, y offset) of the original image inside the black square."""
    side = round(width * side_padding_ratio)             # Python round (banker's) exactly as the reference
    pw, ph = width + 2 * side, height
    L = max(pw, ph)
    return L, (L - pw) // 2 + side, (L - ph) // 2


def letterbox(image: np.ndarray, side_padding_ratio: float = 0.1) -> np.ndarray:
    H, W = image.shape[:2]
    L, x0, y0 = letterbox_geometry(H, W, side_padding_ratio)
    out = np.zeros((L, L) + image.shape[2:], dtype=np.uint8)
    out[y0:y0 + H, x0:x0 + W] = image
    return out


def _linear_coeffs(src: int, dst: int, clamp: bool):
    scale = 1.0 / (np.float64(dst) / np.float64(src))
    d = np.arange(dst)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    c1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int64)
    c0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int64)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), c0, c1


def resize_linear_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 images, bit-exact."""
    H, W = img.shape[:2]
    i = img.astype(np.int64)
    if H == 2 * dst_h and W == 2 * dst_w:
        return ((i[0::2, 0::2] + i[0::2, 1::2] + i[1::2, 0::2] + i[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    squeeze = i.ndim == 2
    if squeeze:
        i = i[:, :, None]
    x0, x1, a0, a1 = _linear_coeffs(W, dst_w, clamp=True)
    y0, y1, b0, b1 = _linear_coeffs(H, dst_h, clamp=False)
    rows = i[:, x0] * a0[None, :, None] + i[:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (rows[y0] >> 4)) >> 16) + ((b1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def resize_nearest_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_NEAREST): src index = min(floor(d * src/dst), src-1)."""
    H, W = img.shape[:2]
    fx = 1.0 / (np.float64(dst_w) / np.float64(W))
    fy = 1.0 / (np.float64(dst_h) / np.float64(H))
    xs = np.minimum(np.floor(np.arange(dst_w) * fx).astype(np.int64), W - 1)
    ys = np.minimum(np.floor(np.arange(dst_h) * fy).astype(np.int64), H - 1)
    return img[ys][:, xs]


def normalize_chw(img_u8: np.ndarray, mean, std) -> np.ndarray:
    """A.Normalize(mean, std, max_pixel_value=255) + ToTensorV2 -> float32 [3,H,W]."""
    m = (np.asarray(mean, dtype=np.float64) * 255.0).astype(np.float32)
    inv = (1.0 / (np.asarray(std, dtype=np.float64) * 255.0)).astype(np.float32)
    x = img_u8.astype(np.float32)
    x = (x - m) * inv
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def preprocess_image(image: np.ndarray, size: int, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225),
                     side_padding_ratio: float = 0.1, bgr: bool = False) -> np.ndarray:
    """The whole image path for one HWC uint8 image -> float32 [3,size,size]."""
    if bgr:
        image = image[:, :, ::-1]
    return normalize_chw(resize_linear_u8(letterbox(image, side_padding_ratio), size, size), mean, std)


def preprocess_mask(mask: np.ndarray, size: int) -> np.ndarray:
    """train_bce_dice.py:148,154: nearest resize, / 255 -> float32 [1,size,size]."""
    return (resize_nearest_u8(mask, size, size).astype(np.float32) / np.float32(255.0))[None]
