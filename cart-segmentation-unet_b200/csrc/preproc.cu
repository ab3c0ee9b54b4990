// Input side on the device: letterbox -> bilinear resize -> normalise -> NCHW fp32, and the
// nearest-neighbour mask resize, for a batch of variable-size uint8 images in ONE launch each.  Replaces, per image on
// the CPU, letterbox_image_with_side_padding (train_bce_dice.py:42-85), cv2.resize INTER_LINEAR / INTER_NEAREST
// (:147-148), A.Resize + A.Normalize + ToTensorV2 (:171-176; create_pseudo_labels_gpu.py:113-117).
//
// The bilinear kernel is OpenCV's 8-bit one, bit for bit: float32 fractions -> 11-bit fixed-point coefficients
// (round to nearest), horizontal pass in int32, vertical pass (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2;
// x coefficients are clamped at the borders, y coefficients are not (rows are clipped instead); an exact 2x
// down-scale is the 2x2 average (a+b+c+d+2)>>2 (OpenCV routes INTER_LINEAR there).  The letterboxed canvas is never
// materialised: taps outside the image read the padding colour (black).
#include "common.cuh"
#include "kernels.cuh"

namespace cs {

struct Axis { int s0, s1, c0, c1; };

// OpenCV's coefficient for destination index d along an axis of `src` source and `dst` destination samples
CS_DEVINL Axis linear_axis(int d, int src, int dst, bool clamp) {
  const double scale = 1.0 / ((double)dst / (double)src);
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (clamp) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  Axis a;
  a.c1 = __float2int_rn(f * 2048.0f);
  a.c0 = __float2int_rn((1.0f - f) * 2048.0f);
  a.s0 = min(max(s, 0), src - 1);
  a.s1 = min(max(s + 1, 0), src - 1);
  return a;
}

// value of the letterboxed canvas at (cy, cx), channel ch
CS_DEVINL int canvas_px(const ImageDesc& im, int cy, int cx, int ch) {
  const int y = cy - im.y0, x = cx - im.x0;
  if (y < 0 || y >= im.height || x < 0 || x >= im.width) return 0;
  return (int)__ldg(im.data + (size_t)y * im.pitch + (size_t)x * 3 + ch);
}

__global__ void __launch_bounds__(256) preproc_images_kernel(const ImageDesc* __restrict__ descs, int S, Norm3 nrm, int bgr,
                                                             float* __restrict__ out) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, b = blockIdx.z;
  if (dx >= S) return;
  const ImageDesc im = descs[b];
  int v[3];
  if (im.canvas_h == 2 * S && im.canvas_w == 2 * S) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      v[c] = (canvas_px(im, 2 * dy, 2 * dx, c) + canvas_px(im, 2 * dy, 2 * dx + 1, c) + canvas_px(im, 2 * dy + 1, 2 * dx, c) +
              canvas_px(im, 2 * dy + 1, 2 * dx + 1, c) + 2) >> 2;
  } else {
    const Axis ax = linear_axis(dx, im.canvas_w, S, true), ay = linear_axis(dy, im.canvas_h, S, false);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int r0 = canvas_px(im, ay.s0, ax.s0, c) * ax.c0 + canvas_px(im, ay.s0, ax.s1, c) * ax.c1;
      const int r1 = canvas_px(im, ay.s1, ax.s0, c) * ax.c0 + canvas_px(im, ay.s1, ax.s1, c) * ax.c1;
      const int o = (((ay.c0 * (r0 >> 4)) >> 16) + ((ay.c1 * (r1 >> 4)) >> 16) + 2) >> 2;
      v[c] = min(max(o, 0), 255);
    }
  }
  const size_t plane = (size_t)S * S;
  float* o = out + (size_t)b * 3 * plane + (size_t)dy * S + dx;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int src_c = bgr ? 2 - c : c;                       // cv2.imread is BGR; the reference flips to RGB (:143)
    o[c * plane] = __fmul_rn(__fsub_rn((float)v[src_c], nrm.mean255[c]), nrm.inv_std255[c]);
  }
}

__global__ void __launch_bounds__(256) preproc_masks_kernel(const ImageDesc* __restrict__ descs, int S,
                                                            float* __restrict__ out) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, b = blockIdx.z;
  if (dx >= S) return;
  const ImageDesc im = descs[b];
  const double fx = 1.0 / ((double)S / (double)im.width), fy = 1.0 / ((double)S / (double)im.height);
  const int sx = min((int)floor(dx * fx), im.width - 1), sy = min((int)floor(dy * fy), im.height - 1);
  const int v = (int)__ldg(im.data + (size_t)sy * im.pitch + sx);
  out[((size_t)b * S + dy) * S + dx] = __fdiv_rn((float)v, 255.0f);           // mask / 255.0 (train_bce_dice.py:154)
}

cudaError_t launch_preproc_images(const ImageDesc* descs, int B, int S, const Norm3& nrm, int bgr, float* out,
                                  cudaStream_t s) {
  dim3 grid((S + 255) / 256, S, B);
  preproc_images_kernel<<<grid, 256, 0, s>>>(descs, S, nrm, bgr, out);
  return launched();
}
cudaError_t launch_preproc_masks(const ImageDesc* descs, int B, int S, float* out, cudaStream_t s) {
  dim3 grid((S + 255) / 256, S, B);
  preproc_masks_kernel<<<grid, 256, 0, s>>>(descs, S, out);
  return launched();
}

}  // namespace cs
