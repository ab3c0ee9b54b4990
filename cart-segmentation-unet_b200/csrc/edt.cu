// Exact Euclidean distance transform -> signed distance map, bit-exact with the reference's
// scipy.ndimage.distance_transform_edt path (src/train_with_boundary_loss.py:191-217).
//
//   fg(pixel)  = ge ? (src >= thr) : (src > thr)
//   sdf(pixel) = +dist(nearest fg pixel)  for background pixels
//                -dist(nearest bg pixel)  for foreground pixels          (then / norm, fp32 divide)
//   images that are all-fg or all-bg give 0 everywhere.
//
// Separable and exact in integers: pass 1 computes, per column, the vertical distance to the nearest
// fg and the nearest bg pixel (uint16 each); pass 2 takes, per pixel, the minimum over the row of
// dx^2 + g(x')^2 against the OPPOSITE class (one transform instead of the reference's two, because at
// every pixel one of its two EDTs is zero).  Squared distances are < 2^24, so int -> float is exact
// and sqrtf / division are IEEE-rounded (this file must not be built with -use_fast_math); that
// equals the reference's float64 sqrt rounded to float32 for every reachable value
// (tests/test_oracle_golden.py::test_sqrt_f32_equals_f64_path_for_all_reachable_values).
#include "edt.cuh"

namespace cs {

size_t sdf_scratch_bytes(int B, int H, int W) { return (size_t)B * H * W * 4 + (size_t)B * 8; }

struct ThresholdPred {
  const float* src; float thr; int ge, H, W;
  __device__ bool operator()(int img, int y, int x) const {
    const float v = __ldg(src + ((size_t)img * H + y) * W + x);
    return ge ? (v >= thr) : (v > thr);
  }
};

// sdf = +dist(nearest fg) on background pixels, -dist(nearest bg) on foreground pixels, / norm; zeros when the image
// is all-fg or all-bg (src/train_with_boundary_loss.py:193-195).
struct SdfEpilogue {
  float* sdf; float norm; int H, W;
  __device__ void operator()(int img, int y, int x, bool fg, int best_sq, bool has_fg, bool has_bg) const {
    float out = 0.f;
    if (has_fg && has_bg) {
      const float dist = __fdiv_rn(__fsqrt_rn((float)best_sq), norm);
      out = fg ? -dist : dist;
    }
    sdf[((size_t)img * H + y) * W + x] = out;
  }
};

// Fallback for images taller than the segmented pass supports: one thread walks one column.
__global__ void sdf_columns_serial_kernel(ThresholdPred pred, int B, int H, int W, ushort2* __restrict__ g,
                                          int* __restrict__ flags) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * W) return;
  const int b = idx / W, x = idx - b * W;
  ushort2* gc = g + (size_t)b * H * W + x;
  int any_fg = 0, any_bg = 0, df = kEdtInf, db = kEdtInf;
  for (int y = 0; y < H; ++y) {
    const bool fg = pred(b, y, x);
    df = fg ? 0 : min(df + 1, kEdtInf);
    db = fg ? min(db + 1, kEdtInf) : 0;
    any_fg |= fg;
    any_bg |= !fg;
    gc[(size_t)y * W] = make_ushort2((unsigned short)df, (unsigned short)db);
  }
  df = kEdtInf; db = kEdtInf;
  for (int y = H - 1; y >= 0; --y) {
    const ushort2 d = gc[(size_t)y * W];
    const bool fg = d.x == 0;
    df = fg ? 0 : min(df + 1, kEdtInf);
    db = fg ? min(db + 1, kEdtInf) : 0;
    gc[(size_t)y * W] = make_ushort2((unsigned short)min((int)d.x, df), (unsigned short)min((int)d.y, db));
  }
  if (any_fg) atomicOr(&flags[2 * b], 1);
  if (any_bg) atomicOr(&flags[2 * b + 1], 1);
}

cudaError_t launch_sdf(const float* src, float thr, int ge, int B, int H, int W, float norm, float* sdf,
                       void* scratch, cudaStream_t s) {
  if (H >= kEdtInf || W >= kEdtInf) return cudaErrorInvalidValue;
  ushort2* g = reinterpret_cast<ushort2*>(scratch);
  int* flags = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(scratch) + (size_t)B * H * W * 4);
  cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)B * 8, s);
  if (e != cudaSuccess) return e;
  const ThresholdPred pred{src, thr, ge, H, W};
  int nseg = 0, rps = 0;
  if (edt_column_geometry(H, &nseg, &rps)) {
    edt_columns_kernel<<<B * ((W + 31) / 32), dim3(32, nseg), 0, s>>>(pred, B, H, W, rps, g, flags);
  } else {
    sdf_columns_serial_kernel<<<(B * W + 63) / 64, 64, 0, s>>>(pred, B, H, W, g, flags);
  }
  if ((e = launched()) != cudaSuccess) return e;
  const SdfEpilogue epi{sdf, norm, H, W};
  const int threads = W >= 256 ? 256 : ((W + 31) / 32) * 32;
  if (W <= kEdtMaxPaddedW) edt_rows_kernel<true><<<B * H, threads, edt_rows_smem(W, true), s>>>(epi, g, flags, H, W);
  else edt_rows_kernel<false><<<B * H, threads, edt_rows_smem(W, false), s>>>(epi, g, flags, H, W);
  return launched();
}

}  // namespace cs
