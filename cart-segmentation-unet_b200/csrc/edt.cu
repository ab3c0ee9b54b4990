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
#include "common.cuh"
#include "kernels.cuh"

namespace cs {

static constexpr int kInf = 30000;   // > any in-image distance (H, W <= 16384); kInf^2 fits int32

size_t sdf_scratch_bytes(int B, int H, int W) { return (size_t)B * H * W * 4 + (size_t)B * 8; }

// One thread per column.  g[(b,y,x)] = {dist to nearest fg in column, dist to nearest bg in column}.
__global__ void sdf_columns_kernel(const float* __restrict__ src, float thr, int ge, int B, int H, int W,
                                   ushort2* __restrict__ g, int* __restrict__ flags) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < B * W;
  const int b = active ? idx / W : 0, x = active ? idx - b * W : 0;
  const float* col = src + (size_t)b * H * W + x;
  ushort2* gc = g + (size_t)b * H * W + x;
  int any_fg = 0, any_bg = 0;
  if (active) {
    int df = kInf, db = kInf;
#pragma unroll 4
    for (int y = 0; y < H; ++y) {
      const float v = __ldg(col + (size_t)y * W);
      const bool fg = ge ? (v >= thr) : (v > thr);
      df = fg ? 0 : min(df + 1, kInf);
      db = fg ? min(db + 1, kInf) : 0;
      any_fg |= fg;
      any_bg |= !fg;
      gc[(size_t)y * W] = make_ushort2((unsigned short)df, (unsigned short)db);
    }
    df = kInf; db = kInf;
    for (int y = H - 1; y >= 0; --y) {
      const ushort2 d = gc[(size_t)y * W];
      const bool fg = d.x == 0;
      df = fg ? 0 : min(df + 1, kInf);
      db = fg ? min(db + 1, kInf) : 0;
      gc[(size_t)y * W] = make_ushort2((unsigned short)min((int)d.x, df), (unsigned short)min((int)d.y, db));
    }
  }
  // per-image "has fg" / "has bg" flags (a warp may straddle two images: reduce per lane's own image)
  if (active) {
    if (any_fg) atomicOr(&flags[2 * b], 1);
    if (any_bg) atomicOr(&flags[2 * b + 1], 1);
  }
}

// One block per image row.  Each thread owns pixels x = tid, tid + blockDim, ... and scans outwards from
// x; the scan stops as soon as dx^2 alone can no longer beat the best candidate (exact pruning).
__global__ void sdf_rows_kernel(const ushort2* __restrict__ g, const int* __restrict__ flags, int H, int W, float norm,
                                float* __restrict__ sdf) {
  extern __shared__ int srow[];                 // [2][W]: squared column distances to fg / to bg
  const int row = blockIdx.x;                   // b*H + y
  const int b = row / H;
  const ushort2* gr = g + (size_t)row * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const ushort2 d = gr[x];
    srow[x] = (int)d.x * (int)d.x;
    srow[W + x] = (int)d.y * (int)d.y;
  }
  __syncthreads();
  const bool degenerate = !(flags[2 * b] && flags[2 * b + 1]);
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    float out = 0.f;
    if (!degenerate) {
      const bool fg = srow[x] == 0;             // distance to nearest fg is 0 <=> the pixel is fg
      const int* opp = fg ? (srow + W) : srow;  // fg pixels look for bg, bg pixels look for fg
      int best = opp[x];
      for (int d = 1; d < W && d * d < best; ++d) {
        const int d2 = d * d;
        if (x - d >= 0) best = min(best, d2 + opp[x - d]);
        if (x + d < W) best = min(best, d2 + opp[x + d]);
      }
      const float dist = __fdiv_rn(__fsqrt_rn((float)best), norm);
      out = fg ? -dist : dist;
    }
    sdf[(size_t)row * W + x] = out;
  }
}

cudaError_t launch_sdf(const float* src, float thr, int ge, int B, int H, int W, float norm, float* sdf,
                       void* scratch, cudaStream_t s) {
  if (H >= kInf || W >= kInf) return cudaErrorInvalidValue;
  ushort2* g = reinterpret_cast<ushort2*>(scratch);
  int* flags = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(scratch) + (size_t)B * H * W * 4);
  cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)B * 8, s);
  if (e != cudaSuccess) return e;
  const int cols = B * W;
  sdf_columns_kernel<<<(cols + 63) / 64, 64, 0, s>>>(src, thr, ge, B, H, W, g, flags);
  e = launched();
  if (e != cudaSuccess) return e;
  const int threads = W >= 256 ? 256 : ((W + 31) / 32) * 32;
  sdf_rows_kernel<<<B * H, threads, 2 * W * sizeof(int), s>>>(g, flags, H, W, norm, sdf);
  return launched();
}

}  // namespace cs
