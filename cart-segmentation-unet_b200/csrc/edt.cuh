// Exact separable Euclidean distance transform, shared by the signed-distance-map kernel (edt.cu) and the Active
// Boundary Loss (abl.cu).  Two passes, integers only:
//   edt_columns_kernel  per column, the vertical distance from every pixel to the nearest "set" and nearest "unset"
//                       pixel (uint16 each).  A column is cut into SEG segments of <= 32 rows; one thread owns one
//                       segment of one column, keeps its predicate bits in a register, exchanges the first / last set
//                       (and unset) row of every segment through shared memory, and then derives each row's distances
//                       with bit scans — no serial walk over the column, one coalesced read and one coalesced write.
//   edt_rows_kernel     per pixel, min over x' of (x - x')^2 + g(x')^2 against the OPPOSITE class with exact pruning
//                       (stop once d^2 alone cannot beat the best candidate); the row is staged in shared memory with
//                       "infinite" padding on both sides so that the scan needs no bounds checks.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace cs {

static constexpr int kEdtInf = 30000;        // > any in-image distance (H, W <= 16384); kEdtInf^2 fits int32
static constexpr int kEdtMaxSegRows = 32;    // rows per thread in the segmented column pass (one bit each)
static constexpr int kEdtMaxSeg = 32;        // segments per column  -> H <= 1024
static constexpr int kEdtMaxPaddedW = 1536;  // rows kernel with padded shared memory: 27 * W bytes <= 48 KB

// Pred: __device__ bool operator()(int img, int y, int x) const   — whether pixel (y, x) of image img is "set".
// flags[2*img] |= 1 if the image has a set pixel, flags[2*img+1] |= 1 if it has an unset pixel.
template <class Pred>
__global__ void edt_columns_kernel(Pred pred, int nimg, int H, int W, int rows_per_seg, ushort2* __restrict__ g,
                                   int* __restrict__ flags) {
  __shared__ short s_first[2][kEdtMaxSeg][32], s_last[2][kEdtMaxSeg][32];
  const int groups = (W + 31) >> 5;
  const int img = blockIdx.x / groups, x = (blockIdx.x - img * groups) * 32 + threadIdx.x;
  const int seg = threadIdx.y, nseg = blockDim.y;
  const int y0 = seg * rows_per_seg;
  const bool active = x < W && img < nimg;
  uint32_t set = 0, valid = 0;
  if (active) {
    for (int r = 0; r < rows_per_seg; ++r) {
      const int y = y0 + r;
      if (y < H) {
        valid |= 1u << r;
        if (pred(img, y, x)) set |= 1u << r;
      }
    }
  }
  const uint32_t bits[2] = {set, ~set & valid};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    s_first[c][seg][threadIdx.x] = bits[c] ? (short)(y0 + __ffs(bits[c]) - 1) : (short)kEdtInf;
    s_last[c][seg][threadIdx.x] = bits[c] ? (short)(y0 + 31 - __clz(bits[c])) : (short)-kEdtInf;
  }
  __syncthreads();
  {                                             // one atomic per warp and class (a warp never straddles two images)
    const unsigned any_set = __ballot_sync(0xffffffffu, active && set != 0);
    const unsigned any_unset = __ballot_sync(0xffffffffu, active && bits[1] != 0);
    if (threadIdx.x == 0 && img < nimg) {
      if (any_set) atomicOr(&flags[2 * img], 1);
      if (any_unset) atomicOr(&flags[2 * img + 1], 1);
    }
  }
  if (!active) return;
  int above[2] = {-kEdtInf, -kEdtInf}, below[2] = {2 * kEdtInf, 2 * kEdtInf};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    for (int s = seg - 1; s >= 0; --s) {
      const int v = s_last[c][s][threadIdx.x];
      if (v >= 0) { above[c] = v; break; }
    }
    for (int s = seg + 1; s < nseg; ++s) {
      const int v = s_first[c][s][threadIdx.x];
      if (v < kEdtInf) { below[c] = v; break; }
    }
  }
  ushort2* gc = g + (size_t)img * H * W + x;
  for (int r = 0; r < rows_per_seg; ++r) {
    const int y = y0 + r;
    if (y >= H) break;
    int d[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t up = bits[c] & (0xffffffffu >> (31 - r));       // rows <= r of this segment
      const uint32_t dn = bits[c] >> r;                              // rows >= r
      const int ya = up ? y0 + 31 - __clz(up) : above[c];
      const int yb = dn ? y + __ffs(dn) - 1 : below[c];
      d[c] = min(min(y - ya, yb - y), kEdtInf);
    }
    gc[(size_t)y * W] = make_ushort2((unsigned short)d[0], (unsigned short)d[1]);
  }
}

// Epi: __device__ void operator()(int img, int y, int x, bool set, int best_sq, bool has_set, bool has_unset) const
//   best_sq = squared distance to the nearest pixel of the opposite class (garbage when that class is absent).
//
// PADDED (W <= kEdtMaxPaddedW): the two rows of squared column distances live in shared memory with Wp = round_up(W, 8)
// "infinite" entries on either side, plus the minimum of every aligned 8-column chunk.  A pixel scans chunks outwards;
// a chunk whose lower bound (nearest column offset)^2 + (chunk minimum) cannot beat the best candidate is skipped with
// one load, the others are read as two 16-byte vectors — exact, because every skipped candidate is >= its bound.
static constexpr int kEdtInfSq = kEdtInf * kEdtInf;

CS_DEVINL int edt_scan_chunk(const int* __restrict__ chunk, int d_first, int step, int best) {
  // candidates (d_first + step*j)^2 + chunk[j], j = 0..7  (step = -1: offsets shrink along the chunk, +1: they grow)
  const int4 a = *reinterpret_cast<const int4*>(chunk), b = *reinterpret_cast<const int4*>(chunk + 4);
  const int v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int d = d_first + step * j;
    best = min(best, d * d + v[j]);
  }
  return best;
}

template <bool PADDED, class Epi>
__global__ void edt_rows_kernel(Epi epi, const ushort2* __restrict__ g, const int* __restrict__ flags, int H, int W) {
  extern __shared__ __align__(16) int srow[];
  const int row = blockIdx.x;                   // img*H + y
  const int img = row / H, y = row - img * H;
  const ushort2* gr = g + (size_t)row * W;
  const int Wp = PADDED ? (W + 7) & ~7 : W;
  const int pitch = PADDED ? 3 * Wp : W, off = PADDED ? Wp : 0;
  int* to_set = srow + off;                     // squared column distance to the nearest set pixel
  int* to_unset = srow + pitch + off;
  int* cmin = srow + 2 * pitch;                 // [2][3*Wp/8] chunk minima (PADDED only)
  if (PADDED) {
    for (int i = threadIdx.x; i < 2 * pitch; i += blockDim.x) srow[i] = kEdtInfSq;
    __syncthreads();
  }
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const ushort2 d = gr[x];
    to_set[x] = (int)d.x * (int)d.x;
    to_unset[x] = (int)d.y * (int)d.y;
  }
  __syncthreads();
  if (PADDED) {
    const int nchunk = pitch >> 3;
    for (int i = threadIdx.x; i < 2 * nchunk; i += blockDim.x) {
      const int* c = srow + (i >= nchunk ? pitch + ((i - nchunk) << 3) : (i << 3));
      int m = c[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) m = min(m, c[j]);
      cmin[i] = m;
    }
    __syncthreads();
  }
  const bool has_set = flags[2 * img] != 0, has_unset = flags[2 * img + 1] != 0;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const bool set = to_set[x] == 0;            // distance to the nearest set pixel is 0 <=> the pixel is set
    int best;
    if (!(set ? has_unset : has_set)) {
      best = kEdtInfSq;                         // the opposite class does not occur in this image: nothing to scan for
    } else if (PADDED) {
      const int* row_opp = srow + (set ? pitch : 0);          // padded coordinates: column x lives at Wp + x
      const int* cm = cmin + (set ? pitch >> 3 : 0);
      const int X = Wp + x, C = X >> 3;
      best = edt_scan_chunk(row_opp + (C << 3), (C << 3) - X, 1, kEdtInfSq);      // own chunk (offsets -j0 .. 7-j0)
      const int maxk = Wp >> 3;
      for (int k = 1; k <= maxk; ++k) {
        const int cl = C - k, cr = C + k;
        const int dl = X - ((cl << 3) + 7), dr = (cr << 3) - X;                   // nearest column of either chunk
        if (min(dl, dr) * min(dl, dr) >= best) break;
        if (dl * dl + cm[cl] < best) best = edt_scan_chunk(row_opp + (cl << 3), dl + 7, -1, best);
        if (dr * dr + cm[cr] < best) best = edt_scan_chunk(row_opp + (cr << 3), dr, 1, best);
      }
    } else {
      const int* opp = (set ? to_unset : to_set) + x;
      best = opp[0];
      for (int d = 1; d < W && d * d < best; ++d) {
        const int d2 = d * d;
        if (x - d >= 0) best = min(best, d2 + opp[-d]);
        if (x + d < W) best = min(best, d2 + opp[d]);
      }
    }
    epi(img, y, x, set, best, has_set, has_unset);
  }
}

// dynamic shared memory of edt_rows_kernel
inline size_t edt_rows_smem(int W, bool padded) {
  if (!padded) return (size_t)2 * W * sizeof(int);
  const size_t Wp = ((size_t)W + 7) & ~(size_t)7;
  return (2 * 3 * Wp + 2 * 3 * Wp / 8) * sizeof(int);
}

// Launch geometry of the segmented column pass (host side).  Returns false when H is too tall for it.
inline bool edt_column_geometry(int H, int* nseg, int* rows_per_seg) {
  int s = 8;
  while (s < kEdtMaxSeg && (H + s - 1) / s > kEdtMaxSegRows) s *= 2;
  const int r = (H + s - 1) / s;
  if (r > kEdtMaxSegRows) return false;
  *nseg = s;
  *rows_per_seg = r;
  return true;
}

}  // namespace cs
