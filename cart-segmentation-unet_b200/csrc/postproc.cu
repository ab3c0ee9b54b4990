// Pseudo-label post-processing on the device — the step right after the inference forward:
//   ensemble_forward      src/data_preprocessing/create_pseudo_labels_gpu.py:201-215  (sum_m w_m * sigmoid(logits_m))
//   threshold + QC scores :294-300  (mask, foreground area, median confidence, mean entropy) — the reference ships
//                         4 B/px of probabilities to the host for these; here 1 B/px of mask and 3 numbers per image
//   clean_mask            src/data_preprocessing/clean_masks.py:12-32   (flood-fill hole filling + largest component)
//   clean_mask_largest_component   src/data_preprocessing/remove_blops.py:14-33
// Connected components: label-equivalence union-find over the whole batch (one int per pixel, atomicMin unions);
// the median is an exact 3-pass radix select on the float bit patterns.
#include "common.cuh"
#include "kernels.cuh"

namespace cs {

static int pp_grid(long long work, int per_block) {
  long long g = (work + per_block - 1) / per_block;
  if (g > 148 * 8) g = 148 * 8;
  return g < 1 ? 1 : (int)g;
}

// ------------------------------------------------------------------------------------------------ ensemble
__global__ void ensemble_accumulate_kernel(const float* __restrict__ logits, float w, long long n, int first,
                                           float* __restrict__ probs) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = __fmul_rn(1.0f / (1.0f + expf(-__ldg(logits + i))), w);     // p.mul_(w)
    probs[i] = first ? p : __fadd_rn(probs[i], p);                              // out_sum.add_(p, alpha=w)
  }
}
cudaError_t launch_ensemble_accumulate(const float* logits, float w, long long n, int first, float* probs, cudaStream_t s) {
  ensemble_accumulate_kernel<<<pp_grid(n, 256 * 4), 256, 0, s>>>(logits, w, n, first, probs);
  return launched();
}

// ------------------------------------------------------------------------------------------------ QC scores
// One block per image, three passes over its probabilities (L2-resident): conf = |p - 0.5| * 2 lies in [0, 1], so its
// float bit pattern is < 2^30 and orders like an unsigned int; the two middle order statistics (numpy's median rule
// for even counts) are found together by an 11 + 11 + 8 bit radix select.  The first pass also produces the mask, the
// foreground count and the entropy sum.  Histogram updates are aggregated per warp (__match_any_sync): confident
// predictions put most pixels into a handful of bins.
static constexpr int kQcThreads = 1024;
static constexpr int kQcBins = 2048;

CS_DEVINL unsigned int conf_bits(float p) { return __float_as_uint(__fmul_rn(fabsf(__fsub_rn(p, 0.5f)), 2.0f)); }

// Run-length aggregation in registers: a thread walks consecutive pixels, whose confidences mostly share their leading
// bits, and touches the shared histogram only when the bin changes.
struct BinRun {
  unsigned int bin, count;
  CS_DEVINL void add(unsigned int* hist, unsigned int b) {
    if (count && b == bin) { ++count; return; }
    if (count) atomicAdd(&hist[bin], count);
    bin = b; count = 1;
  }
  CS_DEVINL void flush(unsigned int* hist) {
    if (count) atomicAdd(&hist[bin], count);
    count = 0;
  }
};

__global__ void __launch_bounds__(kQcThreads) pseudo_qc_kernel(const float* __restrict__ probs, long long n, float thr,
                                                              int mask_value, uint8_t* __restrict__ mask,
                                                              double* __restrict__ stats) {
  __shared__ unsigned int hist[2][kQcBins];                  // [0]: selection of the lower middle rank, [1]: upper
  __shared__ double s_ent[kQcThreads / 32];
  __shared__ unsigned int s_cnt[kQcThreads / 32];
  __shared__ unsigned int s_prefix[2], s_maskbits;
  __shared__ unsigned long long s_k[2];
  const float* p = probs + (size_t)blockIdx.x * n;
  uint8_t* m = mask ? mask + (size_t)blockIdx.x * n : nullptr;
  if (threadIdx.x == 0) {
    s_prefix[0] = s_prefix[1] = 0; s_maskbits = 0;
    s_k[0] = (unsigned long long)((n - 1) / 2);              // np.median: ranks (n-1)/2 and n/2 (0-based)
    s_k[1] = (unsigned long long)(n / 2);
  }
  constexpr int kRun = 8;                                    // consecutive pixels per thread and step
  unsigned int cnt = 0;
  double ent = 0.0;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 19 : (pass == 1 ? 8 : 0);
    const unsigned int bins = pass == 2 ? 256u : 2048u;
    for (int i = threadIdx.x; i < 2 * kQcBins; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    const unsigned int pre0 = s_prefix[0], pre1 = s_prefix[1], mb = s_maskbits;
    const bool split = pre0 != pre1;                         // the two ranks fell into different bins earlier
    BinRun r0{0, 0}, r1{0, 0};
    for (long long i0 = (long long)threadIdx.x * kRun; i0 < n; i0 += (long long)blockDim.x * kRun) {
      const int m_end = (int)min((long long)kRun, n - i0);
      for (int j = 0; j < m_end; ++j) {
        const long long i = i0 + j;
        const float v = __ldg(p + i);
        if (pass == 0) {
          const bool fg = v >= thr;                                      // create_pseudo_labels_gpu.py:294
          cnt += fg ? 1u : 0u;
          if (m) m[i] = fg ? (uint8_t)mask_value : (uint8_t)0;
          const float c = fminf(fmaxf(v, 1e-6f), 0.999999f);             // np.clip(p, eps, 1 - eps) in float32 (:129)
          ent += (double)(-(__fadd_rn(__fmul_rn(c, logf(c)), __fmul_rn(__fsub_rn(1.0f, c), logf(__fsub_rn(1.0f, c))))));
        }
        const unsigned int b = conf_bits(v);
        const unsigned int bin = (b >> shift) & (bins - 1);
        if ((b & mb) == pre0) r0.add(hist[0], bin);
        if (split && (b & mb) == pre1) r1.add(hist[1], bin);
      }
    }
    r0.flush(hist[0]);
    r1.flush(hist[1]);
    __syncthreads();
    if (threadIdx.x < 64) {                                  // warp w locates the bin of rank w: 32 partial sums,
      const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;        // a warp scan, then <= 64 serial steps
      const unsigned int* h = hist[(split && w == 1) ? 1 : 0];
      const unsigned int per = bins >> 5;
      unsigned int part = 0;
      for (unsigned int j = 0; j < per; ++j) part += h[lane * per + j];
      unsigned int incl = part;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned long long kk0 = s_k[w];
      // first lane whose inclusive count exceeds the rank (the last lane if none does: rank beyond the data)
      const unsigned int hit = __ballot_sync(0xffffffffu, (unsigned long long)incl > kk0);
      const int owner = hit ? __ffs(hit) - 1 : 31;
      if (lane == owner) {
        unsigned long long kk = kk0 - (unsigned long long)(incl - part);
        unsigned int bin = lane * per;
        const unsigned int last = bin + per - 1;
        while (bin < last && kk >= h[bin]) { kk -= h[bin]; ++bin; }
        s_k[w] = kk;
        s_prefix[w] |= bin << shift;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_maskbits |= (bins - 1) << shift;
    __syncthreads();
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, o);
  if ((threadIdx.x & 31) == 0) { s_cnt[threadIdx.x >> 5] = cnt; s_ent[threadIdx.x >> 5] = ent; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long c = 0; double e = 0.0;
    for (int w = 0; w < kQcThreads / 32; ++w) { c += s_cnt[w]; e += s_ent[w]; }
    const float a = __uint_as_float(s_prefix[0]), b = __uint_as_float(s_prefix[1]);
    stats[(size_t)blockIdx.x * 4 + 0] = (double)c;                   // foreground pixels (area = count / n)
    stats[(size_t)blockIdx.x * 4 + 1] = (double)__fmul_rn(__fadd_rn(a, b), 0.5f);   // float32 mean of the two middles
    stats[(size_t)blockIdx.x * 4 + 2] = e / (double)n;               // mean entropy
    stats[(size_t)blockIdx.x * 4 + 3] = (double)n;
  }
}
cudaError_t launch_pseudo_qc(const float* probs, int B, long long n, float thr, int mask_value, uint8_t* mask,
                             double* stats, cudaStream_t s) {
  pseudo_qc_kernel<<<B, kQcThreads, 0, s>>>(probs, n, thr, mask_value, mask, stats);
  return launched();
}

// ------------------------------------------------------------------------------------------------ components
CS_DEVINL int uf_find(const int* __restrict__ L, int i) {
  int r = L[i];
  while (r != i) { i = r; r = L[i]; }
  return r;
}
CS_DEVINL void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }     // a > b: hang the larger root below the smaller one
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// cls[i] = 1 foreground / 0 background (written when `in` is given).  L[i] = leftmost pixel of i's horizontal run of
// class `want` inside its 32-pixel lane group (ballot scan), so horizontal connectivity costs no union at all except at
// group boundaries.  The grid-stride loop keeps warps on 32 consecutive flat indices; runs break at row starts.
__global__ void cc_init_kernel(const uint8_t* __restrict__ in, int thr, int want, int W, long long total,
                               uint8_t* __restrict__ cls, int* __restrict__ L) {
  const long long total_ceil = (total + 31) / 32 * 32;
  const unsigned int lane = threadIdx.x & 31;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_ceil; i += (long long)gridDim.x * blockDim.x) {
    const bool inb = i < total;
    int c = 0;
    if (inb) {
      c = in ? ((int)in[i] > thr ? 1 : 0) : (int)cls[i];
      if (in) cls[i] = (uint8_t)c;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, inb && c == want);
    const unsigned int rs = __ballot_sync(0xffffffffu, inb && (i % W) == 0);
    if (!inb) continue;
    int label = (int)i;
    if (c == want) {
      const unsigned int starts = m & (~(m << 1) | rs | 1u);
      const unsigned int upto = starts & (0xffffffffu >> (31 - lane));
      label = (int)(i - (lane - (31 - __clz(upto))));
    }
    L[i] = label;
  }
}
// Unions between pixels of class `want`, only where connectivity is new: a pixel whose left neighbour already links the
// same two runs leaves the work to that neighbour.  conn8 adds the two upper diagonals.
__global__ void cc_merge_kernel(const uint8_t* __restrict__ cls, int want, int conn8, int H, int W, long long total,
                                int* __restrict__ L) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (cls[i] != want) continue;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const bool left = x > 0 && cls[i - 1] == want;
    if (left && (i & 31) == 0) uf_union(L, (int)i, (int)i - 1);          // runs continue across lane groups
    if (y == 0) continue;
    const bool up = cls[i - W] == want;
    const bool upleft = x > 0 && cls[i - W - 1] == want;
    if (up) {
      if (!(left && upleft)) uf_union(L, (int)i, (int)(i - W));
    } else if (conn8) {
      if (upleft && !left) uf_union(L, (int)i, (int)(i - W - 1));
      const bool upright = x + 1 < W && cls[i - W + 1] == want;
      const bool right = x + 1 < W && cls[i + 1] == want;
      if (upright && !right) uf_union(L, (int)i, (int)(i - W + 1));
    }
  }
}
__global__ void cc_compress_kernel(long long total, int* __restrict__ L) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    L[i] = uf_find(L, (int)i);
}
// clean_masks.py:16-22: background the 4-connected flood fill from (0,0) cannot reach becomes foreground;
// a foreground pixel at (0,0) turns the whole image into foreground.
__global__ void cc_fill_holes_kernel(const int* __restrict__ L, long long hw, long long total, uint8_t* __restrict__ cls) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long base = (i / hw) * hw;
    const bool corner_fg = cls[base] != 0;              // the corner pixel itself is never written here
    if (cls[i]) continue;
    if (i == base) continue;                            // the corner itself is reached by definition
    if (corner_fg || L[i] != L[base]) cls[i] = 1;       // filled
  }
}
// per component root: area and the first 2x2 block (block-row major) touching it — OpenCV's label order
__global__ void cc_stats_kernel(const uint8_t* __restrict__ cls, const int* __restrict__ L, int H, int W, long long total,
                                int* __restrict__ area, int* __restrict__ bkey) {
  const int bw = (W + 1) >> 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (!cls[i]) continue;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    atomicAdd(&area[L[i]], 1);
    atomicMin(&bkey[L[i]], (y >> 1) * bw + (x >> 1));
  }
}
__global__ void cc_reset_stats_kernel(long long total, int* __restrict__ area, int* __restrict__ bkey, int B,
                                      unsigned long long* __restrict__ best) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    area[i] = 0;
    bkey[i] = 0x7fffffff;
    if (i < B) best[i] = 0ull;
  }
}
// largest area wins, ties go to the smaller block key: maximise (area << 32) | ~bkey
__global__ void cc_select_kernel(const uint8_t* __restrict__ cls, const int* __restrict__ L, const int* __restrict__ area,
                                 const int* __restrict__ bkey, long long hw, long long total,
                                 unsigned long long* __restrict__ best) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (!cls[i] || L[i] != (int)i) continue;
    const unsigned long long key = ((unsigned long long)(unsigned int)area[i] << 32) | (0xffffffffu - (unsigned int)bkey[i]);
    atomicMax(&best[i / hw], key);
  }
}
__global__ void cc_write_kernel(const uint8_t* __restrict__ cls, const int* __restrict__ L, const int* __restrict__ area,
                                const int* __restrict__ bkey, const unsigned long long* __restrict__ best, long long hw,
                                long long total, int keep_largest, uint8_t* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    bool on = cls[i] != 0;
    if (on && keep_largest) {
      const int r = L[i];
      const unsigned long long key = ((unsigned long long)(unsigned int)area[r] << 32) | (0xffffffffu - (unsigned int)bkey[r]);
      on = key == best[i / hw];
    }
    out[i] = on ? 255 : 0;
  }
}

size_t mask_cleanup_scratch_bytes(int B, int H, int W) {
  const size_t px = (size_t)B * H * W;
  return ((px + 255) & ~(size_t)255) + 3 * ((px * 4 + 255) & ~(size_t)255) + (((size_t)B * 8 + 255) & ~(size_t)255);
}

cudaError_t launch_mask_cleanup(const uint8_t* mask, int B, int H, int W, int bin_thr, int fill_holes, int keep_largest,
                                uint8_t* out, void* scratch, cudaStream_t s) {
  const long long px = (long long)B * H * W, hw = (long long)H * W;
  if (px >= 0x7fffffffLL) return cudaErrorInvalidValue;
  uint8_t* b = static_cast<uint8_t*>(scratch);
  uint8_t* cls = b;
  size_t off = ((size_t)px + 255) & ~(size_t)255;
  const size_t isz = ((size_t)px * 4 + 255) & ~(size_t)255;
  int* L = reinterpret_cast<int*>(b + off); off += isz;
  int* area = reinterpret_cast<int*>(b + off); off += isz;
  int* bkey = reinterpret_cast<int*>(b + off); off += isz;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(b + off);
  const int grid = pp_grid(px, 256);
  cudaError_t e;
#define PP_LAUNCH(...)                              \
  do {                                              \
    __VA_ARGS__;                                    \
    if ((e = launched()) != cudaSuccess) return e;  \
  } while (0)
  PP_LAUNCH(cc_init_kernel<<<grid, 256, 0, s>>>(mask, bin_thr, fill_holes ? 0 : 1, W, px, cls, L));
  if (fill_holes) {
    PP_LAUNCH(cc_merge_kernel<<<grid, 256, 0, s>>>(cls, 0, 0, H, W, px, L));          // background, 4-connected
    PP_LAUNCH(cc_compress_kernel<<<grid, 256, 0, s>>>(px, L));
    PP_LAUNCH(cc_fill_holes_kernel<<<grid, 256, 0, s>>>(L, hw, px, cls));
    PP_LAUNCH(cc_init_kernel<<<grid, 256, 0, s>>>(nullptr, 0, 1, W, px, cls, L));
  }
  if (keep_largest) {
    PP_LAUNCH(cc_merge_kernel<<<grid, 256, 0, s>>>(cls, 1, 1, H, W, px, L));          // foreground, 8-connected
    PP_LAUNCH(cc_compress_kernel<<<grid, 256, 0, s>>>(px, L));
    PP_LAUNCH(cc_reset_stats_kernel<<<grid, 256, 0, s>>>(px, area, bkey, B, best));
    PP_LAUNCH(cc_stats_kernel<<<grid, 256, 0, s>>>(cls, L, H, W, px, area, bkey));
    PP_LAUNCH(cc_select_kernel<<<grid, 256, 0, s>>>(cls, L, area, bkey, hw, px, best));
  }
  PP_LAUNCH(cc_write_kernel<<<grid, 256, 0, s>>>(cls, L, area, bkey, best, hw, px, keep_largest, out));
#undef PP_LAUNCH
  return cudaSuccess;
}

}  // namespace cs
