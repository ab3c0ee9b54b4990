// Active Boundary Loss, binary case — replaces ABL.forward (src/training/losses/abl.py:66-212) with its
// LabelSmoothSoftmaxCEV1 criterion (src/training/losses/label_smooth.py:14-57), as wrapped by BCEDiceABL
// (src/training/train_BCEDice_ABL.py:264-302).
//
// The reference builds index lists with nonzero(), loops on the host until a KL threshold leaves few enough
// pixels (one sync per iteration), and runs scipy's EDT per image on the CPU.  Here everything is a dense
// per-pixel pass on the device, no host synchronisation:
//   1. abl_kl_kernel      KL(self || lower neighbour) + KL(self || right neighbour) of the two-way softmax over the
//                         probabilities (1-p, p)  -> kl map, plus a histogram over the whole ladder of thresholds
//                         1e-5 * 1.2^k the reference's loop could visit; the last block picks the first threshold
//                         that leaves <= H*W/100 pixels (counted over the batch, as the reference does).
//   2. abl_columns/rows   exact EDT of the GT boundary (boundary computed on the fly from the labels), emitted as
//                         max(0, floor(dist) - 1) — the reference's int32-truncated one_hot2dist, clamped.
//   3. abl_forward_kernel 3x3 dilation of (kl > eps), arg-min direction over the 9 neighbours of the distance map,
//                         8 neighbour KLs, label-smoothed cross entropy, distance weight, mean over kept pixels.
//   4. abl_backward_kernel the same traversal writing dlogits (gradient flows through the centre pixel only,
//                         neighbours are detached: abl.py:144-145).
// Reference behaviours kept on purpose are listed in oracle/abl_oracle.py (distance-map batch indexing, scipy's
// EDT of an input without zeros, label smoothing mass 1 - s/8).
#include "edt.cuh"

namespace cs {

static constexpr int kLadder = kAblLadder;      // thresholds 1e-5 * 1.2^k, k < 80 (the last ones exceed any KL)
static constexpr int kPadDist = 100000;         // abl.py:116 max_dis

struct AblResult {                              // lives at the start of the scratch buffer
  double sum;                                   // sum over kept pixels of ce * weight
  unsigned long long kept;                      // pixels with a direction != "stay"
  unsigned long long pred_boundary;             // pixels of the dilated predicted boundary
  float eps;                                    // the threshold chosen
  int k;                                        // its ladder index
  unsigned int counter_a, counter_b;            // last-block tickets
  unsigned int hist[kLadder + 1];               // hist[k] = pixels with exactly k ladder entries below their kl
};

size_t abl_scratch_bytes(int B, int H, int W) {
  const size_t px = (size_t)B * H * W;
  size_t off = (sizeof(AblResult) + 255) & ~(size_t)255;
  off += px * 4;                                // kl map (fp32)
  off += px * 2;                                // distance map (uint16)
  off = (off + 255) & ~(size_t)255;
  off += px * 4;                                // column distances (ushort2)
  off += (size_t)B * 8;                         // per-image flags
  return off + 256;
}

struct AblBuffers {
  AblResult* res; float* kl; unsigned short* dmap; ushort2* g; int* flags;
};
static AblBuffers carve(void* scratch, int B, int H, int W) {
  const size_t px = (size_t)B * H * W;
  uint8_t* b = static_cast<uint8_t*>(scratch);
  AblBuffers a;
  a.res = reinterpret_cast<AblResult*>(b);
  size_t off = (sizeof(AblResult) + 255) & ~(size_t)255;
  a.kl = reinterpret_cast<float*>(b + off); off += px * 4;
  a.dmap = reinterpret_cast<unsigned short*>(b + off); off += px * 2;
  off = (off + 255) & ~(size_t)255;
  a.g = reinterpret_cast<ushort2*>(b + off); off += px * 4;
  a.flags = reinterpret_cast<int*>(b + off);
  return a;
}

// two-way softmax over (1-p, p): log-probabilities l0, l1 and probability s1 of channel 1
struct Soft2 { float l0, l1, s0, s1; };
CS_DEVINL Soft2 soft2_of_logit(float z) {
  const float p = 1.0f / (1.0f + expf(-z));
  const float a0 = 1.0f - p, a1 = p;
  const float m = fmaxf(a0, a1);
  const float e0 = expf(a0 - m), e1 = expf(a1 - m);
  const float sum = e0 + e1, lse = logf(sum);
  Soft2 r;
  r.l0 = a0 - m - lse; r.l1 = a1 - m - lse;
  r.s0 = e0 / sum; r.s1 = e1 / sum;
  return r;
}
// abl.py:14-15 kl_div(a = center, b = other), summed over the two channels
CS_DEVINL float kl2(const Soft2& center, const Soft2& other) {
  return other.s0 * (other.l0 - center.l0) + other.s1 * (other.l1 - center.l1);
}

// ------------------------------------------------------------------------------------------------ 1. KL map
__global__ void __launch_bounds__(256) abl_kl_kernel(const float* __restrict__ logits, int B, int H, int W, float max_n,
                                                     const AblLadder ladder, float* __restrict__ kl,
                                                     AblResult* __restrict__ res) {
  __shared__ unsigned int sh[kLadder + 1];
  __shared__ float c_ladder[kLadder];
  for (int i = threadIdx.x; i <= kLadder; i += blockDim.x) sh[i] = 0;
  for (int i = threadIdx.x; i < kLadder; i += blockDim.x) c_ladder[i] = ladder.v[i];
  __syncthreads();
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const Soft2 self = soft2_of_logit(__ldg(logits + i));
    float ud = 0.f, lr = 0.f;
    if (y + 1 < H) ud = kl2(soft2_of_logit(__ldg(logits + i + W)), self);
    if (x + 1 < W) lr = kl2(soft2_of_logit(__ldg(logits + i + 1)), self);
    const float v = lr + ud;
    kl[i] = v;
    int k = 0;
    while (k < kLadder && v > c_ladder[k]) ++k;
    atomicAdd(&sh[k], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= kLadder; i += blockDim.x)
    if (sh[i]) atomicAdd(&res->hist[i], sh[i]);
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&res->counter_a, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  // count(kl > eps_j) = sum_{k > j} hist[k]; the reference raises j while that count exceeds max_n (float32 compare)
  const volatile unsigned int* h = res->hist;
  unsigned long long above = 0;
  for (int k = 1; k <= kLadder; ++k) above += h[k];
  int j = 0;
  while (j < kLadder - 1 && (float)above > max_n) { ++j; above -= h[j]; }
  res->k = j;
  res->eps = c_ladder[j];
}

// ------------------------------------------------------------------------------------------------ 2. distance maps
CS_DEVINL long long label_of(float t) { return (long long)t; }      // target.long(): truncation (abl.py:177)

// GT boundary (abl.py:94-107): the label differs from the pixel below or the pixel to the right, or is the ignore label.
struct BoundaryPred {
  const float* targets; long long ignore_label; int H, W;
  __device__ bool operator()(int img, int y, int x) const {
    const float* p = targets + ((size_t)img * H + y) * W + x;
    const long long cur = label_of(__ldg(p));
    const long long below = y + 1 < H ? label_of(__ldg(p + W)) : cur;
    const long long right = x + 1 < W ? label_of(__ldg(p + 1)) : cur;
    return below != cur || right != cur || cur == ignore_label;
  }
};

CS_DEVINL int isqrt_floor(int v) {
  int r = (int)sqrtf((float)v);
  while (r * r > v) --r;
  while ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

// faithful: image i feeds map 2i (channel 0: non-boundary pixels get floor(dist to boundary) - 1) and map 2i+1
// (channel 1: boundary pixels get floor(dist to non-boundary) - 1), the layout the reference's torch.cat over [2,H,W]
// maps produces (abl.py:166-167).  Otherwise map i = channel 0 of image i.
struct AblDistEpilogue {
  unsigned short* dmap; int B, H, W, faithful;
  __device__ void operator()(int img, int y, int x, bool bnd, int best_sq, bool has_b, bool has_n) const {
    // scipy's EDT of an input without any zero pixel measures from (row -1, column 0)
    const int best = (bnd ? has_n : has_b) ? best_sq : (y + 1) * (y + 1) + x * x;
    const int v = max(isqrt_floor(best) - 1, 0);
    const int n0 = faithful ? 2 * img : img, n1 = faithful ? 2 * img + 1 : -1;
    if (n0 < B) dmap[((size_t)n0 * H + y) * W + x] = bnd ? 0 : (unsigned short)v;
    if (n1 >= 0 && n1 < B) dmap[((size_t)n1 * H + y) * W + x] = bnd ? (unsigned short)v : 0;
  }
};

// Fallback for images taller than the segmented column pass supports.
__global__ void abl_columns_serial_kernel(BoundaryPred pred, int nimg, int H, int W, ushort2* __restrict__ g,
                                          int* __restrict__ flags) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nimg * W) return;
  const int b = idx / W, x = idx - b * W;
  ushort2* gc = g + (size_t)b * H * W + x;
  int any_b = 0, any_n = 0, db = kEdtInf, dn = kEdtInf;
  for (int y = 0; y < H; ++y) {
    const bool bnd = pred(b, y, x);
    db = bnd ? 0 : min(db + 1, kEdtInf);
    dn = bnd ? min(dn + 1, kEdtInf) : 0;
    any_b |= bnd; any_n |= !bnd;
    gc[(size_t)y * W] = make_ushort2((unsigned short)db, (unsigned short)dn);
  }
  db = kEdtInf; dn = kEdtInf;
  for (int y = H - 1; y >= 0; --y) {
    const ushort2 d = gc[(size_t)y * W];
    const bool bnd = d.x == 0;
    db = bnd ? 0 : min(db + 1, kEdtInf);
    dn = bnd ? min(dn + 1, kEdtInf) : 0;
    gc[(size_t)y * W] = make_ushort2((unsigned short)min((int)d.x, db), (unsigned short)min((int)d.y, dn));
  }
  if (any_b) atomicOr(&flags[2 * b], 1);
  if (any_n) atomicOr(&flags[2 * b + 1], 1);
}

// ------------------------------------------------------------------------------------------------ 3./4. loss
// (d_row, d_col) in the reference's order (abl.py:126-133)
__constant__ int c_dr[9] = {1, -1, 0, 0, -1, 1, -1, 1, 0};
__constant__ int c_dc[9] = {0, 0, -1, 1, 1, 1, -1, -1, 0};

struct AblPixel { bool kept; int dir; float w; float kls[8]; float s1c; float s1n[8]; float p; };

template <bool kNeedNeighbourProbs>
CS_DEVINL bool abl_pixel(const float* __restrict__ logits, const float* __restrict__ kl,
                         const unsigned short* __restrict__ dmap, int H, int W, long long img_off, int y, int x,
                         float eps, float max_clip, AblPixel& px, bool& on_boundary) {
  // dilated predicted boundary: any (kl > eps) in the 3x3 neighbourhood (abl.py:86-91)
  bool pb = false;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) pb |= __ldg(kl + img_off + (long long)yy * W + xx) > eps;
    }
  on_boundary = pb;
  px.kept = false;
  if (!pb) return false;
  int best = 0x7fffffff, dir = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + c_dr[k], xx = x + c_dc[k];
    const int d = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (int)dmap[img_off + (long long)yy * W + xx] : kPadDist;
    if (d < best) { best = d; dir = k; }        // strict <: the first minimum wins (torch.argmin)
  }
  if (dir == 8) return false;
  px.kept = true;
  px.dir = dir;
  const int dc = (int)dmap[img_off + (long long)y * W + x];
  px.w = fminf((float)dc, max_clip) / max_clip;
  const float zc = __ldg(logits + img_off + (long long)y * W + x);
  const Soft2 c = soft2_of_logit(zc);
  px.s1c = c.s1;
  px.p = 1.0f / (1.0f + expf(-zc));
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int yy = min(max(y + c_dr[k], 0), H - 1), xx = min(max(x + c_dc[k], 0), W - 1);   // replicate padding
    const Soft2 nb = soft2_of_logit(__ldg(logits + img_off + (long long)yy * W + xx));
    px.kls[k] = kl2(c, nb);
    if (kNeedNeighbourProbs) px.s1n[k] = nb.s1;
  }
  return true;
}

__global__ void __launch_bounds__(256) abl_forward_kernel(const float* __restrict__ logits, const float* __restrict__ kl,
                                                          const unsigned short* __restrict__ dmap, int B, int H, int W,
                                                          float smoothing, float max_clip, AblResult* __restrict__ res,
                                                          float* __restrict__ loss_out) {
  const float eps = res->eps;
  const float lb_neg = smoothing / 8.0f, lb_pos = 1.0f - smoothing;
  float sum = 0.f;
  unsigned int kept = 0, nb = 0;
  const long long total = (long long)B * H * W, hw = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long img_off = (i / hw) * hw;
    AblPixel px;
    bool on_b;
    const bool k = abl_pixel<false>(logits, kl, dmap, H, W, img_off, y, x, eps, max_clip, px, on_b);
    nb += on_b ? 1u : 0u;
    if (!k) continue;
    float m = px.kls[0];
#pragma unroll
    for (int c = 1; c < 8; ++c) m = fmaxf(m, px.kls[c]);
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) se += expf(px.kls[c] - m);
    const float lse = m + logf(se);
    float ce = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) ce -= (px.kls[c] - lse) * (c == px.dir ? lb_pos : lb_neg);
    sum += ce * px.w;
    ++kept;
  }
  // block reduction -> global accumulators
  __shared__ float s_sum[8];
  __shared__ unsigned int s_kept[8], s_nb[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  sum = warp_sum(sum);
  kept = __reduce_add_sync(0xffffffffu, kept);
  nb = __reduce_add_sync(0xffffffffu, nb);
  if (lane == 0) { s_sum[warp] = sum; s_kept[warp] = kept; s_nb[warp] = nb; }
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    double t = 0.0; unsigned long long k2 = 0, n2 = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += s_sum[w]; k2 += s_kept[w]; n2 += s_nb[w]; }
    if (k2) { atomicAdd(&res->sum, t); atomicAdd(&res->kept, k2); }
    if (n2) atomicAdd(&res->pred_boundary, n2);
    __threadfence();
    is_last = atomicAdd(&res->counter_b, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  const volatile AblResult* r = res;
  const bool valid = r->pred_boundary >= 1;               // abl.py:197-198: otherwise the reference returns None
  // mean over the kept pixels (abl.py:207); an empty selection is 0/0 = NaN exactly as torch's mean of nothing
  loss_out[0] = valid ? (float)(r->sum / (double)r->kept) : 0.f;
  loss_out[1] = valid ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) abl_backward_kernel(const float* __restrict__ logits, const float* __restrict__ kl,
                                                           const unsigned short* __restrict__ dmap, int B, int H, int W,
                                                           float smoothing, float max_clip,
                                                           const AblResult* __restrict__ res,
                                                           const float* __restrict__ grad_out, float* __restrict__ dlogits) {
  const float eps = res->eps;
  const float lb_neg = smoothing / 8.0f, lb_pos = 1.0f - smoothing;
  const float lb_sum = lb_pos + 7.0f * lb_neg;
  const unsigned long long kept = res->kept;
  const float go = (grad_out ? grad_out[0] : 1.f);
  const float scale = (res->pred_boundary >= 1 && kept) ? go / (float)kept : 0.f;
  const long long total = (long long)B * H * W, hw = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long img_off = (i / hw) * hw;
    AblPixel px;
    bool on_b;
    float gz = 0.f;
    if (abl_pixel<true>(logits, kl, dmap, H, W, img_off, y, x, eps, max_clip, px, on_b)) {
      float m = px.kls[0];
#pragma unroll
      for (int c = 1; c < 8; ++c) m = fmaxf(m, px.kls[c]);
      float e[8], se = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) { e[c] = expf(px.kls[c] - m); se += e[c]; }
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float dce = lb_sum * (e[c] / se) - (c == px.dir ? lb_pos : lb_neg);    // d ce / d kl_c
        acc += dce * 2.0f * (px.s1c - px.s1n[c]);                                    // d kl_c / d p(centre)
      }
      gz = scale * px.w * acc * px.p * (1.0f - px.p);
    }
    dlogits[i] = gz;
  }
}

static int grid_1d(long long work, int per_block) {
  long long g = (work + per_block - 1) / per_block;
  if (g > 148 * 8) g = 148 * 8;
  return g < 1 ? 1 : (int)g;
}

cudaError_t launch_abl_forward(const float* logits, const float* targets, int B, int H, int W, const AblLadder& ladder,
                               float max_n, float smoothing, float max_clip, long long ignore_label, int faithful,
                               void* scratch, float* loss_out, cudaStream_t s) {
  if (H >= kEdtInf || W >= kEdtInf) return cudaErrorInvalidValue;
  const AblBuffers a = carve(scratch, B, H, W);
  cudaError_t e = cudaMemsetAsync(a.res, 0, sizeof(AblResult), s);
  if (e != cudaSuccess) return e;
  const long long px = (long long)B * H * W;
  abl_kl_kernel<<<grid_1d(px, 256 * 4), 256, 0, s>>>(logits, B, H, W, max_n, ladder, a.kl, a.res);
  if ((e = launched()) != cudaSuccess) return e;
  const int nimg = faithful ? (B + 1) / 2 : B;
  e = cudaMemsetAsync(a.flags, 0, (size_t)nimg * 8, s);
  if (e != cudaSuccess) return e;
  const BoundaryPred pred{targets, ignore_label, H, W};
  int nseg = 0, rps = 0;
  if (edt_column_geometry(H, &nseg, &rps)) {
    edt_columns_kernel<<<nimg * ((W + 31) / 32), dim3(32, nseg), 0, s>>>(pred, nimg, H, W, rps, a.g, a.flags);
  } else {
    abl_columns_serial_kernel<<<(nimg * W + 63) / 64, 64, 0, s>>>(pred, nimg, H, W, a.g, a.flags);
  }
  if ((e = launched()) != cudaSuccess) return e;
  const AblDistEpilogue epi{a.dmap, B, H, W, faithful};
  const int threads = W >= 256 ? 256 : ((W + 31) / 32) * 32;
  if (W <= kEdtMaxPaddedW) edt_rows_kernel<true><<<nimg * H, threads, edt_rows_smem(W, true), s>>>(epi, a.g, a.flags, H, W);
  else edt_rows_kernel<false><<<nimg * H, threads, edt_rows_smem(W, false), s>>>(epi, a.g, a.flags, H, W);
  if ((e = launched()) != cudaSuccess) return e;
  abl_forward_kernel<<<grid_1d(px, 256), 256, 0, s>>>(logits, a.kl, a.dmap, B, H, W, smoothing, max_clip, a.res, loss_out);
  return launched();
}

cudaError_t launch_abl_backward(const float* logits, int B, int H, int W, float smoothing, float max_clip,
                                const void* scratch, const float* grad_out, float* dlogits, cudaStream_t s) {
  const AblBuffers a = carve(const_cast<void*>(scratch), B, H, W);
  const long long px = (long long)B * H * W;
  abl_backward_kernel<<<grid_1d(px, 256), 256, 0, s>>>(logits, a.kl, a.dmap, B, H, W, smoothing, max_clip, a.res, grad_out,
                                                       dlogits);
  return launched();
}

cudaError_t abl_debug_read(const void* scratch, int B, int H, int W, float* eps, int* k, unsigned long long* kept,
                           unsigned long long* pred_boundary, unsigned short* dmap_out, float* kl_out, cudaStream_t s) {
  const AblBuffers a = carve(const_cast<void*>(scratch), B, H, W);
  AblResult r;
  cudaError_t e = cudaMemcpyAsync(&r, a.res, sizeof(r), cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return e;
  const size_t px = (size_t)B * H * W;
  if (dmap_out && (e = cudaMemcpyAsync(dmap_out, a.dmap, px * 2, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  if (kl_out && (e = cudaMemcpyAsync(kl_out, a.kl, px * 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
  if (eps) *eps = r.eps;
  if (k) *k = r.k;
  if (kept) *kept = r.kept;
  if (pred_boundary) *pred_boundary = r.pred_boundary;
  return cudaSuccess;
}

}  // namespace cs
