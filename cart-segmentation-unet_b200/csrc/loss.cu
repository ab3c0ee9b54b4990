// Fused segmentation losses (reference: BCEDiceLoss train_bce_dice.py:186-199, FocalLoss/FocalDiceLoss
// src/train_with_focalDice.py:195-235, SymmetricBoundaryLoss/CompositeSegLoss
// src/train_with_boundary_loss.py:242-282) and thresholded metrics (train_bce_dice.py:201-232).
//
// One forward pass over (logits, targets[, sdf_gt, sdf_pred]) produces every partial sum the loss
// needs (per-row Dice sums, element-wise BCE/focal sum, the two boundary sums); the last block to
// finish turns them into the scalar.  One backward pass re-reads the inputs and writes dlogits,
// scaled by the incoming grad_output device scalar (no host sync, GradScaler-compatible).
#include "common.cuh"
#include "kernels.cuh"

namespace cs {

enum { S_PT = 0, S_P = 1, S_T = 2, S_ELEM = 3, S_BGT = 4, S_BPRED = 5, S_STRIDE = 8 };

size_t loss_scratch_bytes(int rows) { return (size_t)rows * S_STRIDE * sizeof(double) + 16; }

CS_DEVINL float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
CS_DEVINL float bce_logits(float x, float t) { return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x))); }
CS_DEVINL float pow_gamma(float q, float gamma) {
  if (gamma == 0.f) return 1.f;
  if (gamma == 1.f) return q;
  if (gamma == 2.f) return q * q;
  return powf(q, gamma);
}

template <int NV>
CS_DEVINL void block_reduce_add(float* v, double* dst) {   // dst[i] += sum over block of v[i]
  __shared__ float sred[NV][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) sred[i][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) s += sred[threadIdx.x][w];
    atomicAdd(&dst[threadIdx.x], (double)s);
  }
}

__global__ void __launch_bounds__(256) loss_forward_kernel(LossArgs a) {
  const int row = blockIdx.y;
  const float* x = a.logits + (size_t)row * a.n;
  const float* t = a.targets + (size_t)row * a.n;
  const float* sg = a.sdf_gt ? a.sdf_gt + (size_t)row * a.n : nullptr;
  const float* sp = a.sdf_pred ? a.sdf_pred + (size_t)row * a.n : nullptr;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long n4 = a.n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 tv = __ldg(reinterpret_cast<const float4*>(t) + i);
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), pv = gv;
    if (sg) gv = __ldg(reinterpret_cast<const float4*>(sg) + i);
    if (sp) pv = __ldg(reinterpret_cast<const float4*>(sp) + i);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ts[4] = {tv.x, tv.y, tv.z, tv.w};
    const float gs[4] = {gv.x, gv.y, gv.z, gv.w}, ps[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = sigmoidf_(xs[j]);
      acc[S_PT] += p * ts[j];
      acc[S_P] += p;
      acc[S_T] += ts[j];
      if (a.w_elem != 0.f) {
        const float q = (ts[j] == 1.0f) ? 1.0f - p : p;          // 1 - p_t
        acc[S_ELEM] += a.alpha * pow_gamma(q, a.gamma) * bce_logits(xs[j], ts[j]);
      }
      if (sg) { const float v = p * gs[j]; acc[S_BGT] += a.use_abs ? fabsf(v) : v; }
      if (sp) { const float v = (1.0f - p) * (-ps[j]); acc[S_BPRED] += a.use_abs ? fabsf(v) : v; }
    }
  }
  block_reduce_add<6>(acc, a.stats + (size_t)row * S_STRIDE);

  // ---- last block finalises the scalar(s)
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  unsigned int* counter = reinterpret_cast<unsigned int*>(a.stats + (size_t)a.rows * S_STRIDE);
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    is_last = atomicAdd(counter, 1u) == total - 1;
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  const volatile double* st = a.stats;
  const double n = (double)a.n, rows = (double)a.rows;
  if (a.per_row) {
    for (int r = 0; r < a.rows; ++r) {
      const volatile double* s = st + (size_t)r * S_STRIDE;
      const double dice = (2.0 * s[S_PT] + a.smooth) / (s[S_P] + s[S_T] + a.smooth);
      double v = a.w_dice * (1.0 - dice) + a.w_elem * (a.elem_sum ? s[S_ELEM] : s[S_ELEM] / n);
      v += (a.w_bgt * s[S_BGT] + a.w_bpred * s[S_BPRED]) / n;
      a.loss_out[r] = (float)v;
    }
  } else {
    double elem = 0.0, dice = 0.0, bg = 0.0, bp = 0.0;
    for (int r = 0; r < a.rows; ++r) {
      const volatile double* s = st + (size_t)r * S_STRIDE;
      elem += s[S_ELEM]; bg += s[S_BGT]; bp += s[S_BPRED];
      dice += (2.0 * s[S_PT] + a.smooth) / (s[S_P] + s[S_T] + a.smooth);
    }
    double v = a.w_dice * (1.0 - dice / rows) + a.w_elem * (a.elem_sum ? elem : elem / (rows * n));
    v += (a.w_bgt * bg + a.w_bpred * bp) / (rows * n);
    a.loss_out[0] = (float)v;
  }
}

cudaError_t launch_loss_forward(const LossArgs& a, cudaStream_t s) {
  if (a.n % 4 != 0) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(a.stats, 0, loss_scratch_bytes(a.rows), s);
  if (e != cudaSuccess) return e;
  long long bpr = (a.n / 4 + 256 * 4 - 1) / (256 * 4);
  const long long cap = (148 * 8 + a.rows - 1) / a.rows;
  if (bpr > cap) bpr = cap;
  if (bpr < 1) bpr = 1;
  dim3 grid((unsigned)bpr, (unsigned)a.rows);
  loss_forward_kernel<<<grid, 256, 0, s>>>(a);
  return launched();
}

__global__ void __launch_bounds__(256) loss_backward_kernel(LossArgs a) {
  const int row = blockIdx.y;
  const float* x = a.logits + (size_t)row * a.n;
  const float* t = a.targets + (size_t)row * a.n;
  const float* sg = a.sdf_gt ? a.sdf_gt + (size_t)row * a.n : nullptr;
  const float* sp = a.sdf_pred ? a.sdf_pred + (size_t)row * a.n : nullptr;
  float* dx = a.dlogits + (size_t)row * a.n;
  const double* st = a.stats + (size_t)row * S_STRIDE;
  const float go = a.grad_out ? (a.per_row ? a.grad_out[row] : a.grad_out[0]) : 1.f;
  const double n = (double)a.n, rows = (double)a.rows;
  const double D = st[S_P] + st[S_T] + a.smooth;
  const double twoI = 2.0 * st[S_PT] + a.smooth;
  // d(1 - dice)/dp_i = -(2 t_i D - twoI) / D^2  ->  k_t * t_i + k_0
  const float cd = (float)(a.per_row ? a.w_dice : a.w_dice / rows);
  const float k_t = (float)(-2.0 / D) * cd * go;
  const float k_0 = (float)(twoI / (D * D)) * cd * go;
  const float ce = go * (float)(a.elem_sum ? a.w_elem : (a.per_row ? a.w_elem / n : a.w_elem / (rows * n)));
  const float cb = go * (float)(a.per_row ? 1.0 / n : 1.0 / (rows * n));
  const float cbg = cb * a.w_bgt, cbp = cb * a.w_bpred;
  const long long n4 = a.n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 tv = __ldg(reinterpret_cast<const float4*>(t) + i);
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), pv = gv;
    if (sg) gv = __ldg(reinterpret_cast<const float4*>(sg) + i);
    if (sp) pv = __ldg(reinterpret_cast<const float4*>(sp) + i);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ts[4] = {tv.x, tv.y, tv.z, tv.w};
    const float gs[4] = {gv.x, gv.y, gv.z, gv.w}, ps[4] = {pv.x, pv.y, pv.z, pv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = sigmoidf_(xs[j]);
      const float dp = p * (1.0f - p);
      float gsum = (k_t * ts[j] + k_0) * dp;                      // Dice
      if (a.w_elem != 0.f) {
        float ge;
        if (a.gamma == 0.f) {
          ge = a.alpha * (p - ts[j]);
        } else {
          const bool pos = ts[j] == 1.0f;
          const float q = pos ? 1.0f - p : p;
          const float dq = pos ? -dp : dp;
          const float qg1 = pow_gamma(q, a.gamma - 1.0f);
          ge = a.alpha * (a.gamma * qg1 * dq * bce_logits(xs[j], ts[j]) + qg1 * q * (p - ts[j]));
        }
        gsum += ce * ge;
      }
      if (sg) gsum += cbg * dp * (a.use_abs ? fabsf(gs[j]) : gs[j]);
      if (sp) gsum += cbp * dp * (a.use_abs ? -fabsf(ps[j]) : ps[j]);
      o[j] = gsum;
    }
    reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

cudaError_t launch_loss_backward(const LossArgs& a, cudaStream_t s) {
  if (a.n % 4 != 0) return cudaErrorInvalidValue;
  long long bpr = (a.n / 4 + 256 * 4 - 1) / (256 * 4);
  const long long cap = (148 * 8 + a.rows - 1) / a.rows;
  if (bpr > cap) bpr = cap;
  if (bpr < 1) bpr = 1;
  dim3 grid((unsigned)bpr, (unsigned)a.rows);
  loss_backward_kernel<<<grid, 256, 0, s>>>(a);
  return launched();
}

// ============================================================================ focal loss, reduction = "none"
// FocalLoss(reduction="none") of src/train_with_focalDice.py:214-219: the unreduced map alpha * (1 - p_t)^gamma * BCE and
// its gradient against an element-wise grad_out.  Same per-element arithmetic as the fused kernels above.
__global__ void __launch_bounds__(256) focal_map_forward_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                               long long n, float alpha, float gamma,
                                                               float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = __ldg(x + i), tv = __ldg(t + i);
    const float p = sigmoidf_(xv);
    const float q = (tv == 1.0f) ? 1.0f - p : p;
    out[i] = alpha * pow_gamma(q, gamma) * bce_logits(xv, tv);
  }
}
__global__ void __launch_bounds__(256) focal_map_backward_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                                const float* __restrict__ go, long long n, float alpha,
                                                                float gamma, float* __restrict__ dx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = __ldg(x + i), tv = __ldg(t + i);
    const float p = sigmoidf_(xv);
    float ge;
    if (gamma == 0.f) {
      ge = alpha * (p - tv);
    } else {
      const float dp = p * (1.0f - p);
      const bool pos = tv == 1.0f;
      const float q = pos ? 1.0f - p : p;
      const float dq = pos ? -dp : dp;
      const float qg1 = pow_gamma(q, gamma - 1.0f);
      ge = alpha * (gamma * qg1 * dq * bce_logits(xv, tv) + qg1 * q * (p - tv));
    }
    dx[i] = ge * __ldg(go + i);
  }
}
static int focal_grid(long long n) {
  long long g = (n + 256 * 4 - 1) / (256 * 4);
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}
cudaError_t launch_focal_map_forward(const float* x, const float* t, long long n, float alpha, float gamma, float* out,
                                     cudaStream_t s) {
  focal_map_forward_kernel<<<focal_grid(n), 256, 0, s>>>(x, t, n, alpha, gamma, out);
  return launched();
}
cudaError_t launch_focal_map_backward(const float* x, const float* t, const float* go, long long n, float alpha,
                                      float gamma, float* dx, cudaStream_t s) {
  focal_map_backward_kernel<<<focal_grid(n), 256, 0, s>>>(x, t, go, n, alpha, gamma, dx);
  return launched();
}

// ============================================================================ thresholded metrics
// pred_k = (x >= xs[k]).  The host turns "sigmoid(x) > t" / ">= t" into the exact fp32 bound xs[k]
// (smallest float whose ATen sigmoid passes the test), so masks are bit-identical to the
// reference's sigmoid-then-compare (train_bce_dice.py:209, create_pseudo_labels_gpu.py:294).
template <int KMAX>
__global__ void __launch_bounds__(256) threshold_stats_kernel(const float* __restrict__ logits,
                                                             const float* __restrict__ targets, long long n,
                                                             const float* __restrict__ xs, int K,
                                                             double* __restrict__ counts, double* __restrict__ soft) {
  __shared__ float sx[KMAX];
  __shared__ float sacc[2 * KMAX + 3];
  if (threadIdx.x < KMAX) sx[threadIdx.x] = threadIdx.x < K ? xs[threadIdx.x] : __int_as_float(0x7f800000);
  for (int i = threadIdx.x; i < 2 * KMAX + 3; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int row = blockIdx.y;
  const float* x = logits + (size_t)row * n;
  const float* t = targets + (size_t)row * n;
  float cp[KMAX], ci[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) cp[k] = ci[k] = 0.f;
  float sp = 0.f, stt = 0.f, spt = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = __ldg(x + i), tv = __ldg(t + i);
    const float p = sigmoidf_(xv);
    sp += p; stt += tv; spt += p * tv;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const float pred = xv >= sx[k] ? 1.f : 0.f;
      cp[k] += pred;
      ci[k] += pred * tv;
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const float a = warp_sum(cp[k]), b = warp_sum(ci[k]);
    if ((threadIdx.x & 31) == 0 && k < K) { atomicAdd(&sacc[2 * k], a); atomicAdd(&sacc[2 * k + 1], b); }
  }
  {
    const float a = warp_sum(sp), b = warp_sum(stt), c = warp_sum(spt);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sacc[2 * KMAX], a); atomicAdd(&sacc[2 * KMAX + 1], b); atomicAdd(&sacc[2 * KMAX + 2], c); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) atomicAdd(&counts[(size_t)row * 2 * K + i], (double)sacc[i]);
  if (threadIdx.x < 3) atomicAdd(&soft[(size_t)row * 3 + threadIdx.x], (double)sacc[2 * KMAX + threadIdx.x]);
}

cudaError_t launch_threshold_stats(const float* logits, const float* targets, int rows, long long n, const float* xs,
                                   int K, double* counts, double* soft, cudaStream_t s) {
  if (K < 1 || K > 32) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)rows * 2 * K * sizeof(double), s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(soft, 0, (size_t)rows * 3 * sizeof(double), s);
  if (e != cudaSuccess) return e;
  // per-thread partial counts are exact in fp32 as long as one thread sees < 2^24 pixels
  long long bpr = (n + 256 * 16 - 1) / (256 * 16);
  const long long cap = (148 * 8 + rows - 1) / rows;
  if (bpr > cap) bpr = cap;
  if (bpr < 1) bpr = 1;
  dim3 grid((unsigned)bpr, (unsigned)rows);
  if (K <= 4) threshold_stats_kernel<4><<<grid, 256, 0, s>>>(logits, targets, n, xs, K, counts, soft);
  else if (K <= 16) threshold_stats_kernel<16><<<grid, 256, 0, s>>>(logits, targets, n, xs, K, counts, soft);
  else threshold_stats_kernel<32><<<grid, 256, 0, s>>>(logits, targets, n, xs, K, counts, soft);
  return launched();
}

__global__ void threshold_mask_kernel(const float* __restrict__ logits, long long n4, float xstar,
                                      uint32_t* __restrict__ mask) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(logits) + i);
    mask[i] = (v.x >= xstar ? 1u : 0u) | (v.y >= xstar ? 0x100u : 0u) | (v.z >= xstar ? 0x10000u : 0u) |
              (v.w >= xstar ? 0x1000000u : 0u);
  }
}
cudaError_t launch_threshold_mask(const float* logits, long long n, float xstar, uint8_t* mask, cudaStream_t s) {
  if (n % 4 != 0) return cudaErrorInvalidValue;
  long long g = (n / 4 + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  threshold_mask_kernel<<<(unsigned)g, 256, 0, s>>>(logits, n / 4, xstar, reinterpret_cast<uint32_t*>(mask));
  return launched();
}

}  // namespace cs
