// Bandwidth-bound kernels around the convolutions: weight (un)packing, input im2col, batch-norm
// finalize / apply / backward, ReLU, 2x2 max-pool (+ its backward routing), the 1x1 head.
// All of them are single coalesced passes with 16-byte vectors (8 bf16 channels per thread).
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.cuh"

namespace cs {

static inline int grid_for(long long work, int block, int cap = 148 * 16) {
  long long g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

struct alignas(16) Vec8 { uint32_t w[4]; };
CS_DEVINL void unpack8(const Vec8& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = bf16_lo(v.w[i]); f[2 * i + 1] = bf16_hi(v.w[i]); }
}
CS_DEVINL Vec8 pack8(const float* f) {
  Vec8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.w[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  return v;
}
CS_DEVINL Vec8 ld8(const bf16* p) { return *reinterpret_cast<const Vec8*>(p); }
CS_DEVINL void st8(bf16* p, const Vec8& v) { *reinterpret_cast<Vec8*>(p) = v; }
CS_DEVINL Vec8 ld8_nc(const bf16* p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  Vec8 r;
  r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
  return r;
}

// ============================================================================ weight packing
// Tile of 32 (a) x 32 (b) pairs, T <= 9 taps each.  Thread (b, a) reads the T contiguous taps of its pair (a warp
// covers 32*T contiguous floats), writes out_ab directly (64-byte runs of b) and stages bf16 values in shared memory
// for the transposed copy out_ba (64-byte runs of a).
template <int T>
CS_DEVINL void pack_pairs_tile(bf16 (*tile)[32][34], const float* __restrict__ in, int Na, int Nb, bf16* __restrict__ out_ab,
                               const TapMap& map_ab, bf16* __restrict__ out_ba, const TapMap& map_ba, int tile_x, int tile_y) {
  const int a0 = tile_y * 32, b0 = tile_x * 32;
  const int b = b0 + threadIdx.x;
#pragma unroll
  for (int ai = threadIdx.y; ai < 32; ai += 8) {
    const int a = a0 + ai;
    if (a < Na && b < Nb) {
      const float* src = in + ((size_t)a * Nb + b) * T;
      float v[T];
#pragma unroll
      for (int t = 0; t < T; ++t) v[t] = __ldg(src + t);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const bf16 h = __float2bfloat16(v[t]);
        if (out_ab) out_ab[((size_t)map_ab.v[t] * Na + a) * Nb + b] = h;
        tile[t][threadIdx.x][ai] = h;
      }
    }
  }
  if (!out_ba) return;
  __syncthreads();
  const int a = a0 + threadIdx.x;
#pragma unroll
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int bi = threadIdx.y; bi < 32; bi += 8) {
      if (a < Na && b0 + bi < Nb) out_ba[((size_t)map_ba.v[t] * Nb + b0 + bi) * Na + a] = tile[t][bi][threadIdx.x];
    }
  }
}
template <int T>
__global__ void __launch_bounds__(256) pack_pairs_kernel(const float* __restrict__ in, int Na, int Nb,
                                                         bf16* __restrict__ out_ab, TapMap map_ab,
                                                         bf16* __restrict__ out_ba, TapMap map_ba) {
  __shared__ bf16 tile[T][32][34];                       // [t][b][a], padded: conflict-free both ways
  pack_pairs_tile<T>(tile, in, Na, Nb, out_ab, map_ab, out_ba, map_ba, blockIdx.x, blockIdx.y);
}
cudaError_t launch_pack_pairs(const float* in, int Na, int Nb, int T, bf16* out_ab, TapMap map_ab, bf16* out_ba,
                              TapMap map_ba, cudaStream_t s) {
  dim3 grid((Nb + 31) / 32, (Na + 31) / 32), block(32, 8);
  if (T == 9) pack_pairs_kernel<9><<<grid, block, 0, s>>>(in, Na, Nb, out_ab, map_ab, out_ba, map_ba);
  else if (T == 4) pack_pairs_kernel<4><<<grid, block, 0, s>>>(in, Na, Nb, out_ab, map_ab, out_ba, map_ba);
  else return cudaErrorInvalidValue;
  return launched();
}

// All weight tensors of the network in ONE launch (the per-tensor launches above were 42 launches = 0.29 ms of a 18 ms
// training step): block -> (job, 32x32 tile) through the jobs' tile prefix; the Cin=3 stem rides in the tail blocks.
__global__ void __launch_bounds__(256) pack_batch_kernel(const __grid_constant__ PackBatch pb) {
  __shared__ bf16 tile[9][32][34];
  const int blk = blockIdx.x;
  if (blk >= pb.total_tiles) {                           // stem: W[Cout][Cin<=7][3][3] -> [Cout][64], k = tap*Cin + c
    const int i = (blk - pb.total_tiles) * 256 + threadIdx.y * 32 + threadIdx.x;
    if (i < pb.first_cout * 64) {
      const int co = i >> 6, k = i & 63;
      float v = 0.f;
      if (k < 9 * pb.first_cin) {
        const int tap = k / pb.first_cin, c = k - tap * pb.first_cin;
        v = pb.first_in[((size_t)co * pb.first_cin + c) * 9 + tap];
      }
      pb.first_out[i] = __float2bfloat16(v);
    }
    return;
  }
  int j = 0;
#pragma unroll 1
  while (j + 1 < pb.n && blk >= pb.job[j + 1].tile0) ++j;
  const PackJob& jb = pb.job[j];
  const int local = blk - jb.tile0;
  const int ty = local / jb.tiles_x, tx = local - ty * jb.tiles_x;
  if (jb.T == 9) pack_pairs_tile<9>(tile, jb.in, jb.Na, jb.Nb, jb.out_ab, pb.maps[jb.map_ab], jb.out_ba, pb.maps[jb.map_ba], tx, ty);
  else pack_pairs_tile<4>(reinterpret_cast<bf16(*)[32][34]>(tile), jb.in, jb.Na, jb.Nb, jb.out_ab, pb.maps[jb.map_ab], jb.out_ba,
                          pb.maps[jb.map_ba], tx, ty);
}
cudaError_t launch_pack_batch(PackBatch& pb, cudaStream_t s) {
  int tiles = 0;
  for (int j = 0; j < pb.n; ++j) {
    PackJob& jb = pb.job[j];
    if (jb.T != 9 && jb.T != 4) return cudaErrorInvalidValue;
    jb.tiles_x = (jb.Nb + 31) / 32;
    jb.tile0 = tiles;
    tiles += jb.tiles_x * ((jb.Na + 31) / 32);
  }
  pb.total_tiles = tiles;
  const int first_blocks = pb.first_in ? (pb.first_cout * 64 + 255) / 256 : 0;
  if (tiles + first_blocks == 0) return cudaSuccess;
  pack_batch_kernel<<<tiles + first_blocks, dim3(32, 8), 0, s>>>(pb);
  return launched();
}

__global__ void pack_first_kernel(const float* __restrict__ in, int Cout, int Cin, bf16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 64) return;
  const int co = i >> 6, k = i & 63;
  float v = 0.f;
  if (k < 9 * Cin) {
    const int tap = k / Cin, c = k - tap * Cin;
    v = in[((size_t)co * Cin + c) * 9 + tap];
  }
  out[i] = __float2bfloat16(v);
}
cudaError_t launch_pack_first(const float* in, int Cout, int Cin, bf16* out, cudaStream_t s) {
  pack_first_kernel<<<(Cout * 64 + 255) / 256, 256, 0, s>>>(in, Cout, Cin, out);
  return launched();
}

// dwp[t][pair] -> grad[pair*T + map[t]]: 256 pairs per block staged through shared memory so that both the reads
// (runs of pairs) and the writes (256*T contiguous floats) are coalesced.
template <int T>
__global__ void __launch_bounds__(256) unpack_pairs_kernel(const float* __restrict__ dwp, size_t pairs, TapMap map,
                                                           float* __restrict__ grad) {
  __shared__ float st[256 * T];
  for (size_t p0 = (size_t)blockIdx.x * 256; p0 < pairs; p0 += (size_t)gridDim.x * 256) {
    const size_t p = p0 + threadIdx.x;
    if (p < pairs) {
#pragma unroll
      for (int t = 0; t < T; ++t) st[threadIdx.x * T + map.v[t]] = __ldg(dwp + (size_t)t * pairs + p);
    }
    __syncthreads();
    const size_t n = (pairs - p0 < 256 ? pairs - p0 : 256) * T;
    for (size_t i = threadIdx.x; i < n; i += 256) grad[p0 * T + i] = st[i];
    __syncthreads();
  }
}
cudaError_t launch_unpack_pairs(const float* dwp, int Na, int Nb, int T, TapMap map, float* grad, cudaStream_t s) {
  const size_t pairs = (size_t)Na * Nb;
  const int grid = grid_for((long long)pairs, 256);
  if (T == 9) unpack_pairs_kernel<9><<<grid, 256, 0, s>>>(dwp, pairs, map, grad);
  else if (T == 4) unpack_pairs_kernel<4><<<grid, 256, 0, s>>>(dwp, pairs, map, grad);
  else return cudaErrorInvalidValue;
  return launched();
}

__global__ void unpack_first_kernel(const float* __restrict__ dwp, int Cout, int Cin, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * 9) return;
  const int co = i / (Cin * 9), rem = i - co * Cin * 9;
  const int c = rem / 9, tap = rem - c * 9;
  grad[i] = dwp[co * 64 + tap * Cin + c];
}
cudaError_t launch_unpack_first(const float* dwp, int Cout, int Cin, float* grad, cudaStream_t s) {
  unpack_first_kernel<<<(Cout * Cin * 9 + 255) / 256, 256, 0, s>>>(dwp, Cout, Cin, grad);
  return launched();
}

// ============================================================================ input im2col
// x fp32 NCHW -> col bf16 [pixels][64] with k = (kh*3+kw)*Cin + c (zero padded to 64).  One block builds 128
// consecutive pixels of one image row band: the (3 x (128+2) x Cin) fp32 patch is staged in shared memory with
// coalesced loads, then every thread emits 16-byte groups of 8 k-values so that a warp writes 512 contiguous bytes.
static constexpr int kI2cPix = 128;
__global__ void __launch_bounds__(256) im2col_first_kernel(const float* __restrict__ x, int B, int Cin, int H, int W,
                                                          bf16* __restrict__ col) {
  extern __shared__ float patch[];                       // [Cin][3][kI2cPix + 2]
  const int PW = kI2cPix + 2;
  const int wblocks = (W + kI2cPix - 1) / kI2cPix;
  const int blk = blockIdx.x;
  const int wb = blk % wblocks;
  const int h = (blk / wblocks) % H;
  const int b = blk / (wblocks * H);
  const int w0 = wb * kI2cPix;
  for (int i = threadIdx.x; i < Cin * 3 * PW; i += blockDim.x) {
    const int pw = i % PW, r = (i / PW) % 3, c = i / (3 * PW);
    const int hh = h + r - 1, ww = w0 + pw - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(&x[(((size_t)b * Cin + c) * H + hh) * W + ww]);
    patch[i] = v;
  }
  __shared__ int koff[64];                               // k -> offset of its tap inside the patch (-1: zero padding)
  if (threadIdx.x < 64) {
    const int k = threadIdx.x;
    int off = -1;
    if (k < 9 * Cin) {
      const int tap = k / Cin, c = k - tap * Cin;
      off = (c * 3 + tap / 3) * PW + tap % 3;
    }
    koff[k] = off;
  }
  __syncthreads();
  bf16* dst = col + (((size_t)b * H + h) * W + w0) * 64;
  for (int i = threadIdx.x; i < kI2cPix * 8; i += blockDim.x) {
    const int kg = i & 7, p = i >> 3;
    if (w0 + p >= W) continue;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int off = koff[kg * 8 + j];
      f[j] = off >= 0 ? patch[off + p] : 0.f;
    }
    st8(dst + (size_t)i * 8, pack8(f));
  }
}
cudaError_t launch_im2col_first(const float* x, int B, int Cin, int H, int W, bf16* col, cudaStream_t s) {
  if (9 * Cin > 64) return cudaErrorInvalidValue;
  const int wblocks = (W + kI2cPix - 1) / kI2cPix;
  const size_t smem = (size_t)Cin * 3 * (kI2cPix + 2) * sizeof(float);
  im2col_first_kernel<<<B * H * wblocks, 256, smem, s>>>(x, B, Cin, H, W, col);
  return launched();
}

// ============================================================================ batch norm: statistics -> affine
// One channel: batch statistics -> (scale, shift); `publish` also stores the per-channel vectors the backward pass
// reads and updates the running statistics (momentum, unbiased variance), exactly once per forward.
CS_DEVINL void bn_channel_coef(const BnFinalizeArgs& a, int c, bool publish, float* sc_out, float* sh_out) {
  const double mean = a.sum[c] / a.count;
  double var = a.sq[c] / a.count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double invstd = rsqrt(var + (double)a.eps);
  const float sc = (float)((double)a.gamma[c] * invstd);
  const float sh = (float)((double)a.beta[c] - mean * (double)a.gamma[c] * invstd);
  *sc_out = sc;
  *sh_out = sh;
  if (!publish) return;
  a.scale[c] = sc;
  a.shift[c] = sh;
  a.mean[c] = (float)mean;
  a.invstd[c] = (float)invstd;
  if (a.running_mean) {
    // The conv bias is never added to the activations (train-mode BN subtracts it again); it only
    // shows up in the running mean.
    const double bias = a.conv_bias ? (double)a.conv_bias[c] : 0.0;
    const double m = (double)a.momentum;
    a.running_mean[c] = (float)((1.0 - m) * (double)a.running_mean[c] + m * (mean + bias));
    const double unbiased = a.count > 1.0 ? var * a.count / (a.count - 1.0) : var;
    a.running_var[c] = (float)((1.0 - m) * (double)a.running_var[c] + m * unbiased);
  }
}
// Eval mode: all layers of the network in ONE launch (blockIdx.y = layer) — 18 separate launches of this tiny kernel
// were a tenth of a batch-1 forward.
__global__ void bn_fold_eval_kernel(BnFoldBatch a) {
  const int l = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C[l]) return;
  const float invstd = 1.0f / sqrtf(a.rv[l][c] + a.eps);
  const float sc = a.gamma[l][c] * invstd;
  a.scale[l][c] = sc;
  a.shift[l][c] = a.beta[l][c] + ((a.conv_bias[l] ? a.conv_bias[l][c] : 0.f) - a.rm[l][c]) * sc;
}
cudaError_t launch_bn_fold_eval(const BnFoldBatch& a, cudaStream_t s) {
  if (a.layers < 1 || a.layers > BnFoldBatch::kMax) return cudaErrorInvalidValue;
  int maxC = 0;
  for (int l = 0; l < a.layers; ++l) maxC = a.C[l] > maxC ? a.C[l] : maxC;
  bn_fold_eval_kernel<<<dim3((maxC + 127) / 128, a.layers), 128, 0, s>>>(a);
  return launched();
}

// Statistics -> affine step alone: for the layers whose BatchNorm + ReLU is applied by the CONSUMING convolution
// (conv3_gemm_kernel's transform warps), so that no elementwise pass over the activation exists at all.
__global__ void bn_finalize_kernel(const BnFinalizeArgs fin) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  float sc, sh;
  if (c < fin.C) bn_channel_coef(fin, c, true, &sc, &sh);
  if (c == 0 && fin.num_batches_tracked) *fin.num_batches_tracked += 1;
}
cudaError_t launch_bn_finalize(const BnFinalizeArgs& fin, cudaStream_t s) {
  bn_finalize_kernel<<<(fin.C + 127) / 128, 128, 0, s>>>(fin);
  return launched();
}

// ============================================================================ BN apply + ReLU (+ 2x2 max-pool)
template <bool POOL>
__global__ void __launch_bounds__(256, 4) bn_relu_kernel(const bf16* __restrict__ y, int B, int H, int W, int C, const BnFinalizeArgs fin,
                               bf16* __restrict__ out, int out_pitch, int out_c0, bf16* __restrict__ pooled,
                               const HeadFwd head, const int rev) {
  // Fused statistics -> affine step (was a kernel of its own between the convolution and this pass): every block
  // derives (scale, shift) of all C channels into shared memory; block 0 also publishes them for the backward pass
  // and updates the running statistics.
  extern __shared__ __align__(16) float s_coef[];          // [2][C]
  const float* scale = s_coef;
  const float* shift = s_coef + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) bn_channel_coef(fin, c, blockIdx.x == 0, &s_coef[c], &s_coef[C + c]);
  if (blockIdx.x == 0 && threadIdx.x == 0 && fin.num_batches_tracked) *fin.num_batches_tracked += 1;
  __syncthreads();
  const int cg = C >> 3;
  const int cg_shift = 31 - __clz(cg);
  if (!POOL) {
    // The stride of the grid-stride loop is a multiple of the channel-group count (256 % cg == 0): a thread keeps its
    // channel group, so its coefficients live in registers; four 16-byte loads are in flight before the first use.
    // The tensor is walked from its END to its beginning: the convolution that has just written y finished with the
    // last tiles (still in the 126 MB L2), and the convolution that reads `out` next starts with the first ones, which
    // this pass therefore writes last.  CARTSEG_BN_REVERSE=0 (rev == 0) restores the ascending order for A/B runs.
    const long long total = (long long)B * H * W * cg;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int g = rev ? (int)((total - 1 - i0) % cg) : (int)(i0 % cg);
    float sc[8], sh[8], hw[8];
    *reinterpret_cast<float4*>(sc) = *reinterpret_cast<const float4*>(scale + g * 8);
    *reinterpret_cast<float4*>(sc + 4) = *reinterpret_cast<const float4*>(scale + g * 8 + 4);
    *reinterpret_cast<float4*>(sh) = *reinterpret_cast<const float4*>(shift + g * 8);
    *reinterpret_cast<float4*>(sh + 4) = *reinterpret_cast<const float4*>(shift + g * 8 + 4);
    if (head.logits) {
#pragma unroll
      for (int j = 0; j < 8; ++j) hw[j] = __ldg(head.w + g * 8 + j);
    }
    const float hb = (head.logits && head.b) ? __ldg(head.b) : 0.f;
    constexpr int U = 4;
    for (long long i = i0; i < total; i += U * stride) {
      Vec8 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long ik = rev ? total - 1 - (i + k * stride) : i + k * stride;
        if (i + k * stride < total) v[k] = ld8_nc(y + ik * 8);                 // dense rows: (i / cg) * C + g * 8 == i * 8
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (i + k * stride >= total) break;
        const long long ik = rev ? total - 1 - (i + k * stride) : i + k * stride;
        const long long p = ik >> cg_shift;                  // channel counts are powers of two
        float f[8];
        unpack8(v[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
        const Vec8 stored = pack8(f);
        if (!head.skip_store) st8(out + p * out_pitch + out_c0 + g * 8, stored);
        if (head.logits) {
          // fused 1x1 head (C == 64: the 8 lanes holding one pixel are adjacent): logits = <stored activation, w> + b.
          // total and the stride are multiples of 32 (H, W multiples of 16), so whole warps are in range together.
          float r[8], acc = 0.f;
          unpack8(stored, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc = fmaf(r[j], hw[j], acc);
          acc += __shfl_xor_sync(0xffffffffu, acc, 1);
          acc += __shfl_xor_sync(0xffffffffu, acc, 2);
          acc += __shfl_xor_sync(0xffffffffu, acc, 4);
          if (g == 0) head.logits[p] = acc + hb;
        }
      }
    }
  } else {
    const int H2 = H >> 1, W2 = W >> 1;
    const long long total = (long long)B * H2 * W2 * cg;
    for (long long i_ = blockIdx.x * (long long)blockDim.x + threadIdx.x; i_ < total;
         i_ += (long long)gridDim.x * blockDim.x) {
      const long long i = rev ? total - 1 - i_ : i_;
      const int g = (int)(i % cg);
      const long long q = i / cg;                      // pooled pixel index
      const int w2 = (int)(q % W2);
      const int h2 = (int)((q / W2) % H2);
      const int b = (int)(q / ((long long)W2 * H2));
      float sc[8], sh[8], mx[8];
      *reinterpret_cast<float4*>(sc) = *reinterpret_cast<const float4*>(scale + g * 8);
      *reinterpret_cast<float4*>(sc + 4) = *reinterpret_cast<const float4*>(scale + g * 8 + 4);
      *reinterpret_cast<float4*>(sh) = *reinterpret_cast<const float4*>(shift + g * 8);
      *reinterpret_cast<float4*>(sh + 4) = *reinterpret_cast<const float4*>(shift + g * 8 + 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) mx[j] = 0.f;          // post-ReLU values are >= 0
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const long long p = ((long long)b * H + 2 * h2 + (d >> 1)) * W + 2 * w2 + (d & 1);
        float f[8];
        unpack8(ld8(y + p * C + g * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
        const Vec8 v = pack8(f);
        st8(out + p * out_pitch + out_c0 + g * 8, v);
        float r[8];
        unpack8(v, r);                                  // pool the values as stored (bf16-rounded)
#pragma unroll
        for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], r[j]);
      }
      st8(pooled + q * C + g * 8, pack8(mx));
    }
  }
}
cudaError_t launch_bn_relu(const bf16* y, int B, int H, int W, int C, const BnFinalizeArgs& fin, bf16* out,
                           int out_pitch, int out_c0, bf16* pooled, const HeadFwd& head, cudaStream_t s) {
  const size_t smem = (size_t)2 * C * sizeof(float);
  static const int rev = [] { const char* e = getenv("CARTSEG_BN_REVERSE"); return (e && e[0] == '0') ? 0 : 1; }();
  if (C < 8 || (C & (C - 1))) return cudaErrorInvalidValue;          // power-of-two channel counts (shift instead of divide)
  if (head.logits && (pooled || C != 64 || ((long long)B * H * W) % 4 != 0)) return cudaErrorInvalidValue;
  if (pooled) {
    bn_relu_kernel<true><<<grid_for((long long)B * (H / 2) * (W / 2) * (C / 8), 256), 256, smem, s>>>(
        y, B, H, W, C, fin, out, out_pitch, out_c0, pooled, head, rev);
  } else {
    // four resident blocks per SM (launch bounds; 64 KB of loads in flight per SM), two full waves
    bn_relu_kernel<false><<<grid_for((long long)B * H * W * (C / 8), 256 * 4, 148 * 8), 256, smem, s>>>(y, B, H, W, C, fin, out,
                                                                                            out_pitch, out_c0, pooled, head, rev);
  }
  return launched();
}

__global__ void bn_apply_relu_kernel(const bf16* __restrict__ y, long long P, int C, const float* __restrict__ scale,
                                     const float* __restrict__ shift, bf16* __restrict__ out) {
  const int cg = C >> 3;
  const long long total = P * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    float f[8];
    unpack8(ld8(y + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], __ldg(scale + g * 8 + j), __ldg(shift + g * 8 + j)), 0.f);
    st8(out + i * 8, pack8(f));
  }
}
cudaError_t launch_bn_apply_relu(const bf16* y, long long P, int C, const float* scale, const float* shift, bf16* out,
                                 cudaStream_t s) {
  bn_apply_relu_kernel<<<grid_for(P * (C / 8), 256), 256, 0, s>>>(y, P, C, scale, shift, out);
  return launched();
}

// ============================================================================ BN + ReLU (+ pool/skip) backward
// g_total(pixel) = g[pixel] (+ g_pool[window] if the pixel is the first maximum of its 2x2 window).
// mask = (y*scale + shift > 0), evaluated on the bf16-rounded activation exactly as the forward stored it.
//
// Every thread owns ONE group of 8 channels for the whole kernel (its BN coefficients live in registers) and walks
// over pixels (or 2x2 windows) with 16-byte loads.  Two passes over (g, y): pass 1 reduces sum(gm) and sum(gm*xhat)
// per channel — per-block partials, combined in a fixed order, so the backward pass is run-to-run deterministic —
// pass 2 writes dy.
struct BnCoef { float sc[8], sh[8], mu[8], is[8], hw[8]; };   // hw: head weights (only when the gradient comes from the head)
CS_DEVINL void load8(const float* p, float* d) {
  *reinterpret_cast<float4*>(d) = __ldg(reinterpret_cast<const float4*>(p));
  *reinterpret_cast<float4*>(d + 4) = __ldg(reinterpret_cast<const float4*>(p + 4));
}
CS_DEVINL void load_coef(const BnBwdArgs& a, int g, BnCoef& k) {
  load8(a.scale + g * 8, k.sc);
  load8(a.shift + g * 8, k.sh);
  load8(a.mean + g * 8, k.mu);
  load8(a.invstd + g * 8, k.is);
  if (a.head_dlogits) load8(a.head_w + g * 8, k.hw);
}
// Non-pooled layers, split into a load phase and a compute phase so that a thread can have several units' loads in
// flight before the first use (the kernels run at two blocks per SM: memory parallelism has to come from the thread).
// HEAD: the layer feeds the 1x1 head — its activation gradient dlogits[p] * w[c] (rounded to bf16, as a stand-alone head
// backward would store it) is formed here instead of being written to and read back from HBM twice.  Such a unit holds
// 20 bytes of loads instead of 32, so the HEAD instantiations keep more units in flight (memory parallelism comes from
// the thread: ncu showed the U = 3 head variant at 2.8 TB/s where the two-tensor variant reaches 5.7 TB/s).
template <bool HEAD> struct PlainUnit { Vec8 y, g; };
template <> struct PlainUnit<true> { Vec8 y; float dl; };
template <bool HEAD>
CS_DEVINL void plain_load(const BnBwdArgs& a, int g, long long pix, PlainUnit<HEAD>& u) {
  u.y = ld8_nc(a.y + pix * a.C + g * 8);
  if constexpr (HEAD) u.dl = __ldg(a.head_dlogits + pix);
  else u.g = ld8_nc(a.g + pix * a.g_pitch + a.g_c0 + g * 8);
}
// gm[j] = g[j] where the stored (bf16-rounded) activation is positive, else 0;  xh[j] = (y - mean) * invstd
template <bool HEAD>
CS_DEVINL void plain_compute(const BnCoef& k, const PlainUnit<HEAD>& u, float gm[8], float xh[8], float* act_out = nullptr) {
  float yv[8], t[8], act[8];
  unpack8(u.y, yv);
  if constexpr (HEAD) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = u.dl * k.hw[j];
    unpack8(pack8(o), gm);
  } else {
    unpack8(u.g, gm);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = fmaxf(fmaf(yv[j], k.sc[j], k.sh[j]), 0.f);
  unpack8(pack8(t), act);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    xh[j] = (yv[j] - k.mu[j]) * k.is[j];
    if (!(act[j] > 0.f)) gm[j] = 0.f;
    if (act_out) act_out[j] = act[j];
  }
}

// Pooled layers, register-lean formulation: the four pixels of a 2x2 window stay packed (bf16) and are unpacked one
// channel at a time, so that a thread holds ~100 registers instead of ~170 and two blocks fit on an SM.  `f(d, j, gm,
// xhat)` is called for every pixel d of the window and channel j in the same (j outer, d inner) order for both
// passes; per-channel sums therefore see the pixels in the same order as the wide formulation (bit-identical).
CS_DEVINL float bf16_elem(const Vec8& v, int j) { return (j & 1) ? bf16_hi(v.w[j >> 1]) : bf16_lo(v.w[j >> 1]); }
template <class F>
CS_DEVINL void pooled_unit(const BnBwdArgs& a, const BnCoef& k, int g, long long unit, long long pix[4], F&& f) {
  const int H2 = a.H >> 1, W2 = a.W >> 1;
  const int w2 = (int)(unit % W2);
  const int h2 = (int)((unit / W2) % H2);
  const int b = (int)(unit / ((long long)W2 * H2));
  Vec8 yv8[4], gv8[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    pix[d] = ((long long)b * a.H + 2 * h2 + (d >> 1)) * a.W + 2 * w2 + (d & 1);
    yv8[d] = ld8(a.y + pix[d] * a.C + g * 8);
    gv8[d] = ld8(a.g + pix[d] * a.g_pitch + a.g_c0 + g * 8);
  }
  const Vec8 gp8 = ld8(a.g_pool + unit * a.C + g * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float yv[4], gm[4], act[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      yv[d] = bf16_elem(yv8[d], j);
      gm[d] = bf16_elem(gv8[d], j);
      act[d] = __bfloat162float(__float2bfloat16(fmaxf(fmaf(yv[d], k.sc[j], k.sh[j]), 0.f)));   // as stored by the forward
    }
    int best = 0;
    float bv = act[0];
#pragma unroll
    for (int d = 1; d < 4; ++d)
      if (act[d] > bv) { bv = act[d]; best = d; }
    const float gp = bf16_elem(gp8, j);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float gd = gm[d] + (d == best ? gp : 0.f);
      if (!(act[d] > 0.f)) gd = 0.f;
      f(d, j, gd, (yv[d] - k.mu[j]) * k.is[j]);
    }
  }
}

// Pooled layers, packed formulation (round 2).  The scalar version above costs ~17 instructions per element and the
// pooled kernels were paced by instruction issue, not by HBM (B200, K2: reduce 2.9 TB/s, apply 3.4 TB/s, together 1.5 ms
// of the backward critical path).  Here the two channels of a 32-bit word go through packed fp32 arithmetic
// (fma.rn.f32x2 for activation and xhat, cvt.rn.relu.bf16x2 for the stored activation) and only the window logic — first
// maximum of the bf16-rounded activations, exactly as the forward pooled them, and the ReLU mask — stays scalar: ~11
// instructions per element.  xhat = y * invstd + (-mean * invstd).
struct PoolCoef { uint64_t sc[4], sh[4], is[4], nm[4]; };
CS_DEVINL void load_pool_coef(const BnBwdArgs& a, int g, PoolCoef& k) {
  float sc[8], sh[8], mu[8], is[8];
  load8(a.scale + g * 8, sc); load8(a.shift + g * 8, sh); load8(a.mean + g * 8, mu); load8(a.invstd + g * 8, is);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    k.sc[j] = f32x2(sc[2 * j], sc[2 * j + 1]);
    k.sh[j] = f32x2(sh[2 * j], sh[2 * j + 1]);
    k.is[j] = f32x2(is[2 * j], is[2 * j + 1]);
    k.nm[j] = f32x2(-mu[2 * j] * is[2 * j], -mu[2 * j + 1] * is[2 * j + 1]);
  }
}
// One channel of a window: a[d] = stored activations, g[d] = incoming gradients; on return g[d] = total gradient w.r.t.
// the activation of pixel d (pooled gradient added at the FIRST maximum) with the ReLU mask applied.
CS_DEVINL void pool_route(const float a[4], float g[4], float gp) {
  const float bv = fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]));
  const bool p0 = a[0] == bv;
  const bool p1 = !p0 && a[1] == bv;
  const bool p2 = !p0 && !p1 && a[2] == bv;
  const bool p3 = !(p0 || p1 || p2);
  g[0] = a[0] > 0.f ? (p0 ? g[0] + gp : g[0]) : 0.f;
  g[1] = a[1] > 0.f ? (p1 ? g[1] + gp : g[1]) : 0.f;
  g[2] = a[2] > 0.f ? (p2 ? g[2] + gp : g[2]) : 0.f;
  g[3] = a[3] > 0.f ? (p3 ? g[3] + gp : g[3]) : 0.f;
}
// f(d, j, gd2, xh2): pixel d of the window, 32-bit word j (channels 2j, 2j+1), packed gradient and xhat.
template <class F>
CS_DEVINL void pooled_unit2(const BnBwdArgs& a, const PoolCoef& k, int g, long long unit, long long pix[4], F&& f) {
  // window -> pixels with ONE 32-bit division: b * H + 2 * h2 == 2 * (unit / W2) because H == 2 * H2 (the 64-bit
  // div / mod chain of the scalar version was a fifth of its instructions)
  const unsigned W2 = (unsigned)a.W >> 1;
  const unsigned row = (unsigned)unit / W2, w2 = (unsigned)unit - row * W2;
  const long long p0 = (long long)(2u * row) * a.W + 2u * w2;
  Vec8 yv8[4], gv8[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    pix[d] = p0 + (d >> 1) * a.W + (d & 1);
    yv8[d] = ld8_nc(a.y + pix[d] * a.C + g * 8);
    gv8[d] = ld8_nc(a.g + pix[d] * a.g_pitch + a.g_c0 + g * 8);
  }
  const Vec8 gp8 = ld8_nc(a.g_pool + unit * a.C + g * 8);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float alo[4], ahi[4], glo[4], ghi[4];
    uint64_t xh[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const uint64_t y2 = bf16x2_to_f32x2(yv8[d].w[j]);
      const uint32_t aw = relu_pack_bf16x2(fma_f32x2(y2, k.sc[j], k.sh[j]));   // the activation as the forward stored it
      alo[d] = bf16_lo(aw); ahi[d] = bf16_hi(aw);
      xh[d] = fma_f32x2(y2, k.is[j], k.nm[j]);
      glo[d] = bf16_lo(gv8[d].w[j]); ghi[d] = bf16_hi(gv8[d].w[j]);
    }
    pool_route(alo, glo, bf16_lo(gp8.w[j]));
    pool_route(ahi, ghi, bf16_hi(gp8.w[j]));
#pragma unroll
    for (int d = 0; d < 4; ++d) f(d, j, f32x2(glo[d], ghi[d]), xh[d]);
  }
}

static constexpr int kBnBwdThreads = 256;
static constexpr int kBnBwdMaxBlocks = 148 * 4;

// Partial sums: one row of 2*C floats per (block, row group).  Row groups per block: 8 (one per warp) when a warp
// holds whole pixel rows (C <= 256), else 256 / (C/8) pixel rows of several warps each.  Rows x 2C <= 4096 floats.
static int bn_bwd_rows_per_block(int C) { return C <= 256 ? 8 : kBnBwdThreads / (C / 8); }
size_t bn_bwd_scratch_bytes(int maxC) { return (size_t)kBnBwdMaxBlocks * 4 * maxC * sizeof(float); }

// Grid: two blocks per SM without pooling.  These kernels use no shared memory and <= 85 registers per thread so that
// two of their blocks fit on an SM NEXT TO a resident weight-gradient CTA (192 threads x 49 registers, ~210 KB of
// shared memory): the HBM-bound BN backward of layer L-1 then really runs under the tensor-bound wgrad of layer L
// (cs_unet_backward issues them on two streams).  With three blocks per SM the register file is full and the wgrad
// CTAs cannot be placed until the BN kernel has drained, which serialises the two.
// Measured alternatives (ncu totals of the 14 non-pooled launches of a k2 step, same box): three blocks per SM with 80
// registers and 2 units in flight (spills): apply 1.71 / reduce 1.27 ms vs 1.31 / 1.12 ms for this shape; 4 units in
// flight in the apply kernel (spills at 104 registers): 1.42 ms.
// CARTSEG_POOL_PACKED=0 selects the scalar BN-backward kernels (pooled and plain layers; same-box A/B runs).
static bool pool_packed() {
  static const bool v = [] { const char* e = getenv("CARTSEG_POOL_PACKED"); return !(e && e[0] == '0'); }();
  return v;
}
static int bn_bwd_grid(const BnBwdArgs& a) {
  const int rpb = kBnBwdThreads / (a.C / 8);
  const long long units = a.g_pool ? (long long)a.B * (a.H / 2) * (a.W / 2) : (long long)a.B * a.H * a.W;
  return grid_for(units, rpb * 2, 148 * 2);
}

template <bool POOL, bool HEAD, int U_, int REGS>
__global__ void __launch_bounds__(kBnBwdThreads) __maxnreg__(REGS) bn_bwd_reduce_kernel(BnBwdArgs a) {
  const int cg = a.C >> 3;
  const int g = threadIdx.x % cg, ri = threadIdx.x / cg, rpb = kBnBwdThreads / cg;
  const long long units = POOL ? (long long)a.B * (a.H >> 1) * (a.W >> 1) : (long long)a.B * a.H * a.W;
  BnCoef k;
  load_coef(a, g, k);
  float s1[8], s2[8], s3[8], sb = 0.f;      // s3 / sb: the 1x1 head's weight / bias gradient (HEAD only)
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = s3[j] = 0.f;
  const bool head_grads = HEAD && (a.head_grad_w != nullptr || a.head_grad_b != nullptr);
  if (POOL) {
    for (long long u = (long long)blockIdx.x * rpb + ri; u < units; u += (long long)gridDim.x * rpb) {
      long long pix[4];
      pooled_unit(a, k, g, u, pix, [&](int, int j, float gmv, float xhv) { s1[j] += gmv; s2[j] = fmaf(gmv, xhv, s2[j]); });
    }
  } else {
    constexpr int U = U_;
    const long long stride = (long long)gridDim.x * rpb;
    for (long long u0 = (long long)blockIdx.x * rpb + ri; u0 < units; u0 += U * stride) {
      PlainUnit<HEAD> pu[U];
#pragma unroll
      for (int i = 0; i < U; ++i)
        if (u0 + i * stride < units) plain_load<HEAD>(a, g, u0 + i * stride, pu[i]);
#pragma unroll
      for (int i = 0; i < U; ++i) {
        if (u0 + i * stride >= units) break;
        float gm[8], xh[8];
        if constexpr (HEAD) {
          float act[8];
          plain_compute<HEAD>(k, pu[i], gm, xh, act);
          if (head_grads) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s3[j] = fmaf(pu[i].dl, act[j], s3[j]);
            sb += pu[i].dl;
          }
        } else {
          plain_compute<HEAD>(k, pu[i], gm, xh);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] += gm[j]; s2[j] = fmaf(gm[j], xh[j], s2[j]); }
      }
    }
  }
  // Combine the pixel rows that share a warp with a fixed shuffle pattern (deterministic), then one partial row per
  // row group goes straight to global memory: no shared memory, see bn_bwd_grid.
  int rows_per_block, my_row;
  bool writer;
  if (cg <= 32) {
#pragma unroll
    for (int o = 16; o >= 8; o >>= 1) {
      if (o >= cg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
          s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
      }
    }
    if (head_grads) {
#pragma unroll
      for (int o = 16; o >= 8; o >>= 1) {
        if (o >= cg) {
#pragma unroll
          for (int j = 0; j < 8; ++j) s3[j] += __shfl_xor_sync(0xffffffffu, s3[j], o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
      }
    }
    rows_per_block = kBnBwdThreads / 32;
    my_row = threadIdx.x >> 5;
    writer = (threadIdx.x & 31) < cg;
  } else {
    rows_per_block = rpb;
    my_row = ri;
    writer = true;
  }
  if (writer) {
    // channel-major layout partial[2C][rows]: the finalize kernel reads one channel's partials as a contiguous run
    const size_t rows_total = (size_t)gridDim.x * rows_per_block;
    const size_t r = (size_t)blockIdx.x * rows_per_block + my_row;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a.partial[(size_t)(g * 8 + j) * rows_total + r] = s1[j];
      a.partial[(size_t)(a.C + g * 8 + j) * rows_total + r] = s2[j];
      if (head_grads) a.partial[(size_t)(2 * a.C + g * 8 + j) * rows_total + r] = s3[j];
    }
    if (head_grads && g == 0) a.partial[(size_t)(3 * a.C) * rows_total + r] = sb;
  }
}

template <int REGS>
__global__ void __launch_bounds__(kBnBwdThreads) __maxnreg__(REGS) bn_bwd_pool_reduce_kernel(BnBwdArgs a) {
  const int cg = a.C >> 3;
  const int g = threadIdx.x % cg, ri = threadIdx.x / cg, rpb = kBnBwdThreads / cg;
  const long long units = (long long)a.B * (a.H >> 1) * (a.W >> 1);
  PoolCoef k;
  load_pool_coef(a, g, k);
  uint64_t s1p[4], s2p[4];
  const uint64_t one2 = f32x2(1.f, 1.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) s1p[j] = s2p[j] = 0ull;
  // a.reverse: from the last window to the first — the dgrad that produced g wrote its last tiles last (still in L2)
  for (long long u_ = (long long)blockIdx.x * rpb + ri; u_ < units; u_ += (long long)gridDim.x * rpb) {
    const long long u = a.reverse ? units - 1 - u_ : u_;
    long long pix[4];
    pooled_unit2(a, k, g, u, pix, [&](int, int j, uint64_t gd2, uint64_t xh2) {
      s1p[j] = fma_f32x2(gd2, one2, s1p[j]);
      s2p[j] = fma_f32x2(gd2, xh2, s2p[j]);
    });
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { unpack_f32x2(s1p[j], s1[2 * j], s1[2 * j + 1]); unpack_f32x2(s2p[j], s2[2 * j], s2[2 * j + 1]); }
  // same partial layout as bn_bwd_reduce_kernel (fixed shuffle pattern, one row per row group, channel-major)
  int rows_per_block, my_row;
  bool writer;
  if (cg <= 32) {
#pragma unroll
    for (int o = 16; o >= 8; o >>= 1) {
      if (o >= cg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
          s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
      }
    }
    rows_per_block = kBnBwdThreads / 32;
    my_row = threadIdx.x >> 5;
    writer = (threadIdx.x & 31) < cg;
  } else {
    rows_per_block = rpb;
    my_row = ri;
    writer = true;
  }
  if (writer) {
    const size_t rows_total = (size_t)gridDim.x * rows_per_block;
    const size_t r = (size_t)blockIdx.x * rows_per_block + my_row;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a.partial[(size_t)(g * 8 + j) * rows_total + r] = s1[j];
      a.partial[(size_t)(a.C + g * 8 + j) * rows_total + r] = s2[j];
    }
  }
}

template <int REGS>
__global__ void __launch_bounds__(kBnBwdThreads) __maxnreg__(REGS) bn_bwd_pool_apply_kernel(BnBwdArgs a) {
  const int cg = a.C >> 3;
  const int g = threadIdx.x % cg, ri = threadIdx.x / cg, rpb = kBnBwdThreads / cg;
  const long long units = (long long)a.B * (a.H >> 1) * (a.W >> 1);
  PoolCoef k;
  load_pool_coef(a, g, k);
  uint64_t nc1[4], nc2[4];
  {
    float c1[8], c2[8];
    load8(a.c1 + g * 8, c1);
    load8(a.c2 + g * 8, c2);
#pragma unroll
    for (int j = 0; j < 4; ++j) { nc1[j] = f32x2(-c1[2 * j], -c1[2 * j + 1]); nc2[j] = f32x2(-c2[2 * j], -c2[2 * j + 1]); }
  }
  const uint64_t one2 = f32x2(1.f, 1.f), zero2 = 0ull;
  for (long long u = (long long)blockIdx.x * rpb + ri; u < units; u += (long long)gridDim.x * rpb) {
    long long pix[4];
    Vec8 out[4];
    pooled_unit2(a, k, g, u, pix, [&](int d, int j, uint64_t gd2, uint64_t xh2) {
      // dy = scale * (gd - c1 - xhat * c2)
      uint64_t t = fma_f32x2(xh2, nc2[j], gd2);
      t = fma_f32x2(t, one2, nc1[j]);
      t = fma_f32x2(t, k.sc[j], zero2);
      float lo, hi;
      unpack_f32x2(t, lo, hi);
      out[d].w[j] = pack_bf16x2(lo, hi);
    });
#pragma unroll
    for (int d = 0; d < 4; ++d) st8(a.dy + pix[d] * a.C + g * 8, out[d]);
  }
}

// Plain (no pooling, no head) layers, packed formulation: the ReLU mask is taken from the fp32 pre-activation
// v = fma(y, scale, shift) — the stored activation bf16(max(v, 0)) is positive exactly when v > 2^-134 (round to nearest
// even: 2^-134 is the midpoint between 0 and the smallest bf16 subnormal 2^-133) — so the bf16 round trip of the
// scalar version disappears; activation and xhat are one packed FMA per channel pair.  ~6 instructions per element
// instead of ~12: these kernels share their SMs with a weight-gradient GEMM and are partly issue-bound.
static constexpr float kBf16ReluThreshold = 4.591774807899561e-41f;   // 2^-134
// HEAD (the layer feeding the 1x1 head): the incoming gradient is bf16(dlogits[p] * w[c]) formed on the fly, and the
// head's parameter gradients (sum dlogits * act, sum dlogits) ride along in the reduction — there the bf16-rounded
// activation itself is needed, so it is produced with cvt.rn.relu.bf16x2 and the mask taken from it.
template <bool HEAD> struct PlainUnit2 { Vec8 y, g; };
template <> struct PlainUnit2<true> { Vec8 y; float dl; };
template <bool HEAD>
CS_DEVINL void plain_load2(const BnBwdArgs& a, int g, long long pix, PlainUnit2<HEAD>& u) {
  u.y = ld8_nc(a.y + pix * a.C + g * 8);
  if constexpr (HEAD) u.dl = __ldg(a.head_dlogits + pix);
  else u.g = ld8_nc(a.g + pix * a.g_pitch + a.g_c0 + g * 8);
}
// gm2 / xh2: masked gradient and xhat per channel pair; act2 (HEAD only): the stored activation as fp32 pairs
template <bool HEAD>
CS_DEVINL void plain_compute2(const PoolCoef& k, const uint64_t* hw2, const PlainUnit2<HEAD>& u, uint64_t gm2[4], uint64_t xh2[4],
                              uint64_t* act2) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t y2 = bf16x2_to_f32x2(u.y.w[j]);
    const uint64_t v2 = fma_f32x2(y2, k.sc[j], k.sh[j]);
    xh2[j] = fma_f32x2(y2, k.is[j], k.nm[j]);
    if constexpr (HEAD) {
      const uint32_t aw = relu_pack_bf16x2(v2);
      const float alo = bf16_lo(aw), ahi = bf16_hi(aw);
      float olo, ohi;
      unpack_f32x2(fma_f32x2(f32x2(u.dl, u.dl), hw2[j], 0ull), olo, ohi);
      const uint32_t gw = pack_bf16x2(olo, ohi);             // as a stand-alone head backward would store it
      gm2[j] = f32x2(alo > 0.f ? bf16_lo(gw) : 0.f, ahi > 0.f ? bf16_hi(gw) : 0.f);
      act2[j] = f32x2(alo, ahi);
    } else {
      float vlo, vhi;
      unpack_f32x2(v2, vlo, vhi);
      gm2[j] = f32x2(vlo > kBf16ReluThreshold ? bf16_lo(u.g.w[j]) : 0.f, vhi > kBf16ReluThreshold ? bf16_hi(u.g.w[j]) : 0.f);
    }
  }
}
CS_DEVINL void load_head_w2(const BnBwdArgs& a, int g, uint64_t hw2[4]) {
  float hw[8];
  load8(a.head_w + g * 8, hw);
#pragma unroll
  for (int j = 0; j < 4; ++j) hw2[j] = f32x2(hw[2 * j], hw[2 * j + 1]);
}
template <bool HEAD, int U, int REGS>
__global__ void __launch_bounds__(kBnBwdThreads) __maxnreg__(REGS) bn_bwd_plain_reduce_kernel(BnBwdArgs a) {
  const int cg = a.C >> 3;
  const int g = threadIdx.x % cg, ri = threadIdx.x / cg, rpb = kBnBwdThreads / cg;
  const long long units = (long long)a.B * a.H * a.W;
  PoolCoef k;
  load_pool_coef(a, g, k);
  uint64_t hw2[4] = {0ull, 0ull, 0ull, 0ull};
  if constexpr (HEAD) load_head_w2(a, g, hw2);
  uint64_t s1p[4], s2p[4], s3p[4];
  float sb = 0.f;
  const uint64_t one2 = f32x2(1.f, 1.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) s1p[j] = s2p[j] = s3p[j] = 0ull;
  const bool head_grads = HEAD && (a.head_grad_w != nullptr || a.head_grad_b != nullptr);
  const long long stride = (long long)gridDim.x * rpb;
  for (long long u0 = (long long)blockIdx.x * rpb + ri; u0 < units; u0 += U * stride) {
    PlainUnit2<HEAD> pu[U];
#pragma unroll
    for (int i = 0; i < U; ++i)   // a.reverse: from the last pixel to the first (the tail of g is the freshest in L2)
      if (u0 + i * stride < units) plain_load2<HEAD>(a, g, a.reverse ? units - 1 - (u0 + i * stride) : u0 + i * stride, pu[i]);
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (u0 + i * stride >= units) break;
      uint64_t gm2[4], xh2[4], act2[4];
      plain_compute2<HEAD>(k, hw2, pu[i], gm2, xh2, act2);
#pragma unroll
      for (int j = 0; j < 4; ++j) { s1p[j] = fma_f32x2(gm2[j], one2, s1p[j]); s2p[j] = fma_f32x2(gm2[j], xh2[j], s2p[j]); }
      if constexpr (HEAD) {
        if (head_grads) {
          const uint64_t dl2 = f32x2(pu[i].dl, pu[i].dl);
#pragma unroll
          for (int j = 0; j < 4; ++j) s3p[j] = fma_f32x2(dl2, act2[j], s3p[j]);
          sb += pu[i].dl;
        }
      }
    }
  }
  float s1[8], s2[8], s3[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    unpack_f32x2(s1p[j], s1[2 * j], s1[2 * j + 1]);
    unpack_f32x2(s2p[j], s2[2 * j], s2[2 * j + 1]);
    unpack_f32x2(s3p[j], s3[2 * j], s3[2 * j + 1]);
  }
  // same partial layout as bn_bwd_reduce_kernel (fixed shuffle pattern, one row per row group, channel-major)
  int rows_per_block, my_row;
  bool writer;
  if (cg <= 32) {
#pragma unroll
    for (int o = 16; o >= 8; o >>= 1) {
      if (o >= cg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
          s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
        if (head_grads) {
#pragma unroll
          for (int j = 0; j < 8; ++j) s3[j] += __shfl_xor_sync(0xffffffffu, s3[j], o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
      }
    }
    rows_per_block = kBnBwdThreads / 32;
    my_row = threadIdx.x >> 5;
    writer = (threadIdx.x & 31) < cg;
  } else {
    rows_per_block = rpb;
    my_row = ri;
    writer = true;
  }
  if (writer) {
    const size_t rows_total = (size_t)gridDim.x * rows_per_block;
    const size_t r = (size_t)blockIdx.x * rows_per_block + my_row;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a.partial[(size_t)(g * 8 + j) * rows_total + r] = s1[j];
      a.partial[(size_t)(a.C + g * 8 + j) * rows_total + r] = s2[j];
      if (head_grads) a.partial[(size_t)(2 * a.C + g * 8 + j) * rows_total + r] = s3[j];
    }
    if (head_grads && g == 0) a.partial[(size_t)(3 * a.C) * rows_total + r] = sb;
  }
}
template <bool HEAD, int U, int REGS>
__global__ void __launch_bounds__(kBnBwdThreads) __maxnreg__(REGS) bn_bwd_plain_apply_kernel(BnBwdArgs a) {
  const int cg = a.C >> 3;
  const int g = threadIdx.x % cg, ri = threadIdx.x / cg, rpb = kBnBwdThreads / cg;
  const long long units = (long long)a.B * a.H * a.W;
  PoolCoef k;
  load_pool_coef(a, g, k);
  uint64_t hw2[4] = {0ull, 0ull, 0ull, 0ull};
  if constexpr (HEAD) load_head_w2(a, g, hw2);
  uint64_t nc1[4], nc2[4];
  {
    float c1[8], c2[8];
    load8(a.c1 + g * 8, c1);
    load8(a.c2 + g * 8, c2);
#pragma unroll
    for (int j = 0; j < 4; ++j) { nc1[j] = f32x2(-c1[2 * j], -c1[2 * j + 1]); nc2[j] = f32x2(-c2[2 * j], -c2[2 * j + 1]); }
  }
  const uint64_t one2 = f32x2(1.f, 1.f), zero2 = 0ull;
  const long long stride = (long long)gridDim.x * rpb;
  for (long long u0 = (long long)blockIdx.x * rpb + ri; u0 < units; u0 += U * stride) {
    PlainUnit2<HEAD> pu[U];
#pragma unroll
    for (int i = 0; i < U; ++i)
      if (u0 + i * stride < units) plain_load2<HEAD>(a, g, u0 + i * stride, pu[i]);
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const long long u = u0 + i * stride;
      if (u >= units) break;
      uint64_t gm2[4], xh2[4], act2[4];
      plain_compute2<HEAD>(k, hw2, pu[i], gm2, xh2, act2);
      Vec8 o;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint64_t t = fma_f32x2(xh2[j], nc2[j], gm2[j]);          // dy = scale * (gm - c1 - xhat * c2)
        t = fma_f32x2(t, one2, nc1[j]);
        t = fma_f32x2(t, k.sc[j], zero2);
        float lo, hi;
        unpack_f32x2(t, lo, hi);
        o.w[j] = pack_bf16x2(lo, hi);
      }
      st8(a.dy + u * a.C + g * 8, o);
    }
  }
}

// s1/s2 totals -> per-channel means c1, c2 and the BN parameter gradients.  One block per channel: thread t sums the
// partials of rows t, t+256, ... (contiguous in the channel-major layout) in fp64; warps and then the 8 warp sums are
// combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(BnBwdArgs a, int rows) {
  __shared__ double st[4][8];
  const int c = blockIdx.x;
  const float* p1 = a.partial + (size_t)c * rows;
  const float* p2 = a.partial + (size_t)(a.C + c) * rows;
  // the 1x1 head's parameter gradients ride along (HEAD reduce only): channel c's weight gradient, and block 0 the bias
  const bool head = a.head_dlogits != nullptr && (a.head_grad_w != nullptr || a.head_grad_b != nullptr);
  const float* p3 = a.partial + (size_t)(2 * a.C + c) * rows;
  const float* p4 = a.partial + (size_t)(3 * a.C) * rows;
  double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0;
  for (int r = threadIdx.x; r < rows; r += 256) {
    t1 += (double)__ldg(p1 + r);
    t2 += (double)__ldg(p2 + r);
    if (head) {
      t3 += (double)__ldg(p3 + r);
      if (c == 0) t4 += (double)__ldg(p4 + r);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    t3 += __shfl_xor_sync(0xffffffffu, t3, o);
    t4 += __shfl_xor_sync(0xffffffffu, t4, o);
  }
  if ((threadIdx.x & 31) == 0) {
    st[0][threadIdx.x >> 5] = t1; st[1][threadIdx.x >> 5] = t2; st[2][threadIdx.x >> 5] = t3; st[3][threadIdx.x >> 5] = t4;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  t1 = t2 = t3 = t4 = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { t1 += st[0][k]; t2 += st[1][k]; t3 += st[2][k]; t4 += st[3][k]; }
  if (head) {
    if (a.head_grad_w) a.head_grad_w[c] = (float)t3;
    if (c == 0 && a.head_grad_b) *a.head_grad_b = (float)t4;
  }
  const double inv_n = 1.0 / ((double)a.B * a.H * a.W);
  a.c1[c] = (float)(t1 * inv_n);
  a.c2[c] = (float)(t2 * inv_n);
  if (a.grad_gamma) a.grad_gamma[c] = (float)t2;
  if (a.grad_beta) a.grad_beta[c] = (float)t1;
  if (a.grad_conv_bias) a.grad_conv_bias[c] = 0.f;      // exactly zero: BN removes the bias again
}

cudaError_t launch_bn_bwd_reduce(const BnBwdArgs& a_in, cudaStream_t s) {
  static const int rev = [] { const char* e = getenv("CARTSEG_BN_REVERSE_BWD"); return (e && e[0] == '0') ? 0 : 1; }();
  BnBwdArgs a = a_in;
  a.reverse = rev;
  if ((long long)a.B * a.H * a.W >= (1ll << 31)) return cudaErrorInvalidValue;   // the pooled kernels index windows in 32 bits
  const int cg = a.C / 8;
  if (cg > kBnBwdThreads || kBnBwdThreads % cg) return cudaErrorInvalidValue;
  const int grid = bn_bwd_grid(a);
  if (a.g_pool) {
    if (a.head_dlogits) return cudaErrorInvalidValue;
    if (pool_packed()) bn_bwd_pool_reduce_kernel<128><<<grid, kBnBwdThreads, 0, s>>>(a);
    else bn_bwd_reduce_kernel<true, false, 1, 128><<<grid, kBnBwdThreads, 0, s>>>(a);
  } else if (a.head_dlogits) {
    if (pool_packed()) bn_bwd_plain_reduce_kernel<true, 5, 128><<<grid, kBnBwdThreads, 0, s>>>(a);
    else bn_bwd_reduce_kernel<false, true, 5, 128><<<grid, kBnBwdThreads, 0, s>>>(a);
  } else if (pool_packed()) {
    bn_bwd_plain_reduce_kernel<false, 4, 104><<<grid, kBnBwdThreads, 0, s>>>(a);
  } else {
    bn_bwd_reduce_kernel<false, false, 4, 104><<<grid, kBnBwdThreads, 0, s>>>(a);
  }
  cudaError_t e = launched();
  if (e != cudaSuccess) return e;
  bn_bwd_finalize_kernel<<<a.C, 256, 0, s>>>(a, grid * bn_bwd_rows_per_block(a.C));
  return launched();
}

template <bool POOL, bool HEAD, int U_, int REGS>
__global__ void __launch_bounds__(kBnBwdThreads) __maxnreg__(REGS) bn_bwd_apply_kernel(BnBwdArgs a) {
  const int cg = a.C >> 3;
  const int g = threadIdx.x % cg, ri = threadIdx.x / cg, rpb = kBnBwdThreads / cg;
  const long long units = POOL ? (long long)a.B * (a.H >> 1) * (a.W >> 1) : (long long)a.B * a.H * a.W;
  BnCoef k;
  load_coef(a, g, k);
  float c1[8], c2[8];
  load8(a.c1 + g * 8, c1);
  load8(a.c2 + g * 8, c2);
  if (POOL) {
    for (long long u = (long long)blockIdx.x * rpb + ri; u < units; u += (long long)gridDim.x * rpb) {
      long long pix[4];
      float prev[4];                                     // the even channel of the pair being packed
      Vec8 out[4];
      pooled_unit(a, k, g, u, pix, [&](int d, int j, float gmv, float xhv) {
        const float o = k.sc[j] * (gmv - c1[j] - xhv * c2[j]);
        if (j & 1) out[d].w[j >> 1] = pack_bf16x2(prev[d], o);
        else prev[d] = o;
      });
#pragma unroll
      for (int d = 0; d < 4; ++d) st8(a.dy + pix[d] * a.C + g * 8, out[d]);
    }
  } else {
    constexpr int U = U_;
    const long long stride = (long long)gridDim.x * rpb;
    for (long long u0 = (long long)blockIdx.x * rpb + ri; u0 < units; u0 += U * stride) {
      PlainUnit<HEAD> pu[U];
#pragma unroll
      for (int i = 0; i < U; ++i)
        if (u0 + i * stride < units) plain_load<HEAD>(a, g, u0 + i * stride, pu[i]);
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const long long u = u0 + i * stride;
        if (u >= units) break;
        float gm[8], xh[8], o[8];
        plain_compute<HEAD>(k, pu[i], gm, xh);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = k.sc[j] * (gm[j] - c1[j] - xh[j] * c2[j]);
        st8(a.dy + u * a.C + g * 8, pack8(o));
      }
    }
  }
}
cudaError_t launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t s) {
  const int grid = bn_bwd_grid(a);
  if (a.g_pool && pool_packed()) bn_bwd_pool_apply_kernel<128><<<grid, kBnBwdThreads, 0, s>>>(a);
  else if (a.g_pool) bn_bwd_apply_kernel<true, false, 1, 128><<<grid, kBnBwdThreads, 0, s>>>(a);
  else if (a.head_dlogits && pool_packed()) bn_bwd_plain_apply_kernel<true, 5, 128><<<grid, kBnBwdThreads, 0, s>>>(a);
  else if (a.head_dlogits) bn_bwd_apply_kernel<false, true, 5, 128><<<grid, kBnBwdThreads, 0, s>>>(a);
  else if (pool_packed()) bn_bwd_plain_apply_kernel<false, 3, 104><<<grid, kBnBwdThreads, 0, s>>>(a);
  else bn_bwd_apply_kernel<false, false, 3, 104><<<grid, kBnBwdThreads, 0, s>>>(a);
  return launched();
}

// ============================================================================ 1x1 head (C -> 1)
__global__ void head_fwd_kernel(const bf16* __restrict__ act, long long P, int C, const float* __restrict__ w,
                                const float* __restrict__ b, float* __restrict__ logits) {
  // 8 lanes per pixel, each lane strides over the channel groups
  const int cg = C >> 3;
  const int sub = threadIdx.x & 7;
  const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
  const float bias = b ? __ldg(b) : 0.f;
  for (long long p = blockIdx.x * (long long)(blockDim.x >> 3) + (threadIdx.x >> 3); p < P; p += stride) {
    float acc = 0.f;
    for (int g = sub; g < cg; g += 8) {
      float f[8];
      unpack8(ld8(act + p * C + g * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(f[j], __ldg(w + g * 8 + j), acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0) logits[p] = acc + bias;
  }
}
cudaError_t launch_head_fwd(const bf16* act, long long P, int C, const float* w, const float* b, float* logits,
                            cudaStream_t s) {
  // P is always a multiple of 32 here (H, W multiples of 16), so every warp iterates uniformly
  head_fwd_kernel<<<grid_for(P * 8, 256), 256, 0, s>>>(act, P, C, w, b, logits);
  return launched();
}

__global__ void __launch_bounds__(256) head_bwd_kernel(const bf16* __restrict__ act, const float* __restrict__ dlogits,
                                                      long long P, int C, const float* __restrict__ w,
                                                      bf16* __restrict__ g_act, float* __restrict__ grad_w,
                                                      float* __restrict__ grad_b) {
  extern __shared__ float sacc[];                       // [C + 1]
  for (int i = threadIdx.x; i <= C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cg = C >> 3;
  const int g = threadIdx.x % cg, ri = threadIdx.x / cg, rpb = blockDim.x / cg;
  float wv[8], acc[8], accb = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { wv[j] = __ldg(w + g * 8 + j); acc[j] = 0.f; }
  if (ri < rpb) {
    for (long long p = (long long)blockIdx.x * rpb + ri; p < P; p += (long long)gridDim.x * rpb) {
      const float d = __ldg(dlogits + p);
      float f[8], o[8];
      unpack8(ld8(act + p * C + g * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] = fmaf(d, f[j], acc[j]); o[j] = d * wv[j]; }
      if (g_act) st8(g_act + p * C + g * 8, pack8(o));
      if (g == 0) accb += d;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sacc[g * 8 + j], acc[j]);
    if (g == 0) atomicAdd(&sacc[C], accb);
  }
  __syncthreads();
  if (grad_w)
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&grad_w[i], sacc[i]);
  if (threadIdx.x == 0 && grad_b) atomicAdd(grad_b, sacc[C]);
}
// g_act only (test hook: materialises the activation gradient the fused BN backward forms on the fly)
__global__ void head_grad_act_kernel(const float* __restrict__ dlogits, long long P, int C, const float* __restrict__ w,
                                     bf16* __restrict__ g_act) {
  const int cg = C >> 3;
  const long long total = P * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long p = i / cg;
    const float d = __ldg(dlogits + p);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = d * __ldg(w + g * 8 + j);
    st8(g_act + p * C + g * 8, pack8(o));
  }
}
cudaError_t launch_head_grad_act(const float* dlogits, long long P, int C, const float* w, bf16* g_act, cudaStream_t s) {
  head_grad_act_kernel<<<grid_for(P * (C / 8), 256), 256, 0, s>>>(dlogits, P, C, w, g_act);
  return launched();
}

cudaError_t launch_head_bwd(const bf16* act, const float* dlogits, long long P, int C, const float* w, bf16* g_act,
                            float* grad_w, float* grad_b, cudaStream_t s) {
  const int rpb = 256 / (C / 8);
  if (!grad_w && !grad_b && !g_act) return cudaSuccess;          // frozen head, activation gradient formed elsewhere
  cudaError_t e = cudaSuccess;
  if (grad_w) { e = cudaMemsetAsync(grad_w, 0, C * sizeof(float), s); if (e != cudaSuccess) return e; }
  if (grad_b) { e = cudaMemsetAsync(grad_b, 0, sizeof(float), s); if (e != cudaSuccess) return e; }
  head_bwd_kernel<<<grid_for(P, rpb * 8, 148 * 4), 256, (C + 1) * sizeof(float), s>>>(act, dlogits, P, C, w, g_act,
                                                                                     grad_w, grad_b);
  return launched();
}

// ============================================================================ per-channel sum (convT bias grad)
__global__ void __launch_bounds__(256) channel_sum_kernel(const bf16* __restrict__ gsrc, int pitch, int c0, long long P,
                                                         int C, float* __restrict__ out) {
  extern __shared__ float sacc[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cg = C >> 3;
  const int g = threadIdx.x % cg, ri = threadIdx.x / cg, rpb = blockDim.x / cg;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ri < rpb) {
    for (long long p = (long long)blockIdx.x * rpb + ri; p < P; p += (long long)gridDim.x * rpb) {
      float f[8];
      unpack8(ld8(gsrc + p * pitch + c0 + g * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sacc[g * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&out[i], sacc[i]);
}
cudaError_t launch_channel_sum(const bf16* g, int pitch, int c0, long long P, int C, float* out, cudaStream_t s) {
  if (C / 8 > 256) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(out, 0, C * sizeof(float), s);
  if (e != cudaSuccess) return e;
  const int rpb = 256 / (C / 8);
  channel_sum_kernel<<<grid_for(P, rpb * 8, 148 * 4), 256, C * sizeof(float), s>>>(g, pitch, c0, P, C, out);
  return launched();
}

// Conv-transpose bias gradient from the column sums the dgrad epilogue of dconvL.0 left behind (fp64, the statistics
// machinery of the pixel GEMMs): out[c] = sum[c]; all `n` entries of sum / sq are zeroed again for the next step.
__global__ void stat_to_bias_kernel(double* __restrict__ sum, double* __restrict__ sq, int n, int C, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i < C) out[i] = (float)sum[i];
  sum[i] = 0.0;
  sq[i] = 0.0;
}
cudaError_t launch_stat_to_bias(double* sum, double* sq, int n, int C, float* out, cudaStream_t s) {
  stat_to_bias_kernel<<<(n + 255) / 256, 256, 0, s>>>(sum, sq, n, C, out);
  return launched();
}

// ============================================================================ debug read-back
// NHWC bf16 view (pitch, c0) -> dense fp32 NCHW (tests compare internal tensors stage by stage).
__global__ void nhwc_to_nchw_f32_kernel(const bf16* __restrict__ src, int pitch, int c0, int B, int H, int W, int C,
                                        float* __restrict__ dst) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % C);
    const int b = (int)(i / ((long long)W * H * C));
    dst[i] = __bfloat162float(src[(((long long)b * H + h) * W + w) * pitch + c0 + c]);
  }
}
cudaError_t launch_nhwc_to_nchw_f32(const bf16* src, int pitch, int c0, int B, int H, int W, int C, float* dst,
                                    cudaStream_t s) {
  nhwc_to_nchw_f32_kernel<<<grid_for((long long)B * C * H * W, 256), 256, 0, s>>>(src, pitch, c0, B, H, W, C, dst);
  return launched();
}

}  // namespace cs
