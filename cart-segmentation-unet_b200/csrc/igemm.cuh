// Parameter blocks and host-side launch declarations for the two tcgen05 implicit-GEMM kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch_count.cuh"

namespace cs {

// ---------------------------------------------------------------------------------------------
// "Pixel-major" implicit GEMM:  D[128 pixels, BLOCK_N] += A[pixels, K] * B[K, N]
//   A : NHWC bf16 activations, read as 8(w) x 16(h) pixel patches through 4-D TMA maps
//       (channels, W, H, batch) — out-of-image coordinates are zero-filled by TMA, which is the
//       convolution padding.  One A load per (64-channel chunk, group g) brings a patch that is
//       (R-1) rows taller than the tile; the R vertical taps of the group are 1024-byte-aligned
//       sub-views of it (one 8-pixel image row == one 128B-swizzle atom).
//   B : packed weights [G*R taps][Ntot][K] bf16, K-major, one 2-D TMA map.
//   Used for: conv3x3 fprop and dgrad (G=3 horizontal shifts, R=3), ConvTranspose2d(2,2) fprop
//   (G=1,R=1, N = 4*Cout scattered through 4 strided output maps), its dgrad (G=4 strided input
//   maps, R=1) and the im2col'd first conv (G=1,R=1).
struct PixGemmParams {
  CUtensorMap tmapA[4];
  CUtensorMap tmapB;
  CUtensorMap tmapO[4];
  int G, R;
  int pair;             // 1: CTA-pair kernel (cta_group::2, M = 256); tmapB's box then holds BLOCK_N/2 rows
  int a_map[4], a_dw[4], a_dh[4];
  int a_chan0;          // first input channel (coordinate offset inside the A maps)
  int kchunks;          // K / 64 per tap
  int Ntot;             // rows per tap in B
  int n_blocks;         // Ntot / BLOCK_N
  int tiles_w, tiles_h, batch;
  int H, W;             // extent of the tile grid's image (pixels past it are excluded from statistics)
  int cols_per_map;     // output channels written through one output map (pair kernel: an n-block may span maps)
  int o_blocks_per_map; // n-blocks written through one output map
  int o_chan0;          // first output channel (coordinate offset inside the O maps)
  const float* scale;   // per output channel (index inside its map), may be null
  const float* shift;   // per output channel, may be null
  int relu;
  double* stat_sum;     // per output channel sum / sum of squares of the stored bf16 values
  double* stat_sq;      //   (both null when no statistics are wanted)
  // conv3_gemm_kernel (3x3 convolutions, G = R = 3): ONE (8+2) x (16+2) pixel patch per 64-channel chunk serves all nine
  // taps (the three horizontal taps are 128-byte row offsets into it) instead of three 8 x 18 patches.
  int conv3;            // 1: use conv3_gemm_kernel with tmapA3
  CUtensorMap tmapA3;   // box (64 ch, 10, 18, 1)
  // stem_gemm_kernel (first convolution, Cin <= 7): the im2col rows [pixel][k = (kh*3+kw)*Cin + c] are built in shared
  // memory from the fp32 NCHW image by two builder warps — no im2col matrix in HBM on the forward path.
  const float* stem_x;  // fp32 NCHW input image (non-null selects stem_gemm_kernel)
  int stem_cin;
  CUtensorMap tmapX;    // fp32 (W, H, Cin, B) map over stem_x, box (16, 18, Cin, 1): make_stem_tmap
  // conv3_gemm_kernel only: the A operand is the RAW output of the previous convolution; two transform warps apply that
  // layer's BatchNorm + ReLU (relu(y * in_scale[c] + in_shift[c]), rounded to bf16 exactly as bn_relu_kernel stores it) to
  // every patch in shared memory between the TMA load and the MMAs, and zero the pixels outside the image (the padding is
  // zero AFTER the activation).  Indexed by input channel relative to a_chan0; null = plain operand.
  const float* in_scale;
  const float* in_shift;
  // Eval-mode last layer (Cout == 64, one n-block): the 1x1 head is evaluated in the epilogue on the bf16-rounded
  // activations, logits[pixel] = sum_c act[c] * head_w[c] + head_b, and the activation tile is NOT stored (nothing else
  // reads it in inference).  head_logits: fp32 [B, H, W]; null = plain epilogue.
  const float* head_w;
  const float* head_b;
  float* head_logits;
  // Eval-mode skip producers: the 2x2 max-pooled copy of the output tile (NHWC bf16, pitch Ntot, half resolution) is
  // written from the staged tile in the epilogue — no separate max-pool pass.  H and W must be even.
  void* pool_out;
};

cudaError_t launch_pix_gemm(const PixGemmParams& p, int block_n, int num_sms, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// "Channel-major" implicit GEMM (weight gradients):
//   dW[tap][co, ci] += sum over pixels  dY[pixel, co] * X[pixel + shift(tap), ci]
//   Both operands are NHWC tiles, i.e. MN-major for the MMA (pixels are the K dimension).
//   One CTA owns (128 output channels) x (BLOCK_N input channels) x (R taps of one group g) and a
//   contiguous range of pixel tiles (split-K); partial sums are reduced with red.global.add.f32
//   into a zero-initialised fp32 buffer laid out [G*R][Mtot][Ntot].
struct WgradParams {
  CUtensorMap tmapDY[4];   // indexed by dy_map[g]
  CUtensorMap tmapX[4];    // indexed by x_map[g]
  int G, R;
  int dy_map[4], x_map[4], x_dw[4], x_dh[4];
  int dy_chan0, x_chan0;
  int Mtot, Ntot;          // Cout-like, Cin-like extents of dW
  int m_blocks, n_blocks;  // ceil(Mtot/128), Ntot/BLOCK_N
  int tiles_w, tiles_h, batch;
  int H, W;                // image extent (the X transform zeroes the pixels outside it)
  int splits;              // split-K factor over pixel tiles
  float* dw;               // [G*R][Mtot][Ntot] fp32, accumulated atomically
  // wgrad9_gemm_kernel (3x3 convolutions with Cout == 64): all nine taps in one CTA from ONE (8+2) x (16+2) patch of X
  int nine;                // 1: use wgrad9_gemm_kernel with tmapX9
  CUtensorMap tmapX9;      // box (64 ch, 10, 18, 1)
  // X is the RAW output of the previous convolution: the (otherwise idle) epilogue warps apply its BatchNorm + ReLU to
  // every X patch in shared memory (see PixGemmParams::in_scale).  Indexed by channel relative to x_chan0; null = plain.
  const float* x_scale;
  const float* x_shift;
};

cudaError_t launch_wgrad_gemm(const WgradParams& p, int block_n, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// Host helpers (tensor-map encoding through the driver entry point; no libcuda link dependency).
int make_tmap_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                 const uint32_t box[4]);
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows);
// fp32 NCHW image as (W, H, C, B) with the stem's (16, 18, C, 1) patch box, no swizzle, zero fill outside the image
int make_stem_tmap(CUtensorMap* out, const float* x, int B, int C, int H, int W);

}  // namespace cs
