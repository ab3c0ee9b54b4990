// U-Net plan: workspace layout, TMA descriptors and the forward / backward launch sequences, plus the
// C ABI declared in include/cartseg.h.
//
// Network (reference src/create_testset.py:40-83, logits = final_conv output):
//   conv index  0..9   encoder  conv1.0 conv1.3 conv2.0 ... conv5.3   (level L = i/2 + 1)
//               10..17 decoder  dconv4.0 dconv4.3 ... dconv1.0 dconv1.3
//   up index    0..3   upconv4 upconv3 upconv2 upconv1
// Activations are NHWC bf16.  The skip activations and the conv-transpose outputs are written
// straight into the concat buffers ([pixels][2C]: up-sampled half first, then the skip — :78-81),
// so torch.cat never materialises.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cmath>

#include "../../include/cartseg.h"
#include "igemm.cuh"
#include "kernels.cuh"

namespace cs {

std::atomic<long long> g_kernel_launches{0};
static thread_local char g_err[512] = "";

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}

#define CS_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define CS_TRY(expr)                \
  do {                              \
    int r__ = (expr);               \
    if (r__ != 0) return r__;       \
  } while (0)

static int device_sm_count(int* out);
// SM count used by the single-layer entry points; CARTSEG_LAYER_SMS=n (read on every call) caps it, so that a test can run
// the same layer with different persistent-grid sizes in one process.
static int layer_sm_count(int* out) {
  int r = device_sm_count(out);
  if (r != 0) return r;
  const char* e = getenv("CARTSEG_LAYER_SMS");
  if (e) {
    int n = atoi(e) & ~1;
    if (n >= 2 && n < *out) *out = n;
  }
  return 0;
}
static int device_sm_count(int* out) {
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  int major = 0;
  CS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail("cartseg needs an sm_100a device (found compute capability %d.x); there is no fallback", major);
  CS_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

// --------------------------------------------------------------------------------------------
// NHWC view: `pitch` channels per pixel, the tensor of interest starting at channel c0.
struct View {
  bf16* p = nullptr;
  int pitch = 0, c0 = 0;
};

// TMA map over an NHWC buffer seen as (channels, W, H, B) with a (64, 8, rows, 1) box.
static int nhwc_map(CUtensorMap* m, const bf16* base, int pitch, int B, int H, int W, int rows) {
  const uint64_t dims[4] = {(uint64_t)pitch, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t st[3] = {(uint64_t)pitch * 2, (uint64_t)W * pitch * 2, (uint64_t)H * W * pitch * 2};
  const uint32_t box[4] = {64, 8, (uint32_t)rows, 1};
  int r = make_tmap_4d(m, base, dims, st, box);
  if (r != 0) return fail("cuTensorMapEncodeTiled (NHWC %dx%dx%dx%d, pitch %d) failed: %d", B, H, W, pitch, pitch, r);
  return 0;
}
// Same buffer at twice the resolution, sampled at pixels (2h+i, 2w+j): the conv-transpose scatter /
// gather views.  (H, W) are the COARSE extents; the buffer is [B][2H][2W][pitch].
static int nhwc_map_strided(CUtensorMap* m, const bf16* base, int pitch, int B, int H, int W, int i, int j) {
  const uint64_t dims[4] = {(uint64_t)pitch, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t st[3] = {(uint64_t)2 * pitch * 2, (uint64_t)2 * (2 * W) * pitch * 2,
                          (uint64_t)(2 * H) * (2 * W) * pitch * 2};
  const uint32_t box[4] = {64, 8, 16, 1};
  int r = make_tmap_4d(m, base + ((size_t)i * (2 * W) + j) * pitch, dims, st, box);
  if (r != 0) return fail("cuTensorMapEncodeTiled (strided NHWC) failed: %d", r);
  return 0;
}
static int nhwc_patch_map(CUtensorMap* m, const bf16* base, int pitch, int B, int H, int W, int pw) {
  const uint64_t dims[4] = {(uint64_t)pitch, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t st[3] = {(uint64_t)pitch * 2, (uint64_t)W * pitch * 2, (uint64_t)H * W * pitch * 2};
  const uint32_t box[4] = {64, (uint32_t)pw, 18, 1};
  int r = make_tmap_4d(m, base, dims, st, box);
  if (r != 0) return fail("cuTensorMapEncodeTiled (NHWC patch %dx%dx%dx%d, pitch %d) failed: %d", B, H, W, pitch, pitch, r);
  return 0;
}
// The pixel GEMMs are CTA-pair kernels (cta_group::2): each CTA of the pair stages half of the weight rows.
static bool use_pair() { return true; }

static int weight_map(CUtensorMap* m, const bf16* base, int K, int rows, int block_n) {
  const int box_rows = use_pair() ? block_n / 2 : block_n;
  int r = make_tmap_2d(m, base, (uint64_t)K, (uint64_t)rows, (uint64_t)K * 2, 64, (uint32_t)box_rows);
  if (r != 0) return fail("cuTensorMapEncodeTiled (weights %d x %d) failed: %d", rows, K, r);
  return 0;
}

static int pick_block_n(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }

// 3x3 convolutions go through conv3_gemm_kernel (one activation patch per K-chunk for all nine taps);
// CARTSEG_CONV3=0 selects pix_gemm2_kernel (three patches per K-chunk) for same-box A/B runs.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int conv3_enabled() { static const int v = env_int("CARTSEG_CONV3", 1); return v; }
static int wgrad9_enabled() { static const int v = env_int("CARTSEG_WGRAD9", 1); return v; }
// First convolution through stem_gemm_kernel (im2col rows built in shared memory); CARTSEG_STEM=0 restores
// im2col_first_kernel + the pointwise GEMM on the forward path.
static int stem_enabled() { static const int v = env_int("CARTSEG_STEM", 1); return v; }
// CARTSEG_XFORM=1: BatchNorm + ReLU of every convX.0 / dconvX.0 is applied by its consumer convX.3 (transform warps on the
// forward operand patch and on the weight gradient's X operand, igemm.cu bnrelu_patch) instead of an elementwise pass
// that writes the activation: 3.1 GB less HBM traffic per K2 step and no activation buffer between the two convolutions
// of a DoubleConv.  Default OFF — measured on B200 (round 2, K2, same box): 17.02 ms with it (layers with Cout >= 128
// only: 17.01 ms) against 16.98 ms without.  The 0.74 ms of bn_relu passes it removes come back inside the GEMMs: the
// transform's shared-memory reads and writes go through the same data pipe as the tensor core's operand fetches (ncu:
// LSU 32 % + tensor-core 44 % of the pipe in wgrad<128>, the MMA warp waiting for operands a third of the time;
// conv3<64>, already paced by operand bandwidth, 230 -> 363 us; wgrad<128> +70 us per launch; only the N = 256 layers,
// which have bandwidth to spare, are unaffected).  More transform warps do not help (the pipe, not instruction issue,
// is the limit).  Kept as an option with its own parity tests (tests/test_gpu_layers.py, test_gpu_unet_stages.py).
static int xform_enabled() { static const int v = env_int("CARTSEG_XFORM", 0); return v; }
// Conv-transpose bias gradients (pixel sums of the up-sampled half of the concat gradient) come out of the dgrad epilogue of
// dconvL.0, which writes that tensor (its BN-statistics path: per-channel column sums of the stored bf16 tile, fp64
// atomics) instead of a separate pass that re-reads 0.77 GB per K2 step; CARTSEG_UPBIAS_FUSED=0 restores channel_sum_kernel.
static int upbias_fused() { static const int v = env_int("CARTSEG_UPBIAS_FUSED", 1); return v; }
static int xform_min_cout() { static const int v = env_int("CARTSEG_XFORM_MIN_COUT", 64); return v; }
// TMA map over an NHWC buffer with a (64, pw, 18, 1) box: the whole halo patch of an 8 x 16 pixel tile.
static int nhwc_patch_map(CUtensorMap* m, const bf16* base, int pitch, int B, int H, int W, int pw);

static void pix_common(PixGemmParams& p, int B, int H, int W, int K, int Ntot, int block_n) {
  memset(&p, 0, sizeof(p));
  p.kchunks = K / 64;
  p.Ntot = Ntot;
  p.n_blocks = Ntot / block_n;
  p.tiles_w = (W + 7) / 8;
  p.tiles_h = (H + 15) / 16;
  p.batch = B;
  p.H = H;
  p.W = W;
  p.o_blocks_per_map = p.n_blocks;
  p.cols_per_map = Ntot;
  p.pair = use_pair() ? 1 : 0;
}

// 3x3 convolution as 9 shifted GEMMs: group g = horizontal tap (dw = g-1), r = vertical tap.
static int build_conv3x3(PixGemmParams& p, int* block_n, View in, int K, View out, int N, const bf16* wpack, int B,
                         int H, int W) {
  *block_n = pick_block_n(N);
  pix_common(p, B, H, W, K, N, *block_n);
  p.G = 3;
  p.R = 3;
  for (int g = 0; g < 3; ++g) { p.a_map[g] = 0; p.a_dw[g] = g - 1; p.a_dh[g] = -1; }
  p.a_chan0 = in.c0;
  p.o_chan0 = out.c0;
  CS_TRY(nhwc_map(&p.tmapA[0], in.p, in.pitch, B, H, W, 18));
  CS_TRY(weight_map(&p.tmapB, wpack, K, 9 * N, *block_n));
  CS_TRY(nhwc_map(&p.tmapO[0], out.p, out.pitch, B, H, W, 16));
  if (use_pair() && conv3_enabled()) {
    p.conv3 = 1;
    CS_TRY(nhwc_patch_map(&p.tmapA3, in.p, in.pitch, B, H, W, 10));
  }
  return 0;
}
// Plain [pixels x K] * [K x N] (the im2col'd first conv).
static int build_pointwise(PixGemmParams& p, int* block_n, View in, int K, View out, int N, const bf16* wpack,
                           int B, int H, int W) {
  *block_n = pick_block_n(N);
  pix_common(p, B, H, W, K, N, *block_n);
  p.G = 1;
  p.R = 1;
  p.a_chan0 = in.c0;
  p.o_chan0 = out.c0;
  CS_TRY(nhwc_map(&p.tmapA[0], in.p, in.pitch, B, H, W, 16));
  CS_TRY(weight_map(&p.tmapB, wpack, K, N, *block_n));
  CS_TRY(nhwc_map(&p.tmapO[0], out.p, out.pitch, B, H, W, 16));
  return 0;
}
// ConvTranspose2d(k=2, s=2): out[2h+i, 2w+j, co] = sum_ci in[h, w, ci] * W[ci, co, i, j] (+ bias).
// (H, W) = input extents; `out` has twice the resolution.  wpack = [4][Cout][Cin].
static int build_convT_fprop(PixGemmParams& p, int* block_n, View in, int Cin, View out, int Cout, const bf16* wpack,
                             const float* bias, int B, int H, int W) {
  // the pair kernel lets one n-block span several (i,j) output maps: N = 256 even for Cout = 64
  *block_n = use_pair() ? pick_block_n(4 * Cout) : pick_block_n(Cout);
  // short K (Cin <= 256): the launch is paced by its epilogue and by operand traffic from L2, not by the MMAs -> the
  // N = 128 kernel with its two epilogue groups (up3 fprop 182 -> 134 us; Cin = 512 is slower that way: 75 -> 110 us)
  if (use_pair() && Cin <= 256 && *block_n == 256) *block_n = 128;
  pix_common(p, B, H, W, Cin, 4 * Cout, *block_n);
  p.G = 1;
  p.R = 1;
  p.a_chan0 = in.c0;
  p.o_chan0 = out.c0;
  p.o_blocks_per_map = Cout / *block_n;
  p.cols_per_map = Cout;
  p.shift = bias;
  CS_TRY(nhwc_map(&p.tmapA[0], in.p, in.pitch, B, H, W, 16));
  CS_TRY(weight_map(&p.tmapB, wpack, Cin, 4 * Cout, *block_n));
  for (int ij = 0; ij < 4; ++ij) CS_TRY(nhwc_map_strided(&p.tmapO[ij], out.p, out.pitch, B, H, W, ij >> 1, ij & 1));
  return 0;
}
// Its input gradient: dx[h, w, ci] = sum_{ij, co} dy[2h+i, 2w+j, co] * W[ci, co, i, j].  wpack = [4][Cin][Cout].
static int build_convT_dgrad(PixGemmParams& p, int* block_n, View dy, int Cout, View dx, int Cin, const bf16* wpack,
                             int B, int H, int W) {
  *block_n = pick_block_n(Cin);
  pix_common(p, B, H, W, Cout, Cin, *block_n);
  p.G = 4;
  p.R = 1;
  for (int g = 0; g < 4; ++g) {
    p.a_map[g] = g;
    CS_TRY(nhwc_map_strided(&p.tmapA[g], dy.p, dy.pitch, B, H, W, g >> 1, g & 1));
  }
  p.a_chan0 = dy.c0;
  p.o_chan0 = dx.c0;
  CS_TRY(weight_map(&p.tmapB, wpack, Cout, 4 * Cin, *block_n));
  CS_TRY(nhwc_map(&p.tmapO[0], dx.p, dx.pitch, B, H, W, 16));
  return 0;
}

static int choose_wgrad_splits(int base, int tiles, int sms) {
  // Split-K factor.  One CTA per SM is resident (shared memory), so the grid runs in waves of 148; a CTA costs its
  // share of the pixel tiles plus a fixed prologue / accumulator drain worth ~6 tiles.  The kernel's own time is
  // waves x (tiles per CTA + overhead): 13 splits of a 48-item layer are 624 CTAs = five waves with the last one 22 %
  // full; the minimum is one full wave of long-lived CTAs for every layer of this network.
  // CARTSEG_WGRAD_INVERSION_WEIGHT=w adds w CTA lifetimes to the cost, i.e. prefers several full waves so that a
  // high-priority dgrad that becomes ready while a wgrad wave is resident waits less (CTAs are not preemptible).
  // Measured (tools/trace_backward.py, k2): w = 1 cuts the up-conv dgrads' waiting from 0.3-0.5 ms to nothing, but the
  // weight gradients, now pre-empted at every wave boundary, pile up behind the main stream and the step gets
  // slower (19.5 vs 18.8 ms; w = 2: 20.6 ms).  Default 0.
  const int kSMs = sms, kOverheadTiles = 6;
  static const int kInversionWeight = [] {
    const char* e = getenv("CARTSEG_WGRAD_INVERSION_WEIGHT");
    return e ? atoi(e) : 0;
  }();
  int max_splits = tiles / 16;
  if (max_splits < 1) max_splits = 1;
  if (max_splits > 16 * kSMs) max_splits = 16 * kSMs;
  int splits = 1;
  long long best_cost = -1;
  for (int sp = 1; sp <= max_splits; ++sp) {
    const long long waves = ((long long)base * sp + kSMs - 1) / kSMs;
    const long long cost = (waves + kInversionWeight) * ((tiles + sp - 1) / sp + kOverheadTiles);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = sp; }
  }
  return splits;
}

static void wgrad_common(WgradParams& p, int* block_n, int M, int N, int B, int H, int W, int G, float* dw) {
  memset(&p, 0, sizeof(p));
  *block_n = N % 128 == 0 ? 128 : 64;
  p.G = G;
  p.Mtot = M;
  p.Ntot = N;
  p.m_blocks = (M + 127) / 128;
  p.n_blocks = N / *block_n;
  p.tiles_w = (W + 7) / 8;
  p.tiles_h = (H + 15) / 16;
  p.batch = B;
  p.H = H;
  p.W = W;
  const int tiles = p.tiles_w * p.tiles_h * B;
  const int base = p.m_blocks * p.n_blocks * G;
  p.splits = choose_wgrad_splits(base, tiles, 148);
  p.dw = dw;
}
// dW[kw*3+kh][co][ci] = sum_pixels dy[pixel, co] * x[pixel + (kh-1, kw-1), ci]
static int build_conv3x3_wgrad(WgradParams& p, int* block_n, View x, int Cin, View dy, int Cout, float* dw, int B,
                               int H, int W) {
  wgrad_common(p, block_n, Cout, Cin, B, H, W, 3, dw);
  p.R = 3;
  for (int g = 0; g < 3; ++g) { p.x_dw[g] = g - 1; p.x_dh[g] = -1; }
  p.dy_chan0 = dy.c0;
  p.x_chan0 = x.c0;
  CS_TRY(nhwc_map(&p.tmapDY[0], dy.p, dy.pitch, B, H, W, 16));
  CS_TRY(nhwc_map(&p.tmapX[0], x.p, x.pitch, B, H, W, 18));
  if (Cout == 64 && wgrad9_enabled()) {                  // all nine taps per CTA from one 10 x 18 patch (wgrad9_gemm_kernel)
    p.nine = 1;
    p.n_blocks = Cin / 64;
    p.splits = choose_wgrad_splits(p.n_blocks, p.tiles_w * p.tiles_h * B, 148);
    CS_TRY(nhwc_patch_map(&p.tmapX9, x.p, x.pitch, B, H, W, 10));
  }
  return 0;
}
// dW[co][k] = sum_pixels dy[pixel, co] * col[pixel, k]   (first conv on its im2col matrix)
static int build_pointwise_wgrad(WgradParams& p, int* block_n, View x, int K, View dy, int Cout, float* dw, int B,
                                 int H, int W) {
  wgrad_common(p, block_n, Cout, K, B, H, W, 1, dw);
  p.R = 1;
  p.dy_chan0 = dy.c0;
  p.x_chan0 = x.c0;
  CS_TRY(nhwc_map(&p.tmapDY[0], dy.p, dy.pitch, B, H, W, 16));
  CS_TRY(nhwc_map(&p.tmapX[0], x.p, x.pitch, B, H, W, 16));
  return 0;
}
// dW[ij][ci][co] = sum_pixels x[pixel, ci] * dy[(2h+i, 2w+j), co]; the coarse input x plays the
// "M" operand, the four strided views of dy the "N" operand.
static int build_convT_wgrad(WgradParams& p, int* block_n, View x, int Cin, View dy, int Cout, float* dw, int B,
                             int H, int W) {
  wgrad_common(p, block_n, Cin, Cout, B, H, W, 4, dw);
  p.R = 1;
  for (int g = 0; g < 4; ++g) {
    p.x_map[g] = g;
    CS_TRY(nhwc_map_strided(&p.tmapX[g], dy.p, dy.pitch, B, H, W, g >> 1, g & 1));
  }
  p.dy_chan0 = x.c0;
  p.x_chan0 = dy.c0;
  CS_TRY(nhwc_map(&p.tmapDY[0], x.p, x.pitch, B, H, W, 16));
  return 0;
}

static const TapMap kTapFprop = {{0, 3, 6, 1, 4, 7, 2, 5, 8}};   // (kh,kw) -> kw*3 + kh
static const TapMap kTapDgrad = {{8, 5, 2, 7, 4, 1, 6, 3, 0}};   // (kh,kw) -> (2-kw)*3 + (2-kh)
static const TapMap kTapWgrad = {{0, 3, 6, 1, 4, 7, 2, 5, 8}};   // packed tap g*3+r -> kh*3+kw = r*3+g
static const TapMap kTapIdent = {{0, 1, 2, 3, 4, 5, 6, 7, 8}};

}  // namespace cs

using namespace cs;

// =============================================================================================
//                                         the plan
// =============================================================================================
struct ConvL {
  int cin, cout, H, W;
  long long P;                       // B*H*W
  int pw, pb, pgamma, pbeta;         // parameter indices
  View in, out, g_out, g_in;         // input act, post-ReLU act, grad wrt out, grad wrt in (p == null: none)
  bf16 *y, *dy, *pooled, *g_pool;
  bf16 *wf, *wd;
  double *st_sum, *st_sq;
  double *dg_sum, *dg_sq;            // dconvL.0 only: column sums of its dgrad output (the conv-transpose bias gradient)
  float *bc1, *bc2;
  float *scale, *shift, *mean, *invstd;
  PixGemmParams fp_train, fp_eval, dg;
  WgradParams wg;
  int bn_f, bn_d, bn_w;
  bool act_fused;                    // training: the activation is never stored; the next conv applies BN + ReLU to y itself
  double flops;                      // algorithmic 2*MACs of one pass (fprop == dgrad == wgrad)
};
struct UpL {
  int cin, cout, H, W;               // H, W: input (coarse) extents
  long long P;
  int pw, pb;
  View in, out, g_out, g_in;         // out / g_out live in the concat buffers (pitch 2*cout, c0 = 0)
  bf16 *wf, *wd;
  PixGemmParams fp, dg;
  WgradParams wg;
  int bn_f, bn_d, bn_w;
  double flops;
};

struct cs_unet_plan {
  int B, Cin, H, W;
  int num_sms;
  size_t ws_bytes;
  uint8_t* ws;
  bool bound, forward_done, infer;
  bool eval_act17_missing;           // the last eval forward did not store dconv1.3's activation (head fused in its epilogue)
  ConvL conv[18];
  UpL up[4];
  bf16* col;                         // im2col of the input image [P1][64]
  float* dwp;                        // packed weight-gradient accumulator, shared by all layers
  float* bn_partial;                 // per-block partial sums of the BN backward reduction
  float* dlogits_keep;               // copy of the last dlogits (the head's activation gradient is formed on the fly)
  const float* head_w_keep;          // final_conv.weight of the last backward (device pointer owned by the caller)
  uint8_t* stats_begin;
  size_t stats_bytes;
  // optional per-launch timing of the tensor-core kernels (cs_unet_profile): CUDA events around each GEMM launch
  // internal streams of the backward pass (see cs_unet_backward)
  cudaStream_t s_hi, s_lo;
  cudaEvent_t ev_fork, ev_hi, ev_lo, ev_stage[CS_UNET_NUM_BWD_STAGES];
  bool profiling, no_overlap, deferred_join, join_pending;
  // weight-gradient GEMMs held back until the main stream reaches the encoder's high-resolution levels (see
  // cs_unet_backward): (kind 1 conv / 2 up, layer index, gradient pointers of that launch)
  struct HeldWgrad { int kind, idx, stage; float* gw; float* gb; };
  std::vector<HeldWgrad> held;
  std::vector<HeldWgrad> late;     // wgrads of dconvL.0 that wait for the conv-transpose dgrad right after them
  cudaEvent_t ev_flush;
  const float* x_keep;               // input image of the last training forward (the stem's weight gradient re-reads it)
  std::vector<cudaEvent_t> prof_events;   // pairs (begin, end)
  std::vector<int> prof_class;
  std::vector<double> prof_flops;
  size_t prof_used;
  // developer timeline of the backward pass (cs_unet_trace): an event pair around every launch, on either stream
  bool tracing;
  std::vector<cudaEvent_t> trace_events;
  std::vector<int> trace_label;
  size_t trace_used;
};

namespace {
// kernel classes reported by cs_unet_profile_read
enum { kClsPix256 = 0, kClsPix128, kClsPix64, kClsWgrad128, kClsWgrad64, kClsConv256, kClsConv128, kClsConv64, kClsWgrad9, kNumCls };
static_assert(kNumCls == CS_UNET_NUM_PROFILE_CLASSES, "profile classes");
int pix_class(int bn) { return bn == 256 ? kClsPix256 : (bn == 128 ? kClsPix128 : kClsPix64); }
// 3x3 convolutions run on conv3_gemm_kernel unless CARTSEG_CONV3=0
int conv_class(const PixGemmParams& p, int bn) {
  if (!p.conv3) return pix_class(bn);
  return bn == 256 ? kClsConv256 : (bn == 128 ? kClsConv128 : kClsConv64);
}
int wgrad_class(int bn) { return bn == 128 ? kClsWgrad128 : kClsWgrad64; }
int wgrad_class(const WgradParams& w, int bn) { return w.nine ? kClsWgrad9 : wgrad_class(bn); }

// Runs `launch` between two events on `s` when profiling is on.
template <typename F>
cudaError_t timed(cs_unet_plan* pl, int cls, double flops, cudaStream_t s, F&& launch) {
  if (!pl->profiling) return launch();
  if (pl->prof_events.size() < 2 * (pl->prof_used + 1)) {
    cudaEvent_t a, b;
    cudaError_t e = cudaEventCreate(&a);
    if (e != cudaSuccess) return e;
    e = cudaEventCreate(&b);
    if (e != cudaSuccess) return e;
    pl->prof_events.push_back(a);
    pl->prof_events.push_back(b);
    pl->prof_class.push_back(0);
    pl->prof_flops.push_back(0.0);
  }
  const size_t i = pl->prof_used++;
  pl->prof_class[i] = cls;
  pl->prof_flops[i] = flops;
  cudaError_t e = cudaEventRecord(pl->prof_events[2 * i], s);
  if (e != cudaSuccess) return e;
  e = launch();
  if (e != cudaSuccess) return e;
  return cudaEventRecord(pl->prof_events[2 * i + 1], s);
}
// Runs `launch` between two timing events on `s` when tracing is on (label = kind * 100 + layer index).
template <typename F>
cudaError_t traced(cs_unet_plan* pl, int label, cudaStream_t s, F&& launch) {
  if (!pl->tracing) return launch();
  if (pl->trace_events.size() < 2 * (pl->trace_used + 1)) {
    cudaEvent_t a, b;
    cudaError_t e = cudaEventCreate(&a);
    if (e != cudaSuccess) return e;
    e = cudaEventCreate(&b);
    if (e != cudaSuccess) return e;
    pl->trace_events.push_back(a);
    pl->trace_events.push_back(b);
    pl->trace_label.push_back(0);
  }
  const size_t i = pl->trace_used++;
  pl->trace_label[i] = label;
  cudaError_t e = cudaEventRecord(pl->trace_events[2 * i], s);
  if (e != cudaSuccess) return e;
  e = launch();
  if (e != cudaSuccess) return e;
  return cudaEventRecord(pl->trace_events[2 * i + 1], s);
}
}  // namespace

namespace {

struct Arena {
  uint8_t* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    off = (off + 1023) & ~(size_t)1023;
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

int conv_param_base(int i) { return i < 10 ? (i / 2) * 8 + (i % 2) * 4 : 48 + ((i - 10) / 2) * 8 + ((i - 10) % 2) * 4; }

// Assigns every buffer.  Run once with base == nullptr to size the workspace, once more at bind time.
void layout(cs_unet_plan* pl, uint8_t* base) {
  Arena a{base, 0};
  const int B = pl->B;
  const bool train = !pl->infer;
  int LH[6], LW[6], LC[6];
  long long LP[6];
  for (int L = 1; L <= 5; ++L) {
    LH[L] = pl->H >> (L - 1);
    LW[L] = pl->W >> (L - 1);
    LC[L] = 64 << (L - 1);
    LP[L] = (long long)B * LH[L] * LW[L];
  }
  pl->col = a.take<bf16>(LP[1] * 64);
  bf16* cat[5];
  bf16* gcat[5];
  for (int L = 1; L <= 4; ++L) {
    cat[L] = a.take<bf16>(LP[L] * 2 * LC[L]);
    gcat[L] = train ? a.take<bf16>(LP[L] * 2 * LC[L]) : nullptr;
  }
  size_t dwp_elems = 0;
  for (int i = 0; i < 18; ++i) {
    ConvL& c = pl->conv[i];
    int L;
    if (i < 10) {
      L = i / 2 + 1;
      c.cout = LC[L];
      c.cin = (i % 2) ? LC[L] : (L == 1 ? 64 /* padded K of the im2col'd image */ : LC[L - 1]);
    } else {
      L = 4 - (i - 10) / 2;
      c.cout = LC[L];
      c.cin = (i % 2) ? LC[L] : 2 * LC[L];
    }
    c.H = LH[L];
    c.W = LW[L];
    c.P = LP[L];
    c.flops = 2.0 * (double)c.P * c.cout * (i == 0 ? 9.0 * pl->Cin : 9.0 * c.cin);
    const int pb = conv_param_base(i);
    c.pw = pb; c.pb = pb + 1; c.pgamma = pb + 2; c.pbeta = pb + 3;
    c.y = train ? a.take<bf16>(c.P * c.cout) : nullptr;
    c.dy = train ? a.take<bf16>(c.P * c.cout) : nullptr;
    c.pooled = nullptr;
    c.g_pool = nullptr;
    const bool skip_producer = i < 8 && (i % 2) == 1;          // conv1.3 .. conv4.3
    if (skip_producer) {
      c.out = View{cat[L], 2 * LC[L], LC[L]};
      c.g_out = View{gcat[L], 2 * LC[L], LC[L]};
      c.pooled = a.take<bf16>(LP[L + 1] * LC[L]);
      c.g_pool = train ? a.take<bf16>(LP[L + 1] * LC[L]) : nullptr;
    } else {
      c.out = View{a.take<bf16>(c.P * c.cout), c.cout, 0};
      c.g_out = View{train ? a.take<bf16>(c.P * c.cout) : nullptr, c.cout, 0};
    }
    c.wf = a.take<bf16>((size_t)9 * c.cout * c.cin);
    c.wd = train ? a.take<bf16>((size_t)9 * c.cout * c.cin) : nullptr;
    c.scale = a.take<float>(c.cout);
    c.shift = a.take<float>(c.cout);
    c.mean = a.take<float>(c.cout);
    c.invstd = a.take<float>(c.cout);
    const size_t e = (size_t)(i == 0 ? 1 : 9) * c.cout * c.cin;
    if (e > dwp_elems) dwp_elems = e;
  }
  // inputs / input gradients (wired after all outputs exist)
  for (int i = 0; i < 18; ++i) {
    ConvL& c = pl->conv[i];
    if (i == 0) {
      c.in = View{pl->col, 64, 0};
      c.g_in = View{};
    } else if (i < 10 && (i % 2) == 0) {                        // convL.0, L >= 2: pooled skip of the level above
      c.in = View{pl->conv[i - 1].pooled, pl->conv[i - 1].cout, 0};
      c.g_in = View{pl->conv[i - 1].g_pool, pl->conv[i - 1].cout, 0};
    } else if (i >= 10 && (i % 2) == 0) {                       // dconvL.0: the concat buffer
      const int L = 4 - (i - 10) / 2;
      c.in = View{cat[L], 2 * LC[L], 0};
      c.g_in = View{gcat[L], 2 * LC[L], 0};
    } else {                                                    // X.3: output of X.0
      c.in = pl->conv[i - 1].out;
      c.g_in = pl->conv[i - 1].g_out;
    }
  }
  for (int k = 0; k < 4; ++k) {
    UpL& u = pl->up[k];
    const int L = 4 - k;                                        // output level
    u.cin = LC[L + 1];
    u.cout = LC[L];
    u.H = LH[L + 1];
    u.W = LW[L + 1];
    u.P = LP[L + 1];
    u.flops = 2.0 * (double)u.P * u.cin * u.cout * 4.0;
    u.pw = 40 + 2 * k;
    u.pb = 41 + 2 * k;
    const ConvL& src = k == 0 ? pl->conv[9] : pl->conv[10 + 2 * (k - 1) + 1];
    u.in = src.out;
    u.g_in = src.g_out;
    u.out = View{cat[L], 2 * LC[L], 0};
    u.g_out = View{gcat[L], 2 * LC[L], 0};
    u.wf = a.take<bf16>((size_t)4 * u.cin * u.cout);
    u.wd = train ? a.take<bf16>((size_t)4 * u.cin * u.cout) : nullptr;
    const size_t e = (size_t)4 * u.cin * u.cout;
    if (e > dwp_elems) dwp_elems = e;
  }
  pl->dwp = train ? a.take<float>(dwp_elems) : nullptr;
  pl->bn_partial = train ? reinterpret_cast<float*>(a.take<uint8_t>(bn_bwd_scratch_bytes(1024))) : nullptr;
  pl->dlogits_keep = train ? a.take<float>((size_t)pl->B * pl->H * pl->W) : nullptr;
  // statistics (zeroed once per forward): forward sums and backward sums of every BN
  a.off = (a.off + 1023) & ~(size_t)1023;
  pl->stats_begin = base + a.off;
  const size_t s0 = a.off;
  for (int i = 0; i < 18; ++i) {
    ConvL& c = pl->conv[i];
    c.st_sum = a.take<double>(2 * (size_t)c.cout);
    c.st_sq = c.st_sum + c.cout;
    c.bc1 = a.take<float>(2 * (size_t)c.cout);
    c.bc2 = c.bc1 + c.cout;
    c.dg_sum = c.dg_sq = nullptr;
    if (i >= 10 && (i % 2) == 0) {
      c.dg_sum = a.take<double>(2 * (size_t)c.cin);
      c.dg_sq = c.dg_sum + c.cin;
    }
  }
  pl->stats_bytes = a.off - s0;
  pl->ws_bytes = (a.off + 1023) & ~(size_t)1023;
}

int encode_maps(cs_unet_plan* pl) {
  const int B = pl->B;
  for (int i = 0; i < 18; ++i) {
    ConvL& c = pl->conv[i];
    const View y{c.y, c.cout, 0}, dy{c.dy, c.cout, 0};
    const bool train = !pl->infer;
    if (i == 0) {
      CS_TRY(build_pointwise(c.fp_eval, &c.bn_f, c.in, 64, c.out, c.cout, c.wf, B, c.H, c.W));
      if (train) {
        CS_TRY(build_pointwise(c.fp_train, &c.bn_f, c.in, 64, y, c.cout, c.wf, B, c.H, c.W));
        CS_TRY(build_pointwise_wgrad(c.wg, &c.bn_w, c.in, 64, dy, c.cout, pl->dwp, B, c.H, c.W));
      }
      c.bn_d = 0;
    } else {
      CS_TRY(build_conv3x3(c.fp_eval, &c.bn_f, c.in, c.cin, c.out, c.cout, c.wf, B, c.H, c.W));
      if (train) {
        // convX.3 reads the RAW output of convX.0 and applies its BatchNorm + ReLU in shared memory (forward operand and
        // weight-gradient operand): the activation between the two convolutions of a DoubleConv never exists in HBM
        ConvL& prev = pl->conv[i - 1];
        const bool fused_in = (i % 2) == 1 && xform_enabled() && c.fp_eval.conv3 && c.cout >= xform_min_cout();
        prev.act_fused = fused_in;
        const View xin = fused_in ? View{prev.y, prev.cout, 0} : c.in;
        CS_TRY(build_conv3x3(c.fp_train, &c.bn_f, xin, c.cin, y, c.cout, c.wf, B, c.H, c.W));
        CS_TRY(build_conv3x3(c.dg, &c.bn_d, dy, c.cout, c.g_in, c.cin, c.wd, B, c.H, c.W));
        CS_TRY(build_conv3x3_wgrad(c.wg, &c.bn_w, xin, c.cin, dy, c.cout, pl->dwp, B, c.H, c.W));
        if (fused_in) {
          c.fp_train.in_scale = prev.scale; c.fp_train.in_shift = prev.shift;
          c.wg.x_scale = prev.scale; c.wg.x_shift = prev.shift;
        }
      }
    }
    c.fp_train.stat_sum = c.st_sum;
    c.fp_train.stat_sq = c.st_sq;
    c.fp_eval.scale = c.scale;
    c.fp_eval.shift = c.shift;
    c.fp_eval.relu = 1;
  }
  for (int k = 0; k < 4; ++k) {
    UpL& u = pl->up[k];
    CS_TRY(build_convT_fprop(u.fp, &u.bn_f, u.in, u.cin, u.out, u.cout, u.wf, nullptr, B, u.H, u.W));
    if (!pl->infer) {
      CS_TRY(build_convT_dgrad(u.dg, &u.bn_d, u.g_out, u.cout, u.g_in, u.cin, u.wd, B, u.H, u.W));
      CS_TRY(build_convT_wgrad(u.wg, &u.bn_w, u.in, u.cin, u.g_out, u.cout, pl->dwp, B, u.H, u.W));
    }
  }
  return 0;
}

}  // namespace

// =============================================================================================
//                                          C ABI
// =============================================================================================
extern "C" {

const char* cs_last_error(void) { return g_err; }
int cs_version(void) { return 1; }
long long cs_kernel_launch_count(void) { return g_kernel_launches.load(); }

int cs_unet_plan_create(cs_unet_plan** out, int batch, int in_channels, int height, int width, int inference_only) {
  if (!out) return fail("plan pointer is null");
  *out = nullptr;
  if (batch < 1) return fail("batch must be >= 1 (got %d)", batch);
  if (in_channels < 1 || in_channels > 7) return fail("in_channels must be in [1, 7] (got %d)", in_channels);
  if (height < 16 || width < 16 || height % 16 || width % 16)
    return fail("height and width must be positive multiples of 16 (got %d x %d)", height, width);
  cs_unet_plan* pl = new (std::nothrow) cs_unet_plan();   // value-initialised: all POD members zero
  if (!pl) return fail("out of host memory");
  pl->B = batch; pl->Cin = in_channels; pl->H = height; pl->W = width;
  pl->infer = inference_only != 0;
  layout(pl, nullptr);
  *out = pl;
  return 0;
}

void cs_unet_plan_destroy(cs_unet_plan* plan) {
  if (!plan) return;
  for (cudaEvent_t e : plan->prof_events) cudaEventDestroy(e);
  if (plan->s_hi) {
    cudaStreamDestroy(plan->s_hi);
    cudaStreamDestroy(plan->s_lo);
    cudaEventDestroy(plan->ev_fork);
    cudaEventDestroy(plan->ev_hi);
    cudaEventDestroy(plan->ev_lo);
    cudaEventDestroy(plan->ev_flush);
    for (cudaEvent_t e : plan->ev_stage) cudaEventDestroy(e);
  }
  delete plan;
}

size_t cs_unet_plan_workspace_bytes(const cs_unet_plan* plan) { return plan ? plan->ws_bytes : 0; }

static int ensure_streams(cs_unet_plan* pl);

int cs_unet_plan_bind(cs_unet_plan* pl, void* workspace, size_t bytes) {
  if (!pl) return fail("plan is null");
  if (!workspace || bytes < pl->ws_bytes) return fail("workspace too small: %zu < %zu bytes", bytes, pl->ws_bytes);
  if ((uintptr_t)workspace & 1023) return fail("workspace must be 1024-byte aligned");
  CS_TRY(device_sm_count(&pl->num_sms));
  pl->ws = static_cast<uint8_t*>(workspace);
  layout(pl, pl->ws);
  CS_TRY(encode_maps(pl));
  if (!pl->infer) CS_TRY(ensure_streams(pl));   // not lazily inside cs_unet_backward: illegal during stream capture
  pl->bound = true;
  pl->forward_done = false;
  return 0;
}

int cs_unet_pack_weights(cs_unet_plan* pl, const cs_unet_tensors* t, cs_stream_t stream) {
  if (!pl || !pl->bound) return fail("plan is not bound to a workspace");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // one launch for all 22 weight tensors (was 22 launches per call, twice per training step with the unpacks)
  PackBatch pb;
  memset(&pb, 0, sizeof(pb));
  pb.maps[0] = kTapFprop; pb.maps[1] = kTapDgrad; pb.maps[2] = kTapIdent;
  for (int i = 0; i < 18; ++i) {
    const ConvL& c = pl->conv[i];
    if (!t->param[c.pw]) return fail("parameter %d is null", c.pw);
    if (i == 0) {
      pb.first_in = t->param[c.pw]; pb.first_out = c.wf; pb.first_cout = c.cout; pb.first_cin = pl->Cin;
    } else {
      PackJob& j = pb.job[pb.n++];
      j.in = t->param[c.pw]; j.Na = c.cout; j.Nb = c.cin; j.T = 9;
      j.out_ab = c.wf; j.map_ab = 0; j.out_ba = c.wd; j.map_ba = 1;
    }
  }
  for (int k = 0; k < 4; ++k) {
    const UpL& u = pl->up[k];
    if (!t->param[u.pw]) return fail("parameter %d is null", u.pw);
    // IOHW [ci][co][ij]: dgrad pack [ij][ci][co] (a-major), fprop pack [ij][co][ci] (b-major)
    PackJob& j = pb.job[pb.n++];
    j.in = t->param[u.pw]; j.Na = u.cin; j.Nb = u.cout; j.T = 4;
    j.out_ab = u.wd; j.map_ab = 2; j.out_ba = u.wf; j.map_ba = 2;
  }
  CS_CUDA(launch_pack_batch(pb, s));
  return 0;
}

// CARTSEG_FUSE_HEAD=0 restores the stand-alone 1x1 head kernels (A/B measurements).
static bool fuse_head() {
  static const bool v = [] {
    const char* e = getenv("CARTSEG_FUSE_HEAD");
    return !(e && e[0] == '0');
  }();
  return v;
}

// Eval mode: the head is evaluated in the epilogue of dconv1.3 when that layer runs on the resident-weights Cout = 64
// kernel (always, unless CARTSEG_CONV3=0 / CARTSEG_FUSE_HEAD=0 select the stand-alone kernels).
static bool eval_head_fused(const ConvL& c) { return fuse_head() && c.fp_eval.conv3 && c.bn_f == 64 && c.cin <= 128; }

int cs_unet_forward(cs_unet_plan* pl, const cs_unet_tensors* t, const float* x, int training, float* logits,
                    cs_stream_t stream) {
  if (!pl || !pl->bound) return fail("plan is not bound to a workspace");
  if (!x || !logits) return fail("x / logits is null");
  if (training && pl->infer) return fail("this plan was created inference_only: training-mode forward is not available");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = pl->B;
  if (training) CS_CUDA(cudaMemsetAsync(pl->stats_begin, 0, pl->stats_bytes, s));
  if (!training) {                                           // fold the running statistics of all 18 BN layers
    BnFoldBatch f{};
    f.layers = 18;
    f.eps = 1e-5f;
    for (int i = 0; i < 18; ++i) {
      const ConvL& c = pl->conv[i];
      if (!t->running_mean[i] || !t->running_var[i]) return fail("running statistics of BN %d are null", i);
      f.gamma[i] = t->param[c.pgamma]; f.beta[i] = t->param[c.pbeta]; f.conv_bias[i] = t->param[c.pb];
      f.rm[i] = t->running_mean[i]; f.rv[i] = t->running_var[i];
      f.scale[i] = c.scale; f.shift[i] = c.shift; f.C[i] = c.cout;
    }
    CS_CUDA(launch_bn_fold_eval(f, s));
  }
  const bool stem = stem_enabled() != 0;
  if (!stem) {
    CS_CUDA(launch_im2col_first(x, B, pl->Cin, pl->H, pl->W, pl->col, s));
  } else {
    // The forward path builds the im2col rows inside stem_gemm_kernel.  The matrix itself is only needed by the first
    // convolution's weight gradient: cs_unet_backward writes it on its side stream, off the critical path — which is why
    // x has to stay valid until then (include/cartseg.h).
    PixGemmParams& fp = training ? pl->conv[0].fp_train : pl->conv[0].fp_eval;
    if (fp.stem_x != x) {                                    // the patch map over the caller's image (re-encoded when x moves)
      if ((uintptr_t)x & 15) return fail("x must be 16-byte aligned");
      const int r = make_stem_tmap(&fp.tmapX, x, B, pl->Cin, pl->H, pl->W);
      if (r != 0) return fail("cuTensorMapEncodeTiled (stem image) failed: %d", r);
      fp.stem_x = x; fp.stem_cin = pl->Cin;
    }
    pl->x_keep = training ? x : nullptr;
  }

  auto run_conv = [&](int i) -> int {
    ConvL& c = pl->conv[i];
    if (training) {
      CS_CUDA(traced(pl, 1000 + i, s, [&] { return timed(pl, conv_class(c.fp_train, c.bn_f), c.flops, s, [&] { return launch_pix_gemm(c.fp_train, c.bn_f, pl->num_sms, s); }); }));
      BnFinalizeArgs f{};
      f.sum = c.st_sum; f.sq = c.st_sq; f.count = (double)c.P;
      f.gamma = t->param[c.pgamma]; f.beta = t->param[c.pbeta]; f.conv_bias = t->param[c.pb];
      f.running_mean = t->running_mean[i]; f.running_var = t->running_var[i];
      f.num_batches_tracked = t->num_batches_tracked[i];
      f.momentum = 0.1f; f.eps = 1e-5f;
      f.scale = c.scale; f.shift = c.shift; f.mean = c.mean; f.invstd = c.invstd; f.C = c.cout;
      // the last layer's pass also evaluates the 1x1 head (final_conv) on the activations it has just produced
      // — and does not store them at all: nothing reads the head's input in the training path (the BN backward of this
      // layer recomputes it from y for the head's weight gradient), 0.4 GB less traffic at K2
      const HeadFwd head = (i == 17 && fuse_head()) ? HeadFwd{t->param[80], t->param[81], logits, 1} : HeadFwd{nullptr, nullptr, nullptr, 0};
      if (c.act_fused) CS_CUDA(launch_bn_finalize(f, s));    // the next convolution applies BN + ReLU to y on the fly
      else CS_CUDA(traced(pl, 1100 + i, s, [&] { return launch_bn_relu(c.y, B, c.H, c.W, c.cout, f, c.out.p, c.out.pitch, c.out.c0, c.pooled, head, s); }));
    } else {
      // last layer: the 1x1 head rides in the convolution's epilogue and the activation is never written (nothing reads it
      // in inference): one launch and 0.8 GB of traffic less at B = 64
      const bool head_in_epilogue = i == 17 && eval_head_fused(c);
      c.fp_eval.head_w = head_in_epilogue ? t->param[80] : nullptr;
      c.fp_eval.head_b = head_in_epilogue ? t->param[81] : nullptr;
      c.fp_eval.head_logits = head_in_epilogue ? logits : nullptr;
      // skip producers: the 2x2 max-pooled copy comes out of the same epilogue (a separate max-pool pass over the stored
      // activation cost 0.18 ms of a 4.38 ms batch-64 forward: 4.38 -> 4.20 ms, same box)
      c.fp_eval.pool_out = c.pooled;
      CS_CUDA(timed(pl, conv_class(c.fp_eval, c.bn_f), c.flops, s, [&] { return launch_pix_gemm(c.fp_eval, c.bn_f, pl->num_sms, s); }));
    }
    return 0;
  };
  for (int i = 0; i < 10; ++i) CS_TRY(run_conv(i));
  for (int k = 0; k < 4; ++k) {
    UpL& u = pl->up[k];
    u.fp.shift = t->param[u.pb];
    CS_CUDA(traced(pl, 1200 + k, s, [&] { return timed(pl, pix_class(u.bn_f), u.flops, s, [&] { return launch_pix_gemm(u.fp, u.bn_f, pl->num_sms, s); }); }));
    CS_TRY(run_conv(10 + 2 * k));
    CS_TRY(run_conv(11 + 2 * k));
  }
  const ConvL& last = pl->conv[17];
  const bool head_done = training ? fuse_head() : eval_head_fused(last);
  if (!head_done) CS_CUDA(launch_head_fwd(last.out.p, last.P, 64, t->param[80], t->param[81], logits, s));
  pl->eval_act17_missing = !training && head_done;
  pl->forward_done = training != 0;
  return 0;
}

// Backward stage list: 0 = head, then convs / up-convs in reverse execution order.
//   kind 0: head, 1: conv (idx), 2: up (idx)
static void stage_decode(int stage, int* kind, int* idx) {
  if (stage == 0) { *kind = 0; *idx = 0; return; }
  int s = stage - 1;
  if (s < 12) {                       // decoder: (dconvL.3, dconvL.0, upconvL) for L = 1..4
    const int grp = s / 3, w = s % 3; // grp 0 -> L = 1 (k = 3), grp 3 -> L = 4 (k = 0)
    const int k = 3 - grp;
    if (w == 0) { *kind = 1; *idx = 11 + 2 * k; }
    else if (w == 1) { *kind = 1; *idx = 10 + 2 * k; }
    else { *kind = 2; *idx = k; }
    return;
  }
  *kind = 1;
  *idx = 9 - (s - 12);
}

int cs_unet_stage_params(int stage, int* out_indices, int capacity) {
  if (stage < 0 || stage >= CS_UNET_NUM_BWD_STAGES) return fail("stage %d out of range", stage);
  int kind, idx, n = 0, tmp[6];
  stage_decode(stage, &kind, &idx);
  // the head's parameter gradients are produced by the BN-backward reduction of dconv1.3 (stage 1) unless the
  // stand-alone head kernels are selected (CARTSEG_FUSE_HEAD=0)
  if (kind == 0) { if (!fuse_head()) { tmp[0] = 80; tmp[1] = 81; n = 2; } }
  else if (kind == 1) {
    const int pb = conv_param_base(idx);
    if (idx == 17 && fuse_head()) { tmp[n++] = 80; tmp[n++] = 81; }
    for (int j = 0; j < 4; ++j) tmp[n++] = pb + j;
  }
  else { tmp[0] = 40 + 2 * idx; tmp[1] = 41 + 2 * idx; n = 2; }
  for (int j = 0; j < n && j < capacity; ++j) out_indices[j] = tmp[j];
  return n;
}

// Stream layout of the backward pass.  The chain BN-backward -> dgrad -> BN-backward of the next layer is the
// critical path; the weight-gradient GEMMs hang off it (they only produce parameter gradients).  They are therefore
// issued on a second, lower-priority stream: their tensor-core work overlaps the HBM-bound BN-backward passes of the
// following layers, and whenever both a dgrad and a wgrad are runnable the block scheduler serves the dgrad first.
// Both internal streams are forked from / joined back into the caller's stream with events, so the call stays
// asynchronous and stream-ordered for the caller (and capturable in a CUDA graph).  CARTSEG_OVERLAP=0 disables it.
static bool use_overlap() {
  static const bool v = [] {
    const char* e = getenv("CARTSEG_OVERLAP");
    return !(e && e[0] == '0');
  }();
  return v;
}

static int ensure_streams(cs_unet_plan* pl) {
  if (pl->s_hi) return 0;
  int lo = 0, hi = 0;
  CS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least priority (numerically greatest)
  CS_CUDA(cudaStreamCreateWithPriority(&pl->s_hi, cudaStreamNonBlocking, hi));
  CS_CUDA(cudaStreamCreateWithPriority(&pl->s_lo, cudaStreamNonBlocking, lo));
  CS_CUDA(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
  CS_CUDA(cudaEventCreateWithFlags(&pl->ev_hi, cudaEventDisableTiming));
  CS_CUDA(cudaEventCreateWithFlags(&pl->ev_lo, cudaEventDisableTiming));
  CS_CUDA(cudaEventCreateWithFlags(&pl->ev_flush, cudaEventDisableTiming));
  for (int i = 0; i < CS_UNET_NUM_BWD_STAGES; ++i)
    CS_CUDA(cudaEventCreateWithFlags(&pl->ev_stage[i], cudaEventDisableTiming));
  return 0;
}

// Which weight-gradient GEMMs are held back, and until when.  Two one-CTA-per-SM persistent GEMM kernels cannot share an
// SM, so a tensor-bound wgrad issued next to the tensor-bound dgrads of the deep levels only time-slices with them — and
// because CTAs are not preemptible, every dgrad that becomes ready while a wgrad wave is resident waits for it (the
// round-1 timeline shows the conv-transpose dgrads taking 0.23-0.67 ms instead of 0.07-0.15 ms).  The HBM-bound
// BatchNorm-backward passes of the high-resolution encoder levels at the END of the backward pass, on the other hand,
// do share SMs with a resident wgrad CTA (no shared memory, <= 104 registers) and leave the tensor pipe idle.  So the
// deep, HBM-light wgrads (levels 3-5: conv 4..13, upconv4/3) CAN be enqueued only when the main stream reaches conv
// `flush_conv` (CARTSEG_DEFER_WGRAD=1, CARTSEG_DEFER_FLUSH_CONV=n).  Measured on B200 (round 2, k2,
// profiles/r2_midround_k2_backward_timeline_deferred.txt): the main stream then finishes at 10.1 ms instead of 11.4 ms — every
// conv-transpose dgrad runs in 0.07 ms — but the held wgrads run 1.6-2.5x slower next to the BatchNorm passes they were
// meant to hide under (one issuing warp against 16 memory-bound warps per SM) and nine of them are still queued when the
// main stream ends: 11.60 ms either way, step 18.08 vs 17.93 ms.  Default OFF.
static bool defer_enabled() { static const int v = env_int("CARTSEG_DEFER_WGRAD", 0); return v != 0; }
static int defer_flush_conv() { static const int v = env_int("CARTSEG_DEFER_FLUSH_CONV", 5); return v; }
static bool held_conv(int idx) { return defer_enabled() && idx >= 4 && idx <= 13 && idx > defer_flush_conv(); }
static bool held_up(int k) { return defer_enabled() && k <= 1; }
// Priority inversion at the conv-transposes: dgrad(dconvL.0) is followed at once by the short conv-transpose dgrad, but the
// one-wave wgrad of dconvL.0 (0.3-0.6 ms, not preemptible) takes the SMs as the conv dgrad drains, and the conv-transpose
// dgrad waits for it (round-1 timeline: 0.23-0.67 ms instead of 0.07-0.15 ms).  With CARTSEG_WGRAD_AFTER_UP=1 (default)
// those four wgrads, and the conv-transpose's own wgrad / bias sum, are enqueued behind the conv-transpose dgrad: they
// start when the next BatchNorm-backward passes do.
static bool after_up_enabled() { static const int v = env_int("CARTSEG_WGRAD_AFTER_UP", 1); return v != 0; }
static bool late_conv(int idx) { return after_up_enabled() && idx >= 10 && idx % 2 == 0; }

static int launch_conv_wgrad(cs_unet_plan* pl, int idx, float* gw, cudaStream_t sw) {
  ConvL& c = pl->conv[idx];
  const size_t e = (size_t)(idx == 0 ? 1 : 9) * c.cout * c.cin;
  CS_CUDA(cudaMemsetAsync(pl->dwp, 0, e * sizeof(float), sw));
  CS_CUDA(traced(pl, 400 + idx, sw, [&] { return timed(pl, wgrad_class(c.wg, c.bn_w), c.flops, sw, [&] { return launch_wgrad_gemm(c.wg, c.bn_w, sw); }); }));
  if (idx == 0) CS_CUDA(launch_unpack_first(pl->dwp, c.cout, pl->Cin, gw, sw));
  else CS_CUDA(launch_unpack_pairs(pl->dwp, c.cout, c.cin, 9, kTapWgrad, gw, sw));
  return 0;
}
static int launch_up_wgrad(cs_unet_plan* pl, int idx, float* gw, float* gb, cudaStream_t sw) {
  UpL& u = pl->up[idx];
  if (gb) {
    ConvL& d0 = pl->conv[10 + 2 * idx];                   // dconvL.0, whose dgrad wrote u.g_out
    if (d0.dg.stat_sum) CS_CUDA(traced(pl, 900 + idx, sw, [&] { return launch_stat_to_bias(d0.dg_sum, d0.dg_sq, d0.cin, u.cout, gb, sw); }));
    else CS_CUDA(traced(pl, 900 + idx, sw, [&] { return launch_channel_sum(u.g_out.p, u.g_out.pitch, u.g_out.c0, 4 * u.P, u.cout, gb, sw); }));
  }
  if (gw) {
    CS_CUDA(cudaMemsetAsync(pl->dwp, 0, (size_t)4 * u.cin * u.cout * sizeof(float), sw));
    CS_CUDA(traced(pl, 800 + idx, sw, [&] { return timed(pl, wgrad_class(u.bn_w), u.flops, sw, [&] { return launch_wgrad_gemm(u.wg, u.bn_w, sw); }); }));
    CS_CUDA(launch_unpack_pairs(pl->dwp, u.cin, u.cout, 4, kTapIdent, gw, sw));
  }
  return 0;
}
// Enqueue everything that was held back: the side stream first waits for the main stream's current position.
static int flush_held(cs_unet_plan* pl, cudaStream_t s, cudaStream_t sw, bool overlap) {
  if (pl->held.empty()) return 0;
  if (overlap) {
    CS_CUDA(cudaEventRecord(pl->ev_flush, s));
    CS_CUDA(cudaStreamWaitEvent(sw, pl->ev_flush, 0));
  }
  for (const cs_unet_plan::HeldWgrad& h : pl->held) {
    if (h.kind == 1) CS_TRY(launch_conv_wgrad(pl, h.idx, h.gw, sw));
    else CS_TRY(launch_up_wgrad(pl, h.idx, h.gw, h.gb, sw));
  }
  pl->held.clear();
  return 0;
}

// The side stream waits for the main stream's current position, then runs the wgrads in `late` (always on two streams).
static int flush_late(cs_unet_plan* pl, cudaStream_t s, cudaStream_t sw, cudaEvent_t ev) {
  CS_CUDA(cudaEventRecord(ev, s));
  CS_CUDA(cudaStreamWaitEvent(sw, ev, 0));
  for (const cs_unet_plan::HeldWgrad& h : pl->late) CS_TRY(launch_conv_wgrad(pl, h.idx, h.gw, sw));
  pl->late.clear();
  return 0;
}

int cs_unet_backward_held_stages(int* flush_stage, int* stages, int capacity) {
  // stages whose weight gradients are final only after stage `*flush_stage` has been enqueued (or after the last stage)
  int n = 0;
  *flush_stage = CS_UNET_NUM_BWD_STAGES - 1;
  for (int st = 0; st < CS_UNET_NUM_BWD_STAGES; ++st) {
    int kind, idx;
    stage_decode(st, &kind, &idx);
    if ((kind == 1 && held_conv(idx)) || (kind == 2 && held_up(idx))) {
      if (n < capacity) stages[n] = st;
      ++n;
    }
    if (kind == 1 && idx == defer_flush_conv()) *flush_stage = st;
  }
  return n;
}

int cs_unet_backward(cs_unet_plan* pl, const cs_unet_tensors* t, const float* dlogits, int stage_begin, int stage_end,
                     int frozen_encoder_convs, cs_stream_t stream) {
  if (!pl || !pl->bound) return fail("plan is not bound to a workspace");
  if (!pl->forward_done) return fail("cs_unet_backward needs a preceding training-mode cs_unet_forward");
  if (stage_begin < 0 || stage_end > CS_UNET_NUM_BWD_STAGES || stage_begin > stage_end)
    return fail("bad stage range [%d, %d)", stage_begin, stage_end);
  if (frozen_encoder_convs < 0 || frozen_encoder_convs > 10) return fail("frozen_encoder_convs must be in [0, 10]");
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  cudaStream_t s = caller, sw = caller;      // s: critical path, sw: weight gradients
  const bool overlap = use_overlap() && !pl->no_overlap;
  if (overlap) {
    CS_TRY(ensure_streams(pl));
    s = pl->s_hi;
    sw = pl->s_lo;
    CS_CUDA(cudaEventRecord(pl->ev_fork, caller));
    CS_CUDA(cudaStreamWaitEvent(s, pl->ev_fork, 0));
    CS_CUDA(cudaStreamWaitEvent(sw, pl->ev_fork, 0));
  }
  const int B = pl->B;
  for (int stage = stage_begin; stage < stage_end; ++stage) {
    int kind, idx;
    stage_decode(stage, &kind, &idx);
    if (kind == 0) {
      if (!dlogits) return fail("dlogits is null");
      const ConvL& last = pl->conv[17];
      // The head's input gradient dlogits * w is not materialised: the BN backward of the last layer forms it on the
      // fly (BnBwdArgs::head_dlogits).  What is left of the head backward only produces parameter gradients, so it
      // goes to the weight-gradient stream, off the critical path.
      if (!t->param[80]) return fail("final_conv.weight is null");
      if (!fuse_head()) {
        CS_CUDA(launch_head_bwd(last.out.p, dlogits, last.P, 64, t->param[80], last.g_out.p, t->grad[80], t->grad[81], s));
        continue;
      }
      // The head's parameter gradients come out of the BN-backward reduction of the last layer (stage 1), which
      // recomputes the head's input from y: nothing to launch here.
      if (overlap) {
        CS_CUDA(cudaEventRecord(pl->ev_stage[stage], s));
        CS_CUDA(cudaStreamWaitEvent(sw, pl->ev_stage[stage], 0));
      }
      CS_CUDA(cudaMemcpyAsync(pl->dlogits_keep, dlogits, (size_t)last.P * sizeof(float), cudaMemcpyDeviceToDevice, sw));
      pl->head_w_keep = t->param[80];
    } else if (kind == 1) {
      ConvL& c = pl->conv[idx];
      if (idx == defer_flush_conv()) CS_TRY(flush_held(pl, s, sw, overlap));
      // The im2col matrix of the input image, for the first convolution's weight gradient (last kernel of the pass), is
      // written on the side stream while the main stream is in the bottleneck level: its 0.4 GB of traffic then runs next
      // to tensor-bound, HBM-light kernels.  (Issued at the start of the pass it took 1.0 ms at low priority next to the
      // HBM-bound BatchNorm passes of the full-resolution decoder level and slowed those.)
      if (idx == 9 && pl->x_keep && frozen_encoder_convs == 0 && t->grad[pl->conv[0].pw])
        CS_CUDA(traced(pl, 500, sw, [&] { return launch_im2col_first(pl->x_keep, B, pl->Cin, pl->H, pl->W, pl->col, sw); }));
      if (idx < frozen_encoder_convs) continue;          // nothing below a frozen prefix needs gradients
      BnBwdArgs a{};
      a.g = c.g_out.p; a.g_pitch = c.g_out.pitch; a.g_c0 = c.g_out.c0;
      a.g_pool = c.g_pool; a.y = c.y;
      if (idx == 17 && fuse_head()) {
        if (!dlogits) return fail("dlogits is null");
        a.head_dlogits = dlogits; a.head_w = t->param[80];
        a.head_grad_w = t->grad[80]; a.head_grad_b = t->grad[81];
      }
      a.scale = c.scale; a.shift = c.shift; a.mean = c.mean; a.invstd = c.invstd;
      a.partial = pl->bn_partial; a.c1 = c.bc1; a.c2 = c.bc2; a.dy = c.dy;
      a.grad_gamma = t->grad[c.pgamma]; a.grad_beta = t->grad[c.pbeta]; a.grad_conv_bias = t->grad[c.pb];
      a.B = B; a.H = c.H; a.W = c.W; a.C = c.cout;
      CS_CUDA(traced(pl, 100 + idx, s, [&] { return launch_bn_bwd_reduce(a, s); }));
      CS_CUDA(traced(pl, 200 + idx, s, [&] { return launch_bn_bwd_apply(a, s); }));
      if (t->grad[c.pw]) {
        if (overlap && held_conv(idx)) {
          pl->held.push_back(cs_unet_plan::HeldWgrad{1, idx, stage, t->grad[c.pw], nullptr});
        } else if (overlap && late_conv(idx)) {
          pl->late.push_back(cs_unet_plan::HeldWgrad{1, idx, stage, t->grad[c.pw], nullptr});
        } else {
          if (overlap) {                                  // dy is final: the wgrad may start on the side stream
            CS_CUDA(cudaEventRecord(pl->ev_stage[stage], s));
            CS_CUDA(cudaStreamWaitEvent(sw, pl->ev_stage[stage], 0));
          }
          CS_TRY(launch_conv_wgrad(pl, idx, t->grad[c.pw], sw));
        }
      }
      // Decoder convs always run their dgrad: it is the only writer of the concat-gradient buffer that the conv-transpose
      // weight / bias gradients read.  An encoder conv needs it only if something below it still trains.
      if (c.dg_sum) {                                     // dconvL.0: its dgrad also sums the up-sampled half per channel
        const bool want_bias = upbias_fused() && t->grad[pl->up[(idx - 10) / 2].pb] != nullptr;
        c.dg.stat_sum = want_bias ? c.dg_sum : nullptr;
        c.dg.stat_sq = want_bias ? c.dg_sq : nullptr;
      }
      if (idx >= 10 || (idx > 0 && idx > frozen_encoder_convs))
        CS_CUDA(traced(pl, 300 + idx, s, [&] { return timed(pl, conv_class(c.dg, c.bn_d), c.flops, s, [&] { return launch_pix_gemm(c.dg, c.bn_d, pl->num_sms, s); }); }));
    } else {
      UpL& u = pl->up[idx];
      const bool want = t->grad[u.pb] || t->grad[u.pw];
      const bool after = overlap && after_up_enabled();   // side-stream work of this stage goes behind the dgrad
      if (want && overlap && held_up(idx)) {
        pl->held.push_back(cs_unet_plan::HeldWgrad{2, idx, stage, t->grad[u.pw], t->grad[u.pb]});
      } else if (want && !after) {
        if (overlap) {                                    // g_out of the conv-transpose is final on the main stream
          CS_CUDA(cudaEventRecord(pl->ev_stage[stage], s));
          CS_CUDA(cudaStreamWaitEvent(sw, pl->ev_stage[stage], 0));
        }
        CS_TRY(launch_up_wgrad(pl, idx, t->grad[u.pw], t->grad[u.pb], sw));
      }
      // upconv4's input gradient only feeds conv5.3: nothing reads it when the whole encoder is frozen
      if (!(idx == 0 && frozen_encoder_convs >= 10))
        CS_CUDA(traced(pl, 700 + idx, s, [&] { return timed(pl, pix_class(u.bn_d), u.flops, s, [&] { return launch_pix_gemm(u.dg, u.bn_d, pl->num_sms, s); }); }));
      if (after) {
        CS_TRY(flush_late(pl, s, sw, pl->ev_stage[stage]));
        if (want && !held_up(idx)) CS_TRY(launch_up_wgrad(pl, idx, t->grad[u.pw], t->grad[u.pb], sw));
      }
    }
  }
  if (!pl->late.empty()) CS_TRY(flush_late(pl, s, sw, pl->ev_flush));   // stage range ended between dconvL.0 and its up-conv
  if (stage_end == CS_UNET_NUM_BWD_STAGES) CS_TRY(flush_held(pl, s, sw, overlap));
  if (overlap) {                                          // join: the caller's stream continues after both
    CS_CUDA(cudaEventRecord(pl->ev_hi, s));
    CS_CUDA(cudaEventRecord(pl->ev_lo, sw));
    pl->join_pending = true;
    if (!pl->deferred_join) CS_TRY(cs_unet_backward_wait(pl, stream));
  }
  return 0;
}

int cs_unet_plan_set_sm_limit(cs_unet_plan* pl, int sms) {
  if (!pl || !pl->bound) return fail("plan is not bound to a workspace");
  int all = 0;
  CS_TRY(device_sm_count(&all));
  if (sms <= 0 || sms > all) sms = all;
  pl->num_sms = sms & ~1;                                 // CTA pairs
  if (pl->num_sms < 2) pl->num_sms = 2;
  if (!pl->infer) {                                       // the single-wave weight-gradient grids follow the limit too
    auto refit = [&](WgradParams& w) {
      w.splits = choose_wgrad_splits(w.nine ? w.n_blocks : w.m_blocks * w.n_blocks * w.G, w.tiles_w * w.tiles_h * w.batch,
                                     pl->num_sms);
    };
    for (int i = 0; i < 18; ++i) refit(pl->conv[i].wg);
    for (int k = 0; k < 4; ++k) refit(pl->up[k].wg);
  }
  return 0;
}

int cs_unet_set_deferred_join(cs_unet_plan* pl, int enable) {
  if (!pl) return fail("plan is null");
  pl->deferred_join = enable != 0;
  return 0;
}

int cs_unet_backward_wait(cs_unet_plan* pl, cs_stream_t stream) {
  if (!pl) return fail("plan is null");
  if (!pl->join_pending) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CS_CUDA(cudaStreamWaitEvent(st, pl->ev_hi, 0));
  CS_CUDA(cudaStreamWaitEvent(st, pl->ev_lo, 0));
  return 0;
}

int cs_unet_profile(cs_unet_plan* pl, int enable) {
  if (!pl) return fail("plan is null");
  pl->profiling = enable != 0;
  pl->prof_used = 0;
  return 0;
}

int cs_unet_set_overlap(cs_unet_plan* pl, int enable) {
  if (!pl) return fail("plan is null");
  pl->no_overlap = enable == 0;
  return 0;
}

int cs_unet_trace(cs_unet_plan* pl, int enable) {
  if (!pl) return fail("plan is null");
  pl->tracing = enable != 0;
  pl->trace_used = 0;
  return 0;
}

int cs_unet_trace_read(cs_unet_plan* pl, int capacity, int* labels, double* begin_ms, double* end_ms) {
  if (!pl || !labels || !begin_ms || !end_ms) return fail("cs_unet_trace_read: null pointer");
  const size_t n = pl->trace_used < (size_t)capacity ? pl->trace_used : (size_t)capacity;
  for (size_t i = 0; i < n; ++i) {
    CS_CUDA(cudaEventSynchronize(pl->trace_events[2 * i + 1]));
    float t0 = 0.f, t1 = 0.f;
    CS_CUDA(cudaEventElapsedTime(&t0, pl->trace_events[0], pl->trace_events[2 * i]));
    CS_CUDA(cudaEventElapsedTime(&t1, pl->trace_events[0], pl->trace_events[2 * i + 1]));
    labels[i] = pl->trace_label[i];
    begin_ms[i] = t0;
    end_ms[i] = t1;
  }
  pl->trace_used = 0;
  return (int)n;
}

int cs_unet_profile_read(cs_unet_plan* pl, int n_classes, double* ms, double* flops, long long* launches) {
  if (!pl) return fail("plan is null");
  if (n_classes != kNumCls || !ms || !flops || !launches) return fail("cs_unet_profile_read: expected %d classes", (int)kNumCls);
  for (int i = 0; i < kNumCls; ++i) { ms[i] = 0.0; flops[i] = 0.0; launches[i] = 0; }
  for (size_t i = 0; i < pl->prof_used; ++i) {
    CS_CUDA(cudaEventSynchronize(pl->prof_events[2 * i + 1]));
    float t = 0.f;
    CS_CUDA(cudaEventElapsedTime(&t, pl->prof_events[2 * i], pl->prof_events[2 * i + 1]));
    const int c = pl->prof_class[i];
    ms[c] += t;
    flops[c] += pl->prof_flops[i];
    launches[c] += 1;
  }
  pl->prof_used = 0;
  return 0;
}

// Debug read-back of an internal NHWC bf16 tensor as dense fp32 NCHW (tests only).
//   kind 0: conv raw output y   1: conv activation (post BN+ReLU)   2: grad wrt y   3: grad wrt activation
//        4: pooled activation   5: grad wrt pooled   (index = conv 0..17)
//        6: conv-transpose output   7: grad wrt it   (index = up 0..3)
int cs_unet_debug_read(cs_unet_plan* pl, int kind, int index, int dims_out[4], float* dst, cs_stream_t stream) {
  if (!pl || !pl->bound) return fail("plan is not bound to a workspace");
  View v;
  int C, H, W;
  if (kind >= 0 && kind <= 5) {
    if (index < 0 || index >= 18) return fail("conv index out of range");
    const ConvL& c = pl->conv[index];
    C = c.cout; H = c.H; W = c.W;
    switch (kind) {
      case 0: v = View{c.y, c.cout, 0}; break;
      case 1: v = c.out; break;
      case 2: v = View{c.dy, c.cout, 0}; break;
      case 3: v = c.g_out; break;
      case 4: v = View{c.pooled, c.cout, 0}; H /= 2; W /= 2; break;
      default: v = View{c.g_pool, c.cout, 0}; H /= 2; W /= 2; break;
    }
  } else if (kind == 6 || kind == 7) {
    if (index < 0 || index >= 4) return fail("up index out of range");
    const UpL& u = pl->up[index];
    C = u.cout; H = 2 * u.H; W = 2 * u.W;
    v = kind == 6 ? u.out : u.g_out;
  } else {
    return fail("unknown debug tensor kind %d", kind);
  }
  if (dims_out) { dims_out[0] = pl->B; dims_out[1] = C; dims_out[2] = H; dims_out[3] = W; }
  if (!dst) return 0;
  if (kind == 1 && index == 17 && pl->eval_act17_missing)
    return fail("the last eval-mode forward did not store dconv1.3's activation (the head runs in that layer's epilogue)");
  if (!v.p) return fail("tensor (kind %d, index %d) does not exist in this plan", kind, index);
  if (kind == 1 && !pl->infer && pl->forward_done && ((index == 17 && fuse_head()) || pl->conv[index].act_fused)) {
    // the training path never stores this activation (the head / the next convolution consume y directly): rebuild it
    // from y and the coefficients the forward published
    const ConvL& c = pl->conv[index];
    CS_CUDA(launch_bn_apply_relu(c.y, c.P, c.cout, c.scale, c.shift, c.out.p, static_cast<cudaStream_t>(stream)));
  }
  if (kind == 3 && index == 17 && fuse_head()) {
    // never written by the training path: materialise dlogits * w from the copies kept by the last backward
    if (!pl->head_w_keep) return fail("no backward pass has run on this plan yet");
    CS_CUDA(launch_head_grad_act(pl->dlogits_keep, pl->conv[17].P, 64, pl->head_w_keep, v.p, static_cast<cudaStream_t>(stream)));
  }
  CS_CUDA(launch_nhwc_to_nchw_f32(v.p, v.pitch, v.c0, pl->B, H, W, C, dst, static_cast<cudaStream_t>(stream)));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// SDF / losses / thresholds
// ---------------------------------------------------------------------------------------------
size_t cs_sdf_scratch_bytes(int batch, int height, int width) { return sdf_scratch_bytes(batch, height, width); }

int cs_sdf(const float* src, float thr, int ge, int batch, int height, int width, float norm, float* sdf, void* scratch,
           cs_stream_t stream) {
  if (!src || !sdf || !scratch) return fail("cs_sdf: null pointer");
  if (batch < 1 || height < 1 || width < 1) return fail("cs_sdf: empty input");
  if (!(norm > 0.f)) return fail("cs_sdf: norm must be positive");
  CS_CUDA(launch_sdf(src, thr, ge, batch, height, width, norm, sdf, scratch, static_cast<cudaStream_t>(stream)));
  return 0;
}

size_t cs_loss_scratch_bytes(int rows) { return loss_scratch_bytes(rows); }

static int loss_args(const cs_loss_desc* d, LossArgs* a) {
  if (!d) return fail("loss descriptor is null");
  if (d->rows < 1 || d->n < 4 || d->n % 4) return fail("loss: rows >= 1 and n a positive multiple of 4 required");
  memset(a, 0, sizeof(*a));
  a->rows = d->rows; a->n = d->n;
  a->w_elem = d->w_elem; a->alpha = d->alpha; a->gamma = d->gamma; a->elem_sum = d->elem_sum;
  a->w_dice = d->w_dice; a->smooth = d->smooth;
  a->w_bgt = d->w_bgt; a->w_bpred = d->w_bpred; a->use_abs = d->use_abs; a->per_row = d->per_row;
  return 0;
}

int cs_loss_forward(const cs_loss_desc* d, const float* logits, const float* targets, const float* sdf_gt,
                    const float* sdf_pred, void* scratch, float* loss_out, cs_stream_t stream) {
  LossArgs a;
  CS_TRY(loss_args(d, &a));
  if (!logits || !targets || !scratch || !loss_out) return fail("cs_loss_forward: null pointer");
  a.logits = logits; a.targets = targets; a.sdf_gt = sdf_gt; a.sdf_pred = sdf_pred;
  a.stats = static_cast<double*>(scratch); a.loss_out = loss_out;
  CS_CUDA(launch_loss_forward(a, static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_loss_backward(const cs_loss_desc* d, const float* logits, const float* targets, const float* sdf_gt,
                     const float* sdf_pred, const void* scratch, const float* grad_out, float* dlogits,
                     cs_stream_t stream) {
  LossArgs a;
  CS_TRY(loss_args(d, &a));
  if (!logits || !targets || !scratch || !dlogits) return fail("cs_loss_backward: null pointer");
  a.logits = logits; a.targets = targets; a.sdf_gt = sdf_gt; a.sdf_pred = sdf_pred;
  a.stats = const_cast<double*>(static_cast<const double*>(scratch)); a.grad_out = grad_out; a.dlogits = dlogits;
  CS_CUDA(launch_loss_backward(a, static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_focal_map_forward(const float* logits, const float* targets, long long n, float alpha, float gamma, float* out,
                         cs_stream_t stream) {
  if (!logits || !targets || !out) return fail("cs_focal_map_forward: null pointer");
  if (n < 1) return fail("cs_focal_map_forward: empty input");
  CS_CUDA(launch_focal_map_forward(logits, targets, n, alpha, gamma, out, static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_focal_map_backward(const float* logits, const float* targets, const float* grad_out, long long n, float alpha,
                          float gamma, float* dlogits, cs_stream_t stream) {
  if (!logits || !targets || !grad_out || !dlogits) return fail("cs_focal_map_backward: null pointer");
  if (n < 1) return fail("cs_focal_map_backward: empty input");
  CS_CUDA(launch_focal_map_backward(logits, targets, grad_out, n, alpha, gamma, dlogits, static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_threshold_stats(const float* logits, const float* targets, int rows, long long n, const float* xs, int K,
                       double* counts, double* soft, cs_stream_t stream) {
  if (!logits || !targets || !xs || !counts || !soft) return fail("cs_threshold_stats: null pointer");
  if (rows < 1 || n < 1) return fail("cs_threshold_stats: empty input");
  if (K < 1 || K > 32) return fail("cs_threshold_stats: 1 <= K <= 32 thresholds per call");
  CS_CUDA(launch_threshold_stats(logits, targets, rows, n, xs, K, counts, soft, static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_threshold_mask(const float* logits, long long n, float xstar, uint8_t* mask, cs_stream_t stream) {
  if (!logits || !mask) return fail("cs_threshold_mask: null pointer");
  if (n < 4 || n % 4) return fail("cs_threshold_mask: n must be a positive multiple of 4");
  CS_CUDA(launch_threshold_mask(logits, n, xstar, mask, static_cast<cudaStream_t>(stream)));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Active Boundary Loss
// ---------------------------------------------------------------------------------------------
size_t cs_abl_scratch_bytes(int batch, int height, int width) { return abl_scratch_bytes(batch, height, width); }

static int abl_check(const cs_abl_desc* d) {
  if (!d) return fail("abl descriptor is null");
  if (d->batch < 1 || d->height < 2 || d->width < 2) return fail("abl: batch >= 1 and H, W >= 2 required");
  if (!(d->max_clip_dist > 0.f)) return fail("abl: max_clip_dist must be positive");
  if (d->label_smoothing < 0.f || d->label_smoothing >= 1.f) return fail("abl: label_smoothing must be in [0, 1)");
  return 0;
}

int cs_abl_forward(const cs_abl_desc* d, const float* logits, const float* targets, void* scratch, float* loss_out,
                   cs_stream_t stream) {
  CS_TRY(abl_check(d));
  if (!logits || !targets || !scratch || !loss_out) return fail("cs_abl_forward: null pointer");
  if ((uintptr_t)scratch & 255) return fail("cs_abl_forward: scratch must be 256-byte aligned");
  AblLadder lad;
  for (int i = 0; i < kAblLadder; ++i) lad.v[i] = d->eps_ladder[i];
  for (int i = 1; i < kAblLadder; ++i)
    if (!(lad.v[i] > lad.v[i - 1])) return fail("abl: eps_ladder must be strictly increasing");
  CS_CUDA(launch_abl_forward(logits, targets, d->batch, d->height, d->width, lad, d->max_n, d->label_smoothing,
                             d->max_clip_dist, d->ignore_label, d->per_image_maps ? 0 : 1, scratch, loss_out,
                             static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_abl_backward(const cs_abl_desc* d, const float* logits, const void* scratch, const float* grad_out,
                    float* dlogits, cs_stream_t stream) {
  CS_TRY(abl_check(d));
  if (!logits || !scratch || !dlogits) return fail("cs_abl_backward: null pointer");
  CS_CUDA(launch_abl_backward(logits, d->batch, d->height, d->width, d->label_smoothing, d->max_clip_dist, scratch,
                              grad_out, dlogits, static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_abl_debug_read(const cs_abl_desc* d, const void* scratch, float* eps, int* ladder_index, unsigned long long* kept,
                      unsigned long long* pred_boundary, uint16_t* dist_map_host, float* kl_map_host,
                      cs_stream_t stream) {
  CS_TRY(abl_check(d));
  if (!scratch) return fail("cs_abl_debug_read: null pointer");
  CS_CUDA(abl_debug_read(scratch, d->batch, d->height, d->width, eps, ladder_index, kept, pred_boundary, dist_map_host,
                         kl_map_host, static_cast<cudaStream_t>(stream)));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Pseudo-label post-processing
// ---------------------------------------------------------------------------------------------
int cs_ensemble_accumulate(const float* logits, float weight, long long n, int first, float* probs, cs_stream_t stream) {
  if (!logits || !probs) return fail("cs_ensemble_accumulate: null pointer");
  if (n < 1) return fail("cs_ensemble_accumulate: empty input");
  CS_CUDA(launch_ensemble_accumulate(logits, weight, n, first, probs, static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_pseudo_qc(const float* probs, int batch, long long n, float threshold, int mask_value, uint8_t* mask,
                 double* stats, cs_stream_t stream) {
  if (!probs || !stats) return fail("cs_pseudo_qc: null pointer");
  if (batch < 1 || n < 1) return fail("cs_pseudo_qc: empty input");
  if (mask_value < 1 || mask_value > 255) return fail("cs_pseudo_qc: mask_value must be in 1..255");
  CS_CUDA(launch_pseudo_qc(probs, batch, n, threshold, mask_value, mask, stats, static_cast<cudaStream_t>(stream)));
  return 0;
}

size_t cs_mask_cleanup_scratch_bytes(int batch, int height, int width) {
  return mask_cleanup_scratch_bytes(batch, height, width);
}

int cs_mask_cleanup(const uint8_t* mask, int batch, int height, int width, int bin_threshold, int fill_holes,
                    int keep_largest, uint8_t* out, void* scratch, cs_stream_t stream) {
  if (!mask || !out || !scratch) return fail("cs_mask_cleanup: null pointer");
  if (batch < 1 || height < 1 || width < 1) return fail("cs_mask_cleanup: empty input");
  if ((long long)batch * height * width >= 0x7fffffffLL) return fail("cs_mask_cleanup: more than 2^31 pixels per call");
  if ((uintptr_t)scratch & 255) return fail("cs_mask_cleanup: scratch must be 256-byte aligned");
  CS_CUDA(launch_mask_cleanup(mask, batch, height, width, bin_threshold, fill_holes, keep_largest, out, scratch,
                              static_cast<cudaStream_t>(stream)));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Input side
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(cs_image_desc) == sizeof(ImageDesc), "cs_image_desc and ImageDesc must have the same layout");

int cs_letterbox_geometry(int height, int width, double side_padding_ratio, int* canvas, int* x0, int* y0) {
  if (height < 1 || width < 1 || !canvas || !x0 || !y0) return fail("cs_letterbox_geometry: bad arguments");
  const int side = (int)nearbyint((double)width * side_padding_ratio);   // Python round(): half to even
  const int pw = width + 2 * side;
  const int L = pw > height ? pw : height;
  *canvas = L;
  *x0 = (L - pw) / 2 + side;
  *y0 = (L - height) / 2;
  return 0;
}

int cs_preproc_images(const cs_image_desc* descs_device, int batch, int out_size, const float mean[3],
                      const float std_[3], int bgr, float* out_nchw, cs_stream_t stream) {
  if (!descs_device || !mean || !std_ || !out_nchw) return fail("cs_preproc_images: null pointer");
  if (batch < 1 || batch > 65535 || out_size < 1 || out_size > 65535) return fail("cs_preproc_images: bad batch / size");
  Norm3 n;
  for (int c = 0; c < 3; ++c) {
    if (!(std_[c] > 0.f)) return fail("cs_preproc_images: std must be positive");
    n.mean255[c] = (float)((double)mean[c] * 255.0);
    n.inv_std255[c] = (float)(1.0 / ((double)std_[c] * 255.0));
  }
  CS_CUDA(launch_preproc_images(reinterpret_cast<const ImageDesc*>(descs_device), batch, out_size, n, bgr, out_nchw,
                                static_cast<cudaStream_t>(stream)));
  return 0;
}

int cs_preproc_masks(const cs_image_desc* descs_device, int batch, int out_size, float* out, cs_stream_t stream) {
  if (!descs_device || !out) return fail("cs_preproc_masks: null pointer");
  if (batch < 1 || batch > 65535 || out_size < 1 || out_size > 65535) return fail("cs_preproc_masks: bad batch / size");
  CS_CUDA(launch_preproc_masks(reinterpret_cast<const ImageDesc*>(descs_device), batch, out_size, out,
                               static_cast<cudaStream_t>(stream)));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Single-layer entry points (tests / micro-benchmarks).  scratch = [fprop pack | dgrad pack | fp32 dW pack]
// ---------------------------------------------------------------------------------------------
size_t cs_layer_scratch_bytes(int cin, int cout) {
  const size_t e = (size_t)9 * cin * cout;
  return 2 * ((e * 2 + 1023) & ~(size_t)1023) + ((e * 4 + 1023) & ~(size_t)1023);
}

namespace {
struct LayerScratch {
  bf16 *wf, *wd;
  float* dwp;
};
LayerScratch carve(void* scratch, int cin, int cout) {
  const size_t e = (size_t)9 * cin * cout, hb = (e * 2 + 1023) & ~(size_t)1023;
  uint8_t* b = static_cast<uint8_t*>(scratch);
  return LayerScratch{reinterpret_cast<bf16*>(b), reinterpret_cast<bf16*>(b + hb), reinterpret_cast<float*>(b + 2 * hb)};
}
int check_layer(const void* a, const void* b, const void* scratch, int batch, int h, int w, int cin, int cout) {
  if (!a || !b || !scratch) return fail("layer op: null pointer");
  if ((uintptr_t)scratch & 1023) return fail("layer op: scratch must be 1024-byte aligned");
  if (batch < 1 || h < 1 || w < 1) return fail("layer op: empty input");
  if (cin % 64 || cout % 64 || cin < 64 || cout < 64) return fail("layer op: channels must be multiples of 64");
  return 0;
}
}  // namespace

int cs_conv3x3_fprop(const void* x, int batch, int height, int width, int cin, const float* w_oihw, int cout, void* y,
                     double* stat_sum, double* stat_sq, void* scratch, cs_stream_t stream) {
  CS_TRY(check_layer(x, y, scratch, batch, height, width, cin, cout));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int sms, bn;
  CS_TRY(layer_sm_count(&sms));
  LayerScratch ls = carve(scratch, cin, cout);
  CS_CUDA(launch_pack_pairs(w_oihw, cout, cin, 9, ls.wf, kTapFprop, ls.wd, kTapDgrad, s));
  PixGemmParams p;
  CS_TRY(build_conv3x3(p, &bn, View{(bf16*)x, cin, 0}, cin, View{(bf16*)y, cout, 0}, cout, ls.wf, batch, height, width));
  p.stat_sum = stat_sum;
  p.stat_sq = stat_sq;
  CS_CUDA(launch_pix_gemm(p, bn, sms, s));
  return 0;
}

int cs_conv3x3_fprop_bnrelu(const void* x_raw, const float* in_scale, const float* in_shift, int batch, int height, int width,
                            int cin, const float* w_oihw, int cout, void* y, double* stat_sum, double* stat_sq, void* scratch,
                            cs_stream_t stream) {
  CS_TRY(check_layer(x_raw, y, scratch, batch, height, width, cin, cout));
  if (!in_scale || !in_shift) return fail("layer op: null pointer");
  if (((uintptr_t)in_scale | (uintptr_t)in_shift) & 15) return fail("layer op: in_scale / in_shift must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int sms, bn;
  CS_TRY(layer_sm_count(&sms));
  LayerScratch ls = carve(scratch, cin, cout);
  CS_CUDA(launch_pack_pairs(w_oihw, cout, cin, 9, ls.wf, kTapFprop, ls.wd, kTapDgrad, s));
  PixGemmParams p;
  CS_TRY(build_conv3x3(p, &bn, View{(bf16*)x_raw, cin, 0}, cin, View{(bf16*)y, cout, 0}, cout, ls.wf, batch, height, width));
  if (!p.conv3) return fail("layer op: the fused BN + ReLU operand needs conv3_gemm_kernel (CARTSEG_CONV3=0 is set)");
  p.stat_sum = stat_sum;
  p.stat_sq = stat_sq;
  p.in_scale = in_scale;
  p.in_shift = in_shift;
  CS_CUDA(launch_pix_gemm(p, bn, sms, s));
  return 0;
}

int cs_conv3x3_wgrad_bnrelu(const void* x_raw, const float* x_scale, const float* x_shift, const void* dy, int batch, int height,
                            int width, int cin, int cout, float* dw_oihw, void* scratch, cs_stream_t stream) {
  CS_TRY(check_layer(x_raw, dy, scratch, batch, height, width, cin, cout));
  if (!dw_oihw || !x_scale || !x_shift) return fail("layer op: null pointer");
  if (((uintptr_t)x_scale | (uintptr_t)x_shift) & 15) return fail("layer op: x_scale / x_shift must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int bn;
  LayerScratch ls = carve(scratch, cin, cout);
  WgradParams p;
  CS_TRY(build_conv3x3_wgrad(p, &bn, View{(bf16*)x_raw, cin, 0}, cin, View{(bf16*)dy, cout, 0}, cout, ls.dwp, batch, height, width));
  p.x_scale = x_scale;
  p.x_shift = x_shift;
  CS_CUDA(cudaMemsetAsync(ls.dwp, 0, (size_t)9 * cin * cout * sizeof(float), s));
  CS_CUDA(launch_wgrad_gemm(p, bn, s));
  CS_CUDA(launch_unpack_pairs(ls.dwp, cout, cin, 9, kTapWgrad, dw_oihw, s));
  return 0;
}

int cs_conv3x3_dgrad(const void* dy, int batch, int height, int width, int cin, const float* w_oihw, int cout, void* dx,
                     void* scratch, cs_stream_t stream) {
  CS_TRY(check_layer(dy, dx, scratch, batch, height, width, cin, cout));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int sms, bn;
  CS_TRY(layer_sm_count(&sms));
  LayerScratch ls = carve(scratch, cin, cout);
  CS_CUDA(launch_pack_pairs(w_oihw, cout, cin, 9, ls.wf, kTapFprop, ls.wd, kTapDgrad, s));
  PixGemmParams p;
  CS_TRY(build_conv3x3(p, &bn, View{(bf16*)dy, cout, 0}, cout, View{(bf16*)dx, cin, 0}, cin, ls.wd, batch, height, width));
  CS_CUDA(launch_pix_gemm(p, bn, sms, s));
  return 0;
}

int cs_conv3x3_wgrad(const void* x, const void* dy, int batch, int height, int width, int cin, int cout, float* dw_oihw,
                     void* scratch, cs_stream_t stream) {
  CS_TRY(check_layer(x, dy, scratch, batch, height, width, cin, cout));
  if (!dw_oihw) return fail("layer op: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int bn;
  LayerScratch ls = carve(scratch, cin, cout);
  WgradParams p;
  CS_TRY(build_conv3x3_wgrad(p, &bn, View{(bf16*)x, cin, 0}, cin, View{(bf16*)dy, cout, 0}, cout, ls.dwp, batch, height, width));
  CS_CUDA(cudaMemsetAsync(ls.dwp, 0, (size_t)9 * cin * cout * sizeof(float), s));
  CS_CUDA(launch_wgrad_gemm(p, bn, s));
  CS_CUDA(launch_unpack_pairs(ls.dwp, cout, cin, 9, kTapWgrad, dw_oihw, s));
  return 0;
}

int cs_convT2x2_fprop(const void* x, int batch, int height, int width, int cin, const float* w_iohw, const float* bias,
                      int cout, void* y, int y_pitch, void* scratch, cs_stream_t stream) {
  CS_TRY(check_layer(x, y, scratch, batch, height, width, cin, cout));
  if (y_pitch < cout || y_pitch % 8) return fail("layer op: bad output pitch %d", y_pitch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int sms, bn;
  CS_TRY(layer_sm_count(&sms));
  LayerScratch ls = carve(scratch, cin, cout);
  CS_CUDA(launch_pack_pairs(w_iohw, cin, cout, 4, ls.wd, kTapIdent, ls.wf, kTapIdent, s));
  PixGemmParams p;
  CS_TRY(build_convT_fprop(p, &bn, View{(bf16*)x, cin, 0}, cin, View{(bf16*)y, y_pitch, 0}, cout, ls.wf, bias, batch,
                           height, width));
  CS_CUDA(launch_pix_gemm(p, bn, sms, s));
  return 0;
}

int cs_convT2x2_dgrad(const void* dy, int dy_pitch, int batch, int height, int width, int cin, const float* w_iohw,
                      int cout, void* dx, void* scratch, cs_stream_t stream) {
  CS_TRY(check_layer(dy, dx, scratch, batch, height, width, cin, cout));
  if (dy_pitch < cout || dy_pitch % 8) return fail("layer op: bad gradient pitch %d", dy_pitch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int sms, bn;
  CS_TRY(layer_sm_count(&sms));
  LayerScratch ls = carve(scratch, cin, cout);
  CS_CUDA(launch_pack_pairs(w_iohw, cin, cout, 4, ls.wd, kTapIdent, ls.wf, kTapIdent, s));
  PixGemmParams p;
  CS_TRY(build_convT_dgrad(p, &bn, View{(bf16*)dy, dy_pitch, 0}, cout, View{(bf16*)dx, cin, 0}, cin, ls.wd, batch, height,
                           width));
  CS_CUDA(launch_pix_gemm(p, bn, sms, s));
  return 0;
}

int cs_convT2x2_wgrad(const void* x, const void* dy, int dy_pitch, int batch, int height, int width, int cin, int cout,
                      float* dw_iohw, void* scratch, cs_stream_t stream) {
  CS_TRY(check_layer(x, dy, scratch, batch, height, width, cin, cout));
  if (!dw_iohw) return fail("layer op: null pointer");
  if (dy_pitch < cout || dy_pitch % 8) return fail("layer op: bad gradient pitch %d", dy_pitch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int bn;
  LayerScratch ls = carve(scratch, cin, cout);
  WgradParams p;
  CS_TRY(build_convT_wgrad(p, &bn, View{(bf16*)x, cin, 0}, cin, View{(bf16*)dy, dy_pitch, 0}, cout, ls.dwp, batch, height,
                           width));
  CS_CUDA(cudaMemsetAsync(ls.dwp, 0, (size_t)4 * cin * cout * sizeof(float), s));
  CS_CUDA(launch_wgrad_gemm(p, bn, s));
  CS_CUDA(launch_unpack_pairs(ls.dwp, cin, cout, 4, kTapIdent, dw_iohw, s));
  return 0;
}

}  // extern "C"
