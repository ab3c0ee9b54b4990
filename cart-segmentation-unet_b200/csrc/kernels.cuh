// Launchers for the bandwidth-bound (SIMT) kernels of the U-Net hot path.  All activations are NHWC
// bf16 ("pixel rows" of C channels, optionally embedded in a wider row: pitch + channel offset);
// logits / targets / SDFs are fp32 [B, H*W] exactly as the reference API supplies them.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch_count.cuh"

namespace cs {

typedef __nv_bfloat16 bf16;

// ---- weights -------------------------------------------------------------------------------
// in[(a*Nb + b)*T + t] (fp32)  ->  out_ab[map_ab[t]][a][b]  and  out_ba[map_ba[t]][b][a]  (bf16)
struct TapMap { int v[9]; };
cudaError_t launch_pack_pairs(const float* in, int Na, int Nb, int T, bf16* out_ab, TapMap map_ab, bf16* out_ba,
                              TapMap map_ba, cudaStream_t s);
// The same for every weight tensor of the network in one launch (+ the stem pack below).
struct PackJob {
  const float* in; bf16* out_ab; bf16* out_ba;
  int Na, Nb, T, map_ab, map_ba;      // map_*: index into PackBatch::maps
  int tile0, tiles_x;                 // filled by the launcher
};
struct PackBatch {
  static constexpr int kMaxJobs = 24;
  PackJob job[kMaxJobs];
  TapMap maps[4];
  int n, total_tiles;
  const float* first_in; bf16* first_out; int first_cout, first_cin;   // stem (first_in == NULL: none)
};
cudaError_t launch_pack_batch(PackBatch& pb, cudaStream_t s);
// conv1.0: W[64][Cin<=7][3][3] -> [64][64] with k = (r*3+s)*Cin + c, zero padded
cudaError_t launch_pack_first(const float* in, int Cout, int Cin, bf16* out, cudaStream_t s);
// dWp[t][a][b] (fp32) -> grad[(a*Nb+b)*T + map[t]]
cudaError_t launch_unpack_pairs(const float* dwp, int Na, int Nb, int T, TapMap map, float* grad, cudaStream_t s);
cudaError_t launch_unpack_first(const float* dwp, int Cout, int Cin, float* grad, cudaStream_t s);

// ---- input ---------------------------------------------------------------------------------
// x fp32 NCHW [B,Cin,H,W] -> col bf16 [B*H*W][64], col[p][(r*3+s)*Cin + c] = x[b,c,h+r-1,w+s-1]
cudaError_t launch_im2col_first(const float* x, int B, int Cin, int H, int W, bf16* col, cudaStream_t s);

// ---- batch norm ----------------------------------------------------------------------------
struct BnFinalizeArgs {
  const double* sum; const double* sq; double count;
  const float* gamma; const float* beta; const float* conv_bias;
  float* running_mean; float* running_var; long long* num_batches_tracked;
  float momentum, eps;
  float* scale; float* shift; float* mean; float* invstd;
  int C;
};
// Eval mode: running statistics folded into (scale, shift) for up to 18 layers in one launch.
struct BnFoldBatch {
  static constexpr int kMax = 18;
  const float* gamma[kMax]; const float* beta[kMax]; const float* conv_bias[kMax];
  const float* rm[kMax]; const float* rv[kMax];
  float* scale[kMax]; float* shift[kMax];
  int C[kMax];
  int layers;
  float eps;
};
cudaError_t launch_bn_fold_eval(const BnFoldBatch& a, cudaStream_t s);
// Training forward: (scale, shift) are derived from the batch statistics inside the kernel (fused bn_finalize_train:
// block 0 publishes scale/shift/mean/invstd and updates the running statistics), then
// out[p][out_c0 + c] = relu(y[p][c]*scale[c] + shift[c]); optional 2x2 max-pooled copy (pitch C)
// head.logits != NULL (last layer, C == 64, no pooling): the 1x1 head is evaluated in the same pass,
// logits[p] = sum_c out[p][c] * w[c] + b  on the bf16 values as stored.
struct HeadFwd { const float* w; const float* b; float* logits; int skip_store; };   // skip_store: `out` is not written
// out[p][c] = relu(y[p][c] * scale[c] + shift[c]) from the coefficients a training forward published (test hook:
// materialises the last layer's activation, which the training path never stores)
cudaError_t launch_bn_apply_relu(const bf16* y, long long P, int C, const float* scale, const float* shift, bf16* out,
                                 cudaStream_t s);
// statistics -> (scale, shift, mean, invstd) + running statistics, nothing else (the consumer applies BN + ReLU itself)
cudaError_t launch_bn_finalize(const BnFinalizeArgs& fin, cudaStream_t s);
cudaError_t launch_bn_relu(const bf16* y, int B, int H, int W, int C, const BnFinalizeArgs& fin, bf16* out,
                           int out_pitch, int out_c0, bf16* pooled, const HeadFwd& head, cudaStream_t s);
struct BnBwdArgs {
  const bf16* g; int g_pitch, g_c0;        // gradient w.r.t. the post-ReLU activation
  const float* head_dlogits;               // optional (no pooling): g[p][c] = bf16(head_dlogits[p] * head_w[c]) instead of `g`
  const float* head_w;
  float* head_grad_w; float* head_grad_b;  // optional with head_dlogits: the 1x1 head's parameter gradients
                                           //   grad_w[c] = sum_p dlogits[p] * act[p][c], grad_b = sum_p dlogits[p]
                                           //   (act recomputed from y: the head's input is never stored)
  const bf16* g_pool;                      // optional: gradient w.r.t. the 2x2-pooled activation (pitch C)
  const bf16* y;                           // raw conv output (pitch C)
  const float* scale; const float* shift; const float* mean; const float* invstd;
  float* partial;                          // [blocks][2*C] per-block partial sums (bn_bwd_scratch_bytes)
  float* c1; float* c2;                    // per-channel mean(g*mask), mean(g*mask*xhat)
  bf16* dy;                                // gradient w.r.t. the raw conv output (pitch C)
  float* grad_gamma; float* grad_beta; float* grad_conv_bias;
  int B, H, W, C;
  int reverse;                             // set by the launchers: the reduction walks the tensor from its end (L2 reuse)
};
size_t bn_bwd_scratch_bytes(int maxC);
cudaError_t launch_bn_bwd_reduce(const BnBwdArgs& a, cudaStream_t s);  // + finalize: c1/c2, grad_gamma/beta/bias
cudaError_t launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t s);

// ---- 1x1 head ------------------------------------------------------------------------------
cudaError_t launch_head_fwd(const bf16* act, long long P, int C, const float* w, const float* b, float* logits,
                            cudaStream_t s);
// g_act[p][c] = dlogits[p]*w[c];  grad_w[c] += sum_p dlogits[p]*act[p][c];  grad_b += sum_p dlogits[p]
cudaError_t launch_head_bwd(const bf16* act, const float* dlogits, long long P, int C, const float* w, bf16* g_act,
                            float* grad_w, float* grad_b, cudaStream_t s);   // g_act may be NULL (not materialised)
cudaError_t launch_head_grad_act(const float* dlogits, long long P, int C, const float* w, bf16* g_act, cudaStream_t s);

// ---- conv-transpose bias gradient -------------------------------------------------------------
// grad_b[c] = sum_p g[p][c0 + c]
cudaError_t launch_channel_sum(const bf16* g, int pitch, int c0, long long P, int C, float* out, cudaStream_t s);

// the same from per-channel fp64 column sums left by a pixel-GEMM epilogue: out[c] = sum[c] (c < C); sum / sq[0..n) reset to 0
cudaError_t launch_stat_to_bias(double* sum, double* sq, int n, int C, float* out, cudaStream_t s);

// ---- debug read-back ------------------------------------------------------------------------
cudaError_t launch_nhwc_to_nchw_f32(const bf16* src, int pitch, int c0, int B, int H, int W, int C, float* dst,
                                    cudaStream_t s);

// ---- exact EDT / signed distance map --------------------------------------------------------
// fg(pixel) = ge ? (src >= thr) : (src > thr);  sdf = (+dist to nearest fg | -dist to nearest bg) / norm,
// all-fg / all-bg images give 0.  scratch: B*H*W*4 bytes + B*8 bytes.
size_t sdf_scratch_bytes(int B, int H, int W);
cudaError_t launch_sdf(const float* src, float thr, int ge, int B, int H, int W, float norm, float* sdf,
                       void* scratch, cudaStream_t s);

// ---- fused segmentation losses ---------------------------------------------------------------
struct LossArgs {
  const float* logits; const float* targets;     // [rows][n]
  const float* sdf_gt; const float* sdf_pred;    // optional, [rows][n]
  int rows; long long n;                         // rows = samples (Dice is per row)
  float w_elem;       // weight of the element-wise term  alpha*(1-p_t)^gamma * BCE
  float alpha, gamma; // alpha=1,gamma=0 -> plain BCE
  int elem_sum;       // 1: sum over elements instead of mean
  float w_dice, smooth;
  float w_bgt, w_bpred; int use_abs;
  int per_row;        // 1: loss_out is [rows] (per-sample loss, finetune_for_224.py:208-221)
  double* stats;      // [rows][8] scratch + 1 counter (zeroed by the launcher)
  float* loss_out;
  const float* grad_out; // device scalar (or [rows] if per_row)
  float* dlogits;
};
size_t loss_scratch_bytes(int rows);
cudaError_t launch_loss_forward(const LossArgs& a, cudaStream_t s);
cudaError_t launch_loss_backward(const LossArgs& a, cudaStream_t s);

// FocalLoss(reduction="none"): out[i] = alpha (1-p_t)^gamma BCE_i ; dx[i] = go[i] * d out[i] / d x[i]
cudaError_t launch_focal_map_forward(const float* x, const float* t, long long n, float alpha, float gamma, float* out,
                                     cudaStream_t s);
cudaError_t launch_focal_map_backward(const float* x, const float* t, const float* go, long long n, float alpha,
                                      float gamma, float* dx, cudaStream_t s);

// ---- threshold / metrics ---------------------------------------------------------------------
// For each row and threshold k (given as logit-space bounds xs[k], pred = x >= xs[k]):
//   counts[row][k] = {sum pred, sum pred*t};  soft[row] = {sum p, sum t, sum p*t}
cudaError_t launch_threshold_stats(const float* logits, const float* targets, int rows, long long n, const float* xs,
                                   int K, double* counts, double* soft, cudaStream_t s);
cudaError_t launch_threshold_mask(const float* logits, long long n, float xstar, uint8_t* mask, cudaStream_t s);

// ---- Active Boundary Loss (binary case) ------------------------------------------------------
// src/training/losses/abl.py:66-212 + label_smooth.py:14-57.  ladder: the thresholds the reference's adaptive
// loop can visit (host-computed so that they are the same float32 values); loss_out = {loss, valid}.
static constexpr int kAblLadder = 80;
struct AblLadder { float v[kAblLadder]; };
size_t abl_scratch_bytes(int B, int H, int W);
cudaError_t launch_abl_forward(const float* logits, const float* targets, int B, int H, int W, const AblLadder& ladder,
                               float max_n, float smoothing, float max_clip, long long ignore_label, int faithful,
                               void* scratch, float* loss_out, cudaStream_t s);
cudaError_t launch_abl_backward(const float* logits, int B, int H, int W, float smoothing, float max_clip,
                                const void* scratch, const float* grad_out, float* dlogits, cudaStream_t s);
// test hook (synchronises): threshold chosen, counters, the distance map and the KL map of the last forward
cudaError_t abl_debug_read(const void* scratch, int B, int H, int W, float* eps, int* k, unsigned long long* kept,
                           unsigned long long* pred_boundary, unsigned short* dmap_out, float* kl_out, cudaStream_t s);

// ---- pseudo-label post-processing -------------------------------------------------------------
// probs = (first ? 0 : probs) + w * sigmoid(logits)      (create_pseudo_labels_gpu.py:201-215)
cudaError_t launch_ensemble_accumulate(const float* logits, float w, long long n, int first, float* probs, cudaStream_t s);
// per image: mask = probs >= thr (written as mask_value / 0, may be NULL); stats[b] = {fg pixels, median(|p-0.5|*2),
// mean entropy, n}   (create_pseudo_labels_gpu.py:294-300)
cudaError_t launch_pseudo_qc(const float* probs, int B, long long n, float thr, int mask_value, uint8_t* mask,
                             double* stats, cudaStream_t s);
// fg = mask > bin_thr; optional corner flood-fill hole filling; optional largest 8-connected component; out {0,255}
size_t mask_cleanup_scratch_bytes(int B, int H, int W);
cudaError_t launch_mask_cleanup(const uint8_t* mask, int B, int H, int W, int bin_thr, int fill_holes, int keep_largest,
                                uint8_t* out, void* scratch, cudaStream_t s);

// ---- input side: letterbox + resize + normalise ------------------------------------------------
// One descriptor per image (array in DEVICE memory; same layout as cs_image_desc of include/cartseg.h).
struct ImageDesc {
  const uint8_t* data;          // HWC uint8 (3 channels) or HW uint8 (masks)
  int height, width, pitch;     // pitch in bytes
  int canvas_h, canvas_w;       // size of the (virtual) letterboxed canvas the resize reads
  int x0, y0;                   // position of the image inside the canvas
  int reserved;
};
struct Norm3 { float mean255[3]; float inv_std255[3]; };
cudaError_t launch_preproc_images(const ImageDesc* descs, int B, int S, const Norm3& nrm, int bgr, float* out,
                                  cudaStream_t s);
cudaError_t launch_preproc_masks(const ImageDesc* descs, int B, int S, float* out, cudaStream_t s);

}  // namespace cs
