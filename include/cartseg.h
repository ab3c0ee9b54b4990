/* cartseg.h — C ABI of the B200-native U-Net hot path (libcartseg.so).
 *
 * The reference (endressa/cart-segmentation-unet) has no FFI for this path: the boundary is the
 * Python object protocol `model(x)` / `criterion(logits, targets)` (train_bce_dice.py:308-309,
 * 331-334).  Its only native boundary, pybind `lsr_cpp.lsr_forward/lsr_backward`
 * (src/training/abl_training/losses/lsr_cpp/csrc/lsr_kernel.cu:296-322), sets the conventions kept
 * here: CUDA-only (hard error otherwise, :300-302), work is enqueued on the caller's stream (:220),
 * no host synchronisation.  Each entry point below names the reference code it replaces; the
 * Python host (cart-segmentation-unet_b200/cartseg/_lib.py) binds them with ctypes and exposes them as torch.library ops.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless stated otherwise;
 *   - the caller owns every buffer (including workspaces, sized by the *_bytes queries);
 *   - nothing here allocates device memory, frees it, or synchronises; all work goes to `stream` — so every call can
 *     be captured in a CUDA graph (cs_unet_plan_bind creates the two internal backward streams and their events).
 *     The only exceptions are developer / test hooks that hand results to the HOST and say so:
 *     cs_unet_profile_read, cs_unet_trace_read (wait for their timing events), cs_abl_debug_read (synchronises);
 *   - return value 0 = success, negative = failure; cs_last_error() gives a thread-local message;
 *   - image tensors are fp32 NCHW exactly as the reference's DataLoader yields them; logits,
 *     targets, SDFs and dlogits are fp32 [B,1,H,W]; H and W must be multiples of 16.
 */
#ifndef CARTSEG_H_
#define CARTSEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cs_stream_t; /* cudaStream_t */

const char* cs_last_error(void);
int cs_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py reports it). */
long long cs_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * U-Net  (replaces DoubleConv / UNet, src/create_testset.py:40-83; logits = final_conv output,
 * the trailing sigmoid of :83 is NOT applied — every reference loss takes logits).
 * Parameter order = the 82 trainable tensors in state_dict order (conv1.conv.0.weight, ...
 * final_conv.bias), fp32, reference shapes (OIHW conv, IOHW conv-transpose).
 * BN buffers: 18 layers in state_dict order; running_mean / running_var fp32 [C],
 * num_batches_tracked int64 scalar.
 * ---------------------------------------------------------------------------------------------- */
#define CS_UNET_NUM_PARAMS 82
#define CS_UNET_NUM_BN 18
#define CS_UNET_NUM_BWD_STAGES 23 /* head, then 18 convs and 4 up-convs in reverse execution order */

typedef struct cs_unet_plan cs_unet_plan;

typedef struct cs_unet_tensors {
  const float* param[CS_UNET_NUM_PARAMS];
  float* grad[CS_UNET_NUM_PARAMS];          /* written by cs_unet_backward; may be NULL for forward */
  float* running_mean[CS_UNET_NUM_BN];
  float* running_var[CS_UNET_NUM_BN];
  long long* num_batches_tracked[CS_UNET_NUM_BN];
} cs_unet_tensors;

/* Host-side: lay out activations / packed weights / gradients for one (batch, H, W).
 * inference_only != 0 drops everything the backward pass needs (raw conv outputs, gradient and
 * weight-gradient buffers, dgrad weight packs): such a plan only accepts training == 0 forwards
 * (the pseudo-label path, src/data_preprocessing/create_pseudo_labels_gpu.py:201-215). */
int cs_unet_plan_create(cs_unet_plan** plan, int batch, int in_channels, int height, int width, int inference_only);
void cs_unet_plan_destroy(cs_unet_plan* plan);
size_t cs_unet_plan_workspace_bytes(const cs_unet_plan* plan);
/* Attach a caller-owned device workspace (>= workspace_bytes, 1024-byte aligned) and encode the
 * TMA descriptors that point into it. */
int cs_unet_plan_bind(cs_unet_plan* plan, void* workspace, size_t bytes);
/* fp32 reference-layout weights -> bf16 tap-major packs (fprop + dgrad).  Call after every
 * optimizer step (the Python module tracks parameter versions). */
int cs_unet_pack_weights(cs_unet_plan* plan, const cs_unet_tensors* t, cs_stream_t stream);
/* x fp32 [B,Cin,H,W] -> logits fp32 [B,1,H,W].  training != 0: batch statistics, running buffers
 * updated (momentum 0.1, eps 1e-5), activations kept for backward.  training == 0: running
 * statistics folded into the convolution epilogues.
 * After a training-mode forward, `x` must stay valid and unchanged until the work of the following cs_unet_backward has
 * completed: the first convolution's weight gradient re-reads the image (the forward path builds its im2col rows in
 * shared memory and keeps no copy of it in HBM). */
int cs_unet_forward(cs_unet_plan* plan, const cs_unet_tensors* t, const float* x, int training, float* logits,
                    cs_stream_t stream);
/* Backward of the last training forward.  Stages [stage_begin, stage_end) in reverse execution
 * order; gradients of the parameters owned by those stages are final when the call's work
 * completes, so a data-parallel caller can all-reduce them while later stages run.
 * frozen_encoder_convs: number of leading encoder convs (conv1.0, conv1.3, conv2.0, ...) whose parameters need no
 * gradient — 0 trains everything, 10 freezes the whole encoder (src/train_with_focalDice.py:384-391): their backward
 * stages are skipped altogether.  Any other parameter whose grad[] entry is NULL just loses its own gradient kernel
 * (weight-gradient GEMM / bias sum); gradients still flow through the layer. */
int cs_unet_backward(cs_unet_plan* plan, const cs_unet_tensors* t, const float* dlogits, int stage_begin,
                     int stage_end, int frozen_encoder_convs, cs_stream_t stream);
/* Per-launch timing of the tensor-core kernels (bench.py's roofline): while enabled, every implicit-GEMM launch of
 * cs_unet_forward / cs_unet_backward is bracketed by CUDA events on the launch stream.  cs_unet_profile_read waits for
 * the recorded events and returns, per kernel class, the summed device time (ms), algorithmic FLOPs (2*MACs) and
 * launch count since the last read.  Classes: 0-1 pix_gemm2_kernel N=256 / 128 (conv-transposes), 2 stem_gemm_kernel (first convolution),
 * 3-4 wgrad_gemm_kernel N=128 / 64, 5-7 conv3_gemm_kernel N=256 / 128 / 64 (3x3 convolutions, fprop and dgrad),
 * 8 wgrad9_gemm_kernel (3x3 weight gradients with Cout = 64). */
/* cs_unet_backward runs the weight-gradient GEMMs on an internal lower-priority stream so that they overlap the
 * HBM-bound BatchNorm-backward passes of the following layers (forked from / joined into `stream` with events).
 * cs_unet_set_overlap(plan, 0) serialises everything on the caller's stream (used for per-kernel timing). */
int cs_unet_set_overlap(cs_unet_plan* plan, int enable);
/* By default every cs_unet_backward call makes `stream` wait for its internal streams before returning.  A data-parallel
 * caller that runs the stages bucket by bucket can defer that join (enable = 1): the calls then only enqueue work, and
 * cs_unet_backward_wait(plan, some_stream) makes `some_stream` wait for everything enqueued so far — the communication
 * stream waits per bucket, the caller's stream once after the last stage, so the weight-gradient overlap is not cut at
 * bucket boundaries. */
int cs_unet_set_deferred_join(cs_unet_plan* plan, int enable);
/* The persistent implicit-GEMM kernels launch one CTA per SM and assume a single wave.  When another kernel (an NCCL
 * all-reduce of the data-parallel path) holds some SMs for the whole step, cap the persistent grids at `sms` so that no
 * CTA has to wait for a second wave (sms <= 0 restores the full device). */
int cs_unet_plan_set_sm_limit(cs_unet_plan* plan, int sms);
int cs_unet_backward_wait(cs_unet_plan* plan, cs_stream_t stream);
/* The weight-gradient GEMMs of the deep levels are enqueued late (when the main stream reaches the encoder's level-3
 * stage `*flush_stage`), where they overlap HBM-bound BatchNorm-backward passes instead of time-slicing with tensor-bound
 * dgrads.  A caller that consumes gradients stage by stage (data parallel) must treat the parameter gradients of the
 * returned stages as final only once stage `*flush_stage` has been enqueued.  Returns the number of such stages. */
int cs_unet_backward_held_stages(int* flush_stage, int* stages, int capacity);
#define CS_UNET_NUM_PROFILE_CLASSES 9
int cs_unet_profile(cs_unet_plan* plan, int enable);
int cs_unet_profile_read(cs_unet_plan* plan, int n_classes, double* ms, double* flops, long long* launches);
/* Developer timeline of cs_unet_forward (training mode) / cs_unet_backward: while enabled, the convolution, BatchNorm
 * and conv-transpose launches of the forward pass (kinds 10 conv, 11 BN + ReLU pass, 12 conv-transpose) and every launch
 * of the backward pass (both internal streams) are bracketed by timing events.  cs_unet_trace_read waits for them and returns, per launch, a label (kind * 100 + layer:
 * 1 BN-backward reduce, 2 BN-backward apply, 3 conv dgrad, 4 conv wgrad, 5 im2col of the input image, 6 head, 7 conv-transpose dgrad, 8 its wgrad,
 * 9 its bias gradient) and begin / end times in ms relative to the first launch.  Returns the number of entries. */
int cs_unet_trace(cs_unet_plan* plan, int enable);
int cs_unet_trace_read(cs_unet_plan* plan, int capacity, int* labels, double* begin_ms, double* end_ms);
/* Test hook: copies one internal NHWC bf16 tensor of the plan into `dst` as dense fp32 NCHW (dst == NULL: only
 * report dims_out = {B, C, H, W}).  kind 0..5 index a conv (0..17): raw output, activation, grad wrt raw output,
 * grad wrt activation, pooled activation, grad wrt pooled; kind 6/7 index a conv-transpose (0..3): output, its grad. */
int cs_unet_debug_read(cs_unet_plan* plan, int kind, int index, int dims_out[4], float* dst, cs_stream_t stream);
/* Which parameter indices (into param[]/grad[]) a backward stage finalises; returns the count. */
int cs_unet_stage_params(int stage, int* out_indices, int capacity);

/* ------------------------------------------------------------------------------------------------
 * Signed distance maps (replaces signed_distance_map_np / batch_sdf_from_masks,
 * src/train_with_boundary_loss.py:191-217: exact EDT, sdf = dist_out - dist_in, zeros for
 * all-fg / all-bg, fp32 divide by `norm`).  fg = ge ? src >= thr : src > thr.
 * ---------------------------------------------------------------------------------------------- */
size_t cs_sdf_scratch_bytes(int batch, int height, int width);
int cs_sdf(const float* src, float thr, int ge, int batch, int height, int width, float norm, float* sdf,
           void* scratch, cs_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused losses.  One descriptor covers
 *   BCEDiceLoss            train_bce_dice.py:186-199          (alpha=1, gamma=0)
 *   FocalLoss/FocalDiceLoss src/train_with_focalDice.py:195-235
 *   SymmetricBoundaryLoss / CompositeSegLoss  src/train_with_boundary_loss.py:242-282
 *   BCEDiceLossPerSample   src/finetune_for_224.py:208-221    (per_row = 1)
 * loss = w_elem * mean|sum(alpha (1-p_t)^gamma BCE) + w_dice * (1 - mean_rows dice)
 *        + w_bgt * mean|p*sdf_gt| + w_bpred * mean|(1-p)*(-sdf_pred)|
 * ---------------------------------------------------------------------------------------------- */
typedef struct cs_loss_desc {
  int rows;            /* Dice rows (= B*C for dims (2,3), B for dims (1,2,3)) */
  long long n;         /* elements per row, multiple of 4 */
  float w_elem, alpha, gamma;
  int elem_sum;
  float w_dice, smooth;
  float w_bgt, w_bpred;
  int use_abs;
  int per_row;
} cs_loss_desc;

size_t cs_loss_scratch_bytes(int rows);
/* loss_out: 1 float (or rows floats if per_row).  scratch is re-read by cs_loss_backward. */
int cs_loss_forward(const cs_loss_desc* d, const float* logits, const float* targets, const float* sdf_gt,
                    const float* sdf_pred, void* scratch, float* loss_out, cs_stream_t stream);
/* grad_out: device scalar (or [rows] if per_row), may be NULL (= 1). */
int cs_loss_backward(const cs_loss_desc* d, const float* logits, const float* targets, const float* sdf_gt,
                     const float* sdf_pred, const void* scratch, const float* grad_out, float* dlogits,
                     cs_stream_t stream);

/* FocalLoss(reduction="none") (src/train_with_focalDice.py:214-219): the unreduced map
 * out[i] = alpha * (1 - p_t)^gamma * BCE(logits[i], targets[i]), and its backward against an element-wise grad_out. */
int cs_focal_map_forward(const float* logits, const float* targets, long long n, float alpha, float gamma, float* out,
                         cs_stream_t stream);
int cs_focal_map_backward(const float* logits, const float* targets, const float* grad_out, long long n, float alpha,
                          float gamma, float* dlogits, cs_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Thresholding / metrics (replaces dice_metric, iou_metric, find_best_threshold
 * train_bce_dice.py:201-232; precision_recall_f1 src/train_with_focalDice.py:266-284; the
 * pseudo-label threshold create_pseudo_labels_gpu.py:294).  Thresholds are given as logit-space
 * bounds xs[k]: pred = (x >= xs[k]); the host derives xs[k] so that the result is bit-identical to
 * sigmoid(x) > t (or >= t).
 *   counts[row][k] = {sum pred, sum pred*target}   soft[row] = {sum p, sum t, sum p*t}
 * ---------------------------------------------------------------------------------------------- */
int cs_threshold_stats(const float* logits, const float* targets, int rows, long long n, const float* xs, int K,
                       double* counts, double* soft, cs_stream_t stream);
int cs_threshold_mask(const float* logits, long long n, float xstar, uint8_t* mask, cs_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Active Boundary Loss, binary case.  Replaces ABL.forward
 * (src/training/losses/abl.py:66-212) with its LabelSmoothSoftmaxCEV1 criterion
 * (src/training/losses/label_smooth.py:14-57): the latest training recipe of the reference,
 * src/training/train_BCEDice_ABL.py:264-302.  No host synchronisation: the adaptive KL threshold of
 * abl.py:78-83 is found on the device from a histogram over `eps_ladder` (the float32 values
 * 1e-5 * 1.2^k the reference's loop visits, computed by the host in float64 exactly as Python does),
 * and the scipy EDT of abl.py:16-24 runs as an exact integer EDT on the device.
 *   loss_out[0] = mean over kept boundary pixels of weight * CE (NaN when none is kept, as torch.mean of
 *                 an empty tensor), loss_out[1] = 1 if the predicted boundary is non-empty, else 0 — the case
 *                 in which the reference returns None (abl.py:197-198) and the caller skips the term.
 * per_image_maps == 0 reproduces the reference's distance-map indexing (map of batch entry n = channel n%2
 * of image n/2, abl.py:166-167 + 121); != 0 uses image n's own map.
 * ---------------------------------------------------------------------------------------------- */
#define CS_ABL_LADDER 80
typedef struct cs_abl_desc {
  int batch, height, width;
  float max_n;                 /* float32(H*W*max_N_ratio), abl.py:69 */
  float label_smoothing;       /* 0.2 */
  float max_clip_dist;         /* 20 */
  long long ignore_label;      /* 255 */
  int per_image_maps;
  float eps_ladder[CS_ABL_LADDER];
} cs_abl_desc;
size_t cs_abl_scratch_bytes(int batch, int height, int width);
/* logits, targets: fp32 [B,1,H,W] (targets are truncated to integers like target.long()); scratch 256-byte aligned. */
int cs_abl_forward(const cs_abl_desc* d, const float* logits, const float* targets, void* scratch, float* loss_out,
                   cs_stream_t stream);
/* dlogits = grad_out[0] * d loss / d logits (zero when the loss was invalid); scratch from the forward call. */
int cs_abl_backward(const cs_abl_desc* d, const float* logits, const void* scratch, const float* grad_out,
                    float* dlogits, cs_stream_t stream);
/* Test hook (synchronises `stream`): chosen threshold, counters, and the [B,H,W] distance / KL maps copied to HOST. */
int cs_abl_debug_read(const cs_abl_desc* d, const void* scratch, float* eps, int* ladder_index, unsigned long long* kept,
                      unsigned long long* pred_boundary, uint16_t* dist_map_host, float* kl_map_host, cs_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pseudo-label post-processing: what the reference does on the HOST with the
 * probabilities it copies back at 4 B/px (src/data_preprocessing/create_pseudo_labels_gpu.py:201-215,
 * 294-300) and in its mask clean-up tools (src/data_preprocessing/clean_masks.py:12-32,
 * src/data_preprocessing/remove_blops.py:14-33).
 * ---------------------------------------------------------------------------------------------- */
/* probs[i] = (first ? 0 : probs[i]) + weight * sigmoid(logits[i])   — ensemble_forward, one call per model. */
int cs_ensemble_accumulate(const float* logits, float weight, long long n, int first, float* probs, cs_stream_t stream);
/* Per image b of n pixels: mask = (probs >= threshold) written as mask_value/0 (mask may be NULL), and
 * stats[b] = { foreground pixel count, median(|p - 0.5| * 2) (exact, numpy's even-count rule),
 *              mean binary entropy of clip(p, 1e-6, 1 - 1e-6), n }. */
int cs_pseudo_qc(const float* probs, int batch, long long n, float threshold, int mask_value, uint8_t* mask,
                 double* stats, cs_stream_t stream);
/* fg = mask > bin_threshold.  fill_holes: background the 4-connected flood fill from pixel (0,0) cannot reach becomes
 * foreground (a foreground (0,0) makes everything foreground — cv2.floodFill semantics of clean_masks.py:16-22).
 * keep_largest: only the largest 8-connected component survives (ties: OpenCV's lowest label).  out is {0,255}. */
size_t cs_mask_cleanup_scratch_bytes(int batch, int height, int width);
int cs_mask_cleanup(const uint8_t* mask, int batch, int height, int width, int bin_threshold, int fill_holes,
                    int keep_largest, uint8_t* out, void* scratch, cs_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Input side: what the reference's Dataset does per image on the CPU —
 * letterbox_image_with_side_padding (train_bce_dice.py:42-85), cv2.resize INTER_LINEAR / INTER_NEAREST
 * (:147-148), A.Resize + A.Normalize + ToTensorV2 (:171-176; create_pseudo_labels_gpu.py:113-117) — for a
 * whole batch of variable-size uint8 images in one launch.  The resize is OpenCV's 8-bit bilinear kernel
 * bit for bit; the letterboxed canvas is virtual (taps outside the image read black).
 * ---------------------------------------------------------------------------------------------- */
typedef struct cs_image_desc {
  const uint8_t* data;       /* device pointer: HWC uint8 with 3 channels (images) or HW uint8 (masks) */
  int height, width, pitch;  /* pitch = bytes per row */
  int canvas_h, canvas_w;    /* size of the letterboxed canvas (= height, width when there is none) */
  int x0, y0;                /* where the image sits inside the canvas */
  int reserved;
} cs_image_desc;
/* Host helper: canvas side and offsets of train_bce_dice.py:56-80 (side padding = Python round(width * ratio)). */
int cs_letterbox_geometry(int height, int width, double side_padding_ratio, int* canvas, int* x0, int* y0);
/* descs_device: `batch` descriptors in DEVICE memory.  out_nchw: fp32 [batch,3,out_size,out_size] =
 * (resized - mean*255) * (1/(std*255)); bgr != 0 reverses the channel order first (cv2.imread -> RGB). */
int cs_preproc_images(const cs_image_desc* descs_device, int batch, int out_size, const float mean[3],
                      const float std_[3], int bgr, float* out_nchw, cs_stream_t stream);
/* Masks: nearest-neighbour resize of each HW uint8 image to out_size^2, divided by 255 -> fp32 [batch,1,S,S]. */
int cs_preproc_masks(const cs_image_desc* descs_device, int batch, int out_size, float* out, cs_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Single-layer entry points (unit tests / micro-benchmarks of the tcgen05 kernels).  Activations
 * are NHWC bf16; weights fp32 in the reference layout; `scratch` must hold the packed copies
 * (cs_layer_scratch_bytes).  dw is fp32 in the reference layout.
 * ---------------------------------------------------------------------------------------------- */
size_t cs_layer_scratch_bytes(int cin, int cout);
int cs_conv3x3_fprop(const void* x_nhwc, int batch, int height, int width, int cin, const float* w_oihw, int cout,
                     void* y_nhwc, double* stat_sum, double* stat_sq, void* scratch, cs_stream_t stream);
int cs_conv3x3_dgrad(const void* dy_nhwc, int batch, int height, int width, int cin, const float* w_oihw, int cout,
                     void* dx_nhwc, void* scratch, cs_stream_t stream);
/* The same convolution / weight gradient on relu(x_raw * scale[c] + shift[c]) (rounded to bf16), applied to the operand
 * in shared memory: how convX.3 of a DoubleConv (src/create_testset.py:40-52) consumes the RAW output of convX.0 in
 * training, so that the BatchNorm + ReLU activation between the two convolutions never exists in HBM.  scale / shift:
 * fp32 [cin], 16-byte aligned. */
int cs_conv3x3_fprop_bnrelu(const void* x_raw_nhwc, const float* in_scale, const float* in_shift, int batch, int height,
                            int width, int cin, const float* w_oihw, int cout, void* y_nhwc, double* stat_sum,
                            double* stat_sq, void* scratch, cs_stream_t stream);
int cs_conv3x3_wgrad_bnrelu(const void* x_raw_nhwc, const float* x_scale, const float* x_shift, const void* dy_nhwc,
                            int batch, int height, int width, int cin, int cout, float* dw_oihw, void* scratch,
                            cs_stream_t stream);
int cs_conv3x3_wgrad(const void* x_nhwc, const void* dy_nhwc, int batch, int height, int width, int cin, int cout,
                     float* dw_oihw, void* scratch, cs_stream_t stream);
int cs_convT2x2_fprop(const void* x_nhwc, int batch, int height, int width, int cin, const float* w_iohw,
                      const float* bias, int cout, void* y_nhwc, int y_pitch, void* scratch, cs_stream_t stream);
int cs_convT2x2_dgrad(const void* dy_nhwc, int dy_pitch, int batch, int height, int width, int cin,
                      const float* w_iohw, int cout, void* dx_nhwc, void* scratch, cs_stream_t stream);
int cs_convT2x2_wgrad(const void* x_nhwc, const void* dy_nhwc, int dy_pitch, int batch, int height, int width,
                      int cin, int cout, float* dw_iohw, void* scratch, cs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CARTSEG_H_ */
