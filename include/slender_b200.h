/*
 * slender_b200.h -- C ABI of libslender_b200.so (NVIDIA B200 / sm_100a).
 *
 * This is the drop-in boundary for the dense-head hot path of wanzysky/SlenderObjDet.  Each entry
 * point names the reference interface it replaces (paths relative to the reference checkout;
 * "d2/" = detectron2/detectron2/, "sd/" = slender_det/).  The reference binds its native code with
 * pybind11 (d2/layers/csrc/vision.cpp:76-92); this library is bound with ctypes from
 * slenderobjdet_b200/_lib.py -- plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current device, tensors are contiguous, NCHW
 *     (the reference's layout, d2/layers/csrc/deformable/deform_conv_cuda.cu:312-314, :824-825);
 *   - the caller allocates every output (d2/layers/deform_conv.py:42-46, :89-90, :113, :242-246);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 on success, negative sdb_status on error; sdb_last_error() gives the
 *     thread-local message (the Python shim turns it into RuntimeError, mirroring TORCH_CHECK in
 *     deform_conv_cuda.cu:140-270);
 *   - `offset` and `mask` are always float32 (sampling coordinates stay fp32 even in bf16 mode);
 *     `io_dtype` is the type of x / weight / bias / out / grad tensors.
 */
#ifndef SLENDER_B200_H_
#define SLENDER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_ABI_VERSION 4

typedef enum {
  SDB_OK = 0,
  SDB_ERR_INVALID = -1,      /* bad shape / argument (shape_check, deform_conv_cuda.cu:140-270) */
  SDB_ERR_UNSUPPORTED = -2,  /* valid, but not implemented for this dtype / math mode / shape */
  SDB_ERR_WORKSPACE = -3,    /* workspace NULL or too small */
  SDB_ERR_CUDA = -4          /* a CUDA runtime call or launch failed */
} sdb_status;

typedef enum { SDB_F32 = 0, SDB_BF16 = 1 } sdb_dtype;

/* Arithmetic of the contraction.  FP32: SIMT kernels, fp32 multiply-accumulate, any shape.
 * BF16: tcgen05 tensor-core kernels; sampled values and weights rounded to bf16, fp32
 * accumulate in TMEM.  Requires groups == 1, deformable_groups == 1, C_in % 64 == 0,
 * C_out % 16 == 0, C_out <= 256 (sdb_dcn_supported tells). */
typedef enum {
  SDB_MATH_FP32 = 0,
  SDB_MATH_BF16 = 1,
  /* float32 tensors on tcgen05.mma.kind::tf32 (FORWARD; the backward entry points run the exact FP32 kernels, so
   * gradients keep fp32 accuracy).  The bilinear sample is computed in fp32, then
   *   TF32  : sample and weight rounded to tf32 (10 mantissa bits), one pass: rel ~4e-4;
   *   TF32X3: error-compensated split hi + lo of both operands, three passes (lo*hi + hi*lo + hi*hi): fp32-accurate,
   *           rel ~2e-5 against the fp32 oracle at K = 2304 (what is left is the accumulator's fp32 additions).
   * Require groups == 1, deformable_groups == 1, C_in % 32 == 0, C_out % 16 == 0, C_out <= 256, kH*kW <= 16. */
  SDB_MATH_TF32 = 2,
  SDB_MATH_TF32X3 = 3
} sdb_math;

/* Geometry of one deformable convolution (same meaning as the integer arguments of
 * deform_conv_forward / modulated_deform_conv_forward, d2/layers/csrc/deformable/deform_conv.h:8-112).
 * The same struct serves DCN v1 (mask == NULL at the call) and modulated DCN v2 (mask != NULL). */
typedef struct {
  int32_t N, C_in, H, W;        /* input  [N, C_in, H, W]                                   */
  int32_t C_out, kH, kW;        /* weight [C_out, C_in/groups, kH, kW]                      */
  int32_t sH, sW, pH, pW, dH, dW;
  int32_t groups, deformable_groups;
  /* offset [N, deformable_groups*2*kH*kW, Ho, Wo]: channel 2*(i*kW+j) = dy, +1 = dx
   * (deform_conv_cuda_kernel.cu:257-269); mask [N, deformable_groups*kH*kW, Ho, Wo] (:843-847). */
} sdb_dcn_geom;

const char* sdb_last_error(void);
int sdb_abi_version(void);

/* Output spatial size; returns SDB_ERR_INVALID if it would be <= 0 (deform_conv.py:147-152). */
int sdb_dcn_output_size(const sdb_dcn_geom* g, int32_t* Ho, int32_t* Wo);

/* 1 if (geom, io_dtype, math) can run, 0 otherwise (message in sdb_last_error). */
int sdb_dcn_supported(const sdb_dcn_geom* g, int io_dtype, int math);

typedef enum { SDB_OP_FORWARD = 0, SDB_OP_BACKWARD_DATA = 1, SDB_OP_BACKWARD_WEIGHT = 2 } sdb_dcn_op;

/* Scratch bytes the op needs (0 for SDB_MATH_FP32).  The caller passes a device buffer of at
 * least this size; it replaces the `columns` / `ones` scratch tensors of the reference
 * (deform_conv_cuda.cu:345-352), which are no longer materialised. */
size_t sdb_dcn_workspace_bytes(int op, const sdb_dcn_geom* g, int io_dtype, int math);

/* Bytes of the packed (NHWC bf16) copy of x that BF16 math uses; forward can export it
 * (`x_packed_out`) so the backward calls need not rebuild it (`x_packed`).  0 for FP32 math. */
size_t sdb_dcn_packed_input_bytes(const sdb_dcn_geom* g, int math);
/* Bytes of the saved sampled columns of one problem (sdb_dcn_problem.columns): kH*kW*C_in*2 per output pixel, rounded
 * up to whole 128-pixel tiles.  0 for FP32 math. */
size_t sdb_dcn_columns_bytes(const sdb_dcn_geom* g, int math);

/* Replaces deform_conv_forward (deform_conv.h:116-161 -> deform_conv_cuda.cu:272-438) and
 * modulated_deform_conv_forward (deform_conv.h:263-312 -> deform_conv_cuda.cu:804-927).
 *   out[n,o,h,w] = bias[o] + sum_{c,i,j} W[o,c,i,j] * mask * bilinear(x[n,c], p(h,w,i,j))
 * mask == NULL: v1.  bias == NULL: no bias.  out is overwritten.  x_packed_out may be NULL. */
int sdb_dcn_forward(const void* x, const float* offset, const float* mask, const void* weight,
                    const void* bias, void* out, const sdb_dcn_geom* g, int io_dtype, int math,
                    void* workspace, size_t workspace_bytes, void* x_packed_out, void* stream);

/* Replaces deform_conv_backward_input (deform_conv.h:163-211 -> deform_conv_cuda.cu:440-628) and the
 * grad_input / grad_offset / grad_mask part of modulated_deform_conv_backward (:929-1129).
 *   grad_x      : ACCUMULATED into (caller pre-zeroes, deform_conv.py:89)  -- may be NULL
 *   grad_offset : overwritten (float32)                                     -- may be NULL
 *   grad_mask   : overwritten (float32), v2 only                            -- may be NULL
 * x_packed: optional NHWC-bf16 copy exported by sdb_dcn_forward (BF16 math), else NULL. */
int sdb_dcn_backward_data(const void* x, const float* offset, const float* mask, const void* weight,
                          const void* grad_out, void* grad_x, float* grad_offset, float* grad_mask,
                          const sdb_dcn_geom* g, int io_dtype, int math, void* workspace,
                          size_t workspace_bytes, const void* x_packed, void* stream);

/* Replaces deform_conv_backward_filter (deform_conv.h:213-260 -> deform_conv_cuda.cu:630-802) and the
 * grad_weight / grad_bias part of modulated_deform_conv_backward.
 *   grad_weight += scale * dY . col^T  (ACCUMULATED, as addmm_ beta=1 :770-777; float32 always)
 *   grad_bias   += scale * sum dY      (ACCUMULATED; float32 always; may be NULL) */
int sdb_dcn_backward_weight(const void* x, const float* offset, const float* mask,
                            const void* grad_out, float* grad_weight, float* grad_bias, float scale,
                            const sdb_dcn_geom* g, int io_dtype, int math, void* workspace,
                            size_t workspace_bytes, const void* x_packed, void* stream);

/* ---- whole-head calls: every FPN level x every deformable convolution in ONE launch per kernel --------------------
 * The reference calls its native op once per (level, branch) from a Python loop (reppointsv2.py:728-752:
 * 5 levels x {cls, refine} = 10 forward and 20 backward native calls per step, each re-laying-out the same weights).
 * Here a call takes a TABLE of problems that share the channel counts and the kernel geometry (C_in, C_out, kH, kW,
 * stride, padding, dilation of `g`; g->N/H/W are ignored) and differ in N, H, W and their tensors.  Every kernel runs
 * once over the concatenated 128-pixel tiles of all problems (persistent grid, tile scheduler over the table), so the
 * small pyramid levels fill the machine together instead of occupying 2..66 SMs each.
 *   - weights are re-laid-out into the kernels' operand images ONCE per weight version (sdb_dcn_prepare_weights),
 *     not once per call; a NULL `prepared` makes the call do it in its workspace;
 *   - backward packs grad_out once and runs grad_offset / grad_mask, grad_input and grad_weight / grad_bias from it
 *     (the reference's v2 entry point is one call too, deform_conv.h:314-375);
 *   - problems with the same non-negative `offset_group` sample with the SAME offset (and mask) tensors -- the two
 *     DCNs of a RepPoints level consume one dcn_offset (reppointsv2.py:744-748): the transposed sampling index of
 *     grad_input is built once per group.
 * Semantics: out, grad_offset, grad_mask AND grad_x are overwritten; grad_weight / grad_bias (float32) are
 * accumulated into with `scale` (fold 1/world_size in here for data-parallel training).  Up to SDB_MAX_PROBLEMS
 * problems and SDB_MAX_WEIGHTS weight tensors per call.  SDB_MATH_FP32 loops over the problems with the SIMT kernels. */
#define SDB_MAX_PROBLEMS 16
#define SDB_MAX_WEIGHTS 4
typedef struct {
  int32_t N, H, W;          /* input [N, C_in, H, W] of this problem                                              */
  int32_t weight_id;        /* index into the weights table: the convolution this problem belongs to             */
  int32_t offset_group;     /* >= 0: shares offset / mask (same pointers, same N, H, W) with equal values; -1: own */
  int32_t reserved;
  const void* x;            /* [N, C_in, H, W]                                                                    */
  const float* offset;      /* [N, 2*kH*kW, Ho, Wo].  NULL in EVERY problem of a table (also in the size query) = plain
                               convolution, the zero-offset specialisation used for the head towers (reppointsv2.py:
                               644-675): SDB_MATH_BF16 only, no mask, stride 1 and 'same' padding (2*pad == dil*(k-1));
                               one input row per tap is loaded instead of four, grad_x = the forward kernel on grad_out
                               with the transposed tap-reversed weights (needs C_out % 64 == 0, C_in % 16 == 0)       */
  const float* mask;        /* [N, kH*kW, Ho, Wo] or NULL (v1)                                                    */
  void* out;                /* forward:  [N, C_out, Ho, Wo]                                                       */
  void* x_packed;           /* optional NHWC-bf16 copy of x (sdb_dcn_packed_input_bytes): forward fills it,
                               backward reads it instead of re-packing x; NULL = in the workspace                */
  const void* grad_out;     /* backward: [N, C_out, Ho, Wo]                                                       */
  void* grad_x;             /* backward: [N, C_in, H, W], overwritten; NULL = not wanted                         */
  float* grad_offset;       /* backward: overwritten; NULL = not wanted                                           */
  float* grad_mask;         /* backward: overwritten; NULL = not wanted                                           */
  void* columns;            /* optional buffer of sdb_dcn_columns_bytes: forward saves the sampled columns (the bf16
                               `columns` of deform_conv_cuda.cu:345-352, as tensor-core operand tiles) in it, backward
                               streams them into the weight-gradient GEMM instead of sampling x again; NULL = not
                               saved / sampled again.  Valid for the offsets / mask / x of the forward call only.    */
} sdb_dcn_problem;
typedef struct {
  const void* weight;       /* [C_out, C_in, kH, kW]                                                              */
  const void* bias;         /* [C_out] or NULL                                                                    */
  const void* prepared;     /* operand images from sdb_dcn_prepare_weights, or NULL                               */
  float* grad_weight;       /* backward: float32 [C_out, C_in, kH, kW], accumulated into; NULL = not wanted       */
  float* grad_bias;         /* backward: float32 [C_out], accumulated into; NULL = not wanted                     */
} sdb_dcn_weights;

/* Bytes of / fill the operand images of one weight tensor (0 / no-op for SDB_MATH_FP32).  Call again whenever the
 * weight values change (once per optimiser step); the images are read-only inputs of the calls below. */
size_t sdb_dcn_prepared_weight_bytes(const sdb_dcn_geom* g, int io_dtype, int math);
int sdb_dcn_prepare_weights(const void* weight, const void* bias, const sdb_dcn_geom* g, int io_dtype, int math,
                            void* prepared, void* stream);
/* Scratch bytes of a forward (backward == 0) or backward (backward != 0) call over this table (0 on error). */
size_t sdb_dcn_multi_workspace_bytes(const sdb_dcn_problem* problems, int32_t n_problems, const sdb_dcn_weights* weights,
                                     int32_t n_weights, const sdb_dcn_geom* g, int io_dtype, int math, int backward);
/* Replaces the reference's loop of deform_conv_forward / modulated_deform_conv_forward calls over levels and branches. */
int sdb_dcn_forward_multi(const sdb_dcn_problem* problems, int32_t n_problems, const sdb_dcn_weights* weights,
                          int32_t n_weights, const sdb_dcn_geom* g, int io_dtype, int math, void* workspace,
                          size_t workspace_bytes, void* stream);
/* Replaces the loop of deform_conv_backward_input + deform_conv_backward_filter (or modulated_deform_conv_backward).
 * `flags` (0 = everything in one go) lets a data-parallel trainer start the weight-gradient all-reduce early:
 *   SDB_BWD_WEIGHT_ONLY  only grad_weight / grad_bias (first call: packs grad_out into the workspace);
 *   SDB_BWD_DATA_ONLY    only grad_x / grad_offset / grad_mask;
 *   SDB_BWD_GRAD_PACKED  the workspace still holds the packed grad_out of a previous call on the SAME table
 *                        (second call: SDB_BWD_DATA_ONLY | SDB_BWD_GRAD_PACKED while NCCL reduces the weight grads);
 *   SDB_BWD_NO_GATHER    grad_offset / grad_mask (and, unless DATA_ONLY, the weight gradients) now, the grad_x gather later:
 *                        the dcol tiles and the transposed index stay in the workspace;
 *   SDB_BWD_GATHER_ONLY  only the grad_x gather, from the workspace a SDB_BWD_NO_GATHER call on the SAME table left.
 *                        (Lets a caller place the gather -- an ordinary, non-persistent grid -- wherever its schedule wants it;
 *                        running the all-reduce beside the gather alone was measured slower than beside grad_offset + gather.);
 *   SDB_BWD_BUILD_INDEX  with SDB_BWD_WEIGHT_ONLY: also build the transposed sampling index of the problems that name a grad_x
 *                        (beside the weight-gradient GEMM, where it costs nothing; beside the statically scheduled
 *                        grad_offset kernel it costs ~35 us) and leave it in the workspace;
 *   SDB_BWD_INDEX_READY  with SDB_BWD_DATA_ONLY: use that index instead of building one.
 *                        (For callers whose schedule has room before the collective; in bench.py's two-half step it delays
 *                        the start of the all-reduce and is off by default.) */
enum { SDB_BWD_WEIGHT_ONLY = 1, SDB_BWD_DATA_ONLY = 2, SDB_BWD_GRAD_PACKED = 4, SDB_BWD_NO_GATHER = 8, SDB_BWD_GATHER_ONLY = 16,
       SDB_BWD_BUILD_INDEX = 32, SDB_BWD_INDEX_READY = 64 };
int sdb_dcn_backward_multi(const sdb_dcn_problem* problems, int32_t n_problems, const sdb_dcn_weights* weights,
                           int32_t n_weights, const sdb_dcn_geom* g, int io_dtype, int math, float scale, int flags,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- label assignment (HBM/latency-bound; no tensor cores) ------------------------------------
 * Fuses pairwise_iou (d2/structures/boxes.py:316-348) with Matcher.__call__
 * (d2/modeling/matcher.py:61-126) or TopKMatcher.__call__ (sd/modeling/matchers/topk_matcher.py:38-86).
 * gt [M,4], anchors [X,4] float32 xyxy.  thresholds[n_thresholds] ascending, labels[n_thresholds+1]
 * in {-1,0,1}.  topk > 0: TopKMatcher (every GT's k best anchors get label 1; ties -> lowest anchor
 * index); topk == 0: Matcher, with allow_low_quality != 0 enabling set_low_quality_matches_.
 * Outputs: matches int64 [X], match_labels int8 [X]; M == 0 gives zeros / labels[0].
 * iou_out (optional, [M,X] float32) receives the IoU matrix, bit-identical to pairwise_iou.
 * workspace: sdb_assign_workspace_bytes(M, X, topk). */
size_t sdb_assign_workspace_bytes(int32_t M, int32_t X, int32_t topk);
int sdb_iou_assign(const float* gt, const float* anchors, int32_t M, int32_t X,
                   const float* thresholds, const int8_t* labels, int32_t n_thresholds, int32_t topk,
                   int32_t allow_low_quality, int64_t* matches, int8_t* match_labels, float* iou_out,
                   void* workspace, size_t workspace_bytes, void* stream);
/* Same, on a caller-provided match-quality matrix q [M,X] float32 (the matchers' own signature). */
int sdb_match_quality_assign(const float* q, int32_t M, int32_t X, const float* thresholds,
                             const int8_t* labels, int32_t n_thresholds, int32_t topk,
                             int32_t allow_low_quality, int64_t* matches, int8_t* match_labels,
                             void* workspace, size_t workspace_bytes, void* stream);
/* pairwise_iou alone: iou [N1,N2] float32. */
int sdb_pairwise_iou(const float* boxes1, const float* boxes2, int32_t N1, int32_t N2, float* iou,
                     void* stream);

/* ---- head losses (HBM-bound; fused forward + gradient, sum reduction) --------------------------
 * sigmoid focal loss on logits [R,K] with a class index per row instead of the dense one-hot
 * target the reference builds (sd/modeling/meta_arch/reppoints/reppointsv2.py:294-312,
 * fcos/fcos.py:289-297; formula = fvcore sigmoid_focal_loss_jit, reduction "sum").
 * class_idx[r] in [0,K) marks the foreground class, any other value is background.
 * loss_sum (float32 scalar) is ACCUMULATED into; grad_logits (may be NULL) = grad_scale * dLoss/dx. */
int sdb_sigmoid_focal_loss(const float* logits, const int64_t* class_idx, int64_t R, int32_t K,
                           float alpha, float gamma, float grad_scale, float* loss_sum,
                           float* grad_logits, void* stream);

typedef enum {
  SDB_LOSS_IOU = 0,        /* -log(iou)      sd/layers/iou_loss.py:24-25 */
  SDB_LOSS_LINEAR_IOU = 1, /* 1 - iou        :26-27 */
  SDB_LOSS_GIOU = 2,       /* 1 - giou       :28-29 */
  SDB_LOSS_SMOOTH_L1 = 3,  /* sd/layers/smooth_l1_loss_with_weight.py:3-17 (beta) */
  SDB_LOSS_GIOU_FVCORE = 4 /* fvcore giou_loss, xyxy, eps 1e-7 (meta/heads/anchor_head.py:369-376) */
} sdb_box_loss_kind;
typedef enum { SDB_BOX_LTRB = 0 /* iou_loss */, SDB_BOX_XYXY = 1 /* box_iou_loss */ } sdb_box_form;

/* pred, target [R,4] float32; weight [R] float32 or NULL.  loss_sum ACCUMULATED; grad_pred
 * (may be NULL, [R,4]) = grad_scale * dLoss/dpred. */
int sdb_box_reg_loss(const float* pred, const float* target, const float* weight, int64_t R,
                     int kind, int form, float beta, float grad_scale, float* loss_sum,
                     float* grad_pred, void* stream);

/* compute_centerness_targets (sd/modeling/meta_arch/fcos/utils.py:295-300):
 * out[r] = sqrt( min(l,r)/max(l,r) * min(t,b)/max(t,b) ) for reg_targets [R,4] float32 in (l,t,r,b) order;
 * one rounding per operation in the reference's order, so the result is bit-identical to torch on CPU. */
int sdb_centerness_targets(const float* reg_targets, int64_t R, float* out, void* stream);
/* The FCOSRepPoints module's OWN compute_centerness_targets (fcos/fcos_rpd_s1_topk.py:25-55), which shadows the one
 * above inside that model (losses :288, :291; top-5 selection :117):  out[r] = pow(c, min(w/h, h/w)) with
 * c = min(l,r)/max(l,r) * min(t,b)/max(t,b), w = l + r, h = t + b -- the slender-object exponent. */
int sdb_slender_centerness_targets(const float* reg_targets, int64_t R, float* out, void* stream);

/* ---- GroupNorm + ReLU of the head towers ---------------------------------------
 * The towers are stacks of Conv2d(3x3, bias=False) -> GroupNorm(32, C) -> ReLU(inplace)
 * (sd/modeling/meta_arch/reppoints/reppointsv2.py:644-675, applied per FPN level :733-736; fcos/fcos.py:494-538).
 * One call normalises every tensor of a table (FPN levels x towers): y = relu(group_norm(x, G, gamma, beta, eps)) with
 * torch.nn.functional.group_norm semantics (biased variance over (C/G, H, W) per image, eps inside the square root).
 * Tensors NCHW contiguous, io_dtype float32 or bfloat16; gamma / beta / their gradients float32 [C]; `stats`
 * [N, G, 2] float32 = (mean, rstd), written by the forward and read by the backward.  Backward: grad_x is
 * overwritten (may be NULL), grad_gamma / grad_beta are ACCUMULATED into (summed over images and over every tensor
 * that names the parameter set, in a fixed order); the ReLU mask is recomputed from x.  `relu` == 0 gives plain
 * GroupNorm.  Workspace: sdb_gn_relu_workspace_bytes (same size for both directions). */
typedef struct {
  const void* x;        /* [N, C, H, W] input of the normalisation (the convolution's output) */
  void* y;              /* forward: output */
  const void* grad_y;   /* backward: gradient w.r.t. y */
  void* grad_x;         /* backward: gradient w.r.t. x, or NULL */
  float* stats;         /* [N, G, 2] */
  int32_t N, HW;        /* batch, H * W */
  int32_t param_id;     /* which (gamma, beta) of the parameter table */
  int32_t reserved;
} sdb_gn_tensor;
typedef struct {
  const float* gamma;
  const float* beta;
  float* grad_gamma;    /* backward, may be NULL */
  float* grad_beta;
} sdb_gn_params;
#define SDB_GN_MAX_TENSORS 16
#define SDB_GN_MAX_PARAMS 4
size_t sdb_gn_relu_workspace_bytes(const sdb_gn_tensor* tensors, int32_t n, int32_t C, int32_t G);
int sdb_gn_relu_forward(const sdb_gn_tensor* tensors, int32_t n, const sdb_gn_params* params, int32_t n_params, int32_t C,
                        int32_t G, float eps, int32_t relu, int io_dtype, void* workspace, size_t workspace_bytes,
                        void* stream);
int sdb_gn_relu_backward(const sdb_gn_tensor* tensors, int32_t n, const sdb_gn_params* params, int32_t n_params, int32_t C,
                         int32_t G, float eps, int32_t relu, int io_dtype, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- RepPoints DCN offset construction (first piece) -------------------------
 * dcn_offset = ((1 - gm) * pts.detach() + gm * pts) - dcn_base_offset   (reppointsv2.py:638-642, 742-744;
 * rpd.py:105-110, 624-635), fused into one pass.  pts, out: [N, 2*K, H, W] float32, K = ks*ks points;
 * dcn_base_offset[2k] = k / ks - (ks-1)/2 (y), [2k+1] = k % ks - (ks-1)/2 (x).  flip_xy != 0 first swaps the
 * two channels of every point ((x,y) -> (y,x), rpd.py:628-634); reppointsv2.py does not flip.  The four
 * float32 roundings of the reference expression are kept, so the result is bit-identical to it.
 * Backward: grad_pts = gm * grad_out (with the same channel swap). */
int sdb_reppoints_dcn_offset(const float* pts, int32_t N, int32_t ks, int32_t H, int32_t W, float gradient_mul,
                             int32_t flip_xy, float* out, void* stream);
int sdb_reppoints_dcn_offset_backward(const float* grad_out, int32_t N, int32_t ks, int32_t H, int32_t W,
                                      float gradient_mul, int32_t flip_xy, float* grad_pts, void* stream);

/* ---- RepPoints point assignment (reppointsv2.py:370-428) --------------------
 * Every GT claims the point nearest to its centre (L2 of (point - centre) / (w, h)) among the points of ITS
 * pyramid level, level = clamp(int((log2(w/scale) + log2(h/scale)) / 2), min/max level of the points); a point
 * claimed by several GTs keeps the closest one, ties -> the lowest GT index (the reference's sequential
 * "strictly less than recorded" rule).  The reference loops over GTs on the host (~12 launches each); here:
 * one block per GT + one pass over the points.
 * points [X,2], strides [X] float32; gt [M,4] xyxy float32; gt_labels [M] int64.
 * assigned_bboxes [X,4] float32 (zeros where unassigned), assigned_labels [X] int64 (num_classes where
 * unassigned).  workspace: sdb_point_targets_workspace_bytes(X). */
size_t sdb_point_targets_workspace_bytes(int32_t X);
int sdb_point_targets(const float* points, const float* strides, const float* gt, const int64_t* gt_labels,
                      int32_t X, int32_t M, float scale, int64_t num_classes, float* assigned_bboxes,
                      int64_t* assigned_labels, void* workspace, size_t workspace_bytes, void* stream);

/* ---- FCOS location targets (fcos/utils.py:108-212), one image per call -----
 * For every location: ltrb to every GT, "inside" = min(ltrb) > 0 (center_sampling_radius <= 0) or inside the
 * GT's centre region of radius stride*radius clipped to the box (get_sample_region :108-157, including its
 * "first GT centred at x == 0 means no GT" shortcut), "cared" = size_lo <= max(ltrb) <= size_hi; among the
 * qualifying GTs the one of minimal area (lowest index on ties; INF = 1e8 as in the reference).
 * out_classes[x] = class of that GT or num_classes; out_reg[x] = ltrb to that GT (GT 0 for background rows,
 * as the reference's argmin of an all-INF row gives).  The [X, M, 4] temporaries never exist.
 * locations, sizes_of_interest [X,2] float32; gt [M,4] xyxy float32; gt_classes [M] int64;
 * num_points_per_level / level_strides: HOST arrays of n_levels (<= 8) entries. */
int sdb_fcos_location_targets(const float* locations, const float* sizes_of_interest, const float* gt,
                              const int64_t* gt_classes, int32_t X, int32_t M, const int32_t* num_points_per_level,
                              const float* level_strides, int32_t n_levels, float center_sampling_radius,
                              int64_t num_classes, int64_t* out_classes, float* out_reg, void* stream);

/* compute_topk_targets_for_locations (fcos/utils.py:215-292; the active FCOSRepPoints model's stage 1), one
 * image per call: the targets above plus out_topk[x] (uint8 0/1) = location x is among the `topk` foreground
 * locations of ITS GT with the highest centerness (all of them when the GT has <= topk; ties -> lowest location
 * index).  Regression targets are NOT stride-normalised here (the caller divides, as :284-285 does after the
 * selection).  The reference loops over GTs on the host with a `.sum().item()` sync per GT.
 * workspace: sdb_fcos_topk_workspace_bytes(X); topk <= 16. */
size_t sdb_fcos_topk_workspace_bytes(int32_t X);
int sdb_fcos_topk_location_targets(const float* locations, const float* sizes_of_interest, const float* gt,
                                   const int64_t* gt_classes, int32_t X, int32_t M,
                                   const int32_t* num_points_per_level, const float* level_strides, int32_t n_levels,
                                   float center_sampling_radius, int64_t num_classes, int32_t topk,
                                   int64_t* out_classes, float* out_reg, uint8_t* out_topk, void* workspace,
                                   size_t workspace_bytes, void* stream);

/* Both of the above for a whole batch in ONE call (no per-image host loop): image n uses GT rows
 * gt[n*M_pad .. n*M_pad + gt_counts[n]) of the padded gt [n_images, M_pad, 4] / gt_classes [n_images, M_pad];
 * gt_counts is a DEVICE int32 [n_images] (0 allowed: the image is all background, regression rows measured from a
 * zero box).  topk == 0: no top-k mask (out_topk may be NULL).  centerness_kind: 0 = sqrt form (fcos/utils.py:295-300),
 * 1 = the FCOSRepPoints module's pow form (fcos_rpd_s1_topk.py:25-55, used by ITS top-5 loop :110-121).
 * out_classes [n_images, X], out_reg [n_images, X, 4], out_topk [n_images, X].
 * workspace: n_images * sdb_fcos_topk_workspace_bytes(X) when topk > 0. */
int sdb_fcos_location_targets_batched(const float* locations, const float* sizes_of_interest, const float* gt,
                                      const int64_t* gt_classes, const int32_t* gt_counts, int32_t n_images, int32_t X,
                                      int32_t M_pad, const int32_t* num_points_per_level, const float* level_strides,
                                      int32_t n_levels, float center_sampling_radius, int64_t num_classes, int32_t topk,
                                      int32_t centerness_kind, int64_t* out_classes, float* out_reg, uint8_t* out_topk,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- inference post-processing of a dense point head ---------------------------------------
 * Replaces RepPointsV2.inference / inference_single_image (reppointsv2.py:486-603) for a whole batch, with no host
 * round trip between its steps: per (image, level) score = sigmoid(logit) over the H*W*K cells, the `topk` best cells
 * with score > score_thresh, their boxes decoded from the refined point sets (pts_to_bbox :328-366: transform 0 =
 * "minmax", 1 = "partial_minmax" (first 4 points), 2 = "moment" (mean +- std * exp(moment_transfer[0|1])); x stride,
 * + centre, clamped to [0, w] x [0, h] :558-564); per image the candidates of all levels go through class-aware
 * greedy NMS (IoU > nms_thresh suppresses; detectron2/layers/nms.py:10-29) and the first `max_det` survivors are
 * written in decreasing score order.  Equal scores: lower flat cell index / lower level first (the reference's
 * torch.sort leaves ties implementation-defined).
 *   levels[l].cls [N, K, H, W], .pts [N, 2*num_points, H, W] (the head's NCHW outputs, read in place),
 *   .centers [H*W, 2] (x, y) float32 device tensors; image_sizes: HOST int32 [N][2] = (height, width);
 *   out_boxes [N, max_det, 4], out_scores [N, max_det] float32, out_classes [N, max_det] int64, out_count [N] int32
 *   (device).  out_overflow (device int32, optional): set when more than 4096 cells of one (image, level) tie in the
 *   top 19 bits of their score with the topk-th cell -- the cut is then approximate (never seen on real heads).
 * Limits: n_levels <= 8, n_images <= 64, topk <= 2048, n_levels * topk <= 8192. */
typedef struct {
  const float* cls;
  const float* pts;
  const float* centers;
  int32_t H, W;
  float stride;
} sdb_pp_level;
size_t sdb_points_postprocess_workspace_bytes(int32_t n_levels, int32_t n_images, int32_t topk, float score_thresh);
int sdb_points_postprocess(const sdb_pp_level* levels, int32_t n_levels, int32_t n_images, int32_t num_classes,
                           int32_t num_points, int32_t transform, const float* moment_transfer, const int32_t* image_sizes,
                           float score_thresh, int32_t topk, float nms_thresh, int32_t max_det, float* out_boxes,
                           float* out_scores, int64_t* out_classes, int32_t* out_count, int32_t* out_overflow,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- diagnostics ------------------------------------------------------------------------------
 * Per-kernel timing for bench.py's roofline line.  While enabled, each DCN entry point records a
 * CUDA event pair on ITS stream around its dominant kernel only (the tcgen05 / SIMT main kernel,
 * not the packing helpers).  slot = sdb_dcn_op, where SDB_OP_BACKWARD_DATA (1) is the grad_offset /
 * grad_mask kernel and slot 3 is the grad_input kernel of the same entry point.
 * sdb_profile_read synchronises the recorded
 * events and returns the summed milliseconds and the launch count since the last reset.
 * Not capturable into CUDA graphs; leave disabled (the default) on the training path. */
int sdb_profile_enable(int on);
int sdb_profile_reset(void);
int sdb_profile_read(int slot, float* total_ms, int* launches);
/* Number of kernels this library has launched (or captured into a CUDA graph) since it was loaded. */
long long sdb_launch_count(void);
/* SMs the persistent tensor-core kernels leave free (default 0; process-wide, applies to launches made -- or captured
 * into a CUDA graph -- after the call).  Their CTAs walk a static tile schedule, so a CTA that must wait for an SM held
 * by a co-running kernel delays the whole launch by that wait; a data-parallel caller that overlaps the weight-gradient
 * all-reduce (d2/engine/defaults.py:280-283) with the data-gradient kernels reserves as many SMs as the collective
 * uses CTAs (bench.py: NCCL max_ctas). */
int sdb_set_sm_reserve(int n_sms);
/* The forward kernel (deformable and plain convolution) runs as clusters of two CTAs issuing tcgen05.mma.cta_group::2
 * (M = 256: each CTA gathers its own 128 output pixels and streams half of every weight tile) when C_in % 128 == 0 and
 * C_out % 32 == 0.  on == 0 selects the one-CTA kernel (M = 128) instead -- results are identical; process-wide. */
int sdb_set_forward_pair(int on);
/* The same for the grad_offset / dcol kernel of the backward (C_in % 128 == 0). */
int sdb_set_backward_pair(int on);

#ifdef __cplusplus
}
#endif
#endif /* SLENDER_B200_H_ */
