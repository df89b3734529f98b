/* vnet_b200.h -- C ABI of libvnet_b200.so, the B200-native V-Net engine.
 *
 * The reference (jackyko1991/vnet-tensorflow) has no FFI seam: its hot path is whatever
 * `sess.run(...)` executes inside TensorFlow.  This ABI is the seam a maintainer binds instead
 * (ctypes stub: INTEGRATION.md).  Each entry point names the reference interface it replaces
 * (paths relative to the reference tree):
 *
 *   vnb_create            networks.VNet(...).GetNetwork + build_model_graph     networks.py:209-305, model.py:297-630
 *   vnb_set/get_param     tf.train.Saver restore/save by variable name          model.py:689-699,758-764
 *   vnb_forward           sess.run(['predicted_label/prediction:0','softmax:0']) model.py:914-917
 *   vnb_evaluate_volume   evaluate_single_3D window loop + accumulation + argmax model.py:866-937
 *   vnb_loss              sess.run([summary_op, loss_op]) test step             model.py:784-789
 *   vnb_train_step        sess.run([train_op, summary_op, loss_op])             model.py:743-748
 *   vnb_forward_backward  the gradient half of optimizer.minimize               model.py:660
 *   vnb_apply_gradients   the apply half of optimizer.minimize (+ lr decay)     model.py:641-660
 *   vnb_comm_*            (absent in the reference: single device)              SURVEY.md 2.1
 *
 * Conventions
 *   - All tensors cross the boundary as caller-owned, C-contiguous host buffers in the reference's
 *     layouts: images float32 [N,X,Y,Z,M], labels int32 [N,X,Y,Z] (class indices), logits/softmax
 *     float32 [N,X,Y,Z,K], argmax int64 [N,X,Y,Z]; conv filters [kd,kh,kw,Cin,Cout], transposed-conv
 *     filters [kd,kh,kw,Cout,Cin] (layers2.py:60,66,92).
 *   - The library owns all device memory behind the opaque handle (allocated once for max_batch).
 *   - Every call returns VNB_OK (0) or a negative vnb_status; nothing throws across the boundary;
 *     the message is available from vnb_last_error() (thread local).
 *   - One handle <-> one GPU <-> one host thread.  Data parallelism = one process per GPU.
 *   - There is no CPU fallback: vnb_create fails with VNB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef VNET_B200_H
#define VNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vnb_handle vnb_handle;

typedef enum vnb_status {
  VNB_OK = 0,
  VNB_ERR_INVALID_ARG = -1,
  VNB_ERR_SHAPE = -2,
  VNB_ERR_CUDA = -3,
  VNB_ERR_NCCL = -4,
  VNB_ERR_OOM = -5,
  VNB_ERR_UNSUPPORTED = -6,
  VNB_ERR_INTERNAL = -7
} vnb_status;

/* TrainingSetting.Precision (new optional config key) */
enum { VNB_PREC_FP32 = 0, /* fp32 FMA on CUDA cores, exact-precision parity mode            */
       VNB_PREC_BF16X3 = 1, /* tcgen05 bf16 split (hi/lo) x3, fp32 accumulate: fp32-grade      */
       VNB_PREC_BF16 = 2 }; /* tcgen05 bf16, fp32 accumulate                                    */

/* TrainingSetting.Loss.Name (model.py:495-560), same order as the reference's if-chain */
enum { VNB_LOSS_XENT = 0, VNB_LOSS_WEIGHTED_XENT, VNB_LOSS_SORENSEN, VNB_LOSS_WEIGHTED_SORENSEN,
       VNB_LOSS_JACCARD, VNB_LOSS_WEIGHTED_JACCARD, VNB_LOSS_MIXED_SORENSEN,
       VNB_LOSS_MIXED_WEIGHTED_SORENSEN, VNB_LOSS_MIXED_JACCARD, VNB_LOSS_MIXED_WEIGHTED_JACCARD,
       VNB_LOSS_SORENSEN_FG /* legacy train.py:373-377 --loss_function sorensen: Dice of softmax[...,1] vs the label volume */ };

/* legacy --attention_loss_function (train.py:387-399); needs vnb_set_distmap before a loss / training call */
enum { VNB_ATT_NONE = 0, VNB_ATT_L2 = 1, VNB_ATT_ABS = 2 };

/* TrainingSetting.Optimizer.Name (model.py:649-658) */
enum { VNB_OPT_ADAM = 0, VNB_OPT_SGD = 1, VNB_OPT_MOMENTUM = 2, VNB_OPT_NESTEROV = 3 };

/* which per-variable buffer vnb_set_slot / vnb_get_slot address */
enum { VNB_SLOT_VALUE = 0, VNB_SLOT_GRAD = 1, VNB_SLOT_ADAM_M = 2, VNB_SLOT_ADAM_V = 3 };

typedef struct vnb_config {
  int32_t in_channels;          /* len(Data.ImageFilenames)            model.py:189 */
  int32_t num_classes;          /* len(SegmentationClasses)            model.py:190 */
  int32_t num_channels;         /* Networks.NumChannel                 model.py:212 */
  int32_t num_levels;           /* Networks.NumLevels                  model.py:213 */
  int32_t num_convolutions[8];  /* Networks.NumConvolutions            model.py:214 */
  int32_t bottom_convolutions;  /* Networks.BottomConvolutions         model.py:215 */
  int32_t patch_shape[3];       /* PatchShape [X,Y,Z]                  model.py:198 */
  int32_t max_batch;            /* max(BatchSize, Evaluation BatchSize) model.py:197,234 */
  int32_t precision;            /* VNB_PREC_*                                          */
  int32_t loss;                 /* VNB_LOSS_*                          model.py:222 */
  float loss_weights[8];        /* Loss.Weights                        model.py:223 */
  float loss_alpha;             /* Loss.Alpha                          model.py:224 */
  int32_t optimizer;            /* VNB_OPT_*                           model.py:217 */
  float learning_rate;          /* Optimizer.InitialLearningRate       model.py:218 */
  float decay_factor;           /* Optimizer.Decay.Factor              model.py:219 */
  float decay_steps;            /* Optimizer.Decay.Steps               model.py:220 */
  float momentum;               /* Optimizer.Momentum (Momentum / NesterovMomentum) model.py:653-656 */
  int32_t graph_flavour;        /* 0 = networks.VNet (main.py path), 1 = VNet.py legacy flavour (train.py:271-279) */
  int32_t attention;            /* 1 = --attention: AttentionModule -> (1+softmax)*logits gating -> OutputModule
                                   (train.py:281-312, attention.py:105-114, OutputModule.py:105-114) */
  int32_t attention_loss;       /* VNB_ATT_*                           train.py:387-399 */
  int32_t module_channels;      /* attention.py:41 / OutputModule.py:41 num_channels (0 = 64) */
} vnb_config;

const char* vnb_last_error(void);
const char* vnb_version(void);

int vnb_create(const vnb_config* cfg, int device, vnb_handle** out);
int vnb_destroy(vnb_handle* h);

/* variable inventory, TF names in creation order (vnet/encoder/level_1/conv_1/weights, ...) */
int vnb_num_params(vnb_handle* h, int* count);
int vnb_param_info(vnb_handle* h, int index, const char** tf_name, int* ndim, int64_t dims[5], int* trainable);
int vnb_set_param(vnb_handle* h, const char* tf_name, const void* host, size_t bytes);
int vnb_get_param(vnb_handle* h, const char* tf_name, void* host, size_t bytes);
int vnb_set_slot(vnb_handle* h, const char* tf_name, int slot, const void* host, size_t bytes);
int vnb_get_slot(vnb_handle* h, const char* tf_name, int slot, void* host, size_t bytes);
int vnb_get_step(vnb_handle* h, int64_t* global_step);
int vnb_set_step(vnb_handle* h, int64_t global_step);

/* Batch buffers - `images`, `labels` (and the distance map of vnb_set_distmap) going in, logits / softmax / argmax of
 * vnb_forward coming out - are caller-owned and may live in HOST memory or in DEVICE memory of the handle's GPU (the
 * SURVEY 8(b) "*_dev" variants: the direction is resolved through unified addressing, a batch produced on the GPU is
 * copied device to device).  The copies run on the handle's own stream: a device buffer must be complete before the
 * call (synchronise the stream that produced it), and device outputs are complete when vnb_forward returns.
 * Scalars (loss_out, dice_terms) and parameter buffers are host memory. */
/* inference: any of logits / softmax / argmax may be NULL */
int vnb_forward(vnb_handle* h, const float* images, int n, float* logits, float* softmax, int64_t* argmax);
/* sliding-window evaluation of one padded case (model.py:866-937): volume [X][Y][Z][M] float32 with every extent >=
 * the patch extent; windows in (i,j,k) order, last one clamped, batches of `batch` consecutive windows (batch norm
 * uses each batch's own statistics); outputs (any may be NULL): label int64 [X][Y][Z] = argmax of the un-normalised
 * softmax sums, softmax_sum [X][Y][Z][K], weight [X][Y][Z] = number of windows covering a voxel */
int vnb_evaluate_volume(vnb_handle* h, const float* volume, const int32_t dims[3], const int32_t stride[3], int batch,
                        int64_t* label, float* softmax_sum, float* weight);
/* loss without update; dice_terms (optional) receives [n][K][4] = (I, L, R, xent-sum) per sample/class */
int vnb_loss(vnb_handle* h, const float* images, const int32_t* labels, int n, float* loss_out, double* dice_terms);
/* step metrics, the tf.metrics block of summary_op (model.py:586-626) that the reference fetches with every training
 * step (model.py:743-748), for the batch of the last vnb_loss / vnb_forward_backward / vnb_train_step[_resident] call
 * (n = its batch size; the logits are those of that call's forward pass, i.e. before the weight update).  Integer counts:
 *   confusion [(K+1)][K]   rows = label class (row K: labels outside [0,K), whose one-hot row is all zero), columns =
 *                          argmax class; accuracy and the true/false positive/negative counts of every class follow
 *   auc_hist  [K][2][201]  (optional, may be NULL) per class c >= 1 and truth value (label != c, label == c) the
 *                          histogram of "number of tf.metrics.auc thresholds below softmax[c]" (200 thresholds)
 * vnet_tensorflow_b200/metrics.py turns them into accuracy / sensitivity / specificity / dice / auc as TF computes them. */
#define VNB_AUC_BINS 201
int vnb_read_metrics(vnb_handle* h, int n, uint64_t* confusion, uint64_t* auc_hist);
/* one optimiser step; loss_out may be NULL (then the call does not synchronise) */
int vnb_train_step(vnb_handle* h, const float* images, const int32_t* labels, int n, float dropout_rate,
                   uint64_t seed, float* loss_out);
/* gradients only (left in the VNB_SLOT_GRAD buffers), then the apply half */
int vnb_forward_backward(vnb_handle* h, const float* images, const int32_t* labels, int n, float dropout_rate,
                         uint64_t seed, int update_moving_stats, float* loss_out);
int vnb_apply_gradients(vnb_handle* h);

/* data parallel: one process per GPU; rank 0 creates the id and shares it out of band */
int vnb_comm_unique_id(void* id_out_128_bytes);
int vnb_comm_init(vnb_handle* h, int rank, int world, const void* unique_id_128_bytes);
int vnb_comm_world(vnb_handle* h, int* rank, int* world);
/* synchronised batch norm: batch statistics (sum z, sum z^2) and the two backward sums of every
 * training-mode batch norm are summed over the ranks, which makes `world` ranks with local batch n reproduce one
 * device with batch world*n (the reference's own arithmetic at that batch size).  Off by default: local statistics,
 * no extra exchange.  vnb_comm_sync_bn uses a second NCCL communicator on the compute stream; the callback form
 * hands each [n] double row to the caller on the host (used with gloo in the CPU tests; NULL switches it off). */
int vnb_comm_sync_bn(vnb_handle* h, int on);
typedef int (*vnb_allreduce_fn)(double* host_values, int n, void* user); /* sum in place over the ranks; 0 = ok */
int vnb_set_stats_allreduce(vnb_handle* h, vnb_allreduce_fn fn, void* user, int world);

/* device-resident stepping (benchmark 'value' leg): upload once, then step on the resident batch */
int vnb_upload_batch(vnb_handle* h, const float* images, const int32_t* labels, int n);
int vnb_train_step_resident(vnb_handle* h, int n, float dropout_rate, uint64_t seed);
/* input staging (the feed_dict copy of model.py:743-748 taken off the critical path): vnb_stage_batch enqueues the
 * copy of the NEXT batch into staging buffers on a copy stream and returns; vnb_train_step_staged orders the compute
 * stream after it, moves the batch into the input buffers device to device and runs one optimiser step on it (same
 * arithmetic as vnb_train_step).  Pattern: stage(b0); loop { step_staged(loss_out = NULL); stage(b_next);
 * vnb_read_losses } - the host-to-device copy of b_next then overlaps the step on b.  The copy is asynchronous only
 * from page-locked memory (vnb_host_alloc / vnb_host_free = cudaMallocHost / cudaFreeHost in the context of the handle); such a buffer must stay
 * untouched until the staged step has been issued and a later call synchronised (vnb_read_losses, vnb_sync). */
int vnb_stage_batch(vnb_handle* h, const float* images, const int32_t* labels, int n);
int vnb_train_step_staged(vnb_handle* h, float dropout_rate, uint64_t seed, float* loss_out);
int vnb_host_alloc(vnb_handle* h, size_t bytes, void** out);
int vnb_host_free(void* p);
/* CUDA-event timing on the handle's compute stream: record event 0 / 1, then read the elapsed ms */
int vnb_event_record(vnb_handle* h, int which);
int vnb_event_elapsed_ms(vnb_handle* h, float* ms);
/* per-kernel-class profiling with CUDA events around every convolution launch (5x5x5 fprop/dgrad/wgrad):
 * enable, run steps, then read (sum of device ms, launches, algorithmic FLOPs) for class 0 fprop+dgrad, 1 wgrad */
int vnb_profile_enable(vnb_handle* h, int on);
int vnb_profile_read(vnb_handle* h, int kernel_class, double* ms, int64_t* launches, double* flops);
/* the same records one by one (per-layer roofline table): number of profiled launches since vnb_profile_enable, and
 * launch `index` = its class, device ms, algorithmic FLOPs and a label "scope pass Cin->Cout @DxHxW" (NUL-terminated,
 * truncated to label_bytes); VNB_ERR_INVALID_ARG past the last record */
int vnb_profile_count(vnb_handle* h, int64_t* launches);
int vnb_profile_launch(vnb_handle* h, int64_t index, int* kernel_class, double* ms, double* flops, char* label,
                       size_t label_bytes);

/* attention path (vnb_config.attention = 1): the distance map fed as distmap_placeholder (train.py:176-179,
 * 536) for the next loss / training calls, [n][X][Y][Z] floats in [0,1]; the three loss scalars of the last
 * loss / training call (total_loss_op, loss_op, att_loss_op; train.py:417,351-382,387-399); and
 * softmax_attention of the last forward pass (train.py:288), [n][X][Y][Z][K] */
int vnb_set_distmap(vnb_handle* h, const float* distmap, int n);
int vnb_read_losses(vnb_handle* h, float out_total_seg_att[3]);
/* the two summands of the segmentation loss of the last loss / training call as the reference's summaries name them
 * (model.py:529-530,537-538,545-546,553-554): out[0] = '1.dice' = 1 - dice, out[1] = '2.regularized_xent' =
 * Loss.Alpha * cross entropy (out[0] = 0 for the cross-entropy losses, out[1] = 0 for the pure Dice losses) */
int vnb_read_loss_parts(vnb_handle* h, float out_dice_xent[2]);
int vnb_read_softmax_attention(vnb_handle* h, float* host, size_t bytes, int n);

int vnb_sync(vnb_handle* h);
/* number of kernels this handle has launched so far */
int vnb_gpu_launches(vnb_handle* h, int64_t* count);

/* test / debug hooks ------------------------------------------------------------------------- */
/* copy an intermediate tensor of the last run: kind 0 activation, 1 its gradient (dL/dz for conv
 * units after a backward pass), 2 pre-batch-norm conv output; scope = TF scope of the unit */
int vnb_read_tensor(vnb_handle* h, const char* scope, int kind, float* host, size_t bytes, int n);
/* standalone 5x5x5 convolution ops on host buffers (NDHWC / [125][Cin][Cout]); precision as VNB_PREC_* */
int vnb_op_conv5_fprop(int device, int precision, const float* x, const float* w, const float* bias,
                       const float* residual, float* y, int n, int d, int h, int w_, int cin, int cout);
int vnb_op_conv5_dgrad(int device, int precision, const float* dy, const float* w, float* dx, int n, int d,
                       int h, int w_, int cin, int cout);
int vnb_op_conv5_wgrad(int device, int precision, const float* x, const float* dy, float* dw, int n, int d,
                       int h, int w_, int cin, int cout);
/* the same three ops for the 3x3x3 SAME convolution of the attention / output modules (attention.py:63-92);
 * filters [27][Cin][Cout] */
int vnb_op_conv3_fprop(int device, int precision, const float* x, const float* w, const float* bias,
                       const float* residual, float* y, int n, int d, int h, int w_, int cin, int cout);
int vnb_op_conv3_dgrad(int device, int precision, const float* dy, const float* w, float* dx, int n, int d,
                       int h, int w_, int cin, int cout);
int vnb_op_conv3_wgrad(int device, int precision, const float* x, const float* dy, float* dw, int n, int d,
                       int h, int w_, int cin, int cout);

/* ---- per-op hooks of the remaining hot-path kernels (tests; host buffers in and out) ------------------------------
 * 2x2x2 stride-2 down / transposed-up convolution kernels, layers2.py:65-94 (networks.py:278,292):
 *   op 0 "gather"  : coarse[m][cc] = bias[cc] + sum_{tap,cf} fine[child(m,tap)][cf] * w[tap][cf][cc]   (down fprop, up dgrad)
 *   op 1 "scatter" : fine[child][cf] = bias[cf] + sum_cc coarse[m][cc] * w[tap][cf][cc]                (up fprop, down dgrad)
 *   op 2 "wgrad"   : dw[tap][cf][cc] = sum_m fine[child(m,tap)][cf] * coarse[m][cc]
 * fine [n][2dc][2hc][2wc][cf], coarse [n][dc][hc][wc][cc], w / dw [2][2][2][cf][cc]; `out` is the op's result tensor. */
int vnb_op_k2(int device, int precision, int op, const float* fine, const float* coarse, const float* w, const float* bias,
              float* out, int n, int dc, int hc, int wc, int cf, int cc);
/* tf.layers.batch_normalization(training=True, momentum=.99, epsilon=1e-3) + prelu (networks.py:319, layers2.py:97-99)
 * on a [voxels][c] tensor and its backward pass; alpha / dalpha may be NULL (no activation) */
int vnb_op_bn_fwd(int device, const float* z, const float* gamma, const float* beta, const float* alpha, float* y,
                  double* mean_out, double* var_out, long long voxels, int c);
int vnb_op_bn_bwd(int device, const float* z, const float* dy, const float* gamma, const float* beta, const float* alpha,
                  float* dz, float* dgamma, float* dbeta, float* dalpha, long long voxels, int c);
/* softmax, one-hot, Dice / Jaccard / cross-entropy loss zoo and argmax of model.py:26-92,447,477,495-568 on logits
 * [n][voxels][k] and int32 labels [n][voxels]; `loss` is a VNB_LOSS_* code, terms_out [n][k][4] = (I, L, R, X) */
int vnb_op_softmax_dice_fwd(int device, const float* logits, const int32_t* labels, int n, long long voxels, int k, int loss,
                            const float* weights, float alpha, float* loss_out, float* softmax_out, long long* argmax_out,
                            double* terms_out);
int vnb_op_softmax_dice_bwd(int device, const float* logits, const int32_t* labels, int n, long long voxels, int k, int loss,
                            const float* weights, float alpha, float* dlogits);
/* one tf.train.AdamOptimizer step (model.py:652; epsilon-hat form, beta1 .9, beta2 .999, epsilon 1e-8) in place on flat
 * vectors; t = 1-based step count, lr = exponential_decay value of that step (model.py:649) */
int vnb_op_adam(int device, float* p, const float* g, float* m, float* v, long long count, float lr, long long t);

#ifdef __cplusplus
}
#endif
#endif /* VNET_B200_H */
