// V-Net engine: builds the reference graph (networks.VNet.GetNetwork, networks.py:246-305) as a static
// list of "conv units" over device buffers and runs forward / loss / backward / optimiser steps
// (the sess.run([train_op, loss_op]) of model.py:743-748 and the inference run of model.py:914-917).
//
// A conv unit = { convolution (5^3 | 2^3 down | 2^3 up | 1^3 | tiled input) + bias (+ residual) }
//               -> batch-norm chain (bn_chain.h) -> optional PReLU -> optional dropout.
// Parameters live in one flat fp32 buffer in TF variable-creation order (same names as the reference
// checkpoint) so that the optimiser and the data-parallel all-reduce are single passes.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "bn_chain.h"
#include "conv_ref.cuh"
#include "k2_tiled.cuh"
#include "kernels.cuh"
#include "conv_tc.cuh"
#include "vnb_cuda.h"
#include "wgrad_tc.cuh"
#include "k2_tc.cuh"
#ifndef VNB_EMULATE
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: no link dependency, a no-op unless a profiler injects itself
#endif

namespace vnb {

#define VNB_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " +  \
                               __FILE__ + ":" + std::to_string(__LINE__));                        \
  } while (0)

enum Precision : int { PREC_FP32 = 0, PREC_BF16X3 = 1, PREC_BF16 = 2 };
enum LossName : int {
  LOSS_XENT = 0, LOSS_WEIGHTED_XENT, LOSS_SORENSEN, LOSS_WEIGHTED_SORENSEN, LOSS_JACCARD,
  LOSS_WEIGHTED_JACCARD, LOSS_MIXED_SORENSEN, LOSS_MIXED_WEIGHTED_SORENSEN, LOSS_MIXED_JACCARD,
  LOSS_MIXED_WEIGHTED_JACCARD,
  LOSS_SORENSEN_FG  // legacy train.py:373-377: Dice of softmax[...,1] against the label volume
};
enum AttentionLoss : int { ATT_NONE = 0, ATT_L2 = 1, ATT_ABS = 2 };  // train.py:387-399
enum OptName : int { OPT_ADAM = 0, OPT_SGD = 1, OPT_MOMENTUM = 2, OPT_NESTEROV = 3 };

struct EngineConfig {
  int in_channels = 1, num_classes = 2, num_channels = 16, num_levels = 4;
  int num_convolutions[8] = {1, 2, 3, 3, 0, 0, 0, 0};
  int bottom_convolutions = 3;
  int patch[3] = {64, 64, 64};
  int max_batch = 1;
  int precision = PREC_FP32;
  int loss = LOSS_WEIGHTED_SORENSEN;
  float loss_weights[kMaxClasses] = {1, 1, 1, 1, 1, 1, 1, 1};
  float loss_alpha = 1.0f;
  int optimizer = OPT_ADAM;
  float lr0 = 1e-2f, decay_factor = 0.99f, decay_steps = 100.f;
  float momentum = 0.9f;
  int flavour = 0;  // 0: networks.VNet (live path, model.py:428-438); 1: VNet.py legacy flavour (train.py:271-279)
  int attention = 0;       // 1: AttentionModule -> (1 + softmax) gating -> OutputModule after the V-Net (train.py:281-312)
  int attention_loss = ATT_NONE;
  int module_channels = 64;  // attention.py:41 num_channels
};

// LossCfg of a configuration (Loss.Name / Weights / Alpha of the config JSON, model.py:495-560)
inline LossCfg loss_cfg_of(const EngineConfig& cfg) {
  LossCfg lc;
  lc.K = cfg.num_classes;
  const int l = cfg.loss;
  lc.jaccard = (l == LOSS_JACCARD || l == LOSS_WEIGHTED_JACCARD || l == LOSS_MIXED_JACCARD || l == LOSS_MIXED_WEIGHTED_JACCARD);
  lc.use_dice = !(l == LOSS_XENT || l == LOSS_WEIGHTED_XENT);
  lc.fg_only = (l == LOSS_SORENSEN_FG);
  lc.weighted_dice = (l == LOSS_WEIGHTED_SORENSEN || l == LOSS_WEIGHTED_JACCARD || l == LOSS_MIXED_WEIGHTED_SORENSEN || l == LOSS_MIXED_WEIGHTED_JACCARD);
  lc.use_xent = (l == LOSS_XENT || l == LOSS_WEIGHTED_XENT || (l >= LOSS_MIXED_SORENSEN && l <= LOSS_MIXED_WEIGHTED_JACCARD));
  lc.weighted_xent = (l == LOSS_WEIGHTED_XENT || l == LOSS_MIXED_WEIGHTED_SORENSEN || l == LOSS_MIXED_WEIGHTED_JACCARD);
  lc.xent_alpha = (l >= LOSS_MIXED_SORENSEN && l <= LOSS_MIXED_WEIGHTED_JACCARD) ? cfg.loss_alpha : 1.0f;
  lc.smooth = 1e-5f;
  for (int i = 0; i < kMaxClasses; ++i) lc.w[i] = cfg.loss_weights[i];
  return lc;
}

enum UnitKind : int { U_INPUT_TILE = 0, U_CONV5, U_DOWN, U_UP, U_CONV1, U_ADD, U_CONV3, U_GATE };

struct ParamEntry {
  std::string name;
  int ndim = 0;
  long long dims[5] = {0, 0, 0, 0, 0};
  size_t offset = 0, count = 0;
  bool trainable = true;  // false: moving statistics (state buffer)
};

struct Act {        // one activation tensor and its gradient
  int C = 0;
  Dims dims{0, 0, 0};
  float* a = nullptr;
  float* d = nullptr;
  uint16_t* a_hi = nullptr;  // bf16 copies for the tensor-core path
  uint16_t* a_lo = nullptr;
  uint16_t* d_hi = nullptr;
  uint16_t* d_lo = nullptr;
  bool needs_grad = true;
};

struct Unit {
  std::string scope;
  int kind = U_CONV5;
  int in1 = -1, in2 = -1, res = -1, out = -1;
  int Cin1 = 0, Cin2 = 0, Cout = 0;
  int chain = CH_S;
  bool has_act = false, has_dropout = false;
  bool relu = false;          // activation is ReLU (no alpha parameter): attention.py:51-52
  bool bn_inference = false;  // BN normalises with its moving statistics (train_phase=False, train.py:538-540)
  // parameter offsets (floats) into the flat buffers; -1 = absent
  long long w_off = -1, b_off = -1, alpha_off = -1;
  long long gamma_off[3] = {-1, -1, -1}, beta_off[3] = {-1, -1, -1};
  long long mm_off[3] = {-1, -1, -1}, mv_off[3] = {-1, -1, -1};
  size_t w_count = 0;
  size_t p_lo = 0;                    // first trainable-parameter offset of this unit
  // buffers
  float* z = nullptr;                 // pre-BN tensor [V][Cout]
  double *mean = nullptr, *var = nullptr;
  float *scale = nullptr, *shift = nullptr, *P = nullptr, *Q = nullptr, *S = nullptr;
  // backward plan
  bool res_accumulate = false, in1_accumulate = false, in2_accumulate = false, need_dgrad = true;
  TcConvPlan tc;                      // tensor-core plan (precision != fp32)
  K2TcPlan k2_fprop, k2_dgrad;        // tcgen05 / TMA plans of the 2^3 stride-2 units (precision != fp32)
  K2WgPlan k2_wgrad;
};

// NVTX range around the launches of one unit and pass ("vnet/decoder/level_1/conv_1 fwd"), on with VNB_NVTX=1:
// `ncu --nvtx --nvtx-include "<scope> fwd/"` then captures the kernels of a single layer.
struct NvtxRange {
  bool on;
  NvtxRange(bool enabled, const std::string& scope, const char* pass) : on(enabled) {
#ifndef VNB_EMULATE
    if (on) nvtxRangePushA((scope + pass).c_str());
#else
    (void)scope;
    (void)pass;
#endif
  }
  ~NvtxRange() {
#ifndef VNB_EMULATE
    if (on) nvtxRangePop();
#endif
  }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

class Engine {
 public:
  explicit Engine(const EngineConfig& cfg) : cfg_(cfg) {
    validate();
    VNB_CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
#ifndef VNB_EMULATE
    // Filter gradients are off the critical path of the backward pass (nothing downstream reads them before the
    // optimiser): they run on a second, lower-priority stream so that the HBM-bound batch-norm / 2x2x2 passes of the
    // following units share the SMs with the tensor-bound filter-gradient kernel.  VNB_WGRAD_STREAM=0 disables it.
    const char* ws = getenv("VNB_WGRAD_STREAM");
    if (!(ws && ws[0] == '0') && cfg_.precision != PREC_FP32) {
      int least = 0, greatest = 0;
      VNB_CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      VNB_CUDA_OK(cudaStreamCreateWithPriority(&wg_stream_, cudaStreamNonBlocking, least));
      VNB_CUDA_OK(cudaEventCreateWithFlags(&wg_ready_ev_, cudaEventDisableTiming));
      VNB_CUDA_OK(cudaEventCreateWithFlags(&wg_done_ev_, cudaEventDisableTiming));
      VNB_CUDA_OK(cudaEventCreateWithFlags(&pack_done_ev_, cudaEventDisableTiming));
    }
    // A gradient bucket's all-reduce waits on the main and on the filter-gradient stream itself (default); with
    // VNB_COMM_DUAL_WAIT=0 the main stream joins the filter-gradient stream at every bucket boundary instead, which
    // serialises the overlap of the two streams there.
    const char* dw = getenv("VNB_COMM_DUAL_WAIT");
    comm_waits_wgrad_ = !(dw && dw[0] == '0');
#endif
    // debugging switches (route one pass of the tensor-core modes through the exact-fp32 kernels), read once
    dbg_no_tc_fprop_ = getenv("VNB_DEBUG_NO_TC_FPROP") != nullptr;
    dbg_no_tc_dgrad_ = getenv("VNB_DEBUG_NO_TC_DGRAD") != nullptr;
    dbg_no_tc_wgrad_ = getenv("VNB_DEBUG_NO_TC_WGRAD") != nullptr;
    const char* nv = getenv("VNB_NVTX");
    nvtx_ = nv && nv[0] == '1';
    build_graph();
    allocate();
    init_default_params();
  }
  ~Engine() {
    for (void* p : allocs_) cudaFree(p);
#ifndef VNB_EMULATE
    if (wg_stream_) {
      cudaStreamDestroy(wg_stream_);
      cudaEventDestroy(wg_ready_ev_);
      cudaEventDestroy(wg_done_ev_);
      cudaEventDestroy(pack_done_ev_);
    }
#endif
    if (copy_stream_) {
      cudaStreamDestroy(copy_stream_);
      cudaEventDestroy(staged_ev_);
      cudaEventDestroy(staging_free_ev_);
    }
    cudaStreamDestroy(stream_);
  }

  struct Bucket {  // contiguous slice of the flat gradient buffer, complete once unit `unit_lo` has run backward
    int unit_lo, unit_hi;
    size_t lo, hi;
  };
  const std::vector<Bucket>& buckets() const { return buckets_; }
  void set_grad_hook(std::function<void(int)> hook) { grad_hook_ = std::move(hook); }
  // the filter-gradient stream while it holds work the main stream has not joined yet (communicator, dual-wait mode)
  cudaStream_t pending_wgrad_stream() const { return (comm_waits_wgrad_ && wg_stream_ && wg_pending_) ? wg_stream_ : 0; }
  // Synchronised batch norm: `hook(dev, n)` sums n doubles in place over `world` ranks, ordered after everything
  // enqueued on stream() and before anything enqueued later; an empty hook restores local statistics.
  void set_stats_hook(std::function<void(double*, int)> hook, int world) {
    stats_hook_ = std::move(hook);
    stats_world_ = stats_hook_ ? std::max(world, 1) : 1;
  }
  bool sync_bn() const { return static_cast<bool>(stats_hook_); }

  const std::vector<ParamEntry>& params() const { return entries_; }
  const EngineConfig& config() const { return cfg_; }
  size_t num_trainable() const { return n_train_; }
  cudaStream_t stream() const { return stream_; }
  float* grad_buffer() { return grads_; }
  long long global_step() const { return global_step_; }
  void set_global_step(long long s) { global_step_ = s; }
  void set_grad_scale(float s) { grad_scale_ = s; }

  const ParamEntry& find(const std::string& name) const {
    auto it = index_.find(name);
    if (it == index_.end()) throw std::invalid_argument("unknown variable: " + name);
    return entries_[it->second];
  }
  // kind: 0 value, 1 gradient, 2 Adam m, 3 Adam v
  float* buffer_for(const ParamEntry& e, int kind) {
    if (!e.trainable) {
      if (kind != 0) throw std::invalid_argument("moving statistics have no gradient / slots: " + e.name);
      return state_ + e.offset;
    }
    float* base = kind == 0 ? params_ : kind == 1 ? grads_ : kind == 2 ? adam_m_ : adam_v_;
    return base + e.offset;
  }
  void set_param(const std::string& name, const float* host, size_t bytes, int kind = 0) {
    const ParamEntry& e = find(name);
    if (bytes != e.count * sizeof(float)) throw std::invalid_argument("size mismatch for " + name);
    VNB_CUDA_OK(cudaMemcpyAsync(buffer_for(e, kind), host, bytes, cudaMemcpyHostToDevice, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
    weights_dirty_ = true;
  }
  void get_param(const std::string& name, float* host, size_t bytes, int kind = 0) {
    const ParamEntry& e = find(name);
    if (bytes != e.count * sizeof(float)) throw std::invalid_argument("size mismatch for " + name);
    VNB_CUDA_OK(cudaMemcpyAsync(host, buffer_for(e, kind), bytes, cudaMemcpyDeviceToHost, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
  }

  // ---- public steps ---------------------------------------------------------------------------
  // inference run (model.py:914-917): dropout 0, batch statistics, no moving-average update
  void forward_host(const float* images, int N, float* logits, float* softmax, long long* argmax) {
    check_batch(N);
    upload_images(images, N);
    labelled_n_ = 0;   // the logits no longer belong to the labels on the device
    forward(N, 0.f, 0, false);
    const long long V = voxels(N);
    const int K = cfg_.num_classes;
    if (softmax || argmax) {
      LossCfg lc = loss_cfg();
      dim3 grid(loss_blocks(), N);
      VNB_LAUNCH(softmax_loss_fwd_kernel, grid, 256, 0, stream_, acts_[head_act_].a, (const int32_t*)nullptr,
                 V / N, lc, softmax ? softmax_dev_ : (float*)nullptr, argmax ? argmax_dev_ : (long long*)nullptr,
                 (double*)nullptr);
    }
    if (logits) VNB_CUDA_OK(cudaMemcpyAsync(logits, acts_[head_act_].a, V * K * sizeof(float), cudaMemcpyDefault, stream_));
    if (softmax) VNB_CUDA_OK(cudaMemcpyAsync(softmax, softmax_dev_, V * K * sizeof(float), cudaMemcpyDefault, stream_));
    if (argmax) VNB_CUDA_OK(cudaMemcpyAsync(argmax, argmax_dev_, V * sizeof(long long), cudaMemcpyDefault, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
  }

  // Sliding-window evaluation of one case (model.py:866-937): `volume` is [X][Y][Z][M] with every extent >= the
  // patch extent (the caller pads, as the reference's Padding transform does).  Windows are enumerated in (i, j, k)
  // order with the last one clamped to the border, grouped into batches of `batch` consecutive windows (the grouping
  // matters: batch norm uses the statistics of each evaluation batch), run through the network and
  // accumulated on the device.  Outputs (each optional): label int64 [X][Y][Z], softmax sums [X][Y][Z][K], hit counts.
  void evaluate_volume_host(const float* volume, const int dims[3], const int stride[3], int batch, long long* label,
                            float* sum_out, float* weight_out) {
    const Dims P = acts_[image_act_].dims;
    const int M = cfg_.in_channels, K = cfg_.num_classes;
    check_batch(batch);
    if (dims[0] < P.D || dims[1] < P.H || dims[2] < P.W)
      throw std::invalid_argument("evaluate_volume: the volume must be padded to at least the patch size");
    if (stride[0] < 1 || stride[1] < 1 || stride[2] < 1) throw std::invalid_argument("evaluate_volume: stride must be >= 1");
    WindowGeom wg{dims[0], dims[1], dims[2], P.D, P.H, P.W, M, K};
    const long long V = static_cast<long long>(dims[0]) * dims[1] * dims[2];
    std::vector<int> starts;
    const int pd[3] = {P.D, P.H, P.W};
    int num[3];
    for (int a = 0; a < 3; ++a) num[a] = (dims[a] - pd[a] + stride[a] - 1) / stride[a] + 1;   // ceil((dim-patch)/stride)+1
    for (int i = 0; i < num[0]; ++i)
      for (int j = 0; j < num[1]; ++j)
        for (int k = 0; k < num[2]; ++k) {
          const int idx[3] = {i, j, k};
          for (int a = 0; a < 3; ++a) {
            int st = idx[a] * stride[a];
            if (st + pd[a] > dims[a]) st = dims[a] - pd[a];   // last window clamped (model.py:879-892)
            starts.push_back(st);
          }
        }
    const int nwin = static_cast<int>(starts.size() / 3);
    struct Scratch {
      std::vector<void*> p;
      ~Scratch() {
        for (void* q : p) cudaFree(q);
      }
      void* get(size_t bytes) {
        void* q = nullptr;
        if (cudaMalloc(&q, bytes ? bytes : 1) != cudaSuccess) throw std::runtime_error("CUDA: out of memory in evaluate_volume");
        p.push_back(q);
        return q;
      }
    } sc;
    float* vol_dev = static_cast<float*>(sc.get(V * M * sizeof(float)));
    float* sum_dev = static_cast<float*>(sc.get(V * K * sizeof(float)));
    float* wgt_dev = static_cast<float*>(sc.get(V * sizeof(float)));
    long long* lab_dev = static_cast<long long*>(sc.get(V * sizeof(long long)));
    int* starts_dev = static_cast<int*>(sc.get(starts.size() * sizeof(int)));
    VNB_CUDA_OK(cudaMemcpyAsync(vol_dev, volume, V * M * sizeof(float), cudaMemcpyHostToDevice, stream_));
    VNB_CUDA_OK(cudaMemcpyAsync(starts_dev, starts.data(), starts.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
    VNB_CUDA_OK(cudaMemsetAsync(sum_dev, 0, V * K * sizeof(float), stream_));
    VNB_CUDA_OK(cudaMemsetAsync(wgt_dev, 0, V * sizeof(float), stream_));
    const long long per = static_cast<long long>(P.D) * P.H * P.W;
    const LossCfg lc = loss_cfg();
    for (int b0 = 0; b0 < nwin; b0 += batch) {
      const int nb = std::min(batch, nwin - b0);
      VNB_LAUNCH(window_gather_kernel, grid_for(per * M * nb, 256), 256, 0, stream_, (const float*)vol_dev, wg,
                 (const int*)(starts_dev + 3 * b0), nb, acts_[image_act_].a);
      ++launches_;
      forward(nb, 0.f, 0, false);
      dim3 grid(loss_blocks(), nb);
      VNB_LAUNCH(softmax_loss_fwd_kernel, grid, 256, 0, stream_, acts_[head_act_].a, (const int32_t*)nullptr, per, lc, softmax_dev_,
                 (long long*)nullptr, (double*)nullptr);
      ++launches_;
      // The reference puts the last batch on its work list twice (model.py:903-904 appends the list object that
      // model.py:898-899 already appended), so that batch is run and accumulated twice; its second run produces the same
      // softmax (same windows, same batch statistics), so only the accumulation is repeated here.
      const int reps = (b0 + batch >= nwin) ? 2 : 1;
      for (int rep = 0; rep < reps; ++rep)
        for (int j = 0; j < nb; ++j) {
          const int* st = &starts[3 * (b0 + j)];
          VNB_LAUNCH(window_accumulate_kernel, grid_for(per, 256), 256, 0, stream_, (const float*)(softmax_dev_ + j * per * K), wg,
                     st[0], st[1], st[2], sum_dev, wgt_dev);
          ++launches_;
        }
    }
    if (label) {
      VNB_LAUNCH(volume_argmax_kernel, grid_for(V, 256), 256, 0, stream_, (const float*)sum_dev, V, K, lab_dev);
      ++launches_;
      VNB_CUDA_OK(cudaMemcpyAsync(label, lab_dev, V * sizeof(long long), cudaMemcpyDeviceToHost, stream_));
    }
    if (sum_out) VNB_CUDA_OK(cudaMemcpyAsync(sum_out, sum_dev, V * K * sizeof(float), cudaMemcpyDeviceToHost, stream_));
    if (weight_out) VNB_CUDA_OK(cudaMemcpyAsync(weight_out, wgt_dev, V * sizeof(float), cudaMemcpyDeviceToHost, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
  }

  // loss only (the in-loop test step of model.py:784-789): returns loss, optional dice terms [N][K][4]
  float loss_host(const float* images, const int32_t* labels, int N, double* terms) {
    check_batch(N);
    upload_images(images, N);
    upload_labels(labels, N);
    forward(N, 0.f, 0, false);
    loss_forward(N);
    float loss;
    VNB_CUDA_OK(cudaMemcpyAsync(&loss, loss_dev_, sizeof(float), cudaMemcpyDeviceToHost, stream_));
    if (terms)
      VNB_CUDA_OK(cudaMemcpyAsync(terms, terms_dev_, sizeof(double) * N * cfg_.num_classes * 4, cudaMemcpyDeviceToHost, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
    return loss;
  }

  // forward + loss + backward; gradients left in the flat gradient buffer (no optimiser step)
  void forward_backward_device(int N, float dropout, uint64_t seed, bool update_moving) {
    forward(N, dropout, seed, update_moving);
    loss_forward(N);
    backward(N, dropout, seed);
  }
  void upload_batch(const float* images, const int32_t* labels, int N) {
    check_batch(N);
    upload_images(images, N);
    upload_labels(labels, N);
  }
  // Input staging (the step before the hot path): the next batch is copied into staging buffers on a copy
  // stream while the current step computes; commit_staged() orders the compute stream after that copy, moves the batch
  // into the network's input buffers device to device (microseconds) and frees the staging buffers for the next one.
  void stage_batch(const float* images, const int32_t* labels, int N) {
    check_batch(N);
    const Act& a = acts_[image_act_];
    if (!copy_stream_) {
      VNB_CUDA_OK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
      VNB_CUDA_OK(cudaEventCreateWithFlags(&staged_ev_, cudaEventDisableTiming));
      VNB_CUDA_OK(cudaEventCreateWithFlags(&staging_free_ev_, cudaEventDisableTiming));
      stage_img_ = dev_alloc<float>(voxels_of(a.dims, cfg_.max_batch) * a.C);
      stage_lab_ = dev_alloc<int32_t>(voxels(cfg_.max_batch));
    } else {
      VNB_CUDA_OK(cudaStreamWaitEvent(copy_stream_, staging_free_ev_, 0));   // the previous commit has read the buffers
    }
    VNB_CUDA_OK(cudaMemcpyAsync(stage_img_, images, voxels_of(a.dims, N) * a.C * sizeof(float), cudaMemcpyDefault, copy_stream_));
    VNB_CUDA_OK(cudaMemcpyAsync(stage_lab_, labels, voxels(N) * sizeof(int32_t), cudaMemcpyDefault, copy_stream_));
    VNB_CUDA_OK(cudaEventRecord(staged_ev_, copy_stream_));
    staged_n_ = N;
  }
  int commit_staged() {
    if (staged_n_ == 0) throw std::invalid_argument("no staged batch: call vnb_stage_batch first");
    const Act& a = acts_[image_act_];
    const int N = staged_n_;
    VNB_CUDA_OK(cudaStreamWaitEvent(stream_, staged_ev_, 0));
    VNB_CUDA_OK(cudaMemcpyAsync(a.a, stage_img_, voxels_of(a.dims, N) * a.C * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
    VNB_CUDA_OK(cudaMemcpyAsync(labels_dev_, stage_lab_, voxels(N) * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream_));
    VNB_CUDA_OK(cudaEventRecord(staging_free_ev_, stream_));
    labelled_n_ = N;
    staged_n_ = 0;
    return N;
  }
  // optimiser over the flat buffers; `world` folds the data-parallel mean into the update
  void optimizer_step(int world) {
    const float lr = cfg_.lr0 * std::pow(cfg_.decay_factor, static_cast<float>(global_step_) / cfg_.decay_steps);
    const long long t = global_step_ + 1;
    const float gscale = grad_scale_ / static_cast<float>(world);
    const int blocks = grid_for(n_train_, 256);
    if (cfg_.optimizer == OPT_ADAM) {
      const double b1 = 0.9, b2 = 0.999;
      const float lr_t = static_cast<float>(lr * std::sqrt(1.0 - std::pow(b2, (double)t)) / (1.0 - std::pow(b1, (double)t)));
      VNB_LAUNCH(adam_step_kernel, blocks, 256, 0, stream_, params_, (const float*)grads_, adam_m_, adam_v_,
                 (long long)n_train_, lr_t, 0.9f, 0.999f, 1e-8f, gscale);
    } else if (cfg_.optimizer == OPT_SGD) {
      VNB_LAUNCH(sgd_step_kernel, blocks, 256, 0, stream_, params_, (const float*)grads_, (long long)n_train_, lr, gscale);
    } else {  // Momentum / Nesterov: the accumulator lives in the first Adam slot buffer
      VNB_LAUNCH(momentum_step_kernel, blocks, 256, 0, stream_, params_, (const float*)grads_, adam_m_, (long long)n_train_, lr,
                 cfg_.momentum, cfg_.optimizer == OPT_NESTEROV ? 1 : 0, gscale);
    }
    global_step_ = t;
    weights_dirty_ = true;
  }
  float read_loss() {
    float loss;
    VNB_CUDA_OK(cudaMemcpyAsync(&loss, loss_dev_, sizeof(float), cudaMemcpyDeviceToHost, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
    return loss;
  }
  // single-process training step (model.py:743-748)
  float train_step_host(const float* images, const int32_t* labels, int N, float dropout, uint64_t seed, bool want_loss) {
    upload_batch(images, labels, N);
    forward_backward_device(N, dropout, seed, true);
    optimizer_step(1);
    return want_loss ? read_loss() : 0.f;
  }
  // step metrics of model.py:586-626 for the batch of the last call that was given labels (loss / forward_backward /
  // train step): integer counts, see metrics_kernel.  confusion [(K+1)][K], auc_hist (optional) [K][2][kAucBins].
  void read_metrics(int N, unsigned long long* confusion, unsigned long long* auc_hist) {
    check_batch(N);
    if (N != labelled_n_)
      throw std::invalid_argument("read_metrics: the last call with labels (loss / forward_backward / train step) had a batch of " +
                                  std::to_string(labelled_n_));
    const int K = cfg_.num_classes;
    const size_t n_cm = static_cast<size_t>(K + 1) * K, n_hist = static_cast<size_t>(K) * 2 * kAucBins;
    VNB_CUDA_OK(cudaMemsetAsync(metrics_dev_, 0, (n_cm + n_hist) * sizeof(unsigned long long), stream_));
    const long long V = voxels(N);
    VNB_LAUNCH(metrics_kernel, grid_for(V, 256, 4 * sm_count_), 256, 0, stream_, (const float*)acts_[head_act_].a,
               (const int32_t*)labels_dev_, V, K, auc_hist ? 1 : 0, metrics_dev_, metrics_dev_ + n_cm);
    ++launches_;
    VNB_CUDA_OK(cudaMemcpyAsync(confusion, metrics_dev_, n_cm * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
    if (auc_hist)
      VNB_CUDA_OK(cudaMemcpyAsync(auc_hist, metrics_dev_ + n_cm, n_hist * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
  }
  void sync() { VNB_CUDA_OK(cudaStreamSynchronize(stream_)); }
  // distance map of the attention loss (train.py:176-179 distmap_placeholder), [N][D][H][W] in [0,1]
  void set_distmap(const float* distmap, int N) {
    if (!cfg_.attention) throw std::invalid_argument("set_distmap: the attention path is not enabled");
    check_batch(N);
    VNB_CUDA_OK(cudaMemcpyAsync(distmap_dev_, distmap, voxels(N) * sizeof(float), cudaMemcpyDefault, stream_));
    distmap_valid_ = true;
  }
  void read_losses(float out[3]) {  // total, segmentation, attention (train.py:417)
    VNB_CUDA_OK(cudaMemcpyAsync(out, loss_dev_, 3 * sizeof(float), cudaMemcpyDeviceToHost, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
  }
  // the two summands of a mixed loss as the reference logs them (model.py:529-530): '1.dice' = 1 - dice and
  // '2.regularized_xent' = Loss.Alpha * cross entropy (0 for the pure Dice losses; the whole loss for the x-ent ones)
  void read_loss_parts(float out[2]) {
    float l[4];
    VNB_CUDA_OK(cudaMemcpyAsync(l, loss_dev_, 4 * sizeof(float), cudaMemcpyDeviceToHost, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
    out[0] = l[3];
    out[1] = l[1] - l[3];
  }
  // softmax_attention of the last forward pass (train.py:288), [N][D][H][W][K]
  void read_softmax_attention(float* host, int N) {
    if (!cfg_.attention) throw std::invalid_argument("the attention path is not enabled");
    VNB_CUDA_OK(cudaMemcpyAsync(host, gate_soft_, voxels(N) * cfg_.num_classes * sizeof(float), cudaMemcpyDeviceToHost, stream_));
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
  }

  // debug access for tests: copy an activation (kind 0), its gradient (1) or the unit's pre-BN z (2)
  void read_tensor(const std::string& scope, int kind, float* host, size_t bytes, int N) {
    for (const Unit& u : units_)
      if (u.scope == scope) {
        const Act& a = acts_[u.out];
        const size_t n = static_cast<size_t>(voxels_of(a.dims, N)) * a.C * sizeof(float);
        if (bytes != n) throw std::invalid_argument("read_tensor size mismatch for " + scope);
        const float* src = kind == 0 ? a.a : kind == 1 ? a.d : u.z;
        VNB_CUDA_OK(cudaMemcpyAsync(host, src, n, cudaMemcpyDeviceToHost, stream_));
        VNB_CUDA_OK(cudaStreamSynchronize(stream_));
        return;
      }
    throw std::invalid_argument("unknown scope: " + scope);
  }
  long long gpu_launches() const { return launches_; }

  // ---- device timing ----------------------------------------------------------------------------
  void event_record(int which) {
#ifndef VNB_EMULATE
    if (!timer_ev_[0]) {
      VNB_CUDA_OK(cudaEventCreate(&timer_ev_[0]));
      VNB_CUDA_OK(cudaEventCreate(&timer_ev_[1]));
    }
    VNB_CUDA_OK(cudaEventRecord(timer_ev_[which & 1], stream_));
#else
    (void)which;
#endif
  }
  float event_elapsed_ms() {
    float ms = 0.f;
#ifndef VNB_EMULATE
    VNB_CUDA_OK(cudaEventSynchronize(timer_ev_[1]));
    VNB_CUDA_OK(cudaEventElapsedTime(&ms, timer_ev_[0], timer_ev_[1]));
#endif
    return ms;
  }
  void profile_enable(bool on) {
    profiling_ = on;
    prof_used_ = 0;
  }
  void profile_read(int cls, double* ms, long long* launches, double* flops) {
    double t = 0, f = 0;
    long long n = 0;
#ifndef VNB_EMULATE
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
#endif
    for (size_t i = 0; i < prof_used_; ++i) {
      if (prof_[i].cls != cls) continue;
      float e = 0.f;
#ifndef VNB_EMULATE
      VNB_CUDA_OK(cudaEventElapsedTime(&e, prof_[i].a, prof_[i].b));
#endif
      t += e;
      f += prof_[i].flops;
      ++n;
    }
    *ms = t;
    *launches = n;
    *flops = f;
  }
  // one profiled launch: class, time, algorithmic FLOPs and "scope pass Cin->Cout @DxHxW"; false past the last record
  size_t profile_count() const { return prof_used_; }
  bool profile_launch(size_t i, int* cls, double* ms, double* flops, std::string* label) {
    if (i >= prof_used_) return false;
    float e = 0.f;
#ifndef VNB_EMULATE
    VNB_CUDA_OK(cudaStreamSynchronize(stream_));
    if (wg_stream_) VNB_CUDA_OK(cudaStreamSynchronize(wg_stream_));
    VNB_CUDA_OK(cudaEventElapsedTime(&e, prof_[i].a, prof_[i].b));
#endif
    *cls = prof_[i].cls;
    *ms = e;
    *flops = prof_[i].flops;
    label->clear();
    if (const Unit* u = prof_[i].unit) {
      const Dims& d = acts_[u->out].dims;
      *label = u->scope + " " + prof_[i].pass + " " + std::to_string(u->Cin1 + u->Cin2) + "->" + std::to_string(u->Cout) +
               " @" + std::to_string(d.D) + "x" + std::to_string(d.H) + "x" + std::to_string(d.W);
    }
    return true;
  }
  struct ProfScope {  // brackets one kernel launch with events when profiling is on
    Engine& e;
    long long idx = -1;
    cudaStream_t st;
    ProfScope(Engine& eng, int cls, double flops, cudaStream_t stream = 0, const Unit* unit = nullptr,
              const char* pass = "")
        : e(eng), st(stream ? stream : eng.stream_) {
      if (!e.profiling_) return;
      if (e.prof_used_ == e.prof_.size()) {
        ProfRec r;
#ifndef VNB_EMULATE
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
#endif
        e.prof_.push_back(r);
      }
      idx = static_cast<long long>(e.prof_used_++);
      e.prof_[idx].cls = cls;
      e.prof_[idx].flops = flops;
      e.prof_[idx].unit = unit;
      e.prof_[idx].pass = pass;
#ifndef VNB_EMULATE
      cudaEventRecord(e.prof_[idx].a, st);
#endif
    }
    ~ProfScope() {
#ifndef VNB_EMULATE
      if (idx >= 0) cudaEventRecord(e.prof_[idx].b, st);
#endif
    }
  };

 private:
  // ---- graph construction ---------------------------------------------------------------------
  void validate() {
    const EngineConfig& c = cfg_;
    if (c.num_levels < 1 || c.num_levels > 8) throw std::invalid_argument("num_levels out of range");
    if (c.num_classes < 1 || c.num_classes > kMaxClasses) throw std::invalid_argument("num_classes out of range (1..8)");
    if (c.in_channels < 1 || c.max_batch < 1) throw std::invalid_argument("bad in_channels / max_batch");
    for (int i = 0; i < 3; ++i)
      if (c.patch[i] <= 0 || c.patch[i] % (1 << c.num_levels))
        throw std::invalid_argument("PatchShape must be divisible by 2^NumLevels (odd skip sizes unsupported, SURVEY H7)");
    if (c.num_channels * (1 << c.num_levels) > 256 && (c.num_channels * (1 << c.num_levels)) % 256)
      throw std::invalid_argument("unsupported channel width");
  }
  long long voxels_of(const Dims& d, int N) const { return static_cast<long long>(N) * d.D * d.H * d.W; }
  long long voxels(int N) const { return voxels_of(acts_[head_act_].dims, N); }

  int add_act(int C, Dims dims, bool needs_grad = true) {
    Act a;
    a.C = C;
    a.dims = dims;
    a.needs_grad = needs_grad;
    acts_.push_back(a);
    return static_cast<int>(acts_.size()) - 1;
  }
  long long add_param(const std::string& name, std::initializer_list<long long> dims, bool trainable) {
    ParamEntry e;
    e.name = name;
    e.ndim = static_cast<int>(dims.size());
    size_t cnt = 1;
    int i = 0;
    for (long long d : dims) {
      e.dims[i++] = d;
      cnt *= static_cast<size_t>(d);
    }
    e.count = cnt;
    e.trainable = trainable;
    size_t& cursor = trainable ? n_train_ : n_state_;
    e.offset = cursor;
    cursor += (cnt + 3) & ~size_t(3);  // keep every tensor 16-byte aligned
    index_[name] = entries_.size();
    entries_.push_back(e);
    return static_cast<long long>(e.offset);
  }
  void add_bn(Unit& u, int k, int C, int name_offset = 0) {
    const int idx = k + name_offset;
    const std::string base = u.scope + "/batch_normalization" + (idx == 0 ? "" : "_" + std::to_string(idx));
    u.gamma_off[k] = add_param(base + "/gamma", {C}, true);
    u.beta_off[k] = add_param(base + "/beta", {C}, true);
    u.mm_off[k] = add_param(base + "/moving_mean", {C}, false);
    u.mv_off[k] = add_param(base + "/moving_variance", {C}, false);
  }
  // registers parameters in TF creation order: weights, biases, BNs, alpha
  struct UnitNames {  // explicit variable names (tf.Variable auto-names of the attention / output modules)
    std::string w, b, bn;
  };
  int add_unit(int kind, const std::string& scope, int in1, int in2, int res, int Cout, Dims out_dims,
               int chain, bool act, bool dropout, int bn_name_offset = 0, const UnitNames* names = nullptr) {
    Unit u;
    u.scope = scope;
    u.kind = kind;
    u.in1 = in1;
    u.in2 = in2;
    u.res = res;
    u.Cin1 = acts_[in1].C;
    u.Cin2 = in2 >= 0 ? acts_[in2].C : 0;
    u.Cout = Cout;
    u.chain = chain;
    u.has_act = act;
    u.has_dropout = dropout;
    u.p_lo = n_train_;
    const long long Cin = u.Cin1 + u.Cin2;
    if (names) {  // module unit: ReLU, inference-mode BN, weights / biases are anonymous tf.Variables
      u.relu = act;
      u.bn_inference = true;
      const long long ks = kind == U_CONV3 ? 3 : 1;
      u.w_off = add_param(names->w, {ks, ks, ks, Cin, Cout}, true);
      u.w_count = static_cast<size_t>(ks * ks * ks * Cin * Cout);
      u.b_off = add_param(names->b, {Cout}, true);
      for (const char* leaf : {"/gamma", "/beta"}) {
        const long long off = add_param(names->bn + leaf, {Cout}, true);
        (leaf[1] == 'g' ? u.gamma_off[0] : u.beta_off[0]) = off;
      }
      u.mm_off[0] = add_param(names->bn + "/moving_mean", {Cout}, false);
      u.mv_off[0] = add_param(names->bn + "/moving_variance", {Cout}, false);
      u.out = add_act(Cout, out_dims);
      units_.push_back(u);
      return u.out;
    }
    if (kind == U_GATE) {  // no parameters, no batch norm
      u.out = add_act(Cout, out_dims);
      units_.push_back(u);
      return u.out;
    }
    if (kind == U_CONV5) {
      u.w_off = add_param(scope + "/weights", {5, 5, 5, Cin, Cout}, true);
      u.w_count = 125ull * Cin * Cout;
    } else if (kind == U_DOWN) {
      u.w_off = add_param(scope + "/weights", {2, 2, 2, Cin, Cout}, true);
      u.w_count = 8ull * Cin * Cout;
    } else if (kind == U_UP) {  // layers2.py:92: [2,2,2,Cout,Cin]
      u.w_off = add_param(scope + "/weights", {2, 2, 2, Cout, Cin}, true);
      u.w_count = 8ull * Cin * Cout;
    } else if (kind == U_CONV1) {
      u.w_off = add_param(scope + "/weights", {1, 1, 1, Cin, Cout}, true);
      u.w_count = static_cast<size_t>(Cin) * Cout;
    }
    if (kind != U_INPUT_TILE && kind != U_ADD) u.b_off = add_param(scope + "/biases", {Cout}, true);
    for (int k = 0; k < chain_num_bn(chain); ++k) add_bn(u, k, Cout, bn_name_offset);
    if (act) u.alpha_off = add_param(scope + "/alpha", {Cout}, true);
    u.out = add_act(Cout, out_dims);
    units_.push_back(u);
    return u.out;
  }

  void build_graph() {
    const EngineConfig& c = cfg_;
    Dims full{c.patch[0], c.patch[1], c.patch[2]};
    const int C0 = c.num_channels;
    image_act_ = add_act(c.in_channels, full, false);
    int x;
    if (c.in_channels == 1)  // networks.py:254-259
      x = add_unit(U_INPUT_TILE, "vnet/input_layer", image_act_, -1, -1, C0, full, CH_S, false, false);
    else                      // networks.py:260-266
      x = add_unit(U_CONV5, "vnet/input_layer", image_act_, -1, -1, C0, full, CH_S, true, false);
    const bool legacy = c.flavour == 1;
    // VNet.py:26-40: conv -> BN -> (+ block input on the last conv) -> BN -> act -> dropout.  The add sits
    // between the two batch norms, so the last conv becomes two units: conv+BN0, then add+BN1+act+dropout.
    auto legacy_conv = [&](const std::string& sc, int in1, int in2, int res, int ch, Dims dd) {
      if (res < 0) return add_unit(U_CONV5, sc, in1, in2, -1, ch, dd, CH_2, true, true);
      const int u0 = add_unit(U_CONV5, sc, in1, in2, -1, ch, dd, CH_S, false, false);
      return add_unit(U_ADD, sc, u0, -1, res, ch, dd, CH_S, true, true, 1);
    };
    std::vector<int> features;
    Dims d = full;
    for (int l = 0; l < c.num_levels; ++l) {  // networks.py:270-280
      const int ch = C0 << l;
      const std::string scope = "vnet/encoder/level_" + std::to_string(l + 1);
      const int block_in = x, n = c.num_convolutions[l];
      for (int i = 0; i < n; ++i) {
        const std::string sc = scope + "/conv_" + std::to_string(i + 1);
        if (legacy)
          x = legacy_conv(sc, x, -1, i == n - 1 ? block_in : -1, ch, d);
        else
          x = add_unit(U_CONV5, sc, x, -1, i == n - 1 ? block_in : -1, ch, d, CH_S, true, true);
      }
      features.push_back(x);
      Dims half{d.D / 2, d.H / 2, d.W / 2};
      x = add_unit(U_DOWN, scope + "/down_convolution", x, -1, -1, 2 * ch, half, CH_S, true, false);
      d = half;
    }
    {  // networks.py:282-283
      const int ch = C0 << c.num_levels, block_in = x, n = c.bottom_convolutions;
      for (int i = 0; i < n; ++i) {
        const std::string sc = "vnet/bottom_level/conv_" + std::to_string(i + 1);
        if (legacy)
          x = legacy_conv(sc, x, -1, i == n - 1 ? block_in : -1, ch, d);
        else
          x = add_unit(U_CONV5, sc, x, -1, i == n - 1 ? block_in : -1, ch, d, CH_S, true, true);
      }
    }
    for (int l = c.num_levels - 1; l >= 0; --l) {  // networks.py:285-296
      const int ch = C0 << l;
      const std::string scope = "vnet/decoder/level_" + std::to_string(l + 1);
      const int f = features[l];
      Dims up = acts_[f].dims;
      x = add_unit(U_UP, scope + "/up_convolution", x, -1, -1, ch, up, CH_S, true, false);
      const int n = c.num_convolutions[l];
      if (legacy) {  // VNet.py:43-73: true residual to the up-convolution output
        const int x_up = x;
        if (n == 1) {
          x = legacy_conv(scope + "/conv_1", x_up, f, x_up, ch, up);
        } else {
          x = add_unit(U_CONV5, scope + "/conv_1", x_up, f, -1, ch, up, CH_S, true, true);
          for (int i = 1; i < n; ++i)
            x = legacy_conv(scope + "/conv_" + std::to_string(i + 1), x, -1, i == n - 1 ? x_up : -1, ch, up);
        }
      } else if (n == 1) {  // networks.py:328-340
        x = add_unit(U_CONV5, scope + "/conv_1", x, f, -1, ch, up, CH_T, true, true);
      } else {       // networks.py:342-363
        x = add_unit(U_CONV5, scope + "/conv_1", x, f, -1, ch, up, CH_S, true, true);
        for (int i = 1; i < n; ++i)
          x = add_unit(U_CONV5, scope + "/conv_" + std::to_string(i + 1), x, -1, -1, ch, up, i == n - 1 ? CH_Q : CH_D, true, true);
      }
      d = up;
    }
    head_act_ = add_unit(U_CONV1, "vnet/output_layer", x, -1, -1, c.num_classes, full, CH_S, false, false);  // networks.py:298-303
    if (c.attention) {  // train.py:281-312
      vnet_logits_act_ = head_act_;
      const int att = add_module("attention/AttentionModule", "AttentionModule", head_act_, full);  // attention.py:105-114
      att_logits_act_ = att;
      gate_unit_ = static_cast<int>(units_.size());
      const int masked = add_unit(U_GATE, "masked_vnet", att, head_act_, -1, c.num_classes, full, CH_S, false, false);
      head_act_ = add_module("output/output", "output", masked, full);                               // OutputModule.py:105-114
    }

    // gradient buckets for the data-parallel exchange: ~1/8 of the parameters each, cut at unit
    // boundaries, listed in the order the backward pass completes them (last unit first)
    {
      const size_t target = std::max<size_t>(n_train_ / 8, 1u << 18);
      int hi_unit = static_cast<int>(units_.size());
      size_t hi_off = n_train_;
      for (int ui = static_cast<int>(units_.size()) - 1; ui >= 0; --ui) {
        if (hi_off - units_[ui].p_lo >= target || ui == 0) {
          if (hi_off > units_[ui].p_lo) buckets_.push_back(Bucket{ui, hi_unit, units_[ui].p_lo, hi_off});
          hi_unit = ui;
          hi_off = units_[ui].p_lo;
        }
      }
    }
    // backward plan: walk units in reverse, first writer of each gradient buffer overwrites
    std::vector<bool> written(acts_.size(), false);
    written[head_act_] = true;  // dL/dlogits written by the loss kernel
    for (int ui = static_cast<int>(units_.size()) - 1; ui >= 0; --ui) {
      Unit& u = units_[ui];
      if (u.res >= 0) {
        u.res_accumulate = written[u.res];
        written[u.res] = true;
      }
      u.need_dgrad = (u.kind != U_INPUT_TILE) && acts_[u.in1].needs_grad;
      if (u.need_dgrad) {
        u.in1_accumulate = written[u.in1];
        written[u.in1] = true;
        if (u.in2 >= 0) {
          u.in2_accumulate = written[u.in2];
          written[u.in2] = true;
        }
      }
    }
  }

  // AttentionModule / OutputModule (attention.py:83-114, OutputModule.py:83-114): three residual blocks of
  // {3^3 conv + BN + ReLU, 3^3 conv + BN, 1^3 shortcut conv, add, BN, ReLU} and a 1^3 conv + BN to K classes.
  // `vs` prefixes the anonymous tf.Variables (name scope + variable scope), `bs` the tf.layers BN variables.
  int add_module(const std::string& vs, const std::string& bs, int x, Dims dims) {
    const int nch = cfg_.module_channels;
    int vi = 0, bi = 0;
    auto names = [&](const std::string& sub) {
      UnitNames n;
      auto var = [&]() { return vs + "/" + sub + "/Variable" + (vi == 0 ? std::string() : "_" + std::to_string(vi)); };
      n.w = var();
      ++vi;
      n.b = var();
      ++vi;
      n.bn = bs + "/" + sub + "/batch_normalization" + (bi == 0 ? std::string() : "_" + std::to_string(bi));
      ++bi;
      return n;
    };
    for (int blk = 0; blk < 3; ++blk) {
      const std::string sc = bs + "/encoder/block_" + std::to_string(blk + 1);
      UnitNames n1 = names("encoder"), n2 = names("encoder"), n3 = names("encoder");
      const int c1 = add_unit(U_CONV3, sc + "/conv1", x, -1, -1, nch, dims, CH_S, true, false, 0, &n1);
      const int c2 = add_unit(U_CONV3, sc + "/conv2", c1, -1, -1, nch, dims, CH_S, false, false, 0, &n2);
      x = add_unit(U_CONV1, sc, x, -1, c2, nch, dims, CH_S, true, false, 0, &n3);
    }
    vi = bi = 0;
    UnitNames no = names("output");
    return add_unit(U_CONV1, bs + "/output", x, -1, -1, cfg_.num_classes, dims, CH_S, false, false, 0, &no);
  }

  template <class T>
  T* dev_alloc(size_t count) {
    void* p = nullptr;
    VNB_CUDA_OK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    allocs_.push_back(p);
    return static_cast<T*>(p);
  }

  void allocate() {
    const int NB = cfg_.max_batch;
    params_ = dev_alloc<float>(n_train_);
    grads_ = dev_alloc<float>(n_train_);
    adam_m_ = dev_alloc<float>(n_train_);
    adam_v_ = dev_alloc<float>(n_train_);
    state_ = dev_alloc<float>(n_state_);
    VNB_CUDA_OK(cudaMemset(grads_, 0, n_train_ * sizeof(float)));
    VNB_CUDA_OK(cudaMemset(adam_m_, 0, n_train_ * sizeof(float)));
    VNB_CUDA_OK(cudaMemset(adam_v_, 0, n_train_ * sizeof(float)));
    const bool tc = cfg_.precision != PREC_FP32;
    const bool lo = cfg_.precision == PREC_BF16X3;
    size_t max_w5 = 0;
    int maxC = 1;
    for (Act& a : acts_) {
      const size_t n = static_cast<size_t>(voxels_of(a.dims, NB)) * a.C;
      a.a = dev_alloc<float>(n);
      if (a.needs_grad) a.d = dev_alloc<float>(n);
      if (tc) {
        size_t nh = n;
        if (&a == &acts_[image_act_] && a.C > 1 && a.C % 16) {  // padded bf16 copies of a multi-modal input
          image_cpad_ = (a.C + 15) / 16 * 16;
          nh = static_cast<size_t>(voxels_of(a.dims, NB)) * image_cpad_;
        }
        a.a_hi = dev_alloc<uint16_t>(nh);
        if (lo) a.a_lo = dev_alloc<uint16_t>(nh);
        if (a.needs_grad) {
          a.d_hi = dev_alloc<uint16_t>(n);
          if (lo) a.d_lo = dev_alloc<uint16_t>(n);
        }
      }
      maxC = std::max(maxC, a.C);
    }
    for (Unit& u : units_) {
      const Act& o = acts_[u.out];
      const size_t n = static_cast<size_t>(voxels_of(o.dims, NB)) * o.C;
      if (u.kind == U_GATE) continue;
      u.z = (u.kind == U_INPUT_TILE) ? acts_[u.in1].a : dev_alloc<float>(n);
      u.mean = dev_alloc<double>(u.Cout);
      u.var = dev_alloc<double>(u.Cout);
      u.scale = dev_alloc<float>(u.Cout);
      u.shift = dev_alloc<float>(u.Cout);
      u.P = dev_alloc<float>(u.Cout);
      u.Q = dev_alloc<float>(u.Cout);
      u.S = dev_alloc<float>(u.Cout);
      if (u.kind == U_CONV5 || u.kind == U_CONV3) max_w5 = std::max(max_w5, u.w_count);
    }
    zero_alpha_ = dev_alloc<float>(maxC);
    VNB_CUDA_OK(cudaMemset(zero_alpha_, 0, maxC * sizeof(float)));
    if (cfg_.attention) {
      gate_soft_ = dev_alloc<float>(voxels(NB) * cfg_.num_classes);
      distmap_dev_ = dev_alloc<float>(voxels(NB));
      att_partial_ = dev_alloc<double>(kMaxRedBlocks);
    }
    partial_ = dev_alloc<double>(static_cast<size_t>(kMaxRedBlocks) * 3 * std::max(maxC, 4 * kMaxClasses) + 64);
    sync_stride_ = std::max(maxC, 4);
    sync_buf_ = dev_alloc<double>(static_cast<size_t>(5) * sync_stride_);
    wflip_ = dev_alloc<float>(max_w5);
    labels_dev_ = dev_alloc<int32_t>(voxels(NB));
    metrics_dev_ = dev_alloc<unsigned long long>(static_cast<size_t>(kMaxClasses + 1) * kMaxClasses + static_cast<size_t>(kMaxClasses) * 2 * kAucBins);
    softmax_dev_ = dev_alloc<float>(voxels(NB) * cfg_.num_classes);
    argmax_dev_ = dev_alloc<long long>(voxels(NB));
    loss_partial_ = dev_alloc<double>(static_cast<size_t>(NB) * kLossBlocks * kMaxClasses * 4);
    terms_dev_ = dev_alloc<double>(static_cast<size_t>(NB) * kMaxClasses * 4);
    coef_dev_ = dev_alloc<float>(static_cast<size_t>(NB) * kMaxClasses * 3);
    loss_dev_ = dev_alloc<float>(4);
    if (tc) tc_setup();
  }

  void init_default_params() {  // biases 0, gamma 1, beta 0, moving mean 0 / variance 1, alpha 0.1; weights 0
    std::vector<float> p(n_train_, 0.f), s(n_state_, 0.f);
    for (const ParamEntry& e : entries_) {
      const std::string& nm = e.name;
      auto ends = [&](const char* suf) { const size_t l = strlen(suf); return nm.size() >= l && nm.compare(nm.size() - l, l, suf) == 0; };
      float v = 0.f;
      if (ends("/gamma") || ends("/moving_variance")) v = 1.f;
      if (ends("/alpha")) v = 0.1f;
      float* dst = (e.trainable ? p.data() : s.data()) + e.offset;
      for (size_t i = 0; i < e.count; ++i) dst[i] = v;
    }
    VNB_CUDA_OK(cudaMemcpy(params_, p.data(), n_train_ * sizeof(float), cudaMemcpyHostToDevice));
    VNB_CUDA_OK(cudaMemcpy(state_, s.data(), n_state_ * sizeof(float), cudaMemcpyHostToDevice));
  }

  // ---- helpers --------------------------------------------------------------------------------
  static int grid_for(long long n, int block, int cap = kMaxRedBlocks * 2) {
    long long b = (n + block - 1) / block;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(b, cap)));
  }
  void check_batch(int N) const {
    if (N < 1 || N > cfg_.max_batch) throw std::invalid_argument("batch size outside [1, max_batch]");
  }
  // images / labels / distance maps / forward outputs: caller-owned HOST or DEVICE memory (cudaMemcpyDefault resolves
  // the direction through unified addressing), so a batch produced on the GPU never bounces through the host
  void upload_images(const float* images, int N) {
    const Act& a = acts_[image_act_];
    VNB_CUDA_OK(cudaMemcpyAsync(a.a, images, voxels_of(a.dims, N) * a.C * sizeof(float), cudaMemcpyDefault, stream_));
  }
  void upload_labels(const int32_t* labels, int N) {
    VNB_CUDA_OK(cudaMemcpyAsync(labels_dev_, labels, voxels(N) * sizeof(int32_t), cudaMemcpyDefault, stream_));
    labelled_n_ = N;
  }
  BnParams bn_params(const Unit& u) {
    BnParams bp;
    for (int k = 0; k < 3; ++k) {
      const bool on = k < chain_num_bn(u.chain);
      bp.gamma[k] = on ? params_ + u.gamma_off[k] : nullptr;
      bp.beta[k] = on ? params_ + u.beta_off[k] : nullptr;
      bp.moving_mean[k] = on ? state_ + u.mm_off[k] : nullptr;
      bp.moving_var[k] = on ? state_ + u.mv_off[k] : nullptr;
    }
    return bp;
  }
  static bool v4_ok(int C, long long total) { return C % 4 == 0 && C >= 4 && 256 % (C / 4) == 0 && total < (1LL << 31); }
  static int v4_blocks(long long total4, int C) {  // grid stride must keep (i % C4) fixed per thread: any multiple of 256 threads does
    (void)C;
    // at most four blocks per SM: every block ends with a cross-warp combine and its partial row is re-read by the
    // finalize kernel, so fewer, longer-running blocks are cheaper than the 8 per SM the grid-stride loop would fill.
    // (A fused "last block finalizes" form was measured and dropped: the ticket atomics of ~600 blocks on one address
    // serialise in L2, +15 us per launch against the 8 us of the separate one-block-per-channel finalize kernel.)
    return static_cast<int>(std::max<long long>(1, std::min<long long>((total4 + 256 * 8 - 1) / (256 * 8), kMaxRedBlocks / 2)));
  }
  static RedGeom red_geom(int C, int c0, long long V) {
    RedGeom g;
    g.C = C;
    g.c0 = c0;
    g.CW = std::min(C - c0, 256);
    g.V = V;
    return g;
  }
  static int red_threads(const RedGeom& g) { return (kRedThreads / g.CW) * g.CW; }
  static int red_blocks(const RedGeom& g) {
    const long long per_block = static_cast<long long>(red_threads(g) / g.CW) * 16;  // >= 16 voxels per thread group
    return static_cast<int>(std::max<long long>(1, std::min<long long>((g.V + per_block - 1) / per_block, kMaxRedBlocks)));
  }
  LossCfg loss_cfg() const { return loss_cfg_of(cfg_); }
  static constexpr int kLossBlocks = 296;
  int loss_blocks() const {
    const long long Vn = voxels(1);
    return static_cast<int>(std::max<long long>(1, std::min<long long>((Vn + 2047) / 2048, kLossBlocks)));
  }

  // ---- forward --------------------------------------------------------------------------------
  void run_conv_fprop(Unit& u, int N) {
    const Act& x1 = acts_[u.in1];
    const Act& o = acts_[u.out];
    const float* bias = u.b_off >= 0 ? params_ + u.b_off : nullptr;
    if (u.kind == U_CONV5) {
      if (cfg_.precision != PREC_FP32 && u.tc.fprop.valid && !dbg_no_tc_fprop_) {
        tc_run_fprop(u, N);
        return;
      }
      Conv5Args p;
      p.in1 = x1.a;
      p.in2 = u.in2 >= 0 ? acts_[u.in2].a : nullptr;
      p.C1 = u.Cin1;
      p.C2 = u.Cin2;
      p.w = params_ + u.w_off;
      p.bias = bias;
      p.res = u.res >= 0 ? acts_[u.res].a : nullptr;
      p.out1 = u.z;
      p.out2 = nullptr;
      p.Co1 = u.Cout;
      p.Co2 = 0;
      p.acc1 = p.acc2 = 0;
      p.dims = o.dims;
      p.N = N;
      ProfScope ps(*this, 0, conv5_flops(u, N), 0, &u, "fprop");
      launch_conv5(p);
    } else if (u.kind == U_CONV3) {
      if (cfg_.precision != PREC_FP32 && u.tc.fprop.valid && !dbg_no_tc_fprop_) {
        tc_run_fprop(u, N);
        return;
      }
      Conv5Args p;
      p.in1 = x1.a;
      p.in2 = nullptr;
      p.C1 = u.Cin1;
      p.C2 = 0;
      p.w = params_ + u.w_off;
      p.bias = bias;
      p.res = nullptr;
      p.out1 = u.z;
      p.out2 = nullptr;
      p.Co1 = u.Cout;
      p.Co2 = 0;
      p.acc1 = p.acc2 = 0;
      p.dims = o.dims;
      p.N = N;
      ProfScope ps(*this, 0, conv5_flops(u, N), 0, &u, "fprop");
      launch_conv3(p);
    } else if (u.kind == U_DOWN || u.kind == U_UP) {
      K2Args p{};
      p.w = params_ + u.w_off;
      p.bias = bias;
      p.N = N;
      p.accumulate = 0;
      if (u.kind == U_DOWN) {
        p.fine_in = x1.a;
        p.coarse_out = u.z;
        p.CF = u.Cin1;
        p.CC = u.Cout;
        p.cd = o.dims;
        launch_k2_gather(p, &u.k2_fprop);
      } else {
        p.coarse_in = x1.a;
        p.fine_out = u.z;
        p.CF = u.Cout;
        p.CC = u.Cin1;
        p.cd = x1.dims;
        launch_k2_scatter(p, &u.k2_fprop);
      }
    } else if (u.kind == U_ADD) {
      const long long n = voxels_of(o.dims, N) * u.Cout;
      VNB_LAUNCH(add2_kernel, grid_for(n, 256), 256, 0, stream_, (const float*)x1.a, (const float*)acts_[u.res].a, u.z, n);
      ++launches_;
    } else if (u.kind == U_CONV1 && conv1_general(u)) {
      Conv1Args p;
      p.x = x1.a;
      p.w = params_ + u.w_off;
      p.bias = bias;
      p.res = u.res >= 0 ? acts_[u.res].a : nullptr;
      p.out = u.z;
      p.V = voxels_of(o.dims, N);
      p.K = u.Cin1;
      p.M = u.Cout;
      p.transposed = 0;
      p.accumulate = 0;
      launch_conv1g(p);
    } else if (u.kind == U_CONV1) {
      const long long V = voxels_of(o.dims, N);
      if (conv1_vox_ok(u))
        VNB_LAUNCH(conv1_fprop_vox_kernel, grid_for(V, 256), 256, 0, stream_, (const float*)x1.a,
                   (const float*)(params_ + u.w_off), bias, u.z, V, u.Cin1, u.Cout);
      else
        VNB_LAUNCH(conv1_fprop_kernel, grid_for(V * u.Cout, 256), 256, 0, stream_, (const float*)x1.a,
                   (const float*)(params_ + u.w_off), bias, u.z, V, u.Cin1, u.Cout);
      ++launches_;
    }
  }
  // the small-shape 1^3 kernels serve the V-Net head; anything with a residual or a wide product goes general
  static bool conv1_general(const Unit& u) { return u.res >= 0 || u.Cin1 * u.Cout > 256; }
  // voxel-per-thread head kernels (float4 input rows): every V-Net head (16 -> K)
  static bool conv1_vox_ok(const Unit& u) {
    return u.Cin1 % 4 == 0 && u.Cin1 <= kC1WgMaxCin && u.Cout >= 2 && u.Cout <= kC1MaxK;
  }
  void launch_conv1g(const Conv1Args& p) {
    dim3 grid(static_cast<unsigned>((p.V + 63) / 64), (p.M + 63) / 64);
    VNB_LAUNCH(conv1g_kernel, grid, 256, 0, stream_, p);
    ++launches_;
  }
  void launch_conv3(const Conv5Args& p) {
    const Dims& d = p.dims;
    const int tiles = ((d.W + kC5_TW - 1) / kC5_TW) * ((d.H + kC5_TH - 1) / kC5_TH) * ((d.D + kC5_TD - 1) / kC5_TD);
    dim3 grid(tiles, (p.Co1 + p.Co2 + kC5_CO - 1) / kC5_CO, p.N);
    launch_conv5_attr_once();
    VNB_LAUNCH(conv_ref_kernel<3>, grid, 256, ConvRefGeom<3>::SMEM, stream_, p);
    ++launches_;
  }
  GateArgs gate_args(const Unit& u, int N) {
    GateArgs g;
    g.att = acts_[u.in1].a;
    g.vnet = acts_[u.in2].a;
    g.soft = gate_soft_;
    g.masked = acts_[u.out].a;
    g.distmap = (cfg_.attention_loss != ATT_NONE && distmap_valid_) ? distmap_dev_ : nullptr;
    g.V = voxels_of(acts_[u.out].dims, N);
    g.K = u.Cout;
    g.att_loss = cfg_.attention_loss;
    return g;
  }
  void launch_conv5(const Conv5Args& p) {
    const Dims& d = p.dims;
    const int tiles = ((d.W + kC5_TW - 1) / kC5_TW) * ((d.H + kC5_TH - 1) / kC5_TH) * ((d.D + kC5_TD - 1) / kC5_TD);
    dim3 grid(tiles, (p.Co1 + p.Co2 + kC5_CO - 1) / kC5_CO, p.N);
    launch_conv5_attr_once();
    VNB_LAUNCH(conv5_ref_kernel, grid, 256, kC5_SMEM, stream_, p);
    ++launches_;
  }

  void forward(int N, float dropout, uint64_t seed, bool update_moving) {
    if (cfg_.precision != PREC_FP32) {
      tc_prepare_weights();
      if (image_cpad_ > 0) {
        const Act& im = acts_[image_act_];
        const long long V = voxels_of(im.dims, N);
        VNB_LAUNCH(split_pad_bf16_kernel, grid_for(V * image_cpad_, 256), 256, 0, stream_, (const float*)im.a, V, im.C,
                   image_cpad_, im.a_hi, im.a_lo);
        ++launches_;
      }
    }
    for (size_t ui = 0; ui < units_.size(); ++ui) {
      Unit& u = units_[ui];
      if (cfg_.precision != PREC_FP32) tc_wait_late_packs(static_cast<int>(ui));
      NvtxRange range(nvtx_, u.scope, " fwd");
      const Act& o = acts_[u.out];
      const long long V = voxels_of(o.dims, N);
      int nblk = 1, nq_stride = 2;
      if (u.kind == U_GATE) {  // train.py:295-302; also the attention-loss partial sums
        const GateArgs g = gate_args(u, N);
        att_nblk_ = grid_for(g.V, 256, kMaxRedBlocks);
        VNB_LAUNCH(gate_fwd_kernel, att_nblk_, 256, 0, stream_, g, att_partial_);
        ++launches_;
        continue;
      }
      if (u.bn_inference) {
        run_conv_fprop(u, N);
      } else if (u.kind == U_INPUT_TILE) {
        nblk = grid_for(V, 256, kMaxRedBlocks);
        VNB_LAUNCH(image_stats_kernel, nblk, 256, 0, stream_, (const float*)u.z, V, partial_);
        ++launches_;
      } else {
        fused_stats_blocks_ = 0;
        run_conv_fprop(u, N);
        if (fused_stats_blocks_ > 0) {   // the convolution's epilogue wrote the partial sums (conv_col.cuh)
          nblk = fused_stats_blocks_;
        } else if (v4_ok(u.Cout, V * u.Cout)) {
          const unsigned total4 = static_cast<unsigned>(V * u.Cout / 4);
          nblk = v4_blocks(total4, u.Cout);
          VNB_LAUNCH_PDL(bn_stats_v4_kernel, nblk, 256, 0, stream_, (const float*)u.z, u.Cout, total4, partial_);
          ++launches_;
        } else {
          RedGeom g = red_geom(u.Cout, 0, V);
          nblk = red_blocks(g);
          for (int c0 = 0; c0 < u.Cout; c0 += 256) {
            g = red_geom(u.Cout, c0, V);
            VNB_LAUNCH(bn_stats_kernel, nblk, red_threads(g), 0, stream_, (const float*)u.z, g, partial_);
            ++launches_;
          }
        }
      }
      const double* fin_partial = partial_;
      double count = static_cast<double>(V);
      if (stats_hook_ && !u.bn_inference) {  // global-batch statistics: (sum z, sum z^2) summed over the ranks
        const int PC = u.kind == U_INPUT_TILE ? 1 : u.Cout;
        VNB_LAUNCH(partial_collapse_kernel, nq_stride * PC, 128, 0, stream_, (const double*)partial_, nblk, nq_stride, PC, sync_buf_);
        ++launches_;
        stats_hook_(sync_buf_, nq_stride * PC);
        fin_partial = sync_buf_;
        nblk = 1;
        count *= stats_world_;
      }
      VNB_LAUNCH_PDL(bn_finalize_fwd_kernel, u.Cout, 128, 0, stream_, fin_partial, nblk, nq_stride,
                 u.Cout, count, u.chain, bn_params(u), u.kind == U_INPUT_TILE ? 1 : 0,
                 update_moving ? 1 : 0, u.mean, u.var, u.scale, u.shift, u.bn_inference ? 1 : 0);
      ++launches_;
      ApplyArgs ap;
      ap.z = u.z;
      ap.a = acts_[u.out].a;
      ap.a_hi = acts_[u.out].a_hi;
      ap.a_lo = acts_[u.out].a_lo;
      ap.scale = u.scale;
      ap.shift = u.shift;
      ap.alpha = u.relu ? zero_alpha_ : (u.has_act ? params_ + u.alpha_off : nullptr);
      ap.total = V * u.Cout;
      ap.C = u.Cout;
      ap.tiled_input = u.kind == U_INPUT_TILE ? 1 : 0;
      ap.drop_rate = u.has_dropout ? dropout : 0.f;
      ap.seed = seed;
      ap.unit = static_cast<uint32_t>(ui);
      if (v4_ok(u.Cout, ap.total))
        VNB_LAUNCH_PDL(bn_apply_v4_kernel, grid_for(ap.total / 4, 256), 256, 0, stream_, ap);
      else
        VNB_LAUNCH(bn_apply_kernel, grid_for(ap.total, 256), 256, 0, stream_, ap);
      ++launches_;
    }
  }

  void loss_forward(int N) {
    const LossCfg lc = loss_cfg();
    const long long Vn = voxels(1);
    const int nblk = loss_blocks();
    dim3 grid(nblk, N);
    VNB_LAUNCH(softmax_loss_fwd_kernel, grid, 256, 0, stream_, (const float*)acts_[head_act_].a,
               (const int32_t*)labels_dev_, Vn, lc, (float*)nullptr, (long long*)nullptr, loss_partial_);
    const bool att = cfg_.attention && cfg_.attention_loss != ATT_NONE;
    if (att && !distmap_valid_) throw std::invalid_argument("attention loss configured but no distance map uploaded (vnb_set_distmap)");
    VNB_LAUNCH(loss_finalize_kernel, 1, 64, 0, stream_, (const double*)loss_partial_, N, nblk, Vn, lc, terms_dev_, coef_dev_, loss_dev_,
               (const double*)att_partial_, att ? att_nblk_ : 0, att_elements(N));
    launches_ += 2;
  }

  // number of elements the attention loss averages over (train.py:392,398)
  double att_elements(int N) const {
    return static_cast<double>(voxels(N)) * (cfg_.attention_loss == ATT_ABS ? 2.0 : 1.0);
  }

  // ---- backward -------------------------------------------------------------------------------
  void backward(int N, float dropout, uint64_t seed) {
    if (cfg_.precision != PREC_FP32) tc_wait_late_packs(static_cast<int>(units_.size()));
    const LossCfg lc = loss_cfg();
    const long long Vn = voxels(1);
    dim3 lgrid(loss_blocks(), N);
    VNB_LAUNCH(softmax_loss_bwd_kernel, lgrid, 256, 0, stream_, (const float*)acts_[head_act_].a,
               (const int32_t*)labels_dev_, Vn, lc, (const float*)coef_dev_, 1.0f, acts_[head_act_].d);
    ++launches_;
    for (int ui = static_cast<int>(units_.size()) - 1; ui >= 0; --ui) {
      Unit& u = units_[ui];
      NvtxRange range(nvtx_, u.scope, " bwd");
      Act& o = acts_[u.out];
      const long long V = voxels_of(o.dims, N);
      if (u.kind == U_GATE) {
        const GateArgs g = gate_args(u, N);
        VNB_LAUNCH(gate_bwd_kernel, grid_for(g.V, 256, kMaxRedBlocks), 256, 0, stream_, g, (const float*)o.d, acts_[u.in1].d,
                   u.in1_accumulate ? 1 : 0, acts_[u.in2].d, u.in2_accumulate ? 1 : 0,
                   g.distmap ? static_cast<float>(1.0 / att_elements(N)) : 0.f);
        ++launches_;
        notify_bucket(ui);
        continue;
      }
      BwdArgs b;
      b.z = u.z;
      b.d = o.d;
      b.d_hi = (u.kind == U_CONV5 || u.kind == U_CONV3) ? o.d_hi : nullptr;
      b.d_lo = (u.kind == U_CONV5 || u.kind == U_CONV3) ? o.d_lo : nullptr;
      b.res_grad = u.res >= 0 ? acts_[u.res].d : nullptr;
      b.res_accumulate = u.res_accumulate ? 1 : 0;
      b.res_grad2 = u.kind == U_ADD ? acts_[u.in1].d : nullptr;
      b.res_accumulate2 = u.in1_accumulate ? 1 : 0;
      b.scale = u.scale;
      b.shift = u.shift;
      b.alpha = u.relu ? zero_alpha_ : (u.has_act ? params_ + u.alpha_off : nullptr);
      b.mean = u.mean;
      b.P = u.P;
      b.Q = u.Q;
      b.S = u.S;
      b.C = u.Cout;
      b.tiled_input = u.kind == U_INPUT_TILE ? 1 : 0;
      b.drop_rate = u.has_dropout ? dropout : 0.f;
      b.seed = seed;
      b.unit = static_cast<uint32_t>(ui);
      const bool v4 = v4_ok(u.Cout, V * u.Cout);
      int nblk;
      BnGradPtrs gp;
      for (int k = 0; k < 3; ++k) {
        const bool on = k < chain_num_bn(u.chain);
        gp.dgamma[k] = on ? grads_ + u.gamma_off[k] : nullptr;
        gp.dbeta[k] = on ? grads_ + u.beta_off[k] : nullptr;
      }
      gp.dalpha = (u.has_act && !u.relu) ? grads_ + u.alpha_off : nullptr;
      gp.dbias = u.bn_inference ? grads_ + u.b_off : nullptr;
      if (v4) {
        const unsigned total4 = static_cast<unsigned>(V * u.Cout / 4);
        nblk = v4_blocks(total4, u.Cout);
        VNB_LAUNCH_PDL(bn_bwd_reduce_v4_kernel, nblk, 256, 0, stream_, b, total4, partial_);
        ++launches_;
      } else {
        RedGeom g = red_geom(u.Cout, 0, V);
        nblk = red_blocks(g);
        for (int c0 = 0; c0 < u.Cout; c0 += 256) {
          g = red_geom(u.Cout, c0, V);
          VNB_LAUNCH(bn_bwd_reduce_kernel, nblk, red_threads(g), 0, stream_, b, g, partial_);
          ++launches_;
        }
      }
      const double* fin_partial = partial_;
      const double* gsum = nullptr;
      if (stats_hook_ && !u.bn_inference) {  // R0, R1 of the global batch for dL/dz; parameter gradients stay local
        double* local = sync_buf_;
        double* global = sync_buf_ + static_cast<size_t>(3) * sync_stride_;
        VNB_LAUNCH(partial_collapse_kernel, 3 * u.Cout, 128, 0, stream_, (const double*)partial_, nblk, 3, u.Cout, local);
        ++launches_;
        VNB_CUDA_OK(cudaMemcpyAsync(global, local, sizeof(double) * 2 * u.Cout, cudaMemcpyDeviceToDevice, stream_));
        stats_hook_(global, 2 * u.Cout);
        fin_partial = local;
        gsum = global;
        nblk = 1;
      }
      VNB_LAUNCH_PDL(bn_finalize_bwd_kernel, u.Cout, 128, 0, stream_, fin_partial, nblk, u.Cout,
                 static_cast<double>(V), u.chain, bn_params(u), (const double*)u.var, gp, u.P, u.Q, u.S,
                 u.bn_inference ? 1 : 0, gsum, static_cast<double>(V) * stats_world_);
      ++launches_;
      if (u.kind == U_INPUT_TILE) {  // image needs no gradient
        notify_bucket(ui);
        continue;
      }
      if (v4)
        VNB_LAUNCH_PDL(bn_bwd_apply_v4_kernel, grid_for(V * u.Cout / 4, 256), 256, 0, stream_, b, V * u.Cout);
      else
        VNB_LAUNCH(bn_bwd_apply_kernel, grid_for(V * u.Cout, 256), 256, 0, stream_, b, V * u.Cout);
      ++launches_;
      if (u.kind != U_ADD) run_conv_backward(u, N);
      notify_bucket(ui);
    }
    join_wgrad_stream();   // the optimiser (and the next step's passes) read what the filter-gradient stream wrote
  }
  // stream on which the filter gradient of the current unit is launched: the side stream, ordered after everything
  // the main stream has enqueued so far (dz and its bf16 copies are complete)
  cudaStream_t wgrad_stream_begin() {
#ifndef VNB_EMULATE
    if (wg_stream_ && !profiling_) {   // per-kernel profiling times every launch alone on the main stream
      VNB_CUDA_OK(cudaEventRecord(wg_ready_ev_, stream_));
      VNB_CUDA_OK(cudaStreamWaitEvent(wg_stream_, wg_ready_ev_, 0));
      wg_pending_ = true;
      return wg_stream_;
    }
#endif
    return stream_;
  }
  void join_wgrad_stream() {
#ifndef VNB_EMULATE
    if (wg_stream_ && wg_pending_) {
      VNB_CUDA_OK(cudaEventRecord(wg_done_ev_, wg_stream_));
      VNB_CUDA_OK(cudaStreamWaitEvent(stream_, wg_done_ev_, 0));
      wg_pending_ = false;
    }
#endif
  }
  void notify_bucket(int ui) {
    if (!grad_hook_) return;
    bool any = false;
    for (size_t b = 0; b < buckets_.size(); ++b) any = any || buckets_[b].unit_lo == ui;
    if (any && !comm_waits_wgrad_) join_wgrad_stream();   // the all-reduce of a bucket waits on the main stream only
    for (size_t b = 0; b < buckets_.size(); ++b)
      if (buckets_[b].unit_lo == ui) grad_hook_(static_cast<int>(b));
  }

  void run_conv_backward(Unit& u, int N) {
    const Act& x1 = acts_[u.in1];
    const Act& o = acts_[u.out];
    float* dw = grads_ + u.w_off;
    const float* dz = o.d;
    VNB_CUDA_OK(cudaMemsetAsync(dw, 0, u.w_count * sizeof(float), stream_));
    if (!u.bn_inference) VNB_CUDA_OK(cudaMemsetAsync(grads_ + u.b_off, 0, u.Cout * sizeof(float), stream_));
    if (u.kind == U_CONV3) {
      if (cfg_.precision != PREC_FP32 && u.need_dgrad && u.tc.dgrad.valid && !dbg_no_tc_dgrad_) {
        tc_run_dgrad(u, N);
      } else if (u.need_dgrad) {
        VNB_LAUNCH(flip_transpose_w_kernel, grid_for(static_cast<long long>(u.w_count), 256), 256, 0, stream_,
                   (const float*)(params_ + u.w_off), wflip_, u.Cin1, u.Cout, 27);
        ++launches_;
        Conv5Args p;
        p.in1 = dz;
        p.in2 = nullptr;
        p.C1 = u.Cout;
        p.C2 = 0;
        p.w = wflip_;
        p.bias = nullptr;
        p.res = nullptr;
        p.out1 = x1.d;
        p.Co1 = u.Cin1;
        p.acc1 = u.in1_accumulate ? 1 : 0;
        p.out2 = nullptr;
        p.Co2 = 0;
        p.acc2 = 0;
        p.dims = o.dims;
        p.N = N;
        ProfScope ps(*this, 0, conv5_flops(u, N), 0, &u, "dgrad");
        launch_conv3(p);
      }
      if (cfg_.precision != PREC_FP32 && u.tc.wgrad.valid && !dbg_no_tc_wgrad_) {
        tc_run_wgrad(u, N);
        return;
      }
      Wgrad5Args w;
      w.in1 = x1.a;
      w.in2 = nullptr;
      w.C1 = u.Cin1;
      w.C2 = 0;
      w.dz = dz;
      w.Cout = u.Cout;
      w.dw = dw;
      w.dims = o.dims;
      w.N = N;
      const Dims& d = o.dims;
      const long long ntiles = static_cast<long long>((d.W + kW5_TW - 1) / kW5_TW) * ((d.H + kW5_TH - 1) / kW5_TH) *
                               ((d.D + kW5_TD - 1) / kW5_TD) * N;
      const int pairs = ((u.Cin1 + 15) / 16) * ((u.Cout + 15) / 16);
      long long splits = std::max<long long>(1, std::min<long long>(ntiles, (2 * 1184 + pairs - 1) / pairs));
      w.tiles_per_block = static_cast<int>((ntiles + splits - 1) / splits);
      splits = (ntiles + w.tiles_per_block - 1) / w.tiles_per_block;
      dim3 grid(static_cast<unsigned>(splits), pairs);
      launch_conv5_attr_once();
      {
        ProfScope ps(*this, 1, conv5_flops(u, N), 0, &u, "wgrad");
        VNB_LAUNCH(conv_wgrad_ref_kernel<3>, grid, 256, WgradRefGeom<3>::SMEM, stream_, w);
      }
      ++launches_;
    } else if (u.kind == U_CONV5) {
      const int Cin = u.Cin1 + u.Cin2;
      const bool tc = cfg_.precision != PREC_FP32;
      if (tc && u.need_dgrad && u.tc.dgrad.valid && !dbg_no_tc_dgrad_) {
        tc_run_dgrad(u, N);
      } else if (u.need_dgrad) {
        VNB_LAUNCH(flip_transpose_w_kernel, grid_for(static_cast<long long>(u.w_count), 256), 256, 0, stream_,
                   (const float*)(params_ + u.w_off), wflip_, Cin, u.Cout, 125);
        ++launches_;
        Conv5Args p;
        p.in1 = dz;
        p.in2 = nullptr;
        p.C1 = u.Cout;
        p.C2 = 0;
        p.w = wflip_;
        p.bias = nullptr;
        p.res = nullptr;
        p.out1 = x1.d;
        p.Co1 = u.Cin1;
        p.acc1 = u.in1_accumulate ? 1 : 0;
        p.out2 = u.in2 >= 0 ? acts_[u.in2].d : nullptr;
        p.Co2 = u.Cin2;
        p.acc2 = u.in2_accumulate ? 1 : 0;
        p.dims = o.dims;
        p.N = N;
        ProfScope ps(*this, 0, conv5_flops(u, N), 0, &u, "dgrad");
        launch_conv5(p);
      }
      if (tc && u.tc.wgrad.valid && !dbg_no_tc_wgrad_) {
        tc_run_wgrad(u, N);
        return;
      }
      Wgrad5Args w;
      w.in1 = x1.a;
      w.in2 = u.in2 >= 0 ? acts_[u.in2].a : nullptr;
      w.C1 = u.Cin1;
      w.C2 = u.Cin2;
      w.dz = dz;
      w.Cout = u.Cout;
      w.dw = dw;
      w.dims = o.dims;
      w.N = N;
      const Dims& d = o.dims;
      const long long ntiles = static_cast<long long>((d.W + kW5_TW - 1) / kW5_TW) * ((d.H + kW5_TH - 1) / kW5_TH) *
                               ((d.D + kW5_TD - 1) / kW5_TD) * N;
      const int pairs = ((Cin + 15) / 16) * ((u.Cout + 15) / 16);
      long long splits = std::max<long long>(1, std::min<long long>(ntiles, (2 * 1184 + pairs - 1) / pairs));
      w.tiles_per_block = static_cast<int>((ntiles + splits - 1) / splits);
      splits = (ntiles + w.tiles_per_block - 1) / w.tiles_per_block;
      dim3 grid(static_cast<unsigned>(splits), pairs);
      launch_conv5_attr_once();
      {
        ProfScope ps(*this, 1, conv5_flops(u, N), 0, &u, "wgrad");
        VNB_LAUNCH(conv5_wgrad_ref_kernel, grid, 256, kW5_SMEM, stream_, w);
      }
      ++launches_;
    } else if (u.kind == U_DOWN || u.kind == U_UP) {
      K2Args p{};
      p.w = params_ + u.w_off;
      p.dw = dw;
      p.N = N;
      if (u.kind == U_DOWN) {  // z coarse, x fine
        p.CF = u.Cin1;
        p.CC = u.Cout;
        p.cd = o.dims;
        p.coarse_in = dz;
        p.fine_in = x1.a;
        if (u.need_dgrad) {
          p.fine_out = x1.d;
          p.accumulate = u.in1_accumulate ? 1 : 0;
          launch_k2_scatter(p, &u.k2_dgrad);
        }
      } else {  // z fine, x coarse
        p.CF = u.Cout;
        p.CC = u.Cin1;
        p.cd = x1.dims;
        p.fine_in = dz;
        if (u.need_dgrad) {
          p.coarse_out = x1.d;
          p.accumulate = u.in1_accumulate ? 1 : 0;
          launch_k2_gather(p, &u.k2_dgrad);
        }
        p.coarse_in = x1.a;
      }
      p.bias = nullptr;
      launch_k2_wgrad(p, &u.k2_wgrad);
    } else if (u.kind == U_CONV1 && conv1_general(u)) {
      const long long V = voxels_of(o.dims, N);
      if (u.need_dgrad) {
        Conv1Args p;
        p.x = dz;
        p.w = params_ + u.w_off;  // [Cin][Cout] read as the transposed product
        p.bias = nullptr;
        p.res = nullptr;
        p.out = x1.d;
        p.V = V;
        p.K = u.Cout;
        p.M = u.Cin1;
        p.transposed = 1;
        p.accumulate = u.in1_accumulate ? 1 : 0;
        launch_conv1g(p);
      }
      const int gy = (u.Cin1 + 63) / 64, gz = (u.Cout + 63) / 64;
      long long splits = std::max<long long>(1, std::min<long long>((V + 255) / 256, (4 * 148 + gy * gz - 1) / (gy * gz)));
      const long long vpb = ((V + splits - 1) / splits + 31) / 32 * 32;
      splits = (V + vpb - 1) / vpb;
      dim3 grid(static_cast<unsigned>(splits), gy, gz);
      VNB_LAUNCH(conv1g_wgrad_kernel, grid, 256, 0, stream_, (const float*)x1.a, dz, dw, V, u.Cin1, u.Cout, vpb);
      ++launches_;
    } else if (u.kind == U_CONV1) {
      const long long V = voxels_of(o.dims, N);
      const bool vox = conv1_vox_ok(u);
      if (u.need_dgrad) {
        if (vox)
          VNB_LAUNCH(conv1_dgrad_vox_kernel, grid_for(V, 256), 256, 0, stream_, dz, (const float*)(params_ + u.w_off), x1.d, V,
                     u.Cin1, u.Cout, u.in1_accumulate ? 1 : 0);
        else
          VNB_LAUNCH(conv1_dgrad_kernel, grid_for(V * u.Cin1, 256), 256, 0, stream_, dz, (const float*)(params_ + u.w_off),
                     x1.d, V, u.Cin1, u.Cout, u.in1_accumulate ? 1 : 0);
        ++launches_;
      }
      if (u.Cin1 * u.Cout > 256) throw std::runtime_error("output layer wider than 256 (Cin*K) is not supported");
      if (vox) {
        const int blocks = static_cast<int>(std::min<long long>((V + 255) / 256, 4 * 148));
        VNB_LAUNCH(conv1_wgrad_vox_kernel, blocks, 256, 0, stream_, (const float*)x1.a, dz, dw, V, u.Cin1, u.Cout);
      } else {
        const int vpb = 4096;
        const int blocks = static_cast<int>((V + vpb - 1) / vpb);
        VNB_LAUNCH(conv1_wgrad_kernel, blocks, 256, 0, stream_, (const float*)x1.a, dz, dw, V, u.Cin1, u.Cout, vpb);
      }
      ++launches_;
    }
  }
  static bool k2_tiled_ok(const K2Args& p) { return p.CF % 16 == 0 && p.CC % 16 == 0; }
  void launch_k2_gather(const K2Args& p, K2TcPlan* tcp = nullptr) {
    const long long M = static_cast<long long>(p.N) * p.cd.D * p.cd.H * p.cd.W;
    if (M > 0xFFFFFFFFLL) throw std::invalid_argument("2x2x2 convolution: more than 2^32 coarse voxels per batch");
    if (tcp && tcp->valid && cfg_.precision != PREC_FP32) {   // tcgen05 + TMA gather (k2_tc.cuh)
      k2tc_launch(*tcp, p.N, p.bias, p.accumulate != 0, sm_count_, stream_);
      ++launches_;
      return;
    }
    if (k2_tiled_ok(p)) {
      dim3 grid(static_cast<unsigned>((M + kK2_BM - 1) / kK2_BM), (p.CC + kK2_BN - 1) / kK2_BN);
      if (cfg_.precision != PREC_FP32) VNB_LAUNCH(k2_gather_mma_kernel, grid, 256, 0, stream_, p, M);
      else VNB_LAUNCH(k2_gather_tiled_kernel<false>, grid, 256, 0, stream_, p, M);
    } else {
      VNB_LAUNCH(k2_gather_kernel, grid_for(M * p.CC, 256), 256, 0, stream_, p);
    }
    ++launches_;
  }
  void launch_k2_scatter(const K2Args& p, K2TcPlan* tcp = nullptr) {
    const long long M = static_cast<long long>(p.N) * p.cd.D * p.cd.H * p.cd.W;
    if (M > 0xFFFFFFFFLL) throw std::invalid_argument("2x2x2 convolution: more than 2^32 coarse voxels per batch");
    if (tcp && tcp->valid && cfg_.precision != PREC_FP32) {   // tcgen05 + TMA depth-to-space scatter (k2_tc.cuh)
      k2tc_launch(*tcp, p.N, p.bias, p.accumulate != 0, sm_count_, stream_);
      ++launches_;
      return;
    }
    if (k2_tiled_ok(p)) {
      dim3 grid(static_cast<unsigned>((M + kK2_BM - 1) / kK2_BM), (8 * p.CF + kK2_BN - 1) / kK2_BN);
      if (cfg_.precision != PREC_FP32) VNB_LAUNCH(k2_scatter_mma_kernel, grid, 256, 0, stream_, p, M);
      else VNB_LAUNCH(k2_scatter_tiled_kernel<false>, grid, 256, 0, stream_, p, M);
    } else {
      VNB_LAUNCH(k2_scatter_kernel, grid_for(M * 8 * p.CF, 256), 256, 0, stream_, p);
    }
    ++launches_;
  }
  void launch_k2_wgrad(const K2Args& p, K2WgPlan* tcp = nullptr) {
    const long long M = static_cast<long long>(p.N) * p.cd.D * p.cd.H * p.cd.W;
    if (M > 0xFFFFFFFFLL) throw std::invalid_argument("2x2x2 convolution: more than 2^32 coarse voxels per batch");
    if (tcp && tcp->valid && cfg_.precision != PREC_FP32) {   // tcgen05 filter gradient (k2_tc.cuh)
      k2wg_launch(*tcp, p.N, p.dw, sm_count_, stream_);
      ++launches_;
      return;
    }
    if (k2_tiled_ok(p)) {
      const int gx = (8 * p.CF + kK2_BM - 1) / kK2_BM, gy = (p.CC + kK2_BN - 1) / kK2_BN;
      long long splits = std::max<long long>(1, std::min<long long>((M + 255) / 256, (4 * 148 + gx * gy - 1) / (gx * gy)));
      long long mps = ((M + splits - 1) / splits + kK2_BK - 1) / kK2_BK * kK2_BK;
      splits = (M + mps - 1) / mps;
      dim3 grid(gx, gy, static_cast<unsigned>(splits));
      if (cfg_.precision != PREC_FP32) VNB_LAUNCH(k2_wgrad_mma_kernel, grid, 256, 0, stream_, p, M, mps);
      else VNB_LAUNCH(k2_wgrad_tiled_kernel<false>, grid, 256, 0, stream_, p, M, mps);
    } else {
      const long long outs = 8LL * p.CF * p.CC;
      const int oblocks = static_cast<int>((outs + 255) / 256);
      long long splits = std::max<long long>(1, std::min<long long>(M, (2 * 1184 + oblocks - 1) / oblocks));
      const int vps = static_cast<int>((M + splits - 1) / splits);
      splits = (M + vps - 1) / vps;
      dim3 grid(oblocks, static_cast<unsigned>(splits));
      VNB_LAUNCH(k2_wgrad_kernel, grid, 256, 0, stream_, p, vps);
    }
    ++launches_;
  }
  double conv5_flops(const Unit& u, int N) const {  // 2*MAC of one 5^3 (or 3^3) pass (fprop = dgrad = wgrad)
    return 2.0 * (u.kind == U_CONV3 ? 27.0 : 125.0) * (u.Cin1 + u.Cin2) * u.Cout * static_cast<double>(voxels_of(acts_[u.out].dims, N));
  }
  void launch_conv5_attr_once() {
#ifndef VNB_EMULATE
    static bool attr_set = false;
    if (!attr_set) {
      VNB_CUDA_OK(cudaFuncSetAttribute(conv5_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC5_SMEM));
      VNB_CUDA_OK(cudaFuncSetAttribute(conv5_wgrad_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kW5_SMEM));
      VNB_CUDA_OK(cudaFuncSetAttribute(conv_ref_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvRefGeom<3>::SMEM));
      VNB_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_ref_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgradRefGeom<3>::SMEM));
      attr_set = true;
    }
#endif
  }

  // tensor-core path (conv_tc.cuh / conv_tc_impl.cuh)
  void tc_setup();
  void tc_prepare_weights();
  void tc_wait_late_packs(int ui);
  void tc_run_fprop(Unit& u, int N);
  void tc_run_dgrad(Unit& u, int N);
  void tc_run_wgrad(Unit& u, int N);
  int sm_count_ = 148;
  int fused_stats_blocks_ = 0;   // > 0: the last tc_run_fprop produced the BN partial sums in partial_ (that many rows)
  int image_cpad_ = 0;
  K2PackJob* k2_pack_jobs_dev_ = nullptr;            // weight images of the tcgen05 2^3 kernels (one launch per step)
  int k2_pack_blocks_ = 0, k2_pack_njobs_ = 0;
  PackJob* pack_jobs_dev_[2] = {nullptr, nullptr};   // [0] early forward packs (compute stream), [1] the rest (side stream)
  int pack_blocks_[2] = {0, 0}, pack_njobs_[2] = {0, 0};
  bool pack_built_ = false, late_packs_pending_ = false;
  int pack_first_late_unit_ = 0;
  float* wg_partial_ = nullptr;

  EngineConfig cfg_;
  cudaStream_t stream_ = 0;
  cudaStream_t wg_stream_ = 0;   // filter-gradient side stream (null: everything on stream_)
#ifndef VNB_EMULATE
  cudaEvent_t wg_ready_ev_ = nullptr, wg_done_ev_ = nullptr, pack_done_ev_ = nullptr;
#endif
  bool wg_pending_ = false;
  bool comm_waits_wgrad_ = false;
  bool dbg_no_tc_fprop_ = false, dbg_no_tc_dgrad_ = false, dbg_no_tc_wgrad_ = false;
  bool nvtx_ = false;   // VNB_NVTX=1: NVTX range per unit and pass
  std::vector<ParamEntry> entries_;
  std::map<std::string, size_t> index_;
  std::vector<Act> acts_;
  std::vector<Unit> units_;
  std::vector<void*> allocs_;
  std::vector<Bucket> buckets_;
  std::function<void(int)> grad_hook_;
  std::function<void(double*, int)> stats_hook_;   // synchronised batch norm: sum over the ranks (empty = local statistics)
  int stats_world_ = 1;
  double* sync_buf_ = nullptr;                       // [3][sync_stride_] collapsed local sums + [2][sync_stride_] exchanged sums
  int sync_stride_ = 0;
  size_t n_train_ = 0, n_state_ = 0;
  int image_act_ = -1, head_act_ = -1;
  int vnet_logits_act_ = -1, att_logits_act_ = -1, gate_unit_ = -1;
  float *zero_alpha_ = nullptr, *gate_soft_ = nullptr, *distmap_dev_ = nullptr;
  double* att_partial_ = nullptr;
  int att_nblk_ = 0;
  bool distmap_valid_ = false;
  float *params_ = nullptr, *grads_ = nullptr, *adam_m_ = nullptr, *adam_v_ = nullptr, *state_ = nullptr;
  double* partial_ = nullptr;
  float* wflip_ = nullptr;
  int32_t* labels_dev_ = nullptr;
  unsigned long long* metrics_dev_ = nullptr;   // confusion matrix + AUC histograms of read_metrics
  cudaStream_t copy_stream_ = 0;                  // input staging (stage_batch / commit_staged), created on first use
  cudaEvent_t staged_ev_{}, staging_free_ev_{};
  float* stage_img_ = nullptr;
  int32_t* stage_lab_ = nullptr;
  int staged_n_ = 0;
  int labelled_n_ = 0;                            // batch size of the labels that belong to the logits on the device
  float* softmax_dev_ = nullptr;
  long long* argmax_dev_ = nullptr;
  double *loss_partial_ = nullptr, *terms_dev_ = nullptr;
  float *coef_dev_ = nullptr, *loss_dev_ = nullptr;
  long long global_step_ = 0;
  float grad_scale_ = 1.0f;
  bool weights_dirty_ = true;
  long long launches_ = 0;
  bool profiling_ = false;
  struct ProfRec {
    cudaEvent_t a{}, b{};
    int cls = 0;
    double flops = 0;
    const Unit* unit = nullptr;   // the layer this launch belongs to (per-layer roofline table)
    const char* pass = "";        // "fprop" | "dgrad" | "wgrad"
  };
  std::vector<ProfRec> prof_;
  size_t prof_used_ = 0;
  cudaEvent_t timer_ev_[2] = {};
};

}  // namespace vnb
