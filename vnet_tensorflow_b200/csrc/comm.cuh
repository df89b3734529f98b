// Data-parallel gradient exchange (new functionality: the reference is single-device).
//
// One process per GPU.  The flat gradient buffer is cut into buckets along unit boundaries; as soon
// as the backward pass has produced a bucket (Engine grad hook) it is all-reduced on a side stream
// while the remaining backward kernels keep running on the compute stream; the optimiser step waits
// on the side stream and folds the 1/world mean into its own pass.
//
// Three schedules over the same NCCL communicator (NVLink 5 / NVSwitch transport), VNB_ALLREDUCE=ring|direct|nccl:
//   ring   : hand-rolled reduce-scatter + all-gather built from ncclSend/ncclRecv pairs to the ring neighbours, with our
//            own fp32 accumulate kernel between hops: 2 (W-1) sequential hops per bucket
//   direct : the same reduce-scatter + all-gather with the ring unrolled over the switch: every rank sends chunk c straight
//            to rank c (one grouped exchange), sums the W-1 copies of its own chunk in rank order with one kernel, and
//            sends the result straight to every peer (second grouped exchange).  NVSwitch gives every pair of GPUs full
//            bandwidth, so the 2 (W-1) dependent hops of the ring collapse to two; identical to the ring at W = 2.
//   nccl   : ncclAllReduce (lets NCCL pick NVLS / tree), for comparison
// NCCL is resolved with dlopen so that the library binds to whatever libnccl.so.2 the process
// already carries (PyTorch ships its own) instead of pinning a second copy.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "engine.cuh"

namespace vnb {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;  // NCCL >= 2.18 (optional)
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  static NcclApi& get() {
    static NcclApi api = load();
    return api;
  }
  static NcclApi load() {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error(std::string("NCCL: cannot dlopen libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) throw std::runtime_error(std::string("NCCL: missing symbol ") + n);
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.CommSplit = reinterpret_cast<decltype(a.CommSplit)>(dlsym(h, "ncclCommSplit"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    return a;
  }
};

#define VNB_NCCL_OK(expr)                                                                      \
  do {                                                                                         \
    ncclResult_t r__ = (expr);                                                                 \
    if (r__ != ncclSuccess)                                                                    \
      throw std::runtime_error(std::string("NCCL error: ") + NcclApi::get().GetErrorString(r__) + \
                               " at " + __FILE__ + ":" + std::to_string(__LINE__));            \
  } while (0)

__global__ void accumulate_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] += src[i];
}
// dst += src[0] + src[1] + ... in source order (the direct schedule: every peer's copy of this rank's chunk)
__global__ void accumulate_multi_kernel(float* __restrict__ dst, const float* __restrict__ src, long long stride, int nsrc,
                                        long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = dst[i];
    for (int k = 0; k < nsrc; ++k) s += src[k * stride + i];
    dst[i] = s;
  }
}

class Comm {
 public:
  static void unique_id(void* out128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    VNB_NCCL_OK(NcclApi::get().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
  }

  Comm(int rank, int world, const void* uid, Engine& e) : rank_(rank), world_(world) {
    ncclUniqueId id;
    memcpy(&id, uid, sizeof(id));
    VNB_NCCL_OK(NcclApi::get().CommInitRank(&comm_, world, id, rank));
    VNB_CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    const char* mode = getenv("VNB_ALLREDUCE");
    const std::string m = mode ? mode : "direct";
    schedule_ = m == "nccl" ? 2 : (m == "ring" ? 0 : 1);
    const auto& buckets = e.buckets();
    size_t max_chunk = 0;
    for (const auto& b : buckets) max_chunk = std::max(max_chunk, chunk_elems(b.hi - b.lo));
    events_.resize(buckets.size());
    for (auto& ev : events_) VNB_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    wg_events_.resize(buckets.size());
    for (auto& ev : wg_events_) VNB_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    VNB_CUDA_OK(cudaEventCreateWithFlags(&done_, cudaEventDisableTiming));
    max_chunk_ = std::max<size_t>(max_chunk, 4);
    if (world_ > 1) VNB_CUDA_OK(cudaMalloc(&scratch_, max_chunk_ * static_cast<size_t>(world_ - 1) * sizeof(float)));
    e.set_grad_hook([this, &e](int bucket) { this->on_bucket_ready(e, bucket); });
  }
  ~Comm() {
    if (stats_comm_) NcclApi::get().CommDestroy(stats_comm_);
    if (scratch_) cudaFree(scratch_);
    for (auto ev : events_) cudaEventDestroy(ev);
    for (auto ev : wg_events_) cudaEventDestroy(ev);
    cudaEventDestroy(done_);
    cudaStreamDestroy(stream_);
    if (comm_) NcclApi::get().CommDestroy(comm_);
  }
  int rank() const { return rank_; }
  int world() const { return world_; }

  void begin_step(Engine&) {
    pending_ = 0;
    deferred_.clear();
  }

  // Synchronised batch norm: the per-channel statistic rows are summed with ncclAllReduce on the *compute* stream
  // (they sit on the critical path between a reduction kernel and its finalize kernel), through a communicator of
  // their own so that they never queue behind the gradient buckets on the side stream.  Every rank issues the same
  // sequence of collectives on both communicators (the schedule is data independent).  Collective call.
  void enable_sync_bn(Engine& e, bool on) {
    // Two communicators of one process with collectives in flight on two streams can deadlock when the ranks' GPUs
    // schedule them in different orders (NCCL's documented constraint).  With synchronised batch norm the gradient
    // buckets are therefore not exchanged under the backward pass: they are deferred until the last statistics
    // all-reduce of the step has been issued (finish_allreduce), so at most one communicator is active at any time.
    defer_buckets_ = on && world_ > 1;
    if (!on || world_ == 1) {
      e.set_stats_hook(nullptr, 1);
      return;
    }
    NcclApi& api = NcclApi::get();
    if (!stats_comm_) {
      if (!api.CommSplit) throw std::runtime_error("NCCL: synchronised batch norm needs ncclCommSplit (NCCL >= 2.18)");
      VNB_NCCL_OK(api.CommSplit(comm_, 0, rank_, &stats_comm_, nullptr));
    }
    ncclComm_t c = stats_comm_;
    e.set_stats_hook(
        [c, &e](double* dev, int n) {
          VNB_NCCL_OK(NcclApi::get().AllReduce(dev, dev, static_cast<size_t>(n), ncclDouble, ncclSum, c, e.stream()));
        },
        world_);
  }

  // called from Engine::backward (host side, in stream order) when bucket `bi` is complete
  void on_bucket_ready(Engine& e, int bi) {
    if (world_ == 1) return;
    if (defer_buckets_) {
      deferred_.push_back(bi);
      return;
    }
    exchange_bucket(e, bi);
  }

  void exchange_bucket(Engine& e, int bi) {
    const Engine::Bucket& b = e.buckets()[bi];
    VNB_CUDA_OK(cudaEventRecord(events_[bi], e.stream()));
    VNB_CUDA_OK(cudaStreamWaitEvent(stream_, events_[bi], 0));
    if (cudaStream_t ws = e.pending_wgrad_stream()) {  // dual-wait mode: filter gradients of this bucket still in flight
      VNB_CUDA_OK(cudaEventRecord(wg_events_[bi], ws));
      VNB_CUDA_OK(cudaStreamWaitEvent(stream_, wg_events_[bi], 0));
    }
    float* g = e.grad_buffer() + b.lo;
    const size_t n = b.hi - b.lo;
    if (schedule_ == 0)
      ring_allreduce(g, n);
    else if (schedule_ == 1)
      direct_allreduce(g, n);
    else
      VNB_NCCL_OK(NcclApi::get().AllReduce(g, g, n, ncclFloat, ncclSum, comm_, stream_));
    ++pending_;
  }

  // compute stream waits for all bucket all-reduces before the optimiser reads the gradients
  void finish_allreduce(Engine& e) {
    if (world_ == 1) return;
    for (int bi : deferred_) exchange_bucket(e, bi);   // synchronised batch norm: after the step's last statistics exchange
    deferred_.clear();
    VNB_CUDA_OK(cudaEventRecord(done_, stream_));
    VNB_CUDA_OK(cudaStreamWaitEvent(e.stream(), done_, 0));
  }

 private:
  size_t chunk_elems(size_t n) const { return (((n + world_ - 1) / world_) + 3) & ~size_t(3); }

  // reduce-scatter then all-gather around the ring rank -> rank+1; W-1 hops each
  void ring_allreduce(float* g, size_t n) {
    NcclApi& api = NcclApi::get();
    const int W = world_, next = (rank_ + 1) % W, prev = (rank_ + W - 1) % W;
    const size_t chunk = chunk_elems(n);
    auto range = [&](int c, size_t& off, size_t& len) {
      off = std::min(n, static_cast<size_t>(c) * chunk);
      len = std::min(n - off, chunk);
    };
    for (int s = 0; s < W - 1; ++s) {
      size_t so, sl, ro, rl;
      range(((rank_ - s) % W + W) % W, so, sl);
      range(((rank_ - s - 1) % W + W) % W, ro, rl);
      VNB_NCCL_OK(api.GroupStart());
      if (sl) VNB_NCCL_OK(api.Send(g + so, sl, ncclFloat, next, comm_, stream_));
      if (rl) VNB_NCCL_OK(api.Recv(scratch_, rl, ncclFloat, prev, comm_, stream_));
      VNB_NCCL_OK(api.GroupEnd());
      if (rl) {
        const int blocks = static_cast<int>(std::min<size_t>((rl + 1023) / 1024, 296));
        accumulate_kernel<<<blocks, 256, 0, stream_>>>(g + ro, scratch_, static_cast<long long>(rl));
      }
    }
    for (int s = 0; s < W - 1; ++s) {
      size_t so, sl, ro, rl;
      range(((rank_ + 1 - s) % W + W) % W, so, sl);
      range(((rank_ - s) % W + W) % W, ro, rl);
      VNB_NCCL_OK(api.GroupStart());
      if (sl) VNB_NCCL_OK(api.Send(g + so, sl, ncclFloat, next, comm_, stream_));
      if (rl) VNB_NCCL_OK(api.Recv(g + ro, rl, ncclFloat, prev, comm_, stream_));
      VNB_NCCL_OK(api.GroupEnd());
    }
  }

  // reduce-scatter and all-gather as two all-to-all exchanges (see the header comment)
  void direct_allreduce(float* g, size_t n) {
    NcclApi& api = NcclApi::get();
    const int W = world_;
    const size_t chunk = chunk_elems(n);
    auto range = [&](int c, size_t& off, size_t& len) {
      off = std::min(n, static_cast<size_t>(c) * chunk);
      len = std::min(n - off, chunk);
    };
    size_t mo, ml;
    range(rank_, mo, ml);
    VNB_NCCL_OK(api.GroupStart());
    for (int k = 1; k < W; ++k) {   // peer order rotated by rank: no two ranks start on the same destination
      const int peer = (rank_ + k) % W;
      size_t po, pl;
      range(peer, po, pl);
      if (pl) VNB_NCCL_OK(api.Send(g + po, pl, ncclFloat, peer, comm_, stream_));
      const int idx = peer < rank_ ? peer : peer - 1;   // scratch slots in rank order: fixed summation order
      if (ml) VNB_NCCL_OK(api.Recv(scratch_ + static_cast<size_t>(idx) * max_chunk_, ml, ncclFloat, peer, comm_, stream_));
    }
    VNB_NCCL_OK(api.GroupEnd());
    if (ml) {
      const int blocks = static_cast<int>(std::min<size_t>((ml + 1023) / 1024, 296));
      accumulate_multi_kernel<<<blocks, 256, 0, stream_>>>(g + mo, scratch_, static_cast<long long>(max_chunk_), W - 1,
                                                           static_cast<long long>(ml));
    }
    VNB_NCCL_OK(api.GroupStart());
    for (int k = 1; k < W; ++k) {
      const int peer = (rank_ + k) % W;
      size_t po, pl;
      range(peer, po, pl);
      if (ml) VNB_NCCL_OK(api.Send(g + mo, ml, ncclFloat, peer, comm_, stream_));
      if (pl) VNB_NCCL_OK(api.Recv(g + po, pl, ncclFloat, peer, comm_, stream_));
    }
    VNB_NCCL_OK(api.GroupEnd());
  }

  int rank_, world_;
  ncclComm_t comm_ = nullptr;
  ncclComm_t stats_comm_ = nullptr;
  cudaStream_t stream_ = nullptr;
  int schedule_ = 1;       // 0 ring, 1 direct, 2 ncclAllReduce
  bool defer_buckets_ = false;
  std::vector<int> deferred_;
  size_t max_chunk_ = 4;
  float* scratch_ = nullptr;
  std::vector<cudaEvent_t> events_;
  std::vector<cudaEvent_t> wg_events_;
  cudaEvent_t done_ = nullptr;
  int pending_ = 0;
};

}  // namespace vnb
