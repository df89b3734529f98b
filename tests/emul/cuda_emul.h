// Test-only CUDA-on-CPU shim (NOT part of the product, never linked into libvnet_b200.so).
//
// It lets the CPU test-suite compile the product's kernel sources with g++ (-DVNB_EMULATE) and run
// them block by block, each CUDA thread as a fiber (ucontext), so that index arithmetic, barrier
// phases, reductions and the engine's orchestration can be checked against the oracle in the
// GPU-less container.  Supported: threadIdx/blockIdx/blockDim/gridDim, static and dynamic shared
// memory, __syncthreads, full-warp shuffles, atomicAdd, and the small slice of the runtime API the
// engine uses (malloc / memcpy / memset / streams as no-ops).  The sm_100a async units (mbarrier,
// TMA, tcgen05/TMEM) are modelled in emul_sm100.h following the semantics pinned on hardware by
// tools/probe_tcgen05.cu (profiles/r01_tcgen05_probe.log).
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_e {
  unsigned x, y, z;
};
struct float2 {
  float x, y;
};
struct float4 {
  float x, y, z, w;
};
struct uint2 {
  unsigned x, y;
};
struct uint4 {
  unsigned x, y, z, w;
};
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }

typedef int cudaStream_t;
typedef int cudaError_t;
typedef int cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

namespace emul {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = true;
  uint3_e tid{0, 0, 0};
  int lin = 0;
  uint64_t shfl_seq = 0;
};

struct Block {
  uint3_e bid{0, 0, 0};
  dim3 bdim, gdim;
  int nthreads = 0, live = 0;
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int bar_arrived = 0;
  uint64_t bar_gen = 0;
  uint64_t progress = 0;  // bumped on every state change; used for deadlock detection
  std::vector<uint64_t> slot[2];
  std::vector<uint64_t> warp_count[2];
  std::vector<uint8_t> dyn_smem;
  std::function<void()> body;
  struct NamedBar {
    int arrived = 0;
    uint64_t gen = 0;
  };
  NamedBar named[16];
};

inline Block& blk() {
  static Block b;
  return b;
}
inline Fiber*& cur() {
  static Fiber* f = nullptr;
  return f;
}
static const size_t kStack = 256 * 1024;

inline void yield_() {
  Fiber* f = cur();
  swapcontext(&f->ctx, &blk().sched);
}
inline void yield_wait() { yield_(); }

inline void fiber_entry() {
  Block& b = blk();
  b.body();
  Fiber* f = cur();
  f->done = true;
  b.live--;
  b.progress++;
  if (b.live > 0 && b.bar_arrived == b.live) {  // exited threads no longer count at barriers
    b.bar_arrived = 0;
    b.bar_gen++;
  }
  swapcontext(&f->ctx, &b.sched);
}

inline void run_block(std::function<void()> body) {
  Block& b = blk();
  b.body = body;
  b.live = b.nthreads;
  b.bar_arrived = 0;
  for (auto& nb : b.named) nb = Block::NamedBar();
  if ((int)b.fibers.size() < b.nthreads) b.fibers.resize(b.nthreads);
  int nwarps = (b.nthreads + 31) / 32;
  for (int k = 0; k < 2; ++k) {
    b.slot[k].assign(b.nthreads, 0);
    b.warp_count[k].assign(nwarps, 0);
  }
  for (int i = 0; i < b.nthreads; ++i) {
    Fiber& f = b.fibers[i];
    if (!f.stack) f.stack = static_cast<char*>(malloc(kStack));
    f.done = false;
    f.lin = i;
    f.shfl_seq = 0;
    f.tid.x = i % b.bdim.x;
    f.tid.y = (i / b.bdim.x) % b.bdim.y;
    f.tid.z = i / (b.bdim.x * b.bdim.y);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  int stalled_rounds = 0;
  while (b.live > 0) {
    uint64_t before = b.progress;
    for (int i = 0; i < b.nthreads; ++i) {
      Fiber& f = b.fibers[i];
      if (f.done) continue;
      cur() = &f;
      swapcontext(&b.sched, &f.ctx);
    }
    if (b.progress == before) {
      if (++stalled_rounds > 2000) {
        fprintf(stderr, "[cuda_emul] DEADLOCK: block (%u,%u,%u) made no progress (live=%d, at barrier=%d)\n",
                b.bid.x, b.bid.y, b.bid.z, b.live, b.bar_arrived);
        abort();
      }
    } else {
      stalled_rounds = 0;
    }
  }
  cur() = nullptr;
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F body) {
  Block& b = blk();
  b.bdim = block;
  b.gdim = grid;
  b.nthreads = block.x * block.y * block.z;
  b.dyn_smem.assign(smem + 2048, 0);
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        b.bid = {x, y, z};
        run_block(body);
      }
}

inline void* dyn_smem_ptr() {  // 1024-byte aligned like the real kernels request
  uintptr_t p = reinterpret_cast<uintptr_t>(blk().dyn_smem.data());
  return reinterpret_cast<void*>((p + 1023) & ~uintptr_t(1023));
}

inline void syncthreads() {
  Block& b = blk();
  uint64_t gen = b.bar_gen;
  b.bar_arrived++;
  b.progress++;
  if (b.bar_arrived == b.live) {
    b.bar_arrived = 0;
    b.bar_gen++;
    return;
  }
  while (b.bar_gen == gen) yield_wait();
}

// bar.sync id, count: barrier among `count` threads of the block
inline void named_barrier(int id, int count) {
  Block& b = blk();
  Block::NamedBar& nb = b.named[id & 15];
  uint64_t gen = nb.gen;
  nb.arrived++;
  b.progress++;
  if (nb.arrived == count) {
    nb.arrived = 0;
    nb.gen++;
    return;
  }
  while (nb.gen == gen) yield_wait();
}

template <class T>
inline T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  Block& b = blk();
  Fiber* f = cur();
  int w = f->lin / 32;
  int lanes = std::min(32, b.nthreads - w * 32);
  int buf = f->shfl_seq & 1;
  uint64_t target = (f->shfl_seq / 2 + 1) * (uint64_t)lanes;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b.slot[buf][f->lin] = raw;
  b.warp_count[buf][w]++;
  b.progress++;
  f->shfl_seq++;
  while (b.warp_count[buf][w] < target) yield_wait();
  int src = w * 32 + (src_lane & 31);
  if (src_lane < 0 || src_lane >= lanes) src = f->lin;  // out-of-range source: own value (CUDA semantics)
  T r;
  memcpy(&r, &b.slot[buf][src], sizeof(T));
  return r;
}

}  // namespace emul

#define threadIdx (emul::cur()->tid)
#define blockIdx (emul::blk().bid)
#define blockDim (emul::blk().bdim)
#define gridDim (emul::blk().gdim)
#define warpSize 32

inline void __syncthreads() { emul::syncthreads(); }
inline void __syncwarp(unsigned = 0xffffffffu);  // defined below: rendezvous of the warp's fibers
inline void __threadfence() {}
template <class T>
inline T __shfl_sync(unsigned, T v, int lane) {
  return emul::shfl_from(v, lane);
}
inline void __syncwarp(unsigned) { (void)emul::shfl_from(0, 0); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
  return emul::shfl_from(v, (emul::cur()->lin & 31) ^ m);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, int d) {
  return emul::shfl_from(v, (emul::cur()->lin & 31) + d);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int d) {
  return emul::shfl_from(v, (emul::cur()->lin & 31) - d);
}
template <class T>
inline T atomicAdd(T* p, T v) {
  T o = *p;
  *p = o + v;
  return o;
}
template <class T>
inline T __ldg(const T* p) {
  return *p;
}
template <class T>
inline T __ldcg(const T* p) {
  return *p;
}
inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline float __expf(float x) { return expf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }

// ---- runtime API subset ---------------------------------------------------------------------
inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  return *p ? cudaSuccess : 2;
}
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) {
  return cudaMalloc(reinterpret_cast<void**>(p), n);
}
inline cudaError_t cudaFree(void* p) {
  free(p);
  return cudaSuccess;
}
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  memcpy(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) {
  memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) {
  memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) {
  memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = 0;
  return cudaSuccess;
}
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = 0;
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
