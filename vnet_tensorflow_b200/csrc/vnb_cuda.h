// Build-mode switch: real CUDA (product, sm_100a) or the test-only CPU emulation shim.
#pragma once

#ifdef VNB_EMULATE
#include "cuda_emul.h"  // tests/emul, added to the include path by the test build only
#define VNB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emul::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })
#define VNB_DYN_SMEM(type, name) type* name = static_cast<type*>(emul::dyn_smem_ptr())
#else
#include <cuda_runtime.h>
#define VNB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define VNB_DYN_SMEM(type, name) extern __shared__ __align__(1024) unsigned char name##_raw_[]; \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define VNB_HD __host__ __device__ __forceinline__

// Programmatic dependent launch for the chains of short HBM-bound passes (statistics -> finalize -> apply): a kernel
// launched with VNB_LAUNCH_PDL may become resident while its predecessor in the stream is still running; it calls
// pdl_wait() before it touches anything the predecessor wrote (the wait returns when that grid has completed and its
// writes are visible) and pdl_trigger() to let its own successor do the same.  Saves the launch latency at every
// boundary of such a chain; VNB_NO_PDL=1 launches them the ordinary way.
#ifdef VNB_EMULATE
#define VNB_LAUNCH_PDL VNB_LAUNCH
namespace vnb {
inline void pdl_wait() {}
inline void pdl_trigger() {}
}  // namespace vnb
#else
namespace vnb {
inline bool pdl_enabled() {
  static const bool on = getenv("VNB_NO_PDL") == nullptr;
  return on;
}
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}  // namespace vnb
#define VNB_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) \
  vnb::launch_pdl(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__)
#endif

namespace vnb {

// bf16 <-> fp32 without cuda_bf16.h (bit-exact round-to-nearest-even, NaN preserved)
VNB_HD uint16_t f32_to_bf16(float f) {
  uint32_t u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(f);
#else
  __builtin_memcpy(&u, &f, 4);
#endif
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return static_cast<uint16_t>((u >> 16) | 0x40u);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
VNB_HD float bf16_to_f32(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  __builtin_memcpy(&f, &u, 4);
  return f;
#endif
}

}  // namespace vnb
