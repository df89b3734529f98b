// sm_100a PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld),
// and the shared-memory matrix descriptor + instruction descriptor encoders used by the
// implicit-GEMM convolution kernels. Inline PTX only (no CUTLASS dependency at build time).
//
// Bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables.
#pragma once
#include <cstdint>
#ifdef VNB_EMULATE
#include "emul_sm100.h"  // tests/emul: CPU model of the async units, same function names
#else
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>

namespace sm100 {

typedef CUtensorMap TmaDesc;

// ----------------------------------------------------------------------------------------------
// address helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy smem writes visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Waits are bounded (~4 s of SM clocks): a pipeline bug then traps with a message instead of hanging
// the GPU until the driver watchdog fires.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {
      printf("[vnet_b200] mbarrier wait timed out: block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// bounded wait used by probes/tests: returns false on timeout instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity, uint32_t spins) {
  for (uint32_t i = 0; i < spins; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

// ----------------------------------------------------------------------------------------------
// TMA tiled loads (global -> shared), completion on an mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA tiled stores (shared -> global), bulk-group completion: plain store and element-wise fp32 add (the tensor map's
// data type selects the arithmetic); out-of-range parts of the box are clipped
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tmap)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tmap)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups still READ their shared-memory source (the buffers of the others may be reused)
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// all of this thread's bulk groups have completed (their global writes are performed)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// One lane of a converged warp.  The producer / MMA warps run their loops with all 32 lanes (warp-uniform
// control flow keeps descriptors, barrier addresses and loop state in uniform registers, so the asynchronous
// instructions issue without per-instruction R2UR waterfalls) and only the elected lane issues them.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// tell the compiler a value is warp-uniform (same idiom as a canonical warp index)
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// Barrier wait of a whole role warp: every lane polls the barrier (all lanes observe the same phase in the same
// instruction) and the warp reconverges before it goes on, so the loop state stays warp-uniform and the
// asynchronous instructions that follow are issued once, from converged code.  (Two things that were tried and do
// not work: a leader-only wait followed by __syncwarp -- the warp can stay split and the uniform-datapath instructions
// then execute once per fragment; and a look-ahead mbarrier.test_wait of the next stage -- the predicate is consumed
// by the selp inside the same asm block, so the ~150-cycle query latency lands on the issue path anyway.)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  mbar_wait(bar, parity);
  __syncwarp();
}

// named barrier among `count` threads (count a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, fences, MMA, commit, TMEM loads
// ----------------------------------------------------------------------------------------------
// whole-warp (.sync.aligned) -- writes the TMEM base address to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16/bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32-bit, 16 consecutive columns -> 16 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
}


// ----------------------------------------------------------------------------------------------
// issue-by-one-lane forms.  The role warps run their loops converged; the lane with pred != 0 issues, through a plain
// branch around the instruction: ptxas then keeps the warp-uniform operands in uniform registers.  (Predicating the
// instruction inside the asm block instead makes ptxas move every operand with `@p R2UR.BROADCAST`, a cross-lane
// operation per register; issuing from inside `if (lane == 0)` loops wraps every instruction in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall.)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_f16_ss_if(bool pred, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  if (pred) mma_f16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void mma_commit_if(bool pred, uint32_t bar) {
  if (pred) mma_commit(bar);
}
__device__ __forceinline__ void mbar_expect_tx_if(bool pred, uint32_t bar, uint32_t bytes) {
  if (pred) mbar_expect_tx(bar, bytes);
}
__device__ __forceinline__ void tma_load_2d_if(bool pred, uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  if (pred) tma_load_2d(dst, tmap, bar, c0, c1);
}
__device__ __forceinline__ void tma_load_5d_if(bool pred, uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                               int c3, int c4) {
  if (pred) tma_load_5d(dst, tmap, bar, c0, c1, c2, c3, c4);
}

}  // namespace sm100
#endif  // VNB_EMULATE

namespace sm100 {
// ----------------------------------------------------------------------------------------------
// descriptors
// ----------------------------------------------------------------------------------------------
enum : uint32_t { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };

// shared-memory matrix descriptor (64-bit):
//  [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//  [32,46) stride-dim byte offset>>4 [46,48) version = 1 (sm_100)
//  [49,52) base offset               [61,64) layout type (swizzle)
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t start_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes, uint32_t layout,
                                                            uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((start_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 0x7) << 49;
  d |= static_cast<uint64_t>(layout & 0x7) << 61;
  return d;
}

// instruction descriptor (32-bit) for kind::f16 / kind::tf32, dense, fp32 accumulate:
//  [4,6) c_format (1 = F32)  [7,10) a_format  [10,13) b_format  (0 F16, 1 BF16, 2 TF32)
//  [15] a_major (0 K, 1 MN)  [16] b_major  [17,23) N>>3  [24,29) M>>4
enum : uint32_t { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };
__host__ __device__ __forceinline__ uint32_t make_instr_desc(uint32_t M, uint32_t N, uint32_t fmt,
                                                             uint32_t a_mn_major = 0,
                                                             uint32_t b_mn_major = 0) {
  uint32_t d = 0;
  d |= 1u << 4;  // fp32 accumulator
  d |= (fmt & 7u) << 7;
  d |= (fmt & 7u) << 10;
  d |= (a_mn_major & 1u) << 15;
  d |= (b_mn_major & 1u) << 16;
  d |= ((N >> 3) & 0x3Fu) << 17;
  d |= ((M >> 4) & 0x1Fu) << 24;
  return d;
}

}  // namespace sm100
