// Test-only CPU model of the sm_100a async units used by the implicit-GEMM kernels: mbarrier
// (arrive / expect_tx / parity wait), TMA tiled loads (out-of-range zero fill, 32/64/128-byte swizzle
// on the absolute shared address) and tcgen05 (TMEM accumulators, shared-memory matrix descriptors in
// K-major and MN-major form, .ld).  The semantics are exactly the ones pinned on a real B200 by
// tools/probe_tcgen05.cu (profiles/r01_tcgen05_probe.log): swizzle XOR on absolute address bits,
// base_offset = 0, arbitrary 16-byte start addresses, LBO/SBO as plain byte strides.
// Everything executes synchronously at issue, so this model checks indexing, descriptors, pipeline
// phase logic (deadlocks abort) and the epilogue -- not memory-ordering races.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "cuda_emul.h"

namespace sm100 {

struct TmaDesc {  // what cuTensorMapEncodeTiled would hold
  const uint8_t* base = nullptr;
  int rank = 0;
  uint64_t dims[5] = {1, 1, 1, 1, 1};
  uint64_t strides[5] = {0, 0, 0, 0, 0};  // bytes; strides[0] = element size
  uint32_t box[5] = {1, 1, 1, 1, 1};
  uint32_t elem = 2;
  uint32_t swizzle = 0;  // 0 / 32 / 64 / 128
};

struct EmulState {
  struct Bar {
    uint32_t init = 0, pending = 0, phase = 0;
    int64_t tx = 0;
  };
  std::map<uint32_t, Bar> bars;
  std::vector<float> tmem;  // [128 lanes][512 cols]
  uint8_t* smem_base = nullptr;
  EmulState() : tmem(128 * 512, 0.f) {}
};
inline EmulState& st() {
  static EmulState s;
  return s;
}
// shared "address" = byte offset in the block's dynamic shared memory + 1024 (keeps 1024 alignment)
inline uint32_t smem_u32(const void* p) {
  uint8_t* base = static_cast<uint8_t*>(emul::dyn_smem_ptr());
  return static_cast<uint32_t>(static_cast<const uint8_t*>(p) - base) + 1024u;
}
inline uint8_t* smem_ptr(uint32_t a) { return static_cast<uint8_t*>(emul::dyn_smem_ptr()) + (a - 1024u); }

inline void bar_check(EmulState::Bar& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1u;
    b.pending = b.init;
    emul::blk().progress++;
  }
}
inline void mbar_init(uint32_t bar, uint32_t count) {
  EmulState::Bar b;
  b.init = b.pending = count;
  st().bars[bar] = b;
}
inline void fence_mbar_init() {}
inline void fence_proxy_async_smem() {}
inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  auto& b = st().bars.at(bar);
  b.tx += bytes;
  b.pending--;
  bar_check(b);
}
inline void mbar_arrive(uint32_t bar) {
  auto& b = st().bars.at(bar);
  b.pending--;
  bar_check(b);
}
inline bool elect_one() { return (emul::cur()->lin & 31) == 0; }
inline uint32_t warp_uniform(uint32_t v) { return v; }
inline bool mbar_try_wait(uint32_t bar, uint32_t parity) { return st().bars.at(bar).phase != (parity & 1u); }
inline void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) emul::yield_wait();
}
inline void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  mbar_wait(bar, parity);
  __syncwarp();
}
inline bool mbar_wait_bounded(uint32_t bar, uint32_t parity, uint32_t) {
  mbar_wait(bar, parity);
  return true;
}

inline uint32_t swz_addr(uint32_t a, uint32_t mode) {
  switch (mode) {
    case 128: return a ^ (((a >> 7) & 7u) << 4);
    case 64: return a ^ (((a >> 7) & 3u) << 4);
    case 32: return a ^ (((a >> 7) & 1u) << 4);
    default: return a;
  }
}

inline void tma_prefetch_desc(const void*) {}
inline void tma_load_nd(uint32_t dst, const TmaDesc* t, uint32_t bar, const int* c) {
  const uint32_t e = t->elem;
  uint64_t total = 1;
  for (int i = 0; i < t->rank; ++i) total *= t->box[i];
  uint32_t idx[5] = {0, 0, 0, 0, 0};
  for (uint64_t lin = 0; lin < total; ++lin) {
    uint64_t r = lin;
    bool inb = true;
    uint64_t goff = 0;
    for (int i = 0; i < t->rank; ++i) {
      idx[i] = static_cast<uint32_t>(r % t->box[i]);
      r /= t->box[i];
      const long long g = static_cast<long long>(c[i]) + idx[i];
      if (g < 0 || g >= static_cast<long long>(t->dims[i])) inb = false;
      goff += static_cast<uint64_t>(g < 0 ? 0 : g) * t->strides[i];
    }
    const uint32_t a = swz_addr(dst + static_cast<uint32_t>(lin) * e, t->swizzle);
    if (inb)
      memcpy(smem_ptr(a), t->base + goff, e);
    else
      memset(smem_ptr(a), 0, e);
  }
  auto& b = st().bars.at(bar);
  b.tx -= static_cast<int64_t>(total) * e;
  bar_check(b);
}
inline void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  int c[5] = {c0, c1, 0, 0, 0};
  tma_load_nd(dst, static_cast<const TmaDesc*>(tmap), bar, c);
}
inline void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  int c[5] = {c0, c1, c2, 0, 0};
  tma_load_nd(dst, static_cast<const TmaDesc*>(tmap), bar, c);
}
inline void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  int c[5] = {c0, c1, c2, c3, 0};
  tma_load_nd(dst, static_cast<const TmaDesc*>(tmap), bar, c);
}
// TMA stores: the box is written (or added, fp32) element by element, out-of-range parts clipped; synchronous
inline void tma_store_nd(const TmaDesc* t, uint32_t src, const int* c, bool add) {
  const uint32_t e = t->elem;
  uint64_t total = 1;
  for (int i = 0; i < t->rank; ++i) total *= t->box[i];
  for (uint64_t lin = 0; lin < total; ++lin) {
    uint64_t r = lin;
    bool inb = true;
    uint64_t goff = 0;
    for (int i = 0; i < t->rank; ++i) {
      const uint32_t idx = static_cast<uint32_t>(r % t->box[i]);
      r /= t->box[i];
      const long long g = static_cast<long long>(c[i]) + idx;
      if (g < 0 || g >= static_cast<long long>(t->dims[i])) inb = false;
      goff += static_cast<uint64_t>(g < 0 ? 0 : g) * t->strides[i];
    }
    if (!inb) continue;
    const uint32_t a = swz_addr(src + static_cast<uint32_t>(lin) * e, t->swizzle);
    uint8_t* dst = const_cast<uint8_t*>(t->base) + goff;
    if (add) {
      if (e != 4) {
        fprintf(stderr, "[emul_sm100] TMA reduce-add is modelled for fp32 only\n");
        abort();
      }
      float x, y;
      memcpy(&x, dst, 4);
      memcpy(&y, smem_ptr(a), 4);
      x += y;
      memcpy(dst, &x, 4);
    } else {
      memcpy(dst, smem_ptr(a), e);
    }
  }
}
inline void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  int c[5] = {c0, c1, c2, c3, 0};
  tma_store_nd(static_cast<const TmaDesc*>(tmap), src, c, false);
}
inline void tma_reduce_add_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  int c[5] = {c0, c1, c2, c3, 0};
  tma_store_nd(static_cast<const TmaDesc*>(tmap), src, c, true);
}
inline void tma_store_commit() {}
template <int N>
inline void tma_store_wait_read() {}
inline void tma_store_wait_all() {}
inline void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  int c[5] = {c0, c1, c2, c3, c4};
  tma_load_nd(dst, static_cast<const TmaDesc*>(tmap), bar, c);
}

inline void tmem_alloc(uint32_t smem_slot, uint32_t) {
  if ((emul::cur()->lin & 31) == 0) {
    uint32_t zero = 0;
    memcpy(smem_ptr(smem_slot), &zero, 4);
    // poison: accumulators must be initialised by an accumulate=0 MMA
    for (auto& x : st().tmem) x = __builtin_nanf("");
  }
}
inline void tmem_relinquish() {}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before_sync() {}
inline void tc_fence_after_sync() {}

inline float bf16_bits_to_f32(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// element (mn, k) of an operand described by a shared-memory matrix descriptor
inline float operand_elem(uint64_t desc, bool mn_major, uint32_t mn, uint32_t k) {
  const uint32_t start = static_cast<uint32_t>(desc & 0x3FFF) << 4;
  const uint32_t lbo = static_cast<uint32_t>((desc >> 16) & 0x3FFF) << 4;
  const uint32_t sbo = static_cast<uint32_t>((desc >> 32) & 0x3FFF) << 4;
  const uint32_t layout = static_cast<uint32_t>((desc >> 61) & 7);
  const uint32_t sw = layout == 2 ? 128 : layout == 4 ? 64 : layout == 6 ? 32 : 0;
  uint32_t off;
  if (!mn_major) {
    if (sw)  // rows of `sw` bytes, 8-row groups at SBO, K contiguous inside the row
      off = (mn / 8) * sbo + (mn % 8) * sw + k * 2;
    else     // core matrices: 8 rows x 16 B; K chunks at LBO, 8-row groups at SBO
      off = (mn / 8) * sbo + (mn % 8) * 16 + (k / 8) * lbo + (k % 8) * 2;
  } else {
    const uint32_t atom = sw ? sw / 2 : 8;  // MN elements per atom
    if (sw)  // atom = `sw` bytes of MN x 8 K-rows; MN atoms at LBO, 8-row K groups at SBO
      off = (mn / atom) * lbo + (mn % atom) * 2 + (k / 8) * sbo + (k % 8) * sw;
    else
      off = (mn / 8) * sbo + (mn % 8) * 2 + (k / 8) * lbo + (k % 8) * 16;
  }
  const uint32_t a = swz_addr(start + off, sw);
  uint16_t h;
  memcpy(&h, smem_ptr(a), 2);
  return bf16_bits_to_f32(h);
}
inline void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t N = ((idesc >> 17) & 0x3F) << 3, M = ((idesc >> 24) & 0x1F) << 4;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  const uint32_t col0 = d_tmem & 0xFFFF;
  if (M != 128 || col0 + N > 512) {
    fprintf(stderr, "[emul_sm100] unsupported MMA shape M=%u N=%u col=%u\n", M, N, col0);
    abort();
  }
  std::vector<float> A(128 * 16), B(N * 16);
  for (uint32_t m = 0; m < 128; ++m)
    for (uint32_t k = 0; k < 16; ++k) A[m * 16 + k] = operand_elem(adesc, a_mn, m, k);
  for (uint32_t n = 0; n < N; ++n)
    for (uint32_t k = 0; k < 16; ++k) B[n * 16 + k] = operand_elem(bdesc, b_mn, n, k);
  for (uint32_t m = 0; m < 128; ++m)
    for (uint32_t n = 0; n < N; ++n) {
      float s = 0.f;
      for (uint32_t k = 0; k < 16; ++k) s += A[m * 16 + k] * B[n * 16 + k];
      float& d = st().tmem[m * 512 + col0 + n];
      d = accumulate ? d + s : s;
    }
}
inline void mma_commit(uint32_t bar) { mbar_arrive(bar); }
// predicated forms (see sm100_ptx.cuh)
inline void mma_f16_ss_if(bool pred, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (pred) mma_f16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
}
inline void mma_commit_if(bool pred, uint32_t bar) {
  if (pred) mma_commit(bar);
}
inline void mbar_expect_tx_if(bool pred, uint32_t bar, uint32_t bytes) {
  if (pred) mbar_expect_tx(bar, bytes);
}
inline void tma_load_2d_if(bool pred, uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  if (pred) tma_load_2d(dst, tmap, bar, c0, c1);
}
inline void tma_load_5d_if(bool pred, uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  if (pred) tma_load_5d(dst, tmap, bar, c0, c1, c2, c3, c4);
}
inline void tmem_ld_wait() {}
inline void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  const uint32_t lane = (taddr >> 16) + (emul::cur()->lin & 31), col = taddr & 0xFFFF;
  for (int j = 0; j < 16; ++j) memcpy(&r[j], &st().tmem[lane * 512 + col + j], 4);
}
inline void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  const uint32_t lane = (taddr >> 16) + (emul::cur()->lin & 31), col = taddr & 0xFFFF;
  for (int j = 0; j < 8; ++j) memcpy(&r[j], &st().tmem[lane * 512 + col + j], 4);
}
inline void named_bar_sync(int id, int count) { emul::named_barrier(id, count); }

}  // namespace sm100
