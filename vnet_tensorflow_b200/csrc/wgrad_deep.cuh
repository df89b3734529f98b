// Filter gradient of the 5^3 convolutions of the DEEP levels (>= 128 channels on 16^3 / 8^3 grids: encoder / decoder
// level 4 and the bottom level of networks.VNet; TF autodiff of layers2.py:59-63 under model.py:660).
//
// wgrad5_tc_kernel (wgrad_tc.cuh) folds the kw taps onto GEMM-M with 16-channel atoms: right for the wide, shallow
// levels, but on 8..16-voxel lines it is left with N = 80 MMAs or short bursts, 256 (ci, co) pairs and split-K partials
// of 8-32 MB: 103 TFLOP/s on 256 -> 256 @8^3, 215 on 128 -> 128 @16^3 (bf16x3).  With >= 128 channels each TAP is a full
// tcgen05 GEMM by itself:
//     dw[tap][ci][co] = sum_v X[v + tap - 2][ci] * dZ[v][co]         M = 128 ci, N = 128 or 256 co, K = voxels
// Both operands are MN-major exactly as the NDHWC bf16 (hi, lo) copies sit in HBM: a TMA box (64 channels, W, HT lines)
// lands as [64 voxel rows][128 B] = one SWIZZLE_128B MN-major atom column; the tap shift is the box origin, SAME padding
// is TMA's out-of-range zero fill, and input planes that lie wholly outside the volume are skipped.  A work item is
// (tap, 128-channel ci block, co block, K split); its accumulator lives in TMEM (two buffers: the epilogue of item i
// runs under the MMAs of item i + 1) and is stored straight into dw[tap] (or into a split-K partial that a fixed-order
// kernel sums: deterministic).
//
// 192 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.
#pragma once
#include "conv_tc.cuh"

namespace vnb {

constexpr int kWdThreads = 192;
constexpr int kWdAtom = 64 * 128;   // 64 voxel rows x 64 channels bf16 (8 KB)
constexpr int kWdMaxStages = 4;

struct WdGeom {
  int N, D, H, W;
  int HT, n_hb;              // lines per K tile (W * HT = 64 voxels), line blocks
  int C1, C2, Cout;          // input channels of the two concatenated sources, output channels
  int n_cib, n_cob, NB;      // 128-channel ci blocks, co blocks of NB (128 or 256) channels
  int ksplit, npl;           // K splits per (tap, block); operand planes (2 = hi + lo, three MMA passes)
  int stages, stage_bytes, n_items;
};

__global__ void __launch_bounds__(kWdThreads, 1)
wgrad5_deep_kernel(const __grid_constant__ sm100::TmaDesc x1_hi, const __grid_constant__ sm100::TmaDesc x1_lo,
                   const __grid_constant__ sm100::TmaDesc x2_hi, const __grid_constant__ sm100::TmaDesc x2_lo,
                   const __grid_constant__ sm100::TmaDesc z_hi, const __grid_constant__ sm100::TmaDesc z_lo, const WdGeom g,
                   float* __restrict__ out) {
  using namespace sm100;
  VNB_DYN_SMEM(uint8_t, smem_raw);
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t sm_addr = smem_u32(sm);
  const uint32_t bar_base = sm_addr + static_cast<uint32_t>(g.stages) * g.stage_bytes;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (4 + s); };
  auto accf = [&](int b) { return bar_base + 8u * (8 + b); };
  auto acce = [&](int b) { return bar_base + 8u * (10 + b); };
  const uint32_t slot_addr = bar_base + 8u * 12;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + static_cast<size_t>(g.stages) * g.stage_bytes + 8 * 12);

  const int tid = threadIdx.x;
  const int warp = static_cast<int>(warp_uniform(static_cast<uint32_t>(tid >> 5)));
  if (tid == 0) {
    for (int s = 0; s < kWdMaxStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf(b), 1);
      mbar_init(acce(b), 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(slot_addr, static_cast<uint32_t>(2 * g.NB));
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = warp_uniform(*slot_ptr);
  const int n_ba = g.NB / 64;                                            // B atoms per plane
  const uint32_t a_plane = 2u * kWdAtom, b_plane = static_cast<uint32_t>(n_ba) * kWdAtom;
  const uint32_t b_off = static_cast<uint32_t>(g.npl) * a_plane;         // stage: A planes | B planes

  // item -> (tap, ci block, co block, split); the K tiles of a tap: samples x valid input planes x line blocks
  struct Item {
    int kd, kh, kw, cib, cob, split, d_lo, n_d, t_lo, t_hi;
  };
  auto decode = [&](int item) {
    Item it;
    int x = item;
    const int tap = x % 125;
    x /= 125;
    it.split = x % g.ksplit;
    x /= g.ksplit;
    it.cob = x % g.n_cob;
    it.cib = x / g.n_cob;
    it.kd = tap / 25;
    it.kh = (tap / 5) % 5;
    it.kw = tap % 5;
    // output planes d whose input plane d + kd - 2 exists
    it.d_lo = it.kd < 2 ? 2 - it.kd : 0;
    const int d_hi = it.kd > 2 ? g.D - (it.kd - 2) : g.D;
    it.n_d = d_hi > it.d_lo ? d_hi - it.d_lo : 0;
    const int n_kt = g.N * it.n_d * g.n_hb;
    it.t_lo = static_cast<int>(static_cast<long long>(it.split) * n_kt / g.ksplit);
    it.t_hi = static_cast<int>(static_cast<long long>(it.split + 1) * n_kt / g.ksplit);
    return it;
  };

  if (warp == 0) {
    // ======================= TMA producer =======================
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      const Item it = decode(item);
      const int ci0 = it.cib * 128;
      const bool src1 = ci0 < g.C1;
      const int cx = src1 ? ci0 : ci0 - g.C1;
      const TmaDesc* xh = src1 ? &x1_hi : &x2_hi;
      const TmaDesc* xl = src1 ? &x1_lo : &x2_lo;
      for (int t = it.t_lo; t < it.t_hi; ++t) {
        int x = t;
        const int hb = x % g.n_hb;
        x /= g.n_hb;
        const int d = it.d_lo + x % it.n_d, n = x / it.n_d;
        mbar_wait_warp(empty(s), ph ^ 1u);
        if (leader) {
          const uint32_t st = sm_addr + static_cast<uint32_t>(s) * g.stage_bytes;
          mbar_expect_tx(full(s), static_cast<uint32_t>(g.npl) * (a_plane + b_plane));
          const int h0 = hb * g.HT;
          for (int a = 0; a < 2; ++a) {
            tma_load_5d(st + a * kWdAtom, xh, full(s), cx + 64 * a, it.kw - 2, h0 + it.kh - 2, d + it.kd - 2, n);
            if (g.npl == 2) tma_load_5d(st + a_plane + a * kWdAtom, xl, full(s), cx + 64 * a, it.kw - 2, h0 + it.kh - 2, d + it.kd - 2, n);
          }
          for (int b = 0; b < n_ba; ++b) {
            const int co = it.cob * g.NB + 64 * b;
            tma_load_5d(st + b_off + b * kWdAtom, &z_hi, full(s), co, 0, h0, d, n);
            if (g.npl == 2) tma_load_5d(st + b_off + b_plane + b * kWdAtom, &z_lo, full(s), co, 0, h0, d, n);
          }
        }
        __syncwarp();
        if (++s == g.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const bool leader = elect_one();
    const uint32_t idesc = make_instr_desc(128, static_cast<uint32_t>(g.NB), FMT_BF16, 1, 1);
    const uint64_t desc0 = make_smem_desc(0, kWdAtom, 1024, SWZ_128B);   // MN-major: 64-channel atoms at LBO, 8 K rows per SBO
    int s = 0;
    uint32_t ph = 0, n_it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++n_it) {
      const Item it = decode(item);
      const uint32_t buf = n_it & 1u, use = n_it >> 1;
      mbar_wait_warp(acce(buf), (use & 1u) ^ 1u);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem + buf * static_cast<uint32_t>(g.NB);
      for (int t = it.t_lo; t < it.t_hi; ++t) {
        mbar_wait_warp(full(s), ph);
        tc_fence_after_sync();
        const uint32_t st = sm_addr + static_cast<uint32_t>(s) * g.stage_bytes;
        const uint64_t dah = desc0 + (st >> 4), dal = dah + (a_plane >> 4);
        const uint64_t dbh = desc0 + ((st + b_off) >> 4), dbl = dbh + (b_plane >> 4);
        const uint32_t first = t != it.t_lo ? 1u : 0u;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {   // 64 voxel rows = four K = 16 steps of 2 KB
          const uint64_t o = static_cast<uint64_t>(ks * 128);
          if (g.npl == 2) {
            mma_f16_ss_if(leader, d_tmem, dal + o, dbh + o, idesc, (first | static_cast<uint32_t>(ks)) != 0 ? 1u : 0u);
            mma_f16_ss_if(leader, d_tmem, dah + o, dbl + o, idesc, 1u);
            mma_f16_ss_if(leader, d_tmem, dah + o, dbh + o, idesc, 1u);
          } else {
            mma_f16_ss_if(leader, d_tmem, dah + o, dbh + o, idesc, (first | static_cast<uint32_t>(ks)) != 0 ? 1u : 0u);
          }
        }
        mma_commit_if(leader, empty(s));
        __syncwarp();
        if (++s == g.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      mma_commit_if(leader, accf(buf));
      __syncwarp();
    }
  } else {
    // ======================= epilogue: accumulator row ci -> out[split][tap][ci][co block] =======================
    const int lane = tid & 31, q = warp & 3;
    const int Cin = g.C1 + g.C2;
    uint32_t n_it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++n_it) {
      const Item it = decode(item);
      const uint32_t buf = n_it & 1u, use = n_it >> 1;
      mbar_wait(accf(buf), use & 1u);
      tc_fence_after_sync();
      const int tap = (it.kd * 5 + it.kh) * 5 + it.kw;
      const int ci = it.cib * 128 + q * 32 + lane;
      float* dst = out + ((static_cast<size_t>(it.split) * 125 + tap) * Cin + ci) * g.Cout + it.cob * g.NB;
      const uint32_t t_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * static_cast<uint32_t>(g.NB);
      const bool empty_item = it.t_hi <= it.t_lo;   // no K tile: the accumulator was never written
      for (int j = 0; j < g.NB / 16; ++j) {
        uint32_t v[16];
        tmem_ld16(t_addr + static_cast<uint32_t>(j) * 16u, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<float4*>(dst + j * 16)[i] =
              empty_item ? make_float4(0.f, 0.f, 0.f, 0.f)
                         : make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                       __uint_as_float(v[4 * i + 3]));
      }
      tc_fence_before_sync();
      mbar_arrive(acce(buf));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, static_cast<uint32_t>(2 * g.NB));
}

// dw[i] = sum over the K splits, fixed order
__global__ void __launch_bounds__(256) wgrad5_deep_reduce_kernel(const float4* __restrict__ partial, int splits, size_t n4,
                                                                 float4* __restrict__ dw) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 s = partial[i];
    for (int k = 1; k < splits; ++k) {
      const float4 p = partial[static_cast<size_t>(k) * n4 + i];
      s.x += p.x;
      s.y += p.y;
      s.z += p.z;
      s.w += p.w;
    }
    dw[i] = s;
  }
}

struct WdPlan {
  bool valid = false;
  WdGeom g{};
  sm100::TmaDesc x1_hi, x1_lo, x2_hi, x2_lo, z_hi, z_lo;
  size_t smem = 0;
  size_t partial_floats = 0;   // 0: the kernel stores straight into dw
};

inline bool wd_enabled() { return getenv("VNB_WG_NO_DEEP") == nullptr; }

inline bool wd_plan_geometry(WdPlan& pl, int N, int D, int H, int W, int C1, int C2, int Cout, bool split3, int sms) {
  if (!wd_enabled()) return false;
  if (C1 <= 0 || C1 % 128 != 0 || C2 % 128 != 0 || Cout % 128 != 0) return false;
  if (W > 64 || 64 % W != 0 || W < 4) return false;
  const int HT = 64 / W;
  if (H % HT != 0) return false;
  // the shallow levels stay with wgrad5_tc_kernel: this mapping re-reads X and dZ once per tap
  if (static_cast<long long>(D) * H * W > 16 * 16 * 16) return false;
  WdGeom& g = pl.g;
  g = WdGeom{};
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.HT = HT;
  g.n_hb = H / HT;
  g.C1 = C1; g.C2 = C2; g.Cout = Cout;
  g.n_cib = (C1 + C2) / 128;
  g.NB = Cout % 256 == 0 ? 256 : 128;
  g.n_cob = Cout / g.NB;
  g.npl = split3 ? 2 : 1;
  g.stage_bytes = g.npl * (2 + g.NB / 64) * kWdAtom;
  g.stages = std::min(kWdMaxStages, (227 * 1024 - 2048) / g.stage_bytes);
  if (g.stages < 2) return false;
  pl.smem = static_cast<size_t>(g.stages) * g.stage_bytes + 2048;
  // K splits only when the items would leave most SMs idle (measured on 128 -> 128 @16^3, 125 items on 148 SMs: one
  // split 77 us, two 85 us, four 89 us -- the partials and their reduce cost more than the idle SMs)
  const int base_items = 125 * g.n_cib * g.n_cob;
  const int n_kt_min = N * (D - 2) * g.n_hb;
  g.ksplit = 1;
  while (base_items * g.ksplit * 2 <= sms && g.ksplit < 4 && n_kt_min / (g.ksplit * 2) >= 8) g.ksplit *= 2;
  if (const char* e = getenv("VNB_WD_KSPLIT")) g.ksplit = std::max(1, std::min({atoi(e), 8, n_kt_min}));   // tests
  g.n_items = base_items * g.ksplit;
  pl.partial_floats = g.ksplit > 1 ? static_cast<size_t>(g.ksplit) * 125 * (C1 + C2) * Cout : 0;
  return true;
}

// NDHWC bf16 tensor -> (C, W, H, D, N) map with box (64, W, HT, 1, 1), SWIZZLE_128B
inline void wd_encode_act(sm100::TmaDesc* out, const uint16_t* base, int N, int D, int H, int W, int C, int HT) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
  const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
  const uint32_t box[5] = {64, (uint32_t)W, (uint32_t)HT, 1, 1};
  tma_encode(out, base, 5, dims, str, box, 128);
}

inline void wd_encode_plan(WdPlan& pl, int N, const uint16_t* x1_hi, const uint16_t* x1_lo, const uint16_t* x2_hi, const uint16_t* x2_lo,
                           const uint16_t* z_hi, const uint16_t* z_lo) {
  const WdGeom& g = pl.g;
  wd_encode_act(&pl.x1_hi, x1_hi, N, g.D, g.H, g.W, g.C1, g.HT);
  wd_encode_act(&pl.x1_lo, x1_lo ? x1_lo : x1_hi, N, g.D, g.H, g.W, g.C1, g.HT);
  if (g.C2 > 0) {
    wd_encode_act(&pl.x2_hi, x2_hi, N, g.D, g.H, g.W, g.C2, g.HT);
    wd_encode_act(&pl.x2_lo, x2_lo ? x2_lo : x2_hi, N, g.D, g.H, g.W, g.C2, g.HT);
  } else {
    pl.x2_hi = pl.x1_hi;
    pl.x2_lo = pl.x1_lo;
  }
  wd_encode_act(&pl.z_hi, z_hi, N, g.D, g.H, g.W, g.Cout, g.HT);
  wd_encode_act(&pl.z_lo, z_lo ? z_lo : z_hi, N, g.D, g.H, g.W, g.Cout, g.HT);
}

// N <= the plan's batch; `partial` holds pl.partial_floats floats when the plan splits K.  Returns the launches made.
inline int wd_launch(const WdPlan& pl, int N, float* partial, float* dw, int sms, cudaStream_t stream) {
  auto kfn = wgrad5_deep_kernel;
#ifndef VNB_EMULATE
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      throw std::runtime_error("CUDA: cannot reserve shared memory for wgrad5_deep_kernel");
    attr = true;
  }
#endif
  WdGeom g = pl.g;
  g.N = N;
  const int grid = std::max(1, std::min(g.n_items, sms));
  float* out = g.ksplit > 1 ? partial : dw;
  VNB_LAUNCH(kfn, grid, kWdThreads, pl.smem, stream, pl.x1_hi, pl.x1_lo, pl.x2_hi, pl.x2_lo, pl.z_hi, pl.z_lo, g, out);
  if (g.ksplit == 1) return 1;
  const size_t n4 = static_cast<size_t>(125) * (g.C1 + g.C2) * g.Cout / 4;
  const int blocks = static_cast<int>(std::min<size_t>((n4 + 255) / 256, 148 * 8));
  VNB_LAUNCH(wgrad5_deep_reduce_kernel, blocks, 256, 0, stream, (const float4*)partial, g.ksplit, n4, (float4*)dw);
  return 2;
}

}  // namespace vnb
