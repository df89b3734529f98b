// tcgen05 filter-gradient kernel for the 5x5x5 convolutions (TF autodiff of layers2.py:59-63 under
// optimizer.minimize, model.py:660):   dW[kd][kh][kw][ci][co] = sum_v X[v + (kd,kh,kw) - 2][ci] * dZ[v][co]
//
// GEMM with K = voxels along w.  Both operands are read straight from NDHWC bf16 tiles ([voxel rows]
// [16 channels] = 32-byte rows, SWIZZLE_32B) as MN-major UMMA operands, and two tap axes are folded
// into the MMA by *overlapping atoms* (descriptor semantics pinned on hardware by
// tools/probe_tcgen05.cu, test mn_fold_*):
//   A^T: M = 8 atoms x 16 ci, atom j = the X line shifted by j voxels (LBO = one 32-byte row) -> kw = j
//   B  : N = 5 atoms x 16 co, atom l = the dZ line l lines further down (LBO = line pitch)     -> kh = 4 - l
//   D_kd[(j,ci)][(l,co)] += sum_w X[dx][hx][w + j - 2][ci] * dZ[dx - kd + 2][hx - 2 + l][w][co]
// so one M=128 x N=80 x K=16 MMA advances 25 taps at once (5 of the 8 kw atoms are useful), and the five
// kd planes use five accumulators = 400 of the 512 TMEM columns, resident for the CTA's lifetime.
// A CTA owns one (16-ci chunk, 16-co chunk) pair and a strided share of the voxel slabs (split-K);
// partial filter gradients go to a scratch buffer and are summed in fixed order (deterministic).
//
// warp0 = TMA producer, warp1 = MMA issuer, warps 2-5 = final TMEM read-out.
#pragma once
#include "conv_tc.cuh"

namespace vnb {

constexpr int kWgThreads = 192;

struct WgGeom {
  int N, D, H, W;
  int Wr;            // W rounded up to the MMA K step (16 voxels): boxes are Wr wide, TMA zero-fills past the line end
  int C1, C2, Cout;
  int HT;            // X lines per work item
  int n_hb;          // ceil(H / HT)
  int n_ci, n_co;    // 16-channel chunks
  int splits;        // CTAs per (ci,co) pair
  int z_stages;
  int xt_bytes, zt_bytes;  // per-plane tile sizes (1024-aligned)
  int npl;           // 1 (bf16) or 2 (hi/lo)
  long long* dbg;    // development counters (tools/kbench.cu), null in the product: per CTA {loop cycles, cycles
                     // waiting on TMA data, MMAs issued, wall ns}
};

// KS = 3 serves the attention / output module convolutions with the same scheme: 3 useful kw atoms of 8,
// N = 3 x 16, three kd accumulators.
//
// Loop order ("rolling X planes"): a CTA walks runs of consecutive dZ planes dz for a fixed (sample, line block).
// Step dz loads ONE dZ tile (HT + KS-1 lines) and ONE new X tile (plane dz + R); the X tiles of planes
// dz-R .. dz+R stay resident in a ring of KS+1 slots and accumulator kd pairs dZ plane dz with X plane dz + kd - R.
// Compared with re-loading the KS dZ planes of every X slab this cuts the L2 -> SMEM traffic ~4x (it was the
// bound of the 16-channel layers at 128^3: 6 TB/s over the 148 SMs).  Planes outside the volume are TMA zero fill.
template <int NSPLIT, int KS>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad5_tc_kernel(const __grid_constant__ TmaDesc x1_hi, const __grid_constant__ TmaDesc x1_lo,
                 const __grid_constant__ TmaDesc x2_hi, const __grid_constant__ TmaDesc x2_lo,
                 const __grid_constant__ TmaDesc z_hi, const __grid_constant__ TmaDesc z_lo, const WgGeom g,
                 float* __restrict__ partial /* [split][pair][KS^3][16][16] */) {
  using namespace sm100;
  constexpr int NPL = NSPLIT == 3 ? 2 : 1;
  constexpr int RC = KS / 2, NB = KS * 16, TAPS = KS * KS * KS;
  constexpr int XS = KS + 1;   // X ring slots: KS live planes + one being prefetched
  VNB_DYN_SMEM(uint8_t, smem_raw);
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t sm_addr = smem_u32(sm);
  const uint32_t x_ring = sm_addr;                                      // XS x NPL x xt_bytes
  const uint32_t z_ring = x_ring + static_cast<uint32_t>(XS) * NPL * g.xt_bytes;   // z_stages x NPL x zt_bytes
  const uint32_t bar_off = static_cast<uint32_t>(XS) * NPL * g.xt_bytes + static_cast<uint32_t>(g.z_stages) * NPL * g.zt_bytes;
  const uint32_t bar_base = sm_addr + bar_off;
  auto xfull = [&](int s) { return bar_base + 8u * s; };            // [8]
  auto xempty = [&](int s) { return bar_base + 8u * (8 + s); };     // [8]
  auto zfull = [&](int s) { return bar_base + 8u * (16 + s); };     // [8]
  auto zempty = [&](int s) { return bar_base + 8u * (24 + s); };    // [8]
  const uint32_t done_bar = bar_base + 8u * 32;
  const uint32_t slot_addr = bar_base + 8u * 33;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + bar_off + 8 * 33);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = static_cast<int>(warp_uniform(static_cast<uint32_t>(tid >> 5)));
  const int pair = blockIdx.x / g.splits, split = blockIdx.x % g.splits;
  const int ci_chunk = pair / g.n_co, co_chunk = pair % g.n_co;
  // work units u = ((n * n_hb + hb) * D + dz); this CTA owns the contiguous run [u0, u1)
  const long long U = static_cast<long long>(g.N) * g.n_hb * g.D;
  const long long u0 = U * split / g.splits, u1 = U * (split + 1) / g.splits;
  const int lpm = g.W == 8 ? 2 : 1;                // X lines covered by one K = 16 step
  const int ksteps = g.Wr * lpm / 16;              // MMA k-steps per group of `lpm` lines
  const uint32_t x_pitch = static_cast<uint32_t>(g.Wr + 8) * 32u;  // bytes between X lines in smem
  const uint32_t z_pitch = static_cast<uint32_t>(g.Wr) * 32u;

  if (tid == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
      mbar_init(zfull(s), 1);
      mbar_init(zempty(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(slot_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = warp_uniform(*slot_ptr);

  // warps 0 and 1 run their loops with all lanes (uniform control flow and descriptors); one elected lane waits on
  // the barriers and issues the TMA / MMA / commit instructions
  if (warp == 0) {
    const bool leader = elect_one();
    {
      const bool src1 = ci_chunk * 16 < g.C1;
      const int xc = src1 ? ci_chunk * 16 : ci_chunk * 16 - g.C1;
      const TmaDesc* xh = src1 ? &x1_hi : &x2_hi;
      const TmaDesc* xl = src1 ? &x1_lo : &x2_lo;
      int zs = 0;
      uint32_t zph = 0;
      long long xload = 0;   // running X-plane load index: slot = xload % XS, phase = (xload / XS) & 1
      const uint32_t x_tx = static_cast<uint32_t>(g.HT) * (g.Wr + 8) * 32u * NPL;
      const uint32_t z_tx = static_cast<uint32_t>(g.HT + KS - 1) * g.Wr * 32u * NPL;
      for (long long u = u0; u < u1; ++u) {
        const int dz = static_cast<int>(u % g.D);
        const int hb = static_cast<int>((u / g.D) % g.n_hb);
        const int n = static_cast<int>(u / (static_cast<long long>(g.D) * g.n_hb));
        const int h0 = hb * g.HT;
        const bool chain_start = (u == u0) || dz == 0;
        for (int pl = chain_start ? dz - RC : dz + RC; pl <= dz + RC; ++pl) {
          const int xs = static_cast<int>(xload % XS);
          const uint32_t xph = static_cast<uint32_t>((xload / XS) & 1);
          mbar_wait_warp(xempty(xs), xph ^ 1u);
          mbar_expect_tx_if(leader, xfull(xs), x_tx);
          tma_load_5d_if(leader, x_ring + (xs * NPL) * g.xt_bytes, xh, xfull(xs), xc, -RC, h0, pl, n);
          if (NSPLIT == 3) tma_load_5d_if(leader, x_ring + (xs * NPL + 1) * g.xt_bytes, xl, xfull(xs), xc, -RC, h0, pl, n);
          ++xload;
        }
        mbar_wait_warp(zempty(zs), zph ^ 1u);
        mbar_expect_tx_if(leader, zfull(zs), z_tx);
        tma_load_5d_if(leader, z_ring + (zs * NPL) * g.zt_bytes, &z_hi, zfull(zs), co_chunk * 16, 0, h0 - RC, dz, n);
        if (NSPLIT == 3) tma_load_5d_if(leader, z_ring + (zs * NPL + 1) * g.zt_bytes, &z_lo, zfull(zs), co_chunk * 16, 0, h0 - RC, dz, n);
        if (++zs == g.z_stages) {
          zs = 0;
          zph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    {
      const uint32_t idesc = make_instr_desc(128, NB, FMT_BF16, 1, 1);
      const uint32_t sbo_a = lpm == 2 ? x_pitch : 256u;   // K rows 8..15: next line (W = 8) or next 8 voxels
      const uint32_t sbo_b = 256u;                         // dZ lines are contiguous, so both cases are +256 B
      int zs = 0;
      uint32_t zph = 0;
      long long xlo = 0;     // load index of X plane dz - RC of the current step
      long long x_seen = -1; // newest X-plane load whose barrier this warp has already observed
      VNB_DBG_DECL;
      for (long long u = u0; u < u1; ++u) {
        const int dz = static_cast<int>(u % g.D);
        const bool chain_end = (u + 1 == u1) || dz + 1 == g.D;
        VNB_DBG_WAIT(mbar_wait_warp(zfull(zs), zph));
        tc_fence_after_sync();
        const uint32_t za_hi = z_ring + (zs * NPL) * g.zt_bytes;
        const uint64_t db0 = make_smem_desc(za_hi, z_pitch, sbo_b, SWZ_32B);
        for (int kd = 0; kd < KS; ++kd) {
          const long long xi = xlo + kd;   // X plane dz + kd - RC
          const int xs = static_cast<int>(xi % XS);
          if (xi > x_seen) {   // planes of earlier steps were waited for then (only the newest plane of a step is new)
            VNB_DBG_WAIT(mbar_wait_warp(xfull(xs), static_cast<uint32_t>((xi / XS) & 1)));
            tc_fence_after_sync();
            x_seen = xi;
          }
          const uint32_t xa_hi = x_ring + (xs * NPL) * g.xt_bytes;
          const uint32_t d_addr = tmem + kd * NB;
          // only the start-address field changes between MMAs: one base descriptor per operand, 64-bit adds after
          const uint64_t da0 = make_smem_desc(xa_hi, 32, sbo_a, SWZ_32B);
          uint32_t acc = (u == u0) ? 0u : 1u;
          VNB_DBG_COUNT((NSPLIT == 3 ? 3 : 1) * ((g.HT + lpm - 1) / lpm) * ksteps);
          if (leader) {   // one branch around the whole burst: the MMAs of a (plane, kd) pair issue back to back
            // running descriptors; a line advances A by x_pitch and B by z_pitch, a k-step both by 512 B
            uint64_t da_t = da0, db_t = db0;
            const uint32_t a_line16 = (static_cast<uint32_t>(lpm) * x_pitch) >> 4, b_line16 = (static_cast<uint32_t>(lpm) * z_pitch) >> 4;
            const uint32_t a_lo16 = static_cast<uint32_t>(g.xt_bytes) >> 4, b_lo16 = static_cast<uint32_t>(g.zt_bytes) >> 4;
            for (int t = 0; t < g.HT; t += lpm) {
              uint64_t da = da_t, db = db_t;
              for (int ks = 0; ks < ksteps; ++ks) {
                mma_f16_ss(d_addr, da, db, idesc, acc);
                if (NSPLIT == 3) {
                  mma_f16_ss(d_addr, da + a_lo16, db, idesc, 1u);
                  mma_f16_ss(d_addr, da, db + b_lo16, idesc, 1u);
                }
                acc = 1u;
                da += 32;  // next 16 voxels: 16 rows x 32 B = 512 B (only taken when lpm == 1)
                db += 32;
              }
              da_t += a_line16;
              db_t += b_line16;
            }
          }
        }
        mma_commit_if(leader, zempty(zs));
        if (++zs == g.z_stages) {
          zs = 0;
          zph ^= 1u;
        }
        // X plane dz - RC is dead after this step; at the end of a run of planes so are the other KS-1
        const int dead = chain_end ? KS : 1;
        for (int k = 0; k < dead; ++k) mma_commit_if(leader, xempty(static_cast<int>((xlo + k) % XS)));
        xlo += dead;
      }
      mma_commit_if(leader, done_bar);
      if (leader) {
        VNB_DBG_STORE(g.dbg, done_bar, 0);
      }
    }
  } else {
    // read-out: thread = accumulator row m = (kw slot j, ci); columns n = (l, co); kh = KS-1 - l
    const int q = warp & 3, m = q * 32 + lane;
    const int j = m / 16, ci = m % 16;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
    const bool has_work = u0 < u1;
    float* out = partial + (static_cast<size_t>(split) * (g.n_ci * g.n_co) + pair) * (TAPS * 256);
    for (int kd = 0; kd < KS; ++kd)
      for (int l = 0; l < KS; ++l) {
        uint32_t v[16];
        if (has_work) {
          tmem_ld16(tmem + (static_cast<uint32_t>(q * 32) << 16) + kd * NB + l * 16, v);
          tmem_ld_wait();
        } else {
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (j < KS) {
          const int tap = (kd * KS + (KS - 1 - l)) * KS + j;
          float4* o = reinterpret_cast<float4*>(out + (tap * 16 + ci) * 16);
          for (int i = 0; i < 4; ++i)
            o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                               __uint_as_float(v[4 * i + 3]));
        }
      }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// dw[tap][ci][co] = sum_split partial[split][pair(ci/16, co/16)][tap][ci%16][co%16]   (fixed order: deterministic)
// A block reduces 256 / nsub consecutive outputs per trip: thread group s (of nsub = 1, 2, 4 or 8) sums splits
// s, s+nsub, ... with coalesced reads of one split each, then the group sums are added in group order.  The
// summation tree depends only on (splits, nsub), both fixed by the layer geometry.
__global__ void __launch_bounds__(256) wgrad5_reduce_kernel(const float* __restrict__ partial, int splits, int n_ci, int n_co,
                                                            int Cin, int Cout, float* __restrict__ dw, int taps, int nsub) {
  // Cin = real input channels of dw; the GEMM may have run on a zero-padded multiple of 16 (n_ci chunks)
  __shared__ float red[256];
  const unsigned total = static_cast<unsigned>(taps) * Cin * Cout;   // < 2^31 (largest filter: 8.2 M elements)
  const int pairs = n_ci * n_co;
  const unsigned per = 256u / nsub;
  const unsigned lane = threadIdx.x % per, sub = threadIdx.x / per;
  const size_t split_stride = static_cast<size_t>(pairs) * (static_cast<size_t>(taps) * 256);
  for (unsigned base = blockIdx.x * per; base < total; base += gridDim.x * per) {
    const unsigned i = base + lane;
    float s = 0.f;
    if (i < total) {
      const unsigned co = i % Cout, r = i / Cout;
      const unsigned ci = r % Cin, tap = r / Cin;
      const unsigned pair = (ci / 16) * n_co + co / 16;
      const size_t off = (static_cast<size_t>(pair) * taps + tap) * 256 + (ci % 16) * 16 + co % 16;
      for (int sp = sub; sp < splits; sp += nsub) s += partial[static_cast<size_t>(sp) * split_stride + off];
    }
    if (nsub == 1) {
      if (i < total) dw[i] = s;
      continue;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    if (sub == 0 && i < total) {
      float t = 0.f;
      for (int w = 0; w < nsub; ++w) t += red[w * per + lane];
      dw[i] = t;
    }
    __syncthreads();
  }
}

struct WgPlan {
  bool valid = false;
  WgGeom g{};
  int KS = 5;
  TmaDesc x1_hi, x1_lo, x2_hi, x2_lo, z_hi, z_lo;
  size_t smem = 0;
  size_t partial_floats = 0;
};

inline bool wg_plan_geometry(WgPlan& pl, int N, int D, int H, int W, int C1, int C2, int Cout, bool split3, int sms,
                             int ks = 5) {
  pl.KS = ks;
  if (C1 % 16 || C2 % 16 || Cout % 16 || C1 <= 0) return false;
  if (W < 8) return false;
  const int Wr = W == 8 ? 8 : (W + 15) / 16 * 16;
  WgGeom& g = pl.g;
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.Wr = Wr;
  g.C1 = C1; g.C2 = C2; g.Cout = Cout;
  g.npl = split3 ? 2 : 1;
  int ht = std::max(2, 512 / Wr);
  if (split3) ht = std::max(2, ht / 2);
  ht = std::min(ht, H);
  if (W == 8 && (ht % 2)) return false;
  if (ht + ks - 1 > 256 || Wr + 8 > 256) return false;
  const int xslots = ks + 1;
  const int ht_min = W == 8 ? 2 : 1;
  for (;; ht = std::max(ht_min, ht / 2)) {   // shrink the line block until the X ring + two dZ stages fit
    g.HT = ht;
    g.xt_bytes = ((ht * (Wr + 8) * 32 + 1023) / 1024) * 1024;
    g.zt_bytes = (((ht + ks - 1) * Wr * 32 + 1023) / 1024) * 1024;
    const int budget = 224 * 1024 - xslots * g.npl * g.xt_bytes - 2048;
    g.z_stages = std::min(8, budget / (g.npl * g.zt_bytes));
    if (g.z_stages >= 2 || ht <= ht_min) break;
  }
  if (g.z_stages < 2) return false;
  if (W == 8 && (g.HT % 2)) return false;
  g.n_hb = (H + g.HT - 1) / g.HT;
  g.n_ci = (C1 + C2) / 16;
  g.n_co = Cout / 16;
  const int pairs = g.n_ci * g.n_co;
  const int items = N * D * g.n_hb;
  // one CTA per SM (the rings fill shared memory): never spill a partial second wave of CTAs
  g.splits = std::max(1, std::min(items, sms / pairs));
  pl.smem = static_cast<size_t>(xslots) * g.npl * g.xt_bytes + static_cast<size_t>(g.z_stages) * g.npl * g.zt_bytes + 512 + 1024;
  pl.partial_floats = static_cast<size_t>(g.splits) * pairs * (ks * ks * ks) * 256;
  return true;
}

struct TcConvPlan {
  TcKernelPlan fprop, dgrad;
  WgPlan wgrad;
};

// x tensors: box (16, W+8, HT, 1, 1); dz tensor: box (16, W, HT+4, 1, 1); all SWIZZLE_32B
inline void wg_encode_act(TmaDesc* out, const uint16_t* base, int N, int D, int H, int W, int C, int bw, int bh) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
  const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
  const uint32_t box[5] = {16, (uint32_t)bw, (uint32_t)bh, 1, 1};
  tma_encode(out, base, 5, dims, str, box, 32);
}

inline void wg_encode_plan(WgPlan& pl, int Nmax, const uint16_t* x1_hi, const uint16_t* x1_lo, const uint16_t* x2_hi,
                           const uint16_t* x2_lo, const uint16_t* z_hi, const uint16_t* z_lo) {
  const WgGeom& g = pl.g;
  wg_encode_act(&pl.x1_hi, x1_hi, Nmax, g.D, g.H, g.W, g.C1, g.Wr + 8, g.HT);
  wg_encode_act(&pl.x1_lo, x1_lo ? x1_lo : x1_hi, Nmax, g.D, g.H, g.W, g.C1, g.Wr + 8, g.HT);
  if (g.C2 > 0) {
    wg_encode_act(&pl.x2_hi, x2_hi, Nmax, g.D, g.H, g.W, g.C2, g.Wr + 8, g.HT);
    wg_encode_act(&pl.x2_lo, x2_lo ? x2_lo : x2_hi, Nmax, g.D, g.H, g.W, g.C2, g.Wr + 8, g.HT);
  } else {
    pl.x2_hi = pl.x1_hi;
    pl.x2_lo = pl.x1_lo;
  }
  wg_encode_act(&pl.z_hi, z_hi, Nmax, g.D, g.H, g.W, g.Cout, g.Wr, g.HT + pl.KS - 1);
  wg_encode_act(&pl.z_lo, z_lo ? z_lo : z_hi, Nmax, g.D, g.H, g.W, g.Cout, g.Wr, g.HT + pl.KS - 1);
}

template <int NSPLIT, int KS>
inline void wg_launch_inst(const WgPlan& pl, const WgGeom& g, int grid, float* partial, cudaStream_t stream) {
  auto kfn = wgrad5_tc_kernel<NSPLIT, KS>;
#ifndef VNB_EMULATE
  static bool attr = false;
  if (!attr && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    throw std::runtime_error("CUDA: cannot reserve shared memory for wgrad5_tc_kernel");
  attr = true;
#endif
  VNB_LAUNCH(kfn, grid, kWgThreads, pl.smem, stream, pl.x1_hi, pl.x1_lo, pl.x2_hi, pl.x2_lo, pl.z_hi, pl.z_lo, g, partial);
}

inline void wg_launch(const WgPlan& pl, int N, bool split3, float* partial, float* dw, cudaStream_t stream, int cin_real = 0) {
  WgGeom g = pl.g;
  g.N = N;
  const int pairs = g.n_ci * g.n_co;
  const int grid = pairs * g.splits;
  if (pl.KS == 3) {
    if (split3) wg_launch_inst<3, 3>(pl, g, grid, partial, stream);
    else wg_launch_inst<1, 3>(pl, g, grid, partial, stream);
  } else {
    if (split3) wg_launch_inst<3, 5>(pl, g, grid, partial, stream);
    else wg_launch_inst<1, 5>(pl, g, grid, partial, stream);
  }
  const int cin = cin_real > 0 ? cin_real : g.C1 + g.C2;
  const int taps = pl.KS * pl.KS * pl.KS;
  const long long total = static_cast<long long>(taps) * cin * g.Cout;
  const int nsub = g.splits >= 32 ? 8 : g.splits >= 8 ? 4 : g.splits >= 4 ? 2 : 1;
  const int per = 256 / nsub;
  const int blocks = static_cast<int>(std::min<long long>((total + per - 1) / per, 148 * 16));
  VNB_LAUNCH(wgrad5_reduce_kernel, blocks, 256, 0, stream, (const float*)partial, g.splits, g.n_ci, g.n_co, cin, g.Cout, dw, taps, nsub);
}

}  // namespace vnb
