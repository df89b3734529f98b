// tcgen05 5x5x5 convolution for the 16 -> 16 channel layers at full resolution (networks.py:264,316,333: the input
// layer / encoder level 1 / decoder level 1 of networks.VNet, and their input gradients) -- "plane-sliding columns with
// resident weights".
//
// Why a second kernel: conv5_tc_kernel re-loads, for every item of three output lines, the 7 input lines of each of the
// five kd planes plus all 25 weight tiles -- 405 KB of L2 -> shared-memory traffic per 225 MMAs = 35 B/clk/SM on all
// 148 SMs, 81 % of what the L2 delivers chip-wide (6.3 KB/clk): measured, the MMA warp waits for TMA data 22 % of the
// time and an N = 80 MMA costs 86 cycles instead of its 52-cycle floor (profiles/r02_kbench_ct16.txt).  Here
//   * the 25 (kd, kh) weight tiles of the 16 x 16 filter (hi and lo: 128 KB) stay in shared memory for the CTA's lifetime;
//   * a work item is a COLUMN: one 128-row tile position (n, line block) walked along d.  Each step loads ONE input plane
//     tile (tile lines + 4 halo lines) and feeds the five output planes it contributes to (kd = p - d + 2), whose
//     accumulators live in five of the six 80-column TMEM slots; the sixth is being drained by the epilogue.  An input
//     plane is therefore loaded once per column instead of five times: 40 KB per 75 MMAs = 10 B/clk/SM.
// Same GEMM formulation as conv5_tc_kernel (kw folded into N = 5 * 16, SAME padding by TMA zero fill, shift-sum epilogue,
// bf16x3 = hi*hi + lo*hi + hi*lo); wider layers are served as one launch per 16-channel (slice, k-chunk) pair, the later
// k-chunks accumulating into the output.
//
// warp0 = TMA producer, warp1 = MMA issuer (+ TMEM alloc), warps 2-5 = epilogue.
#pragma once
#include "conv_tc.cuh"

namespace vnb {

constexpr int kColThreads = 192;
constexpr int kColSlots = 6;                       // 80-column accumulator slots
constexpr int kColWTile = 80 * 32;                 // one (kd, kh) weight tile: 80 rows x 16 bf16
constexpr int kColEpiBytes = 2 * 128 * kTcEpiRowPad * 4;   // two kw slices staged at a time

struct ColGeom {
  int N, D, H, W;
  int lpt;           // lines per 128-row tile (128 / W)
  int n_hb;          // line blocks per plane
  int ds, n_seg;     // planes per column segment, segments per column
  int n_items;
  int n_a;           // A ring depth
  int a_stage_bytes; // per operand plane, 1024-aligned
};

struct ColArgs {
  ColGeom g;
  int cch;             // first channel of the 16-channel k-chunk inside the A tensor
  int wrow;            // first row of this (slice, k-chunk)'s 25 weight tiles in the packed weights, in tiles of 80 rows:
  int wrow_step;       //   tile (kd, kh) sits at row (wrow + (kd * 5 + kh) * wrow_step) * 80
  const float* bias;   // 16 values or nullptr
  const float* res;    // residual [V][res_stride] (already offset to the slice) or nullptr
  int res_stride;
  float* out;          // [V][out_stride], already offset to the slice's first channel
  int out_stride;
  int accumulate;
  long long* dbg = nullptr;
};

template <int NSPLIT>
struct ColCfg {
  static constexpr int NPL = NSPLIT == 3 ? 2 : 1;
  static constexpr int W_BYTES = 25 * kColWTile;                 // per operand plane (64 000 B)
};

template <int NSPLIT>
__global__ void __launch_bounds__(kColThreads, 1)
conv5_col_kernel(const __grid_constant__ sm100::TmaDesc a_hi, const __grid_constant__ sm100::TmaDesc a_lo,
                 const __grid_constant__ sm100::TmaDesc w_hi, const __grid_constant__ sm100::TmaDesc w_lo, const ColArgs p) {
  using namespace sm100;
  using Cfg = ColCfg<NSPLIT>;
  constexpr int NPL = Cfg::NPL, NB = 80, RC = 2, KS = 5;
  VNB_DYN_SMEM(uint8_t, smem_raw);
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t sm_addr = smem_u32(sm);
  const ColGeom& g = p.g;
  // layout: weights [NPL][25 tiles] | A ring [n_a][NPL][a_stage] | epilogue staging | barriers
  const uint32_t w_base = sm_addr;
  const uint32_t w_bytes = ((NPL * Cfg::W_BYTES + 1023) / 1024) * 1024;
  const uint32_t a_ring = w_base + w_bytes;
  const uint32_t epi_off = w_bytes + static_cast<uint32_t>(g.n_a) * NPL * g.a_stage_bytes;
  float* epi = reinterpret_cast<float*>(sm + epi_off);
  const uint32_t bar_base = sm_addr + epi_off + kColEpiBytes;
  auto afull = [&](int s) { return bar_base + 8u * s; };          // [4]
  auto aempty = [&](int s) { return bar_base + 8u * (4 + s); };   // [4]
  const uint32_t wfull = bar_base + 8u * 8;
  const uint32_t tfull0 = bar_base + 8u * 9, tempty0 = bar_base + 8u * (9 + kColSlots);
  const uint32_t slot_addr = bar_base + 8u * (9 + 2 * kColSlots);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + epi_off + kColEpiBytes + 8 * (9 + 2 * kColSlots));

  const int tid = threadIdx.x;
  const int warp = static_cast<int>(warp_uniform(static_cast<uint32_t>(tid >> 5)));
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
    }
    mbar_init(wfull, 1);
    for (int b = 0; b < kColSlots; ++b) {
      mbar_init(tfull0 + 8u * b, 1);
      mbar_init(tempty0 + 8u * b, 128);
    }
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
  const uint32_t a_rows = static_cast<uint32_t>((g.lpt + KS - 1) * g.W);   // rows loaded per A stage

  // item -> (n, line block, plane segment); steps of an item = input planes [max(d0 - 2, 0), min(d0 + dn + 1, D - 1)]
  auto decode = [&](int item, int& n, int& hb, int& d0, int& dn) {
    const int seg = item % g.n_seg;
    int x = item / g.n_seg;
    hb = x % g.n_hb;
    n = x / g.n_hb;
    d0 = seg * g.ds;
    dn = g.D - d0 < g.ds ? g.D - d0 : g.ds;
  };

  if (warp == 0) {
    // ======================= TMA producer =======================
    const bool leader = elect_one();
    if (leader) {   // the filter: 25 tiles per operand plane, once
      mbar_expect_tx(wfull, NPL * Cfg::W_BYTES);
      for (int t = 0; t < 25; ++t) {
        const int row = (p.wrow + t * p.wrow_step) * NB;
        tma_load_2d(w_base + t * kColWTile, &w_hi, wfull, 0, row);
        if (NSPLIT == 3) tma_load_2d(w_base + Cfg::W_BYTES + t * kColWTile, &w_lo, wfull, 0, row);
      }
    }
    __syncwarp();
    int as = 0;
    uint32_t aph = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      int n, hb, d0, dn;
      decode(item, n, hb, d0, dn);
      const int h0 = hb * g.lpt;
      const int p_lo = d0 - RC > 0 ? d0 - RC : 0, p_hi = d0 + dn - 1 + RC < g.D - 1 ? d0 + dn - 1 + RC : g.D - 1;
      for (int pl = p_lo; pl <= p_hi; ++pl) {
        mbar_wait_warp(aempty(as), aph ^ 1u);
        if (leader) {
          const uint32_t dst = a_ring + static_cast<uint32_t>(as) * NPL * g.a_stage_bytes;
          mbar_expect_tx(afull(as), NPL * a_rows * 32u);
          tma_load_5d(dst, &a_hi, afull(as), p.cch, 0, h0 - RC, pl, n);
          if (NSPLIT == 3) tma_load_5d(dst + g.a_stage_bytes, &a_lo, afull(as), p.cch, 0, h0 - RC, pl, n);
        }
        if (++as == g.n_a) {
          as = 0;
          aph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const bool leader = elect_one();
    const uint32_t idesc = make_instr_desc(128, NB, FMT_BF16);
    const uint64_t desc0 = make_smem_desc(0, 16, 256, SWZ_32B);
    const uint64_t dw0 = desc0 + (w_base >> 4);
    constexpr uint32_t w_lo16 = Cfg::W_BYTES >> 4, w_tile16 = kColWTile >> 4;
    const uint32_t a_lo16 = static_cast<uint32_t>(g.a_stage_bytes) >> 4, kh_step16 = static_cast<uint32_t>(g.W * 32) >> 4;
    int as = 0;
    uint32_t aph = 0;
    uint32_t gi0 = 0;   // running output-plane index of this CTA at the item's first plane: slot = gi % 6, use = gi / 6
    VNB_DBG_DECL;
    mbar_wait_warp(wfull, 0);
    tc_fence_after_sync();
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      int n, hb, d0, dn;
      decode(item, n, hb, d0, dn);
      const int p_lo = d0 - RC > 0 ? d0 - RC : 0, p_hi = d0 + dn - 1 + RC < g.D - 1 ? d0 + dn - 1 + RC : g.D - 1;
      for (int pl = p_lo; pl <= p_hi; ++pl) {
        // output planes fed by this input plane
        const int dlo = pl - RC > d0 ? pl - RC : d0, dhi = pl + RC < d0 + dn - 1 ? pl + RC : d0 + dn - 1;
        uint32_t d_addr[KS], first[KS];
        uint64_t dwk[KS];
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          const int d = dlo + j;
          const uint32_t gi = gi0 + static_cast<uint32_t>(d - d0);
          d_addr[j] = tmem + (gi % kColSlots) * NB;
          first[j] = (pl == (d - RC > 0 ? d - RC : 0)) ? 1u : 0u;      // first input plane of output plane d
          dwk[j] = dw0 + static_cast<uint64_t>(static_cast<uint32_t>((pl - d + RC) * KS) * w_tile16);   // tile (kd, kh = 0)
          if (d <= dhi && first[j]) {   // a fresh accumulator: wait until the epilogue has drained the slot's previous use
            VNB_DBG_WAITP(p.dbg, dbg_wait2, mbar_wait_warp(tempty0 + 8u * (gi % kColSlots), ((gi / kColSlots) & 1u) ^ 1u));
          }
        }
        tc_fence_after_sync();
        VNB_DBG_WAITP(p.dbg, dbg_wait, mbar_wait_warp(afull(as), aph));
        tc_fence_after_sync();
        const uint64_t da0 = desc0 + ((a_ring + static_cast<uint32_t>(as) * NPL * g.a_stage_bytes) >> 4);
        const int nd = dhi - dlo + 1;
        VNB_DBG_COUNT((NSPLIT == 3 ? 3 : 1) * KS * nd);
        if (leader) {
          uint64_t da = da0;
#pragma unroll
          for (int kh = 0; kh < KS; ++kh) {
#pragma unroll
            for (int j = 0; j < KS; ++j) {
              if (j < nd) {
                const uint64_t db = dwk[j] + static_cast<uint64_t>(kh * w_tile16);
                const uint32_t acc = (kh == 0 && first[j]) ? 0u : 1u;
                mma_f16_ss(d_addr[j], da, db, idesc, acc);
                if (NSPLIT == 3) {
                  mma_f16_ss(d_addr[j], da + a_lo16, db, idesc, 1u);
                  mma_f16_ss(d_addr[j], da, db + w_lo16, idesc, 1u);
                }
              }
            }
            da += kh_step16;
          }
          mma_commit(afull(as) + 32u);   // aempty: the stage is free once these MMAs have read it
          // output planes whose last input plane this was: d = pl - 2, and at the last plane of the volume the rest
#pragma unroll
          for (int j = 0; j < KS; ++j) {
            const int d = dlo + j;
            if (j < nd && pl == (d + RC < g.D - 1 ? d + RC : g.D - 1)) {
              const uint32_t gi = gi0 + static_cast<uint32_t>(d - d0);
              mma_commit(tfull0 + 8u * (gi % kColSlots));
            }
          }
        }
        __syncwarp();
        if (++as == g.n_a) {
          as = 0;
          aph ^= 1u;
        }
      }
      gi0 += static_cast<uint32_t>(dn);
    }
    if (leader && gi0 > 0) {
      const uint32_t gl = gi0 - 1;
      (void)gl;
      VNB_DBG_STORE(p.dbg, tfull0 + 8u * (gl % kColSlots), (gl / kColSlots) & 1u);
    }
  } else {
    // ======================= epilogue (warps 2..5 = TMEM lane quarters) =======================
    const int lane = tid & 31;
    const int q = warp & 3;
    const int r = q * 32 + lane;             // row inside the 128-row tile
    const int lt = r / g.W, w = r % g.W;     // line inside the tile, voxel inside the line
    uint32_t gi = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      int n, hb, d0, dn;
      decode(item, n, hb, d0, dn);
      const int gh = hb * g.lpt + lt;
      const bool valid = lt < g.lpt && gh < g.H;
      for (int d = d0; d < d0 + dn; ++d, ++gi) {
        const uint32_t slot = gi % kColSlots;
        mbar_wait(tfull0 + 8u * slot, (gi / kColSlots) & 1u);
        tc_fence_after_sync();
        const uint32_t t_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + slot * NB;
        float acc[16];
        uint32_t v[KS][16];
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) tmem_ld16(t_addr + kw * 16, v[kw]);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(tempty0 + 8u * slot);   // the accumulator is in registers: hand the slot back
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(v[RC][i]);
        // shift-sum over the kw slices, two slices staged at a time: y[w] = sum_kw D[w + kw - 2][kw]
#pragma unroll
        for (int round = 0; round < 2; ++round) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int kw = round == 0 ? s : 3 + s;
            float4* dst = reinterpret_cast<float4*>(epi + (s * 128 + r) * kTcEpiRowPad);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_float4(__uint_as_float(v[kw][4 * i]), __uint_as_float(v[kw][4 * i + 1]),
                                   __uint_as_float(v[kw][4 * i + 2]), __uint_as_float(v[kw][4 * i + 3]));
          }
          named_bar_sync(1, 128);
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int kw = round == 0 ? s : 3 + s;
            const int ws = w + kw - RC;
            if (ws >= 0 && ws < g.W) {
              const float4* src = reinterpret_cast<const float4*>(epi + (s * 128 + r + kw - RC) * kTcEpiRowPad);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 f = src[i];
                acc[4 * i] += f.x;
                acc[4 * i + 1] += f.y;
                acc[4 * i + 2] += f.z;
                acc[4 * i + 3] += f.w;
              }
            }
          }
          named_bar_sync(1, 128);
        }
        if (valid) {
          const long long vox = ((static_cast<long long>(n) * g.D + d) * g.H + gh) * g.W + w;
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += p.bias[i];
          }
          if (p.res) {
            const float4* rs = reinterpret_cast<const float4*>(p.res + vox * p.res_stride);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 f = rs[i];
              acc[4 * i] += f.x;
              acc[4 * i + 1] += f.y;
              acc[4 * i + 2] += f.z;
              acc[4 * i + 3] += f.w;
            }
          }
          float4* o = reinterpret_cast<float4*>(p.out + vox * p.out_stride);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 f = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
            if (p.accumulate) {
              const float4 old = o[i];
              f.x += old.x;
              f.y += old.y;
              f.z += old.z;
              f.w += old.w;
            }
            o[i] = f;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// plan: column geometry for a [N][D][H][W] activation (the caller has checked the channel structure)
inline bool col_plan_geometry(ColGeom& g, size_t& smem, int N, int D, int H, int W, bool split3, int sms) {
  if (W < 8 || W > 128 || 128 % W != 0) return false;
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.lpt = 128 / W;
  g.n_hb = (H + g.lpt - 1) / g.lpt;
  const int npl = split3 ? 2 : 1;
  g.a_stage_bytes = (((g.lpt + 4) * W * 32 + 1023) / 1024) * 1024;
  const int w_bytes = ((npl * 25 * kColWTile + 1023) / 1024) * 1024;
  const int fixed = w_bytes + kColEpiBytes + 256 + 1024;
  g.n_a = std::min(4, (227 * 1024 - fixed) / (npl * g.a_stage_bytes));
  if (g.n_a < 2) return false;
  smem = static_cast<size_t>(fixed) + static_cast<size_t>(g.n_a) * npl * g.a_stage_bytes;
  // plane segments: the shortest columns (>= 16 planes: each segment re-loads four halo planes) that leave the last wave
  // of the persistent CTAs at least 90 % full
  const int cols = N * g.n_hb;
  g.ds = D;
  g.n_seg = 1;
  double best = -1.0;
  for (int nseg = 1; nseg <= std::max(1, D / 16); ++nseg) {
    const int ds = (D + nseg - 1) / nseg;
    const int ns = (D + ds - 1) / ds;
    const long long items = static_cast<long long>(cols) * ns;
    const long long waves = (items + sms - 1) / sms;
    const double eff = static_cast<double>(items) / static_cast<double>(waves * sms) * ds / (ds + 4.0);
    if (eff > best + 1e-9) {
      best = eff;
      g.ds = ds;
      g.n_seg = ns;
    }
  }
  g.n_items = cols * g.n_seg;
  return true;
}

// 16-channel tensors on both sides (one or two of each): the layers of networks.VNet at full resolution
inline bool col_plan_try(TcKernelPlan& pl, int N, int D, int H, int W, int C1, int C2, int Co1, int Co2, bool split3, int sms) {
  if (getenv("VNB_TC_NO_COL")) return false;
  // at most two 16-channel chunks on either side (one launch per slice and k-chunk; wider layers are better served by
  // the N = 160 instances of conv5_tc_kernel, which share an activation tile between the output slices)
  auto chunks_ok = [](int a, int b) { return a > 0 && a % 16 == 0 && b % 16 == 0 && a + b <= 32; };
  if (!chunks_ok(C1, C2) || !chunks_ok(Co1, Co2)) return false;
  ColGeom cg{};
  size_t smem = 0;
  if (!col_plan_geometry(cg, smem, N, D, H, W, split3, sms)) return false;
  pl.col = true;
  pl.CT = 16;
  pl.KC = 16;
  pl.smem = smem;
  pl.cg.lpt = cg.lpt; pl.cg.n_hb = cg.n_hb; pl.cg.ds = cg.ds; pl.cg.n_seg = cg.n_seg; pl.cg.n_a = cg.n_a;
  pl.cg.a_stage_bytes = cg.a_stage_bytes;
  TcGeom& g = pl.g;   // the fields the weight packing and the tensor-map encoding read
  g = TcGeom{};
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.C1 = C1; g.C2 = C2; g.Co1 = Co1; g.Co2 = Co2;
  g.T = 1; g.bh = cg.lpt; g.bd = 1;
  g.LP = W; g.lpt = cg.lpt; g.tile_rows = 128; g.halo = 0; g.Wt = W; g.n_wb = 1;
  g.n_hb = cg.n_hb; g.n_db = D;
  g.n_slices = (Co1 + Co2) / 16;
  g.n_kc = (C1 + C2) / 16;
  g.n_items = cg.n_items;
  g.resident = 1;   // tensor-map boxes carry the tile lines plus the four halo lines
  g.a_stage_bytes = cg.a_stage_bytes;
  g.n_a = cg.n_a;
  g.n_b = 0;
  return true;
}

template <int NSPLIT>
inline void col_launch_inst(const sm100::TmaDesc& a_hi, const sm100::TmaDesc& a_lo, const sm100::TmaDesc& w_hi,
                            const sm100::TmaDesc& w_lo, const ColArgs& a, size_t smem, int sms, cudaStream_t stream) {
  auto kfn = conv5_col_kernel<NSPLIT>;
#ifndef VNB_EMULATE
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      throw std::runtime_error("CUDA: cannot reserve shared memory for conv5_col_kernel");
    attr = true;
  }
#endif
  const int grid = std::max(1, std::min(a.g.n_items, sms));
  VNB_LAUNCH(kfn, grid, kColThreads, smem, stream, a_hi, a_lo, w_hi, w_lo, a);
}

inline int col_launch(const TcKernelPlan& pl, const TcArgs& a, bool split3, int sms, cudaStream_t stream) {
  const TcGeom& g = pl.g;
  ColGeom cg{};
  cg.N = a.g.N; cg.D = g.D; cg.H = g.H; cg.W = g.W;
  cg.lpt = pl.cg.lpt; cg.n_hb = pl.cg.n_hb; cg.ds = pl.cg.ds; cg.n_seg = pl.cg.n_seg; cg.n_a = pl.cg.n_a;
  cg.a_stage_bytes = pl.cg.a_stage_bytes;
  cg.n_items = cg.N * cg.n_hb * cg.n_seg;
  const int kc1 = g.C1 / 16, ctot = g.Co1 + g.Co2;
  int launches = 0;
  for (int s = 0; s < g.n_slices; ++s)
    for (int kc = 0; kc < g.n_kc; ++kc) {
      ColArgs c;
      c.g = cg;
      const bool src1 = kc < kc1;
      c.cch = (src1 ? kc : kc - kc1) * 16;
      c.wrow = s * 25 * g.n_kc + kc;
      c.wrow_step = g.n_kc;
      const int co = s * 16;
      c.bias = (kc == 0 && a.bias) ? a.bias + co : nullptr;
      c.res = (kc == 0 && a.res) ? a.res + co : nullptr;
      c.res_stride = ctot;
      int acc;
      if (co < g.Co1) {
        c.out = a.out1 + co;
        c.out_stride = g.Co1;
        acc = a.acc1;
      } else {
        c.out = a.out2 + (co - g.Co1);
        c.out_stride = g.Co2;
        acc = a.acc2;
      }
      c.accumulate = kc > 0 ? 1 : acc;   // later k-chunks add to what the first one stored
      c.dbg = a.dbg;
      const sm100::TmaDesc& ah = src1 ? pl.a1_hi : pl.a2_hi;
      const sm100::TmaDesc& al = src1 ? pl.a1_lo : pl.a2_lo;
      if (split3) col_launch_inst<3>(ah, al, pl.w_hi, pl.w_lo, c, pl.smem, sms, stream);
      else col_launch_inst<1>(ah, al, pl.w_hi, pl.w_lo, c, pl.smem, sms, stream);
      ++launches;
    }
  return launches;
}

}  // namespace vnb
