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
constexpr int kColSlots = 3;                       // 160-column accumulator slots, one per pair of output planes
constexpr int kColWTile = 80 * 32;                 // one (kd, kh) weight tile: 80 rows x 16 bf16
constexpr int kColEpiBytes = 2 * 4 * 96 * 4 + 1024;        // warp-boundary exchange of the shift-sum epilogue, double-buffered (3 KB) + pad

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
  // batch-norm statistics of the stored values, folded into the epilogue (networks.py:319: tf.layers.batch_normalization
  // right after the convolution): per-CTA (sum z, sum z^2) rows partial[(cta * 2 + q) * stats_c + stats_c0 + channel],
  // the layout bn_finalize_fwd_kernel reads.  nullptr: no statistics (input gradients, all but the last k-chunk).
  double* stats = nullptr;
  int stats_c = 0, stats_c0 = 0;
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
      for (int t = 0; t < 25; ++t) {   // shared-memory order [kh][kd]: the tiles of taps kd, kd + 1 are adjacent (N = 160 operand)
        const int kd = t % KS, kh = t / KS;
        const int row = (p.wrow + (kd * KS + kh) * p.wrow_step) * NB;
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
    // Output planes are accumulated in PAIRS (pb, pb + 1) sharing one 160-column slot: columns [0, 80) belong to the top
    // plane pt = pb + 1, columns [80, 160) to pb.  At input plane p the pair takes taps kd0 = p - pt + 2 (top) and kd0 + 1
    // (bottom), whose weight tiles are adjacent in shared memory ([kh][kd] order), so one N = 160 MMA feeds both planes:
    // tensor-pipe bound (80 cycles) where two N = 80 MMAs cost ~65 each.  Where only one of the two taps exists (the first
    // and last input plane of a pair, volume / segment borders) an N = 80 MMA writes that half alone, and the first MMA
    // of every (pair, step) is issued as two halves because the halves start accumulating at different steps.
    const bool leader = elect_one();
    const uint32_t idesc80 = make_instr_desc(128, 80, FMT_BF16), idesc160 = make_instr_desc(128, 160, FMT_BF16);
    const uint64_t desc0 = make_smem_desc(0, 16, 256, SWZ_32B);
    const uint64_t dw0 = desc0 + (w_base >> 4);
    constexpr uint32_t w_lo16 = Cfg::W_BYTES >> 4, w_tile16 = kColWTile >> 4;
    const uint32_t a_lo16 = static_cast<uint32_t>(g.a_stage_bytes) >> 4, kh_step16 = static_cast<uint32_t>(g.W * 32) >> 4;
    int as = 0;
    uint32_t aph = 0;
    uint32_t gi0 = 0;   // running pair index of this CTA at the item's first pair: slot = gi % 3, use = gi / 3
    VNB_DBG_DECL;
    mbar_wait_warp(wfull, 0);
    tc_fence_after_sync();
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      int n, hb, d0, dn;
      decode(item, n, hb, d0, dn);
      const int npairs = (dn + 1) >> 1;
      const int p_lo = d0 - RC > 0 ? d0 - RC : 0, p_hi = d0 + dn - 1 + RC < g.D - 1 ? d0 + dn - 1 + RC : g.D - 1;
      for (int pl = p_lo; pl <= p_hi; ++pl) {
        VNB_DBG_WAITP(p.dbg, dbg_wait, mbar_wait_warp(afull(as), aph));
        tc_fence_after_sync();
        const uint64_t da0 = desc0 + ((a_ring + static_cast<uint32_t>(as) * NPL * g.a_stage_bytes) >> 4);
        // pairs whose planes this input plane feeds: pb in [pl - 3, pl + 2]
        int k_first = pl - 3 - d0;
        k_first = k_first > 0 ? (k_first + 1) >> 1 : 0;
        uint32_t done_mask = 0;
        for (int j = 0; j < 3; ++j) {
          const int k = k_first + j;
          const int pb = d0 + 2 * k, pt = pb + 1;
          if (k >= npairs || pb > pl + 2) break;
          const bool top = pt < d0 + dn;
          const int kd0 = pl - pt + 2;                      // tap of the top plane; the bottom plane takes kd0 + 1
          const bool v0 = top && kd0 >= 0 && kd0 <= 4, v1 = kd0 + 1 >= 0 && kd0 + 1 <= 4;
          if (!v0 && !v1) continue;
          const uint32_t f0 = (v0 && pl == (pt - RC > 0 ? pt - RC : 0)) ? 1u : 0u;   // first input plane of the top plane
          const uint32_t f1 = (v1 && pl == (pb - RC > 0 ? pb - RC : 0)) ? 1u : 0u;
          const uint32_t gi = gi0 + static_cast<uint32_t>(k), slot = gi % kColSlots;
          if (f1) {   // the pair starts here: its slot must have been drained (the two older pairs of this step go first)
            VNB_DBG_WAITP(p.dbg, dbg_wait2, mbar_wait_warp(tempty0 + 8u * slot, ((gi / kColSlots) & 1u) ^ 1u));
            tc_fence_after_sync();
          }
          const uint32_t d_top = tmem + slot * 160u, d_bot = d_top + 80u;
          const uint64_t db_top = dw0 + static_cast<uint64_t>(static_cast<uint32_t>(kd0 < 0 ? 0 : kd0) * w_tile16);   // tile (kd0, kh = 0)
          const uint64_t db_bot = dw0 + static_cast<uint64_t>(static_cast<uint32_t>(kd0 + 1 > 4 ? 4 : kd0 + 1) * w_tile16);
          VNB_DBG_COUNT((NSPLIT == 3 ? 3 : 1) * KS * ((v0 ? 1 : 0) + (v1 ? 1 : 0)));
          if (leader) {
            uint64_t da = da0;
            if (v0 && v1) {
#pragma unroll
              for (int kh = 0; kh < KS; ++kh) {
                const uint64_t b0 = db_top + static_cast<uint64_t>(kh * KS * w_tile16);
                if (kh == 0) {
                  mma_f16_ss(d_top, da, b0, idesc80, f0 ^ 1u);
                  mma_f16_ss(d_bot, da, b0 + w_tile16, idesc80, f1 ^ 1u);
                } else {
                  mma_f16_ss(d_top, da, b0, idesc160, 1u);
                }
                if (NSPLIT == 3) {
                  mma_f16_ss(d_top, da + a_lo16, b0, idesc160, 1u);
                  mma_f16_ss(d_top, da, b0 + w_lo16, idesc160, 1u);
                }
                da += kh_step16;
              }
            } else {
              const uint32_t d_addr = v0 ? d_top : d_bot, first = v0 ? f0 : f1;
              const uint64_t bb = v0 ? db_top : db_bot;
#pragma unroll
              for (int kh = 0; kh < KS; ++kh) {
                const uint64_t b0 = bb + static_cast<uint64_t>(kh * KS * w_tile16);
                mma_f16_ss(d_addr, da, b0, idesc80, (kh == 0 && first) ? 0u : 1u);
                if (NSPLIT == 3) {
                  mma_f16_ss(d_addr, da + a_lo16, b0, idesc80, 1u);
                  mma_f16_ss(d_addr, da, b0 + w_lo16, idesc80, 1u);
                }
                da += kh_step16;
              }
            }
          }
          __syncwarp();
          // last input plane of the pair: that of its top plane (of the bottom plane when the segment ends on it)
          const int last_plane = top ? pt : pb;
          if (pl == (last_plane + RC < g.D - 1 ? last_plane + RC : g.D - 1)) done_mask |= 1u << slot;
        }
        if (leader) {
          mma_commit(afull(as) + 32u);   // aempty: the stage is free once these MMAs have read it
#pragma unroll
          for (uint32_t sl = 0; sl < kColSlots; ++sl)
            if (done_mask & (1u << sl)) mma_commit(tfull0 + 8u * sl);
        }
        __syncwarp();
        if (++as == g.n_a) {
          as = 0;
          aph ^= 1u;
        }
      }
      gi0 += static_cast<uint32_t>(npairs);
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
    const bool ok0 = w - 2 >= 0, ok1 = w - 1 >= 0, ok3 = w + 1 < g.W, ok4 = w + 2 < g.W;   // kw neighbours inside the line
    uint32_t gi = 0, ti = 0;                 // running pair / plane counters of this CTA
    float st1[16], st2[16];                  // this row's running sum z, sum z^2 over the CTA's tiles (BN statistics)
#pragma unroll
    for (int i = 0; i < 16; ++i) st1[i] = st2[i] = 0.f;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      int n, hb, d0, dn;
      decode(item, n, hb, d0, dn);
      const int gh = hb * g.lpt + lt;
      const bool valid = lt < g.lpt && gh < g.H;
      const int npairs = (dn + 1) >> 1;
      for (int k = 0; k < npairs; ++k, ++gi) {
        const uint32_t slot = gi % kColSlots;
        mbar_wait(tfull0 + 8u * slot, (gi / kColSlots) & 1u);
        tc_fence_after_sync();
        const int pb = d0 + 2 * k;
        const bool top = pb + 1 < d0 + dn;
        for (int half = 1; half >= (top ? 0 : 1); --half, ++ti) {   // bottom plane (columns 80..159) first
          const int d = half ? pb : pb + 1;
          const uint32_t t_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + slot * 160u + static_cast<uint32_t>(half) * 80u;
          float acc[16];
          uint32_t v[KS][16];
#pragma unroll
          for (int kw = 0; kw < KS; ++kw) tmem_ld16(t_addr + kw * 16, v[kw]);
          tmem_ld_wait();
          if (half == (top ? 0 : 1)) {   // both planes of the pair are in registers: hand the slot back
            tc_fence_before_sync();
            mbar_arrive(tempty0 + 8u * slot);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(v[RC][i]);
          // shift-sum over the kw slices, y[w] = sum_kw D[w + kw - 2][kw]: row r needs slice kw of row r + kw - 2, i.e. of the
          // lane two / one below or above -- warp shuffles, plus a 1.5 KB exchange for the two rows either side of a warp
          // boundary (double-buffered by plane parity: one named barrier per plane)
          float* xch = epi + (ti & 1u) * (4 * 96);
          float* mine = xch + q * 96;   // [0,32): slice 0 of lanes 30, 31  [32,48): slice 1 of lane 31  [48,80): slice 4 of lanes 0, 1
                                        // [80,96): slice 3 of lane 0
          if (lane >= 30) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mine[(lane - 30) * 16 + i] = __uint_as_float(v[0][i]);
          }
          if (lane == 31) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mine[32 + i] = __uint_as_float(v[1][i]);
          }
          if (lane <= 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mine[48 + lane * 16 + i] = __uint_as_float(v[4][i]);
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mine[80 + i] = __uint_as_float(v[3][i]);
          }
          named_bar_sync(1, 128);
          const float* below = xch + (q - 1) * 96;   // previous warp = the rows below this warp's first row
          const float* above = xch + (q + 1) * 96;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a0 = __shfl_up_sync(0xffffffffu, __uint_as_float(v[0][i]), 2);
            float a1 = __shfl_up_sync(0xffffffffu, __uint_as_float(v[1][i]), 1);
            float a3 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[3][i]), 1);
            float a4 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[4][i]), 2);
            if (lane < 2) a0 = (q > 0 && ok0) ? below[lane * 16 + i] : 0.f;
            if (lane < 1) a1 = (q > 0 && ok1) ? below[32 + i] : 0.f;
            if (lane > 30) a3 = (q < 3 && ok3) ? above[80 + i] : 0.f;
            if (lane > 29) a4 = (q < 3 && ok4) ? above[48 + (lane - 30) * 16 + i] : 0.f;
            acc[i] += (ok0 ? a0 : 0.f) + (ok1 ? a1 : 0.f) + (ok3 ? a3 : 0.f) + (ok4 ? a4 : 0.f);
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
              st1[4 * i] += f.x;
              st1[4 * i + 1] += f.y;
              st1[4 * i + 2] += f.z;
              st1[4 * i + 3] += f.w;
              st2[4 * i] += f.x * f.x;
              st2[4 * i + 1] += f.y * f.y;
              st2[4 * i + 2] += f.z * f.z;
              st2[4 * i + 3] += f.w * f.w;
            }
          }
        }
      }
    }
    if (p.stats) {   // the CTA's row of the two-stage per-channel reduction: butterfly inside the warp, four warp sums in double
      named_bar_sync(1, 128);          // the exchange buffers of the last tile are no longer read
      float* red = epi;                // [4 warps][32 values]
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a = st1[i], b = st2[i];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, off);
          b += __shfl_xor_sync(0xffffffffu, b, off);
        }
        if (lane == 0) {
          red[q * 32 + i] = a;
          red[q * 32 + 16 + i] = b;
        }
      }
      named_bar_sync(1, 128);
      if (r < 32) {
        const double t = static_cast<double>(red[r]) + static_cast<double>(red[32 + r]) + static_cast<double>(red[64 + r]) +
                         static_cast<double>(red[96 + r]);
        const int qn = r >> 4, c = r & 15;
        p.stats[(static_cast<size_t>(blockIdx.x) * 2 + qn) * p.stats_c + p.stats_c0 + c] = t;
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
  // 16 -> 16, 32 -> 16 and 16 -> 32 channels: one launch per 16-channel slice / k-chunk, at most two.  (32 -> 32 would be
  // four launches that each re-read the activations: measured 0.41 ms against 0.30 ms of conv5_tc_kernel's N = 160
  // instance, which shares an activation tile between the output slices.)
  auto chunks_ok = [](int a, int b) { return a > 0 && a % 16 == 0 && b % 16 == 0 && a + b <= 32; };
  if (!chunks_ok(C1, C2) || !chunks_ok(Co1, Co2)) return false;
  if (((C1 + C2) / 16) * ((Co1 + Co2) / 16) > 2) return false;
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
      if (a.stats && kc == g.n_kc - 1) {   // the values this launch stores are the layer's output: fold the BN statistics in
        c.stats = a.stats;
        c.stats_c = ctot;
        c.stats_c0 = co;
      }
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
