// tcgen05 implicit-GEMM 5x5x5 convolution (fprop and, with flipped/transposed weights, dgrad).
//
// Replaces tf.nn.convolution 5^3 'SAME' + bias (+ residual) -- layers2.py:59-63 via networks.py:264,
// 316,333,346,356 -- and its input gradient (TF autodiff, model.py:660), on the 5th-gen tensor cores.
//
// Formulation ("kw-folded line tiles", im2col-free):
//   * GEMM M = 128 consecutive *input* voxels = whole W-lines (rows r = (d,h,w), w fastest), loaded by
//     ONE 5-D TMA box per (kd,kh) tap pair straight from the NDHWC bf16 activation tensor; SAME
//     padding in d/h comes from TMA out-of-range zero fill, the 5 kw taps are folded into the GEMM N
//     dimension:  D[r][kw*CT + co] = sum_{kd,kh,ci} X[line(r)+(kd,kh)][w(r)][ci] * W[kd][kh][kw][ci][co]
//     so one A tile feeds N = 5*CT columns (80 / 160) instead of CT -- the skinny-N problem of
//     Cout = 16/32 layers (65 % of the FLOPs) becomes an N = 80/160 MMA.
//   * accumulators stay in TMEM over all 25*Cin/KC k-steps; the epilogue re-aligns the 5 kw slices
//     (y[w] = sum_kw D[w+kw-2][kw]) through shared memory, adds bias / residual and stores fp32.
//   * warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps2-5 (and 6-9 in single-pass
//     mode) = epilogue; the producer / issuer warps run their loops converged and one elected lane issues
//     (sm100_ptx.cuh).  Shared-memory rings with full/empty mbarriers; TMEM is a ring of per-tile accumulator
//     slots (tmem_full/tmem_empty per slot) so the epilogue of one item overlaps the main loop of the next;
//     persistent CTAs, one per SM.
//   * precision: bf16 operands, fp32 accumulate.  NSPLIT = 3 runs hi*hi + lo*hi + hi*lo on
//     (hi, lo) bf16 splits of activations and weights = fp32-grade products ("bf16x3").
#pragma once
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "sm100_ptx.cuh"
#include "vnb_cuda.h"

namespace vnb {

using sm100::TmaDesc;

// development counters of the MMA-issuing thread (WgGeom::dbg, TcArgs::dbg); compiled out of the emulation build
#if !defined(VNB_EMULATE) && defined(VNB_KB_COUNTERS)
__device__ __forceinline__ unsigned long long vnb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VNB_DBG_DECL long long dbg_t0 = clock64(), dbg_wait = 0, dbg_wait2 = 0, dbg_n = 0; unsigned long long dbg_g0 = vnb_globaltimer()
#define VNB_DBG_WAITP(ptr, ctr, stmt) do { if (ptr) { const long long t_ = clock64(); stmt; ctr += clock64() - t_; } else { stmt; } } while (0)
#define VNB_DBG_WAIT(stmt) VNB_DBG_WAITP(g.dbg, dbg_wait, stmt)
#define VNB_DBG_COUNT(n) dbg_n += (n)
#define VNB_DBG_STORE(ptr, bar, par) do { if (ptr) { mbar_wait(bar, par); long long* d_ = (ptr) + 8 * blockIdx.x; d_[0] = clock64() - dbg_t0; d_[1] = dbg_wait; \
  d_[2] = dbg_n; d_[3] = static_cast<long long>(vnb_globaltimer() - dbg_g0); d_[4] = dbg_wait2; } } while (0)
#else
#define VNB_DBG_DECL
#define VNB_DBG_WAIT(stmt) stmt
#define VNB_DBG_WAITP(ptr, ctr, stmt) stmt
#define VNB_DBG_COUNT(n)
#define VNB_DBG_STORE(ptr, bar, par)
#endif

constexpr int kTcStages = 4;
constexpr int kTcSlots = 6;                 // barrier pairs reserved for accumulator tile slots (>= any NSLOT)
constexpr int kTcThreads = 192;             // 2 control warps + 4 epilogue warps (one epilogue group)
constexpr int kTcEpiRowPad = 20;  // floats per staged row (16 + 4: conflict-free float4 rows)
constexpr int kTcEpiBytes = 5 * 128 * kTcEpiRowPad * 4;

// KS = filter extent: 5 for the V-Net convolutions, 3 for the attention / output module convolutions
// (attention.py:83-92, tf.pad 1 + VALID == SAME)
template <int CT, int TMAX, int KC, int NSPLIT, int KS = 5>
struct TcCfg {
  static constexpr int ROWB = KC * 2;                       // bytes per smem row (= swizzle span)
  static constexpr int NB = KS * CT;                        // GEMM N (kw-folded)
  static constexpr int R = KS / 2;
  static constexpr int NPL = NSPLIT == 3 ? 2 : 1;           // operand planes (hi, lo)
  // TMEM: the 512 columns are a ring of NSLOT accumulator tile slots of NB columns each; an item takes T <= TMAX
  // consecutive slots (one per 128-row tile), the epilogue hands every slot back as soon as its tile is stored.
  // CT = 16: six slots = two items of three tiles; CT = 32: three slots, items of one or two tiles.
  static constexpr int NSLOT = 512 / NB > 6 ? 6 : 512 / NB;
  static constexpr int TSTREAM = CT == 32 ? 1 : TMAX;       // tiles per item of the streaming (non-resident) pipeline
  static constexpr int A_BYTES = TSTREAM * 128 * ROWB;
  static constexpr int B_BYTES = ((NB * ROWB + 1023) / 1024) * 1024;
  static constexpr int STAGE_BYTES = NPL * (A_BYTES + B_BYTES);
  // single-pass bf16 is epilogue-bound on the Cout = 16 layers: it gets one epilogue group (4 warps, own
  // staging buffer) per TMEM accumulator buffer; the 3-pass mode is MMA-bound and keeps one group
  static constexpr int EG = NSPLIT == 3 ? 1 : 2;
  static constexpr int THREADS = 64 + 128 * EG;
  static constexpr int EPI_BYTES = EG * kTcEpiBytes;
  static constexpr int SMEM_BYTES = kTcStages * STAGE_BYTES + EPI_BYTES + 256 + 1024;
  static constexpr uint32_t LAYOUT = ROWB == 32 ? sm100::SWZ_32B : ROWB == 64 ? sm100::SWZ_64B : sm100::SWZ_128B;
  static_assert(TMAX <= NSLOT && NSLOT * NB <= 512, "TMEM overflow");
  static_assert(NB % 16 == 0 && NB <= 256, "invalid UMMA N");
};

struct TcGeom {
  int N, D, H, W;      // activation extents
  int C1, C2;          // channels of the two concatenated inputs (C2 may be 0)
  int Co1, Co2;        // output channel split (Co2 may be 0)
  int T;               // 128-row MMA tiles per work item (<= TMAX)
  int bh, bd;          // box extent in lines (h) and planes (d): bh*bd == T*lpt
  // row geometry of a work item's A tile.  A "line" occupies LP consecutive rows: LP = W (whole lines, no halo:
  // the kw neighbours of a voxel outside the line are SAME padding) or, for W > 128, LP = Wt + KS - 1 rows of a
  // Wt-wide segment with its halo (TMA zero-fills what lies outside the volume).  MMA tile t starts at row
  // t*tile_rows (tile_rows = lpt*LP <= 128) and always spans 128 rows; rows past lpt lines are ignored.
  int LP, lpt, tile_rows;
  int halo;            // 1: segment mode
  int Wt, n_wb;        // segment width and segments per line (W, 1 without halo)
  int n_hb, n_db;      // blocks per sample along h and d
  int n_slices;        // (Co1+Co2) / CT
  int n_kc;            // (C1+C2) / KC
  int n_items;
  // resident-lines mode (bd == 1): one A load per (kd, k-chunk) holds bh+4 lines and serves all five kh taps
  int resident;
  int a_stage_bytes;   // per operand plane, 1024-aligned
  int n_a, n_b;        // ring depths
};

struct TcArgs {
  TcGeom g;
  const float* bias;   // [Co1+Co2] or nullptr
  const float* res;    // [V][Co1+Co2] or nullptr
  float* out1;
  float* out2;
  int acc1, acc2;
  double* stats = nullptr;    // column kernel only (conv_col.cuh): per-CTA batch-norm partial sums of the output
  long long* dbg = nullptr;   // development counters (tools/kbench.cu): per CTA {loop cycles, cycles waiting on TMA
                              // data, MMAs, wall ns, cycles waiting for a free accumulator}
};

// epilogue shared by both pipeline modes: TMEM -> registers -> shift-sum over the 5 kw slices -> bias /
// residual -> fp32 store.  Runs on warps 2..5 (128 threads = 128 TMEM lanes).
template <int CT, int TMAX, int KC, int NSPLIT, int KS>
__device__ __forceinline__ void conv5_tc_epilogue(const TcArgs& p, float* epi_base, uint32_t tmem, uint32_t tfull0, uint32_t tempty0) {
  using Cfg = TcCfg<CT, TMAX, KC, NSPLIT, KS>;
  constexpr int RC = KS / 2;  // centre tap
  using namespace sm100;
  const TcGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int eg = (warp - 2) >> 2;                       // epilogue group of this warp
  float* epi = epi_base + eg * (kTcEpiBytes / 4);
  const int bar_id = 1 + eg;
  auto tfull_bar = [&](int b) { return tfull0 + 8u * b; };
  auto tempty_bar = [&](int b) { return tempty0 + 8u * b; };
  {
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;     // row inside a 128-row tile
    const int Ctot = g.Co1 + g.Co2;
    int j = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++j) {
      if (Cfg::EG == 2 && (j & 1) != eg) continue;      // two groups: alternate items
      const int slice = item % g.n_slices;
      int x = item / g.n_slices;
      const int wb = x % g.n_wb;
      x /= g.n_wb;
      const int hb = x % g.n_hb;
      x /= g.n_hb;
      const int db = x % g.n_db;
      const int n = x / g.n_db;
      for (int t = 0; t < g.T; ++t) {
        const int gi = j * g.T + t;                     // running tile index of this CTA -> accumulator slot, use count
        const int slot = gi % Cfg::NSLOT;
        const uint32_t use = static_cast<uint32_t>(gi / Cfg::NSLOT);
        mbar_wait(tfull_bar(slot), use & 1u);
        tc_fence_after_sync();
        const int lt = r / g.LP, i = r % g.LP;          // line inside this MMA tile, row inside the line
        const int line = t * g.lpt + lt;
        const int w = g.halo ? wb * g.Wt + i - RC : i;
        const int gh = hb * g.bh + line % g.bh, gd = db * g.bd + line / g.bh;
        const bool valid = lt < g.lpt && line < g.bh * g.bd && gh < g.H && gd < g.D && w >= 0 && w < g.W &&
                           (!g.halo || (i >= RC && i < g.Wt + RC));
        const long long vox = ((static_cast<long long>(n) * g.D + gd) * g.H + gh) * g.W + w;
        const uint32_t t_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + slot * Cfg::NB;
#pragma unroll 1
        for (int cc = 0; cc < CT / 16; ++cc) {
          float acc[16];
          {  // all five kw slices in flight, one wait; the centre slice (kw = 2) never leaves registers
            uint32_t v[KS][16];
#pragma unroll
            for (int kw = 0; kw < KS; ++kw) tmem_ld16(t_addr + kw * CT + cc * 16, v[kw]);
            tmem_ld_wait();
#pragma unroll
            for (int kw = 0; kw < KS; ++kw) {
              if (kw == RC) continue;
              float4* dst = reinterpret_cast<float4*>(epi + (kw * 128 + r) * kTcEpiRowPad);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_float4(__uint_as_float(v[kw][4 * i]), __uint_as_float(v[kw][4 * i + 1]),
                                     __uint_as_float(v[kw][4 * i + 2]), __uint_as_float(v[kw][4 * i + 3]));
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(v[RC][i]);
          }
          named_bar_sync(bar_id, 128);
#pragma unroll
          for (int kw = 0; kw < KS; ++kw) {
            if (kw == RC) continue;
            const int ws = w + kw - RC;
            if (ws >= 0 && ws < g.W) {
              const float4* src = reinterpret_cast<const float4*>(epi + (kw * 128 + r + kw - RC) * kTcEpiRowPad);
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
          named_bar_sync(bar_id, 128);
          if (valid) {
            const int co = slice * CT + cc * 16;  // first of 16 output channels handled here
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] += p.bias[co + i];
            }
            if (p.res) {
              const float4* rs = reinterpret_cast<const float4*>(p.res + vox * Ctot + co);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 f = rs[i];
                acc[4 * i] += f.x;
                acc[4 * i + 1] += f.y;
                acc[4 * i + 2] += f.z;
                acc[4 * i + 3] += f.w;
              }
            }
            float4* o;
            int accumulate;
            if (co < g.Co1) {
              o = reinterpret_cast<float4*>(p.out1 + vox * g.Co1 + co);
              accumulate = p.acc1;
            } else {
              o = reinterpret_cast<float4*>(p.out2 + vox * g.Co2 + (co - g.Co1));
              accumulate = p.acc2;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 f = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
              if (accumulate) {
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
        tc_fence_before_sync();
        mbar_arrive(tempty_bar(slot));  // 128 arrivals hand the tile slot back to the MMA warp
      }
    }
  }
}

// Resident-lines pipeline (plans with bd == 1): ring A holds, per (kd, k-chunk), the bh+4 input lines that
// all five kh taps of the item read (the kh shift is a start-address offset of kh*W rows inside the swizzled
// tile); ring B streams the packed weights per (kd, kh, k-chunk).  Cuts the activation traffic L2->SMEM
// from 25 to 5*(bh+4)/bh line loads per output line.
template <int CT, int TMAX, int KC, int NSPLIT, int KS>
__device__ __forceinline__ void conv5_tc_resident(const TmaDesc& a1_hi, const TmaDesc& a1_lo, const TmaDesc& a2_hi,
                                                  const TmaDesc& a2_lo, const TmaDesc& w_hi, const TmaDesc& w_lo,
                                                  const TcArgs& p, uint8_t* sm, uint32_t sm_addr) {
  using Cfg = TcCfg<CT, TMAX, KC, NSPLIT, KS>;
  constexpr int RC = KS / 2;
  using namespace sm100;
  const TcGeom& g = p.g;
  const int tid = threadIdx.x;
  const int warp = static_cast<int>(warp_uniform(static_cast<uint32_t>(tid >> 5)));
  const uint32_t a_ring = sm_addr;
  const uint32_t b_ring = a_ring + static_cast<uint32_t>(g.n_a) * Cfg::NPL * g.a_stage_bytes;
  const uint32_t epi_off = static_cast<uint32_t>(g.n_a) * Cfg::NPL * g.a_stage_bytes + static_cast<uint32_t>(g.n_b) * Cfg::NPL * Cfg::B_BYTES;
  float* epi = reinterpret_cast<float*>(sm + epi_off);
  const uint32_t bar_base = sm_addr + epi_off + Cfg::EPI_BYTES;
  auto afull = [&](int s) { return bar_base + 8u * s; };          // [4]
  auto aempty = [&](int s) { return bar_base + 8u * (4 + s); };   // [4]
  auto bfull = [&](int s) { return bar_base + 8u * (8 + s); };    // [4]
  auto bempty = [&](int s) { return bar_base + 8u * (12 + s); };  // [4]
  const uint32_t tfull0 = bar_base + 8u * 16, tempty0 = bar_base + 8u * (16 + kTcSlots);   // [kTcSlots] each
  const uint32_t slot_addr = bar_base + 8u * (16 + 2 * kTcSlots);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + epi_off + Cfg::EPI_BYTES + 8 * (16 + 2 * kTcSlots));
  const int kc1 = g.C1 / KC;
  const int n_it = KS * KS * g.n_kc;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    for (int b = 0; b < Cfg::NSLOT; ++b) {
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
  const uint32_t a_rows = static_cast<uint32_t>((g.bh + KS - 1) * g.LP);   // rows loaded per A stage

  // warps 0 and 1 run their loops with all lanes (warp-uniform state); one elected lane waits and issues
  if (warp == 0) {
    const bool leader = elect_one();
    {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      uint32_t afull_cur = afull(0), bfull_cur = bfull(0), a_addr_cur = a_ring, b_addr_cur = b_ring;
      const uint32_t a_stage_all = static_cast<uint32_t>(Cfg::NPL * g.a_stage_bytes);
      for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
        const int slice = item % g.n_slices;
        int x = item / g.n_slices;
        const int wb = x % g.n_wb;
        x /= g.n_wb;
        const int hb = x % g.n_hb;
        x /= g.n_hb;
        const int db = x % g.n_db;
        const int n = x / g.n_db;
        const int h0 = hb * g.bh, d0 = db;
        const int w0 = g.halo ? wb * g.Wt - RC : 0;
        // running ring state (barrier address, destination, weight row) instead of indices: the producer's scalar
        // work per stage adds to the refill latency of a B slot
        const int brow_item = slice * n_it * Cfg::NB, brow_kh = g.n_kc * Cfg::NB;
        for (int kd = 0; kd < KS; ++kd)
          for (int kc = 0; kc < g.n_kc; ++kc) {
            mbar_wait_warp(afull_cur + 32u, aph ^ 1u);   // aempty
            const bool src1 = kc < kc1;
            const int cch = (src1 ? kc : kc - kc1) * KC;
            if (leader) {
              mbar_expect_tx(afull_cur, Cfg::NPL * a_rows * Cfg::ROWB);
              tma_load_5d(a_addr_cur, src1 ? &a1_hi : &a2_hi, afull_cur, cch, w0, h0 - RC, d0 + kd - RC, n);
              if (NSPLIT == 3)
                tma_load_5d(a_addr_cur + g.a_stage_bytes, src1 ? &a1_lo : &a2_lo, afull_cur, cch, w0, h0 - RC, d0 + kd - RC, n);
            }
            afull_cur += 8u;
            a_addr_cur += a_stage_all;
            if (++as == g.n_a) {
              as = 0;
              aph ^= 1u;
              afull_cur = afull(0);
              a_addr_cur = a_ring;
            }
            int brow = brow_item + (kd * KS * g.n_kc + kc) * Cfg::NB;
#pragma unroll
            for (int kh = 0; kh < KS; ++kh) {
              mbar_wait_warp(bfull_cur + 32u, bph ^ 1u);   // bempty
              if (leader) {
                mbar_expect_tx(bfull_cur, Cfg::NPL * Cfg::NB * Cfg::ROWB);
                tma_load_2d(b_addr_cur, &w_hi, bfull_cur, 0, brow);
                if (NSPLIT == 3) tma_load_2d(b_addr_cur + Cfg::B_BYTES, &w_lo, bfull_cur, 0, brow);
              }
              brow += brow_kh;
              bfull_cur += 8u;
              b_addr_cur += Cfg::NPL * Cfg::B_BYTES;
              if (++bs == g.n_b) {
                bs = 0;
                bph ^= 1u;
                bfull_cur = bfull(0);
                b_addr_cur = b_ring;
              }
            }
          }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    {
      // The issue loop is the critical path of the kernel (tools/probe_tcgen05 `rates`, tools/kbench): the tensor pipe
      // only reaches its floor when MMAs arrive back to back, and every scalar instruction between two MMAs is
      // exposed.  Hence: descriptors are advanced incrementally (they differ only in the 14-bit start-address field,
      // in units of 16 bytes), the kh / tile / k-slice loops are fully unrolled, ring state is a running barrier
      // address + descriptor instead of an index, and nothing but the barrier wait sits between two stages.
      const uint32_t idesc = make_instr_desc(128, Cfg::NB, FMT_BF16);
      const uint64_t desc0 = make_smem_desc(0, 16, 8 * Cfg::ROWB, Cfg::LAYOUT);
      const uint32_t a_stride16 = static_cast<uint32_t>(Cfg::NPL * g.a_stage_bytes) >> 4, a_lo16 = static_cast<uint32_t>(g.a_stage_bytes) >> 4;
      constexpr uint32_t b_stride16 = (Cfg::NPL * Cfg::B_BYTES) >> 4, b_lo16 = Cfg::B_BYTES >> 4;
      const uint32_t kh_step16 = static_cast<uint32_t>(g.LP * Cfg::ROWB) >> 4, t_step16 = static_cast<uint32_t>(g.tile_rows * Cfg::ROWB) >> 4;
      const uint64_t da_slot0 = desc0 + (a_ring >> 4), db_slot0 = desc0 + (b_ring >> 4);
      const uint32_t afull0 = afull(0), bfull0 = bfull(0);   // aempty(s) = afull(s) + 32, bempty(s) = bfull(s) + 32
      uint64_t da_cur = da_slot0, db_cur = db_slot0;
      uint32_t afull_cur = afull0, bfull_cur = bfull0;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int j = 0;
      VNB_DBG_DECL;
      for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++j) {
        // accumulator slots of the item's tiles: running tile index gi -> slot gi % NSLOT, use count gi / NSLOT
        uint32_t d_tile[TMAX], tslot[TMAX];
#pragma unroll
        for (int t = 0; t < TMAX; ++t) {
          const int gi = j * g.T + t;
          tslot[t] = static_cast<uint32_t>(gi % Cfg::NSLOT);
          d_tile[t] = tmem + tslot[t] * Cfg::NB;
          if (t < g.T) {
            const uint32_t use = static_cast<uint32_t>(gi / Cfg::NSLOT);
            VNB_DBG_WAITP(p.dbg, dbg_wait2, mbar_wait_warp(tempty0 + 8u * tslot[t], (use & 1u) ^ 1u));
          }
        }
        tc_fence_after_sync();
        uint32_t acc0 = 0u;   // 0 for the first MMA of every tile of the item, 1 afterwards
        const int n_ak = KS * g.n_kc;
        for (int ak = 0; ak < n_ak; ++ak) {   // one A stage per (kd, k-chunk)
          VNB_DBG_WAITP(p.dbg, dbg_wait, mbar_wait_warp(afull_cur, aph));
          tc_fence_after_sync();
          uint64_t da_kh = da_cur;
#pragma unroll
          for (int kh = 0; kh < KS; ++kh) {
            VNB_DBG_WAITP(p.dbg, dbg_wait, mbar_wait_warp(bfull_cur, bph));
            tc_fence_after_sync();
            if (leader) {
#pragma unroll
              for (int t = 0; t < TMAX; ++t) {
                if (t < g.T) {
                  const uint64_t da_t = da_kh + static_cast<uint64_t>(static_cast<uint32_t>(t) * t_step16);
                  const uint32_t d_addr = d_tile[t];
#pragma unroll
                  for (int ks = 0; ks < KC / 16; ++ks) {
                    VNB_DBG_COUNT(NSPLIT == 3 ? 3 : 1);
                    mma_f16_ss(d_addr, da_t + 2 * ks, db_cur + 2 * ks, idesc, ks == 0 ? acc0 : 1u);
                    if (NSPLIT == 3) {
                      mma_f16_ss(d_addr, da_t + a_lo16 + 2 * ks, db_cur + 2 * ks, idesc, 1u);
                      mma_f16_ss(d_addr, da_t + 2 * ks, db_cur + b_lo16 + 2 * ks, idesc, 1u);
                    }
                  }
                }
              }
              mma_commit(bfull_cur + 32u);   // bempty: the slot is free once these MMAs have read it
            }
            acc0 = 1u;
            da_kh += kh_step16;
            db_cur += b_stride16;
            bfull_cur += 8u;
            if (++bs == g.n_b) {
              bs = 0;
              bph ^= 1u;
              db_cur = db_slot0;
              bfull_cur = bfull0;
            }
          }
          mma_commit_if(leader, afull_cur + 32u);   // aempty
          da_cur += a_stride16;
          afull_cur += 8u;
          if (++as == g.n_a) {
            as = 0;
            aph ^= 1u;
            da_cur = da_slot0;
            afull_cur = afull0;
          }
        }
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
          if (t < g.T) mma_commit_if(leader, tfull0 + 8u * tslot[t]);   // the item's accumulators are complete
      }
      if (leader && j > 0) {
        const int gl = j * g.T - 1;   // last tile issued
        (void)gl;
        VNB_DBG_STORE(p.dbg, tfull0 + 8u * (gl % Cfg::NSLOT), (gl / Cfg::NSLOT) & 1);
      }
    }
  } else {
    conv5_tc_epilogue<CT, TMAX, KC, NSPLIT, KS>(p, epi, tmem, tfull0, tempty0);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int CT, int TMAX, int KC, int NSPLIT, int KS>
__global__ void __launch_bounds__(64 + 128 * (NSPLIT == 3 ? 1 : 2), 1)
conv5_tc_kernel(const __grid_constant__ TmaDesc a1_hi, const __grid_constant__ TmaDesc a1_lo,
                const __grid_constant__ TmaDesc a2_hi, const __grid_constant__ TmaDesc a2_lo,
                const __grid_constant__ TmaDesc w_hi, const __grid_constant__ TmaDesc w_lo, const TcArgs p) {
  using Cfg = TcCfg<CT, TMAX, KC, NSPLIT, KS>;
  constexpr int RC = KS / 2;
  using namespace sm100;
  VNB_DYN_SMEM(uint8_t, smem_raw);
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t sm_addr = smem_u32(sm);
  float* epi = reinterpret_cast<float*>(sm + kTcStages * Cfg::STAGE_BYTES);
  const uint32_t bar_base = sm_addr + kTcStages * Cfg::STAGE_BYTES + Cfg::EPI_BYTES;
  // barriers: full[0..S), empty[S..2S), tmem_full[kTcSlots], tmem_empty[kTcSlots]; then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kTcStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kTcStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kTcStages + kTcSlots + b); };
  const uint32_t slot_addr = bar_base + 8u * (2 * kTcStages + 2 * kTcSlots);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + kTcStages * Cfg::STAGE_BYTES + Cfg::EPI_BYTES + 8 * (2 * kTcStages + 2 * kTcSlots));

  const int tid = threadIdx.x;
  const int warp = static_cast<int>(warp_uniform(static_cast<uint32_t>(tid >> 5)));
  const TcGeom& g = p.g;
  const int n_it = KS * KS * g.n_kc;
  const int kc1 = g.C1 / KC;
  if (g.resident) {
    conv5_tc_resident<CT, TMAX, KC, NSPLIT, KS>(a1_hi, a1_lo, a2_hi, a2_lo, w_hi, w_lo, p, sm, sm_addr);
    return;
  }

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < Cfg::NSLOT; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
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

  if (warp == 0) {
    // ======================= TMA producer (whole warp loops, one elected lane issues) =======================
    const bool leader = elect_one();
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
        const int slice = item % g.n_slices;
        int x = item / g.n_slices;
        const int wb = x % g.n_wb;
        x /= g.n_wb;
        const int hb = x % g.n_hb;
        x /= g.n_hb;
        const int db = x % g.n_db;
        const int n = x / g.n_db;
        const int h0 = hb * g.bh, d0 = db * g.bd;
        const int w0 = g.halo ? wb * g.Wt - RC : 0;
        for (int it = 0; it < n_it; ++it) {
          const int kc = it % g.n_kc, kh = (it / g.n_kc) % KS, kd = it / (KS * g.n_kc);
          mbar_wait_warp(empty_bar(stage), phase ^ 1u);
          const uint32_t st_addr = sm_addr + stage * Cfg::STAGE_BYTES;
          const uint32_t a_bytes = static_cast<uint32_t>(g.bh * g.bd * g.LP) * Cfg::ROWB;
          const bool src1 = kc < kc1;
          const int cch = (src1 ? kc : kc - kc1) * KC;
          const int brow = (slice * n_it + it) * Cfg::NB;
          mbar_expect_tx_if(leader, full_bar(stage), Cfg::NPL * (a_bytes + Cfg::NB * Cfg::ROWB));
          tma_load_5d_if(leader, st_addr, src1 ? &a1_hi : &a2_hi, full_bar(stage), cch, w0, h0 + kh - RC, d0 + kd - RC, n);
          tma_load_2d_if(leader, st_addr + Cfg::NPL * Cfg::A_BYTES, &w_hi, full_bar(stage), 0, brow);
          if (NSPLIT == 3) {
            tma_load_5d_if(leader, st_addr + Cfg::A_BYTES, src1 ? &a1_lo : &a2_lo, full_bar(stage), cch, w0, h0 + kh - RC, d0 + kd - RC, n);
            tma_load_2d_if(leader, st_addr + Cfg::NPL * Cfg::A_BYTES + Cfg::B_BYTES, &w_lo, full_bar(stage), 0, brow);
          }
          if (++stage == kTcStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp loops, one elected lane issues) =======================
    const bool leader = elect_one();
    {
      const uint32_t idesc = make_instr_desc(128, Cfg::NB, FMT_BF16);
      int stage = 0;
      uint32_t phase = 0;
      int j = 0;
      VNB_DBG_DECL;
      for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++j) {
        uint32_t d_tile[Cfg::TSTREAM], tslot[Cfg::TSTREAM];
#pragma unroll
        for (int t = 0; t < Cfg::TSTREAM; ++t) {
          const int gi = j * g.T + t;
          tslot[t] = static_cast<uint32_t>(gi % Cfg::NSLOT);
          d_tile[t] = tmem + tslot[t] * Cfg::NB;
          if (t < g.T) {
            const uint32_t use = static_cast<uint32_t>(gi / Cfg::NSLOT);
            VNB_DBG_WAITP(p.dbg, dbg_wait2, mbar_wait_warp(tempty_bar(tslot[t]), (use & 1u) ^ 1u));
          }
        }
        tc_fence_after_sync();
        for (int it = 0; it < n_it; ++it) {
          VNB_DBG_WAITP(p.dbg, dbg_wait, mbar_wait_warp(full_bar(stage), phase));
          tc_fence_after_sync();
          // descriptors differ between MMAs only in the start-address field (bits 0-13, units of 16 B):
          // build one base per operand per stage and advance it with a single 64-bit add
          const uint32_t a_hi = sm_addr + stage * Cfg::STAGE_BYTES;
          const uint64_t da_hi0 = make_smem_desc(a_hi, 16, 8 * Cfg::ROWB, Cfg::LAYOUT);
          const uint64_t da_lo0 = da_hi0 + (Cfg::A_BYTES >> 4);
          const uint64_t db_hi0 = da_hi0 + ((Cfg::NPL * Cfg::A_BYTES) >> 4);
          const uint64_t db_lo0 = db_hi0 + (Cfg::B_BYTES >> 4);
#pragma unroll
          for (int t = 0; t < Cfg::TSTREAM; ++t) {
            if (t >= g.T) continue;
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
              const uint64_t aoff = static_cast<uint64_t>((static_cast<uint32_t>(t * g.tile_rows) * Cfg::ROWB + ks * 32) >> 4);
              const uint64_t boff = static_cast<uint64_t>((ks * 32) >> 4);
              const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
              const uint32_t d_addr = d_tile[t];
              VNB_DBG_COUNT(NSPLIT == 3 ? 3 : 1);
              mma_f16_ss_if(leader, d_addr, da_hi0 + aoff, db_hi0 + boff, idesc, acc);
              if (NSPLIT == 3) {
                mma_f16_ss_if(leader, d_addr, da_lo0 + aoff, db_hi0 + boff, idesc, 1u);
                mma_f16_ss_if(leader, d_addr, da_hi0 + aoff, db_lo0 + boff, idesc, 1u);
              }
            }
          }
          mma_commit_if(leader, empty_bar(stage));  // smem slot reusable once these MMAs have read it
          if (++stage == kTcStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
#pragma unroll
        for (int t = 0; t < Cfg::TSTREAM; ++t)
          if (t < g.T) mma_commit_if(leader, tfull_bar(tslot[t]));  // accumulators of this item complete
      }
      if (leader && j > 0) {
        const int gl = j * g.T - 1;
        (void)gl;
        VNB_DBG_STORE(p.dbg, tfull_bar(gl % Cfg::NSLOT), (gl / Cfg::NSLOT) & 1);
      }
    }
  } else {
    conv5_tc_epilogue<CT, TMAX, KC, NSPLIT, KS>(p, epi, tmem, tfull_bar(0), tempty_bar(0));
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 TF filter [125][Cin][Cout] -> bf16 (hi, lo) GEMM-B tiles
//   packed[slice][it = (kd*5+kh)*n_kc + kc][n = kw*CT + col][k = KC]   (K-major rows of KC elements)
// fprop : col = output channel, k = input channel, taps as stored
// dgrad : col = input channel (the conv's "output"), k = output channel, taps flipped (124 - tap)
// ---------------------------------------------------------------------------------------------
// One block per (slice, it): the [5*CT][KC] tile is gathered through shared memory so that both the fp32
// reads (runs of CT or KC contiguous floats) and the bf16 writes (whole tile contiguous) are coalesced.
__global__ void __launch_bounds__(256) pack_w5_kernel(const float* __restrict__ w, int Cin, int Cout, int dgrad, int CT, int KC,
                                                      uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int Cin_gemm, int KS = 5) {
  // Cin_gemm >= Cin: input channels seen by the GEMM (zero rows beyond the real Cin; fprop only)
  __shared__ float tile[192 * 33];         // [nrow][k] with a padded pitch of KC + 1
  const int Kin = dgrad ? Cout : Cin_gemm;  // GEMM K channels
  const int n_kc = Kin / KC, n_it = KS * KS * n_kc, NB = KS * CT, pitch = KC + 1;
  const int it = blockIdx.x % n_it, slice = blockIdx.x / n_it;
  const int kc = it % n_kc, kh = (it / n_kc) % KS, kd = it / (KS * n_kc);
  const int elems = NB * KC;
  for (int e = threadIdx.x; e < elems; e += blockDim.x) {
    int nrow, k;
    if (!dgrad) {  // source rows: fixed (tap, kch), CT contiguous output channels
      const int cl = e % CT, k_ = (e / CT) % KC, kw = e / (CT * KC);
      nrow = kw * CT + cl;
      k = k_;
      const int tap = (kd * KS + kh) * KS + kw;
      const int kch = kc * KC + k;
      tile[nrow * pitch + k] = kch < Cin ? w[(static_cast<long long>(tap) * Cin + kch) * Cout + slice * CT + cl] : 0.f;
    } else {       // source rows: fixed (tap, input channel), KC contiguous output channels
      k = e % KC;
      nrow = e / KC;
      const int kw = nrow / CT, col = slice * CT + nrow % CT;
      const int tap = KS * KS * KS - 1 - ((kd * KS + kh) * KS + kw);
      tile[nrow * pitch + k] = w[(static_cast<long long>(tap) * Cin + col) * Cout + kc * KC + k];
    }
  }
  __syncthreads();
  const long long base = static_cast<long long>(blockIdx.x) * elems;
  for (int e = threadIdx.x; e < elems; e += blockDim.x) {
    const float v = tile[(e / KC) * pitch + e % KC];
    const uint16_t h = f32_to_bf16(v);
    hi[base + e] = h;
    if (lo) lo[base + e] = f32_to_bf16(v - bf16_to_f32(h));
  }
}

// all layers' packs in one launch: blockIdx.x is mapped to (job, tile) through a prefix table
struct PackJob {
  const float* w;
  uint16_t* hi;
  uint16_t* lo;
  int Cin, Cout, dgrad, CT, KC, Cin_gemm;
  int first_block;   // prefix sum of tiles
  int KS;            // filter extent (5 or 3)
};
__device__ __forceinline__ void pack_w5_tile(const PackJob& j, int tile_idx, float* tile);

__global__ void __launch_bounds__(256) pack_w5_multi_kernel(const PackJob* __restrict__ jobs, int njobs) {
  __shared__ float tile[192 * 33];
  int lo = 0, hi = njobs - 1;
  const int b = blockIdx.x;
  while (lo < hi) {  // last job with first_block <= b
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= b) lo = mid; else hi = mid - 1;
  }
  const PackJob j = jobs[lo];
  pack_w5_tile(j, b - j.first_block, tile);
}

__device__ __forceinline__ void pack_w5_tile(const PackJob& j, int tile_idx, float* tile) {
  const int Kin = j.dgrad ? j.Cout : j.Cin_gemm;
  const int KC = j.KC, CT = j.CT, Cin = j.Cin, Cout = j.Cout, KS = j.KS;
  const int n_kc = Kin / KC, n_it = KS * KS * n_kc, NB = KS * CT, pitch = KC + 1;
  const int it = tile_idx % n_it, slice = tile_idx / n_it;
  const int kc = it % n_kc, kh = (it / n_kc) % KS, kd = it / (KS * n_kc);
  const int elems = NB * KC;
  for (int e = threadIdx.x; e < elems; e += blockDim.x) {
    if (!j.dgrad) {
      const int cl = e % CT, k = (e / CT) % KC, kw = e / (CT * KC);
      const int tap = (kd * KS + kh) * KS + kw, kch = kc * KC + k;
      tile[(kw * CT + cl) * pitch + k] = kch < Cin ? j.w[(static_cast<long long>(tap) * Cin + kch) * Cout + slice * CT + cl] : 0.f;
    } else {
      const int k = e % KC, nrow = e / KC;
      const int kw = nrow / CT, col = slice * CT + nrow % CT;
      const int tap = KS * KS * KS - 1 - ((kd * KS + kh) * KS + kw);
      tile[nrow * pitch + k] = j.w[(static_cast<long long>(tap) * Cin + col) * Cout + kc * KC + k];
    }
  }
  __syncthreads();
  const long long base = static_cast<long long>(tile_idx) * elems;
  for (int e = threadIdx.x; e < elems; e += blockDim.x) {
    const float v = tile[(e / KC) * pitch + e % KC];
    const uint16_t h = f32_to_bf16(v);
    j.hi[base + e] = h;
    if (j.lo) j.lo[base + e] = f32_to_bf16(v - bf16_to_f32(h));
  }
}

// ---------------------------------------------------------------------------------------------
// host helpers: tensor maps and launch plans
// ---------------------------------------------------------------------------------------------
inline void tma_encode(TmaDesc* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, uint32_t swizzle_bytes, uint32_t elem_bytes = 2) {
#ifdef VNB_EMULATE
  TmaDesc d;
  d.base = static_cast<const uint8_t*>(base);
  d.rank = rank;
  d.elem = elem_bytes;
  d.swizzle = swizzle_bytes;
  for (int i = 0; i < rank; ++i) {
    d.dims[i] = dims[i];
    d.strides[i] = i == 0 ? elem_bytes : strides_bytes[i - 1];
    d.box[i] = box[i];
  }
  *out = d;
#else
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f)
      throw std::runtime_error("CUDA: cuTensorMapEncodeTiled unavailable");
    fn = reinterpret_cast<EncodeFn>(f);
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("CUDA: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
#endif
}

// NDHWC bf16 activation [N][D][H][W][C] -> 5-D map (C, W, H, D, N) with box (KC, W, bh, bd, 1)
inline void tma_encode_act(TmaDesc* out, const uint16_t* base, int N, int D, int H, int W, int C, int KC, int bh, int bd, int bw = 0) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
  const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
  const uint32_t box[5] = {(uint32_t)KC, (uint32_t)(bw > 0 ? bw : W), (uint32_t)bh, (uint32_t)bd, 1};
  tma_encode(out, base, 5, dims, str, box, KC * 2);
}
// packed weights [rows][KC] -> 2-D map, box (KC, NB)
inline void tma_encode_w(TmaDesc* out, const uint16_t* base, long long rows, int KC, int NB) {
  const uint64_t dims[2] = {(uint64_t)KC, (uint64_t)rows};
  const uint64_t str[1] = {(uint64_t)KC * 2};
  const uint32_t box[2] = {(uint32_t)KC, (uint32_t)NB};
  tma_encode(out, base, 2, dims, str, box, KC * 2);
}

struct ColPlanGeom {    // conv_col.cuh: plane-sliding columns with resident weights (16 -> 16 channel layers)
  int lpt = 0, n_hb = 0, ds = 0, n_seg = 0, n_a = 0, a_stage_bytes = 0;
};

struct TcKernelPlan {   // one launch of conv5_tc_kernel (or, col: one conv5_col_kernel launch per 16-channel slice and k-chunk)
  bool valid = false;
  bool col = false;
  ColPlanGeom cg{};
  int CT = 0, KC = 0, KS = 5;
  TcGeom g{};
  TmaDesc a1_hi, a1_lo, a2_hi, a2_lo, w_hi, w_lo;
  uint16_t* wp_hi = nullptr;   // packed weights
  uint16_t* wp_lo = nullptr;
  size_t wp_elems = 0;
  size_t smem = 0;             // dynamic shared memory of the launch
};

// geometry for a [N][D][H][W] activation, kernel-side channel counts (C1+C2 in, Co1+Co2 out)
inline bool col_plan_try(TcKernelPlan& pl, int N, int D, int H, int W, int C1, int C2, int Co1, int Co2, bool split3, int sms);

inline bool tc_plan_geometry(TcKernelPlan& pl, int N, int D, int H, int W, int C1, int C2, int Co1, int Co2, bool split3,
                             int ks = 5, int sms = 148) {
  auto mult = [](int v, int m) { return v % m == 0; };
  pl.KS = ks;
  pl.col = false;
  if (ks == 5 && col_plan_try(pl, N, D, H, W, C1, C2, Co1, Co2, split3, sms)) return true;
  if (mult(C1, 32) && mult(C2, 32) && mult(Co1, 32) && mult(Co2, 32) && Co1 > 0) {
    pl.CT = 32;
    pl.KC = 32;
  } else if (mult(C1, 16) && mult(C2, 16) && mult(Co1, 16) && mult(Co2, 16) && mult(Co1 + Co2, 32) && Co1 > 0 && C1 > 0 &&
             ks == 5 && !getenv("VNB_TC_NO_CT32K16")) {
    // 32 output channels over 16-channel k-chunks (the input gradient of decoder level 1: 16 -> 16 + 16): one N = 160
    // slice instead of two N = 80 slices that would each re-read the activation tile
    pl.CT = 32;
    pl.KC = 16;
  } else if (mult(C1, 16) && mult(C2, 16) && mult(Co1, 16) && mult(Co2, 16) && Co1 > 0 && C1 > 0) {
    pl.CT = 16;
    pl.KC = 16;
  } else {
    return false;
  }
  if (W < 1) return false;
  // row geometry (see TcGeom): whole lines when a line fits one 128-row MMA tile, haloed segments otherwise
  int LP, lpt, halo, Wt, n_wb;
  if (W <= 128) {
    halo = 0; Wt = W; n_wb = 1; LP = W; lpt = 128 / W;
  } else {
    halo = 1;
    n_wb = (W + (128 - ks + 1) - 1) / (128 - ks + 1);
    Wt = (W + n_wb - 1) / n_wb;
    LP = Wt + ks - 1;
    lpt = 1;
  }
  if (LP > 256) return false;
  const int tmax = pl.CT == 16 ? 3 : 2;          // template TMAX of the kernel instances
  const int tstream = pl.CT == 16 ? 3 : 1;      // the streaming pipeline of the CT = 32 instances keeps one tile
  for (int pass = 0; pass < 2; ++pass)   // pass 0: boxes that tile the volume exactly; pass 1: accept unused rows
  for (int T = (getenv("VNB_TC_TMAX") ? std::min(tmax, atoi(getenv("VNB_TC_TMAX"))) : tmax); T >= 1; --T) {
    const int lines = T * lpt;
    int bh, bd;
    if (lines <= H) {
      bh = lines;
      bd = 1;
    } else {
      const bool exact = lines % H == 0 && lines / H <= D && D % (lines / H) == 0;
      if (pass == 0 && !exact) continue;
      bh = H;
      bd = std::min(lines / H, D);
    }
    if (bh > 256 || bd > 256) continue;
    TcGeom& g = pl.g;
    g.N = N; g.D = D; g.H = H; g.W = W;
    g.C1 = C1; g.C2 = C2; g.Co1 = Co1; g.Co2 = Co2;
    g.T = T; g.bh = bh; g.bd = bd;
    g.LP = LP; g.lpt = lpt; g.tile_rows = lpt * LP; g.halo = halo; g.Wt = Wt; g.n_wb = n_wb;
    g.n_hb = (H + bh - 1) / bh;
    g.n_db = (D + bd - 1) / bd;
    g.n_slices = (Co1 + Co2) / pl.CT;
    g.n_kc = (C1 + C2) / pl.KC;
    g.n_items = N * g.n_db * g.n_hb * g.n_wb * g.n_slices;
    // shared-memory plan: resident-lines pipeline when the box is a single d-plane and the rings fit
    const int npl = split3 ? 2 : 1;
    const int rowb = pl.KC * 2;
    const int b_bytes = ((ks * pl.CT * rowb + 1023) / 1024) * 1024;
    const int tmax_rows = tstream * 128;
    g.resident = 0;
    g.a_stage_bytes = 0;
    g.n_a = g.n_b = 0;
    const size_t epi_bytes = static_cast<size_t>(split3 ? 1 : 2) * kTcEpiBytes;
    pl.smem = static_cast<size_t>(kTcStages) * npl * (tmax_rows * rowb + b_bytes) + epi_bytes + 256 + 1024;
    if (bd == 1 && !getenv("VNB_TC_NO_RESIDENT")) {
      // rows touched: loads fill (bh + ks - 1) lines, the last MMA tile reads 128 rows from its start row
      const int rows = std::max((bh + ks - 1) * LP, (ks - 1) * LP + (T - 1) * g.tile_rows + 128);
      const int a_stage = ((rows * rowb + 1023) / 1024) * 1024;
      for (int nb = 4; nb >= 2; --nb) {
        const size_t need = 2ull * npl * a_stage + static_cast<size_t>(nb) * npl * b_bytes + epi_bytes + 256 + 1024;
        if (need <= 227 * 1024) {
          g.resident = 1;
          g.a_stage_bytes = a_stage;
          g.n_a = 2;
          g.n_b = nb;
          pl.smem = need;
          break;
        }
      }
    }
    if (T > tstream && !g.resident) continue;   // more tiles than the streaming pipeline holds: try a smaller item
    if (pl.CT == 32 && T == 2 && !getenv("VNB_TC_TMAX")) {
      // Two tiles per weight tile halve the weight traffic (measured per-tile cost 0.9x with three MMA passes,
      // 0.78x with one) but double the item size: keep one tile when the coarser items leave more of the last
      // wave of persistent CTAs idle than that buys.
      const double f = split3 ? 0.9 : 0.78;
      const long long items2 = g.n_items, items1 = 2LL * g.n_items;   // one-tile items of the same layer (exact boxes)
      const double cost2 = static_cast<double>((items2 + sms - 1) / sms) * 2.0 * f;
      const double cost1 = static_cast<double>((items1 + sms - 1) / sms);
      if (cost1 < cost2) continue;
    }
    return true;
  }
  return false;
}

template <int CT, int TMAX, int KC, int NSPLIT, int KS = 5>
inline void tc_launch_inst(const TcKernelPlan& pl, const TcArgs& a, int sms, cudaStream_t stream) {
  using Cfg = TcCfg<CT, TMAX, KC, NSPLIT, KS>;
  auto kfn = conv5_tc_kernel<CT, TMAX, KC, NSPLIT, KS>;
#ifndef VNB_EMULATE
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      throw std::runtime_error("CUDA: cannot reserve shared memory for conv5_tc_kernel");
    attr = true;
  }
#endif
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "streaming pipeline exceeds shared memory");
  const int grid = std::max(1, std::min(a.g.n_items, sms));
  VNB_LAUNCH(kfn, grid, Cfg::THREADS, pl.smem, stream, pl.a1_hi, pl.a1_lo, pl.a2_hi, pl.a2_lo, pl.w_hi, pl.w_lo, a);
}

inline int col_launch(const TcKernelPlan& pl, const TcArgs& a, bool split3, int sms, cudaStream_t stream);

// returns the number of kernels launched
inline int tc_launch(const TcKernelPlan& pl, const TcArgs& a, bool split3, int sms, cudaStream_t stream) {
  if (pl.col) return col_launch(pl, a, split3, sms, stream);
  if (pl.KS == 3) {
    if (pl.CT == 16) {
      if (split3) tc_launch_inst<16, 3, 16, 3, 3>(pl, a, sms, stream);
      else tc_launch_inst<16, 3, 16, 1, 3>(pl, a, sms, stream);
    } else {
      if (split3) tc_launch_inst<32, 2, 32, 3, 3>(pl, a, sms, stream);
      else tc_launch_inst<32, 2, 32, 1, 3>(pl, a, sms, stream);
    }
    return 1;
  }
  if (pl.CT == 16) {
    if (split3) tc_launch_inst<16, 3, 16, 3>(pl, a, sms, stream);
    else tc_launch_inst<16, 3, 16, 1>(pl, a, sms, stream);
  } else if (pl.KC == 16) {
    if (split3) tc_launch_inst<32, 2, 16, 3>(pl, a, sms, stream);
    else tc_launch_inst<32, 2, 16, 1>(pl, a, sms, stream);
  } else {
    if (split3) tc_launch_inst<32, 2, 32, 3>(pl, a, sms, stream);
    else tc_launch_inst<32, 2, 32, 1>(pl, a, sms, stream);
  }
  return 1;
}

}  // namespace vnb

#include "conv_col.cuh"
