// tcgen05 filter-gradient kernel for the 5x5x5 convolutions (TF autodiff of layers2.py:59-63 under
// optimizer.minimize, model.py:660):   dW[kd][kh][kw][ci][co] = sum_v X[v + (kd,kh,kw) - 2][ci] * dZ[v][co]
//
// GEMM with K = voxels along w.  Both operands are read straight from NDHWC bf16 tiles ([voxel rows]
// [16 channels] = 32-byte rows, SWIZZLE_32B) as MN-major UMMA operands, and two tap axes are folded
// into the MMA by *overlapping atoms* (descriptor semantics pinned on hardware by
// tools/probe_tcgen05.cu, test mn_fold_*).  With "P" the M-side tensor and "Q" the N-side tensor:
//   A^T: M = 8 atoms x 16 channels of P, atom j = the P line shifted by j voxels (LBO = one 32-byte row) -> kw
//   B  : N = KS line atoms x NCH chunks x 16 channels of Q; the Q tile is stored [line][chunk][w][16], so atom
//        a = line * NCH + chunk sits a * (Wr * 32 B) after the first one (uniform LBO) and advancing the sliding line
//        window by one line moves the start address by NCH atoms                                                   -> kh
//   D_r[(j, cp)][(l, ch, cq)] += sum_w P[q + r - R][h][w + j - R][cp] * Q[q][h - R + l][w][ch, cq]      r = plane offset -> kd
// One M=128 x N x K=16 MMA advances KS*KS taps of NCH chunk pairs at once (KS of the 8 shift atoms are useful).
//   NCH = 1: N = 80, the five plane offsets use five accumulators (400 TMEM columns); MMAs of this shape are bound by
//            their shared-memory operand fetch (52 cycles floor, 58 measured, against N/2 = 40 on the tensor pipe).
//   NCH = 2: N = 160 = tensor-pipe bound (80 cycles for twice the work).  Three accumulators fill TMEM (480 columns), so
//            the plane offsets are split over two CTA groups, r in {0,1,2} and {3,4}, whose split-K shares are sized 3 : 2.
// Roles: normally P = X (a 16-ci chunk) and Q = dZ (co chunks).  When Cout has a single chunk but Cin has an even number
// (decoder level 1: 32 -> 16) the roles are swapped -- P = dZ, Q = X -- which mirrors the tap indices (kd = KS-1-r,
// kh = l, kw = KS-1-j).  If the two Q chunks of a group then come from the two concatenated inputs, the Q tile is filled
// by one TMA box per (line, chunk) instead of one box per tile.
// W segments: the K (w) extent of a line may be cut into n_wb segments (P boxes carry the +-R halo) so that the rings fit.
// A CTA owns one (P chunk, Q chunk group) pair, one plane-offset group and a contiguous share of the voxel slabs
// (split-K); partial filter gradients go to a scratch buffer and are summed in fixed order (deterministic).
//
// warp0 = TMA producer, warp1 = MMA issuer, warps 2-5 = final TMEM read-out.
#pragma once
#include "conv_tc.cuh"
#include "wgrad_deep.cuh"

namespace vnb {

constexpr int kWgThreads = 192;

struct WgGeom {
  int N, D, H, W;
  int Wr;            // segment width rounded up to the MMA K step (16 voxels): boxes are Wr wide, TMA zero-fills past the line end
  int n_wb;          // segments per line
  int C1, C2, Cout;
  int HT;            // P lines per work item
  int n_hb;          // ceil(H / HT)
  int n_ci, n_co;    // 16-channel chunks of X and dZ
  int swap;          // 0: P = X, Q = dZ;  1: P = dZ, Q = X
  int nch;           // Q chunks per CTA (1 or 2) == template NCH
  int n_pc, n_qg;    // P chunks, Q chunk groups
  int q_split;       // one TMA box per (line, chunk) of the Q tile (the chunks of a group come from different tensors)
  int ngroups;       // plane-offset groups (1 or 2)
  int r0[2], nr[2];  // first plane offset and count per group
  int splits[2];     // CTAs per pair and group
  int xs;            // P ring slots (max nr + 1)
  int z_stages;
  int xt_bytes, zt_bytes;  // per-plane tile sizes (1024-aligned)
  int npl;           // 1 (bf16) or 2 (hi/lo)
  long long* dbg;    // development counters (tools/kbench.cu), null in the product: per CTA {loop cycles, cycles
                     // waiting on TMA data, MMAs issued, wall ns}
};

// KS = 3 serves the attention / output module convolutions with the same scheme: 3 useful kw atoms of 8,
// N = 3 x 16 x NCH, three plane offsets in one group.
//
// Loop order ("rolling P planes"): a CTA walks runs of consecutive Q planes q for a fixed (sample, line block, segment).
// Step q loads ONE Q tile (HT + KS-1 lines) and ONE new P tile (plane q + r0 + nr - 1 - R); the P tiles of planes
// q + r0 - R .. q + r0 + nr - 1 - R stay resident in a ring and accumulator r pairs Q plane q with P plane q + r - R.
// Planes outside the volume are TMA zero fill.
template <int NSPLIT, int KS, int NCH>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad5_tc_kernel(const __grid_constant__ TmaDesc p1_hi, const __grid_constant__ TmaDesc p1_lo,
                 const __grid_constant__ TmaDesc p2_hi, const __grid_constant__ TmaDesc p2_lo,
                 const __grid_constant__ TmaDesc q1_hi, const __grid_constant__ TmaDesc q1_lo,
                 const __grid_constant__ TmaDesc q2_hi, const __grid_constant__ TmaDesc q2_lo, const WgGeom g,
                 float* __restrict__ partial /* [slot][pair(ci/16, co/16)][KS^3][16][16] */) {
  using namespace sm100;
  constexpr int NPL = NSPLIT == 3 ? 2 : 1;
  constexpr int RC = KS / 2, NB = KS * 16 * NCH, TAPS = KS * KS * KS;
  constexpr int NRMAX = (512 / NB) < KS ? (512 / NB) : KS;   // accumulators of NB columns that fit TMEM
  constexpr int XS = NRMAX + 1;                               // P ring slots: NRMAX live planes + one being prefetched (== g.xs)
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
  // CTA -> (pair, plane-offset group, split)
  const int per_pair = g.splits[0] + g.splits[1];
  const int pair = blockIdx.x / per_pair;
  int split = blockIdx.x % per_pair;
  const int grp = split >= g.splits[0] ? 1 : 0;
  if (grp) split -= g.splits[0];
  const int nsplit = g.splits[grp], r0 = g.r0[grp], nr = g.nr[grp];
  const int pc = pair / g.n_qg, qg = pair % g.n_qg;
  // work units u = (((n * n_hb + hb) * n_wb + wb) * D + q); this CTA owns the contiguous run [u0, u1)
  // (32-bit on purpose: the role warps divide by these in their loops and 64-bit divisions sit on the issue path)
  const int U = g.N * g.n_hb * g.n_wb * g.D;
  const int u0 = static_cast<int>(static_cast<long long>(U) * split / nsplit);
  const int u1 = static_cast<int>(static_cast<long long>(U) * (split + 1) / nsplit);
  const int lpm = g.W == 8 ? 2 : 1;                // P lines covered by one K = 16 step
  const int ksteps = g.Wr * lpm / 16;              // MMA k-steps per group of `lpm` lines
  const uint32_t x_pitch = static_cast<uint32_t>(g.Wr + 8) * 32u;      // bytes between P lines in smem
  const uint32_t q_atom = static_cast<uint32_t>(g.Wr) * 32u;           // bytes between N atoms (chunks, then lines)
  const uint32_t q_pitch = q_atom * NCH;                               // bytes between Q lines

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
      // P chunk: channel offset inside its source tensor
      const int pch_all = pc * 16;
      const bool psrc1 = g.swap || pch_all < g.C1;
      const int pch = psrc1 ? pch_all : pch_all - g.C1;
      const TmaDesc* ph = psrc1 ? &p1_hi : &p2_hi;
      const TmaDesc* pl_ = psrc1 ? &p1_lo : &p2_lo;
      // Q chunks qg*NCH .. +NCH-1: source tensor and chunk coordinate inside it
      const TmaDesc* qh[NCH];
      const TmaDesc* ql[NCH];
      int qcoord[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int qc = qg * NCH + ch;
        const bool s1 = !g.swap || qc * 16 < g.C1;
        qh[ch] = s1 ? &q1_hi : &q2_hi;
        ql[ch] = s1 ? &q1_lo : &q2_lo;
        qcoord[ch] = s1 ? qc : qc - g.C1 / 16;
      }
      int zs = 0;
      uint32_t zph = 0;
      uint32_t xload = 0;    // running P-plane load index: slot = xload % XS, phase = (xload / XS) & 1
      const uint32_t x_tx = static_cast<uint32_t>(g.HT) * (g.Wr + 8) * 32u * NPL;
      const uint32_t z_tx = static_cast<uint32_t>(g.HT + KS - 1) * g.Wr * 32u * NCH * NPL;
      for (int u = u0; u < u1; ++u) {
        const int q = u % g.D;
        int x = u / g.D;
        const int wb = x % g.n_wb;
        x /= g.n_wb;
        const int hb = x % g.n_hb;
        const int n = x / g.n_hb;
        const int h0 = hb * g.HT, w0 = wb * g.Wr;
        const bool chain_start = (u == u0) || q == 0;
        const int p_lo = q + r0 - RC, p_hi = p_lo + nr - 1;
        for (int pl = chain_start ? p_lo : p_hi; pl <= p_hi; ++pl) {
          const int xs = static_cast<int>(xload % XS);
          const uint32_t xph = (xload / XS) & 1u;
          mbar_wait_warp(xempty(xs), xph ^ 1u);
          mbar_expect_tx_if(leader, xfull(xs), x_tx);
          tma_load_5d_if(leader, x_ring + (xs * NPL) * g.xt_bytes, ph, xfull(xs), pch, w0 - RC, h0, pl, n);
          if (NSPLIT == 3) tma_load_5d_if(leader, x_ring + (xs * NPL + 1) * g.xt_bytes, pl_, xfull(xs), pch, w0 - RC, h0, pl, n);
          ++xload;
        }
        mbar_wait_warp(zempty(zs), zph ^ 1u);
        mbar_expect_tx_if(leader, zfull(zs), z_tx);
        const uint32_t zdst = z_ring + (zs * NPL) * g.zt_bytes;
        const int plane = n * g.D + q;   // Q maps merge (D, N): every Q plane read is a real plane
        if (!g.q_split) {   // one box: (16, Wr, NCH chunks, HT + KS-1 lines, 1)
          tma_load_5d_if(leader, zdst, qh[0], zfull(zs), 0, w0, qcoord[0], h0 - RC, plane);
          if (NSPLIT == 3) tma_load_5d_if(leader, zdst + g.zt_bytes, ql[0], zfull(zs), 0, w0, qcoord[0], h0 - RC, plane);
        } else {            // one box per (line, chunk): (16, Wr, 1, 1, 1)
          for (int l = 0; l < g.HT + KS - 1; ++l)
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
              const uint32_t off = static_cast<uint32_t>(l * NCH + ch) * q_atom;
              tma_load_5d_if(leader, zdst + off, qh[ch], zfull(zs), 0, w0, qcoord[ch], h0 - RC + l, plane);
              if (NSPLIT == 3) tma_load_5d_if(leader, zdst + g.zt_bytes + off, ql[ch], zfull(zs), 0, w0, qcoord[ch], h0 - RC + l, plane);
            }
        }
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
      const uint32_t sbo_b = lpm == 2 ? q_pitch : 256u;
      int zs = 0;
      uint32_t zph = 0;
      uint32_t xlo = 0;      // load index of P plane q + r0 - RC of the current step
      int x_seen = -1;       // newest P-plane load whose barrier this warp has already observed
      int q = u0 % g.D;      // running plane index of the unit (no division in the loop)
      VNB_DBG_DECL;
      for (int u = u0; u < u1; ++u) {
        const bool chain_end = (u + 1 == u1) || q + 1 == g.D;
        VNB_DBG_WAIT(mbar_wait_warp(zfull(zs), zph));
        tc_fence_after_sync();
        const uint32_t za_hi = z_ring + (zs * NPL) * g.zt_bytes;
        const uint64_t db0 = make_smem_desc(za_hi, q_atom, sbo_b, SWZ_32B);
        for (int rr = 0; rr < nr; ++rr) {
          const uint32_t xi = xlo + rr;   // P plane q + r0 + rr - RC
          const int xs = static_cast<int>(xi % XS);
          if (static_cast<int>(xi) > x_seen) {   // planes of earlier steps were waited for then (only the newest plane of a step is new)
            VNB_DBG_WAIT(mbar_wait_warp(xfull(xs), (xi / XS) & 1u));
            tc_fence_after_sync();
            x_seen = static_cast<int>(xi);
          }
          const uint32_t xa_hi = x_ring + (xs * NPL) * g.xt_bytes;
          const uint32_t d_addr = tmem + rr * NB;
          // only the start-address field changes between MMAs: one base descriptor per operand, 64-bit adds after
          const uint64_t da0 = make_smem_desc(xa_hi, 32, sbo_a, SWZ_32B);
          uint32_t acc = (u == u0) ? 0u : 1u;
          VNB_DBG_COUNT((NSPLIT == 3 ? 3 : 1) * ((g.HT + lpm - 1) / lpm) * ksteps);
          if (leader) {   // one branch around the whole burst: the MMAs of a (plane, r) pair issue back to back
            // running descriptors; a line advances A by x_pitch and B by q_pitch, a k-step both by 512 B
            uint64_t da_t = da0, db_t = db0;
            const uint32_t a_line16 = (static_cast<uint32_t>(lpm) * x_pitch) >> 4, b_line16 = (static_cast<uint32_t>(lpm) * q_pitch) >> 4;
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
        // the oldest P plane is dead after this step; at the end of a run of planes so are the others
        const int dead = chain_end ? nr : 1;
        for (int k = 0; k < dead; ++k) mma_commit_if(leader, xempty(static_cast<int>((xlo + k) % XS)));
        xlo += dead;
        if (++q == g.D) q = 0;
      }
      mma_commit_if(leader, done_bar);
      if (leader) {
        VNB_DBG_STORE(g.dbg, done_bar, 0);
      }
    }
  } else {
    // read-out: thread = accumulator row m = (shift atom j, P channel cm); columns n = (line atom l, chunk ch, Q channel)
    const int qd = warp & 3, m = qd * 32 + lane;
    const int j = m / 16, cm = m % 16;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
    const bool has_work = u0 < u1;
    const int slot = grp ? g.splits[0] + split : split;
    float* out = partial + static_cast<size_t>(slot) * (g.n_ci * g.n_co) * (TAPS * 256);
    for (int rr = 0; rr < nr; ++rr)
      for (int l = 0; l < KS; ++l)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          uint32_t v[16];
          if (has_work) {
            tmem_ld16(tmem + (static_cast<uint32_t>(qd * 32) << 16) + rr * NB + (l * NCH + ch) * 16, v);
            tmem_ld_wait();
          } else {
            for (int i = 0; i < 16; ++i) v[i] = 0u;
          }
          if (j < KS) {
            const int r = r0 + rr, qc = qg * NCH + ch;
            if (!g.swap) {   // P = X chunk pc (ci = cm), Q = dZ chunk qc (co = column)
              const int tap = (r * KS + (KS - 1 - l)) * KS + j;
              float4* o = reinterpret_cast<float4*>(out + (static_cast<size_t>(pc * g.n_co + qc) * TAPS + tap) * 256 + cm * 16);
              for (int i = 0; i < 4; ++i)
                o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
            } else {         // P = dZ chunk pc (co = cm), Q = X chunk qc (ci = column): mirrored taps
              const int tap = ((KS - 1 - r) * KS + l) * KS + (KS - 1 - j);
              float* o = out + (static_cast<size_t>(qc * g.n_co + pc) * TAPS + tap) * 256 + cm;
              for (int i = 0; i < 16; ++i) o[i * 16] = __uint_as_float(v[i]);
            }
          }
        }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// dw[tap][ci][co] = sum over the slots of the tap's plane-offset group of partial[slot][pair(ci/16, co/16)][tap][ci%16][co%16]
// (fixed order: deterministic).  A block reduces 256 / nsub consecutive outputs per trip: thread group s (of nsub = 1, 2, 4
// or 8) sums slots s, s+nsub, ... with coalesced reads of one slot each, then the group sums are added in group order.  The
// summation tree depends only on (splits, nsub), both fixed by the layer geometry.
struct WgReduceGeom {
  int n_ci, n_co, Cin, Cout, taps, ks, nsub;
  int swap, ngroups, r_split;   // plane offsets r >= r_split belong to group 1
  int splits[2];
};
__global__ void __launch_bounds__(256) wgrad5_reduce_kernel(const float* __restrict__ partial, WgReduceGeom g, float* __restrict__ dw) {
  // Cin = real input channels of dw; the GEMM may have run on a zero-padded multiple of 16 (n_ci chunks)
  __shared__ float red[256];
  const unsigned total = static_cast<unsigned>(g.taps) * g.Cin * g.Cout;   // < 2^31 (largest filter: 8.2 M elements)
  const int pairs = g.n_ci * g.n_co;
  const unsigned per = 256u / g.nsub;
  const unsigned lane = threadIdx.x % per, sub = threadIdx.x / per;
  const size_t split_stride = static_cast<size_t>(pairs) * (static_cast<size_t>(g.taps) * 256);
  for (unsigned base = blockIdx.x * per; base < total; base += gridDim.x * per) {
    const unsigned i = base + lane;
    float s = 0.f;
    if (i < total) {
      const unsigned co = i % g.Cout, r_ = i / g.Cout;
      const unsigned ci = r_ % g.Cin, tap = r_ / g.Cin;
      const unsigned pair = (ci / 16) * g.n_co + co / 16;
      const size_t off = (static_cast<size_t>(pair) * g.taps + tap) * 256 + (ci % 16) * 16 + co % 16;
      const int kd = static_cast<int>(tap) / (g.ks * g.ks);
      const int r = g.swap ? g.ks - 1 - kd : kd;
      const int grp = (g.ngroups == 2 && r >= g.r_split) ? 1 : 0;
      const int first = grp ? g.splits[0] : 0, cnt = g.splits[grp];
      // fixed order, four loads in flight: slots sub, sub + nsub, ... in four interleaved running sums
      const float* src = partial + static_cast<size_t>(first) * split_stride + off;
      const size_t st = static_cast<size_t>(g.nsub) * split_stride;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int sp = sub;
      for (; sp + 3 * g.nsub < cnt; sp += 4 * g.nsub) {
        const float* q = src + static_cast<size_t>(sp) * split_stride;
        s0 += q[0];
        s1 += q[st];
        s2 += q[2 * st];
        s3 += q[3 * st];
      }
      for (; sp < cnt; sp += g.nsub) s0 += src[static_cast<size_t>(sp) * split_stride];
      s = (s0 + s1) + (s2 + s3);
    }
    if (g.nsub == 1) {
      if (i < total) dw[i] = s;
      continue;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    if (sub == 0 && i < total) {
      float t = 0.f;
      for (int w = 0; w < g.nsub; ++w) t += red[w * per + lane];
      dw[i] = t;
    }
    __syncthreads();
  }
}

struct WgPlan {
  bool valid = false;
  WgGeom g{};
  int KS = 5;
  TmaDesc p1_hi, p1_lo, p2_hi, p2_lo, q1_hi, q1_lo, q2_hi, q2_lo;
  size_t smem = 0;
  size_t partial_floats = 0;
};

inline bool wg_plan_geometry(WgPlan& pl, int N, int D, int H, int W, int C1, int C2, int Cout, bool split3, int sms,
                             int ks = 5) {
  pl.KS = ks;
  if (C1 % 16 || C2 % 16 || Cout % 16 || C1 <= 0) return false;
  if (W < 8) return false;
  WgGeom& g = pl.g;
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.C1 = C1; g.C2 = C2; g.Cout = Cout;
  g.npl = split3 ? 2 : 1;
  g.n_ci = (C1 + C2) / 16;
  g.n_co = Cout / 16;
  // N = 160 (two Q chunks) whenever one of the two tensors has an even number of 16-channel chunks
  g.nch = 1;
  g.swap = 0;
  if (!getenv("VNB_WG_NCH1") && W >= 16) {   // 8-voxel lines: a (plane, r) burst is 12 MMAs, the wider N does not pay
    if (g.n_co % 2 == 0) {
      g.nch = 2;
    } else if (g.n_ci % 2 == 0) {
      g.nch = 2;
      g.swap = 1;
    }
  }
  g.n_pc = g.swap ? g.n_co : g.n_ci;
  g.n_qg = (g.swap ? g.n_ci : g.n_co) / g.nch;
  g.q_split = (g.swap && g.nch == 2 && C2 > 0 && (C1 / 16) % 2 != 0) ? 1 : 0;
  // plane-offset groups: as many accumulators of N columns as TMEM holds
  const int nb = ks * 16 * g.nch;
  const int nr_max = std::min(ks, 512 / nb);
  g.ngroups = (ks + nr_max - 1) / nr_max;
  if (g.ngroups > 2) return false;
  g.r0[0] = 0;
  g.nr[0] = g.ngroups == 1 ? ks : nr_max;
  g.r0[1] = g.nr[0];
  g.nr[1] = ks - g.nr[0];
  g.xs = g.nr[0] + 1;
  // line block / W segment: the candidate with the most rows per step whose rings fit (P ring + at least two Q stages)
  const int ht_min = W == 8 ? 2 : 1;
  long long best = -1;
  for (int nwb = 1; nwb <= 8; nwb *= 2) {
    if (nwb > 1 && (g.nch == 1 || W % (nwb * 16) != 0)) continue;   // NCH = 1 keeps the whole-line geometry of round 1
    const int wseg = W / nwb;
    const int Wr = W == 8 ? 8 : (wseg + 15) / 16 * 16;
    if (Wr + 8 > 256) continue;
    int ht = std::max(2, 512 / Wr);
    if (split3) ht = std::max(2, ht / 2);
    ht = std::min(ht, H);
    if (W == 8 && (ht % 2)) continue;
    if (ht + ks - 1 > 256) continue;
    int xt = 0, zt = 0, zst = 0;
    for (;; ht = std::max(ht_min, ht / 2)) {   // shrink the line block until the P ring + two Q stages fit
      xt = ((ht * (Wr + 8) * 32 + 1023) / 1024) * 1024;
      zt = (((ht + ks - 1) * Wr * 32 * g.nch + 1023) / 1024) * 1024;
      const int budget = 224 * 1024 - g.xs * g.npl * xt - 2048;
      zst = budget > 0 ? std::min(8, budget / (g.npl * zt)) : 0;
      if (zst >= 2 || ht <= ht_min) break;
    }
    if (zst < 2) continue;
    if (W == 8 && (ht % 2)) continue;
    const long long score = static_cast<long long>(ht) * Wr;
    if (score > best) {
      best = score;
      g.Wr = Wr; g.n_wb = nwb; g.HT = ht; g.xt_bytes = xt; g.zt_bytes = zt; g.z_stages = zst;
    }
  }
  if (best < 0) return false;
  g.n_hb = (H + g.HT - 1) / g.HT;
  const int pairs = g.n_pc * g.n_qg;
  const long long items = static_cast<long long>(N) * D * g.n_hb * g.n_wb;
  if (g.ngroups == 1) {
    // one CTA per SM (the rings fill shared memory): never spill a partial second wave of CTAs
    g.splits[0] = static_cast<int>(std::max<long long>(1, std::min<long long>(items, sms / pairs)));
    g.splits[1] = 0;
  } else {   // group shares proportional to their plane offsets, so both finish together
    const int c = static_cast<int>(std::max<long long>(2, std::min<long long>(2 * items, sms / pairs)));
    int s0 = (c * g.nr[0] + ks / 2) / ks;
    s0 = std::max(1, std::min(c - 1, s0));
    g.splits[0] = static_cast<int>(std::min<long long>(s0, items));
    g.splits[1] = static_cast<int>(std::min<long long>(c - s0, items));
  }
  pl.smem = static_cast<size_t>(g.xs) * g.npl * g.xt_bytes + static_cast<size_t>(g.z_stages) * g.npl * g.zt_bytes + 512 + 1024;
  pl.partial_floats = static_cast<size_t>(g.splits[0] + g.splits[1]) * (g.n_ci * g.n_co) * (ks * ks * ks) * 256;
  return true;
}

struct TcConvPlan {
  TcKernelPlan fprop, dgrad;
  WgPlan wgrad;
  WdPlan wgrad_deep;   // per-tap GEMM form for the deep levels (wgrad_deep.cuh); preferred when valid
};

// P tensors: box (16, Wr+8, HT, 1, 1) over (C, W, H, D, N); SWIZZLE_32B
inline void wg_encode_act(TmaDesc* out, const uint16_t* base, int N, int D, int H, int W, int C, int bw, int bh) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
  const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
  const uint32_t box[5] = {16, (uint32_t)bw, (uint32_t)bh, 1, 1};
  tma_encode(out, base, 5, dims, str, box, 32);
}
// Q tensors: the channel axis is cut into (16, chunk) and the (D, N) planes are merged, so that one box
// (16, Wr, chunks, lines, 1) lands as [line][chunk][w][16] in shared memory
inline void wg_encode_q(TmaDesc* out, const uint16_t* base, int N, int D, int H, int W, int C, int bw, int chunks, int lines) {
  const uint64_t dims[5] = {16, (uint64_t)W, (uint64_t)(C / 16), (uint64_t)H, (uint64_t)D * N};
  const uint64_t str[4] = {(uint64_t)C * 2, 32, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[5] = {16, (uint32_t)bw, (uint32_t)chunks, (uint32_t)lines, 1};
  tma_encode(out, base, 5, dims, str, box, 32);
}

inline void wg_encode_plan(WgPlan& pl, int Nmax, const uint16_t* x1_hi, const uint16_t* x1_lo, const uint16_t* x2_hi,
                           const uint16_t* x2_lo, const uint16_t* z_hi, const uint16_t* z_lo) {
  const WgGeom& g = pl.g;
  const int qlines = g.q_split ? 1 : g.HT + pl.KS - 1, qchunks = g.q_split ? 1 : g.nch;
  if (!g.swap) {
    wg_encode_act(&pl.p1_hi, x1_hi, Nmax, g.D, g.H, g.W, g.C1, g.Wr + 8, g.HT);
    wg_encode_act(&pl.p1_lo, x1_lo ? x1_lo : x1_hi, Nmax, g.D, g.H, g.W, g.C1, g.Wr + 8, g.HT);
    if (g.C2 > 0) {
      wg_encode_act(&pl.p2_hi, x2_hi, Nmax, g.D, g.H, g.W, g.C2, g.Wr + 8, g.HT);
      wg_encode_act(&pl.p2_lo, x2_lo ? x2_lo : x2_hi, Nmax, g.D, g.H, g.W, g.C2, g.Wr + 8, g.HT);
    } else {
      pl.p2_hi = pl.p1_hi;
      pl.p2_lo = pl.p1_lo;
    }
    wg_encode_q(&pl.q1_hi, z_hi, Nmax, g.D, g.H, g.W, g.Cout, g.Wr, qchunks, qlines);
    wg_encode_q(&pl.q1_lo, z_lo ? z_lo : z_hi, Nmax, g.D, g.H, g.W, g.Cout, g.Wr, qchunks, qlines);
    pl.q2_hi = pl.q1_hi;
    pl.q2_lo = pl.q1_lo;
  } else {
    wg_encode_act(&pl.p1_hi, z_hi, Nmax, g.D, g.H, g.W, g.Cout, g.Wr + 8, g.HT);
    wg_encode_act(&pl.p1_lo, z_lo ? z_lo : z_hi, Nmax, g.D, g.H, g.W, g.Cout, g.Wr + 8, g.HT);
    pl.p2_hi = pl.p1_hi;
    pl.p2_lo = pl.p1_lo;
    wg_encode_q(&pl.q1_hi, x1_hi, Nmax, g.D, g.H, g.W, g.C1, g.Wr, qchunks, qlines);
    wg_encode_q(&pl.q1_lo, x1_lo ? x1_lo : x1_hi, Nmax, g.D, g.H, g.W, g.C1, g.Wr, qchunks, qlines);
    if (g.C2 > 0) {
      wg_encode_q(&pl.q2_hi, x2_hi, Nmax, g.D, g.H, g.W, g.C2, g.Wr, qchunks, qlines);
      wg_encode_q(&pl.q2_lo, x2_lo ? x2_lo : x2_hi, Nmax, g.D, g.H, g.W, g.C2, g.Wr, qchunks, qlines);
    } else {
      pl.q2_hi = pl.q1_hi;
      pl.q2_lo = pl.q1_lo;
    }
  }
}

template <int NSPLIT, int KS, int NCH>
inline void wg_launch_inst(const WgPlan& pl, const WgGeom& g, int grid, float* partial, cudaStream_t stream) {
  auto kfn = wgrad5_tc_kernel<NSPLIT, KS, NCH>;
#ifndef VNB_EMULATE
  static bool attr = false;
  if (!attr && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    throw std::runtime_error("CUDA: cannot reserve shared memory for wgrad5_tc_kernel");
  attr = true;
#endif
  VNB_LAUNCH(kfn, grid, kWgThreads, pl.smem, stream, pl.p1_hi, pl.p1_lo, pl.p2_hi, pl.p2_lo, pl.q1_hi, pl.q1_lo, pl.q2_hi, pl.q2_lo, g,
             partial);
}

inline void wg_launch(const WgPlan& pl, int N, bool split3, float* partial, float* dw, cudaStream_t stream, int cin_real = 0) {
  WgGeom g = pl.g;
  g.N = N;
  const int pairs = g.n_pc * g.n_qg;
  const int grid = pairs * (g.splits[0] + g.splits[1]);
  if (pl.KS == 3) {
    if (g.nch == 2) {
      if (split3) wg_launch_inst<3, 3, 2>(pl, g, grid, partial, stream);
      else wg_launch_inst<1, 3, 2>(pl, g, grid, partial, stream);
    } else {
      if (split3) wg_launch_inst<3, 3, 1>(pl, g, grid, partial, stream);
      else wg_launch_inst<1, 3, 1>(pl, g, grid, partial, stream);
    }
  } else {
    if (g.nch == 2) {
      if (split3) wg_launch_inst<3, 5, 2>(pl, g, grid, partial, stream);
      else wg_launch_inst<1, 5, 2>(pl, g, grid, partial, stream);
    } else {
      if (split3) wg_launch_inst<3, 5, 1>(pl, g, grid, partial, stream);
      else wg_launch_inst<1, 5, 1>(pl, g, grid, partial, stream);
    }
  }
  WgReduceGeom rg;
  rg.n_ci = g.n_ci;
  rg.n_co = g.n_co;
  rg.Cin = cin_real > 0 ? cin_real : g.C1 + g.C2;
  rg.Cout = g.Cout;
  rg.ks = pl.KS;
  rg.taps = pl.KS * pl.KS * pl.KS;
  rg.swap = g.swap;
  rg.ngroups = g.ngroups;
  rg.r_split = g.r0[1];
  rg.splits[0] = g.splits[0];
  rg.splits[1] = g.splits[1];
  const int smin = g.ngroups == 2 ? std::min(g.splits[0], g.splits[1]) : g.splits[0];
  rg.nsub = smin >= 32 ? 8 : smin >= 8 ? 4 : smin >= 4 ? 2 : 1;
  const long long total = static_cast<long long>(rg.taps) * rg.Cin * g.Cout;
  const int per = 256 / rg.nsub;
  const int blocks = static_cast<int>(std::min<long long>((total + per - 1) / per, 148 * 16));
  VNB_LAUNCH(wgrad5_reduce_kernel, blocks, 256, 0, stream, (const float*)partial, rg, dw);
}

}  // namespace vnb
