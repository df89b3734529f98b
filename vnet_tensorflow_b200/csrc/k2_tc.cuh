// tcgen05 / TMA form of the 2x2x2 stride-2 down convolution and the transposed up convolution
// (layers2.py:65-94 <- networks.py:278,292) and of their input gradients.  Both are GEMMs over the non-overlapping
// 2x2x2 blocks whose "im2col" is a pure re-indexing, so the TMA unit does the gather / the depth-to-space scatter:
//
//   gather   coarse[m][cc]           = sum_{kd,kh,(kw,cf)} fine[child(m; kd,kh,kw)][cf] * w[kd][kh][kw][cf][cc]
//            (down fprop, up dgrad)    M = 128 coarse voxels, K = 8 CF, N = CC
//   scatter  fine[child(m; tap)][cf] = sum_cc coarse[m][cc] * w[tap][cf][cc]
//            (up fprop, down dgrad)    M = 128 coarse voxels, K = CC,   N = 8 CF  (blocks of <= 256 columns)
//
// The fine tensor [N][2Dc][2Hc][2Wc][CF] is addressed through FOUR 4-D tensor maps, one per (kd, kh): dimensions
// ((kw,cf) = 2 CF contiguous floats, ow, oh, (n,od)) with the (kd, kh) offset folded into the base address, so that a
// box (32 floats, ow_t, oh_t, od_t) with ow_t * oh_t * od_t = 128 is exactly the [128 rows][32 K-elements] slab of the
// gather's A operand, and the same box shape is the TMA *store* of 32 accumulator columns of the scatter (the
// depth-to-space interleave happens in the tensor map's strides).  The coarse tensor uses one map of the same shape.
//
// The activations stay fp32 in HBM (they are read exactly once here, 4 B/element, the kernels are HBM-bound), and
// the products are fp32-exact to rounding: a converter warpgroup splits every fp32 chunk that TMA landed in shared
// memory into THREE bf16 pieces (hi, mid, lo: 3 x 8 = the 24 significant bits of an fp32, an exact decomposition) as
// K-major SWIZZLE_64B planes, and the MMA warp issues the six products down to 2^-16 relative weight (mid*mid, lo*hi,
// hi*lo, mid*hi, hi*mid, hi*hi; kind::f16, fp32 accumulate in TMEM) -- the tensor pipe is < 20 % busy either way, so
// the extra passes are free and the error (~1e-7 relative) is below the 3xTF32 mma.sync kernels these replace.  The
// weights are pre-split once per optimiser step (k2tc_pack_multi_kernel) into the [K chunk][hi|mid|lo][N][32] images
// the B loads fetch.  (The filter-gradient kernel below keeps the two-piece split: its slots are converted in place.)
//
// Roles of the 320 threads: warp 0 TMA producer, warp 1 MMA issuer (owns TMEM), warps 2-5 converters (thread = row),
// warps 6-9 epilogue (TMEM -> registers -> swizzled staging -> TMA store, or TMA reduce-add when the destination
// already holds a gradient).  Two accumulator buffers in TMEM: the epilogue of item i runs under the loads,
// conversions and MMAs of item i + 1.
#pragma once
#include "conv_tc.cuh"

namespace vnb {

constexpr int kK2TcThreads = 320;
constexpr int kK2TcA32 = 128 * 128;      // fp32 A chunk as landed by TMA: 128 rows x 32 floats, SWIZZLE_128B
constexpr int kK2TcA16 = 128 * 64;       // one bf16 plane of it: 128 rows x 32 bf16, SWIZZLE_64B (K-major)
constexpr int kK2TcOutBuf = 128 * 128;   // epilogue staging: 128 rows x 32 floats, SWIZZLE_128B
constexpr int kK2TcMaxStages = 4;
constexpr int kK2TcBarBytes = 256;
constexpr int kK2TcBiasBytes = 1024;   // bias table in shared memory (<= 256 distinct channels)

struct K2TcGeom {
  int ow_t, oh_t, od_t;          // tile = ow_t x oh_t x od_t = 128 coarse voxels (od over the merged (n, od) axis)
  int n_tw, n_th, n_td;          // tiles per axis
  int n_kc, in_cpm;              // K chunks of 32 fp32 elements; chunks per input tensor map
  int NB, n_nb, Ntot, out_cpm;   // accumulator columns per item, N blocks, GEMM N, 32-column chunks per output map
  int n_items, stages, stage_bytes, tmem_cols;
  int bias_mod, accumulate;      // bias index = column % bias_mod; accumulate: TMA reduce-add instead of store
};

// (a, b) -> packed bf16 pairs hi = bf16(x) and lo = bf16(x - hi), a in the low half-word (the lower address)
__device__ __forceinline__ void k2tc_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
#else
  const uint16_t ha = f32_to_bf16(a), hb = f32_to_bf16(b);
  hi = static_cast<uint32_t>(ha) | (static_cast<uint32_t>(hb) << 16);
  lo = static_cast<uint32_t>(f32_to_bf16(a - bf16_to_f32(ha))) | (static_cast<uint32_t>(f32_to_bf16(b - bf16_to_f32(hb))) << 16);
#endif
}

// (a, b) -> packed bf16 pairs of the exact three-piece split x = hi + mid + lo
__device__ __forceinline__ void k2tc_split3(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(rb), "f"(ra));
  const float sa = ra - __uint_as_float(mid << 16), sb = rb - __uint_as_float(mid & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(sb), "f"(sa));
#else
  auto piece = [](float& x) {
    const uint16_t h = f32_to_bf16(x);
    x -= bf16_to_f32(h);
    return static_cast<uint32_t>(h);
  };
  const uint32_t ha = piece(a), hb = piece(b), ma = piece(a), mb = piece(b), la = piece(a), lb = piece(b);
  hi = ha | (hb << 16);
  mid = ma | (mb << 16);
  lo = la | (lb << 16);
#endif
}

__global__ void __launch_bounds__(kK2TcThreads, 1)
k2_tc_kernel(const __grid_constant__ sm100::TmaDesc in0, const __grid_constant__ sm100::TmaDesc in1,
             const __grid_constant__ sm100::TmaDesc in2, const __grid_constant__ sm100::TmaDesc in3,
             const __grid_constant__ sm100::TmaDesc out0, const __grid_constant__ sm100::TmaDesc out1,
             const __grid_constant__ sm100::TmaDesc out2, const __grid_constant__ sm100::TmaDesc out3,
             const __grid_constant__ sm100::TmaDesc bmap, const K2TcGeom g, const float* __restrict__ bias) {
  using namespace sm100;
  VNB_DYN_SMEM(uint8_t, smem_raw);
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t sm_addr = smem_u32(sm);
  // layout: stages x {A fp32 | A hi | A mid | A lo | B hi | B mid | B lo} | 2 staging buffers | barriers | bias table
  const uint32_t out_off = static_cast<uint32_t>(g.stages) * g.stage_bytes;
  const uint32_t bar_base = sm_addr + out_off + 2 * kK2TcOutBuf;
  auto full = [&](int s) { return bar_base + 8u * s; };            // TMA landed (A fp32 + B)
  auto conv = [&](int s) { return bar_base + 8u * (4 + s); };      // bf16 planes written (128 converter threads)
  auto empty = [&](int s) { return bar_base + 8u * (8 + s); };     // the stage's MMAs have completed
  auto accf = [&](int b) { return bar_base + 8u * (12 + b); };     // accumulator complete
  auto acce = [&](int b) { return bar_base + 8u * (14 + b); };     // accumulator drained (128 epilogue threads)
  const uint32_t slot_addr = bar_base + 8u * 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + out_off + 2 * kK2TcOutBuf + 8 * 16);

  // the bias sits in shared memory: a global load per column on the epilogue's critical path cost the up convolution
  // 118 us instead of 62 (every chunk waited a round trip through the L2 the stream itself saturates)
  float* bias_s = reinterpret_cast<float*>(sm + out_off + 2 * kK2TcOutBuf + kK2TcBarBytes);

  const int tid = threadIdx.x;
  const int warp = static_cast<int>(warp_uniform(static_cast<uint32_t>(tid >> 5)));
  if (bias != nullptr && tid < g.bias_mod) bias_s[tid] = bias[tid];
  if (tid == 0) {
    for (int s = 0; s < kK2TcMaxStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), 128);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf(b), 1);
      mbar_init(acce(b), 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(slot_addr, static_cast<uint32_t>(g.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = warp_uniform(*slot_ptr);
  const uint32_t b_plane = static_cast<uint32_t>(g.NB) * 64u;   // bytes of one B plane (NB rows x 32 bf16)

  // item -> (N block, tile); the tile's box origin in (ow, oh, (n,od))
  auto decode = [&](int item, int& nb, int& c1, int& c2, int& c3) {
    nb = item % g.n_nb;
    int t = item / g.n_nb;
    c1 = (t % g.n_tw) * g.ow_t;
    t /= g.n_tw;
    c2 = (t % g.n_th) * g.oh_t;
    c3 = (t / g.n_th) * g.od_t;
  };

  if (warp == 0) {
    // ======================= TMA producer =======================
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      int nb, c1, c2, c3;
      decode(item, nb, c1, c2, c3);
      for (int kc = 0; kc < g.n_kc; ++kc) {
        mbar_wait_warp(empty(s), ph ^ 1u);
        if (leader) {
          const uint32_t st = sm_addr + static_cast<uint32_t>(s) * g.stage_bytes;
          const int mi = kc / g.in_cpm, c0 = (kc % g.in_cpm) * 32;
          const TmaDesc* im = mi == 0 ? &in0 : mi == 1 ? &in1 : mi == 2 ? &in2 : &in3;
          mbar_expect_tx(full(s), kK2TcA32 + 3u * b_plane);
          tma_load_4d(st, im, full(s), c0, c1, c2, c3);
          const int brow = kc * 3 * g.Ntot + nb * g.NB;
          for (int pc = 0; pc < 3; ++pc)
            tma_load_2d(st + kK2TcA32 + 3 * kK2TcA16 + pc * b_plane, &bmap, full(s), 0, brow + pc * g.Ntot);
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
    const uint32_t idesc = make_instr_desc(128, static_cast<uint32_t>(g.NB), FMT_BF16);
    const uint64_t desc0 = make_smem_desc(0, 16, 512, SWZ_64B);
    int s = 0;
    uint32_t ph = 0, it = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      mbar_wait_warp(acce(buf), (use & 1u) ^ 1u);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem + buf * static_cast<uint32_t>(g.NB);
      for (int kc = 0; kc < g.n_kc; ++kc) {
        mbar_wait_warp(conv(s), ph);
        tc_fence_after_sync();
        const uint32_t a_hi = sm_addr + static_cast<uint32_t>(s) * g.stage_bytes + kK2TcA32;
        const uint32_t b_hi = a_hi + 3 * kK2TcA16;
        const uint64_t a16 = kK2TcA16 >> 4, b16 = b_plane >> 4;
        const uint64_t dah = desc0 + (a_hi >> 4), dam = dah + a16, dal = dam + a16;
        const uint64_t dbh = desc0 + (b_hi >> 4), dbm = dbh + b16, dbl = dbm + b16;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {   // two K = 16 steps per 32-element chunk: start address + 32 bytes
          const uint64_t o = static_cast<uint64_t>(ks * 2);
          mma_f16_ss_if(leader, d_tmem, dam + o, dbm + o, idesc, (kc | ks) != 0 ? 1u : 0u);   // smallest terms first
          mma_f16_ss_if(leader, d_tmem, dal + o, dbh + o, idesc, 1u);
          mma_f16_ss_if(leader, d_tmem, dah + o, dbl + o, idesc, 1u);
          mma_f16_ss_if(leader, d_tmem, dam + o, dbh + o, idesc, 1u);
          mma_f16_ss_if(leader, d_tmem, dah + o, dbm + o, idesc, 1u);
          mma_f16_ss_if(leader, d_tmem, dah + o, dbh + o, idesc, 1u);
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
  } else if (warp < 6) {
    // ======================= converters: fp32 chunk -> bf16 (hi, lo) K-major SWIZZLE_64B planes =======================
    const int r = tid - 64;                       // row of the 128-row tile
    const uint32_t sw128 = static_cast<uint32_t>(r & 7), sw64 = static_cast<uint32_t>((r >> 1) & 3);
    const uint32_t src_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t dst_off = kK2TcA32 + static_cast<uint32_t>(r >> 3) * 512u + static_cast<uint32_t>(r & 7) * 64u;
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      for (int kc = 0; kc < g.n_kc; ++kc) {
        mbar_wait(full(s), ph);
        uint8_t* st = sm + static_cast<size_t>(s) * g.stage_bytes;
        float4 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = *reinterpret_cast<const float4*>(st + src_off + ((static_cast<uint32_t>(j) ^ sw128) << 4));
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // eight K elements = one 16-byte unit of each plane
          const float4 a = x[2 * q], b = x[2 * q + 1];
          uint4 h, m, l;
          k2tc_split3(a.x, a.y, h.x, m.x, l.x);
          k2tc_split3(a.z, a.w, h.y, m.y, l.y);
          k2tc_split3(b.x, b.y, h.z, m.z, l.z);
          k2tc_split3(b.z, b.w, h.w, m.w, l.w);
          const uint32_t o = dst_off + ((static_cast<uint32_t>(q) ^ sw64) << 4);
          *reinterpret_cast<uint4*>(st + o) = h;
          *reinterpret_cast<uint4*>(st + o + kK2TcA16) = m;
          *reinterpret_cast<uint4*>(st + o + 2 * kK2TcA16) = l;
        }
        fence_proxy_async_smem();   // the planes are read by the tensor core through the async proxy
        mbar_arrive(conv(s));
        if (++s == g.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ======================= epilogue: TMEM -> (+ bias) -> swizzled staging -> TMA store / reduce-add =======================
    const int lane = tid & 31, q = warp & 3;
    const int r = q * 32 + lane;
    const bool issuer = tid == 192;
    const uint32_t sw128 = static_cast<uint32_t>(r & 7);
    const int nchunks = g.NB / 32;
    uint32_t it = 0, jc = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
      int nb, c1, c2, c3;
      decode(item, nb, c1, c2, c3);
      const uint32_t buf = it & 1u, use = it >> 1;
      mbar_wait(accf(buf), use & 1u);
      tc_fence_after_sync();
      for (int j = 0; j < nchunks; ++j, ++jc) {
        const uint32_t t_addr = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * static_cast<uint32_t>(g.NB) + static_cast<uint32_t>(j) * 32u;
        uint32_t v[2][16];
        tmem_ld16(t_addr, v[0]);
        tmem_ld16(t_addr + 16, v[1]);
        tmem_ld_wait();
        if (j == nchunks - 1) {   // the accumulator is in registers: hand the buffer back to the MMA warp
          tc_fence_before_sync();
          mbar_arrive(acce(buf));
        }
        const int n0 = nb * g.NB + j * 32;
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(v[i >> 4][i & 15]);
        if (bias) {   // bias_mod % 16 == 0 and n0 % 32 == 0: a group of four columns never straddles the wrap
          int bi = n0 % g.bias_mod;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + bi);
            y[4 * jj] += b4.x;
            y[4 * jj + 1] += b4.y;
            y[4 * jj + 2] += b4.z;
            y[4 * jj + 3] += b4.w;
            bi += 4;
            bi = bi >= g.bias_mod ? bi - g.bias_mod : bi;
          }
        }
        // staging buffer jc & 1: its previous TMA store (two chunks ago) must have finished reading
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        uint8_t* sb = sm + out_off + (jc & 1u) * kK2TcOutBuf + static_cast<uint32_t>(r) * 128u;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          *reinterpret_cast<float4*>(sb + ((static_cast<uint32_t>(jj) ^ sw128) << 4)) =
              make_float4(y[4 * jj], y[4 * jj + 1], y[4 * jj + 2], y[4 * jj + 3]);
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (issuer) {
          const int jg = n0 / 32, mi = jg / g.out_cpm, c0 = (jg % g.out_cpm) * 32;
          const TmaDesc* om = mi == 0 ? &out0 : mi == 1 ? &out1 : mi == 2 ? &out2 : &out3;
          const uint32_t src = sm_addr + out_off + (jc & 1u) * kK2TcOutBuf;
          if (g.accumulate) tma_reduce_add_4d(om, src, c0, c1, c2, c3);
          else tma_store_4d(om, src, c0, c1, c2, c3);
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, static_cast<uint32_t>(g.tmem_cols));
}

// ---------------------------------------------------------------------------------------------
// weight images: w [8 CF][CC] fp32 (TF layout [2][2][2][CF][CC]; the transposed convolution's [2][2][2][out][in] is the
// same array) -> bf16 (hi, mid, lo) B operands, K-major rows of 32 elements:
//   gather  image  [kc < 8CF/32][hi|mid|lo][cc][kk]     = w[kc*32 + kk][cc]
//   scatter image  [kc < CC/32][hi|mid|lo][n < 8CF][kk] = w[n][kc*32 + kk]
// ---------------------------------------------------------------------------------------------
struct K2PackJob {
  const float* w;
  uint16_t* img_gather;
  uint16_t* img_scatter;
  int CF, CC, first_block, n_blocks;
};

__global__ void __launch_bounds__(256) k2tc_pack_multi_kernel(const K2PackJob* __restrict__ jobs, int njobs) {
  int ji = 0;
  while (ji + 1 < njobs && static_cast<int>(blockIdx.x) >= jobs[ji + 1].first_block) ++ji;
  const K2PackJob j = jobs[ji];
  const int rows = 8 * j.CF, total = rows * j.CC;
  for (int e = (static_cast<int>(blockIdx.x) - j.first_block) * 256 + static_cast<int>(threadIdx.x); e < total; e += j.n_blocks * 256) {
    const int row = e / j.CC, col = e % j.CC;
    float v = j.w[e];
    uint16_t pc[3];
    for (int k = 0; k < 3; ++k) {   // exact three-piece split
      pc[k] = f32_to_bf16(v);
      v -= bf16_to_f32(pc[k]);
    }
    if (j.img_gather) {
      const size_t o = (static_cast<size_t>(row / 32) * 3 * j.CC + col) * 32 + row % 32;
      for (int k = 0; k < 3; ++k) j.img_gather[o + static_cast<size_t>(k) * j.CC * 32] = pc[k];
    }
    if (j.img_scatter) {
      const size_t o = (static_cast<size_t>(col / 32) * 3 * rows + row) * 32 + col % 32;
      for (int k = 0; k < 3; ++k) j.img_scatter[o + static_cast<size_t>(k) * rows * 32] = pc[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side: plans
// ---------------------------------------------------------------------------------------------
struct K2TcPlan {
  bool valid = false;
  bool scatter = false;
  K2TcGeom g{};
  sm100::TmaDesc in[4], out[4], b;
  uint16_t* img = nullptr;   // weight image of this direction (8 CF CC x 3 bf16)
  size_t smem = 0;
  int Dc = 0;
  // what the tensor maps were encoded for: a launch with another batch size re-encodes them, so that the (n, od) axis
  // ends at the batch and rows beyond it are out of range (zero-filled loads, clipped stores)
  int enc_N = 0, CF = 0, CC = 0;
  Dims cd{};
  const float* fine = nullptr;
  const float* coarse = nullptr;
};

inline bool k2tc_enabled() { return getenv("VNB_K2_NO_TC") == nullptr; }

// tile shape: powers of two with product 128 that waste the fewest rows on this grid (ties: the widest ow_t)
inline void k2tc_tile_shape(int Wc, int Hc, long long ND, int& ow_t, int& oh_t, int& od_t) {
  double best = 1e300;
  for (int w = 128; w >= 1; w >>= 1)
    for (int h = 128 / w; h >= 1; h >>= 1) {
      const int d = 128 / (w * h);
      auto pad = [](long long v, int t) { return static_cast<double>((v + t - 1) / t * t); };
      const double vol = pad(Wc, w) * pad(Hc, h) * pad(ND, d);
      if (vol < best * (1.0 - 1e-12)) {
        best = vol;
        ow_t = w;
        oh_t = h;
        od_t = d;
      }
    }
}

// geometry for coarse dims (Dc, Hc, Wc), batch N (the tensor maps are encoded for the plan's N; a launch may use fewer)
inline bool k2tc_plan_geometry(K2TcPlan& pl, bool scatter, int N, Dims cd, int CF, int CC) {
  if (CF <= 0 || CC <= 0 || CF % 16 != 0 || CC % 32 != 0) return false;
  K2TcGeom& g = pl.g;
  g = K2TcGeom{};
  pl.scatter = scatter;
  pl.Dc = cd.D;
  k2tc_tile_shape(cd.W, cd.H, static_cast<long long>(N) * cd.D, g.ow_t, g.oh_t, g.od_t);
  g.n_tw = (cd.W + g.ow_t - 1) / g.ow_t;
  g.n_th = (cd.H + g.oh_t - 1) / g.oh_t;
  g.n_td = static_cast<int>((static_cast<long long>(N) * cd.D + g.od_t - 1) / g.od_t);
  if (!scatter) {
    if (CC > 256) return false;
    g.n_kc = CF / 4;        // 8 CF / 32
    g.in_cpm = CF / 16;     // 2 CF / 32 chunks per (kd, kh) map
    g.Ntot = CC;
    g.NB = CC;
    // few tiles (the deepest level: 8 tiles of 1024 coarse voxels): narrower column blocks give more, shorter items
    // (the A chunks are then loaded and split once per block; everything sits in L2 at that size)
    while (g.NB > 64 && g.NB % 64 == 0 && static_cast<long long>(g.n_tw) * g.n_th * g.n_td * (CC / g.NB) * 4 <= 148) g.NB /= 2;
    g.out_cpm = CC / 32;    // every output chunk through map 0
    g.bias_mod = CC;
  } else {
    g.n_kc = CC / 32;
    g.in_cpm = g.n_kc;      // one input map
    g.Ntot = 8 * CF;
    g.NB = g.Ntot % 256 == 0 ? 256 : 128;
    g.out_cpm = CF / 16;
    g.bias_mod = CF;
  }
  g.n_nb = g.Ntot / g.NB;
  g.stage_bytes = ((kK2TcA32 + 3 * kK2TcA16 + 3 * g.NB * 64 + 1023) / 1024) * 1024;
  if (g.bias_mod > 256 || g.bias_mod % 16 != 0) return false;
  const int fixed = 2 * kK2TcOutBuf + kK2TcBarBytes + kK2TcBiasBytes + 1024;
  g.stages = std::min(kK2TcMaxStages, (227 * 1024 - fixed) / g.stage_bytes);
  if (g.stages < 2) return false;
  pl.smem = static_cast<size_t>(fixed) + static_cast<size_t>(g.stages) * g.stage_bytes;
  g.tmem_cols = 32;
  while (g.tmem_cols < 2 * g.NB) g.tmem_cols *= 2;
  g.n_items = g.n_tw * g.n_th * g.n_td * g.n_nb;
  return true;
}

// fine tensor [N][2Dc][2Hc][2Wc][CF] -> the (kd, kh) map: ((kw,cf), ow, oh, (n,od)), box (32, ow_t, oh_t, od_t)
inline void k2tc_encode_fine(sm100::TmaDesc* out, const float* fine, int N, Dims cd, int CF, int kd, int kh, const K2TcGeom& g) {
  const uint64_t Wf = 2ull * cd.W, Hf = 2ull * cd.H;
  const float* base = fine + (static_cast<uint64_t>(kd) * Hf + kh) * Wf * CF;
  const uint64_t dims[4] = {2ull * CF, (uint64_t)cd.W, (uint64_t)cd.H, (uint64_t)N * cd.D};
  const uint64_t str[3] = {2ull * CF * 4, 2ull * Wf * CF * 4, 2ull * Hf * Wf * CF * 4};
  const uint32_t box[4] = {32, (uint32_t)g.ow_t, (uint32_t)g.oh_t, (uint32_t)g.od_t};
  tma_encode(out, base, 4, dims, str, box, 128, 4);
}
inline void k2tc_encode_coarse(sm100::TmaDesc* out, const float* coarse, int N, Dims cd, int CC, const K2TcGeom& g) {
  const uint64_t dims[4] = {(uint64_t)CC, (uint64_t)cd.W, (uint64_t)cd.H, (uint64_t)N * cd.D};
  const uint64_t str[3] = {(uint64_t)CC * 4, (uint64_t)cd.W * CC * 4, (uint64_t)cd.H * cd.W * CC * 4};
  const uint32_t box[4] = {32, (uint32_t)g.ow_t, (uint32_t)g.oh_t, (uint32_t)g.od_t};
  tma_encode(out, coarse, 4, dims, str, box, 128, 4);
}

// tensor maps of a plan: gather reads `fine`, writes `coarse`; scatter reads `coarse`, writes `fine`
inline void k2tc_encode_plan(K2TcPlan& pl, int N, Dims cd, int CF, int CC, const float* fine, const float* coarse) {
  pl.enc_N = N;
  pl.cd = cd;
  pl.CF = CF;
  pl.CC = CC;
  pl.fine = fine;
  pl.coarse = coarse;
  sm100::TmaDesc* f = pl.scatter ? pl.out : pl.in;
  sm100::TmaDesc* c = pl.scatter ? pl.in : pl.out;
  for (int t = 0; t < 4; ++t) k2tc_encode_fine(&f[t], fine, N, cd, CF, t >> 1, t & 1, pl.g);
  k2tc_encode_coarse(&c[0], coarse, N, cd, CC, pl.g);
  c[1] = c[2] = c[3] = c[0];
  tma_encode_w(&pl.b, pl.img, static_cast<long long>(pl.g.n_kc) * 3 * pl.g.Ntot, 32, pl.g.NB);
}

// launch for a batch of N (<= the plan's); `accumulate`: add into the destination (TMA reduce-add)
inline void k2tc_launch(K2TcPlan& pl, int N, const float* bias, bool accumulate, int sms, cudaStream_t stream) {
  if (N != pl.enc_N) k2tc_encode_plan(pl, N, pl.cd, pl.CF, pl.CC, pl.fine, pl.coarse);
  auto kfn = k2_tc_kernel;
#ifndef VNB_EMULATE
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      throw std::runtime_error("CUDA: cannot reserve shared memory for k2_tc_kernel");
    attr = true;
  }
#endif
  K2TcGeom g = pl.g;
  g.n_td = static_cast<int>((static_cast<long long>(N) * pl.Dc + g.od_t - 1) / g.od_t);
  g.n_items = g.n_tw * g.n_th * g.n_td * g.n_nb;
  g.accumulate = accumulate ? 1 : 0;
  const int grid = std::max(1, std::min(g.n_items, sms));
  VNB_LAUNCH(kfn, grid, kK2TcThreads, pl.smem, stream, pl.in[0], pl.in[1], pl.in[2], pl.in[3], pl.out[0], pl.out[1], pl.out[2],
             pl.out[3], pl.b, g, bias);
}

// ---------------------------------------------------------------------------------------------
// Filter gradient of both layers on the tensor cores:
//   dw[(kd,kh,kw,cf)][cc] = sum_m fine[child(m; kd,kh,kw)][cf] * coarse[m][cc]
// M = 128 filter rows r = (kd,kh,kw,cf) (one "M block": four 32-row chunks, each a (kd,kh) box column range), N = CC,
// K = the voxels.  Both operands are MN-major exactly as they sit in memory: a TMA box [128 voxels][32 floats] lands in a
// 16 KB slot, the converter warpgroup rewrites the slot IN PLACE as two [128 voxels][32 bf16] SWIZZLE_64B planes (hi | lo;
// every thread reads its row, one named barrier, every thread writes), and such a plane is one MN-major SWIZZLE_64B atom
// column (32 MN elements x 128 K rows): the four A slots / the CC/32 B slots of a set are the atoms at LBO = 16 KB, a
// K = 16 step advances the start address by two 8-row groups (1 KB).  A CTA owns one M block and a contiguous range of
// voxel tiles, accumulates them all into one TMEM accumulator (24 MMAs per tile: 8 K steps x 3 split passes) and adds
// the [128][min(CC, 64)] result to dw with fp32 atomics (wider layers: one CTA group per 64-column block) (dw is zeroed by the caller; like the mma.sync kernel it replaces this
// is the step's only run-to-run non-determinism, last-bit level).
// ---------------------------------------------------------------------------------------------
constexpr int kK2WgSlot = 16384;
constexpr int kK2WgMaxSets = 3;

struct K2WgGeom {
  int ow_t, oh_t, od_t, n_tw, n_th, n_td;
  int n_tiles, n_mb, splits;     // voxel tiles, M blocks (8 CF / 128), tile ranges per (M block, column block)
  int in_cpm;                    // 32-row chunks per (kd,kh) map = CF / 16
  int CC, NBw, n_cb, n_bc;       // coarse channels; GEMM N per CTA = min(CC, 64); column blocks; B chunks per CTA = NBw / 32
  int sets, set_bytes, tmem_cols;
};

__global__ void __launch_bounds__(kK2TcThreads, 1)
k2_wgrad_tc_kernel(const __grid_constant__ sm100::TmaDesc f0, const __grid_constant__ sm100::TmaDesc f1,
                   const __grid_constant__ sm100::TmaDesc f2, const __grid_constant__ sm100::TmaDesc f3,
                   const __grid_constant__ sm100::TmaDesc cmap, const K2WgGeom g, float* __restrict__ dw) {
  using namespace sm100;
  VNB_DYN_SMEM(uint8_t, smem_raw);
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t sm_addr = smem_u32(sm);
  const uint32_t bar_off = static_cast<uint32_t>(g.sets) * g.set_bytes;
  const uint32_t bar_base = sm_addr + bar_off;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto conv = [&](int s) { return bar_base + 8u * (4 + s); };
  auto empty = [&](int s) { return bar_base + 8u * (8 + s); };
  const uint32_t accf = bar_base + 8u * 12;
  const uint32_t slot_addr = bar_base + 8u * 13;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + bar_off + 8 * 13);

  const int tid = threadIdx.x;
  const int warp = static_cast<int>(warp_uniform(static_cast<uint32_t>(tid >> 5)));
  if (tid == 0) {
    for (int s = 0; s < kK2WgMaxSets; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), 128);
      mbar_init(empty(s), 1);
    }
    mbar_init(accf, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(slot_addr, static_cast<uint32_t>(g.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = warp_uniform(*slot_ptr);

  const int n_pairs = g.n_mb * g.n_cb;
  const int mb = (static_cast<int>(blockIdx.x) % n_pairs) % g.n_mb, cb = (static_cast<int>(blockIdx.x) % n_pairs) / g.n_mb;
  const int sp = static_cast<int>(blockIdx.x) / n_pairs;
  const int t_lo = static_cast<int>(static_cast<long long>(sp) * g.n_tiles / g.splits);
  const int t_hi = static_cast<int>(static_cast<long long>(sp + 1) * g.n_tiles / g.splits);
  const int n_slots = 4 + g.n_bc;

  if (warp == 0) {
    // ======================= TMA producer: per tile four fine boxes (this M block's chunks) + the coarse boxes =======================
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int t = t_lo; t < t_hi; ++t) {
      int x = t;
      const int c1 = (x % g.n_tw) * g.ow_t;
      x /= g.n_tw;
      const int c2 = (x % g.n_th) * g.oh_t;
      const int c3 = (x / g.n_th) * g.od_t;
      mbar_wait_warp(empty(s), ph ^ 1u);
      if (leader) {
        const uint32_t st = sm_addr + static_cast<uint32_t>(s) * g.set_bytes;
        mbar_expect_tx(full(s), static_cast<uint32_t>(n_slots) * kK2WgSlot);
        for (int c = 0; c < 4; ++c) {
          const int rc = mb * 4 + c, mi = rc / g.in_cpm, c0 = (rc % g.in_cpm) * 32;
          const TmaDesc* fm = mi == 0 ? &f0 : mi == 1 ? &f1 : mi == 2 ? &f2 : &f3;
          tma_load_4d(st + c * kK2WgSlot, fm, full(s), c0, c1, c2, c3);
        }
        for (int j = 0; j < g.n_bc; ++j) tma_load_4d(st + (4 + j) * kK2WgSlot, &cmap, full(s), cb * g.NBw + j * 32, c1, c2, c3);
      }
      __syncwarp();
      if (++s == g.sets) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const bool leader = elect_one();
    const uint32_t idesc = make_instr_desc(128, static_cast<uint32_t>(g.NBw), FMT_BF16, 1, 1);
    const uint64_t desc0 = make_smem_desc(0, kK2WgSlot, 512, SWZ_64B);   // MN-major: atoms at LBO, 8-row K groups at SBO
    int s = 0;
    uint32_t ph = 0;
    for (int t = t_lo; t < t_hi; ++t) {
      mbar_wait_warp(conv(s), ph);
      tc_fence_after_sync();
      const uint32_t a0 = sm_addr + static_cast<uint32_t>(s) * g.set_bytes, b0 = a0 + 4 * kK2WgSlot;
      const uint64_t dah = desc0 + (a0 >> 4), dal = dah + (kK2TcA16 >> 4);
      const uint64_t dbh = desc0 + (b0 >> 4), dbl = dbh + (kK2TcA16 >> 4);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t o = static_cast<uint64_t>(ks * 64);   // 16 K rows = 1024 bytes
        mma_f16_ss_if(leader, tmem, dal + o, dbh + o, idesc, (t != t_lo || ks != 0) ? 1u : 0u);
        mma_f16_ss_if(leader, tmem, dah + o, dbl + o, idesc, 1u);
        mma_f16_ss_if(leader, tmem, dah + o, dbh + o, idesc, 1u);
      }
      mma_commit_if(leader, empty(s));
      __syncwarp();
      if (++s == g.sets) {
        s = 0;
        ph ^= 1u;
      }
    }
    mma_commit_if(leader, accf);
    __syncwarp();
  } else if (warp < 6) {
    // ======================= converters: every slot in place, fp32 rows -> (hi | lo) bf16 planes =======================
    const int r = tid - 64;
    const uint32_t sw128 = static_cast<uint32_t>(r & 7), sw64 = static_cast<uint32_t>((r >> 1) & 3);
    const uint32_t src_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t dst_off = static_cast<uint32_t>(r >> 3) * 512u + static_cast<uint32_t>(r & 7) * 64u;
    int s = 0;
    uint32_t ph = 0;
    for (int t = t_lo; t < t_hi; ++t) {
      mbar_wait(full(s), ph);
      for (int c = 0; c < n_slots; ++c) {
        uint8_t* st = sm + static_cast<size_t>(s) * g.set_bytes + static_cast<size_t>(c) * kK2WgSlot;
        float4 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = *reinterpret_cast<const float4*>(st + src_off + ((static_cast<uint32_t>(j) ^ sw128) << 4));
        named_bar_sync(1, 128);   // every row of the slot is in registers before the planes overwrite it
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = x[2 * q], b = x[2 * q + 1];
          uint4 h, l;
          k2tc_split2(a.x, a.y, h.x, l.x);
          k2tc_split2(a.z, a.w, h.y, l.y);
          k2tc_split2(b.x, b.y, h.z, l.z);
          k2tc_split2(b.z, b.w, h.w, l.w);
          const uint32_t o = dst_off + ((static_cast<uint32_t>(q) ^ sw64) << 4);
          *reinterpret_cast<uint4*>(st + o) = h;
          *reinterpret_cast<uint4*>(st + o + kK2TcA16) = l;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(conv(s));
      if (++s == g.sets) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else if (t_hi > t_lo) {
    // ======================= epilogue: dw[(mb, row)][cc] += accumulator =======================
    const int lane = tid & 31, q = warp & 3;
    const int row = mb * 128 + q * 32 + lane;
    mbar_wait(accf, 0);
    tc_fence_after_sync();
    float* dst = dw + static_cast<size_t>(row) * g.CC + cb * g.NBw;
    for (int j = 0; j < g.NBw / 16; ++j) {
      uint32_t v[16];
      tmem_ld16(tmem + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j) * 16u, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) atomicAdd(dst + j * 16 + i, __uint_as_float(v[i]));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, static_cast<uint32_t>(g.tmem_cols));
}

struct K2WgPlan {
  bool valid = false;
  K2WgGeom g{};
  sm100::TmaDesc f[4], c;
  size_t smem = 0;
  int Dc = 0;
  int enc_N = 0, CF = 0, CC = 0;   // see K2TcPlan: the voxels beyond the batch must read as zeros here
  Dims cd{};
  const float* fine = nullptr;
  const float* coarse = nullptr;
};

inline bool k2wg_plan_geometry(K2WgPlan& pl, int N, Dims cd, int CF, int CC) {
  if (CF <= 0 || CC <= 0 || CF % 16 != 0 || CC % 32 != 0) return false;
  K2WgGeom& g = pl.g;
  g = K2WgGeom{};
  pl.Dc = cd.D;
  k2tc_tile_shape(cd.W, cd.H, static_cast<long long>(N) * cd.D, g.ow_t, g.oh_t, g.od_t);
  g.n_tw = (cd.W + g.ow_t - 1) / g.ow_t;
  g.n_th = (cd.H + g.oh_t - 1) / g.oh_t;
  g.n_mb = CF / 16;
  g.in_cpm = CF / 16;
  g.CC = CC;
  g.NBw = std::min(CC, 64);
  if (CC % g.NBw != 0) return false;
  g.n_cb = CC / g.NBw;
  g.n_bc = g.NBw / 32;
  g.set_bytes = (4 + g.n_bc) * kK2WgSlot;
  g.sets = std::min(kK2WgMaxSets, (227 * 1024 - kK2TcBarBytes - 1024) / g.set_bytes);
  if (g.sets < 2) return false;
  pl.smem = static_cast<size_t>(g.sets) * g.set_bytes + kK2TcBarBytes + 1024;
  g.tmem_cols = 32;
  while (g.tmem_cols < g.NBw) g.tmem_cols *= 2;
  return true;
}

inline void k2wg_encode_plan(K2WgPlan& pl, int N, Dims cd, int CF, int CC, const float* fine, const float* coarse) {
  pl.enc_N = N;
  pl.cd = cd;
  pl.CF = CF;
  pl.CC = CC;
  pl.fine = fine;
  pl.coarse = coarse;
  K2TcGeom tg{};
  tg.ow_t = pl.g.ow_t;
  tg.oh_t = pl.g.oh_t;
  tg.od_t = pl.g.od_t;
  for (int t = 0; t < 4; ++t) k2tc_encode_fine(&pl.f[t], fine, N, cd, CF, t >> 1, t & 1, tg);
  k2tc_encode_coarse(&pl.c, coarse, N, cd, CC, tg);
}

// dw must be zero (or hold what the result is added to)
inline void k2wg_launch(K2WgPlan& pl, int N, float* dw, int sms, cudaStream_t stream) {
  if (N != pl.enc_N) k2wg_encode_plan(pl, N, pl.cd, pl.CF, pl.CC, pl.fine, pl.coarse);
  auto kfn = k2_wgrad_tc_kernel;
#ifndef VNB_EMULATE
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      throw std::runtime_error("CUDA: cannot reserve shared memory for k2_wgrad_tc_kernel");
    attr = true;
  }
#endif
  K2WgGeom g = pl.g;
  g.n_td = static_cast<int>((static_cast<long long>(N) * pl.Dc + g.od_t - 1) / g.od_t);
  g.n_tiles = g.n_tw * g.n_th * g.n_td;
  const int n_pairs = g.n_mb * g.n_cb;
  g.splits = std::max(1, std::min(g.n_tiles, sms / n_pairs));
  VNB_LAUNCH(kfn, n_pairs * g.splits, kK2TcThreads, pl.smem, stream, pl.f[0], pl.f[1], pl.f[2], pl.f[3], pl.c, g, dw);
}

}  // namespace vnb
