// Register-blocked fp32 tile kernels for the 2x2x2 stride-2 down / up convolutions
// (layers2.py:65-94 <- networks.py:278,292) and their gradients.  These layers are 0.7 % of the FLOPs
// but touch every activation once, so they are written as implicit GEMMs over the non-overlapping
// 2x2x2 blocks (no im2col buffer):
//   gather  C[m][cc]        = sum_{tap,cf} fine[child(m,tap)][cf] * w[tap][cf][cc]      (down fprop, up dgrad)
//   scatter fine[child][cf] = sum_cc coarse[m][cc] * w[tap][cf][cc]                     (up fprop, down dgrad)
//   wgrad   dw[tap][cf][cc] = sum_m fine[child(m,tap)][cf] * coarse[m][cc]              (both)
// 64x64 output tile per 256-thread block, K chunks of 16 through shared memory.  Inner product, by template flag:
//   MMA = false : exact fp32 FMA, 4x4 outputs per thread (the fp32 parity path)
//   MMA = true  : warp-level mma.sync m16n8k8 TF32 on (big, small) splits of both operands, small*big + big*small +
//                 big*big with fp32 accumulate ("3xTF32": fp32-grade products, like the bf16x3 convolutions), fed by
//                 a 4-stage cp.async pipeline (k2_*_mma_kernel): the kernels are bound by the activation traffic
//                 instead of the FMA pipe / the latency of one load per thread.  Used by the tensor-core modes.
#pragma once
#include "conv_ref.cuh"

namespace vnb {

constexpr int kK2_BM = 64, kK2_BN = 64, kK2_BK = 16, kK2_PAD = 8;   // pitch 72: conflict-free MMA fragment reads

#ifndef VNB_EMULATE
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#else
// CPU model of the warp-level MMA (tests): lanes publish their fragments, rendezvous, then each lane forms its four
// outputs.  Fragment layout of mma.m16n8k8 (g = lane / 4, q = lane % 4):
//   a0 (g, q)  a1 (g+8, q)  a2 (g, q+4)  a3 (g+8, q+4);  b0 (k=q, n=g)  b1 (k=q+4, n=g);  c0,c1 (g, 2q+{0,1})  c2,c3 (g+8, ..)
inline void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  static uint32_t fa[64][32][4], fb[64][32][2];
  const int lin = emul::cur()->lin, w = (lin / 32) % 64, lane = lin & 31;
  for (int i = 0; i < 4; ++i) fa[w][lane][i] = a[i] & 0xFFFFE000u;   // TF32 keeps 10 mantissa bits
  for (int i = 0; i < 2; ++i) fb[w][lane][i] = b[i] & 0xFFFFE000u;
  __syncwarp();
  const int g = lane >> 2, q = lane & 3;
  auto A = [&](int row, int k) {
    const uint32_t u = fa[w][(row & 7) * 4 + (k & 3)][(row >> 3) + 2 * (k >> 2)];
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  auto B = [&](int k, int n) {
    const uint32_t u = fb[w][n * 4 + (k & 3)][k >> 2];
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  for (int h = 0; h < 2; ++h)
    for (int j = 0; j < 2; ++j) {
      float sum = c[2 * h + j];
      for (int k = 0; k < 8; ++k) sum += A(g + 8 * h, k) * B(k, 2 * q + j);
      c[2 * h + j] = sum;
    }
  __syncwarp();
}
#endif

// big / small TF32 split of an fp32 value: big keeps the top 11 significant bits, small = x - big is exact
__device__ __forceinline__ void tf32_split(float x, uint32_t& big, uint32_t& small) {
#if defined(__CUDA_ARCH__)
  big = __float_as_uint(x) & 0xFFFFE000u;
  small = __float_as_uint(x - __uint_as_float(big));
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  big = u & 0xFFFFE000u;
  float fb, fs;
  memcpy(&fb, &big, 4);
  fs = x - fb;
  memcpy(&small, &fs, 4);
#endif
}

// accumulators of one thread: FMA path acc[i][j] = (row ty*4+i, col tx*4+j); MMA path: warp tile 32 x 16 at
// (wm*32, wn*16), v[mi*8 + ni*4 + e] = MMA tile (mi, ni) element e
struct K2Acc {
  float v[16];
};

template <bool MMA>
__device__ __forceinline__ void k2_tile_product(const float (*As)[kK2_BM + kK2_PAD], const float (*Bs)[kK2_BN + kK2_PAD], int t,
                                                K2Acc& acc);

// Fine-grid index arithmetic.  A coarse voxel m = ((n*Dc + od)*Hc + oh)*Wc + ow owns the 2x2x2 fine block whose
// first voxel (tap 0) is k2_fine_base(m); tap = (a, b, c) adds k2_tap_offset.  Voxel counts fit 32 bits (the engine
// checks this when it builds the graph), so the decomposition uses 32-bit divisions and is hoisted out of the K loops.
__device__ __forceinline__ long long k2_fine_base(const K2Args& p, long long m) {
  unsigned o = static_cast<unsigned>(m);
  const unsigned W = p.cd.W, H = p.cd.H, D = p.cd.D;
  const unsigned ow = o % W;
  o /= W;
  const unsigned oh = o % H;
  o /= H;
  const unsigned od = o % D;
  const unsigned n = o / D;
  return ((static_cast<long long>(n) * (2 * D) + 2 * od) * (2 * H) + 2 * oh) * (2 * W) + 2 * ow;
}
__device__ __forceinline__ long long k2_tap_offset(const K2Args& p, int tap) {
  return (static_cast<long long>(tap >> 2) * (2 * p.cd.H) + ((tap >> 1) & 1)) * (2 * p.cd.W) + (tap & 1);
}
__device__ __forceinline__ long long k2_child(const K2Args& p, long long m, int tap) {
  return k2_fine_base(p, m) + k2_tap_offset(p, tap);
}

template <>
__device__ __forceinline__ void k2_tile_product<false>(const float (*As)[kK2_BM + kK2_PAD], const float (*Bs)[kK2_BN + kK2_PAD],
                                                       int t, K2Acc& acc) {
  const int tx = t % 16, ty = t / 16;
#pragma unroll
  for (int k = 0; k < kK2_BK; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc.v[i * 4 + j] += a[i] * b[j];
  }
}

template <>
__device__ __forceinline__ void k2_tile_product<true>(const float (*As)[kK2_BM + kK2_PAD], const float (*Bs)[kK2_BN + kK2_PAD],
                                                      int t, K2Acc& acc) {
  const int warp = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
#pragma unroll
  for (int ks = 0; ks < kK2_BK; ks += 8) {
    uint32_t ab[2][4], as_[2][4], bb[2][2], bs[2][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int m = wm * 32 + mi * 16 + g;
      tf32_split(As[ks + q][m], ab[mi][0], as_[mi][0]);
      tf32_split(As[ks + q][m + 8], ab[mi][1], as_[mi][1]);
      tf32_split(As[ks + q + 4][m], ab[mi][2], as_[mi][2]);
      tf32_split(As[ks + q + 4][m + 8], ab[mi][3], as_[mi][3]);
    }
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
      const int n = wn * 16 + ni * 8 + g;
      tf32_split(Bs[ks + q][n], bb[ni][0], bs[ni][0]);
      tf32_split(Bs[ks + q + 4][n], bb[ni][1], bs[ni][1]);
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        float(&c)[4] = *reinterpret_cast<float(*)[4]>(&acc.v[mi * 8 + ni * 4]);
        mma_tf32_16x8x8(c, as_[mi], bb[ni]);   // small terms first
        mma_tf32_16x8x8(c, ab[mi], bs[ni]);
        mma_tf32_16x8x8(c, ab[mi], bb[ni]);
      }
  }
}

// A thread's outputs are 4 rows x 2 pairs of horizontally adjacent columns (both inner-product paths):
// k2_thread_rows_cols gives the tile-local rows / first columns, k2_pair(acc, ri, ci, v0, v1) the values.
template <bool MMA>
__device__ __forceinline__ void k2_thread_rows_cols(int t, int (&rows)[4], int (&cols)[2]) {
  if (MMA) {
    const int warp = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) rows[ri] = wm * 32 + (ri >> 1) * 16 + g + 8 * (ri & 1);
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) cols[ci] = wn * 16 + ci * 8 + 2 * q;
  } else {
    const int tx = t % 16, ty = t / 16;
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) rows[ri] = ty * 4 + ri;
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) cols[ci] = tx * 4 + 2 * ci;
  }
}
template <bool MMA>
__device__ __forceinline__ void k2_pair(const K2Acc& acc, int ri, int ci, float& v0, float& v1) {
  const int i = MMA ? (ri >> 1) * 8 + ci * 4 + 2 * (ri & 1) : ri * 4 + 2 * ci;
  v0 = acc.v[i];
  v1 = acc.v[i + 1];
}
template <bool MMA, class F>
__device__ __forceinline__ void k2_for_each_pair(const K2Acc& acc, int t, F&& f) {
  int rows[4], cols[2];
  k2_thread_rows_cols<MMA>(t, rows, cols);
#pragma unroll
  for (int ri = 0; ri < 4; ++ri)
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      float v0, v1;
      k2_pair<MMA>(acc, ri, ci, v0, v1);
      f(rows[ri], cols[ci], v0, v1);
    }
}

template <bool MMA>
__device__ __forceinline__ void k2_gather_epilogue(const K2Args& p, long long M, long long m0, int n0, int t, const K2Acc& acc) {
  k2_for_each_pair<MMA>(acc, t, [&](int row, int col, float v0, float v1) {
    const long long m = m0 + row;
    const int n = n0 + col;
    if (m >= M || n >= p.CC) return;
    float2 o = make_float2(v0, v1);
    if (p.bias) {
      o.x += p.bias[n];
      o.y += p.bias[n + 1];
    }
    float2* dst = reinterpret_cast<float2*>(p.coarse_out + m * p.CC + n);
    if (p.accumulate) {
      const float2 old = *dst;
      o.x += old.x;
      o.y += old.y;
    }
    *dst = o;
  });
}

// the fine-grid index is split into a per-row base and a per-column tap offset, each computed once
template <bool MMA>
__device__ __forceinline__ void k2_scatter_epilogue(const K2Args& p, long long M, long long m0, int n0, int t, const K2Acc& acc) {
  const int NN = 8 * p.CF;
  int rows[4], cols[2];
  k2_thread_rows_cols<MMA>(t, rows, cols);
  long long fbase[4], toff[2];
  int cfs[2];
#pragma unroll
  for (int ri = 0; ri < 4; ++ri) fbase[ri] = m0 + rows[ri] < M ? k2_fine_base(p, m0 + rows[ri]) : -1;
#pragma unroll
  for (int ci = 0; ci < 2; ++ci) {
    const int n = n0 + cols[ci];
    cfs[ci] = -1;
    toff[ci] = 0;
    if (n < NN) {   // CF % 4 == 0: a pair never straddles two taps
      cfs[ci] = n % p.CF;
      toff[ci] = k2_tap_offset(p, n / p.CF);
    }
  }
#pragma unroll
  for (int ri = 0; ri < 4; ++ri)
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      if (fbase[ri] < 0 || cfs[ci] < 0) continue;
      float2 o;
      k2_pair<MMA>(acc, ri, ci, o.x, o.y);
      if (p.bias) {
        o.x += p.bias[cfs[ci]];
        o.y += p.bias[cfs[ci] + 1];
      }
      float2* dst = reinterpret_cast<float2*>(p.fine_out + (fbase[ri] + toff[ci]) * p.CF + cfs[ci]);
      if (p.accumulate) {
        const float2 old = *dst;
        o.x += old.x;
        o.y += old.y;
      }
      *dst = o;
    }
}

template <bool MMA>
__device__ __forceinline__ void k2_wgrad_epilogue(const K2Args& p, int r0, int n0, int t, const K2Acc& acc) {
  const int RR = 8 * p.CF;
  k2_for_each_pair<MMA>(acc, t, [&](int row, int col, float v0, float v1) {
    const int rr = r0 + row, n = n0 + col;
    if (rr >= RR || n >= p.CC) return;   // CC % 4 == 0: n + 1 < CC as well
    atomicAdd(p.dw + static_cast<long long>(rr) * p.CC + n, v0);
    atomicAdd(p.dw + static_cast<long long>(rr) * p.CC + n + 1, v1);
  });
}

// grid: (ceil(M/64), ceil(CC/64)); requires CF % 16 == 0, CC % 4 == 0
// Software pipelined: the global loads of K chunk i+1 are in flight while chunk i is multiplied out of the other
// shared-memory buffer (one barrier per chunk).
template <bool MMA>
__global__ void __launch_bounds__(256) k2_gather_tiled_kernel(K2Args p, long long M) {
  __shared__ float As[2][kK2_BK][kK2_BM + kK2_PAD];
  __shared__ float Bs[2][kK2_BK][kK2_BN + kK2_PAD];
  const int t = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * kK2_BM;
  const int n0 = blockIdx.y * kK2_BN;
  K2Acc acc = {};
  const int arow = t / 4, akq = (t % 4) * 4;   // A loader: one float4 per thread
  const int bk = t / 16, bn4 = (t % 16) * 4;   // B loader
  const long long am = m0 + arow;
  const long long abase = am < M ? k2_fine_base(p, am) : 0;
  const int chunks_per_tap = p.CF / kK2_BK, n_it = 8 * chunks_per_tap;
  float4 av, bv;
  auto fetch = [&](int it) {
    const int tap = it / chunks_per_tap, cf0 = (it % chunks_per_tap) * kK2_BK;
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (am < M) av = *reinterpret_cast<const float4*>(p.fine_in + (abase + k2_tap_offset(p, tap)) * p.CF + cf0 + akq);
    if (n0 + bn4 < p.CC) bv = *reinterpret_cast<const float4*>(p.w + (static_cast<long long>(tap) * p.CF + cf0 + bk) * p.CC + n0 + bn4);
  };
  auto stash = [&](int buf) {
    As[buf][akq + 0][arow] = av.x;
    As[buf][akq + 1][arow] = av.y;
    As[buf][akq + 2][arow] = av.z;
    As[buf][akq + 3][arow] = av.w;
    *reinterpret_cast<float4*>(&Bs[buf][bk][bn4]) = bv;
  };
  fetch(0);
  stash(0);
  __syncthreads();
  for (int it = 0; it < n_it; ++it) {
    if (it + 1 < n_it) fetch(it + 1);
    k2_tile_product<MMA>(As[it & 1], Bs[it & 1], t, acc);
    if (it + 1 < n_it) stash((it + 1) & 1);
    __syncthreads();
  }
  k2_gather_epilogue<MMA>(p, M, m0, n0, t, acc);
}

// grid: (ceil(M/64), ceil(8*CF/64)); requires CC % 16 == 0, CF % 4 == 0
template <bool MMA>
__global__ void __launch_bounds__(256) k2_scatter_tiled_kernel(K2Args p, long long M) {
  __shared__ float As[2][kK2_BK][kK2_BM + kK2_PAD];
  __shared__ float Bs[2][kK2_BK][kK2_BN + kK2_PAD];
  const int t = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * kK2_BM;
  const int n0 = blockIdx.y * kK2_BN;  // n = tap*CF + cf
  const int NN = 8 * p.CF;
  K2Acc acc = {};
  const int row = t / 4, kq = (t % 4) * 4;
  const int n_it = p.CC / kK2_BK;
  float4 av, bv;
  auto fetch = [&](int it) {
    const int cc0 = it * kK2_BK;
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + row < M) av = *reinterpret_cast<const float4*>(p.coarse_in + (m0 + row) * p.CC + cc0 + kq);
    if (n0 + row < NN) bv = *reinterpret_cast<const float4*>(p.w + static_cast<long long>(n0 + row) * p.CC + cc0 + kq);
  };
  auto stash = [&](int buf) {
    As[buf][kq + 0][row] = av.x; As[buf][kq + 1][row] = av.y; As[buf][kq + 2][row] = av.z; As[buf][kq + 3][row] = av.w;
    Bs[buf][kq + 0][row] = bv.x; Bs[buf][kq + 1][row] = bv.y; Bs[buf][kq + 2][row] = bv.z; Bs[buf][kq + 3][row] = bv.w;
  };
  fetch(0);
  stash(0);
  __syncthreads();
  for (int it = 0; it < n_it; ++it) {
    if (it + 1 < n_it) fetch(it + 1);
    k2_tile_product<MMA>(As[it & 1], Bs[it & 1], t, acc);
    if (it + 1 < n_it) stash((it + 1) & 1);
    __syncthreads();
  }
  k2_scatter_epilogue<MMA>(p, M, m0, n0, t, acc);
}

// grid: (ceil(8*CF/64), ceil(CC/64), splits); fp32 atomics into pre-zeroed dw; CF % 4 == 0, CC % 4 == 0
template <bool MMA>
__global__ void __launch_bounds__(256) k2_wgrad_tiled_kernel(K2Args p, long long M, long long m_per_split) {
  __shared__ float As[2][kK2_BK][kK2_BM + kK2_PAD];   // [k = voxel][r = (tap,cf) row]
  __shared__ float Bs[2][kK2_BK][kK2_BN + kK2_PAD];   // [k = voxel][n = cc]
  const int t = threadIdx.x;
  const int r0 = blockIdx.x * kK2_BM, n0 = blockIdx.y * kK2_BN;
  const int RR = 8 * p.CF;
  const long long mb = static_cast<long long>(blockIdx.z) * m_per_split;
  const long long me = mb + m_per_split < M ? mb + m_per_split : M;
  K2Acc acc = {};
  const int lk = t / 16, l4 = (t % 16) * 4;
  const int r = r0 + l4;
  const int tap = r < RR ? r / p.CF : 0, cf = r < RR ? r % p.CF : 0;
  const int n_it = static_cast<int>((me - mb + kK2_BK - 1) / kK2_BK);
  const long long tap_off = k2_tap_offset(p, tap);
  float4 av, bv;
  auto fetch = [&](int it) {
    const long long m = mb + static_cast<long long>(it) * kK2_BK + lk;
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < me) {
      if (r < RR) av = *reinterpret_cast<const float4*>(p.fine_in + (k2_fine_base(p, m) + tap_off) * p.CF + cf);
      if (n0 + l4 < p.CC) bv = *reinterpret_cast<const float4*>(p.coarse_in + m * p.CC + n0 + l4);
    }
  };
  auto stash = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lk][l4]) = av;
    *reinterpret_cast<float4*>(&Bs[buf][lk][l4]) = bv;
  };
  if (n_it > 0) {
    fetch(0);
    stash(0);
  }
  __syncthreads();
  for (int it = 0; it < n_it; ++it) {
    if (it + 1 < n_it) fetch(it + 1);
    k2_tile_product<MMA>(As[it & 1], Bs[it & 1], t, acc);
    if (it + 1 < n_it) stash((it + 1) & 1);
    __syncthreads();
  }
  k2_wgrad_epilogue<MMA>(p, r0, n0, t, acc);
}


// -------------------------------------------------------------------------------------------------
// Tensor-core forms: 4-stage cp.async pipeline (16-byte global -> shared copies, no register staging), 3xTF32
// mma.sync inner product.  Shared-memory tiles per stage, chosen so that both the copies and the fragment reads
// are conflict-free:
//   "row-major"  T[r][k], pitch 20 floats  (r = GEMM row or column, k contiguous: source rows are K-contiguous)
//   "k-major"    T[k][r], pitch 72 floats  (source rows are contiguous along the GEMM row / column)
// -------------------------------------------------------------------------------------------------
constexpr int kK2_STAGES = 4, kK2_PR = 20, kK2_PK = 72;
constexpr int kK2_TILE_FLOATS = 64 * kK2_PR > 16 * kK2_PK ? 64 * kK2_PR : 16 * kK2_PK;   // 1280

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool valid) {
#ifndef VNB_EMULATE
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int bytes = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
#else
  for (int i = 0; i < 4; ++i) smem_dst[i] = valid ? gmem_src[i] : 0.f;
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef VNB_EMULATE
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef VNB_EMULATE
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// element (row r, k) of an A / B tile in either layout
template <bool KMAJOR>
__device__ __forceinline__ float k2_tile_at(const float* T, int r, int k) {
  return KMAJOR ? T[k * kK2_PK + r] : T[r * kK2_PR + k];
}

// m_valid / n_valid: rows / columns of the 64 x 64 tile that exist; warps whose 32 x 16 sub-tile lies outside skip
template <bool A_KMAJOR, bool B_KMAJOR>
__device__ __forceinline__ void k2_tile_product_mma(const float* As, const float* Bs, int t, K2Acc& acc, int m_valid, int n_valid) {
  const int warp = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  if (wm * 32 >= m_valid || wn * 16 >= n_valid) return;
#pragma unroll
  for (int ks = 0; ks < kK2_BK; ks += 8) {
    uint32_t ab[2][4], as_[2][4], bb[2][2], bs[2][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int m = wm * 32 + mi * 16 + g;
      tf32_split(k2_tile_at<A_KMAJOR>(As, m, ks + q), ab[mi][0], as_[mi][0]);
      tf32_split(k2_tile_at<A_KMAJOR>(As, m + 8, ks + q), ab[mi][1], as_[mi][1]);
      tf32_split(k2_tile_at<A_KMAJOR>(As, m, ks + q + 4), ab[mi][2], as_[mi][2]);
      tf32_split(k2_tile_at<A_KMAJOR>(As, m + 8, ks + q + 4), ab[mi][3], as_[mi][3]);
    }
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
      const int n = wn * 16 + ni * 8 + g;
      tf32_split(k2_tile_at<B_KMAJOR>(Bs, n, ks + q), bb[ni][0], bs[ni][0]);
      tf32_split(k2_tile_at<B_KMAJOR>(Bs, n, ks + q + 4), bb[ni][1], bs[ni][1]);
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        float(&c)[4] = *reinterpret_cast<float(*)[4]>(&acc.v[mi * 8 + ni * 4]);
        mma_tf32_16x8x8(c, as_[mi], bb[ni]);   // small terms first
        mma_tf32_16x8x8(c, ab[mi], bs[ni]);
        mma_tf32_16x8x8(c, ab[mi], bb[ni]);
      }
  }
}

// generic pipeline driver: issue(it, As_stage, Bs_stage) enqueues the copies of K chunk `it`
template <bool A_KMAJOR, bool B_KMAJOR, class Issue>
__device__ __forceinline__ void k2_mma_mainloop(float* smem, int n_it, int t, K2Acc& acc, int m_valid, int n_valid, Issue&& issue) {
  auto As = [&](int s) { return smem + s * 2 * kK2_TILE_FLOATS; };
  auto Bs = [&](int s) { return smem + s * 2 * kK2_TILE_FLOATS + kK2_TILE_FLOATS; };
#pragma unroll
  for (int s = 0; s < kK2_STAGES - 1; ++s) {
    if (s < n_it) issue(s, As(s), Bs(s));
    cp_async_commit();
  }
  for (int it = 0; it < n_it; ++it) {
    cp_async_wait<kK2_STAGES - 2>();   // chunk `it` has landed (one group per chunk, committed in order)
    __syncthreads();                   // ... for every thread; and everyone is done reading chunk it-1
    const int nx = it + kK2_STAGES - 1;
    if (nx < n_it) issue(nx, As(nx % kK2_STAGES), Bs(nx % kK2_STAGES));   // refills the buffer of chunk it-1
    cp_async_commit();
    k2_tile_product_mma<A_KMAJOR, B_KMAJOR>(As(it % kK2_STAGES), Bs(it % kK2_STAGES), t, acc, m_valid, n_valid);
  }
  cp_async_wait<0>();
}

// grid: (ceil(M/64), ceil(CC/64)); CF % 16 == 0, CC % 4 == 0
__global__ void __launch_bounds__(256) k2_gather_mma_kernel(K2Args p, long long M) {
  __shared__ float4 smem4[kK2_STAGES * 2 * kK2_TILE_FLOATS / 4];   // float4: 16-byte alignment for cp.async
  float* smem = reinterpret_cast<float*>(smem4);
  const int t = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * kK2_BM;
  const int n0 = blockIdx.y * kK2_BN;
  K2Acc acc = {};
  const int arow = t / 4, akq = (t % 4) * 4;   // A: row-major [m][k], one 16-byte copy per thread
  const int bk = t / 16, bn4 = (t % 16) * 4;   // B: k-major [k][n]
  const long long am = m0 + arow;
  const bool a_ok = am < M, b_ok = n0 + bn4 < p.CC;
  const long long abase = a_ok ? k2_fine_base(p, am) : 0;
  const int chunks_per_tap = p.CF / kK2_BK;
  k2_mma_mainloop<false, true>(smem, 8 * chunks_per_tap, t, acc, 64, p.CC - n0, [&](int it, float* As, float* Bs) {
    const int tap = it / chunks_per_tap, cf0 = (it % chunks_per_tap) * kK2_BK;
    cp_async16(As + arow * kK2_PR + akq, p.fine_in + (a_ok ? (abase + k2_tap_offset(p, tap)) * p.CF + cf0 + akq : 0), a_ok);
    cp_async16(Bs + bk * kK2_PK + bn4, p.w + (b_ok ? (static_cast<long long>(tap) * p.CF + cf0 + bk) * p.CC + n0 + bn4 : 0), b_ok);
  });
  k2_gather_epilogue<true>(p, M, m0, n0, t, acc);
}

// grid: (ceil(M/64), ceil(8*CF/64)); CC % 16 == 0, CF % 4 == 0
__global__ void __launch_bounds__(256) k2_scatter_mma_kernel(K2Args p, long long M) {
  __shared__ float4 smem4[kK2_STAGES * 2 * kK2_TILE_FLOATS / 4];   // float4: 16-byte alignment for cp.async
  float* smem = reinterpret_cast<float*>(smem4);
  const int t = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * kK2_BM;
  const int n0 = blockIdx.y * kK2_BN;  // n = tap*CF + cf
  const int NN = 8 * p.CF;
  K2Acc acc = {};
  const int row = t / 4, kq = (t % 4) * 4;     // A [m][k] and B [n][k]: both row-major
  const bool a_ok = m0 + row < M, b_ok = n0 + row < NN;
  k2_mma_mainloop<false, false>(smem, p.CC / kK2_BK, t, acc, 64, NN - n0, [&](int it, float* As, float* Bs) {
    const int cc0 = it * kK2_BK;
    cp_async16(As + row * kK2_PR + kq, p.coarse_in + (a_ok ? (m0 + row) * p.CC + cc0 + kq : 0), a_ok);
    cp_async16(Bs + row * kK2_PR + kq, p.w + (b_ok ? static_cast<long long>(n0 + row) * p.CC + cc0 + kq : 0), b_ok);
  });
  k2_scatter_epilogue<true>(p, M, m0, n0, t, acc);
}

// grid: (ceil(8*CF/64), ceil(CC/64), splits); fp32 atomics into pre-zeroed dw; CF % 4 == 0, CC % 4 == 0
__global__ void __launch_bounds__(256) k2_wgrad_mma_kernel(K2Args p, long long M, long long m_per_split) {
  __shared__ float4 smem4[kK2_STAGES * 2 * kK2_TILE_FLOATS / 4];   // float4: 16-byte alignment for cp.async
  float* smem = reinterpret_cast<float*>(smem4);
  const int t = threadIdx.x;
  const int r0 = blockIdx.x * kK2_BM, n0 = blockIdx.y * kK2_BN;
  const int RR = 8 * p.CF;
  const long long mb = static_cast<long long>(blockIdx.z) * m_per_split;
  const long long me = mb + m_per_split < M ? mb + m_per_split : M;
  K2Acc acc = {};
  const int lk = t / 16, l4 = (t % 16) * 4;    // A [k = voxel][r] and B [k = voxel][n]: both k-major
  const int r = r0 + l4;
  const int tap = r < RR ? r / p.CF : 0, cf = r < RR ? r % p.CF : 0;
  const long long tap_off = k2_tap_offset(p, tap);
  const int n_it = static_cast<int>((me - mb + kK2_BK - 1) / kK2_BK);
  k2_mma_mainloop<true, true>(smem, n_it, t, acc, RR - r0, p.CC - n0, [&](int it, float* As, float* Bs) {
    const long long m = mb + static_cast<long long>(it) * kK2_BK + lk;
    const bool a_ok = m < me && r < RR, b_ok = m < me && n0 + l4 < p.CC;
    cp_async16(As + lk * kK2_PK + l4, p.fine_in + (a_ok ? (k2_fine_base(p, m) + tap_off) * p.CF + cf : 0), a_ok);
    cp_async16(Bs + lk * kK2_PK + l4, p.coarse_in + (b_ok ? m * p.CC + n0 + l4 : 0), b_ok);
  });
  k2_wgrad_epilogue<true>(p, r0, n0, t, acc);
}

}  // namespace vnb
