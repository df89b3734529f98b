// Register-blocked fp32 tile kernels for the 2x2x2 stride-2 down / up convolutions
// (layers2.py:65-94 <- networks.py:278,292) and their gradients.  These layers are 0.7 % of the FLOPs
// but touch every activation once, so they are written as implicit GEMMs over the non-overlapping
// 2x2x2 blocks (no im2col buffer):
//   gather  C[m][cc]        = sum_{tap,cf} fine[child(m,tap)][cf] * w[tap][cf][cc]      (down fprop, up dgrad)
//   scatter fine[child][cf] = sum_cc coarse[m][cc] * w[tap][cf][cc]                     (up fprop, down dgrad)
//   wgrad   dw[tap][cf][cc] = sum_m fine[child(m,tap)][cf] * coarse[m][cc]              (both)
// 64x64 output tile per 256-thread block, 4x4 outputs per thread, K chunks of 16 through shared memory.
#pragma once
#include "conv_ref.cuh"

namespace vnb {

constexpr int kK2_BM = 64, kK2_BN = 64, kK2_BK = 16, kK2_PAD = 4;

__device__ __forceinline__ long long k2_child(const K2Args& p, long long m, int tap) {
  // m: flat coarse voxel index over [N][Dc][Hc][Wc]; returns flat fine voxel index
  long long o = m;
  const int ow = static_cast<int>(o % p.cd.W);
  o /= p.cd.W;
  const int oh = static_cast<int>(o % p.cd.H);
  o /= p.cd.H;
  const int od = static_cast<int>(o % p.cd.D);
  const long long n = o / p.cd.D;
  const int fd = 2 * od + (tap >> 2), fh = 2 * oh + ((tap >> 1) & 1), fw = 2 * ow + (tap & 1);
  return ((n * (2 * p.cd.D) + fd) * (2 * p.cd.H) + fh) * (2 * p.cd.W) + fw;
}

__device__ __forceinline__ void k2_tile_fma(const float (*As)[kK2_BM + kK2_PAD], const float (*Bs)[kK2_BN + kK2_PAD],
                                            int ty, int tx, float (&acc)[4][4]) {
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
      for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
  }
}

// grid: (ceil(M/64), ceil(CC/64)); requires CF % 16 == 0, CC % 4 == 0
// Software pipelined: the global loads of K chunk i+1 are in flight while chunk i is multiplied out of the other
// shared-memory buffer (one barrier per chunk).
__global__ void __launch_bounds__(256) k2_gather_tiled_kernel(K2Args p, long long M) {
  __shared__ float As[2][kK2_BK][kK2_BM + kK2_PAD];
  __shared__ float Bs[2][kK2_BK][kK2_BN + kK2_PAD];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const long long m0 = static_cast<long long>(blockIdx.x) * kK2_BM;
  const int n0 = blockIdx.y * kK2_BN;
  float acc[4][4] = {};
  const int arow = t / 4, akq = (t % 4) * 4;   // A loader: one float4 per thread
  const int bk = t / 16, bn4 = (t % 16) * 4;   // B loader
  const long long am = m0 + arow;
  const int chunks_per_tap = p.CF / kK2_BK, n_it = 8 * chunks_per_tap;
  float4 av, bv;
  auto fetch = [&](int it) {
    const int tap = it / chunks_per_tap, cf0 = (it % chunks_per_tap) * kK2_BK;
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (am < M) av = *reinterpret_cast<const float4*>(p.fine_in + k2_child(p, am, tap) * p.CF + cf0 + akq);
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
    k2_tile_fma(As[it & 1], Bs[it & 1], ty, tx, acc);
    if (it + 1 < n_it) stash((it + 1) & 1);
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= p.CC) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (p.bias) {
      const float4 b = *reinterpret_cast<const float4*>(p.bias + n);
      o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
    }
    float4* dst = reinterpret_cast<float4*>(p.coarse_out + m * p.CC + n);
    if (p.accumulate) {
      const float4 old = *dst;
      o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
    }
    *dst = o;
  }
}

// grid: (ceil(M/64), ceil(8*CF/64)); requires CC % 16 == 0, CF % 4 == 0
__global__ void __launch_bounds__(256) k2_scatter_tiled_kernel(K2Args p, long long M) {
  __shared__ float As[2][kK2_BK][kK2_BM + kK2_PAD];
  __shared__ float Bs[2][kK2_BK][kK2_BN + kK2_PAD];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const long long m0 = static_cast<long long>(blockIdx.x) * kK2_BM;
  const int n0 = blockIdx.y * kK2_BN;  // n = tap*CF + cf
  const int NN = 8 * p.CF;
  float acc[4][4] = {};
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
    k2_tile_fma(As[it & 1], Bs[it & 1], ty, tx, acc);
    if (it + 1 < n_it) stash((it + 1) & 1);
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= NN) return;
  const int tap = n / p.CF, cf = n % p.CF;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (p.bias) {
      const float4 b = *reinterpret_cast<const float4*>(p.bias + cf);
      o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
    }
    float4* dst = reinterpret_cast<float4*>(p.fine_out + k2_child(p, m, tap) * p.CF + cf);
    if (p.accumulate) {
      const float4 old = *dst;
      o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
    }
    *dst = o;
  }
}

// grid: (ceil(8*CF/64), ceil(CC/64), splits); fp32 atomics into pre-zeroed dw; CF % 4 == 0, CC % 4 == 0
__global__ void __launch_bounds__(256) k2_wgrad_tiled_kernel(K2Args p, long long M, long long m_per_split) {
  __shared__ float As[2][kK2_BK][kK2_BM + kK2_PAD];   // [k = voxel][r = (tap,cf) row]
  __shared__ float Bs[2][kK2_BK][kK2_BN + kK2_PAD];   // [k = voxel][n = cc]
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int r0 = blockIdx.x * kK2_BM, n0 = blockIdx.y * kK2_BN;
  const int RR = 8 * p.CF;
  const long long mb = static_cast<long long>(blockIdx.z) * m_per_split;
  const long long me = mb + m_per_split < M ? mb + m_per_split : M;
  float acc[4][4] = {};
  const int lk = t / 16, l4 = (t % 16) * 4;
  const int r = r0 + l4;
  const int tap = r < RR ? r / p.CF : 0, cf = r < RR ? r % p.CF : 0;
  const int n_it = static_cast<int>((me - mb + kK2_BK - 1) / kK2_BK);
  float4 av, bv;
  auto fetch = [&](int it) {
    const long long m = mb + static_cast<long long>(it) * kK2_BK + lk;
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < me) {
      if (r < RR) av = *reinterpret_cast<const float4*>(p.fine_in + k2_child(p, m, tap) * p.CF + cf);
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
    k2_tile_fma(As[it & 1], Bs[it & 1], ty, tx, acc);
    if (it + 1 < n_it) stash((it + 1) & 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = r0 + ty * 4 + i;
    if (rr >= RR) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.CC) atomicAdd(p.dw + static_cast<long long>(rr) * p.CC + n, acc[i][j]);
    }
  }
}

}  // namespace vnb
