// fp32 CUDA-core convolution kernels ("exact" precision mode and the on-device anchor that the
// tcgen05 implicit-GEMM kernels are validated against at sizes the CPU oracle cannot reach).
//
// Reference ops restated (paths relative to /root/reference):
//   layers2.convolution 5x5x5 stride 1 SAME   layers2.py:59-63  <- networks.py:264,316,333,346,356
//   layers2.down_convolution 2x2x2 stride 2   layers2.py:78-84  <- networks.py:278
//   layers2.up_convolution (conv3d_transpose) layers2.py:65-74,88-94 <- networks.py:292
//   1x1x1 output convolution                  networks.py:302
// plus their input-gradient (dgrad) and filter-gradient (wgrad) forms (TF autodiff of model.py:660).
//
// Layout: activations NDHWC fp32; 5^3 filters [125][Cin][Cout] (= TF [kd,kh,kw,Cin,Cout]);
// 2^3 filters [8][Cfine][Ccoarse] (TF down: [2,2,2,Cin,Cout], TF up: [2,2,2,Cout,Cin] -- both have
// the fine-resolution channel first, so one set of kernels serves both directions).
#pragma once
#include "vnb_cuda.h"

namespace vnb {

struct Dims {  // spatial extents of one sample
  int D, H, W;
};

// -------------------------------------------------------------------------------------------------
// 5x5x5 stride-1 SAME convolution, forward form.  Input = channel concat of (in1, in2)
// (tf.concat at networks.py:325 is never materialised).  Output channels [0,Co1) go to out1 and
// [Co1, Co1+Co2) to out2 (used by dgrad of the concat convolution); each output may accumulate.
// dgrad is this kernel run on dz with flipped+transposed weights (see flip_transpose_w_kernel).
// -------------------------------------------------------------------------------------------------
constexpr int kC5_TD = 4, kC5_TH = 4, kC5_TW = 16;         // output tile, one voxel per thread
constexpr int kC5_CK = 8;                                   // input channels per smem chunk
constexpr int kC5_CO = 16;                                  // output channels per thread
template <int KS> struct ConvRefGeom {                      // KS = 5 (V-Net, layers2.py:59-63) or 3 (attention.py:63-81)
  static constexpr int R = KS / 2, TAPS = KS * KS * KS;
  static constexpr int HH = kC5_TH + KS - 1, HW = kC5_TW + KS - 1;
  static constexpr int HV = (kC5_TD + KS - 1) * HH * HW;  // halo voxels (1280 for KS = 5)
  static constexpr int XS = HV + 1;                         // padded channel stride (bank conflicts)
  static constexpr size_t SMEM = (static_cast<size_t>(kC5_CK) * XS + TAPS * kC5_CK * kC5_CO) * sizeof(float);
};
constexpr size_t kC5_SMEM = ConvRefGeom<5>::SMEM;

struct Conv5Args {
  const float* in1;
  const float* in2;   // may be nullptr
  int C1, C2;         // channels of in1 / in2
  const float* w;     // [KS^3][C1+C2][Cout]
  const float* bias;  // [Cout] or nullptr
  const float* res;   // residual [V][Cout] added in the epilogue, or nullptr
  float* out1;
  float* out2;        // may be nullptr
  int Co1, Co2;       // Cout = Co1 + Co2
  int acc1, acc2;     // accumulate into existing contents
  Dims dims;
  int N;
};

template <int KS>
__global__ void __launch_bounds__(256) conv_ref_kernel(Conv5Args p) {
  using G = ConvRefGeom<KS>;
  VNB_DYN_SMEM(float, smem);
  float* xs = smem;                       // [CK][XS]
  float* ws = smem + kC5_CK * G::XS;      // [TAPS][CK][CO]
  const int D = p.dims.D, H = p.dims.H, W = p.dims.W;
  const int tw_n = (W + kC5_TW - 1) / kC5_TW, th_n = (H + kC5_TH - 1) / kC5_TH;
  int tile = blockIdx.x;
  const int tw0 = (tile % tw_n) * kC5_TW;
  tile /= tw_n;
  const int th0 = (tile % th_n) * kC5_TH;
  const int td0 = (tile / th_n) * kC5_TD;
  const int co0 = blockIdx.y * kC5_CO;
  const int n = blockIdx.z;
  const int t = threadIdx.x;
  const int lw = t % kC5_TW, lh = (t / kC5_TW) % kC5_TH, ld = t / (kC5_TW * kC5_TH);
  const int Cin = p.C1 + p.C2, Cout = p.Co1 + p.Co2;
  const long long sample = static_cast<long long>(D) * H * W;

  float acc[kC5_CO];
#pragma unroll
  for (int j = 0; j < kC5_CO; ++j) acc[j] = 0.f;

  for (int cb = 0; cb < Cin; cb += kC5_CK) {
    __syncthreads();
    // halo tile: xs[ci][hv], zero outside the volume (SAME padding) and beyond Cin
    for (int i = t; i < G::HV * kC5_CK; i += 256) {
      const int ci = i % kC5_CK, hv = i / kC5_CK;
      const int hw = hv % G::HW, hh = (hv / G::HW) % G::HH, hd = hv / (G::HW * G::HH);
      const int gd = td0 + hd - G::R, gh = th0 + hh - G::R, gw = tw0 + hw - G::R, cg = cb + ci;
      float val = 0.f;
      if (cg < Cin && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
        const long long vox = n * sample + (static_cast<long long>(gd) * H + gh) * W + gw;
        val = cg < p.C1 ? p.in1[vox * p.C1 + cg] : p.in2[vox * p.C2 + (cg - p.C1)];
      }
      xs[ci * G::XS + hv] = val;
    }
    for (int i = t; i < G::TAPS * kC5_CK * kC5_CO; i += 256) {
      const int co = i % kC5_CO, ci = (i / kC5_CO) % kC5_CK, tap = i / (kC5_CO * kC5_CK);
      const int cg = cb + ci;
      ws[i] = (cg < Cin && co0 + co < Cout) ? p.w[(static_cast<long long>(tap) * Cin + cg) * Cout + co0 + co] : 0.f;
    }
    __syncthreads();
    for (int kd = 0; kd < KS; ++kd)
      for (int kh = 0; kh < KS; ++kh)
        for (int kw = 0; kw < KS; ++kw) {
          const int tap = (kd * KS + kh) * KS + kw;
          const int hv = ((ld + kd) * G::HH + (lh + kh)) * G::HW + lw + kw;
#pragma unroll
          for (int ci = 0; ci < kC5_CK; ++ci) {
            const float x = xs[ci * G::XS + hv];
            const float* wr = ws + (tap * kC5_CK + ci) * kC5_CO;
#pragma unroll
            for (int j = 0; j < kC5_CO; ++j) acc[j] += x * wr[j];
          }
        }
  }
  const int gd = td0 + ld, gh = th0 + lh, gw = tw0 + lw;
  if (gd >= D || gh >= H || gw >= W) return;
  const long long vox = n * sample + (static_cast<long long>(gd) * H + gh) * W + gw;
#pragma unroll
  for (int j = 0; j < kC5_CO; ++j) {
    const int co = co0 + j;
    if (co >= Cout) break;
    float y = acc[j];
    if (p.bias) y += p.bias[co];
    if (p.res) y += p.res[vox * Cout + co];
    if (co < p.Co1) {
      float* o = p.out1 + vox * p.Co1 + co;
      *o = p.acc1 ? *o + y : y;
    } else {
      float* o = p.out2 + vox * p.Co2 + (co - p.Co1);
      *o = p.acc2 ? *o + y : y;
    }
  }
}
#define conv5_ref_kernel conv_ref_kernel<5>

// wd[tap'][co][ci] = w[TAPS-1 - tap'][ci][co]: dgrad(dz) = conv(dz, wd) (flipped taps, swapped channels)
__global__ void flip_transpose_w_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cin, int Cout, int taps) {
  const long long total = static_cast<long long>(taps) * Cin * Cout;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    const int co = static_cast<int>((i / Cin) % Cout);
    const int tap = static_cast<int>(i / (static_cast<long long>(Cin) * Cout));
    wd[i] = w[(static_cast<long long>(taps - 1 - tap) * Cin + ci) * Cout + co];
  }
}

// -------------------------------------------------------------------------------------------------
// KS^3 filter gradient: dw[tap][ci][co] += sum_v x[v + off(tap)][ci] * dz[v][co]
// grid: (voxel splits, (Cin/16 blocks) * (Cout/16 blocks)); thread = one (ci, co) pair with all
// KS^3 taps in registers; block partial sums are added to dw with fp32 atomics (dw pre-zeroed).
// -------------------------------------------------------------------------------------------------
constexpr int kW5_TD = 2, kW5_TH = 4, kW5_TW = 16;
constexpr int kW5_TV = kW5_TD * kW5_TH * kW5_TW;                    // 128
template <int KS> struct WgradRefGeom {
  static constexpr int R = KS / 2, TAPS = KS * KS * KS;
  static constexpr int HH = kW5_TH + KS - 1, HW = kW5_TW + KS - 1;
  static constexpr int HV = (kW5_TD + KS - 1) * HH * HW;          // 960 for KS = 5
  static constexpr size_t SMEM = (static_cast<size_t>(HV) * 16 + kW5_TV * 16) * sizeof(float);
};
constexpr size_t kW5_SMEM = WgradRefGeom<5>::SMEM;

struct Wgrad5Args {
  const float* in1;
  const float* in2;
  int C1, C2;
  const float* dz;  // [V][Cout]
  int Cout;
  float* dw;        // [KS^3][C1+C2][Cout], pre-zeroed
  Dims dims;
  int N;
  int tiles_per_block;
};

template <int KS>
__global__ void __launch_bounds__(256) conv_wgrad_ref_kernel(Wgrad5Args p) {
  using G = WgradRefGeom<KS>;
  VNB_DYN_SMEM(float, smem);
  float* xs = smem;                 // [HV][16 ci]
  float* ds = smem + G::HV * 16;    // [TV][16 co]
  const int D = p.dims.D, H = p.dims.H, W = p.dims.W;
  const int Cin = p.C1 + p.C2, Cout = p.Cout;
  const int co_blocks = (Cout + 15) / 16;
  const int ci0 = (blockIdx.y / co_blocks) * 16, co0 = (blockIdx.y % co_blocks) * 16;
  const int t = threadIdx.x, ci = t / 16, co = t % 16;
  const int tw_n = (W + kW5_TW - 1) / kW5_TW, th_n = (H + kW5_TH - 1) / kW5_TH, td_n = (D + kW5_TD - 1) / kW5_TD;
  const long long tiles_per_sample = static_cast<long long>(tw_n) * th_n * td_n;
  const long long ntiles = tiles_per_sample * p.N;
  const long long sample = static_cast<long long>(D) * H * W;
  float acc[G::TAPS];
#pragma unroll
  for (int k = 0; k < G::TAPS; ++k) acc[k] = 0.f;

  const long long first = static_cast<long long>(blockIdx.x) * p.tiles_per_block;
  for (long long tile = first; tile < first + p.tiles_per_block && tile < ntiles; ++tile) {
    const int n = static_cast<int>(tile / tiles_per_sample);
    long long r = tile % tiles_per_sample;
    const int tw0 = static_cast<int>(r % tw_n) * kW5_TW;
    r /= tw_n;
    const int th0 = static_cast<int>(r % th_n) * kW5_TH;
    const int td0 = static_cast<int>(r / th_n) * kW5_TD;
    __syncthreads();
    for (int i = t; i < G::HV * 16; i += 256) {
      const int c = i % 16, hv = i / 16;
      const int hw = hv % G::HW, hh = (hv / G::HW) % G::HH, hd = hv / (G::HW * G::HH);
      const int gd = td0 + hd - G::R, gh = th0 + hh - G::R, gw = tw0 + hw - G::R, cg = ci0 + c;
      float val = 0.f;
      if (cg < Cin && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
        const long long vox = n * sample + (static_cast<long long>(gd) * H + gh) * W + gw;
        val = cg < p.C1 ? p.in1[vox * p.C1 + cg] : p.in2[vox * p.C2 + (cg - p.C1)];
      }
      xs[i] = val;
    }
    for (int i = t; i < kW5_TV * 16; i += 256) {
      const int c = i % 16, tv = i / 16;
      const int lw = tv % kW5_TW, lh = (tv / kW5_TW) % kW5_TH, ld = tv / (kW5_TW * kW5_TH);
      const int gd = td0 + ld, gh = th0 + lh, gw = tw0 + lw;
      float val = 0.f;
      if (co0 + c < Cout && gd < D && gh < H && gw < W) {
        const long long vox = n * sample + (static_cast<long long>(gd) * H + gh) * W + gw;
        val = p.dz[vox * Cout + co0 + c];
      }
      ds[i] = val;
    }
    __syncthreads();
    for (int tv = 0; tv < kW5_TV; ++tv) {
      const float g = ds[tv * 16 + co];
      const int lw = tv % kW5_TW, lh = (tv / kW5_TW) % kW5_TH, ld = tv / (kW5_TW * kW5_TH);
      const float* xb = xs + ((ld * G::HH + lh) * G::HW + lw) * 16 + ci;
#pragma unroll
      for (int kd = 0; kd < KS; ++kd)
#pragma unroll
        for (int kh = 0; kh < KS; ++kh)
#pragma unroll
          for (int kw = 0; kw < KS; ++kw)
            acc[(kd * KS + kh) * KS + kw] += xb[((kd * G::HH + kh) * G::HW + kw) * 16] * g;
    }
  }
  if (ci0 + ci < Cin && co0 + co < Cout) {
#pragma unroll
    for (int k = 0; k < G::TAPS; ++k)
      atomicAdd(p.dw + (static_cast<long long>(k) * Cin + ci0 + ci) * Cout + co0 + co, acc[k]);
  }
}
#define conv5_wgrad_ref_kernel conv_wgrad_ref_kernel<5>

// -------------------------------------------------------------------------------------------------
// 2x2x2 stride-2 kernels.  `fine` is [N][2Dc][2Hc][2Wc][CF], `coarse` is [N][Dc][Hc][Wc][CC],
// w is [8][CF][CC] with tap = (a*2 + b)*2 + d (a,b,d = offsets along D,H,W).
//   gather : coarse[o][cc] (+)= sum_{tap,cf} fine[2o+tap][cf] * w[tap][cf][cc] (+ bias[cc])   down fprop, up dgrad
//   scatter: fine[2i+tap][cf] (+)= sum_cc coarse[i][cc] * w[tap][cf][cc] (+ bias[cf])         up fprop, down dgrad
//   wgrad  : dw[tap][cf][cc] += sum_{n,i} fine[2i+tap][cf] * coarse[i][cc]
// -------------------------------------------------------------------------------------------------
struct K2Args {
  const float* fine_in;
  const float* coarse_in;
  float* fine_out;
  float* coarse_out;
  const float* w;
  float* dw;
  const float* bias;
  int CF, CC;
  Dims cd;  // coarse dims
  int N;
  int accumulate;
};

__global__ void k2_gather_kernel(K2Args p) {
  const long long total = static_cast<long long>(p.N) * p.cd.D * p.cd.H * p.cd.W * p.CC;
  const int Hf = 2 * p.cd.H, Wf = 2 * p.cd.W, Df = 2 * p.cd.D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % p.CC);
    long long o = i / p.CC;
    const int ow = static_cast<int>(o % p.cd.W);
    o /= p.cd.W;
    const int oh = static_cast<int>(o % p.cd.H);
    o /= p.cd.H;
    const int od = static_cast<int>(o % p.cd.D);
    const int n = static_cast<int>(o / p.cd.D);
    float s = p.bias ? p.bias[cc] : 0.f;
    for (int tap = 0; tap < 8; ++tap) {
      const int fd = 2 * od + (tap >> 2), fh = 2 * oh + ((tap >> 1) & 1), fw = 2 * ow + (tap & 1);
      const float* f = p.fine_in + (((static_cast<long long>(n) * Df + fd) * Hf + fh) * Wf + fw) * p.CF;
      const float* wr = p.w + static_cast<long long>(tap) * p.CF * p.CC + cc;
      for (int cf = 0; cf < p.CF; ++cf) s += f[cf] * wr[static_cast<long long>(cf) * p.CC];
    }
    p.coarse_out[i] = p.accumulate ? p.coarse_out[i] + s : s;
  }
}

__global__ void k2_scatter_kernel(K2Args p) {
  const int Hf = 2 * p.cd.H, Wf = 2 * p.cd.W, Df = 2 * p.cd.D;
  const long long total = static_cast<long long>(p.N) * Df * Hf * Wf * p.CF;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cf = static_cast<int>(i % p.CF);
    long long o = i / p.CF;
    const int fw = static_cast<int>(o % Wf);
    o /= Wf;
    const int fh = static_cast<int>(o % Hf);
    o /= Hf;
    const int fd = static_cast<int>(o % Df);
    const int n = static_cast<int>(o / Df);
    const int tap = ((fd & 1) << 2) | ((fh & 1) << 1) | (fw & 1);
    const float* c = p.coarse_in + (((static_cast<long long>(n) * p.cd.D + (fd >> 1)) * p.cd.H + (fh >> 1)) * p.cd.W + (fw >> 1)) * p.CC;
    const float* wr = p.w + (static_cast<long long>(tap) * p.CF + cf) * p.CC;
    float s = p.bias ? p.bias[cf] : 0.f;
    for (int cc = 0; cc < p.CC; ++cc) s += c[cc] * wr[cc];
    p.fine_out[i] = p.accumulate ? p.fine_out[i] + s : s;
  }
}

// grid: (ceil(8*CF*CC / 256), voxel splits); fp32 atomics into pre-zeroed dw
__global__ void k2_wgrad_kernel(K2Args p, int voxels_per_split) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= 8LL * p.CF * p.CC) return;
  const int cc = static_cast<int>(e % p.CC);
  const int cf = static_cast<int>((e / p.CC) % p.CF);
  const int tap = static_cast<int>(e / (static_cast<long long>(p.CC) * p.CF));
  const int Hf = 2 * p.cd.H, Wf = 2 * p.cd.W, Df = 2 * p.cd.D;
  const long long V = static_cast<long long>(p.N) * p.cd.D * p.cd.H * p.cd.W;
  const long long v0 = static_cast<long long>(blockIdx.y) * voxels_per_split;
  float s = 0.f;
  for (long long v = v0; v < v0 + voxels_per_split && v < V; ++v) {
    long long o = v;
    const int ow = static_cast<int>(o % p.cd.W);
    o /= p.cd.W;
    const int oh = static_cast<int>(o % p.cd.H);
    o /= p.cd.H;
    const int od = static_cast<int>(o % p.cd.D);
    const int n = static_cast<int>(o / p.cd.D);
    const int fd = 2 * od + (tap >> 2), fh = 2 * oh + ((tap >> 1) & 1), fw = 2 * ow + (tap & 1);
    s += p.fine_in[(((static_cast<long long>(n) * Df + fd) * Hf + fh) * Wf + fw) * p.CF + cf] * p.coarse_in[v * p.CC + cc];
  }
  atomicAdd(p.dw + e, s);
}

// -------------------------------------------------------------------------------------------------
// 1x1x1 convolution (output layer, networks.py:302): w [Cin][K]
// -------------------------------------------------------------------------------------------------
__global__ void conv1_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ z, long long V, int Cin, int K) {
  const long long total = V * K;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const long long v = i / K;
    float s = bias ? bias[k] : 0.f;
    for (int c = 0; c < Cin; ++c) s += x[v * Cin + c] * w[c * K + k];
    z[i] = s;
  }
}
__global__ void conv1_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w,
                                   float* __restrict__ dx, long long V, int Cin, int K, int accumulate) {
  const long long total = V * Cin;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cin);
    const long long v = i / Cin;
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += dz[v * K + k] * w[c * K + k];
    dx[i] = accumulate ? dx[i] + s : s;
  }
}
// dw[c][k] += sum_v x[v][c] * dz[v][k].  256 threads = (Cin*K pairs) x (256 / pairs voxel lanes);
// lanes stride over the block's voxel chunk, partial sums are combined through shared memory.
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                          float* __restrict__ dw, long long V, int Cin, int K,
                                                          int voxels_per_block) {
  __shared__ float red[256];
  const int pairs = Cin * K;
  const int lanes = 256 / pairs;  // pairs <= 256 (checked by the engine)
  const int t = threadIdx.x, pr = t % pairs, ln = t / pairs;
  const int c = pr / K, k = pr % K;
  const long long v0 = static_cast<long long>(blockIdx.x) * voxels_per_block;
  const long long v1 = v0 + voxels_per_block < V ? v0 + voxels_per_block : V;
  float s = 0.f;
  if (ln < lanes)
    for (long long v = v0 + ln; v < v1; v += lanes) s += x[v * Cin + c] * dz[v * K + k];
  red[t] = ln < lanes ? s : 0.f;
  __syncthreads();
  if (t < pairs) {
    float tot = 0.f;
    for (int l = 0; l < lanes; ++l) tot += red[l * pairs + t];
    atomicAdd(dw + t, tot);
  }
}

// Voxel-per-thread forms of the three head kernels for Cin % 4 == 0, Cin <= 32, K <= 4 (every V-Net head): one thread
// streams a voxel's Cin inputs as float4 (warps read contiguous memory), the [Cin][K] filter sits in shared memory.
// Same arithmetic order over c as the scalar kernels above.
constexpr int kC1MaxCin = 32, kC1MaxK = 4;
__global__ void __launch_bounds__(256) conv1_fprop_vox_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ z, long long V,
                                                              int Cin, int K) {
  __shared__ float ws[kC1MaxCin * kC1MaxK];
  for (int i = threadIdx.x; i < Cin * K; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int C4 = Cin / 4;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc[kC1MaxK];
#pragma unroll
    for (int k = 0; k < kC1MaxK; ++k) acc[k] = (bias && k < K) ? bias[k] : 0.f;
    const float4* xv = reinterpret_cast<const float4*>(x + v * Cin);
    for (int c4 = 0; c4 < C4; ++c4) {
      const float4 f = xv[c4];
      const float xs[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < kC1MaxK; ++k)
          if (k < K) acc[k] += xs[j] * ws[(c4 * 4 + j) * K + k];
    }
    if (K == 2) {
      *reinterpret_cast<float2*>(z + v * 2) = make_float2(acc[0], acc[1]);
    } else if (K == 4) {
      *reinterpret_cast<float4*>(z + v * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int k = 0; k < kC1MaxK; ++k)
        if (k < K) z[v * K + k] = acc[k];
    }
  }
}
__global__ void __launch_bounds__(256) conv1_dgrad_vox_kernel(const float* __restrict__ dz, const float* __restrict__ w,
                                                              float* __restrict__ dx, long long V, int Cin, int K, int accumulate) {
  __shared__ float ws[kC1MaxCin * kC1MaxK];
  for (int i = threadIdx.x; i < Cin * K; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int C4 = Cin / 4;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float g[kC1MaxK];
#pragma unroll
    for (int k = 0; k < kC1MaxK; ++k) g[k] = k < K ? dz[v * K + k] : 0.f;
    float4* out = reinterpret_cast<float4*>(dx + v * Cin);
    for (int c4 = 0; c4 < C4; ++c4) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float sacc = 0.f;
#pragma unroll
        for (int k = 0; k < kC1MaxK; ++k)
          if (k < K) sacc += g[k] * ws[(c4 * 4 + j) * K + k];
        o[j] = sacc;
      }
      float4 r = make_float4(o[0], o[1], o[2], o[3]);
      if (accumulate) {
        const float4 old = out[c4];
        r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
      }
      out[c4] = r;
    }
  }
}
// dw[c][k] += sum_v x[v][c] dz[v][k] for Cin <= 16: 16 x 4 register accumulators per thread over a strided voxel set
// (all indices static after unrolling), a warp butterfly, a shared-memory sum over the 8 warps and one atomic per
// (c, k) and block
constexpr int kC1WgMaxCin = 16;
__global__ void __launch_bounds__(256) conv1_wgrad_vox_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                              float* __restrict__ dw, long long V, int Cin, int K) {
  __shared__ float red[8][kC1WgMaxCin * kC1MaxK];
  float acc[kC1WgMaxCin][kC1MaxK];
#pragma unroll
  for (int c = 0; c < kC1WgMaxCin; ++c)
#pragma unroll
    for (int k = 0; k < kC1MaxK; ++k) acc[c][k] = 0.f;
  const int C4 = Cin / 4;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float g[kC1MaxK];
#pragma unroll
    for (int k = 0; k < kC1MaxK; ++k) g[k] = k < K ? dz[v * K + k] : 0.f;
    const float4* xv = reinterpret_cast<const float4*>(x + v * Cin);
#pragma unroll
    for (int c4 = 0; c4 < kC1WgMaxCin / 4; ++c4) {
      if (c4 < C4) {
        const float4 f = xv[c4];
        const float xs[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < kC1MaxK; ++k) acc[c4 * 4 + j][k] += xs[j] * g[k];   // g[k] = 0 for k >= K
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < kC1WgMaxCin; ++c)
#pragma unroll
    for (int k = 0; k < kC1MaxK; ++k) {
      float sacc = acc[c][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (lane == 0) red[warp][c * kC1MaxK + k] = sacc;
    }
  __syncthreads();
  if (threadIdx.x < Cin * K) {
    const int c = threadIdx.x / K, k = threadIdx.x % K;
    float tot = 0.f;
    for (int wq = 0; wq < 8; ++wq) tot += red[wq][c * kC1MaxK + k];
    atomicAdd(dw + threadIdx.x, tot);
  }
}

// -------------------------------------------------------------------------------------------------
// General 1x1x1 convolution (shortcut branch and output layer of the attention / output modules,
// attention.py:98-100,111): a [V][Cin] x [Cin][Cout] matrix product, fp32 FMA.
//   fprop : out[v][co]  = sum_ci x[v][ci] w[ci][co] + bias[co] + res[v][co]
//   dgrad : same kernel with `transposed` = 1 (w read as [Cout][Cin]) and `accumulate`
//   wgrad : dw[ci][co] += sum_v x[v][ci] dz[v][co]   (dw pre-zeroed, fp32 atomics across voxel splits)
// Block = 64 voxels x 64 outputs; thread = one voxel x 16 consecutive outputs.
// -------------------------------------------------------------------------------------------------
struct Conv1Args {
  const float* x;     // [V][K]   (K = reduction width)
  const float* w;     // [K][M], or [M][K] when transposed
  const float* bias;  // [M] or nullptr
  const float* res;   // [V][M] or nullptr
  float* out;         // [V][M]
  long long V;
  int K, M;
  int transposed, accumulate;
};

__global__ void __launch_bounds__(256) conv1g_kernel(Conv1Args p) {
  __shared__ float xs[64][65];
  __shared__ float ws[64][64];
  const int t = threadIdx.x, lv = t % 64, m0 = blockIdx.y * 64 + (t / 64) * 16;
  const long long v0 = static_cast<long long>(blockIdx.x) * 64, v = v0 + lv;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int kb = 0; kb < p.K; kb += 64) {
    const int kn = p.K - kb < 64 ? p.K - kb : 64;
    __syncthreads();
    for (int i = t; i < 64 * kn; i += 256) {
      const int k = i % kn, r = i / kn;
      xs[r][k] = (v0 + r < p.V) ? p.x[(v0 + r) * p.K + kb + k] : 0.f;
    }
    for (int i = t; i < 64 * 64; i += 256) {
      const int m = i % 64, k = i / 64, mg = blockIdx.y * 64 + m;
      float val = 0.f;
      if (k < kn && mg < p.M)
        val = p.transposed ? p.w[static_cast<long long>(mg) * p.K + kb + k] : p.w[static_cast<long long>(kb + k) * p.M + mg];
      ws[k][m] = val;
    }
    __syncthreads();
    const int mo = (t / 64) * 16;
    for (int k = 0; k < kn; ++k) {
      const float xv = xs[lv][k];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += xv * ws[k][mo + j];
    }
  }
  if (v >= p.V) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int m = m0 + j;
    if (m >= p.M) break;
    float y = acc[j];
    if (p.bias) y += p.bias[m];
    if (p.res) y += p.res[v * p.M + m];
    float* o = p.out + v * p.M + m;
    *o = p.accumulate ? *o + y : y;
  }
}

// grid (voxel splits, ceil(Cin/64), ceil(Cout/64)); thread (ty, tx) owns a 4x4 block of the 64x64 tile
__global__ void __launch_bounds__(256) conv1g_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                           float* __restrict__ dw, long long V, int Cin, int Cout,
                                                           long long voxels_per_block) {
  __shared__ float xs[32][64];
  __shared__ float ds[32][64];
  const int t = threadIdx.x, ty = t / 16, tx = t % 16;
  const int ci0 = blockIdx.y * 64, co0 = blockIdx.z * 64;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const long long vb = static_cast<long long>(blockIdx.x) * voxels_per_block;
  const long long ve = vb + voxels_per_block < V ? vb + voxels_per_block : V;
  for (long long v0 = vb; v0 < ve; v0 += 32) {
    __syncthreads();
    for (int i = t; i < 32 * 64; i += 256) {
      const int c = i % 64, r = i / 64;
      const bool in = v0 + r < ve;
      xs[r][c] = (in && ci0 + c < Cin) ? x[(v0 + r) * Cin + ci0 + c] : 0.f;
      ds[r][c] = (in && co0 + c < Cout) ? dz[(v0 + r) * Cout + co0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      float xa[4], db[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xa[a] = xs[r][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) db[b] = ds[r][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] += xa[a] * db[b];
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int ci = ci0 + ty * 4 + a, co = co0 + tx * 4 + b;
      if (ci < Cin && co < Cout) atomicAdd(dw + static_cast<long long>(ci) * Cout + co, acc[a][b]);
    }
}

}  // namespace vnb
