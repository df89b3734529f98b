// HBM-bound kernels of the V-Net hot path: per-channel batch-norm statistics, the fused
// BN-chain + PReLU + dropout apply pass, their backward passes, softmax + Dice/Jaccard/x-ent
// reductions (forward, argmax, backward), and the multi-tensor Adam/SGD step.
//
// Reference semantics restated (paths relative to /root/reference):
//   tf.layers.batch_normalization(training=True)     networks.py:259..361  -> bn_stats / bn_finalize / bn_apply
//   prelu                                           layers2.py:97-99      -> bn_apply (fused)
//   tf.nn.dropout(rate)                             networks.py:321..363  -> bn_apply (fused, counter RNG)
//   softmax / one_hot / dice_coe / x-ent / argmax   model.py:26-92,447,477,495-568 -> softmax_loss_*
//   exponential_decay + Adam/SGD                    model.py:641-660      -> optimizer_step
//
// Layout: activations are NDHWC fp32, i.e. a [V][C] matrix per tensor with V = N*D*H*W voxels and
// channels fastest.  All per-channel reductions are two-stage and deterministic: each block writes a
// partial in double precision, a single-block finalize kernel sums the partials in index order.
#pragma once
#include "bn_chain.h"
#include "vnb_cuda.h"

namespace vnb {

constexpr int kRedThreads = 256;   // upper bound; actual block = (256 / CW) * CW threads
constexpr int kMaxRedBlocks = 1184; // 8 blocks per SM on 148 SMs
constexpr float kBnEps = 1e-3f;     // networks.py:259 epsilon=0.001
constexpr float kBnMomentum = 0.99f;

// ---------------------------------------------------------------------------------------------
// counter-based dropout RNG: keep-mask is a pure function of (seed, unit, element index), so the
// backward pass regenerates it instead of storing it.  tf.nn.dropout keeps where u >= rate.
// ---------------------------------------------------------------------------------------------
// One 32-bit hash (lowbias32) yields two 16-bit uniforms, so a float4 of activations costs two hashes.
// keep(idx) <=> u16 >= round(rate * 65536); the scale uses the quantised rate so E[mask * scale] = 1.
VNB_HD uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
VNB_HD uint32_t dropout_key(uint64_t seed, uint32_t unit) {
  return hash32(static_cast<uint32_t>(seed) ^ hash32(static_cast<uint32_t>(seed >> 32) + 0x9E3779B1u * (unit + 1u)));
}
// 32 random bits shared by elements idx with equal idx >> 1
VNB_HD uint32_t dropout_bits(uint32_t key, uint64_t idx) {
  const uint64_t pair = idx >> 1;
  return hash32(static_cast<uint32_t>(pair) * 0x9E3779B1u + static_cast<uint32_t>(pair >> 32) * 0x85EBCA77u + key);
}
VNB_HD uint32_t dropout_threshold(float rate) { return static_cast<uint32_t>(rate * 65536.0f + 0.5f); }
VNB_HD float dropout_keep_scale(float rate) {
  return rate > 0.f ? 65536.0f / (65536.0f - static_cast<float>(dropout_threshold(rate))) : 1.0f;
}
VNB_HD bool dropout_keep(uint32_t key, uint64_t idx, uint32_t thresh) {
  const uint32_t b = dropout_bits(key, idx);
  return ((idx & 1u) ? (b >> 16) : (b & 0xFFFFu)) >= thresh;
}

// ---------------------------------------------------------------------------------------------
// block-level per-channel combine. Thread t owns channel (t % CW); acc[] are its private sums.
// partial layout: [block][NQ][C] doubles.
// ---------------------------------------------------------------------------------------------
template <int NQ>
__device__ __forceinline__ void block_channel_combine(const double (&acc)[NQ], int CW, int C, int c0,
                                                      double* __restrict__ partial) {
  __shared__ double red[NQ][kRedThreads];
  const int t = threadIdx.x, nt = blockDim.x;
  for (int q = 0; q < NQ; ++q) red[q][t] = acc[q];
  __syncthreads();
  if (t < CW) {
    for (int q = 0; q < NQ; ++q) {
      double s = 0.0;
      for (int g = t; g < nt; g += CW) s += red[q][g];
      partial[(static_cast<size_t>(blockIdx.x) * NQ + q) * C + c0 + t] = s;
    }
  }
}

struct RedGeom {  // how a [V][C] tensor (or a channel window of it) is spread over a reduction grid
  int C, c0, CW;  // total channels, window start, window width (CW <= 256, blockDim % CW == 0)
  long long V;
};

// element loop used by every per-channel reduction: thread t -> channel c0 + t % CW,
// voxels (blockIdx*G + t / CW) + k * gridDim*G, G = blockDim / CW voxel groups per block
#define VNB_CHANNEL_LOOP(geom, v, c)                                              \
  const int c = (geom).c0 + static_cast<int>(threadIdx.x) % (geom).CW;            \
  const long long vstep_ = static_cast<long long>(gridDim.x) * (blockDim.x / (geom).CW); \
  for (long long v = static_cast<long long>(blockIdx.x) * (blockDim.x / (geom).CW) + threadIdx.x / (geom).CW; \
       v < (geom).V; v += vstep_)

// ---------------------------------------------------------------------------------------------
// BN forward statistics: partial[blk][2][C] = (sum z, sum z^2)
// ---------------------------------------------------------------------------------------------
__global__ void bn_stats_kernel(const float* __restrict__ z, RedGeom g, double* __restrict__ partial) {
  double acc[2] = {0.0, 0.0};
  float s = 0.f, s2 = 0.f;
  int cnt = 0;
  VNB_CHANNEL_LOOP(g, v, c) {
    const float x = z[v * g.C + c];
    s += x;
    s2 += x * x;
    if (++cnt == 64) {  // flush fp32 running sums into double every 64 elements
      acc[0] += s;
      acc[1] += s2;
      s = s2 = 0.f;
      cnt = 0;
    }
  }
  acc[0] += s;
  acc[1] += s2;
  block_channel_combine<2>(acc, g.CW, g.C, g.c0, partial);
}

// input layer, in_channels == 1 (networks.py:254-259): statistics of the single-channel image
__global__ void image_stats_kernel(const float* __restrict__ img, long long V, double* __restrict__ partial) {
  double acc[2] = {0.0, 0.0};
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double x = img[v];
    acc[0] += x;
    acc[1] += x * x;
  }
  __shared__ double red[2][kRedThreads];
  red[0][threadIdx.x] = acc[0];
  red[1][threadIdx.x] = acc[1];
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (unsigned i = 0; i < blockDim.x; ++i) {
      a += red[0][i];
      b += red[1][i];
    }
    partial[blockIdx.x * 2 + 0] = a;
    partial[blockIdx.x * 2 + 1] = b;
  }
}

// ---------------------------------------------------------------------------------------------
// Synchronised batch norm (data parallel, optional): the per-block partial sums partial[blk][nq][PC] are collapsed
// to one row out[nq][PC] in a fixed order; the engine sums that row over the ranks and hands it to the finalize
// kernels as a one-block partial array.  One block per (quantity, channel).
// ---------------------------------------------------------------------------------------------
__global__ void partial_collapse_kernel(const double* __restrict__ partial, int nblk, int nq, int PC,
                                        double* __restrict__ out) {
  const int qc = blockIdx.x;
  __shared__ double col_red[128];
  double s = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) s += partial[static_cast<size_t>(b) * nq * PC + qc];
  col_red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double t = 0.0;
  for (unsigned i = 0; i < blockDim.x; ++i) t += col_red[i];
  out[qc] = t;
}

// ---------------------------------------------------------------------------------------------
// BN forward finalize: partial sums -> mu, sigma^2 -> chain closed form -> scale/shift; optional
// moving-average update (only train_op runs UPDATE_OPS, model.py:665-666).
// One thread per channel. `single_channel_stats`: all channels share partial column 0 (tiled input).
// ---------------------------------------------------------------------------------------------
struct BnParams {       // device pointers into the flat parameter / state buffers
  const float* gamma[3];
  const float* beta[3];
  float* moving_mean[3];
  float* moving_var[3];
};

// per-channel forward finalize: (sum z, sum z^2) -> mu, sigma^2 -> chain closed form -> scale / shift (+ UPDATE_OPS)
__device__ __forceinline__ void bn_fwd_channel(int c, double s, double s2, double count, int chain, const BnParams& bp,
                                               int update_moving, double* __restrict__ mean_out, double* __restrict__ var_out,
                                               float* __restrict__ scale_out, float* __restrict__ shift_out) {
  const double mu = s / count;
  double var = s2 / count - mu * mu;
  if (var < 0.0) var = 0.0;
  double gam[3] = {1, 1, 1}, bet[3] = {0, 0, 0};
  const int nbn = chain_num_bn(chain);
  for (int k = 0; k < nbn; ++k) {
    gam[k] = bp.gamma[k][c];
    bet[k] = bp.beta[k][c];
  }
  const ChainOut o = chain_eval(chain, var, gam, bet, static_cast<double>(kBnEps));
  mean_out[c] = mu;
  var_out[c] = var;
  const double A = o.A.v, B = bet[o.beta_idx];
  scale_out[c] = static_cast<float>(A);
  shift_out[c] = static_cast<float>(B - A * mu);
  if (update_moving) {
    for (int k = 0; k < nbn; ++k) {
      const float bm = static_cast<float>((o.mean_is_mu[k] ? mu : 0.0) + o.bn_mean[k]);
      const float bv = static_cast<float>(o.bn_var[k]);
      // assign_moving_average: var -= (var - value) * (1 - momentum)
      bp.moving_mean[k][c] -= (bp.moving_mean[k][c] - bm) * (1.0f - kBnMomentum);
      bp.moving_var[k][c] -= (bp.moving_var[k][c] - bv) * (1.0f - kBnMomentum);
    }
  }
}

__global__ void bn_finalize_fwd_kernel(const double* __restrict__ partial, int nblk, int nq_stride, int C,
                                       double count, int chain, BnParams bp, int single_channel_stats,
                                       int update_moving, double* __restrict__ mean_out,
                                       double* __restrict__ var_out, float* __restrict__ scale_out,
                                       float* __restrict__ shift_out, int inference = 0) {
  pdl_trigger();   // see vnb_cuda.h: the next short pass may become resident now ...
  pdl_wait();      // ... and this one reads nothing of its predecessor before that grid has completed
  const int c = blockIdx.x;  // one block per channel; threads sum the per-block partials in a fixed order
  if (inference) {  // training=False: normalise with the moving statistics (attention.py:69 fed train_phase=False)
    if (threadIdx.x != 0) return;
    const double mu = bp.moving_mean[0][c], var = bp.moving_var[0][c];
    const double A = static_cast<double>(bp.gamma[0][c]) * rsqrt(var + static_cast<double>(kBnEps));
    mean_out[c] = mu;
    var_out[c] = var;
    scale_out[c] = static_cast<float>(A);
    shift_out[c] = static_cast<float>(static_cast<double>(bp.beta[0][c]) - A * mu);
    return;
  }
  const int pc = single_channel_stats ? 0 : c;
  const int PC = single_channel_stats ? 1 : C;
  __shared__ double fin_red[2][128];
  double ps = 0.0, ps2 = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    ps += partial[(static_cast<size_t>(b) * nq_stride + 0) * PC + pc];
    ps2 += partial[(static_cast<size_t>(b) * nq_stride + 1) * PC + pc];
  }
  fin_red[0][threadIdx.x] = ps;
  fin_red[1][threadIdx.x] = ps2;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double s = 0.0, s2 = 0.0;
  for (unsigned i = 0; i < blockDim.x; ++i) {
    s += fin_red[0][i];
    s2 += fin_red[1][i];
  }
  bn_fwd_channel(c, s, s2, count, chain, bp, update_moving, mean_out, var_out, scale_out, shift_out);
}

// ---------------------------------------------------------------------------------------------
// fused apply: a = dropout(prelu(scale*z + shift)); optional bf16 hi (+lo) copies for the
// tensor-core convolution path (a ~= hi + lo, lo = bf16(a - hi)).
// ---------------------------------------------------------------------------------------------
struct ApplyArgs {
  const float* z;        // [V][C] pre-BN tensor; for the tiled input layer: [V][1] image
  float* a;              // [V][C] output activation (fp32)
  uint16_t* a_hi;        // optional bf16 copies (nullptr = skip)
  uint16_t* a_lo;
  const float* scale;    // [C]
  const float* shift;    // [C]
  const float* alpha;    // [C] or nullptr (no activation: input layer M==1, output layer)
  long long total;       // V*C
  int C;
  int tiled_input;       // z has one channel, broadcast over C (tf.tile, networks.py:258)
  float drop_rate;       // 0 = identity
  uint64_t seed;
  uint32_t unit;
};

__global__ void bn_apply_kernel(ApplyArgs p) {
  const float keep_scale = dropout_keep_scale(p.drop_rate);
  const uint32_t dkey = dropout_key(p.seed, p.unit), dthr = dropout_threshold(p.drop_rate);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % p.C);
    const float zin = p.tiled_input ? p.z[i / p.C] : p.z[i];
    float y = p.scale[c] * zin + p.shift[c];
    if (p.alpha) y = y > 0.f ? y : p.alpha[c] * y;  // max(0,y) + alpha*min(0,y)
    if (p.drop_rate > 0.f) y = dropout_keep(dkey, static_cast<uint64_t>(i), dthr) ? y * keep_scale : 0.f;
    p.a[i] = y;
    if (p.a_hi) {
      const uint16_t hi = f32_to_bf16(y);
      p.a_hi[i] = hi;
      if (p.a_lo) p.a_lo[i] = f32_to_bf16(y - bf16_to_f32(hi));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward of the fused apply + BN chain.
//   d    = dL/da (after the consumers accumulated into it)
//   g    = dL/dyhat = dropout_bwd(d) * prelu'(yhat)          yhat = scale*z + shift
//   R0 = sum g, R1 = sum g*(z - mu), Ralpha = sum dropout_bwd(d) * min(yhat, 0)
// ---------------------------------------------------------------------------------------------
struct BwdArgs {
  const float* z;       // [V][C] (or [V][1] for the tiled input layer)
  float* d;             // [V][C] in: dL/da ; out (apply kernel): dL/dz, in place
  uint16_t* d_hi;       // optional bf16 copies of dL/dz for the tensor-core dgrad/wgrad
  uint16_t* d_lo;
  float* res_grad;      // optional: residual branch gradient target (block input's dL/da)
  int res_accumulate;   // 0: write, 1: add
  float* res_grad2;     // optional second target (the summand of an add unit)
  int res_accumulate2;
  const float* scale;
  const float* shift;
  const float* alpha;   // or nullptr
  const double* mean;   // [C]
  const float* P;       // [C] coefficients from bn_finalize_bwd (apply kernel only)
  const float* Q;
  const float* S;
  int C;
  int tiled_input;
  float drop_rate;
  uint64_t seed;
  uint32_t unit;
};

__device__ __forceinline__ float bwd_g(const BwdArgs& p, long long i, int c, float zin, float& yhat, float& dd) {
  yhat = p.scale[c] * zin + p.shift[c];
  dd = p.d[i];
  if (p.drop_rate > 0.f)
    dd = dropout_keep(dropout_key(p.seed, p.unit), static_cast<uint64_t>(i), dropout_threshold(p.drop_rate))
             ? dd * dropout_keep_scale(p.drop_rate) : 0.f;
  if (!p.alpha) return dd;
  // TF gradients of maximum(0,y)/minimum(0,y): 1 for y>0, alpha for y<0, 0 at the tie
  return yhat > 0.f ? dd : (yhat < 0.f ? dd * p.alpha[c] : 0.f);
}

__global__ void bn_bwd_reduce_kernel(BwdArgs p, RedGeom g, double* __restrict__ partial) {
  double acc[3] = {0.0, 0.0, 0.0};
  VNB_CHANNEL_LOOP(g, v, c) {
    const long long i = v * g.C + c;
    const float zin = p.tiled_input ? p.z[v] : p.z[i];
    float yhat, dd;
    const float gg = bwd_g(p, i, c, zin, yhat, dd);
    acc[0] += gg;
    acc[1] += static_cast<double>(gg) * (static_cast<double>(zin) - p.mean[c]);
    if (p.alpha && yhat < 0.f) acc[2] += static_cast<double>(dd) * yhat;
  }
  block_channel_combine<3>(acc, g.CW, g.C, g.c0, partial);
}

struct BnGradPtrs {  // where the parameter gradients go (flat gradient buffer), nullptr = none
  float* dgamma[3];
  float* dbeta[3];
  float* dalpha;
  float* dbias;  // only with inference-mode BN (batch statistics make the conv bias gradient exactly 0)
};

// per-channel backward finalize: (R0, R1, Ralpha) -> parameter gradients and the dL/dz coefficients P, Q, S
__device__ __forceinline__ void bn_bwd_channel(int c, double R0, double R1, double Ra, int C, double count, int chain,
                                               const BnParams& bp, const double* __restrict__ var, const BnGradPtrs& gp,
                                               float* __restrict__ P, float* __restrict__ Q, float* __restrict__ S, int inference,
                                               const double* __restrict__ gsum, double gcount) {
  double gam[3] = {1, 1, 1}, bet[3] = {0, 0, 0};
  const int nbn = chain_num_bn(chain);
  for (int k = 0; k < nbn; ++k) {
    gam[k] = bp.gamma[k][c];
    bet[k] = bp.beta[k][c];
  }
  const ChainOut o = chain_eval(chain, var[c], gam, bet, static_cast<double>(kBnEps));
  if (inference) {  // statistics are constants: dz = A*g, dgamma = R1*rsqrt(var+eps), dbeta = R0, dbias = A*R0
    P[c] = static_cast<float>(o.A.v);
    Q[c] = 0.f;
    S[c] = 0.f;
    if (gp.dgamma[0]) gp.dgamma[0][c] = static_cast<float>(R1 * o.A.d[1]);
    if (gp.dbeta[0]) gp.dbeta[0][c] = static_cast<float>(R0);
    if (gp.dbias) gp.dbias[c] = static_cast<float>(o.A.v * R0);
    if (gp.dalpha) gp.dalpha[c] = static_cast<float>(Ra);
    return;
  }
  // synchronised batch norm: the statistics belong to the global batch, so dL/dz takes R0, R1 and the count summed
  // over the ranks (gsum = [2][C]); the parameter gradients below stay local sums - the gradient exchange averages them
  const double G0 = gsum ? gsum[c] : R0, G1 = gsum ? gsum[C + c] : R1, gc = gsum ? gcount : count;
  P[c] = static_cast<float>(o.A.v);
  Q[c] = static_cast<float>(-o.A.v * G0 / gc);
  S[c] = static_cast<float>(2.0 * o.A.d[0] * G1 / gc);
  for (int k = 0; k < nbn; ++k) {
    if (gp.dgamma[k]) gp.dgamma[k][c] = static_cast<float>(R1 * o.A.d[1 + k]);
    if (gp.dbeta[k]) gp.dbeta[k][c] = (k == o.beta_idx) ? static_cast<float>(R0) : 0.f;
  }
  if (gp.dalpha) gp.dalpha[c] = static_cast<float>(Ra);
}

__global__ void bn_finalize_bwd_kernel(const double* __restrict__ partial, int nblk, int C, double count,
                                       int chain, BnParams bp, const double* __restrict__ var,
                                       BnGradPtrs gp, float* __restrict__ P, float* __restrict__ Q,
                                       float* __restrict__ S, int inference = 0,
                                       const double* __restrict__ gsum = nullptr, double gcount = 0.0) {
  pdl_trigger();   // see vnb_cuda.h: the next short pass may become resident now ...
  pdl_wait();      // ... and this one reads nothing of its predecessor before that grid has completed
  const int c = blockIdx.x;  // one block per channel
  __shared__ double fin_red[3][128];
  double p0 = 0, p1 = 0, p2 = 0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    p0 += partial[(static_cast<size_t>(b) * 3 + 0) * C + c];
    p1 += partial[(static_cast<size_t>(b) * 3 + 1) * C + c];
    p2 += partial[(static_cast<size_t>(b) * 3 + 2) * C + c];
  }
  fin_red[0][threadIdx.x] = p0;
  fin_red[1][threadIdx.x] = p1;
  fin_red[2][threadIdx.x] = p2;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double R0 = 0, R1 = 0, Ra = 0;
  for (unsigned i = 0; i < blockDim.x; ++i) {
    R0 += fin_red[0][i];
    R1 += fin_red[1][i];
    Ra += fin_red[2][i];
  }
  bn_bwd_channel(c, R0, R1, Ra, C, count, chain, bp, var, gp, P, Q, S, inference, gsum, gcount);
}

// dL/dz = P*g + Q + S*(z - mu), in place over d; optional residual-branch gradient and bf16 copies
__global__ void bn_bwd_apply_kernel(BwdArgs p, long long total) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % p.C);
    const float zin = p.z[i];
    float yhat, dd;
    const float gg = bwd_g(p, i, c, zin, yhat, dd);
    const float dz = p.P[c] * gg + p.Q[c] + p.S[c] * (zin - static_cast<float>(p.mean[c]));
    p.d[i] = dz;
    if (p.res_grad) p.res_grad[i] = p.res_accumulate ? p.res_grad[i] + dz : dz;
    if (p.res_grad2) p.res_grad2[i] = p.res_accumulate2 ? p.res_grad2[i] + dz : dz;
    if (p.d_hi) {
      const uint16_t hi = f32_to_bf16(dz);
      p.d_hi[i] = hi;
      if (p.d_lo) p.d_lo[i] = f32_to_bf16(dz - bf16_to_f32(hi));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// softmax + loss reductions (model.py:447,477,26-92,495-560) and argmax (model.py:568)
// partial layout: [n][blk][K][4] = (I, L, R, X) with X = sum_v cw[t_v] * (-log p_{t_v}) [t==c slot]
// ---------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 8;

struct LossCfg {
  int K;
  int jaccard;            // 0 sorensen, 1 jaccard
  int weighted_dice;      // use class weights in the dice term
  int use_dice;           // dice term present
  int use_xent;           // cross-entropy term present
  int weighted_xent;      // class-weighted cross entropy
  int fg_only;            // legacy train.py:373-377 'sorensen': Dice of channel 1 against the label volume only
  float xent_alpha;       // multiplier of the x-ent term (Loss.Alpha for mixed_*, 1 for pure)
  float smooth;           // 1e-5
  float w[kMaxClasses];   // Loss.Weights
};

__global__ void softmax_loss_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                        long long Vn /*voxels per sample*/, LossCfg cfg,
                                        float* __restrict__ softmax_out, long long* __restrict__ argmax_out,
                                        double* __restrict__ partial) {
  const int K = cfg.K, n = blockIdx.y;
  double acc[kMaxClasses][4];
  for (int c = 0; c < K; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < Vn;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long gv = static_cast<long long>(n) * Vn + v;
    float x[kMaxClasses];
    float mx = -3.4e38f;
    int am = 0;
    for (int c = 0; c < K; ++c) {
      x[c] = logits[gv * K + c];
      if (x[c] > mx) {  // strict > : lowest index wins ties (tf.argmax)
        mx = x[c];
        am = c;
      }
    }
    float se = 0.f;
    for (int c = 0; c < K; ++c) {
      x[c] = expf(x[c] - mx);
      se += x[c];
    }
    const float inv = 1.0f / se;
    const int t = labels ? labels[gv] : -1;
    for (int c = 0; c < K; ++c) {
      const float pc = x[c] * inv;
      if (softmax_out) softmax_out[gv * K + c] = pc;
      const float tc = (t == c) ? 1.f : 0.f;
      acc[c][0] += pc * tc;
      acc[c][1] += cfg.jaccard ? pc * pc : pc;
      acc[c][2] += tc;  // t*t == t
      if (t == c) acc[c][3] += -static_cast<double>(logf(pc > 1e-38f ? pc : 1e-38f));
    }
    if (argmax_out) argmax_out[gv] = am;
  }
  if (!partial) return;
  __shared__ double red[kRedThreads];
  for (int c = 0; c < K; ++c)
    for (int q = 0; q < 4; ++q) {
      red[threadIdx.x] = acc[c][q];
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (unsigned i = 0; i < blockDim.x; ++i) s += red[i];
        partial[((static_cast<size_t>(n) * gridDim.x + blockIdx.x) * K + c) * 4 + q] = s;
      }
      __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Step metrics: the tf.metrics block of summary_op (model.py:586-626), which the reference fetches in the same
// sess.run as train_op (model.py:743-748).  Everything there is a function of integer counts over the batch:
//   confusion[(K+1)][K]: rows = label class (row K collects labels outside [0,K), whose one-hot row is all zero),
//                        columns = argmax class (first maximum).  accuracy, true/false positives/negatives,
//                        sensitivity / specificity / dice per class follow on the host (tf.metrics semantics).
//   auc_hist[K][2][kAucBins]: per class and per truth value (label != c | label == c) the histogram of
//                        bin = number of tf.metrics.auc thresholds strictly below softmax[c]  (200 thresholds:
//                        -1e-7, j/199 for j = 1..198, 1 + 1e-7, compared in fp32 as `prediction > threshold`);
//                        the confusion matrix at every threshold is a suffix sum of it.
// HBM-bound: (K + 1) * 4 B read per voxel; block-local integer histograms in shared memory, integer atomics to the
// global counters (exact, order independent).
// ---------------------------------------------------------------------------------------------
constexpr int kAucThresholds = 200;             // tf.metrics.auc default num_thresholds
constexpr int kAucBins = kAucThresholds + 1;

__global__ void __launch_bounds__(256) metrics_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                                      long long V, int K, int want_auc,
                                                      unsigned long long* __restrict__ confusion,
                                                      unsigned long long* __restrict__ auc_hist) {
  __shared__ unsigned int cm[(kMaxClasses + 1) * kMaxClasses];
  __shared__ unsigned int hist[kMaxClasses * 2 * kAucBins];
  __shared__ float thr[kAucThresholds];
  for (int i = threadIdx.x; i < (kMaxClasses + 1) * kMaxClasses; i += blockDim.x) cm[i] = 0u;
  for (int i = threadIdx.x; i < kMaxClasses * 2 * kAucBins; i += blockDim.x) hist[i] = 0u;
  for (int j = threadIdx.x; j < kAucThresholds; j += blockDim.x)
    thr[j] = j == 0 ? -1e-7f
                    : (j == kAucThresholds - 1 ? static_cast<float>(1.0 + 1e-7)
                                               : static_cast<float>(static_cast<double>(j) / (kAucThresholds - 1)));
  __syncthreads();
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float x[kMaxClasses];
    float mx = -3.4e38f;
    int am = 0;
    for (int c = 0; c < K; ++c) {
      x[c] = logits[v * K + c];
      if (x[c] > mx) {  // strict > : lowest index wins ties (tf.argmax)
        mx = x[c];
        am = c;
      }
    }
    const int t = labels[v];
    atomicAdd(&cm[((t >= 0 && t < K) ? t : K) * K + am], 1u);
    if (!want_auc) continue;
    float se = 0.f;   // same arithmetic as softmax_loss_fwd_kernel: the probabilities are the ones vnb_forward returns
    for (int c = 0; c < K; ++c) {
      x[c] = expf(x[c] - mx);
      se += x[c];
    }
    const float inv = 1.0f / se;
    for (int c = 1; c < K; ++c) {   // model.py:601-603 skips class 0
      const float pc = x[c] * inv;
      int b = static_cast<int>(pc * static_cast<float>(kAucThresholds - 1)) + 1;   // first guess, then exact against the table
      b = b < 0 ? 0 : (b > kAucThresholds ? kAucThresholds : b);
      while (b < kAucThresholds && pc > thr[b]) ++b;
      while (b > 0 && !(pc > thr[b - 1])) --b;
      atomicAdd(&hist[(c * 2 + (t == c ? 1 : 0)) * kAucBins + b], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (K + 1) * K; i += blockDim.x)
    if (cm[i]) atomicAdd(&confusion[i], static_cast<unsigned long long>(cm[i]));
  if (want_auc)
    for (int i = threadIdx.x; i < K * 2 * kAucBins; i += blockDim.x)
      if (hist[i]) atomicAdd(&auc_hist[i], static_cast<unsigned long long>(hist[i]));
}

// single block: sums partials, evaluates the configured loss and d(loss)/d(I,L,X) per (n,c)
// terms_out [N][K][4]; coef_out [N][K][3] = (dLoss/dI, dLoss/dL, dLoss/dX-per-voxel-weight)
__global__ void loss_finalize_kernel(const double* __restrict__ partial, int N, int nblk, long long Vn,
                                     LossCfg cfg, double* __restrict__ terms_out, float* __restrict__ coef_out,
                                     float* __restrict__ loss_out, const double* __restrict__ att_partial = nullptr,
                                     int att_nblk = 0, double att_div = 1.0) {
  const int K = cfg.K;
  const int pairs = N * K;
  for (int pr = threadIdx.x; pr < pairs; pr += blockDim.x) {
    const int n = pr / K, c = pr % K;
    for (int q = 0; q < 4; ++q) {
      double s = 0.0;
      for (int b = 0; b < nblk; ++b) s += partial[((static_cast<size_t>(n) * nblk + b) * K + c) * 4 + q];
      terms_out[(static_cast<size_t>(n) * K + c) * 4 + q] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double loss = 0.0, dice_part = 0.0;
  const double s = cfg.smooth;
  for (int i = 0; i < pairs * 3; ++i) coef_out[i] = 0.f;
  if (cfg.use_dice) {
    double dice = 0.0;
    if (cfg.fg_only) {  // train.py:373-377: dice_coe(softmax[...,1:2], labels) averaged over the batch only
      for (int n = 0; n < N; ++n) {
        const double* t = terms_out + (static_cast<size_t>(n) * K + 1) * 4;
        const double num = 2.0 * t[0] + s, den = t[1] + t[2] + s;
        dice += num / den / N;
        float* co = coef_out + (static_cast<size_t>(n) * K + 1) * 3;
        co[0] = static_cast<float>(-(2.0 / den) / N);
        co[1] = static_cast<float>((num / (den * den)) / N);
      }
    } else if (cfg.weighted_dice) {  // model.py:70-75
      for (int n = 0; n < N; ++n) {
        double num = 0.0, den = 0.0;
        for (int c = 0; c < K; ++c) {
          const double* t = terms_out + (static_cast<size_t>(n) * K + c) * 4;
          num += 2.0 * cfg.w[c] * t[0] + s;
          den += cfg.w[c] * (t[1] + t[2]) + s;
        }
        dice += num / den / N;
        for (int c = 0; c < K; ++c) {
          float* co = coef_out + (static_cast<size_t>(n) * K + c) * 3;
          co[0] = static_cast<float>(-(2.0 * cfg.w[c] / den) / N);          // d(1-dice)/dI
          co[1] = static_cast<float>((num * cfg.w[c] / (den * den)) / N);   // d(1-dice)/dL
        }
      }
    } else {  // model.py:82-83
      for (int n = 0; n < N; ++n)
        for (int c = 0; c < K; ++c) {
          const double* t = terms_out + (static_cast<size_t>(n) * K + c) * 4;
          const double num = 2.0 * t[0] + s, den = t[1] + t[2] + s;
          dice += num / den / pairs;
          float* co = coef_out + (static_cast<size_t>(n) * K + c) * 3;
          co[0] = static_cast<float>(-(2.0 / den) / pairs);
          co[1] = static_cast<float>((num / (den * den)) / pairs);
        }
    }
    loss += 1.0 - dice;
    dice_part = 1.0 - dice;
  }
  if (cfg.use_xent) {  // reduce_mean over all N*V voxels (model.py:91,496)
    double xs = 0.0;
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < K; ++c) {
        const double wc = cfg.weighted_xent ? cfg.w[c] : 1.0;
        xs += wc * terms_out[(static_cast<size_t>(n) * K + c) * 4 + 3];
        coef_out[(static_cast<size_t>(n) * K + c) * 3 + 2] =
            static_cast<float>(cfg.xent_alpha * wc / (static_cast<double>(N) * Vn));
      }
    loss += cfg.xent_alpha * xs / (static_cast<double>(N) * Vn);
  }
  double att = 0.0;  // attention loss (train.py:383-399), partial sums from gate_fwd_kernel
  for (int b = 0; b < att_nblk; ++b) att += att_partial[b];
  att /= att_div;
  loss_out[0] = static_cast<float>(loss + att);  // train.py:417 total_loss_op
  loss_out[1] = static_cast<float>(loss);
  loss_out[2] = static_cast<float>(att);
  loss_out[3] = static_cast<float>(dice_part);   // summary '1.dice' of the mixed losses (model.py:529,537,545,553)
}

// dL/dlogits, written to `dlogits` (the output layer's activation-gradient buffer)
__global__ void softmax_loss_bwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                        long long Vn, LossCfg cfg, const float* __restrict__ coef,
                                        float grad_scale, float* __restrict__ dlogits) {
  const int K = cfg.K, n = blockIdx.y;
  float cI[kMaxClasses], cL[kMaxClasses], cX[kMaxClasses];
  for (int c = 0; c < K; ++c) {
    cI[c] = coef[(static_cast<size_t>(n) * K + c) * 3 + 0];
    cL[c] = coef[(static_cast<size_t>(n) * K + c) * 3 + 1];
    cX[c] = coef[(static_cast<size_t>(n) * K + c) * 3 + 2];
  }
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < Vn;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long gv = static_cast<long long>(n) * Vn + v;
    float p[kMaxClasses];
    float mx = -3.4e38f;
    for (int c = 0; c < K; ++c) {
      p[c] = logits[gv * K + c];
      mx = p[c] > mx ? p[c] : mx;
    }
    float se = 0.f;
    for (int c = 0; c < K; ++c) {
      p[c] = expf(p[c] - mx);
      se += p[c];
    }
    const float inv = 1.0f / se;
    const int t = labels[gv];
    float dp[kMaxClasses];
    float dot = 0.f;
    for (int c = 0; c < K; ++c) {
      p[c] *= inv;
      dp[c] = (t == c ? cI[c] : 0.f) + cL[c] * (cfg.jaccard ? 2.f * p[c] : 1.f);
      dot += p[c] * dp[c];
    }
    const bool valid = t >= 0 && t < K;
    for (int c = 0; c < K; ++c) {
      float gl = p[c] * (dp[c] - dot);
      if (valid) gl += cX[t] * (p[c] - (t == c ? 1.f : 0.f));
      dlogits[gv * K + c] = gl * grad_scale;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Attention gating of the legacy trainer (train.py:281-302,383-399):
//   s = softmax(logits_attention);  logits_masked = (1 + s) * logits_vnet
//   attention loss  l2: mean(100 * (s[...,1] - distmap)^2)     abs: mean(|s - [1 - distmap, distmap]|)
// forward writes s and logits_masked and the per-block attention-loss partial sums; backward
// turns dL/dlogits_masked (+ the attention-loss term) into dL/dlogits_attention and dL/dlogits_vnet.
// ---------------------------------------------------------------------------------------------
struct GateArgs {
  const float* att;      // [V][K] logits_attention
  const float* vnet;     // [V][K] logits_vnet
  float* soft;           // [V][K] softmax_attention
  float* masked;         // [V][K] logits_masked
  const float* distmap;  // [V] or nullptr
  long long V;
  int K;
  int att_loss;          // 0 none, 1 l2, 2 abs
};

__device__ __forceinline__ void gate_softmax(const float* __restrict__ a, int K, float (&s)[kMaxClasses]) {
  float mx = -3.4e38f;
  for (int c = 0; c < K; ++c) {
    s[c] = a[c];
    mx = s[c] > mx ? s[c] : mx;
  }
  float se = 0.f;
  for (int c = 0; c < K; ++c) {
    s[c] = expf(s[c] - mx);
    se += s[c];
  }
  const float inv = 1.0f / se;
  for (int c = 0; c < K; ++c) s[c] *= inv;
}

__global__ void gate_fwd_kernel(GateArgs p, double* __restrict__ partial) {
  double acc = 0.0;
  const int K = p.K;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < p.V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s[kMaxClasses];
    gate_softmax(p.att + v * K, K, s);
    for (int c = 0; c < K; ++c) {
      p.soft[v * K + c] = s[c];
      p.masked[v * K + c] = (1.0f + s[c]) * p.vnet[v * K + c];
    }
    if (p.distmap && p.att_loss == 1) {
      const float e = s[1] - p.distmap[v];
      acc += static_cast<double>(e * e * 100.0f);
    } else if (p.distmap && p.att_loss == 2) {
      const float d = p.distmap[v];
      acc += static_cast<double>(fabsf(s[0] - (1.0f - d)) + fabsf(s[1] - d));
    }
  }
  __shared__ double red[kRedThreads];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (unsigned i = 0; i < blockDim.x; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}

// att_coef = 1 / (number of elements the attention loss averages over)
__global__ void gate_bwd_kernel(GateArgs p, const float* __restrict__ dmasked, float* __restrict__ datt, int acc_att,
                                float* __restrict__ dvnet, int acc_vnet, float att_coef) {
  const int K = p.K;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < p.V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s[kMaxClasses], ds[kMaxClasses];
    for (int c = 0; c < K; ++c) s[c] = p.soft[v * K + c];
    float dot = 0.f;
    for (int c = 0; c < K; ++c) {
      const float dm = dmasked[v * K + c], x = p.vnet[v * K + c];
      ds[c] = x * dm;
      const float dv = (1.0f + s[c]) * dm;
      float* o = dvnet + v * K + c;
      *o = acc_vnet ? *o + dv : dv;
    }
    if (p.distmap && p.att_loss == 1) {
      ds[1] += 200.0f * (s[1] - p.distmap[v]) * att_coef;
    } else if (p.distmap && p.att_loss == 2) {
      const float d = p.distmap[v];
      const float e0 = s[0] - (1.0f - d), e1 = s[1] - d;
      ds[0] += (e0 > 0.f ? 1.f : (e0 < 0.f ? -1.f : 0.f)) * att_coef;
      ds[1] += (e1 > 0.f ? 1.f : (e1 < 0.f ? -1.f : 0.f)) * att_coef;
    }
    for (int c = 0; c < K; ++c) dot += s[c] * ds[c];
    for (int c = 0; c < K; ++c) {
      const float g = s[c] * (ds[c] - dot);
      float* o = datt + v * K + c;
      *o = acc_att ? *o + g : g;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// optimiser over the flat parameter buffer (model.py:649-660).  Adam in TF's epsilon-hat form:
//   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m += (g-m)(1-b1);  v += (g^2-v)(1-b2);
//   p -= lr_t * m / (sqrt(v) + eps)
// `gscale` folds the data-parallel gradient average (1/world) into the same pass.
// ---------------------------------------------------------------------------------------------
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                                 float gscale) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.0f - b2);
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, long long n, float lr,
                                float gscale) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    p[i] -= lr * g[i] * gscale;
}

// [V][C] fp32 -> [V][Cpad] bf16 (hi, lo) with zero padding channels: multi-modal network input for the
// tensor-core input convolution (which wants channel multiples of 16)
__global__ void split_pad_bf16_kernel(const float* __restrict__ x, long long V, int C, int Cpad, uint16_t* __restrict__ hi,
                                      uint16_t* __restrict__ lo) {
  const long long total = V * Cpad;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const float f = c < C ? x[(i / Cpad) * C + c] : 0.f;
    const uint16_t h = f32_to_bf16(f);
    hi[i] = h;
    if (lo) lo[i] = f32_to_bf16(f - bf16_to_f32(h));
  }
}

// tf.train.MomentumOptimizer: accum = momentum*accum + g; p -= lr*accum   (use_nesterov: p -= lr*(g + momentum*accum))
__global__ void momentum_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ accum, long long n,
                                     float lr, float momentum, int nesterov, float gscale) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gscale;
    const float a = momentum * accum[i] + gi;
    accum[i] = a;
    p[i] -= nesterov ? lr * (gi + momentum * a) : lr * a;
  }
}

// fp32 -> bf16 hi/lo split of an arbitrary buffer (used for the network input and packed weights)
__global__ void split_bf16_kernel(const float* __restrict__ x, long long n, uint16_t* __restrict__ hi,
                                  uint16_t* __restrict__ lo) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float f = x[i];
    const uint16_t h = f32_to_bf16(f);
    hi[i] = h;
    if (lo) lo[i] = f32_to_bf16(f - bf16_to_f32(h));
  }
}

}  // namespace vnb

// =================================================================================================
// Vectorised (4 channels per thread, 128-bit accesses) variants of the hot per-element passes.
// Preconditions (checked by the engine, which falls back to the scalar kernels otherwise):
//   C % 4 == 0, (C/4) divides 256, total elements < 2^31.
// =================================================================================================
namespace vnb {

VNB_HD uint32_t pack_bf16x2(float a, float b) {
  return static_cast<uint32_t>(f32_to_bf16(a)) | (static_cast<uint32_t>(f32_to_bf16(b)) << 16);
}

__device__ __forceinline__ void store_hi_lo4(uint16_t* hi, uint16_t* lo, unsigned i4, const float (&y)[4]) {
  uint16_t h[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = f32_to_bf16(y[k]);
  reinterpret_cast<uint2*>(hi)[i4] = make_uint2(h[0] | (static_cast<uint32_t>(h[1]) << 16), h[2] | (static_cast<uint32_t>(h[3]) << 16));
  if (lo)
    reinterpret_cast<uint2*>(lo)[i4] = make_uint2(pack_bf16x2(y[0] - bf16_to_f32(h[0]), y[1] - bf16_to_f32(h[1])),
                                                  pack_bf16x2(y[2] - bf16_to_f32(h[2]), y[3] - bf16_to_f32(h[3])));
}

__global__ void __launch_bounds__(256) bn_apply_v4_kernel(ApplyArgs p) {
  pdl_trigger();   // see vnb_cuda.h: the next short pass may become resident now ...
  pdl_wait();      // ... and this one reads nothing of its predecessor before that grid has completed
  const unsigned total4 = static_cast<unsigned>(p.total / 4), C4 = static_cast<unsigned>(p.C / 4);
  const float keep_scale = dropout_keep_scale(p.drop_rate);
  const uint32_t dkey = dropout_key(p.seed, p.unit), dthr = dropout_threshold(p.drop_rate);
  const float4* __restrict__ z4 = reinterpret_cast<const float4*>(p.z);
  float4* __restrict__ a4 = reinterpret_cast<float4*>(p.a);
#pragma unroll 4
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const unsigned c = (i & (C4 - 1u)) * 4;  // C4 divides 256, hence a power of two
    float z[4];
    if (p.tiled_input) {
      const float v = p.z[i / C4];
      z[0] = z[1] = z[2] = z[3] = v;
    } else {
      const float4 f = z4[i];
      z[0] = f.x; z[1] = f.y; z[2] = f.z; z[3] = f.w;
    }
    const float4 sc = *reinterpret_cast<const float4*>(p.scale + c);
    const float4 sh = *reinterpret_cast<const float4*>(p.shift + c);
    float y[4] = {sc.x * z[0] + sh.x, sc.y * z[1] + sh.y, sc.z * z[2] + sh.z, sc.w * z[3] + sh.w};
    if (p.alpha) {
      const float4 al = *reinterpret_cast<const float4*>(p.alpha + c);
      y[0] = y[0] > 0.f ? y[0] : al.x * y[0];
      y[1] = y[1] > 0.f ? y[1] : al.y * y[1];
      y[2] = y[2] > 0.f ? y[2] : al.z * y[2];
      y[3] = y[3] > 0.f ? y[3] : al.w * y[3];
    }
    if (p.drop_rate > 0.f) {
      const uint32_t b0 = dropout_bits(dkey, static_cast<uint64_t>(i) * 4), b1 = dropout_bits(dkey, static_cast<uint64_t>(i) * 4 + 2);
      y[0] = (b0 & 0xFFFFu) >= dthr ? y[0] * keep_scale : 0.f;
      y[1] = (b0 >> 16) >= dthr ? y[1] * keep_scale : 0.f;
      y[2] = (b1 & 0xFFFFu) >= dthr ? y[2] * keep_scale : 0.f;
      y[3] = (b1 >> 16) >= dthr ? y[3] * keep_scale : 0.f;
    }
    a4[i] = make_float4(y[0], y[1], y[2], y[3]);
    if (p.a_hi) store_hi_lo4(p.a_hi, p.a_lo, i, y);
  }
}

// g for 4 consecutive channels (same semantics as bwd_g)
__device__ __forceinline__ void bwd_g4(const BwdArgs& p, unsigned i, unsigned c, const float (&z)[4], float (&g)[4],
                                       float (&yhat)[4], float (&dd)[4]) {
  const float4 sc = *reinterpret_cast<const float4*>(p.scale + c);
  const float4 sh = *reinterpret_cast<const float4*>(p.shift + c);
  const float4 d4 = reinterpret_cast<const float4*>(p.d)[i];
  yhat[0] = sc.x * z[0] + sh.x; yhat[1] = sc.y * z[1] + sh.y; yhat[2] = sc.z * z[2] + sh.z; yhat[3] = sc.w * z[3] + sh.w;
  dd[0] = d4.x; dd[1] = d4.y; dd[2] = d4.z; dd[3] = d4.w;
  if (p.drop_rate > 0.f) {
    const float inv = dropout_keep_scale(p.drop_rate);
    const uint32_t dkey = dropout_key(p.seed, p.unit), dthr = dropout_threshold(p.drop_rate);
    const uint32_t b0 = dropout_bits(dkey, static_cast<uint64_t>(i) * 4), b1 = dropout_bits(dkey, static_cast<uint64_t>(i) * 4 + 2);
    dd[0] = (b0 & 0xFFFFu) >= dthr ? dd[0] * inv : 0.f;
    dd[1] = (b0 >> 16) >= dthr ? dd[1] * inv : 0.f;
    dd[2] = (b1 & 0xFFFFu) >= dthr ? dd[2] * inv : 0.f;
    dd[3] = (b1 >> 16) >= dthr ? dd[3] * inv : 0.f;
  }
  if (p.alpha) {
    const float4 al = *reinterpret_cast<const float4*>(p.alpha + c);
    const float a[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = yhat[k] > 0.f ? dd[k] : (yhat[k] < 0.f ? dd[k] * a[k] : 0.f);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = dd[k];
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_v4_kernel(BwdArgs p, long long total) {
  pdl_trigger();   // see vnb_cuda.h: the next short pass may become resident now ...
  pdl_wait();      // ... and this one reads nothing of its predecessor before that grid has completed
  const unsigned total4 = static_cast<unsigned>(total / 4), C4 = static_cast<unsigned>(p.C / 4);
  const float4* __restrict__ z4 = reinterpret_cast<const float4*>(p.z);
#pragma unroll 2
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const unsigned c = (i & (C4 - 1u)) * 4;
    const float4 zf = z4[i];
    const float z[4] = {zf.x, zf.y, zf.z, zf.w};
    float g[4], yhat[4], dd[4];
    bwd_g4(p, i, c, z, g, yhat, dd);
    const float4 P = *reinterpret_cast<const float4*>(p.P + c);
    const float4 Q = *reinterpret_cast<const float4*>(p.Q + c);
    const float4 S = *reinterpret_cast<const float4*>(p.S + c);
    float dz[4];
    dz[0] = P.x * g[0] + Q.x + S.x * (z[0] - static_cast<float>(p.mean[c]));
    dz[1] = P.y * g[1] + Q.y + S.y * (z[1] - static_cast<float>(p.mean[c + 1]));
    dz[2] = P.z * g[2] + Q.z + S.z * (z[2] - static_cast<float>(p.mean[c + 2]));
    dz[3] = P.w * g[3] + Q.w + S.w * (z[3] - static_cast<float>(p.mean[c + 3]));
    reinterpret_cast<float4*>(p.d)[i] = make_float4(dz[0], dz[1], dz[2], dz[3]);
    if (p.res_grad) {
      float4* r = reinterpret_cast<float4*>(p.res_grad) + i;
      float4 o = make_float4(dz[0], dz[1], dz[2], dz[3]);
      if (p.res_accumulate) {
        const float4 old = *r;
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *r = o;
    }
    if (p.res_grad2) {
      float4* r = reinterpret_cast<float4*>(p.res_grad2) + i;
      float4 o = make_float4(dz[0], dz[1], dz[2], dz[3]);
      if (p.res_accumulate2) {
        const float4 old = *r;
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *r = o;
    }
    if (p.d_hi) store_hi_lo4(p.d_hi, p.d_lo, i, dz);
  }
}

// z = a + b (the residual add that sits between the two batch norms of the legacy flavour, VNet.py:33-34)
__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ z, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    z[i] = a[i] + b[i];
}

// per-channel reductions, 4 channels per thread. blockDim = 256, C4 = C/4 divides 256.
// partial layout identical to the scalar kernels: [block][NQ][C] doubles.
template <int NQ>
__device__ __forceinline__ void block_channel_combine_v4(float (&acc)[NQ][4], unsigned C4, int C,
                                                         double* __restrict__ partial) {
  const unsigned t = threadIdx.x;
  if (C4 <= 32u) {
    // Thread t owns channel quad t % C4 = lane % C4 (C4 is a power of two): a butterfly over the lane offsets >= C4
    // leaves every lane with the warp's sum of its quad (fixed tree), then the eight warp sums are added in warp order
    // in double.  (The earlier form -- C4 threads walking 256 / C4 shared-memory entries each for every quantity -- made
    // the tail of a block as long as its main loop on the 16- and 32-channel tensors.)
    __shared__ float wred[8][NQ * 4][32];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v = acc[q][k];
        for (unsigned off = 16u; off >= C4; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, static_cast<int>(off));
        acc[q][k] = v;
      }
    const unsigned lane = t & 31u, w = t >> 5;
    if (lane < C4) {
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) wred[w][q * 4 + k][lane] = acc[q][k];
    }
    __syncthreads();
    for (unsigned idx = t; idx < C4 * NQ * 4u; idx += 256u) {   // one (quantity, channel) per thread and trip
      const unsigned quad = idx % C4, qk = idx / C4;
      double s = 0.0;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) s += static_cast<double>(wred[ww][qk][quad]);
      partial[(static_cast<size_t>(blockIdx.x) * NQ + qk / 4) * C + quad * 4 + (qk & 3u)] = s;
    }
    return;
  }
  __shared__ float red[NQ * 4][256];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int k = 0; k < 4; ++k) red[q * 4 + k][t] = acc[q][k];
  __syncthreads();
  if (t < C4) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double s = 0.0;
        for (unsigned g = t; g < 256; g += C4) s += red[q * 4 + k][g];
        partial[(static_cast<size_t>(blockIdx.x) * NQ + q) * C + t * 4 + k] = s;
      }
  }
}

__global__ void __launch_bounds__(256) bn_stats_v4_kernel(const float* __restrict__ z, int C, unsigned total4,
                                                          double* __restrict__ partial) {
  pdl_trigger();   // see vnb_cuda.h: the next short pass may become resident now ...
  pdl_wait();      // ... and this one reads nothing of its predecessor before that grid has completed
  const unsigned C4 = static_cast<unsigned>(C / 4);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  // thread t keeps channel quad t % C4 because the stride gridDim*256 is a multiple of C4
#pragma unroll 4
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total4; i += gridDim.x * 256u) {
    const float4 f = reinterpret_cast<const float4*>(z)[i];
    acc[0][0] += f.x; acc[0][1] += f.y; acc[0][2] += f.z; acc[0][3] += f.w;
    acc[1][0] += f.x * f.x; acc[1][1] += f.y * f.y; acc[1][2] += f.z * f.z; acc[1][3] += f.w * f.w;
  }
  block_channel_combine_v4<2>(acc, C4, C, partial);
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_v4_kernel(BwdArgs p, unsigned total4, double* __restrict__ partial) {
  pdl_trigger();   // see vnb_cuda.h: the next short pass may become resident now ...
  pdl_wait();      // ... and this one reads nothing of its predecessor before that grid has completed
  const unsigned C4 = static_cast<unsigned>(p.C / 4);
  const unsigned c = (threadIdx.x % C4) * 4;
  const float mu[4] = {static_cast<float>(p.mean[c]), static_cast<float>(p.mean[c + 1]), static_cast<float>(p.mean[c + 2]),
                       static_cast<float>(p.mean[c + 3])};
  float acc[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 2
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total4; i += gridDim.x * 256u) {
    float z[4];
    if (p.tiled_input) {
      const float v = p.z[i / C4];
      z[0] = z[1] = z[2] = z[3] = v;
    } else {
      const float4 zf = reinterpret_cast<const float4*>(p.z)[i];
      z[0] = zf.x; z[1] = zf.y; z[2] = zf.z; z[3] = zf.w;
    }
    float g[4], yhat[4], dd[4];
    bwd_g4(p, i, c, z, g, yhat, dd);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[0][k] += g[k];
      acc[1][k] += g[k] * (z[k] - mu[k]);
      if (p.alpha && yhat[k] < 0.f) acc[2][k] += dd[k] * yhat[k];
    }
  }
  block_channel_combine_v4<3>(acc, C4, p.C, partial);
}

}  // namespace vnb

// =================================================================================================
// Sliding-window evaluation (model.py:866-937): window gather, softmax accumulation, final argmax.
// The volume, the un-normalised softmax sums and the hit counts stay on the device for a whole case.
// =================================================================================================
namespace vnb {

struct WindowGeom {
  int X, Y, Z;      // (padded) volume extents
  int PX, PY, PZ;   // patch extents
  int M, K;         // modalities, classes
};

// images[b][x][y][z][m] = volume[sx+x][sy+y][sz+z][m] for the windows of one batch (model.py:895-903)
__global__ void window_gather_kernel(const float* __restrict__ vol, WindowGeom g, const int* __restrict__ starts, int nwin,
                                     float* __restrict__ images) {
  const long long per = static_cast<long long>(g.PX) * g.PY * g.PZ * g.M;
  const long long total = per * nwin;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per);
    long long r = i % per;
    const int m = static_cast<int>(r % g.M);
    r /= g.M;
    const int z = static_cast<int>(r % g.PZ);
    r /= g.PZ;
    const int y = static_cast<int>(r % g.PY);
    const int x = static_cast<int>(r / g.PY);
    const int sx = starts[3 * b], sy = starts[3 * b + 1], sz = starts[3 * b + 2];
    images[i] = vol[((static_cast<long long>(sx + x) * g.Y + sy + y) * g.Z + sz + z) * g.M + m];
  }
}

// softmax_np[c][window] += softmax[j,...,c]; weight_np[window] += 1 (model.py:919-929) for ONE window: windows of a
// batch overlap, so they are accumulated by consecutive launches in window order -- the same order of float
// additions per voxel as the reference's NumPy loop
__global__ void window_accumulate_kernel(const float* __restrict__ softmax, WindowGeom g, int sx, int sy, int sz,
                                         float* __restrict__ sum, float* __restrict__ weight) {
  const long long per = static_cast<long long>(g.PX) * g.PY * g.PZ;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < per;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int z = static_cast<int>(r % g.PZ);
    r /= g.PZ;
    const int y = static_cast<int>(r % g.PY);
    const int x = static_cast<int>(r / g.PY);
    const long long v = (static_cast<long long>(sx + x) * g.Y + sy + y) * g.Z + sz + z;
    for (int c = 0; c < g.K; ++c) sum[v * g.K + c] += softmax[i * g.K + c];
    weight[v] += 1.0f;
  }
}

// label = argmax_c of the accumulated (un-normalised) sums, lowest index on ties (np.argmax, model.py:934)
__global__ void volume_argmax_kernel(const float* __restrict__ sum, long long V, int K, long long* __restrict__ label) {
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float mx = sum[v * K];
    int am = 0;
    for (int c = 1; c < K; ++c) {
      const float x = sum[v * K + c];
      if (x > mx) {
        mx = x;
        am = c;
      }
    }
    label[v] = am;
  }
}

}  // namespace vnb
