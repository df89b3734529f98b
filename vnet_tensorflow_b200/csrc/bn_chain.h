// Per-channel closed form of the reference's batch-norm chains.
//
// Every conv unit of networks.VNet feeds its (bias + optional residual) output z through one of four
// chains of *training-mode* batch norms (reference networks.py, tf.layers.batch_normalization with
// momentum .99, eps 1e-3, batch statistics always):
//
//   CH_S  plain           y = BN0(z)                                      networks.py:319,279,293,303,347
//   CH_Q  "x + BN(x)"     li = BN0(z); y = BN1(z + li)                    networks.py:358-361 (last conv)
//   CH_D  dead BN         li = BN0(z) (unused); y = BN1(z)                networks.py:358,361 (middle convs)
//   CH_T  triple          u = BN0(z); li = BN1(u); y = BN2(u + li)        networks.py:334-337 (n == 1)
//   CH_2  double          y = BN1(BN0(z))                                 VNet.py:31-35 (legacy flavour, non-last convs)
//
// All intermediate tensors are per-channel affine functions  s*(z - mu) + m  of z, so the whole chain
// collapses to  y = A*(z - mu) + B  with A = A(sigma^2, gamma_k) and B = beta_last, and each BN's own
// batch mean / variance (needed for the moving-average UPDATE_OPS, model.py:665-666) follows in
// closed form.  The backward pass needs dA/dsigma^2 and dA/dgamma_k; they are obtained with forward
// mode dual numbers so the four chain formulas are written exactly once.
//
//   dL/dz_i      = A*(g_i - R0/n) + (2/n) * dA/dsigma^2 * R1 * (z_i - mu)      R0 = sum g, R1 = sum g*(z-mu)
//   dL/dgamma_k  = R1 * dA/dgamma_k          dL/dbeta_last = R0          (other betas: exactly 0)
//
// with g = dL/dy.  Host+device, double precision (per-channel work only).
#pragma once
#include "vnb_cuda.h"

namespace vnb {

enum ChainType : int { CH_S = 0, CH_Q = 1, CH_D = 2, CH_T = 3, CH_2 = 4 };

VNB_HD int chain_num_bn(int type) { return type == CH_S ? 1 : (type == CH_T ? 3 : 2); }

struct Dual {  // value + partials w.r.t. (sigma^2, gamma0, gamma1, gamma2)
  double v;
  double d[4];
};
VNB_HD Dual dual_const(double c) {
  Dual r;
  r.v = c;
  for (int i = 0; i < 4; ++i) r.d[i] = 0.0;
  return r;
}
VNB_HD Dual dual_var(double c, int which) {
  Dual r = dual_const(c);
  r.d[which] = 1.0;
  return r;
}
VNB_HD Dual operator+(const Dual& a, const Dual& b) {
  Dual r;
  r.v = a.v + b.v;
  for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
VNB_HD Dual operator*(const Dual& a, const Dual& b) {
  Dual r;
  r.v = a.v * b.v;
  for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
VNB_HD Dual dual_rsqrt_eps(const Dual& x, double eps) {  // (x + eps)^(-1/2)
  Dual r;
#if defined(__CUDA_ARCH__)
  r.v = rsqrt(x.v + eps);
#else
  r.v = 1.0 / __builtin_sqrt(x.v + eps);
#endif
  const double k = -0.5 * r.v * r.v * r.v;
  for (int i = 0; i < 4; ++i) r.d[i] = k * x.d[i];
  return r;
}

struct ChainOut {
  Dual A;            // scale on (z - mu), with partials
  int beta_idx;      // which BN's beta is the additive constant B
  double bn_mean[3]; // batch mean seen by BN k   (mu_z contribution added by the caller: see below)
  double bn_var[3];  // batch (biased) variance seen by BN k
  bool mean_is_mu[3];  // true: batch mean of BN k is mu_z (+ bn_mean[k] offset), false: bn_mean[k] only
};

// gamma/beta: the chain's BN parameters in graph order (unused entries ignored).
VNB_HD ChainOut chain_eval(int type, double sigma2, const double gamma[3], const double beta[3],
                           double eps) {
  ChainOut o;
  const Dual s2 = dual_var(sigma2, 0);
  const Dual g0 = dual_var(gamma[0], 1), g1 = dual_var(gamma[1], 2), g2 = dual_var(gamma[2], 3);
  const Dual r0 = dual_rsqrt_eps(s2, eps);
  for (int k = 0; k < 3; ++k) {
    o.bn_mean[k] = 0.0;
    o.bn_var[k] = 0.0;
    o.mean_is_mu[k] = false;
  }
  o.bn_var[0] = sigma2;
  o.mean_is_mu[0] = true;
  if (type == CH_S) {
    o.A = g0 * r0;
    o.beta_idx = 0;
  } else if (type == CH_D) {
    o.A = g1 * r0;
    o.beta_idx = 1;
    o.bn_var[1] = sigma2;
    o.mean_is_mu[1] = true;
  } else if (type == CH_Q) {
    const Dual a = dual_const(1.0) + g0 * r0;    // z + BN0(z) = a*(z-mu) + (mu + beta0)
    const Dual vt = a * a * s2;
    o.A = g1 * a * dual_rsqrt_eps(vt, eps);
    o.beta_idx = 1;
    o.bn_var[1] = vt.v;
    o.mean_is_mu[1] = true;
    o.bn_mean[1] = beta[0];
  } else if (type == CH_2) {
    const Dual s0 = g0 * r0;                      // u = s0*(z-mu) + beta0
    const Dual vu = s0 * s0 * s2;
    o.A = g1 * s0 * dual_rsqrt_eps(vu, eps);
    o.beta_idx = 1;
    o.bn_mean[1] = beta[0];
    o.bn_var[1] = vu.v;
  } else {  // CH_T
    const Dual s0 = g0 * r0;                      // u = s0*(z-mu) + beta0
    const Dual vu = s0 * s0 * s2;
    const Dual r1 = dual_rsqrt_eps(vu, eps);
    const Dual st = s0 * (dual_const(1.0) + g1 * r1);  // u + BN1(u) = st*(z-mu) + beta0 + beta1
    const Dual vt = st * st * s2;
    o.A = g2 * st * dual_rsqrt_eps(vt, eps);
    o.beta_idx = 2;
    o.bn_mean[1] = beta[0];
    o.bn_var[1] = vu.v;
    o.bn_mean[2] = beta[0] + beta[1];
    o.bn_var[2] = vt.v;
  }
  return o;
}

}  // namespace vnb
