// fx_models.cuh -- D(theta), D'(theta), D''(theta) for the tagged diffusivity
// models, in closed form, evaluated in registers (fp64).
//
// The reference obtains D' and D'' by jax autodiff of the model definitions
// (src/frontx/_boltzmann.py:29, src/frontx/models.py:95-100); here they are
// hand-derived logarithmic-derivative forms that share one log per base
// (log Se, log(1-Se)) and use exp() for every power, so that a van Genuchten
// evaluation costs 2 log + 3 exp instead of 5+ pow calls.
//
//   Se = (theta - theta_r)/(theta_s - theta_r)          models.py:17-24
//   LETd            D = Dwt Se^L/(Se^L + E (1-Se)^T)    models.py:62-66
//   Brooks-Corey    D = Ks/(alpha n dT) Se^(1/n+l+1)    models.py:160-174 with D=K/C (:88-100)
//   van Genuchten   D = (1-m)Ks/(alpha m dT) Se^(l-1/m) (1-v)^2/v, v=(1-Se^(1/m))^m   :217-253
//   LETxs           D = Ks/(alpha dT) kr P Q (Ls/(1-Se)+Ts/Se)/(P+Q)^2               :298-315
#pragma once

#include "../../include/frontx_b200.h"

namespace fx {

constexpr int NDERIVED = 10;

// Per-problem derived constants, computed once when a thread picks a problem up.
template <int MODEL>
__device__ __forceinline__ void derive_params(const double* __restrict__ p, double (&d)[NDERIVED]) {
#pragma unroll
  for (int k = 0; k < NDERIVED; k++) d[k] = 0.0;
  if constexpr (MODEL == FX_CONSTANT) {
    d[0] = p[0];
  } else if constexpr (MODEL == FX_POWER_LAW) {
    d[0] = p[0];  // a
    d[1] = p[1];  // k
    d[2] = p[2];  // eps
    // pow(theta<0, k) is real only for integer k: keep pow's sign rule
    double kr = rint(p[1]);
    bool is_int = (kr == p[1]) && fabs(p[1]) < 9.0e15;
    bool is_odd = is_int && (fmod(fabs(kr), 2.0) == 1.0);
    d[3] = is_int ? (is_odd ? -1.0 : 1.0) : __longlong_as_double(0x7ff8000000000000LL);
  } else if constexpr (MODEL == FX_LOG) {
    d[0] = p[0];
    d[1] = p[1];
  } else if constexpr (MODEL == FX_BROOKS_COREY) {
    double n = p[0], l = p[1], Ks = p[2], alpha = p[3], tr = p[4], ts = p[5];
    double dT = ts - tr;
    d[0] = 1.0 / n + l + 1.0;        // exponent e
    d[1] = Ks / (alpha * n * dT);    // Cc
    d[2] = tr;
    d[3] = dT;
    d[4] = 1.0 / dT;
  } else if constexpr (MODEL == FX_VAN_GENUCHTEN) {
    double m = p[1], l = p[2], Ks = p[3], alpha = p[4], tr = p[5], ts = p[6];
    double dT = ts - tr;
    d[0] = 1.0 / m;
    d[1] = m;
    d[2] = l - 1.0 / m;                          // a
    d[3] = (1.0 - m) * Ks / (alpha * m * dT);    // Cc  (1/n = 1-m)
    d[4] = tr;
    d[5] = dT;
    d[6] = 1.0 / dT;
  } else if constexpr (MODEL == FX_LETD) {
    double tr = p[4], ts = p[5];
    double dT = ts - tr;
    d[0] = p[0];       // L
    d[1] = log(p[1]);  // ln E
    d[2] = p[2];       // T
    d[3] = p[3];       // Dwt
    d[4] = tr;
    d[5] = dT;
    d[6] = 1.0 / dT;
  } else if constexpr (MODEL == FX_LETXS) {
    double Ks = p[6], alpha = p[7], tr = p[8], ts = p[9];
    double dT = ts - tr;
    d[0] = p[0];        // Lw
    d[1] = log(p[1]);   // ln Ew
    d[2] = p[2];        // Tw
    d[3] = p[3];        // Ls
    d[4] = log(p[4]);   // ln Es
    d[5] = p[5];        // Ts
    d[6] = Ks / (alpha * dT);  // Cc
    d[7] = tr;
    d[8] = dT;
    d[9] = 1.0 / dT;
  }
}

// D, D' and (when want2) D'' at theta.  NaN outside the model's domain, like
// pow() of a negative base in the reference.
template <int MODEL>
__device__ __forceinline__ void eval_D(const double (&d)[NDERIVED], double theta, bool want2,
                                       double& D, double& D1, double& D2) {
  D2 = 0.0;
  if constexpr (MODEL == FX_CONSTANT) {
    D = d[0];
    D1 = 0.0;
  } else if constexpr (MODEL == FX_POWER_LAW) {
    double a = d[0], k = d[1], eps = d[2];
    double mag = exp(k * log(fabs(theta)));
    double sgn = (theta < 0.0) ? d[3] : 1.0;  // NaN for non-integer k
    double pw = a * sgn * mag;
    double it = 1.0 / theta;
    D = pw + eps;
    D1 = k * pw * it;
    if (want2) D2 = (k - 1.0) * D1 * it;
  } else if constexpr (MODEL == FX_LOG) {
    double it = 1.0 / theta;
    D = d[0] + d[1] * log(theta);
    D1 = d[1] * it;
    if (want2) D2 = -D1 * it;
  } else if constexpr (MODEL == FX_BROOKS_COREY) {
    double e = d[0], Cc = d[1], tr = d[2], dT = d[3], idT = d[4];
    double Se = (theta - tr) / dT;
    double L = log(Se);
    double iSe = idT / Se;  // (1/Se) * dSe/dtheta
    D = Cc * exp(e * L);
    D1 = D * e * iSe;
    if (want2) D2 = D1 * (e - 1.0) * iSe;
  } else if constexpr (MODEL == FX_VAN_GENUCHTEN) {
    double im = d[0], m = d[1], a = d[2], Cc = d[3], tr = d[4], dT = d[5], idT = d[6];
    double Se = (theta - tr) / dT;
    double L = log(Se);
    // w = 1 - Se^(1/m) has condition number ~1e7 at b = theta_s - 1e-7 (the reference's own
    // test point); form it without cancellation so that implementations agree to ~1e-15
    double w = -expm1(im * L);
    double u = 1.0 - w;            // Se^(1/m)
    double v = exp(m * log(w));    // (1 - Se^(1/m))^m
    double omv = 1.0 - v;
    double iSe = 1.0 / Se, iw = 1.0 / w, iomv = 1.0 / omv;
    double g = exp(a * L) * omv * omv / v;   // Se^a (1-v)^2 / v
    double q = -u * iw * iSe;                // v'/v
    double r = (1.0 + v) * iomv;
    double lg1 = a * iSe - q * r;            // (ln g)'
    D = Cc * g;
    D1 = D * lg1 * idT;
    if (want2) {
      double q1 = q * iSe * (im * iw - 1.0);           // q'
      double r1 = 2.0 * q * v * iomv * iomv;           // r'
      double lg2 = -a * iSe * iSe - q1 * r - q * r1;   // (ln g)''
      D2 = D * (lg2 + lg1 * lg1) * idT * idT;
    }
  } else if constexpr (MODEL == FX_LETD) {
    double Lx = d[0], lnE = d[1], T = d[2], Dwt = d[3], tr = d[4], dT = d[5], idT = d[6];
    double Se = (theta - tr) / dT;
    double oms = 1.0 - Se;
    double L1 = log(Se), L2 = log(oms);
    double R = exp(lnE + T * L2 - Lx * L1);  // E (1-Se)^T / Se^L
    double iSe = 1.0 / Se, ioms = 1.0 / oms;
    double rho = -T * ioms - Lx * iSe;       // R'/R
    double qq = 1.0 / (1.0 + R);
    double R1 = R * rho;
    D = Dwt * qq;
    D1 = -Dwt * R1 * qq * qq * idT;
    if (want2) {
      double rho1 = -T * ioms * ioms + Lx * iSe * iSe;
      double R2 = R * (rho1 + rho * rho);
      D2 = -Dwt * (R2 - 2.0 * R1 * R1 * qq) * qq * qq * idT * idT;
    }
  } else if constexpr (MODEL == FX_LETXS) {
    double Lw = d[0], lnEw = d[1], Tw = d[2], Ls = d[3], lnEs = d[4], Ts = d[5], Cc = d[6];
    double tr = d[7], dT = d[8], idT = d[9];
    double Se = (theta - tr) / dT;
    double oms = 1.0 - Se;
    double L1 = log(Se), L2 = log(oms);
    double R = exp(lnEw + Tw * L2 - Lw * L1);  // Ew (1-Se)^Tw / Se^Lw  -> kr = 1/(1+R)
    double S = exp(lnEs + Ts * L1 - Ls * L2);  // Q/P = Es Se^Ts / (1-Se)^Ls
    double iSe = 1.0 / Se, ioms = 1.0 / oms;
    double kr = 1.0 / (1.0 + R), qs = 1.0 / (1.0 + S);
    double A = R * kr, B = S * qs;
    double sig = Ls * ioms + Ts * iSe;
    double sig1 = Ls * ioms * ioms - Ts * iSe * iSe;
    double rho = -Tw * ioms - Lw * iSe;
    double isig = 1.0 / sig;
    double G1 = -rho * A + sig * (1.0 - 2.0 * B) + sig1 * isig;  // (ln D)' wrt Se
    D = Cc * kr * B * qs * sig;
    D1 = D * G1 * idT;
    if (want2) {
      double rho1 = -Tw * ioms * ioms + Lw * iSe * iSe;
      double sig2 = 2.0 * (Ls * ioms * ioms * ioms + Ts * iSe * iSe * iSe);
      double A1 = rho * A * kr;  // A' = rho A (1-A), 1-A = kr
      double B1 = sig * B * qs;  // B' = sig B (1-B)
      double G2 = -rho1 * A - rho * A1 + sig1 * (1.0 - 2.0 * B) - 2.0 * sig * B1 +
                  (sig2 * sig - sig1 * sig1) * isig * isig;
      D2 = D * (G2 + G1 * G1) * idT * idT;
    }
  }
}

}  // namespace fx
