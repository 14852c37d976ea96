// fx_models.cuh -- D(theta), D'(theta), D''(theta) for the tagged diffusivity
// models, in closed form, evaluated in registers (fp64).
//
// The reference obtains D' and D'' by jax autodiff of the model definitions
// (src/frontx/_boltzmann.py:29, src/frontx/models.py:95-100); here they are
// hand-derived logarithmic-derivative forms that share one log per base
// (log Se, log(1-Se)) and use exp() for every power, so that a van Genuchten
// evaluation costs 2 log + 3 exp + ONE reciprocal instead of 5+ pow calls and 7 divisions.
//
//   Se = (theta - theta_r)/(theta_s - theta_r)          models.py:17-24
//   LETd            D = Dwt Se^L/(Se^L + E (1-Se)^T)    models.py:62-66
//   Brooks-Corey    D = Ks/(alpha n dT) Se^(1/n+l+1)    models.py:160-174 with D=K/C (:88-100)
//   van Genuchten   D = (1-m)Ks/(alpha m dT) Se^(l-1/m) (1-v)^2/v, v=(1-Se^(1/m))^m   :217-253
//   LETxs           D = Ks/(alpha dT) kr P Q (Ls/(1-Se)+Ts/Se)/(P+Q)^2               :298-315
//
// Everything here is FX_HD (host + device) and uses fx_math.cuh only, so the CPU
// twin (oracle/fx_twin.cpp) evaluates bit-identical values; see fx_math.cuh.
#pragma once

#include "../../include/frontx_b200.h"
#include "fx_math.cuh"

namespace fx {

constexpr int NDERIVED = 12;

// Per-problem derived constants, computed once when a thread picks a problem up.
template <int MODEL>
FX_HD void derive_params(const double* p, double (&d)[NDERIVED], const MathTab& T) {
#pragma unroll
  for (int k = 0; k < NDERIVED; k++) d[k] = 0.0;
  if constexpr (MODEL == FX_CONSTANT) {
    d[0] = p[0];
    d[1] = fx_rcp(p[0]);
  } else if constexpr (MODEL == FX_POWER_LAW) {
    d[0] = p[0];  // a
    d[1] = p[1];  // k
    d[2] = p[2];  // eps
    // pow(theta<0, k) is real only for integer k: keep pow's sign rule
    double kr = rint(p[1]);
    bool is_int = (kr == p[1]) && fabs(p[1]) < 9.0e15;
    bool is_odd = is_int && (kr * 0.5 != rint(kr * 0.5));
    d[3] = is_int ? (is_odd ? -1.0 : 1.0) : fx_nan();
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
    d[5] = 1.0 / d[1];
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
    d[7] = 1.0 / d[3];
    d[8] = p[7];  // form of 1 - Se^(1/m): 0 the reference's literal subtraction, else cancellation-free
  } else if constexpr (MODEL == FX_LETD) {
    double tr = p[4], ts = p[5];
    double dT = ts - tr;
    d[0] = p[0];             // L
    d[1] = fx_log(p[1], T);  // ln E
    d[2] = p[2];             // T
    d[3] = p[3];             // Dwt
    d[4] = tr;
    d[5] = dT;
    d[6] = 1.0 / dT;
    d[7] = 1.0 / p[3];
  } else if constexpr (MODEL == FX_TABULATED) {
    d[0] = p[0];  // the table's address, bit-cast
    d[1] = p[1];  // number of knots
  } else if constexpr (MODEL == FX_LETXS) {
    double Ks = p[6], alpha = p[7], tr = p[8], ts = p[9];
    double dT = ts - tr;
    d[0] = p[0];              // Lw
    d[1] = fx_log(p[1], T);   // ln Ew
    d[2] = p[2];              // Tw
    d[3] = p[3];              // Ls
    d[4] = fx_log(p[4], T);   // ln Es
    d[5] = p[5];              // Ts
    d[6] = Ks / (alpha * dT);  // Cc
    d[7] = tr;
    d[8] = dT;
    d[9] = 1.0 / dT;
    d[10] = 1.0 / d[6];
  }
}

// D, D', (when want2) D'' and 1/D at theta.  NaN outside the model's domain, like pow() of a
// negative base in the reference.  All quotients of one evaluation are formed from ONE
// reciprocal (of the product of the denominators, unwound by multiplications): an fp64
// reciprocal costs about ten multiply-adds on sm_100a.
template <int MODEL>
FX_HD void eval_D(const double (&d)[NDERIVED], double theta, bool want2, double& D, double& D1,
                  double& D2, double& iD, const MathTab& T) {
  D2 = 0.0;
  if constexpr (MODEL == FX_CONSTANT) {
    D = d[0];
    D1 = 0.0;
    iD = d[1];
  } else if constexpr (MODEL == FX_POWER_LAW) {
    double a = d[0], k = d[1], eps = d[2];
    double mag = fx_exp(k * fx_log(fabs(theta), T), T);
    double sgn = (theta < 0.0) ? d[3] : 1.0;  // NaN for non-integer k
    double pw = a * sgn * mag;
    D = pw + eps;
    double ip = fx_rcp(theta * D);
    double it = D * ip;
    iD = theta * ip;
    D1 = k * pw * it;
    if (want2) D2 = (k - 1.0) * D1 * it;
  } else if constexpr (MODEL == FX_LOG) {
    D = fma(d[1], fx_log(theta, T), d[0]);
    double ip = fx_rcp(theta * D);
    double it = D * ip;
    iD = theta * ip;
    D1 = d[1] * it;
    if (want2) D2 = -D1 * it;
  } else if constexpr (MODEL == FX_BROOKS_COREY) {
    double e = d[0], Cc = d[1], tr = d[2], dT = d[3], idT = d[4], iCc = d[5];
    double Se = div_by(theta - tr, dT, idT);
    double L = fx_log(Se, T);
    double E = fx_exp(e * L, T);   // Se^e
    double ip = fx_rcp(Se * E);
    double iSe = idT * (E * ip);   // (1/Se) * dSe/dtheta
    D = Cc * E;
    iD = iCc * (Se * ip);
    D1 = D * e * iSe;
    if (want2) D2 = D1 * (e - 1.0) * iSe;
  } else if constexpr (MODEL == FX_VAN_GENUCHTEN) {
    double im = d[0], m = d[1], a = d[2], Cc = d[3], tr = d[4], dT = d[5], idT = d[6], iCc = d[7];
    double Se = div_by(theta - tr, dT, idT);
    double L = fx_log(Se, T);
    // w = 1 - Se^(1/m) has condition number ~1e7 at b = theta_s - 1e-7 (the reference's own test
    // point).  By default it is the LITERAL subtraction 1 - fx_exp(log(Se)/m), as models.py:245,253
    // write it (the rounding of Se^(1/m) is then ~1e-9 of w there, and differs between pow()
    // implementations); params[7] != 0 asks for the cancellation-free -expm1(log(Se)/m) instead.
    double w = -fx_expm1(im * L, T, d[8] != 0.0);
    double u = 1.0 - w;                     // Se^(1/m)
    double v = fx_exp(m * fx_log(w, T), T); // (1 - Se^(1/m))^m
    double omv = 1.0 - v;
    double E = fx_exp(a * L, T);            // Se^a
    // reciprocals of Se, w, 1-v, v, E from one
    double p2 = Se * w, p3 = p2 * omv, p4 = p3 * v;
    double R = fx_rcp(p4 * E);
    double iE = R * p4;
    R *= E;
    double iv = R * p3;
    R *= v;
    double iomv = R * p2;
    R *= omv;
    double iw = R * Se, iSe = R * w;
    double q = -u * iw * iSe;               // v'/v
    double r = (1.0 + v) * iomv;
    double lg1 = fma(a, iSe, -q * r);       // (ln g)',  g = Se^a (1-v)^2 / v
    D = Cc * (E * omv * omv * iv);
    iD = iCc * (iE * iomv * iomv * v);
    D1 = D * lg1 * idT;
    if (want2) {
      double q1 = q * iSe * fma(im, iw, -1.0);            // q'
      double r1 = 2.0 * q * v * iomv * iomv;              // r'
      double lg2 = fma(-a * iSe, iSe, -fma(q1, r, q * r1));  // (ln g)''
      D2 = D * fma(lg1, lg1, lg2) * idT * idT;
    }
  } else if constexpr (MODEL == FX_LETD) {
    double Lx = d[0], lnE = d[1], Tx = d[2], Dwt = d[3], tr = d[4], dT = d[5], idT = d[6], iDwt = d[7];
    double Se = div_by(theta - tr, dT, idT);
    double oms = 1.0 - Se;
    double L1 = fx_log(Se, T), L2 = fx_log(oms, T);
    double R = fx_exp(fma(Tx, L2, fma(-Lx, L1, lnE)), T);  // E (1-Se)^T / Se^L
    double opR = 1.0 + R;
    double p2 = Se * oms;
    double ip = fx_rcp(p2 * opR);
    double qq = ip * p2;                                   // 1/(1+R)
    double ir = ip * opR;
    double iSe = ir * oms, ioms = ir * Se;
    double rho = fma(-Tx, ioms, -Lx * iSe);                // R'/R
    double R1 = R * rho;
    D = Dwt * qq;
    iD = iDwt * opR;
    D1 = -Dwt * R1 * qq * qq * idT;
    if (want2) {
      double rho1 = fma(-Tx * ioms, ioms, Lx * iSe * iSe);
      double R2 = R * fma(rho, rho, rho1);
      D2 = -Dwt * fma(-2.0 * R1 * R1, qq, R2) * qq * qq * idT * idT;
    }
  } else if constexpr (MODEL == FX_TABULATED) {
    // D = -(o_1 (I - c))/2 with o(theta) the PCHIP of the profile's inverse, o_k its k-th derivative and I its
    // antiderivative (interpolated.py:59-67): the local cubic of scipy/interpax's PPoly and what jax.grad makes of
    // it.  Knots: theta ascending | o | do/dtheta | I - c, n doubles each.
    const double* tab = reinterpret_cast<const double*>((uintptr_t)f64_bits(d[0]));
    const long long n = (long long)d[1];
    D = D1 = iD = fx_nan();
    D2 = want2 ? fx_nan() : 0.0;
    if (tab != nullptr && n >= 2) {  // (a lane that never held a problem evaluates with an all-zero block)
      const double *x = tab, *y = tab + n, *s = tab + 2 * n, *A = tab + 3 * n;
      if (theta >= x[0] && theta <= x[n - 1]) {  // extrapolate=False; a NaN theta fails both
        long long lo = 0, hi = n;  // searchsorted(x, theta, side="right") - 1, clipped to [0, n-2]
        while (lo < hi) {
          const long long mid = (lo + hi) >> 1;
          if (x[mid] <= theta) lo = mid + 1; else hi = mid;
        }
        long long k = lo - 1;
        k = k < 0 ? 0 : (k > n - 2 ? n - 2 : k);
        const double x0 = x[k], y0 = y[k], d0 = s[k];
        const double h = x[k + 1] - x0, m = (y[k + 1] - y0) / h;
        const double t = (d0 + s[k + 1] - 2.0 * m) / h;
        const double c0 = t / h, c1 = (m - d0) / h - t;
        const double u = theta - x0;
        const double o = ((c0 * u + c1) * u + d0) * u + y0;
        const double o1 = (3.0 * c0 * u + 2.0 * c1) * u + d0;
        const double o2 = 6.0 * c0 * u + 2.0 * c1;
        const double I = A[k] + (((c0 * 0.25 * u + c1 * (1.0 / 3.0)) * u + d0 * 0.5) * u + y0) * u;
        D = -(o1 * I) * 0.5;
        D1 = -(o2 * I + o1 * o) * 0.5;
        if (want2) D2 = -((6.0 * c0) * I + 2.0 * (o2 * o) + o1 * o1) * 0.5;
        iD = fx_rcp(D);
      }
    }
  } else if constexpr (MODEL == FX_LETXS) {
    double Lw = d[0], lnEw = d[1], Tw = d[2], Ls = d[3], lnEs = d[4], Ts = d[5], Cc = d[6];
    double tr = d[7], dT = d[8], idT = d[9], iCc = d[10];
    double Se = div_by(theta - tr, dT, idT);
    double oms = 1.0 - Se;
    double L1 = fx_log(Se, T), L2 = fx_log(oms, T);
    double R = fx_exp(fma(Tw, L2, fma(-Lw, L1, lnEw)), T);  // Ew (1-Se)^Tw / Se^Lw  -> kr = 1/(1+R)
    double S = fx_exp(fma(Ts, L1, fma(-Ls, L2, lnEs)), T);  // Q/P = Es Se^Ts / (1-Se)^Ls
    double opR = 1.0 + R, opS = 1.0 + S;
    // reciprocals of Se, 1-Se, 1+R, 1+S from one
    double p2 = Se * oms, p3 = p2 * opR;
    double Rr = fx_rcp(p3 * opS);
    double qs = Rr * p3;
    Rr *= opS;
    double kr = Rr * p2;
    Rr *= opR;
    double ioms = Rr * Se, iSe = Rr * oms;
    double A = R * kr, B = S * qs;
    double sig = fma(Ls, ioms, Ts * iSe);
    double sig1 = fma(Ls * ioms, ioms, -Ts * iSe * iSe);
    double rho = fma(-Tw, ioms, -Lw * iSe);
    D = Cc * kr * B * qs * sig;
    double ip = fx_rcp(sig * D);
    double isig = D * ip;
    iD = sig * ip;
    (void)iCc;
    double omb2 = fma(-2.0, B, 1.0);
    double G1 = fma(-rho, A, fma(sig, omb2, sig1 * isig));  // (ln D)' wrt Se
    D1 = D * G1 * idT;
    if (want2) {
      double rho1 = fma(-Tw * ioms, ioms, Lw * iSe * iSe);
      double sig2 = 2.0 * fma(Ls * ioms * ioms, ioms, Ts * iSe * iSe * iSe);
      double A1 = rho * A * kr;  // A' = rho A (1-A), 1-A = kr
      double B1 = sig * B * qs;  // B' = sig B (1-B)
      double lsig2 = fma(sig2, sig, -sig1 * sig1) * isig * isig;  // (ln sig)''
      double G2 = fma(-rho1, A, fma(-rho, A1, fma(sig1, omb2, fma(-2.0 * sig, B1, lsig2))));
      D2 = D * fma(G1, G1, G2) * idT * idT;
    }
  }
}

}  // namespace fx
