// fx_arith.cuh -- the floating-point arithmetic of one Kvaerno5 step attempt, the
// chord iteration, the step-size controller, the event and the bisection update,
// as small host+device functions.
//
// The CUDA kernel (fx_solve.cuh) drives these from its flattened one-RHS-per-trip
// state machine; the CPU twin (oracle/fx_twin.cpp, test infrastructure) drives
// the SAME functions from plain nested loops.  Because every multiply-add is an
// explicit fma() and both sides are built with contraction off, the two agree bit
// for bit, which is what lets the GPU tests demand identical decision counters.
//
// Semantics restated (see oracle/fx_oracle.c for the literal, independent version):
//   diffrax Kvaerno5 / AbstractESDIRK stage loop, chord root finder with one Jacobian
//   per attempt, PIDController(rtol, atol) as an I-controller, Hairer initial step,
//   boolean event of src/frontx/_forward.py:76-78, optimistix Bisection(:104-112).
#pragma once

#include <math.h>

#include "fx_math.cuh"

namespace fx {

constexpr double GAMMA = 0.26;  // Kvaerno5 diagonal

// Kvaerno5 (Kvaerno 2004; diffrax/_solver/kvaerno5.py)
struct Tableau {
  double a[7][6];     // strictly lower rows
  double pred[7][6];  // stage predictors
  double berr[8];     // b_sol - b_embedded (7 used)
  double c[8];        // nodes (7 used)
};

#define FX_TABLEAU_INIT                                                                              \
  {                                                                                                  \
    {{0, 0, 0, 0, 0, 0},                                                                             \
     {0.26, 0, 0, 0, 0, 0},                                                                          \
     {0.13, 0.84033320996790809, 0, 0, 0, 0},                                                        \
     {0.22371961478320505, 0.47675532319799699, -0.06470895363112615, 0, 0, 0},                      \
     {0.16648564323248321, 0.10450018841591720, 0.03631482272098715, -0.13090704451073998, 0, 0},    \
     {0.13855640231268224, 0, -0.04245337201752043, 0.02446657898003141, 0.61943039072480676, 0},    \
     {0.13659751177640291, 0, -0.05496908796538376, -0.04118626728321046, 0.62993304899016403,       \
      0.06962479448202728}},                                                                         \
        {{0, 0, 0, 0, 0, 0},                                                                         \
         {1.0, 0, 0, 0, 0, 0},                                                                       \
         {-1.366025403784441, 2.3660254037844357, 0, 0, 0, 0},                                       \
         {-0.19650552613122207, 0.8113579546496623, 0.38514757148155954, 0, 0, 0},                   \
         {0.10375304369958693, 0.937994698066431, -0.04174774176601781, 0, 0, 0},                    \
         {-0.17281112873898072, 0.6235784481025847, 0.5492326806363959, 0, 0, 0},                    \
         {0.13855640231268224, 0, -0.04245337201752043, 0.02446657898003141, 0.61943039072480676,    \
          0.26}},                                                                                    \
        {0.13659751177640291 - 0.13855640231268224, 0.0,                                             \
         -0.05496908796538376 - (-0.04245337201752043), -0.04118626728321046 - 0.02446657898003141,  \
         0.62993304899016403 - 0.61943039072480676, 0.06962479448202728 - 0.26, 0.26, 0.0},          \
    {                                                                                                \
      0.0, 0.52, 1.230333209967908, 0.8957659843500759, 0.43639360985864756, 1.0, 1.0, 0.0           \
    }                                                                                                \
  }

struct Tol {
  double rtol, atol;
};

// root mean square of (a, b), squared: the controller's and the chord iteration's norm is
// sqrt(mean(x^2)) (diffrax rms_norm); every test made on it is made on its square here,
// which needs no square root
FX_HD double ms2(double a, double b) { return 0.5 * fma(a, a, b * b); }

// right-hand side of src/frontx/_boltzmann.py:27-33: [p, -((o/2 + D' p)/D) p]  (+ k/o radial term);
// iD = 1/D comes from the model evaluation, which shares one reciprocal among its quotients
struct Rhs {
  double F0, F1, iD, num;
};
FX_HD Rhs rhs_eval(double D1, double iD, double te, double ye1, int radial_k) {
  Rhs r;
  r.iD = iD;
  r.num = fma(D1, ye1, 0.5 * te);  // o/2 + D' p
  double br = r.num * r.iD;
  if (radial_k) br += (double)radial_k / te;
  r.F0 = ye1;
  r.F1 = -br * ye1;
  return r;
}

// inverse of the chord matrix gamma*h*df/dy - I at the evaluation point (2x2, closed form)
struct Chord {
  double m00, m01, m10, m11;
};
FX_HD Chord chord_matrix(double D, double D1, double D2, const Rhs& r, double te, double ye1, double h,
                         int radial_k) {
  double J10 = -ye1 * fma(D2 * ye1, D, -r.num * D1) * r.iD * r.iD;
  double J11 = -fma(2.0 * D1, ye1, 0.5 * te) * r.iD;
  if (radial_k) J11 -= (double)radial_k / te;
  double g = KC(K_GAMMA) * h;
  double a11 = fma(g, J11, -1.0);
  double a10 = g * J10;
  double idet = fx_rcp(fma(-g, a10, -a11));  // det of [[-1, g],[a10, a11]]
  Chord M;
  M.m00 = a11 * idet;
  M.m01 = -g * idet;
  M.m10 = -a10 * idet;
  M.m11 = -idet;
  return M;
}

// one chord iteration on the stage derivative f: f -= M (F - f); returns the SQUARE of the
// scaled step size ||diff / (atol + rtol |f|)||_rms, from one shared reciprocal
FX_HD double chord_update(const Chord& M, double F0, double F1, double& f0, double& f1, Tol tol) {
  double r0 = F0 - f0, r1 = F1 - f1;
  double dd0 = fma(M.m00, r0, M.m01 * r1);
  double dd1 = fma(M.m10, r0, M.m11 * r1);
  f0 -= dd0;
  f1 -= dd1;
  double s0 = fma(tol.rtol, fabs(f0), tol.atol), s1 = fma(tol.rtol, fabs(f1), tol.atol);
  double ir = fx_rcp(s0 * s1);
  return ms2(dd0 * s1 * ir, dd1 * s0 * ir);
}

// chord termination after `iter` iterations (optimistix Chord.terminate, max 10 steps), on
// the squares S = size^2, P = prev^2 of the last two step sizes.  With rate = size/prev and
// factor = size*rate/(1-rate) = size^2/(prev-size):
//   small     = size < 1e-13                      <=>  S < 1e-26
//   diverged  = !isfinite(rate) or rate > 2       <=>  !(S <= 4 P) or !(P > 0) or !(S < inf)
//   converged = 0 < factor < kappa = 1e-2         <=>  L > 0 and L^2 > 4 kappa^2 S^3,
//               L = kappa^2 (P - S) - S^2         (squaring kappa*prev > S + kappa*size twice)
// so the iteration needs neither a division nor a square root.
FX_HD void chord_decide(int iter, double S, double P, bool& done, bool& ok) {
  done = false;
  ok = false;
  if (iter >= 2) {
    bool small = S < KC(K_1EM26);
    bool diverged = !(S <= 4.0 * P) || !(P > 0.0) || !(S < fx_inf());
    double L = fma(KC(K_1EM4), P - S, -S * S);
    bool converged = (L > 0.0) && (L * L > KC(K_4EM4) * S * S * S);
    done = small || diverged || converged;
    ok = small || (converged && !diverged);
    if (!done && iter >= 10) done = true;
  }
}

// Hairer initial step, first half: h0 and d1 from f(t0, y0)
FX_HD void init_h0(double y00, double y01, double F0, double F1, Tol tol, double& h0, double& d1) {
  double sc0 = fma(fabs(y00), tol.rtol, tol.atol), sc1 = fma(fabs(y01), tol.rtol, tol.atol);
  double i0 = fx_rcp(sc0), i1 = fx_rcp(sc1);
  double d0 = sqrt(ms2(y00 * i0, y01 * i1));
  d1 = sqrt(ms2(F0 * i0, F1 * i1));
  bool cond = (d0 < KC(K_1EM5)) || (d1 < KC(K_1EM5));
  h0 = cond ? KC(K_1EM6) : KC(K_0P01) * (d0 * fx_rcp(d1));
}
// second half: dt0 from f(t0 + h0, y0 + h0 f0); error order 5
FX_HD double init_dt(double y00, double y01, double f00, double f01, double F0, double F1, double h0,
                     double d1, Tol tol, const MathTab& T) {
  double sc0 = fma(fabs(y00), tol.rtol, tol.atol), sc1 = fma(fabs(y01), tol.rtol, tol.atol);
  double d2 = sqrt(ms2((F0 - f00) * fx_rcp(sc0), (F1 - f01) * fx_rcp(sc1))) * fx_rcp(h0);
  double max_d = fmax(d1, d2);
  double h1 = (max_d <= KC(K_1EM15)) ? fmax(KC(K_1EM6), h0 * KC(K_1EM3))
                                     : fx_exp(KC(K_0P2) * fx_log(KC(K_0P01) * fx_rcp(max_d), T), T);
  return fmin(100.0 * h0, h1);
}

// y_partial and predictor of implicit stage `stage` (1..6), its node, and the first
// evaluation point.  fs(j, c) returns component c of stage derivative j.
template <class FS>
FX_HD void stage_prepare(const Tableau& tab, int stage, FS fs, double y00, double y01, double h, double t0,
                         double& yp0, double& yp1, double& f0, double& f1, double& te) {
  double a0 = 0.0, a1 = 0.0, p0 = 0.0, p1 = 0.0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j < stage) {
      double g0 = fs(j, 0), g1 = fs(j, 1);
      double aa = tab.a[stage][j], pp = tab.pred[stage][j];
      a0 = fma(aa, g0, a0);
      a1 = fma(aa, g1, a1);
      p0 = fma(pp, g0, p0);
      p1 = fma(pp, g1, p1);
    }
  }
  yp0 = fma(h, a0, y00);
  yp1 = fma(h, a1, y01);
  f0 = p0;
  f1 = p1;
  te = fma(tab.c[stage], h, t0);
}

FX_HD double stage_point(double yp, double h, double f) { return fma(KC(K_GAMMA), h * f, yp); }

// end of a step attempt: stiffly accurate y1, embedded error, I-controller
// (PIDController(rtol, atol) with pcoeff = dcoeff = 0: factor = 0.9 err^(-1/5) clipped to
// [0.2 or 1, 10]; accept iff err < 1).  Works on E2 = err^2: err < 1 <=> E2 < 1 and
// err^(-1/5) = E2^(-1/10), so no square root is taken.  A failed implicit solve or a NaN error
// estimate is err = inf: rejected, dt <- 0.2 dt, with no arithmetic on infinities.
struct StepEnd {
  double y10, y11, dt_next;
  bool keep;
};
template <class FS>
FX_HD StepEnd step_finish(const Tableau& tab, bool ok, FS fs, double y00, double y01, double yp0, double yp1,
                          double f0, double f1, double h, Tol tol, const MathTab& T) {
  StepEnd s;
  s.y10 = y00;
  s.y11 = y01;
  s.keep = false;
  s.dt_next = h * KC(K_0P2);
  if (!ok) return s;
  s.y10 = stage_point(yp0, h, f0);
  s.y11 = stage_point(yp1, h, f1);
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int j = 0; j < 7; j++) {
    a0 = fma(tab.berr[j], fs(j, 0), a0);
    a1 = fma(tab.berr[j], fs(j, 1), a1);
  }
  double e0 = h * a0, e1 = h * a1;
  bool nan1 = (s.y10 != s.y10) || (s.y11 != s.y11);
  double q0 = nan1 ? y00 : s.y10, q1 = nan1 ? y01 : s.y11;
  double sc0 = fma(fmax(fabs(y00), fabs(q0)), tol.rtol, tol.atol);
  double sc1 = fma(fmax(fabs(y01), fabs(q1)), tol.rtol, tol.atol);
  double ir = fx_rcp(sc0 * sc1);
  double E2 = ms2(e0 * sc1 * ir, e1 * sc0 * ir);
  if (!(E2 < fx_inf())) return s;  // inf or NaN error estimate: reject and shrink
  s.keep = E2 < 1.0;
  double factor = 10.0;
  if (E2 > 0.0) {
    factor = KC(K_0P9) * fx_exp(-KC(K_0P1) * fx_log(E2, T), T);
    double fmin_ = s.keep ? 1.0 : KC(K_0P2);
    if (!(factor >= fmin_)) factor = fmin_;
    if (factor > 10.0) factor = 10.0;
  }
  s.dt_next = h * factor;
  return s;
}

// boolean event of src/frontx/_forward.py:76-78
FX_HD bool event_fired(double direction, double theta, double dtheta, double i, double itol) {
  return (direction * dtheta <= 0.0) || (direction * theta > direction * (i - itol));
}

// bisection on d_dob (optimistix.Bisection(rtol=inf, atol=itol, flip="detect",
// expand_if_necessary=True); src/frontx/_forward.py:104-112)
struct Bisect {
  double lower, upper, d_cur;
  int phase, lo_neg, n_bis, n_exp;  // phase 0: lower end, 1: upper end, 2: midpoints

  FX_HD void start(double D_b, double b, double i) {
    lower = 0.0;
    upper = (i - b) / (2.0 * sqrt(D_b));  // _forward.py:110
    phase = 0;
    lo_neg = 0;
    n_bis = 0;
    n_exp = 0;
    d_cur = lower;
  }
  // consume the residual of the shot at d_cur; returns -1 (shoot again at d_cur) or an fx_result
  FX_HD int next(double res, double itol, int max_bisect, int max_expansions) {
    if (phase == 0) {
      lo_neg = res < 0.0;
      phase = 1;
      d_cur = upper;
      return -1;
    }
    if (phase == 1) {
      int hi_neg = res < 0.0;
      if (lo_neg != hi_neg) {
        phase = 2;
        n_bis = 0;
        d_cur = fma(0.5, upper - lower, lower);
        return -1;
      }
      if (n_exp >= max_expansions) return 2;  // FX_NO_BRACKET
      double w = upper - lower;               // expand_if_necessary
      lower -= w;
      upper += w;
      n_exp++;
      phase = 0;
      d_cur = lower;
      return -1;
    }
    if ((res < 0.0) == (bool)lo_neg) lower = d_cur; else upper = d_cur;
    n_bis++;
    double scale = itol + fx_inf() * fabs(d_cur);  // atol + rtol*|y| with rtol = inf
    if ((fabs(upper - lower) < scale) && (fabs(res) < itol)) return 0;  // FX_SUCCESSFUL
    if (n_bis >= max_bisect) return 1;                                 // FX_MAX_STEPS_REACHED
    d_cur = fma(0.5, upper - lower, lower);
    return -1;
  }
};

}  // namespace fx
