// fx_solve.cuh -- the batched shooting-solve kernel (one thread per problem).
//
// Replaces, for a batch, the whole of frontx.solve (src/frontx/_forward.py:60-122):
//   bisection on d_dob (optimistix.Bisection, :104-112)
//     > one IVP per shot (diffrax.diffeqsolve, :84-95) with the boolean event (:76-78)
//       > Kvaerno5 ESDIRK step attempts under the I-controller (rtol 1e-3, atol 1e-6, :92)
//         > 6 implicit stages, each a chord iteration on the 2-vector stage derivative
//           > one right-hand side evaluation (src/frontx/_boltzmann.py:27-33)
//
// Design (B200, FP64-pipe bound; no tensor cores, nothing here is a contraction):
//   * One thread owns one problem from first shot to converged bracket.  The
//     five nested loops above are flattened into ONE loop whose trip is exactly
//     one right-hand-side evaluation (D, D' -- and D'' on the trip that also
//     builds the chord Jacobian).  Everything else (chord update, stage advance,
//     error estimate / controller, event, bisection update, fetching the next
//     problem) is a short predicated block after it, so lanes that are in
//     different stages, steps, shots or problems still execute the expensive
//     model evaluation convergently.
//   * Persistent threads: a finished lane takes the next problem from a global
//     counter (one atomicAdd per warp per trip at most), so a lane stuck in a
//     4096-attempt crawl does not hold back its warp-mates' throughput.
//   * The 7 stage derivatives live in shared memory, one column per thread
//     (conflict-free), because the stage index is a per-lane runtime value; the
//     Butcher tableau and predictor rows sit beside them.
//   * Kernels are compiled per model (template parameter), so the model switch
//     costs nothing and registers hold only that model's derived constants;
//     mixed batches are bucketed by model on the host (fx_api.cu).
//   * Output: 64-byte summary + 32-byte counter record per problem, and one
//     32-byte knot (o, theta, theta', theta'') per accepted step.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fx_models.cuh"

namespace fx {

constexpr int SOLVE_BLOCK = 128;

// Kvaerno5 (Kvaerno 2004; diffrax/_solver/kvaerno5.py): gamma = 0.26
struct Tableau {
  double a[7][6];     // strictly lower rows
  double pred[7][6];  // stage predictors
  double berr[8];     // b_sol - b_embedded (7 used)
  double c[8];        // nodes (7 used)
};

__constant__ Tableau c_tab = {
    {{0, 0, 0, 0, 0, 0},
     {0.26, 0, 0, 0, 0, 0},
     {0.13, 0.84033320996790809, 0, 0, 0, 0},
     {0.22371961478320505, 0.47675532319799699, -0.06470895363112615, 0, 0, 0},
     {0.16648564323248321, 0.10450018841591720, 0.03631482272098715, -0.13090704451073998, 0, 0},
     {0.13855640231268224, 0, -0.04245337201752043, 0.02446657898003141, 0.61943039072480676, 0},
     {0.13659751177640291, 0, -0.05496908796538376, -0.04118626728321046, 0.62993304899016403,
      0.06962479448202728}},
    {{0, 0, 0, 0, 0, 0},
     {1.0, 0, 0, 0, 0, 0},
     {-1.366025403784441, 2.3660254037844357, 0, 0, 0, 0},
     {-0.19650552613122207, 0.8113579546496623, 0.38514757148155954, 0, 0, 0},
     {0.10375304369958693, 0.937994698066431, -0.04174774176601781, 0, 0, 0},
     {-0.17281112873898072, 0.6235784481025847, 0.5492326806363959, 0, 0, 0},
     {0.13855640231268224, 0, -0.04245337201752043, 0.02446657898003141, 0.61943039072480676, 0.26}},
    {0.13659751177640291 - 0.13855640231268224, 0.0, -0.05496908796538376 - (-0.04245337201752043),
     -0.04118626728321046 - 0.02446657898003141, 0.62993304899016403 - 0.61943039072480676,
     0.06962479448202728 - 0.26, 0.26, 0.0},
    {0.0, 0.52, 1.230333209967908, 0.8957659843500759, 0.43639360985864756, 1.0, 1.0, 0.0}};

struct SolveArgs {
  long long n;               // problems in this launch when counts == nullptr
  const unsigned long long* counts;  // per-model bucket sizes (device), or nullptr
  int slot;                  // this launch's bucket
  const int* perm;           // bucket-major list of problem indices, or nullptr
  const double* params;      // [N][12]
  const double* b;           // [N]
  const double* i;           // [N]
  double* summary;           // [N][8]
  int* counters;             // [N][8]
  double* knots;             // [N][k_max][4] or nullptr
  int k_max;
  unsigned long long* queue; // work counter (zeroed before launch)
  double itol, rtol, atol, ob;
  int max_bisect, max_ode_steps, max_expansions, radial_k;
};

enum Mode : int { M_FETCH = 0, M_DB = 1, M_INIT0 = 2, M_INIT1 = 3, M_CHORD = 4, M_DONE = 5 };
enum Phase : int { PH_LOWER = 0, PH_UPPER = 1, PH_MID = 2 };

__device__ __forceinline__ double rms2(double a, double b) {
  return sqrt(fma(a, a, b * b)) * 0.70710678118654752440;
}

template <int MODEL, bool RADIAL>
__global__ void __launch_bounds__(SOLVE_BLOCK) shoot_bisect_kernel(const SolveArgs A) {
  constexpr double GAMMA = 0.26;
  __shared__ Tableau tab;
  __shared__ double fs[7][2][SOLVE_BLOCK];  // stage derivatives, one column per thread

  const int tid = threadIdx.x;
  {
    const double* src = reinterpret_cast<const double*>(&c_tab);
    double* dst = reinterpret_cast<double*>(&tab);
    for (int k = tid; k < (int)(sizeof(Tableau) / sizeof(double)); k += SOLVE_BLOCK) dst[k] = src[k];
  }
  __syncthreads();

  const unsigned lane = tid & 31;
  const unsigned FULL = 0xffffffffu;

  // bucket geometry: model `slot` owns perm[off .. off+n)
  long long n_work = A.n;
  const int* perm = A.perm;
  unsigned long long* queue = A.queue;
  if (A.counts) {
    long long off = 0;
    for (int k = 0; k < A.slot; k++) off += (long long)A.counts[k];
    n_work = (long long)A.counts[A.slot];
    perm += off;
    queue += A.slot;
  }

  // ---- per-lane state ---------------------------------------------------
  int mode = M_FETCH;
  long long prob = -1;
  double dpar[NDERIVED];
#pragma unroll
  for (int k = 0; k < NDERIVED; k++) dpar[k] = 0.0;
  double bnd_b = 0, bnd_i = 0, direction = 0, D_b = 0;
  // bisection
  double lower = 0, upper = 0, d_cur = 0;
  int phase = PH_LOWER, lo_neg = 0, n_bis = 0;
  // integrator
  double t0 = 0, t1 = 0, h = 0, y00 = 0, y01 = 0;  // step [t0,t1], h = t1 - t0, state at t0
  double yp0 = 0, yp1 = 0;                 // y_partial of the current stage
  double f0 = 0, f1 = 0;                   // chord iterate (stage derivative)
  double m00 = 0, m01 = 0, m10 = 0, m11 = 0;  // inverse of gamma*h*J - I
  double dsz = 0;                          // previous chord step size
  double te = 0, ye0 = 0, ye1 = 0;         // where the next RHS is evaluated
  int stage = 1, iter = 0, att_shot = 0;
  // counters
  int n_shots = 0, n_exp = 0, n_att = 0, n_rej = 0, n_rhs = 0, n_jac = 0, n_knots = 0;
  double* knot_row = nullptr;

  for (;;) {
    // ---- fetch ----------------------------------------------------------
    unsigned need = __ballot_sync(FULL, mode == M_FETCH);
    if (need) {
      int leader = __ffs(need) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(queue, (unsigned long long)__popc(need));
      base = __shfl_sync(FULL, base, leader);
      if (mode == M_FETCH) {
        long long w = (long long)base + __popc(need & ((1u << lane) - 1u));
        if (w < n_work) {
          prob = perm ? (long long)perm[w] : w;
          const double* p = A.params + prob * FX_NPARAM;
          derive_params<MODEL>(p, dpar);
          bnd_b = A.b[prob];
          bnd_i = A.i[prob];
          direction = (bnd_i > bnd_b) ? 1.0 : ((bnd_i < bnd_b) ? -1.0 : 0.0);
          knot_row = A.knots ? A.knots + (size_t)prob * A.k_max * FX_KNOT : nullptr;
          n_shots = n_exp = n_att = n_rej = n_rhs = n_jac = n_knots = 0;
          mode = M_DB;
          te = A.ob;
          ye0 = bnd_b;
          ye1 = 0.0;
        } else {
          mode = M_DONE;
        }
      }
    }
    if (!__any_sync(FULL, mode != M_DONE)) break;

    // ---- the one expensive thing per trip: a right-hand-side evaluation ---
    const bool wantJ = (mode == M_CHORD) && (stage == 1) && (iter == 0);
    double D, D1, D2;
    eval_D<MODEL>(dpar, ye0, wantJ, D, D1, D2);
    const double iD = 1.0 / D;
    const double num = fma(D1, ye1, 0.5 * te);  // o/2 + D' p
    double br = num * iD;
    if (RADIAL) br += (double)A.radial_k / te;
    const double F0 = ye1;
    const double F1 = -br * ye1;

    bool att_end = false, att_ok = false, shot_end = false, occurred = false;
    bool start_shot = false, prep = false;

    if (mode == M_DB) {
      // upper = (i - b)/(2 sqrt(D(b)))   _forward.py:110
      D_b = D;
      lower = 0.0;
      upper = (bnd_i - bnd_b) / (2.0 * sqrt(D));
      phase = PH_LOWER;
      d_cur = lower;
      start_shot = true;
    } else if (mode == M_INIT0) {
      // solver.init + first half of the Hairer initial-step heuristic
      n_rhs++;
      fs[0][0][tid] = F0;
      fs[0][1][tid] = F1;
      if (knot_row && A.k_max > 0) *reinterpret_cast<double4*>(knot_row) = make_double4(t0, y00, y01, F1);
      n_knots = 1;
      double sc0 = fma(fabs(y00), A.rtol, A.atol), sc1 = fma(fabs(y01), A.rtol, A.atol);
      double d0 = rms2(y00 / sc0, y01 / sc1);
      double d1 = rms2(F0 / sc0, F1 / sc1);
      bool cond = (d0 < 1e-5) || (d1 < 1e-5);
      double h0 = cond ? 1e-6 : 0.01 * (d0 / d1);
      h = h0;     // parked until INIT1
      dsz = d1;   // parked
      te = t0 + h0;
      ye0 = fma(h0, F0, y00);
      ye1 = fma(h0, F1, y01);
      mode = M_INIT1;
    } else if (mode == M_INIT1) {
      n_rhs++;
      double h0 = h, d1 = dsz;
      double sc0 = fma(fabs(y00), A.rtol, A.atol), sc1 = fma(fabs(y01), A.rtol, A.atol);
      double d2 = rms2((F0 - fs[0][0][tid]) / sc0, (F1 - fs[0][1][tid]) / sc1) / h0;
      double max_d = fmax(d1, d2);
      double h1 = (max_d <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : exp(0.2 * log(0.01 / max_d));
      double dt = fmin(100.0 * h0, h1);
      t1 = t0 + dt;
      h = t1 - t0;
      stage = 1;
      prep = true;
      mode = M_CHORD;
    } else if (mode == M_CHORD) {
      n_rhs++;
      if (wantJ) {
        // chord matrix gamma*h*df/dy - I at the first implicit stage's predicted point
        n_jac++;
        double J10 = -ye1 * (D2 * ye1 * D - num * D1) * iD * iD;
        double J11 = -fma(2.0 * D1, ye1, 0.5 * te) * iD;
        if (RADIAL) J11 -= (double)A.radial_k / te;
        double g = GAMMA * h;
        double a11 = fma(g, J11, -1.0);
        double a10 = g * J10;
        double idet = 1.0 / (-a11 - g * a10);  // det of [[-1, g],[a10, a11]]
        m00 = a11 * idet;
        m01 = -g * idet;
        m10 = -a10 * idet;
        m11 = -idet;
      }
      double r0 = F0 - f0, r1 = F1 - f1;
      double dd0 = fma(m00, r0, m01 * r1);
      double dd1 = fma(m10, r0, m11 * r1);
      f0 -= dd0;
      f1 -= dd1;
      double s0 = fma(A.rtol, fabs(f0), A.atol), s1 = fma(A.rtol, fabs(f1), A.atol);
      double size = rms2(dd0 / s0, dd1 / s1);
      iter++;
      bool done = false, ok = false;
      if (iter >= 2) {
        double rate = size / dsz;
        double factor = size * rate / (1.0 - rate);
        bool small = size < 1e-13;
        bool diverged = !isfinite(rate) || (rate > 2.0);
        bool converged = (factor > 0.0) && (factor < 1e-2);
        done = small || diverged || converged;
        ok = small || (converged && !diverged);
        if (!done && iter >= 10) { done = true; ok = false; }
      }
      dsz = size;
      if (done) {
        if (ok) {
          fs[stage][0][tid] = f0;
          fs[stage][1][tid] = f1;
          if (stage < 6) { stage++; prep = true; }
          else { att_end = true; att_ok = true; }
        } else {
          att_end = true;
        }
      } else {
        ye0 = fma(GAMMA, h * f0, yp0);
        ye1 = fma(GAMMA, h * f1, yp1);
      }
    }

    // ---- end of a step attempt: error estimate, controller, event ---------
    if (att_end) {
      n_att++;
      att_shot++;
      double y10 = y00, y11 = y01, e0 = INFINITY, e1 = INFINITY;
      if (att_ok) {
        y10 = fma(GAMMA, h * f0, yp0);
        y11 = fma(GAMMA, h * f1, yp1);
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
          a0 = fma(tab.berr[j], fs[j][0][tid], a0);
          a1 = fma(tab.berr[j], fs[j][1][tid], a1);
        }
        e0 = h * a0;
        e1 = h * a1;
        if (isnan(e0)) e0 = INFINITY;
        if (isnan(e1)) e1 = INFINITY;
      }
      bool nan1 = isnan(y10) || isnan(y11);
      double q0 = nan1 ? y00 : y10, q1 = nan1 ? y01 : y11;
      double sc0 = fma(fmax(fabs(y00), fabs(q0)), A.rtol, A.atol);
      double sc1 = fma(fmax(fabs(y01), fabs(q1)), A.rtol, A.atol);
      double err = rms2(e0 / sc0, e1 / sc1);
      bool keep = err < 1.0;
      double inv = (err == 0.0) ? INFINITY : 1.0 / err;
      double factor = 0.9 * exp(0.2 * log(inv));
      double fmin_ = keep ? 1.0 : 0.2;
      if (!(factor >= fmin_)) factor = fmin_;
      if (factor > 10.0) factor = 10.0;
      double dt_next = h * factor;
      if (keep) {
        t0 = t1;
        y00 = y10;
        y01 = y11;
        fs[0][0][tid] = f0;  // FSAL
        fs[0][1][tid] = f1;
        if (knot_row && n_knots < A.k_max) {
          double4* dst = reinterpret_cast<double4*>(knot_row + (size_t)n_knots * FX_KNOT);
          *dst = make_double4(t0, y00, y01, f1);
        }
        n_knots++;
        if ((direction * y01 <= 0.0) || (direction * y00 > direction * (bnd_i - A.itol))) {
          shot_end = true;
          occurred = true;
        }
      } else {
        n_rej++;
      }
      if (!shot_end) {
        if (att_shot >= A.max_ode_steps) {
          shot_end = true;
        } else {
          t1 = t0 + dt_next;
          h = t1 - t0;
          stage = 1;
          prep = true;
        }
      }
    }

    // ---- end of a shot: residual, bisection bookkeeping --------------------
    if (shot_end) {
      double res = occurred ? (y00 - bnd_i) : direction * INFINITY;  // _forward.py:97-101
      n_shots++;
      int result = -1;
      if (phase == PH_LOWER) {
        lo_neg = res < 0.0;
        phase = PH_UPPER;
        d_cur = upper;
      } else if (phase == PH_UPPER) {
        int hi_neg = res < 0.0;
        if (lo_neg != hi_neg) {
          phase = PH_MID;
          n_bis = 0;
          d_cur = lower + 0.5 * (upper - lower);
        } else if (n_exp >= A.max_expansions) {
          result = FX_NO_BRACKET;
        } else {  // expand_if_necessary
          double w = upper - lower;
          lower -= w;
          upper += w;
          n_exp++;
          phase = PH_LOWER;
          d_cur = lower;
        }
      } else {
        if ((res < 0.0) == (bool)lo_neg) lower = d_cur; else upper = d_cur;
        n_bis++;
        double scale = A.itol + INFINITY * fabs(d_cur);  // rtol = inf
        if ((fabs(upper - lower) < scale) && (fabs(res) < A.itol)) result = FX_SUCCESSFUL;
        else if (n_bis >= A.max_bisect) result = FX_MAX_STEPS_REACHED;
        else d_cur = lower + 0.5 * (upper - lower);
      }
      if (result < 0) {
        start_shot = true;
      } else {
        double2* s = reinterpret_cast<double2*>(A.summary + (size_t)prob * FX_NSUM);
        s[0] = make_double2(d_cur, t0);
        s[1] = make_double2(y00, -2.0 * D_b * d_cur);
        s[2] = make_double2(D_b, res);
        s[3] = make_double2(lower, upper);
        int4* c = reinterpret_cast<int4*>(A.counters + (size_t)prob * FX_NCNT);
        c[0] = make_int4(result, n_shots, n_exp, n_att);
        c[1] = make_int4(n_rej, n_rhs, n_jac, n_knots);
        mode = M_FETCH;
      }
    }

    if (start_shot) {
      t0 = A.ob;
      y00 = bnd_b;
      y01 = d_cur;
      att_shot = 0;
      te = t0;
      ye0 = y00;
      ye1 = y01;
      mode = M_INIT0;
    }

    // ---- set up an implicit stage: y_partial, predictor, node --------------
    if (prep) {
      double a0 = 0.0, a1 = 0.0, p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (j < stage) {
          double g0 = fs[j][0][tid], g1 = fs[j][1][tid];
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
      iter = 0;
      te = fma(tab.c[stage], h, t0);
      ye0 = fma(GAMMA, h * f0, yp0);
      ye1 = fma(GAMMA, h * f1, yp1);
    }
  }
}

}  // namespace fx
