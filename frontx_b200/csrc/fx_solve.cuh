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
//     mixed batches are bucketed by model on the device (fx_api.cu).
//   * Output: 64-byte summary + 32-byte counter record per problem, and one
//     32-byte knot (o, theta, theta', theta'') per accepted step.
//   * The arithmetic itself is in fx_arith.cuh / fx_models.cuh / fx_math.cuh
//     (host+device, explicit fma, built with -fmad=false) so that the CPU twin
//     used by the tests reproduces every decision bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fx_arith.cuh"
#include "fx_models.cuh"

namespace fx {

constexpr int SOLVE_BLOCK = 128;

__constant__ Tableau c_tab = FX_TABLEAU_INIT;

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

template <int MODEL, bool RADIAL>
__global__ void __launch_bounds__(SOLVE_BLOCK) shoot_bisect_kernel(const SolveArgs A) {
  __shared__ Tableau tab;
  __shared__ MathTab mt;                     // log / exp lookup tables (lanes index them divergently)
  __shared__ double fsm[7][2][SOLVE_BLOCK];  // stage derivatives, one column per thread

  const int tid = threadIdx.x;
  {
    const double* src = reinterpret_cast<const double*>(&c_tab);
    double* dst = reinterpret_cast<double*>(&tab);
    for (int k = tid; k < (int)(sizeof(Tableau) / sizeof(double)); k += SOLVE_BLOCK) dst[k] = src[k];
    src = reinterpret_cast<const double*>(&g_math_tab_dev);
    dst = reinterpret_cast<double*>(&mt);
    for (int k = tid; k < (int)(sizeof(MathTab) / sizeof(double)); k += SOLVE_BLOCK) dst[k] = src[k];
  }
  __syncthreads();

  const unsigned lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  const Tol tol = {A.rtol, A.atol};
  const int radial_k = RADIAL ? A.radial_k : 0;
  auto fs = [&](int j, int c) -> double { return fsm[j][c][tid]; };

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
  Bisect bis;
  bis.start(1.0, 0.0, 0.0);
  double t0 = 0, t1 = 0, h = 0, y00 = 0, y01 = 0;  // step [t0,t1], h = t1 - t0, state at t0
  double yp0 = 0, yp1 = 0;                 // y_partial of the current stage
  double f0 = 0, f1 = 0;                   // chord iterate (stage derivative)
  Chord M = {0, 0, 0, 0};                  // inverse of gamma*h*J - I
  double dsz = 0;                          // square of the previous chord step size
  double te = 0, ye0 = 0, ye1 = 0;         // where the next RHS is evaluated
  int stage = 1, iter = 0, att_shot = 0;
  int n_shots = 0, n_att = 0, n_rej = 0, n_rhs = 0, n_jac = 0, n_knots = 0;
  int rhs_mark = 0, rej_mark = 0;          // counters at the start of the current shot
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
          derive_params<MODEL>(p, dpar, mt);
          bnd_b = A.b[prob];
          bnd_i = A.i[prob];
          direction = (bnd_i > bnd_b) ? 1.0 : ((bnd_i < bnd_b) ? -1.0 : 0.0);
          knot_row = A.knots ? A.knots + (size_t)prob * A.k_max * FX_KNOT : nullptr;
          n_shots = n_att = n_rej = n_rhs = n_jac = n_knots = 0;
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
    double D, D1, D2, iD;
    eval_D<MODEL>(dpar, ye0, wantJ, D, D1, D2, iD, mt);
    const Rhs R = rhs_eval(D1, iD, te, ye1, radial_k);

    bool att_end = false, att_ok = false, shot_end = false, occurred = false;
    bool start_shot = false, prep = false;

    if (mode == M_DB) {
      D_b = D;
      bis.start(D, bnd_b, bnd_i);
      start_shot = true;
    } else if (mode == M_INIT0) {
      // solver.init + first half of the Hairer initial-step heuristic
      n_rhs++;
      fsm[0][0][tid] = R.F0;
      fsm[0][1][tid] = R.F1;
      if (knot_row && A.k_max > 0) *reinterpret_cast<double4*>(knot_row) = make_double4(t0, y00, y01, R.F1);
      n_knots = 1;
      double h0, d1;
      init_h0(y00, y01, R.F0, R.F1, tol, h0, d1);
      h = h0;     // parked until INIT1
      dsz = d1;   // parked
      te = t0 + h0;
      ye0 = fma(h0, R.F0, y00);
      ye1 = fma(h0, R.F1, y01);
      mode = M_INIT1;
    } else if (mode == M_INIT1) {
      n_rhs++;
      double dt = init_dt(y00, y01, fs(0, 0), fs(0, 1), R.F0, R.F1, h, dsz, tol, mt);
      t1 = t0 + dt;
      h = t1 - t0;
      stage = 1;
      prep = true;
      mode = M_CHORD;
    } else if (mode == M_CHORD) {
      n_rhs++;
      if (wantJ) {
        n_jac++;
        M = chord_matrix(D, D1, D2, R, te, ye1, h, radial_k);
      }
      double size = chord_update(M, R.F0, R.F1, f0, f1, tol);
      iter++;
      bool done, ok;
      chord_decide(iter, size, dsz, done, ok);
      dsz = size;
      if (done) {
        if (ok) {
          fsm[stage][0][tid] = f0;
          fsm[stage][1][tid] = f1;
          if (stage < 6) { stage++; prep = true; }
          else { att_end = true; att_ok = true; }
        } else {
          att_end = true;
        }
      } else {
        ye0 = stage_point(yp0, h, f0);
        ye1 = stage_point(yp1, h, f1);
      }
    }

    // ---- end of a step attempt: error estimate, controller, event ---------
    if (att_end) {
      n_att++;
      att_shot++;
      StepEnd s = step_finish(tab, att_ok, fs, y00, y01, yp0, yp1, f0, f1, h, tol, mt);
      if (s.keep) {
        t0 = t1;
        y00 = s.y10;
        y01 = s.y11;
        fsm[0][0][tid] = f0;  // FSAL
        fsm[0][1][tid] = f1;
        if (knot_row && n_knots < A.k_max) {
          double4* dst = reinterpret_cast<double4*>(knot_row + (size_t)n_knots * FX_KNOT);
          *dst = make_double4(t0, y00, y01, f1);
        }
        n_knots++;
        if (event_fired(direction, y00, y01, bnd_i, A.itol)) {
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
          t1 = t0 + s.dt_next;
          h = t1 - t0;
          stage = 1;
          prep = true;
        }
      }
    }

    // ---- end of a shot: residual, bisection bookkeeping --------------------
    if (shot_end) {
      double res = occurred ? (y00 - bnd_i) : direction * fx_inf();  // _forward.py:97-101
      n_shots++;
      double d_shot = bis.d_cur;
      const int rhs_shot = n_rhs - rhs_mark, rej_shot = n_rej - rej_mark;
      int result = bis.next(res, A.itol, A.max_bisect, A.max_expansions);
      // A shot is a pure function of d_dob.  Once the bracket has collapsed to two adjacent
      // doubles without |res| < itol, the midpoint is one of its ends and every remaining
      // bisection step repeats the shot just made: replay its residual and counters
      // instead of integrating again (the knot row it wrote is the same too).
      while (result < 0 && bis.d_cur == d_shot) {
        n_shots++;
        n_att += att_shot;
        n_rej += rej_shot;
        n_rhs += rhs_shot;
        n_jac += att_shot;
        result = bis.next(res, A.itol, A.max_bisect, A.max_expansions);
      }
      if (result < 0) {
        start_shot = true;
      } else {
        double2* s = reinterpret_cast<double2*>(A.summary + (size_t)prob * FX_NSUM);
        s[0] = make_double2(d_shot, t0);
        s[1] = make_double2(y00, -2.0 * D_b * d_shot);
        s[2] = make_double2(D_b, res);
        s[3] = make_double2(bis.lower, bis.upper);
        int4* c = reinterpret_cast<int4*>(A.counters + (size_t)prob * FX_NCNT);
        c[0] = make_int4(result, n_shots, bis.n_exp, n_att);
        c[1] = make_int4(n_rej, n_rhs, n_jac, n_knots);
        mode = M_FETCH;
      }
    }

    if (start_shot) {
      t0 = A.ob;
      y00 = bnd_b;
      y01 = bis.d_cur;
      att_shot = 0;
      rhs_mark = n_rhs;
      rej_mark = n_rej;
      te = t0;
      ye0 = y00;
      ye1 = y01;
      mode = M_INIT0;
    }

    // ---- set up an implicit stage: y_partial, predictor, node --------------
    if (prep) {
      stage_prepare(tab, stage, fs, y00, y01, h, t0, yp0, yp1, f0, f1, te);
      iter = 0;
      ye0 = stage_point(yp0, h, f0);
      ye1 = stage_point(yp1, h, f1);
    }
  }
}

}  // namespace fx
