/*
 * fx_twin.cpp -- CPU twin of the CUDA kernel's ARITHMETIC, driven by plain nested loops.
 *
 * TEST INFRASTRUCTURE ONLY (see fx_oracle.h): used by tests/ and smoke() as a
 * checker; the product never links or calls it.
 *
 * Why a second CPU implementation: the shooting problem amplifies last-bit
 * differences in D(theta) (libm pow vs any other pow) into different
 * accept/reject and chord-convergence decisions in a few percent of problems, so
 * the literal restatement (fx_oracle.c: libm, autodiff-style jets, LU) can only be
 * compared with the GPU statistically.  This twin is a HOST BUILD of the
 * kernel's own arithmetic headers (frontx_b200/csrc/fx_math.cuh, fx_models.cuh,
 * fx_arith.cuh are host+device, explicit fma, contraction off on both sides),
 * called from the textbook loop nest
 *     bisection > shot > step attempt > stage > chord iteration
 * instead of the kernel's flattened one-RHS-per-trip state machine.  It must agree
 * with the GPU BIT FOR BIT (tests/test_gpu_parity.py), which checks the state
 * machine, the work queue, the shared-memory staging and the output paths; the
 * arithmetic itself is checked against the independent literal oracle on the CPU
 * (tests/test_twin_vs_literal.py).
 *
 * Control flow follows /root/reference/src/frontx/_forward.py:60-122 exactly as
 * fx_oracle.c does.
 */
#include <pthread.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstring>

#include "../frontx_b200/csrc/fx_arith.cuh"
#include "../frontx_b200/csrc/fx_models.cuh"

namespace {

const fx::Tableau h_tab = FX_TABLEAU_INIT;
const fx::MathTab& MT = fx::g_math_tab_host;

struct Options {
  double itol, rtol, atol, ob;
  int max_bisect, max_ode_steps, max_expansions, radial_k;
};

struct Counters {
  int n_shots = 0, n_att = 0, n_rej = 0, n_rhs = 0, n_jac = 0, n_knots = 0;
};

struct Shot {
  double residual, t_end, theta_end;
};

template <int MODEL>
Shot shoot(const double (&dpar)[fx::NDERIVED], double b, double i, double d_dob, const Options& o,
           Counters& c, int k_max, double* knots) {
  using namespace fx;
  const Tol tol = {o.rtol, o.atol};
  const double direction = (i > b) ? 1.0 : ((i < b) ? -1.0 : 0.0);
  double fsv[7][2];
  auto fs = [&](int j, int k) -> double { return fsv[j][k]; };
  double t0 = o.ob, y00 = b, y01 = d_dob;
  double D, D1, D2, iD;

  // solver.init and the initial step size (two right-hand sides)
  eval_D<MODEL>(dpar, y00, false, D, D1, D2, iD, MT);
  Rhs R = rhs_eval(D1, iD, t0, y01, o.radial_k);
  c.n_rhs++;
  fsv[0][0] = R.F0;
  fsv[0][1] = R.F1;
  if (knots && k_max > 0) { knots[0] = t0; knots[1] = y00; knots[2] = y01; knots[3] = R.F1; }
  c.n_knots = 1;
  double h0, d1;
  init_h0(y00, y01, R.F0, R.F1, tol, h0, d1);
  eval_D<MODEL>(dpar, fma(h0, R.F0, y00), false, D, D1, D2, iD, MT);
  Rhs R1 = rhs_eval(D1, iD, t0 + h0, fma(h0, R.F1, y01), o.radial_k);
  c.n_rhs++;
  double dt = init_dt(y00, y01, fsv[0][0], fsv[0][1], R1.F0, R1.F1, h0, d1, tol, MT);
  double t1 = t0 + dt;

  bool occurred = false;
  for (int att = 0; att < o.max_ode_steps;) {
    double h = t1 - t0;
    double yp0 = 0, yp1 = 0, f0 = 0, f1 = 0, te = 0;
    Chord M = {0, 0, 0, 0};
    bool ok = true;
    for (int stage = 1; stage <= 6 && ok; stage++) {
      stage_prepare(h_tab, stage, fs, y00, y01, h, t0, yp0, yp1, f0, f1, te);
      double prev = 0.0;
      for (int iter = 0;;) {
        double ye0 = stage_point(yp0, h, f0), ye1 = stage_point(yp1, h, f1);
        bool wantJ = (stage == 1) && (iter == 0);
        eval_D<MODEL>(dpar, ye0, wantJ, D, D1, D2, iD, MT);
        Rhs Rs = rhs_eval(D1, iD, te, ye1, o.radial_k);
        c.n_rhs++;
        if (wantJ) {
          M = chord_matrix(D, D1, D2, Rs, te, ye1, h, o.radial_k);
          c.n_jac++;
        }
        double size = chord_update(M, Rs.F0, Rs.F1, f0, f1, tol);
        iter++;
        bool done, good;
        chord_decide(iter, size, prev, done, good);
        prev = size;
        if (done) {
          ok = good;
          break;
        }
      }
      if (ok) {
        fsv[stage][0] = f0;
        fsv[stage][1] = f1;
      }
    }
    att++;
    c.n_att++;
    StepEnd s = step_finish(h_tab, ok, fs, y00, y01, yp0, yp1, f0, f1, h, tol, MT);
    if (s.keep) {
      t0 = t1;
      y00 = s.y10;
      y01 = s.y11;
      fsv[0][0] = f0;  // FSAL
      fsv[0][1] = f1;
      if (knots && c.n_knots < k_max) {
        double* k = knots + (size_t)FX_KNOT * c.n_knots;
        k[0] = t0; k[1] = y00; k[2] = y01; k[3] = f1;
      }
      c.n_knots++;
      if (event_fired(direction, y00, y01, i, o.itol)) {
        occurred = true;
        break;
      }
    } else {
      c.n_rej++;
    }
    t1 = t0 + s.dt_next;
  }
  Shot out;
  out.residual = occurred ? (y00 - i) : direction * fx_inf();
  out.t_end = t0;
  out.theta_end = y00;
  return out;
}

template <int MODEL>
int solve_one(const double* p, double b, double i, const Options& o, double* summary, int32_t* counters,
              int k_max, double* knots) {
  using namespace fx;
  double dpar[NDERIVED];
  derive_params<MODEL>(p, dpar, MT);
  double D, D1, D2, iD;
  eval_D<MODEL>(dpar, b, false, D, D1, D2, iD, MT);
  const double D_b = D;
  Bisect bis;
  bis.start(D_b, b, i);
  Counters c;
  int result;
  Shot last;
  double d_shot;
  do {
    d_shot = bis.d_cur;
    last = shoot<MODEL>(dpar, b, i, d_shot, o, c, k_max, knots);
    c.n_shots++;
    result = bis.next(last.residual, o.itol, o.max_bisect, o.max_expansions);
  } while (result < 0);
  summary[0] = d_shot;
  summary[1] = last.t_end;
  summary[2] = last.theta_end;
  summary[3] = -2.0 * D_b * d_shot;
  summary[4] = D_b;
  summary[5] = last.residual;
  summary[6] = bis.lower;
  summary[7] = bis.upper;
  counters[0] = result;
  counters[1] = c.n_shots;
  counters[2] = bis.n_exp;
  counters[3] = c.n_att;
  counters[4] = c.n_rej;
  counters[5] = c.n_rhs;
  counters[6] = c.n_jac;
  counters[7] = c.n_knots;
  return result;
}

// Flow-rate boundary (north-star C4; fronts.solve_flowrate semantics): the
// inlet value b is unknown, the inlet flux density q = Qb/(angle ob height) is given, so every
// shot starts from (b, d_dob = -q/D(b)) and the bisection runs on b inside [b_lo, b_hi].
template <int MODEL>
int solve_one_flow(const double* p, double q, double i, double b_lo, double b_hi, const Options& o,
                   double* summary, int32_t* counters, int k_max, double* knots) {
  using namespace fx;
  double dpar[NDERIVED];
  derive_params<MODEL>(p, dpar, MT);
  Bisect bis;
  bis.lower = b_lo;
  bis.upper = b_hi;
  bis.phase = 0;
  bis.lo_neg = 0;
  bis.n_bis = 0;
  bis.n_exp = 0;
  bis.d_cur = b_lo;
  Counters c;
  int result;
  Shot last;
  double b_shot, D_b = 0.0, d_dob = 0.0;
  do {
    b_shot = bis.d_cur;
    double D, D1, D2, iD;
    eval_D<MODEL>(dpar, b_shot, false, D, D1, D2, iD, MT);
    D_b = D;
    d_dob = -q * iD;
    last = shoot<MODEL>(dpar, b_shot, i, d_dob, o, c, k_max, knots);
    c.n_shots++;
    result = bis.next(last.residual, o.itol, o.max_bisect, 0);
  } while (result < 0);
  summary[0] = b_shot;
  summary[1] = last.t_end;
  summary[2] = last.theta_end;
  summary[3] = d_dob;
  summary[4] = D_b;
  summary[5] = last.residual;
  summary[6] = bis.lower;
  summary[7] = bis.upper;
  counters[0] = result;
  counters[1] = c.n_shots;
  counters[2] = bis.n_exp;
  counters[3] = c.n_att;
  counters[4] = c.n_rej;
  counters[5] = c.n_rhs;
  counters[6] = c.n_jac;
  counters[7] = c.n_knots;
  return result;
}

int flow_dispatch(int model, const double* p, double q, double i, double b_lo, double b_hi, const Options& o,
                  double* summary, int32_t* counters, int k_max, double* knots) {
  switch (model) {
    case FX_CONSTANT: return solve_one_flow<FX_CONSTANT>(p, q, i, b_lo, b_hi, o, summary, counters, k_max, knots);
    case FX_POWER_LAW: return solve_one_flow<FX_POWER_LAW>(p, q, i, b_lo, b_hi, o, summary, counters, k_max, knots);
    case FX_BROOKS_COREY: return solve_one_flow<FX_BROOKS_COREY>(p, q, i, b_lo, b_hi, o, summary, counters, k_max, knots);
    case FX_VAN_GENUCHTEN: return solve_one_flow<FX_VAN_GENUCHTEN>(p, q, i, b_lo, b_hi, o, summary, counters, k_max, knots);
    case FX_LETXS: return solve_one_flow<FX_LETXS>(p, q, i, b_lo, b_hi, o, summary, counters, k_max, knots);
    case FX_LETD: return solve_one_flow<FX_LETD>(p, q, i, b_lo, b_hi, o, summary, counters, k_max, knots);
    case FX_LOG: return solve_one_flow<FX_LOG>(p, q, i, b_lo, b_hi, o, summary, counters, k_max, knots);
    case FX_TABULATED: return solve_one_flow<FX_TABULATED>(p, q, i, b_lo, b_hi, o, summary, counters, k_max, knots);
    default: counters[0] = -1; return -1;
  }
}

int solve_dispatch(int model, const double* p, double b, double i, const Options& o, double* summary,
                   int32_t* counters, int k_max, double* knots) {
  switch (model) {
    case FX_CONSTANT: return solve_one<FX_CONSTANT>(p, b, i, o, summary, counters, k_max, knots);
    case FX_POWER_LAW: return solve_one<FX_POWER_LAW>(p, b, i, o, summary, counters, k_max, knots);
    case FX_BROOKS_COREY: return solve_one<FX_BROOKS_COREY>(p, b, i, o, summary, counters, k_max, knots);
    case FX_VAN_GENUCHTEN: return solve_one<FX_VAN_GENUCHTEN>(p, b, i, o, summary, counters, k_max, knots);
    case FX_LETXS: return solve_one<FX_LETXS>(p, b, i, o, summary, counters, k_max, knots);
    case FX_LETD: return solve_one<FX_LETD>(p, b, i, o, summary, counters, k_max, knots);
    case FX_LOG: return solve_one<FX_LOG>(p, b, i, o, summary, counters, k_max, knots);
    case FX_TABULATED: return solve_one<FX_TABULATED>(p, b, i, o, summary, counters, k_max, knots);
    default: counters[0] = -1; return -1;
  }
}

template <int MODEL>
void D_one(const double* p, double theta, double* out) {
  double d[fx::NDERIVED];
  fx::derive_params<MODEL>(p, d, MT);
  double iD;
  fx::eval_D<MODEL>(d, theta, true, out[0], out[1], out[2], iD, MT);
}

struct Batch {
  int64_t n;
  const int32_t* model;
  const double *params, *b, *i;
  const double *q = nullptr, *b_hi = nullptr;  // flow-rate boundary: flux density and upper bracket (b is the lower)
  Options o;
  double* summary;
  int32_t* counters;
  int k_max;
  double* knots;
  std::atomic<long long> next{0};
};

void* worker(void* arg) {
  Batch* B = static_cast<Batch*>(arg);
  for (;;) {
    long long j0 = B->next.fetch_add(4);
    if (j0 >= B->n) break;
    long long j1 = j0 + 4 < B->n ? j0 + 4 : B->n;
    for (long long j = j0; j < j1; j++) {
      double* kn = B->knots ? B->knots + (size_t)FX_KNOT * B->k_max * j : nullptr;
      if (B->q)
        flow_dispatch(B->model[j], B->params + (size_t)FX_NPARAM * j, B->q[j], B->i[j], B->b[j], B->b_hi[j], B->o,
                      B->summary + (size_t)FX_NSUM * j, B->counters + (size_t)FX_NCNT * j, B->k_max, kn);
      else
        solve_dispatch(B->model[j], B->params + (size_t)FX_NPARAM * j, B->b[j], B->i[j], B->o,
                       B->summary + (size_t)FX_NSUM * j, B->counters + (size_t)FX_NCNT * j, B->k_max, kn);
    }
  }
  return nullptr;
}

}  // namespace

extern "C" {

void fxt_D(int model, const double* p, double theta, double* out) {
  switch (model) {
    case FX_CONSTANT: D_one<FX_CONSTANT>(p, theta, out); break;
    case FX_POWER_LAW: D_one<FX_POWER_LAW>(p, theta, out); break;
    case FX_BROOKS_COREY: D_one<FX_BROOKS_COREY>(p, theta, out); break;
    case FX_VAN_GENUCHTEN: D_one<FX_VAN_GENUCHTEN>(p, theta, out); break;
    case FX_LETXS: D_one<FX_LETXS>(p, theta, out); break;
    case FX_LETD: D_one<FX_LETD>(p, theta, out); break;
    case FX_LOG: D_one<FX_LOG>(p, theta, out); break;
    case FX_TABULATED: D_one<FX_TABULATED>(p, theta, out); break;
    default: out[0] = out[1] = out[2] = fx::fx_nan();
  }
}

double fxt_log(double x) { return fx::fx_log(x, MT); }
double fxt_exp(double x) { return fx::fx_exp(x, MT); }
double fxt_expm1(double x) { return fx::fx_expm1(x, MT); }

int fxt_solve_flowrate_batch(int64_t n, const int32_t* model, const double* params, const double* q,
                             const double* i, const double* b_lo, const double* b_hi, double itol,
                             int max_bisect, int max_ode_steps, double rtol, double atol, int radial_k, double ob,
                             double* summary, int32_t* counters, int k_max, double* knots, int nthreads);

int fxt_solve_batch(int64_t n, const int32_t* model, const double* params, const double* b, const double* i,
                    double itol, int max_bisect, int max_ode_steps, double rtol, double atol,
                    int max_expansions, int radial_k, double ob, double* summary, int32_t* counters,
                    int k_max, double* knots, int nthreads) {
  Batch B;
  B.n = n;
  B.model = model;
  B.params = params;
  B.b = b;
  B.i = i;
  B.o = Options{itol, rtol, atol, ob, max_bisect, max_ode_steps, max_expansions, radial_k};
  B.summary = summary;
  B.counters = counters;
  B.k_max = k_max;
  B.knots = knots;
  if (nthreads <= 0) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    nthreads = c > 0 ? (int)c : 1;
  }
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < nthreads; t++)
    if (pthread_create(&th[started], nullptr, worker, &B) == 0) started++;
  worker(&B);
  for (int t = 0; t < started; t++) pthread_join(th[t], nullptr);
  return 0;
}

int fxt_solve_flowrate_batch(int64_t n, const int32_t* model, const double* params, const double* q,
                             const double* i, const double* b_lo, const double* b_hi, double itol,
                             int max_bisect, int max_ode_steps, double rtol, double atol, int radial_k, double ob,
                             double* summary, int32_t* counters, int k_max, double* knots, int nthreads) {
  Batch B;
  B.n = n;
  B.model = model;
  B.params = params;
  B.b = b_lo;
  B.i = i;
  B.q = q;
  B.b_hi = b_hi;
  B.o = Options{itol, rtol, atol, ob, max_bisect, max_ode_steps, 0, radial_k};
  B.summary = summary;
  B.counters = counters;
  B.k_max = k_max;
  B.knots = knots;
  if (nthreads <= 0) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    nthreads = c > 0 ? (int)c : 1;
  }
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < nthreads; t++)
    if (pthread_create(&th[started], nullptr, worker, &B) == 0) started++;
  worker(&B);
  for (int t = 0; t < started; t++) pthread_join(th[t], nullptr);
  return 0;
}

}  // extern "C"
