/*
 * fx_oracle.h -- CPU restatement (plain C) of the frontx shooting solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under frontx_b200/ may include, link or
 * execute this.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py use it, and only as the checker / the
 * CPU arm -- never as the product path.
 *
 * PARITY UNPINNED at the north star's 1e-6: the reference (gerlero/frontx,
 * pure Python on jax/diffrax/optimistix/interpax) cannot be imported here or
 * on the GPU box (no jax wheels, no network), and it ships no golden arrays.
 * What pins this oracle:
 *   - the reference's own known-answer tests (tests/test_solve.py:38-42 Philip
 *     exact exp(-o) to 1e-3; :148-156 unsolvable flag; tests/test_inverse.py:
 *     29-62), reproduced in tests/test_oracle_kat.py;
 *   - 40-digit mpmath values of D, D', D'' evaluated from the reference's
 *     model DEFINITIONS (src/frontx/models.py) -- tests/golden/;
 *   - an independent scipy (Radau + brentq) solution of the same boundary
 *     value problem, which stands in for the reference's `fronts` cross-check
 *     (tests/test_solve.py:141-145, abs 5e-2).
 * The arithmetic of the path lives in un-vendored dependencies (diffrax
 * >=0.7,<0.8: Kvaerno5, PIDController, Event, dense output; optimistix
 * >=0.0.10,<0.1: Bisection, chord root finder; lineax: LU); their published
 * algorithms are restated here, anchored on the reference's call sites
 * src/frontx/_forward.py:60-122 and src/frontx/_boltzmann.py:15-35.
 */
#ifndef FX_ORACLE_H
#define FX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FXO_NPARAM 12

/* model ids and parameter-block layouts (doubles) */
enum {
  FXO_CONSTANT = 0,       /* p[0]=D0                                                  */
  FXO_POWER_LAW = 1,      /* D = a*theta^k + eps;   p = {a, k, eps}                   */
  FXO_BROOKS_COREY = 2,   /* p = {n, l, Ks, alpha, theta_r, theta_s}                  */
  FXO_VAN_GENUCHTEN = 3,  /* p = {n, m, l, Ks, alpha, theta_r, theta_s, form}         */
  FXO_LETXS = 4,          /* p = {Lw, Ew, Tw, Ls, Es, Ts, Ks, alpha, theta_r, theta_s} */
  FXO_LETD = 5,           /* p = {L, E, T, Dwt, theta_r, theta_s}                     */
  FXO_LOG = 6,            /* D = c0 + c1*ln(theta); p = {c0, c1}  (Philip 1960 test)  */
  FXO_TABULATED = 7       /* InterpolatedSolution.D: p = {HOST pointer (bit-cast) to 4 rows of n doubles --
                             theta ascending, o, do/dtheta, Iodtheta - c --, n}               */
};

/* summary[8]  : d_dob, oi, theta(oi), S0=-2 D(b) d_dob, D(b), residual, lower, upper
 * counters[8] : result, n_shots, n_expansions, n_attempts, n_rejected, n_rhs, n_jac, n_knots
 * result      : 0 successful, 1 bisection max_steps_reached, 2 bracket expansion exhausted
 * knots       : [k_max][5] = o, theta, dtheta/do, F0, F1  (F = stage derivative kept for
 *               the cubic Hermite dense output; see fxo_eval)                         */
#define FXO_NSUM 8
#define FXO_NCNT 8
#define FXO_KNOT 5

typedef struct {
  double rtol;          /* 1e-3  (_forward.py:92) */
  double atol;          /* 1e-6  (_forward.py:92) */
  int max_ode_steps;    /* 4096  (diffrax default max_steps) */
  int max_expansions;   /* bound on bracket expansions (U6)  */
  int radial_k;         /* 0 planar (reference); 1 cylindrical/polar, 2 spherical (fronts) */
  double ob;            /* start of integration (0 for the reference) */
  /* Switches for the dependency semantics that cannot be verified without the reference's
   * executable.  0 = the reading this oracle takes by default. */
  int u4_failure_aborts;   /* U4: 1 = a failed implicit stage solve ends the shot (no event) instead of
                              being a rejected step with dt <- 0.2 dt */
  int u5_or_termination;   /* U5: 1 = Bisection stops when EITHER |upper-lower| < atol + rtol|y| OR
                              |residual| < atol (default: both) */
  int u6_expand_rule;      /* U6: bracket expansion: 0 both ends out by the width, 1 both ends out by
                              half the width, 2 upper end out by the width */
} fxo_options;

void fxo_default_options(fxo_options *opt);

/* van Genuchten: params[7] == 0 (default) forms 1 - Se^(1/m) by the literal (cancelling)
 * subtraction of models.py:245,253; params[7] != 0 forms the same value without cancellation.
 * See the comment in D_jet(). */

/* diagnostics summed over all solves since fxo_diag_reset(): how often each uncertain branch ran */
enum { FXO_DIAG_STAGE_FAIL = 0, /* attempts whose implicit solve failed (U4 path)          */
       FXO_DIAG_ERROR_REJECT,   /* attempts rejected by the error test (err >= 1)         */
       FXO_DIAG_NAN_ERROR,      /* attempts with a NaN error estimate (turned into inf)   */
       FXO_DIAG_EXPANSION,      /* bracket expansions (U6 path)                           */
       FXO_NDIAG };
void fxo_diag_reset(void);
void fxo_diag_get(int64_t *out /*[FXO_NDIAG]*/);

/* D, D', D'' from the reference DEFINITIONS by 3rd-order Taylor jets. out[3]. */
void fxo_D(int model, const double *p, double theta, double *out);

/* right-hand side f(o,[theta,p]) -- _boltzmann.py:27-33 */
void fxo_rhs(int model, const double *p, int radial_k, double o, const double *y, double *f);

/* one shot (_forward.py:80-102): returns residual; fills optional knots/counters */
double fxo_shoot(int model, const double *p, double b, double i, double itol,
                 double d_dob, const fxo_options *opt,
                 int k_max, double *knots, int *n_knots, int32_t *counters);

/* full solve (_forward.py:60-122) */
int fxo_solve(int model, const double *p, double b, double i, double itol,
              int max_bisect, const fxo_options *opt,
              double *summary, int32_t *counters, int k_max, double *knots);

/* batch driver, OpenMP over problems (nthreads<=0: all cores) */
int fxo_solve_batch(int64_t n, const int32_t *model, const double *params,
                    const double *b, const double *i, double itol, int max_bisect,
                    const fxo_options *opt, double *summary, int32_t *counters,
                    int k_max, double *knots, int nthreads);

/* dense output (Solution.__call__/d_do, _forward.py:25-37): clip, searchsorted, Hermite */
void fxo_eval(int n_knots, const double *knots, int64_t q, const double *o,
              double *theta, double *dtheta_do);

/* inverse path (InterpolatedSolution, _inverse/interpolated.py:18-73) for one
 * profile with theta strictly monotone; see fx_oracle_inverse.c */
int fxo_inverse(int64_t npts, const double *o, const double *theta,
                double *D_at_knots, double *sorptivity_pchip, double *sorptivity_trapz);

int fxo_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
