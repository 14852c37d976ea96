/*
 * frontx_b200.h -- C ABI of libfrontx_b200.so: the B200 (sm_100a) batched
 * Boltzmann-transformed shooting solve behind frontx.solve.
 *
 * The reference (gerlero/frontx 0.1.0) is pure Python on jax/diffrax and has
 * no FFI of its own; its boundary for this path is the Python call
 *     frontx.solve(D, *, b, i, itol=1e-3, max_steps=100, throw=True) -> Solution
 *         (src/frontx/_forward.py:60-72)
 * plus the Solution protocol (src/frontx/_forward.py:17-57,
 * src/frontx/_boltzmann.py:101-161) and InterpolatedSolution
 * (src/frontx/_inverse/interpolated.py:18-73).  An arbitrary Python callable D
 * cannot cross a C ABI, so models are TAGGED: (model id, 12-double parameter
 * block) -- the six families the north star names plus the two closed forms
 * the reference's own tests use.  Each entry point below names the reference
 * interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA-runtime types in signatures
 *     (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - `*_batch` entry points take DEVICE pointers owned by the caller and are
 *     asynchronous on `stream`; scratch (queue, per-problem bisection state) is taken
 *     from and returned to the stream-ordered memory pool inside the call;
 *   - `*_host` entry points take HOST pointers, stage through device memory the
 *     library owns, and return after the results are back in host memory;
 *   - return 0 on success, negative on error; fx_last_error() gives the text
 *     (thread-local).  There is no CPU fallback: without a CUDA device every
 *     compute entry point returns FX_ERR_CUDA.
 */
#ifndef FRONTX_B200_H
#define FRONTX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FX_VERSION 100 /* 0.1.0 */

#define FX_NPARAM 12 /* doubles per problem in `params` */
#define FX_NSUM 8    /* doubles per problem in `summary` */
#define FX_NCNT 8    /* int32 per problem in `counters` */
#define FX_KNOT 4    /* doubles per dense-output knot: o, theta, dtheta/do, d2theta/do2 */

/* Diffusivity models.  Parameter-block layouts (unused slots ignored):
 *   FX_CONSTANT       {D0}                                   D = D0            (fronts.D.constant)
 *   FX_POWER_LAW      {a, k, eps}                            D = a*theta^k+eps (fronts.D.power_law;
 *                                                            tests/test_solve.py:151 D=theta)
 *   FX_BROOKS_COREY   {n, l, Ks, alpha, theta_r, theta_s}    src/frontx/models.py:126-174
 *   FX_VAN_GENUCHTEN  {n, m, l, Ks, alpha, theta_r, theta_s, form}  src/frontx/models.py:177-253
 *                     (m = 1 - 1/n; form = 0: 1 - Se^(1/m) by the literal subtraction of :245,253,
 *                      form = 1: the same value without cancellation, -expm1(log(Se)/m))
 *   FX_LETXS          {Lw, Ew, Tw, Ls, Es, Ts, Ks, alpha, theta_r, theta_s}  models.py:256-315
 *   FX_LETD           {L, E, T, Dwt, theta_r, theta_s}       src/frontx/models.py:35-66
 *   FX_LOG            {c0, c1}      D = c0 + c1*ln(theta)    tests/test_solve.py:40 (Philip 1960)
 *   FX_TABULATED      {table, n}    D = -(do/dtheta (Iodtheta - c))/2 of a measured profile: what
 *                                   InterpolatedSolution.D is (src/frontx/_inverse/interpolated.py:59-67), so that it
 *                                   can be fed back into solve (tests/test_inverse.py:42-48).  `table` is a DEVICE
 *                                   pointer, bit-cast into the double, to 4 rows of n doubles: theta ascending, o,
 *                                   do/dtheta at the knots, Iodtheta - c at the knots (fx_unique_batch +
 *                                   fx_inverse_batch produce them); it must outlive the call.  NaN outside
 *                                   [theta_0, theta_n-1] (extrapolate=False).
 */
enum fx_model {
  FX_CONSTANT = 0,
  FX_POWER_LAW = 1,
  FX_BROOKS_COREY = 2,
  FX_VAN_GENUCHTEN = 3,
  FX_LETXS = 4,
  FX_LETD = 5,
  FX_LOG = 6,
  FX_TABULATED = 7,
  FX_NUM_MODELS = 8
};

/* per-problem status in counters[.][0]; mirrors src/frontx/_forward.py:114-122:
 * 0 -> RESULTS.successful, anything else -> RESULTS.max_steps_reached */
enum fx_result {
  FX_SUCCESSFUL = 0,
  FX_MAX_STEPS_REACHED = 1, /* bisection used max_bisect steps without |residual| < itol */
  FX_NO_BRACKET = 2,        /* bracket expansion exhausted without a sign change          */
  FX_INVALID_MODEL = -1,    /* model id outside enum fx_model: the problem was not solved  */
  FX_UNFINISHED = -3        /* only if a launch ended early on an internal queue error     */
};

enum fx_error {
  FX_OK = 0,
  FX_ERR_ARG = -1,
  FX_ERR_CUDA = -2,
  FX_ERR_MODEL = -3
};

/* summary[j]  = {d_dob, oi, theta(oi), S0 = -2 D(b) d_dob, D(b), residual, lower, upper}
 * counters[j] = {result, n_shots, n_expansions, n_attempts, n_rejected, n_rhs, n_jac, n_knots}
 *   (n_attempts .. n_jac are summed over all shots; n_knots is the LAST shot's, which is the
 *    one the reference returns as Solution -- _forward.py:109,115)                          */

typedef struct fx_solve_options {
  double itol;           /* 1e-3   _forward.py:69  */
  int32_t max_bisect;    /* 100    _forward.py:70  */
  int32_t max_ode_steps; /* 4096   diffrax.diffeqsolve default max_steps */
  double rtol;           /* 1e-3   _forward.py:92  */
  double atol;           /* 1e-6   _forward.py:92  */
  int32_t max_expansions; /* 64    bound on Bisection(expand_if_necessary=True) */
  int32_t radial_k;      /* 0 planar (the reference); 1 cylindrical/polar, 2 spherical */
  double ob;             /* 0 for the reference; >0 only with radial_k>0 */
  uint32_t model_mask;   /* optional hint: bit m set = problems of enum fx_model m may occur.  0 = unknown
                            (every model is launched; empty ones exit at once).  A problem whose model is
                            outside a non-zero mask is not solved and reports FX_INVALID_MODEL.           */
} fx_solve_options;

void fx_default_solve_options(fx_solve_options *opt);

/* Replaces frontx.solve (src/frontx/_forward.py:60-122) for a batch of n
 * independent problems: Kvaerno5 ESDIRK shots (diffrax, :84-95) with the boolean
 * event (:76-78), bisection on d_dob (:104-112).  Persistent lanes pop single
 * shots from a device-wide queue (fx_solve.cuh); idle lanes also run the next levels of a
 * problem's bisection tree speculatively (same results; with dense output the accepted shot is
 * then integrated once more to write its knots).  Safe to call from several host threads on
 * distinct streams.
 * k_max = 0 skips dense output; otherwise knots is [n][k_max][4] and
 * rows hold the last shot's accepted steps (n_knots may exceed k_max: the table
 * is then truncated and Solution evaluation refuses it).  All pointers device. */
int fx_solve_batch(int64_t n, const int32_t *model_id, const double *params /*[n][12]*/,
                   const double *b /*[n]*/, const double *i /*[n]*/,
                   const fx_solve_options *opt /*host; NULL = defaults*/,
                   double *summary /*[n][8]*/, int32_t *counters /*[n][8]*/,
                   int32_t k_max, double *knots /*[n][k_max][4] or NULL*/, void *stream);

/* Same, HOST pointers in and out (H2D, kernel, D2H inside the call). */
int fx_solve_batch_host(int64_t n, const int32_t *model_id, const double *params,
                        const double *b, const double *i, const fx_solve_options *opt,
                        double *summary, int32_t *counters, int32_t k_max, double *knots);

/* Radial problems with a FLOW-RATE boundary (north-star config C4).  Not in the reference:
 * semantics of fronts.solve_flowrate on the reference's algorithm.
 * The ODE has the radial term k/o (opt->radial_k = 1 cylindrical/polar, 2 spherical), the
 * boundary is at opt->ob > 0, the inlet value b is unknown and every shot starts from
 * (b, d_dob = -q/D(b)) with q[j] = Qb/(angle*ob*height) the imposed inlet flux density.  The
 * bisection of src/frontx/_forward.py:104-112 runs on b inside [b_lo, b_hi] (not widened:
 * FX_NO_BRACKET if the residuals at both ends have one sign) until |theta(oi) - i| < itol.
 *   summary[j] = {b, oi, theta(oi), d_dob = -q/D(b), D(b), residual, b_lower, b_upper}
 * counters, k_max and knots as in fx_solve_batch.  All pointers device. */
int fx_solve_flowrate_batch(int64_t n, const int32_t *model_id, const double *params /*[n][12]*/,
                            const double *q /*[n]*/, const double *i /*[n]*/,
                            const double *b_lo /*[n]*/, const double *b_hi /*[n]*/,
                            const fx_solve_options *opt /*host; NULL = defaults, cylindrical, ob=1e-6*/,
                            double *summary /*[n][8]*/, int32_t *counters /*[n][8]*/,
                            int32_t k_max, double *knots /*[n][k_max][4] or NULL*/, void *stream);

/* Replaces Solution.__call__ / Solution.d_do (src/frontx/_forward.py:25-37):
 * clip o to [0, oi], locate the step, cubic Hermite dense output.  o_query is
 * [n][q] (per_problem=1) or [q] shared by all problems (per_problem=0).
 * theta / dtheta_do are [n][q]; either may be NULL.  All pointers device. */
int fx_eval_batch(int64_t n, const int32_t *counters /*[n][8]*/, const double *knots, int32_t k_max,
                  int64_t q, const double *o_query, int32_t per_problem, double *theta,
                  double *dtheta_do, void *stream);

/* Replaces, for a batch of solved candidates, ScaledSolution.fitting_data (src/frontx/_inverse/fit.py:45-82:
 * the one-parameter least-squares fit of D0 in theta(o/sqrt(D0)) to m data points) and the cost of fit()
 * (:145-151: mean(((sol(o) - theta)/sigma)^2), inf for a candidate whose solve failed).
 *   counters, summary, knots, k_max  as written by fx_solve_batch (dense output required)
 *   o, theta [m], weight [m] = 1/sigma or NULL   the data, shared by all candidates
 *   mode   0: D0 = 1;  1: fit D0 (start (o[m-1]/oi)^2, :66);  2: D0 given in D0[n] (e.g. from the sorptivity, :36-43)
 *   D0 [n] in/out, cost [n], status [n] (0 converged, 1 max steps or failed solve, 2 no descent direction)
 * All pointers device. */
int fx_scaled_fit_batch(int64_t n, const int32_t *counters, const double *summary, const double *knots,
                        int32_t k_max, int64_t m, const double *o, const double *theta, const double *weight,
                        int32_t mode, double *D0, double *cost, int32_t *status, void *stream);

/* D, dD/dtheta, d2D/dtheta2 at arbitrary theta (the callable the reference keeps
 * as Solution.D and uses in flux/sorptivity, src/frontx/_boltzmann.py:151-161).
 * theta is [n][q]; out is [n][q][3]. All pointers device. */
int fx_diffusivity_batch(int64_t n, const int32_t *model_id, const double *params, int64_t q,
                         const double *theta, double *out, void *stream);

/* Replaces InterpolatedSolution(o, theta) (src/frontx/_inverse/interpolated.py:18-73)
 * for `batch` profiles of `npts` points each, theta strictly monotone in o:
 *   D_at_knots [batch][npts]  InterpolatedSolution.D at every data point, caller's order  (:59-67)
 *   sorptivity [batch][4]     {PCHIP Int theta do over the data (:69-73), trapezoid Int theta do
 *                              (_inverse/sorptivity.py:16-19), PCHIP-extrapolated Int_0^o[0], 0};
 *                              optional (NULL to skip: the forward interpolant's slopes are then
 *                              not formed at all, a quarter of the kernel's time)
 *   status     [batch]        0 ok, 1 theta not strictly monotone (D is then meaningless)
 *   do_dtheta, Iodtheta [batch][npts], optional (NULL to skip): the inverse PCHIP's knot
 *                              slopes and antiderivative values (minus its value at theta=i),
 *                              which fx_inverse_query_batch needs.
 * All pointers device; asynchronous on stream. */
int fx_inverse_batch(int64_t batch, int64_t npts, const double *o, const double *theta,
                     double *D_at_knots, double *sorptivity, int32_t *status,
                     double *do_dtheta, double *Iodtheta, void *stream);

/* InterpolatedSolution.D at arbitrary theta (interpolated.py:59-67): interval search in the
 * profile's theta, cubic pieces of derivative() and antiderivative(); NaN outside the data
 * range (extrapolate=False, :47).  theta_query / D_out are [batch][q].  Device pointers. */
int fx_inverse_query_batch(int64_t batch, int64_t npts, const double *o, const double *theta,
                           const double *do_dtheta, const double *Iodtheta, int64_t q,
                           const double *theta_query, double *D_out, void *stream);

/* InterpolatedSolution.__call__ (interpolated.py:30,52-57): the forward PCHIP y(x) (x = o
 * strictly ascending, y = theta) at arbitrary x_query[batch][q], end polynomials
 * extrapolated (PchipInterpolator default). Device pointers. */
int fx_pchip_eval_batch(int64_t batch, int64_t npts, const double *x, const double *y, int64_t q,
                        const double *x_query, double *y_out, void *stream);

/* Value, derivative and antiderivative of a profile's PCHIP y(x) at arbitrary x, from the knot slopes
 * and knot antiderivative values fx_inverse_batch leaves behind (its do_dtheta / Iodtheta outputs; call
 * it with the roles of o and theta swapped to get those of the FORWARD interpolant theta(o)).  Replaces
 * InterpolatedSolution.d_do (src/frontx/_boltzmann.py:130-135 = the derivative of the forward PCHIP) and
 * the PPoly antiderivative behind InterpolatedSolution.sorptivity(o) (interpolated.py:69-73).
 *   x strictly monotone (either direction); anti[k] = Int from the LAST point to knot k, may be NULL
 *   x_query [batch][q] (per_profile=1) or [q]; value / derivative / antiderivative [batch][q], any NULL
 *   extrapolate != 0: the end cubics continue beyond the data (PchipInterpolator/PPoly default), else NaN.
 * All pointers device. */
int fx_pchip_query_batch(int64_t batch, int64_t npts, const double *x, const double *y, const double *slope,
                         const double *anti, int64_t q, const double *x_query, int32_t per_profile,
                         int32_t extrapolate, double *value, double *derivative, double *antiderivative,
                         void *stream);

/* jnp.unique(theta, return_index=True); o = o[indices] (interpolated.py:44-45) per profile: theta sorted
 * ascending, equal values collapsed onto their FIRST occurrence (-0.0 == +0.0, NaN last).  The compacted
 * rows are written to the front of o_out / theta_out [batch][npts] (and the original indices to index_out,
 * may be NULL); n_unique[batch] receives their lengths.  The sort is CUB's stable radix sort on an
 * order-preserving key; a profile that is already strictly monotone needs none of this (fx_inverse_batch
 * takes it in either direction).  All pointers device. */
int fx_unique_batch(int64_t batch, int64_t npts, const double *o, const double *theta, double *o_out,
                    double *theta_out, int32_t *index_out, int64_t *n_unique, void *stream);

/* measured FP64 FMA throughput of the current device, FLOP/s (roofline denominator) */
int fx_measure_fp64_peak(double *flops_per_s, void *stream);

/* Kernel timing for the roofline: when on, fx_solve_batch brackets each per-model
 * shoot_bisect_kernel launch with CUDA events on the stream it runs on;
 * fx_last_solve_kernel_ms waits for and returns the last such duration. */
void fx_set_kernel_timing(int on);
int fx_last_solve_kernel_ms(int model, float *ms);

/* number of kernel launches issued by this library in this process so far */
int64_t fx_launch_count(void);

const char *fx_last_error(void);
int fx_version(void);
int fx_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
