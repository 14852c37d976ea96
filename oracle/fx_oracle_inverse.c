/*
 * fx_oracle_inverse.c -- CPU restatement of frontx.InterpolatedSolution (the
 * profile -> D(theta) inverse) and frontx.sorptivity.
 *
 * TEST INFRASTRUCTURE ONLY (see fx_oracle.h).
 *
 * Follows /root/reference/src/frontx/_inverse/interpolated.py:18-73 and
 * src/frontx/_inverse/sorptivity.py:8-19.  The arithmetic lives in interpax
 * (>=0.3.8,<0.4; PchipInterpolator, PPoly.derivative/antiderivative), a JAX port
 * of scipy.interpolate's; scipy 1.18 IS installed here, so this restatement is
 * PINNED against scipy.interpolate.PchipInterpolator in tests/test_oracle_inverse.py
 * and against the reference's analytic tests (tests/test_inverse.py:29-62).
 * Sequential, scipy's evaluation order (power sums, running constants).
 */
#include "fx_oracle.h"

#include <math.h>
#include <stdlib.h>

static double sgn(double v) { return (double)((v > 0.0) - (v < 0.0)); }

/* scipy.interpolate.PchipInterpolator._edge_case */
static double edge_case(double h0, double h1, double m0, double m1) {
  double d = ((2.0 * h0 + h1) * m0 - h0 * m1) / (h0 + h1);
  if (sgn(d) != sgn(m0)) return 0.0;
  if ((sgn(m0) != sgn(m1)) && (fabs(d) > 3.0 * fabs(m0))) return 3.0 * m0;
  return d;
}

/* scipy.interpolate.PchipInterpolator._find_derivatives */
static void pchip_slopes(int64_t n, const double *x, const double *y, double *d) {
  if (n == 2) {
    d[0] = d[1] = (y[1] - y[0]) / (x[1] - x[0]);
    return;
  }
  for (int64_t k = 1; k < n - 1; k++) {
    double h0 = x[k] - x[k - 1], h1 = x[k + 1] - x[k];
    double m0 = (y[k] - y[k - 1]) / h0, m1 = (y[k + 1] - y[k]) / h1;
    if ((sgn(m0) != sgn(m1)) || m0 == 0.0 || m1 == 0.0) {
      d[k] = 0.0;
    } else {
      double w1 = 2.0 * h1 + h0, w2 = h1 + 2.0 * h0;
      double whmean = (w1 / m0 + w2 / m1) / (w1 + w2);
      d[k] = 1.0 / whmean;
    }
  }
  {
    double h0 = x[1] - x[0], h1 = x[2] - x[1];
    d[0] = edge_case(h0, h1, (y[1] - y[0]) / h0, (y[2] - y[1]) / h1);
    double g0 = x[n - 1] - x[n - 2], g1 = x[n - 2] - x[n - 3];
    d[n - 1] = edge_case(g0, g1, (y[n - 1] - y[n - 2]) / g0, (y[n - 2] - y[n - 3]) / g1);
  }
}

/* CubicHermiteSpline coefficients of interval k (highest power first) */
static void hermite_coeffs(const double *x, const double *y, const double *d, int64_t k, double *c) {
  double h = x[k + 1] - x[k];
  double slope = (y[k + 1] - y[k]) / h;
  double t = (d[k] + d[k + 1] - 2.0 * slope) / h;
  c[0] = t / h;
  c[1] = (slope - d[k]) / h - t;
  c[2] = d[k];
  c[3] = y[k];
}

/* value at local coordinate u of the antiderivative of interval k (constant excluded);
 * scipy evaluate_poly1 sums increasing powers */
static double anti_local(const double *c, double u) {
  double a3 = c[3], a2 = c[2] / 2.0, a1 = c[1] / 3.0, a0 = c[0] / 4.0;
  double z = u, res = 0.0;
  res += a3 * z; z *= u;
  res += a2 * z; z *= u;
  res += a1 * z; z *= u;
  res += a0 * z;
  return res;
}
static double deriv_local(const double *c, double u) {
  return c[2] + 2.0 * c[1] * u + 3.0 * c[0] * (u * u);
}

typedef struct { double th; int64_t idx; } sortkey_t;
static int cmp_key(const void *a, const void *b) {
  const sortkey_t *p = (const sortkey_t *)a, *q = (const sortkey_t *)b;
  if (p->th < q->th) return -1;
  if (p->th > q->th) return 1;
  return (p->idx > q->idx) - (p->idx < q->idx);
}

/* returns 0; D_at_knots in the caller's order (NaN for duplicates dropped by unique) */
int fxo_inverse(int64_t n, const double *o, const double *theta, double *D_at_knots,
                double *sorptivity_pchip, double *sorptivity_trapz) {
  if (n < 2) return -1;
  /* ---- forward interpolant theta(o) and its antiderivative (interpolated.py:30,69-73) ---- */
  double *df = (double *)malloc(sizeof(double) * (size_t)n);
  pchip_slopes(n, o, theta, df);
  double I_total = 0.0, c[4];
  for (int64_t k = 0; k < n - 1; k++) {
    hermite_coeffs(o, theta, df, k, c);
    I_total += anti_local(c, o[k + 1] - o[k]);
  }
  hermite_coeffs(o, theta, df, 0, c);
  double I_at_0 = anti_local(c, 0.0 - o[0]); /* extrapolated when o[0] != 0 */
  double oi = o[n - 1], i_far = theta[n - 1];
  if (sorptivity_pchip) *sorptivity_pchip = (I_total - I_at_0) - i_far * (oi - 0.0);
  if (sorptivity_trapz) { /* raw trapezoid of theta over o; callers add the (0,b) point and -i*oi */
    double s = 0.0;
    for (int64_t k = 0; k < n - 1; k++) s += (o[k + 1] - o[k]) * (theta[k + 1] + theta[k]) / 2.0;
    *sorptivity_trapz = s;
  }
  free(df);

  /* ---- theta, idx = unique(theta, return_index=True); o = o[idx]  (:44-45) ---- */
  sortkey_t *keys = (sortkey_t *)malloc(sizeof(sortkey_t) * (size_t)n);
  for (int64_t k = 0; k < n; k++) { keys[k].th = theta[k]; keys[k].idx = k; }
  qsort(keys, (size_t)n, sizeof(sortkey_t), cmp_key);
  int64_t m = 0;
  for (int64_t k = 0; k < n; k++)
    if (m == 0 || keys[k].th != keys[m - 1].th) keys[m++] = keys[k];
  double *xs = (double *)malloc(sizeof(double) * (size_t)m);
  double *ys = (double *)malloc(sizeof(double) * (size_t)m);
  double *d = (double *)malloc(sizeof(double) * (size_t)m);
  double *C = (double *)malloc(sizeof(double) * (size_t)m);
  for (int64_t k = 0; k < m; k++) { xs[k] = keys[k].th; ys[k] = o[keys[k].idx]; }
  for (int64_t k = 0; k < n; k++) D_at_knots[k] = NAN;
  if (m < 2) { free(keys); free(xs); free(ys); free(d); free(C); return 1; }

  /* ---- inverse = PchipInterpolator(theta, o); derivative(); antiderivative()  (:47-49) ---- */
  pchip_slopes(m, xs, ys, d);
  C[0] = 0.0;
  for (int64_t k = 0; k < m - 1; k++) {
    hermite_coeffs(xs, ys, d, k, c);
    C[k + 1] = C[k] + anti_local(c, xs[k + 1] - xs[k]); /* PPoly fix_continuity */
  }
  /* c = Iodtheta(i), i = theta[-1]  (:40,50): locate i among the sorted knots */
  double c_i = NAN;
  for (int64_t k = 0; k < m; k++)
    if (xs[k] == i_far) c_i = C[k];

  /* ---- D(theta_k) = -(do_dtheta * (Iodtheta - c))/2  (:59-67) ---- */
  for (int64_t k = 0; k < m; k++) {
    double der, anti;
    if (k < m - 1) {
      der = d[k];
      anti = C[k];
    } else { /* last knot: interval m-2 evaluated at its right end */
      hermite_coeffs(xs, ys, d, m - 2, c);
      double u = xs[m - 1] - xs[m - 2];
      der = deriv_local(c, u);
      anti = C[m - 2] + anti_local(c, u);
    }
    D_at_knots[keys[k].idx] = -(der * (anti - c_i)) / 2.0;
  }
  free(keys); free(xs); free(ys); free(d); free(C);
  return 0;
}
