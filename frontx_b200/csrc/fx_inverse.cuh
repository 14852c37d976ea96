// fx_inverse.cuh -- profile -> D(theta) inverse as a single-pass scan kernel.
//
// Replaces InterpolatedSolution(o, theta).D evaluated at the knots and
// InterpolatedSolution.sorptivity (src/frontx/_inverse/interpolated.py:18-73;
// interpax.PchipInterpolator / PPoly.derivative / PPoly.antiderivative are ports
// of scipy's) and the trapezoid of frontx.sorptivity (_inverse/sorptivity.py:16-19):
//
//   inverse interpolant  o(theta): PCHIP on theta sorted ascending (:44-47)
//   D(theta_k) = -( do/dtheta(theta_k) * Int_i^theta_k o dtheta ) / 2        (:59-67)
//
// For theta strictly monotone in o the sort is a reversal, and Int_i^theta is a
// running sum of per-interval cubic-Hermite integrals taken from the far-field
// end (the last point, theta = i) towards the boundary.  All formulas below are
// written in that TRAVERSAL order r = 0..N-1 (r = 0 is the caller's last point)
// with SIGNED steps, which is term-by-term the sorted-order formula for a
// decreasing profile and its exact mirror for an increasing one.
//
// Algorithmic traffic: 16 B/point read (o, theta), 8 B/point written (D); every point is
// read once (plus a 3-point halo per warp), measured DRAM traffic equals it.
//
// Every WARP is on its own (there is no CTA barrier in the kernel): it takes tickets, in order, for
// tiles of 256 consecutive points of one profile, each THREAD owning INV_ITEMS = 8 consecutive points
// that it keeps in registers:
//   the warp stages its points with coalesced cp.async, a tile ahead (request_warp / collect_chunk)
//   steps (d theta, d o) of the chunk's intervals and the two either side
//   PCHIP slope do/dtheta at the chunk's points (and dtheta/do when the sorptivity is wanted)
//   per-interval cubic-Hermite integrals; running sum in the thread, one warp scan of the thread totals
//   publish the tile aggregate; whoever ARRIVES last in a group of 32 tiles publishes the group total,
//   whoever closes the last group of a supergroup of 32 groups the supergroup total
//   go on to the next two tiles, then add up what lies before the tile in its profile in a fixed
//   three-level order: earlier supergroup totals, earlier group totals of its supergroup, earlier
//   aggregates of its group (tickets are handed out in order, so every earlier tile is held by a
//   warp that publishes before it waits for anything: no deadlock, no serial chain, and the
//   association depends on the tile index alone: bit-reproducible run to run)
//   D = -d * (prefix + local)/2, written coalesced in the caller's point order.
//
// Arithmetic: an fp64 division costs ~25 issue slots on sm_100a (scripts/ubench_fp64.cu), and
// scipy's formulas written literally take 12 per point, which made the first version of this
// kernel issue-bound at 13 % of the HBM roofline; the second (secants shared through shuffles,
// striped shared-memory tile, 4 reciprocals per point) reached 21 % at 363 instructions per point.
// Here the harmonic mean is cleared of its divisions algebraically (pchip_interior_steps): ONE
// reciprocal per point and slope, no secant formed, no shuffle outside the scan.
//
// The kernel is a latency chain -- publish -> group total -> look-back -> next publish -- and every
// choice below that looks odd was measured on C5 (10^3 profiles x 10^6 points); DESIGN.md 3.4 has the
// list.  The CTA-tiled form of the same kernel (1024-point tiles, one barrier per tile, two-level
// look-back) ran at 6.2 ms where this one runs at 5.7 ms.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "fx_math.cuh"

namespace fx {

constexpr int INV_BLOCK = 128;
constexpr int INV_ITEMS = 8;                      // consecutive points owned by one thread
constexpr int INV_NW = INV_BLOCK / 32;            // independent warps per CTA
constexpr int INV_NAGG = 4;  // inverse integral, forward PCHIP integral, trapezoid, flagged-bad count
constexpr int INV_GROUP = 32; // tiles per look-back group (one warp reads a group)
#ifndef FX_INV_CTAS
#define FX_INV_CTAS 4  // resident CTAs per SM the register budget is set for
#endif
// compute_chunk: how many of a thread's nine slope steps may be in flight at once (D alone / with the
// sorptivity sums, which carry two slopes per step); measured flat between 1 and 9 on C5
#ifndef FX_INV_LAG
#define FX_INV_LAG 3
#endif
#ifndef FX_INV_LAG_SORP
#define FX_INV_LAG_SORP 2
#endif
#ifndef FX_INV_DEFER
#define FX_INV_DEFER 2
#endif
// 1: a tile finds its direction (theta rising or falling along the traversal) from its own points and
// publishes it as +1 / -1 (0: not monotone); the profile is monotone iff the signs of all its tiles add up
// to +-(number of tiles that hold an interval).  0: every tile loads the profile's last two values from
// global memory to learn the direction (two dependent L2 round trips per tile).
#ifndef FX_INV_TILESIGN
#define FX_INV_TILESIGN 1
#endif
constexpr int INV_DEFER = FX_INV_DEFER;  // a tile's output is written this many tiles later
static_assert(INV_DEFER >= 1 && INV_DEFER <= 2, "the kernel loop's exit test looks at the first and last deferred tile only");

__device__ __forceinline__ double sgn(double v) { return (double)((v > 0.0) - (v < 0.0)); }

// scipy PchipInterpolator._find_derivatives, interior point: harmonic mean of the
// neighbouring secants, weighted by the steps
__device__ __forceinline__ double pchip_interior(double h0, double h1, double m0, double m1) {
  if ((sgn(m0) != sgn(m1)) || (m0 == 0.0) || (m1 == 0.0)) return 0.0;
  double w1 = 2.0 * h1 + h0;
  double w2 = h1 + 2.0 * h0;
  double whmean = (w1 / m0 + w2 / m1) / (w1 + w2);
  return 1.0 / whmean;
}
// scipy PchipInterpolator._edge_case: one-sided three-point estimate, shape preserving
__device__ __forceinline__ double pchip_edge(double h0, double h1, double m0, double m1) {
  double d = ((2.0 * h0 + h1) * m0 - h0 * m1) / (h0 + h1);
  if (sgn(d) != sgn(m0)) return 0.0;
  if ((sgn(m0) != sgn(m1)) && (fabs(d) > 3.0 * fabs(m0))) return 3.0 * m0;
  return d;
}

// 1/x: hardware seed (relative error 2^-23) and one cubic correction, y(1 + e + e^2): the remaining
// error, e^3 plus two roundings, is below 1.5 ulp.  A zero or non-finite argument gives inf/NaN,
// which every caller discards by a select.
__device__ __forceinline__ double rcp_lean(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  return fma(y, e, y);
}

// Interior PCHIP slope of y(x) at a point from the steps either side, (h0, g0) = (dx, dy) of the
// interval before it and (h1, g1) of the one after.  scipy's weighted harmonic mean
//   (w1 + w2) / (w1/m0 + w2/m1),  m = g/h,  w1 = 2 h1 + h0,  w2 = h1 + 2 h0
// cleared of its three divisions: (w1 + w2) g0 g1 / (w1 h0 g1 + w2 h1 g0) -- one reciprocal per point
// and no secant is ever formed.  `same` says the two secants have the same strict sign.
__device__ __forceinline__ double pchip_interior_steps(double h0, double g0, double h1, double g1, bool same) {
  const double w1 = fma(2.0, h1, h0);
  const double w2 = fma(2.0, h0, h1);
  const double num = ((w1 + w2) * g0) * g1;
  const double den = fma(w2 * h1, g0, (w1 * h0) * g1);
  const double v = num * rcp_lean(den);
  return same ? v : 0.0;
}

__device__ __forceinline__ double warp_incl_scan(double v, unsigned lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double u = __shfl_up_sync(0xffffffffu, v, d);
    if ((int)lane >= d) v += u;
  }
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Aggregates are published without flags: the scratch is preset to an all-ones pattern (a NaN no
// computation produces) and a reader spins until the words it needs differ from it.  One L2 round
// trip when the value is there instead of three (flag, fence, data), no fence on the reader's
// side (each 8-byte access is atomic): the look-back sits on every tile's critical path.
__device__ __forceinline__ bool is_unset(double v) { return __double_as_longlong(v) == -1LL; }
// ... so nothing published may BE that pattern: a quiet NaN keeps its payload through additions, and a
// profile that holds the all-ones NaN (a buffer preset to 0xff, a missing-data marker) would otherwise
// publish totals that read as "not there yet" and leave every later tile of the profile spinning.
__device__ __forceinline__ double published(double v) {
  return is_unset(v) ? __longlong_as_double(0x7ff8000000000000LL) : v;
}
// A value that is not there yet is polled at a short fixed pause.  The waits sit on the kernel's
// critical chain (publish -> group total -> look-back -> next publish): pauses of 128 ns and more,
// or a growing pause, were measured two to three times slower.
#ifndef FX_INV_NAP
#define FX_INV_NAP 20
#endif
__device__ __forceinline__ void wait_load4(const double* src, double (&v)[4]) {
  const unsigned nap = FX_INV_NAP;
  for (;;) {
#pragma unroll
    for (int c = 0; c < 4; c++) v[c] = __ldcg(src + c);
    if (!(is_unset(v[0]) || is_unset(v[1]) || is_unset(v[2]) || is_unset(v[3]))) break;
    __nanosleep(nap);
  }
}

__device__ __forceinline__ double wait_load1(const double* src) {
  double v;
  const unsigned nap = FX_INV_NAP;
  for (;;) {
    v = __ldcg(src);
    if (!is_unset(v)) break;
    __nanosleep(nap);
  }
  return v;
}

constexpr int INV_NP = INV_ITEMS + 3;  // points j = -1 .. 9 held by a thread, index a = j + 1
constexpr int INV_NI = INV_ITEMS + 2;  // intervals (j, j+1), j = -1 .. 8, index b = j + 1

// cp.async: global -> shared without passing through registers, so a tile's points can be requested
// a whole tile ahead at no register cost.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, bool ok) {  // !ok: zero-fill, nothing read
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// A WARP owns 32 * INV_ITEMS = 256 consecutive points of the tile and stages them itself, as 16-byte
// chunks of two points: lane l copies chunk 32 q + l, so one instruction moves 512 contiguous bytes
// (whole sectors, no reliance on L1), and thread u (in memory order, u = 31 - lane) then reads ITS
// chunks 4u-1 .. 4u+4.  The rows are padded by one chunk so that both the copy and the 16-byte
// reads at a 64-byte stride are free of bank conflicts.  Only the warp reads
// what it staged: cp.async.wait_all + __syncwarp, no CTA barrier.
static_assert(INV_ITEMS == 8, "the chunk swizzle is laid out for four chunks per thread");
constexpr int INV_CH = INV_ITEMS / 2;      // chunks per thread
constexpr int INV_WCH = 32 * INV_CH;       // chunks per warp; two more hold the halo either side
constexpr int INV_WPTS = 32 * INV_ITEMS;   // points per warp
// one 16-byte pad per 128-byte row: every address below is a per-thread constant plus an immediate
__device__ __forceinline__ int swz(int c) { return c + (c >> 3); }
constexpr int INV_WCHP = INV_WCH + INV_WCH / 8;  // padded chunks per warp

// Request the warp's points of one array.  wbase is the profile's memory index of the warp's lowest
// point (the one furthest along the traversal); chunk c holds memory indices wbase + 2c, + 2c + 1,
// chunk INV_WCH the two below wbase and chunk INV_WCH + 1 the two from wbase + 256.
// EDGE = false: every one of them exists (the bulk of a long profile); A16: chunks are 16-byte
// aligned (even npts, aligned arrays).  EDGE = true: first/last tiles and short profiles, guarded.
template <bool EDGE, bool A16>
__device__ __forceinline__ void request_warp(const double* __restrict__ g, long long n, long long wbase,
                                             double2* buf, unsigned lane) {
  const double* base = g + wbase;
  if (!EDGE && A16) {
#pragma unroll
    for (int q = 0; q < INV_CH; q++) {
      const int c = 32 * q + (int)lane;
      cp_async16(&buf[swz(c)], base + 2 * c);
    }
    // the halo; the profile's first tile has nothing above its first warp (and does not use it)
    if (lane == 0 || (lane == 1 && wbase + INV_WPTS < n))
      cp_async16(&buf[INV_WCHP + lane], lane == 0 ? base - 2 : base + INV_WPTS);
  } else {
#pragma unroll
    for (int q = 0; q <= INV_ITEMS; q++) {
      int e = 32 * q + (int)lane;          // element of the warp's run, memory order
      double* dst = reinterpret_cast<double*>(&buf[swz(e >> 1)]) + (e & 1);
      if (q == INV_ITEMS) {                // the halo: elements -2, -1 and 256
        e = lane == 2 ? INV_WPTS : (int)lane - 2;
        dst = lane == 2 ? &buf[INV_WCHP + 1].x : reinterpret_cast<double*>(&buf[INV_WCHP]) + lane;
      }
      const long long mi = wbase + e;
      const bool ok = EDGE ? (mi >= 0 && mi < n) : (mi < n);
      if (q < INV_ITEMS || lane < 3) cp_async8(dst, ok ? base + e : g, ok);
    }
  }
}
// ... and each thread picks up its own: x[a] is point j = a - 1 of the thread
__device__ __forceinline__ void collect_chunk(const double2* buf, unsigned lane, double (&x)[INV_NP]) {
  const int u = 31 - (int)lane;
  {
    const double2 v = (u == 0) ? buf[INV_WCHP] : buf[swz(INV_CH * u - 1)];
    x[INV_NP - 1] = v.x;
    x[INV_NP - 2] = v.y;
  }
#pragma unroll
  for (int i = 0; i < INV_CH; i++) {
    const double2 v = buf[swz(INV_CH * u) + i];  // the thread's four chunks share a row
    x[INV_ITEMS - 2 * i] = v.x;
    x[INV_ITEMS - 1 - 2 * i] = v.y;
  }
  x[0] = ((u == 31) ? buf[INV_WCHP + 1] : buf[swz(INV_CH * u + INV_CH)]).x;
}

// What one thread makes of its points:
//   d[j]          do/dtheta at point j = 0..8 (8 closes the thread's last interval)
//   I[j]          Int o dtheta over the interval (j, j+1), j = 0..7
//   sumF, sumT    the thread's share of the forward PCHIP and trapezoid integrals of theta do
//   mono          whether every interval of the thread steps in theta the way the profile does (up)
// EDGE = true applies the end-point formulas and drops what lies beyond the profile.
template <bool EDGE, bool SORP>
__device__ __forceinline__ void compute_chunk(const double* __restrict__ th_g, const double* __restrict__ o_g,
                                              long long n, long long rb, const double (&x)[INV_NP],
                                              const double (&y)[INV_NP], double (&d)[INV_ITEMS + 1],
                                              double (&I)[INV_ITEMS], double& sumF, double& sumT, bool& up,
                                              bool& mono) {
  // One pass along the thread's points, each step using what the one before left behind: the interval
  // after point j (its steps and a number with the sign of its secant -- steps of 1e-77 would have to
  // meet for that product to underflow), the slope at j, the integral over the interval before j.
  // The empty asm statements pin this order: left to itself the compiler interleaves all nine slopes
  // and needs 170 registers for it.
  sumF = 0.0;
  sumT = 0.0;
  mono = true;
  double hc = x[1] - x[0], gc = y[1] - y[0];  // interval before point 0
  double sc = hc * gc;
  double dprev = 0.0, dfprev = 0.0;
#pragma unroll
  for (int j = 0; j <= INV_ITEMS; j++) {
    double xn = x[j + 2], yn = y[j + 2];
    {  // step j starts when the slope of step j - LAG is known: LAG steps in flight
      constexpr int LAG = SORP ? FX_INV_LAG_SORP : FX_INV_LAG;
      const double dep = j >= LAG ? d[j - LAG] : 0.0;
      asm volatile("" : "+d"(xn), "+d"(yn) : "d"(dep));
    }
    const double hn = xn - x[j + 1], gn = yn - y[j + 1];  // interval after point j
    const double sn = hn * gn;
    const bool same = sc * sn > 0.0;  // the two secants have the same strict sign
    double dj = pchip_interior_steps(hc, gc, hn, gn, same);
    double dfj = 0.0;
    if (SORP) dfj = pchip_interior_steps(gc, hc, gn, hn, same);
    if (!EDGE && j == 0) {
      // the profile's first point in an otherwise ordinary tile: one-sided, intervals (0,1) and (1,2).
      // Every later tile of the profile waits for this tile's aggregate, so it takes the fast path too.
      if (rb == 0) {
        const double h2 = x[3] - x[2], g2 = y[3] - y[2];
        dj = pchip_edge(hn, h2, gn / hn, g2 / h2);
        if (SORP) dfj = pchip_edge(gn, g2, hn / gn, h2 / g2);
      }
    }
    if (EDGE) {
      const long long r = rb + j;
      const bool live = r < n, has_left = r >= 1, has_right = r <= n - 2;
      if (!live) {
        dj = 0.0;
        dfj = 0.0;
      } else if (!(has_left && has_right)) {
        if (n == 2) {
          dj = has_right ? gn / hn : gc / hc;
          if (SORP) dfj = has_right ? hn / gn : hc / gc;
        } else if (has_right) {  // r == 0, one-sided: intervals (0,1) and (1,2), both staged
          const int a2 = j + 3 < INV_NP ? j + 3 : INV_NP - 1;  // r == 0 is the j = 0 of one thread
          const double h2 = x[a2] - x[a2 - 1], g2 = y[a2] - y[a2 - 1];
          dj = pchip_edge(hn, h2, gn / hn, g2 / h2);
          if (SORP) dfj = pchip_edge(gn, g2, hn / gn, h2 / g2);
        } else {  // r == n-1, one-sided, mirrored: intervals (n-2,n-1) and (n-3,n-2)
          const double x0 = th_g[0], x1 = th_g[1], x2 = th_g[2];
          const double y0 = o_g[0], y1 = o_g[1], y2 = o_g[2];
          const double a1 = x0 - x1, a2_ = x1 - x2, c1 = y0 - y1, c2 = y1 - y2;
          dj = pchip_edge(a1, a2_, c1 / a1, c2 / a2_);
          if (SORP) dfj = pchip_edge(c1, c2, a1 / c1, a2_ / c2);
        }
      }
    }
    d[j] = dj;
    if (j >= 1) {  // cubic-Hermite integral over the interval (j-1, j), whose steps are (hc, gc)
      double v = hc * fma(0.5, y[j] + y[j + 1], (dprev - dj) * (hc * (1.0 / 12.0)));
      bool exists = true;
      if (EDGE) exists = rb + j - 1 <= n - 2;
      v = exists ? v : 0.0;
#if FX_INV_TILESIGN
      if (j == 1) up = hc > 0.0;  // the thread's own direction: its first interval (the tile compares the threads')
#endif
      if (exists) mono = mono && (up ? hc > 0.0 : hc < 0.0);
      if (SORP) {
        const double trap = gc * (0.5 * (x[j] + x[j + 1]));
        if (exists) {
          sumT += trap;
          sumF += fma((dfprev - dfj) * gc, gc * (1.0 / 12.0), trap);
        }
      }
      I[j - 1] = v;
    }
    hc = hn;
    gc = gn;
    sc = sn;
    dprev = dj;
    dfprev = dfj;
  }
}

// Int_0^{o_first} of the forward PCHIP's first cubic extrapolated back to o = 0 (scipy extrapolate=True)
__device__ __forceinline__ double pchip_extrapolated_head(const double* o_g, const double* th_g, long long n) {
  double ext = 0.0;
  double o0 = o_g[0];
  if (o0 != 0.0 && n >= 2) {
    double o1 = o_g[1], y0 = th_g[0], y1 = th_g[1];
    double hh = o1 - o0, m = (y1 - y0) / hh;
    double d0, d1;
    if (n == 2) {
      d0 = d1 = m;
    } else {
      double o2 = o_g[2], y2 = th_g[2];
      double h1 = o2 - o1, m1 = (y2 - y1) / h1;
      d0 = pchip_edge(hh, h1, m, m1);
      d1 = (n == 3) ? pchip_edge(h1, hh, m1, m) : pchip_interior(hh, h1, m, m1);
    }
    double tt = (d0 + d1 - 2.0 * m) / hh;
    double c0 = tt / hh, c1 = (m - d0) / hh - tt, c2 = d0, c3 = y0;
    double u = 0.0 - o0;  // antiderivative at o=0 relative to o0
    double anti = (((c0 * 0.25 * u + c1 * (1.0 / 3.0)) * u + c2 * 0.5) * u + c3) * u;
    ext = -anti;
  }
  return ext;
}

// Three fixed look-back levels: aggregates of the earlier tiles of the group of 32, totals of the earlier
// groups of the supergroup of 32 groups, totals of the earlier supergroups.  Whoever ARRIVES last in a
// group (supergroup) publishes its total.  (Letting the tile that is last by index wait for the others was
// measured at 450 ms: it waits, in line, for tiles that other warps hold as their NEXT ticket, and
// everything behind its own next ticket waits with it.)
struct InvArgs {
  long long batch, npts, ntiles;  // tiles of INV_WPTS points per profile
  double inv_ntiles;              // 1 / ntiles: a ticket is split into (profile, tile) without a 64-bit division
  const double* o;
  const double* theta;
  double* D;
  double* sorp;
  int* status;
  double* slope;
  double* anti;
  double* agg;    // [batch][ntiles][INV_NAGG]
  double* gtot;   // [batch][ngroups][INV_NAGG], full groups only
  double* stot;   // [batch][nsuper][INV_NAGG], full supergroups only
  long long ngroups, nsuper;
  int* gcount;    // [batch][ngroups] arrivals
  int* scount;    // [batch][nsuper] closed groups
  unsigned long long* ticket;
};

template <int NC, bool BATCHED>
__device__ __forceinline__ void look_back3(const InvArgs& A, long long p, long long t, unsigned lane,
                                           double (&part)[NC]) {
  const long long G = t / INV_GROUP, S = G / INV_GROUP;
  const long long k = G * INV_GROUP + lane, g = S * INV_GROUP + lane;
  const double* sa = A.agg + (p * A.ntiles + k) * INV_NAGG;
  const double* sg = A.gtot + (p * A.ngroups + g) * INV_NAGG;
  const double* ss = A.stot + (p * A.nsuper + lane) * INV_NAGG;
  const bool need_a = k < t, need_g = g < G, need_s = (long long)lane < S;
  double va[NC], vg[NC], vs[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) {  // all three levels requested at once
    vs[c] = need_s ? __ldcg(ss + c) : 0.0;
    vg[c] = need_g ? __ldcg(sg + c) : 0.0;
    va[c] = need_a ? __ldcg(sa + c) : 0.0;
  }
#pragma unroll
  for (int c = 0; c < NC; c++) {
    if (is_unset(vs[c])) vs[c] = wait_load1(ss + c);
    part[c] = vs[c];
  }
  // Profiles beyond 32 supergroups (8.4e6 points).  Two forms of the same sum, same order of additions:
  // four loads in flight, or one at a time.  This loop never runs on C5, yet which form is compiled in moves
  // the whole kernel by 8 % (with the sorptivity sums: 7.0 vs 7.6 ms) or 20 % (D alone: 7.1 vs 5.7 ms) --
  // ptxas schedules the surrounding code differently -- so each instantiation gets the form it is faster with.
  if (BATCHED) {
    for (long long si = lane + 32; si < S; si += 128) {  // four
      double v[4][NC];                                    // loads in flight, added in the order of the index
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double* src = A.stot + (p * A.nsuper + si + 32 * q) * INV_NAGG;
#pragma unroll
        for (int c = 0; c < NC; c++) v[q][c] = (si + 32 * q < S) ? __ldcg(src + c) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double* src = A.stot + (p * A.nsuper + si + 32 * q) * INV_NAGG;
#pragma unroll
        for (int c = 0; c < NC; c++) {
          if (is_unset(v[q][c])) v[q][c] = wait_load1(src + c);
          part[c] += v[q][c];
        }
      }
    }
  } else {
    for (long long si = lane + 32; si < S; si += 32) {
      const double* src = A.stot + (p * A.nsuper + si) * INV_NAGG;
#pragma unroll
      for (int c = 0; c < NC; c++) part[c] += wait_load1(src + c);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; c++) {
    if (is_unset(vg[c])) vg[c] = wait_load1(sg + c);
    part[c] += vg[c];
    if (is_unset(va[c])) va[c] = wait_load1(sa + c);
    part[c] += va[c];
  }
#pragma unroll
  for (int c = 0; c < NC; c++) part[c] = warp_sum(part[c]);
}

template <bool SORP, bool A16>
__global__ void __launch_bounds__(INV_BLOCK, FX_INV_CTAS) inverse_scan_kernel(const InvArgs A) {
  __shared__ double2 s_in[2][INV_NW][INV_WCHP + 2];  // the next tile's (theta, o), each warp's own points
  extern __shared__ double2 s_dyn[];  // [INV_DEFER][2][INV_NW][INV_WCH]: slopes and in-tile prefixes of the deferred tiles

  const int tid = threadIdx.x;
  const unsigned lane = tid & 31, wid = tid >> 5;
  auto s_dd = [&](int slot) { return s_dyn + ((size_t)(slot * 2 + 0) * INV_NW + wid) * INV_WCH; };
  auto s_ee = [&](int slot) { return s_dyn + ((size_t)(slot * 2 + 1) * INV_NW + wid) * INV_WCH; };
  const long long total_tiles = A.batch * A.ntiles;
  const long long n = A.npts;
  auto bulk_tile = [&](long long t) {  // the points up to r0+256+1 exist, r0+256 is not the last
    const long long r0 = t * INV_WPTS;
    return r0 + INV_WPTS + 2 <= n && n > 2;
  };
  // Lane 0 takes the ticket with a bare atom (atomicAdd's warp-aggregated form waits for the answer on
  // the spot); nothing touches the answer until the tile in hand is done, then every lane decodes it.
  auto take_ticket = [&]() {
    unsigned long long tk = 0;
    if (lane == 0) asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(tk) : "l"(A.ticket) : "memory");
    return tk;
  };
  long long p = 0, t = 0;
#ifndef FX_INV_FASTDIV
#define FX_INV_FASTDIV 1
#endif
  auto decode_ticket = [&](unsigned long long raw) {
    const long long tk = (long long)__shfl_sync(0xffffffffu, raw, 0);
#if FX_INV_FASTDIV
    // tk / ntiles by a double-precision reciprocal and one correction step (exact: tickets stay far below
    // 2^52); the 64-bit integer division is a ~70-instruction subroutine, once per tile and thread
    long long pq = (long long)((double)tk * A.inv_ntiles);
    long long tr = tk - pq * A.ntiles;
    if (tr < 0) { pq--; tr += A.ntiles; }
    if (tr >= A.ntiles) { pq++; tr -= A.ntiles; }
    const bool live = tk < total_tiles;
    p = live ? pq : A.batch;
    t = live ? tr : tk - A.batch * A.ntiles;
#else
    p = tk < total_tiles ? tk / A.ntiles : A.batch;
    t = tk - p * A.ntiles;
#endif
  };
  decode_ticket(take_ticket());
  bool have = p < A.batch;
  double dir_a = 0.0, dir_b = 0.0;
  auto request_tile = [&]() {
    const long long wbase = n - t * INV_WPTS - INV_WPTS;
    const double* th_g = A.theta + p * n;
    const double* o_g = A.o + p * n;
#if !FX_INV_TILESIGN
    dir_a = th_g[n - 2];
    dir_b = th_g[n - 1];
#endif
    if (bulk_tile(t)) {
      request_warp<false, A16>(th_g, n, wbase, s_in[0][wid], lane);
      request_warp<false, A16>(o_g, n, wbase, s_in[1][wid], lane);
    } else {
      request_warp<true, false>(th_g, n, wbase, s_in[0][wid], lane);
      request_warp<true, false>(o_g, n, wbase, s_in[1][wid], lane);
    }
  };
  if (have) request_tile();

  long long pv_p[INV_DEFER], pv_t[INV_DEFER];
#pragma unroll
  for (int q = 0; q < INV_DEFER; q++) pv_p[q] = -1, pv_t[q] = 0;
  int slot = 0;

  for (;;) {
    if (!have && pv_p[0] < 0 && pv_p[INV_DEFER - 1] < 0) break;
    unsigned long long next_raw = ~0ull >> 1;  // "none left"
    if (have) next_raw = take_ticket();  // one tile ahead; looked at only after this tile's arithmetic
    double d[INV_ITEMS + 1], excl[INV_ITEMS];
    int arrived = 0;
    if (have) {
      const long long rb = t * INV_WPTS + (long long)lane * INV_ITEMS;
      const double* th_g = A.theta + p * n;
      const double* o_g = A.o + p * n;
      const double dir_step = dir_a - dir_b;
      double x[INV_NP], y[INV_NP];
      cp_async_wait_all();
      __syncwarp();
      collect_chunk(s_in[0][wid], lane, x);
      collect_chunk(s_in[1][wid], lane, y);
      double I[INV_ITEMS], sumF, sumT;
      bool mono;
      bool up = dir_step > 0.0;
      if (bulk_tile(t)) compute_chunk<false, SORP>(th_g, o_g, n, rb, x, y, d, I, sumF, sumT, up, mono);
      else compute_chunk<true, SORP>(th_g, o_g, n, rb, x, y, d, I, sumF, sumT, up, mono);
#if FX_INV_TILESIGN
      // +1: every interval of the tile rises along the traversal, -1: every one falls, 0: neither
      // (threads beyond the profile's end hold no interval and do not vote)
      const bool has = rb <= n - 2;
      const unsigned hasm = __ballot_sync(0xffffffffu, has);
      const unsigned upm = __ballot_sync(0xffffffffu, has && up);
      const bool allmono = __all_sync(0xffffffffu, !has || mono);
      const double tile_sign = (allmono && hasm) ? (upm == hasm ? 1.0 : (upm == 0u ? -1.0 : 0.0)) : 0.0;
#else
      const bool anybad = __any_sync(0xffffffffu, !(dir_step != 0.0) || !mono);
      const double tile_sign = anybad ? 1.0 : 0.0;
#endif

      double run = 0.0;
#pragma unroll
      for (int j = 0; j < INV_ITEMS; j++) {
        excl[j] = run;
        run += I[j];
      }
      const double wincl = warp_incl_scan(run, lane);
      const double toff = wincl - run;
#pragma unroll
      for (int j = 0; j < INV_ITEMS; j++) excl[j] += toff;
      const double carry = __shfl_sync(0xffffffffu, wincl, 31);
      double f = 0.0, z = 0.0;
      if (SORP) {
        f = warp_sum(sumF);  // the two sorptivity integrands (fixed tree)
        z = warp_sum(sumT);
      }
      // ---- publish; the last tile of a group (supergroup) closes it ---------------------------------
      if (lane == 0) {
        double* ag = A.agg + (p * A.ntiles + t) * INV_NAGG;
        __stcg(ag + 0, published(carry));
        __stcg(ag + 1, published(f));
        __stcg(ag + 2, published(z));
        __stcg(ag + 3, tile_sign);
      }
      // whoever arrives last in a full group closes it (nobody ever waits for a particular tile here);
      // the count comes back while the next tile is being requested
      if (lane == 0 && t / INV_GROUP < A.ngroups) arrived = atomicAdd(A.gcount + p * A.ngroups + t / INV_GROUP, 1);
    }

    // ---- the points of the next tile are requested before anything else is waited for ---------------
    const long long this_p = p, this_t = t;
    const bool had = have;
    decode_ticket(next_raw);
    have = p < A.batch;
    if (have) request_tile();

    // ---- close the group, and the supergroup, if this warp's tile was the last of it to arrive --------
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (had && arrived == INV_GROUP - 1) {
      const long long G = this_t / INV_GROUP;
      double tot[INV_NAGG];
      wait_load4(A.agg + (this_p * A.ntiles + G * INV_GROUP + lane) * INV_NAGG, tot);
#pragma unroll
      for (int c = 0; c < INV_NAGG; c++) tot[c] = warp_sum(tot[c]);
      int arrived2 = 0;
      if (lane == 0) {
        double* gt = A.gtot + (this_p * A.ngroups + G) * INV_NAGG;
#pragma unroll
        for (int c = 0; c < INV_NAGG; c++) __stcg(gt + c, published(tot[c]));
        if (G / INV_GROUP < A.nsuper) arrived2 = atomicAdd(A.scount + this_p * A.nsuper + G / INV_GROUP, 1);
      }
      arrived2 = __shfl_sync(0xffffffffu, arrived2, 0);
      if (arrived2 == INV_GROUP - 1) {
        const long long S = G / INV_GROUP;
        wait_load4(A.gtot + (this_p * A.ngroups + S * INV_GROUP + lane) * INV_NAGG, tot);
#pragma unroll
        for (int c = 0; c < INV_NAGG; c++) tot[c] = warp_sum(tot[c]);
        if (lane == 0) {
          double* st = A.stot + (this_p * A.nsuper + S) * INV_NAGG;
#pragma unroll
          for (int c = 0; c < INV_NAGG; c++) __stcg(st + c, published(tot[c]));
        }
      }
    }

    // ---- the deferred tile -----------------------------------------------------------------------------
    if (pv_p[0] >= 0) {
      const long long p2 = pv_p[0], t2 = pv_t[0];
      const bool closes = (t2 == A.ntiles - 1);  // the last tile of a profile closes its sums
      double prefix, tot[INV_NAGG];
      if (closes) {
        look_back3<INV_NAGG, SORP>(A, p2, t2, lane, tot);
        prefix = tot[0];
      } else {
        double one[1];
        look_back3<1, SORP>(A, p2, t2, lane, one);
        prefix = one[0];
      }
      const long long wb2 = n - t2 * INV_WPTS - INV_WPTS;  // may be < 0 in the last tile
      __syncwarp();
      const int xl = ((int)lane & ~7) | (((int)lane & 7) ^ (((int)lane >> 3) & 3));
      const double2* sd = s_dd(slot) + xl;
      const double2* se = s_ee(slot) + xl;
      const long long at0 = p2 * n + wb2 + 2 * (long long)lane;
      if (A16 && wb2 >= 0 && !A.slope && !A.anti) {
        double2* Dp = reinterpret_cast<double2*>(A.D + at0);
#pragma unroll
        for (int q = 0; q < INV_CH; q++) {
          const double2 dd = sd[q * 32], ee = se[q * 32];
          Dp[32 * q] = make_double2(-(dd.x * (prefix + ee.x)) * 0.5, -(dd.y * (prefix + ee.y)) * 0.5);
        }
      } else {
#pragma unroll
        for (int q = 0; q < INV_CH; q++) {
          const long long mi = wb2 + 2 * (32 * q + (long long)lane);
          const double2 dd = sd[q * 32], ee = se[q * 32];
          const double ix = prefix + ee.x, iy = prefix + ee.y;
          const double2 Dv = make_double2(-(dd.x * ix) * 0.5, -(dd.y * iy) * 0.5);
          const long long at = at0 + 64 * q;
          if (A16) {
            if (mi >= 0) *reinterpret_cast<double2*>(A.D + at) = Dv;
          } else {
            if (mi >= 0) A.D[at] = Dv.x;
            if (mi + 1 >= 0) A.D[at + 1] = Dv.y;
          }
          if (A.slope) {
            if (mi >= 0) A.slope[at] = dd.x;
            if (mi + 1 >= 0) A.slope[at + 1] = dd.y;
          }
          if (A.anti) {
            if (mi >= 0) A.anti[at] = ix;
            if (mi + 1 >= 0) A.anti[at + 1] = iy;
          }
        }
      }
      if (closes && lane == 0) {
        const double* th_g = A.theta + p2 * n;
        const double* o_g = A.o + p2 * n;
        const double* own = A.agg + (p2 * A.ntiles + t2) * INV_NAGG;
        const double ssum = tot[3] + __ldcg(own + 3);
#if FX_INV_TILESIGN
        // the last tile holds no interval when it holds a single point
        const double voters = (double)(A.ntiles - (((n - 1) % INV_WPTS) == 0 ? 1 : 0));
        A.status[p2] = (fabs(ssum) == voters) ? 0 : 1;
#else
        A.status[p2] = (ssum != 0.0) ? 1 : 0;
#endif
        if (SORP) {
          double F = tot[1] + __ldcg(own + 1), Z = tot[2] + __ldcg(own + 2);
          double* out = A.sorp + p2 * 4;
          out[0] = -F;  // traversal runs from o_last down to o_first: Int_{o_first}^{o_last} theta do = -sum
          out[1] = -Z;
          out[2] = pchip_extrapolated_head(o_g, th_g, n);
          out[3] = 0.0;
        }
      }
    }

    if (had) {
      __syncwarp();  // the lanes' reads of the deferred tile above come before these writes
#pragma unroll
      for (int i = 0; i < INV_CH; i++) {  // chunk 4u + i holds the thread's points j = 7-2i, 6-2i
        const int u = 31 - (int)lane;
        const int c = INV_CH * u + (i ^ ((u >> 1) & 3));
        s_dd(slot)[c] = make_double2(d[INV_ITEMS - 1 - 2 * i], d[INV_ITEMS - 2 - 2 * i]);
        s_ee(slot)[c] = make_double2(excl[INV_ITEMS - 1 - 2 * i], excl[INV_ITEMS - 2 - 2 * i]);
      }
    }
#pragma unroll
    for (int q = 0; q + 1 < INV_DEFER; q++) pv_p[q] = pv_p[q + 1], pv_t[q] = pv_t[q + 1];
    pv_p[INV_DEFER - 1] = had ? this_p : -1;
    pv_t[INV_DEFER - 1] = this_t;
    slot = (slot + 1 == INV_DEFER) ? 0 : slot + 1;
  }
}

// InterpolatedSolution.D(theta_q) for arbitrary theta_q: one thread per query.
__global__ void inverse_query_kernel(long long batch, long long n, const double* __restrict__ o,
                                     const double* __restrict__ theta, const double* __restrict__ slope,
                                     const double* __restrict__ anti, long long q,
                                     const double* __restrict__ tq, double* __restrict__ out) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  long long total = batch * q;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    long long p = g / q;
    const double* th = theta + p * n;
    const double* oo = o + p * n;
    const double* dd = slope + p * n;
    const double* aa = anti + p * n;
    double x = tq[g];
    bool decreasing = th[0] > th[n - 1];
    double xmin = decreasing ? th[n - 1] : th[0], xmax = decreasing ? th[0] : th[n - 1];
    double res = nan;
    if (x >= xmin && x <= xmax) {
      // sorted ascending index k <-> caller index j = decreasing ? n-1-k : k
      // interval = searchsorted(x_sorted, x, side='right') - 1, clipped to [0, n-2]
      long long lo = 0, hi = n;
      while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        double v = th[decreasing ? n - 1 - mid : mid];
        if (v <= x) lo = mid + 1; else hi = mid;
      }
      long long k = lo - 1;
      if (k < 0) k = 0;
      if (k > n - 2) k = n - 2;
      long long j0 = decreasing ? n - 1 - k : k, j1 = decreasing ? j0 - 1 : j0 + 1;
      double x0 = th[j0], x1 = th[j1], y0 = oo[j0], y1 = oo[j1], d0 = dd[j0], d1 = dd[j1];
      double h = x1 - x0, m = (y1 - y0) / h;
      double t = (d0 + d1 - 2.0 * m) / h;
      double c0 = t / h, c1 = (m - d0) / h - t, c2 = d0, c3 = y0;
      double u = x - x0;
      double der = (3.0 * c0 * u + 2.0 * c1) * u + c2;
      double ant = (((c0 * 0.25 * u + c1 * (1.0 / 3.0)) * u + c2 * 0.5) * u + c3) * u;
      // aa[] holds Int_i^theta at the knots; for an increasing profile i is the largest theta
      res = -(der * (aa[j0] + ant)) * 0.5;
    }
    out[g] = res;
  }
}

// Value, derivative and antiderivative of the PCHIP y(x) of a profile at arbitrary x, from the knot slopes
// and knot antiderivative values the scan kernel leaves behind (x strictly monotone, either direction;
// anti[k] = Int from the caller's LAST point to knot k).  extrapolate != 0: the end cubics continue beyond
// the data (scipy/interpax PchipInterpolator and PPoly default); else NaN there.  One thread per query.
__global__ void pchip_query_kernel(long long batch, long long n, const double* __restrict__ x,
                                   const double* __restrict__ y, const double* __restrict__ slope,
                                   const double* __restrict__ anti, long long q, const double* __restrict__ xq,
                                   int per_profile, int extrapolate, double* __restrict__ val,
                                   double* __restrict__ der, double* __restrict__ ant) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  long long total = batch * q;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    long long p = g / q;
    const double* xs = x + p * n;
    const double* ys = y + p * n;
    const double* dd = slope + p * n;
    const double* aa = anti ? anti + p * n : nullptr;
    const double v = xq[per_profile ? g : g - p * q];
    const bool decreasing = xs[0] > xs[n - 1];
    const double xmin = decreasing ? xs[n - 1] : xs[0], xmax = decreasing ? xs[0] : xs[n - 1];
    double rv = nan, rd = nan, ra = nan;
    if (v == v && (extrapolate || (v >= xmin && v <= xmax))) {
      long long lo = 0, hi = n;  // searchsorted(x_ascending, v, side='right') - 1, clipped to [0, n-2]
      while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (xs[decreasing ? n - 1 - mid : mid] <= v) lo = mid + 1; else hi = mid;
      }
      long long k = lo - 1;
      if (k < 0) k = 0;
      if (k > n - 2) k = n - 2;
      const long long j0 = decreasing ? n - 1 - k : k, j1 = decreasing ? j0 - 1 : j0 + 1;
      const double x0 = xs[j0], x1 = xs[j1], y0 = ys[j0], y1 = ys[j1], d0 = dd[j0], d1 = dd[j1];
      const double h = x1 - x0, m = (y1 - y0) / h;
      const double t = (d0 + d1 - 2.0 * m) / h;
      const double c0 = t / h, c1 = (m - d0) / h - t;
      const double u = v - x0;
      rv = ((c0 * u + c1) * u + d0) * u + y0;
      rd = (3.0 * c0 * u + 2.0 * c1) * u + d0;
      if (aa) ra = aa[j0] + (((c0 * 0.25 * u + c1 * (1.0 / 3.0)) * u + d0 * 0.5) * u + y0) * u;
    }
    if (val) val[g] = rv;
    if (der) der[g] = rd;
    if (ant) ant[g] = ra;
  }
}

// ---- jnp.unique(theta, return_index=True) (src/frontx/_inverse/interpolated.py:44-45) ----------------
// order-preserving 64-bit key of a double: -0.0 is +0.0, negatives below positives, NaN last
__global__ void unique_keys_kernel(long long n, const double* __restrict__ v, unsigned long long* __restrict__ key,
                                   int* __restrict__ idx) {
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v[g] + 0.0);
    key[g] = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    idx[g] = (int)g;
  }
}
__global__ void unique_gather_kernel(const int* __restrict__ count, const int* __restrict__ idx,
                                     const double* __restrict__ a, const double* __restrict__ b,
                                     double* __restrict__ a_out, double* __restrict__ b_out, int* __restrict__ idx_out,
                                     long long* __restrict__ n_out) {
  const long long m = *count;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < m; g += (long long)gridDim.x * blockDim.x) {
    const int j = idx[g];
    a_out[g] = a[j];
    b_out[g] = b[j];
    if (idx_out) idx_out[g] = j;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = m;
}

// slope of the PCHIP y(x) at knot j, read straight from global memory (x ascending or descending)
__device__ __forceinline__ double pchip_slope_global(const double* __restrict__ x,
                                                     const double* __restrict__ y, long long j, long long n) {
  if (n == 2) return (y[1] - y[0]) / (x[1] - x[0]);
  if (j == 0) {
    double h0 = x[1] - x[0], h1 = x[2] - x[1];
    return pchip_edge(h0, h1, (y[1] - y[0]) / h0, (y[2] - y[1]) / h1);
  }
  if (j == n - 1) {
    double h0 = x[n - 1] - x[n - 2], h1 = x[n - 2] - x[n - 3];
    return pchip_edge(h0, h1, (y[n - 1] - y[n - 2]) / h0, (y[n - 2] - y[n - 3]) / h1);
  }
  double h0 = x[j] - x[j - 1], h1 = x[j + 1] - x[j];
  return pchip_interior(h0, h1, (y[j] - y[j - 1]) / h0, (y[j + 1] - y[j]) / h1);
}

// InterpolatedSolution.__call__: the forward PCHIP theta(o) at arbitrary o, extrapolating
// with the end polynomials (scipy/interpax PchipInterpolator default). x strictly ascending.
__global__ void pchip_eval_kernel(long long batch, long long n, const double* __restrict__ x,
                                  const double* __restrict__ y, long long q,
                                  const double* __restrict__ xq, double* __restrict__ out) {
  long long total = batch * q;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    long long p = g / q;
    const double* xs = x + p * n;
    const double* ys = y + p * n;
    double v = xq[g];
    long long lo = 0, hi = n;
    while (lo < hi) {  // searchsorted(x, v, side='right')
      long long mid = (lo + hi) >> 1;
      if (xs[mid] <= v) lo = mid + 1; else hi = mid;
    }
    long long k = lo - 1;
    if (k < 0) k = 0;
    if (k > n - 2) k = n - 2;
    double x0 = xs[k], x1 = xs[k + 1], y0 = ys[k], y1 = ys[k + 1];
    double d0 = pchip_slope_global(xs, ys, k, n), d1 = pchip_slope_global(xs, ys, k + 1, n);
    double h = x1 - x0, m = (y1 - y0) / h;
    double t = (d0 + d1 - 2.0 * m) / h;
    double c0 = t / h, c1 = (m - d0) / h - t;
    double u = v - x0;
    out[g] = ((c0 * u + c1) * u + d0) * u + y0;
  }
}

inline cudaError_t inverse_launch(long long batch, long long npts, const double* o, const double* theta,
                                       double* D, double* sorp, int* status, double* slope, double* anti,
                                       int sms, cudaStream_t s, int* launches) {
  InvArgs a;
  a.batch = batch;
  a.npts = npts;
  a.ntiles = (npts + INV_WPTS - 1) / INV_WPTS;
  a.inv_ntiles = 1.0 / (double)a.ntiles;
  a.o = o;
  a.theta = theta;
  a.D = D;
  a.sorp = sorp;
  a.status = status;
  a.slope = slope;
  a.anti = anti;
  a.ngroups = a.ntiles / INV_GROUP;   // full groups
  a.nsuper = a.ngroups / INV_GROUP;   // full supergroups
  const size_t tiles = (size_t)batch * (size_t)a.ntiles;
  const size_t entries = tiles + (size_t)batch * (size_t)(a.ngroups + a.nsuper) + 1;
  const size_t agg_bytes = entries * INV_NAGG * sizeof(double);
  const size_t cnt_bytes = ((size_t)batch * (size_t)(a.ngroups + a.nsuper) * sizeof(int) + 15) / 8 * 8;
  char* scratch = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&scratch, agg_bytes + cnt_bytes + 64, s);
  if (e != cudaSuccess) return e;
  a.agg = reinterpret_cast<double*>(scratch);
  a.gtot = a.agg + tiles * INV_NAGG;
  a.stot = a.gtot + (size_t)batch * (size_t)a.ngroups * INV_NAGG;
  a.gcount = reinterpret_cast<int*>(scratch + agg_bytes);
  a.scount = a.gcount + (size_t)batch * (size_t)a.ngroups;
  a.ticket = reinterpret_cast<unsigned long long*>(scratch + agg_bytes + cnt_bytes);
  struct Guard {  // the scratch goes back to the pool, in stream order, on every way out
    char* p;
    cudaStream_t s;
    ~Guard() { cudaFreeAsync(p, s); }
  } guard{scratch, s};
  e = cudaMemsetAsync(scratch, 0xff, agg_bytes, s);  // the "unset" pattern (see is_unset)
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(scratch + agg_bytes, 0, cnt_bytes + 64, s);
  if (e != cudaSuccess) return e;
  const bool a16 = (npts % 2 == 0) && ((((uintptr_t)o | (uintptr_t)theta | (uintptr_t)D) & 15) == 0);
  void (*kern)(const InvArgs) =
      sorp ? (a16 ? inverse_scan_kernel<true, true> : inverse_scan_kernel<true, false>)
           : (a16 ? inverse_scan_kernel<false, true> : inverse_scan_kernel<false, false>);
  const size_t dyn = (size_t)INV_DEFER * 2 * INV_NW * INV_WCH * sizeof(double2);
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int bps = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, INV_BLOCK, dyn);
  if (e != cudaSuccess) return e;
  if (bps < 1) bps = 1;
  const long long cap = (long long)sms * bps;
  const long long want = ((long long)tiles + INV_NW - 1) / INV_NW;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  kern<<<grid, INV_BLOCK, dyn, s>>>(a);
  *launches = 1;
  return cudaGetLastError();
}

}  // namespace fx
