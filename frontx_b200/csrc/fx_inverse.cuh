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
// read once (plus a 3-point halo per tile), measured DRAM traffic equals it.  One CTA owns a tile
// of TILE consecutive points of one profile:
//   load tile+halo (coalesced) -> shared memory
//   PCHIP slopes of o(theta) and of theta(o) for the tile (3-point stencils)
//   per-interval integrals; block scan (warp shuffles) of the inverse ones
//   publish the tile aggregate, go on to the next tile, then add up the aggregates of all
//   EARLIER tiles of the profile in a fixed two-level order (tickets are handed out in order,
//   so every earlier tile is resident or finished: no deadlock, no serial chain, and the
//   result is bit-reproducible run to run)
//   D = -d * (prefix + local)/2, written coalesced in the caller's point order.
//
// Arithmetic: an fp64 division costs ~25 issue slots on sm_100a (scripts/ubench_fp64.cu), and
// scipy's formulas written literally take 12 per point, which made the first version of this
// kernel issue-bound at 13 % of the HBM roofline.  Here every interval's two secants
// (do/dtheta and dtheta/do, reciprocals of one another) are formed once, by the branch-free
// reciprocal of fx_math.cuh, and shared with the neighbouring point through a warp shuffle; the
// harmonic mean of the interior slope, 1/((w1/m0 + w2/m1)/(w1+w2)), becomes
// (w1+w2)/(w1*im0 + w2*im1) with im the other direction's secant: 4 reciprocals per point.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fx_math.cuh"

namespace fx {

constexpr int INV_BLOCK = 128;
constexpr int INV_ITEMS = 4;
constexpr int INV_TILE = INV_BLOCK * INV_ITEMS;
constexpr int INV_HALO_L = 2;
constexpr int INV_NAGG = 4;  // inverse integral, forward PCHIP integral, trapezoid, flagged-bad count
constexpr int INV_GROUP = 32; // tiles per look-back group (one warp reads a group)

struct InvArgs {
  long long batch, npts, ntiles;      // tiles per profile
  const double* o;                    // [batch][npts]
  const double* theta;                // [batch][npts]
  double* D;                          // [batch][npts]
  double* sorp;                       // [batch][4]: PCHIP Int theta do, trapezoid Int theta do, extrapolated Int_0^o0, spare
  int* status;                        // [batch]
  double* slope;                      // [batch][npts] do/dtheta at the knots, or nullptr
  double* anti;                       // [batch][npts] Int_i^theta o dtheta at the knots, or nullptr
  double* agg;                        // [batch][ntiles][INV_NAGG] tile aggregates
  double* gtot;                       // [batch][ngroups][INV_NAGG] totals of the full groups of INV_GROUP tiles
  int* gcount;                        // [batch][ngroups] tiles of the group that have published
  long long ngroups;                  // ceil(ntiles / INV_GROUP)
  unsigned long long* ticket;
};

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

// 1/x by the hardware seed and two Newton steps (correctly rounded for ordinary arguments; a
// zero or non-finite step only occurs in a profile that is flagged non-monotone anyway)
__device__ __forceinline__ double rcp_lean(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// interior slope from the two secants m0, m1 and their reciprocals im0, im1 (the secants of the
// inverse function): scipy's weighted harmonic mean with one reciprocal; zero unless the secants
// have the same strict sign
__device__ __forceinline__ double pchip_interior_r(double h0, double h1, double m0, double m1, double im0,
                                                   double im1) {
  double w1 = fma(2.0, h1, h0);
  double w2 = fma(2.0, h0, h1);
  double v = (w1 + w2) * rcp_lean(fma(w1, im0, w2 * im1));
  // same strict sign <=> positive product (secants are O(1e-9 .. 1e9) here: no underflow to zero)
  return (m0 * m1 > 0.0) ? v : 0.0;
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
__device__ __forceinline__ void wait_load4(const double* src, double (&v)[4]) {
  for (;;) {
#pragma unroll
    for (int c = 0; c < 4; c++) v[c] = __ldcg(src + c);
    if (!(is_unset(v[0]) || is_unset(v[1]) || is_unset(v[2]) || is_unset(v[3]))) break;
    __nanosleep(20);
  }
}

__global__ void __launch_bounds__(INV_BLOCK, 8) inverse_scan_kernel(const InvArgs A) {
  constexpr int NS = INV_TILE + INV_HALO_L + 2;  // points held: halo 2 left, 2 right (1 used + pad)
  __shared__ double s_th[NS], s_o[NS];
  __shared__ double s_d[2][INV_TILE + 1];  // do/dtheta at the tile's points; two buffers: a tile's
                                           // output is written one tile later (see below)
  __shared__ double s_df[INV_TILE + 1];    // dtheta/do
  __shared__ double s_warp[2][INV_BLOCK / 32];
  __shared__ double s_scan[INV_ITEMS][INV_BLOCK / 32];
  __shared__ double s_off[INV_ITEMS * (INV_BLOCK / 32) + 1];
  constexpr int NWT = INV_ITEMS * (INV_BLOCK / 32);  // warp totals per tile
  static_assert(NWT <= 32, "the warp totals of a tile are scanned by one warp");
  __shared__ double s_red[INV_NAGG];
  __shared__ unsigned long long s_ticket;
  __shared__ int s_bad;

  const int tid = threadIdx.x;
  const unsigned lane = tid & 31, wid = tid >> 5;
  const long long total_tiles = A.batch * A.ntiles;
  const long long n = A.npts;
  const bool two = (n == 2);

  // The prefix of a tile needs the aggregates of every earlier tile of its profile, which other
  // CTAs are computing at the same moment: waiting for them right away left 45 % of the warp
  // time at that barrier.  So a tile publishes its aggregate, the CTA goes on to load and scan
  // its NEXT tile, and only then resolves the previous tile's prefix and writes its output.
  // Aggregates are published before anything is waited for, so this cannot deadlock.
  long long pv_p = -1, pv_t = 0;  // the deferred tile
  double pv_excl[INV_ITEMS] = {0.0, 0.0, 0.0, 0.0};
  int pv_buf = 0, buf = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_ticket = atomicAdd(A.ticket, 1ull);
      s_bad = 0;
    }
    __syncthreads();
    const long long ticket = (long long)s_ticket;
    const bool have = ticket < total_tiles;
    if (!have && pv_p < 0) break;
    long long p = 0, t = 0;
    double excl[INV_ITEMS] = {0.0, 0.0, 0.0, 0.0};

    if (have) {
      p = ticket / A.ntiles;
      t = ticket - p * A.ntiles;
      const long long r0 = t * INV_TILE;
      const double* th_g = A.theta + p * n;
      const double* o_g = A.o + p * n;
      double* sd = s_d[buf];
      const double dir_step = th_g[n - 2] - th_g[n - 1];  // issued early: needed after the slopes

      // ---- load tile + halo in traversal order (r -> caller index n-1-r) ----
      // s[q] holds traversal point r0 - 2 + q, i.e. memory index jtop - q with jtop = n+1-r0
      {
        const double* th_top = th_g + (n + 1 - r0);
        const double* o_top = o_g + (n + 1 - r0);
        if (r0 >= INV_HALO_L && r0 + INV_TILE + 2 <= n) {  // every point of tile and halo exists
          for (int q = tid; q < NS; q += INV_BLOCK) {
            s_th[q] = th_top[-q];
            s_o[q] = o_top[-q];
          }
        } else {
          const long long qlo = INV_HALO_L - r0, qhi = n + INV_HALO_L - r0;  // valid q in [qlo, qhi)
          for (int q = tid; q < NS; q += INV_BLOCK) {
            const bool ok = q >= qlo && q < qhi;
            s_th[q] = ok ? th_top[-q] : 0.0;
            s_o[q] = ok ? o_top[-q] : 0.0;
          }
        }
      }
      __syncthreads();

      // ---- slopes do/dtheta (inverse interpolant) and dtheta/do (forward) at r in [r0, r0+TILE] --
      // Each lane forms the secants of the interval to the RIGHT of its point; the interval to the
      // left is the previous lane's (one shuffle; lane 0 of a warp recomputes it).
      const long long left = n - r0;                      // points of the profile from r0 on
      const int npt = left > INV_TILE + 1 ? INV_TILE + 1 : (int)left;  // points e in [0, npt) exist
      const bool first_tile = (r0 == 0);
      const bool ends_here = left <= INV_TILE + 1;        // the profile's last point is e = npt-1
      auto slope_at = [&](int e, bool shuffled) {
        const int q = e + INV_HALO_L;
        const bool live = e < npt;
        const bool has_right = live && !(ends_here && e == npt - 1);
        double hx1 = 0.0, ho1 = 0.0, m1 = 0.0, mf1 = 0.0;  // interval (q, q+1)
        if (has_right) {
          hx1 = s_th[q + 1] - s_th[q];
          ho1 = s_o[q + 1] - s_o[q];
          m1 = ho1 * rcp_lean(hx1);
          mf1 = hx1 * rcp_lean(ho1);
        }
        double hx0, ho0, m0, mf0;
        if (shuffled) {
          hx0 = __shfl_up_sync(0xffffffffu, hx1, 1);
          ho0 = __shfl_up_sync(0xffffffffu, ho1, 1);
          m0 = __shfl_up_sync(0xffffffffu, m1, 1);
          mf0 = __shfl_up_sync(0xffffffffu, mf1, 1);
        }
        const bool has_left = live && !(first_tile && e == 0);
        if ((!shuffled || lane == 0) && has_left) {  // interval (q-1, q)
          hx0 = s_th[q] - s_th[q - 1];
          ho0 = s_o[q] - s_o[q - 1];
          m0 = ho0 * rcp_lean(hx0);
          mf0 = hx0 * rcp_lean(ho0);
        }
        if (live) {
          double d, df;
          if (has_left && has_right) {
            d = pchip_interior_r(hx0, hx1, m0, m1, mf0, mf1);
            df = pchip_interior_r(ho0, ho1, mf0, mf1, m0, m1);
          } else if (two) {
            d = has_right ? m1 : m0;
            df = has_right ? mf1 : mf0;
          } else if (has_right) {  // r == 0, one-sided: intervals (q, q+1) and (q+1, q+2)
            double hx2 = s_th[q + 2] - s_th[q + 1], ho2 = s_o[q + 2] - s_o[q + 1];
            d = pchip_edge(hx1, hx2, m1, ho2 / hx2);
            df = pchip_edge(ho1, ho2, mf1, hx2 / ho2);
          } else {  // r == n-1, one-sided, mirrored: intervals (q-1, q) and (q-2, q-1)
            double hxm = s_th[q - 1] - s_th[q - 2], hom = s_o[q - 1] - s_o[q - 2];
            d = pchip_edge(hx0, hxm, m0, hom / hxm);
            df = pchip_edge(ho0, hom, mf0, hxm / hom);
          }
          sd[e] = d;
          s_df[e] = df;
        } else {
          sd[e] = 0.0;
          s_df[e] = 0.0;
        }
      };
#pragma unroll
      for (int it = 0; it < INV_ITEMS; it++) slope_at(it * INV_BLOCK + tid, true);
      if (tid == 0) slope_at(INV_TILE, false);  // the point that closes the tile's last interval
      __syncthreads();

      // ---- per-interval integrals and the block scan (one barrier) -----------------
      double incl[INV_ITEMS], mine[INV_ITEMS];
      double sumF = 0.0, sumT = 0.0;
      int bad = 0;
      const bool up = dir_step > 0.0, flat = !(dir_step != 0.0);  // profile direction, from interval r=0
#pragma unroll
      for (int it = 0; it < INV_ITEMS; it++) {
        int e = it * INV_BLOCK + tid;
        long long r = r0 + e;
        double I = 0.0;
        if (r < n - 1) {
          int q = e + INV_HALO_L;
          double hx = s_th[q + 1] - s_th[q];  // signed step in theta
          double ho = s_o[q + 1] - s_o[q];    // signed step in o
          if (flat || !(up ? hx > 0.0 : hx < 0.0)) bad = 1;
          double trap = ho * (s_th[q] + s_th[q + 1]) * 0.5;
          I = hx * (s_o[q] + s_o[q + 1]) * 0.5 + hx * hx * (sd[e] - sd[e + 1]) * (1.0 / 12.0);
          sumF += trap + ho * ho * (s_df[e] - s_df[e + 1]) * (1.0 / 12.0);
          sumT += trap;
        }
        mine[it] = I;
        incl[it] = warp_incl_scan(I, lane);
        if (lane == 31) s_scan[it][wid] = incl[it];
      }
      sumF = warp_sum(sumF);  // the two sorptivity integrands (fixed tree)
      sumT = warp_sum(sumT);
      if (lane == 0) {
        s_warp[0][wid] = sumF;
        s_warp[1][wid] = sumT;
      }
      if (bad) atomicOr(&s_bad, 1);
      __syncthreads();
      // exclusive offsets of the ITEMS x 8 warp totals (item-major = traversal order): one warp scans them
      if (wid == 0) {
        const double v = ((int)lane < NWT) ? (&s_scan[0][0])[lane] : 0.0;
        const double inc = warp_incl_scan(v, lane);
        if ((int)lane < NWT) s_off[lane] = inc - v;
        if (lane == 31) s_off[NWT] = inc;
      }
      __syncthreads();
      const double carry = s_off[NWT];
#pragma unroll
      for (int it = 0; it < INV_ITEMS; it++) excl[it] = s_off[it * (INV_BLOCK / 32) + wid] + (incl[it] - mine[it]);

      // ---- publish the tile aggregate; close the group if this is its last tile to arrive ------
      if (wid == 0) {
        if (lane == 0) {
          double f = 0.0, z = 0.0;
          for (int w = 0; w < INV_BLOCK / 32; w++) {
            f += s_warp[0][w];
            z += s_warp[1][w];
          }
          double* ag = A.agg + (p * A.ntiles + t) * INV_NAGG;
          __stcg(ag + 0, carry);
          __stcg(ag + 1, f);
          __stcg(ag + 2, z);
          __stcg(ag + 3, (double)s_bad);
          __threadfence();  // ordered before the arrival count below
        }
        const long long g = t / INV_GROUP, g0 = g * INV_GROUP;
        const bool full = g < A.ntiles / INV_GROUP;
        int arrived = 0;
        if (lane == 0 && full) arrived = atomicAdd(A.gcount + p * A.ngroups + g, 1);
        arrived = __shfl_sync(0xffffffffu, arrived, 0);
        if (full && arrived == INV_GROUP - 1) {
          double tot[INV_NAGG];
          wait_load4(A.agg + (p * A.ntiles + g0 + lane) * INV_NAGG, tot);
#pragma unroll
          for (int c = 0; c < INV_NAGG; c++) tot[c] = warp_sum(tot[c]);
          if (lane == 0) {
            double* gt = A.gtot + (p * A.ngroups + g) * INV_NAGG;
#pragma unroll
            for (int c = 0; c < INV_NAGG; c++) __stcg(gt + c, tot[c]);
          }
        }
      }
    }

    // ---- the deferred tile: prefix over the earlier tiles of its profile, then its output --------
    // Two fixed levels, no chain: whoever publishes the last aggregate of a group of INV_GROUP
    // tiles also publishes the group total (above); a tile adds the totals of all earlier groups
    // and the aggregates of its own group's earlier tiles.  O(ntiles/INV_GROUP + INV_GROUP) reads
    // and -- the association being fixed by the tile index alone -- bit-reproducible run to run.
    if (pv_p >= 0) {
      const long long p2 = pv_p, t2 = pv_t;
      if (wid == 0) {
        const long long g = t2 / INV_GROUP, g0 = g * INV_GROUP;
        double part[INV_NAGG] = {0.0, 0.0, 0.0, 0.0};
        for (long long gi = lane; gi < g; gi += 32) {
          double v[INV_NAGG];
          wait_load4(A.gtot + (p2 * A.ngroups + gi) * INV_NAGG, v);
#pragma unroll
          for (int c = 0; c < INV_NAGG; c++) part[c] += v[c];
        }
        const long long k = g0 + lane;
        if (k < t2) {
          double v[INV_NAGG];
          wait_load4(A.agg + (p2 * A.ntiles + k) * INV_NAGG, v);
#pragma unroll
          for (int c = 0; c < INV_NAGG; c++) part[c] += v[c];
        }
#pragma unroll
        for (int c = 0; c < INV_NAGG; c++) part[c] = warp_sum(part[c]);
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < INV_NAGG; c++) s_red[c] = part[c];
        }
      }
      __syncthreads();
      const double prefix = s_red[0];
      const long long r0 = t2 * INV_TILE;
      const double* sd = s_d[pv_buf];

      // D at the tile's points, caller order: traversal point r0 + e is memory index top - e
      const long long top = p2 * n + (n - 1 - r0);
      const long long left = n - r0;
      const int npt = left > INV_TILE ? INV_TILE : (int)left;
      double* D_top = A.D + top;
#pragma unroll
      for (int it = 0; it < INV_ITEMS; it++) {
        const int e = it * INV_BLOCK + tid;
        if (e < npt) {
          const double Iod = prefix + pv_excl[it];
          D_top[-e] = -(sd[e] * Iod) * 0.5;
          if (A.slope) A.slope[top - e] = sd[e];
          if (A.anti) A.anti[top - e] = Iod;
        }
      }

      // the last tile of a profile closes its sorptivity sums and status
      if (t2 == A.ntiles - 1 && tid == 0) {
        const double* th_g = A.theta + p2 * n;
        const double* o_g = A.o + p2 * n;
        const double* own = A.agg + (p2 * A.ntiles + t2) * INV_NAGG;
        double F = s_red[1] + __ldcg(own + 1), Z = s_red[2] + __ldcg(own + 2), nbad = s_red[3] + __ldcg(own + 3);
        // traversal runs from o_last down to o_first: Int_{o_first}^{o_last} theta do = -sum
        double* out = A.sorp + p2 * 4;
        out[0] = -F;
        out[1] = -Z;
        // extrapolation of the first forward cubic from o_first back to 0 (scipy extrapolate=True)
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
          ext = -anti;  // Int_0^{o0}
        }
        out[2] = ext;
        out[3] = 0.0;
        A.status[p2] = (nbad != 0.0) ? 1 : 0;
      }
    }

    if (have) {
      pv_p = p;
      pv_t = t;
#pragma unroll
      for (int it = 0; it < INV_ITEMS; it++) pv_excl[it] = excl[it];
      pv_buf = buf;
      buf ^= 1;
    } else {
      pv_p = -1;
    }
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

inline cudaError_t inverse_launch(long long batch, long long npts, const double* o,
                                  const double* theta, double* D, double* sorp, int* status,
                                  double* slope, double* anti,
                                  int sms, cudaStream_t s, int* launches) {
  InvArgs a;
  a.batch = batch;
  a.npts = npts;
  a.ntiles = (npts + INV_TILE - 1) / INV_TILE;
  a.o = o;
  a.theta = theta;
  a.D = D;
  a.sorp = sorp;
  a.status = status;
  a.slope = slope;
  a.anti = anti;
  size_t tiles = (size_t)batch * (size_t)a.ntiles;
  a.ngroups = (a.ntiles + INV_GROUP - 1) / INV_GROUP;
  size_t groups = (size_t)batch * (size_t)a.ngroups;
  size_t agg_bytes = (tiles + groups) * INV_NAGG * sizeof(double);  // tile aggregates, then group totals
  size_t cnt_bytes = (groups * sizeof(int) + 7) / 8 * 8;            // arrival counts of the groups
  char* scratch = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&scratch, agg_bytes + cnt_bytes + 64, s);
  if (e != cudaSuccess) return e;
  a.agg = reinterpret_cast<double*>(scratch);
  a.gtot = a.agg + tiles * INV_NAGG;
  a.gcount = reinterpret_cast<int*>(scratch + agg_bytes);
  a.ticket = reinterpret_cast<unsigned long long*>(scratch + agg_bytes + cnt_bytes);
  e = cudaMemsetAsync(scratch, 0xff, agg_bytes, s);  // the "unset" pattern (see is_unset)
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(scratch + agg_bytes, 0, cnt_bytes + 64, s);
  if (e != cudaSuccess) return e;
  int bps = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, inverse_scan_kernel, INV_BLOCK, 0);
  if (e != cudaSuccess) return e;
  if (bps < 1) bps = 1;
  long long cap = (long long)sms * bps;
  int grid = (int)((long long)tiles < cap ? (long long)tiles : cap);
  if (grid < 1) grid = 1;
  inverse_scan_kernel<<<grid, INV_BLOCK, 0, s>>>(a);
  *launches = 1;
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return cudaFreeAsync(scratch, s);
}

}  // namespace fx
