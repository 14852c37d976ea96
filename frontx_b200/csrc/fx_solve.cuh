// fx_solve.cuh -- the batched shooting-solve kernel: persistent lanes, a device-wide queue
// of SHOTS, speculative bisection when lanes would otherwise idle.
//
// Replaces, for a batch, the whole of frontx.solve (src/frontx/_forward.py:60-122):
//   bisection on d_dob (optimistix.Bisection, :104-112)
//     > one IVP per shot (diffrax.diffeqsolve, :84-95) with the boolean event (:76-78)
//       > Kvaerno5 ESDIRK step attempts under the I-controller (rtol 1e-3, atol 1e-6, :92)
//         > 6 implicit stages, each a chord iteration on the 2-vector stage derivative
//           > one right-hand side evaluation (src/frontx/_boltzmann.py:27-33)
//
// Design (B200, FP64-pipe / issue bound; no tensor cores, nothing here is a contraction):
//   * The unit of scheduling is ONE SHOT (an initial-value problem for one d_dob), not one
//     problem.  A problem is ~18 sequential shots (102 when the bisection fails), a shot is
//     ~1300 right-hand sides; at 10^5 problems per GPU there is barely more than one problem
//     per resident lane, so problem-granular scheduling leaves a third of the lanes idle in
//     a second round and the whole launch then waits for the 0.5 % of problems that need 5x
//     the shots.  Here the bisection state of every problem lives in global memory (in the
//     caller's own summary/counters rows plus 8 bytes of scratch) and lanes are workers: a
//     lane pops a shot, integrates it, folds the residual into that problem's bisection,
//     pushes the problem's next shot and pops the oldest waiting one (FIFO, so all problems
//     advance at the same rate and finish together).
//   * Speculative bisection: when fewer problems remain than lanes, the next 2..5 levels of
//     a problem's bisection tree (3..31 midpoints, each a pure function of the bracket) are
//     pushed as independent shots, sized so that the batches of all unfinished problems just
//     fill the lanes; whoever finishes the last shot of a batch walks the tree with the real
//     residuals, consuming up to 5 bisection steps per round.  Results, counters and the
//     accepted d_dob are identical to sequential bisection (only shots on the taken path are
//     counted); the wasted shots run on lanes that had nothing else to do.  This is what
//     shortens the critical path of the failing problems from ~55 sequential shots to ~25.
//   * Inside a shot the four nested loops (step attempts > stages > chord iterations > RHS)
//     are flattened into ONE loop whose trip is exactly one right-hand-side evaluation (D, D',
//     1/D -- and D'' on the trip that also builds the chord Jacobian), so lanes in different
//     stages, steps, shots or problems execute the expensive model evaluation convergently;
//     everything else is a short predicated block after it.
//   * Parking: a lane whose six stages converged waits up to WAIT_TRIPS trips for WAIT_QUORUM
//     lanes of its warp, so that the ~200-instruction attempt-end block runs for groups.
//   * Registers bound the resident warps: the 7 stage derivatives (the stage index is a per-lane
//     runtime value), the model constants and every per-lane value that is not needed on each
//     trip live in shared memory, one column per thread (conflict-free), beside the Butcher
//     tableau, the predictor rows and the log/exp lookup tables: 96 registers, 5 CTAs per SM.
//   * Kernels are compiled per model (template parameter); mixed batches are bucketed by
//     model on the device (fx_api.cu) and each bucket has its own queue.
//   * Output: 64-byte summary + 32-byte counter record per problem, and one 32-byte knot
//     (o, theta, theta', theta'') per accepted step of the last shot.
//   * The arithmetic itself is in fx_arith.cuh / fx_models.cuh / fx_math.cuh (host+device,
//     explicit fma, built with -fmad=false) so that the CPU twin used by the tests reproduces
//     every decision bit for bit, whatever order the queue ran the shots in.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fx_arith.cuh"
#include "fx_models.cuh"

namespace fx {

// Scheduling knobs (results never depend on them; scripts/exp_variants.py builds and times variants):
#ifndef FX_SOLVE_BLOCK
#define FX_SOLVE_BLOCK 128
#endif
#ifndef FX_SOLVE_CTAS
#define FX_SOLVE_CTAS 5      // <= 96 registers per thread: 5 warps per SM sub-partition
#endif
#ifndef FX_SPEC_MAX_DEPTH
#define FX_SPEC_MAX_DEPTH 5  // up to 2^depth - 1 speculative shots per problem and round
#endif
#ifndef FX_SPARSE_NAP
#define FX_SPARSE_NAP 2000   // ns between the polls of an idle warp in the end phase
#endif
#ifndef FX_SPARSE_POLL_MASK
#define FX_SPARSE_POLL_MASK 63u // a busy warp with free lanes polls every (mask + 1)-th trip in the end phase
#endif
constexpr int SOLVE_BLOCK = FX_SOLVE_BLOCK;
constexpr int SOLVE_CTAS_PER_SM = FX_SOLVE_CTAS;
constexpr int SPEC_MAX_DEPTH = FX_SPEC_MAX_DEPTH;
static_assert(SPEC_MAX_DEPTH >= 1 && SPEC_MAX_DEPTH <= 7, "a ring entry carries the tree node in 8 bits");
constexpr int SPEC_NODES = 1 << SPEC_MAX_DEPTH;  // entries per speculation slot (entry 0 = header)
constexpr int REPLAY_NODE = 255;                 // ring code of "integrate the accepted shot again, for its knots"
constexpr int FX_IN_FLIGHT = -3;                 // counters[.][0] while a problem is unfinished: if a launch ever
                                                 // ended early (ctl->abort), such problems do not read as solved

__constant__ Tableau c_tab = FX_TABLEAU_INIT;

// queue control block of one model bucket (set by init_queues_kernel)
struct QueueCtl {
  unsigned long long head;  // next ticket to hand out
  unsigned long long tail;  // next ticket to reserve for a push (tickets >= n_work live in the ring)
  long long avail;          // pushed and not yet claimed
  long long remaining;      // unfinished problems
  unsigned int abort;       // internal inconsistency (a ring entry never arrived): everyone leaves
  unsigned int pad_[7];
};

// result of one shot: what the bisection and the counters need from it
struct ShotResult {
  double res, t_end, y_end;
  int att, rej, rhs, knots;
};

// one node of a speculation slot (48 bytes, 16-byte aligned parts).  Entry 0 of a slot is
// its header: k.x = shots of the batch still running, k.y = depth of the batch.
struct SpecEntry {
  int4 k;  // att, rej, rhs, knots
  double res, t_end, y_end, pad_;
};

struct SolveArgs {
  const unsigned long long* counts;  // per-model bucket sizes (device)
  int slot;                  // this launch's bucket
  int ring_region;           // which of the ring's per-bucket regions it owns (buckets absent from the batch have none)
  int full_grid;             // 1: the buckets of this batch run one after another, each on every CTA of its grid
  const int* perm;           // bucket-major list of problem indices
  const double* params;      // [N][12]
  const double* b;           // [N]  boundary value; with a flow-rate boundary: lower end of the bracket on b
  const double* i;           // [N]
  const double* q;           // [N]  flow-rate boundary only: inlet flux density Qb/(angle ob height)
  const double* b_hi;        // [N]  flow-rate boundary only: upper end of the bracket on b
  double* summary;           // [N][8]; rows double as bisection state while a problem is in flight
  int* counters;             // [N][8]; likewise
  double* knots;             // [N][k_max][4] or nullptr
  int k_max;
  QueueCtl* ctl;             // [FX_NUM_MODELS + 1]
  unsigned long long* ring;  // task ring shared by the buckets (each owns a region)
  long long ring_pad;        // spare entries per bucket beyond its problem count
  int2* bstate;              // [N]: .x = phase | lo_neg << 2 | n_bis << 3, .y = speculation slot + 1
  SpecEntry* spec;           // [spec_slots][SPEC_NODES]
  unsigned int* spec_next;   // bump allocator of speculation slots (shared by the buckets)
  unsigned int spec_slots;
  double itol, rtol, atol, ob;
  int max_bisect, max_ode_steps, max_expansions, radial_k;
};

enum Mode : int { M_FETCH = 0, M_DB = 1, M_INIT0 = 2, M_INIT1 = 3, M_CHORD = 4, M_DONE = 5, M_WAIT = 6 };

// A lane whose six stages have converged parks (M_WAIT) until WAIT_QUORUM lanes of its warp are
// parked or it has waited WAIT_TRIPS trips: the error estimate / controller block (~200
// instructions, needed by one lane in 22 per trip) then runs once for the group instead of on
// four trips out of five for one or two lanes, and the group starts its next attempt -- and
// builds its chord Jacobians -- together.  Measured on C2: -15 % warp instructions per trip for
// 5 % idle lane-trips.  The arithmetic is untouched, so results do not depend on the grouping.
constexpr int WAIT_QUORUM = 6;
constexpr int WAIT_TRIPS = 4;
constexpr int WAIT_MIN_WORKING = 16;

__device__ __forceinline__ unsigned long long ring_entry(int prob, int node, unsigned long long lap) {
  return ((unsigned long long)((lap + 1) & 0xffffffull) << 40) | ((unsigned long long)(node & 0xff) << 32) |
         (unsigned long long)(unsigned int)prob;
}

// push `count` shots of problem `prob` (nodes node0 .. node0+count-1) onto this bucket's ring.
// The caller has made the problem's state visible (__threadfence) before calling.
__device__ __forceinline__ void push_tasks(QueueCtl* ctl, unsigned long long* ring, long long n_work,
                                           long long cap, int prob, int node0, int count) {
  unsigned long long t = atomicAdd(&ctl->tail, (unsigned long long)count);
  for (int k = 0; k < count; k++) {
    unsigned long long r = t + k - (unsigned long long)n_work;
    unsigned long long lap = r / (unsigned long long)cap, idx = r - lap * (unsigned long long)cap;
    __stcg(&ring[idx], ring_entry(prob, node0 + k, lap));
  }
  __threadfence();
  atomicAdd((unsigned long long*)&ctl->avail, (unsigned long long)count);
}

// midpoint of speculation node `node` (heap index, root 1) of the bracket [lower, upper]:
// the same fma the sequential bisection evaluates after taking that path
__device__ __forceinline__ double node_midpoint(double lower, double upper, int node) {
  int depth = 31 - __clz(node);
  double lo = lower, hi = upper;
  double mid = fma(0.5, hi - lo, lo);
  for (int l = depth - 1; l >= 0; l--) {
    if ((node >> l) & 1) lo = mid; else hi = mid;
    mid = fma(0.5, hi - lo, lo);
  }
  return mid;
}

// derived constants actually used by each model (rows of s_dpar)
template <int MODEL>
struct DerivedCount {
  static constexpr int value = MODEL == FX_CONSTANT ? 2 : MODEL == FX_POWER_LAW ? 4 : MODEL == FX_LOG ? 2 :
                               MODEL == FX_TABULATED ? 2 :
                               MODEL == FX_BROOKS_COREY ? 6 : MODEL == FX_VAN_GENUCHTEN ? 9 :
                               MODEL == FX_LETD ? 8 : 11;
};

struct BlockCtx {
  long long n_work, ring_cap, n_lanes;
  const int* perm;
  QueueCtl* ctl;
  unsigned long long* ring;
};

// GEOM: 0 planar (the reference's frontx.solve), 1 radial term k/o with the boundary at ob > 0,
// 2 radial with a flow-rate boundary (north-star C4, fronts.solve_flowrate semantics): the
// bisection runs on the boundary VALUE b, every shot starting from (b, d_dob = -q/D(b)).
template <int MODEL, int GEOM>
__global__ void __launch_bounds__(SOLVE_BLOCK, SOLVE_CTAS_PER_SM) shoot_bisect_kernel(const SolveArgs A) {
  constexpr int NDP = DerivedCount<MODEL>::value;
  constexpr bool RADIAL = GEOM >= 1;
  constexpr bool FLOW = GEOM == 2;
  __shared__ Tableau tab;
  __shared__ MathTab mt;                     // log / exp lookup tables (lanes index them divergently)
  __shared__ __align__(16) double fsm[7][SOLVE_BLOCK][2];  // stage derivatives: one 16-byte pair per stage and thread
  // Per-lane state that is not needed on every trip lives in shared memory (one column per
  // thread, conflict-free) so that the registers hold only the hot integrator state:
  // registers, not arithmetic, bound the resident warps of this kernel.
  __shared__ double s_dpar[NDP][SOLVE_BLOCK];  // the model's derived constants
  __shared__ double s_cold[FLOW ? 6 : 5][SOLVE_BLOCK];  // b, i, direction, bisection value of the shot, t1, (q)
  __shared__ int s_cnt[4][SOLVE_BLOCK];        // attempts, rejections, knots of the shot; RHS count where a frozen attempt ended (-1: none)
  __shared__ BlockCtx ctx;                     // bucket geometry (same for every thread)

  const int tid = threadIdx.x;
  {
    const double* src = reinterpret_cast<const double*>(&c_tab);
    double* dst = reinterpret_cast<double*>(&tab);
    for (int k = tid; k < (int)(sizeof(Tableau) / sizeof(double)); k += SOLVE_BLOCK) dst[k] = src[k];
    src = reinterpret_cast<const double*>(&g_math_tab_dev);
    dst = reinterpret_cast<double*>(&mt);
    for (int k = tid; k < (int)(sizeof(MathTab) / sizeof(double)); k += SOLVE_BLOCK) dst[k] = src[k];
  }
#pragma unroll
  for (int k = 0; k < NDP; k++) s_dpar[k][tid] = 0.0;  // every lane evaluates the model on every trip, work or not
  __syncthreads();

  const unsigned lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  const Tol tol = {A.rtol, A.atol};
  const int radial_k = RADIAL ? A.radial_k : 0;
  auto fs = [&](int j, int c) -> double { return fsm[j][tid][c]; };

  // bucket geometry: model `slot` owns perm[off .. off+n_work) and a ring region of its own
  // A mixed batch runs one persistent grid per model side by side: each keeps a share of its
  // CTAs proportional to its bucket (the rest exit at once), so that together they fill the SMs.
  if (tid == 0) {
    long long off = 0, total = 0;
    for (int k = 0; k < A.slot; k++) off += (long long)A.counts[k];
    for (int k = 0; k <= FX_NUM_MODELS; k++) total += (long long)A.counts[k];
    ctx.n_work = (long long)A.counts[A.slot];
    long long share = total > 0 ? (ctx.n_work * (long long)gridDim.x + total - 1) / total : 0;
    if (ctx.n_work > 0 && share < 1) share = 1;
    // Side by side, the shares follow the buckets' problem COUNTS, not their work (Brooks-Corey takes 1.6x the
    // right-hand sides of LETd in config C3), so one bucket finishes first and its SMs idle: fine for a batch
    // that is one latency chain per problem anyway, wasteful for a large one.  Large mixed batches therefore
    // run their buckets one after another (fx_api.cu), each on the whole grid.
    if (A.full_grid) share = gridDim.x;
    if ((long long)blockIdx.x >= share) ctx.n_work = 0;  // this CTA is not needed
    ctx.perm = A.perm + off;
    ctx.ctl = A.ctl + A.slot;
    ctx.ring = A.ring + off + (long long)A.ring_region * A.ring_pad;
    ctx.ring_cap = ctx.n_work + A.ring_pad;
    ctx.n_lanes = share * SOLVE_BLOCK;
  }
  __syncthreads();
  if (ctx.n_work == 0) return;
  // Speculative shots do not write knots; when the bisection ends on one, the accepted shot is
  // integrated once more (node = -1) to write the dense output -- one shot more than the ~18 of a solve.
  const bool can_speculate = A.spec_slots > 0;

  // ---- per-lane state (registers) ---------------------------------------
  int mode = M_FETCH;
  int prob = 0, node = 0;                  // the shot this lane is integrating
  double t0 = 0, h = 0, y00 = 0, y01 = 0;  // step [t0,t1], h = t1 - t0, state at t0
  double yp0 = 0, yp1 = 0;                 // y_partial of the current stage
  double f0 = 0, f1 = 0;                   // chord iterate (stage derivative)
  Chord M = {0, 0, 0, 0};                  // inverse of gamma*h*J - I
  double dsz = 0;                          // square of the previous chord step size
  double te = 0, ye0 = 0, ye1 = 0;         // where the next RHS is evaluated
  int stage = 1, iter = 0, rhs_shot = 0;
  unsigned poll = 0, idle_ns = 1000u;
  bool sparse = false;                     // few problems left: spread shots over warps
  // ---- per-lane state (shared memory columns) ----------------------------
  double& bnd_b = s_cold[0][tid];
  double& bnd_i = s_cold[1][tid];
  double& direction = s_cold[2][tid];
  double& d_shot = s_cold[3][tid];
  double& t1 = s_cold[4][tid];
  int& att_shot = s_cnt[0][tid];
  int& rej_shot = s_cnt[1][tid];
  int& n_knots = s_cnt[2][tid];
  int& rhs_mark = s_cnt[3][tid];

  for (;;) {
    // ---- fetch: pop shots for the lanes that have none -------------------------------------
    const unsigned need = __ballot_sync(FULL, mode == M_FETCH);
    const bool busy = __any_sync(FULL, mode != M_FETCH && mode != M_DONE);
    // A warp that is already integrating takes new shots into its free lanes at no issue cost,
    // so it polls often; a fully idle warp backs off (1 -> 16 us), which makes late, sparse
    // work settle into few full warps instead of many nearly empty ones.
    // In the last rounds (few problems left, each a critical path) the opposite holds: a shot
    // runs fastest in a warp of its own, so busy warps stop taking work and idle ones poll fast.
    if (need && (!busy || (poll++ & (sparse ? FX_SPARSE_POLL_MASK : 1u)) == 0)) {
      // The probe is three loads of warp-uniform addresses (one transaction each) and two votes: an idle warp's
      // poll must cost next to nothing, because in the end phase the polls of the idle warps share the issue
      // slots with the few shots that are the launch's critical path.  Only when something is waiting does a
      // leader claim it.
      QueueCtl* ctl = ctx.ctl;
      const long long rem = *(volatile long long*)&ctl->remaining;
      const unsigned int ab = *(volatile unsigned int*)&ctl->abort;
      const long long av = *(volatile long long*)&ctl->avail;
      const bool over = __any_sync(FULL, rem <= 0 || ab != 0u);
      const bool some = !over && __any_sync(FULL, av > 0);
      sparse = __any_sync(FULL, rem * 16 < ctx.n_lanes);
      long long got = over ? -1 : 0;
      unsigned long long h0 = 0;
      if (some) {
        const int leader = __ffs(need) - 1;
        const int want = __popc(need);
        if ((int)lane == leader) {
          long long old = (long long)atomicAdd((unsigned long long*)&ctl->avail,
                                               (unsigned long long)(-(long long)want));
          got = old >= want ? want : (old > 0 ? old : 0);
          if (got < want) atomicAdd((unsigned long long*)&ctl->avail, (unsigned long long)(want - got));
          if (got > 0) h0 = atomicAdd(&ctl->head, (unsigned long long)got);
        }
        got = __shfl_sync(FULL, got, leader);
        h0 = __shfl_sync(FULL, h0, leader);
      }
      if (mode == M_FETCH) {
        const int rank = __popc(need & ((1u << lane) - 1u));
        if (got < 0) {
          mode = M_DONE;
        } else if (rank < got) {
          const unsigned long long ticket = h0 + rank;
          const bool fresh = ticket < (unsigned long long)ctx.n_work;
          bool lost = false;
          if (fresh) {
            prob = ctx.perm[ticket];
            node = 0;
          } else {
            const unsigned long long* ring = ctx.ring;
            const unsigned long long ring_cap = (unsigned long long)ctx.ring_cap;
            unsigned long long r = ticket - (unsigned long long)ctx.n_work;
            unsigned long long lap = r / ring_cap, idx = r - lap * ring_cap;
            const unsigned long long seq = (lap + 1) & 0xffffffull;
            unsigned long long e = __ldcg(&ring[idx]);
            unsigned spins = 0;
            while ((e >> 40) != seq) {
              __nanosleep(64);
              e = __ldcg(&ring[idx]);
              if (++spins > (1u << 22)) {  // the matching push never arrived: give up loudly
                atomicExch(&ctx.ctl->abort, 1u);
                lost = true;
                break;
              }
            }
            __threadfence();
            prob = (int)(unsigned int)(e & 0xffffffffull);
            node = (int)((e >> 32) & 0xff);
            if (node == REPLAY_NODE) node = -1;
          }
          if (lost) {
            mode = M_DONE;
          } else {
            {
              double dpar[NDERIVED];
              derive_params<MODEL>(A.params + (size_t)prob * FX_NPARAM, dpar, mt);
#pragma unroll
              for (int k = 0; k < NDP; k++) s_dpar[k][tid] = dpar[k];
            }
            bnd_b = A.b[prob];
            bnd_i = A.i[prob];
            direction = (bnd_i > bnd_b) ? 1.0 : ((bnd_i < bnd_b) ? -1.0 : 0.0);
            if (FLOW && fresh) {
              // bracket [b_lo, b_hi] given by the caller; first shot at its lower end
              double* S = A.summary + (size_t)prob * FX_NSUM;
              __stcg(S + 6, bnd_b);
              __stcg(S + 7, A.b_hi[prob]);
              int4* c = reinterpret_cast<int4*>(A.counters + (size_t)prob * FX_NCNT);
              __stcg(c, make_int4(FX_IN_FLIGHT, 0, 0, 0));
              __stcg(c + 1, make_int4(0, 0, 0, 0));
              __stcg(&A.bstate[prob], make_int2(0, 0));
              d_shot = bnd_b;
            }
            if (fresh && !FLOW) {
              mode = M_DB;  // D(b) first: the bracket needs it
              te = A.ob;
              ye0 = bnd_b;
              ye1 = 0.0;
            } else if (FLOW) {
              if (!fresh) {
                const double* S = A.summary + (size_t)prob * FX_NSUM;
                const double lower = __ldcg(S + 6), upper = __ldcg(S + 7);
                const int phase = __ldcg(&A.bstate[prob].x) & 3;
                d_shot = (node < 0) ? __ldcg(S + 0)
                         : (phase == 0) ? lower : (phase == 1) ? upper : node_midpoint(lower, upper, node ? node : 1);
              }
              s_cold[FLOW ? 5 : 0][tid] = A.q[prob];
              mode = M_DB;  // D(b) of this shot's b: d_dob = -q/D(b)
              te = A.ob;
              ye0 = d_shot;
              ye1 = 0.0;
            } else {
              const double* S = A.summary + (size_t)prob * FX_NSUM;
              const double lower = __ldcg(S + 6), upper = __ldcg(S + 7);
              const int phase = __ldcg(&A.bstate[prob].x) & 3;
              d_shot = (node < 0) ? __ldcg(S + 0)
                       : (phase == 0) ? lower : (phase == 1) ? upper : node_midpoint(lower, upper, node ? node : 1);
              t0 = A.ob;
              y00 = bnd_b;
              y01 = d_shot;
              att_shot = rej_shot = rhs_shot = 0;
              rhs_mark = -1;
              te = t0;
              ye0 = y00;
              ye1 = y01;
              mode = M_INIT0;
            }
          }
        }
      }
      if (!busy) {
        if (got == 0) {
          __nanosleep(sparse ? (unsigned)FX_SPARSE_NAP : idle_ns);
          if (idle_ns < 16000u) idle_ns *= 2;
        } else {
          idle_ns = 1000u;
        }
      }
    }
    if (!__any_sync(FULL, mode != M_DONE)) break;
    if (!__any_sync(FULL, mode != M_FETCH && mode != M_DONE)) continue;

    // ---- the one expensive thing per trip: a right-hand-side evaluation ---
    const bool wantJ = (mode == M_CHORD) && (stage == 1) && (iter == 0);
    double D, D1, D2, iD;
    {
      double dpar[NDERIVED];
#pragma unroll
      for (int k = 0; k < NDERIVED; k++) dpar[k] = (k < NDP) ? s_dpar[k < NDP ? k : 0][tid] : 0.0;
      eval_D<MODEL>(dpar, ye0, wantJ, D, D1, D2, iD, mt);
    }
    const Rhs R = rhs_eval(D1, iD, te, ye1, radial_k);

    bool att_end = false, att_ok = false, shot_end = false, occurred = false;
    bool prep = false;

    if (FLOW && mode == M_DB) {
      // flow-rate boundary: this shot starts from (b, -q/D(b)) with b the bisection value
      bnd_b = d_shot;
      direction = (bnd_i > bnd_b) ? 1.0 : ((bnd_i < bnd_b) ? -1.0 : 0.0);
      t0 = A.ob;
      y00 = bnd_b;
      y01 = -s_cold[FLOW ? 5 : 0][tid] * iD;
      att_shot = rej_shot = rhs_shot = 0;
      rhs_mark = -1;
      te = t0;
      ye0 = y00;
      ye1 = y01;
      mode = M_INIT0;
    } else if (mode == M_DB) {
      // fresh problem: bracket [0, (i-b)/(2 sqrt(D(b)))], first shot at its lower end
      Bisect bis;
      bis.start(D, bnd_b, bnd_i);
      double* S = A.summary + (size_t)prob * FX_NSUM;
      __stcg(S + 4, D);
      __stcg(S + 6, bis.lower);
      __stcg(S + 7, bis.upper);
      int4* c = reinterpret_cast<int4*>(A.counters + (size_t)prob * FX_NCNT);
      __stcg(c, make_int4(FX_IN_FLIGHT, 0, 0, 0));
      __stcg(c + 1, make_int4(0, 0, 0, 0));
      __stcg(&A.bstate[prob], make_int2(0, 0));
      d_shot = bis.lower;
      t0 = A.ob;
      y00 = bnd_b;
      y01 = d_shot;
      att_shot = rej_shot = rhs_shot = 0;
      rhs_mark = -1;
      te = t0;
      ye0 = y00;
      ye1 = y01;
      mode = M_INIT0;
    } else if (mode == M_INIT0) {
      // solver.init + first half of the Hairer initial-step heuristic
      rhs_shot++;
      fsm[0][tid][0] = R.F0;
      fsm[0][tid][1] = R.F1;
      if (A.knots && node <= 0 && A.k_max > 0)
        *reinterpret_cast<double4*>(A.knots + (size_t)prob * A.k_max * FX_KNOT) = make_double4(t0, y00, y01, R.F1);
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
      rhs_shot++;
      double dt = init_dt(y00, y01, fs(0, 0), fs(0, 1), R.F0, R.F1, h, dsz, tol, mt);
      t1 = t0 + dt;
      h = t1 - t0;
      stage = 1;
      prep = true;
      mode = M_CHORD;
    } else if (mode == M_CHORD) {
      rhs_shot++;
      if (wantJ) M = chord_matrix(D, D1, D2, R, te, ye1, h, radial_k);
      double size = chord_update(M, R.F0, R.F1, f0, f1, tol);
      iter++;
      bool done, ok;
      chord_decide(iter, size, dsz, done, ok);
      dsz = size;
      if (done) {
        if (ok) {
          fsm[stage][tid][0] = f0;
          fsm[stage][tid][1] = f1;
          if (stage < 6) { stage++; prep = true; }
          else { mode = M_WAIT; iter = 0; }  // iter counts the trips parked
        } else {
          att_end = true;
        }
      } else {
        ye0 = stage_point(yp0, h, f0);
        ye1 = stage_point(yp1, h, f1);
      }
    }

    // ---- release the parked lanes when enough of them have gathered ---------
    {
      const unsigned parked = __ballot_sync(FULL, mode == M_WAIT);
      if (parked) {
        const bool overdue = __any_sync(FULL, mode == M_WAIT && iter >= WAIT_TRIPS);
        // parking pays only in a full warp (throughput-bound); in a sparse one it is pure latency
        const int working = __popc(__ballot_sync(FULL, mode == M_CHORD || mode == M_INIT0 || mode == M_INIT1));
        if (mode == M_WAIT) {
          if (__popc(parked) >= WAIT_QUORUM || overdue || working < WAIT_MIN_WORKING) {
            att_end = true;
            att_ok = true;
            mode = M_CHORD;
          } else {
            iter++;
          }
        }
      }
    }

    // ---- end of a step attempt: error estimate, controller, event ---------
    if (att_end) {
      att_shot++;
      StepEnd s = step_finish(tab, att_ok, fs, y00, y01, yp0, yp1, f0, f1, h, tol, mt);
      // A FROZEN integrator: if this attempt leaves (t0, y0, f(t0, y0), h) bit for bit as it found them, the next
      // attempt -- a pure function of that state -- repeats it exactly, and so does every one after it, until
      // max_ode_steps ends the shot without an event.  That is how a too-steep shot dies against a sharp front:
      // within a few hundred attempts either the step underflows (t0 + dt == t0, h = 0: every further attempt is
      // "accepted" and moves nothing) or it settles at one ulp of t0 with the error test rejecting it and the
      // controller proposing 0.85 ulp, which rounds back to one ulp -- and the remaining ~3700 of diffrax's 4096
      // attempts are copies.  They are counted, not integrated: same residual, same counters, same knots (the CPU
      // twin runs them all; the tests demand bit equality).
      // The test must cost nothing where it is false -- it runs once per attempt, in a block few lanes of a warp
      // execute together: a frozen step is zero or one ulp of t0, which the exponents of h and t0 tell (two integer
      // instructions); only then is the state compared.  The copies' right-hand sides are those of one attempt:
      // the first frozen attempt marks where the next one -- a copy -- starts, the second knows the count.
      bool frozen = false;
      int rhs_att = 0;
#ifndef FX_FROZEN_FF
#define FX_FROZEN_FF 1
#endif
      if (FX_FROZEN_FF && (__double2hiint(t0) - __double2hiint(h)) >= (51 << 20)) {
        const double nt0 = s.keep ? t1 : t0;
        // a rejected attempt changes nothing but the step; an accepted one must reproduce t0, y0 and f(t0, y0)
        const bool same = (((nt0 + s.dt_next) - nt0) == h) &&
                          (!s.keep || (t1 == t0 && s.y10 == y00 && s.y11 == y01 && f0 == fsm[0][tid][0] && f1 == fsm[0][tid][1]));
        if (same && rhs_mark >= 0) {
          frozen = true;
          rhs_att = rhs_shot - rhs_mark;
        }
        rhs_mark = same ? rhs_shot : -1;
      }
      if (s.keep) {
        t0 = t1;
        y00 = s.y10;
        y01 = s.y11;
        fsm[0][tid][0] = f0;  // FSAL
        fsm[0][tid][1] = f1;
        if (A.knots && node <= 0 && n_knots < A.k_max) {
          double* knot_row = A.knots + (size_t)prob * A.k_max * FX_KNOT;
          *reinterpret_cast<double4*>(knot_row + (size_t)n_knots * FX_KNOT) = make_double4(t0, y00, y01, f1);
        }
        n_knots++;
        if (event_fired(direction, y00, y01, bnd_i, A.itol)) {
          shot_end = true;
          occurred = true;
        }
      } else {
        rej_shot++;
      }
      if (frozen && !shot_end && att_shot < A.max_ode_steps) {
        const int left = A.max_ode_steps - att_shot;  // the copies
        att_shot += left;
        rhs_shot += left * rhs_att;
        if (s.keep) {
          if (A.knots && node <= 0) {
            double* knot_row = A.knots + (size_t)prob * A.k_max * FX_KNOT;
            for (int k = n_knots; k < A.k_max && k < n_knots + left; k++)
              *reinterpret_cast<double4*>(knot_row + (size_t)k * FX_KNOT) = make_double4(t0, y00, y01, f1);
          }
          n_knots += left;
        } else {
          rej_shot += left;
        }
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

    // ---- end of a shot: fold it into the problem's bisection ----------------
    if (shot_end && node < 0) {
      // the accepted shot, integrated again for its dense output: the problem is done
      __threadfence();
      atomicAdd((unsigned long long*)&ctx.ctl->remaining, (unsigned long long)(-1LL));
      mode = M_FETCH;
    } else if (shot_end) {
      const double res = occurred ? (y00 - bnd_i) : direction * fx_inf();  // _forward.py:97-101
      double* S = A.summary + (size_t)prob * FX_NSUM;
      int4* C = reinterpret_cast<int4*>(A.counters + (size_t)prob * FX_NCNT);
      int2 bs = __ldcg(&A.bstate[prob]);
      SpecEntry* slot = (node > 0) ? A.spec + (size_t)(bs.y - 1) * SPEC_NODES : nullptr;
      bool walker = true;
      int depth = 1;
      if (node > 0) {
        // one of a speculative batch: record the result; the lane that completes the batch walks it
        SpecEntry* e = slot + node;
        __stcg(&e->k, make_int4(att_shot, rej_shot, rhs_shot, n_knots));
        __stcg(&e->res, res);
        __stcg(&e->t_end, t0);
        __stcg(&e->y_end, y00);
        depth = __ldcg(&slot->k.y);
        __threadfence();
        walker = atomicSub(&slot->k.x, 1) == 1;
        if (walker) __threadfence();
      }
      if (walker) {
        Bisect bis;
        bis.lower = __ldcg(S + 6);
        bis.upper = __ldcg(S + 7);
        const double D_b = __ldcg(S + 4);
        const int4 c0 = __ldcg(C), c1 = __ldcg(C + 1);
        bis.phase = bs.x & 3;
        bis.lo_neg = (bs.x >> 2) & 1;
        bis.n_bis = bs.x >> 3;
        bis.n_exp = c0.z;
        int n_shots = c0.y, n_att = c0.w, n_rej = c1.x, n_rhs = c1.y;
        int cur = node > 0 ? 1 : 0;  // tree node being consumed (0: the ordinary single shot)
        bis.d_cur = (node > 0) ? node_midpoint(bis.lower, bis.upper, 1) : d_shot;
        ShotResult e = {res, t0, y00, att_shot, rej_shot, rhs_shot, n_knots};
        int result = -1;
        double d_last = d_shot;
        for (int level = 0;; level++) {
          if (node > 0) {
            if (cur != node) {
              const SpecEntry* q = slot + cur;
              const int4 k = __ldcg(&q->k);
              e.res = __ldcg(&q->res);
              e.t_end = __ldcg(&q->t_end);
              e.y_end = __ldcg(&q->y_end);
              e.att = k.x; e.rej = k.y; e.rhs = k.z; e.knots = k.w;
            } else {
              e = ShotResult{res, t0, y00, att_shot, rej_shot, rhs_shot, n_knots};
            }
          }
          d_last = bis.d_cur;
          const bool went_right = ((e.res < 0.0) == (bool)bis.lo_neg);
          n_shots++;
          n_att += e.att;
          n_rej += e.rej;
          n_rhs += e.rhs;
          result = bis.next(e.res, A.itol, A.max_bisect, A.max_expansions);
          // A shot is a pure function of d_dob.  Once the bracket has collapsed to two adjacent
          // doubles without |res| < itol, the midpoint is one of its ends and every remaining
          // bisection step repeats the shot just made: replay its residual and counters.
          while (result < 0 && bis.d_cur == d_last) {
            n_shots++;
            n_att += e.att;
            n_rej += e.rej;
            n_rhs += e.rhs;
            result = bis.next(e.res, A.itol, A.max_bisect, A.max_expansions);
          }
          if (result >= 0 || level + 1 >= depth) break;
          cur = 2 * cur + (went_right ? 1 : 0);
        }
        if (result >= 0) {
          double2* s2 = reinterpret_cast<double2*>(S);
          __stcg(s2 + 0, make_double2(d_last, e.t_end));
          __stcg(s2 + 1, make_double2(e.y_end, -2.0 * D_b * d_last));
#ifdef FX_DEBUG_TIMES
          {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            e.res = (double)now;  // developer experiment: finish time in the residual slot
          }
#endif
          __stcg(s2 + 2, make_double2(D_b, e.res));
          __stcg(s2 + 3, make_double2(bis.lower, bis.upper));
          __stcg(C, make_int4(result, n_shots, bis.n_exp, n_att));
          __stcg(C + 1, make_int4(n_rej, n_rhs, n_att, e.knots));
          if (A.knots && node > 0) {
            // the knots in memory are those of an earlier ordinary shot: run the accepted one again
            __threadfence();
            push_tasks(ctx.ctl, ctx.ring, ctx.n_work, ctx.ring_cap, prob, REPLAY_NODE, 1);
          } else {
            atomicAdd((unsigned long long*)&ctx.ctl->remaining, (unsigned long long)(-1LL));
          }
        } else {
          // not converged: save the bisection state and queue the problem's next shot(s)
          __stcg(S + 6, bis.lower);
          __stcg(S + 7, bis.upper);
          __stcg(C, make_int4(FX_IN_FLIGHT, n_shots, bis.n_exp, n_att));
          __stcg(C + 1, make_int4(n_rej, n_rhs, n_att, e.knots));
          // Speculation depth: batches of 2^d - 1 shots, sized so that the batches of all
          // unfinished problems together fill the resident lanes (rho = lanes per problem);
          // the fractional level is rounded per problem and shot by a hash, which keeps the
          // choice a pure function of (problem, shot, problems remaining).
          int d = 1;
          if (can_speculate && bis.phase == 2) {
            long long rem = *(volatile long long*)&ctx.ctl->remaining;
            if (rem < 1) rem = 1;
            const long long rho = ctx.n_lanes * 256 / rem;  // lanes per unfinished problem, in 1/256
            while (d < SPEC_MAX_DEPTH && ((2LL << d) - 1) * 256 <= rho) d++;
            if (d < SPEC_MAX_DEPTH) {
              const long long lo = ((1LL << d) - 1) * 256, hi = ((2LL << d) - 1) * 256;
              const unsigned u = (((unsigned)prob * 2654435761u) ^ ((unsigned)n_shots * 40503u)) >> 24;
              if ((rho - lo) * 256 > (hi - lo) * (long long)u) d++;
            }
            // Ring entries pushed and not yet popped must never exceed its capacity (the bucket's problems
            // + one spare entry per resident lane).  Batches sized to fill the lanes stay far below it; as
            // insurance no batch is pushed into a ring that is already half full.
#ifndef FX_RING_GUARD
#define FX_RING_GUARD 1
#endif
            if (FX_RING_GUARD && d > 1) {
              const unsigned long long tl = *(volatile unsigned long long*)&ctx.ctl->tail;
              unsigned long long hd = *(volatile unsigned long long*)&ctx.ctl->head;
              if (hd < (unsigned long long)ctx.n_work) hd = (unsigned long long)ctx.n_work;
              if (tl > hd && (long long)(tl - hd) > ctx.ring_cap / 2) d = 1;
            }
            if (d > 1 && bs.y == 0) {
              unsigned int sl = atomicAdd(A.spec_next, 1u);
              if (sl < A.spec_slots) bs.y = (int)sl + 1; else d = 1;
            }
          }
          bs.x = bis.phase | (bis.lo_neg << 2) | (bis.n_bis << 3);
          __stcg(&A.bstate[prob], bs);
          if (d > 1) {
            SpecEntry* hdr = A.spec + (size_t)(bs.y - 1) * SPEC_NODES;
            __stcg(&hdr->k, make_int4((1 << d) - 1, d, 0, 0));
          }
          __threadfence();
          if (d > 1) push_tasks(ctx.ctl, ctx.ring, ctx.n_work, ctx.ring_cap, prob, 1, (1 << d) - 1);
          else push_tasks(ctx.ctl, ctx.ring, ctx.n_work, ctx.ring_cap, prob, 0, 1);
        }
      }
      mode = M_FETCH;
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

// sets each bucket's queue control block from the bucket sizes (one thread per bucket)
__global__ void init_queues_kernel(const unsigned long long* counts, QueueCtl* ctl) {
  int m = threadIdx.x;
  if (m <= FX_NUM_MODELS) {
    long long n = (long long)counts[m];
    ctl[m].head = 0;
    ctl[m].tail = (unsigned long long)n;
    ctl[m].avail = n;
    ctl[m].remaining = n;
    ctl[m].abort = 0;
  }
}

}  // namespace fx
