// fx_api.cu -- the C ABI of libfrontx_b200.so (include/frontx_b200.h) and the
// host-side launch logic: device-side bucketing of a mixed batch by model, one
// persistent launch per model on forked streams, the dense-output evaluation
// and diffusivity kernels, and a DFMA micro-benchmark for the FP64 roofline.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/frontx_b200.h"
#include "fx_inverse.cuh"
#include "fx_solve.cuh"

// A mixed batch of at least this many problems per resident lane runs its model buckets one after another
// (see solve_impl); below it they run side by side.  Scheduling only: results do not depend on it.
#ifndef FX_SEQUENTIAL_FACTOR
#define FX_SEQUENTIAL_FACTOR 3
#endif

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof g_err, fmt, detail);
  return code;
}

#define FX_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return fail(FX_ERR_CUDA, #call ": %s", cudaGetErrorString(e__)); \
  } while (0)

struct DeviceInfo {
  int sms = 0;
  bool ok = false;
};

int device_info(DeviceInfo* out) {
  int dev = 0;
  FX_CUDA(cudaGetDevice(&dev));
  static std::mutex mu;
  static DeviceInfo cache[64];
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 0 || dev >= 64) return fail(FX_ERR_CUDA, "device index out of range%s");
  if (!cache[dev].ok) {
    FX_CUDA(cudaDeviceGetAttribute(&cache[dev].sms, cudaDevAttrMultiProcessorCount, dev));
    // the per-call scratch comes from the stream-ordered pool: keep freed blocks cached across
    // synchronisations instead of returning them to the driver after every call
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    cache[dev].ok = true;
  }
  *out = cache[dev];
  return FX_OK;
}

// ---- bucketing of a mixed batch by model -----------------------------------
// scratch layout (unsigned long long): counts[16] | cursors[16] | speculation-slot allocator
constexpr int NSLOT = FX_NUM_MODELS + 1;  // one bucket per model + one for problems that are not solved
constexpr int SCR_COUNTS = 0, SCR_CURSORS = 16, SCR_WORDS = 32;
static_assert(NSLOT <= 16, "scratch layout");

// `mask` = the models the caller says are present (bit m; all seven when it does not know): a problem
// whose model is outside it, or is no model at all, is parked in the last slot and flagged, never solved
__device__ __forceinline__ int bucket_of(int m, unsigned mask) {
  return (m < 0 || m >= FX_NUM_MODELS || !((mask >> m) & 1u)) ? FX_NUM_MODELS : m;
}

__global__ void count_models_kernel(long long n, const int* __restrict__ model_id, unsigned mask,
                                    unsigned long long* scratch) {
  __shared__ unsigned int local[NSLOT];
  if (threadIdx.x < NSLOT) local[threadIdx.x] = 0;
  __syncthreads();
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    atomicAdd(&local[bucket_of(model_id[j], mask)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < NSLOT && local[threadIdx.x])
    atomicAdd(&scratch[SCR_COUNTS + threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

__global__ void scatter_models_kernel(long long n, const int* __restrict__ model_id, unsigned mask,
                                      unsigned long long* scratch, int* __restrict__ perm,
                                      int* __restrict__ counters) {
  // bucket-major placement; order inside a bucket is by warp-aggregated cursor
  long long base[NSLOT];
  long long acc = 0;
  for (int k = 0; k < NSLOT; k++) {
    base[k] = acc;
    acc += (long long)scratch[SCR_COUNTS + k];
  }
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const int m = bucket_of(model_id[j], mask);
    if (m == FX_NUM_MODELS) counters[j * FX_NCNT + 0] = FX_INVALID_MODEL;  // never solved
    unsigned long long pos = atomicAdd(&scratch[SCR_CURSORS + m], 1ull);
    perm[base[m] + (long long)pos] = (int)j;
  }
}

template <int MODEL, int GEOM>
int launch_geom(const fx::SolveArgs& a, int grid, cudaStream_t s) {
  fx::shoot_bisect_kernel<MODEL, GEOM><<<grid, fx::SOLVE_BLOCK, 0, s>>>(a);
  g_launches++;
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

// geom: 0 planar, 1 radial, 2 radial with a flow-rate boundary (template parameter GEOM)
template <int MODEL>
int launch_model(const fx::SolveArgs& a, int geom, int grid, cudaStream_t s) {
  if (geom == 2) return launch_geom<MODEL, 2>(a, grid, s);
  if (geom == 1) return launch_geom<MODEL, 1>(a, grid, s);
  return launch_geom<MODEL, 0>(a, grid, s);
}

template <int MODEL, int GEOM>
int occupancy_geom(int* blocks) {
  // the kernel keeps its per-lane state in shared memory: ask for the largest carve-out
  FX_CUDA(cudaFuncSetAttribute(fx::shoot_bisect_kernel<MODEL, GEOM>,
                               cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  FX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, fx::shoot_bisect_kernel<MODEL, GEOM>,
                                                        fx::SOLVE_BLOCK, 0));
  return FX_OK;
}

template <int MODEL>
int occupancy_of(int geom, int* blocks) {
  if (geom == 2) return occupancy_geom<MODEL, 2>(blocks);
  if (geom == 1) return occupancy_geom<MODEL, 1>(blocks);
  return occupancy_geom<MODEL, 0>(blocks);
}

int launch_by_id(int model, const fx::SolveArgs& a, int geom, long long n_upper, int sms,
                 cudaStream_t s) {
  int bps = 1, rc = FX_OK;
  switch (model) {
    case FX_CONSTANT: rc = occupancy_of<FX_CONSTANT>(geom, &bps); break;
    case FX_POWER_LAW: rc = occupancy_of<FX_POWER_LAW>(geom, &bps); break;
    case FX_BROOKS_COREY: rc = occupancy_of<FX_BROOKS_COREY>(geom, &bps); break;
    case FX_VAN_GENUCHTEN: rc = occupancy_of<FX_VAN_GENUCHTEN>(geom, &bps); break;
    case FX_LETXS: rc = occupancy_of<FX_LETXS>(geom, &bps); break;
    case FX_LETD: rc = occupancy_of<FX_LETD>(geom, &bps); break;
    case FX_LOG: rc = occupancy_of<FX_LOG>(geom, &bps); break;
    case FX_TABULATED: rc = occupancy_of<FX_TABULATED>(geom, &bps); break;
    default: return fail(FX_ERR_MODEL, "unknown model id%s");
  }
  if (rc) return rc;
  if (bps < 1) bps = 1;
  long long want = (n_upper + fx::SOLVE_BLOCK - 1) / fx::SOLVE_BLOCK;
  long long cap = (long long)sms * bps;  // persistent grid: every CTA resident
  // whole CTAs per SM, the same number on every SM (the work queue hands lanes beyond n nothing)
  long long even = (want + sms - 1) / sms * sms;
  int grid = (int)(even < cap ? even : cap);
  if (grid < 1) grid = 1;
  switch (model) {
    case FX_CONSTANT: return launch_model<FX_CONSTANT>(a, geom, grid, s);
    case FX_POWER_LAW: return launch_model<FX_POWER_LAW>(a, geom, grid, s);
    case FX_BROOKS_COREY: return launch_model<FX_BROOKS_COREY>(a, geom, grid, s);
    case FX_VAN_GENUCHTEN: return launch_model<FX_VAN_GENUCHTEN>(a, geom, grid, s);
    case FX_LETXS: return launch_model<FX_LETXS>(a, geom, grid, s);
    case FX_LETD: return launch_model<FX_LETD>(a, geom, grid, s);
    case FX_LOG: return launch_model<FX_LOG>(a, geom, grid, s);
    default: return launch_model<FX_TABULATED>(a, geom, grid, s);
  }
}

// Helper streams and events that run the per-model launches of one call side by side.  One set per
// (device, caller stream): two host threads solving on two streams share nothing, and a set's mutex is
// held for the whole fork/join enqueue section of a call so that record/wait pairs never interleave.
struct Fork {
  int dev = -1;
  cudaStream_t owner = nullptr;
  cudaStream_t streams[FX_NUM_MODELS] = {};
  cudaEvent_t begin = nullptr, done[FX_NUM_MODELS] = {};
  cudaEvent_t k0[FX_NUM_MODELS] = {}, k1[FX_NUM_MODELS] = {};  // kernel timing (fx_set_kernel_timing)
  bool timed[FX_NUM_MODELS] = {};
  std::mutex mu;
};

std::atomic<int> g_kernel_timing{0};
std::mutex g_fork_mu;
std::vector<std::unique_ptr<Fork>> g_forks;
Fork* g_last_timed[64][FX_NUM_MODELS] = {};  // per device: the set that holds the latest timed launch of a model
constexpr size_t MAX_FORKS = 64;             // beyond that many (device, stream) pairs, sets are shared

int get_fork(cudaStream_t owner, Fork** out) {
  int dev = 0;
  FX_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_fork_mu);
  Fork* same_dev = nullptr;
  for (auto& f : g_forks) {
    if (f->dev != dev) continue;
    if (f->owner == owner) { *out = f.get(); return FX_OK; }
    same_dev = f.get();
  }
  if (g_forks.size() >= MAX_FORKS && same_dev) { *out = same_dev; return FX_OK; }
  std::unique_ptr<Fork> f(new Fork);
  f->dev = dev;
  f->owner = owner;
  for (int m = 0; m < FX_NUM_MODELS; m++) {
    FX_CUDA(cudaStreamCreateWithFlags(&f->streams[m], cudaStreamNonBlocking));
    FX_CUDA(cudaEventCreateWithFlags(&f->done[m], cudaEventDisableTiming));
    FX_CUDA(cudaEventCreate(&f->k0[m]));
    FX_CUDA(cudaEventCreate(&f->k1[m]));
  }
  FX_CUDA(cudaEventCreateWithFlags(&f->begin, cudaEventDisableTiming));
  *out = f.get();
  g_forks.push_back(std::move(f));
  return FX_OK;
}

// stream-ordered scratch of one call, released on every way out of it
struct Scratch {
  cudaStream_t s;
  void* p = nullptr;
  explicit Scratch(cudaStream_t s_) : s(s_) {}
  ~Scratch() { if (p) cudaFreeAsync(p, s); }
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
};

int check_options(const fx_solve_options* in, fx_solve_options* o) {
  if (in) *o = *in; else fx_default_solve_options(o);
  if (!(o->itol > 0.0) || o->max_bisect < 1 || o->max_ode_steps < 1 || !(o->rtol > 0.0) ||
      !(o->atol >= 0.0) || o->max_expansions < 0)
    return fail(FX_ERR_ARG, "invalid solve options%s");
  if (o->radial_k < 0 || o->radial_k > 2) return fail(FX_ERR_ARG, "radial_k must be 0, 1 or 2%s");
  if (o->radial_k > 0 && !(o->ob > 0.0)) return fail(FX_ERR_ARG, "radial problems need ob > 0%s");
  if (o->model_mask >> FX_NUM_MODELS) return fail(FX_ERR_ARG, "model_mask has bits beyond the known models%s");
  return FX_OK;
}

// ---- dense output evaluation (Solution.__call__/d_do, _forward.py:25-37) -----
__global__ void eval_kernel(long long n, const int* __restrict__ counters,
                            const double* __restrict__ knots, int k_max, long long q,
                            const double* __restrict__ o_query, int per_problem,
                            double* __restrict__ theta, double* __restrict__ dtheta) {
  long long total = n * q;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    long long j = g / q, c = g - j * q;
    int nk = counters[j * FX_NCNT + 7];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double th = nan, dth = nan;
    if (nk >= 1 && nk <= k_max) {
      const double4* K = reinterpret_cast<const double4*>(knots + (size_t)j * k_max * FX_KNOT);
      double t = o_query[per_problem ? g : c];
      double4 first = K[0], last = K[nk - 1];
      t = fmin(fmax(t, first.x), last.x);  // jnp.clip(o, 0, oi); NaN query stays NaN below
      if (o_query[per_problem ? g : c] != o_query[per_problem ? g : c]) t = nan;
      if (nk == 1) {
        th = first.y;
        dth = first.z;
      } else {
        // index = clip(searchsorted(ts, t, side="left") - 1, 0, nk-2)
        int lo = 0, hi = nk;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (K[mid].x < t) lo = mid + 1; else hi = mid;
        }
        int idx = min(max(lo - 1, 0), nk - 2);
        double4 a = K[idx], bb = K[idx + 1];
        double dt = bb.x - a.x;
        double s = (t - a.x) / ((dt == 0.0) ? 1.0 : dt);
        {  // theta: k = dt * theta'
          double k0 = dt * a.z, k1 = dt * bb.z;
          double ca = k0 + k1 + 2.0 * a.y - 2.0 * bb.y;
          double cb = -2.0 * k0 - k1 - 3.0 * a.y + 3.0 * bb.y;
          th = fma(fma(fma(ca, s, cb), s, k0), s, a.y);
        }
        {  // dtheta/do is the SECOND STATE COMPONENT's interpolant: k = dt * theta''
          double k0 = dt * a.w, k1 = dt * bb.w;
          double ca = k0 + k1 + 2.0 * a.z - 2.0 * bb.z;
          double cb = -2.0 * k0 - k1 - 3.0 * a.z + 3.0 * bb.z;
          dth = fma(fma(fma(ca, s, cb), s, k0), s, a.z);
        }
      }
    }
    if (theta) theta[g] = th;
    if (dtheta) dtheta[g] = dth;
  }
}

// ---- ScaledSolution.fitting_data + the cost of fit(), a batch of candidates at a time -----------
// (src/frontx/_inverse/fit.py:45-82 and :145-151).  A candidate's solved profile theta(o) is
// compared with m data points through the one-parameter rescaling o -> o/sqrt(D0); D0 is fitted by
// damped Gauss-Newton on u = ln D0 (the reference runs optimistix Levenberg-Marquardt on D0 itself,
// same minimiser), then cost = mean(((theta(o_k/sqrt(D0)) - theta_k) w_k)^2).  One warp per
// candidate, data points strided over its lanes, fixed shuffle-tree reductions.
__device__ __forceinline__ void hermite_theta(const double4* K, int nk, double x, double& th, double& dth) {
  // theta and d theta/dx of the dense-output cubic at x, clipped to the table (jnp.clip, _forward.py:30)
  const double4 first = K[0], last = K[nk - 1];
  const bool inside = (x > first.x) && (x < last.x);
  x = fmin(fmax(x, first.x), last.x);
  if (nk == 1) {
    th = first.y;
    dth = 0.0;
    return;
  }
  int lo = 0, hi = nk;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (K[mid].x < x) lo = mid + 1; else hi = mid;
  }
  int idx = min(max(lo - 1, 0), nk - 2);
  double4 a = K[idx], b = K[idx + 1];
  double dt = b.x - a.x;
  double idt = (dt == 0.0) ? 0.0 : 1.0 / dt;
  double s = (x - a.x) * idt;
  double k0 = dt * a.z, k1 = dt * b.z;
  double ca = k0 + k1 + 2.0 * a.y - 2.0 * b.y;
  double cb = -2.0 * k0 - k1 - 3.0 * a.y + 3.0 * b.y;
  th = fma(fma(fma(ca, s, cb), s, k0), s, a.y);
  dth = inside ? fma(fma(3.0 * ca, s, 2.0 * cb), s, k0) * idt : 0.0;
}

__device__ __forceinline__ double warp_total(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// mode 0: D0 = 1 (cost of the unscaled solution); 1: fit D0 to the data; 2: D0 given in D0_io
__global__ void scaled_fit_kernel(long long n, const int* __restrict__ counters,
                                  const double* __restrict__ summary, const double* __restrict__ knots,
                                  int k_max, int m, const double* __restrict__ o, const double* __restrict__ theta,
                                  const double* __restrict__ weight, int mode, double* __restrict__ D0_io,
                                  double* __restrict__ cost, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  for (long long j = warp; j < n; j += nwarps) {
    const int nk = counters[j * FX_NCNT + 7];
    const bool solved = counters[j * FX_NCNT + 0] == 0 && nk >= 1 && nk <= k_max;
    if (!solved) {  // fit.py:145-151: a failed solve costs inf
      if (lane == 0) {
        cost[j] = inf;
        status[j] = 1;
        if (mode != 2) D0_io[j] = 1.0;
      }
      continue;
    }
    const double4* K = reinterpret_cast<const double4*>(knots + (size_t)j * k_max * FX_KNOT);
    const double oi = summary[j * FX_NSUM + 1];
    auto sums = [&](double D0, double& c, double& g, double& h) {
      // c = sum r^2, g = sum r dr/du, h = sum (dr/du)^2 with u = ln D0, r = (theta(o/sqrt D0) - theta_k) w_k
      const double isq = rsqrt(D0);
      double sc = 0.0, sg = 0.0, sh = 0.0;
      for (int k = lane; k < m; k += 32) {
        double th, dth;
        const double x = o[k] * isq;
        hermite_theta(K, nk, x, th, dth);
        const double w = weight ? weight[k] : 1.0;
        const double r = (th - theta[k]) * w;
        const double ju = -0.5 * dth * x * w;
        sc = fma(r, r, sc);
        sg = fma(r, ju, sg);
        sh = fma(ju, ju, sh);
      }
      c = warp_total(sc);
      g = warp_total(sg);
      h = warp_total(sh);
    };
    double D0 = 1.0;
    int st = 0;
    double c, g, h;
    if (mode == 2) D0 = D0_io[j];
    if (mode == 1) {
      const double ratio = o[m - 1] / oi;  // fit.py:66: y0 = (o[-1]/original.oi)^2
      double u = 2.0 * log(ratio > 0.0 ? ratio : 1.0);
      double lam = 1e-3;
      sums(exp(u), c, g, h);
      st = 1;  // max_steps_reached unless the loop converges
      for (int it = 0; it < 200; it++) {
        if (!(h > 0.0)) { st = (g == 0.0) ? 0 : 2; break; }  // flat: nothing to fit / no direction
        // trust region: D0 changes by at most e^1.5 per step.  Far from the minimum only a few data points
        // see the front, h is tiny and the raw Gauss-Newton step can jump to D0 = inf (where the curve is the
        // constant b and the cost happens to be lower than at a poor start)
        const double du = fmax(-1.5, fmin(1.5, -g / (h * (1.0 + lam))));
        double c2, g2, h2;
        sums(exp(u + du), c2, g2, h2);
        if (c2 <= c) {
          u += du;
          const bool small = fabs(du) <= 1e-13 + 1e-11 * fabs(u) || (c - c2) <= 1e-16 * c;
          c = c2; g = g2; h = h2;
          lam = fmax(lam * 0.2, 1e-12);
          if (small) { st = 0; break; }
        } else {
          lam *= 10.0;
          if (lam > 1e12) { st = 0; break; }  // cannot decrease any further: at the minimum to rounding
        }
      }
      // Polish: the minimiser is the root of g(u) = sum r dr/du, which the sums give analytically; a few secant
      // steps from the last Gauss-Newton iterate pin it to rounding (the damped iteration stops once the COST
      // no longer decreases, which leaves u good to ~1e-8 only).
      if (st == 0 && h > 0.0) {
        double u0 = u, g0 = g, u1 = u - g / h, c1, g1, h1;
        sums(exp(u1), c1, g1, h1);
        for (int it = 0; it < 12 && g1 != g0 && g1 != 0.0; it++) {
          const double u2 = u1 - g1 * (u1 - u0) / (g1 - g0);
          if (!(fabs(u2 - u) < 1e-4)) break;  // stay at the minimum the iteration found
          u0 = u1; g0 = g1;
          u1 = u2;
          sums(exp(u1), c1, g1, h1);
          if (fabs(u1 - u0) <= 4e-16 * fmax(1.0, fabs(u1))) break;
        }
        if (fabs(g1) <= fabs(g) && fabs(u1 - u) < 1e-4) { u = u1; c = c1; }
      }
      D0 = exp(u);
    } else {
      sums(D0, c, g, h);
    }
    if (lane == 0) {
      D0_io[j] = D0;
      cost[j] = c / (double)m;
      status[j] = st;
    }
  }
}

// ---- D, D', D'' at arbitrary theta (Solution.D for flux/sorptivity) ----------
template <int MODEL>
__device__ void diffusivity_one(const double* p, double th, double* out) {
  double d[fx::NDERIVED];
  fx::derive_params<MODEL>(p, d, fx::g_math_tab_dev);
  double D, D1, D2;
  double iD;
  fx::eval_D<MODEL>(d, th, true, D, D1, D2, iD, fx::g_math_tab_dev);
  out[0] = D;
  out[1] = D1;
  out[2] = D2;
}

__global__ void diffusivity_kernel(long long n, const int* __restrict__ model_id,
                                   const double* __restrict__ params, long long q,
                                   const double* __restrict__ theta, double* __restrict__ out) {
  long long total = n * q;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    long long j = g / q;
    const double* p = params + j * FX_NPARAM;
    double th = theta[g];
    double* o = out + g * 3;
    switch (model_id[j]) {
      case FX_CONSTANT: diffusivity_one<FX_CONSTANT>(p, th, o); break;
      case FX_POWER_LAW: diffusivity_one<FX_POWER_LAW>(p, th, o); break;
      case FX_BROOKS_COREY: diffusivity_one<FX_BROOKS_COREY>(p, th, o); break;
      case FX_VAN_GENUCHTEN: diffusivity_one<FX_VAN_GENUCHTEN>(p, th, o); break;
      case FX_LETXS: diffusivity_one<FX_LETXS>(p, th, o); break;
      case FX_LETD: diffusivity_one<FX_LETD>(p, th, o); break;
      case FX_LOG: diffusivity_one<FX_LOG>(p, th, o); break;
      case FX_TABULATED: diffusivity_one<FX_TABULATED>(p, th, o); break;
      default: o[0] = o[1] = o[2] = __longlong_as_double(0x7ff8000000000000LL);
    }
  }
}

// ---- FP64 FMA throughput (roofline denominator) ------------------------------
__global__ void dfma_peak_kernel(double* sink, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
#pragma unroll 1
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) sink[0] = s;
}

// flow-rate boundary: fills summary[3] = d_dob = -q/D(b) and summary[4] = D(b) at the accepted b
template <int MODEL>
__device__ void flow_finalize_one(const double* p, double q, double* S) {
  double d[fx::NDERIVED];
  fx::derive_params<MODEL>(p, d, fx::g_math_tab_dev);
  double D, D1, D2, iD;
  fx::eval_D<MODEL>(d, S[0], false, D, D1, D2, iD, fx::g_math_tab_dev);
  S[3] = -q * iD;
  S[4] = D;
}

__global__ void flow_finalize_kernel(long long n, const int* __restrict__ model_id,
                                     const double* __restrict__ params, const double* __restrict__ q,
                                     double* summary, const int* __restrict__ counters) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    if (counters[j * FX_NCNT] < 0) continue;  // invalid model id: never solved
    const double* p = params + j * FX_NPARAM;
    double* S = summary + j * FX_NSUM;
    switch (model_id[j]) {
      case FX_CONSTANT: flow_finalize_one<FX_CONSTANT>(p, q[j], S); break;
      case FX_POWER_LAW: flow_finalize_one<FX_POWER_LAW>(p, q[j], S); break;
      case FX_BROOKS_COREY: flow_finalize_one<FX_BROOKS_COREY>(p, q[j], S); break;
      case FX_VAN_GENUCHTEN: flow_finalize_one<FX_VAN_GENUCHTEN>(p, q[j], S); break;
      case FX_LETXS: flow_finalize_one<FX_LETXS>(p, q[j], S); break;
      case FX_LETD: flow_finalize_one<FX_LETD>(p, q[j], S); break;
      case FX_LOG: flow_finalize_one<FX_LOG>(p, q[j], S); break;
      case FX_TABULATED: flow_finalize_one<FX_TABULATED>(p, q[j], S); break;
      default: break;
    }
  }
}

int solve_impl(int64_t n, const int32_t* model_id, const double* params, const double* b,
               const double* i, const double* q, const double* b_hi, const fx_solve_options* opt_in,
               double* summary, int32_t* counters, int32_t k_max, double* knots, void* stream);

int grid_for(long long total, int block, int sms) {
  long long want = (total + block - 1) / block;
  long long cap = (long long)sms * 16;
  long long g = want < cap ? want : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" {

void fx_default_solve_options(fx_solve_options* o) {
  o->itol = 1e-3;
  o->max_bisect = 100;
  o->max_ode_steps = 4096;
  o->rtol = 1e-3;
  o->atol = 1e-6;
  o->max_expansions = 64;
  o->radial_k = 0;
  o->ob = 0.0;
  o->model_mask = 0;
}

const char* fx_last_error(void) { return g_err; }
int fx_version(void) { return FX_VERSION; }
long long fx_launch_count_ll(void) { return g_launches.load(); }
int64_t fx_launch_count(void) { return (int64_t)g_launches.load(); }

void fx_set_kernel_timing(int on) { g_kernel_timing.store(on ? 1 : 0); }

int fx_last_solve_kernel_ms(int model, float* ms) {
  if (model < 0 || model >= FX_NUM_MODELS || !ms) return fail(FX_ERR_ARG, "bad model id or null pointer%s");
  int dev = 0;
  FX_CUDA(cudaGetDevice(&dev));
  Fork* fk = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_fork_mu);
    if (dev >= 0 && dev < 64) fk = g_last_timed[dev][model];
  }
  if (!fk) return fail(FX_ERR_ARG, "no timed launch recorded: call fx_set_kernel_timing(1) first%s");
  std::lock_guard<std::mutex> lk(fk->mu);
  if (!fk->timed[model]) return fail(FX_ERR_ARG, "no timed launch recorded: call fx_set_kernel_timing(1) first%s");
  FX_CUDA(cudaEventSynchronize(fk->k1[model]));
  FX_CUDA(cudaEventElapsedTime(ms, fk->k0[model], fk->k1[model]));
  return FX_OK;
}

int fx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int fx_solve_batch(int64_t n, const int32_t* model_id, const double* params, const double* b,
                   const double* i, const fx_solve_options* opt_in, double* summary,
                   int32_t* counters, int32_t k_max, double* knots, void* stream) {
  return solve_impl(n, model_id, params, b, i, nullptr, nullptr, opt_in, summary, counters, k_max, knots, stream);
}

int fx_solve_flowrate_batch(int64_t n, const int32_t* model_id, const double* params, const double* q,
                            const double* i, const double* b_lo, const double* b_hi,
                            const fx_solve_options* opt_in, double* summary, int32_t* counters,
                            int32_t k_max, double* knots, void* stream) {
  if (n > 0 && (!q || !b_lo || !b_hi)) return fail(FX_ERR_ARG, "null pointer argument%s");
  fx_solve_options o;
  if (opt_in) o = *opt_in; else { fx_default_solve_options(&o); o.radial_k = 1; o.ob = 1e-6; }
  if (o.radial_k < 1) return fail(FX_ERR_ARG, "a flow-rate boundary needs radial_k = 1 or 2 and ob > 0%s");
  o.max_expansions = 0;  // the bracket on b is the caller's; it is not widened
  return solve_impl(n, model_id, params, b_lo, i, q, b_hi, &o, summary, counters, k_max, knots, stream);
}

}  // extern "C"

namespace {

int solve_impl(int64_t n, const int32_t* model_id, const double* params, const double* b,
               const double* i, const double* q, const double* b_hi, const fx_solve_options* opt_in,
               double* summary, int32_t* counters, int32_t k_max, double* knots, void* stream) {
  if (n < 0 || k_max < 0) return fail(FX_ERR_ARG, "negative size%s");
  if (n == 0) return FX_OK;
  if (!model_id || !params || !b || !i || !summary || !counters)
    return fail(FX_ERR_ARG, "null pointer argument%s");
  if (k_max > 0 && !knots) return fail(FX_ERR_ARG, "k_max > 0 needs a knots buffer%s");
  if (n > 2147483647LL) return fail(FX_ERR_ARG, "n exceeds 2^31-1 problems per call%s");
  fx_solve_options o;
  int rc = check_options(opt_in, &o);
  if (rc) return rc;
  DeviceInfo di;
  if ((rc = device_info(&di))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  Fork* fk = nullptr;
  if ((rc = get_fork(s, &fk))) return rc;

  // The caller may say which models the batch holds (opt->model_mask): only those get a ring region and a
  // launch.  Without the hint every model is launched and the empty buckets' grids exit at once.
  const unsigned mask = o.model_mask ? o.model_mask : ((1u << FX_NUM_MODELS) - 1u);
  int n_regions = 0, region_of[FX_NUM_MODELS];
  for (int m = 0; m < FX_NUM_MODELS; m++) region_of[m] = ((mask >> m) & 1u) ? n_regions++ : 0;

  // Scratch, ONE stream-ordered allocation: bucket counters and cursors, the speculation-slot allocator,
  // the queue control blocks, 8 bytes of bisection state per problem, the shot ring (one region per
  // bucket present: its problems + one spare entry per resident lane) -- all of that zeroed by one
  // memset -- then the bucket-major permutation and the speculation slots.
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const long long max_lanes = (long long)di.sms * 2048;
  const long long ring_pad = max_lanes + 64;
  const size_t ring_words = (size_t)n + (size_t)n_regions * (size_t)ring_pad;
  // one slot per problem that ever speculates (it keeps it): ~116 MB of pool memory whatever the depth
  long long spec_slots = (long long)di.sms * (16384 / fx::SPEC_NODES) + 8;
  if (spec_slots > n) spec_slots = n;
  const size_t off_ctl = up((SCR_WORDS + 8) * sizeof(unsigned long long));
  const size_t off_bstate = off_ctl + up(NSLOT * sizeof(fx::QueueCtl));
  const size_t off_ring = off_bstate + up((size_t)n * sizeof(int2));
  const size_t off_perm = off_ring + up(ring_words * sizeof(unsigned long long));
  const size_t off_spec = off_perm + up((size_t)n * sizeof(int));
  const size_t total = off_spec + (size_t)spec_slots * fx::SPEC_NODES * sizeof(fx::SpecEntry);
  Scratch scr(s);
  FX_CUDA(cudaMallocAsync(&scr.p, total, s));
  char* base = static_cast<char*>(scr.p);
  unsigned long long* scratch = reinterpret_cast<unsigned long long*>(base);
  fx::QueueCtl* ctl = reinterpret_cast<fx::QueueCtl*>(base + off_ctl);
  int2* bstate = reinterpret_cast<int2*>(base + off_bstate);
  unsigned long long* ring = reinterpret_cast<unsigned long long*>(base + off_ring);
  int* perm = reinterpret_cast<int*>(base + off_perm);
  fx::SpecEntry* spec = spec_slots > 0 ? reinterpret_cast<fx::SpecEntry*>(base + off_spec) : nullptr;
  FX_CUDA(cudaMemsetAsync(base, 0, off_perm, s));
  int g = grid_for(n, 256, di.sms);
  count_models_kernel<<<g, 256, 0, s>>>(n, model_id, mask, scratch);
  g_launches++;
  scatter_models_kernel<<<g, 256, 0, s>>>(n, model_id, mask, scratch, perm, counters);
  g_launches++;
  fx::init_queues_kernel<<<1, 32, 0, s>>>(scratch + SCR_COUNTS, ctl);
  g_launches++;
  FX_CUDA(cudaGetLastError());

  fx::SolveArgs a;
  memset(&a, 0, sizeof a);
  a.counts = scratch + SCR_COUNTS;
  a.perm = perm;
  a.params = params;
  a.b = b;
  a.i = i;
  a.q = q;
  a.b_hi = b_hi;
  a.summary = summary;
  a.counters = counters;
  a.knots = k_max > 0 ? knots : nullptr;
  a.k_max = k_max;
  a.ctl = ctl;
  a.ring = ring;
  a.ring_pad = ring_pad;
  a.bstate = bstate;
  a.spec = spec;
  a.spec_next = reinterpret_cast<unsigned int*>(scratch + SCR_WORDS);
  a.spec_slots = (unsigned int)spec_slots;
  a.itol = o.itol;
  a.rtol = o.rtol;
  a.atol = o.atol;
  a.ob = o.ob;
  a.max_bisect = o.max_bisect;
  a.max_ode_steps = o.max_ode_steps;
  a.max_expansions = o.max_expansions;
  a.radial_k = o.radial_k;

  {
    // One persistent launch per model present.  A batch of a few problems per lane is a latency chain per
    // problem (~18 sequential shots): its buckets run SIDE BY SIDE on forked streams, each on a share of the
    // CTAs.  A large mixed batch is throughput: its buckets run ONE AFTER ANOTHER on the caller's stream, each on
    // the whole grid, so that no SM idles because one model's problems are cheaper than another's.
    std::lock_guard<std::mutex> lk(fk->mu);
    const bool timing = g_kernel_timing.load() != 0;
    const int geom = q ? 2 : (o.radial_k > 0 ? 1 : 0);
    const bool sequential = n_regions > 1 && n >= (long long)FX_SEQUENTIAL_FACTOR * di.sms * fx::SOLVE_CTAS_PER_SM * fx::SOLVE_BLOCK;
    a.full_grid = sequential ? 1 : 0;
    if (!sequential) FX_CUDA(cudaEventRecord(fk->begin, s));
    for (int m = 0; m < FX_NUM_MODELS; m++) {
      if (!((mask >> m) & 1u)) continue;
      cudaStream_t ks = sequential ? s : fk->streams[m];
      if (!sequential) FX_CUDA(cudaStreamWaitEvent(ks, fk->begin, 0));
      a.slot = m;
      a.ring_region = region_of[m];
      if (timing) FX_CUDA(cudaEventRecord(fk->k0[m], ks));
      if ((rc = launch_by_id(m, a, geom, n, di.sms, ks))) return rc;
      if (timing) FX_CUDA(cudaEventRecord(fk->k1[m], ks));
      fk->timed[m] = timing;
      if (!sequential) {
        FX_CUDA(cudaEventRecord(fk->done[m], ks));
        FX_CUDA(cudaStreamWaitEvent(s, fk->done[m], 0));
      }
    }
    if (timing) {
      std::lock_guard<std::mutex> lk2(g_fork_mu);
      int dev = fk->dev;
      if (dev >= 0 && dev < 64)
        for (int m = 0; m < FX_NUM_MODELS; m++)
          if ((mask >> m) & 1u) g_last_timed[dev][m] = fk;
    }
  }
  if (q) {
    // flow-rate boundary: d_dob = -q/D(b) and D(b) at the accepted b, by the same evaluation the shots used
    flow_finalize_kernel<<<grid_for(n, 256, di.sms), 256, 0, s>>>(n, model_id, params, q, summary, counters);
    g_launches++;
    FX_CUDA(cudaGetLastError());
  }
  return FX_OK;  // ~Scratch returns the scratch to the pool in stream order
}

}  // namespace

extern "C" {

int fx_solve_batch_host(int64_t n, const int32_t* model_id, const double* params, const double* b,
                        const double* i, const fx_solve_options* opt, double* summary,
                        int32_t* counters, int32_t k_max, double* knots) {
  if (n < 0 || k_max < 0) return fail(FX_ERR_ARG, "negative size%s");
  if (n == 0) return FX_OK;
  if (!model_id || !params || !b || !i || !summary || !counters)
    return fail(FX_ERR_ARG, "null pointer argument%s");
  if (k_max > 0 && !knots) return fail(FX_ERR_ARG, "k_max > 0 needs a knots buffer%s");
  // one stream per host thread and device, kept: its fork set (get_fork) is then reused call after call
  thread_local cudaStream_t t_stream[64] = {};
  int dev = 0;
  FX_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(FX_ERR_CUDA, "device index out of range%s");
  if (!t_stream[dev]) FX_CUDA(cudaStreamCreateWithFlags(&t_stream[dev], cudaStreamNonBlocking));
  cudaStream_t s = t_stream[dev];
  int32_t *d_model = nullptr, *d_cnt = nullptr;
  double *d_par = nullptr, *d_b = nullptr, *d_i = nullptr, *d_sum = nullptr, *d_knots = nullptr;
  size_t nn = (size_t)n;
  int rc = FX_OK;
  auto body = [&]() -> int {
    FX_CUDA(cudaMallocAsync((void**)&d_model, nn * sizeof(int32_t), s));
    FX_CUDA(cudaMallocAsync((void**)&d_par, nn * FX_NPARAM * sizeof(double), s));
    FX_CUDA(cudaMallocAsync((void**)&d_b, nn * sizeof(double), s));
    FX_CUDA(cudaMallocAsync((void**)&d_i, nn * sizeof(double), s));
    FX_CUDA(cudaMallocAsync((void**)&d_sum, nn * FX_NSUM * sizeof(double), s));
    FX_CUDA(cudaMallocAsync((void**)&d_cnt, nn * FX_NCNT * sizeof(int32_t), s));
    if (k_max > 0) FX_CUDA(cudaMallocAsync((void**)&d_knots, nn * k_max * FX_KNOT * sizeof(double), s));
    FX_CUDA(cudaMemcpyAsync(d_model, model_id, nn * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemcpyAsync(d_par, params, nn * FX_NPARAM * sizeof(double), cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemcpyAsync(d_b, b, nn * sizeof(double), cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemcpyAsync(d_i, i, nn * sizeof(double), cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemsetAsync(d_cnt, 0, nn * FX_NCNT * sizeof(int32_t), s));
    int r = fx_solve_batch(n, d_model, d_par, d_b, d_i, opt, d_sum, d_cnt, k_max, d_knots, s);
    if (r) return r;
    FX_CUDA(cudaMemcpyAsync(summary, d_sum, nn * FX_NSUM * sizeof(double), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(counters, d_cnt, nn * FX_NCNT * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (k_max > 0)
      FX_CUDA(cudaMemcpyAsync(knots, d_knots, nn * k_max * FX_KNOT * sizeof(double), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    return FX_OK;
  };
  rc = body();
  void* bufs[] = {d_model, d_par, d_b, d_i, d_sum, d_cnt, d_knots};
  for (void* p : bufs)
    if (p) cudaFreeAsync(p, s);
  cudaStreamSynchronize(s);
  return rc;
}

int fx_eval_batch(int64_t n, const int32_t* counters, const double* knots, int32_t k_max, int64_t q,
                  const double* o_query, int32_t per_problem, double* theta, double* dtheta_do,
                  void* stream) {
  if (n < 0 || q < 0 || k_max < 1) return fail(FX_ERR_ARG, "bad sizes%s");
  if (n == 0 || q == 0) return FX_OK;
  if (!counters || !knots || !o_query) return fail(FX_ERR_ARG, "null pointer argument%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  eval_kernel<<<grid_for(n * q, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      n, counters, knots, k_max, q, o_query, per_problem, theta, dtheta_do);
  g_launches++;
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int fx_scaled_fit_batch(int64_t n, const int32_t* counters, const double* summary, const double* knots,
                        int32_t k_max, int64_t m, const double* o, const double* theta, const double* weight,
                        int32_t mode, double* D0, double* cost, int32_t* status, void* stream) {
  if (n < 0 || m < 1 || k_max < 1 || mode < 0 || mode > 2) return fail(FX_ERR_ARG, "bad sizes or mode%s");
  if (n == 0) return FX_OK;
  if (m > 2147483647LL) return fail(FX_ERR_ARG, "too many data points%s");
  if (!counters || !summary || !knots || !o || !theta || !D0 || !cost || !status)
    return fail(FX_ERR_ARG, "null pointer argument%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  scaled_fit_kernel<<<grid_for(n * 32, 128, di.sms), 128, 0, (cudaStream_t)stream>>>(
      n, counters, summary, knots, k_max, (int)m, o, theta, weight, mode, D0, cost, status);
  g_launches++;
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int fx_diffusivity_batch(int64_t n, const int32_t* model_id, const double* params, int64_t q,
                         const double* theta, double* out, void* stream) {
  if (n < 0 || q < 0) return fail(FX_ERR_ARG, "bad sizes%s");
  if (n == 0 || q == 0) return FX_OK;
  if (!model_id || !params || !theta || !out) return fail(FX_ERR_ARG, "null pointer argument%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  diffusivity_kernel<<<grid_for(n * q, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      n, model_id, params, q, theta, out);
  g_launches++;
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int fx_inverse_batch(int64_t batch, int64_t npts, const double* o, const double* theta,
                     double* D_at_knots, double* sorptivity, int32_t* status, double* do_dtheta,
                     double* Iodtheta, void* stream) {
  if (batch < 0 || npts < 0) return fail(FX_ERR_ARG, "bad sizes%s");
  if (batch == 0) return FX_OK;
  if (npts < 2) return fail(FX_ERR_ARG, "a profile needs at least 2 points%s");
  if (!o || !theta || !D_at_knots || !status)
    return fail(FX_ERR_ARG, "null pointer argument%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  int launches = 0;
  cudaError_t e = fx::inverse_launch(batch, npts, o, theta, D_at_knots, sorptivity, status,
                                     do_dtheta, Iodtheta, di.sms, (cudaStream_t)stream, &launches);
  g_launches += launches;
  if (e != cudaSuccess) return fail(FX_ERR_CUDA, "inverse_launch: %s", cudaGetErrorString(e));
  return FX_OK;
}

int fx_inverse_query_batch(int64_t batch, int64_t npts, const double* o, const double* theta,
                           const double* do_dtheta, const double* Iodtheta, int64_t q,
                           const double* theta_query, double* D_out, void* stream) {
  if (batch < 0 || npts < 2 || q < 0) return fail(FX_ERR_ARG, "bad sizes%s");
  if (batch == 0 || q == 0) return FX_OK;
  if (!o || !theta || !do_dtheta || !Iodtheta || !theta_query || !D_out)
    return fail(FX_ERR_ARG, "null pointer argument%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  fx::inverse_query_kernel<<<grid_for(batch * q, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      batch, npts, o, theta, do_dtheta, Iodtheta, q, theta_query, D_out);
  g_launches++;
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int fx_pchip_eval_batch(int64_t batch, int64_t npts, const double* x, const double* y, int64_t q,
                        const double* x_query, double* y_out, void* stream) {
  if (batch < 0 || npts < 2 || q < 0) return fail(FX_ERR_ARG, "bad sizes%s");
  if (batch == 0 || q == 0) return FX_OK;
  if (!x || !y || !x_query || !y_out) return fail(FX_ERR_ARG, "null pointer argument%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  fx::pchip_eval_kernel<<<grid_for(batch * q, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      batch, npts, x, y, q, x_query, y_out);
  g_launches++;
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int fx_pchip_query_batch(int64_t batch, int64_t npts, const double* x, const double* y, const double* slope,
                         const double* anti, int64_t q, const double* x_query, int32_t per_profile,
                         int32_t extrapolate, double* value, double* derivative, double* antiderivative,
                         void* stream) {
  if (batch < 0 || npts < 2 || q < 0) return fail(FX_ERR_ARG, "bad sizes%s");
  if (batch == 0 || q == 0) return FX_OK;
  if (!x || !y || !slope || !x_query) return fail(FX_ERR_ARG, "null pointer argument%s");
  if (antiderivative && !anti) return fail(FX_ERR_ARG, "the antiderivative needs the knot antiderivative values%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  fx::pchip_query_kernel<<<grid_for(batch * q, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      batch, npts, x, y, slope, anti, q, x_query, per_profile, extrapolate, value, derivative, antiderivative);
  g_launches++;
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int fx_unique_batch(int64_t batch, int64_t npts, const double* o, const double* theta, double* o_out,
                    double* theta_out, int32_t* index_out, int64_t* n_unique, void* stream) {
  if (batch < 0 || npts < 1) return fail(FX_ERR_ARG, "bad sizes%s");
  if (batch == 0) return FX_OK;
  if (npts > 2147483647LL) return fail(FX_ERR_ARG, "a profile exceeds 2^31-1 points%s");
  if (!o || !theta || !o_out || !theta_out || !n_unique) return fail(FX_ERR_ARG, "null pointer argument%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)npts;
  // stable LSD radix sort of (key(theta), index): equal thetas keep the caller's order, so the first of
  // every run is the first occurrence, which is what jnp.unique(return_index=True) returns
  size_t tmp_sort = 0, tmp_sel = 0;
  FX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const unsigned long long*)nullptr,
                                          (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, n, 0, 64, s));
  FX_CUDA(cub::DeviceSelect::UniqueByKey(nullptr, tmp_sel, (const unsigned long long*)nullptr, (const int*)nullptr,
                                         (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, n, s));
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t kb = up((size_t)n * sizeof(unsigned long long)), ib = up((size_t)n * sizeof(int));
  const size_t tb = up(tmp_sort > tmp_sel ? tmp_sort : tmp_sel);
  Scratch scr(s);
  FX_CUDA(cudaMallocAsync(&scr.p, 3 * kb + 3 * ib + tb + 256, s));
  char* base = static_cast<char*>(scr.p);
  unsigned long long* key0 = reinterpret_cast<unsigned long long*>(base);
  unsigned long long* key1 = reinterpret_cast<unsigned long long*>(base + kb);
  unsigned long long* key2 = reinterpret_cast<unsigned long long*>(base + 2 * kb);
  int* idx0 = reinterpret_cast<int*>(base + 3 * kb);
  int* idx1 = reinterpret_cast<int*>(base + 3 * kb + ib);
  int* idx2 = reinterpret_cast<int*>(base + 3 * kb + 2 * ib);
  void* tmp = base + 3 * kb + 3 * ib;
  int* count = reinterpret_cast<int*>(base + 3 * kb + 3 * ib + tb);
  const int g = grid_for(npts, 256, di.sms);
  for (int64_t p = 0; p < batch; p++) {
    const double* th = theta + p * npts;
    fx::unique_keys_kernel<<<g, 256, 0, s>>>(npts, th, key0, idx0);
    size_t t1 = tmp_sort, t2 = tmp_sel;
    FX_CUDA(cub::DeviceRadixSort::SortPairs(tmp, t1, key0, key1, idx0, idx1, n, 0, 64, s));
    FX_CUDA(cub::DeviceSelect::UniqueByKey(tmp, t2, key1, idx1, key2, idx2, count, n, s));
    fx::unique_gather_kernel<<<g, 256, 0, s>>>(count, idx2, o + p * npts, th, o_out + p * npts, theta_out + p * npts,
                                              index_out ? index_out + p * npts : nullptr,
                                              reinterpret_cast<long long*>(n_unique) + p);
    g_launches += 2;  // ours; the sort and the selection are CUB's
    FX_CUDA(cudaGetLastError());
  }
  return FX_OK;
}

int fx_measure_fp64_peak(double* flops_per_s, void* stream) {
  if (!flops_per_s) return fail(FX_ERR_ARG, "null pointer argument%s");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  double* sink = nullptr;
  FX_CUDA(cudaMallocAsync((void**)&sink, sizeof(double), s));
  cudaEvent_t e0, e1;
  FX_CUDA(cudaEventCreate(&e0));
  FX_CUDA(cudaEventCreate(&e1));
  const int block = 256, grid = di.sms * 8, iters = 4096;
  dfma_peak_kernel<<<grid, block, 0, s>>>(sink, 64, 1.0);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    FX_CUDA(cudaEventRecord(e0, s));
    dfma_peak_kernel<<<grid, block, 0, s>>>(sink, iters, 1.0);
    FX_CUDA(cudaEventRecord(e1, s));
    FX_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    FX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  g_launches += 6;
  double flops = 2.0 * 64.0 * (double)iters * (double)block * (double)grid;
  *flops_per_s = flops / ((double)best * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  FX_CUDA(cudaFreeAsync(sink, s));
  return FX_OK;
}

}  // extern "C"
