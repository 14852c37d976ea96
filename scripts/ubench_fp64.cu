// ubench_fp64.cu -- FP64 pipe micro-benchmarks on sm_100a (developer tool, not product code).
// Measures, per SM sub-partition, the dependent-issue latency and the throughput of DFMA,
// fp64 division and reciprocal as a function of resident warps and ILP, so that the
// occupancy / ILP targets of shoot_bisect_kernel are chosen from measurements.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_fp64 ubench_fp64.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

template <int ILP>
__global__ void dfma_chain(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 1e-9 + k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += x[k];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void ddiv_chain(double* out, int iters, double a) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = 1.0 + threadIdx.x * 1e-9 + k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < ILP; k++) x[k] = a / x[k];
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += x[k];
  if (s == 123.456) out[0] = s;
}

// mixed: one DFMA + n_int integer ops per trip, to see whether ALU instructions issue for free
template <int NINT>
__global__ void mixed_chain(double* out, int iters, double a, double b, int c) {
  double x = threadIdx.x * 1e-9, y = 1.0 + x;
  int z = threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
      x = fma(x, a, b);
      y = fma(y, a, b);
#pragma unroll
      for (int k = 0; k < NINT; k++) z = (z ^ c) + (z >> 3);
    }
  }
  if (x + y + z == 123.456) out[0] = x;
}

template <class F>
float time_it(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ghz = clk_khz * 1e-6;
  printf("%s SMs %d clock %.3f GHz (nominal; cycles below assume it)\n", p.name, p.multiProcessorCount, ghz);
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  int nsm = p.multiProcessorCount;
  // warps per SM sub-partition: block = 128*w threads, one block per SM => w warps per SMSP
  printf("\nDFMA: cycles per dependent DFMA per warp (lat) and DFMA lanes/clk/SM (thr)\n");
  printf("%6s %4s %10s %10s\n", "w/SMSP", "ILP", "cyc/dfma", "lanes/clk");
  for (int w : {1, 2, 3, 4, 6, 8}) {
    auto run = [&](auto kern, int ilp) {
      float ms = time_it([&] { kern<<<nsm, 128 * w>>>(out, iters, 0.999, 1e-3); });
      double cycles = ms * 1e-3 * ghz * 1e9;
      double n_per_warp = (double)iters * 16 * ilp;
      printf("%6d %4d %10.2f %10.1f\n", w, ilp, cycles / (iters * 16.0), n_per_warp * 4 * w * 32 / cycles);
    };
    run(dfma_chain<1>, 1);
    run(dfma_chain<2>, 2);
    run(dfma_chain<4>, 4);
  }
  printf("\nfp64 division a/x: cycles per dependent division per warp; divisions/clk/SM\n");
  for (int w : {1, 2, 4, 8}) {
    auto run = [&](auto kern, int ilp) {
      float ms = time_it([&] { kern<<<nsm, 128 * w>>>(out, iters / 4, 1.2345); });
      double cycles = ms * 1e-3 * ghz * 1e9;
      double n_per_warp = (double)(iters / 4) * 4 * ilp;
      printf("%6d %4d %10.2f %10.2f\n", w, ilp, cycles / (iters / 4 * 4.0), n_per_warp * 4 * w * 32 / cycles);
    };
    run(ddiv_chain<1>, 1);
    run(ddiv_chain<2>, 2);
    run(ddiv_chain<4>, 4);
  }
  printf("\nmixed: 2 DFMA + 2*NINT ALU ops per trip, 4 warps/SMSP: cycles per trip per SMSP\n");
  {
    auto run = [&](auto kern, int nint) {
      float ms = time_it([&] { kern<<<nsm, 512>>>(out, iters, 0.999, 1e-3, 12345); });
      double cycles = ms * 1e-3 * ghz * 1e9;
      printf("NINT %d: %.2f cycles per (2 DFMA + %d ALU) x 4 warps\n", nint, cycles / (iters * 16.0), 2 * nint);
    };
    run(mixed_chain<0>, 0);
    run(mixed_chain<1>, 1);
    run(mixed_chain<2>, 2);
    run(mixed_chain<4>, 4);
  }
  return 0;
}
