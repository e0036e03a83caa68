// Dependent-issue latencies that bound the one-CTA Cholesky (chol_small.cuh): DFMA, rsqrt, 1/x, sqrt, LDS round trip,
// __syncthreads with 256 threads.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency_probe fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void probe(double* out, long long* cyc, double seed, int n) {
  __shared__ double sm[256];
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  sm[threadIdx.x] = x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = fma(x, y, 1e-9);
    else if (OP == 1) x = rsqrt(x) + 1.5;
    else if (OP == 2) x = 1.0 / x + 1.5;
    else if (OP == 3) x = sqrt(x) + 1.5;
    else if (OP == 4) { sm[threadIdx.x] = x; x = sm[(threadIdx.x + 1) & 255] + 1e-9; }
    else if (OP == 5) { __syncthreads(); x += 1e-9; }
    else if (OP == 6) x = log(x) + 3.0;
    else if (OP == 7) x = exp(-x) + 0.5;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA dependent", "rsqrt + DADD", "1/x + DADD", "sqrt + DADD", "STS+LDS round trip + DADD", "__syncthreads (256 thr) + DADD", "log + DADD", "exp + DADD"};
  const int n = 4096;
  for (int threads : {32, 256}) {
    for (int op = 0; op < 8; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: probe<0><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 1: probe<1><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 2: probe<2><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 3: probe<3><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 4: probe<4><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 5: probe<5><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 6: probe<6><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 7: probe<7><<<1, threads>>>(out, cyc, 1.3, n); break;
        }
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("threads=%3d  %-32s %7.1f cycles per iteration\n", threads, names[op], (double)c / n);
    }
  }
  return 0;
}
