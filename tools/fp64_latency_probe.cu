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
    else if (OP == 8) {        // one dependent DMMA.8x8x4 (accumulator chain)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(x), "+d"(y) : "d"(1e-9), "d"(1e-9));
    } else if (OP == 9) {      // DMMA whose A operand is the previous DMMA's result (operand chain)
      double c0 = 0.0, c1 = 0.0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(x), "d"(1e-9));
      x = c0 + 1.0;
    } else if (OP == 10) {     // DMMA -> 64-bit shuffle -> DMMA round (the panel solve's layout conversion)
      double c0 = 0.0, c1 = 0.0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(x), "d"(1e-9));
      x = __shfl_sync(0xffffffffu, c0, (threadIdx.x & 28) | ((threadIdx.x & 3) >> 1)) + 1.0;
    }
  }
  if (OP >= 11) {      // 8x8 register Cholesky with rsqrt pivots (the diagonal-block chain of chol_small.cuh); OP 12: + row solve
    double D[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int c2 = 0; c2 <= c; ++c2) D[c][c2] = (c == c2) ? 9.0 + x : 0.5 + 0.01 * (c + c2);
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
      double inv[8], xr[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) xr[c] = D[7][c] + 1.0;
      int bad = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        double dj = D[c][c];
        if (OP <= 12) { if (!(dj > 0.0 && dj < INFINITY)) dj = 1.0; }
        else if (!(dj > 0.0 && dj < INFINITY)) bad = c + 1;              // recorded, not substituted: off the chain
        const double iv = rsqrt(dj), sq = dj * iv;
        inv[c] = iv; D[c][c] = sq;
        if (OP >= 13 && c + 1 < 8) {                                      // the next pivot's inputs first
          D[c + 1][c] *= iv;
          D[c + 1][c + 1] = fma(-D[c + 1][c], D[c + 1][c], D[c + 1][c + 1]);
        }
#pragma unroll
        for (int c2 = c + 1 + (OP >= 13 ? 1 : 0); c2 < 8; ++c2) D[c2][c] *= iv;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
          for (int c3 = c + 1; c3 <= c2; ++c3)
            if (!(OP >= 13 && c2 == c + 1 && c3 == c + 1)) D[c2][c3] = fma(-D[c2][c], D[c3][c], D[c2][c3]);
        if (OP == 12 || OP == 14) {
          double v = xr[c];
#pragma unroll
          for (int c2 = 0; c2 < c; ++c2) v = fma(-xr[c2], D[c][c2], v);
          xr[c] = v * iv;
        }
      }
      // feed the result back so the next factorisation depends on this one, and keep the matrix positive definite
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int c2 = 0; c2 <= c; ++c2) D[c][c2] = (c == c2) ? 9.0 + 1e-3 * D[c][c2] + ((OP == 12 || OP == 14) ? 1e-9 * xr[c] : 0.0) + (bad ? 1.0 : 0.0) : 0.5 + 1e-3 * D[c][c2];
    }
    x = D[7][7] + D[3][1];
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA dependent", "rsqrt + DADD", "1/x + DADD", "sqrt + DADD", "STS+LDS round trip + DADD", "__syncthreads (256 thr) + DADD", "log + DADD", "exp + DADD", "DMMA.8x8x4 accumulator chain", "DMMA -> DADD -> DMMA (operand chain)", "DMMA -> SHFL.64 -> DADD -> DMMA", "8x8 register Cholesky (rsqrt pivots)", "8x8 register Cholesky + one row solve", "8x8 Cholesky, next pivot first, no select", "8x8 Cholesky, next pivot first + row solve"};
  const int n = 4096;
  for (int threads : {32, 256}) {
    for (int op = 0; op < 15; ++op) {
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
          case 8: probe<8><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 9: probe<9><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 10: probe<10><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 11: probe<11><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 12: probe<12><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 13: probe<13><<<1, threads>>>(out, cyc, 1.3, n); break;
          case 14: probe<14><<<1, threads>>>(out, cyc, 1.3, n); break;
        }
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("threads=%3d  %-32s %7.1f cycles per iteration\n", threads, names[op], (double)c / n);
    }
  }
  return 0;
}
