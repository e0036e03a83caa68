// One-CTA Cholesky of a packed lower-triangular matrix held in shared memory, blocked so that the CTA meets at
// two barriers per 8 columns instead of per column, with the forward substitution of one right-hand side riding
// along as an extra row.  Used per objective evaluation by the one-restart-per-CTA log-likelihood paths
// (gpUtils._nll, reference gpUtils.py:46-80: "kernel build + Cholesky per evaluation").
//
//   for each block column J (8 wide):
//     phase 1  left-looking update  A[i, J] -= L[i, :J0] . L[J, :J0]^T  for the rows i >= J0 and the rhs row:
//              independent dot products, one (row, column) item per thread pass, no read/write overlap
//              (reads touch finished columns < J0 only);
//     phase 2  every participating thread factors the 8x8 diagonal block redundantly in registers (36 doubles,
//              no communication), then owns one row below it -- or the rhs row -- and solves its 8 entries
//              against the block, also in registers.
// Packed rows make the thread-per-row accesses conflict-free: consecutive triangular numbers i(i+1)/2 are
// distinct modulo 16 (and 32) over any 16 (32) consecutive rows.
// All multiply-adds are explicit fma() so the routine rounds identically in translation units compiled with
// and without -fmad.
#pragma once
#include <cuda_runtime.h>

namespace apgp {

constexpr int CHOL_B = 8;

// -DAPGP_PROF: thread 0 accumulates SM cycles per phase into g_prof (tools/profile_optimizers.py --prof)
#ifdef APGP_PROF
__device__ long long g_prof[16];
#define PROF_T(var) long long var = clock64()
#define PROF_ADD(slot, t0) do { if (threadIdx.x == 0) g_prof[slot] += clock64() - (t0); } while (0)
#else
#define PROF_T(var) do {} while (0)
#define PROF_ADD(slot, t0) do {} while (0)
#endif

// K: packed lower triangle (row i at i(i+1)/2), overwritten by L (entries of the diagonal blocks are written back
//    only when store_diag).  r: [N] rhs in, z = L^{-1} r out.  diag: [N] out, L_ii.  *badflag is set to 1 (never
//    cleared here) when a pivot is not positive and finite.  NT = blockDim.x.  Ends with a barrier.
template <int NT>
__device__ __forceinline__ void chol_packed_blocked(double* __restrict__ K, double* __restrict__ r,
                                                    double* __restrict__ diag, int N, int* badflag, bool store_diag) {
  const int tid = threadIdx.x;
  for (int J0 = 0; J0 < N; J0 += CHOL_B) {
    const int bw = (N - J0 < CHOL_B) ? (N - J0) : CHOL_B;
    PROF_T(t_p1);
    if (J0 > 0) {
      // items = (row, column) dot products of length J0, one item per thread pass; 32-bit index math
      // (N < 256, so i(i+1)/2 < 2^15).  (Splitting short item lists over 2-8 lanes + shuffles measured slower.)
      const int nitems = (N - J0 + 1) * CHOL_B;                 // rows J0..N-1 and the rhs row, CHOL_B columns each
      for (int it = tid; it < nitems; it += NT) {
        const int i = J0 + (it >> 3), c = it & (CHOL_B - 1);
        if (c >= bw || (i < N && J0 + c > i)) continue;
        const double* Li = (i < N) ? K + i * (i + 1) / 2 : r;
        const double* Lc = K + (J0 + c) * (J0 + c + 1) / 2;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;             // J0 is a multiple of 8
#pragma unroll 2
        for (int k = 0; k < J0; k += 4) {
          s0 = fma(Li[k], Lc[k], s0); s1 = fma(Li[k + 1], Lc[k + 1], s1);
          s2 = fma(Li[k + 2], Lc[k + 2], s2); s3 = fma(Li[k + 3], Lc[k + 3], s3);
        }
        double* dst = (i < N) ? K + i * (i + 1) / 2 + J0 + c : r + J0 + c;
        *dst -= ((s0 + s1) + (s2 + s3));
      }
      PROF_ADD(4, t_p1);
      __syncthreads();
    }
    PROF_ADD(0, t_p1);
    PROF_T(t_p2);
    const int nbelow = N - J0 - bw;                              // rows under the diagonal block; + 1 rhs row (needs N < NT)
    double D[CHOL_B][CHOL_B];
    if (tid <= nbelow) {
      double inv[CHOL_B];
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c)
#pragma unroll
        for (int c2 = 0; c2 <= c; ++c2)
          D[c][c2] = (c < bw) ? K[(size_t)(J0 + c) * (J0 + c + 1) / 2 + J0 + c2] : ((c == c2) ? 1.0 : 0.0);
      PROF_ADD(8, t_p2);
      PROF_T(t_f);
      bool bad = false;
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c) {
        double dj = D[c][c];
        if (!(dj > 0.0 && dj < INFINITY)) { bad = true; dj = 1.0; }
        const double iv = rsqrt(dj);
        inv[c] = iv;
        D[c][c] = dj * iv;
#pragma unroll
        for (int c2 = c + 1; c2 < CHOL_B; ++c2) D[c2][c] *= iv;
#pragma unroll
        for (int c2 = c + 1; c2 < CHOL_B; ++c2)
#pragma unroll
          for (int c3 = c + 1; c3 <= c2; ++c3) D[c2][c3] = fma(-D[c2][c], D[c3][c], D[c2][c3]);
      }
      PROF_ADD(9, t_f);
      PROF_T(t_r);
      double* row = (tid < nbelow) ? K + (size_t)(J0 + bw + tid) * (J0 + bw + tid + 1) / 2 + J0 : r + J0;
      double x[CHOL_B];
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c) x[c] = (c < bw) ? row[c] : 0.0;
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c) {
        double v = x[c];
#pragma unroll
        for (int c2 = 0; c2 < c; ++c2) v = fma(-x[c2], D[c][c2], v);
        x[c] = v * inv[c];
      }
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c) if (c < bw) row[c] = x[c];
      PROF_ADD(10, t_r);
      if (tid == nbelow) {                                       // the rhs-row thread also publishes the pivots
        if (bad) *badflag = 1;
#pragma unroll
        for (int c = 0; c < CHOL_B; ++c) if (c < bw) diag[J0 + c] = D[c][c];
      }
    }
    PROF_ADD(5, t_p2);
    __syncthreads();
    PROF_ADD(1, t_p2);
    if (store_diag && tid == nbelow) {                           // everyone has read the unfactored block: write L back
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c)
#pragma unroll
        for (int c2 = 0; c2 <= c; ++c2)
          if (c < bw) K[(size_t)(J0 + c) * (J0 + c + 1) / 2 + J0 + c2] = D[c][c2];
    }
  }
  if (store_diag) __syncthreads();
}

}  // namespace apgp
