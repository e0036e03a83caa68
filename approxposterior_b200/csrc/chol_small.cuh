// One-CTA Cholesky of a packed lower-triangular matrix held in shared memory, blocked so that the CTA meets at
// two barriers per 8 columns instead of per column, with the forward substitution of one right-hand side riding
// along as an extra row.  Used per objective evaluation by the one-restart-per-CTA log-likelihood paths
// (gpUtils._nll, reference gpUtils.py:46-80: "kernel build + Cholesky per evaluation").
//
//   for each block column J (8 wide):
//     phase 1  left-looking update  A[i, J] -= L[i, :J0] . L[J, :J0]^T  for the rows i >= J0 and the rhs rows:
//              independent dot products, one (row, column) item per thread pass, no read/write overlap
//              (reads touch finished columns < J0 only).  It is applied in two instalments: the contribution of
//              the columns finished before block J-1 is subtracted by the warps that would otherwise idle during
//              phase 2 of block J-1 (look-ahead); only the last 8 columns' contribution is left for everyone at
//              the top of block J;
//     phase 2  every participating thread factors the 8x8 diagonal block redundantly in registers (36 doubles,
//              no communication), then owns one row below it -- or a rhs row -- and solves its 8 entries
//              against the block, also in registers.
// Packed rows make the thread-per-row accesses conflict-free: consecutive triangular numbers i(i+1)/2 are
// distinct modulo 16 (and 32) over any 16 (32) consecutive rows.
// All multiply-adds are explicit fma() so the routine rounds identically in translation units compiled with
// and without -fmad.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace apgp {

constexpr int CHOL_B = 8;

// -DAPGP_PROF: thread 0 accumulates SM cycles per phase into g_prof (tools/profile_optimizers.py --prof)
#if defined(APGP_PROF) && !defined(APGP_PROF_CG)
__device__ long long g_prof[16];
#define PROF_T(var) long long var = clock64()
#define PROF_ADD(slot, t0) do { if (threadIdx.x == 0) g_prof[slot] += clock64() - (t0); } while (0)
#else
#define PROF_T(var) do {} while (0)
#define PROF_ADD(slot, t0) do {} while (0)
#endif

// K: packed lower triangle (row i at i(i+1)/2), overwritten by L (entries of the diagonal blocks are written back
//    only when store_diag).  r: nrhs right-hand sides, row q at r + q*ldr, [N] each: rhs in, L^{-1} rhs out (with
//    the N unit vectors as right-hand sides the rows come back as the columns of L^{-1}).  diag: [N] out, L_ii.
//    *badflag receives the 1-based index of the first pivot that is not positive and finite (never cleared here).  NT = blockDim.x >= N + nrhs,
//    nrhs >= 1.  Ends with a barrier.
//    FAST_PIVOT: pivots by rsqrt (<= 1 ulp, ~75 cycles) instead of IEEE sqrt + divide (~190 cycles): used by the
//    optimiser objectives, where the pivot chain is the critical path; the factorisation behind predict keeps the
//    correctly rounded pair (its errors are amplified by cond(K) into alpha and L^{-1}).
// Left-looking update of block column JB on the FP64 tensor pipe:
//   A[i][JB + c] -= sum_{k0 <= k < k1} L[i][k] L[JB + c][k]     for the rows i >= JB and the rhs rows, c < bwB
// as one DMMA.8x8x4 chain per 8 rows (two accumulator pairs, so consecutive k4 steps do not wait on each other): an
// m8 tile of rows times the 8 columns of the block, k running over the finished columns.  The per-(row, column) FMA
// dot products this replaces were 14.4 k of the 36 k cycles of a factorisation at N = 70 (profiles/r01_optimizers.md).
// warp / nwarps: the calling warps' index and count; k0, k1 multiples of 8.  Packed rows: row i at K + i (i + 1) / 2;
// rhs row q at r + q * ldr.  Entries above the diagonal inside the block (JB + c > i) are not stored and not written.
__device__ __forceinline__ void chol_update_dmma(double* __restrict__ K, double* __restrict__ r, int N, int nrhs, int ldr,
                                                 int JB, int bwB, int k0, int k1, int warp, int nwarps) {
  const int lane = threadIdx.x & 31, rsub = lane >> 2, kq = lane & 3;
  const int nmat = N - JB, nrows = nmat + nrhs, ntile = (nrows + 7) >> 3;
  const bool bvalid = rsub < bwB;
  const double* Lc = K + (JB + rsub) * (JB + rsub + 1) / 2;
  for (int mt = warp; mt < ntile; mt += nwarps) {
    const int vr = mt * 8 + rsub;
    double* Li = nullptr;
    if (vr < nmat) Li = K + (JB + vr) * (JB + vr + 1) / 2;
    else if (vr < nrows) Li = r + (vr - nmat) * ldr;
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
    for (int k = k0; k < k1; k += 8) {
      const double a0 = Li ? Li[k + kq] : 0.0, b0 = bvalid ? Lc[k + kq] : 0.0;
      const double a1 = Li ? Li[k + 4 + kq] : 0.0, b1 = bvalid ? Lc[k + 4 + kq] : 0.0;
      dmma884(c0, c1, a0, b0);
      dmma884(e0, e1, a1, b1);
    }
    c0 += e0; c1 += e1;
    if (Li) {
      const int col = 2 * kq;
      const bool tri = vr < nmat;                      // matrix rows store only columns <= their own index
      if (col < bwB && !(tri && col > vr)) Li[JB + col] -= c0;
      if (col + 1 < bwB && !(tri && col + 1 > vr)) Li[JB + col + 1] -= c1;
    }
  }
}

// The 8x8 diagonal block in registers: D (lower triangle) -> L, inv = 1 / pivots, and ONE row x solved against the block
// on the way (x[c] = (x[c] - sum_{c2<c} x[c2] L[c][c2]) / L[c][c], as soon as pivot c is known: the factorisation is a
// dependent chain that leaves most issue slots empty).  The chain is  pivot -> scale L[c+1][c] -> update D[c+1][c+1] ->
// next pivot: those two operations are issued FIRST after every pivot, and a pivot that is not positive and finite is
// only recorded (bad = 1-based index of the first one), not replaced, so no compare/select sits on the chain either --
// the NaNs stay inside a factorisation that is reported as failed.  Measured (tools/fp64_latency_probe.cu, one warp):
// 1277 cycles per block in program order with the select, 922 in this order, 1036 with the row solve riding along.
// Every entry sees the same operations in the same order as before: results are bit-identical.
template <bool FAST_PIVOT>
__device__ __forceinline__ void chol_block8(double (&D)[CHOL_B][CHOL_B], double (&inv)[CHOL_B], double (&x)[CHOL_B],
                                            int& bad, int J0) {
#pragma unroll
  for (int c = 0; c < CHOL_B; ++c) {
    const double dj = D[c][c];
    if (!(dj > 0.0 && dj < INFINITY) && !bad) bad = J0 + c + 1;
    double iv, sq;
    if (FAST_PIVOT) { iv = rsqrt(dj); sq = dj * iv; }
    else { sq = sqrt(dj); iv = 1.0 / sq; }
    inv[c] = iv;
    D[c][c] = sq;
    if (c + 1 < CHOL_B) {                                      // the next pivot's inputs first
      D[c + 1][c] *= iv;
      D[c + 1][c + 1] = fma(-D[c + 1][c], D[c + 1][c], D[c + 1][c + 1]);
    }
#pragma unroll
    for (int c2 = c + 2; c2 < CHOL_B; ++c2) D[c2][c] *= iv;
#pragma unroll
    for (int c2 = c + 2; c2 < CHOL_B; ++c2)
#pragma unroll
      for (int c3 = c + 1; c3 <= c2; ++c3) D[c2][c3] = fma(-D[c2][c], D[c3][c], D[c2][c3]);
    double v = x[c];
#pragma unroll
    for (int c2 = 0; c2 < c; ++c2) v = fma(-x[c2], D[c][c2], v);
    x[c] = v * iv;
  }
}

template <int NT, int BAR>
__device__ __forceinline__ void chol_sync() {
  if (BAR == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;\n" :: "r"(BAR), "r"(NT) : "memory");
}

template <int NT, bool FAST_PIVOT, int BAR = 0>
__device__ __forceinline__ void chol_packed_blocked_classic(double* __restrict__ K, double* __restrict__ r,
                                                    double* __restrict__ diag, int N, int* badflag, bool store_diag,
                                                    int nrhs, int ldr) {
  const int tid = threadIdx.x;
  for (int J0 = 0; J0 < N; J0 += CHOL_B) {
    const int bw = (N - J0 < CHOL_B) ? (N - J0) : CHOL_B;
    PROF_T(t_p1);
    if (J0 > 0) {
      chol_update_dmma(K, r, N, nrhs, ldr, J0, bw, 0, J0, tid >> 5, NT / 32);
      PROF_ADD(4, t_p1);
      chol_sync<NT, BAR>();
    }
    PROF_ADD(0, t_p1);
    PROF_T(t_p2);
    const int nbelow = N - J0 - bw;                              // rows under the diagonal block; then the rhs rows
    double D[CHOL_B][CHOL_B];
    if (tid < nbelow + nrhs) {
      double inv[CHOL_B];
      {
        const double* Dr = K + J0 * (J0 + 1) / 2 + J0;          // row J0 + c of the block starts c (J0 + 1) + c (c - 1) / 2 further
#pragma unroll
        for (int c = 0; c < CHOL_B; ++c) {
#pragma unroll
          for (int c2 = 0; c2 <= c; ++c2) D[c][c2] = (c < bw) ? Dr[c2] : ((c == c2) ? 1.0 : 0.0);
          Dr += J0 + c + 1;
        }
      }
      PROF_ADD(8, t_p2);
      PROF_T(t_f);
      int bad = 0;                                                // 1-based index of the first failing pivot
      double* row = (tid < nbelow) ? K + (J0 + bw + tid) * (J0 + bw + tid + 1) / 2 + J0
                                   : r + (tid - nbelow) * ldr + J0;
      double x[CHOL_B];
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c) x[c] = (c < bw) ? row[c] : 0.0;
      chol_block8<FAST_PIVOT>(D, inv, x, bad, J0);              // factor the block, solve this thread's row on the way
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c) if (c < bw) row[c] = x[c];
      PROF_ADD(9, t_f);
      if (tid == nbelow) {                                       // the first rhs-row thread also publishes the pivots
        if (bad && *badflag == 0) *badflag = bad;
#pragma unroll
        for (int c = 0; c < CHOL_B; ++c) if (c < bw) diag[J0 + c] = D[c][c];
      }
    }
    PROF_ADD(5, t_p2);
    chol_sync<NT, BAR>();
    PROF_ADD(1, t_p2);
    if (store_diag && tid == nbelow) {                           // everyone has read the unfactored block: write L back
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c)
#pragma unroll
        for (int c2 = 0; c2 <= c; ++c2)
          if (c < bw) K[(size_t)(J0 + c) * (J0 + c + 1) / 2 + J0 + c2] = D[c][c2];
    }
  }
  if (store_diag) chol_sync<NT, BAR>();
}


template <int NT, bool FAST_PIVOT>
__device__ __forceinline__ void chol_packed_blocked_lookahead(double* __restrict__ K, double* __restrict__ r,
                                                    double* __restrict__ diag, int N, int* badflag, bool store_diag,
                                                    int nrhs, int ldr) {
  const int tid = threadIdx.x;
  // A[i][JB + c] -= sum_{k0 <= k < k1} L[i][k] L[JB + c][k]  for the rows i >= JB and the rhs rows, columns c < bwB of
  // block JB, by the threads t0 <= tid < t0 + nthr.  One (row, column) item per thread pass; 32-bit index math
  // (N < 256, so i(i+1)/2 < 2^15); k0, k1 multiples of 8.  (Splitting short item lists over 2-8 lanes + shuffles
  // measured slower.)
  auto update = [&](int JB, int bwB, int k0, int k1, int t0, int nthr) {
    chol_update_dmma(K, r, N, nrhs, ldr, JB, bwB, k0, k1, (tid - t0) >> 5, nthr / 32);
  };
  for (int J0 = 0; J0 < N; J0 += CHOL_B) {
    const int bw = (N - J0 < CHOL_B) ? (N - J0) : CHOL_B;
    PROF_T(t_p1);
    if (J0 > 0) {                                                // everyone: the last 8 finished columns -> this block
      update(J0, bw, J0 - CHOL_B, J0, 0, NT);
      PROF_ADD(4, t_p1);
      __syncthreads();
    }
    PROF_ADD(0, t_p1);
    PROF_T(t_p2);
    const int nbelow = N - J0 - bw;                              // rows under the diagonal block; then the rhs rows
    // look-ahead: block J+1 receives the contribution of the columns finished BEFORE this block (k < J0) from the
    // warps that have no row in phase 2; it touches columns >= J0 + 8 only, phase 2 columns J0 .. J0 + 7
    const int J1 = J0 + CHOL_B;
    const int bw1 = (N - J1 < CHOL_B) ? (N - J1) : CHOL_B;
    const int rowT = (nbelow + nrhs + 31) & ~31;                 // first thread of the first idle warp
    const bool ahead = J1 < N && J0 > 0;                         // callers guarantee idle warps (N + nrhs <= NT - 160)
    double D[CHOL_B][CHOL_B];
    if (ahead && tid >= rowT) update(J1, bw1, 0, J0, rowT, NT - rowT);
    if (tid < nbelow + nrhs) {
      double inv[CHOL_B];
      {
        const double* Dr = K + J0 * (J0 + 1) / 2 + J0;          // row J0 + c of the block starts c (J0 + 1) + c (c - 1) / 2 further
#pragma unroll
        for (int c = 0; c < CHOL_B; ++c) {
#pragma unroll
          for (int c2 = 0; c2 <= c; ++c2) D[c][c2] = (c < bw) ? Dr[c2] : ((c == c2) ? 1.0 : 0.0);
          Dr += J0 + c + 1;
        }
      }
      PROF_ADD(8, t_p2);
      PROF_T(t_f);
      int bad = 0;                                                // 1-based index of the first failing pivot
      double* row = (tid < nbelow) ? K + (J0 + bw + tid) * (J0 + bw + tid + 1) / 2 + J0
                                   : r + (tid - nbelow) * ldr + J0;
      double x[CHOL_B];
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c) x[c] = (c < bw) ? row[c] : 0.0;
      chol_block8<FAST_PIVOT>(D, inv, x, bad, J0);              // factor the block, solve this thread's row on the way
#pragma unroll
      for (int c = 0; c < CHOL_B; ++c) if (c < bw) row[c] = x[c];
      PROF_ADD(9, t_f);
      if (tid == nbelow) {                                       // the first rhs-row thread also publishes the pivots
        if (bad && *badflag == 0) *badflag = bad;
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


// Dispatcher.  Small systems (N + nrhs <= 96 rows: at least five of the eight warps have no row in phase 2) take the
// look-ahead form, measured -6 % / -10 % / -16 % per factorisation at N = 50 / 70 / 90; larger ones keep the single
// full-length update at the top of each block (splitting it cost +15 % at N = 200 and +6 % in the 64x64 diagonal
// kernel with its 65 right-hand sides).
template <int NT, bool FAST_PIVOT = true>
__device__ __forceinline__ void chol_packed_blocked(double* __restrict__ K, double* __restrict__ r,
                                                    double* __restrict__ diag, int N, int* badflag, bool store_diag,
                                                    int nrhs = 1, int ldr = 0) {
  if (N + nrhs <= NT - 160) chol_packed_blocked_lookahead<NT, FAST_PIVOT>(K, r, diag, N, badflag, store_diag, nrhs, ldr);
  else chol_packed_blocked_classic<NT, FAST_PIVOT>(K, r, diag, N, badflag, store_diag, nrhs, ldr);
}

// The classic form run by the first NT threads of a larger CTA (threadIdx.x < NT), meeting at named barrier BAR: the
// rest of the CTA keeps working (chol_group.cuh factors a 64x64 diagonal block on one half of the CTA while the other
// half runs trailing updates).
template <int NT, bool FAST_PIVOT, int BAR>
__device__ __forceinline__ void chol_packed_blocked_sub(double* __restrict__ K, double* __restrict__ r,
                                                        double* __restrict__ diag, int N, int* badflag, bool store_diag,
                                                        int nrhs = 1, int ldr = 0) {
  chol_packed_blocked_classic<NT, FAST_PIVOT, BAR>(K, r, diag, N, badflag, store_diag, nrhs, ldr);
}

}  // namespace apgp
