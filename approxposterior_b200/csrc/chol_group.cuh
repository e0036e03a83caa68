// Fused log-likelihood of ONE hyper-parameter vector for training sets too large for one CTA's shared memory
// (N > ~220): covariance build, blocked right-looking Cholesky with look-ahead, forward substitution and the
// log-likelihood reductions in a single kernel, executed by a thread-block CLUSTER of C CTAs (C = 1, 2, 4, 8, 16) that
// keeps the matrix in L2-resident global memory and meets only at two release/acquire flags per 64-column block.
//
// Replaces, per evaluation of gpUtils._nll (reference gpUtils.py:46-80: george.GP.log_likelihood = kernel build +
// Cholesky), the 3 Np/64 + 2 dependent launches of the tiled path in factor.cu (build_K, then chol_diag / chol_panel /
// chol_update per block column, then loglik_finish), whose cost was one launch gap plus one serial 64x64 pivot chain per
// block column regardless of the batch size.
//
// Work split.  The lower triangle of K is cut into 64x64 tiles; tile (i, j) belongs to CTA  (i(i+1)/2 + j) mod C  for
// the whole factorisation (owner computes: build, every trailing update, the panel solve, and -- for diagonal tiles --
// the 64x64 factorisation).  Only PANEL results cross CTAs, so step k needs two flags:
//   A_k  "diagonal block k factored":  D_k^{-1} and z_k are in global memory        (one writer, flagA)
//   B_k  "panel k solved":             every L_ik = A_ik D_k^{-T} is in global memory (C writers, cntB)
// Look-ahead: after B_k a CTA first updates its tiles of column k+1 (the owner of (k+1, k+1) then factors it at once
// and raises A_{k+1}), and only then applies step k to the rest of its trailing tiles -- that deferred work overlaps
// the next block's serial pivot chain on the owner.
//
// Inside a CTA (256 threads): two tile workers of 4 warps each (warp tile 32x32, DMMA.8x8x4 from padded shared tiles
// staged with cp.async.cg, so a worker's loads overlap the other's MMAs); the 64x64 diagonal factorisation uses all
// 256 threads (chol_small.cuh: 8-wide register-blocked, with r_k and the 64 unit vectors riding along as right-hand
// sides so that z_k = D^{-1} r_k and D^{-1} itself come out of the same pass).
//
// The flags only ever grow (epoch-based targets), so repeated evaluations by the same cluster -- the device
// optimisers call this once per objective evaluation -- need no reset.  All partial sums are added in block order:
// every CTA of the cluster returns the SAME bits, which the replicated optimiser state relies on.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "common.cuh"
#include "chol_small.cuh"

namespace apgp {

constexpr int CG_T = 64;                 // tile edge
constexpr int CG_LD = 68;                // padded shared leading dimension (conflict-free DMMA fragment loads)
constexpr int CG_THREADS = 256;
constexpr int CG_WT = 128;               // threads per tile worker
constexpr int CG_TILE = CG_T * CG_LD;    // doubles per staged tile
constexpr size_t CG_SMEM_DOUBLES = (size_t)4 * CG_TILE;          // 2 workers x (A, B)
constexpr int CG_DIAG_LDR = CG_T + 1;
// the diagonal factorisation aliases the tile buffers: packed block + (1 + 64) right-hand sides + pivots
static_assert((size_t)(CG_T * (CG_T + 1) / 2 + (CG_T + 1) * CG_DIAG_LDR + CG_T) <= CG_SMEM_DOUBLES, "diag scratch fits");

struct CholGroup {
  int N, Np, nb, d;
  int C, rank;                           // cluster size, this CTA's rank in it
  const double* X;                       // [N][d] training inputs (global, read-only for the kernel's lifetime)
  const double* y;                       // [N]
  double* K;                             // this restart's [Np][Np] row-major workspace (lower tiles used)
  double* Dinv;                          // [nb][64][64]
  double* r;                             // [Np]
  double* part;                          // [2][nb][4]: per block sum z^2, sum log diag, bad pivot (double-buffered by epoch)
  unsigned long long* flagA;             // sync words of this restart (monotone)
  unsigned long long* cntB;
};

// per-restart global workspace of a batch of R problems, carved from one allocation of cg_ws_bytes(Np, R)
struct GroupWs {
  double* K; double* Dinv; double* r; double* part; unsigned long long* flags;
  size_t sK, sDinv, sR, sPart;           // strides (in doubles) between restarts; flags: 2 words per restart
};
inline size_t cg_ws_bytes(int Np, int R) {
  const size_t nb = Np / CG_T;
  return (size_t)R * ((size_t)Np * Np + nb * CG_T * CG_T + Np + 2 * nb * 4) * sizeof(double) + ((size_t)R * 16 + 255) / 256 * 256 + 256;
}
inline GroupWs cg_ws_carve(void* ws_bytes, int Np, int R) {
  const size_t nb = Np / CG_T;
  GroupWs ws;
  char* base = static_cast<char*>(ws_bytes);
  ws.flags = reinterpret_cast<unsigned long long*>(base); base += ((size_t)R * 16 + 255) / 256 * 256;
  ws.K = reinterpret_cast<double*>(base); ws.sK = (size_t)Np * Np; base += (size_t)R * ws.sK * 8;
  ws.Dinv = reinterpret_cast<double*>(base); ws.sDinv = nb * CG_T * CG_T; base += (size_t)R * ws.sDinv * 8;
  ws.r = reinterpret_cast<double*>(base); ws.sR = Np; base += (size_t)R * ws.sR * 8;
  ws.part = reinterpret_cast<double*>(base); ws.sPart = 2 * nb * 4;
  return ws;
}
__device__ __forceinline__ CholGroup cg_make(const GroupWs& ws, int rr, int N, int Np, int d, int C, int rank,
                                             const double* X, const double* y) {
  CholGroup g;
  g.N = N; g.Np = Np; g.nb = Np / CG_T; g.d = d; g.C = C; g.rank = rank; g.X = X; g.y = y;
  g.K = ws.K + rr * ws.sK; g.Dinv = ws.Dinv + rr * ws.sDinv; g.r = ws.r + rr * ws.sR; g.part = ws.part + rr * ws.sPart;
  g.flagA = ws.flags + 2 * (size_t)rr; g.cntB = g.flagA + 1;
  return g;
}
__device__ __forceinline__ int cg_cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return (int)r; }

__device__ __forceinline__ int cg_owner(int i, int j, int C) { return (i * (i + 1) / 2 + j) % C; }

// ---- cross-CTA flags ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long cg_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// all threads of the CTA have finished their global writes (call with the whole CTA)
__device__ __forceinline__ void cg_publish_store(unsigned long long* p, unsigned long long v) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.u64 [%0], %1;\n" :: "l"(p), "l"(v) : "memory"); }
}
__device__ __forceinline__ void cg_publish_add(unsigned long long* p) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); asm volatile("red.release.gpu.global.add.u64 [%0], 1;\n" :: "l"(p) : "memory"); }
}
__device__ __forceinline__ void cg_wait_ge(const unsigned long long* p, unsigned long long v) {
  if (threadIdx.x == 0) { while (cg_ld_acquire(p) < v) { } }
  __syncthreads();
}

// ---- tile staging: 64x64 doubles, global row-major (ld) -> shared [r][CG_LD], 16-byte cp.async through L2 ------
__device__ __forceinline__ void cg_stage_tile(double* dst, const double* src, int ld, int wtid) {
#pragma unroll 4
  for (int e = wtid; e < CG_T * (CG_T / 2); e += CG_WT) {
    const int r = e >> 5, c2 = (e & 31) * 2;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                 :: "r"(smem_u32(dst + r * CG_LD + c2)), "l"(src + (size_t)r * ld + c2) : "memory");
  }
}
__device__ __forceinline__ void cg_stage_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

struct CgAcc { double v[4][4][2]; };

// acc = As(64x64, [m][k]) * Bs(64x64, [n][k])^T ; warp (wm, wn) of the worker owns rows wm*32.., cols wn*32..
__device__ __forceinline__ void cg_tile_mma(const double* As, const double* Bs, CgAcc& acc, int wm, int wn, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc.v[i][j][0] = 0.0; acc.v[i][j][1] = 0.0; }
  const double* ap = As + (wm * 32 + (lane >> 2)) * CG_LD + (lane & 3);
  const double* bp = Bs + (wn * 32 + (lane >> 2)) * CG_LD + (lane & 3);
#pragma unroll 4
  for (int k4 = 0; k4 < CG_T / 4; ++k4) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = ap[i * 8 * CG_LD + k4 * 4]; b[i] = bp[i * 8 * CG_LD + k4 * 4]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
  }
}

// Panel solve by substitution: one thread owns one row of a 64x64 panel tile and overwrites it with
//   x = a L^{-T}   (L: the factored diagonal block, row-major in shared memory with leading dimension CG_LD;
//                   inv: reciprocal pivots 1 / L_cc)
// in 8-wide column blocks, exactly the structure of chol_small.cuh: the contribution of the finished columns goes into 8
// independent accumulators (no dependent chain), the 8x8 triangular remainder is solved in registers.  The row lives in
// registers for the whole solve (64 doubles); L is read as warp-wide broadcasts.  ~2000 FMAs per row.
__device__ __forceinline__ void cg_trsm_row(double (&x)[CG_T], const double* __restrict__ Ls, const double* __restrict__ inv) {
#pragma unroll
  for (int J = 0; J < CG_T / 8; ++J) {
    double u[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) u[cc] = x[8 * J + cc];
#pragma unroll
    for (int c2 = 0; c2 < 8 * J; ++c2) {
      const double xv = x[c2];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) u[cc] = fma(-xv, Ls[(8 * J + cc) * CG_LD + c2], u[cc]);
    }
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      double v = u[cc];
#pragma unroll
      for (int c3 = 0; c3 < cc; ++c3) v = fma(-x[8 * J + c3], Ls[(8 * J + cc) * CG_LD + 8 * J + c3], v);
      x[8 * J + cc] = v * inv[8 * J + cc];
    }
  }
}

// One evaluation.  hyp (shared or global, [3 + d]): mean, amplitude, noise variance, 1/M_0 .. 1/M_{d-1}.
// sm: CG_SMEM_DOUBLES doubles of dynamic shared memory (16-byte aligned).  epoch: number of evaluations this cluster
// has completed on these flags (identical in every CTA).  Returns the log-likelihood (-inf when not positive definite
// or not finite) to every thread of every CTA of the cluster.  Ends with a CTA barrier.
template <bool FAST_PIVOT>
__device__ double chol_group_loglik(const CholGroup& g, const double* hyp, double* sm, unsigned long long epoch) {
  const int tid = threadIdx.x, w = tid >> 7, wtid = tid & 127, warp = (tid >> 5) & 3, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int nb = g.nb, Np = g.Np, C = g.C, me = g.rank, d = g.d;
  double* As = sm + (size_t)w * 2 * CG_TILE;
  double* Bs = As + CG_TILE;
  const unsigned long long baseA = epoch * (unsigned long long)(nb + 1);
  const unsigned long long baseB = epoch * (unsigned long long)(nb - 1) * C;       // one increment per CTA per panel
  double* part = g.part + (size_t)(epoch & 1ull) * nb * 4;
  const double mean = hyp[0], amp = hyp[1], noise = hyp[2];
  __shared__ int s_bad;

  // ---- build: every CTA forms its own tiles; the owner of (i, 0) also initialises r_i = y_i - mean --------------
  {
    int t = 0;
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j <= i; ++j) {
        if (cg_owner(i, j, C) != me) continue;
        if ((t++ & 1) != w) continue;
        // stage the two row blocks of X in the worker's buffers: rows i*64.. in As, rows j*64.. in Bs ([row][d])
        for (int e = wtid; e < CG_T * d; e += CG_WT) {
          const int rr = e / d, c = e - rr * d;
          const int gi = i * CG_T + rr, gj = j * CG_T + rr;
          As[e] = (gi < g.N) ? g.X[(size_t)gi * d + c] : 0.0;
          Bs[e] = (gj < g.N) ? g.X[(size_t)gj * d + c] : 0.0;
        }
        named_bar_sync(1 + w, CG_WT);
        double* Kt = g.K + (size_t)i * CG_T * Np + (size_t)j * CG_T;
        for (int e = wtid; e < CG_T * CG_T; e += CG_WT) {
          const int rr = e >> 6, cc = e & 63;
          const int gi = i * CG_T + rr, gj = j * CG_T + cc;
          double v;
          if (gi < g.N && gj < g.N) {
            double s = 0.0;
            for (int c = 0; c < d; ++c) { const double df = As[rr * d + c] - Bs[cc * d + c]; s += df * df * hyp[3 + c]; }
            v = amp * exp(-0.5 * s);
            if (gi == gj) v += noise;
          } else {
            v = (gi == gj) ? 1.0 : 0.0;
          }
          Kt[(size_t)rr * Np + cc] = v;
        }
        if (j == 0 && wtid < CG_T) {
          const int gi = i * CG_T + wtid;
          g.r[gi] = (gi < g.N) ? (g.y[gi] - mean) : 0.0;
        }
        named_bar_sync(1 + w, CG_WT);
      }
  }
  __syncthreads();

  for (int k = 0; k < nb; ++k) {
    // ---- (1) diagonal block k: factor, invert, z_k, partial sums -- by its owner, all 256 threads -------------------
    if (cg_owner(k, k, C) == me) {
      double* S = sm;                                   // packed lower triangle
      double* R = S + CG_T * (CG_T + 1) / 2;            // [CG_DIAG_LDR]: r_k -> z_k
      double* dg = R + CG_DIAG_LDR;                     // [64] pivots
      double* A = g.K + (size_t)k * CG_T * Np + (size_t)k * CG_T;
      double* rk = g.r + k * CG_T;
      if (tid == 0) s_bad = 0;
      for (int e = tid; e < CG_T * CG_T; e += CG_THREADS) {
        const int i = e >> 6, j = e & 63;
        if (j <= i) S[i * (i + 1) / 2 + j] = A[(size_t)i * Np + j];          // own tile: written by this CTA only
      }
      if (tid < CG_T) R[tid] = __ldcg(rk + tid);                              // r_k was updated by other CTAs' panels
      __syncthreads();
      chol_packed_blocked<CG_THREADS, FAST_PIVOT>(S, R, dg, CG_T, &s_bad, true, 1, CG_DIAG_LDR);
      // the factored block goes back row-major (the panel solves read it), its reciprocal pivots into the block's
      // slot of Dinv -- no explicit inverse: the panels are solved by substitution (cg_trsm_row)
      double* Dg = g.Dinv + (size_t)k * CG_T * CG_T;
      for (int e = tid; e < CG_T * CG_T; e += CG_THREADS) {
        const int i = e >> 6, j = e & 63;
        A[(size_t)i * Np + j] = (j <= i) ? S[i * (i + 1) / 2 + j] : 0.0;
      }
      if (tid < CG_T) { Dg[tid] = 1.0 / dg[tid]; rk[tid] = R[tid]; }
      if (tid < 32) {                                                          // block partials, fixed order
        double zz = R[tid] * R[tid] + R[tid + 32] * R[tid + 32];
        double lg = log(dg[tid]) + log(dg[tid + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { zz += __shfl_xor_sync(0xffffffffu, zz, o); lg += __shfl_xor_sync(0xffffffffu, lg, o); }
        if (tid == 0) { part[k * 4 + 0] = zz; part[k * 4 + 1] = lg; part[k * 4 + 2] = s_bad ? 1.0 : 0.0; }
      }
      if (C > 1) cg_publish_store(g.flagA, baseA + k + 1); else __syncthreads();
    }
    // ---- (2) deferred trailing work of step k-1 (columns >= k+1) overlaps the owner's pivot chain ------------------
    if (k > 0) {
      int t = 0;
      for (int j = k + 1; j < nb; ++j)
        for (int i = j; i < nb; ++i) {
          if (cg_owner(i, j, C) != me) continue;
          if ((t++ & 1) != w) continue;
          cg_stage_tile(As, g.K + (size_t)i * CG_T * Np + (size_t)(k - 1) * CG_T, Np, wtid);
          cg_stage_tile(Bs, g.K + (size_t)j * CG_T * Np + (size_t)(k - 1) * CG_T, Np, wtid);
          cg_stage_wait();
          named_bar_sync(1 + w, CG_WT);
          CgAcc acc;
          cg_tile_mma(As, Bs, acc, wm, wn, lane);
          double* Ct = g.K + (size_t)i * CG_T * Np + (size_t)j * CG_T;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int rr = wm * 32 + a * 8 + (lane >> 2), cc = wn * 32 + b * 8 + 2 * (lane & 3);
              double2* p = reinterpret_cast<double2*>(Ct + (size_t)rr * Np + cc);
              double2 o = *p; o.x -= acc.v[a][b][0]; o.y -= acc.v[a][b][1]; *p = o;
            }
          named_bar_sync(1 + w, CG_WT);
        }
      __syncthreads();
    }
    if (k == nb - 1) break;
    // ---- (3) panel k: L_ik = A_ik L_kk^{-T} by substitution (one thread per row), r_i -= L_ik z_k --------------------
    if (C > 1) cg_wait_ge(g.flagA, baseA + k + 1);
    {
      // every worker keeps its own copy of L_kk (row-major, ld CG_LD) in its first buffer; reciprocal pivots and z_k
      // in its second
      int mine = 0;
      for (int i = k + 1; i < nb; ++i) mine += (cg_owner(i, k, C) == me);
      if (mine > 0) {
        cg_stage_tile(As, g.K + (size_t)k * CG_T * Np + (size_t)k * CG_T, Np, wtid);
        if (wtid < CG_T) { Bs[wtid] = __ldcg(g.Dinv + (size_t)k * CG_T * CG_T + wtid); Bs[CG_T + wtid] = __ldcg(g.r + k * CG_T + wtid); }
        cg_stage_wait();
        named_bar_sync(1 + w, CG_WT);
        // the worker's two halves (64 threads each) take alternate tiles of the CTA's share
        const int half = wtid >> 6, row = wtid & 63;
        int t = 0;
        for (int i = k + 1; i < nb; ++i) {
          if (cg_owner(i, k, C) != me) continue;
          if ((t++ & 3) != 2 * w + half) continue;
          double* Arow = g.K + (size_t)(i * CG_T + row) * Np + (size_t)k * CG_T;
          double x[CG_T];
#pragma unroll
          for (int m = 0; m < CG_T / 2; ++m) {
            const double2 v = __ldcg(reinterpret_cast<const double2*>(Arow) + m);
            x[2 * m] = v.x; x[2 * m + 1] = v.y;
          }
          cg_trsm_row(x, As, Bs);
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int m = 0; m < CG_T / 2; ++m) {
            double2 v; v.x = x[2 * m]; v.y = x[2 * m + 1];
            reinterpret_cast<double2*>(Arow)[m] = v;
            s0 = fma(v.x, Bs[CG_T + 2 * m], s0); s1 = fma(v.y, Bs[CG_T + 2 * m + 1], s1);
          }
          double* ri = g.r + i * CG_T + row;
          *ri = __ldcg(ri) - (s0 + s1);
        }
        named_bar_sync(1 + w, CG_WT);                      // the worker's buffers are free again
      }
    }
    if (C > 1) { cg_publish_add(g.cntB); cg_wait_ge(g.cntB, baseB + (unsigned long long)(k + 1) * C); } else __syncthreads();
    // ---- (4) urgent part of step k: column k+1 (the next diagonal block and the next panel's inputs) --------------
    {
      int t = 0;
      const int j = k + 1;
      for (int i = j; i < nb; ++i) {
        if (cg_owner(i, j, C) != me) continue;
        if ((t++ & 1) != w) continue;
        cg_stage_tile(As, g.K + (size_t)i * CG_T * Np + (size_t)k * CG_T, Np, wtid);
        cg_stage_tile(Bs, g.K + (size_t)j * CG_T * Np + (size_t)k * CG_T, Np, wtid);
        cg_stage_wait();
        named_bar_sync(1 + w, CG_WT);
        CgAcc acc;
        cg_tile_mma(As, Bs, acc, wm, wn, lane);
        double* Ct = g.K + (size_t)i * CG_T * Np + (size_t)j * CG_T;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int rr = wm * 32 + a * 8 + (lane >> 2), cc = wn * 32 + b * 8 + 2 * (lane & 3);
            double2* p = reinterpret_cast<double2*>(Ct + (size_t)rr * Np + cc);
            double2 o = *p; o.x -= acc.v[a][b][0]; o.y -= acc.v[a][b][1]; *p = o;
          }
        named_bar_sync(1 + w, CG_WT);
      }
    }
    __syncthreads();
  }

  // ---- log-likelihood: block partials in block order (identical bits in every CTA) -----------------------------------
  if (C > 1) cg_wait_ge(g.flagA, baseA + nb);
  double zz = 0.0, lg = 0.0, bad = 0.0;
  for (int k = 0; k < nb; ++k) { zz += __ldcg(part + k * 4); lg += __ldcg(part + k * 4 + 1); bad += __ldcg(part + k * 4 + 2); }
  double ll = -0.5 * zz - lg - 0.5 * g.N * 1.8378770664093454836;
  if (bad != 0.0 || !(ll == ll) || !(fabs(ll) < INFINITY)) ll = -INFINITY;
  __syncthreads();
  return ll;
}

}  // namespace apgp
