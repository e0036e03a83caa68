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
static_assert((size_t)(CG_T * (CG_T + 1) / 2 + (CG_T + 1) * CG_DIAG_LDR + CG_T) <= (size_t)2 * CG_TILE, "diag scratch fits in ONE worker's buffers");

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

// ---- helpers of the v2 schedule ---------------------------------------------------------------------------------
// worker-scope publication: the 128 threads of one worker have finished their global writes
__device__ __forceinline__ void cg_publish_store_w(unsigned long long* p, unsigned long long v, int bar, int wtid) {
  named_bar_sync(bar, CG_WT);
  if (wtid == 0) { __threadfence(); asm volatile("st.release.gpu.global.u64 [%0], %1;\n" :: "l"(p), "l"(v) : "memory"); }
}

// tile (i, j) from its row-major index in the lower triangle, L = i (i + 1) / 2 + j
__device__ __forceinline__ void cg_unrank(int L, int& i, int& j) {
  int r = (int)((sqrtf(8.0f * (float)L + 1.0f) - 1.0f) * 0.5f);
  while ((r + 1) * (r + 2) / 2 <= L) ++r;
  while (r * (r + 1) / 2 > L) --r;
  i = r; j = L - r * (r + 1) / 2;
}

// C(i, j) -= A(i, kcol) * A(j, kcol)^T for one 64x64 tile, by one worker.  The C fragments are fetched into registers
// while the cp.async operand copies are in flight, so the read-modify-write adds no exposed latency after the MMAs.
__device__ __forceinline__ void cg_update_tile(const CholGroup& g, int i, int j, int kcol, double* As, double* Bs, int w,
                                               int wtid, int wm, int wn, int lane) {
  const int Np = g.Np;
  cg_stage_tile(As, g.K + (size_t)i * CG_T * Np + (size_t)kcol * CG_T, Np, wtid);
  cg_stage_tile(Bs, g.K + (size_t)j * CG_T * Np + (size_t)kcol * CG_T, Np, wtid);
  double* Ct = g.K + (size_t)i * CG_T * Np + (size_t)j * CG_T;
  double2 c[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int rr = wm * 32 + a * 8 + (lane >> 2), cc = wn * 32 + b * 8 + 2 * (lane & 3);
      c[a][b] = *reinterpret_cast<const double2*>(Ct + (size_t)rr * Np + cc);
    }
  cg_stage_wait();
  named_bar_sync(1 + w, CG_WT);
  CgAcc acc;
  cg_tile_mma(As, Bs, acc, wm, wn, lane);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int rr = wm * 32 + a * 8 + (lane >> 2), cc = wn * 32 + b * 8 + 2 * (lane & 3);
      double2 o = c[a][b];
      o.x -= acc.v[a][b][0]; o.y -= acc.v[a][b][1];
      *reinterpret_cast<double2*>(Ct + (size_t)rr * Np + cc) = o;
    }
  named_bar_sync(1 + w, CG_WT);                            // the worker's buffers are free again
}

// Explicit inverse of a 64x64 lower-triangular factor held packed in shared memory, by the 128 threads of worker 0
// (named barrier `bar`): column j of the inverse is the forward substitution of e_j, carried by the thread pair
// (2 j, 2 j + 1) which splits every dot product.  Xc: [64][CG_DIAG_LDR] scratch, column j of the inverse in row j.
__device__ __forceinline__ void cg_tri_inverse64(const double* __restrict__ S, const double* __restrict__ dg,
                                                 double* __restrict__ Xc, int wtid, int bar) {
  const int j = wtid >> 1, h = wtid & 1;
  double* x = Xc + j * CG_DIAG_LDR;
  // every lane runs all 64 steps (the pair shuffle and the warp barrier need the whole warp): lanes whose column
  // starts below row i just idle through the early steps
  for (int i = 0; i < CG_T; ++i) {
    const double* Li = S + i * (i + 1) / 2;
    double s0 = 0.0, s1 = 0.0;
    if (i > j) {
      int c = j + h;
      for (; c + 2 < i; c += 4) { s0 = fma(Li[c], x[c], s0); s1 = fma(Li[c + 2], x[c + 2], s1); }
      for (; c < i; c += 2) s0 = fma(Li[c], x[c], s0);
    }
    double s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (i >= j && h == 0) x[i] = (((i == j) ? 1.0 : 0.0) - s) / dg[i];
    __syncwarp();                                          // the pair shares x through shared memory
  }
  named_bar_sync(bar, CG_WT);
}

// One evaluation.  hyp (shared or global, [3 + d]): mean, amplitude, noise variance, 1/M_0 .. 1/M_{d-1}.
// sm: CG_SMEM_DOUBLES doubles of dynamic shared memory (16-byte aligned).  epoch: number of evaluations this cluster
// has completed on these flags (identical in every CTA).  Returns the log-likelihood (-inf when not positive definite
// or not finite) to every thread of every CTA of the cluster.  Ends with a CTA barrier.
//
// Schedule of iteration k inside a CTA (two workers of 128 threads, w0 and w1):
//   (1+2) the owner's w0 factors diagonal block k (nrhs = 1: z_k rides along), inverts it (cg_tri_inverse64) and
//         raises A_k, while w1 -- and w0 as soon as it is done, and both workers of every other CTA -- pull the
//         DEFERRED trailing tiles of step k-1 (columns >= k+1) from a shared-memory task counter;
//   (3)   panel k after A_k;  B_k;
//   (4)   urgent tiles of step k: column k+1 only (so that block k+1 can be factored first thing in iteration k+1).
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
  const int ntiles = nb * (nb + 1) / 2;
  __shared__ int s_bad, s_next, s_task[2];

  // ---- build: every CTA forms its own tiles; the owner of (i, 0) also initialises r_i = y_i - mean --------------
  {
    int t = 0;
    for (int L = me; L < ntiles; L += C, ++t) {
      if ((t & 1) != w) continue;
      int i, j;
      cg_unrank(L, i, j);
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
    const bool own_diag = cg_owner(k, k, C) == me;
    // task counter of the deferred phase: first tile index of row k+1 that belongs to this CTA
    if (tid == 0) {
      const int L0 = (k + 1) * (k + 2) / 2;
      s_next = (L0 <= me) ? 0 : (L0 - me + C - 1) / C;
    }
    __syncthreads();
    // ---- (1) diagonal block k on worker 0 of its owner ----------------------------------------------------------------
    if (own_diag && w == 0) {
      double* S = sm;                                   // packed lower triangle (worker 0's buffers)
      double* R = S + CG_T * (CG_T + 1) / 2;            // [CG_DIAG_LDR] r_k -> z_k
      double* dg = R + CG_DIAG_LDR;                     // [64] pivots
      double* Xc = dg + CG_T;                           // [64][CG_DIAG_LDR] columns of the inverse
      double* A = g.K + (size_t)k * CG_T * Np + (size_t)k * CG_T;
      double* rk = g.r + k * CG_T;
      if (wtid == 0) s_bad = 0;
      for (int e = wtid; e < CG_T * CG_T; e += CG_WT) {
        const int i = e >> 6, j = e & 63;
        if (j <= i) S[i * (i + 1) / 2 + j] = A[(size_t)i * Np + j];          // own tile: written by this CTA only
      }
      if (wtid < CG_T) R[wtid] = __ldcg(rk + wtid);                           // r_k was updated by other CTAs' panels
      named_bar_sync(3, CG_WT);
      chol_packed_blocked_sub<CG_WT, FAST_PIVOT, 3>(S, R, dg, CG_T, &s_bad, true, 1, CG_DIAG_LDR);
      cg_tri_inverse64(S, dg, Xc, wtid, 3);
      double* Dg = g.Dinv + (size_t)k * CG_T * CG_T;
      for (int e = wtid; e < CG_T * CG_T; e += CG_WT) {
        const int i = e >> 6, j = e & 63;
        A[(size_t)i * Np + j] = (j <= i) ? S[i * (i + 1) / 2 + j] : 0.0;
        Dg[e] = (j <= i) ? Xc[j * CG_DIAG_LDR + i] : 0.0;                     // D^{-1}[i][j] = (column j)[i]
      }
      if (wtid < CG_T) rk[wtid] = R[wtid];
      if (wtid < 32) {                                                         // block partials, fixed order
        double zz = R[wtid] * R[wtid] + R[wtid + 32] * R[wtid + 32];
        double lg = log(dg[wtid]) + log(dg[wtid + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { zz += __shfl_xor_sync(0xffffffffu, zz, o); lg += __shfl_xor_sync(0xffffffffu, lg, o); }
        if (wtid == 0) { part[k * 4 + 0] = zz; part[k * 4 + 1] = lg; part[k * 4 + 2] = s_bad ? 1.0 : 0.0; }
      }
      if (C > 1) cg_publish_store_w(g.flagA, baseA + k + 1, 3, wtid); else named_bar_sync(3, CG_WT);
    }
    // ---- (2) deferred trailing tiles of step k-1 (columns >= k+1), pulled from the task counter ----------------------
    if (k > 0) {
      for (;;) {
        if (wtid == 0) {
          int L = -1;
          for (;;) {
            const int t = atomicAdd(&s_next, 1);
            const int cand = me + t * C;
            if (cand >= ntiles) break;
            int i, j;
            cg_unrank(cand, i, j);
            if (j >= k + 1) { L = cand; break; }
          }
          s_task[w] = L;
        }
        named_bar_sync(1 + w, CG_WT);
        const int L = s_task[w];
        named_bar_sync(1 + w, CG_WT);                      // everyone has read the slot before the leader reuses it
        if (L < 0) break;
        int i, j;
        cg_unrank(L, i, j);
        cg_update_tile(g, i, j, k - 1, As, Bs, w, wtid, wm, wn, lane);
      }
    }
    __syncthreads();
    if (k == nb - 1) break;
    // ---- (3) panel k: L_ik = A_ik D_k^{-T}, r_i -= L_ik z_k -----------------------------------------------------------
    if (C > 1) cg_wait_ge(g.flagA, baseA + k + 1);
    {
      int t = 0;
      for (int i = k + 1; i < nb; ++i) {
        if (cg_owner(i, k, C) != me) continue;
        if ((t++ & 1) != w) continue;
        double* Aik = g.K + (size_t)i * CG_T * Np + (size_t)k * CG_T;
        cg_stage_tile(As, Aik, Np, wtid);
        cg_stage_tile(Bs, g.Dinv + (size_t)k * CG_T * CG_T, CG_T, wtid);
        cg_stage_wait();
        named_bar_sync(1 + w, CG_WT);
        CgAcc acc;
        cg_tile_mma(As, Bs, acc, wm, wn, lane);
        named_bar_sync(1 + w, CG_WT);                                          // everyone has read As/Bs
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int rr = wm * 32 + a * 8 + (lane >> 2), cc = wn * 32 + b * 8 + 2 * (lane & 3);
            double2 o; o.x = acc.v[a][b][0]; o.y = acc.v[a][b][1];
            *reinterpret_cast<double2*>(Aik + (size_t)rr * Np + cc) = o;
            As[rr * CG_LD + cc] = o.x; As[rr * CG_LD + cc + 1] = o.y;          // keep L_ik for the rhs update
          }
        if (wtid < CG_T) Bs[wtid] = __ldcg(g.r + k * CG_T + wtid);             // z_k
        named_bar_sync(1 + w, CG_WT);
        if (wtid < CG_T) {
          double s = 0.0;
#pragma unroll 8
          for (int c = 0; c < CG_T; ++c) s = fma(As[wtid * CG_LD + c], Bs[c], s);
          double* ri = g.r + i * CG_T + wtid;
          *ri = __ldcg(ri) - s;
        }
        named_bar_sync(1 + w, CG_WT);
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
        cg_update_tile(g, i, j, k, As, Bs, w, wtid, wm, wn, lane);
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
