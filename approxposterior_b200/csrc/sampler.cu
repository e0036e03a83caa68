// Device-resident affine-invariant ensemble sampler (Goodman & Weare stretch move, red/blue
// split) over the GP-surrogate posterior mean.  One CTA owns one independent ensemble for the
// whole chain; ensembles never talk to each other, so a grid of CTAs needs no grid-wide sync.
//
// Replaces emcee.EnsembleSampler(...).sample(...) as driven from reference approx.py:839-847
// with log_prob_fn = ApproxPosterior._gpll (approx.py:148-189):
//   lnprob(q) = -inf                          if q non-finite or outside the (box) prior
//             = m + k*(q)^T alpha             otherwise (mean-only predict, approx.py:178-180)
//   blob      = lnprior (nan when rejected)   (approx.py:188, blobs_dtype approx.py:843)
// Per iteration (emcee 3.0.x RedBlueMove/StretchMove): shuffle colours, then for each colour
//   zz = ((a-1)u+1)^2/a ; q = c[r] - (c[r]-s) zz ; accept iff (d-1)ln zz + lp(q) - lp(s) > ln u'.
// Random draws come from Philox4x32-10, or from replay buffers recorded by the CPU oracle so
// that chains can be compared draw-for-draw.
#include "apgp_internal.h"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace apgp {
namespace {
namespace cg = cooperative_groups;

struct Smem {
  double* xs;      // [d+1][Npad] (alphaA is row d) or null when it does not fit
  double* coords;  // [nw][d]
  double* lp;      // [nw]
  double* blob;    // [nw]
  double* q;       // [nw][d]   proposals (initial pass evaluates all nw walkers)
  double* sq;      // [nw][d]   the same rows scaled for the kernel distance
  double* nlp;     // [nw]
  const double* etab;  // [256] 2^(j/256) for exp_neg256
  int* ok;         // [nw]
  // everything random about a step is independent of the chain's state, so it is drawn for SB steps at a time by
  // all threads in parallel (the per-step critical path is then: propose, evaluate, accept)
  unsigned long long* key;  // [SB][nw] random keys for the colouring
  int* colour;     // [SB][nw]
  int* list;       // [SB][2][Ns] walkers of colour 0 / colour 1, in walker order
  int* rint;       // [SB][2][Ns] partner index into the complementary list
  double* zz;      // [SB][2][Ns] stretch factor
  double* fac;     // [SB][2][Ns] (d-1) ln zz
  double* logu;    // [SB][2][Ns] ln u of the accept test
};

// lnprob for the `np` rows of sm.q.  A warp takes two rows per pass (they share every training-set load),
// lanes stride over the training points; sm.sq holds the rows pre-scaled by sqrt(1/(2 M_i)).
// inner product loop of eval_rows over the training points; XS is the [d+1][Npad] SoA block (row d = alphaA).
// Two query rows x two training points per trip: every operand load is shared by at least two evaluations,
// and all addressing is 32-bit offset arithmetic (the kernel is issue-bound, not FP64-bound).
// D > 0: the dimension is a compile-time constant (the distance loop unrolls and the two query rows live in registers);
// D = 0: generic run-time d.
template <int D>
__device__ __forceinline__ void eval_pair(const double* __restrict__ xs, const double* __restrict__ al,
                                          const double* __restrict__ q0s, const double* __restrict__ q1s,
                                          const double* __restrict__ etab, int N, int Npad, int d_rt, int lane,
                                          double& acc0, double& acc1) {
  const int d = D ? D : d_rt;
  double q0r[D ? D : 1], q1r[D ? D : 1];
  if (D) {
#pragma unroll
    for (int c = 0; c < (D ? D : 1); ++c) { q0r[c] = q0s[c]; q1r[c] = q1s[c]; }
  }
  const double* q0 = D ? q0r : q0s;
  const double* q1 = D ? q1r : q1s;
  int j = lane;
  // four training points per trip: eight independent exponential chains per thread (the loop is bound by the dependent
  // latency of the chain, ~10 FP64 operations deep, not by issue slots or the FP64 pipe: ncu 50 % / 71 % with two
  // points per trip).  The per-lane summation order is unchanged (j, j+32, j+64, j+96, ...), so results are
  // bit-identical to the two-point loop.
  for (; j + 96 < N; j += 128) {
    double s00 = 0.0, s01 = 0.0, s02 = 0.0, s03 = 0.0, s10 = 0.0, s11 = 0.0, s12 = 0.0, s13 = 0.0;
    int off = j;
#pragma unroll
    for (int c = 0; c < d; ++c, off += Npad) {
      const double xa = xs[off], xb = xs[off + 32], xc = xs[off + 64], xd = xs[off + 96];
      const double qa = q0[c], qb = q1[c];
      double t;
      t = xa - qa; s00 = fma(t, t, s00);
      t = xb - qa; s01 = fma(t, t, s01);
      t = xc - qa; s02 = fma(t, t, s02);
      t = xd - qa; s03 = fma(t, t, s03);
      t = xa - qb; s10 = fma(t, t, s10);
      t = xb - qb; s11 = fma(t, t, s11);
      t = xc - qb; s12 = fma(t, t, s12);
      t = xd - qb; s13 = fma(t, t, s13);
    }
    const double a0 = al[j], a1 = al[j + 32], a2 = al[j + 64], a3 = al[j + 96];
    const double e00 = exp_neg256<false>(s00, etab), e01 = exp_neg256<false>(s01, etab), e02 = exp_neg256<false>(s02, etab), e03 = exp_neg256<false>(s03, etab);
    const double e10 = exp_neg256<false>(s10, etab), e11 = exp_neg256<false>(s11, etab), e12 = exp_neg256<false>(s12, etab), e13 = exp_neg256<false>(s13, etab);
    acc0 = fma(e00, a0, acc0); acc0 = fma(e01, a1, acc0); acc0 = fma(e02, a2, acc0); acc0 = fma(e03, a3, acc0);
    acc1 = fma(e10, a0, acc1); acc1 = fma(e11, a1, acc1); acc1 = fma(e12, a2, acc1); acc1 = fma(e13, a3, acc1);
  }
  for (; j + 32 < N; j += 64) {
    double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;      // s[row][point]
    int off = j;
#pragma unroll
    for (int c = 0; c < d; ++c, off += Npad) {
      const double xa = xs[off], xb = xs[off + 32];
      const double qa = q0[c], qb = q1[c];
      double t;
      t = xa - qa; s00 = fma(t, t, s00);
      t = xb - qa; s01 = fma(t, t, s01);
      t = xa - qb; s10 = fma(t, t, s10);
      t = xb - qb; s11 = fma(t, t, s11);
    }
    const double a0 = al[j], a1 = al[j + 32];
    acc0 = fma(exp_neg256<false>(s00, etab), a0, acc0); acc0 = fma(exp_neg256<false>(s01, etab), a1, acc0);
    acc1 = fma(exp_neg256<false>(s10, etab), a0, acc1); acc1 = fma(exp_neg256<false>(s11, etab), a1, acc1);
  }
  for (; j < N; j += 32) {
    double s0 = 0.0, s1 = 0.0;
    int off = j;
#pragma unroll
    for (int c = 0; c < d; ++c, off += Npad) {
      const double x = xs[off];
      const double d0 = x - q0[c], d1 = x - q1[c];
      s0 = fma(d0, d0, s0); s1 = fma(d1, d1, s1);
    }
    const double a = al[j];
    acc0 = fma(exp_neg256<false>(s0, etab), a, acc0);
    acc1 = fma(exp_neg256<false>(s1, etab), a, acc1);
  }
}

// rows [r0, np) of sm.q (r0 even)
// nslice > 0: "split training set" mode -- sm.xs holds only this CTA's slice ([d+1][slice_pad], nslice points) and the
// raw partial sums (no mean, no finiteness gate) go to sm.nlp for the cluster-wide reduction that follows.
__device__ void eval_rows(const SamplerParams& p, const Smem& sm, int np, int r0 = 0, int nslice = 0, int slice_pad = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int d = p.d, Npad = nslice ? slice_pad : p.Npad;
  const int Neff = nslice ? nslice : p.N;
  for (int i0 = r0 + 2 * warp; i0 < np; i0 += 2 * nwarps) {
    const int i1 = (i0 + 1 < np) ? i0 + 1 : i0;
    const int ok0 = sm.ok[i0], ok1 = sm.ok[i1];
    __syncwarp();
    double acc0 = 0.0, acc1 = 0.0;
    if (ok0 | ok1) {
      const double* q0 = sm.sq + i0 * d;
      const double* q1 = sm.sq + i1 * d;
      if (sm.xs) {
        const double* al = sm.xs + d * Npad;
        switch (d) {                       // the shared-memory path is the hot one: specialise small dimensions (the kernel
                                           // is capped at 64 registers by its 1024-thread launch bound: larger d would spill)
          case 1: eval_pair<1>(sm.xs, al, q0, q1, sm.etab, Neff, Npad, d, lane, acc0, acc1); break;
          case 2: eval_pair<2>(sm.xs, al, q0, q1, sm.etab, Neff, Npad, d, lane, acc0, acc1); break;
          case 3: eval_pair<3>(sm.xs, al, q0, q1, sm.etab, Neff, Npad, d, lane, acc0, acc1); break;
          default: eval_pair<0>(sm.xs, al, q0, q1, sm.etab, Neff, Npad, d, lane, acc0, acc1); break;
        }
      } else {
        eval_pair<0>(p.Xs, p.alphaA, q0, q1, sm.etab, p.N, Npad, d, lane, acc0, acc1);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
      }
    }
    if (lane < 2) {
      const int i = lane ? i1 : i0;
      const int oki = lane ? ok1 : ok0;
      if (nslice) {
        if (lane == 0 || i1 != i0) sm.nlp[i] = lane ? acc1 : acc0;          // partial sum over this CTA's slice
      } else {
        const double mu = p.mean + (lane ? acc1 : acc0);
        const bool fin = oki && (mu == mu) && (fabs(mu) < INFINITY);
        if (lane == 0 || i1 != i0) { sm.nlp[i] = fin ? mu : -INFINITY; sm.ok[i] = fin ? 1 : 0; }
      }
    }
  }
}

// prior gate on the raw row + its scaled copy for eval_rows
__device__ __forceinline__ int stage_row(const SamplerParams& p, const Smem& sm, int i) {
  int ok = 1;
  for (int c = 0; c < p.d; ++c) {
    const double v = sm.q[i * p.d + c];
    ok = ok && (v == v) && (v >= p.lo[c]) && (v <= p.hi[c]);
    sm.sq[i * p.d + c] = v * p.qscale[c];
  }
  return ok;
}

// C = CTAs per ensemble (a thread-block cluster when C > 1).  A single ensemble is the reference's own usage
// (one emcee.EnsembleSampler of 20*ndim walkers, mcmcUtils.py:75) and would otherwise live on ONE SM: with a cluster
// every CTA keeps a full replica of the ensemble state and of the pre-drawn randomness, proposes / evaluates /
// accepts its slice of each half-step's walkers, writes the accepted moves into every replica through distributed
// shared memory, and the cluster meets at one hardware cluster barrier per half-step.  Same draws, same arithmetic
// per walker: chains are bit-identical to C = 1.
// tsplit > 0: the scaled training set does not fit one CTA's shared memory, so it is PARTITIONED over the cluster instead
// (tsplit = points per CTA, padded): every CTA keeps its slice resident, evaluates ALL proposals of a half-step against
// it, the partial sums meet in every CTA through distributed shared memory (one cluster barrier per half-step) and are
// added in rank order -- every replica then takes the same accept decisions locally, with no remote state writes.
__global__ void __launch_bounds__(1024) sampler_kernel(const __grid_constant__ SamplerParams p, int xs_in_smem, int SB,
                                                       int C, int tsplit) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int e = blockIdx.x / C, rank = blockIdx.x % C, tid = threadIdx.x;
  cg::cluster_group cluster = cg::this_cluster();
  const int nw = p.nwalk, d = p.d, Ns = nw / 2, Npad = p.Npad;
  const long W = (long)p.nens * nw;
  Smem sm;
  __shared__ double etab_s[256];
  if (tid < 256) etab_s[tid] = exp2((double)tid * (1.0 / 256.0));
  if (blockDim.x < 256) for (int j = blockDim.x + tid; j < 256; j += blockDim.x) etab_s[j] = exp2((double)j * (1.0 / 256.0));
  sm.etab = etab_s;
  double* f = reinterpret_cast<double*>(raw);
  sm.xs = nullptr;
  double* parts = nullptr;                   // split mode: [2][C][nw] partial sums from every CTA of the cluster
  const int n_lo = tsplit * rank;
  const int nslice = tsplit ? max(0, min(tsplit, p.N - n_lo)) : 0;
  if (xs_in_smem) { sm.xs = f; f += (size_t)(d + 1) * Npad; }
  if (tsplit) { sm.xs = f; f += (size_t)(d + 1) * tsplit; parts = f; f += (size_t)2 * C * nw; }
  sm.coords = f; f += nw * d;
  sm.lp = f; f += nw;
  sm.blob = f; f += nw;
  sm.q = f; f += nw * d;
  sm.sq = f; f += nw * d;
  sm.nlp = f; f += nw;
  sm.zz = f; f += (size_t)SB * nw;
  sm.fac = f; f += (size_t)SB * nw;
  sm.logu = f; f += (size_t)SB * nw;
  sm.key = reinterpret_cast<unsigned long long*>(f); f += (size_t)SB * nw;
  int* ip = reinterpret_cast<int*>(f);
  sm.colour = ip; ip += (size_t)SB * nw;
  sm.list = ip; ip += (size_t)SB * nw;
  sm.rint = ip; ip += (size_t)SB * nw;
  sm.ok = ip; ip += nw;

  if (xs_in_smem) {
    for (int idx = tid; idx < (d + 1) * Npad; idx += blockDim.x)
      sm.xs[idx] = (idx < d * Npad) ? p.Xs[idx] : p.alphaA[idx - d * Npad];
  }
  if (tsplit) {
    for (int idx = tid; idx < (d + 1) * tsplit; idx += blockDim.x) {
      const int c = idx / tsplit, t = idx - c * tsplit;
      double v = 0.0;
      if (t < nslice) v = (c < d) ? p.Xs[(size_t)c * Npad + n_lo + t] : p.alphaA[n_lo + t];
      sm.xs[idx] = v;
    }
  }
  int hs = 0;                                // exchanges done so far (parity selects the parts buffer)
  // split mode: publish my partial sums of rows [0, np) to every CTA, meet, add in rank order, gate
  auto combine = [&](int np) {
    double* mine = parts + (size_t)(hs & 1) * C * nw;
    for (int i = tid; i < np; i += blockDim.x) {
      const double v = sm.nlp[i];
      for (int r = 0; r < C; ++r) cluster.map_shared_rank(mine, r)[rank * nw + i] = v;
    }
    cluster.sync();
    for (int i = tid; i < np; i += blockDim.x) {
      double mu = p.mean;
      for (int r = 0; r < C; ++r) mu += mine[r * nw + i];
      const bool fin = sm.ok[i] && (mu == mu) && (fabs(mu) < INFINITY);
      sm.nlp[i] = fin ? mu : -INFINITY; sm.ok[i] = fin ? 1 : 0;
    }
    ++hs;
    __syncthreads();
  };
  for (int idx = tid; idx < nw * d; idx += blockDim.x) {
    double v = p.p0[(size_t)e * nw * d + idx];
    sm.coords[idx] = v; sm.q[idx] = v;
  }
  __syncthreads();
  for (int w = tid; w < nw; w += blockDim.x) sm.ok[w] = stage_row(p, sm, w);
  __syncthreads();
  if (tsplit) cluster.sync();                // every CTA's parts buffers exist before anyone writes into a peer
  eval_rows(p, sm, nw, 0, tsplit ? (nslice ? nslice : -1) : 0, tsplit);
  __syncthreads();
  if (tsplit) combine(nw);
  for (int w = tid; w < nw; w += blockDim.x) { sm.lp[w] = sm.nlp[w]; sm.blob[w] = sm.ok[w] ? p.lnprior_const : NAN; }
  __syncthreads();
  if (C > 1) cluster.sync();                 // every replica initialised before anyone writes into a peer
  // my slice of each half-step's Ns proposals (even boundaries: a warp evaluates rows in pairs) and of the walkers
  // whose chain entries I store
  const int i_lo = tsplit ? 0 : ((Ns * rank) / C) & ~1;
  const int i_hi = tsplit ? Ns : ((rank == C - 1) ? Ns : (((Ns * (rank + 1)) / C) & ~1));
  const int w_lo = (nw * rank) / C, w_hi = (nw * (rank + 1)) / C;

  Philox rng; rng.k0 = (uint32_t)p.seed; rng.k1 = (uint32_t)(p.seed >> 32);
  const bool replay = p.r_zz != nullptr;
  const int nst_tot = p.nsteps_total ? p.nsteps_total : p.nsteps;     // draws are indexed by the chain's global step

  for (int step0 = 0; step0 < p.nsteps; step0 += SB) {
    const int sb = min(SB, p.nsteps - step0);
    // ---------------------------------------------------------------- randomness of the next sb steps, in parallel
    if (replay) {
      for (int idx = tid; idx < sb * nw; idx += blockDim.x) {
        const int s = idx / nw, w = idx - s * nw;
        sm.colour[idx] = p.r_inds[((size_t)e * nst_tot + p.step_base + step0 + s) * nw + w];
      }
    } else {
      // uniformly random half/half colouring (same law as emcee's shuffle of arange(nw) % 2): every walker
      // draws a key, the nw/2 smallest keys are colour 0.  Fully parallel -- no serial Fisher-Yates.
      for (int idx = tid; idx < sb * nw; idx += blockDim.x) {
        const int s = idx / nw, w = idx - s * nw;
        uint32_t o[4]; rng.gen((uint32_t)e, (uint32_t)(p.step_base + step0 + s), 0x10000u, (uint32_t)w, o);
        sm.key[idx] = ((uint64_t)o[0] << 32) | o[1];
      }
      __syncthreads();
      for (int idx = tid; idx < sb * nw; idx += blockDim.x) {
        const int s = idx / nw, w = idx - s * nw;
        const unsigned long long* ks = sm.key + (size_t)s * nw;
        const uint64_t kw = ks[w];
        int rank = 0;
        for (int v = 0; v < nw; ++v) { const uint64_t kv = ks[v]; rank += (kv < kw) || (kv == kw && v < w); }
        sm.colour[idx] = rank < Ns ? 0 : 1;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < sb * nw; idx += blockDim.x) {     // position of w among the walkers of its colour
      const int s = idx / nw, w = idx - s * nw;
      const int* cs = sm.colour + (size_t)s * nw;
      const int cw = cs[w];
      int pos = 0;
      for (int v = 0; v < w; ++v) pos += (cs[v] == cw);
      sm.list[((size_t)s * 2 + cw) * Ns + pos] = w;
    }
    for (int idx = tid; idx < sb * 2 * Ns; idx += blockDim.x) {   // stretch factor, partner, accept threshold
      const int s = idx / (2 * Ns), rem = idx - s * 2 * Ns, split = rem / Ns, i = rem - split * Ns;
      const int step = p.step_base + step0 + s;
      double zz, lu; int r;
      if (replay) {
        const size_t off = (((size_t)e * nst_tot + step) * 2 + split) * Ns + i;
        zz = p.r_zz[off]; r = p.r_rint[off]; lu = p.r_logu[off];
      } else {
        uint32_t o[4], o2[4];
        rng.gen((uint32_t)e, (uint32_t)step, (uint32_t)split, (uint32_t)i, o);
        rng.gen((uint32_t)e, (uint32_t)step, (uint32_t)(split + 2), (uint32_t)i, o2);
        const double u = u01_from_bits(o[0], o[1]);
        const double t = (p.a - 1.0) * u + 1.0;
        zz = t * t / p.a;
        lu = log(u01_from_bits(o[2], o[3]));
        r = (int)(((uint64_t)o2[0] * (uint64_t)Ns) >> 32);
      }
      sm.zz[idx] = zz; sm.rint[idx] = r; sm.logu[idx] = lu; sm.fac[idx] = (d - 1.0) * log(zz);
    }
    __syncthreads();

    // ---------------------------------------------------------------- the sb steps themselves
    for (int s = 0; s < sb; ++s) {
      const int step = step0 + s;
      for (int split = 0; split < 2; ++split) {
        const int* sidx = sm.list + ((size_t)s * 2 + split) * Ns;          // walkers being moved
        const int* cidx = sm.list + ((size_t)s * 2 + (split ^ 1)) * Ns;    // the complementary half
        const size_t boff = ((size_t)s * 2 + split) * Ns;
        for (int i = i_lo + tid; i < i_hi; i += blockDim.x) {
          const double zz = sm.zz[boff + i];
          const double* cs = sm.coords + cidx[sm.rint[boff + i]] * d;
          const double* ss = sm.coords + sidx[i] * d;
          for (int c = 0; c < d; ++c) sm.q[i * d + c] = cs[c] - (cs[c] - ss[c]) * zz;
          sm.ok[i] = stage_row(p, sm, i);
        }
        __syncthreads();
        eval_rows(p, sm, i_hi, i_lo, tsplit ? (nslice ? nslice : -1) : 0, tsplit);
        __syncthreads();
        if (tsplit) combine(Ns);
        for (int i = i_lo + tid; i < i_hi; i += blockDim.x) {
          const int j = sidx[i];
          const double diff = sm.fac[boff + i] + sm.nlp[i] - sm.lp[j];
          if (diff > sm.logu[boff + i]) {
            const double nl = sm.nlp[i], bl = sm.ok[i] ? p.lnprior_const : NAN;
            if (C > 1 && !tsplit) {
              for (int r = 0; r < C; ++r) {                    // the move goes into every replica of the state
                double* rc = cluster.map_shared_rank(sm.coords, r);
                for (int c = 0; c < d; ++c) rc[j * d + c] = sm.q[i * d + c];
                cluster.map_shared_rank(sm.lp, r)[j] = nl;
                cluster.map_shared_rank(sm.blob, r)[j] = bl;
              }
            } else {
              for (int c = 0; c < d; ++c) sm.coords[j * d + c] = sm.q[i * d + c];
              sm.lp[j] = nl;
              sm.blob[j] = bl;
            }
            if (!tsplit || rank == 0) atomicAdd(&p.naccept[(size_t)e * nw + j], 1);
          }
        }
        if (C > 1 && !tsplit) cluster.sync(); else __syncthreads();
      }
      if ((step + 1) % p.thin == 0) {
        const long srow = (step + 1) / p.thin - 1;
        for (int idx = w_lo * d + tid; idx < w_hi * d; idx += blockDim.x)
          p.chain[(srow * W + (size_t)e * nw) * d + idx] = sm.coords[idx];
        for (int w = w_lo + tid; w < w_hi; w += blockDim.x) {
          p.logp[srow * W + (size_t)e * nw + w] = sm.lp[w];
          p.blob[srow * W + (size_t)e * nw + w] = sm.blob[w];
        }
        if (C > 1 && !tsplit) cluster.sync(); // my reads of the replica finish before a peer's next accepted move lands
      }
      // no sync needed: the next writes to coords/lp/blob happen after two more __syncthreads
    }
  }
  if (p.final_state && rank == 0) {
    __syncthreads();
    for (int idx = tid; idx < nw * d; idx += blockDim.x) p.final_state[(size_t)e * nw * d + idx] = sm.coords[idx];
  }
  if (C > 1) cluster.sync();                 // no replica may disappear while a peer can still write into it
}

}  // namespace

int launch_sampler(const SamplerParams& p, cudaStream_t st, int* launches) {
  if (p.nwalk < 2 || (p.nwalk & 1) || p.nwalk > 1024 || p.nens < 1) return (int)cudaErrorInvalidValue;
  // shared memory: per-walker state, then SB steps' worth of pre-drawn randomness (44 bytes per walker per step),
  // then -- if it still fits -- the scaled training set
  const size_t cap = 200 * 1024;
  const size_t state = (size_t)p.nwalk * (3 * p.d + 3) * 8 + (size_t)p.nwalk * 4 + 64;
  const size_t per_step = (size_t)p.nwalk * 44;
  const size_t xs_bytes = (size_t)(p.d + 1) * p.Npad * 8;
  if (state + per_step > cap) return (int)cudaErrorInvalidValue;
  int xs_in_smem = (state + per_step + xs_bytes <= cap) ? 1 : 0;
  // the training set does not fit one CTA: partition it over a cluster of Cs CTAs (each keeps N / Cs points resident and
  // the partial sums meet through distributed shared memory) -- unless even 8 slices do not fit, or the caller forbids it
  int split = 0, Cs = 1;
  if (!xs_in_smem && !getenv("APGP_SAMPLER_NO_SPLIT")) {
    // the smallest slice count that fits, then -- while the grid stays within half the GPU (a cluster barrier per
    // half-step costs more when the SMs are oversubscribed) -- as many slices as possible: measured at N = 2000, d = 10,
    // 200 walkers (tools/bench_sampler_split.py): 1 ensemble 100 us per step streamed from L2, 88 / 51 / 33 us with
    // 2 / 4 / 8 slices; 16 ensembles 202 us streamed, 90 / 52 / 66 us
    for (int c = 2; c <= 8; c <<= 1) {
      const int per = (((p.N + c - 1) / c) + 1) & ~1;
      const size_t need = state + per_step + (size_t)(p.d + 1) * per * 8 + (size_t)2 * c * p.nwalk * 8;
      if (need > cap) continue;
      if (split && (long long)p.nens * c > 74) break;
      split = per; Cs = c;
    }
    if (const char* sv = getenv("APGP_SAMPLER_SPLIT")) {          // A/B runs: force the number of slices
      const int c = atoi(sv);
      if (c == 2 || c == 4 || c == 8) {
        const int per = (((p.N + c - 1) / c) + 1) & ~1;
        if (state + per_step + (size_t)(p.d + 1) * per * 8 + (size_t)2 * c * p.nwalk * 8 <= cap) { split = per; Cs = c; }
      }
    }
  }
  const size_t resident = xs_in_smem ? xs_bytes : (split ? (size_t)(p.d + 1) * split * 8 + (size_t)2 * Cs * p.nwalk * 8 : 0);
  const size_t room = cap - state - resident;
  int SB = (int)(room / per_step);
  if (SB > 32) SB = 32;
  if (SB > p.nsteps) SB = p.nsteps;
  if (SB < 1) SB = 1;
  const size_t smem = state + (size_t)SB * per_step + resident;
  static PerDeviceOnce attr;
  if (attr.needed()) {
    cudaError_t e = cudaFuncSetAttribute(sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr.mark();
  }
  int Ns = p.nwalk / 2;
  // CTAs per ensemble: clusters only pay when there are few ensembles (the grid would not fill the GPU anyway) and a
  // half-step carries enough work to amortise the cluster barrier: >= 8 proposals per CTA and >= 6000 kernel
  // evaluations per CTA per half-step (measured, tools/bench_single_ensemble.py: 100 walkers at N = 500 / 1000 gain
  // 1.3x / 1.45x with C = 4, 40 walkers at N = 90 lose 20 %).  APGP_SAMPLER_CLUSTER overrides (1, 2, 4, 8) for A/B runs.
  int C = 1;
  if (split) {
    C = Cs;                                   // the slices ARE the cluster
  } else {
    if (p.nens <= 32) {
      for (int c = 8; c >= 2; c >>= 1)
        if (Ns / c >= 8 && (long long)Ns * p.N / c >= 6000) { C = c; break; }
    }
    if (const char* cv = getenv("APGP_SAMPLER_CLUSTER")) { int c = atoi(cv); if (c == 1 || c == 2 || c == 4 || c == 8) C = c; }
    while (C > 1 && Ns / C < 2) C >>= 1;
  }
  // one warp per pair of proposals of a CTA's slice (a warp evaluates two rows per pass), 8..32 warps: a single large
  // ensemble lives on few SMs, so its parallelism is warps (APGP_SAMPLER_WARPS overrides, for A/B runs)
  const int pairs = split ? (Ns + 1) / 2 : ((Ns + C - 1) / C + 1) / 2;      // split mode: every CTA evaluates all proposals
  int nwarps = Ns < 8 ? (Ns < 1 ? 1 : Ns) : (pairs < 8 ? 8 : (pairs > 32 ? 32 : pairs));
  if (const char* wv = getenv("APGP_SAMPLER_WARPS")) { int w = atoi(wv); if (w >= 1 && w <= 32) nwarps = w; }
  if (p.step_base == 0) cudaMemsetAsync(p.naccept, 0, sizeof(int) * (size_t)p.nens * p.nwalk, st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.nens * C)); cfg.blockDim = dim3((unsigned)(nwarps * 32));
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = (unsigned)C; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, sampler_kernel, p, xs_in_smem, SB, C, split);
  if (le != cudaSuccess && C > 1) {
    // the cluster could not be scheduled (e.g. a partitioned GPU): one CTA per ensemble gives the same chains
    (void)cudaGetLastError();
    C = 1;
    const int pairs1 = (Ns + 1) / 2;
    nwarps = Ns < 8 ? (Ns < 1 ? 1 : Ns) : (pairs1 < 8 ? 8 : (pairs1 > 32 ? 32 : pairs1));
    cfg.gridDim = dim3((unsigned)p.nens); cfg.blockDim = dim3((unsigned)(nwarps * 32));
    attrs[0].val.clusterDim.x = 1;
    // (a split launch falls back to streaming the training set from L2; SB was sized for the smaller footprint)
    le = cudaLaunchKernelEx(&cfg, sampler_kernel, p, xs_in_smem, SB, C, 0);
  }
  if (le != cudaSuccess) return (int)le;
  if (launches) ++*launches;
  return (int)cudaGetLastError();
}

}  // namespace apgp
