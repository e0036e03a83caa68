// Integrated autocorrelation time of a device-resident chain: emcee's estimator (emcee.autocorr.integrated_time,
// 3.0.x) as reference mcmcUtils.estimateBurnin consumes it (mcmcUtils.py:198: tau = sampler.get_autocorr_time(tol=0),
// then iburn = int(2 max tau), ithin = int(0.5 min tau)), without copying the chain to the host.
//
//   per dimension k:  f(t) = mean over walkers w of  acf_w(t) / acf_w(0),
//                     acf_w(t) = sum_{s=0}^{n-1-t} (x_w(s) - mean_w)(x_w(s+t) - mean_w)     (linear, zero padded)
//                     tau(M)  = 2 sum_{t<=M} f(t) - 1 ;  window = smallest M with M >= c tau(M)  (Sokal, c = 5)
//
// emcee evaluates acf_w with one zero-padded FFT per walker over ALL n lags and then looks at the first few hundred.
// Here the lags are summed directly, and only as many as the window needs: lags [0, T) first, T doubling until the
// window condition is met inside the evaluated range (or T = n).  Cost n T per series instead of n log n with a
// library FFT and a host round trip; T ~ 256 for the chains approxposterior produces (tau ~ 30-60).
//
// Work split.  A series (walker w, dimension k) is staged in shared memory with its mean removed.  When the grid is
// large enough (many walkers) a CTA walks through G walkers of one dimension one after the other, normalises each
// acf by its lag 0 and keeps the running sum over its walkers in registers; when there are few walkers (the README's 20)
// every series is split into `nchunk` s-ranges handled by different CTAs, which store un-normalised partial sums.
// A second kernel adds the partials IN A FIXED ORDER (chunks, then walkers), so results do not depend on scheduling.
#include "apgp_internal.h"

namespace apgp {
namespace {

constexpr int AC_THREADS = 256;
constexpr int AC_LPT = 4;                       // lags per thread: a CTA covers AC_THREADS * AC_LPT = 1024 lags per pass
constexpr size_t AC_SMEM_MAX = 216 * 1024;

struct AcParams {
  const double* chain;     // [n_total][W][d]
  long long off;           // first step used
  int thin;                // step stride
  int n;                   // series length after discard / thin
  int W, d;
  int T;                   // lags evaluated: [0, T)
  int G;                   // walkers per CTA (nchunk == 1) else 1
  int nchunk;              // s-range chunks per series
  int ngroups;             // ceil(W / G)
  double* partial;         // nchunk == 1: [d][ngroups][T] normalised sums over the group's walkers
                           // nchunk  > 1: [d][W][nchunk][T] un-normalised partial acf
};

// grid: (ngroups * nchunk, d)
__global__ void __launch_bounds__(AC_THREADS) acf_partial_kernel(const AcParams p) {
  extern __shared__ __align__(16) double xs[];           // [n + T] series, mean removed, zero tail
  __shared__ double red[AC_THREADS / 32];
  __shared__ double s_mean, s_c0;
  const int tid = threadIdx.x, k = blockIdx.y;
  const int grp = blockIdx.x / p.nchunk, chunk = blockIdx.x - grp * p.nchunk;
  const int n = p.n, T = p.T;
  const int s_lo = (int)((long long)n * chunk / p.nchunk), s_hi = (int)((long long)n * (chunk + 1) / p.nchunk);
  const size_t stride = (size_t)p.W * p.d * p.thin;
  double accsum[4][AC_LPT];                               // up to 4 passes of 1024 lags per CTA (T <= 4096 per launch)
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < AC_LPT; ++b) accsum[a][b] = 0.0;
  const int w0 = grp * p.G, w1 = (w0 + p.G < p.W) ? w0 + p.G : p.W;
  for (int w = w0; w < w1; ++w) {
    const double* src = p.chain + (size_t)p.off * p.W * p.d + (size_t)w * p.d + k;
    double part = 0.0;
    for (int s = tid; s < n; s += AC_THREADS) { const double v = src[(size_t)s * stride]; xs[s] = v; part += v; }
    for (int s = n + tid; s < n + T + AC_LPT; s += AC_THREADS) xs[s] = 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int i = 0; i < AC_THREADS / 32; ++i) t += red[i]; s_mean = t / n; }
    __syncthreads();
    const double mean = s_mean;
    double c0p = 0.0;
    for (int s = tid; s < n; s += AC_THREADS) { const double v = xs[s] - mean; xs[s] = v; c0p = fma(v, v, c0p); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c0p += __shfl_xor_sync(0xffffffffu, c0p, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = c0p;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int i = 0; i < AC_THREADS / 32; ++i) t += red[i]; s_c0 = t; }
    __syncthreads();
    const double inv_c0 = (p.nchunk == 1) ? 1.0 / s_c0 : 1.0;      // a walker that never moved gives inf/nan, as emcee does
    // lags t = pass*1024 + tid*4 + {0..3}: acc[j] = sum_s x[s] x[s + t + j]; the zero tail makes the upper limit n for all
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int t0 = pass * AC_THREADS * AC_LPT + tid * AC_LPT;
      if (pass * AC_THREADS * AC_LPT >= T) break;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if (t0 < T) {
        const int hi = (s_hi < n - t0) ? s_hi : n - t0;            // terms beyond n - t0 are zero anyway
        const double* xa = xs, * xb = xs + t0;
        int s = s_lo;
        double b0 = (s < hi + 3) ? xb[s] : 0.0, b1 = (s < hi + 3) ? xb[s + 1] : 0.0, b2 = (s < hi + 3) ? xb[s + 2] : 0.0;
        for (; s < hi; ++s) {
          const double a = xa[s], b3 = xb[s + 3];
          a0 = fma(a, b0, a0); a1 = fma(a, b1, a1); a2 = fma(a, b2, a2); a3 = fma(a, b3, a3);
          b0 = b1; b1 = b2; b2 = b3;
        }
      }
      accsum[pass][0] += a0 * inv_c0; accsum[pass][1] += a1 * inv_c0; accsum[pass][2] += a2 * inv_c0; accsum[pass][3] += a3 * inv_c0;
    }
    __syncthreads();                                              // the next walker overwrites xs
  }
  double* out = (p.nchunk == 1) ? p.partial + ((size_t)k * p.ngroups + grp) * T
                                : p.partial + (((size_t)k * p.W + grp) * p.nchunk + chunk) * T;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass)
#pragma unroll
    for (int j = 0; j < AC_LPT; ++j) {
      const int t = pass * AC_THREADS * AC_LPT + tid * AC_LPT + j;
      if (t < T) out[t] = accsum[pass][j];
    }
}

// f[k][t] = (1/W) sum over walkers (fixed order) of the normalised acf.  grid (ceil(T/256), d)
__global__ void acf_reduce_kernel(const AcParams p, double* __restrict__ f) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (t >= p.T) return;
  double s = 0.0;
  if (p.nchunk == 1) {
    const double* src = p.partial + (size_t)k * p.ngroups * p.T + t;
    for (int g = 0; g < p.ngroups; ++g) s += src[(size_t)g * p.T];
  } else {
    for (int w = 0; w < p.W; ++w) {
      const double* src = p.partial + ((size_t)k * p.W + w) * p.nchunk * p.T;
      double a = 0.0, c0 = 0.0;
      for (int c = 0; c < p.nchunk; ++c) { a += src[(size_t)c * p.T + t]; c0 += src[(size_t)c * p.T]; }
      s += a / c0;
    }
  }
  f[(size_t)k * p.T + t] = s / p.W;
}

PerDeviceOnce g_ac_attr;

}  // namespace

bool autocorr_fits(int n, int T) { return ((size_t)n + T + AC_LPT + 8) * sizeof(double) <= AC_SMEM_MAX && T <= 4 * AC_THREADS * AC_LPT; }

size_t autocorr_partial_bytes(int W, int d, int T, int num_sms, int* G_out, int* nchunk_out) {
  // enough CTAs to fill the GPU: with many walkers group them, with few split each series into s-chunks
  const long series = (long)W * d;
  int G = 1, nchunk = 1;
  if (series >= 8L * num_sms) { G = (int)(series / (8L * num_sms)); if (G < 1) G = 1; if (G > 64) G = 64; }
  else { nchunk = (int)((4L * num_sms + series - 1) / series); if (nchunk > 64) nchunk = 64; if (nchunk < 1) nchunk = 1; }
  *G_out = G; *nchunk_out = nchunk;
  const size_t ngroups = ((size_t)W + G - 1) / G;
  return (nchunk == 1 ? (size_t)d * ngroups * T : (size_t)d * W * nchunk * T) * sizeof(double);
}

// f_dev [d][T] <- walker-averaged normalised autocorrelation for lags [0, T)
int launch_autocorr(const double* chain, long long off, int thin, int n, int W, int d, int T, int num_sms, double* partial,
                    double* f_dev, cudaStream_t st, int* launches) {
  if (!autocorr_fits(n, T)) return (int)cudaErrorInvalidValue;
  if (g_ac_attr.needed()) {
    cudaError_t e = cudaFuncSetAttribute(acf_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AC_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    g_ac_attr.mark();
  }
  AcParams p;
  p.chain = chain; p.off = off; p.thin = thin; p.n = n; p.W = W; p.d = d; p.T = T;
  (void)autocorr_partial_bytes(W, d, T, num_sms, &p.G, &p.nchunk);
  p.ngroups = (W + p.G - 1) / p.G;
  p.partial = partial;
  const size_t smem = ((size_t)n + T + AC_LPT + 8) * sizeof(double);
  acf_partial_kernel<<<dim3((unsigned)(p.ngroups * p.nchunk), (unsigned)d), AC_THREADS, smem, st>>>(p);
  acf_reduce_kernel<<<dim3((unsigned)((T + 255) / 256), (unsigned)d), 256, 0, st>>>(p, f_dev);
  if (launches) *launches += 2;
  return (int)cudaGetLastError();
}

}  // namespace apgp
