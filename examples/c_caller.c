/* A plain-C caller of libapgp (no Python, no C++): the boundary of include/apgp.h used the way a maintainer's FFI
 * would use it.  Builds a small GP, factorises it, evaluates mean / variance / BAPE utility for a batch of queries
 * from host buffers, and checks the results: the log-likelihood against the value the CPU oracle gives for the same
 * (generated) inputs -- tests/test_z_c_caller.py recomputes that constant -- and that the GP interpolates its targets.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/c_caller.c -Lapproxposterior_b200 -lapgp \
 *       -Wl,-rpath,$PWD/approxposterior_b200 -lm -o /tmp/c_caller && /tmp/c_caller
 *
 * Exit status: 0 = ok; 3 = no usable CUDA device (the library has no CPU fallback and says so); 1 = wrong results. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "apgp.h"

#define EXPECT_LOGLIK 131.72132402644814   /* oracle.GPOracle.log_likelihood on the inputs generated below */

static unsigned long long lcg_state = 0x9E3779B97F4A7C15ull;
static double uniform(double lo, double hi) {
  lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
  return lo + (hi - lo) * (double)(lcg_state >> 11) / 9007199254740992.0;
}

int main(void) {
  enum { N = 100, D = 2, Q = 1000 };
  static double X[N * D], y[N], Xq[(Q + N) * D], mu[Q + N], var[Q + N], util[Q + N];
  apgp_handle* h = NULL;
  apgp_predict_opts o;
  double log_metric[D] = {1.0, 1.0}, logdet = 0.0, loglik = 0.0, ymax = -1e300, err = 0.0;
  int info = 0, i, st;

  printf("libapgp version %d\n", apgp_version());
  st = apgp_create(&h, 0);
  if (st != APGP_OK) {
    fprintf(stderr, "apgp_create failed (%d): %s\n", st, apgp_last_error());
    return 3;
  }
  for (i = 0; i < N; ++i) {
    X[i * D] = uniform(-5.0, 5.0); X[i * D + 1] = uniform(-5.0, 5.0);
    y[i] = -0.125 * (X[i * D] * X[i * D] + X[i * D + 1] * X[i * D + 1]);
    if (y[i] > ymax) ymax = y[i];
  }
  for (i = 0; i < Q * D; ++i) Xq[i] = uniform(-5.0, 5.0);
  memcpy(Xq + Q * D, X, sizeof(X));                       /* the training points themselves close the batch */
  if (apgp_set_training(h, X, y, N, D, 1) != APGP_OK || apgp_set_hyper(h, -3.0, 1.0, log_metric, -12.0) != APGP_OK) {
    fprintf(stderr, "setup failed: %s\n", apgp_last_error());
    return 1;
  }
  st = apgp_factorize(h, &logdet, &loglik, &info);
  if (st != APGP_OK) {
    fprintf(stderr, "apgp_factorize: status %d, info %d: %s\n", st, info, apgp_last_error());
    return 1;
  }
  memset(&o, 0, sizeof(o));
  o.want_var = 1; o.utility = APGP_UTIL_BAPE; o.has_box = 1; o.ybest = ymax;
  for (i = 0; i < D; ++i) { o.lo[i] = -5.0; o.hi[i] = 5.0; }
  st = apgp_predict(h, Xq, Q + N, mu, var, util, &o, 1);
  if (st != APGP_OK) {
    fprintf(stderr, "apgp_predict: %s\n", apgp_last_error());
    return 1;
  }
  for (i = 0; i < N; ++i) {
    const double e = fabs(mu[Q + i] - y[i]);
    if (e > err) err = e;
    if (!(var[Q + i] >= 0.0 && var[Q + i] < 1e-3)) { fprintf(stderr, "variance at a training point: %g\n", var[Q + i]); return 1; }
  }
  for (i = 0; i < Q; ++i)
    if (!(var[i] >= 0.0) || mu[i] != mu[i]) { fprintf(stderr, "query %d: mu %g var %g\n", i, mu[i], var[i]); return 1; }
  printf("log|K| = %.6f, log-likelihood = %.6f, max |mu(x_i) - y_i| = %.3g, kernels launched = %lld\n", logdet, loglik, err,
         apgp_launch_count(h));
  apgp_destroy(h);
  /* white noise e^-12 times |alpha| <= ~1e2 bounds the misfit at the training points; the oracle gives 9.0e-4 */
  if (!(err < 1e-2)) { fprintf(stderr, "the GP does not interpolate its training set\n"); return 1; }
  if (!(fabs(loglik - EXPECT_LOGLIK) <= 1e-6 * fabs(EXPECT_LOGLIK))) {
    fprintf(stderr, "log-likelihood %.12f, expected %.12f\n", loglik, EXPECT_LOGLIK);
    return 1;
  }
  printf("c_caller ok\n");
  return 0;
}
