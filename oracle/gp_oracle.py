"""NumPy/SciPy fp64 restatement of the george.GP surface approxposterior calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Model (george 0.3.x semantics, anchored on the reference call sites):
  kernel      k(x, x') = A * exp(-0.5 * sum_i (x_i - x'_i)^2 / M_i)
              ``ExpSquaredKernel(metric=M, ndim)``: M_i is a *squared* length
              scale, parameters are log M_i  (reference gpUtils.py:160).
              ``a * kernel`` stores log_constant = log(a / ndim) and evaluates
              to A = ndim * exp(log_constant)            (reference gpUtils.py:164-165;
              pinned by tests/test_InitGP.py:43 and tests/test_GPUtil.py:50).
  mean        constant, fitted                            (reference gpUtils.py:176)
  white noise frozen log-variance, added to diag(K)       (reference gpUtils.py:176-177)
  parameter vector  [mean, (log_constant), log M_0 .. log M_{d-1}]
                                                          (tests/test_InitGP.py:43,76)
"""
import numpy as np
from scipy.linalg import cholesky, cho_solve

TINY = 1.25e-12  # george's default yerr; TINY**2 is added to the diagonal (numerically a no-op)


class GPOracle(object):
    """Duck type of george.GP restricted to ExpSquared (x constant) + constant mean."""

    def __init__(self, ndim, metric, amp=None, mean=0.0, white_noise=-12.0):
        self.ndim = int(ndim)
        metric = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if metric.size == 1 and self.ndim > 1:
            metric = np.full(self.ndim, float(metric[0]))
        assert metric.size == self.ndim and np.all(metric > 0)
        self.log_M = np.log(metric)
        self.fit_amp = amp is not None
        # george: a*kernel -> ConstantKernel(log_constant=log(a/ndim))
        self.log_c = np.log(float(amp) / self.ndim) if self.fit_amp else None
        self.mean = float(mean)
        self.white_noise = float(white_noise)
        self.computed = False
        self._dirty = True
        self._x = None
        self._L = None
        self._alpha = None
        self._alpha_y = None

    # ---- parameter bookkeeping (george.modeling.ModelSet) --------------------
    def get_parameter_names(self):
        names = ["mean:value"]
        if self.fit_amp:
            names.append("kernel:k1:log_constant")
            pre = "kernel:k2:metric:"
        else:
            pre = "kernel:metric:"
        names += [pre + "log_M_%d_%d" % (i, i) for i in range(self.ndim)]
        return tuple(names)

    def get_parameter_vector(self):
        p = [self.mean]
        if self.fit_amp:
            p.append(self.log_c)
        return np.concatenate([p, self.log_M]).astype(np.float64)

    def set_parameter_vector(self, p):
        p = np.asarray(p, dtype=np.float64).ravel()
        if p.size != len(self):
            raise ValueError("dimension mismatch")
        self.mean = float(p[0])
        k = 1
        if self.fit_amp:
            self.log_c = float(p[1])
            k = 2
        self.log_M = p[k:].copy()
        self._dirty = True
        self._alpha = None

    def __len__(self):
        return 1 + (1 if self.fit_amp else 0) + self.ndim

    @property
    def amplitude(self):
        return self.ndim * np.exp(self.log_c) if self.fit_amp else 1.0

    # ---- kernel --------------------------------------------------------------
    def _kernel(self, x1, x2):
        with np.errstate(over="ignore"):
            w = np.exp(-self.log_M)                      # 1 / M_i
        d2 = np.zeros((x1.shape[0], x2.shape[0]))
        for i in range(self.ndim):
            diff = x1[:, i][:, None] - x2[:, i][None, :]
            d2 += diff * diff * w[i]
        return self.amplitude * np.exp(-0.5 * d2)

    def _parse(self, t):
        t = np.asarray(t, dtype=np.float64)
        if t.ndim == 0:
            t = t.reshape(1, 1)
        elif t.ndim == 1:
            t = t.reshape(-1, 1)
        if t.ndim != 2 or t.shape[1] != self.ndim:
            raise ValueError("dimension mismatch")
        return np.ascontiguousarray(t)

    # ---- factorisation (george BasicSolver: scipy cholesky + cho_solve) -------
    def compute(self, x, yerr=TINY):
        self._x = self._parse(x)
        self._yerr2 = float(yerr) ** 2
        self._dirty = True
        self.recompute()

    def recompute(self, quiet=False):
        if not self._dirty and self.computed:
            return True
        if self._x is None:
            raise RuntimeError("You need to compute the model first")
        try:
            K = self._kernel(self._x, self._x)
            K[np.diag_indices_from(K)] += self._yerr2 + np.exp(self.white_noise)
            if not np.all(np.isfinite(K)):
                raise np.linalg.LinAlgError("non-finite covariance")
            self._L = cholesky(K, lower=True, overwrite_a=True, check_finite=False)
            self.log_determinant = 2.0 * np.sum(np.log(np.diag(self._L)))
            self._const = -0.5 * (self._x.shape[0] * np.log(2.0 * np.pi) + self.log_determinant)
        except (ValueError, np.linalg.LinAlgError):
            self.computed = False
            if quiet:
                return False
            raise
        self.computed = True
        self._dirty = False
        self._alpha = None
        return True

    def apply_inverse(self, b):
        return cho_solve((self._L, True), b, check_finite=False)

    def _compute_alpha(self, y):
        if self._alpha is not None and self._alpha_y is y:      # george caches alpha per y as well
            return self._alpha
        yv = np.asarray(y, dtype=np.float64).ravel()
        if yv.size != self._x.shape[0]:
            raise ValueError("dimension mismatch")
        alpha = self.apply_inverse(yv - self.mean)
        if isinstance(y, np.ndarray) and not y.flags.writeable:
            self._alpha, self._alpha_y = alpha, y                 # only immutable targets are safe to key by identity
        return alpha

    # ---- george.GP.predict(y, t, return_cov=False, return_var=...) ------------
    def predict(self, y, t, return_cov=False, return_var=False):
        if return_cov:
            raise NotImplementedError("only the forms approxposterior uses are restated")
        self.recompute()
        xs = self._parse(t)
        alpha = self._compute_alpha(y)
        Kxs = self._kernel(xs, self._x)
        mu = Kxs @ alpha + self.mean
        if not return_var:
            return mu
        KinvKxs = self.apply_inverse(Kxs.T)
        var = np.full(xs.shape[0], self.amplitude)          # k** carries no white noise
        var -= np.sum(Kxs.T * KinvKxs, axis=0)
        return mu, var

    # ---- george.GP.log_likelihood / grad_log_likelihood -----------------------
    def log_likelihood(self, y, quiet=False):
        if not self.recompute(quiet=quiet):
            return -np.inf
        r = np.asarray(y, dtype=np.float64).ravel() - self.mean
        ll = self._const - 0.5 * np.dot(r, self.apply_inverse(r))
        return ll if np.isfinite(ll) else -np.inf

    def grad_log_likelihood(self, y, quiet=False):
        if not self.recompute(quiet=quiet):
            return np.zeros(len(self))
        alpha = self._compute_alpha(y)
        N = self._x.shape[0]
        Kinv = self.apply_inverse(np.eye(N))
        A = np.outer(alpha, alpha) - Kinv
        Kk = self._kernel(self._x, self._x)                 # noise-free
        g = [np.sum(alpha)]                                 # d/d mean
        if self.fit_amp:
            g.append(0.5 * np.sum(A * Kk))                  # dK/dlog_c = K_kernel
        w = np.exp(-self.log_M)
        for i in range(self.ndim):
            diff = self._x[:, i][:, None] - self._x[:, i][None, :]
            g.append(0.5 * np.sum(A * Kk * (0.5 * diff * diff * w[i])))
        return np.asarray(g)


def default_gp_oracle(theta, y, white_noise=-12, fitAmp=False, rng=np.random):
    """Restates reference gpUtils.defaultGP (gpUtils.py:114-181) on the oracle GP.

    RNG consumption: one ``randn(ndim)`` for the initial metric (gpUtils.py:156).
    """
    theta = np.asarray(theta).squeeze()
    y = np.asarray(y).squeeze()
    ndim = 1 if theta.ndim <= 1 else theta.shape[-1]
    initialMetric = np.fabs(rng.randn(ndim))
    gp = GPOracle(ndim, initialMetric, amp=(np.var(y) if fitAmp else None),
                  mean=np.median(y), white_noise=white_noise)
    gp.compute(theta)
    return gp
