"""Kernel objects mirroring the slice of ``george.kernels`` that approxposterior builds.

Reference: ``gpUtils.defaultGP`` (gpUtils.py:160-173) constructs
``ExpSquaredKernel(metric=initialMetric, ndim=ndim)``, optionally ``np.var(y) * kernel``;
``order != None`` adds a ``LinearKernel`` which is out of scope for the B200 engine.

george semantics kept here:
  * ``ExpSquaredKernel``: k = exp(-0.5 * sum_i dx_i^2 / M_i); parameters are ``log M_i``.
  * ``a * kernel`` -> ``ConstantKernel(log_constant=log(a/ndim)) * kernel`` whose value is
    ``ndim * exp(log_constant)`` (pinned by reference tests/test_InitGP.py:43, test_GPUtil.py:50).
"""
import numpy as np

__all__ = ["ExpSquaredKernel", "ConstantKernel", "Product", "LinearKernel"]


class _Kernel(object):
    def __rmul__(self, a):
        return Product(ConstantKernel(log_constant=np.log(float(a) / self.ndim), ndim=self.ndim), self)

    __mul__ = __rmul__

    def __add__(self, other):
        raise NotImplementedError("kernel sums (e.g. + LinearKernel, gpUtils.py:169-173) are outside the "
                                  "B200 engine's scope: ExpSquared (x constant) only")

    __radd__ = __add__

    def __len__(self):
        return len(self.get_parameter_vector())


class ExpSquaredKernel(_Kernel):
    def __init__(self, metric, ndim=1):
        self.ndim = int(ndim)
        metric = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if metric.size == 1 and self.ndim > 1:
            metric = np.full(self.ndim, float(metric[0]))
        if metric.ndim != 1 or metric.size != self.ndim:
            raise ValueError("only axis-aligned metrics are supported (vector of length ndim)")
        if np.any(metric <= 0):
            raise ValueError("metric must be positive")
        self.log_M = np.log(metric)

    def get_parameter_names(self):
        return tuple("metric:log_M_%d_%d" % (i, i) for i in range(self.ndim))

    def get_parameter_vector(self):
        return self.log_M.copy()

    def set_parameter_vector(self, v):
        v = np.asarray(v, dtype=np.float64).ravel()
        if v.size != self.ndim:
            raise ValueError("dimension mismatch")
        self.log_M = v.copy()

    # pieces the engine consumes
    @property
    def amplitude(self):
        return 1.0

    @property
    def fit_amp(self):
        return False


class ConstantKernel(_Kernel):
    def __init__(self, log_constant, ndim=1):
        self.ndim = int(ndim)
        self.log_constant = float(log_constant)

    def get_parameter_names(self):
        return ("log_constant",)

    def get_parameter_vector(self):
        return np.array([self.log_constant])

    def set_parameter_vector(self, v):
        self.log_constant = float(np.asarray(v).ravel()[0])


class Product(_Kernel):
    """``ConstantKernel * ExpSquaredKernel`` -- the only product approxposterior builds."""

    def __init__(self, k1, k2):
        if not (isinstance(k1, ConstantKernel) and isinstance(k2, ExpSquaredKernel)):
            raise NotImplementedError("only ConstantKernel * ExpSquaredKernel is supported")
        self.k1, self.k2 = k1, k2
        self.ndim = k2.ndim

    def get_parameter_names(self):
        return tuple("k1:" + n for n in self.k1.get_parameter_names()) + \
            tuple("k2:" + n for n in self.k2.get_parameter_names())

    def get_parameter_vector(self):
        return np.concatenate([self.k1.get_parameter_vector(), self.k2.get_parameter_vector()])

    def set_parameter_vector(self, v):
        v = np.asarray(v, dtype=np.float64).ravel()
        self.k1.set_parameter_vector(v[:1])
        self.k2.set_parameter_vector(v[1:])

    @property
    def log_M(self):
        return self.k2.log_M

    @property
    def amplitude(self):
        return self.ndim * np.exp(self.k1.log_constant)

    @property
    def fit_amp(self):
        return True


def LinearKernel(*args, **kwargs):
    raise NotImplementedError("LinearKernel (defaultGP(order=...), gpUtils.py:169-173) is outside the B200 "
                              "engine's scope")
