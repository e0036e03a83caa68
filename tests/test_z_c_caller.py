"""A non-Python caller of the C-ABI: examples/c_caller.c is compiled as strict C99 against include/apgp.h (the header
must be valid C, not only C++), linked with libapgp.so and executed.  Without a GPU the program must fail loudly
(no CPU fallback); on a B200 it runs factorise + fused predict from host buffers and checks its results against the
oracle's known answer, which the CPU half recomputes here.  (The file sorts last on purpose: the driver runs the suite
with -x, and a build-environment problem of this stand-alone program must not cut the kernel parity tests short.)"""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "c_caller.c")


def _build(tmp_path):
    exe = str(tmp_path / "c_caller")
    libdir = os.path.join(ROOT, "approxposterior_b200")
    if not os.path.isfile(os.path.join(libdir, "libapgp.so")):
        pytest.fail("libapgp.so is not built: run __graft_entry__.build()")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
                    "-L", libdir, "-lapgp", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True, capture_output=True)
    return exe


def _env():
    """libapgp.so names libcudart.so.12 without a run path (a Python process has it loaded already or finds it through
    LD_LIBRARY_PATH): give the stand-alone program the same directories explicitly."""
    env = dict(os.environ)
    dirs = [d for d in env.get("LD_LIBRARY_PATH", "").split(":") if d]
    try:
        import nvidia.cuda_runtime as rt
        dirs.append(os.path.join(os.path.dirname(rt.__file__), "lib"))
    except Exception:
        pass
    dirs.append("/usr/local/cuda/lib64")
    env["LD_LIBRARY_PATH"] = ":".join(dirs)
    return env


def _lcg_inputs():
    """The generator of examples/c_caller.c."""
    state = [0x9E3779B97F4A7C15]

    def uniform(lo, hi):
        state[0] = (state[0] * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
        return lo + (hi - lo) * float(state[0] >> 11) / 9007199254740992.0
    X = np.zeros((100, 2))
    for i in range(100):
        X[i, 0] = uniform(-5.0, 5.0)
        X[i, 1] = uniform(-5.0, 5.0)
    y = -0.125 * np.sum(X * X, axis=1)
    return X, y


def test_c_caller_known_answer_is_the_oracles():
    from oracle import GPOracle
    X, y = _lcg_inputs()
    g = GPOracle(2, np.exp([1.0, 1.0]), mean=-3.0, white_noise=-12.0)
    g.compute(X)
    want = float(g.log_likelihood(y))
    const = float(re.search(r"#define EXPECT_LOGLIK\s+([0-9.eE+-]+)", open(SRC).read()).group(1))
    assert abs(const - want) <= 1e-9 * abs(want)
    mu, var = g.predict(y, X, return_var=True)
    assert np.abs(mu - y).max() < 1e-2 and var.max() < 1e-3          # the margins the C program checks with


def test_c_caller_compiles_as_c99_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked run")
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=_env())
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_caller_runs_on_the_gpu(tmp_path):
    try:
        exe = _build(tmp_path)
    except (OSError, subprocess.CalledProcessError) as e:          # no gcc / linker cannot resolve the CUDA runtime
        pytest.skip("cannot build a stand-alone C program here: %s" % (getattr(e, "stderr", b"") or e))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=_env())
    if r.returncode == 127 or "error while loading shared libraries" in r.stderr:
        pytest.skip("loader cannot resolve the CUDA runtime outside Python: " + r.stderr.strip())
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "c_caller ok" in r.stdout
