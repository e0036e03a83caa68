"""bench.py's CPU-runnable legs honour the driver's JSON contract (the `--impl reference` arm of every configuration is
what the driver runs beside the GPU arm): one line on stdout, the same metric / unit / config as the GPU arm, a
`cpu_baseline` describing the run, zero-byte `e2e`, and all host cores used whatever the launcher exported."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("cfg,metric_has", [("cfg3", "N=2048, d=5"), ("cfg5", "N=512, d=5"), ("cfg2", "65536 walkers")])
def test_reference_arm_line(cfg, metric_has):
    # torchrun exports OMP_NUM_THREADS=1: the arm must not inherit it (round 1's N > 1 ratios were void because it did)
    d = _run(["--config", cfg, "--impl", "reference", "--steps", "1", "--warmup", "1"], {"OMP_NUM_THREADS": "1"})
    assert d["impl"] == "reference" and metric_has in d["metric"] and d["value"] > 0
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    if cfg in ("cfg3", "cfg5", "cfg2"):                       # cfg2: one ensemble per core, side by side
        assert cb["cores"] == (os.cpu_count() or 1)
    if cfg == "cfg3":                                         # the driver's reference record carries both halves of the metric
        b = d["bape_iteration"]
        assert "unavailable" not in b, b
        assert b["value"] > 0 and b["higher_is_better"] is False and b["kind"] in ("reference", "port")


def test_other_ranks_of_the_reference_arm_stay_silent():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--gpus", "2"],
                         capture_output=True, text=True, env=dict(os.environ, RANK="1", WORLD_SIZE="2"), timeout=300, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
