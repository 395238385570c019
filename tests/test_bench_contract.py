"""bench.py's reference arm (the CPU oracle timed on the host cores) on a small workload: one JSON line with the keys the
driver reads, `impl: reference`, and -- under a multi-rank launch -- only rank 0 prints.  No GPU needed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--impl", "reference", "--steps", "2", "--warmup", "1", "--n-map", "30000", "--n-scan", "3000", "--iters", "5"]


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return [ln for ln in p.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "scans/s" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["value"] == d["cpu_baseline"]["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []
