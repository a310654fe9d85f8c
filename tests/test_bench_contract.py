"""bench.py's reference arm (the CPU implementation of the path, timed on the host cores) must print ONE JSON line with the
contract's keys — CPU; also under a torchrun-style environment where only rank 0 may print."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))


def test_reference_arm_prints_the_contract_line():
    out = _run({})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bpr_train_samples_per_s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""
