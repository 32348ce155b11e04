"""bench.py's reference arm (the reference's CPU implementation of the path timed on the host cores) runs without a GPU:
check the JSON contract the driver parses - one line, the measured arm's metric / unit / config, `impl`, `cpu_baseline`, `e2e` -
and that ranks other than 0 exit quietly."""
import json
import os
import subprocess
import sys

import vst_b200  # noqa: F401  (puts the repo root on sys.path via conftest)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reconet_1080p_infer_frames_per_s" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1.0          # one frame per step
    assert "1920x1080" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    from oracle import ref_loader

    assert cb["kind"] == ref_loader.find_root(prefer_copy=True)[1] and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
