"""bench.py contract of the reference (CPU) arm, exercised on a toy model so that it runs in seconds: one JSON line
with the keys the driver reads; non-zero ranks print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--tiny", "--steps", "2",
                           "--warmup", "1", "--gpus", "1", *extra], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_contract_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "model", "deberta.py")) or \
        os.path.isfile("/root/reference/model/deberta.py")
    # the UNMODIFIED reference when its model files are available (baseline/install_ref.py), else the oracle port
    assert cb["kind"] == ("reference" if staged else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_other_workloads():
    for wl, unit in (("mc", "questions/s"), ("bert", "samples/s")):
        r = _run({"RANK": "0", "WORLD_SIZE": "1"}, "--workload", wl)
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([l for l in r.stdout.splitlines() if l.strip().startswith("{")][0])
        assert d["impl"] == "reference" and d["unit"] == unit and d["value"] > 0 and d["cpu_baseline"]["sample"]
