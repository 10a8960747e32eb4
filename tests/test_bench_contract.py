"""bench.py contract on the CPU: the reference arm (`--impl reference`: the unmodified reference's own
Cython generator from oracle/_ref on one host core, the oracle's OpenMP restatement when the reference is
not installed) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, PYLBM_B200_CPU_BUDGET_S="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "3", "--workload", "d3q19_lid_256"], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "MLUPS" and line["unit"] == "MLUPS"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["steps"] == 2 and line["warmup"] == 3 and line["n_gpus"] == 1
    installed = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "pylbm"))
    assert line["cpu_baseline"]["kind"] == ("reference" if installed else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    if installed:
        assert line["cpu_baseline"]["cores"] == 1 and "generator='cython'" in line["cpu_baseline"]["sample"]
        assert line["cpu_port"]["kind"] == "port"
    # the same `config` keys as the device arm prints (the driver compares the two dictionaries)
    assert set(line["config"]) == {"workload", "case", "n", "storage", "arithmetic", "l2"}
    assert set(line["run"]) == {"api", "parallelism", "halo", "setup_s"}
    assert line["cpu_baseline"]["value"] == line["value"] and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
