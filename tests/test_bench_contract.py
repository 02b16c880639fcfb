"""bench.py's reference arm runs on the CPU (it times the reference's own implementation), so the shape of
the JSON line can be checked without a GPU on a tiny scene."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line(ref):
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scene", "pyramid_1k", "--steps", "3",
                                   "--warmup", "1", "--settle", "2"], cwd=ROOT, timeout=300).decode().strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "constraint_iterations_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["vs_baseline"] is None


def test_non_zero_ranks_of_the_reference_arm_stay_silent(ref):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--scene", "pyramid_10"],
                                  cwd=ROOT, env=env, timeout=120).decode().strip()
    assert out == ""
