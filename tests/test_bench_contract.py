"""bench.py's reference arm runs on the CPU (it times the reference's own implementation), so the shape of
the JSON line can be checked without a GPU on a tiny scene."""
import glob
import json
import os
import subprocess
import sys

import pytest

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
    assert line["steps"] >= 10, "the reference arm times at least ten steps"
    ran = line["executed_iterations"]
    assert 1 <= ran[0] <= 20 and 1 <= ran[1] <= 20 and line["executed_constraint_iterations_per_sec"] > 0


def test_non_zero_ranks_of_the_reference_arm_stay_silent(ref):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--scene", "pyramid_10"],
                                  cwd=ROOT, env=env, timeout=120).decode().strip()
    assert out == ""


@pytest.mark.gpu
def test_gpu_arm_prints_one_recomputable_contract_line():
    """bench.py's own arm on a small scene: every key of the round contract, and numbers that can be recomputed from the
    line's own fields (value from joints / time; roofline.frac from the SURVEY 8(d) bytes, kernel time and peak)."""
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--scene", "pyramid_10k", "--steps", "4", "--warmup", "3", "--settle", "5",
                                   "--no-parity"], cwd=ROOT, timeout=600).decode().strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "clocks", "roofline", "roofline_radix", "cpu_baseline", "broadphase_pairs_per_sec",
                "executed_constraint_iterations_per_sec", "relaxed_constraint_iterations_per_sec"):
        assert key in line, key
    assert line["metric"] == "constraint_iterations_per_sec" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["dtype"] == "f32" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert line["warmup"] >= 3 and line["gpu_launches"] > 0 and "workload" in line["config"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] < line["value"]
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "formula", "kernel_forms_launched"):
        assert key in roof, key
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    # SURVEY 8(d) only: joints x 128 + relaxed impulse x 196 + relaxed displacement x 136, over the kernel time
    bytes_ = roof["joints"] * 128 + roof["active_joint_iterations"][0] * 196 + roof["active_joint_iterations"][1] * 136
    assert abs(bytes_ / roof["algorithmic_bytes_per_launch"] - 1) < 1e-6
    assert abs(bytes_ / (roof["kernel_ms"] * 1e-3) / 1e9 / roof["achieved"] - 1) < 1e-6
    assert roof["kernel_forms_launched"] == {"3": 4}, "the default iteration kernel of the resident pipeline is the strip-local one"
    # value = joints x 40 / time of the timed steps
    assert abs(line["value"] * line["ms_per_step"] * 1e-3 / 40 / roof["joints"] - 1) < 0.02
    clocks = line["clocks"]
    assert clocks["sm_mhz"] and clocks["sm_max_mhz"]
    cpu = line["cpu_baseline"]
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in cpu, key
