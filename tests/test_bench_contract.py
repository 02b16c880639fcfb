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


def test_non_zero_ranks_of_the_reference_arm_stay_silent(ref):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--scene", "pyramid_10"],
                                  cwd=ROOT, env=env, timeout=120).decode().strip()
    assert out == ""


def _recorded(pattern):
    return sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))


@pytest.mark.parametrize("path", _recorded("bench_r1j.json") + _recorded("bench_*_r1i.json") + _recorded("bench_*gpu_spanning_r1h.json"),
                         ids=os.path.basename)
def test_recorded_bench_lines_keep_the_contract(path):
    """The bench lines committed under profiles/ (written by bench.py on a B200) carry every key of the round
    contract, with consistent values."""
    line = json.load(open(path))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["metric"] == "constraint_iterations_per_sec" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["dtype"] == "f32" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert line["warmup"] >= 3 and line["gpu_launches"] > 0 and "workload" in line["config"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] < line["value"]
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roof, key
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    clocks = line["clocks"]
    assert clocks["sm_mhz"] and clocks["sm_max_mhz"] and not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value = joints x 40 x ranks / time of the timed steps
    assert abs(line["value"] * line["ms_per_step"] * 1e-3 / (line["n_gpus"] * 40) / line["roofline"]["joints"] - 1) < 0.02
    if line["cpu_baseline"] is not None:
        for key in ("value", "unit", "cores", "kind", "sample"):
            assert key in line["cpu_baseline"], key
    if line["n_gpus"] > 1:
        assert line["spanning"]["replicas_identical"] is True and line["spanning"]["ranks"] == line["n_gpus"]
