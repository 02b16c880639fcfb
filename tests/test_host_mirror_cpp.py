"""Source compatibility of the host mirror: a program written the way the reference's demo uses the
physics API compiles against phyx_b200/host unchanged (CPU), and — on a GPU — the same source built
against the REFERENCE's headers prints the same line."""
import os
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "host_api_demo.cpp")
HOST = os.path.join(ROOT, "phyx_b200", "host")


def build_against_mirror(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "phyx_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", HOST])
    exe = str(tmp_path / "demo_mirror")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", f"-I{HOST}", f"-I{os.path.join(ROOT, 'include')}", SRC,
                           os.path.join(HOST, "phyx_host.cpp"), f"-L{os.path.join(ROOT, 'phyx_b200')}", "-lphyx_b200",
                           f"-Wl,-rpath,{os.path.join(ROOT, 'phyx_b200')}", "-lpthread", "-o", exe])
    return exe


def test_demo_style_program_compiles_against_the_mirror(tmp_path):
    assert os.path.exists(build_against_mirror(tmp_path))


@pytest.mark.gpu
def test_demo_style_program_matches_the_reference_build(tmp_path):
    exe = build_against_mirror(tmp_path)
    ours = subprocess.check_output([exe, "12", "40"]).decode().strip()
    want = open(os.path.join(ROOT, "tests", "golden", "host_api_demo_12_40.txt")).read().strip()
    assert ours == want
