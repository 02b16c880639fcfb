"""CPU-side checks of the C ABI: the library builds/loads, exports exactly what include/phyx_b200.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT
from phyx_b200 import capi


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "phyx_b200", "csrc")])
    return capi.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "phyx_b200.h")).read()
    return sorted(set(re.findall(r"PHYX_B200_API\s+[\w\s\*]+?\b(phyx_b200_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH]).decode()
    exported = set(re.findall(r" T (phyx_b200_\w+)", out))
    assert exported == set(header_symbols())


def test_struct_sizes_match_header():
    assert C.sizeof(capi.SolveConfig) == 16
    assert C.sizeof(capi.SolveStats) == 8 * 4 + 2 * 8 + 7 * 4 + 4  # 4 B tail padding (int64 alignment)
    assert C.sizeof(capi.BroadphaseStats) == 8 + 8 + 12 + 4  # 3 floats + tail padding


def test_struct_layouts_match_the_compiled_header(tmp_path):
    """sizeof / offsetof of every struct that crosses the ABI, from the header itself (gcc), against the ctypes mirrors and the
    host mirror's use of them: a field added on one side only would smash a caller's stack."""
    structs = {"phyx_b200_solve_config": capi.SolveConfig, "phyx_b200_solve_stats": capi.SolveStats,
               "phyx_b200_broadphase_stats": capi.BroadphaseStats, "phyx_b200_step_info": capi.StepInfo}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "phyx_b200.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in ct._fields_:
            lines.append(f'printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for field, _ in ct._fields_:
            assert int(got[f"{cname}.{field}"]) == getattr(ct, field).offset, f"{cname}.{field}"


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "phyx_b200.h")).read()
    for cite in ("src/World.cpp:39-70", "src/Collider.cpp:251-284", "src/Collider.cpp:296-366", "src/Solver.cpp:17-119", "src/RigidBody.h:12-58"):
        assert cite in text


def test_no_cpu_fallback_without_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    status = lib.phyx_b200_create(0, C.byref(h))
    assert status == 3 and not h.value  # PHYX_B200_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.phyx_b200_last_error()
    with pytest.raises(capi.PhyxError):
        capi.Context(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under phyx_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "phyx_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|phyx_oracle|libphyx_ref", src):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
