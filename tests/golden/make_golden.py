"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

The reference (zeux/phyx @ 327b6c96) ships no tests or golden vectors, so the pins
are produced here: the unmodified reference, compiled by oracle/Makefile into
oracle/_ref/libphyx_ref_strict.so (strict IEEE build: -O2 -ffp-contract=off, asserts on), is
stepped through its public stage functions and its inputs/outputs are recorded.

Run (in the authoring container, where /root/reference is mounted):
    make -C oracle ref && python tests/golden/make_golden.py

Outputs (committed):
    solve_<scene>_s<step>.npz   bodies/joints/contact points BEFORE Solver::SolveJoints and the
                                reference's outputs for Solve_AVX2 / Solve_SSE2 / Solve_Scalar
                                (Island_Single, 20+20 iterations) plus the joint order used
    stages_<scene>_s<step>.npz  per-stage captures: IntegrateVelocity, UpdateBroadphase, the full
                                overlapping-pair list, IntegratePosition
    radix.npz                   radixFloat / radixSort3 known answers
    trajectory.json             FNV-1a hashes of the body state after steps {1,10,100} per scene
                                and mode (detects oracle/compiler drift)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refpy  # noqa: E402
from phyx_b200 import scenes, types as T  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DT = 1.0 / 60.0


def fnv1a(data: bytes) -> str:
    h = 0xCBF29CE484222325
    for chunk in np.frombuffer(data, dtype=np.uint8).reshape(-1):
        h = ((h ^ int(chunk)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def state_bytes(bodies):
    """Body state used for hashes: pos, basis, velocities (bit patterns)."""
    cols = [bodies[f].reshape(bodies.shape[0], -1) for f in ("pos", "xVector", "yVector", "velocity", "angularVelocity")]
    return np.ascontiguousarray(np.concatenate(cols, axis=1)).tobytes()


def capture(scene_name, steps):
    w = refpy.RefWorld(scenes.make(scene_name), "strict")
    for step in range(max(steps) + 1):
        if step in steps:
            b_start = w.bodies()
            w.step_staged(mask=0x01)  # IntegrateVelocity
            b_vel = w.bodies()
            w.step_staged(mask=0x3E | refpy.SAFE_PAIRS)  # broadphase .. RefreshContactJoints
            entries = w.broadphase()
            pairs = refpy.all_pairs(b_vel)
            b0, j0, cp = w.bodies(), w.joints(), w.contact_points()
            out = {"bodies": b0, "joints": j0, "contact_points": cp}
            for tag, mode in (("avx2", T.SOLVE_AVX2), ("sse2", T.SOLVE_SSE2), ("scalar", T.SOLVE_SCALAR)):
                rb, rj, idx = refpy.solve_joints(b0, j0, cp, solve=mode)
                out[f"bodies_{tag}"] = rb
                out[f"joints_{tag}"] = rj
                out[f"order_{tag}"] = idx
            np.savez_compressed(os.path.join(OUT, f"solve_{scene_name}_s{step}.npz"), **out)
            w.step_staged(mask=0x40)  # SolveJoints (AVX2)
            b_solved = w.bodies()
            w.step_staged(mask=0x80)  # IntegratePosition
            b_end = w.bodies()
            np.savez_compressed(
                os.path.join(OUT, f"stages_{scene_name}_s{step}.npz"),
                bodies_start=b_start,
                bodies_after_velocity=b_vel,
                broadphase=entries,
                pairs=pairs,
                bodies_solved=b_solved,
                bodies_end=b_end,
            )
        else:
            w.step()


def trajectories():
    res = {}
    for scene_name in ("pyramid_10", "pyramid_1k", "stack_1k", "islands_8x10"):
        for tag, mode in (("avx2", T.SOLVE_AVX2), ("sse2", T.SOLVE_SSE2), ("scalar", T.SOLVE_SCALAR)):
            w = refpy.RefWorld(scenes.make(scene_name), "strict")
            rec = {}
            for step in range(1, 101):
                w.step(solve=mode)
                if step in (1, 10, 100):
                    rec[str(step)] = {
                        "hash": fnv1a(state_bytes(w.bodies())),
                        "joints": int(len(w.joints())),
                        "manifolds": int(len(w.manifolds())),
                    }
            res[f"{scene_name}/{tag}"] = rec
    with open(os.path.join(OUT, "trajectory.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


def radix():
    vals = np.array(
        [0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 3.5, -3.5, 1e7, -1e7, np.inf, -np.inf, 482.5, -482.49866, 2.0**-149, -(2.0**-149)],
        dtype=np.float32,
    )
    keys = np.array([refpy.radix_float(v) for v in vals], dtype=np.uint32)
    # deterministic pseudo-random keys with many ties (LCG, no RNG module)
    x = np.zeros(5000, dtype=np.uint64)
    s = 12345
    for i in range(x.shape[0]):
        s = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        x[i] = (s >> 33) % (1 << 32) if i % 3 else (s >> 33) % 64
    x = x.astype(np.uint32)
    np.savez_compressed(os.path.join(OUT, "radix.npz"), floats=vals, keys=keys, sort_in=x, sort_out=refpy.radix_sort3(x))


if __name__ == "__main__":
    assert refpy.available("strict"), "build oracle/_ref first: make -C oracle ref"
    capture("pyramid_10", (0, 5, 30))
    capture("stack_1k", (0, 40))
    capture("pyramid_1k", (30,))
    trajectories()
    radix()
    print("golden fixtures written to", OUT)

# host_api_demo_12_40.txt: tests/cpp/host_api_demo.cpp compiled against the REFERENCE's headers and sources
# (same flags as the strict oracle build) and run as `demo 12 40`:
#   REF=/root/reference/src; g++ -O2 -ffp-contract=off -std=c++11 -mavx2 -mfma -include cstdio -include cstring \
#     -include cassert -DMICROPROFILE_ENABLED=0 '-DMicroProfileOnThreadExit()=do{}while(0)' -I$REF -I$REF/microprofile \
#     tests/cpp/host_api_demo.cpp $REF/World.cpp $REF/Solver.cpp $REF/Collider.cpp $REF/base/WorkQueue.cpp -lpthread -o demo_ref
