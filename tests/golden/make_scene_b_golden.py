#!/usr/bin/env python
"""Generates tests/golden/scene_b_attempts.{npz,json}: 8 whole grasp attempts into the 40-object pile of the reference's default
scene (UR5gripper_2_finger_many_objects.xml) on the fp64 CPU oracle.

TEST INFRASTRUCTURE.  Env i: reset by the live rule of GraspingEnv.py:418-430 with RandomState(20000 + i), 500 sub-steps for the
pile to form, then `move_and_grasp` (GraspingEnv.py:205-386) at the object nearest to the table centre with rotation i % 6.
The state before the attempt is stored, so that a replayer (tests/test_scale_parity_gpu.py) runs the same attempt from the
same state at any env index of a large batch.

Contact-rich attempts can hinge on rounding (a finger that just catches or just misses an object edge): r02c's build, whose tree
Cholesky rounds differently, closed the gripper in 173 instead of 301 sub-steps on one of the first fixture's records - on all 128
copies alike.  A record is therefore kept only if its outcome (reward and all 12 phase counters) is the same on the oracle, on the
oracle with one object coordinate of the start state moved by one unit in the last place (up and down), and on the oracle built
with -O3 -ffp-contract=fast: 16 candidate environments, the first 8 robust ones are kept.

  python tests/golden/make_scene_b_golden.py      (about a minute on 8 cores: the oracle does ~100 sub-steps/s on this scene)
"""
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "scene_b_attempts")
N = 8
CANDIDATES = 16


def attempt(args):
    """(env i, variant): variant 0 baseline, +1 / -1 = one-ulp perturbation of the first object's x before the attempt; the process
    environment selects the oracle build (GRASP_ORACLE_SO)"""
    i, variant, so = args
    if so:
        os.environ["GRASP_ORACLE_SO"] = so
    from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob
    from oracle.oracle_py import OracleEnv
    from tests.common import HOME, reset_qpos_scene_b

    blob = load_scene_blob("B")
    A, _ = load_scene("B")
    o = OracleEnv(blob)
    o.reset(reset_qpos_scene_b(A, i))
    o.move_group("All", HOME + np.array([0, 0, 0, 0, 0, 0, 0.001]), 1e-9, 499)
    q, v = o.qpos.copy(), o.qvel.copy()
    pos = q[8:288].reshape(40, 7)[:, :3]
    on_table = (pos[:, 2] > 0.85) & (pos[:, 2] < 1.0)
    k = int((np.linalg.norm(pos[:, :2] - np.array([0.0, -0.6]), axis=1) + (~on_table) * 10).argmin())
    coords = np.array([pos[k, 0], pos[k, 1], pos[k, 2] + 0.02])
    q2 = q.copy()
    if variant:
        q2[8] = np.nextafter(q2[8], np.inf if variant > 0 else -np.inf)
    o.reset(q2, v)
    s0 = o.substeps
    r, info = o.move_and_grasp(coords, i % 6, 0.91)
    rec = {"env": i, "rot": i % 6, "coords": [float(t) for t in coords], "reward": int(r), "info": [int(t) for t in info],
           "substeps": int(o.substeps - s0), "arm_qpos": [float(t) for t in o.qpos[:8]], "_q0": q, "_v0": v, "_q1": o.qpos.copy()}
    o.close()
    return rec


def main():
    import subprocess

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "fast"])
    fast = os.path.join(ROOT, "oracle", "libgrasp_oracle_fast.so")
    jobs = [(i, var, so) for i in range(CANDIDATES) for var, so in ((0, None), (1, None), (-1, None), (0, fast))]
    with Pool(min(len(jobs), os.cpu_count() or 1), maxtasksperchild=1) as pool:
        res = pool.map(attempt, jobs, chunksize=1)
    recs = []
    for i in range(CANDIDATES):
        four = res[4 * i:4 * i + 4]
        robust = all((r["reward"], r["info"]) == (four[0]["reward"], four[0]["info"]) for r in four)
        print("env", i, "robust" if robust else "rounding-sensitive", [(r["reward"], r["info"][5], r["info"][7]) for r in four])
        if robust and len(recs) < N:
            recs.append(four[0])
    assert len(recs) == N, len(recs)
    arrays = {"qpos0": np.array([r.pop("_q0") for r in recs]), "qvel0": np.array([r.pop("_v0") for r in recs]),
              "qpos1": np.array([r.pop("_q1") for r in recs])}
    json.dump({"about": "8 grasp attempts into the 40-object pile on the fp64 CPU oracle (tests/golden/make_scene_b_golden.py), outcome "
                        "insensitive to a one-ulp perturbation and to the -O3 -ffp-contract=fast build", "table_height": 0.91, "records": recs},
              open(OUT + ".json", "w"), separators=(",", ":"))
    np.savez_compressed(OUT + ".npz", **arrays)
    print([(r["env"], r["reward"], r["substeps"]) for r in recs])


if __name__ == "__main__":
    main()
