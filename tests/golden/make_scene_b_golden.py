#!/usr/bin/env python
"""Generates tests/golden/scene_b_attempts.{npz,json}: 8 whole grasp attempts into the 40-object pile of the reference's default
scene (UR5gripper_2_finger_many_objects.xml) on the fp64 CPU oracle.

TEST INFRASTRUCTURE.  Env i: reset by the live rule of GraspingEnv.py:418-430 with RandomState(20000 + i), 500 sub-steps for the
pile to form, then `move_and_grasp` (GraspingEnv.py:205-386) at the object nearest to the table centre with rotation i % 6.
The state before the attempt is stored, so that a replayer (tests/test_scale_parity_gpu.py) runs the same attempt from the
same state at any env index of a large batch.

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


def run(i):
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
    o.reset(q, v)
    s0 = o.substeps
    r, info = o.move_and_grasp(coords, i % 6, 0.91)
    rec = {"env": i, "rot": i % 6, "coords": [float(t) for t in coords], "reward": int(r), "info": [int(t) for t in info],
           "substeps": int(o.substeps - s0), "arm_qpos": [float(t) for t in o.qpos[:8]], "_q0": q, "_v0": v, "_q1": o.qpos.copy()}
    o.close()
    return rec


def main():
    with Pool(min(N, os.cpu_count() or 1)) as pool:
        recs = pool.map(run, range(N))
    arrays = {"qpos0": np.array([r.pop("_q0") for r in recs]), "qvel0": np.array([r.pop("_v0") for r in recs]),
              "qpos1": np.array([r.pop("_q1") for r in recs])}
    json.dump({"about": "8 grasp attempts into the 40-object pile on the fp64 CPU oracle (tests/golden/make_scene_b_golden.py)",
               "table_height": 0.91, "records": recs}, open(OUT + ".json", "w"), separators=(",", ":"))
    np.savez_compressed(OUT + ".npz", **arrays)
    print([(r["reward"], r["substeps"]) for r in recs])


if __name__ == "__main__":
    main()
