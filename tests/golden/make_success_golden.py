#!/usr/bin/env python
"""Generates tests/golden/success_64.npz + success_64.json: 64 single grasp attempts on the fp64 CPU oracle, at least 48 of them
REWARDED (the 256-action replay holds only 9 successful grasps - a thin sample for "success flag bit-exact").

TEST INFRASTRUCTURE.  Environments 100 .. 100+E-1 of the 6-object scene (reset RandomState(20000 + i), settled 1000 ms) run
`GraspEnv.step` (GraspingEnv.py:62-156 restated on the oracle, the same code path as make_replay_golden.py) with actions aimed at
object tops (what a trained agent does).  Every record keeps the oracle's full state before the action, so a replayer can run
each attempt on its own from that state: the first 48 rewarded records and the first 16 executed-but-unrewarded ones are kept.

  python tests/golden/make_success_golden.py
"""
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "success_64")
N_ENVS, N_STEPS, W, H, TABLE_HEIGHT = 48, 16, 200, 200, 0.91
N_REWARDED, N_FAILED = 48, 16


def run_env(i):
    from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob
    from oracle.oracle_py import OracleEnv
    from tests.common import reset_qpos_scene_a

    blob = load_scene_blob("A")
    A, _ = load_scene("A")
    cam = int(np.asarray(A["cam_top_down"]).ravel()[0])
    o = OracleEnv(blob)
    o.reset(reset_qpos_scene_a(A, i))
    o.stay(1000)
    rng = np.random.RandomState(30000 + i)
    rec = []
    for k in range(N_STEPS):
        q, v = o.qpos.copy(), o.qvel.copy()
        o.reset(q, v)
        _, depth = o.render(cam, W, H)
        ys, xs = np.where((depth[60:140, 40:160] < 1.09 - 0.005) & (depth[60:140, 40:160] > 1.09 - 0.06))
        if not len(ys):
            break
        j = rng.randint(len(ys))
        x, y = int(xs[j]) + 40, int(ys[j]) + 60
        a = [y * W + x, int(rng.randint(0, 6))]
        d = float(depth[y][x])
        c = o.pixel_2_world(x, y, d, cam, W, H)
        if c[2] < 0.8 or c[1] > -0.3:
            continue
        reward, info = o.move_and_grasp(c, a[1], TABLE_HEIGHT)
        rec.append({"env": i, "step": k, "action": a, "depth": d, "coords": [float(t) for t in c], "reward": int(reward),
                    "info": [int(t) for t in info], "arm_qpos": [float(t) for t in o.qpos[:8]], "_q0": q, "_v0": v, "_q1": o.qpos.copy()})
    o.close()
    return rec


def main():
    with Pool(min(N_ENVS, os.cpu_count() or 1)) as pool:
        recs = [r for e in pool.map(run_env, range(100, 100 + N_ENVS)) for r in e]
    good = [r for r in recs if r["reward"] == 1][:N_REWARDED]
    bad = [r for r in recs if r["reward"] == 0][:N_FAILED]
    assert len(good) == N_REWARDED and len(bad) == N_FAILED, (len(good), len(bad), len(recs))
    keep = good + bad
    arrays = {"qpos0": np.array([r.pop("_q0") for r in keep]), "qvel0": np.array([r.pop("_v0") for r in keep]),
              "qpos1": np.array([r.pop("_q1") for r in keep])}
    json.dump({"about": "64 single grasp attempts on the fp64 CPU oracle from stored states (tests/golden/make_success_golden.py); "
                        "records 0..47 rewarded, 48..63 executed without reward", "table_height": TABLE_HEIGHT, "records": keep},
              open(OUT + ".json", "w"), separators=(",", ":"))
    np.savez_compressed(OUT + ".npz", **arrays)
    print(f"{len(recs)} attempts run, {sum(r['reward'] for r in recs)} rewarded; kept {len(good)} + {len(bad)}")


if __name__ == "__main__":
    main()
