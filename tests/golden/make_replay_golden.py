#!/usr/bin/env python
"""Generates tests/golden/replay_256.json: the fixed 256-action replay of BASELINE.json's north star
("grasp-success bit-parity to the CPU reference on a fixed 256-action replay").

TEST INFRASTRUCTURE.  16 environments of the 6-object scene x 16 consecutive `GraspEnv.step` calls each, run on the
fp64 CPU oracle exactly the way the reference's step() does it (GraspingEnv.py:62-156): depth lookup at the action pixel
in the current observation, pixel_2_world, the "bad depth" gate, move_and_grasp, re-render.  Env i is reset with
RandomState(20000 + i) (scene-A rule) and settled for 1000 ms; its actions come from RandomState(30000 + i):
  40 %  a pixel on top of an object lying on the table (what a trained agent does),
  45 %  a pixel of the table region of the image (the agent's random-action filter, Grasping_Agent_multidiscrete.py:267-279),
  15 %  any pixel of the image (exercises the skip gate: floor, robot base);
rotation uniform over the 6 classes.  The actions are frozen in the fixture, so the GPU replay does not depend on its
own observations for WHAT to do - only for the depth it reads at the given pixel.

Re-synchronised replay.  Contact-rich rigid-body dynamics is chaotic: two fp64 implementations that differ only in
summation order (the GPU reduces across lanes, uses FMA) separate after a few thousand contact sub-steps, so a free-running
16-step comparison measures chaos, not correctness (measured: first differences after 9-14 consecutive attempts).  The
fixture therefore stores the oracle's full state (qpos, qvel) before every action (replay_256_states.npz), the oracle
itself restarts every step from that state (`reset(qpos, qvel)`: zero warm start, PID memory at the current angles), and a
replayer loads the same state before the same action: 256 independent whole-attempt comparisons along one fixed trajectory.

  python tests/golden/make_replay_golden.py            # rewrites replay_256.json
  python tests/golden/make_replay_golden.py --check    # re-runs the oracle and compares with the committed file
"""
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "replay_256.json")
OUT_STATES = os.path.join(HERE, "replay_256_states.npz")
N_ENVS, N_STEPS, W, H, TABLE_HEIGHT = 16, 16, 200, 200, 0.91


def choose_action(rng, depth):
    u = rng.uniform()
    if u < 0.40:
        # object tops: 5 mm .. 60 mm above the table top (camera 1.09 m above it), inside the table region of the image
        ys, xs = np.where((depth[60:140, 40:160] < 1.09 - 0.005) & (depth[60:140, 40:160] > 1.09 - 0.06))
        if len(ys):
            k = rng.randint(len(ys))
            x, y = int(xs[k]) + 40, int(ys[k]) + 60
        else:
            x, y = int(rng.randint(40, 160)), int(rng.randint(60, 140))
    elif u < 0.85:
        x, y = int(rng.randint(40, 160)), int(rng.randint(60, 140))
    else:
        x, y = int(rng.randint(0, W)), int(rng.randint(0, H))
    return [y * W + x, int(rng.randint(0, 6))]


def replay_env(i, actions=None, states=None, perturb_ulp=0):
    """Runs env i on the oracle.  `actions` given: replay them; None: choose them from the oracle's own observations.
    `states` given ((qpos [T,nq], qvel [T,nv])): every step starts from the stored state instead of the oracle's own.
    `perturb_ulp`: move the x coordinate of the first object by one unit in the last place after the settle (rounding-
    sensitivity experiment of tools/free_run_replay.py; 0 for the fixture)."""
    from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob
    from oracle.oracle_py import OracleEnv
    from tests.common import reset_qpos_scene_a

    blob = load_scene_blob("A")
    A, _ = load_scene("A")
    cam = int(np.asarray(A["cam_top_down"]).ravel()[0])
    o = OracleEnv(blob)
    o.reset(reset_qpos_scene_a(A, i))
    o.stay(1000)
    if perturb_ulp:
        o.qpos[8] = np.nextafter(o.qpos[8], np.inf if perturb_ulp > 0 else -np.inf)
    rng = np.random.RandomState(30000 + i)
    rec = []
    for k in range(N_STEPS):
        # synchronisation point: restart from (qpos, qvel) - the oracle's own, or the stored ones
        q, v = (o.qpos.copy(), o.qvel.copy()) if states is None else (states[0][k], states[1][k])
        o.reset(q, v)
        _, depth = o.render(cam, W, H)
        a = choose_action(rng, depth) if actions is None else actions[k]
        x, y = a[0] % W, a[0] // W
        d = float(depth[y][x])
        c = o.pixel_2_world(x, y, d, cam, W, H)
        executed = not (c[2] < 0.8 or c[1] > -0.3)  # GraspingEnv.py:124
        s0 = o.substeps
        reward, info = 0, [0] * 12
        if executed:
            reward, info = o.move_and_grasp(c, a[1], TABLE_HEIGHT)
        rec.append({"action": a, "depth": d, "coords": [float(v) for v in c], "executed": bool(executed), "reward": int(reward),
                    "info": [int(v) for v in info], "substeps": int(o.substeps - s0), "arm_qpos": [float(v) for v in o.qpos[:8]],
                    "_q0": q, "_v0": v, "_q1": o.qpos.copy()})
    o.close()
    return rec


def generate(actions=None, states=None):
    """-> (json-able dict, state arrays {"qpos0" [E,T,nq], "qvel0" [E,T,nv], "qpos1" [E,T,nq] after the step})"""
    with Pool(min(N_ENVS, os.cpu_count() or 1)) as pool:
        envs = pool.starmap(replay_env, [(i, None if actions is None else actions[i],
                                          None if states is None else (states["qpos0"][i], states["qvel0"][i])) for i in range(N_ENVS)])
    arrays = {"qpos0": np.array([[s.pop("_q0") for s in e] for e in envs]), "qvel0": np.array([[s.pop("_v0") for s in e] for e in envs]),
              "qpos1": np.array([[s.pop("_q1") for s in e] for e in envs])}
    return {"about": "fixed 256-action replay on the fp64 CPU oracle (tests/golden/make_replay_golden.py); scene A, env i reset with "
                     "RandomState(20000+i) and settled 1000 ms, 16 consecutive GraspEnv.step calls per env",
            "n_envs": N_ENVS, "n_steps": N_STEPS, "table_height": TABLE_HEIGHT, "envs": envs}, arrays


def compare(old, new, tol_q=1e-9, verbose=True):
    bad = 0
    for i, (eo, en) in enumerate(zip(old["envs"], new["envs"])):
        for k, (so, sn) in enumerate(zip(eo, en)):
            dq = np.abs(np.array(so["arm_qpos"]) - np.array(sn["arm_qpos"])).max()
            if (so["reward"], so["executed"], so["info"]) != (sn["reward"], sn["executed"], sn["info"]) or dq > tol_q:
                bad += 1
                if verbose:
                    print("differs: env", i, "step", k, so["reward"], sn["reward"], so["info"], sn["info"], "arm dq %.2e" % dq)
    return bad


def main():
    if "--check" in sys.argv:
        # teacher-forced from the stored states: what a replayer (the GPU test, or an oracle built with other compiler flags via
        # GRASP_ORACLE_SO) has to reproduce
        old = json.load(open(OUT))
        new, _ = generate([[s["action"] for s in e] for e in old["envs"]], dict(np.load(OUT_STATES)))
        bad = compare(old, new, 1e-4 if os.environ.get("GRASP_ORACLE_SO") else 1e-9)
        print("replay_256.json:", "reproduced" if not bad else f"{bad} of 256 records differ")
        sys.exit(1 if bad else 0)
    g, arrays = generate()
    json.dump(g, open(OUT, "w"), separators=(",", ":"))
    np.savez_compressed(OUT_STATES, **arrays)
    flat = [s for e in g["envs"] for s in e]
    print(f"wrote {OUT}: {len(flat)} actions, {sum(s['executed'] for s in flat)} executed, {sum(s['reward'] for s in flat)} successful grasps, "
          f"{sum(s['substeps'] for s in flat)} sub-steps")


if __name__ == "__main__":
    main()
