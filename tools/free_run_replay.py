#!/usr/bin/env python
"""FREE-RUNNING version of the fixed 256-action replay (tests/golden/replay_256.json): 16 environments x 16 consecutive
GraspEnv.step calls with the frozen actions of the fixture, every environment continuing from ITS OWN state (the test
tests/test_replay_gpu.py::test_fixed_256_action_replay_matches_oracle restarts every step from the oracle's state instead).

Contact-rich rigid-body dynamics amplifies rounding differences, so two correct fp64 implementations that differ in summation
order / FMA contraction / factorisation algorithm agree for a number of attempts and then separate.  This tool measures that
number for several "replayers" against the committed oracle trajectory, so that a modelling bug (which would show in the
re-synchronised replay as well) can be told from rounding:

  --backend gpu                       the CUDA engine (GE_LIB selects another build, e.g. the -fmad=false one)
  --backend oracle                    the CPU oracle again (GRASP_ORACLE_SO selects another build, e.g. -O3 -ffp-contract=fast)
  --backend oracle --perturb-ulp 1    the same oracle build, one object coordinate moved by ONE unit in the last place at the start

Each step the replayer re-initialises its own state the way the fixture's generator did (`reset(qpos, qvel)`: zero warm start, PID
memory at the current angles), so the only difference to the fixture is where the state comes from.  Output: JSON report with
records that agree (executed flag, reward, 12 phase counters), rewards that agree, the first differing step per env, and the
state difference to the oracle's trajectory before every step.

  python tools/free_run_replay.py --backend gpu --out gpurun_out/free_run_gpu.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)


def run_gpu(g, st):
    from mujoco_rl_ur5_b200.batched_env import BatchedGraspEnv

    n, T = g["n_envs"], g["n_steps"]
    env = BatchedGraspEnv(n, "A", 0)
    env.engine.set_state(st["qpos0"][:, 0], st["qvel0"][:, 0])
    out = [[None] * T for _ in range(n)]
    for k in range(T):
        q, v = env.engine.get_state()
        qh = q.cpu().numpy().copy()
        env.engine.set_state(q.clone(), v.clone())  # the generator's reset(qpos, qvel) at every step, from the replayer's OWN state
        env.current_observation = None
        actions = np.array([g["envs"][i][k]["action"] for i in range(n)], dtype=np.int32)
        _, reward, _, info = env.step(actions)
        executed = info["executed"].cpu().numpy().astype(bool)
        ginfo = env.engine.grasp_info().cpu().numpy()
        for i in range(n):
            out[i][k] = {"executed": bool(executed[i]), "reward": int(reward[i]), "info": ginfo[i].tolist() if executed[i] else [0] * 12, "q0": qh[i]}
    env.close()
    return out


def _oracle_env(args):
    import make_replay_golden as mk

    i, actions, ulp = args
    rec = mk.replay_env(i, actions, None, perturb_ulp=ulp)
    return [{"executed": r["executed"], "reward": r["reward"], "info": r["info"], "q0": r["_q0"]} for r in rec]


def run_oracle(g, ulp):
    from multiprocessing import Pool

    n = g["n_envs"]
    with Pool(min(n, os.cpu_count() or 1)) as pool:
        return pool.map(_oracle_env, [(i, [s["action"] for s in g["envs"][i]], ulp) for i in range(n)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gpu", choices=["gpu", "oracle"])
    ap.add_argument("--perturb-ulp", type=int, default=0)
    ap.add_argument("--label", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    g = json.load(open(os.path.join(GOLD, "replay_256.json")))
    st = np.load(os.path.join(GOLD, "replay_256_states.npz"))
    n, T = g["n_envs"], g["n_steps"]
    res = run_gpu(g, st) if a.backend == "gpu" else run_oracle(g, a.perturb_ulp)
    agree = reward_agree = 0
    first_diff, growth = [], np.zeros((n, T))
    for i in range(n):
        fd = None
        for k in range(T):
            s, r = g["envs"][i][k], res[i][k]
            same = (r["executed"], r["reward"]) == (s["executed"], s["reward"]) and (not s["executed"] or r["info"] == s["info"])
            agree += same
            reward_agree += r["reward"] == s["reward"] and r["executed"] == s["executed"]
            if not same and fd is None:
                fd = k
            growth[i, k] = float(np.abs(np.asarray(r["q0"]) - st["qpos0"][i, k]).max())
        first_diff.append(fd)
    label = a.label or (a.backend + (f"+{a.perturb_ulp}ulp" if a.perturb_ulp else "") +
                        (" GE_LIB=" + os.path.basename(os.environ["GE_LIB"]) if a.backend == "gpu" and os.environ.get("GE_LIB") else "") +
                        (" GRASP_ORACLE_SO=" + os.path.basename(os.environ["GRASP_ORACLE_SO"]) if a.backend == "oracle" and os.environ.get("GRASP_ORACLE_SO") else ""))
    report = {"replayer": label, "records": n * T, "records_identical": int(agree), "rewards_identical": int(reward_agree),
              "envs_identical_to_the_end": int(sum(f is None for f in first_diff)), "first_differing_step_per_env": first_diff,
              "median_state_diff_before_step": [float(np.median(growth[:, k])) for k in range(T)],
              "max_state_diff_before_step": [float(growth[:, k].max()) for k in range(T)]}
    print(json.dumps(report))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
