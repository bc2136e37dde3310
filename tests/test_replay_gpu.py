"""The north star's acceptance replay: "grasp-success flag bit-exact on a fixed 256-action replay, joint angles within 1e-4".

tests/golden/replay_256.json holds 16 environments x 16 consecutive GraspEnv.step calls run on the fp64 CPU oracle
(tests/golden/make_replay_golden.py: the reference's step(), GraspingEnv.py:62-156, restated on the oracle).  Here the
same frozen actions go through the product's public batched API (BatchedGraspEnv.step: host actions -> depth lookup in
the device-rendered observation -> pixel_2_world -> gate -> whole grasp attempt on the device -> re-render) and every
record is compared: the depth read at the action pixel (1e-4 m), the executed / skipped decision (exact), the reward
(bit-exact), the per-phase sub-step counts of the grasp program (exact) and the arm joint angles after the attempt
(1e-4 rad).  Two environments are also replayed live on the oracle so a stale fixture cannot hide a regression.
"""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay_256.json")


def _load():
    return json.load(open(GOLD))


def test_replay_fixture_is_what_the_oracle_produces_today():
    """CPU: re-run the first 3 steps of two environments on the oracle and compare with the committed fixture."""
    import sys

    sys.path.insert(0, os.path.dirname(GOLD))
    import make_replay_golden as mk

    g = _load()
    assert g["n_envs"] * g["n_steps"] == 256 and len(g["envs"]) == 16 and all(len(e) == 16 for e in g["envs"])
    old_steps = mk.N_STEPS
    mk.N_STEPS = 3
    try:
        for i in (0, 11):
            rec = mk.replay_env(i, [s["action"] for s in g["envs"][i][:3]])
            for sn, so in zip(rec, g["envs"][i]):
                assert (sn["executed"], sn["reward"], sn["info"]) == (so["executed"], so["reward"], so["info"])
                assert abs(sn["depth"] - so["depth"]) < 1e-7
                assert np.abs(np.array(sn["arm_qpos"]) - np.array(so["arm_qpos"])).max() < 1e-9
    finally:
        mk.N_STEPS = old_steps


def test_replay_fixture_exercises_every_outcome():
    """the replay is only a parity statement if it contains skipped actions, failed attempts and successful grasps"""
    flat = [s for e in _load()["envs"] for s in e]
    assert sum(not s["executed"] for s in flat) >= 10
    assert sum(s["reward"] for s in flat) >= 5
    assert sum(s["info"][11] for s in flat) >= 20       # fingers closed on something at the table
    assert sum(s["info"][3] == 2 for s in flat) >= 1    # grasp height not reached within 300 sub-steps
    assert sum(s["substeps"] for s in flat) > 400_000


@pytest.mark.gpu
def test_fixed_256_action_replay_matches_oracle():
    from mujoco_rl_ur5_b200.batched_env import BatchedGraspEnv

    g = _load()
    n, T = g["n_envs"], g["n_steps"]
    env = BatchedGraspEnv(n, "A", 0, seed_base=20000, settle_ms=1000)
    env.reset()
    bad = []
    n_reward, n_exec = 0, 0
    worst_q, worst_d = 0.0, 0.0
    for k in range(T):
        actions = np.array([g["envs"][i][k]["action"] for i in range(n)], dtype=np.int32)
        depth_before = env.current_observation["depth"] if env.current_observation is not None else env.get_observation()["depth"]
        depth_before = depth_before.cpu().numpy()
        obs, reward, done, info = env.step(actions)
        executed = info["executed"].cpu().numpy().astype(bool)
        ginfo = env.engine.grasp_info().cpu().numpy()
        qpos = env.engine.get_state()[0].cpu().numpy()
        for i in range(n):
            s = g["envs"][i][k]
            x, y = s["action"][0] % 200, s["action"][0] // 200
            dd = abs(float(depth_before[i, y, x]) - s["depth"])
            worst_d = max(worst_d, dd)
            if dd > 1e-4:
                bad.append(("depth", i, k, float(depth_before[i, y, x]), s["depth"]))
            if bool(executed[i]) != s["executed"]:
                bad.append(("executed", i, k, bool(executed[i]), s["executed"]))
                continue
            if int(reward[i]) != s["reward"]:
                bad.append(("reward", i, k, int(reward[i]), s["reward"], ginfo[i].tolist(), s["info"]))
            if s["executed"]:
                n_exec += 1
                if ginfo[i].tolist() != s["info"]:
                    bad.append(("phase steps", i, k, ginfo[i].tolist(), s["info"]))
            dq = float(np.abs(qpos[i][:8] - np.array(s["arm_qpos"])).max())
            worst_q = max(worst_q, dq)
            if dq > 1e-4:
                bad.append(("arm_qpos", i, k, dq))
            n_reward += int(reward[i])
        assert not done.any()
    env.close()
    report = {"actions": n * T, "executed": n_exec, "successful_grasps": n_reward, "max_abs_arm_angle_diff": worst_q,
              "max_abs_depth_diff_at_action_pixel": worst_d, "mismatches": [list(map(str, b)) for b in bad]}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(GOLD))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        json.dump(report, open(os.path.join(out, "replay_256_report.json"), "w"), indent=1)
    except OSError:
        pass
    print("replay:", {k: v for k, v in report.items() if k != "mismatches"})
    assert not bad, bad[:10]
