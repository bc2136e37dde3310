"""The reference-facing façade on the GPU: `gym.make("gym_grasper:Grasper-v0")` -> GraspEnv / MJ_Controller, driven like
example_agent.py (reference: example_agent.py:8-26) and checked against the oracle for one attempt."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "mujoco_rl_ur5_b200", "compat")


@pytest.fixture(scope="module")
def env():
    sys.path.insert(0, COMPAT)
    import gym

    e = gym.make("gym_grasper:Grasper-v0", file="/UR5+gripper/UR5gripper_2_finger.xml", show_obs=False, render=False, quiet=True)
    yield e
    e.close()
    sys.path.remove(COMPAT)


def test_gym_surface(env):
    assert list(env.action_space.nvec) == [40000, 6]
    assert env.TABLE_HEIGHT == 0.91 and env.IMAGE_WIDTH == 200
    cam = env.model.camera_name2id("top_down")
    assert np.allclose(env.model.cam_pos0[cam], [0, -0.6, 2.0])
    a = env.action_space.sample()
    assert 0 <= a[0] < 40000 and 0 <= a[1] < 6
    env.print_info()


def test_reset_step_like_example_agent(env, scene_a):
    from oracle.oracle_py import OracleEnv
    from tests.common import object_positions

    blob, A, _ = scene_a
    np.random.seed(20)
    obs = env.reset()
    assert obs["rgb"].shape == (200, 200, 3) and obs["rgb"].dtype == np.uint8 and obs["depth"].shape == (200, 200)
    assert float(obs["depth"].min()) < 1.09  # something (objects / arm) is above the table top, 1.09 m below the camera
    q0 = env.data.qpos.copy()
    v0 = env.data.qvel.copy()
    # aim at the first box through the pixel interface
    p = object_positions(A, q0)[0]
    px, py = env.controller.world_2_pixel(np.array([p[0], p[1], p[2] + 0.02]))
    obs2, reward, done, info = env.step([int(py) * 200 + int(px), 0])
    assert reward in (0, 1) and done is False and obs2["depth"].shape == (200, 200)
    # the same attempt on the oracle, from the same state, with the coordinates the env derived from its depth image
    d = obs["depth"][int(py)][int(px)]
    coords = env.controller.pixel_2_world(int(px), int(py), d)
    o = OracleEnv(blob)
    o.reset(q0, v0)
    r, oinfo = o.move_and_grasp(coords, 0, 0.91)
    assert r == reward and list(env.last_grasp_info) == oinfo
    assert np.abs(env.data.qpos[:8] - o.qpos[:8]).max() < 1e-4
    o.close()


def test_controller_methods(env):
    c = env.controller
    assert c.groups["Arm"] == [0, 1, 2, 3, 4] and c.groups["Gripper"] == [6]
    r = c.move_ee([0.0, -0.6, 1.1], max_steps=1000, tolerance=0.05, quiet=True)
    assert r == "success" and c.last_steps > 1
    r = c.open_gripper(quiet=True)
    assert r == "success" or r.startswith("max")
    r = c.move_ee([5.0, 5.0, 5.0], max_steps=10, quiet=True)
    assert r.startswith("No")
    c.actuators[0][4].Kp = 10.0
    assert c.actuators[0][4].Kp == 10.0
    c.stay(20)
    xyz = c.pixel_2_world(136, 80, 1.11)
    assert np.abs(xyz - [-0.16551974, -0.50804459, 0.88999999]).max() < 1e-6


def test_open_loop_actuation_matches_oracle(env, scene_a):
    """MJ_Controller.actuate_joint_group + sim.step() (MujocoController.py:256-267, :611): open-loop motor values, bare sub-steps"""
    from oracle.oracle_py import OracleEnv

    blob, A, _ = scene_a
    np.random.seed(21)
    env.reset()
    q0, v0 = env.data.qpos.copy(), env.data.qvel.copy()
    env.set_state(q0, v0)
    c = env.controller
    c.actuate_joint_group("Arm", [0.5, -0.25, 0.3, 0.1, -0.1])
    c.actuate_joint_group("Gripper", [0.2])
    assert np.allclose(np.asarray(env.data.ctrl), [0.5, -0.25, 0.3, 0.1, -0.1, 0.0, 0.2])
    for _ in range(25):
        env.sim.step()
    o = OracleEnv(blob)
    o.reset(q0, v0)
    o.ctrl[:] = [0.5, -0.25, 0.3, 0.1, -0.1, 0.0, 0.2]
    o.step(25)
    assert np.abs(env.data.qpos - o.qpos).max() < 1e-9
    assert np.abs(env.data.qvel - o.qvel).max() < 1e-7
    o.close()
    c.actuate_joint_group("Arm", [0.0] * 4)  # wrong length: printed and swallowed like the reference
    c.move_group_to_joint_target(quiet=True, max_steps=50)  # the next closed-loop movement takes the controls back
    env.sim.data.ctrl[:] = 0


def test_default_scene_is_the_reference_default():
    """GraspEnv() with no file = UR5gripper_2_finger_many_objects.xml (GraspingEnv.py:30); GRASP_SCENE=A is an announced override"""
    from mujoco_rl_ur5_b200 import grasp_env

    assert grasp_env._scene_key(None) == ("B", False) and grasp_env._scene_key(grasp_env.REFERENCE_DEFAULT_FILE) == ("B", False)
    assert grasp_env._scene_key("/UR5+gripper/UR5gripper_2_finger.xml") == ("A", False)
    e = grasp_env.GraspEnv(quiet=True, show_obs=False)
    assert e.scene == "B" and e.engine.size(1) == 248
    e.close()


def test_record_grasps_writes_the_side_camera_picture(env, tmp_path):
    """SURVEY 8f.4 on the GPU (GraspingEnv.py:329-335): the phase-by-phase attempt of `record_grasps=True` from a state where the oracle's
    attempt is rewarded (tests/golden/success_64) must hold the object at the drop pose and write the 1000 x 1000 `side` picture."""
    import json

    from PIL import Image

    gold = os.path.join(ROOT, "tests", "golden")
    recs = json.load(open(os.path.join(gold, "success_64.json")))["records"]
    z = np.load(os.path.join(gold, "success_64.npz"))
    made = []
    for i in range(4):  # rewarded records; the recorded variant issues the movements one by one, so its step counts may differ
        env.set_state(z["qpos0"][i], z["qvel0"][i])
        grasped, png = env.move_and_grasp_recorded(recs[i]["coords"], recs[i]["action"][1], directory=str(tmp_path))
        if grasped:
            made.append(png)
    assert len(made) >= 3, made
    im = np.asarray(Image.open(made[0]))
    assert im.shape == (1000, 1000, 3) and im.std() > 5  # a real picture, not a blank frame
