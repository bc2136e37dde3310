"""GraspEnv façade: the reference's gym environment (gym_grasper/envs/GraspingEnv.py:25-489) backed by the CUDA engine.

Same constructor keywords, `reset()`, `step(action, record_grasps=False, markers=False, action_info="no info")`, `close()`,
`print_info()`, `action_space` (MultiDiscrete([W*H, 6])), `model`, `TABLE_HEIGHT`, `IMAGE_WIDTH/HEIGHT`,
`current_observation`, `controller` — so that `example_agent.py` and `Grasping_Agent_multidiscrete.py` run unchanged with
`mujoco_rl_ur5_b200/compat` on PYTHONPATH.  One GraspEnv = one environment of a BatchedEngine (n_envs = 1); use
`BatchedGraspEnv` for throughput.

Deviations from the reference, all documented in DESIGN.md: deterministic `stay` / PID period, analytic IK
(A.4), no viewer / cv2 windows.  The default scene is the reference's (`file="/UR5+gripper/UR5gripper_2_finger_many_objects.xml"`,
GraspingEnv.py:30); `file="/UR5+gripper/UR5gripper_2_finger.xml"` (or GRASP_SCENE=A in the environment, announced on stdout)
selects the 6-object scene.
"""
import copy
import math
import os
from collections import defaultdict

import numpy as np

from .batched_env import HOME, ROTATIONS, scene_b_reset_qpos
from .controller import MJ_Controller, _Model, colored
from .engine import BatchedEngine
from .model.scene import load_scene, load_scene_blob

try:
    from gym import spaces, utils
except Exception:  # pragma: no cover - the shim lives in mujoco_rl_ur5_b200/compat
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat"))
    from gym import spaces, utils


REFERENCE_DEFAULT_FILE = "/UR5+gripper/UR5gripper_2_finger_many_objects.xml"  # GraspingEnv.py:30


def _scene_key(file):
    """(scene key, came from the GRASP_SCENE override?)  The MJCFs are compiled into assets/scene_{a,b}.blob; `file` selects one by
    name exactly as the reference's argument selects the XML."""
    if file is None or str(file) == REFERENCE_DEFAULT_FILE:
        ov = os.environ.get("GRASP_SCENE")
        if ov in ("A", "B"):
            return ov, True
        return "B", False
    name = os.path.basename(str(file))
    return ("B" if "many_objects" in name else "A"), False


class GraspEnv(utils.EzPickle):
    metadata = {"render.modes": ["rgb_array"], "video.frames_per_second": 500}

    def __init__(self, file=REFERENCE_DEFAULT_FILE, image_width=200, image_height=200, show_obs=True, demo=False, render=False, device=0,
                 quiet=False, strict_reference_quirks=True):
        self.initialized = False
        self.IMAGE_WIDTH = image_width
        self.IMAGE_HEIGHT = image_height
        self.rotations = dict(ROTATIONS)
        self.action_space_type = "multidiscrete"
        self.step_called = 0
        utils.EzPickle.__init__(self)
        self.scene, overridden = _scene_key(file)
        if overridden:
            print("GraspEnv: GRASP_SCENE={} overrides the reference's default scene ({})".format(self.scene, REFERENCE_DEFAULT_FILE))
        # reference quirks kept by default (Q5: reset() leaves current_observation stale)
        self.strict_reference_quirks = bool(strict_reference_quirks)
        self.arrays, self.names = load_scene(self.scene)
        self.engine = BatchedEngine(load_scene_blob(self.scene), 1, device)
        self.model = _Model(self.arrays, self.names)
        self.frame_skip = 1
        self.dt = self.model.opt.timestep * self.frame_skip
        self._set_action_space()
        self.observation_space = spaces.Dict({"rgb": spaces.Box(0, 255, (image_height, image_width, 3), np.uint8),
                                              "depth": spaces.Box(0, np.inf, (image_height, image_width), np.float32)})
        self.controller = MJ_Controller(self.model, None, None, engine=self.engine, scene=self.scene)
        self.sim = self.controller.sim
        self.data = self.sim.data
        self.viewer = None
        # MujocoEnv.__init__ steps once before the controller exists (GraspingEnv.py:80-85): zero images
        self.current_observation = defaultdict()
        self.current_observation["rgb"] = np.zeros((self.IMAGE_WIDTH, self.IMAGE_HEIGHT, 3))
        self.current_observation["depth"] = np.zeros((self.IMAGE_WIDTH, self.IMAGE_HEIGHT))
        self.step_called = 1
        self.initialized = True
        self.grasp_counter = 0
        self.show_observations = show_obs
        self.demo_mode = demo
        self.TABLE_HEIGHT = 0.91
        self.render = render  # NB: like the reference, this shadows the gym render() method with a bool (GraspingEnv.py:57)
        self.quiet = quiet
        self._episode = 0
        self.seed_base = 20000

    def __repr__(self):
        return f"GraspEnv(obs height={self.IMAGE_HEIGHT}, obs_width={self.IMAGE_WIDTH}, AS={self.action_space_type})"

    def _print(self, *a, **k):
        if not self.quiet:
            print(*a, **k)

    def _set_action_space(self):
        self.action_space = spaces.MultiDiscrete([self.IMAGE_HEIGHT * self.IMAGE_WIDTH, len(self.rotations)])
        return self.action_space

    # ------------------------------------------------------------------ step (GraspingEnv.py:62-156)
    def step(self, action, record_grasps=False, markers=False, action_info="no info"):
        done = False
        info = {}
        if self.step_called == 1:
            self.current_observation = self.get_observation(show=False)
        x = int(action[0]) % self.IMAGE_WIDTH
        y = int(action[0]) // self.IMAGE_WIDTH
        rotation = int(action[1])
        depth = self.current_observation["depth"][y][x]
        coordinates = self.controller.pixel_2_world(pixel_x=x, pixel_y=y, depth=depth, height=self.IMAGE_HEIGHT, width=self.IMAGE_WIDTH)
        self._print(colored("Action ({}): Pixel X: {}, Pixel Y: {}, Rotation: {} ({} deg)".format(action_info, x, y, rotation, self.rotations[rotation]),
                            color="blue", attrs=["bold"]))
        self._print(colored("Transformed into world coordinates: {}".format(coordinates[:2]), color="blue", attrs=["bold"]))
        if coordinates[2] < 0.8 or coordinates[1] > -0.3:
            self._print(colored("Skipping execution due to bad depth value!", color="red", attrs=["bold"]))
            reward = 0
        else:
            grasped_something = self.move_and_grasp(coordinates, rotation, render=self.render, record_grasps=record_grasps, markers=markers)
            reward = 1 if grasped_something else 0
            self._print(colored("Reward received during step: {}".format(reward), color="yellow", attrs=["bold"]))
        self.current_observation = self.get_observation(show=self.show_observations)
        self.step_called += 1
        return self.current_observation, reward, done, info

    def rotate_wrist_3_joint_to_value(self, degrees):
        self.controller.current_target_joint_values[5] = math.radians(degrees)
        return self.controller.move_group_to_joint_target(tolerance=0.05, max_steps=500, render=self.render, quiet=True)

    def transform_height(self, height_action, depth_height):
        return np.round(self.TABLE_HEIGHT + height_action * (0.1) / self.action_space.nvec[1], decimals=3)

    def move_and_grasp_recorded(self, coordinates, rotation, directory="."):
        """`record_grasps=True` (GraspingEnv.py:329-335): the reference photographs a successful grasp with the `side` camera at
        1000 x 1000 while the object still hangs in the closed gripper, i.e. between the final finger check and the re-opening, and
        writes `Grasp_{n}.png`.  The one-kernel grasp program cannot stop there, so this variant issues the same movements one by one
        through the controller (each still a device loop) and takes the picture at that point.  Returns (grasped, png path or None)."""
        c = self.controller
        xyz = np.asarray(coordinates, dtype=np.float64)
        above, centre, drop = [xyz[0], xyz[1], 1.1], [0.0, -0.6, 1.1], [0.6, 0.0, 1.15]
        kw = dict(quiet=True, render=False)
        res = c.move_ee(above, max_steps=1000, tolerance=0.05, **kw)
        if res.startswith("No"):
            res = c.move_ee(centre, max_steps=1000, tolerance=0.05, **kw)
        holding = False
        if not res.startswith("max"):  # stuck on the way: no attempt (GraspingEnv.py:240-246)
            self.rotate_wrist_3_joint_to_value(self.rotations[rotation])
            c.open_gripper(half=True, **kw)
            low = [xyz[0], xyz[1], max(self.TABLE_HEIGHT, xyz[2] - 0.01)]
            if not c.move_ee(low, max_steps=300, tolerance=0.01, **kw).startswith("max"):
                c.stay(100)
                holding = c.grasp(**kw)
        c.actuators[0][4].Kp = 10.0
        c.move_ee(centre, max_steps=1000, tolerance=0.05, **kw)
        c.move_ee(drop, max_steps=1200, tolerance=0.01, **kw)
        check = c.close_gripper(max_steps=1000, **kw) if holding else "Skipped"
        grasped = bool(holding and check.startswith("max"))
        png = None
        if grasped:
            rgb, _ = c.get_image_data(width=1000, height=1000, camera="side")
            self.grasp_counter += 1
            png = os.path.join(directory, "Grasp_{}.png".format(self.grasp_counter))
            from PIL import Image  # (the reference: cv.imwrite of the colour-swapped array = the same RGB picture)

            Image.fromarray(np.ascontiguousarray(rgb)).save(png)
        c.open_gripper(**kw)
        if grasped:
            c.stay(200)
        self.rotate_wrist_3_joint_to_value(0)
        c.actuators[0][4].Kp = 20.0
        self._print(colored("Successful grasp!", color="green", attrs=["bold"]) if grasped else colored("Did not grasp anything.", color="red", attrs=["bold"]))
        return grasped, png

    def move_and_grasp(self, coordinates, rotation, render=False, record_grasps=False, markers=False, plot=False):
        """Whole 11-phase attempt as ONE device program (ge_grasp), then the reference's result summary."""
        if record_grasps:
            return self.move_and_grasp_recorded(coordinates, rotation)[0]
        eng = self.engine
        self.controller._push_targets()
        eng.grasp(np.asarray(coordinates, dtype=np.float64).reshape(1, 3), np.array([rotation], dtype=np.int32), self.TABLE_HEIGHT)
        eng.run()
        _, _, reward, _ = eng.results()
        inf = eng.grasp_info()[0].cpu().numpy()
        self.controller.current_target_joint_values = eng.get_targets()[0].cpu().numpy().copy()
        self.controller.actuators[0][4]._kp = 20.0
        code = {0: "No valid joint angles received, could not move EE to position.", 1: "success", 2: "max. steps reached"}
        r1 = int(inf[0]) % 10
        self._print("Results: ")
        self._print("Move to pre grasp position: ".ljust(40, " "), ("Center" if inf[0] >= 10 else "Above target") if r1 == 1 else "Failed", ",", inf[1], "steps")
        self._print("Rotate gripper: ".ljust(40, " "), "success" if r1 != 2 else "Skipped", ",", inf[2], "steps")
        self._print("Move to grasping position:".ljust(40, " "), code.get(int(inf[3]), "Skipped") if r1 != 2 else "Skipped", ",", inf[4], "steps")
        self._print("Grasped anything?: ".ljust(40, " "), bool(inf[11]))
        self._print("Move to center: ".ljust(40, " "), inf[6], "steps")
        self._print("Move to drop position: ".ljust(40, " "), inf[7], "steps")
        self._print("Final finger check: ".ljust(40, " "), "Object in the gripper" if int(reward[0]) else "Nothing in the gripper")
        self._print("Open gripper: ".ljust(40, " "), inf[9], "steps")
        self.last_grasp_info = inf
        if int(reward[0]):
            self._print(colored("Successful grasp!", color="green", attrs=["bold"]))
            return True
        self._print(colored("Did not grasp anything.", color="red", attrs=["bold"]))
        return False

    # ------------------------------------------------------------------ observation / reset (GraspingEnv.py:390-477)
    def get_observation(self, show=True):
        rgb, depth = self.controller.get_image_data(width=self.IMAGE_WIDTH, height=self.IMAGE_HEIGHT, show=False)
        depth = self.controller.depth_2_meters(depth)
        observation = defaultdict()
        observation["rgb"] = rgb
        observation["depth"] = depth
        return observation

    def reset(self):
        return self.reset_model()

    def reset_model(self, show_obs=True):
        """Object randomisation drawn from numpy's GLOBAL RNG like the reference (Q8): the live 40-object rule of
        GraspingEnv.py:418-430 for the many-objects scene, the IT4 rule (:435-463) for the 6-object scene."""
        if self.scene == "B":
            q = scene_b_reset_qpos(self.arrays)
            return self._finish_reset(q)
        q = np.array(self.arrays["qpos0"], dtype=np.float64).copy()
        q[:7] = HOME
        q[7] = HOME[6]
        nobj = (len(q) - 8) // 7
        # draw order of the reference loop: for j in [rot, x, y, z]: boxes then balls (no draws for rot / z)
        xs = np.array([np.random.uniform(low=-0.25, high=0.25) for _ in range(nobj)])
        ys = np.array([np.random.uniform(low=-0.17, high=0.17) for _ in range(nobj)])
        for i in range(nobj):
            a = 8 + 7 * i
            q[a], q[a + 1], q[a + 2] = xs[i], ys[i], 0.0
            q[a + 3:a + 7] = [1, 0, 0, 0]
        return self._finish_reset(q)

    def _finish_reset(self, q):
        self.set_state(q, np.zeros(int(self.arrays["nv"][0])))
        self.controller.set_group_joint_target(group="All", target=q[self.controller.actuated_joint_ids])
        self.controller.stay(1000, render=self.render)
        if self.demo_mode:
            self.controller.stay(5000, render=self.render)
        obs = self.get_observation(show=self.show_observations)
        if not self.strict_reference_quirks:
            # the reference does NOT do this (quirk Q5, GraspingEnv.py:87-88,477): the first step of every episode but the first reads
            # its depth from the previous episode's last image
            self.current_observation = obs
        return obs

    def set_state(self, qpos, qvel):
        self.engine.set_state(np.asarray(qpos, dtype=np.float64).reshape(1, -1), np.asarray(qvel, dtype=np.float64).reshape(1, -1))
        self.controller.current_target_joint_values = self.engine.get_targets()[0].cpu().numpy().copy()
        for i in range(7):
            self.controller.controller_list[i]._kp = float(self.arrays["pid_kp"][i])

    def close(self):
        self.engine.close()

    def print_info(self):
        print("Model timestep:", self.model.opt.timestep)
        print("Set number of frames skipped: ", self.frame_skip)
        print("dt = timestep * frame_skip: ", self.dt)
        print("Frames per second = 1/dt: ", self.metadata["video.frames_per_second"])
        print("Actionspace: ", self.action_space)
        print("Observation space:", self.observation_space)
