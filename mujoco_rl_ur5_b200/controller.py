"""MJ_Controller façade: the reference's controller surface (gym_grasper/controller/MujocoController.py:21-829) on top of
the CUDA engine.  It is a single-environment VIEW (env 0 of a BatchedEngine); every method that moves the robot sets up a
movement with the C-ABI and lets `ge_run` execute the PID -> mj_step loop on the GPU.

Return conventions are the reference's: movement methods return "success", "max. steps reached: {n}" or
"No valid joint angles received, could not move EE to position." (MujocoController.py:362,376,463).
Plot / marker / viewer arguments are accepted and ignored (diagnostics, out of scope).
"""
from collections import defaultdict

import numpy as np

from .engine import GROUPS, BatchedEngine
from .model.scene import load_scene, load_scene_blob

try:  # the reference prints coloured status lines; fall back to plain text when termcolor is absent
    from termcolor import colored
except Exception:  # pragma: no cover
    def colored(text, *a, **k):
        return text


class _PIDProxy:
    """Looks like simple_pid.PID for the attributes the reference touches (Kp mutable: GraspingEnv.py:282,347)."""

    def __init__(self, ctrl, index, kp, kd, limit):
        self._c, self._i = ctrl, index
        self._kp, self.Ki, self.Kd = float(kp), 0.0, float(kd)
        self.output_limits = (-float(limit), float(limit))
        self.sample_time = 0.0001

    @property
    def Kp(self):
        return self._kp

    @Kp.setter
    def Kp(self, v):
        self._kp = float(v)
        self._c.engine.set_gain(self._i, float(v))

    @property
    def tunings(self):
        return (self._kp, self.Ki, self.Kd)

    @property
    def setpoint(self):
        return float(self._c.current_target_joint_values[self._i])

    @setpoint.setter
    def setpoint(self, v):
        self._c.current_target_joint_values[self._i] = v


class _CtrlView:
    """`sim.data.ctrl`: the 7 actuator controls of this environment.  Reads come from the device (after a movement: the last PID
    outputs, MujocoController.py:327); item assignment is the open-loop write of actuate_joint_group / toss_it_from_the_ellbow
    (MujocoController.py:256-267, :605-612) and goes straight to the device row."""

    def __init__(self, ctrl):
        self._c = ctrl

    def _read(self):
        return self._c.engine.get_ctrl()[self._c.env_index].cpu().numpy()

    def __len__(self):
        return 7

    def __iter__(self):
        return iter(self._read())

    def __array__(self, dtype=None, copy=None):
        a = self._read()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return self._read()[k]

    def __setitem__(self, k, v):
        eng, e = self._c.engine, self._c.env_index
        full = eng.get_ctrl()
        row = full[e].cpu().numpy()
        row[k] = v
        full[e] = eng.torch.as_tensor(row, dtype=full.dtype, device=full.device)
        mask = np.zeros(eng.n_envs, np.uint8)
        mask[e] = 1
        eng.set_ctrl(full, mask)

    def __repr__(self):
        return repr(self._read())


class _Data:
    def __init__(self, ctrl):
        self._c = ctrl
        self._ctrl = _CtrlView(ctrl)

    @property
    def qpos(self):
        return self._c.engine.get_state()[0][self._c.env_index].cpu().numpy()

    @property
    def qvel(self):
        return self._c.engine.get_state()[1][self._c.env_index].cpu().numpy()

    @property
    def ctrl(self):
        return self._ctrl

    @ctrl.setter
    def ctrl(self, v):
        self._ctrl[:] = v

    @property
    def body_xpos(self):
        return self._c.engine.body_xpos()[self._c.env_index].cpu().numpy()


class _Sim:
    def __init__(self, ctrl):
        self.data = _Data(ctrl)
        self._c = ctrl

    def step(self):
        """mujoco_py MjSim.step(): ONE bare mj_step with sim.data.ctrl as it stands (MujocoController.py:379, :611) - no PID
        evaluation; the closed loop is move_group_to_joint_target."""
        mask = np.zeros(self._c.engine.n_envs, np.uint8)
        mask[self._c.env_index] = 1
        self._c.engine.step_open_loop(1, mask)

    def render(self, width=200, height=200, camera_name="top_down", depth=True):
        rgb, d = self._c.get_image_data(camera=camera_name, width=width, height=height)
        return rgb, d


class _Model:
    """mjModel attributes the reference reads (MujocoController.py:86-134,737-758; Grasping_Agent_multidiscrete.py:130-135)."""

    def __init__(self, arrays, names):
        A = arrays
        self._names = names
        self.nbody, self.njnt, self.ncam = int(A["nbody"][0]), int(A["njnt"][0]), int(A["ncam"][0])
        self.cam_pos0, self.cam_mat0, self.cam_fovy = A["cam_pos0"], A["cam_mat0"], A["cam_fovy"]
        self.jnt_range = A["jnt_range"]
        self.actuator_ctrlrange = A["actuator_ctrlrange"]
        self.actuator_trnid = np.stack([A["actuator_jntid"], -np.ones_like(A["actuator_jntid"])], axis=1)

        class _O:
            pass

        self.opt = _O()
        self.opt.timestep = float(A["opt_timestep"][0])
        self.stat = _O()
        self.stat.extent = float(A["stat_extent"][0])
        self.vis = _O()
        self.vis.map = _O()
        self.vis.map.znear, self.vis.map.zfar = float(A["vis_znear"][0]), float(A["vis_zfar"][0])
        self._jnt_qposadr = A["jnt_qposadr"]
        self._jnt_type = A["jnt_type"]

    def body_name2id(self, n):
        return self._names["body"].index(n)

    def body_id2name(self, i):
        return self._names["body"][i]

    def joint_name2id(self, n):
        return self._names["joint"].index(n)

    def joint_id2name(self, i):
        return self._names["joint"][i]

    def actuator_id2name(self, i):
        return self._names["actuator"][i]

    def camera_name2id(self, n):
        return self._names["camera"].index(n)

    def camera_id2name(self, i):
        return self._names["camera"][i]

    def get_joint_qpos_addr(self, name):
        j = self.joint_name2id(name)
        a = int(self._jnt_qposadr[j])
        n = {0: 7, 1: 4, 2: 1, 3: 1}[int(self._jnt_type[j])]
        return a if n == 1 else (a, a + n)


class MJ_Controller(object):
    def __init__(self, model=None, simulation=None, viewer=None, engine=None, scene="A", env_index=0, device=0):
        if engine is None:
            engine = BatchedEngine(load_scene_blob(scene), 1, device)
        self.engine = engine
        self.env_index = env_index
        self.arrays, self.names = load_scene(scene)
        self.model = model if model is not None else _Model(self.arrays, self.names)
        self.sim = simulation if simulation is not None else _Sim(self)
        self.viewer = viewer
        self.groups = defaultdict(list)
        self.groups["All"] = list(range(7))
        self.groups["Arm"] = list(range(5))
        self.groups["Gripper"] = [6]
        self.create_lists()
        self.actuated_joint_ids = np.array([i[2] for i in self.actuators])
        self.reached_target = False
        self.current_output = np.zeros(7)
        self.image_counter = 0
        self.cam_matrix = None
        self.cam_init = False
        self.last_movement_steps = 0
        self.current_carthesian_target = None

    # ------------------------------------------------------------------ set-up (MujocoController.py:53-77,136-254)
    def create_group(self, group_name, idx_list):
        try:
            assert len(idx_list) <= 7, "Too many joints specified!"
            assert group_name not in self.groups.keys(), "A group with name {} already exists!".format(group_name)
            assert np.max(idx_list) <= 7, "List contains invalid actuator ID (too high)"
            self.groups[group_name] = idx_list
            print("Created new control group '{}'.".format(group_name))
        except Exception as e:
            print(e)
            print("Could not create a new group.")

    def create_lists(self):
        A = self.arrays
        self.controller_list = [_PIDProxy(self, i, A["pid_kp"][i], A["pid_kd"][i], A["pid_lim"][i]) for i in range(7)]
        self.current_target_joint_values = self.engine.get_targets()[self.env_index].cpu().numpy().copy()
        self.actuators = []
        for i in range(7):
            j = int(A["actuator_jntid"][i])
            self.actuators.append([i, self.names["actuator"][i], j, self.names["joint"][j], self.controller_list[i]])

    def _mask(self, group):
        assert group in self.groups.keys(), "No group with name {} exists!".format(group)
        m = 0
        for i in self.groups[group]:
            m |= 1 << i
        return m

    def _push_targets(self):
        t = self.engine.get_targets()
        t[self.env_index] = self.engine.torch.as_tensor(self.current_target_joint_values, dtype=t.dtype, device=t.device)
        self.engine.set_targets(t)

    def _finish(self):
        self.engine.run()
        res, steps, _, _ = self.engine.results()
        r, s = int(res[self.env_index]), int(steps[self.env_index])
        self.last_movement_steps = s
        self.current_target_joint_values = self.engine.get_targets()[self.env_index].cpu().numpy().copy()
        if r == 1:
            return "success"
        if r == 2:
            return "max. steps reached: {}".format(s - 1)
        if r == 3:
            return "No valid joint angles received, could not move EE to position."
        return ""

    # ------------------------------------------------------------------ movements
    def actuate_joint_group(self, group, motor_values):
        """Open-loop write of the group's motor controls (MujocoController.py:256-267); they act in the sim.step() calls that follow
        and are overwritten by the next PID evaluation of move_group_to_joint_target, as in the reference."""
        try:
            assert group in self.groups.keys(), "No group with name {} exists!".format(group)
            assert len(motor_values) == len(self.groups[group]), "Invalid number of actuator values!"
            for i, v in enumerate(self.groups[group]):
                self.sim.data.ctrl[v] = motor_values[i]
        except Exception as e:
            print(e)
            print("Could not actuate requested joint group.")

    def move_group_to_joint_target(self, group="All", target=None, tolerance=0.1, max_steps=10000, plot=False, marker=False,
                                   render=True, quiet=False):
        try:
            mask = self._mask(group)
            if target is not None:
                assert len(target) == len(self.groups[group]), "Mismatching target dimensions for group {}!".format(group)
                for i, v in enumerate(self.groups[group]):
                    self.current_target_joint_values[v] = target[i]
            self._push_targets()
            self.engine.move_group(mask, None, tolerance, max_steps)
            result = self._finish()
            if not quiet:
                if result == "success" and target is not None:
                    print(colored("Joint values for group {} within requested tolerance! ({} steps)".format(group, self.last_movement_steps),
                                  color="green", attrs=["bold"]))
                elif result.startswith("max"):
                    print(colored("Max number of steps reached: {}".format(max_steps), color="red", attrs=["bold"]))
            return result
        except Exception as e:
            print(e)
            print("Could not move to requested joint target.")

    def set_group_joint_target(self, group, target):
        idx = self.groups[group]
        try:
            assert len(target) == len(idx), "Length of the target must match the number of actuated joints in the group."
            self.current_target_joint_values[idx] = target
            self._push_targets()
        except Exception as e:
            print(e)
            print(f"Could not set new group joint target for group {group}")

    def open_gripper(self, half=False, **kwargs):
        kwargs = {k: v for k, v in kwargs.items() if k in ("quiet", "render", "plot", "marker")}
        return self.move_group_to_joint_target(group="Gripper", target=[0.0 if half else 0.4], max_steps=1000, tolerance=0.05, **kwargs)

    def close_gripper(self, **kwargs):
        return self.move_group_to_joint_target(group="Gripper", target=[-0.4], tolerance=0.01, **kwargs)

    def grasp(self, **kwargs):
        result = self.close_gripper(max_steps=300, **kwargs)
        return result != "success"

    def move_ee(self, ee_position, **kwargs):
        joint_angles = self.ik(ee_position)
        if joint_angles is not None:
            result = self.move_group_to_joint_target(group="Arm", target=joint_angles, **kwargs)
        else:
            result = "No valid joint angles received, could not move EE to position."
            self.last_movement_steps = 0
        return result

    def ik(self, ee_position):
        try:
            assert len(ee_position) == 3, "Invalid EE target! Please specify XYZ-coordinates in a list of length 3."
            self.current_carthesian_target = np.array(ee_position, dtype=np.float64).copy()
            xyz = np.tile(self.current_carthesian_target, (self.engine.n_envs, 1))
            q5, ok = self.engine.ik(xyz)
            if bool(ok[self.env_index]):
                return q5[self.env_index].cpu().numpy()
            print("Failed to find IK solution.")
            return None
        except Exception as e:
            print(e)
            print("Could not find an inverse kinematics solution.")

    def stay(self, duration, render=True):
        self._push_targets()
        self.engine.stay(int(duration))
        self.engine.run()

    # ------------------------------------------------------------------ camera (MujocoController.py:708-806)
    def get_image_data(self, show=False, camera="top_down", width=200, height=200):
        cam = self.model.camera_name2id(camera)
        rgb, depth_m = self.engine.render(cam, width, height)
        rgb, depth_m = rgb[self.env_index].cpu().numpy(), depth_m[self.env_index].cpu().numpy()
        # the reference returns the raw GL depth buffer in [0,1]; invert depth_2_meters so that depth_2_meters(get_image_data()) is metric
        ext = self.model.stat.extent
        near, far = self.model.vis.map.znear * ext, self.model.vis.map.zfar * ext
        depth = (1.0 - near / depth_m.astype(np.float64)) / (1.0 - near / far)
        return rgb, depth.astype(np.float32)

    def depth_2_meters(self, depth):
        extend = self.model.stat.extent
        near = self.model.vis.map.znear * extend
        far = self.model.vis.map.zfar * extend
        return near / (1 - depth * (1 - near / far))

    def create_camera_data(self, width, height, camera):
        cam_id = self.model.camera_name2id(camera)
        fovy = self.model.cam_fovy[cam_id]
        f = 0.5 * height / np.tan(fovy * np.pi / 360)
        self.cam_matrix = np.array(((f, 0, width / 2), (0, f, height / 2), (0, 0, 1)))
        self.cam_rot_mat = np.reshape(self.model.cam_mat0[cam_id], (3, 3))
        self.cam_pos = self.model.cam_pos0[cam_id]
        self.cam_init = True

    def world_2_pixel(self, world_coordinate, width=200, height=200, camera="top_down"):
        if not self.cam_init:
            self.create_camera_data(width, height, camera)
        hom_pixel = self.cam_matrix @ self.cam_rot_mat @ (world_coordinate - self.cam_pos)
        pixel = hom_pixel[:2] / hom_pixel[2]
        return np.round(pixel[0]).astype(int), np.round(pixel[1]).astype(int)

    def pixel_2_world(self, pixel_x, pixel_y, depth, width=200, height=200, camera="top_down"):
        if not self.cam_init:
            self.create_camera_data(width, height, camera)
        pixel_coord = np.array([pixel_x, pixel_y, 1]) * (-depth)
        pos_c = np.linalg.inv(self.cam_matrix) @ pixel_coord
        pos_w = np.linalg.inv(self.cam_rot_mat) @ (pos_c + self.cam_pos)
        return pos_w

    # ------------------------------------------------------------------ diagnostics kept as no-ops
    def show_model_info(self):
        print("\nNumber of bodies: {}".format(self.model.nbody))
        for i in range(self.model.nbody):
            print("Body ID: {}, Body Name: {}".format(i, self.model.body_id2name(i)))
        print("\nNumber of joints: {}".format(self.model.njnt))
        for i in range(7):
            print("Actuator ID: {}, Actuator Name: {}, Controlled Joint: {}".format(i, self.actuators[i][1], self.actuators[i][3]))

    def add_marker(self, *a, **k):
        pass

    def fill_plot_list(self, *a, **k):
        pass

    def create_joint_angle_plot(self, *a, **k):
        pass

    def display_current_values(self):
        print("Current joint values:", self.sim.data.qpos[self.actuated_joint_ids])

    @property
    def last_steps(self):
        return self.last_movement_steps
