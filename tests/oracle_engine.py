"""TEST INFRASTRUCTURE: `OracleEngine` — the BatchedEngine surface (mujoco_rl_ur5_b200/engine.py) implemented on the CPU oracle for
ONE environment, so that the host-side façade (GraspEnv, MJ_Controller) and the reference's unchanged scripts can be exercised
in the `-m "not gpu"` suite.  It is never importable from the product (tests/ only) and is not a fallback: the product's
BatchedEngine still raises without CUDA.  Movements are queued by move_* / stay / grasp and executed by run(), like ge_run.
"""
import numpy as np
import torch

from oracle.oracle_py import OracleEnv


class OracleEngine:
    def __init__(self, blob, n_envs=1, device=0):
        assert n_envs == 1, "the oracle is a single-environment implementation"
        self.torch = torch
        self.o = OracleEnv(blob)
        self.n_envs = 1
        self.device = torch.device("cpu")
        self.nq, self.nv, self.nbody, self.ngeom, self.nu = self.o.nq, self.o.nv, self.o.nbody, self.o.ngeom, 7
        self._pending = None
        self._result, self._steps, self._reward, self._info = 0, 0, 0, [0] * 12
        self._kp0 = self.o.kp.copy()
        self.launched = 0

    def size(self, what):
        return [self.nq, self.nv, self.nbody, self.ngeom, 7, 1, 128, 0, 1, 1 if self.nv > 100 else 0][what]

    def close(self):
        if self.o is not None:
            self.o.close()
            self.o = None

    # state
    def set_state(self, qpos, qvel=None, env_mask=None):
        q = np.asarray(torch.as_tensor(qpos).cpu(), dtype=np.float64).reshape(-1)
        v = None if qvel is None else np.asarray(torch.as_tensor(qvel).cpu(), dtype=np.float64).reshape(-1)
        self.o.reset(q, v)
        self.o.kp[:] = self._kp0
        self._result = self._steps = self._reward = 0

    def get_state(self):
        return torch.tensor(self.o.qpos.copy())[None], torch.tensor(self.o.qvel.copy())[None]

    def body_xpos(self):
        self.o.fk()
        return torch.tensor(self.o.field("xpos").reshape(1, self.nbody, 3))

    def set_gain(self, actuator, value):
        self.o.kp[actuator] = float(np.asarray(value).reshape(-1)[0])

    def set_targets(self, target):
        self.o.target[:] = np.asarray(torch.as_tensor(target).cpu(), dtype=np.float64).reshape(-1)[:7]

    def get_targets(self):
        return torch.tensor(self.o.target.copy())[None]

    def set_ctrl(self, ctrl, env_mask=None):
        self.o.ctrl[:] = np.asarray(torch.as_tensor(ctrl).cpu(), dtype=np.float64).reshape(-1)[:7]

    def get_ctrl(self):
        return torch.tensor(self.o.ctrl.copy())[None]

    def step_open_loop(self, substeps=1, env_mask=None):
        self.o.step(int(substeps))

    # movements
    def move_group(self, group="All", target=None, tolerance=0.1, max_steps=10000, env_mask=None):
        mask = OracleEnv.GROUPS[group] if isinstance(group, str) else int(group)
        tg = None
        if target is not None:
            t = np.asarray(torch.as_tensor(target).cpu(), dtype=np.float64).reshape(-1)
            ids = [i for i in range(7) if mask >> i & 1]
            tg = t[ids] if len(t) == 7 else t
        self._pending = ("group", mask, tg, tolerance, max_steps)

    def move_ee(self, xyz, tolerance=0.05, max_steps=1000, env_mask=None):
        self._pending = ("ee", np.asarray(torch.as_tensor(xyz).cpu(), dtype=np.float64).reshape(3), tolerance, max_steps)

    def stay(self, ms, env_mask=None):
        self._pending = ("stay", int(ms))

    def grasp(self, coords, rot, table_height=0.91, env_mask=None):
        self._pending = ("grasp", np.asarray(torch.as_tensor(coords).cpu(), dtype=np.float64).reshape(3), int(np.asarray(torch.as_tensor(rot).cpu()).reshape(-1)[0]), table_height)

    def run(self, max_substeps=0):
        p, self._pending = self._pending, None
        if p is None:
            return 0
        self.launched += 1
        if p[0] == "group":
            self._result, self._steps = self.o.move_group(p[1], p[2], p[3], p[4])
        elif p[0] == "ee":
            r, self._steps = self.o.move_ee(p[1], p[2], p[3])
            self._result = 3 if r == 0 else r
        elif p[0] == "stay":
            self.o.stay(p[1])
        else:
            self._reward, self._info = self.o.move_and_grasp(p[1], p[2], p[3])
        return 0

    def results(self):
        return (torch.tensor([self._result], dtype=torch.int32), torch.tensor([self._steps], dtype=torch.int32),
                torch.tensor([self._reward], dtype=torch.uint8), torch.tensor([self.o.substeps], dtype=torch.int64))

    def grasp_info(self):
        return torch.tensor([self._info], dtype=torch.int32)

    def status(self):
        return torch.zeros(1, dtype=torch.int32)

    # camera / IK
    def ik(self, xyz):
        q = self.o.ik(np.asarray(torch.as_tensor(xyz).cpu(), dtype=np.float64).reshape(-1)[:3])
        ok = q is not None
        return torch.tensor(q if ok else np.zeros(5))[None], torch.tensor([1 if ok else 0], dtype=torch.uint8)

    def pixel_2_world(self, px, py, depth, cam=1, width=200, height=200):
        f = lambda a: float(np.asarray(torch.as_tensor(a).cpu()).reshape(-1)[0])
        return torch.tensor(self.o.pixel_2_world(f(px), f(py), f(depth), cam, width, height))[None]

    def render(self, cam=1, width=200, height=200, out=None):
        rgb, depth = self.o.render(cam, width, height)
        return torch.tensor(rgb)[None], torch.tensor(depth)[None]

    def counters(self):
        return self.launched, self.launched
