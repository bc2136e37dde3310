"""TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/grasp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def lib():
    global _LIB
    if _LIB is None:
        # GRASP_ORACLE_SO: another build of the same source (e.g. -ffp-contract=fast) for rounding-sensitivity experiments
        so = os.environ.get("GRASP_ORACLE_SO")
        if not so:
            so = os.path.join(_DIR, "libgrasp_oracle.so")
            if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_DIR, "grasp_oracle.c")):
                build()
        L = C.CDLL(so)
        P, D, I = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_model_load.restype = P
        L.orc_model_load.argtypes = [C.c_char_p, C.c_int64]
        L.orc_model_free.argtypes = [P]
        L.orc_model_size.argtypes = [P, C.c_int]
        L.orc_env_create.restype = P
        L.orc_env_create.argtypes = [P]
        L.orc_env_free.argtypes = [P]
        for f in ("orc_forward", "orc_step", "orc_fk"):
            getattr(L, f).argtypes = [P]
        for f in ("orc_qpos", "orc_qvel", "orc_qacc_ws", "orc_ctrl", "orc_target", "orc_kp", "orc_last_input"):
            getattr(L, f).restype = D
            getattr(L, f).argtypes = [P]
        L.orc_substeps.restype = C.c_long
        L.orc_substeps.argtypes = [P]
        L.orc_set_dt_pid.argtypes = [P, C.c_double]
        L.orc_set_solver.argtypes = [P, C.c_int]
        L.orc_reset.argtypes = [P, D, D]
        L.orc_field.restype = D
        L.orc_field.argtypes = [P, C.c_char_p, I]
        for f in ("orc_ncon", "orc_nefc", "orc_solver_iter"):
            getattr(L, f).argtypes = [P]
        L.orc_contact.argtypes = [P, C.c_int, D]
        L.orc_move_group.argtypes = [P, C.c_int, D, C.c_int, C.c_double, C.c_int, I]
        L.orc_stay.argtypes = [P, C.c_int]
        L.orc_ik.argtypes = [P, D, D]
        L.orc_move_ee.argtypes = [P, D, C.c_double, C.c_int, I]
        L.orc_move_and_grasp.argtypes = [P, D, C.c_int, C.c_double, I]
        L.orc_pixel_2_world.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, D]
        L.orc_depth_2_meters.restype = C.c_double
        L.orc_depth_2_meters.argtypes = [P, C.c_double]
        L.orc_render.argtypes = [P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
        L.orc_newton_stats.argtypes = [C.POINTER(C.c_long), C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def newton_stats(reset=True):
    """process-wide solver statistics of the oracle (profiling aid): dict with solves, Hessian builds, builds whose active set did not
    change since the previous build of the same solve, line-search iterations, constraint rows, solves ending after 1 / 2 / 3+ builds"""
    out = (C.c_long * 8)()
    lib().orc_newton_stats(out, 1 if reset else 0)
    keys = ["solves", "builds", "builds_same_active_set", "linesearch_iterations", "rows", "solves_1_build", "solves_2_builds", "solves_3plus_builds"]
    return dict(zip(keys, [int(v) for v in out]))


class OracleEnv:
    """One fp64 CPU environment."""

    GROUPS = {"All": 0x7F, "Arm": 0x1F, "Gripper": 0x40}

    def __init__(self, blob: bytes):
        L = lib()
        self.L = L
        self.model = L.orc_model_load(blob, len(blob))
        assert self.model, "oracle could not load the model blob"
        self.env = L.orc_env_create(self.model)
        self.nq, self.nv, self.nbody, self.ngeom, self.nM, self.njnt = [L.orc_model_size(self.model, k) for k in range(6)]

    def close(self):
        if self.env:
            self.L.orc_env_free(self.env)
            self.L.orc_model_free(self.model)
            self.env = None

    def _view(self, fn, n):
        return np.ctypeslib.as_array(getattr(self.L, fn)(self.env), shape=(n,))

    @property
    def qpos(self):
        return self._view("orc_qpos", self.nq)

    @property
    def qvel(self):
        return self._view("orc_qvel", self.nv)

    @property
    def qacc_warmstart(self):
        return self._view("orc_qacc_ws", self.nv)

    @property
    def ctrl(self):
        return self._view("orc_ctrl", 7)

    @property
    def target(self):
        return self._view("orc_target", 7)

    @property
    def kp(self):
        return self._view("orc_kp", 7)

    @property
    def last_input(self):
        return self._view("orc_last_input", 7)

    @property
    def substeps(self):
        return int(self.L.orc_substeps(self.env))

    def reset(self, qpos, qvel=None):
        qpos = np.ascontiguousarray(qpos, np.float64)
        qv = None if qvel is None else _dp(np.ascontiguousarray(qvel, np.float64))
        self.L.orc_reset(self.env, _dp(qpos), qv)

    def set_solver(self, name):
        self.L.orc_set_solver(self.env, {"newton": 0, "pgs": 1}[name])

    def forward(self):
        self.L.orc_forward(self.env)

    def step(self, n=1):
        for _ in range(n):
            self.L.orc_step(self.env)

    def fk(self):
        self.L.orc_fk(self.env)

    def field(self, name):
        n = C.c_int(0)
        p = self.L.orc_field(self.env, name.encode(), C.byref(n))
        if not p or n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    @property
    def ncon(self):
        return self.L.orc_ncon(self.env)

    @property
    def nefc(self):
        return self.L.orc_nefc(self.env)

    @property
    def solver_iter(self):
        return self.L.orc_solver_iter(self.env)

    def contacts(self):
        out = np.zeros((self.ncon, 16))
        for i in range(self.ncon):
            self.L.orc_contact(self.env, i, _dp(out[i]))
        return out

    def dense_M(self, dof_parentid, dof_Madr, which="qM"):
        q = self.field(which)
        nv = self.nv
        M = np.zeros((nv, nv))
        for i in range(nv):
            j, k = i, 0
            while j >= 0:
                M[i, j] = M[j, i] = q[dof_Madr[i] + k]
                j = dof_parentid[j]
                k += 1
        return M

    def move_group(self, group="All", target=None, tolerance=0.1, max_steps=10000):
        mask = self.GROUPS[group] if isinstance(group, str) else int(group)
        steps = C.c_int(0)
        if target is None:
            r = self.L.orc_move_group(self.env, mask, None, 0, tolerance, max_steps, C.byref(steps))
        else:
            t = np.ascontiguousarray(target, np.float64)
            r = self.L.orc_move_group(self.env, mask, _dp(t), 1, tolerance, max_steps, C.byref(steps))
        return r, steps.value

    def stay(self, ms):
        self.L.orc_stay(self.env, int(ms))

    def ik(self, xyz):
        q5 = np.zeros(5)
        ok = self.L.orc_ik(self.env, _dp(np.ascontiguousarray(xyz, np.float64)), _dp(q5))
        return (q5 if ok else None)

    def move_ee(self, xyz, tolerance=0.05, max_steps=1000):
        steps = C.c_int(0)
        r = self.L.orc_move_ee(self.env, _dp(np.ascontiguousarray(xyz, np.float64)), tolerance, max_steps, C.byref(steps))
        return r, steps.value

    def move_and_grasp(self, coords, rot, table_height=0.91):
        info = (C.c_int * 12)()
        r = self.L.orc_move_and_grasp(self.env, _dp(np.ascontiguousarray(coords, np.float64)), int(rot), table_height, info)
        return int(r), list(info)

    def pixel_2_world(self, px, py, depth, cam=1, W=200, H=200):
        out = np.zeros(3)
        self.L.orc_pixel_2_world(self.model, cam, W, H, float(px), float(py), float(depth), _dp(out))
        return out

    def depth_2_meters(self, gl_depth):
        return float(self.L.orc_depth_2_meters(self.model, float(gl_depth)))

    def render(self, cam=1, W=200, H=200):
        rgb = np.zeros((H, W, 3), np.uint8)
        depth = np.zeros((H, W), np.float32)
        self.L.orc_render(self.env, cam, W, H, rgb.ctypes.data_as(C.POINTER(C.c_uint8)), depth.ctypes.data_as(C.POINTER(C.c_float)))
        return rgb, depth
