#!/usr/bin/env python
"""Generates tests/golden/camera_math.json by running the REFERENCE's own camera math
(gym_grasper/controller/MujocoController.py: depth_2_meters :729-740, create_camera_data :742-759, world_2_pixel :761-781,
pixel_2_world :783-806) in this container.  The module imports mujoco_py / simple_pid / ikpy / termcolor / matplotlib at
import time; those are stubbed (none of them is touched by the four methods), the methods run unmodified against a fake
`self.model` carrying the camera constants of the scene (top_down: pos (0,-0.6,2.0), identity rotation, fovy 45).

Run here (needs /root/reference):  python tests/golden/make_golden.py
"""
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("GRASP_REFERENCE_DIR", "/root/reference")
for name in ["mujoco_py", "simple_pid", "ikpy", "ikpy.chain", "termcolor", "matplotlib", "matplotlib.pyplot", "pyquaternion"]:
    m = types.ModuleType(name)
    sys.modules[name] = m
sys.modules["simple_pid"].PID = object
sys.modules["termcolor"].colored = lambda s, *a, **k: s
sys.modules["ikpy"].chain = sys.modules["ikpy.chain"]
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["pyquaternion"].Quaternion = object
sys.path.insert(0, REF)
import importlib.util  # noqa: E402

# load the controller module by path: importing the `gym_grasper` package would pull in gym + the env (mujoco_py GL context)
_spec = importlib.util.spec_from_file_location("ref_MujocoController", os.path.join(REF, "gym_grasper/controller/MujocoController.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
MJ_Controller = _mod.MJ_Controller


class _NS:
    pass


def fake_self(extent):
    s = _NS()
    s.model = _NS()
    s.model.stat = _NS()
    s.model.stat.extent = extent
    s.model.vis = _NS()
    s.model.vis.map = _NS()
    s.model.vis.map.znear, s.model.vis.map.zfar = 0.05, 50.0
    s.model.cam_fovy = np.array([45.0, 45.0, 45.0])
    s.model.cam_pos0 = np.array([[2, 2, 2.7], [0, -0.6, 2.0], [0.8, -0.6, 1.0]])
    s.model.cam_mat0 = np.array([np.eye(3).reshape(9)] * 3)
    s.model.camera_name2id = lambda n: {"main1": 0, "top_down": 1, "side": 2}[n]
    s.cam_init = False
    s.cam_matrix = None
    s.create_camera_data = lambda w, h, c: MJ_Controller.create_camera_data(s, w, h, c)
    return s


def main():
    rng = np.random.RandomState(1234)
    s = fake_self(1.756707635790221)
    px = rng.randint(0, 200, 32)
    py = rng.randint(0, 200, 32)
    d = rng.uniform(0.8, 2.0, 32)
    px[0], py[0], d[0] = 136, 80, 1.11  # media/console.png known answer
    p2w = [MJ_Controller.pixel_2_world(s, int(a), int(b), float(c)).tolist() for a, b, c in zip(px, py, d)]
    w2p = [[int(v) for v in MJ_Controller.world_2_pixel(s, np.array(w))] for w in p2w]
    gl = rng.uniform(0.0, 0.9999, 32)
    d2m = MJ_Controller.depth_2_meters(s, gl).tolist()
    out = dict(extent=s.model.stat.extent, px=px.tolist(), py=py.tolist(), depth=d.tolist(), pixel_2_world=p2w, world_2_pixel=w2p,
               gl_depth=gl.tolist(), depth_2_meters=d2m, cam_matrix=s.cam_matrix.tolist())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "camera_math.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, "console.png KAT:", p2w[0])


if __name__ == "__main__":
    main()
