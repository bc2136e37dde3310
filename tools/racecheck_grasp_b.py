#!/usr/bin/env python
"""Grasp attempts in the 40-object scene for compute-sanitizer (racecheck / memcheck / synccheck): the states and targets of
tests/golden/scene_b_attempts (settled piles), the first `steps` iterations of the attempt (approach, descent into the pile, gripper closing).
usage: compute-sanitizer --tool racecheck python tools/racecheck_grasp_b.py [envs<=8] [steps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_rl_ur5_b200.engine import BatchedEngine
from mujoco_rl_ur5_b200.model.scene import load_scene_blob

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 700
recs = json.load(open(os.path.join(ROOT, "tests", "golden", "scene_b_attempts.json")))["records"]
z = np.load(os.path.join(ROOT, "tests", "golden", "scene_b_attempts.npz"))
eng = BatchedEngine(load_scene_blob("B"), n, 0)
eng.set_state(z["qpos0"][:n], z["qvel0"][:n])
eng.grasp(np.array([r["coords"] for r in recs[:n]]), np.array([r["rot"] for r in recs[:n]], dtype=np.int32), 0.91)
eng.run(steps)
print("status", eng.status().cpu().numpy(), "phase info", eng.grasp_info().cpu().numpy()[:, :6].tolist())
eng.close()
