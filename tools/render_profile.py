#!/usr/bin/env python
"""4096-env reset + two renders of scene A - the command ncu captures for k_render.   usage: [ncu ...] python tools/render_profile.py [envs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mujoco_rl_ur5_b200.batched_env import BatchedGraspEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedGraspEnv(n, "A", 0)
env.reset()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
for r in range(3):
    a.record()
    obs = env.get_observation()
    b.record()
    torch.cuda.synchronize()
    print(f"render {r}: {a.elapsed_time(b):.3f} ms for {n} envs (not a bench value when run under ncu)")
