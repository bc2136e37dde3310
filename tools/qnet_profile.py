#!/usr/bin/env python
"""Q-net forward alone (64 synthetic images, seeded weights) — the command ncu captures for the conv / glue kernels.
usage: [ncu ...] python tools/qnet_profile.py [images] [reps] [chunk]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mujoco_rl_ur5_b200.qnet import QNetForward, make_torch_qnet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 64
torch.manual_seed(0)
qf = QNetForward(make_torch_qnet(6).state_dict(), 0, max_batch=chunk)
g = torch.Generator(device="cuda").manual_seed(1)
obs = {"rgb": torch.randint(0, 256, (n, 200, 200, 3), dtype=torch.uint8, device="cuda", generator=g),
       "depth": 1.0 + 0.1 * torch.rand((n, 200, 200), device="cuda", generator=g)}
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
for r in range(reps):
    a.record()
    act, val = qf.greedy(qf.forward(qf.obs_to_state(obs, 1.1)))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"rep {r}: {ms:.3f} ms, {n / ms * 1e3:.0f} images/s, {n * 41.99424e9 / ms / 1e9:.1f} TFLOP/s (not a bench value when run under ncu)")
