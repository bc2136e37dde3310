#!/usr/bin/env python
"""One learner update on a fixed synthetic batch (12 transitions, the reference's BATCH_SIZE) - the command ncu captures for the learner kernels.
usage: [ncu ...] python tools/learn_profile.py [batch] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mujoco_rl_ur5_b200.qnet_learn import QNetLearner

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator().manual_seed(7)
st = torch.rand(B, 4, 200, 200, generator=g).cuda()
ac = torch.randint(0, 6 * 200 * 200, (B, 1), generator=g).cuda()
rw = (torch.rand(B, 1, generator=g) < 0.3).float().cuda()
L = QNetLearner(seed=0)
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
for r in range(reps):
    a.record()
    loss = L.learn_step(st, ac, rw)
    b.record()
    torch.cuda.synchronize()
    print(f"rep {r}: {a.elapsed_time(b):.3f} ms per update, loss {loss:.5f} (not a bench value when run under ncu)")
