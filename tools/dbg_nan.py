"""debug: which pipeline stage first produces a non-finite value under load (GE_NANCHECK=1), scene B, CTA-per-env build"""
import sys, os
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("GE_NANCHECK", "1")
import torch
from mujoco_rl_ur5_b200.engine import BatchedEngine
from mujoco_rl_ur5_b200.batched_env import scene_b_reset_qpos, HOME
from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
A, _ = load_scene("B")
eng = BatchedEngine(load_scene_blob("B"), N, 0)
eng.set_state(np.stack([scene_b_reset_qpos(A, 20000 + i) for i in range(N)]))
tgt = np.tile(HOME + np.array([0.2, 0.1, -0.1, 0.1, 0.1, 0.3, -0.1]), (N, 1))
for chunk in range(steps // 10):
    eng.move_group("All", tgt, 1e-9, 9)
    eng.run()
    st = eng.status().cpu().numpy()
    sub = eng.results()[3].cpu().numpy()
    bad = np.nonzero(st)[0]
    print(f"after {10*(chunk+1)} sub-steps: {len(bad)} envs flagged; bits histogram", {b: int(((st >> b) & 1).sum()) for b in range(16) if ((st >> b) & 1).any()},
          "first flagged envs", bad[:12].tolist(), "their substeps", sub[bad[:12]].tolist(), flush=True)
    if len(bad) > 20:
        break
q, v = eng.get_state()
print("non-finite qpos envs:", int((~torch.isfinite(q)).any(dim=1).sum()), "qvel:", int((~torch.isfinite(v)).any(dim=1).sum()))
