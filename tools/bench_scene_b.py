"""scene-B throughput probe: N envs, reset + settle (free fall -> pile), then K sub-steps in the piled state"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from mujoco_rl_ur5_b200.engine import BatchedEngine
from mujoco_rl_ur5_b200.batched_env import scene_b_reset_qpos, HOME
from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
A, _ = load_scene("B")
eng = BatchedEngine(load_scene_blob("B"), N, 0)
q = np.stack([scene_b_reset_qpos(A, 20000 + i) for i in range(N)])
eng.set_state(q)
tgt = np.tile(HOME + np.array([0.2, 0.1, -0.1, 0.1, 0.1, 0.3, -0.1]), (N, 1))
def timed(steps, label):
    eng.move_group("All", tgt, 1e-9, steps - 1)
    torch.cuda.synchronize(); t = time.time()
    eng.run()
    torch.cuda.synchronize(); dt = time.time() - t
    st = eng.status().cpu().numpy()
    print(f"{label}: {N} envs x {steps} sub-steps in {dt:.2f} s = {N*steps/dt:,.0f} sub-steps/s; overflow envs {int((st!=0).sum())}", flush=True)
timed(100, "free fall (0..100)")
timed(200, "impacts (100..300)")
timed(200, "pile (300..500)")
if len(sys.argv) > 3 and sys.argv[3] == "profile":  # under `ncu --profile-from-start off`: only the K piled sub-steps are captured
    torch.cuda.profiler.start()
timed(K, "settled pile")
if len(sys.argv) > 3 and sys.argv[3] == "profile":
    torch.cuda.profiler.stop()
