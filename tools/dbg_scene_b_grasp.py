"""debug helper: one whole grasp attempt on scene B, GPU program vs oracle"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle.oracle_py import OracleEnv
from mujoco_rl_ur5_b200.engine import BatchedEngine
from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob
from tests.common import reset_qpos_scene_b, HOME

A, n = load_scene("B")
blob = load_scene_blob("B")
o = OracleEnv(blob)
o.reset(reset_qpos_scene_b(A, 0))
t = time.time()
o.move_group("All", HOME + np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.001]), 1e-9, 499)
print("oracle settle 500:", round(time.time() - t, 1), "s ncon", o.ncon, flush=True)
q, v = o.qpos.copy(), o.qvel.copy()
# aim at the object closest to the table centre
pos = q[8:288].reshape(40, 7)[:, :3]
on = (pos[:, 2] > 0.85) & (pos[:, 2] < 1.0)
d = np.linalg.norm(pos[:, :2] - np.array([0.0, -0.6]), axis=1) + (~on) * 10
k = int(d.argmin())
coords = np.array([pos[k, 0], pos[k, 1], pos[k, 2] + 0.02])
print("target object", k, coords)
eng = BatchedEngine(blob, 1, 0)
eng.set_state(q[None], v[None])
t = time.time()
eng.grasp(coords[None], np.array([1], dtype=np.int32), 0.91); eng.run()
print("gpu grasp:", round(time.time() - t, 1), "s", flush=True)
_, _, reward, total = eng.results()
info = eng.grasp_info().cpu().numpy()[0]
gq = eng.get_state()[0][0].cpu().numpy()
o.reset(q, v)
t = time.time()
r, oinfo = o.move_and_grasp(coords, 1, 0.91)
print("oracle grasp:", round(time.time() - t, 1), "s")
print("reward gpu/orc", int(reward[0]), r)
print("info gpu", info.tolist()); print("info orc", oinfo)
print("arm dq", np.abs(gq[:8] - o.qpos[:8]).max(), "all dq", np.abs(gq - o.qpos).max(), "status", eng.status().cpu().numpy())
