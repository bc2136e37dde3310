import sys
sys.path.insert(0, ".")
from mujoco_rl_ur5_b200.engine import BatchedEngine
from mujoco_rl_ur5_b200.model.scene import load_scene_blob
e = BatchedEngine(load_scene_blob(sys.argv[1] if len(sys.argv) > 1 else "A"), 8, 0)
print("sizes", [e.L.ge_size(e.h, k) for k in range(10)])
