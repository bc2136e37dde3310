"""Scene registry: compiles the reference's MJCF scenes into model blobs and attaches the constants that the
reference keeps in Python (PID gains, IK chain, material colours).

The compiled blobs are committed under mujoco_rl_ur5_b200/assets/ because /root/reference does not exist on
the GPU box; `python -m mujoco_rl_ur5_b200.model.scene` regenerates them from the reference assets.
"""
import os

import numpy as np

from .blob import pack, unpack
from .mjcf import compile_mjcf, model_to_blob_arrays

ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
REFERENCE_DIR = os.environ.get("GRASP_REFERENCE_DIR", "/root/reference")
SCENES = {
    "A": "UR5+gripper/UR5gripper_2_finger.xml",            # 6 objects, condim 4 (BASELINE configs 1-4)
    "B": "UR5+gripper/UR5gripper_2_finger_many_objects.xml",  # 40 free objects, condim 6 (GraspingEnv.py:30 default)
    # BASELINE config 3 "fixed-size objects" = the reference's iteration 1, "objects of equal size" (README.md:20): the 6-object
    # scene with every graspable object turned into the 4 cm cube of box_1 (UR5gripper_2_finger.xml:238)
    "A1": "UR5+gripper/UR5gripper_2_finger.xml",
}
OBJECT_GEOMS_A = ("box_1", "box_2", "box_3", "ball_1", "ball_2", "ball_3")


def _equal_size_objects(root):
    n = 0
    for g in root.iter("geom"):
        if g.attrib.get("name") in OBJECT_GEOMS_A:
            g.attrib["type"], g.attrib["size"] = "box", "0.02 0.02 0.02"
            n += 1
    assert n == len(OBJECT_GEOMS_A), n


TRANSFORMS = {"A1": _equal_size_objects}

# PID gains after p_scale=3, d_scale=0.1 (MujocoController.py:157-235); Ki = 0 everywhere
PID_KP = np.array([7, 10, 5, 7, 5, 5, 2.5]) * 3.0
PID_KD = np.array([1.1, 1.0, 0.5, 0.1, 0.1, 0.1, 0.0]) * 0.1
PID_LIM = np.array([2.0, 2.0, 2.0, 1.0, 1.0, 1.0, 1.0])
# ur5_gripper.urdf chain (lines 61-234): d1 shoulder height, d4 lateral offset 0.13585-0.1197+0.093,
# a1 upper arm, a2 forearm, d5 wrist_2->wrist_3, d6 wrist_3->ee
IK_CHAIN = np.array([0.089159, 0.13585 - 0.1197 + 0.093, 0.425, 0.39225, 0.09465, 0.0823])
IK_LOWER = np.array([-3.14159265, -3.14159265, -3.14159265, -3.14159265, -3.14159265])
IK_UPPER = np.array([3.14159265, -0.9, 3.14159265, 3.14159265, 3.14159265])  # lift upper=-0.9: urdf:94
IK_OFFSET = np.array([0.0, -0.005, 0.16])  # ee_link -> grasp centre (MujocoController.py:493)
# flat material colours standing in for the reference's textures (UR5gripper_2_finger.xml:74-88)
MATERIAL_RGB = {"ur5_mat": (0.45, 0.45, 0.45), "gripper_mat": (0.45, 0.45, 0.45), "bench_mat": (0.62, 0.62, 0.64),
                "tablecube": (0.55, 0.40, 0.26), "geom": (0.8, 0.6, 0.4), "floor_mat": (0.2, 0.3, 0.4)}


def compile_scene(key, reference_dir=None):
    path = os.path.join(reference_dir or REFERENCE_DIR, SCENES[key])
    M = compile_mjcf(path, TRANSFORMS.get(key))
    names = M["_names"]
    M["pid_kp"], M["pid_kd"], M["pid_lim"] = PID_KP, PID_KD, PID_LIM
    M["ik_chain"], M["ik_lower"], M["ik_upper"], M["ik_offset"] = IK_CHAIN, IK_LOWER, IK_UPPER, IK_OFFSET
    M["ik_base_body"] = names["body"].index("base_link")
    M["ee_body"] = names["body"].index("ee_link")
    # base_link is welded to the world: its world position is a model constant (used by the IK, MujocoController.py:488)
    from .rigid import Kinematics
    M["ik_base_pos"] = Kinematics(M).forward(M["qpos0"])["xpos"][M["ik_base_body"]]
    M["cam_top_down"] = names["camera"].index("top_down")
    mats = M.pop("_materials", None)
    rgba = np.array(M["geom_rgba"], dtype=np.float64)
    if mats:
        for g, mat in enumerate(mats):
            if mat in MATERIAL_RGB:
                rgba[g, :3] = MATERIAL_RGB[mat]
    M["geom_rgba"] = rgba
    return M


def blob_path(key):
    return os.path.join(ASSET_DIR, f"scene_{key.lower()}.blob")


def names_path(key):
    return os.path.join(ASSET_DIR, f"scene_{key.lower()}.names.json")


def load_scene_blob(key="A"):
    """bytes of the committed blob for scene `key` (compiled on the fly if missing and the reference exists)."""
    p = blob_path(key)
    if not os.path.exists(p):
        build_assets([key])
    with open(p, "rb") as f:
        return f.read()


def load_scene(key="A"):
    """(dict of arrays, names dict) for scene `key`."""
    import json

    arrays = unpack(load_scene_blob(key))
    with open(names_path(key)) as f:
        names = json.load(f)
    return arrays, names


def build_assets(keys=("A", "B", "A1")):
    import json

    os.makedirs(ASSET_DIR, exist_ok=True)
    for k in keys:
        M = compile_scene(k)
        with open(blob_path(k), "wb") as f:
            f.write(pack(model_to_blob_arrays(M)))
        with open(names_path(k), "w") as f:
            json.dump(M["_names"], f, indent=0)
        print(f"scene {k}: nq={M['nq']} nv={M['nv']} ngeom={M['ngeom']} npair={M['npair']} -> {blob_path(k)}")


if __name__ == "__main__":
    build_assets()
