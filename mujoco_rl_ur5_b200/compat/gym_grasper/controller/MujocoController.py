from mujoco_rl_ur5_b200.controller import MJ_Controller  # noqa: F401
