from mujoco_rl_ur5_b200.grasp_env import GraspEnv  # noqa: F401
