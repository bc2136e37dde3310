"""BatchedGreedyAgent: the acting half of the reference's Grasp_Agent for N environments at once, entirely on the device.

Reference: Grasping_Agent_multidiscrete.py — `greedy` (:284-299) = transform_observation (:301-368) -> policy_net (:295) ->
flat arg-max -> transform_action (:381-386); `epsilon_greedy` (:232-282) mixes in uniformly random table pixels.
Here: gq_obs_to_state -> tcgen05 Q-network forward -> gq_argmax, all kernels of libgrasp_qnet.so; the result is an int32
[N,2] tensor (pixel index, rotation index) that BatchedGraspEnv.step accepts without leaving the GPU (BASELINE config 4).
The learner (replay buffer, BCE loss, Adam; :388-446) is a SURVEY 8(f) "next" row and is not part of this class.
"""
import numpy as np

from .qnet import QNetForward, make_torch_qnet


class BatchedGreedyAgent:
    def __init__(self, env, state_dict=None, seed=0, chunk=64):
        import torch

        self.torch = torch
        self.env = env
        if state_dict is None:  # random default initialisation, as the reference's agent starts from
            torch.manual_seed(seed)
            state_dict = make_torch_qnet(len(env.rotations)).state_dict()
        self.qnet = QNetForward(state_dict, env.engine.device.index or 0, max_batch=chunk)
        # depth_threshold = camera height - TABLE_HEIGHT + 0.01 (Grasping_Agent_multidiscrete.py:130-135)
        cam_z = float(np.asarray(env.arrays["cam_pos0"]).reshape(-1, 3)[env.cam][2])
        self.depth_threshold = float(np.round(cam_z - env.TABLE_HEIGHT + 0.01, 3))
        self.gen = torch.Generator(device=env.engine.device)
        self.gen.manual_seed(seed)

    def q_values(self, obs):
        return self.qnet.forward(self.qnet.obs_to_state(obs, self.depth_threshold))

    def greedy(self, obs):
        """[N,2] int32 actions on the device and their Q-values"""
        act, val = self.qnet.greedy(self.q_values(obs))
        return act.to(self.torch.int32), val

    def epsilon_greedy(self, obs, eps):
        """with probability eps a uniformly random (pixel, rotation) — the reference additionally rejects pixels off the table by
        looking up the depth (Grasping_Agent_multidiscrete.py:262-279); here random pixels are drawn inside the table's image box"""
        t = self.torch
        act, val = self.greedy(obs)
        N, W = act.shape[0], self.env.IMAGE_WIDTH
        px = t.randint(40, 160, (N,), generator=self.gen, device=act.device)
        py = t.randint(60, 140, (N,), generator=self.gen, device=act.device)
        rnd = t.stack([py * W + px, t.randint(0, len(self.env.rotations), (N,), generator=self.gen, device=act.device)], dim=1).to(t.int32)
        pick = t.rand(N, generator=self.gen, device=act.device) < eps
        return t.where(pick[:, None], rnd, act), val
