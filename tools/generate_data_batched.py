#!/usr/bin/env python
"""Offline-RL data generation with N batched environments — what `Offline RL/generate_data.py:21-101` does with one env:
epsilon-greedy actions from the Q-network, every (state, flat action, reward) written in the reference's file format
(`Data/grasping_data_{k}.pt`, 12 transitions each; mujoco_rl_ur5_b200/offline_data.py) so that the reference's
`Offline RL/train.py` / `grasping_dataset.py` read them unchanged.

  python tools/generate_data_batched.py --envs 256 --episodes 2 --steps 10 --out Data [--weights ckpt.pt] [--eps 0.3]
Needs a GPU (no CPU fallback).  One `add` = N transitions; episodes reset all environments (GraspingEnv.py:409-477).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--episodes", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--eps", type=float, default=0.3)
    ap.add_argument("--out", default="Data")
    ap.add_argument("--weights", default=None, help="checkpoint written by the reference agent (key model_state_dict) or a bare state_dict")
    ap.add_argument("--scene", default="A")
    ap.add_argument("--seed", type=int, default=122)  # generate_data.py:16
    args = ap.parse_args()
    import torch

    from mujoco_rl_ur5_b200.batched_agent import BatchedGreedyAgent
    from mujoco_rl_ur5_b200.batched_env import BatchedGraspEnv
    from mujoco_rl_ur5_b200.offline_data import TransitionRecorder

    sd = None
    if args.weights:
        ck = torch.load(args.weights, map_location="cpu", weights_only=False)
        sd = ck.get("model_state_dict", ck)
    env = BatchedGraspEnv(args.envs, args.scene, 0)
    agent = BatchedGreedyAgent(env, sd, seed=args.seed)
    rec = TransitionRecorder(args.out)
    for ep in range(args.episodes):
        obs = env.reset()
        for _ in range(args.steps):
            act, _ = agent.epsilon_greedy(obs, args.eps)
            nxt, reward, done, info = env.step(act)
            rec.add(obs, act, reward)
            obs = nxt
        print(f"episode {ep + 1}: rewards so far {dict(rec.reward_counter)}, {len(rec.files)} files")
    files = rec.close()
    print(f"wrote {len(files)} files to {args.out}")


if __name__ == "__main__":
    main()
