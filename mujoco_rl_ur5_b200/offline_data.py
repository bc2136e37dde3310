"""Offline-RL wire format of the reference, read and written straight from the batched engine.

The reference collects transitions with `Offline RL/generate_data.py:34-101` and stores them with `torch.save` as a dict

    {"states":  [ {"rgb": uint8 ndarray [H,W,3], "depth": float32 ndarray [H,W] (metres)}, ... ],
     "actions": [ int, ... ],     # flat arg-max index over the [6, H, W] Q-map = rot * H*W + y * W + x  (action.item(), :67)
     "rewards": [ int, ... ]}     # 0 / 1                                                                (:76)

in files `Data/grasping_data_{k}.pt` of FILE_SIZE = 12 transitions (:19, :84-85); `unite_data.py:9-30` concatenates them into
one `..._total_of_{n}_transitions.pt`, `extract_positives.py:10-22` filters reward == 1, and `grasping_dataset.py:13-17` reads the
three lists back.  This module produces and consumes exactly those files, so data recorded from N batched environments
trains the reference's `Offline RL/train.py` unchanged, and the reference's recorded data replays into the batched tools.

Host-side only (numpy + torch.save / torch.load): the arrays come from `BatchedGraspEnv` observations (device tensors are
copied to the host once per `add`), nothing here touches the oracle.
"""
import os
import time
from collections import defaultdict

import numpy as np

FILE_SIZE = 12  # generate_data.py:19


def flat_action(env_action, width=200, height=200):
    """(pixel index, rotation index) as `GraspEnv.step` takes it -> the flat Q-map index the agent stores
    (inverse of Grasp_Agent.transform_action, Grasping_Agent_multidiscrete.py:381-386)."""
    a = np.asarray(env_action, dtype=np.int64)
    return a[..., 1] * (width * height) + a[..., 0]


def env_action(flat, width=200, height=200):
    """flat Q-map index -> (pixel index, rotation index): Grasp_Agent.transform_action"""
    f = np.asarray(flat, dtype=np.int64)
    return np.stack([f % (width * height), f // (width * height)], axis=-1)


def _to_numpy(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def pack_transitions(rgb, depth, actions, rewards):
    """arrays with a leading transition axis -> the reference's dict of three Python lists (each state its own pair of arrays,
    plain ints for actions / rewards, as `torch.save(save_dic, ...)` sees them in generate_data.py:84)."""
    rgb, depth = _to_numpy(rgb), _to_numpy(depth)
    actions, rewards = _to_numpy(actions).astype(np.int64).ravel(), _to_numpy(rewards).astype(np.int64).ravel()
    n = len(actions)
    if not (len(rgb) == len(depth) == len(rewards) == n):
        raise ValueError("states, actions and rewards must have the same length")
    if rgb.dtype != np.uint8 or rgb.ndim != 4 or rgb.shape[-1] != 3:
        raise ValueError("rgb must be uint8 [n, H, W, 3]")
    out = defaultdict(list)
    out["states"] = [{"rgb": rgb[i].copy(), "depth": depth[i].astype(np.float32)} for i in range(n)]
    out["actions"] = [int(a) for a in actions]
    out["rewards"] = [int(r) for r in rewards]
    return out


def save_transitions(path, rgb, depth, actions, rewards):
    import torch

    d = pack_transitions(rgb, depth, actions, rewards)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(d, path)
    return len(d["actions"])


def _torch_load(path):
    """torch.load of a (trusted) data file: pickled dicts of numpy arrays need weights_only=False on torch >= 2.6; older torch has
    no such argument"""
    import torch

    try:
        return torch.load(path, weights_only=False)
    except TypeError:  # torch < 1.13
        return torch.load(path)


def load_transitions(path):
    """reads a reference data file (any of generate_data / unite_data / extract_positives outputs) ->
    (rgb u8 [n,H,W,3], depth f32 [n,H,W], actions i64 [n], rewards i64 [n]).  The reference's generator can leave the lists
    of the last file with different lengths (a state is appended before its action exists, generate_data.py:43-47,96-100);
    the common prefix is returned."""
    import torch

    d = _torch_load(path)
    states, actions, rewards = d["states"], d["actions"], d["rewards"]
    n = min(len(states), len(actions), len(rewards))
    if n == 0:
        return np.zeros((0, 0, 0, 3), np.uint8), np.zeros((0, 0, 0), np.float32), np.zeros(0, np.int64), np.zeros(0, np.int64)
    rgb = np.stack([np.asarray(s["rgb"]).astype(np.uint8) for s in states[:n]])
    depth = np.stack([np.asarray(s["depth"], dtype=np.float32) for s in states[:n]])
    return rgb, depth, np.asarray([int(a) for a in actions[:n]], np.int64), np.asarray([int(r) for r in rewards[:n]], np.int64)


def unite(paths, out_path=None):
    """unite_data.py: concatenate data files in the given order; returns the dict (and writes it if `out_path` is given;
    `out_path` may contain `{n}` for the transition count and `{time}` for the reference's %d_%m_%y_%H_%M stamp)."""
    import torch

    final = defaultdict(list)
    for p in paths:
        d = _torch_load(p)
        for k in ("states", "actions", "rewards"):
            final[k] += d[k]
    if out_path:
        torch.save(final, out_path.format(n=len(final["actions"]), time=time.strftime("%d_%m_%y_%H_%M", time.localtime())))
    return final


def extract_positives(data):
    """extract_positives.py:10-22: the transitions with reward == 1"""
    out = defaultdict(list)
    for i, r in enumerate(data["rewards"]):
        if r == 1:
            for k in ("states", "actions", "rewards"):
                out[k].append(data[k][i])
    return out


class TransitionRecorder:
    """Collects (state before the action, flat action, reward) from a batched environment and writes the reference's files.

        rec = TransitionRecorder("Data")
        obs = env.reset()
        for _ in range(steps):
            actions = agent.act(obs)                    # [N, 2] (pixel, rot) as BatchedGraspEnv.step takes them
            nxt, reward, done, info = env.step(actions)
            rec.add(obs, actions, reward)               # N transitions
            obs = nxt
        rec.close()                                     # flushes the last (shorter) file

    Files are `grasping_data_{k}.pt`, k = 1, 2, ... with `file_size` transitions each (generate_data.py:83), in env-major order
    within one `add` (env 0 first)."""

    def __init__(self, directory, file_size=FILE_SIZE, width=200, height=200, first_index=1):
        self.dir, self.file_size, self.W, self.H = directory, int(file_size), width, height
        self.k = first_index
        self.rgb, self.depth, self.act, self.rew = [], [], [], []
        self.n_pending = 0
        self.files = []
        self.reward_counter = defaultdict(int)  # generate_data.py:38,77

    def add(self, obs, env_actions, rewards):
        rgb, depth = _to_numpy(obs["rgb"]), _to_numpy(obs["depth"])
        if rgb.ndim == 3:
            rgb, depth = rgb[None], depth[None]
        a = flat_action(np.asarray(_to_numpy(env_actions)).reshape(-1, 2), self.W, self.H)
        r = _to_numpy(rewards).astype(np.int64).ravel()
        if not (len(rgb) == len(depth) == len(a) == len(r)):
            raise ValueError("observation batch, actions and rewards disagree in length")
        self.rgb.append(rgb.astype(np.uint8, copy=True)); self.depth.append(depth.astype(np.float32, copy=True))
        self.act.append(a); self.rew.append(r)
        for v in r:
            self.reward_counter[str(int(v))] += 1
        self.n_pending += len(a)
        self._flush(final=False)

    def _flush(self, final):
        if self.n_pending == 0 or (self.n_pending < self.file_size and not final):
            return
        rgb, depth = np.concatenate(self.rgb), np.concatenate(self.depth)
        act, rew = np.concatenate(self.act), np.concatenate(self.rew)
        i = 0
        while self.n_pending - i >= self.file_size or (final and i < self.n_pending):
            j = min(i + self.file_size, self.n_pending)
            path = os.path.join(self.dir, f"grasping_data_{self.k}.pt")
            save_transitions(path, rgb[i:j], depth[i:j], act[i:j], rew[i:j])
            self.files.append(path)
            self.k += 1
            i = j
        self.rgb, self.depth, self.act, self.rew = ([rgb[i:]], [depth[i:]], [act[i:]], [rew[i:]]) if i < self.n_pending else ([], [], [], [])
        self.n_pending -= i

    def close(self):
        self._flush(final=True)
        return self.files
