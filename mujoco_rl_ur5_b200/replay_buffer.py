"""Device-resident replay buffer with the reference's sampling semantics (first half of the learner).

Reference: `Modules.ReplayBuffer` (Modules.py:28-55) — a Python list of `simple_Transition(state, action, reward)` used as a
ring (`position` wraps at `size`), `random.seed(20)` at construction, and
    sample(B) = random.sample(memory, B - 1) + [memory[position - 1]]       (uniform without replacement + the newest)
The agent pushes one transition per env step (Grasping_Agent_multidiscrete.py:552) and samples BATCH_SIZE = 12 in `learn()`
(:401) once the buffer holds 2 * BATCH_SIZE.

Here the transitions live in preallocated tensors on one device (states [size,4,H,W], flat actions, rewards), a batched
environment pushes N transitions at once, and `sample` draws the SAME indices the reference would: `random.sample` picks
positions through `randbelow` only, so sampling `range(len)` with the same generator state returns the indices of the
elements the reference's list sampling returns.  By default the generator is Python's global `random` module seeded with
20 exactly like the reference (the unchanged agent draws its epsilon from the same stream); pass `rng=random.Random(20)` for
a private stream (a generator passed in is used as it is, never re-seeded).  Storage and gathers are tensor plumbing (any torch device); there is no kernel here.
"""
import random as _random


class DeviceReplayBuffer:
    def __init__(self, size, state_shape=(4, 200, 200), device="cuda", state_dtype=None, rng=None, seed=20):
        import torch

        self.torch = torch
        self.size = int(size)
        self.device = torch.device(device)
        self.states = torch.zeros((self.size,) + tuple(state_shape), dtype=state_dtype or torch.float32, device=self.device)
        self.actions = torch.zeros(self.size, dtype=torch.int64, device=self.device)
        self.rewards = torch.zeros(self.size, dtype=torch.float32, device=self.device)
        self.position = 0
        self.length = 0
        # Modules.py:33 seeds the GLOBAL stream with 20; a caller-supplied generator keeps its own state, seed=None skips seeding
        self.rng = rng if rng is not None else _random
        if rng is None and seed is not None:
            self.rng.seed(seed)

    def __len__(self):
        return self.length

    def push(self, state, action, reward):
        """one transition: state [4,H,W] or [1,4,H,W] (the agent's tensor), action / reward scalars or 1-element tensors"""
        t = self.torch
        s = state if isinstance(state, t.Tensor) else t.as_tensor(state)
        self.push_batch(s.reshape((1,) + tuple(self.states.shape[1:])), t.as_tensor(action).reshape(1), t.as_tensor(reward).reshape(1))

    def push_batch(self, states, actions, rewards):
        """N transitions in env order = N consecutive `push` calls of the reference (oldest entries are overwritten on wrap-around)"""
        t = self.torch
        n = int(states.shape[0])
        states = states.to(self.device, self.states.dtype)
        actions = t.as_tensor(actions).to(self.device, t.int64).reshape(n)
        rewards = t.as_tensor(rewards).to(self.device, t.float32).reshape(n)
        if n > self.size:  # only the last `size` survive; the write position still advances by n
            skip = n - self.size
            states, actions, rewards = states[skip:], actions[skip:], rewards[skip:]
            self.position = (self.position + skip) % self.size
            n = self.size
        idx = (self.position + t.arange(n, device=self.device)) % self.size
        self.states.index_copy_(0, idx, states)
        self.actions.index_copy_(0, idx, actions)
        self.rewards.index_copy_(0, idx, rewards)
        self.length = min(self.size, self.length + n)
        self.position = (self.position + n) % self.size

    def sample_indices(self, batch_size):
        idx = self.rng.sample(range(self.length), batch_size - 1)
        idx.append((self.position - 1) % self.length)  # memory[position - 1]: the newest (Python's -1 when position == 0)
        return idx

    def sample(self, batch_size):
        """-> (states [B,4,H,W], actions [B,1] int64, rewards [B,1] f32), in the reference's order (random ones, then the newest)"""
        t = self.torch
        idx = t.as_tensor(self.sample_indices(batch_size), dtype=t.int64, device=self.device)
        return self.states.index_select(0, idx), self.actions.index_select(0, idx)[:, None], self.rewards.index_select(0, idx)[:, None]

    def get(self, index):
        return self.states[index], self.actions[index], self.rewards[index]
