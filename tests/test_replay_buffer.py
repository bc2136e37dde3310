"""DeviceReplayBuffer against the reference's `Modules.ReplayBuffer` (Modules.py:28-55): same ring behaviour and the
same sampled transitions from the same `random` stream."""
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

from mujoco_rl_ur5_b200.replay_buffer import DeviceReplayBuffer

REF = "/root/reference"


def _drive(buf_push, buf_sample, buf_len, n_push=57, size=20, B=5):
    """the agent's pattern: one push per step, a sample as soon as 2 B transitions are stored (Grasping_Agent_multidiscrete.py:396-401)"""
    out = []
    for k in range(n_push):
        buf_push(k)
        if buf_len() >= 2 * B:
            out.append(buf_sample(B))
    return out


def _ours(size=20, B=5, n_push=57, rng=None):
    b = DeviceReplayBuffer(size, state_shape=(1, 2, 2), device="cpu", rng=rng)
    return _drive(lambda k: b.push(torch.full((1, 1, 2, 2), float(k)), torch.tensor([[k * 3]]), torch.tensor([[k % 2]])),
                  lambda B_: [int(v) for v in b.sample(B_)[0][:, 0, 0, 0]], lambda: len(b), n_push, size, B)


def test_ring_and_sampling_golden():
    # generated with the reference class (test below) and frozen here so that the check also runs where /root/reference is absent
    got = _ours()
    assert got[0] == [2, 4, 1, 6, 9] and got[-1] == [47, 46, 49, 54, 56]  # = Modules.ReplayBuffer(20, simple=True) driven the same way
    again = _ours()
    assert got == again  # seed 20 at construction -> reproducible
    flat = [v for s in got for v in s]
    assert all(s[-1] == 9 + i for i, s in enumerate(got))             # the newest transition is always the last element
    assert all(len(set(s[:-1])) == 4 for s in got)                     # B - 1 distinct random picks
    assert min(got[-1]) >= 57 - 20                                     # after wrap-around only the last `size` pushes survive
    assert max(flat) == 56


def test_push_batch_equals_consecutive_pushes():
    a = DeviceReplayBuffer(8, state_shape=(1,), device="cpu", rng=random.Random(20))
    b = DeviceReplayBuffer(8, state_shape=(1,), device="cpu", rng=random.Random(20))
    k = 0
    for n in (3, 4, 5, 11, 2):
        s = torch.arange(k, k + n, dtype=torch.float32)[:, None]
        a.push_batch(s, torch.arange(k, k + n), torch.ones(n))
        for i in range(n):
            b.push(s[i], k + i, 1.0)
        k += n
        assert a.position == b.position and len(a) == len(b)
        assert torch.equal(a.states, b.states) and torch.equal(a.actions, b.actions)
        assert a.sample_indices(4) == b.sample_indices(4)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "Modules.py")), reason="the reference checkout is not on this machine")
def test_same_samples_as_the_reference_class():
    import importlib.util

    pytest.importorskip("torchvision")
    saved = sys.modules.get("prettytable")
    sys.modules["prettytable"] = types.SimpleNamespace(PrettyTable=object)
    try:
        spec = importlib.util.spec_from_file_location("ref_modules", os.path.join(REF, "Modules.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("prettytable", None)
        else:
            sys.modules["prettytable"] = saved
    for size, B, n in ((20, 5, 57), (2000, 12, 300), (30, 12, 100)):
        rb = mod.ReplayBuffer(size, simple=True)  # seeds the global `random` with 20
        ref = _drive(lambda k: rb.push(torch.full((1, 1, 2, 2), float(k)), torch.tensor([[k * 3]]), torch.tensor([[k % 2]])),
                     lambda B_: [int(t.state[0, 0, 0, 0]) for t in rb.sample(B_)], lambda: len(rb), n, size, B)
        ours = _ours(size, B, n)  # rng=None: the global `random` module, re-seeded with 20 by the constructor
        assert ours == ref


def test_a_callers_generator_is_not_reseeded():
    """the constructor used to call rng.seed(20) on a generator passed in, and on the global module whatever `seed`"""
    r = random.Random(12345)
    expect = random.Random(12345).random()
    DeviceReplayBuffer(4, state_shape=(1,), device="cpu", rng=r)
    assert r.random() == expect
    random.seed(777)
    expect = random.Random(777).random()
    DeviceReplayBuffer(4, state_shape=(1,), device="cpu", seed=None)
    assert random.random() == expect
