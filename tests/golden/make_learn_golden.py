#!/usr/bin/env python
"""Generates tests/golden/qnet_learn_golden.json: one `learn()` step of the reference agent (Grasping_Agent_multidiscrete.py:388-446)
run here on CPU in fp32 with the REFERENCE's own Modules.MULTIDISCRETE_RESNET (Modules.py:308-311):
    q_pred = policy_net(state_batch).view(B, -1).gather(1, action_batch); loss = binary_cross_entropy(q_pred, reward) / 1
    loss.backward(); Adam(lr 0.001, weight_decay 0.00002).step()          (:139-156, :426-446; GAMMA = 0 branch)
on a fixed synthetic batch (B = 3 - BatchNorm in training mode normalises over the whole batch here, exactly as in the reference's
learn(), unlike the batch-1 acting forward).  Stored: the loss, per-parameter gradient checksums, and parameter checksums after the step.
It is the parity target of the learner's backward kernels (not built yet) and pins `qnet.make_torch_qnet` as their fp32
comparison network.  `prettytable` (absent here) is stubbed; nothing else is touched.  Run:  python tests/golden/make_learn_golden.py"""
import json
import os
import sys
import types

import torch
import torch.nn.functional as F

REF = os.environ.get("GRASP_REFERENCE_DIR", "/root/reference")
B = 3


def batch():
    g = torch.Generator().manual_seed(5)
    state = torch.rand((B, 4, 200, 200), generator=g)
    action = torch.tensor([[3 * 40000 + 123 * 200 + 77], [0 * 40000 + 20 * 200 + 150], [5 * 40000 + 199 * 200 + 199]])  # flat rot*H*W + y*W + x
    reward = torch.tensor([[1.0], [0.0], [1.0]])
    return state, action, reward


def learn_step(net):
    """the arithmetic of learn() for NUMBER_ACCUMULATIONS_BEFORE_UPDATE = 1, GAMMA = 0"""
    opt = torch.optim.Adam(net.parameters(), lr=0.001, weight_decay=0.00002)
    opt.zero_grad()
    state, action, reward = batch()
    q_pred = net(state).view(B, -1).gather(1, action)
    loss = F.binary_cross_entropy(q_pred, reward) / 1
    loss.backward()
    grads = {k: [float(p.grad.double().sum()), float(p.grad.double().abs().sum())] for k, p in net.named_parameters()}
    opt.step()
    after = {k: [float(p.detach().double().sum()), float(p.detach().double().abs().sum())] for k, p in net.named_parameters()}
    return float(loss.detach()), [float(v) for v in q_pred.detach().reshape(-1)], grads, after


def main():
    sys.modules["prettytable"] = types.SimpleNamespace(PrettyTable=object)
    sys.path.insert(0, REF)
    import Modules  # the reference file, unmodified

    torch.manual_seed(0)
    net = Modules.MULTIDISCRETE_RESNET(number_actions_dim_2=6)
    loss, q_pred, grads, after = learn_step(net)
    gold = dict(batch=B, loss=loss, q_pred=q_pred, grads=grads, params_after_adam_step=after, lr=0.001, weight_decay=0.00002)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "qnet_learn_golden.json")
    json.dump(gold, open(path, "w"))
    print("wrote", path, "loss %.6f" % loss, "q_pred", q_pred)


if __name__ == "__main__":
    main()
