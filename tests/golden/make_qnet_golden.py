#!/usr/bin/env python
"""Generates tests/golden/qnet_golden.json with the REFERENCE's own Modules.MULTIDISCRETE_RESNET (Modules.py:308-311) run here
on CPU in fp32: parameter checksums after torch.manual_seed(0) default initialisation, and the sigmoid output of one
training-mode batch-1 forward (the way the agent calls it) at 256 fixed positions + the flat arg-max.
`prettytable` (absent here) is stubbed; nothing else is touched.  Run:  python tests/golden/make_qnet_golden.py"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GRASP_REFERENCE_DIR", "/root/reference")
sys.modules["prettytable"] = types.SimpleNamespace(PrettyTable=object)
sys.path.insert(0, REF)
import Modules  # noqa: E402  (the reference file, unmodified)

torch.manual_seed(0)
net = Modules.MULTIDISCRETE_RESNET(number_actions_dim_2=6)
sd = {k: v.clone() for k, v in net.state_dict().items()}  # before the forward (BN running stats change in train mode)
g = torch.Generator().manual_seed(1)
state = torch.rand((1, 4, 200, 200), generator=g)
with torch.no_grad():
    out = net(state)  # train mode (never .eval()): BatchNorm uses the statistics of this one image
flat = out.reshape(-1)
pos = np.random.RandomState(7).choice(flat.numel(), 256, replace=False)
gold = dict(
    keys=list(sd.keys()),
    checksums={k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items() if v.dtype.is_floating_point},
    n_params=int(sum(p.numel() for p in net.parameters())),
    out_shape=list(out.shape), positions=pos.tolist(), values=[float(flat[i]) for i in pos], argmax=int(flat.argmax()), max=float(flat.max()),
    mean=float(flat.mean()),
)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "qnet_golden.json")
json.dump(gold, open(path, "w"))
print("wrote", path, "params", gold["n_params"], "out", gold["out_shape"], "argmax", gold["argmax"], "max %.6f" % gold["max"])
