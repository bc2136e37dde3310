"""debug helper: per-contact differences GPU vs oracle for scene B states"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle.oracle_py import OracleEnv
from mujoco_rl_ur5_b200.engine import BatchedEngine
from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob
from tests.common import reset_qpos_scene_b

A, n = load_scene("B")
blob = load_scene_blob("B")
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 0
TGT = np.array([0, -1.57, 1.57, -1.57, -1.57, 0.0, 0.3]) + np.array([0.2, 0.1, -0.1, 0.1, 0.1, 0.3, -0.1])
eng = BatchedEngine(blob, 2, 0)
np.set_printoptions(precision=6, suppress=True, linewidth=220)
for env in range(2):
    o = OracleEnv(blob)
    o.reset(reset_qpos_scene_b(A, env))
    if settle:
        o.move_group("All", TGT, 1e-7, settle)
    q, v = o.qpos.copy(), o.qvel.copy()
    o.reset(q, v); o.forward()
    qq = np.stack([q, q]); vv = np.stack([v, v])
    eng.set_state(qq, vv)
    gc = eng.debug_forward(0, "contact").reshape(-1, 16)
    oc = o.contacts()
    print("env", env, "ncon", len(gc), o.ncon)
    for i in range(min(len(gc), len(oc))):
        d = np.abs(gc[i, :13] - oc[i, :13]).max()
        if d > 1e-9 or not np.array_equal(gc[i, 13:], oc[i, 13:]):
            g1, g2 = int(oc[i, 13]), int(oc[i, 14])
            print(" contact", i, "geoms", g1, g2, "types", A["geom_type"][g1], A["geom_type"][g2], "maxdiff", d)
            print("   gpu", gc[i]); print("   orc", oc[i])
    for f in ("qacc", "qfrc_constraint"):
        g, r = eng.debug_forward(0, f), o.field(f)
        print(" ", f, "maxdiff", np.abs(g - r).max(), "scale", np.abs(r).max(), "niter gpu", eng.debug_forward(0, "niter"), "orc", o.solver_iter)
    # trajectory
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if steps:
        o.reset(q, v)
        eng.set_state(qq, vv)
        eng.move_group("All", np.stack([TGT, TGT]), 1e-7, steps); eng.run()
        o.move_group("All", TGT, 1e-7, steps)
        gq, gv = eng.get_state()
        dq = np.abs(gq[0].cpu().numpy() - o.qpos)
        print("  traj", steps, "max dq", dq.max(), "at", int(dq.argmax()), "status", eng.status().cpu().numpy())
