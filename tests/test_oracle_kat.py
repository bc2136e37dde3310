"""CPU tests: the oracle against every known-answer vector the reference offers for this path and against
physics invariants.  The arithmetic of `sim.step()` lives in MuJoCo (absent, unpinned), so these pins are WEAK: parity with
the real mujoco_py binary stays unpinned (see oracle/grasp_oracle.c header and DESIGN.md)."""
import json
import os

import numpy as np
import pytest

from tests.common import HOME, object_positions, reset_qpos_scene_a

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture()
def orc(scene_a):
    from oracle.oracle_py import OracleEnv

    o = OracleEnv(scene_a[0])
    yield o
    o.close()


def test_camera_math_golden(orc):
    """pixel_2_world / depth_2_meters against outputs of the REFERENCE's own methods (tests/golden/make_golden.py),
    including the media/console.png known answer (136, 80, 1.11) -> (-0.16551974, -0.50804459, 0.88999999)."""
    g = json.load(open(os.path.join(GOLD, "camera_math.json")))
    for px, py, d, w in zip(g["px"], g["py"], g["depth"], g["pixel_2_world"]):
        assert np.abs(orc.pixel_2_world(px, py, d) - np.array(w)).max() < 1e-12
    assert np.abs(orc.pixel_2_world(136, 80, 1.11) - [-0.16551974, -0.50804459, 0.88999999]).max() < 2e-8  # 8 printed digits
    for gl, m in zip(g["gl_depth"], g["depth_2_meters"]):
        assert abs(orc.depth_2_meters(gl) - m) < 1e-9 * max(1.0, m)


def test_ik_known_answer(orc, scene_a):
    """IK KAT of SURVEY 8(c).2: target (0,-0.6,0.95) -> [-1.7522, -1.2222, 1.6600, -2.0086, -1.5708] (ikpy-equivalent,
    agreement <= 1e-3 rad), and forward kinematics of the MJCF arm puts ee_link on the offset target with its x axis down."""
    blob, A, names = scene_a
    q5 = orc.ik([0.0, -0.6, 0.95])
    assert np.abs(q5 - [-1.7522, -1.2222, 1.6600, -2.0086, -1.5708]).max() < 1e-3
    q = reset_qpos_scene_a(A, 0)
    q[:5] = q5
    orc.reset(q)
    ee = names["body"].index("ee_link")
    xpos = orc.field("xpos").reshape(-1, 3)[ee]
    xmat = orc.field("xmat").reshape(-1, 3, 3)[ee]
    assert np.abs(xpos - (np.array([0.0, -0.6, 0.95]) + [0, -0.005, 0.16])).max() < 1e-9
    assert np.abs(xmat[:, 0] - [0, 0, -1]).max() < 1e-9
    assert orc.ik([5.0, 5.0, 5.0]) is None  # unreachable -> "No valid joint angles"


def test_mass_matrix_matches_jacobian_formula(orc, scene_a):
    """CRBA in the oracle vs the independent Jacobian-sum formula of the model compiler, random configuration."""
    from mujoco_rl_ur5_b200.model.rigid import Kinematics

    blob, A, _ = scene_a
    rng = np.random.RandomState(3)
    q = reset_qpos_scene_a(A, 1)
    q[:8] += rng.uniform(-0.5, 0.5, 8)
    for i in range(6):
        quat = rng.normal(size=4)
        q[8 + 7 * i + 3:8 + 7 * i + 7] = quat / np.linalg.norm(quat)
    orc.reset(q)
    orc.forward()
    M = orc.dense_M(A["dof_parentid"], A["dof_Madr"])
    M2 = Kinematics(A).mass_matrix(q)
    assert np.abs(M - M2).max() < 1e-12 * np.abs(M2).max() + 1e-13
    assert np.linalg.eigvalsh(M).min() > 0


def test_free_fall(orc, scene_a):
    """objects dropped from the stack fall with g = 9.81: z(t) = -1/2 g t^2 under semi-implicit Euler (exact recurrence)"""
    blob, A, _ = scene_a
    q = np.array(A["qpos0"]).copy()
    q[:7] = HOME
    q[7] = 0.3
    for i in range(6):  # spread the objects so that they do not touch while falling
        q[8 + 7 * i] = -0.2 + 0.08 * i
        q[8 + 7 * i + 2] = 0.3
    orc.reset(q)
    n, h = 50, 0.002
    orc.step(n)
    z = object_positions(A, orc.qpos)[:, 2] - (object_positions(A, q)[:, 2])
    expect = -9.81 * h * h * n * (n + 1) / 2  # v_k = -g h k, z_n = sum_k h v_k
    assert np.abs(z - expect).max() < 1e-9


def test_saturated_slew_rate(orc, scene_a):
    """media/plot_1.png: under saturated control the joints slew at gear*ctrl_max/damping*h per sub-step:
    101*2/65*0.002 = 0.0062 rad (pan, lift, elbow), 101*1/45*0.002 = 0.0045 rad (wrist_1)."""
    blob, A, _ = scene_a
    q = reset_qpos_scene_a(A, 0)
    orc.reset(q)
    orc.target[:] = HOME + np.array([1.5, -1.0, -1.0, 1.2, 0, 0, 0])
    rates = []
    prev = orc.qpos[:4].copy()
    for k in range(120):
        orc.move_group("All", None, 1e-9, 0)  # PID only: steps > max_steps ends the loop before sim.step()
        orc.step()
        if k >= 100:
            rates.append(np.abs(orc.qpos[:4] - prev))
        prev = orc.qpos[:4].copy()
    r = np.mean(rates, axis=0)
    pan_rate, wrist_rate = 101 * 2 / 65 * 0.002, 101 * 1 / 45 * 0.002
    assert abs(r[0] - pan_rate) < 0.05 * pan_rate, r          # pan: no gravity load
    assert abs(r[3] - wrist_rate) < 0.02 * wrist_rate, r      # wrist_1: light link
    assert np.all(np.abs(r[1:3] - pan_rate) < 0.25 * pan_rate), r  # lift / elbow: same motor, gravity-loaded (plot slope is approximate)


def test_objects_rest_on_table(orc, scene_a):
    """after the 1000 ms settle every object rests on the table top (z = 0.91) at its half-height + the contact margin"""
    blob, A, _ = scene_a
    orc.reset(reset_qpos_scene_a(A, 0))
    orc.stay(1000)
    pos = object_positions(A, orc.qpos)
    half = np.array([A["geom_size"][30 + i][0] for i in range(6)])
    assert np.abs(pos[:, 2] - (0.91 + half + 1e-3)).max() < 2e-4
    assert np.abs(orc.qvel[8:]).max() < 1e-3
    assert np.abs(orc.qpos[:7] - HOME).max() < 0.02  # PD control holds the arm against gravity within the sag


def test_contact_forces_carry_the_weight_at_rest(orc, scene_a):
    """Newton's third law at rest: the generalised constraint force on every resting object is exactly its weight (vertical), with no
    horizontal component and no torque (friction rows inactive), and the solved acceleration is zero - an analytic fixed point of
    the soft-contact model that holds whatever the impedance parameters are"""
    blob, A, _ = scene_a
    orc.reset(reset_qpos_scene_a(A, 0))
    orc.stay(1000)
    orc.forward()
    qfc, qacc = orc.field("qfrc_constraint"), orc.field("qacc")
    g = abs(float(np.asarray(A["opt_gravity"]).ravel()[2]))
    first = int(np.asarray(A["nbody"]).ravel()[0]) - 6
    for b in range(first, first + 6):
        d, m = int(A["body_dofadr"][b]), float(A["body_mass"][b])
        assert abs(qfc[d + 2] - m * g) < 1e-8 * m * g + 1e-10, (b, qfc[d + 2], m * g)
        assert np.abs(qfc[d:d + 2]).max() < 1e-8 and np.abs(qfc[d + 3:d + 6]).max() < 1e-8
        assert np.abs(qacc[d:d + 6]).max() < 1e-8
    con = orc.contacts()
    assert len(con) >= 6 and (con[:, 0] < 1e-3 + 1e-9).all()  # every reported contact is inside the 1 mm margin


def test_newton_satisfies_kkt_where_pgs_does_not(orc, scene_a):
    """The MJCF sets no solver -> MuJoCo's Newton (UR5gripper_2_finger.xml:19-22).  On this model PGS with the MJCF budget of
    100 iterations is far from converged at first finger contact, Newton reaches the optimum (solver decision)."""
    blob, A, _ = scene_a
    orc.reset(reset_qpos_scene_a(A, 0))
    orc.stay(1000)
    p = object_positions(A, orc.qpos)[0]
    orc.move_ee([p[0], p[1], 1.1], 0.05, 1000)
    orc.move_group("Gripper", [0.0], 0.05, 1000)
    orc.move_ee([p[0], p[1], p[2] + 0.01], 0.01, 300)
    orc.stay(100)
    orc.target[6] = -0.4
    finger_bodies = {14, 16}
    for _ in range(60):  # close until a finger touches the box
        orc.move_group("Gripper", None, 1e-9, 0)
        orc.forward()
        if any(int(A["geom_bodyid"][int(c[14])]) in finger_bodies for c in orc.contacts()):
            break
        orc.step()
    else:
        pytest.fail("fingers never touched the box")
    nv = 44
    viol = {}
    for solver in ("newton", "pgs"):
        orc.set_solver(solver)
        orc.forward()
        J = orc.field("efc_J").reshape(-1, nv)
        R, aref, f = orc.field("efc_R"), orc.field("efc_aref"), orc.field("efc_force")
        M = orc.dense_M(A["dof_parentid"], A["dof_Madr"])
        res = J @ np.linalg.solve(M, J.T @ f) + R * f + J @ orc.field("qacc_smooth") - aref
        # KKT of the dual: residual = 0 where f != 0 (and for the equality row 0), residual >= 0 where f = 0
        active = (f != 0) | (np.arange(len(f)) == 0)
        viol[solver] = max(np.abs(res[active]).max(initial=0.0), np.maximum(0.0, -res[~active]).max(initial=0.0)) / max(1.0, np.abs(aref).max())
    orc.set_solver("newton")
    assert viol["newton"] < 1e-8, viol
    # run the gripper closing with both solvers from the same state: Newton clamps the box (gripper blocked -> "max. steps"),
    # PGS at the MJCF's 100-iteration budget has not converged on the stiff pyramid rows and ejects the box
    q, v = orc.qpos.copy(), orc.qvel.copy()
    box0 = object_positions(A, q)[0]
    outcome = {}
    for solver in ("newton", "pgs"):
        orc.set_solver(solver)
        orc.reset(q, v)
        r, _ = orc.move_group("Gripper", [-0.4], 0.01, 300)
        outcome[solver] = (r, float(np.abs(object_positions(A, orc.qpos)[0] - box0).max()), float(np.abs(orc.qvel).max()))
    orc.set_solver("newton")
    assert outcome["newton"][0] == 2 and outcome["newton"][1] < 0.02, outcome
    assert outcome["pgs"][1] > 0.2 or outcome["pgs"][2] > 50.0, outcome


def test_grasp_attempt_runs_and_is_deterministic(scene_a):
    from oracle.oracle_py import OracleEnv

    blob, A, _ = scene_a
    out = []
    for _ in range(2):
        o = OracleEnv(blob)
        o.reset(reset_qpos_scene_a(A, 0))
        o.stay(1000)
        p = object_positions(A, o.qpos)[0]
        r, info = o.move_and_grasp([p[0], p[1], p[2] + 0.02], 0)
        out.append((r, info, o.qpos.copy()))
        o.close()
    assert out[0][0] == out[1][0] == 1  # the 4 cm box is picked and carried to the drop bin
    assert out[0][1] == out[1][1] and np.array_equal(out[0][2], out[1][2])
    assert out[0][1][1] in range(200, 500)  # pre-grasp phase length in the range media/console.png shows (362)


def test_newton_converges_in_about_one_iteration_with_warm_start(scene_a):
    """solver statistics over one whole grasp attempt (what the GPU kernel work is planned against): with
    qacc_warmstart the primal Newton solver needs one Hessian factorisation in > 85 % of the sub-steps and the active set always
    changes between two factorisations of one solve (so there is nothing to reuse inside a solve)"""
    from oracle.oracle_py import OracleEnv, newton_stats

    blob, A, _ = scene_a
    o = OracleEnv(blob)
    o.reset(reset_qpos_scene_a(A, 3))
    o.stay(1000)
    newton_stats()
    p = object_positions(A, o.qpos)[0]
    o.move_and_grasp([p[0], p[1], p[2] + 0.02], 0, 0.91)
    s = newton_stats()
    o.close()
    assert s["solves"] > 1000 and s["builds"] / s["solves"] < 1.5
    assert s["solves_1_build"] / s["solves"] > 0.85
    assert s["builds_same_active_set"] <= 0.02 * s["builds"]
    assert 1.0 <= s["linesearch_iterations"] / s["builds"] < 6.0


def test_render_statistics_match_the_reference_mean_and_std(scene_a):
    """the reference's pickled `mean_and_std` (normalize.py over 100 resets of real MuJoCo renders) holds the depth mean
    1.5318247 m and std 0.4265042 m of the 200x200 top-down observation (float32 => taken after depth_2_meters) and the RGB means
    108.3 / 120.3 / 132.3.  The oracle's ray-cast observation of the 6-object scene reproduces the depth statistics to 1e-3 (camera pose,
    fovy, table extent and depth conversion all enter); colour is flat-shaded here, so the RGB means are only required within 15 %"""
    from oracle.oracle_py import OracleEnv

    blob, A, _ = scene_a
    cam = int(np.asarray(A["cam_top_down"]).ravel()[0])
    depth, rgb = [], []
    for i in range(6):
        o = OracleEnv(blob)
        o.reset(reset_qpos_scene_a(A, i))
        o.stay(1000)
        r, d = o.render(cam, 200, 200)
        depth.append(d); rgb.append(r)
        o.close()
    depth, rgb = np.stack(depth).astype(np.float64), np.stack(rgb).astype(np.float64)
    assert abs(depth.mean() - 1.5318247) < 2e-3, depth.mean()
    assert abs(depth.std() - 0.4265042) < 2e-3, depth.std()
    ref_rgb = np.array([108.29777875, 120.32914675, 132.30339475])
    assert (np.abs(rgb.mean(axis=(0, 1, 2)) / ref_rgb - 1) < 0.15).all(), rgb.mean(axis=(0, 1, 2))


# ---------------------------------------------------------------------------------------------- media/console.png
def _console_attempt(blob, A):
    """the attempt of media/console.png — pixel (136, 80) -> (-0.16551974, -0.50804459, 0.88999999), nothing there — on the oracle,
    started like a mid-episode step (the arm returns from the drop pose of a previous attempt)"""
    from oracle.oracle_py import OracleEnv
    from tests.common import reset_qpos_scene_a

    o = OracleEnv(blob)
    o.reset(reset_qpos_scene_a(A, 0))
    o.stay(1000)
    o.move_and_grasp([0.2, -0.7, 0.92], 0, 0.91)
    r, info = o.move_and_grasp([-0.16551974, -0.50804459, 0.88999999], 0, 0.91)
    o.close()
    return r, info


def test_console_png_plausibility(scene_a):
    """The reference's screenshot shows, for an empty spot of the table: pre-grasp 362 steps, grasp position 136, centre 202, drop 631,
    open 33, reward 0.  The oracle is inside a factor 2 of every phase it can be compared on, and agrees on the outcome.
    (What cannot match is documented in test_console_png_exact_step_counts.)"""
    blob, A, _ = scene_a
    r, info = _console_attempt(blob, A)
    assert r == 0 and info[11] == 0                      # "Did not grasp anything." / "Grasped anything?: False"
    assert info[0] == 1 and 362 / 2 <= info[1] <= 362 * 2  # "Above target , 362 steps"
    assert 202 / 2 <= info[6] <= 202 * 2                 # "Move to center: success , 202 steps"
    assert 33 / 2 <= info[9] <= 33 * 2                   # "Open gripper: success , 33 steps"


@pytest.mark.xfail(strict=True, reason="media/console.png predates the current reference code: it has no 'Rotate gripper' line, which "
                   "GraspingEnv.py:357-360 prints on every attempt, and reports 'success, 136 steps' for a descent to the table surface that "
                   "the current flow (target z = max(TABLE_HEIGHT, z - 0.01), tolerance 0.01, 300 steps, GraspingEnv.py:259-262) cannot reach: "
                   "the fingers touch the table first.  Tried and recorded in DESIGN.md 2: link masses from the MJCF <inertial> tags instead of "
                   "inertiafromgeom (same counts: 311/301/110/1201), tolerances 0.02-0.1 for the drop move (254-288 steps, never 631).  No "
                   "hypothesis reproduces 362/136/202/631/33, so the screenshot cannot pin the physics.")
def test_console_png_exact_step_counts(scene_a):
    blob, A, _ = scene_a
    _, info = _console_attempt(blob, A)
    assert [info[1], info[4], info[6], info[7], info[9]] == [362, 136, 202, 631, 33]
