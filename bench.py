#!/usr/bin/env python
"""bench.py — MuJoCo-equivalent sub-steps/s of the batched grasp engine (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: every busy environment advances `--chunk` iterations of the
reference control loop (7 PID evaluations + one mj_step, MujocoController.py:318-382) while running the grasp program
(GraspingEnv.py:205-386); environments whose attempt finished get the next synthetic waypoint before the next step.
Workload: scene A (UR5gripper_2_finger.xml), 4096 envs per GPU, physics + control only in the timed `value`
(BASELINE configs[1]/[2] scene and env count); `e2e` goes through BatchedGraspEnv.step (host actions in, pixel_2_world,
whole attempts, 200x200 RGB-D render, host rewards out).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
Under torchrun every rank owns one GPU and its own 4096 envs (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mujoco_substeps_per_sec"
UNIT = "substeps/s"
STATE_BYTES_PER_SUBSTEP = 2 * (50 + 44 + 44 + 14) * 8  # fp64 read+write of qpos, qvel, qacc_warmstart, 7 PID inputs, 7 targets (scene A)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 8:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}
        return out


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def _oracle_worker(args):
    seed, n_attempts, budget_s = args
    from mujoco_rl_ur5_b200.batched_env import scene_a_reset_qpos
    from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob
    from oracle.oracle_py import OracleEnv

    blob = load_scene_blob("A")
    A, _ = load_scene("A")
    o = OracleEnv(blob)
    o.reset(scene_a_reset_qpos(A, seed))
    o.stay(1000)
    rng = np.random.RandomState(30000 + seed)
    s0, t0 = o.substeps, time.perf_counter()
    for _ in range(n_attempts):
        xyz = [rng.uniform(-0.2, 0.2), rng.uniform(-0.75, -0.45), 0.92]
        o.move_and_grasp(xyz, int(rng.randint(0, 6)), 0.91)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    n = o.substeps - s0
    o.close()
    return n, dt


def cpu_baseline_single(budget_s=10.0):
    n, dt = _oracle_worker((20000, 1000, budget_s))
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle restatement (not mujoco_py), 1 env, synthetic-waypoint grasp attempts for {dt:.1f} s = {n} sub-steps"}


_W = {}


def _ref_init():
    from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob
    from oracle.oracle_py import OracleEnv

    _W["blob"] = load_scene_blob("A")
    _W["A"] = load_scene("A")[0]
    _W["env"] = OracleEnv(_W["blob"])
    _W["episode"] = 0


def _ref_work(args):
    """one worker = one oracle env (one host thread): reset + settle + grasp attempts until `budget_s` of wall time is used"""
    worker, budget_s = args
    from mujoco_rl_ur5_b200.batched_env import scene_a_reset_qpos

    o = _W["env"]
    t0 = time.perf_counter()
    s0 = o.substeps
    seed = 20000 + worker + 1000 * _W["episode"]
    _W["episode"] += 1
    o.reset(scene_a_reset_qpos(_W["A"], seed))
    o.stay(1000)
    rng = np.random.RandomState(30000 + seed)
    while time.perf_counter() - t0 < budget_s:
        xyz = [rng.uniform(-0.2, 0.2), rng.uniform(-0.75, -0.45), 0.92]
        o.move_and_grasp(xyz, int(rng.randint(0, 6)), 0.91)
    return o.substeps - s0


def run_reference(args):
    """--impl reference: the reference's own mujoco_py path cannot be installed, so this arm times the CPU oracle
    restatement with one single-env process per host thread — every worker resets, settles and runs grasp attempts for a
    bounded wall-time budget per step; value = sub-steps of all workers / wall time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    budget = args.ref_step_seconds
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores, initializer=_ref_init) as pool:
        total_n, total_t = 0, 0.0
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_work, [(w, budget) for w in range(cores)], chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                total_n += sum(res)
                total_t += dt
    value = total_n / total_t
    sample = f"{cores} single-env oracle processes, each step = reset + settle + grasp attempts for >= {budget:.1f} s wall per worker, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": "scene A (UR5gripper_2_finger.xml) grasp attempts, synthetic waypoints; CPU oracle restatement of the reference path (not mujoco_py)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from mujoco_rl_ur5_b200.batched_env import BatchedGraspEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = args.envs
    env = BatchedGraspEnv(N, "A", local, env_index_offset=rank * N)
    eng = env.engine
    env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(30000 + rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def assign_idle():
        idle = (eng.busy() == 0)
        u = torch.rand((N, 3), generator=gen, device=dev, dtype=torch.float64)
        coords = torch.stack([-0.2 + 0.4 * u[:, 0], -0.75 + 0.3 * u[:, 1], torch.full((N,), 0.92, device=dev, dtype=torch.float64)], dim=1)
        rot = torch.randint(0, 6, (N,), generator=gen, device=dev, dtype=torch.int32)
        eng.grasp(coords, rot, 0.91, env_mask=idle.to(torch.uint8))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident stepping
    for _ in range(args.warmup):
        assign_idle()
        eng.run_async(args.chunk)
    barrier()
    sub0 = env.total_substeps()
    l0, s0 = eng.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    evs = []
    kern_evs = []
    for _ in range(args.steps):
        flush.fill_(1.0)  # L2 flush between timed iterations (outside the event pairs)
        a, b, c = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        assign_idle()
        b.record()
        eng.run_async(args.chunk)
        c.record()
        evs.append((a, c))
        kern_evs.append((b, c))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(c) for a, c in evs)
    kern_ms = sum(b.elapsed_time(c) for b, c in kern_evs)
    sub1 = env.total_substeps()
    l1, s1 = eng.counters()
    nsub = sub1 - sub0
    tms = torch.tensor([ms, kern_ms], dtype=torch.float64, device=dev)
    tn = torch.tensor([nsub], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
    ms_max, kern_ms_max, nsub_all = float(tms[0]), float(tms[1]), float(tn[0])
    value = nsub_all / (ms_max * 1e-3)

    # ---- e2e: the public batched gym API with host buffers
    env.reset()
    rng = np.random.RandomState(30000 + rank)

    def table_actions():
        # random pixels over the table region of the image (the agent's random-action filter, Grasping_Agent_multidiscrete.py:267-279)
        px = rng.randint(40, 160, N)
        py = rng.randint(60, 140, N)
        return np.stack([py * 200 + px, rng.randint(0, 6, N)], axis=1)

    env.step(table_actions())  # warm-up attempt
    barrier()
    e_sub0 = env.total_substeps()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        env.step(table_actions())
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    e_n = env.total_substeps() - e_sub0
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    ten = torch.tensor([e_n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(ten, op=dist.ReduceOp.SUM)
    e2e_value = float(ten[0]) / float(te[0])
    # ---- a16: Q-network forward (bf16 tcgen05) on the observations of the envs, chunked; extra object, not the headline
    qinfo = None
    if args.qnet_images > 0:
        from mujoco_rl_ur5_b200.qnet import QNetForward, make_torch_qnet

        torch.manual_seed(0)
        qf = QNetForward(make_torch_qnet(6).state_dict(), local, max_batch=args.qnet_chunk)
        nimg = min(args.qnet_images, N)
        obs = env.current_observation
        sub = {"rgb": obs["rgb"][:nimg], "depth": obs["depth"][:nimg]}
        for _ in range(2):
            act, _ = qf.greedy(qf.forward(qf.obs_to_state(sub, 1.1)))
        barrier()
        l_q0 = qf.launches
        q0, q1 = torch.cuda.Event(True), torch.cuda.Event(True)
        q0.record()
        for _ in range(args.qnet_reps):
            act, _ = qf.greedy(qf.forward(qf.obs_to_state(sub, 1.1)))
        q1.record()
        barrier()
        qms = q0.elapsed_time(q1) / args.qnet_reps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf = nimg * 41.99424e9 / (qms * 1e-3) / 1e12
        pk = float(peaks.get("bf16_tflops_sustained", 1400.0))
        qinfo = {"images": nimg, "chunk": args.qnet_chunk, "ms_per_batch": qms, "images_per_s": nimg / (qms * 1e-3), "tflops": tf,
                 "roofline": {"bound": "tensor", "achieved": tf, "peak": pk, "unit": "TFLOP/s", "frac": tf / pk,
                              "note": "41.99 GFLOP per 200x200 image (SURVEY 8d); peak = " + ("measured sustained cuBLAS bf16" if peaks else "fallback")},
                 "kernels_per_batch": (qf.launches - l_q0) // max(args.qnet_reps, 1)}
    # ---- scene B (the reference's default 40-object scene): extra object, not the headline.  Envs are reset by
    # the reference rule, 500 untimed sub-steps let the objects fall and pile up, then `scene_b_steps` sub-steps are timed.
    sbinfo = None
    if args.scene_b_envs > 0:
        from mujoco_rl_ur5_b200.batched_env import HOME, scene_b_reset_qpos
        from mujoco_rl_ur5_b200.engine import BatchedEngine
        from mujoco_rl_ur5_b200.model.scene import load_scene, load_scene_blob

        nb = args.scene_b_envs
        Ab, _ = load_scene("B")
        engb = BatchedEngine(load_scene_blob("B"), nb, local)
        engb.set_state(np.stack([scene_b_reset_qpos(Ab, 20000 + rank * nb + i) for i in range(nb)]))
        tgt = np.tile(HOME + np.array([0.2, 0.1, -0.1, 0.1, 0.1, 0.3, -0.1]), (nb, 1))
        engb.move_group("All", tgt, 1e-9, 499)
        engb.run()
        barrier()
        b0, b1 = torch.cuda.Event(True), torch.cuda.Event(True)
        b0.record()
        engb.move_group("All", tgt, 1e-9, args.scene_b_steps - 1)
        engb.run()
        b1.record()
        barrier()
        tb = torch.tensor([b0.elapsed_time(b1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        flagged = int((engb.status() != 0).sum().item())
        sbinfo = {"workload": f"scene B UR5gripper_2_finger_many_objects.xml (40 free objects, condim 6, nv 248), {nb} envs/GPU, "
                              f"{args.scene_b_steps} PID+physics sub-steps in the piled-up state (after 500 untimed sub-steps)",
                  "value": world * nb * args.scene_b_steps / (float(tb[0]) * 1e-3), "unit": UNIT, "envs_per_gpu": nb,
                  "ms": float(tb[0]), "workspace": "HBM rows (engine variant hbm)", "envs_flagged": flagged}
        if rank == 0 and args.cpu_seconds > 0:
            from oracle.oracle_py import OracleEnv  # cpu_baseline leg: the checker timed beside the product, never on its path

            o = OracleEnv(load_scene_blob("B"))
            o.reset(scene_b_reset_qpos(Ab, 20000))
            o.move_group("All", tgt[0], 1e-9, 249)
            t0 = time.time()
            o.move_group("All", tgt[0], 1e-9, 79)
            dtb = time.time() - t0
            o.close()
            sbinfo["cpu_baseline"] = {"value": 80 / dtb, "unit": UNIT, "cores": 1, "kind": "port",
                                      "sample": "fp64 oracle, one env, 80 sub-steps after 250 sub-steps of settling"}
        engb.close()
    status = eng.status()
    status_or = int(torch.bitwise_or(status[0], status.max()).item()) if N else 0
    n_flag = int((status != 0).sum().item())

    if rank == 0:
        peak, which = measured_peak()
        per_launch_sub = nsub / max(args.steps, 1)
        achieved = (nsub * STATE_BYTES_PER_SUBSTEP) / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_k_run_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = cpu_baseline_single(args.cpu_seconds)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"scene A UR5gripper_2_finger.xml, {N} envs/GPU, physics+PID+grasp-program sub-steps, synthetic waypoints; "
                                   f"step = {args.chunk} loop iterations per busy env; L2 flushed (256 MiB fill) between timed steps",
                       "envs_per_gpu": N, "chunk": args.chunk, "substeps_per_step": per_launch_sub, "solver": "newton", "parallelism": f"env-shard x{world}"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "note": f"k_run; algorithmic bytes = {STATE_BYTES_PER_SUBSTEP} B/env-sub-step (fp64 state rows), peak = {which} copy bandwidth; "
                                 "this path is fp64-ALU/latency bound, not HBM bound (DESIGN.md)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": env.h2d_bytes_per_step, "d2h_bytes_per_step": env.d2h_bytes_per_step,
                    "steps": args.e2e_steps, "note": "BatchedGraspEnv.step: pinned host actions -> pixel_2_world -> full grasp attempts -> RGB-D render -> host rewards"},
            "gpu_launches": int(l1 - l0), "substep_kernel_launches": int(s1 - s0),
            "clocks": clocks, "env_status_flags": {"envs_flagged": n_flag, "or": status_or}, "qnet": qinfo, "scene_b": sbinfo,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--chunk", type=int, default=256, help="control-loop iterations per busy env and step")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--qnet-images", type=int, default=256, help="images for the Q-net forward leg (0 = skip)")
    ap.add_argument("--qnet-chunk", type=int, default=64)
    ap.add_argument("--qnet-reps", type=int, default=3)
    ap.add_argument("--scene-b-envs", type=int, default=1024, help="environments of the 40-object scene for the scene_b leg (0 = skip)")
    ap.add_argument("--scene-b-steps", type=int, default=100)
    ap.add_argument("--ref-step-seconds", type=float, default=3.0, help="--impl reference: wall-time budget per worker and step")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
