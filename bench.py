#!/usr/bin/env python
"""bench.py -- AGV-steps/s (batched scenes, incl. state encode) on N B200s.

  python bench.py --gpus N --steps K --warmup W [--workload c3|c2|c5] [--impl reference]

A "step" is one agv_rollout launch: TICKS (default 16) passes of Scene.run_mode's loop
(Scene.py:111-164) for every scene resident on the GPU.  In every tick each created AGV
takes its turn in list order = limited-visual state encode (bf16, written to HBM) + A*
guidance (Manhattan-path row DP / BFS on the state matrix) + Explorer.execute_action with
collision test / task FSM / reward (PROTO_INTELLIGENT, POLICY_GUIDED, OBS_CLIP, auto_reset).
Workload c3 = BASELINE.json configs[2]: synthetic 64x64 RMFS layout, 64 AGVs/scene,
16384 scenes per GPU (weak scaling).  Prints ONE JSON line on rank 0.

The timed windows are stationary: the scene state is snapshotted once (agv_get_state) and
restored (agv_set_state) before every window, so `value`, `e2e` and `e2e_closed` time the
same turns; every window is repeated and the median is reported.  After the headline the
default run appends two sub-records: `train` (BASELINE configs[3]: AG-DQN on the README 9x9
layout with the CNN in the loop, NCCL gradient all-reduce when launched on several ranks) and
`expert` (configs[4]: A_star-mode rollouts of 2^20 AGVs on a 128x128 layout pushed into the
device replay buffer).
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "reinforcement-learning-for-multi-agv-pathfinding_b200"

METRIC = "AGV-steps/sec (batched envs, incl. state encode)"
UNIT = "AGV-steps/s"


# ----------------------------------------------------------------------------- workloads
def rect_layout(w, h, nx, ny, ps):
    """Layout geometry of the reference (Layout.py:29-30,83-102) -- bench-local copy so
    that the measured path never imports oracle/."""
    W = nx * (w + 1) + 1
    H = ny * (h + 1) + 1 + 3
    g = np.zeros((H, W), dtype=np.uint8)
    for iy in range(ny):
        for dy in range(h):
            for ix in range(nx):
                for dx in range(w):
                    g[iy * (h + 1) + dy + 1, ix * (w + 1) + dx + 1] = 1
    for i in range(ps):
        g[H - 2, int(W * (i + 1) / (ps + 1) + 1) - 1] = 2
    return g


def layout_128():
    """128x128 layout_list (the rect generator cannot make 128: nx*(w+1)=127 is prime,
    SURVEY 8a): 6x2 islands with 1-cell aisles, last rows track, picking on row 127."""
    g = np.zeros((128, 128), dtype=np.uint8)
    for iy in range(41):
        for dy in range(2):
            for ix in range(18):
                for dx in range(6):
                    g[iy * 3 + dy + 1, ix * 7 + dx + 1] = 1
    for i in range(16):
        g[126, int(128 * (i + 1) / 17 + 1) - 1] = 2
    return g


README_9x9 = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 1, 1, 1, 1, 1, 1, 0, 0], [0, 1, 1, 1, 1, 1, 1, 0, 0],
                       [0, 0, 0, 0, 0, 0, 1, 1, 0], [0, 0, 0, 2, 0, 0, 1, 1, 0], [0, 0, 0, 0, 0, 0, 1, 1, 0],
                       [0, 1, 1, 1, 1, 1, 1, 0, 0], [0, 1, 1, 1, 1, 1, 1, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0]],
                      dtype=np.uint8)

WORKLOADS = {
    # name: (grid fn, AGVs/scene, scenes/GPU, description)
    "c3": (lambda: rect_layout(6, 2, 9, 20, 8), 64, 16384,
           "configs[2]: synthetic 64x64 RMFS rect(6,2,9,20,8), 64 AGVs/scene, 16384 scenes/GPU"),
    "c2": (lambda: README_9x9, 4, 4096, "configs[1]: README 9x9 layout_list, 4 AGVs/scene, 4096 scenes/GPU"),
    "c5": (layout_128, 64, 16384, "configs[4]: 128x128 layout_list, 64 AGVs/scene, 16384 scenes/GPU (2^20 AGVs)"),
}


def synth_tasks(grid, n_env, seed):
    """Layout.__create_task's distribution (each storage station once, in random order,
    paired with a uniformly random picking station; Layout.py:104-116) drawn with numpy
    instead of Python's `random` -- synthetic data."""
    ss = np.argwhere(grid == 1)  # (row, col)
    ps = np.argwhere(grid == 2)
    rng = np.random.default_rng(seed)
    T = len(ss)
    out = np.empty((n_env, T, 4), dtype=np.uint8)
    for e in range(n_env):
        perm = rng.permutation(T)
        pj = rng.integers(0, len(ps), T)
        out[e, :, 0] = ss[perm, 1] + 1
        out[e, :, 1] = ss[perm, 0] + 1
        out[e, :, 2] = ps[pj, 1] + 1
        out[e, :, 3] = ps[pj, 0] + 1
    return out


def max_training_steps(W, H):
    return int((W / 2 + H) * 3 * ((W * H) / 2.5))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
class CpuArm:
    """Times the CPU port of the reference path (oracle/agv_oracle.c, pinned against the
    live Python reference) on the host cores: one scene per thread, same protocol as the
    GPU step (encode + guided BFS + execute per turn, auto-reset)."""

    def __init__(self, workload, n_threads, seed=0):
        from oracle import oracle as orc  # the one place bench.py may execute oracle/
        self.orc = orc
        gridf, N, _, _ = WORKLOADS[workload]
        grid = gridf()
        tasks = synth_tasks(grid, n_threads, seed + 17)
        self.n = n_threads
        self.envs = [orc.Env(grid, tasks[i], N) for i in range(n_threads)]

    def batch(self, nticks):
        """every thread advances its scene by nticks; returns (turns, seconds)"""
        orc = self.orc
        out = [0] * self.n

        def run(i):
            out[i] = self.envs[i].run(orc.PROTO_INTELLIGENT, orc.POLICY_GUIDED, nticks, encode_P=3)[0]

        th = [threading.Thread(target=run, args=(i,)) for i in range(self.n)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return sum(out), time.perf_counter() - t0


def _pyref_worker(q_out, root, workload, seed, seconds):
    """one process = one core: the UNMODIFIED Python reference (Scene / Explorer / Layout / StateManager /
    astar, imported from the vendored copy with pygame / matplotlib stubbed) driven in 'intelligent' mode by
    an A*-guidance controller that builds the limited-visual state of the acting AGV and takes the first
    FindPathAstar action (what one GPU turn does: encode + guidance + execute_action);
    counts the turns taken in `seconds` of wall time."""
    import contextlib
    import io
    import random
    os.environ["AGV_REFERENCE_ROOT"] = root
    sys.path.insert(0, ROOT)
    from oracle import ref_harness as rh
    gridf, N, _, _ = WORKLOADS[workload]
    grid = gridf()
    R = rh.load()
    steps = 0
    t_end = time.perf_counter() + seconds

    class Stop(Exception):
        pass

    class Guided:
        def __init__(self):
            self.sm = R.StateManager()

        def choose_action(self, all_info, name):
            nonlocal steps
            if time.perf_counter() >= t_end:
                raise Stop()
            obs, cp, tp, valid = self.sm.create_state(all_info, name, obs_clip=True, padding_size=3)
            f = R.astar.FindPathAstar(valid, (cp[0] - 1, cp[1] - 1), (tp[0] - 1, tp[1] - 1))
            found, _pl, _pm, al = f.run_astar_method()
            steps += 1
            return 4 if (found is False or len(al) == 0) else {"UP": 0, "RIGHT": 1, "DOWN": 2, "LEFT": 3}[al[0]]

        def store_info(self, all_info, reward, is_end, name):
            pass

    ep = 0
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        while time.perf_counter() < t_end:
            scene = rh.make_scene(N, layout_list=grid.tolist(), seed=seed * 1000 + ep, flags={})
            try:
                scene.run_game(control_pattern="intelligent", smart_controller=Guided(), render=False)
            except Stop:
                pass
            ep += 1
    q_out.put((steps, time.perf_counter() - t0))


class PyRefArm:
    """The reference's own CPU implementation of the path: one worker process per host core, each running
    whole episodes of the unmodified Python scene (baseline/_ref, vendored by build(); /root/reference in the
    build container)."""

    def __init__(self, workload, n_procs):
        self.workload, self.n = workload, n_procs
        self.root = None
        for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
            if os.path.isdir(os.path.join(cand, "src", "multiAGVscene")):
                self.root = cand
                break

    def available(self):
        return self.root is not None

    def sample(self, seconds, seed=0):
        """(AGV-steps, wall seconds) of n processes each working for `seconds`"""
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        ps = [ctx.Process(target=_pyref_worker, args=(q, self.root, self.workload, seed * 100 + i, seconds))
              for i in range(self.n)]
        t0 = time.perf_counter()
        for p in ps:
            p.start()
        res = [q.get() for _ in ps]
        for p in ps:
            p.join()
        return sum(r[0] for r in res), time.perf_counter() - t0, max(r[1] for r in res)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- helpers
def median(xs):
    xs = sorted(xs)
    n = len(xs)
    return xs[n // 2] if n % 2 else 0.5 * (xs[n // 2 - 1] + xs[n // 2])


def config_dict(desc, grid, N, E, ticks, policy, obs):
    """the workload description -- identical in our arm and in the reference arm"""
    H, W = grid.shape
    return {"workload": desc, "H": int(H), "W": int(W), "agvs_per_scene": N, "scenes_per_gpu": E,
            "tasks_per_scene": int((grid == 1).sum()), "ticks_per_step": ticks,
            "protocol": ("intelligent + A* guidance (matrix b)" if policy == "guided" else "A_star control mode (matrix a)"),
            "obs": "clip 3x7x7 bf16" if obs == "clip" else "full 3x%dx%d bf16" % (H, W)}


def load_json(name):
    try:
        with open(os.path.join(ROOT, name)) as f:
            return json.load(f)
    except Exception:
        return {}


# ----------------------------------------------------------------------------- sub-records
def sub_train(pkg, lib, du, dev, local_rank, world, steps, matched=False):
    """BASELINE configs[3]: AG-DQN training on the README 9x9 layout (4 AGVs per scene), CNN in the loop
    (sub-step protocol), device PER, DDQN learner in CUDA graphs; with several ranks the scenes and the
    replay shard per rank and the flat gradient is all-reduced over NCCL once per update.
    matched=True: the reference's schedule -- one batch-32 update per stored transition (MADQN.py:113-171) --
    on 32 scenes, so that the learner work per AGV-step is like for like."""
    import torch
    import torch.distributed as dist
    gridf, N, E, desc = WORKLOADS["c2"]
    grid = gridf()
    if matched:
        E = 32
    rank = dist.get_rank() if world > 1 else 0
    sc = pkg.BatchedScene(grid, E, N, tasks=synth_tasks(grid, E, seed=2000 + rank), P=3, auto_reset=True, device=local_rank)
    sc.reset()
    M = importlib.import_module(PKG + ".algorithm.MADQN_structure.MADQN")
    if matched:
        agent = M.BatchedAgent(sc, memory_size=50000, batch_size=32, epsilon=0.8, ddp=world > 1, seed=rank,
                               updates_per_transition=1.0)
    else:
        agent = M.BatchedAgent(sc, memory_size=1 << 20, batch_size=256, epsilon=0.8, ddp=world > 1, seed=rank)
    for _ in range(3 if matched else 2 * N + 4):
        agent.tick(train=True)
    sc.counters()
    u0, s0 = agent.learn_step_counter, agent.transitions_stored()
    l0 = lib.agv_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    du.barrier(dev)
    ev0.record()
    for _ in range(steps):
        agent.tick(train=True)
    ev1.record()
    du.barrier(dev)
    ms, turns = du.reduce_timing(ev0.elapsed_time(ev1), sc.counters()[0], dev)
    updates, stored = agent.learn_step_counter - u0, agent.transitions_stored() - s0
    rec = {"workload": ("AG-DQN training, reference schedule (1 batch-32 update per stored transition): "
                        if matched else "AG-DQN training: ") + desc.replace("4096 scenes", "%d scenes" % E),
           "value": turns / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "turns_per_step": turns / steps, "scenes_per_gpu": E, "n_gpus": world,
           "learner_updates": updates, "transitions_stored_rank0": stored,
           "updates_per_stored_transition": updates / max(1, stored), "batch_size": agent.batch_size,
           "gpu_launches": int(lib.agv_launch_count() - l0)}
    if world > 1:
        # the one collective of the system, timed alone: all-reduce of the flat fp32 gradient
        n = 20
        ev0.record()
        for _ in range(n):
            dist.all_reduce(agent._flat)
        ev1.record()
        torch.cuda.synchronize()
        rec["collective"] = {"op": "nccl all_reduce fp32", "elements": int(agent._flat.numel()),
                             "per_update": 1, "ms_per_all_reduce": ev0.elapsed_time(ev1) / n,
                             "overlap": "issued on a side stream; the next sub-step's encode / act forward runs meanwhile"}
    sc.close()
    return rec


def sub_expert(pkg, lib, du, dev, local_rank, world, launches, ticks_per_launch=8):
    """BASELINE configs[4]: Expert-mode rollout collection -- A_star control mode on a 128x128 layout,
    2^20 AGVs per GPU, every turn's (state, expert action, reward, next state) pushed into the device replay"""
    import torch
    gridf, N, E, desc = WORKLOADS["c5"]
    grid = gridf()
    rank = int(os.environ.get("RANK", "0"))
    sc = pkg.BatchedScene(grid, E, N, tasks=synth_tasks(grid, E, seed=3000 + rank), P=3, auto_reset=True, device=local_rank)
    sc.reset()
    rp = pkg.DeviceReplay(1 << 22, (3, sc.K, sc.K), device=local_rank)
    sc.collect_expert(rp, 2 * N, ticks_per_launch=ticks_per_launch)
    sc.counters()
    l0 = lib.agv_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    du.barrier(dev)
    ev0.record()
    pushed = sc.collect_expert(rp, launches * ticks_per_launch, ticks_per_launch=ticks_per_launch)
    ev1.record()
    du.barrier(dev)
    ms, turns = du.reduce_timing(ev0.elapsed_time(ev1), int(pushed.item()), dev)
    rec = {"workload": "Expert rollouts into device replay: " + desc, "value": turns / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms / launches, "steps": launches, "ticks_per_step": ticks_per_launch,
           "turns_per_step": turns / launches, "scenes_per_gpu": E, "agvs_per_gpu": E * N, "n_gpus": world,
           "replay_capacity": 1 << 22, "replay_entries_rank0": rp.info()[0],
           "row": "clip obs + next obs bf16, action, reward, done",
           "kernel": lib.agv_last_tick_kernel().decode(), "gpu_launches": int(lib.agv_launch_count() - l0)}
    sc.close()
    rp.close()
    return rec


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=0, help="override scenes per GPU")
    ap.add_argument("--agvs", type=int, default=0, help="override AGVs per scene (tuning)")
    ap.add_argument("--ticks", type=int, default=16, help="ticks per step (one agv_rollout launch); 0 = one agv_tick per step")
    ap.add_argument("--reps", type=int, default=5, help="repetitions of every timed window (the median is reported)")
    ap.add_argument("--spinup", type=int, default=-1,
                    help="untimed ticks before warm-up (default: 6*AGVs -- the scenes all start at Scene.init and the "
                         "workload swings for ~250 ticks before it is stationary: tuning/drift.py)")
    ap.add_argument("--mode", default="env", choices=["env", "train", "train-matched", "expert", "substep"],
                    help="env: the headline (+ train / expert sub-records); train / expert: one sub-record alone; "
                         "substep: the per-AGV callback protocol (encode -> step per AGV index) replayed from one CUDA "
                         "graph per tick")
    ap.add_argument("--obs", default="clip", choices=["clip", "full"],
                    help="clip: limited visual 3x7x7 (AG-DQN, headline); full: 3xHxW (PG/AC/DQN controllers)")
    ap.add_argument("--policy", default="guided", choices=["guided", "astar"],
                    help="guided: intelligent protocol + A* guidance on matrix b (headline); "
                         "astar: Scene A_star control mode (matrix a, no episode end)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the train / expert sub-records")
    ap.add_argument("--shaped", action="store_true", help="settings.Sparse_Reward shaped reward (a second path search per turn)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--given", action="store_true",
                    help="device-timed leg with POLICY_GIVEN and an all -1 action tensor resident on the device (what the "
                         "e2e legs run, minus the copies)")
    ap.add_argument("--ref-seconds", type=float, default=4.0, help="--impl reference: seconds of Python scene per step and core")
    args = ap.parse_args()
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    gridf, N, E_default, desc = WORKLOADS[args.workload]
    if args.agvs > 0:  # tuning: other congestion levels on the same layout
        N = args.agvs
        desc += " [AGVs per scene overridden: %d]" % N
    E = args.scenes or E_default
    grid = gridf()
    H, W = grid.shape
    K = max(0, args.ticks)
    cfg = config_dict(desc, grid, N, E, max(1, K), args.policy, args.obs)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        nthr = host_threads()
        # the oracle C port (pinned against the live Python reference), one scene per host thread
        tps = 50 if args.workload != "c2" else 2000
        arm = CpuArm(args.workload, nthr)
        arm.batch(2 * N + 8)
        pr = [arm.batch(tps) for _ in range(3)]
        port_v = sum(p[0] for p in pr) / max(1e-9, sum(p[1] for p in pr))
        port = {"value": port_v, "unit": UNIT, "cores": nthr, "kind": "port",
                "sample": "%d host threads x 1 scene x %d ticks x 3 (after %d spin-up ticks), oracle C port" % (nthr, tps, 2 * N + 8)}
        py = PyRefArm(args.workload, nthr)
        if py.available():
            # the reference's own implementation: one process per core, every step = a bounded sample of
            # `ref_seconds` of whole episodes from Scene.init (the early part of an episode: AGVs spawn one
            # per tick, so fewer are active than in the GPU run's steady state; the cost per AGV-step -- one
            # FindPathAstar call + two limited-visual encodes -- is the same)
            py.sample(min(2.0, args.ref_seconds))  # warm-up (page cache, imports): one short sample
            res = [py.sample(args.ref_seconds, seed=1 + i) for i in range(steps)]
            turns, secs = sum(r[0] for r in res), sum(r[2] for r in res)
            v = turns / max(secs, 1e-9)
            base = {"value": v, "unit": UNIT, "cores": nthr, "kind": "reference",
                    "sample": "unmodified Python Scene/Explorer/StateManager/astar from %s: %d processes x %.1f s of "
                              "'intelligent' episodes with an A*-guidance controller (one limited-visual encode + "
                              "one FindPathAstar per AGV-step, like the GPU step) per step, %d steps, %d AGV-steps" % (
                                  os.path.relpath(py.root, ROOT) if py.root.startswith(ROOT) else py.root, nthr,
                                  args.ref_seconds, steps, turns)}
            ms_step = 1e3 * secs / steps
        else:
            v, base, ms_step = port_v, port, 1e3 * sum(p[1] for p in pr) / 3
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "python float / int lists" if base["kind"] == "reference" else "u8/int32",
                "data": "synthetic", "config": cfg, "cpu_baseline": base, "cpu_port": port,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    lib = pkg._lib.lib()
    du = importlib.import_module(PKG + ".dist_util")

    def finish(line):
        if rank == 0 and line is not None:
            print(json.dumps(line))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    base_line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                 "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic"}

    if args.mode in ("train", "train-matched", "expert"):
        if args.mode == "expert":
            rec = sub_expert(pkg, lib, du, dev, local_rank, world, max(1, steps // 8))
            dt = "u8/u64 bitboards, bf16 obs"
        else:
            rec = sub_train(pkg, lib, du, dev, local_rank, world, steps, matched=args.mode == "train-matched")
            dt = "bf16 autocast CNN, u64 bitboards"
        line = dict(base_line, value=rec["value"], ms_per_step=rec["ms_per_step"], dtype=dt,
                    config={"workload": rec["workload"], "mode": args.mode}, gpu_launches=rec["gpu_launches"], detail=rec)
        return finish(line)

    tasks = synth_tasks(grid, E, seed=1000 + rank)
    sc = pkg.BatchedScene(grid, E, N, tasks=tasks, P=3, auto_reset=True, device=local_rank, shaped_reward=args.shaped)
    del tasks
    sc.reset()
    PROTO, POL, OBS = pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP
    if args.policy == "astar":
        PROTO, POL = pkg.PROTO_ASTAR, pkg.POLICY_ASTAR
    if args.obs == "full":
        OBS = pkg.OBS_FULL
    out = {}

    if args.mode == "substep":
        # the network-in-the-loop decomposition without a network: begin_tick, then for every AGV
        # index k: agv_encode(k) + agv_step_turn(k, action = -1 -> A* guidance), end_tick.  2N+2
        # launches per tick, captured once into a CUDA graph and replayed (launch-latency bound).
        minus1 = torch.full((E,), -1, dtype=torch.int8, device=dev)

        def tick_substeps():
            sc.begin_tick()
            for k in range(N):
                sc.encode(k, OBS)
                sc.step_turn(k, PROTO, minus1)
            sc.end_tick()

        for _ in range(max(3, warmup)):
            tick_substeps()
        torch.cuda.synchronize()
        results = {}
        for label in ("eager", "graph"):
            if label == "graph":
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    tick_substeps()
                run = g.replay
            else:
                run = tick_substeps
            for _ in range(3):
                run()
            sc.counters()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            du.barrier(dev)
            ev0.record()
            for _ in range(steps):
                run()
            ev1.record()
            du.barrier(dev)
            ms, turns = du.reduce_timing(ev0.elapsed_time(ev1), sc.counters()[0], dev)
            results[label] = (turns / (ms * 1e-3), ms / steps, turns / steps)
        v, msps, tps = results["graph"]
        line = dict(base_line, value=v, ms_per_step=msps, dtype="u8/u64 bitboards, bf16 obs",
                    config={"workload": "sub-step protocol (CUDA graph): " + desc, "mode": "substep",
                            "launches_per_tick": 2 * N + 2, "agvs_per_scene": N, "scenes_per_gpu": E},
                    eager={"value": results["eager"][0], "ms_per_step": results["eager"][1]},
                    gpu_launches=(2 * N + 2) * steps, turns_per_step=tps)
        return finish(line)

    # ------------------------------------------------------------------ headline: fused rollout steps
    act_dev = torch.full((max(1, K), E, N) if K > 0 else (E, N), -1, dtype=torch.int8, device=dev) if args.given else None

    def step():
        pol = pkg.POLICY_GIVEN if args.given else POL
        if K > 0:
            sc.rollout(K, PROTO, pol, OBS, actions=act_dev, want_lens=False, out=out)
        else:
            sc.tick(PROTO, pol, OBS, actions=act_dev, want_lens=False, out=out)

    KK = max(1, K)
    spin = args.spinup if args.spinup >= 0 else 6 * N
    for _ in range((spin + KK - 1) // KK):
        step()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    snap = sc.get_state()  # every timed window below starts from this state: same turns in every leg

    def timed(fn, k, after=None):
        """restore the snapshot, one untimed call, then k calls bracketed by barrier + synchronize with CUDA
        events on the launching stream; `after` runs before the closing event (joins work queued on other
        streams into the timed region); returns (ms max over ranks, AGV turns over all ranks, launches)"""
        sc.set_state(*snap)
        fn()
        if after is not None:
            after()
        sc.counters()
        l0 = lib.agv_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        du.barrier(dev)
        ev0.record()
        for _ in range(k):
            fn()
        if after is not None:
            after()
        ev1.record()
        du.barrier(dev)
        ms = ev0.elapsed_time(ev1)
        turns = sc.counters()[0]
        launches = lib.agv_launch_count() - l0
        ms, turns = du.reduce_timing(ms, turns, dev)
        return ms, turns, launches

    def windows(fn, after=None, reps=None):
        res = [timed(fn, steps, after) for _ in range(reps or max(1, args.reps))]
        vals = [t / (m * 1e-3) for m, t, _ in res]
        mid = sorted(range(len(res)), key=lambda i: vals[i])[len(res) // 2]
        return res[mid], vals

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    (ms, turns, launches), vals = windows(step)
    value = turns / (ms * 1e-3)

    # ---- end to end through the C-ABI host-buffer entry points (agv_rollout_host / agv_tick_host): per step
    # H2D of the [ticks, E, N] action tensor from pinned memory (all -1 = "use A* guidance", the AG-DQN
    # epsilon-gate's other branch) and D2H of act / reward / done.  Open loop: the copies of neighbouring
    # steps overlap the rollout (double-buffered, library-owned copy streams) and the last copy-backs are
    # joined inside the timed region.  Closed loop: every step's results are on the host before the next step
    # is issued -- what a host-side policy needs.
    guided = args.policy == "guided"
    act_host = torch.full((KK, E, N) if K > 0 else (E, N), -1, dtype=torch.int8).pin_memory()

    def step_host(join):
        pol = pkg.POLICY_GIVEN if guided else POL
        if K > 0:
            sc.rollout_host(K, PROTO, pol, OBS, actions_host=act_host if guided else None, join=join)
        else:
            sc.tick_host(PROTO, pol, OBS, actions_host=act_host if guided else None, join=join)
        if join:
            torch.cuda.current_stream().synchronize()

    (ms_e, turns_e, _), vals_e = windows(lambda: step_host(False), after=sc.host_join)
    e2e_value = turns_e / (ms_e * 1e-3)
    (ms_c, turns_c, _), vals_c = windows(lambda: step_host(True), reps=min(3, max(1, args.reps)))
    clk = clocks.stop() if rank == 0 else None

    # ---- rooflines of the dominant kernel (the rollout kernel is the only kernel of ours in the step)
    peaks = load_json("MEASURED_PEAKS.json")
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    obs_bytes = 3 * 7 * 7 * 2 if args.obs == "clip" else 3 * H * W * 2
    unit_bytes = obs_bytes + 32 + 1 + 4 + 1                  # 332 B per AGV-step (clip)
    per_launch_units = turns / steps / world
    dur_s = ms * 1e-3 / steps
    ach = per_launch_units * unit_bytes / dur_s / 1e9
    prof = load_json(os.path.join("profiles", "rollout_%s.json" % args.workload))
    traffic = prof.get("dram_bytes_per_launch") if prof.get("ticks_per_launch") == KK else None
    sm_mhz = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    issue = None
    # (the committed profile is of the headline configuration only: guided policy, clip obs, sparse reward, the
    # profile's ticks per launch -- other configurations run other kernels / instruction mixes)
    if (prof.get("warp_instr_per_agv_step") and prof.get("ticks_per_launch") == KK and guided and args.obs == "clip"
            and not args.shaped):
        # binding roof of the limited-visual rollout: instruction issue of the scene warps' dependent chains.
        # achieved = warp instructions per launch (ncu smsp__inst_executed.sum per AGV-step of the committed
        # profile x the AGV-steps of this launch) / launch duration; peak = 1 warp instruction per cycle and SM
        # sub-partition
        wi = prof["warp_instr_per_agv_step"] * per_launch_units
        issue_peak = n_sm * 4 * sm_mhz * 1e6
        issue = {"achieved": wi / dur_s / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instr/s",
                 "frac": wi / dur_s / issue_peak, "warp_instr_per_agv_step": prof["warp_instr_per_agv_step"],
                 "source": prof.get("source")}

    line = None
    if rank == 0:
        line = dict(base_line, value=value, ms_per_step=ms / steps, dtype="u8/u64 bitboards, bf16 obs",
                    config=dict(cfg, spinup_ticks=spin, window="state restored before every window (agv_set_state); "
                                "median of %d windows of %d steps" % (max(1, args.reps), steps),
                                l2="outputs larger than L2: %.0f MB of observation records written per step" % (
                                    KK * E * N * obs_bytes / 1e6)),
                    turns_per_step=turns / steps, windows=[round(v, 1) for v in vals],
                    e2e={"value": e2e_value, "unit": UNIT,
                         "h2d_bytes_per_step": (KK * E * N * world) if guided else 0,
                         "d2h_bytes_per_step": KK * E * N * 6 * world, "ms_per_step": ms_e / steps,
                         "turns_per_step": turns_e / steps, "windows": [round(v, 1) for v in vals_e],
                         "note": "agv_rollout_host: pinned host buffers, H2D of the [ticks,E,N] action tensor + rollout + "
                                 "D2H of act/reward/done every step; open loop (the copies of steps k-1 / k+1 overlap the "
                                 "rollout of step k; last copy-backs joined inside the timed region)"},
                    e2e_closed={"value": turns_c / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c / steps,
                                "turns_per_step": turns_c / steps, "windows": [round(v, 1) for v in vals_c],
                                "note": "same call, but every step's results are on the host (stream synchronised) "
                                        "before the next step is issued"},
                    gpu_launches=int(launches), clocks=clk,
                    roofline={"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                              "traffic": traffic, "kernel": lib.agv_last_tick_kernel().decode(),
                              "bytes_per_unit": unit_bytes, "peak_source": peak_src, "issue": issue,
                              "note": ("HBM is not the binding roof of the limited-visual rollout (SURVEY 8(d): the 1e9 "
                                       "target is 0.6 % of the HBM roof); what binds is the latency of the sequential "
                                       "per-scene turn chains on the issue / ALU pipes: see roofline.issue and DESIGN.md 3")
                              if args.obs == "clip" else "full-map obs: 3*H*W bf16 + 38 B per AGV-step, HBM-write-bound"})
    sc.close()
    del sc, out
    torch.cuda.empty_cache()

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        nthr = host_threads()
        tps = 50 if args.workload != "c2" else 2000
        arm = CpuArm(args.workload, nthr)  # bounded sample: ~cpu-seconds of CPU work
        arm.batch(2 * N + 8)
        t1 = arm.batch(tps)[1]
        calls = int(max(1, min(400, args.cpu_seconds / max(t1, 1e-3))))
        res = [arm.batch(tps) for _ in range(calls)]
        c_turns, c_secs = sum(r[0] for r in res), sum(r[1] for r in res)
        line["cpu_baseline"] = {"value": c_turns / c_secs, "unit": UNIT, "cores": nthr, "kind": "port",
                                "sample": "%d threads x 1 scene x %d ticks x %d calls (%.1f s, %d AGV turns), oracle C "
                                          "port of the reference path; the unmodified Python reference is timed by "
                                          "--impl reference" % (nthr, tps, calls, c_secs, c_turns)}
        del arm

    if not args.no_sub and args.workload == "c3" and args.obs == "clip" and not args.shaped:
        # BASELINE configs[3] / configs[4] in the driver's view (every rank takes part: the training
        # sub-record all-reduces the gradient over NCCL)
        sub = {}
        sub["train"] = sub_train(pkg, lib, du, dev, local_rank, world, 10)
        sub["train_matched"] = sub_train(pkg, lib, du, dev, local_rank, world, 4, matched=True)
        torch.cuda.empty_cache()
        sub["expert"] = sub_expert(pkg, lib, du, dev, local_rank, world, 2)
        if line is not None:
            line.update(sub)
    return finish(line)


if __name__ == "__main__":
    sys.exit(main())
