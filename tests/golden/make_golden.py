"""Generates tests/golden/*.json by running the LIVE, unmodified Python reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The fixtures are committed; tests never need the reference checkout again.
Everything here is test infrastructure (see oracle/agv_oracle.c header).
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle.oracle import README_9x9  # noqa: E402

ACT = {"UP": 0, "RIGHT": 1, "DOWN": 2, "LEFT": 3, "STOP": 4}


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote %s (%d bytes)" % (name, os.path.getsize(path)))


def gen_bfs():
    R = rh.load()
    cases = []
    # the reference's only own known-answer test: astar.py:169-186
    demo = [[1] * 7 for _ in range(12)]
    demo[2][0] = 0
    demo[4][4] = demo[4][5] = demo[4][6] = 0
    demo[5][4] = 0
    demo[6][4] = demo[6][5] = 0
    demo[10][1] = demo[10][4] = 0
    f = R.astar.FindPathAstar([list(map(float, r)) for r in demo], (0, 3), (5, 5))
    found, pl, _pm, al = f.run_astar_method()
    cases.append({"name": "astar_demo", "valid": demo, "start": [0, 3], "target": [5, 5], "found": bool(found),
                  "actions": [ACT[a] for a in al], "path": [list(p) for p in pl]})
    rng = random.Random(1234)
    for n in range(400):
        H, W = rng.randint(2, 14), rng.randint(2, 14)
        dens = rng.choice([0.0, 0.1, 0.2, 0.3, 0.4])
        valid = [[0 if rng.random() < dens else 1 for _ in range(W)] for _ in range(H)]
        s = (rng.randrange(W), rng.randrange(H))
        t = (rng.randrange(W), rng.randrange(H)) if rng.random() > 0.03 else s
        f = R.astar.FindPathAstar([list(r) for r in valid], s, t)
        found, _pl, _pm, al = f.run_astar_method()
        cases.append({"valid": valid, "start": list(s), "target": list(t), "found": bool(found),
                      "actions": [ACT[a] for a in al]})
    dump("bfs.json", cases)


def fingerprint(names, actions, infos):
    """SURVEY App. B.4 fingerprint"""
    pos = [[[v[0], v[1]] for v in info] for info in infos]
    return hashlib.sha256(json.dumps([names, actions, pos]).encode()).hexdigest()[:16]


def gen_astar_traj():
    out = []
    cfgs = [("C1", dict(rect=(4, 2, 2, 2, 2)), 4, s, 2000) for s in (0, 1, 2)]
    cfgs += [("C2", dict(layout_list=README_9x9.tolist()), 4, 0, 2000),
             ("C2_gridlock", dict(layout_list=README_9x9.tolist()), 4, 1, 120),
             ("rect_3_2_2_2_2_n1", dict(rect=(3, 2, 2, 2, 2)), 1, 7, 2000),
             ("rect_2_1_2_2_1_n6", dict(rect=(2, 1, 2, 2, 1)), 6, 3, 400)]
    for name, kw, n, seed, cap in cfgs:
        scene = rh.make_scene(n, seed=seed, **kw)
        grid = [list(r) for r in scene.layout.layout_original]
        tasks = [list(t) for t in scene.layout.task_list]
        names, actions, infos, ticks, finished = rh.run_astar(scene, cap)
        out.append({"name": name, "seed": seed, "n_agv": n, "grid": grid, "tasks": tasks,
                    "veh": [int(v[3:]) - 1 for v in names], "actions": actions,
                    "pos": [[[v[0], v[1]] for v in info] for info in infos],
                    "ticks": ticks, "finished": finished, "sha": fingerprint(names, actions, infos),
                    "final": rh.snapshot(scene)})
        print(name, seed, "ticks", ticks, "steps", len(actions), "stops", actions.count(4), out[-1]["sha"])
    dump("astar_traj.json", out)


def gen_intelligent():
    out = []
    rng = random.Random(99)
    layouts = [dict(rect=(1, 1, 1, 1, 1)), dict(rect=(2, 1, 2, 1, 1)), dict(rect=(2, 2, 2, 2, 2)),
               dict(rect=(4, 2, 2, 2, 2)), dict(layout_list=README_9x9.tolist()), dict(rect=(1, 2, 2, 1, 2))]
    variants = [({}, 0.9, 4), ({}, 0.97, 4), ({}, 1.0, 4), ({"Sparse_Reward": True}, 0.9, 4),
                ({"Explorer_Always_Loaded": True}, 0.9, 4), ({"Explorer_Always_Empty": True}, 0.95, 5),
                ({}, 0.8, 5), ({"Sparse_Reward": True}, 1.0, 4)]
    ep = 0
    for li, kw in enumerate(layouts):
        for vi, (flags, pg, na) in enumerate(variants):
            for n in (1, 2, 4):
                if (li + vi + n) % 3 == 0 and not (li == 0):
                    continue  # thin out
                seed = 1000 + ep
                scene = rh.make_scene(n, seed=seed, flags=flags, **kw)
                ctl = rh.RecordingController(scene, random.Random(seed), p_guided=pg, n_actions=na,
                                             record_obs=(ep % 17 == 0 and li < 5))
                rt = rh.run_intelligent(scene, ctl)
                turns = ctl.turns
                if len(turns) > 700:
                    ep += 1
                    continue  # keep the fixture small; long episodes are covered by the live tests
                out.append({"ep": ep, "n_agv": n, "flags": flags,
                            "grid": [list(r) for r in scene.layout.layout_original],
                            "tasks": [list(t) for t in scene.layout.task_list],
                            "max_steps": scene.max_training_steps,
                            "running_time": rt if isinstance(rt, int) else -1,
                            "turns": turns})
                ep += 1
    # a max-steps ending: a single empty AGV that mostly STOPs on the smallest layout
    scene = rh.make_scene(1, seed=5, rect=(1, 1, 1, 1, 1), flags={})

    class Stopper(rh.RecordingController):
        def choose_action(self, all_info, name):
            a = super().choose_action(all_info, name)
            t = self.turns[-1]
            if len(self.turns) > 3:
                t["a"] = 4
                return 4
            return a
    ctl = Stopper(scene, random.Random(5), p_guided=1.0, n_actions=4)
    rt = rh.run_intelligent(scene, ctl)
    out.append({"ep": ep, "n_agv": 1, "flags": {}, "grid": [list(r) for r in scene.layout.layout_original],
                "tasks": [list(t) for t in scene.layout.task_list], "max_steps": scene.max_training_steps,
                "running_time": rt, "turns": ctl.turns})
    ends = {}
    for e in out:
        t = e["turns"][-1]
        ends[(t["r"], t["end"])] = ends.get((t["r"], t["end"]), 0) + 1
    print("intelligent episodes", len(out), "turns", sum(len(e["turns"]) for e in out), "endings", ends)
    dump("intelligent.json", out)


def gen_per():
    """PER.py SumTree/Memory with injected uniforms (random.uniform patched)."""
    R = rh.load()
    PER = R.PER
    out = []
    rng = random.Random(7)
    for cap, n_add, nb in ((8, 5, 4), (8, 13, 4), (64, 64, 32), (100, 250, 32), (50, 37, 16)):
        mem = PER.Memory(cap)
        errs = [rng.random() * rng.choice([0.01, 1.0, 5.0]) for _ in range(n_add)]
        for k, e in enumerate(errs):
            mem.add(e, k)
        us = [rng.random() for _ in range(nb)]
        it = iter(us)
        orig = PER.random.uniform
        PER.random.uniform = lambda a, b: a + (b - a) * next(it)
        try:
            batch, idxs, isw = mem.sample(nb)
        finally:
            PER.random.uniform = orig
        out.append({"capacity": cap, "errors": errs, "uniforms": us, "batch": [int(b) for b in batch],
                    "tree_idx": [int(i) for i in idxs], "total": float(mem.tree.total()),
                    "n_entries": int(mem.tree.n_entries)})
    dump("per.json", out)


def gen_expert_csv():
    """ExpertManager.create_data_by_self of the live reference -> tests/golden/expert_ref_*.csv"""
    import contextlib, io, signal, shutil, importlib
    rh.load()
    EM = importlib.import_module("src.algorithm.Manager.ExpertManager")
    meta = []
    for tag, rect, n, seed, times in (("a", (2, 1, 2, 1, 1), 1, 11, 3), ("b", (2, 2, 2, 1, 2), 2, 4, 2)):
        random.seed(seed)
        scene = rh.make_scene(n, rect=rect, flags={})   # make_scene(seed=None) keeps the stream
        tmp = "/tmp/_expert_ref_%s.csv" % tag
        if os.path.exists(tmp):
            os.remove(tmp)
        ex = EM.Expert(scene, *rect, n, filename=tmp)
        orig = scene.run_game
        scene.run_game = lambda control_pattern="A_star", **kw: orig(control_pattern=control_pattern, render=False)
        signal.alarm(60)
        with contextlib.redirect_stdout(io.StringIO()):
            ex.create_data_by_self(times=times)
        signal.alarm(0)
        dst = os.path.join(HERE, "expert_ref_%s.csv" % tag)
        shutil.copy(tmp, dst)
        meta.append({"tag": tag, "rect": list(rect), "n_agv": n, "seed": seed, "times": times,
                     "bytes": os.path.getsize(dst)})
        print("wrote", dst, os.path.getsize(dst))
    dump("expert_ref_meta.json", meta)


if __name__ == "__main__":
    gen_bfs()
    gen_astar_traj()
    gen_intelligent()
    gen_per()
    gen_expert_csv()
