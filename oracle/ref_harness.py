"""Headless import of the unmodified Python reference.

TEST INFRASTRUCTURE ONLY.  Sources: /root/reference in the build container, or its copy under the
git-ignored baseline/_ref/ (made by __graft_entry__.build(), travels to the GPU box).  Used by
tests/golden/make_golden*.py (which write the committed fixtures), by the tests that drive the
reference's own controllers against this package's Scene, and by `bench.py --impl reference`.
Recipe: SURVEY.md App. B.1.
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys
import types

def _default_root():
    """the live checkout in the build container, else the copy vendored by __graft_entry__.build()"""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in ("/root/reference", os.path.join(here, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "src", "multiAGVscene")):
            return cand
    return "/root/reference"


REF_ROOT = os.environ.get("AGV_REFERENCE_ROOT") or _default_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "multiAGVscene"))


_mods = None


def load():
    """Returns a namespace with the reference classes.  pygame / matplotlib are
    replaced by empty stub modules (Scene.py:2 imports pygame at module top)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    for m in ("pygame", "matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for p in (REF_ROOT, os.path.join(REF_ROOT, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):
        from multiAGVscene.Layout import Layout
        from multiAGVscene.Explorer import Explorer
        from multiAGVscene.Scene import Scene
        import utils.astar as astar
        import utils.settings as settings
        from src.algorithm.Manager.StateManager import StateManager
        import importlib
        PER = importlib.import_module("src.algorithm.MADQN_structure.PER")
    ns = types.SimpleNamespace(Layout=Layout, Explorer=Explorer, Scene=Scene, astar=astar, settings=settings,
                               StateManager=StateManager, PER=PER)
    _mods = ns
    return ns


class TickCap(Exception):
    pass


def make_scene(n_agv, rect=None, layout_list=None, task_list=None, seed=None, flags=None):
    """Layout -> Explorers -> Scene like src/main.py:25-41.  `flags` sets
    SuperParas class attributes (read at Explorer.__init__ / in check_action)."""
    R = load()
    flags = flags or {}
    for k in ("Explorer_Always_Loaded", "Explorer_Always_Empty", "Sparse_Reward"):
        setattr(R.settings.SuperParas, k, bool(flags.get(k, False)))
    if seed is not None:
        random.seed(seed)
    if layout_list is not None:
        layout = R.Layout(layout_list=[list(map(int, r)) for r in layout_list], task_list=task_list)
    else:
        w, h, nx, ny, ps = rect
        layout = R.Layout(w, h, nx, ny, ps, task_list=task_list)
    group = [R.Explorer(layout, veh_name="veh%d" % (i + 1), icon_name="veh%d" % (i + 1)) for i in range(n_agv)]
    scene = R.Scene(layout, group)
    return scene


def snapshot(scene):
    """[[x, y, tx, ty, loaded] for created AGVs] from Scene.create_info"""
    info = scene.create_info()
    return [[v[1][0], v[1][1], v[2][0], v[2][1], int(bool(v[3]))] for v in info[1:]]


def run_astar(scene, max_ticks):
    """Scene.run_game('A_star', render=False) with a tick cap (A_star mode never
    terminates on gridlock).  check_new_veh is called exactly once
    per tick, so it carries the cap.  Returns (names, actions, infos, ticks, finished)."""
    orig = scene.check_new_veh
    state = {"ticks": 0}

    def capped():
        state["ticks"] += 1
        r = orig()
        if state["ticks"] >= max_ticks:
            raise TickCap()
        return r

    scene.check_new_veh = capped
    rec = {"names": [], "actions": [], "infos": []}
    # record through the explorers so a capped run still yields its log
    for ex in scene.explorer_group:
        def wrap(ex=ex, orig_exec=ex.execute_action):
            def f(input_action, all_info=[], explorer_group=None):
                rec["names"].append(ex.explorer_name)
                rec["actions"].append(ex.dir.action_str_value(input_action))
                rec["infos"].append(snapshot(scene))
                return orig_exec(input_action, all_info, explorer_group)
            return f
        ex.execute_action = wrap()
    finished = True
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            scene.run_game(control_pattern="A_star", render=False)
        except TickCap:
            finished = False
    return rec["names"], rec["actions"], rec["infos"], scene.running_time, finished


class RecordingController:
    """smart_controller for Scene.run_game('intelligent'): mixes AG-DQN guidance
    (StateManager valid matrix + FindPathAstar, MADQN.py:192-200) with random
    actions and records everything the parity tests need."""

    def __init__(self, scene, rng, p_guided=0.85, n_actions=4, record_obs=False, P=3):
        R = load()
        self.R = R
        self.scene = scene
        self.rng = rng
        self.p_guided = p_guided
        self.n_actions = n_actions
        self.sm = R.StateManager()
        self.record_obs = record_obs
        self.P = P
        self.turns = []

    def choose_action(self, all_info, name):
        R = self.R
        obs_c, cp, tp, valid = self.sm.create_state(all_info, name, obs_clip=True, padding_size=self.P)
        f = R.astar.FindPathAstar(valid, (cp[0] - 1, cp[1] - 1), (tp[0] - 1, tp[1] - 1))
        found, _pl, _pm, al = f.run_astar_method()
        guided = 4 if (found is False or len(al) == 0) else {"UP": 0, "RIGHT": 1, "DOWN": 2, "LEFT": 3}[al[0]]
        if self.rng.random() < self.p_guided and guided < self.n_actions:
            a = guided
        else:
            a = self.rng.randrange(self.n_actions)
        t = {"i": int(name[3:]) - 1, "a": a, "guided": guided, "glen": len(al), "before": snapshot_info(all_info)}
        if self.record_obs:
            obs_f, _, _, _ = self.sm.create_state(all_info, name, obs_clip=False)
            t["obs_clip"] = [[[int(v) for v in row] for row in ch] for ch in obs_c]
            t["obs_full"] = [[[int(v) for v in row] for row in ch] for ch in obs_f]
        self.turns.append(t)
        return a

    def store_info(self, all_info, reward, is_end, name):
        t = self.turns[-1]
        t["r"] = reward
        t["end"] = int(bool(is_end))
        t["after"] = snapshot_info(all_info)


def snapshot_info(all_info):
    return [[v[1][0], v[1][1], v[2][0], v[2][1], int(bool(v[3]))] for v in all_info[1:]]


def run_intelligent(scene, controller):
    with contextlib.redirect_stdout(io.StringIO()):
        rt = scene.run_game(control_pattern="intelligent", smart_controller=controller, render=False)
    return rt
