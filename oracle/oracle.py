"""ctypes front-end of the CPU oracle (oracle/agv_oracle.c) + numpy restatements.

TEST INFRASTRUCTURE ONLY -- see the header of agv_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.

Citations are relative to the reference checkout.
"""
from __future__ import annotations

import ctypes as C
import os
import random
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libagv_oracle.so")

UP, RIGHT, DOWN, LEFT, STOP = 0, 1, 2, 3, 4
PROTO_ASTAR, PROTO_INTELLIGENT = 0, 1
POLICY_GIVEN, POLICY_ASTAR, POLICY_GUIDED = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "agv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i, l = C.c_void_p, C.c_int, C.c_long
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i, i, i, i, vp, vp, i, i, i, l]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_tasks.argtypes = [vp, vp]
        L.orc_reset.argtypes = [vp]
        L.orc_valid_a.argtypes = [vp, i, vp]
        L.orc_valid_b.argtypes = [vp, i, vp]
        L.orc_bfs_literal.restype = i
        L.orc_bfs_literal.argtypes = [vp, i, i, i, i, i, i, vp, vp, vp, i]
        L.orc_bfs_field.restype = i
        L.orc_bfs_field.argtypes = [vp, i, i, i, i, i, i, vp, vp, vp, vp]
        L.orc_action_astar.restype = i
        L.orc_action_astar.argtypes = [vp, i, vp]
        L.orc_action_guided.restype = i
        L.orc_action_guided.argtypes = [vp, i, vp]
        L.orc_turn_state.restype = i
        L.orc_turn_state.argtypes = [vp, i]
        L.orc_execute.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp]
        L.orc_spawn.restype = i
        L.orc_spawn.argtypes = [vp]
        L.orc_begin_tick.argtypes = [vp]
        L.orc_tick.restype = i
        L.orc_tick.argtypes = [vp, i, i, vp, vp, vp, vp, vp]
        L.orc_tick_obs.restype = i
        L.orc_tick_obs.argtypes = [vp, i, i, vp, vp, vp, vp, vp, vp, i, i, vp, vp]
        L.orc_encode.argtypes = [vp, i, i, i, vp]
        L.orc_snapshot.restype = i
        L.orc_snapshot.argtypes = [vp, vp]
        L.orc_get_state.argtypes = [vp, vp, vp]
        L.orc_run.restype = l
        L.orc_run.argtypes = [vp, i, i, l, i, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------
# Layout geometry (src/multiAGVscene/Layout.py:15-132), restated with numpy.
# ----------------------------------------------------------------------------
def rect_layout(w, h, nx, ny, ps):
    """Rect layout grid codes.  W = nx*(w+1)+1, H = ny*(h+1)+1+3 (Layout.py:29-30);
    storage at (ix*(w+1)+dx+2, iy*(h+1)+dy+2) (92-102); picking i at
    x=int(W*(i+1)/(ps+1)+1), y=H-1 (83-90); picking overrides storage (125-128)."""
    W = nx * (w + 1) + 1
    H = ny * (h + 1) + 1 + 3
    g = np.zeros((H, W), dtype=np.uint8)
    for iy in range(ny):
        for dy in range(h):
            y = iy * (h + 1) + dy + 2
            for ix in range(nx):
                for dx in range(w):
                    x = ix * (w + 1) + dx + 2
                    g[y - 1, x - 1] = 1
    for i in range(ps):
        x = int(W * (i + 1) / (ps + 1) + 1)
        y = H - 1
        g[y - 1, x - 1] = 2
    return g


def stations(grid):
    """storage / picking station lists in the reference's order: rect layouts and
    layout_list layouts both enumerate row-major (Layout.py:66-81,92-102); rect
    picking stations are ordered by i which is also row-major on row H-1."""
    ss = [(int(x) + 1, int(y) + 1) for y, x in zip(*np.nonzero(grid == 1))]
    ps = [(int(x) + 1, int(y) + 1) for y, x in zip(*np.nonzero(grid == 2))]
    return ss, ps


def make_tasks(grid, rng: random.Random):
    """Layout.__create_task (Layout.py:104-116): pop a random storage station,
    pair with a random picking station, until none are left.  Same call order on
    the RNG (randint for ss, then randint for ps)."""
    ss, ps = stations(grid)
    tasks = []
    while True:
        i = rng.randint(0, len(ss) - 1)
        j = rng.randint(0, len(ps) - 1)
        s = ss.pop(i)
        tasks.append(s + ps[j])
        if not ss:
            break
    return tasks


def max_training_steps(W, H):
    """Scene.py:24"""
    return int((W / 2 + H) * 3 * ((W * H) / 2.5))


README_9x9 = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0],
                       [0, 1, 1, 1, 1, 1, 1, 0, 0],
                       [0, 1, 1, 1, 1, 1, 1, 0, 0],
                       [0, 0, 0, 0, 0, 0, 1, 1, 0],
                       [0, 0, 0, 2, 0, 0, 1, 1, 0],
                       [0, 0, 0, 0, 0, 0, 1, 1, 0],
                       [0, 1, 1, 1, 1, 1, 1, 0, 0],
                       [0, 1, 1, 1, 1, 1, 1, 0, 0],
                       [0, 0, 0, 0, 0, 0, 0, 0, 0]], dtype=np.uint8)  # README.md:73-81


class Env:
    """One scene (Layout + Explorer group + Scene) of the oracle."""

    def __init__(self, grid, tasks, n_agv, always_loaded=False, always_empty=False, shaped=False,
                 max_steps=None):
        self.grid = np.ascontiguousarray(grid, dtype=np.uint8)
        self.H, self.W = self.grid.shape
        self.N = int(n_agv)
        self.tasks = np.ascontiguousarray(np.asarray(tasks, dtype=np.int32).reshape(-1, 4))
        self.T = self.tasks.shape[0]
        if max_steps is None:
            max_steps = max_training_steps(self.W, self.H)
        self.max_steps = int(max_steps)
        self._h = lib().orc_create(self.H, self.W, self.N, self.T, _p(self.grid), _p(self.tasks),
                                   int(always_loaded), int(always_empty), int(shaped), self.max_steps)
        self.reset()

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self):
        lib().orc_reset(self._h)

    def set_tasks(self, tasks):
        t = np.ascontiguousarray(np.asarray(tasks, dtype=np.int32).reshape(-1, 4))
        assert t.shape[0] == self.T
        self.tasks = t
        lib().orc_set_tasks(self._h, _p(t))

    def valid_a(self, i):
        out = np.zeros((self.H, self.W), dtype=np.uint8)
        lib().orc_valid_a(self._h, i, _p(out))
        return out

    def valid_b(self, i):
        out = np.zeros((self.H, self.W), dtype=np.uint8)
        lib().orc_valid_b(self._h, i, _p(out))
        return out

    def action_astar(self, i):
        ln = C.c_int(0)
        a = lib().orc_action_astar(self._h, i, C.byref(ln))
        return a, ln.value

    def action_guided(self, i):
        ln = C.c_int(0)
        a = lib().orc_action_guided(self._h, i, C.byref(ln))
        return a, ln.value

    def turn_state(self, i):
        return lib().orc_turn_state(self._h, i)

    def begin_tick(self):
        lib().orc_begin_tick(self._h)

    def execute(self, i, action, proto):
        ri, ie, su, pl = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        rf = C.c_double(0)
        lib().orc_execute(self._h, i, int(action), proto, C.byref(ri), C.byref(rf), C.byref(ie), C.byref(su),
                          C.byref(pl))
        return rf.value, bool(ie.value), ri.value, bool(su.value), pl.value

    def spawn(self):
        return lib().orc_spawn(self._h)

    def tick(self, proto, policy, given=None):
        act = np.zeros(self.N, dtype=np.int8)
        rew = np.zeros(self.N, dtype=np.float64)
        end = np.zeros(self.N, dtype=np.uint8)
        plen = np.zeros(self.N, dtype=np.int32)
        g = None
        if given is not None:
            g = np.ascontiguousarray(given, dtype=np.int8)
        n = lib().orc_tick(self._h, proto, policy, _p(g) if g is not None else None, _p(act), _p(rew), _p(end),
                           _p(plen))
        return n, act, rew, end, plen

    def tick_obs(self, proto, policy, given=None, obs_kind=0, P=3, want_next=False):
        """(turns, act, rew, end, plen, slen, obs, next_obs); obs are float32 [N,3,..]"""
        act = np.zeros(self.N, dtype=np.int8)
        rew = np.zeros(self.N, dtype=np.float64)
        end = np.zeros(self.N, dtype=np.uint8)
        plen = np.zeros(self.N, dtype=np.int32)
        slen = np.zeros(self.N, dtype=np.int32)
        K = 2 * P + 1
        shp = None if obs_kind == 0 else ((self.N, 3, K, K) if obs_kind == 1 else (self.N, 3, self.H, self.W))
        obs = np.zeros(shp, dtype=np.float32) if shp else None
        nxt = np.zeros(shp, dtype=np.float32) if (shp and want_next) else None
        g = np.ascontiguousarray(given, dtype=np.int8) if given is not None else None
        n = lib().orc_tick_obs(self._h, proto, policy, _p(g) if g is not None else None, _p(act), _p(rew), _p(end),
                               _p(plen), _p(slen), obs_kind, P, _p(obs) if obs is not None else None,
                               _p(nxt) if nxt is not None else None)
        return n, act, rew, end, plen, slen, obs, nxt

    def encode(self, i, clip=False, P=3):
        K = 2 * P + 1
        out = np.zeros((3, K, K) if clip else (3, self.H, self.W), dtype=np.float32)
        lib().orc_encode(self._h, i, int(clip), P, _p(out))
        return out

    def snapshot(self):
        out = np.zeros((self.N, 5), dtype=np.int32)
        n = lib().orc_snapshot(self._h, _p(out))
        return out[:n].copy()

    def state(self):
        hdr = np.zeros(6, dtype=np.int64)
        agv = np.zeros((self.N, 8), dtype=np.int32)
        lib().orc_get_state(self._h, _p(hdr), _p(agv))
        return hdr, agv

    def run(self, proto, policy, nticks, encode_P=None):
        cs = C.c_ulong(0)
        obs = None
        if encode_P is not None:
            K = 2 * encode_P + 1
            obs = np.zeros(3 * K * K, dtype=np.float32)
        n = lib().orc_run(self._h, proto, policy, int(nticks), int(encode_P or 0),
                          _p(obs) if obs is not None else None, C.byref(cs))
        return n, cs.value


def bfs_field(valid, start, target):
    """(action, found, len) with 0-based (x, y) start/target, distance-field rule."""
    v = np.ascontiguousarray(valid, dtype=np.uint8)
    H, W = v.shape
    found, ln = C.c_int(0), C.c_int(0)
    dist = np.zeros(H * W, dtype=np.int32)
    queue = np.zeros(H * W, dtype=np.int32)
    a = lib().orc_bfs_field(_p(v), H, W, start[0], start[1], target[0], target[1], C.byref(found), C.byref(ln),
                            _p(dist), _p(queue))
    return a, bool(found.value), ln.value


def bfs_literal(valid, start, target):
    """(action, found, action_list) following astar.py:73-153 step by step."""
    v = np.ascontiguousarray(valid, dtype=np.uint8)
    H, W = v.shape
    found, ln = C.c_int(0), C.c_int(0)
    acts = np.zeros(H * W + 4, dtype=np.int8)
    a = lib().orc_bfs_literal(_p(v), H, W, start[0], start[1], target[0], target[1], C.byref(found), C.byref(ln),
                              _p(acts), acts.size)
    return a, bool(found.value), acts[:ln.value].tolist()


# ----------------------------------------------------------------------------
# Prioritised replay (src/algorithm/MADQN_structure/PER.py:6-112), numpy float64.
# ----------------------------------------------------------------------------
class SumTree:
    def __init__(self, capacity):
        self.capacity = capacity
        self.tree = np.zeros(2 * capacity - 1)
        self.write = 0
        self.n_entries = 0

    def add(self, p):
        """PER.py:41-52; returns the data slot written"""
        slot = self.write
        self.update(slot + self.capacity - 1, p)
        self.write += 1
        if self.write >= self.capacity:
            self.write = 0
        if self.n_entries < self.capacity:
            self.n_entries += 1
        return slot

    def update(self, idx, p):
        """PER.py:55-59 with _propagate 16-22 unrolled"""
        change = p - self.tree[idx]
        self.tree[idx] = p
        while idx != 0:
            idx = (idx - 1) // 2
            self.tree[idx] += change

    def retrieve(self, s):
        """PER.py:25-35"""
        idx = 0
        while True:
            left = 2 * idx + 1
            if left >= len(self.tree):
                return idx
            if s <= self.tree[left]:
                idx = left
            else:
                s = s - self.tree[left]
                idx = left + 1

    def total(self):
        return self.tree[0]


def per_priority(error, e=0.01, a=0.6):
    """Memory._get_priority, PER.py:79-80"""
    return (np.abs(error) + e) ** a


def per_sample(tree: SumTree, n, uniforms01):
    """Memory.sample (PER.py:86-108) with the random.uniform(a,b) draws injected as
    u in [0,1): random.uniform(a, b) = a + (b-a)*random() (CPython)."""
    segment = tree.total() / n
    slots, tree_idx, prios = [], [], []
    for i in range(n):
        a = segment * i
        b = segment * (i + 1)
        s = a + (b - a) * uniforms01[i]
        idx = tree.retrieve(s)
        tree_idx.append(idx)
        slots.append(idx - tree.capacity + 1)
        prios.append(tree.tree[idx])
    return np.array(slots), np.array(tree_idx), np.array(prios)
