"""BatchedScene -- E independent scenes (Layout + Explorers + Scene) resident on one GPU.

Thin host mirror of the C ABI (include/agv_b200.h): torch only provides device
memory and the stream.  Reference semantics: src/multiAGVscene/Scene.py:98-175,
Explorer.py:131-208, src/algorithm/Manager/StateManager.py:14-41.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def max_training_steps(W: int, H: int) -> int:
    """Scene.max_training_steps, Scene.py:24"""
    return int((W / 2 + H) * 3 * ((W * H) / 2.5))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class BatchedScene:
    def __init__(self, grid, n_env: int, n_agv: int, tasks=None, P: int = 3, always_loaded: bool = False,
                 always_empty: bool = False, shaped_reward: bool = False, auto_reset: bool = False,
                 max_steps: int | None = None, device: int | None = None):
        if not torch.cuda.is_available():
            raise L.AgvError("BatchedScene needs a CUDA device; there is no CPU fallback")
        self.lib = L.lib()
        self.grid = np.ascontiguousarray(grid, dtype=np.uint8)
        self.H, self.W = self.grid.shape
        self.E, self.N, self.P = int(n_env), int(n_agv), int(P)
        self.K = 2 * self.P + 1
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.max_steps = max_training_steps(self.W, self.H) if max_steps is None else int(max_steps)
        t = None
        if tasks is not None:
            t = np.asarray(tasks)
            if t.ndim == 2:
                t = np.broadcast_to(t[None], (self.E,) + t.shape)
            assert t.shape[0] == self.E and t.shape[2] == 4
        self.T = 0 if t is None else int(t.shape[1])
        self.cfg = L.AgvConfig(H=self.H, W=self.W, E=self.E, N=self.N, T=self.T, P=self.P,
                               always_loaded=int(always_loaded), always_empty=int(always_empty),
                               shaped_reward=int(shaped_reward), auto_reset=int(auto_reset),
                               max_steps=self.max_steps, device=self.device_index, reserved=0)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.empty(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        L.check(self.lib.agv_create(C.byref(self.cfg), self.grid.ctypes.data_as(C.c_void_p), C.byref(h)), "agv_create")
        self._h = h
        self._buf = {}
        if t is not None:
            self.set_tasks(t)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.lib.agv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _get(self, name, shape, dtype):
        b = self._buf.get(name)
        if b is None or tuple(b.shape) != tuple(shape) or b.dtype != dtype:
            # zero-initialised once: observation records of AGVs without a turn are never written
            b = torch.zeros(shape, dtype=dtype, device=self.device)
            self._buf[name] = b
        return b

    def obs_shape(self, obs_kind):
        return (3, self.K, self.K) if obs_kind == L.OBS_CLIP else (3, self.H, self.W)

    # ------------------------------------------------------------------ setup
    def set_tasks(self, tasks, env_begin: int = 0):
        """Layout.task_list per scene: [n, T, 4] (sx, sy, px, py), 1-based."""
        t = np.ascontiguousarray(np.asarray(tasks), dtype=np.uint8)
        assert t.ndim == 3 and t.shape[1] == self.T and t.shape[2] == 4
        L.check(self.lib.agv_set_tasks(self._h, t.ctypes.data_as(C.c_void_p), env_begin, t.shape[0], self._stream()),
                "agv_set_tasks")

    def reset(self, mask=None):
        """Scene.init() + create_explorer of AGV 0 for the masked scenes (all if None)."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        L.check(self.lib.agv_reset(self._h, _ptr(m), self._stream()), "agv_reset")

    # ------------------------------------------------------------------ fused tick
    def tick(self, proto, policy, obs_kind=L.OBS_NONE, actions=None, want_next=False, want_lens=True,
             out=None):
        """One Scene.run_mode loop pass for every scene.  Returns a dict of device
        tensors indexed [scene, agv] (buffers are reused between calls)."""
        E, N = self.E, self.N
        o = {} if out is None else out
        o.setdefault("act", self._get("act", (E, N), torch.int8))
        o.setdefault("reward", self._get("reward", (E, N), torch.float32))
        o.setdefault("done", self._get("done", (E, N), torch.uint8))
        if want_lens:
            o.setdefault("plen", self._get("plen", (E, N), torch.int16))
            o.setdefault("slen", self._get("slen", (E, N), torch.int16))
        if obs_kind != L.OBS_NONE:
            shp = (E, N) + self.obs_shape(obs_kind)
            o.setdefault("obs", self._get("obs%d" % obs_kind, shp, torch.bfloat16))
            if want_next:
                o.setdefault("next_obs", self._get("next%d" % obs_kind, shp, torch.bfloat16))
        a = None
        if policy == L.POLICY_GIVEN:
            a = torch.as_tensor(actions, device=self.device).to(torch.int8).contiguous()
            assert tuple(a.shape) == (E, N)
        L.check(self.lib.agv_tick(self._h, proto, policy, obs_kind, _ptr(a), _ptr(o["act"]), _ptr(o["reward"]),
                                  _ptr(o["done"]), _ptr(o.get("plen")), _ptr(o.get("slen")), _ptr(o.get("obs")),
                                  _ptr(o.get("next_obs")), self._stream()), "agv_tick")
        return o

    def tick_host(self, proto, policy, obs_kind=L.OBS_NONE, actions_host=None, want_next=False, join=True):
        """Same through HOST (pinned) buffers for actions in and act/reward/done out.  The copy-back
        runs on the library's own stream and overlaps the next tick; with `join` (default) the
        caller's stream waits for it right away, otherwise call host_join() before reading."""
        E, N = self.E, self.N
        hb = self._buf.setdefault("host", {})
        if not hb:
            hb["act"] = torch.empty((E, N), dtype=torch.int8).pin_memory()
            hb["reward"] = torch.empty((E, N), dtype=torch.float32).pin_memory()
            hb["done"] = torch.empty((E, N), dtype=torch.uint8).pin_memory()
        obs = nxt = None
        if obs_kind != L.OBS_NONE:
            shp = (E, N) + self.obs_shape(obs_kind)
            obs = self._get("obs%d" % obs_kind, shp, torch.bfloat16)
            if want_next:
                nxt = self._get("next%d" % obs_kind, shp, torch.bfloat16)
        ah = None
        if policy == L.POLICY_GIVEN:
            ah = actions_host
            assert ah.dtype == torch.int8 and tuple(ah.shape) == (E, N) and ah.is_contiguous()
        L.check(self.lib.agv_tick_host(self._h, proto, policy, obs_kind, _ptr(ah), _ptr(hb["act"]),
                                       _ptr(hb["reward"]), _ptr(hb["done"]), _ptr(obs), _ptr(nxt), self._stream()),
                "agv_tick_host")
        if join:
            self.host_join()
        return {"act": hb["act"], "reward": hb["reward"], "done": hb["done"], "obs": obs, "next_obs": nxt}

    def host_join(self):
        """The current stream waits for the outstanding copy-backs of tick_host."""
        L.check(self.lib.agv_tick_host_join(self._h, self._stream()), "agv_tick_host_join")

    # ------------------------------------------------------------------ multi-tick rollout
    def rollout(self, n_ticks: int, proto, policy, obs_kind=L.OBS_NONE, actions=None, want_next=False,
                want_lens=True, out=None):
        """n_ticks Scene.run_mode loop passes for every scene in one launch (agv_rollout): the same
        results as n_ticks calls of tick(), in rings indexed [tick, scene, agv].  `actions`
        (POLICY_GIVEN) is an open-loop [n_ticks, E, N] int8 tensor (negative = A* guidance)."""
        K, E, N = int(n_ticks), self.E, self.N
        o = {} if out is None else out
        o.setdefault("act", self._get("r_act%d" % K, (K, E, N), torch.int8))
        o.setdefault("reward", self._get("r_reward%d" % K, (K, E, N), torch.float32))
        o.setdefault("done", self._get("r_done%d" % K, (K, E, N), torch.uint8))
        if want_lens:
            o.setdefault("plen", self._get("r_plen%d" % K, (K, E, N), torch.int16))
            o.setdefault("slen", self._get("r_slen%d" % K, (K, E, N), torch.int16))
        if obs_kind != L.OBS_NONE:
            shp = (K, E, N) + self.obs_shape(obs_kind)
            o.setdefault("obs", self._get("r_obs%d_%d" % (obs_kind, K), shp, torch.bfloat16))
            if want_next:
                o.setdefault("next_obs", self._get("r_next%d_%d" % (obs_kind, K), shp, torch.bfloat16))
        a = None
        if policy == L.POLICY_GIVEN:
            a = torch.as_tensor(actions, device=self.device).to(torch.int8).contiguous()
            assert tuple(a.shape) == (K, E, N)
        L.check(self.lib.agv_rollout(self._h, K, proto, policy, obs_kind, _ptr(a), _ptr(o["act"]), _ptr(o["reward"]),
                                     _ptr(o["done"]), _ptr(o.get("plen")), _ptr(o.get("slen")), _ptr(o.get("obs")),
                                     _ptr(o.get("next_obs")), self._stream()), "agv_rollout")
        return o

    def rollout_host(self, n_ticks: int, proto, policy, obs_kind=L.OBS_NONE, actions_host=None, want_next=False,
                     join=True):
        """rollout() through HOST (pinned) buffers: [n_ticks, E, N] actions in, act / reward / done out
        (agv_rollout_host); observations stay on the device."""
        K, E, N = int(n_ticks), self.E, self.N
        hb = self._buf.setdefault("host_r%d" % K, {})
        if not hb:
            hb["act"] = torch.empty((K, E, N), dtype=torch.int8).pin_memory()
            hb["reward"] = torch.empty((K, E, N), dtype=torch.float32).pin_memory()
            hb["done"] = torch.empty((K, E, N), dtype=torch.uint8).pin_memory()
        obs = nxt = None
        if obs_kind != L.OBS_NONE:
            shp = (K, E, N) + self.obs_shape(obs_kind)
            obs = self._get("r_obs%d_%d" % (obs_kind, K), shp, torch.bfloat16)
            if want_next:
                nxt = self._get("r_next%d_%d" % (obs_kind, K), shp, torch.bfloat16)
        ah = None
        if policy == L.POLICY_GIVEN:
            ah = actions_host
            assert ah.dtype == torch.int8 and tuple(ah.shape) == (K, E, N) and ah.is_contiguous()
        L.check(self.lib.agv_rollout_host(self._h, K, proto, policy, obs_kind, _ptr(ah), _ptr(hb["act"]),
                                          _ptr(hb["reward"]), _ptr(hb["done"]), _ptr(obs), _ptr(nxt), self._stream()),
                "agv_rollout_host")
        if join:
            self.host_join()
        return {"act": hb["act"], "reward": hb["reward"], "done": hb["done"], "obs": obs, "next_obs": nxt}

    def collect_expert(self, replay, n_ticks: int, obs_kind=L.OBS_CLIP, ticks_per_launch: int = 8):
        """Expert-mode rollout collection (ExpertManager.create_data_by_self, batched): `n_ticks`
        passes of the A_star control mode for every scene, `ticks_per_launch` at a time in one
        agv_rollout launch; the (state, A* action) pair of every AGV turn -- reward and next state
        too, so the same rows also serve DQN-style learners -- goes from the rollout's rings into the
        device replay buffer with ONE push per launch (rows in [tick, scene, agv] order, i.e. exactly
        the order of per-tick pushes).  Returns the number of AGV turns pushed (device tensor)."""
        turns = torch.zeros((), dtype=torch.int64, device=self.device)
        done_ticks = 0
        while done_ticks < n_ticks:
            K = min(int(ticks_per_launch), n_ticks - done_ticks)
            o = self.rollout(K, L.PROTO_ASTAR, L.POLICY_ASTAR, obs_kind, want_next=True, want_lens=False)
            rows = K * self.E * self.N
            replay.push(o["act"].view(-1), o["reward"].view(-1), o["done"].view(-1), o["obs"].view(rows, -1),
                        o["next_obs"].view(rows, -1), None)
            turns += (o["act"] >= 0).sum()
            done_ticks += K
        return turns

    # ------------------------------------------------------------------ sub-step protocol
    def begin_tick(self):
        L.check(self.lib.agv_begin_tick(self._h, self._stream()), "agv_begin_tick")

    def end_tick(self):
        L.check(self.lib.agv_end_tick(self._h, self._stream()), "agv_end_tick")

    def encode(self, k, obs_kind=L.OBS_CLIP, post=False, out=None):
        shp = (self.E,) + self.obs_shape(obs_kind)
        obs = out if out is not None else self._get("sub_obs%d_%d" % (obs_kind, int(post)), shp, torch.bfloat16)
        will = self._get("will_act%d" % int(post), (self.E,), torch.uint8)
        L.check(self.lib.agv_encode(self._h, k, obs_kind, int(post), _ptr(obs), _ptr(will), self._stream()),
                "agv_encode")
        return obs, will

    def bfs_action(self, k, matrix=L.MATRIX_STATE):
        act = self._get("bfs_act", (self.E,), torch.int8)
        plen = self._get("bfs_plen", (self.E,), torch.int16)
        L.check(self.lib.agv_bfs_action(self._h, k, matrix, _ptr(act), _ptr(plen), self._stream()), "agv_bfs_action")
        return act, plen

    def step_turn(self, k, proto, actions):
        a = torch.as_tensor(actions, device=self.device).to(torch.int8).contiguous()
        assert tuple(a.shape) == (self.E,)
        rew = self._get("st_rew", (self.E,), torch.float32)
        done = self._get("st_done", (self.E,), torch.uint8)
        acted = self._get("st_acted", (self.E,), torch.uint8)
        slen = self._get("st_slen", (self.E,), torch.int16)
        self.last_action = self._get("st_act", (self.E,), torch.int8)  # action executed (BFS action for -1)
        L.check(self.lib.agv_step_turn(self._h, k, proto, _ptr(a), _ptr(rew), _ptr(done), _ptr(acted), _ptr(slen),
                                       _ptr(self.last_action), self._stream()), "agv_step_turn")
        return rew, done, acted, slen

    # ------------------------------------------------------------------ state access
    def get_state(self):
        hdr = np.zeros((self.E, 6), dtype=np.int64)
        agv = np.zeros((self.E, self.N, 8), dtype=np.int32)
        L.check(self.lib.agv_get_state(self._h, hdr.ctypes.data_as(C.c_void_p), agv.ctypes.data_as(C.c_void_p),
                                       self._stream()), "agv_get_state")
        return hdr, agv

    def set_state(self, hdr, agv):
        hdr = np.ascontiguousarray(hdr, dtype=np.int64)
        agv = np.ascontiguousarray(agv, dtype=np.int32)
        assert hdr.shape == (self.E, 6) and agv.shape == (self.E, self.N, 8)
        L.check(self.lib.agv_set_state(self._h, hdr.ctypes.data_as(C.c_void_p), agv.ctypes.data_as(C.c_void_p),
                                       self._stream()), "agv_set_state")

    def dump_frame(self, e: int = 0):
        """Off-device frame of scene `e` for a viewer (the optional rendering hook of SURVEY 8f rank 4;
        the pygame renderer itself, Scene.py:177-339, is out of scope): the `all_info` list of
        Scene.create_info (Scene.py:341-352) -- [layout codes, [name, [x, y], [tx, ty], loaded], ...] for the
        created AGVs -- plus the tick counter and the episode-over flag."""
        hdr, agv = self.get_state()
        info = [self.grid.tolist()]
        for j in range(int(hdr[e, 4])):
            x, y, tx, ty, _stage, loaded = (int(v) for v in agv[e, j, :6])
            info.append(["veh%d" % (j + 1), [x, y], [tx, ty], bool(loaded)])
        return {"all_info": info, "running_time": int(hdr[e, 0]), "task_finished": bool(hdr[e, 3]), "over": bool(hdr[e, 5])}

    def counters(self):
        """(AGV turns, episodes ended, invalid action codes) since the last call."""
        out = (C.c_uint64 * 3)()
        L.check(self.lib.agv_read_counters(self._h, out, self._stream()), "agv_read_counters")
        return int(out[0]), int(out[1]), int(out[2])


def bfs_batch(valid, start, target, device=None):
    """FindPathAstar on B matrices at once: valid uint8 [B,H,W], start/target int32
    [B,2] 0-based (x, y).  Returns (action int8[B], pathlen int32[B], found uint8[B])."""
    lib = L.lib()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    v = torch.as_tensor(valid, device=dev).to(torch.uint8).contiguous()
    B, H, W = v.shape
    s = torch.as_tensor(start, device=dev).to(torch.int32).contiguous()
    t = torch.as_tensor(target, device=dev).to(torch.int32).contiguous()
    act = torch.empty(B, dtype=torch.int8, device=dev)
    ln = torch.empty(B, dtype=torch.int32, device=dev)
    fnd = torch.empty(B, dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    L.check(lib.agv_bfs_batch(B, H, W, _ptr(v), _ptr(s), _ptr(t), _ptr(act), _ptr(ln), _ptr(fnd), st), "agv_bfs_batch")
    return act, ln, fnd


def bfs_field(valid, target, device=None):
    """Target-rooted distance fields: uint16 [B,H,W] (0xFFFF unreachable) as int32."""
    lib = L.lib()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    v = torch.as_tensor(valid, device=dev).to(torch.uint8).contiguous()
    B, H, W = v.shape
    t = torch.as_tensor(target, device=dev).to(torch.int32).contiguous()
    dist = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    L.check(lib.agv_bfs_field(B, H, W, _ptr(v), _ptr(t), _ptr(dist), st), "agv_bfs_field")
    return dist.to(torch.int32) & 0xFFFF
