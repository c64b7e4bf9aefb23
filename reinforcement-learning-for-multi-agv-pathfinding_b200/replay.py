"""DeviceReplay -- host mirror of the agv_replay_* C ABI (device-resident replay buffer
with the reference's PER sum-tree; src/algorithm/MADQN_structure/PER.py:6-112)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class DeviceReplay:
    def __init__(self, capacity: int, obs_shape, device: int | None = None):
        if not torch.cuda.is_available():
            raise L.AgvError("DeviceReplay needs a CUDA device; there is no CPU fallback")
        self.lib = L.lib()
        self.capacity = int(capacity)
        self.obs_shape = tuple(int(v) for v in obs_shape)
        self.obs_elems = int(np.prod(self.obs_shape)) if self.obs_shape else 0
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.empty(1, device=self.device)
        h = C.c_void_p()
        L.check(self.lib.agv_replay_create(self.device_index, self.capacity, self.obs_elems, C.byref(h)),
                "agv_replay_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.agv_replay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def push(self, act, reward=None, done=None, obs=None, next_obs=None, priority=None):
        """Memory.add for every row with act >= 0, in row order.  Tensors are device
        tensors: act int8[rows], reward f32, done u8, obs/next_obs bf16[rows, ...],
        priority f64[rows] = (|error|+0.01)**0.6 (None -> 1.0)."""
        act = act.contiguous().view(-1)
        n = act.numel()
        for t, dt in ((act, torch.int8), (reward, torch.float32), (done, torch.uint8), (obs, torch.bfloat16),
                      (next_obs, torch.bfloat16), (priority, torch.float64)):
            assert t is None or (t.dtype == dt and t.is_cuda and t.is_contiguous()), (dt, t.dtype if t is not None else None)
        if obs is not None:
            assert obs.numel() == n * self.obs_elems
        L.check(self.lib.agv_replay_push(self._h, n, _ptr(act), _ptr(reward), _ptr(done), _ptr(obs), _ptr(next_obs),
                                         _ptr(priority), self._stream()), "agv_replay_push")

    def _alloc_batch(self, n):
        dev = self.device
        return dict(slot=torch.empty(n, dtype=torch.int64, device=dev),
                    prio=torch.empty(n, dtype=torch.float64, device=dev),
                    obs=torch.empty((n,) + self.obs_shape, dtype=torch.bfloat16, device=dev),
                    act=torch.empty(n, dtype=torch.int8, device=dev),
                    reward=torch.empty(n, dtype=torch.float32, device=dev),
                    next_obs=torch.empty((n,) + self.obs_shape, dtype=torch.bfloat16, device=dev),
                    done=torch.empty(n, dtype=torch.uint8, device=dev))

    def sample(self, n: int, u01):
        """Memory.sample(n) with the uniforms injected (u01 float64[n] in [0,1))."""
        u = torch.as_tensor(u01, dtype=torch.float64, device=self.device).contiguous()
        assert u.numel() == n
        b = self._alloc_batch(n)
        L.check(self.lib.agv_replay_sample(self._h, n, _ptr(u), _ptr(b["slot"]), _ptr(b["prio"]), _ptr(b["obs"]),
                                           _ptr(b["act"]), _ptr(b["reward"]), _ptr(b["next_obs"]), _ptr(b["done"]),
                                           self._stream()), "agv_replay_sample")
        return b

    def gather(self, slots):
        s = torch.as_tensor(slots, dtype=torch.int64, device=self.device).contiguous()
        n = s.numel()
        b = self._alloc_batch(n)
        b["slot"] = s
        L.check(self.lib.agv_replay_gather(self._h, n, _ptr(s), _ptr(b["obs"]), _ptr(b["act"]), _ptr(b["reward"]),
                                           _ptr(b["next_obs"]), _ptr(b["done"]), self._stream()), "agv_replay_gather")
        return b

    def update(self, tree_idx, priority):
        """Memory.update: tree indices (slot + capacity - 1) and new priorities"""
        i = torch.as_tensor(tree_idx, dtype=torch.int64, device=self.device).contiguous()
        p = torch.as_tensor(priority, dtype=torch.float64, device=self.device).contiguous()
        L.check(self.lib.agv_replay_update(self._h, i.numel(), _ptr(i), _ptr(p), self._stream()), "agv_replay_update")

    def info(self):
        """(n_entries, write cursor, SumTree.total())"""
        out = (C.c_int64 * 2)()
        tot = C.c_double(0)
        L.check(self.lib.agv_replay_info(self._h, out, C.byref(tot), self._stream()), "agv_replay_info")
        return int(out[0]), int(out[1]), float(tot.value)
