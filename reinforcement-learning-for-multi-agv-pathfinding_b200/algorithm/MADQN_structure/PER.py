"""Prioritised replay with the reference's interface, stored on the GPU.

Reference: src/algorithm/MADQN_structure/PER.py:6-112 (SumTree, Memory).  Samples are
the AG-DQN transitions (s, a, r, s_, is_done) of MADQN.py:129; they live in the
device ring buffer (bf16 observations, exact for the 0/1 state planes) and the
float64 sum-tree is walked on the device.  The uniforms of Memory.sample come from
Python's `random` exactly like the reference (one random() per segment).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from ...replay import DeviceReplay


class SumTree:
    """view of the device sum-tree (heap layout, leaf of slot s = s + capacity - 1)"""

    def __init__(self, capacity, replay=None):
        self.capacity = capacity
        self._replay = replay

    @property
    def n_entries(self):
        return self._replay.info()[0] if self._replay is not None else 0

    @property
    def write(self):
        return self._replay.info()[1] if self._replay is not None else 0

    def total(self):
        return self._replay.info()[2] if self._replay is not None else 0.0


class Memory:  # stored as (s, a, r, s_, done)
    e = 0.01
    a = 0.6
    beta = 0.4
    beta_increment_per_sampling = 0.001

    def __init__(self, capacity, device=None):
        self.capacity = capacity
        self._device = device
        self._replay = None
        self.tree = SumTree(capacity)

    def _get_priority(self, error):
        return (np.abs(error) + self.e) ** self.a

    def _ensure(self, obs_shape):
        if self._replay is None:
            self._replay = DeviceReplay(self.capacity, obs_shape, device=self._device)
            self.tree = SumTree(self.capacity, self._replay)

    def add(self, error, sample):
        """Memory.add (PER.py:82-84): one transition"""
        s, a, r, s_, done = sample
        s = np.asarray(s)
        self._ensure(s.shape)
        dev = self._replay.device
        p = torch.tensor([float(self._get_priority(np.asarray(error, dtype=np.float64)))], dtype=torch.float64,
                         device=dev)
        self._replay.push(torch.tensor([int(a)], dtype=torch.int8, device=dev),
                          torch.tensor([float(r)], dtype=torch.float32, device=dev),
                          torch.tensor([int(done)], dtype=torch.uint8, device=dev),
                          torch.as_tensor(s, dtype=torch.float32).to(dev).to(torch.bfloat16).reshape(1, -1).contiguous(),
                          torch.as_tensor(np.asarray(s_), dtype=torch.float32).to(dev).to(torch.bfloat16)
                          .reshape(1, -1).contiguous(), p)

    def add_batch(self, errors, act, reward, done, obs, next_obs):
        """batched Memory.add of device tensors (rows with act < 0 are skipped)"""
        self._ensure(tuple(obs.shape[1:]))
        pr = (errors.to(torch.float64).abs() + self.e) ** self.a
        self._replay.push(act, reward, done, obs, next_obs, pr.contiguous())

    def sample(self, n):
        """Memory.sample (PER.py:86-108) -> (batch, idxs, is_weight)"""
        self.beta = np.min([1., self.beta + self.beta_increment_per_sampling])
        u = [random.random() for _ in range(n)]  # random.uniform(a, b) == a + (b-a)*random()
        b = self._replay.sample(n, u)
        slots = b["slot"].cpu().numpy()
        idxs = [int(s) + self.capacity - 1 for s in slots]
        priorities = b["prio"].cpu().numpy()
        obs = b["obs"].float().cpu().numpy()
        nxt = b["next_obs"].float().cpu().numpy()
        act = b["act"].cpu().numpy()
        rew = b["reward"].cpu().numpy()
        done = b["done"].cpu().numpy()
        batch = [(obs[i], int(act[i]), float(rew[i]), nxt[i], int(done[i])) for i in range(n)]
        n_entries, _, total = self._replay.info()
        sampling_probabilities = priorities / total
        is_weight = np.power(n_entries * sampling_probabilities, -self.beta)
        is_weight /= is_weight.max()
        return batch, idxs, is_weight

    def sample_device(self, n):
        """same draw, but the batch stays on the device (dict of tensors)"""
        self.beta = np.min([1., self.beta + self.beta_increment_per_sampling])
        return self._replay.sample(n, [random.random() for _ in range(n)])

    def update(self, idx, error):
        """Memory.update (PER.py:110-112)"""
        p = self._get_priority(np.asarray(error, dtype=np.float64))
        self._replay.update([int(idx)], [float(p)])
