"""StateManager -- observation builder with the reference's interface, run on the GPU.

Reference: src/algorithm/Manager/StateManager.py:14-192.  create_state(all_info,
this_veh, obs_clip, padding_size) takes the snapshot list produced by
Scene.create_info ([layout, [name, [x, y], [tx, ty], loaded], ...]); the snapshot is
loaded into a scratch device scene (agv_set_state) and the CUDA encoder
(agv_encode) emits the three planes [current, target, valid path].  Values are 0/1,
so the bf16 device planes are exact.  create_state_batch encodes many snapshots in one
launch (what Expert.form_point needs).
"""
from __future__ import annotations

import numpy as np

from ... import _lib as L
from ...batched import BatchedScene


class StateManager:
    def __init__(self):
        self.padding_size = 0
        self._scenes = {}

    # ------------------------------------------------------------------ device scratch scenes
    def _scene(self, layout, n_env, n_veh, P):
        codes = np.asarray(layout)
        codes = np.where(codes == 1, 1, np.where((codes == 2) | (codes == 3), 2, 0)).astype(np.uint8)
        # code 3 (obstacle) is impassable like a picking station for the valid plane
        # (StateManager.py:109-114); it can never be a target
        key = (codes.tobytes(), codes.shape, n_env, n_veh, P)
        sc = self._scenes.get(key)
        if sc is None:
            if len(self._scenes) > 8:
                self._scenes.pop(next(iter(self._scenes))).close()
            sc = BatchedScene(codes, n_env, n_veh, tasks=None, P=P)
            self._scenes[key] = sc
        return sc

    @staticmethod
    def all_info_analysis(all_info, this_veh):
        """StateManager.py:122-133"""
        current_place, target_place, occupied_place, occupied_target, veh_loaded = 0, 0, [], [], False
        for one_veh in all_info[1:]:
            if one_veh[0] == this_veh:
                current_place, target_place, veh_loaded = one_veh[1], one_veh[2], one_veh[3]
            else:
                occupied_place.append(one_veh[1])
                occupied_target.append(one_veh[2])
        return current_place, target_place, occupied_place, occupied_target, veh_loaded

    def _encode(self, items, obs_clip, P):
        """items: list of (all_info, veh_name) sharing one layout.  Returns
        (obs_clip_or_None, obs_full) as float32 numpy [B,3,..]"""
        layout = items[0][0][0]
        B = len(items)
        n_veh = max(len(ai) - 1 for ai, _ in items)
        sc = self._scene(layout, B, n_veh, P)
        hdr = np.zeros((B, 6), dtype=np.int64)
        agv = np.ones((B, n_veh, 8), dtype=np.int32)
        agv[:, :, 4:] = 0
        which = np.zeros(B, dtype=np.int64)
        for b, (ai, name) in enumerate(items):
            vehs = ai[1:]
            hdr[b, 4] = len(vehs)
            found = False
            for j, v in enumerate(vehs):
                agv[b, j, 0:2] = v[1]
                agv[b, j, 2:4] = v[2]
                agv[b, j, 5] = 1 if v[3] else 0
                if v[0] == name:
                    which[b], found = j, True
            if not found:
                raise KeyError("vehicle %r is not in all_info" % (name,))
        sc.set_state(hdr, agv)
        full = np.zeros((B, 3, sc.H, sc.W), dtype=np.float32)
        clip = np.zeros((B, 3, sc.K, sc.K), dtype=np.float32) if obs_clip else None
        for k in np.unique(which):
            sel = which == k
            o, _ = sc.encode(int(k), L.OBS_FULL)
            full[sel] = o.float().cpu().numpy()[sel]
            if obs_clip:
                o, _ = sc.encode(int(k), L.OBS_CLIP)
                clip[sel] = o.float().cpu().numpy()[sel]
        return clip, full

    # ------------------------------------------------------------------ reference API
    def create_state(self, all_info, this_veh, obs_clip=False, padding_size=3):
        """-> (obs, current_place, target_place, valid_path_matrix) as nested lists
        (StateManager.py:14-41)"""
        self.padding_size = padding_size
        current_place, target_place, _, _, _ = self.all_info_analysis(all_info, this_veh)
        clip, full = self._encode([(all_info, this_veh)], obs_clip, padding_size)
        valid_path_matrix = full[0, 2].astype(np.float64).tolist()
        obs = (clip[0] if obs_clip else full[0]).astype(np.float64).tolist()
        return obs, current_place, target_place, valid_path_matrix

    def create_state_batch(self, all_infos, vehs, obs_clip=False, padding_size=3):
        """many snapshots at once -> float32 numpy [B,3,K,K] or [B,3,H,W]"""
        self.padding_size = padding_size
        clip, full = self._encode(list(zip(all_infos, vehs)), obs_clip, padding_size)
        return clip if obs_clip else full

    def create_path_matrix(self, layout, veh_loaded, current_place, target_place, occupied_place):
        """StateManager.py:43-61 -> (valid_path_matrix, forbidden_path_matrix)"""
        info = [layout, ["_self", list(current_place), list(target_place), bool(veh_loaded)]]
        info += [["_o%d" % i, list(p), [1, 1], False] for i, p in enumerate(occupied_place or [])]
        _, full = self._encode([(info, "_self")], False, max(self.padding_size, 0))
        valid = full[0, 2].astype(np.float64)
        return valid.tolist(), (1.0 - valid).tolist()

    def create_position_matrix(self, layout, current_place, target_place):
        """StateManager.py:82-90"""
        info = [layout, ["_self", list(current_place), list(target_place), False]]
        _, full = self._encode([(info, "_self")], False, max(self.padding_size, 0))
        return full[0, 0].astype(np.float64).tolist(), full[0, 1].astype(np.float64).tolist()
