"""Expert -- A*-labelled trajectory collection, the reference's CSV wire format, and
behavioural-cloning batches.

Reference: src/algorithm/Manager/ExpertManager.py:15-179.  Same constructor, file
naming, CSV header/row layout (`layout,info,veh_name,action`; one row per trajectory)
and sampling protocol, so files written here feed the reference's pre_training and
vice versa.  Collection is batched: all `times` trajectories run at once on the GPU
with the fused A_star tick (agv_tick, PROTO_ASTAR + POLICY_ASTAR); the per-turn
snapshots the CSV needs are assembled from two state dumps per tick (an AGV only
changes its own record during its own turn).  BC batches are encoded by the CUDA
encoder through StateManager.create_state_batch.
"""
from __future__ import annotations

import csv
import json
import os
import random

import numpy as np
import torch

from ... import _lib as L
from ...batched import BatchedScene
from .StateManager import StateManager as stateManager

csv.field_size_limit(500 * 1024 * 1024)


class Expert:
    def __init__(self, multi_agv_scene=None, ss_x_width=3, ss_y_width=2, ss_x_num=2, ss_y_num=2, ps_num=2,
                 explorer_num=1, filename=None):
        self.storage_path = os.path.join(os.path.split(os.path.abspath(__file__))[0], "ExpertData")
        if filename is None:
            file_name = self.create_file_name([ss_x_width, ss_y_width, ss_x_num, ss_y_num, ps_num, explorer_num])
        else:
            file_name = filename
        self.file_name = file_name if os.path.isabs(file_name) else os.path.join(self.storage_path, file_name)
        self.trajectory = []
        self.headers = ["layout", "info", "veh_name", "action"]
        self.multi_agv_scene = multi_agv_scene
        self.stateManager = stateManager()
        self.expert_size = self.count_lines()
        self.sample_mode_list = ["points", "trajectory"]
        self.sample_mode = self.sample_mode_list[0]
        self.batch_size = 512
        self.lr = 0.0005
        self.all_data = []
        self.max_ticks = None       # cap per trajectory (A_star mode can gridlock)
        self.truncated = 0          # trajectories that hit the cap in the last collection

    # ------------------------------------------------------------------ collection
    def collect(self, times):
        """Runs `times` A_star-mode trajectories at once.  Returns a list of
        (layout, states_json, names, actions, finished) in trajectory order."""
        scene = self.multi_agv_scene
        layout = scene.layout
        names = [ex.explorer_name for ex in scene.explorer_group]
        N = len(names)
        task_sets = []
        for _ in range(times):  # one Scene.init() per trajectory draws its task list (ExpertManager.py:42)
            layout.init()
            task_sets.append(layout.task_array())
        tasks = np.stack(task_sets)
        dev = BatchedScene(layout.codes, times, N, tasks=tasks, P=3,
                           always_loaded=bool(scene.explorer_group[0].always_loaded),
                           always_empty=bool(scene.explorer_group[0].always_empty), auto_reset=False,
                           max_steps=scene.max_training_steps)
        dev.reset()
        cap = self.max_ticks if self.max_ticks is not None else scene.max_training_steps
        states = [[] for _ in range(times)]
        vehs = [[] for _ in range(times)]
        acts = [[] for _ in range(times)]
        hdr0, agv0 = dev.get_state()
        tick = 0
        while tick < cap and not hdr0[:, 5].all():
            out = dev.tick(L.PROTO_ASTAR, L.POLICY_ASTAR, want_lens=False)
            act = out["act"].cpu().numpy()
            hdr1, agv1 = dev.get_state()
            for e in range(times):
                if hdr0[e, 5]:
                    continue
                nc = int(hdr0[e, 4])
                for i in range(N):
                    if act[e, i] < 0:
                        continue
                    # snapshot before AGV i's turn: AGVs j < i show their end-of-tick record
                    snap = []
                    for j in range(nc):
                        r = agv1[e, j] if j < i else agv0[e, j]
                        snap.append([names[j], [int(r[0]), int(r[1])], [int(r[2]), int(r[3])], bool(r[5])])
                    states[e].append(json.dumps(snap))
                    vehs[e].append(names[i])
                    acts[e].append(int(act[e, i]))
            hdr0, agv0 = hdr1, agv1
            tick += 1
        finished = hdr0[:, 5].astype(bool)
        self.truncated = int((~finished).sum())
        dev.close()
        lay = layout.layout_original
        return [(lay, states[e], vehs[e], acts[e], bool(finished[e])) for e in range(times)]

    def create_data_by_self(self, times=50):
        """ExpertManager.py:35-52: header + one CSV row per trajectory"""
        os.makedirs(os.path.dirname(self.file_name), exist_ok=True)
        self.check_csv()
        rows = self.collect(times)
        with open(self.file_name, mode="a", encoding="utf-8", newline="") as csvfile:
            writer = csv.writer(csvfile)
            writer.writerow(self.headers)
            for lay, st, nm, ac, _ in rows:
                writer.writerow([lay, st, nm, ac])
        self.expert_size = self.count_lines()
        print("Collecting finished, line number is " + str(self.expert_size))

    def create_data_by_rl(self):
        pass

    # ------------------------------------------------------------------ sampling
    def sample_data(self, sample_n=1):
        """ExpertManager.py:57-67: one random trajectory row -> (states, actions)"""
        if not self.all_data:
            self.all_data = self.read_all_data_csv()
        order = random.randint(1, self.expert_size)
        row = self.all_data[order]
        kind = "P" if self.sample_mode == self.sample_mode_list[0] else "T"
        return self.analyse_data(row, sample_type=kind)

    def analyse_data(self, sample_data, sample_type="T"):
        layout = json.loads(sample_data[0])
        all_info = self.data_restore(sample_data[1], layout)
        veh_name = json.loads(sample_data[2].replace("'", '"'))
        action = json.loads(sample_data[3])
        if sample_type == "T":
            return self.form_trajectory(all_info, veh_name, action)
        return self.form_point(all_info, veh_name, action)

    def form_trajectory(self, all_info, veh_name, action):
        obs = self.stateManager.create_state_batch(all_info, veh_name)
        return [o.astype(np.float64) for o in obs], action

    def form_point(self, all_info, veh_name, action):
        """ExpertManager.py:89-100: batch_size random steps of the trajectory (the batch size
        shrinks permanently when a trajectory is shorter, like the reference)"""
        if len(all_info) <= self.batch_size:
            self.batch_size = len(all_info)
        orders = random.sample(range(0, len(all_info)), self.batch_size)
        obs = self.stateManager.create_state_batch([all_info[i] for i in orders], [veh_name[i] for i in orders])
        return [o.astype(np.float64) for o in obs], [action[i] for i in orders]

    @staticmethod
    def data_restore(data, layout):
        """CSV `info` cell (the str() of a list of JSON strings) -> list of all_info
        snapshots [layout, [name, [x, y], [tx, ty], loaded], ...] (ExpertManager.py:102-114)"""
        swapped = data.replace('"', "\x00").replace("'", '"').replace("\x00", "'")
        out = []
        for item in json.loads(swapped):
            snap = json.loads(item.replace("'", '"'))
            snap.insert(0, layout)
            out.append(snap)
        return out

    def pre_training(self, policy_network, device, lr_=0.0005):
        """one behavioural-cloning step: loss = mean(-log pi(a*|s)), Adam (ExpertManager.py:116-132)"""
        optimizer = torch.optim.Adam(policy_network.parameters(), lr=lr_)
        expert_s, expert_a = self.sample_data(sample_n=1)
        obs = torch.from_numpy(np.stack(expert_s)).float().to(device)
        actions = torch.LongTensor(np.array(expert_a)).unsqueeze(1).to(device)
        log_probs = torch.log(policy_network(obs).gather(1, actions)).to(device)
        bc_loss = torch.mean(-log_probs)
        optimizer.zero_grad()
        bc_loss.backward()
        optimizer.step()
        return bc_loss

    # ------------------------------------------------------------------ csv plumbing
    def write_data_csv(self, state_action=None):
        with open(self.file_name, mode="a", encoding="utf-8", newline="") as csvfile:
            csv.writer(csvfile).writerow(self.headers if state_action is None else state_action)

    def read_data_csv(self, read_line=1):
        return self.read_all_data_csv()[read_line]

    def read_all_data_csv(self):
        with open(self.file_name, "r", encoding="utf-8") as csvfile:
            return list(csv.reader(csvfile))

    def clear_csv(self):
        with open(self.file_name, mode="w", encoding="utf-8", newline="") as csvfile:
            csvfile.truncate()

    def check_csv(self):
        if os.path.exists(self.file_name):
            print("csv file already exists, recover it")
            self.clear_csv()

    @staticmethod
    def create_file_name(paras):
        return "expert_data_" + "_".join(str(p) for p in paras) + ".csv"

    def count_lines(self):
        if not os.path.exists(self.file_name):
            return 0
        with open(self.file_name, "r", encoding="utf-8") as csvfile:
            return sum(1 for _ in csv.reader(csvfile)) - 1
