"""MADQNAgentController -- the AG-DQN link between scene and agent, with the reference's interface.

Reference: src/algorithm/MADQN_structure/Controller.py:18-197.  Same constructor, attributes and
methods (create_agent, self_init, model_run, choose_action, store_info, check_determination,
save_neural_network, save_log), so main.py-style drivers keep working.  Two ways to run:

  * `rmfs_scene` is the facade Scene (multiAGVscene/Scene.py, one scene): model_run() is the reference's
    episode loop -- Scene.init(); Scene.run_game("intelligent", smart_controller=self) -- with the per-AGV
    callbacks choose_action / store_info driving the one-decision-at-a-time Agent (MADQN.py).
  * `rmfs_scene` is a BatchedScene (thousands of scenes): model_run() trains the BatchedAgent, whose
    sub-step protocol is the same callback sequence as tensor calls (MADQN.py: BatchedAgent.tick); an
    "episode" of the log is then `ticks_per_log` ticks of every scene.

control_mode "train_NN" creates fresh networks; "use_NN" loads policy_net.pt / target_net.pt from
`storage_path` -- whole-module pickles written by this class OR by the reference (algorithm/checkpoint.py).
matplotlib plots of the reference's model_run tail (SaveManager.draw_picture, plt.show) are out of scope.
"""
from __future__ import annotations

import datetime
import os

import numpy as np
import torch

from ..checkpoint import load_module, save_module
from ..Manager.StateManager import StateManager as stateManager
from .MADQN import Agent, BatchedAgent, Net


class VehObj:
    """per-AGV scratch of the callback protocol (Controller.py:185-197)"""

    def __init__(self, this_veh):
        self.veh_name = this_veh
        self.obs_current = 0
        self.obs_next = 0
        self.obs_forbidden_matrix = 0
        self.action = []
        self.reward = 0
        self.is_end = False
        self.last_state = 0
        self.last_state_store = False


class MADQNAgentController:
    """a link between environment and algorithm"""

    def __init__(self, rmfs_scene, map_xdim, map_ydim, max_task, control_mode=1, state_number=4, expert_guiding=False):
        self.control_mode = control_mode
        self.state_number = state_number
        self.rmfs_model = rmfs_scene
        self.stateManager = stateManager()
        self.storage_path = os.path.join(os.path.split(os.path.abspath(__file__))[0], "network_picture")
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.batched = not hasattr(rmfs_scene, "run_game")  # a BatchedScene has no per-AGV callback loop
        self.agent = None
        self.create_agent(map_xdim, map_ydim)
        # training parameters (Controller.py:36-48)
        self.simulation_times = 150
        self.max_value = max_task * 3 - 1
        self.max_value_times = 0
        self.duration_times = 100
        self.lr_start_decay = False
        self.lifelong_reward = []
        self.action_length_record = 0
        self.reward_acc = 0
        self.veh_group = []
        self.logs = []
        self.ticks_per_log = 64  # batched mode: ticks of every scene per logged "episode"

    # ------------------------------------------------------------------ networks
    def create_agent(self, map_xdim, map_ydim):
        """create / load the two networks and the agent (Controller.py:50-63)"""
        n_act = getattr(self.rmfs_model, "action_number", 4)
        if self.control_mode == "use_NN":
            policy_net = load_module(os.path.join(self.storage_path, "policy_net.pt"))
            target_net = load_module(os.path.join(self.storage_path, "target_net.pt"))
        else:  # "train_NN"
            policy_net = Net(self.state_number, n_act, map_xdim, map_ydim)
            target_net = Net(self.state_number, n_act, map_xdim, map_ydim)
        if self.batched:
            self.agent = BatchedAgent(self.rmfs_model, policy_net, target_net, memory_size=1 << 20, batch_size=256)
            if self.control_mode == "use_NN":
                self.agent.epsilon = 1.0
        else:
            self.agent = Agent(policy_net, target_net, self.device)

    def self_init(self):
        self.reward_acc = 0
        self.veh_group = []
        self.action_length_record = 0

    # ------------------------------------------------------------------ main loop
    def model_run(self):
        start_time = datetime.datetime.now()
        print("Training starts at" + str(start_time))
        for i_episode in range(self.simulation_times):
            self.self_init()
            if self.batched:
                running_time = self._batched_episode()
            else:
                self.rmfs_model.init()
                running_time = self.rmfs_model.run_game(control_pattern="intelligent", smart_controller=self, render=True)
            self.lifelong_reward.append(int(self.reward_acc))
            log = 'i_episode: {},\t reward_accu: {}, \t action_length: {},\t running times: {}'.format(
                i_episode, self.reward_acc, self.action_length_record, running_time)
            self.logs.append(log)
            print(log)
            if self.lr_start_decay:
                if hasattr(self.agent, "change_learning_rate"):
                    self.agent.change_learning_rate(times=100)
                self.agent.change_explore_rate(times=100)
            if i_episode % 100 == 0:
                self.save_neural_network(auto=True)
            if self.check_determination(self.reward_acc):
                break
        self.save_neural_network(auto=False)
        self.logs.append(self.lifelong_reward)
        self.save_log()
        print(datetime.datetime.now() - start_time)

    def _batched_episode(self):
        """`ticks_per_log` ticks of every scene with the network in the loop; the logged reward is the mean
        accumulated reward per scene, so that max_value keeps its meaning"""
        ag, sc = self.agent, self.rmfs_model
        r0 = float(ag.reward_acc.item())
        train = self.control_mode != "use_NN"
        turns = 0
        for _ in range(self.ticks_per_log):
            turns += int(ag.tick(train=train))
        self.reward_acc = (float(ag.reward_acc.item()) - r0) / sc.E
        self.action_length_record = turns
        return self.ticks_per_log

    # ------------------------------------------------------------------ callbacks of Scene.run_mode
    def _veh(self, this_veh, create=False):
        for veh in self.veh_group:
            if veh.veh_name == this_veh:
                return veh
        if not create:
            return None
        veh = VehObj(this_veh)
        self.veh_group.append(veh)
        return veh

    def choose_action(self, all_info, this_veh):
        """Scene.py:146 -> action for `this_veh` (Controller.py:105-126)"""
        veh_obj = self._veh(this_veh, create=True)
        obs, cp, tp, valid_path_matrix = self.stateManager.create_state(all_info, this_veh, obs_clip=True)
        obs = np.array(obs)
        veh_obj.obs_current = obs
        epsilon = 1. if self.control_mode == "use_NN" else 0.
        action = self.agent.choose_action(obs, current_place=cp, target_place=tp, valid_path_matrix=valid_path_matrix,
                                          epsilon=epsilon)[0]
        veh_obj.action.append(action)
        self.action_length_record += 1
        return action

    def store_info(self, all_info, reward, is_end, this_veh):
        """Scene.py:156 (Controller.py:128-145)"""
        self.reward_acc += reward
        if self.control_mode == "use_NN":
            return
        veh_obj = self._veh(this_veh)
        obs, _cp, _tp, _valid = self.stateManager.create_state(all_info, this_veh, obs_clip=True)
        veh_obj.obs_next, veh_obj.reward = np.array(obs), reward
        self.agent.store_transition(veh_obj.obs_current, veh_obj.action[-1], veh_obj.reward, veh_obj.obs_next,
                                    1 if is_end else 0)

    def check_determination(self, reward_accu):
        """end training after `duration_times` consecutive episodes at the maximum reward; the first one
        starts the decay of learning / exploring rate (Controller.py:147-159)"""
        if int(reward_accu) >= self.max_value:
            self.lr_start_decay = True
            self.max_value_times += 1
        else:
            self.max_value_times = 0
        return self.max_value_times == self.duration_times

    # ------------------------------------------------------------------ persistence
    def save_neural_network(self, auto=False):
        """policy_net[_auto].pt / target_net[_auto].pt, loadable by the reference's use_NN"""
        if self.control_mode == "use_NN":
            return
        os.makedirs(self.storage_path, exist_ok=True)
        tag = "_auto" if auto else ""
        if hasattr(self.agent, "finish_update"):
            self.agent.finish_update()
        save_module(self.agent.policy_network, os.path.join(self.storage_path, "policy_net%s.pt" % tag))
        save_module(self.agent.target_network, os.path.join(self.storage_path, "target_net%s.pt" % tag))

    def save_log(self):
        if self.control_mode == "use_NN":
            return
        os.makedirs(self.storage_path, exist_ok=True)
        with open(os.path.join(self.storage_path, "logs.txt"), 'w') as f:
            for one_log in self.logs:
                f.write(str(one_log))
                f.write("\r\n")
