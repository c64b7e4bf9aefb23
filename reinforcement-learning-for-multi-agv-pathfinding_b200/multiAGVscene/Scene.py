"""Scene -- tick loop, turn order, control modes of ONE scene, backed by the CUDA path.

Drop-in for the reference's src/multiAGVscene/Scene.py (logic part: 15-24, 56, 61-65,
82-175, 341-366): same constructor, attributes, run_game(control_pattern,
smart_controller, render) protocol and return values.  Rendering (pygame) is out of
scope: `render` is accepted and ignored.  This class drives a 1-scene BatchedScene
through the sub-step C ABI (agv_begin_tick / agv_bfs_action / agv_step_turn /
agv_end_tick) so that the callback protocol (choose_action / store_info, one AGV at
a time) is reproduced exactly; thousands of scenes at once use BatchedScene directly.
"""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from ..batched import BatchedScene
from ..utils.settings import SuperParas as SuperPara
from ..utils.utils import ColorBox, Direction as Dir, Working


class Scene:
    def __init__(self, layout, explorer_group):
        self.layout = layout
        self.control_pattern = ""
        self.clock = None
        self.running_time = 0
        self.FPS = SuperPara.FPS
        self.x_width = layout.scene_x_width
        self.y_width = layout.scene_y_width
        self.max_training_steps = int((self.x_width / 2 + self.y_width) * 3 * ((self.x_width * self.y_width) / 2.5))
        self.color_box = ColorBox()
        self.explorer_group = explorer_group
        if len(explorer_group) == 0:
            print("WARNING: the number of veh is zero")
            return
        self.working_manager = Working()
        self.dir = Dir()
        self.action_number = self.dir.action_num()
        self.smart_controller = None
        self.render = True
        self.manual_actions = None  # optional iterator of action names for "manual" mode (no keyboard here)
        for i, ex in enumerate(explorer_group):
            ex._attach(self, i)
        n = len(explorer_group)
        self._dev = BatchedScene(layout.codes, 1, n, tasks=layout.task_array()[None], P=3,
                                 always_loaded=bool(explorer_group[0].always_loaded),
                                 always_empty=bool(explorer_group[0].always_empty),
                                 shaped_reward=bool(SuperPara.Sparse_Reward), auto_reset=False,
                                 max_steps=self.max_training_steps)
        self._tasks_uploaded = list(layout.task_list)
        self._hdr = None
        self._fresh = False   # device state == "Scene.init() done, no AGV created"
        self._device_init()

    # ------------------------------------------------------------------ device sync
    def _device_init(self):
        """Scene.init on the device WITHOUT creating AGV 0 (run_game does that)."""
        if list(self.layout.task_list) != self._tasks_uploaded:
            self._dev.set_tasks(self.layout.task_array()[None])
            self._tasks_uploaded = list(self.layout.task_list)
        hdr, agv = self._dev.get_state()
        hdr[:] = 0
        agv[0, :, 0:2] = 1
        agv[0, :, 4:8] = 0
        for i, ex in enumerate(self.explorer_group):  # targets survive Explorer.init
            agv[0, i, 2:4] = ex.target_position
        self._dev.set_state(hdr, agv)
        self._fresh = True
        self._pull()

    def _pull(self):
        hdr, agv = self._dev.get_state()
        self._hdr = hdr[0]
        n_created = int(hdr[0, 4])
        for i, ex in enumerate(self.explorer_group):
            ex._refresh(agv[0, i], i < n_created)
        lay = self.layout
        cursor, n_done = int(hdr[0, 1]), int(hdr[0, 2])
        lay.task_finished = bool(hdr[0, 3])
        # task_arrangement = [[order], [veh], [done]] (Layout.py:50); veh names / done flags are
        # reconstructed from the AGV records: an assigned task is done unless an AGV still holds it
        held = {ex.task_order: ex.explorer_name for ex in self.explorer_group
                if ex.has_created and not (ex.all_assigned and ex.task_stage == 0)}
        prev = lay.task_arrangement
        names = list(prev[1][:cursor]) + [None] * max(0, cursor - len(prev[1]))
        for order, name in held.items():
            if order < cursor and names[order] is None:
                names[order] = name
        done = [0 if o in held else 1 for o in range(cursor)]
        if sum(done) != n_done:  # an idle AGV keeps its last (finished) order
            done = [1] * cursor
            for o in sorted(held, reverse=True):
                if sum(done) > n_done and o < cursor:
                    done[o] = 0
        lay.task_arrangement = [list(range(cursor)), names, done]

    def _create_explorer(self, index):
        if index != int(self._hdr[4]):
            raise L.AgvError("AGVs are created in list order (Scene.py:354-366): next is %d" % int(self._hdr[4]))
        if index == 0 and self._fresh:
            self._dev.reset()  # Scene.init state + create_explorer of AGV 0 (Scene.py:86)
            self._fresh = False
        else:
            raise L.AgvError("later AGVs are created by check_new_veh")
        self._pull()

    # ------------------------------------------------------------------ reference API
    def init(self):
        """Scene.py:61-65"""
        self.layout.init()
        for ex in self.explorer_group:
            ex.init()
        self._device_init()

    def render_init(self):
        pass

    def render_step(self):
        pass

    def create_info(self):
        """[layout_original, [name, [x, y], [tx, ty], loaded] per created AGV] (Scene.py:341-352)"""
        info = [self.layout.layout_original]
        for ex in self.explorer_group:
            if not ex.has_created:
                break
            info.append([ex.explorer_name, list(ex.current_place), list(ex.target_position), ex.loaded])
        return info

    def check_new_veh(self):
        """Scene.py:354-366: returns the index of the AGV created this tick (0 = none)"""
        before = int(self._hdr[4])
        self._dev.end_tick()
        self._pull()
        return before if int(self._hdr[4]) > before else 0

    def _turn_state(self, i):
        """Scene.py:124-134 on the mirrored state"""
        ex = self.explorer_group[i]
        if not ex.has_created:
            return -1
        if self.layout.task_finished:
            return -2
        return 0 if ex.all_assigned else 1

    def _bfs_action(self, i, matrix):
        hdr_over = int(self._hdr[5])
        if hdr_over:
            raise L.AgvError("the episode is over; call Scene.init()")
        act, _ = self._dev.bfs_action(i, matrix)
        a = int(act.cpu()[0])
        return 4 if a < 0 else a

    def _valid_matrix(self, i, matrix):
        """matrix (a)/(b) of AGV i as nested int lists, via the device encoder / BFS field"""
        import torch
        from ..batched import bfs_field
        ex = self.explorer_group[i]
        if matrix == L.MATRIX_STATE:
            obs, _ = self._dev.encode(i, L.OBS_FULL)
            return [[int(v) for v in row] for row in obs[0, 2].float().cpu().numpy()]
        # matrix (a) differs from (b) only by the parked / own cells: rebuild it
        # from the layout codes and the mirrored positions -- pure data layout, no search
        codes = self.layout.codes
        v = (codes == 0) | ((codes == 1) & (not ex.loaded))
        v = v.astype(np.int64)
        v[ex.target_position[1] - 1, ex.target_position[0] - 1] = 1
        if len(self.explorer_group) > 1:
            for o in self.explorer_group:
                v[o.current_place[1] - 1, o.current_place[0] - 1] = 0
        return v.tolist()

    def _execute(self, i, action, proto):
        rew, done, acted, slen = self._dev.step_turn(i, proto, np.array([action], dtype=np.int8))
        acted_v = int(acted.cpu()[0])
        r32, d, sl = float(rew.cpu()[0]), bool(int(done.cpu()[0])), int(slen.cpu()[0]) & 0xFFFF
        self._pull()
        if not acted_v:
            raise L.AgvError("AGV %d cannot take a turn now (not created, idle or episode over)" % i)
        if sl != 0xFFFF:  # shaped reward: the reference's float64 expression (Explorer.py:243-248)
            max_length = 0.5 * (len(self.layout.layout_original) + len(self.layout.layout_original[0]))
            reward = 0.001 * (max_length - sl) / max_length
        else:
            reward = int(r32)
        if proto == L.PROTO_ASTAR:
            # Explorer.execute_action still reports the failed move to its caller
            d = reward == -1 and sl == 0xFFFF
        return reward, d

    def run_game(self, control_pattern="manual", smart_controller=None, render=True):
        """Scene.py:82-96"""
        self.render = render
        self.control_pattern = control_pattern
        if self._fresh:
            self.explorer_group[0].create_explorer()
        if control_pattern == "manual":
            self.run_mode(control_pattern)
        if control_pattern in ["A_star", "auto"]:
            return self.run_mode(control_pattern)
        if control_pattern == "intelligent":
            return self.run_mode(control_pattern, smart_controller)

    def run_mode(self, running_type="manual", smart_controller=None, max_ticks=None):
        """Scene.py:98-175.  `max_ticks` (extension) caps the loop: A_star mode can gridlock
        forever in the reference."""
        self.running_time = 0
        if running_type == "manual" and len(self.explorer_group) > 1:
            print("WARNING: manual mode can only control one AGV")
            return
        self.smart_controller = smart_controller
        expert_info, expert_name, expert_action = [], [], []
        proto = L.PROTO_INTELLIGENT if running_type == "intelligent" else L.PROTO_ASTAR
        while True:
            self.running_time += 1
            self._dev.begin_tick()
            key_action = ""
            if running_type == "manual":
                if self.manual_actions is None:
                    raise L.AgvError("manual mode needs Scene.manual_actions (an iterator of action names); "
                                     "there is no keyboard in the headless build")
                key_action = next(self.manual_actions, None)
                if key_action is None:
                    return
            for i, explorer in enumerate(self.explorer_group):
                st = self._turn_state(i)
                if st == -1:
                    break
                if st == -2:
                    self._mark_over()
                    return expert_info, expert_name, expert_action
                if st == 0:
                    continue
                all_info = []
                input_action = ""
                if running_type in ["A_star", "auto"]:
                    input_action = explorer.find_path_astar(self.explorer_group)
                    expert_info.append(self.create_info())
                    expert_name.append(explorer.explorer_name)
                    expert_action.append(self.dir.action_str_value(input_action))
                if running_type == "manual":
                    input_action = key_action
                if running_type == "intelligent":
                    all_info = self.create_info()
                    input_action = self.smart_controller.choose_action(all_info, explorer.explorer_name)
                    input_action = self.dir.action_value_str(int(input_action))
                if running_type == "intelligent":
                    reward, is_end = explorer.execute_action(input_action, all_info, self.explorer_group)
                    self.smart_controller.store_info(self.create_info(), reward, is_end, explorer.explorer_name)
                    if is_end:
                        return self.running_time
                elif input_action != "":
                    explorer.execute_action(input_action)
            self.check_new_veh()
            if max_ticks is not None and self.running_time >= max_ticks:
                return expert_info, expert_name, expert_action

    def _mark_over(self):
        hdr, agv = self._dev.get_state()
        hdr[0, 5] = 1
        self._dev.set_state(hdr, agv)
        self._pull()
