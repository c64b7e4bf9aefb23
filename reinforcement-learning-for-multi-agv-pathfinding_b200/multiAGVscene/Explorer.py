"""Explorer -- one AGV of a scene: a host-side VIEW of AGV i in the device scene.

Same constructor, attributes and method names as the reference
(src/multiAGVscene/Explorer.py:13-302).  All logic (task FSM, move legality, reward,
valid matrix, "A*") runs in the CUDA kernels behind the owning Scene; the attributes
here are refreshed from the device after every call.
"""
from __future__ import annotations

from ..utils.utils import Direction as Dir
from ..utils.utils import Working
from ..utils.settings import SuperParas as SuperPara
from .. import _lib as L


class Explorer:
    def __init__(self, layout, veh_name="veh1", icon_name="veh1"):
        self.layout = layout
        self.dir = Dir()
        self.working_manager = Working()
        self.explorer_name = veh_name
        self.icon_name = icon_name
        self.target_position = [1, 1]
        # flags are read when the AGV is built, like the reference (Explorer.py:36-37)
        self.always_loaded = SuperPara.Explorer_Always_Loaded
        self.always_empty = SuperPara.Explorer_Always_Empty
        self.action_distribution = [0, 0, 0, 0, 0]
        self.action_str = self.dir.value_str[1]
        self._scene = None  # set by Scene.__init__
        self._index = -1
        self.init()

    def init(self):
        """Explorer.py:45-58 (target_position is deliberately not reset)"""
        self.last_place = [1, 1]
        self.current_place = [1, 1]
        self.task_order = 0
        self.task_stage = 0
        self.running_state = "Start"
        self.loaded = False
        self.has_created = False
        self.all_assigned = False
        self.action_list = []
        self.Working = False
        self.working_type = 0
        self.time_counting = 0

    # ------------------------------------------------------------------ plumbing
    def _attach(self, scene, index):
        self._scene, self._index = scene, index

    def _need_scene(self):
        if self._scene is None:
            raise L.AgvError("Explorer %r is not part of a Scene yet: its state lives on the GPU and is "
                             "created by Scene(layout, explorer_group)" % self.explorer_name)
        return self._scene

    def _refresh(self, row, created):
        """row = (x, y, tx, ty, stage, loaded, idle, order) from agv_get_state"""
        self.current_place = [int(row[0]), int(row[1])]
        self.target_position = [int(row[2]), int(row[3])]
        self.task_stage = int(row[4])
        self.loaded = bool(row[5])
        self.all_assigned = bool(row[6])
        self.task_order = int(row[7])
        self.has_created = bool(created)

    def get_icon(self):
        return "multiAGVscene/icons/%s_%d.png" % (self.explorer_name, self.task_stage + 1)

    @property
    def icon_path(self):
        return self.get_icon()

    # ------------------------------------------------------------------ reference API
    def create_explorer(self):
        """Explorer.py:64-66 -- only the scene can create AGVs (in list order)"""
        self._need_scene()._create_explorer(self._index)

    def action_format(self, input_action):
        value = self.dir.action_str_value(input_action) if isinstance(input_action, str) else input_action
        return value, self.dir.action_value_str(value)

    def action_logical(self, action_value):
        return list(Dir.DELTA[action_value])

    def execute_action(self, input_action, all_info=[], explorer_group=None):
        """Explorer.py:131-149.  `all_info` empty -> no AGV-AGV test (A_star / manual
        protocol), otherwise the intelligent protocol's collision test against the created
        AGVs.  Returns (reward, is_end) with the reference's types."""
        value, self.action_str = self.action_format(input_action)
        self.action_distribution = [0, 0, 0, 0, 0]
        self.action_distribution[value] = 1
        proto = L.PROTO_INTELLIGENT if len(all_info) > 0 else L.PROTO_ASTAR
        reward, is_end = self._need_scene()._execute(self._index, value, proto)
        self.action_list.append(input_action)
        return reward, is_end

    def create_valid_matrix(self, explorer_group=None):
        """Explorer.py:261-284 (matrix a) as nested lists of ints"""
        return self._need_scene()._valid_matrix(self._index, L.MATRIX_ASTAR)

    def find_path_astar(self, explorer_group=None):
        """Explorer.py:286-302 -> action name ("STOP" when no path)"""
        a = self._need_scene()._bfs_action(self._index, L.MATRIX_ASTAR)
        return self.dir.action_value_str(a)
