"""Layout -- grid geometry, station lists and the task list of one RMFS scene.

Host-side setup with the reference's constructor, attributes and RNG call order
(reference: src/multiAGVscene/Layout.py:15-135); the arrays it produces are what
agv_create / agv_set_tasks upload.  Coordinates are 1-based (x = column, y = row).
"""
from __future__ import annotations

import random
from collections import namedtuple

import numpy as np

StorageStation = namedtuple("ss", "x_position y_position")
PickingStation = namedtuple("ps", "x_position y_position")

TRACK, STORAGE, PICKING = 0, 1, 2


class Layout:
    def __init__(self, storage_station_x_width=3, storage_station_y_width=2, storage_station_x_num=2,
                 storage_station_y_num=2, picking_station_number=2, layout_list=None, task_list=None):
        if layout_list is None:
            self.storage_station_x_width = storage_station_x_width
            self.storage_station_y_width = storage_station_y_width
            self.storage_station_interval = 1
            self.storage_station_bottom = 3
            self.storage_station_x_num = storage_station_x_num
            self.storage_station_y_num = storage_station_y_num
            self.picking_station_number = picking_station_number
            codes = self._rect_codes()
        else:
            codes = np.asarray(layout_list)
            # only 1 and 2 are stations; every other code is track (Layout.py:66-81,118-132)
            codes = np.where(codes == STORAGE, STORAGE, np.where(codes == PICKING, PICKING, TRACK))
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.scene_y_width, self.scene_x_width = (int(v) for v in self.codes.shape)
        # row-major enumeration == the reference's list order for both constructors
        self.storage_station_list = [StorageStation(int(x) + 1, int(y) + 1)
                                     for y, x in zip(*np.nonzero(self.codes == STORAGE))]
        self.picking_station_list = [PickingStation(int(x) + 1, int(y) + 1)
                                     for y, x in zip(*np.nonzero(self.codes == PICKING))]
        if layout_list is None:
            self.picking_station_list = self._rect_picking()
        self.task_list = task_list
        self.task_is_fixed = task_list is not None
        self.task_arrangement = [[], [], []]
        self.task_finished = False
        self.layout, self.layout_original = [], []
        self.init()

    # rect generator: W = nx*(w+1)+1, H = ny*(h+1)+1+3 (Layout.py:29-30)
    def _rect_codes(self):
        w, h = self.storage_station_x_width, self.storage_station_y_width
        nx, ny = self.storage_station_x_num, self.storage_station_y_num
        gap, bottom = self.storage_station_interval, self.storage_station_bottom
        W = nx * (w + gap) + gap
        H = ny * (h + gap) + gap + bottom
        self.scene_x_width, self.scene_y_width = W, H
        g = np.zeros((H, W), dtype=np.uint8)
        xs = (np.arange(nx)[:, None] * (w + gap) + np.arange(w)[None, :] + gap).ravel()
        ys = (np.arange(ny)[:, None] * (h + gap) + np.arange(h)[None, :] + gap).ravel()
        g[np.ix_(ys, xs)] = STORAGE
        for p in self._rect_picking():
            g[p.y_position - 1, p.x_position - 1] = PICKING  # picking overrides (Layout.py:125-128)
        return g

    def _rect_picking(self):
        # x = int(W*(i+1)/(ps+1)+1), y = H-1 (Layout.py:83-90); duplicates are kept
        n = self.picking_station_number
        return [PickingStation(int(self.scene_x_width * (i + 1) / (n + 1) + 1), self.scene_y_width - 1)
                for i in range(n)]

    def init(self):
        """new random task list (unless fixed) and a fresh layout matrix (Layout.py:56-64)"""
        if not self.task_is_fixed:
            self.task_list = self._draw_tasks()
        self.task_finished = False
        self.task_arrangement = [[], [], []]
        self.layout = [[int(v) for v in row] for row in self.codes]
        self.layout_original = [list(row) for row in self.layout]

    def _draw_tasks(self):
        """every storage station exactly once in random order, each paired with a random
        picking station; consumes `random` exactly like Layout.py:104-116 (randint for the
        storage index, then randint for the picking index, per task)."""
        remaining = list(self.storage_station_list)
        picking = self.picking_station_list
        tasks = []
        while remaining:
            i = random.randint(0, len(remaining) - 1)
            j = random.randint(0, len(picking) - 1)
            tasks.append(tuple(remaining.pop(i)) + tuple(picking[j]))
        return tasks

    def change_layout(self, x_dim, y_dim, value):
        self.layout[y_dim][x_dim] = value

    # ---- arrays for the device
    def task_array(self):
        return np.asarray(self.task_list, dtype=np.uint8).reshape(-1, 4)
