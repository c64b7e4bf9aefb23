"""FindPathAstar -- the reference's path finder interface on the device BFS.

Reference: src/utils/astar.py:53-153.  The reference's "A*" pops a FIFO open list
(all_cost is never assigned, astar.py:14,86) and re-parents on ties (108-117): its
result is the lexicographically largest shortest path under LEFT < UP < DOWN < RIGHT.
Here the target-rooted distance field comes from the CUDA kernel (agv_bfs_field); the
path is then read off the field by taking, at every cell, the first of RIGHT, DOWN, UP,
LEFT that is one step closer to the target.
"""
from __future__ import annotations

import numpy as np

from ..batched import bfs_field

_PRIO = (("RIGHT", 1, 0), ("DOWN", 0, 1), ("UP", 0, -1), ("LEFT", -1, 0))


class FindPathAstar:
    def __init__(self, world_map, start_pos, target_pos):
        self.world_map = np.array(world_map)
        self.start_pos = (int(start_pos[0]), int(start_pos[1]))
        self.target_pos = (int(target_pos[0]), int(target_pos[1]))
        self.find_target = False
        self.path_list = []
        self.path_map = 0
        self.action_list = []

    def run_astar_method(self):
        """-> (find_target, path_list [target..start], path_map, action_list)"""
        wm = self.world_map
        H, W = wm.shape
        valid = (wm == 1.).astype(np.uint8)
        (sx, sy), (tx, ty) = self.start_pos, self.target_pos
        self.find_target, self.path_list, self.action_list, self.path_map = False, [], [], 0
        if (sx, sy) == (tx, ty):
            self.find_target = True
            path = [(sx, sy)]
        else:
            dist = bfs_field(valid[None], np.array([[tx, ty]], dtype=np.int32))[0].cpu().numpy()
            INF = 0xFFFF

            def d(x, y):
                return int(dist[y, x]) if (0 <= x < W and 0 <= y < H and valid[y, x]) else INF

            best = min(d(sx + dx, sy + dy) for _, dx, dy in _PRIO)
            if best == INF:
                return self.find_target, self.path_list, self.path_map, self.action_list
            self.find_target = True
            path = [(sx, sy)]
            x, y, want = sx, sy, best
            while True:
                for name, dx, dy in _PRIO:
                    if d(x + dx, y + dy) == want:
                        x, y = x + dx, y + dy
                        self.action_list.append(name)
                        path.append((x, y))
                        break
                if want == 0:
                    break
                want -= 1
        self.path_list = path[::-1]
        self.path_map = wm.copy()
        for (x, y) in self.path_list:
            self.path_map[y][x] = -1
        return self.find_target, self.path_list, self.path_map, self.action_list
