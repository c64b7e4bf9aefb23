"""B200-native batched multi-AGV (RMFS) environment.

Drop-in for the hot path of stoneMan349/Reinforcement-learning-for-multi-AGV-pathfinding
(env step + state encode + BFS "A*" guidance + replay): hand-written sm_100a CUDA
kernels behind a C ABI (include/agv_b200.h), with Python mirrors of the reference's
Layout / Explorer / Scene / StateManager / astar / PER interfaces on top.
There is no CPU fallback: the product fails loudly when the CUDA library is missing.
"""
from . import _lib
from ._lib import (AgvError, UP, RIGHT, DOWN, LEFT, STOP, NO_TURN, PROTO_ASTAR, PROTO_INTELLIGENT, POLICY_GIVEN,
                   POLICY_ASTAR, POLICY_GUIDED, MATRIX_ASTAR, MATRIX_STATE, OBS_NONE, OBS_CLIP, OBS_FULL)
from .batched import BatchedScene, bfs_batch, bfs_field, max_training_steps
from .replay import DeviceReplay
from .multiAGVscene import Layout, Explorer, Scene
from .algorithm.Manager.StateManager import StateManager
from .algorithm.Manager.ExpertManager import Expert
from .algorithm.MADQN_structure.PER import Memory, SumTree
from .utils.astar import FindPathAstar
from .utils.utils import Direction
from .utils.settings import SuperParas

__all__ = ["BatchedScene", "DeviceReplay", "bfs_batch", "bfs_field", "max_training_steps", "AgvError", "Layout",
           "Explorer", "Scene", "StateManager", "Expert", "Memory", "SumTree", "FindPathAstar", "Direction",
           "SuperParas"]
