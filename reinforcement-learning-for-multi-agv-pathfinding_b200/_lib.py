"""ctypes binding of include/agv_b200.h (the C-ABI drop-in boundary).

There is no CPU fallback: if the CUDA library is missing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGV_B200_LIB", os.path.join(HERE, "libagv_b200.so"))  # override = tuning builds

OK = 0
UP, RIGHT, DOWN, LEFT, STOP, NO_TURN = 0, 1, 2, 3, 4, -1
PROTO_ASTAR, PROTO_INTELLIGENT = 0, 1
POLICY_GIVEN, POLICY_ASTAR, POLICY_GUIDED = 0, 1, 2
MATRIX_ASTAR, MATRIX_STATE = 0, 1
OBS_NONE, OBS_CLIP, OBS_FULL = 0, 1, 2


class AgvConfig(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("E", C.c_int32), ("N", C.c_int32), ("T", C.c_int32),
                ("P", C.c_int32), ("always_loaded", C.c_int32), ("always_empty", C.c_int32),
                ("shaped_reward", C.c_int32), ("auto_reset", C.c_int32), ("max_steps", C.c_int64),
                ("device", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/agv_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
SYMBOLS = {
    "agv_create": (C.c_int, [C.POINTER(AgvConfig), _vp, C.POINTER(_vp)]),
    "agv_destroy": (C.c_int, [_vp]),
    "agv_set_tasks": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "agv_reset": (C.c_int, [_vp, _vp, _vp]),
    "agv_tick": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_tick_host": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_tick_host_join": (C.c_int, [_vp, _vp]),
    "agv_rollout": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_rollout_host": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_begin_tick": (C.c_int, [_vp, _vp]),
    "agv_encode": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "agv_bfs_action": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "agv_step_turn": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_end_tick": (C.c_int, [_vp, _vp]),
    "agv_get_state": (C.c_int, [_vp, _vp, _vp, _vp]),
    "agv_set_state": (C.c_int, [_vp, _vp, _vp, _vp]),
    "agv_read_counters": (C.c_int, [_vp, _vp, _vp]),
    "agv_bfs_batch": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_bfs_field": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "agv_replay_create": (C.c_int, [_i32, _i64, _i32, C.POINTER(_vp)]),
    "agv_replay_destroy": (C.c_int, [_vp]),
    "agv_replay_push": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_replay_sample": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_replay_update": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "agv_replay_gather": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "agv_replay_info": (C.c_int, [_vp, _vp, _vp, _vp]),
    "agv_strerror": (C.c_char_p, [C.c_int]),
    "agv_last_cuda_error": (C.c_int, []),
    "agv_launch_count": (C.c_uint64, []),
    "agv_version": (C.c_char_p, []),
    "agv_last_tick_kernel": (C.c_char_p, []),
    "agv_rollout_profile": (C.c_int, [_vp, _vp, _vp]),
}

_lib = None


class AgvError(RuntimeError):
    pass


def lib():
    """Loads libagv_b200.so (built in-tree by build.py).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AgvError("CUDA extension %s is missing: run `python __graft_entry__.py` (build()) first; "
                           "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)  # AttributeError if the symbol is not exported
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != OK:
        L = lib()
        msg = L.agv_strerror(rc).decode()
        if rc == -2:
            msg += " (cudaError %d)" % L.agv_last_cuda_error()
        raise AgvError("%s failed: %s" % (what or "agv call", msg))
