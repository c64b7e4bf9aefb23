import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "reinforcement-learning-for-multi-agv-pathfinding_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "live: needs the live reference checkout (build container only)")


def load_pkg():
    return importlib.import_module(PKG_NAME)


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(params=["async", "lanes", "warp", "rollout"])
def tick_impl(request, monkeypatch):
    """Runs a fused-tick test once per implementation of agv_tick: decoupled lanes
    (agv_async.cuh, the default), lock-step lanes (agv_lanes.cuh), warp-per-scene (k_tick) and the
    persistent rollout kernel with n_ticks = 1 (agv_rollout.cuh).
    The library reads AGV_TICK_IMPL at every call."""
    monkeypatch.setenv("AGV_TICK_IMPL", request.param)
    if request.param == "rollout":
        monkeypatch.setenv("AGV_ROLLOUT_IMPL", "persistent")  # also on the small maps agv_rollout would not take
    return request.param
