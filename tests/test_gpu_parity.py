"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and
the committed live-reference fixtures.  Bit-exact for every integer quantity;
rewards are compared as float32(reference float64)."""
import random

import numpy as np
import pytest

from conftest import golden, load_pkg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import torch
    assert torch.cuda.is_available()
    return load_pkg()


def _cmp_tick(pkg, sc, envs, proto, policy, obs_kind, P, given=None, want_next=False, tag=""):
    out = sc.tick(proto, policy, obs_kind, actions=given, want_next=want_next)
    act = out["act"].cpu().numpy()
    rew = out["reward"].cpu().numpy()
    done = out["done"].cpu().numpy()
    plen = out["plen"].cpu().numpy().astype(np.int32) & 0xFFFF
    slen = out["slen"].cpu().numpy().astype(np.int32) & 0xFFFF
    obs = out["obs"].float().cpu().numpy() if obs_kind else None
    nxt = out["next_obs"].float().cpu().numpy() if (obs_kind and want_next) else None
    hdr, agv = sc.get_state()
    turns = 0
    for e, env in enumerate(envs):
        g = None if given is None else np.asarray(given)[e]
        n, a, r, d, pl, sl, ob, nx = env.tick_obs(proto, policy, given=g, obs_kind=obs_kind, P=P,
                                                  want_next=want_next)
        turns += n
        assert np.array_equal(act[e], a), (tag, e, act[e], a)
        assert np.array_equal(rew[e], r.astype(np.float32)), (tag, e, rew[e], r)
        assert np.array_equal(done[e], d), (tag, e)
        assert np.array_equal(plen[e], pl), (tag, e, plen[e], pl)
        assert np.array_equal(slen[e], sl), (tag, e, slen[e], sl)
        took = a >= 0  # observation records are only written for AGVs that took a turn
        if obs is not None:
            assert np.array_equal(obs[e][took], ob[took]), (tag, e)
        if nxt is not None:
            assert np.array_equal(nxt[e][took], nx[took]), (tag, e)
        oh, oa = env.state()
        assert np.array_equal(hdr[e], oh), (tag, e, hdr[e], oh)
        assert np.array_equal(agv[e], oa), (tag, e)
    return turns


def test_bfs_fixtures(pkg):
    """agv_bfs_batch against FindPathAstar outputs recorded from the live reference"""
    cases = golden("bfs.json")
    by_shape = {}
    for c in cases:
        v = np.array(c["valid"], dtype=np.uint8)
        by_shape.setdefault(v.shape, []).append((v, c))
    n = 0
    for shape, lst in by_shape.items():
        valid = np.stack([v for v, _ in lst])
        start = np.array([c["start"] for _, c in lst], dtype=np.int32)
        target = np.array([c["target"] for _, c in lst], dtype=np.int32)
        act, ln, fnd = pkg.bfs_batch(valid, start, target)
        act, ln, fnd = act.cpu().numpy(), ln.cpu().numpy(), fnd.cpu().numpy()
        for k, (_, c) in enumerate(lst):
            first = c["actions"][0] if c["actions"] else 4
            assert act[k] == first and ln[k] == len(c["actions"]) and bool(fnd[k]) == c["found"], (shape, k)
            n += 1
    assert n == len(cases)


def test_bfs_fixtures_at_baseline_sizes(pkg):
    """agv_bfs_batch against the live reference's FindPathAstar at 64x64 / 128x128
    (tests/golden/bfs_large.json)"""
    cases = golden("bfs_large.json")
    by_shape = {}
    for c in cases:
        v = np.array([[int(ch) for ch in row] for row in c["valid"]], dtype=np.uint8)
        by_shape.setdefault(v.shape, []).append((v, c))
    assert set(by_shape) == {(64, 64), (128, 128)}
    for shape, lst in by_shape.items():
        valid = np.stack([v for v, _ in lst])
        start = np.array([c["start"] for _, c in lst], dtype=np.int32)
        target = np.array([c["target"] for _, c in lst], dtype=np.int32)
        act, ln, fnd = pkg.bfs_batch(valid, start, target)
        act, ln, fnd = act.cpu().numpy(), ln.cpu().numpy(), fnd.cpu().numpy()
        dist = pkg.bfs_field(valid, target).cpu().numpy()
        for k, (_, c) in enumerate(lst):
            first = c["actions"][0] if c["actions"] else 4
            assert act[k] == first and ln[k] == len(c["actions"]) and bool(fnd[k]) == c["found"], c["name"]
            # the distance field decreases by one along the reference's whole action list
            x, y = c["start"]
            for j, a in enumerate(c["actions"]):
                dx, dy = [(0, -1), (1, 0), (0, 1), (-1, 0)][a]
                x, y = x + dx, y + dy
                assert dist[k, y, x] == len(c["actions"]) - 1 - j, (c["name"], j)


@pytest.mark.parametrize("H,W", [(9, 9), (32, 32), (33, 64), (64, 64), (64, 65), (20, 128), (65, 40), (128, 128),
                                  (1, 7), (7, 1), (128, 3)])
def test_bfs_random_vs_oracle(pkg, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    B = 96
    valid = (rng.random((B, H, W)) > rng.choice([0.0, 0.1, 0.25, 0.35], size=(B, 1, 1))).astype(np.uint8)
    start = np.stack([rng.integers(0, W, B), rng.integers(0, H, B)], axis=1).astype(np.int32)
    target = np.stack([rng.integers(0, W, B), rng.integers(0, H, B)], axis=1).astype(np.int32)
    target[:3] = start[:3]
    act, ln, fnd = pkg.bfs_batch(valid, start, target)
    act, ln, fnd = act.cpu().numpy(), ln.cpu().numpy(), fnd.cpu().numpy()
    dist = pkg.bfs_field(valid, target).cpu().numpy()
    for b in range(B):
        a, f, l = orc.bfs_field(valid[b], tuple(start[b]), tuple(target[b]))
        assert (act[b], bool(fnd[b]), ln[b]) == (a, f, l), (b, start[b], target[b])
        # distance field: 0 at a valid target, monotone along the chosen action
        tx, ty = target[b]
        if valid[b, ty, tx]:
            assert dist[b, ty, tx] == 0
            if l > 0:
                sx, sy = start[b]
                dx, dy = [(0, -1), (1, 0), (0, 1), (-1, 0)][a]
                assert dist[b, sy + dy, sx + dx] == l - 1
        else:
            assert (dist[b] == 0xFFFF).all()


@pytest.mark.usefixtures("tick_impl")
@pytest.mark.parametrize("idx", range(7))
def test_astar_trajectories_fixture(pkg, idx):
    """Scene.run_game('A_star') trajectories of the live reference"""
    c = golden("astar_traj.json")[idx]
    grid = np.array(c["grid"], dtype=np.uint8)
    tasks = np.array(c["tasks"], dtype=np.uint8)[None]
    sc = pkg.BatchedScene(grid, 1, c["n_agv"], tasks=tasks)
    sc.reset()
    veh, actions = [], []
    for t in range(c["ticks"] + 2):
        out = sc.tick(pkg.PROTO_ASTAR, pkg.POLICY_ASTAR)
        a = out["act"].cpu().numpy()[0]
        for i in range(c["n_agv"]):
            if a[i] >= 0:
                veh.append(i)
                actions.append(int(a[i]))
        hdr, agv = sc.get_state()
        if hdr[0, 5] or (not c["finished"] and hdr[0, 0] >= c["ticks"]):
            break
    assert veh == c["veh"] and actions == c["actions"]
    assert int(hdr[0, 0]) == c["ticks"]
    snap = [[int(v) for v in agv[0, j, [0, 1, 2, 3, 5]]] for j in range(int(hdr[0, 4]))]
    assert snap == c["final"]


def test_intelligent_fixture_substep(pkg):
    """the callback protocol (encode -> guidance BFS -> step_turn -> encode) turn by turn
    against the live-reference recordings: obs, guided action, reward, is_end, snapshots"""
    eps = golden("intelligent.json")
    n_turns = n_obs = 0
    for ep in eps:
        flags = ep["flags"]
        grid = np.array(ep["grid"], dtype=np.uint8)
        N = ep["n_agv"]
        sc = pkg.BatchedScene(grid, 1, N, tasks=np.array(ep["tasks"], dtype=np.uint8)[None], P=3,
                              always_loaded=flags.get("Explorer_Always_Loaded", False),
                              always_empty=flags.get("Explorer_Always_Empty", False),
                              shaped_reward=flags.get("Sparse_Reward", False), max_steps=ep["max_steps"])
        sc.reset()
        turns = ep["turns"]
        k = 0
        M = 0.5 * (grid.shape[0] + grid.shape[1])
        while k < len(turns):
            sc.begin_tick()
            ended = False
            for i in range(N):
                obs, will = sc.encode(i, pkg.OBS_CLIP)
                if not int(will.cpu()[0]):
                    continue
                t = turns[k]
                assert t["i"] == i
                ga, gl = sc.bfs_action(i, pkg.MATRIX_STATE)
                assert int(ga.cpu()[0]) == t["guided"] and (int(gl.cpu()[0]) & 0xFFFF) == t["glen"]
                if "obs_clip" in t:
                    assert obs.float().cpu().numpy()[0].astype(int).tolist() == t["obs_clip"]
                    of, _ = sc.encode(i, pkg.OBS_FULL)
                    assert of.float().cpu().numpy()[0].astype(int).tolist() == t["obs_full"]
                    n_obs += 1
                rew, done, acted, slen = sc.step_turn(i, pkg.PROTO_INTELLIGENT, np.array([t["a"]], dtype=np.int8))
                assert int(acted.cpu()[0]) == 1
                r = float(rew.cpu()[0])
                sl = int(slen.cpu()[0]) & 0xFFFF
                if sl != 0xFFFF:  # shaped reward: rebuild the float64 value exactly like Explorer.py:243-248
                    assert 0.001 * (M - sl) / M == t["r"]
                    assert r == np.float32(t["r"])
                else:
                    assert r == t["r"]
                assert int(done.cpu()[0]) == t["end"]
                hdr, agv = sc.get_state()
                snap = [[int(v) for v in agv[0, j, [0, 1, 2, 3, 5]]] for j in range(int(hdr[0, 4]))]
                assert snap == t["after"], (ep["ep"], k)
                k += 1
                n_turns += 1
                if t["end"]:
                    ended = True
                    break
            if ended:
                break
            sc.end_tick()
        assert k == len(turns)
        if ep["running_time"] >= 0:
            assert int(sc.get_state()[0][0, 0]) == ep["running_time"]
        sc.close()
    assert n_turns > 6000 and n_obs > 50


@pytest.mark.usefixtures("tick_impl")
def test_intelligent_fixture_fused_given(pkg):
    """same recordings through the fused tick with the recorded actions"""
    eps = golden("intelligent.json")
    for ep in eps[::3]:
        flags = ep["flags"]
        grid = np.array(ep["grid"], dtype=np.uint8)
        N = ep["n_agv"]
        kw = dict(always_loaded=flags.get("Explorer_Always_Loaded", False),
                  always_empty=flags.get("Explorer_Always_Empty", False),
                  shaped=flags.get("Sparse_Reward", False), max_steps=ep["max_steps"])
        sc = pkg.BatchedScene(grid, 1, N, tasks=np.array(ep["tasks"], dtype=np.uint8)[None], P=3,
                              always_loaded=kw["always_loaded"], always_empty=kw["always_empty"],
                              shaped_reward=kw["shaped"], max_steps=ep["max_steps"])
        sc.reset()
        turns = ep["turns"]
        k = 0
        while k < len(turns):
            given = np.full((1, N), 4, dtype=np.int8)
            kk, last = k, -1
            while kk < len(turns) and turns[kk]["i"] > last:
                given[0, turns[kk]["i"]] = turns[kk]["a"]
                last = turns[kk]["i"]
                kk += 1
            out = sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GIVEN, actions=given)
            act = out["act"].cpu().numpy()[0]
            rew = out["reward"].cpu().numpy()[0]
            done = out["done"].cpu().numpy()[0]
            for i in range(N):
                if act[i] >= 0:
                    t = turns[k]
                    assert (t["i"], t["a"]) == (i, act[i])
                    assert rew[i] == np.float32(t["r"]) and done[i] == t["end"]
                    k += 1
            if sc.get_state()[0][0, 5]:
                break
        assert k == len(turns)
        sc.close()


CASES = [
    # name, grid, N, E, ticks, proto, policy, obs_kind, flags
    ("c1_guided_clip", lambda: orc.rect_layout(4, 2, 2, 2, 2), 4, 16, 120, 1, 2, 1, {}),
    ("c2_guided_full", lambda: orc.README_9x9, 4, 16, 100, 1, 2, 2, {}),
    ("c1_astar", lambda: orc.rect_layout(4, 2, 2, 2, 2), 6, 8, 150, 0, 1, 1, {}),
    ("c3_guided_clip", lambda: orc.rect_layout(6, 2, 9, 20, 8), 64, 6, 140, 1, 2, 1, {}),
    ("c3_astar_full", lambda: orc.rect_layout(6, 2, 9, 20, 8), 40, 3, 60, 0, 1, 2, {}),
    ("c3_shaped", lambda: orc.rect_layout(6, 2, 9, 20, 8), 24, 4, 60, 1, 2, 1, {"shaped": True}),
    ("wide_128", lambda: orc.rect_layout(6, 2, 18, 41, 9), 48, 3, 100, 1, 2, 1, {}),
    ("tall_40x128", lambda: orc.rect_layout(2, 2, 13, 41, 3), 33, 3, 80, 1, 2, 1, {}),
    ("always_loaded", lambda: orc.rect_layout(3, 2, 3, 3, 2), 5, 8, 80, 1, 2, 1, {"always_loaded": True}),
    ("always_empty", lambda: orc.rect_layout(3, 2, 3, 3, 2), 5, 8, 80, 1, 2, 2, {"always_empty": True}),
    ("single_agv", lambda: orc.rect_layout(3, 2, 2, 2, 2), 1, 8, 200, 1, 2, 1, {}),
    ("more_agv_than_tasks", lambda: orc.rect_layout(1, 1, 2, 1, 1), 6, 8, 80, 0, 1, 1, {}),
]


@pytest.mark.usefixtures("tick_impl")
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_tick_vs_oracle(pkg, case):
    name, gridf, N, E, ticks, proto, policy, obs_kind, flags = case
    grid = gridf()
    assert grid.shape[0] <= 128 and grid.shape[1] <= 128
    tasks = np.array([orc.make_tasks(grid, random.Random(100 + e)) for e in range(E)], dtype=np.uint8)
    sc = pkg.BatchedScene(grid, E, N, tasks=tasks, P=3, always_loaded=flags.get("always_loaded", False),
                          always_empty=flags.get("always_empty", False), shaped_reward=flags.get("shaped", False))
    sc.reset()
    envs = [orc.Env(grid, tasks[e], N, always_loaded=flags.get("always_loaded", False),
                    always_empty=flags.get("always_empty", False), shaped=flags.get("shaped", False))
            for e in range(E)]
    turns = 0
    for t in range(ticks):
        # restart finished / failed scenes so that the run keeps exercising turns
        hdr, _ = sc.get_state()
        over = hdr[:, 5].astype(bool)
        if over.any():
            sc.reset(over.astype(np.uint8))
            for e in np.nonzero(over)[0]:
                envs[e].reset()
        turns += _cmp_tick(pkg, sc, envs, proto, policy, obs_kind, 3, want_next=(t % 3 == 0), tag=(name, t))
    assert turns > ticks
    c_turns, c_eps, c_bad = sc.counters()
    assert c_turns == turns and c_bad == 0
    sc.close()


@pytest.mark.usefixtures("tick_impl")
def test_fused_given_random_actions_vs_oracle(pkg):
    """random (mostly illegal) open-loop actions incl. the guided fallback (-1) and STOP"""
    grid = orc.rect_layout(4, 2, 2, 2, 2)
    E, N = 32, 4
    rng = np.random.default_rng(3)
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    for proto in (0, 1):
        sc = pkg.BatchedScene(grid, E, N, tasks=tasks, P=2)
        sc.reset()
        envs = [orc.Env(grid, tasks[e], N) for e in range(E)]
        for t in range(150):
            hdr, _ = sc.get_state()
            over = hdr[:, 5].astype(bool)
            if over.any():
                sc.reset(over.astype(np.uint8))
                for e in np.nonzero(over)[0]:
                    envs[e].reset()
            given = rng.choice(np.array([-1, -1, -1, -1, -1, -1, 0, 1, 2, 3, 4], dtype=np.int8), size=(E, N))
            _cmp_tick(pkg, sc, envs, proto, 0, 1, 2, given=given, want_next=True, tag=("given", proto, t))
        sc.close()


@pytest.mark.usefixtures("tick_impl")
def test_auto_reset_and_substep_equals_fused(pkg):
    """auto_reset=1: sub-step protocol (begin/encode/bfs/step/end) == fused guided tick"""
    grid = orc.README_9x9
    E, N = 24, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(7 * e)) for e in range(E)], dtype=np.uint8)
    a = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
    b = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
    a.reset()
    b.reset()
    import torch
    for t in range(120):
        out = a.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP, want_next=True)
        b.begin_tick()
        for k in range(N):
            obs, will = b.encode(k, pkg.OBS_CLIP)
            act, plen = b.bfs_action(k, pkg.MATRIX_STATE)
            took = out["act"][:, k] >= 0
            assert torch.equal(will.bool(), took)
            assert torch.equal(obs[took], out["obs"][:, k][took])
            assert not obs[~took].any()  # the sub-step encoder zero-fills scenes without a turn
            assert torch.equal(act, out["act"][:, k])
            rew, done, acted, slen = b.step_turn(k, pkg.PROTO_INTELLIGENT, act.clamp(min=0))
            assert torch.equal(acted.bool(), out["act"][:, k] >= 0)
            assert torch.equal(rew, out["reward"][:, k]) and torch.equal(done, out["done"][:, k])
            nobs, did = b.encode(k, pkg.OBS_CLIP, post=True)
            assert torch.equal(nobs[took], out["next_obs"][:, k][took])
        b.end_tick()
        ha, ga = a.get_state()
        hb, gb = b.get_state()
        assert np.array_equal(ha, hb) and np.array_equal(ga, gb)
    assert a.counters()[1] > 0  # some episodes ended and were auto-reset
    a.close()
    b.close()


def test_invalid_arguments(pkg):
    import ctypes as C
    L = pkg._lib
    lib = L.lib()
    grid = np.zeros((4, 4), dtype=np.uint8)
    h = C.c_void_p()
    cfg = L.AgvConfig(H=4, W=4, E=1, N=65, T=0, P=3, max_steps=10, device=0)
    assert lib.agv_create(C.byref(cfg), grid.ctypes.data_as(C.c_void_p), C.byref(h)) == -1
    cfg = L.AgvConfig(H=200, W=4, E=1, N=1, T=0, P=3, max_steps=10, device=0)
    assert lib.agv_create(C.byref(cfg), grid.ctypes.data_as(C.c_void_p), C.byref(h)) == -4
    bad = grid.copy()
    bad[1, 1] = 3
    cfg = L.AgvConfig(H=4, W=4, E=1, N=1, T=0, P=3, max_steps=10, device=0)
    assert lib.agv_create(C.byref(cfg), bad.ctypes.data_as(C.c_void_p), C.byref(h)) == -5
    # invalid action codes are data: counted and treated as STOP, never a crash
    sc = pkg.BatchedScene(orc.rect_layout(2, 1, 1, 1, 1), 2, 1, tasks=np.array([[[2, 2, 2, 4], [3, 2, 2, 4]]] * 2))
    sc.reset()
    sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GIVEN, actions=np.full((2, 1), 9, dtype=np.int8))
    assert sc.counters()[2] == 2
    with pytest.raises(L.AgvError):
        sc.set_tasks(np.zeros((2, 2, 4), dtype=np.uint8))  # coordinates out of range
    sc.close()


@pytest.mark.usefixtures("tick_impl")
@pytest.mark.parametrize("P", [1, 2, 5])
def test_clip_window_sizes(pkg, P):
    """limited visual with padding sizes other than the default 3 (K = 3, 5, 11)"""
    grid = orc.rect_layout(3, 2, 3, 3, 2)
    E, N = 6, 5
    tasks = np.array([orc.make_tasks(grid, random.Random(50 + e)) for e in range(E)], dtype=np.uint8)
    sc = pkg.BatchedScene(grid, E, N, tasks=tasks, P=P)
    sc.reset()
    envs = [orc.Env(grid, tasks[e], N) for e in range(E)]
    for t in range(60):
        hdr, _ = sc.get_state()
        over = hdr[:, 5].astype(bool)
        if over.any():
            sc.reset(over.astype(np.uint8))
            for e in np.nonzero(over)[0]:
                envs[e].reset()
        _cmp_tick(pkg, sc, envs, 1, 2, 1, P, want_next=True, tag=("P", P, t))
    sc.close()


@pytest.mark.usefixtures("tick_impl")
def test_edge_scenes(pkg):
    """no tasks at all, one task for many AGVs, a 1-row corridor, an early max_steps ending"""
    # T = 0: AGV 0 is created idle, task_finished is raised at creation, the next tick returns
    grid = orc.rect_layout(2, 1, 1, 1, 1)
    sc = pkg.BatchedScene(grid, 2, 3, tasks=None)
    sc.reset()
    envs = [orc.Env(grid, np.zeros((0, 4), dtype=np.int32), 3) for _ in range(2)]
    for t in range(3):
        _cmp_tick(pkg, sc, envs, 1, 2, 0, 3, tag=("T0", t))
    assert sc.get_state()[0][:, 5].all()
    sc.close()
    # corridor 1 x 12 with a storage cell and a picking station, 3 AGVs (they block each other)
    grid = np.zeros((1, 12), dtype=np.uint8)
    grid[0, 6] = 1
    grid[0, 11] = 2
    tasks = np.array([[[7, 1, 12, 1]]] * 4, dtype=np.uint8)
    for proto, policy in ((0, 1), (1, 2)):
        sc = pkg.BatchedScene(grid, 4, 3, tasks=tasks)
        sc.reset()
        envs = [orc.Env(grid, tasks[e], 3) for e in range(4)]
        for t in range(40):
            _cmp_tick(pkg, sc, envs, proto, policy, 2, 3, want_next=True, tag=("corridor", proto, t))
        sc.close()
    # max_steps = 5: every episode ends by the step cap (Scene.py:155)
    grid = orc.rect_layout(2, 1, 2, 1, 1)
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(4)], dtype=np.uint8)
    sc = pkg.BatchedScene(grid, 4, 2, tasks=tasks, max_steps=5)
    sc.reset()
    envs = [orc.Env(grid, tasks[e], 2, max_steps=5) for e in range(4)]
    for t in range(7):
        _cmp_tick(pkg, sc, envs, 1, 2, 1, 3, tag=("cap", t))
    hdr, _ = sc.get_state()
    assert hdr[:, 5].all() and (hdr[:, 0] <= 5).all()
    sc.close()


@pytest.mark.usefixtures("tick_impl")
def test_injected_overlapping_state(pkg):
    """agv_set_state with two created AGVs on one cell (reachable in the reference through the
    A_star protocol with hand-fed actions): the slow-path ballots must agree with the oracle"""
    grid = orc.rect_layout(4, 2, 2, 2, 2)
    E, N = 4, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    rng = np.random.default_rng(11)
    for proto in (0, 1):
        sc = pkg.BatchedScene(grid, E, N, tasks=tasks)
        sc.reset()
        envs = [orc.Env(grid, tasks[e], N) for e in range(E)]
        # A_star protocol, hand-fed actions: AGV 0 steps RIGHT to (2,1) and waits, AGV 1 spawns at
        # (1,1) and follows onto the same cell (no AGV-AGV test in this protocol), then both wander
        script = [[1, 4, 4, 4], [4, 1, 4, 4]] + [list(rng.integers(0, 5, size=N)) for _ in range(6)]
        for t, row in enumerate(script):
            given = np.tile(np.array(row, dtype=np.int8), (E, 1))
            _cmp_tick(pkg, sc, envs, 0, 0, 1, 3, given=given, tag=("ov-setup", t))
            if t == 1:
                hdr, agv = sc.get_state()
                assert (agv[:, 0, :2] == agv[:, 1, :2]).all() and (hdr[:, 4] >= 2).all()
        hdr, agv = sc.get_state()
        sc.set_state(hdr, agv)  # round trip through agv_set_state keeps the (overlapping) state
        # now continue with the guided policy under either protocol from the overlapping state
        for t in range(25):
            _cmp_tick(pkg, sc, envs, proto, 2, 1, 3, want_next=True, tag=("ov", proto, t))
        sc.close()


def test_c_abi_from_plain_c(pkg, tmp_path):
    """the C ABI driven by a plain C program (no Python / torch in the loop)"""
    import os
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "abi_smoke")
    pkg_dir = os.path.dirname(pkg._lib.LIB_PATH)
    cuda = "/usr/local/cuda"
    subprocess.check_call(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + cuda + "/include", "-L" + pkg_dir,
                           "-L" + cuda + "/lib64", "-lagv_b200", "-lcudart", "-Wl,-rpath," + pkg_dir,
                           "-Wl,-rpath," + cuda + "/lib64"])
    grid = orc.rect_layout(4, 2, 2, 2, 2)
    E, N, ticks = 2, 4, 50
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    T = tasks.shape[1]
    inp = "%d %d %d %d %d %d\n" % (grid.shape[0], grid.shape[1], E, N, T, ticks)
    inp += " ".join(str(int(v)) for v in grid.ravel()) + "\n" + " ".join(str(int(v)) for v in tasks.ravel()) + "\n"
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.strip().splitlines()
    envs = [orc.Env(grid, tasks[e], N) for e in range(E)]
    turns = 0
    for t in range(ticks):
        exp = []
        for env in envs:
            n, a, r, d, pl, sl, _, _ = env.tick_obs(orc.PROTO_INTELLIGENT, orc.POLICY_GUIDED)
            turns += n
            exp += ["%d:%g" % (int(a[i]), float(np.float32(r[i]))) for i in range(N)]
        assert out[t].split() == exp, t
    assert out[-1].startswith("turns %d " % turns)


@pytest.mark.usefixtures("tick_impl")
def test_output_buffers_are_not_overrun(pkg):
    """compute-sanitizer is closed on this pool, so out-of-bounds stores are hunted with guard
    bands: every output tensor of agv_tick / agv_encode is a window into a larger buffer filled
    with a sentinel; the bands around the window must be untouched (odd E, N and unaligned
    record sizes on purpose)."""
    import ctypes as C
    import torch
    L = pkg._lib
    for (gridf, E, N, P, kind) in ((lambda: orc.rect_layout(4, 2, 2, 2, 2), 7, 3, 3, 1),
                                   (lambda: orc.README_9x9, 5, 5, 2, 2),
                                   (lambda: orc.rect_layout(6, 2, 9, 20, 8), 3, 37, 3, 1),
                                   (lambda: orc.rect_layout(6, 2, 9, 20, 8), 2, 9, 3, 2)):
        grid = gridf()
        tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
        sc = pkg.BatchedScene(grid, E, N, tasks=tasks, P=P, auto_reset=True)
        sc.reset()
        G = 4096  # guard elements on each side
        R = int(np.prod(sc.obs_shape(kind)))

        def guarded(n, dtype, fill):
            buf = torch.full((n + 2 * G,), fill, dtype=dtype, device=sc.device)
            return buf, buf[G:G + n]

        bufs = {k: guarded(n, dt, f) for k, (n, dt, f) in {
            "act": (E * N, torch.int8, 77), "reward": (E * N, torch.float32, 1234.5),
            "done": (E * N, torch.uint8, 201), "plen": (E * N, torch.int16, 12345), "slen": (E * N, torch.int16, 12345),
            "obs": (E * N * R, torch.bfloat16, 7.0), "next_obs": (E * N * R, torch.bfloat16, 7.0)}.items()}
        out = {k: v[1].view((E, N) + (sc.obs_shape(kind) if "obs" in k else ())) for k, v in bufs.items()}
        for t in range(2 * N + 6):
            sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, kind, want_next=True, out=out)
        # sub-step encoders into a guarded [E, R] window
        eb, ew = guarded(E * R, torch.bfloat16, 7.0)
        wb, ww = guarded(E, torch.uint8, 201)
        st = C.c_void_p(torch.cuda.current_stream(sc.device).cuda_stream)
        sc.begin_tick()
        for k in range(N):
            L.check(sc.lib.agv_encode(sc._h, k, kind, 0, C.c_void_p(ew.data_ptr()), C.c_void_p(ww.data_ptr()), st))
        torch.cuda.synchronize()
        for name, (buf, win) in list(bufs.items()) + [("enc", (eb, ew)), ("will", (wb, ww))]:
            fill = buf[0].item()
            assert (buf[:G] == fill).all() and (buf[-G:] == fill).all(), (name, E, N, kind)
        assert set(out["obs"].float().unique().tolist()) <= {0.0, 1.0, 7.0}
        sc.close()


def test_tick_host_equals_tick(pkg):
    """agv_tick_host (pinned host buffers, overlapped copy-back) returns what agv_tick returns"""
    import torch
    grid = orc.rect_layout(4, 2, 2, 2, 2)
    E, N = 64, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(100 + s)) for s in range(E)], dtype=np.uint8)
    a = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
    b = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
    a.reset(); b.reset()
    rng = np.random.default_rng(5)
    hist = []
    for t in range(40):
        acts = rng.integers(-1, 5, size=(E, N)).astype(np.int8)
        o = a.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GIVEN, pkg.OBS_CLIP, actions=acts)
        ah = torch.from_numpy(acts).pin_memory()
        # join=False on odd steps: the result is only read after the explicit join below
        oh = b.tick_host(pkg.PROTO_INTELLIGENT, pkg.POLICY_GIVEN, pkg.OBS_CLIP, actions_host=ah, join=(t % 2 == 0))
        b.host_join()
        torch.cuda.synchronize()
        assert np.array_equal(o["act"].cpu().numpy(), oh["act"].numpy()), t
        assert np.array_equal(o["reward"].cpu().numpy(), oh["reward"].numpy()), t
        assert np.array_equal(o["done"].cpu().numpy(), oh["done"].numpy()), t
        took = o["act"].cpu().numpy() >= 0
        assert np.array_equal(o["obs"].float().cpu().numpy()[took], oh["obs"].float().cpu().numpy()[took]), t
        hist.append(int(took.sum()))
    assert sum(hist) > 0
    ha, aa = a.get_state(); hb, ab = b.get_state()
    assert np.array_equal(ha, hb) and np.array_equal(aa, ab)
    a.close(); b.close()


def _random_layout(rng, H, W):
    """random track / storage / picking codes with an aisle structure (every storage cell touches a
    track), row 0 and column 0 kept as track so that the spawn cell (1,1) is free"""
    g = np.zeros((H, W), dtype=np.uint8)
    for r in range(2, H - 1):
        for cc in range(2, W - 1):
            if (r % 3 != 0) and (cc % (2 + int(rng.integers(2, 6)))) != 0 and rng.random() < 0.8:
                g[r, cc] = 1
    for cc in rng.choice(np.arange(2, W - 1), size=min(4, W - 3), replace=False):
        g[H - 1, cc] = 2
    return g


@pytest.mark.parametrize("seed", range(12))
def test_cross_implementation_fuzz(pkg, seed, monkeypatch):
    """random map sizes (so that every (NW, RPL) kernel variant, scene counts that do not fill a warp and
    AGV counts from 1 to 64 occur), random open-loop / guided actions in both protocols, with and without
    next_obs: the four fused-tick implementations must produce identical outputs and scene states"""
    import torch
    rng = np.random.default_rng(1000 + seed)
    H, W = [(40, 40), (64, 64), (33, 100), (100, 40), (128, 128), (65, 65), (34, 34), (48, 70), (64, 30), (90, 90),
            (36, 128), (128, 36)][seed]
    grid = _random_layout(rng, H, W)
    stor = np.argwhere(grid == 1); pick = np.argwhere(grid == 2)
    E = int(rng.integers(1, 70)); N = int(rng.choice([1, 2, 7, 33, 64])); T = int(rng.integers(3, 40))
    tasks = np.zeros((E, T, 4), dtype=np.uint8)
    for e in range(E):
        s = stor[rng.integers(0, len(stor), size=T)]; p = pick[rng.integers(0, len(pick), size=T)]
        tasks[e, :, 0] = s[:, 1] + 1; tasks[e, :, 1] = s[:, 0] + 1; tasks[e, :, 2] = p[:, 1] + 1; tasks[e, :, 3] = p[:, 0] + 1
    proto = int(seed % 2)
    policy = [pkg.POLICY_GIVEN, pkg.POLICY_GUIDED, pkg.POLICY_ASTAR][seed % 3]
    want_next = bool(seed % 4 == 1)
    results = {}
    monkeypatch.setenv("AGV_ROLLOUT_IMPL", "persistent")
    for impl in ("warp", "lanes", "async", "rollout"):
        monkeypatch.setenv("AGV_TICK_IMPL", impl)
        sc = pkg.BatchedScene(grid, E, N, tasks=tasks, P=3, auto_reset=True)
        sc.reset()
        arng = np.random.default_rng(77 + seed)
        trace = []
        for t in range(60):
            given = None
            if policy == pkg.POLICY_GIVEN:
                given = arng.choice(np.array([-1, -1, -1, -1, -1, 0, 1, 2, 3, 4], dtype=np.int8), size=(E, N))
            o = sc.tick(proto, policy, pkg.OBS_CLIP, actions=given, want_next=want_next)
            act = o["act"].cpu().numpy()
            took = act >= 0
            rec = [act.copy(), o["reward"].cpu().numpy().copy(), o["done"].cpu().numpy().copy(),
                   (o["plen"].cpu().numpy().astype(np.int32) & 0xFFFF)[took], o["obs"].float().cpu().numpy()[took]]
            if want_next:
                rec.append(o["next_obs"].float().cpu().numpy()[took])
            trace.append(rec)
        hdr, agv = sc.get_state()
        results[impl] = (trace, hdr, agv)
        sc.close()
    ref = results["warp"]
    for impl in ("lanes", "async", "rollout"):
        tr, hdr, agv = results[impl]
        for t, (a, b) in enumerate(zip(ref[0], tr)):
            for q, (x, y) in enumerate(zip(a, b)):
                assert np.array_equal(x, y), (impl, seed, t, q)
        assert np.array_equal(ref[1], hdr) and np.array_equal(ref[2], agv), (impl, seed)
    assert sum(int((r[0] >= 0).sum()) for r in ref[0]) > 0
