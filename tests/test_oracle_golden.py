"""Pins the CPU oracle (oracle/) against fixtures produced by the live reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib
import json

import numpy as np
import pytest

from conftest import golden
from oracle import oracle as orc


def test_bfs_matches_reference_fixtures():
    cases = golden("bfs.json")
    assert cases[0]["name"] == "astar_demo"
    # the reference's own KAT (astar.py:169-186): R R R D D D D R R R U U L
    assert cases[0]["actions"] == [1, 1, 1, 2, 2, 2, 2, 1, 1, 1, 0, 0, 3]
    for c in cases:
        v = np.array(c["valid"], dtype=np.uint8)
        a_l, f_l, acts = orc.bfs_literal(v, c["start"], c["target"])
        a_f, f_f, ln = orc.bfs_field(v, c["start"], c["target"])
        assert f_l == c["found"] and f_f == c["found"]
        assert acts == c["actions"]
        assert ln == len(c["actions"])
        first = c["actions"][0] if c["actions"] else 4
        assert a_l == first and a_f == first


def _large_cases():
    for c in golden("bfs_large.json"):
        v = np.array([[int(ch) for ch in row] for row in c["valid"]], dtype=np.uint8)
        yield c, v


def test_bfs_matches_reference_fixtures_at_baseline_sizes():
    """63 FindPathAstar runs of the live reference at 64x64 (bench workload c3 matrices with
    AGV obstacles, random grids, walls) and 128x128 (c5): closes the parity chain
    reference -> literal restatement -> distance-field rule at BASELINE sizes."""
    n64 = n128 = 0
    for c, v in _large_cases():
        assert v.shape == (c["H"], c["W"])
        a_l, f_l, acts = orc.bfs_literal(v, c["start"], c["target"])
        a_f, f_f, ln = orc.bfs_field(v, c["start"], c["target"])
        assert f_l == c["found"] and f_f == c["found"], c["name"]
        assert acts == c["actions"], c["name"]
        assert ln == len(c["actions"]), c["name"]
        first = c["actions"][0] if c["actions"] else 4
        assert a_l == first and a_f == first, c["name"]
        n64 += c["H"] == 64
        n128 += c["H"] == 128
    assert n64 >= 50 and n128 >= 5


def test_bfs_field_equals_literal_large_grids():
    rng = np.random.default_rng(5)
    for n in range(60):
        H, W = int(rng.integers(20, 70)), int(rng.integers(20, 70))
        v = (rng.random((H, W)) > rng.choice([0.0, 0.15, 0.3, 0.38])).astype(np.uint8)
        s = (int(rng.integers(W)), int(rng.integers(H)))
        t = (int(rng.integers(W)), int(rng.integers(H)))
        a_l, f_l, acts = orc.bfs_literal(v, s, t)
        a_f, f_f, ln = orc.bfs_field(v, s, t)
        assert (a_l, f_l, len(acts)) == (a_f, f_f, ln)


@pytest.mark.parametrize("idx", range(7))
def test_astar_trajectories(idx):
    c = golden("astar_traj.json")[idx]
    env = orc.Env(np.array(c["grid"], dtype=np.uint8), c["tasks"], c["n_agv"])
    veh, actions, pos = [], [], []
    ticks = 0
    while not env.state()[0][5] and ticks < c["ticks"] + 5:
        hdr, agv = env.state()
        # replay turn by turn so that the snapshot before every turn can be compared
        env.begin_tick()
        returned = False
        for i in range(env.N):
            st = env.turn_state(i)
            if st == -1:
                break
            if st == -2:
                returned = True
                break
            if st == 0:
                continue
            a, _ = env.action_astar(i)
            veh.append(i)
            actions.append(a)
            pos.append(env.snapshot()[:, :2].tolist())
            env.execute(i, a, orc.PROTO_ASTAR)
        ticks += 1
        if returned:
            break
        env.spawn()
        if not c["finished"] and ticks >= c["ticks"]:
            break
    assert veh == c["veh"]
    assert actions == c["actions"]
    assert pos == c["pos"]
    assert ticks == c["ticks"]
    names = ["veh%d" % (v + 1) for v in veh]
    sha = hashlib.sha256(json.dumps([names, actions, pos]).encode()).hexdigest()[:16]
    assert sha == c["sha"]
    assert env.snapshot().tolist() == c["final"]


def test_astar_tick_driver_equals_turn_replay():
    c = golden("astar_traj.json")[0]
    env = orc.Env(np.array(c["grid"], dtype=np.uint8), c["tasks"], c["n_agv"])
    acts = []
    for _ in range(c["ticks"] + 3):
        n, act, rew, end, plen = env.tick(orc.PROTO_ASTAR, orc.POLICY_ASTAR)
        acts += [int(a) for a in act if a >= 0]
        if env.state()[0][5]:
            break
    assert acts == c["actions"]
    assert int(env.state()[0][0]) == c["ticks"]


def _replay_intelligent(ep, check_obs=True):
    flags = ep["flags"]
    env = orc.Env(np.array(ep["grid"], dtype=np.uint8), ep["tasks"], ep["n_agv"],
                  always_loaded=flags.get("Explorer_Always_Loaded", False),
                  always_empty=flags.get("Explorer_Always_Empty", False),
                  shaped=flags.get("Sparse_Reward", False), max_steps=ep["max_steps"])
    assert env.max_steps == orc.max_training_steps(env.W, env.H)
    turns = ep["turns"]
    k = 0
    while k < len(turns):
        env.begin_tick()
        ended = False
        for i in range(env.N):
            st = env.turn_state(i)
            if st == -1:
                break
            assert st != -2
            if st == 0:
                continue
            t = turns[k]
            assert t["i"] == i
            assert env.snapshot().tolist() == t["before"]
            ga, gl = env.action_guided(i)
            assert ga == t["guided"] and gl == t["glen"]
            if "obs_clip" in t and check_obs:
                assert env.encode(i, clip=True, P=3).astype(int).tolist() == t["obs_clip"]
                assert env.encode(i, clip=False).astype(int).tolist() == t["obs_full"]
            r, end, ri, su, pl = env.execute(i, t["a"], orc.PROTO_INTELLIGENT)
            assert r == t["r"], (k, r, t["r"])  # exact, also for the float64 shaped reward
            assert int(end) == t["end"]
            assert env.snapshot().tolist() == t["after"]
            k += 1
            if end:
                ended = True
                break
        if ended:
            break
        env.spawn()
    assert k == len(turns)
    assert turns[-1]["end"] == 1
    if ep["running_time"] >= 0:
        assert int(env.state()[0][0]) == ep["running_time"]


def test_intelligent_episodes():
    eps = golden("intelligent.json")
    n_obs = 0
    for ep in eps:
        _replay_intelligent(ep)
        n_obs += sum(1 for t in ep["turns"] if "obs_clip" in t)
    assert n_obs > 50


def test_intelligent_tick_driver_given_actions():
    """orc_tick(POLICY_GIVEN) == turn-by-turn replay"""
    eps = golden("intelligent.json")
    for ep in eps[::5]:
        flags = ep["flags"]
        env = orc.Env(np.array(ep["grid"], dtype=np.uint8), ep["tasks"], ep["n_agv"],
                      always_loaded=flags.get("Explorer_Always_Loaded", False),
                      always_empty=flags.get("Explorer_Always_Empty", False),
                      shaped=flags.get("Sparse_Reward", False), max_steps=ep["max_steps"])
        turns = ep["turns"]
        k = 0
        while k < len(turns) and not env.state()[0][5]:
            # actions for this tick: next turns in order, indexed by AGV
            given = np.full(env.N, 4, dtype=np.int8)
            kk = k
            seen = set()
            while kk < len(turns) and turns[kk]["i"] not in seen and (not seen or turns[kk]["i"] > max(seen)):
                given[turns[kk]["i"]] = turns[kk]["a"]
                seen.add(turns[kk]["i"])
                kk += 1
            n, act, rew, end, plen = env.tick(orc.PROTO_INTELLIGENT, orc.POLICY_GIVEN, given)
            for i in range(env.N):
                if act[i] >= 0:
                    assert turns[k]["i"] == i and turns[k]["a"] == act[i]
                    assert rew[i] == turns[k]["r"] and end[i] == turns[k]["end"]
                    k += 1
        assert k == len(turns)


def test_per_sumtree():
    for c in golden("per.json"):
        tree = orc.SumTree(c["capacity"])
        for e in c["errors"]:
            tree.add(orc.per_priority(e))
        assert tree.n_entries == c["n_entries"]
        assert tree.total() == c["total"]
        slots, tidx, pr = orc.per_sample(tree, len(c["uniforms"]), c["uniforms"])
        assert tidx.tolist() == c["tree_idx"]
        # the reference stored the insertion counter k as data: slot -> last k written there
        last = {}
        for k in range(len(c["errors"])):
            last[k % c["capacity"]] = k
        assert [last[s] for s in slots.tolist()] == c["batch"]


def test_layout_geometry():
    g = orc.rect_layout(4, 2, 2, 2, 2)
    assert g.shape == (10, 11)
    ss, ps = orc.stations(g)
    assert ps == [(4, 9), (8, 9)] and len(ss) == 32
    c = golden("astar_traj.json")[0]
    assert g.tolist() == c["grid"]
    import random
    rng = random.Random(0)
    assert [list(t) for t in orc.make_tasks(g, rng)] == c["tasks"]
    assert orc.max_training_steps(11, 10) == int((11 / 2 + 10) * 3 * ((11 * 10) / 2.5))
