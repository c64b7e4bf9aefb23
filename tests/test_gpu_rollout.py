"""agv_rollout (persistent multi-tick kernel, csrc/agv_rollout.cuh): a K-tick rollout must equal K single
fused ticks bit for bit -- outputs of every tick, observation records of every turn, final scene state --
and, through them, the CPU oracle and the live-reference fixtures."""
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


@pytest.fixture(autouse=True, params=["persistent", "default"])
def rollout_impl(request, monkeypatch):
    """every test runs twice: with the persistent kernel forced on every map shape, and with the library's
    own choice (maps up to 32 x 64 take the per-tick path inside agv_rollout)"""
    if request.param == "persistent":
        monkeypatch.setenv("AGV_ROLLOUT_IMPL", "persistent")
    return request.param


def _np(o, key):
    return o[key].cpu().numpy()


def _compare_rollout_with_ticks(pkg, grid, E, N, tasks, K, rounds, proto, policy, obs_kind, want_next, seed=0,
                                auto_reset=True, given_choices=None, **kw):
    """scene A: `rounds` rollouts of K ticks; scene B: rounds*K single ticks.  Everything must match."""
    a = pkg.BatchedScene(grid, E, N, tasks=tasks, P=3, auto_reset=auto_reset, **kw)
    b = pkg.BatchedScene(grid, E, N, tasks=tasks, P=3, auto_reset=auto_reset, **kw)
    a.reset(); b.reset()
    rng = np.random.default_rng(seed)
    turns = 0
    for r in range(rounds):
        given = None
        if policy == pkg.POLICY_GIVEN:
            given = rng.choice(np.array(given_choices or [-1, -1, -1, -1, -1, 0, 1, 2, 3, 4], dtype=np.int8), size=(K, E, N))
        o = a.rollout(K, proto, policy, obs_kind, actions=given, want_next=want_next)
        ra = {k: _np(o, k) for k in ("act", "reward", "done")}
        ra["plen"] = _np(o, "plen").astype(np.int32) & 0xFFFF
        ra["slen"] = _np(o, "slen").astype(np.int32) & 0xFFFF
        if obs_kind:
            ra["obs"] = o["obs"].float().cpu().numpy()
            if want_next:
                ra["next_obs"] = o["next_obs"].float().cpu().numpy()
        for k in range(K):
            ob = b.tick(proto, policy, obs_kind, actions=None if given is None else given[k], want_next=want_next)
            act = _np(ob, "act")
            took = act >= 0
            turns += int(took.sum())
            assert np.array_equal(ra["act"][k], act), (r, k)
            assert np.array_equal(ra["reward"][k], _np(ob, "reward")), (r, k)
            assert np.array_equal(ra["done"][k], _np(ob, "done")), (r, k)
            assert np.array_equal(ra["plen"][k], _np(ob, "plen").astype(np.int32) & 0xFFFF), (r, k)
            assert np.array_equal(ra["slen"][k], _np(ob, "slen").astype(np.int32) & 0xFFFF), (r, k)
            if obs_kind:
                assert np.array_equal(ra["obs"][k][took], ob["obs"].float().cpu().numpy()[took]), (r, k)
                if want_next:
                    assert np.array_equal(ra["next_obs"][k][took], ob["next_obs"].float().cpu().numpy()[took]), (r, k)
        ha, ga = a.get_state()
        hb, gb = b.get_state()
        assert np.array_equal(ha, hb) and np.array_equal(ga, gb), r
    ca, cb = a.counters(), b.counters()
    assert ca == cb and ca[0] == turns
    a.close(); b.close()
    return turns


ROLL_CASES = [
    # name, grid, N, E, K, rounds, proto, policy, obs_kind, want_next
    ("c1_guided_clip", lambda: orc.rect_layout(4, 2, 2, 2, 2), 4, 40, 16, 8, 1, 2, 1, True),
    ("c2_9x9_guided", lambda: orc.README_9x9, 4, 70, 25, 4, 1, 2, 1, False),
    ("c1_astar_noobs", lambda: orc.rect_layout(4, 2, 2, 2, 2), 6, 9, 30, 4, 0, 1, 0, False),
    ("c3_guided_clip", lambda: orc.rect_layout(6, 2, 9, 20, 8), 64, 37, 12, 12, 1, 2, 1, False),
    ("c3_guided_clip_next", lambda: orc.rect_layout(6, 2, 9, 20, 8), 64, 8, 10, 10, 1, 2, 1, True),
    ("c3_astar_clip", lambda: orc.rect_layout(6, 2, 9, 20, 8), 40, 5, 20, 4, 0, 1, 1, True),
    ("c3_given", lambda: orc.rect_layout(6, 2, 9, 20, 8), 24, 33, 8, 8, 1, 0, 1, True),
    ("c3_given_astar_proto", lambda: orc.rect_layout(6, 2, 9, 20, 8), 16, 12, 8, 6, 0, 0, 1, True),
    ("wide_128", lambda: orc.rect_layout(6, 2, 18, 41, 9), 48, 5, 10, 6, 1, 2, 1, False),
    ("tall_40x128", lambda: orc.rect_layout(2, 2, 13, 41, 3), 33, 3, 16, 4, 1, 2, 1, True),
    ("n7_not_multiple_of_8", lambda: orc.rect_layout(5, 2, 6, 11, 4), 7, 21, 12, 5, 1, 2, 1, True),
    ("single_agv", lambda: orc.rect_layout(3, 2, 2, 2, 2), 1, 8, 50, 3, 1, 2, 1, False),
    ("more_agv_than_tasks", lambda: orc.rect_layout(1, 1, 2, 1, 1), 6, 8, 20, 3, 0, 1, 1, False),
]


@pytest.mark.parametrize("case", ROLL_CASES, ids=[c[0] for c in ROLL_CASES])
def test_rollout_equals_single_ticks(pkg, case):
    name, gridf, N, E, K, rounds, proto, policy, obs_kind, want_next = case
    grid = gridf()
    tasks = np.array([orc.make_tasks(grid, random.Random(100 + e)) for e in range(E)], dtype=np.uint8)
    turns = _compare_rollout_with_ticks(pkg, grid, E, N, tasks, K, rounds, proto, policy, obs_kind, want_next, seed=5)
    assert turns > K * rounds


@pytest.mark.parametrize("shape", [1, 2], ids=["light", "heavy"])
@pytest.mark.parametrize("proto,policy", [(1, 2), (0, 1)], ids=["guided", "astar_mode"])
@pytest.mark.parametrize("layout", [(6, 2, 18, 41, 9), (6, 2, 9, 20, 8)], ids=["127x127", "64x64"])
def test_rollout_both_launch_shapes(pkg, monkeypatch, layout, shape, proto, policy):
    """the persistent kernel has two launch shapes per map class (BFS workers per CTA x CTAs per SM), picked by
    a load hint (agv_rollout.cu: launch_shape); either must reproduce the single ticks bit-exactly"""
    monkeypatch.setenv("AGV_ROLLOUT_SHAPE", str(shape))
    grid = orc.rect_layout(*layout)
    E, N, K = 40, 48, 6
    tasks = np.array([orc.make_tasks(grid, random.Random(300 + e)) for e in range(E)], dtype=np.uint8)
    turns = _compare_rollout_with_ticks(pkg, grid, E, N, tasks, K, 6, proto, policy, pkg.OBS_CLIP, True, seed=9)
    assert turns > K * 6


def test_rollout_without_auto_reset_and_dead_scenes(pkg):
    """episodes that end stay over for the rest of the rollout (no auto-reset): remaining ticks do nothing"""
    grid = orc.rect_layout(4, 2, 2, 2, 2)
    E, N = 48, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    _compare_rollout_with_ticks(pkg, grid, E, N, tasks, 40, 3, pkg.PROTO_INTELLIGENT, pkg.POLICY_GIVEN, pkg.OBS_CLIP,
                                True, seed=9, auto_reset=False, given_choices=[-1, -1, -1, -1, -1, -1, -1, 0, 1, 2, 3])


def test_rollout_scene_queue(pkg, monkeypatch):
    """more scenes than resident slots (AGV_ROLLOUT_CPS=1: one CTA per SM): scenes are pulled from the
    heaviest-first queue as slots free up; the schedule must not change any result"""
    import torch
    monkeypatch.setenv("AGV_ROLLOUT_CPS", "1")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    grid = orc.rect_layout(5, 2, 6, 11, 4)  # 37 x 37: lane-per-scene variant <1, 2>
    E, N = sms * 32 + 1500, 8
    rng = random.Random(3)
    base = [orc.make_tasks(grid, random.Random(s)) for s in range(64)]
    tasks = np.array([base[rng.randrange(64)] for _ in range(E)], dtype=np.uint8)
    turns = _compare_rollout_with_ticks(pkg, grid, E, N, tasks, 12, 3, pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED,
                                        pkg.OBS_CLIP, False, seed=1)
    assert turns > E * 12


def test_rollout_vs_oracle_and_reference_fixture(pkg):
    """Scene.run_game('A_star') trajectories recorded from the live reference, replayed as ONE rollout"""
    for c in golden("astar_traj.json"):
        grid = np.array(c["grid"], dtype=np.uint8)
        tasks = np.array(c["tasks"], dtype=np.uint8)[None]
        sc = pkg.BatchedScene(grid, 1, c["n_agv"], tasks=tasks)
        sc.reset()
        K = c["ticks"]
        o = sc.rollout(K, pkg.PROTO_ASTAR, pkg.POLICY_ASTAR)
        act = o["act"].cpu().numpy()[:, 0]
        veh, actions = [], []
        for k in range(K):
            for i in range(c["n_agv"]):
                if act[k, i] >= 0:
                    veh.append(i)
                    actions.append(int(act[k, i]))
        assert veh == c["veh"] and actions == c["actions"], c["name"]
        hdr, agv = sc.get_state()
        assert int(hdr[0, 0]) == c["ticks"]
        snap = [[int(v) for v in agv[0, j, [0, 1, 2, 3, 5]]] for j in range(int(hdr[0, 4]))]
        assert snap == c["final"]
        sc.close()


def test_rollout_vs_oracle_c3(pkg):
    """64x64 / 64 AGVs: a 48-tick rollout with observations against the CPU oracle, turn by turn"""
    grid = orc.rect_layout(6, 2, 9, 20, 8)
    E, N, K = 5, 64, 48
    tasks = np.array([orc.make_tasks(grid, random.Random(40 + e)) for e in range(E)], dtype=np.uint8)
    sc = pkg.BatchedScene(grid, E, N, tasks=tasks, P=3, auto_reset=True)
    sc.reset()
    envs = [orc.Env(grid, tasks[e], N) for e in range(E)]
    for r in range(3):
        o = sc.rollout(K, pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP, want_next=True)
        act, rew, done = _np(o, "act"), _np(o, "reward"), _np(o, "done")
        plen = _np(o, "plen").astype(np.int32) & 0xFFFF
        obs, nxt = o["obs"].float().cpu().numpy(), o["next_obs"].float().cpu().numpy()
        for k in range(K):
            for e, env in enumerate(envs):
                if env.state()[0][5]:
                    env.reset()
                n, a, rr, d, pl, sl, ob, nx = env.tick_obs(orc.PROTO_INTELLIGENT, orc.POLICY_GUIDED, obs_kind=1, P=3,
                                                          want_next=True)
                took = a >= 0
                assert np.array_equal(act[k, e], a) and np.array_equal(rew[k, e], rr.astype(np.float32)), (r, k, e)
                assert np.array_equal(done[k, e], d) and np.array_equal(plen[k, e], pl), (r, k, e)
                assert np.array_equal(obs[k, e][took], ob[took]) and np.array_equal(nxt[k, e][took], nx[took]), (r, k, e)
        hdr, agv = sc.get_state()
        for e, env in enumerate(envs):
            oh, oa = env.state()
            assert np.array_equal(hdr[e], oh) and np.array_equal(agv[e], oa), (r, e)
    sc.close()


def test_rollout_fallback_configurations(pkg):
    """shaped reward / full-map observations are outside the persistent kernel: agv_rollout runs them as
    n_ticks fused ticks -- still identical to single ticks"""
    grid = orc.rect_layout(5, 2, 6, 11, 4)
    E, N = 6, 12
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    _compare_rollout_with_ticks(pkg, grid, E, N, tasks, 6, 3, pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP,
                                True, shaped_reward=True)
    _compare_rollout_with_ticks(pkg, grid, E, N, tasks, 6, 2, pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_FULL,
                                False)


def test_rollout_host_equals_rollout(pkg):
    import torch
    grid = orc.rect_layout(5, 2, 6, 11, 4)
    E, N, K = 40, 16, 9
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    a = pkg.BatchedScene(grid, E, N, tasks=tasks, P=3, auto_reset=True)
    b = pkg.BatchedScene(grid, E, N, tasks=tasks, P=3, auto_reset=True)
    a.reset(); b.reset()
    rng = np.random.default_rng(2)
    for r in range(5):
        given = torch.from_numpy(rng.choice(np.array([-1, -1, -1, 0, 1, 2, 3], dtype=np.int8), size=(K, E, N))).pin_memory()
        oh = a.rollout_host(K, pkg.PROTO_INTELLIGENT, pkg.POLICY_GIVEN, pkg.OBS_CLIP, actions_host=given, join=True)
        torch.cuda.synchronize()
        od = b.rollout(K, pkg.PROTO_INTELLIGENT, pkg.POLICY_GIVEN, pkg.OBS_CLIP, actions=given)
        for key in ("act", "reward", "done"):
            assert np.array_equal(oh[key].numpy(), od[key].cpu().numpy()), (r, key)
        took = od["act"].cpu().numpy() >= 0
        assert np.array_equal(oh["obs"].float().cpu().numpy()[took], od["obs"].float().cpu().numpy()[took])
    a.close(); b.close()
