"""GPU tests of the reference-facing host layer (Layout / Explorer / Scene / StateManager /
FindPathAstar / PER Memory / Expert) against fixtures recorded from the live reference."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, golden, load_pkg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


def _make_scene(pkg, grid, tasks, n, flags=None):
    flags = flags or {}
    SP = pkg.SuperParas
    SP.Explorer_Always_Loaded = bool(flags.get("Explorer_Always_Loaded", False))
    SP.Explorer_Always_Empty = bool(flags.get("Explorer_Always_Empty", False))
    SP.Sparse_Reward = bool(flags.get("Sparse_Reward", False))
    try:
        layout = pkg.Layout(layout_list=grid, task_list=[tuple(t) for t in tasks])
        group = [pkg.Explorer(layout, veh_name="veh%d" % (i + 1), icon_name="veh%d" % (i + 1)) for i in range(n)]
        return pkg.Scene(layout, group)
    finally:
        SP.Explorer_Always_Loaded = SP.Explorer_Always_Empty = SP.Sparse_Reward = False


@pytest.mark.parametrize("idx", [0, 3, 5, 6])
def test_scene_astar_mode_matches_reference(pkg, idx):
    c = golden("astar_traj.json")[idx]
    scene = _make_scene(pkg, c["grid"], c["tasks"], c["n_agv"])
    assert c["finished"]
    info, names, actions = scene.run_game(control_pattern="A_star", render=False)
    assert actions == c["actions"]
    assert [int(n[3:]) - 1 for n in names] == c["veh"]
    pos = [[v[1] for v in i[1:]] for i in info]
    assert pos == c["pos"]
    sha = hashlib.sha256(json.dumps([names, actions, pos]).encode()).hexdigest()[:16]
    assert sha == c["sha"]  # SURVEY App. B.4 fingerprints
    assert scene.running_time == c["ticks"]
    assert info[0][0] == c["grid"]


def test_scene_astar_gridlock_cap(pkg):
    c = golden("astar_traj.json")[4]
    scene = _make_scene(pkg, c["grid"], c["tasks"], c["n_agv"])
    scene.explorer_group[0].create_explorer()
    info, names, actions = scene.run_mode("A_star", max_ticks=c["ticks"])
    assert actions == c["actions"] and not c["finished"]


class ReplayController:
    """returns the actions the live reference's controller chose and checks everything the
    scene hands to the callbacks"""

    def __init__(self, turns):
        self.turns, self.k = turns, 0

    def choose_action(self, all_info, name):
        t = self.turns[self.k]
        assert name == "veh%d" % (t["i"] + 1)
        snap = [[v[1][0], v[1][1], v[2][0], v[2][1], int(bool(v[3]))] for v in all_info[1:]]
        assert snap == t["before"], self.k
        return t["a"]

    def store_info(self, all_info, reward, is_end, name):
        t = self.turns[self.k]
        snap = [[v[1][0], v[1][1], v[2][0], v[2][1], int(bool(v[3]))] for v in all_info[1:]]
        assert snap == t["after"], self.k
        assert reward == t["r"] and type(reward) is type(t["r"]), (self.k, reward, t["r"])
        assert bool(is_end) == bool(t["end"])
        self.k += 1


def test_scene_intelligent_mode_callbacks(pkg):
    eps = golden("intelligent.json")
    n = 0
    for ep in eps[::4]:
        scene = _make_scene(pkg, ep["grid"], ep["tasks"], ep["n_agv"], ep["flags"])
        assert scene.max_training_steps == ep["max_steps"]
        ctl = ReplayController(ep["turns"])
        rt = scene.run_game(control_pattern="intelligent", smart_controller=ctl, render=True)
        assert ctl.k == len(ep["turns"])
        assert rt == ep["running_time"]
        n += ctl.k
        # Scene.init() starts a fresh episode with the same fixed task list
        scene.init()
        assert scene.create_info() == [scene.layout.layout_original]
        ctl2 = ReplayController(ep["turns"])
        assert scene.run_game("intelligent", ctl2, render=False) == ep["running_time"]
    assert n > 1000


def test_state_manager_matches_reference(pkg):
    sm = pkg.StateManager()
    n = 0
    for ep in golden("intelligent.json"):
        names = ["veh%d" % (i + 1) for i in range(ep["n_agv"])]
        for t in ep["turns"]:
            if "obs_clip" not in t:
                continue
            info = [ep["grid"]] + [[names[j], [v[0], v[1]], [v[2], v[3]], bool(v[4])] for j, v in enumerate(t["before"])]
            obs, cp, tp, valid = sm.create_state(info, names[t["i"]], obs_clip=True, padding_size=3)
            assert np.array(obs).astype(int).tolist() == t["obs_clip"]
            obs_f, cp, tp, valid = sm.create_state(info, names[t["i"]])
            assert np.array(obs_f).astype(int).tolist() == t["obs_full"]
            assert np.array(valid).astype(int).tolist() == t["obs_full"][2]
            assert cp == [t["before"][t["i"]][0], t["before"][t["i"]][1]]
            assert isinstance(obs[0][0][0], float)
            n += 1
    assert n > 50
    # batch form == single form
    ep = golden("intelligent.json")[0]
    names = ["veh%d" % (i + 1) for i in range(ep["n_agv"])]
    infos = [[ep["grid"]] + [[names[j], [v[0], v[1]], [v[2], v[3]], bool(v[4])] for j, v in enumerate(t["before"])]
             for t in ep["turns"][:20]]
    vehs = [names[t["i"]] for t in ep["turns"][:20]]
    b = sm.create_state_batch(infos, vehs)
    for k in range(len(infos)):
        assert np.array_equal(b[k], np.array(sm.create_state(infos[k], vehs[k])[0], dtype=np.float32))


def test_find_path_astar_matches_reference(pkg):
    NAMES = ["UP", "RIGHT", "DOWN", "LEFT", "STOP"]
    cases = golden("bfs.json")
    for c in cases[:150]:
        f = pkg.FindPathAstar([list(map(float, r)) for r in c["valid"]], tuple(c["start"]), tuple(c["target"]))
        found, path, pmap, acts = f.run_astar_method()
        assert bool(found) == c["found"]
        assert [NAMES.index(a) for a in acts] == c["actions"]
        if "path" in c:  # the reference's own KAT, astar.py:169-186
            assert [list(p) for p in path] == c["path"]
            assert pmap[3][0] == -1 and pmap[5][5] == -1


def test_per_memory_matches_reference(pkg):
    import torch
    for c in golden("per.json"):
        mem = pkg.Memory(c["capacity"])
        for k, e in enumerate(c["errors"]):
            s = np.full((3, 7, 7), k % 2, dtype=np.float64)
            mem.add(e, (s, k % 4, float(k), s, k % 2))
        assert mem.tree.n_entries == c["n_entries"]
        assert mem.tree.total() == c["total"]  # float64 sums in the reference's order
        it = iter(c["uniforms"])
        orig = random.random
        random.random = lambda: next(it)
        try:
            batch, idxs, isw = mem.sample(len(c["uniforms"]))
        finally:
            random.random = orig
        assert idxs == c["tree_idx"]
        assert [int(b[2]) for b in batch] == c["batch"]  # reward carries the insertion counter k
        assert all(b[1] == int(b[2]) % 4 and b[4] == int(b[2]) % 2 and b[0].shape == (3, 7, 7) for b in batch)
        assert isw.max() == 1.0
        # Memory.update changes the total like the reference formula
        t0 = mem.tree.total()
        mem.update(idxs[0], 2.0)
        tree = orc.SumTree(c["capacity"])
        for e in c["errors"]:
            tree.add(orc.per_priority(e))
        tree.update(idxs[0], orc.per_priority(2.0))
        assert mem.tree.total() == tree.total() and t0 != mem.tree.total()


def test_replay_large_push_and_gather(pkg):
    """the rebuild path (pushes > 2048 rows): ring wrap, skipped rows, totals, gather"""
    import torch
    dev = torch.device("cuda", 0)
    cap, R = 5000, 12
    rp = pkg.DeviceReplay(cap, (R,))
    rng = np.random.default_rng(0)
    tree = np.zeros(cap)
    store = {}
    write = 0
    for it in range(4):
        n = 3000
        act = rng.integers(-1, 4, n).astype(np.int8)
        rew = rng.random(n).astype(np.float32)
        done = rng.integers(0, 2, n).astype(np.uint8)
        obs = rng.integers(0, 2, (n, R)).astype(np.float32)
        pr = rng.random(n) + 0.1
        rp.push(torch.tensor(act, device=dev), torch.tensor(rew, device=dev), torch.tensor(done, device=dev),
                torch.tensor(obs, device=dev).to(torch.bfloat16), torch.tensor(obs[::-1].copy(), device=dev).to(torch.bfloat16),
                torch.tensor(pr, device=dev))
        for i in np.nonzero(act >= 0)[0]:
            tree[write] = pr[i]
            store[write] = (act[i], rew[i], done[i], obs[i], obs[n - 1 - i])
            write = (write + 1) % cap
    ne, w, total = rp.info()
    assert ne == min(cap, len(store)) and w == write
    assert abs(total - tree.sum()) < 1e-9 * tree.sum()
    slots = np.array(sorted(store))[::37]
    b = rp.gather(slots)
    for q, s in enumerate(slots):
        a, r, d, o, o2 = store[int(s)]
        assert int(b["act"][q]) == a and float(b["reward"][q]) == r and int(b["done"][q]) == d
        assert np.array_equal(b["obs"][q].float().cpu().numpy(), o) and np.array_equal(b["next_obs"][q].float().cpu().numpy(), o2)
    # stratified sampling == the reference's descent (PER.py:25-35) on the same heap-layout tree
    # (for a capacity that is not a power of two the leaves are NOT in slot order)
    u = rng.random(64)
    sb = rp.sample(64, u)
    ot = orc.SumTree(cap)
    ot.tree[cap - 1:] = tree
    for v in range(cap - 2, -1, -1):
        ot.tree[v] = ot.tree[2 * v + 1] + ot.tree[2 * v + 2]
    slots_ref, _, prio_ref = orc.per_sample(ot, 64, u)
    assert sb["slot"].cpu().numpy().tolist() == slots_ref.tolist()
    assert np.array_equal(sb["prio"].cpu().numpy(), prio_ref)
    rp.close()


def test_expert_collection_writes_the_reference_csv(pkg, tmp_path):
    meta = golden("expert_ref_meta.json")
    for m in meta:
        random.seed(m["seed"])
        layout = pkg.Layout(*m["rect"])
        group = [pkg.Explorer(layout, veh_name="veh%d" % (i + 1), icon_name="veh%d" % (i + 1))
                 for i in range(m["n_agv"])]
        scene = pkg.Scene(layout, group)
        out = str(tmp_path / ("expert_%s.csv" % m["tag"]))
        ex = pkg.Expert(scene, *m["rect"], m["n_agv"], filename=out)
        ex.create_data_by_self(times=m["times"])
        ref = open(os.path.join(GOLDEN, "expert_ref_%s.csv" % m["tag"]), "rb").read()
        assert open(out, "rb").read() == ref  # byte-identical wire format and content
        assert ex.truncated == 0 and ex.expert_size == m["times"]
        # behavioural cloning from the file: sample -> encode on the GPU -> one Adam step
        import torch
        net = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * layout.scene_y_width * layout.scene_x_width, 5),  # 5: expert STOPs index 4
                                  torch.nn.Softmax(dim=1)).cuda()
        s, a = ex.sample_data()
        assert len(s) == len(a) and s[0].shape == (3, layout.scene_y_width, layout.scene_x_width)
        loss = ex.pre_training(net, torch.device("cuda"))
        assert torch.isfinite(loss)


def test_batched_agdqn_agent(pkg):
    """AG-DQN with the network in the loop: epsilon=0 reproduces the fused guided tick exactly;
    a short training run fills the device replay and takes finite optimiser steps"""
    import importlib
    import torch
    M = importlib.import_module("reinforcement-learning-for-multi-agv-pathfinding_b200.algorithm.MADQN_structure.MADQN")
    grid = orc.README_9x9
    E, N = 64, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(3 * e)) for e in range(E)], dtype=np.uint8)
    a = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
    b = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
    a.reset()
    b.reset()
    torch.manual_seed(0)
    agent = M.BatchedAgent(a, memory_size=4096, batch_size=32, epsilon=0.0)
    for t in range(40):
        n = agent.tick(train=False)
        out = b.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED)
        assert int(n) == int((out["act"] >= 0).sum())
        ha, ga = a.get_state()
        hb, gb = b.get_state()
        assert np.array_equal(ha, hb) and np.array_equal(ga, gb)
    # training: epsilon 0.8, every sub-step stores E transitions and takes one optimiser step
    agent = M.BatchedAgent(a, memory_size=4096, batch_size=32, epsilon=0.8, seed=1)
    w0 = agent.policy_network.fc4.weight.detach().clone()
    turns = sum(int(agent.tick(train=True)) for _ in range(12))
    n_entries, _, total = agent.memory.info()
    assert turns > 0 and 0 < n_entries <= 4096 and total > 0
    # learning starts once the replay really holds 100 entries (MADQN.py:131): within the first two ticks here
    assert 10 * N <= agent.learn_step_counter <= 12 * N
    assert agent.transitions_stored() == n_entries
    import math
    losses = agent.loss_value
    assert len(losses) == agent.learn_step_counter and all(math.isfinite(l) and abs(l) <= 0.5 for l in losses)
    assert not torch.equal(w0, agent.policy_network.fc4.weight)
    # stored actions are network-indexable (a guidance STOP is never stored)
    bt = agent.memory.gather(torch.arange(n_entries))
    assert int(bt["act"].min()) >= 0 and int(bt["act"].max()) <= 3
    assert set(bt["obs"].float().unique().tolist()) <= {0.0, 1.0}
    # the reference Net layout: 1.74 M parameters, 2304-wide flatten
    assert sum(p.numel() for p in M.Net().parameters()) == 1736516
    assert M.Net()(torch.zeros(2, 3, 7, 7)).shape == (2, 4)
    a.close()
    b.close()


def test_expert_rollouts_into_device_replay(pkg):
    """A_star-mode turns land in the replay buffer in (scene, AGV) order with the expert action"""
    import torch
    grid = orc.rect_layout(4, 2, 2, 2, 2)
    E, N = 8, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    sc = pkg.BatchedScene(grid, E, N, tasks=tasks)
    sc.reset()
    rp = pkg.DeviceReplay(4096, (3, 7, 7))
    envs = [orc.Env(grid, tasks[e], N) for e in range(E)]
    exp = []
    for t in range(20):
        for env in envs:
            n, a, r, d, pl, sl, ob, nx = env.tick_obs(orc.PROTO_ASTAR, orc.POLICY_ASTAR, obs_kind=1, P=3, want_next=True)
            for i in range(N):
                if a[i] >= 0:
                    exp.append((int(a[i]), float(r[i]), ob[i], nx[i]))
    pushed = int(sc.collect_expert(rp, 20))
    assert pushed == len(exp) and rp.info()[0] == len(exp)
    b = rp.gather(torch.arange(len(exp)))
    act = b["act"].cpu().numpy()
    rew = b["reward"].cpu().numpy()
    obs = b["obs"].float().cpu().numpy()
    nxt = b["next_obs"].float().cpu().numpy()
    for q, (a, r, ob, nx) in enumerate(exp):
        assert act[q] == a and rew[q] == np.float32(r)
        assert np.array_equal(obs[q], ob) and np.array_equal(nxt[q], nx)
    sc.close()
    rp.close()


def test_q_values_within_bf16_tolerance(pkg):
    """north star: Q-values within a stated bf16/fp32 tolerance.  Reference side: the states the
    live reference's StateManager produced (fixture), fp32 forward on the CPU.  Our side: the CUDA
    encoder's bf16 states, bf16-autocast channels-last forward on the GPU, same weights.
    Tolerance: |dQ| <= 2e-2 * max|Q|.  The states themselves are bit-identical."""
    import importlib
    import torch
    M = importlib.import_module("reinforcement-learning-for-multi-agv-pathfinding_b200.algorithm.MADQN_structure.MADQN")
    torch.manual_seed(0)
    net = M.Net()
    ref_obs, infos, vehs = [], [], []
    for ep in golden("intelligent.json"):
        names = ["veh%d" % (i + 1) for i in range(ep["n_agv"])]
        for t in ep["turns"]:
            if "obs_clip" in t and ep["grid"] == golden("intelligent.json")[0]["grid"]:
                ref_obs.append(t["obs_clip"])
                infos.append([ep["grid"]] + [[names[j], [v[0], v[1]], [v[2], v[3]], bool(v[4])]
                                             for j, v in enumerate(t["before"])])
                vehs.append(names[t["i"]])
    assert len(ref_obs) >= 8
    x_ref = torch.tensor(np.array(ref_obs, dtype=np.float32))
    with torch.no_grad():
        q_ref = net(x_ref)
    ours = pkg.StateManager().create_state_batch(infos, vehs, obs_clip=True)
    assert np.array_equal(ours, x_ref.numpy())  # integer part: exact
    gnet = M.Net().cuda()
    gnet.load_state_dict(net.state_dict())
    gnet = gnet.to(memory_format=torch.channels_last)
    xb = torch.tensor(ours).cuda().to(torch.bfloat16)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        q_bf16 = gnet(xb.contiguous(memory_format=torch.channels_last)).float().cpu()
    with torch.no_grad():
        q_fp32 = gnet(xb.float()).cpu()
    scale = q_ref.abs().max()
    assert (q_fp32 - q_ref).abs().max() <= 1e-4 * scale          # fp32 on the GPU vs fp32 on the CPU
    assert (q_bf16 - q_ref).abs().max() <= 2e-2 * scale          # bf16 autocast, stated tolerance
