"""The reference's training loops drive this package unchanged (BASELINE north star), checkpoints interchange
with the reference, and the batched learner is the reference's update (8f rank 3).

The reference sources come from the git-ignored baseline/_ref/ (vendored by __graft_entry__.build(); travels to
the GPU box) or /root/reference; tests that need them skip when neither is there."""
import contextlib
import importlib
import io
import math
import os
import random
import sys
import types

import numpy as np
import pytest

from conftest import load_pkg, PKG_NAME
from oracle import oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not rh.available(), reason="reference sources not vendored (baseline/_ref)")


@pytest.fixture(scope="module")
def pkg():
    import torch
    assert torch.cuda.is_available()
    return load_pkg()


def _ref_controller_modules():
    """imports the reference's controller modules untouched (gym is only used by its CartPole demos)"""
    rh.load()
    for m in ("gym",):
        sys.modules.setdefault(m, types.ModuleType(m))
    with contextlib.redirect_stdout(io.StringIO()):
        madqn = importlib.import_module("algorithm.MADQN_structure.Controller")
        pg = importlib.import_module("algorithm.PG_structure.Controller")
    return madqn, pg


def _seed_all(seed):
    import torch
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _run_reference_controller(kind, scene, layout, episodes, seed):
    """the UNMODIFIED reference controller on `scene` (the reference's Scene or this package's facade);
    networks on the CPU in both runs so that the two runs are arithmetic-identical"""
    import torch
    madqn, pg = _ref_controller_modules()
    _seed_all(seed)
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: False   # only for the controller's device choice
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            if kind == "madqn":
                ctl = madqn.MADQNAgentController(scene, layout.scene_x_width, layout.scene_y_width,
                                                 len(layout.storage_station_list), control_mode="train_NN",
                                                 state_number=3)
            else:
                ctl = pg.PGAgentController(scene, layout.scene_x_width, layout.scene_y_width,
                                           len(layout.storage_station_list), control_mode="train_NN",
                                           state_number=3, expert_guiding=False)
    finally:
        torch.cuda.is_available = real
    log = []
    with contextlib.redirect_stdout(io.StringIO()):
        for ep in range(episodes):
            ctl.self_init() if hasattr(ctl, "self_init") else None
            scene.init()
            try:
                rt = scene.run_game(control_pattern="intelligent", smart_controller=ctl, render=False)
                err = None
            except IndexError as e:   # the reference's own defect: a guidance STOP indexes the 4-wide Q row
                rt, err = None, "IndexError"
            log.append((rt if isinstance(rt, int) else None, float(ctl.reward_acc), int(getattr(ctl, "action_length_record", 0)), err))
    return log


@needs_ref
@pytest.mark.parametrize("kind", ["madqn", "pg"])
def test_reference_controllers_drive_the_facade_scene(pkg, kind):
    """AG-DQN (MADQNAgentController) and PG (PGAgentController) of the reference, source untouched, run the
    same episodes against the reference's Scene and against this package's Scene (only the Layout / Explorer /
    Scene imports swapped): running_time, accumulated reward and action count of every
    episode must be identical."""
    R = rh.load()
    rect, n_agv, episodes = (2, 2, 2, 1, 1), 3, 4
    random.seed(11)
    ref_layout = R.Layout(*rect)
    task_list = list(ref_layout.task_list)
    fixed = R.Layout(*rect, task_list=task_list)
    ref_group = [R.Explorer(fixed, veh_name="veh%d" % (i + 1), icon_name="veh%d" % (i + 1)) for i in range(n_agv)]
    ref_scene = R.Scene(fixed, ref_group)
    ours_layout = pkg.Layout(*rect, task_list=task_list)
    ours_group = [pkg.Explorer(ours_layout, veh_name="veh%d" % (i + 1), icon_name="veh%d" % (i + 1)) for i in range(n_agv)]
    ours_scene = pkg.Scene(ours_layout, ours_group)
    assert np.array_equal(np.array(ours_layout.layout_original), np.array(fixed.layout_original))
    a = _run_reference_controller(kind, ref_scene, fixed, episodes, seed=5)
    b = _run_reference_controller(kind, ours_scene, ours_layout, episodes, seed=5)
    assert a == b, (a, b)
    assert sum(x[2] for x in a) > episodes  # turns were taken


def test_madqn_controller_mirror_trains_and_reloads(pkg, tmp_path):
    """this package's MADQNAgentController over the facade Scene: train_NN episodes, save, use_NN reload"""
    import torch
    C = importlib.import_module(PKG_NAME + ".algorithm.MADQN_structure.Controller")
    _seed_all(3)
    layout = pkg.Layout(2, 2, 2, 1, 1)
    group = [pkg.Explorer(layout, veh_name="veh%d" % (i + 1), icon_name="veh%d" % (i + 1)) for i in range(2)]
    scene = pkg.Scene(layout, group)
    ctl = C.MADQNAgentController(scene, layout.scene_x_width, layout.scene_y_width, len(layout.storage_station_list),
                                 control_mode="train_NN", state_number=3)
    ctl.storage_path = str(tmp_path)
    ctl.simulation_times = 3
    with contextlib.redirect_stdout(io.StringIO()):
        ctl.model_run()
    assert len(ctl.lifelong_reward) == 3 and os.path.exists(tmp_path / "policy_net.pt") and os.path.exists(tmp_path / "logs.txt")
    w = {k: v.detach().cpu().clone() for k, v in ctl.agent.policy_network.state_dict().items()}
    ctl2 = C.MADQNAgentController.__new__(C.MADQNAgentController)
    ctl2.control_mode, ctl2.state_number, ctl2.rmfs_model = "use_NN", 3, scene
    ctl2.storage_path, ctl2.batched = str(tmp_path), False
    ctl2.device = torch.device("cuda")
    ctl2.create_agent(layout.scene_x_width, layout.scene_y_width)
    for k, v in ctl2.agent.policy_network.state_dict().items():
        assert torch.equal(v.cpu(), w[k]), k


def test_checkpoint_round_trip_and_resume(pkg, tmp_path):
    """BatchedAgent.save writes reference-format whole-module pickles; load restores weights, Adagrad
    accumulators and counters; a resumed agent continues bit-identically"""
    import torch
    M = importlib.import_module(PKG_NAME + ".algorithm.MADQN_structure.MADQN")
    grid = orc.README_9x9
    E, N = 48, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)

    def fresh(seed):
        torch.manual_seed(0)
        sc = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
        sc.reset()
        return sc, M.BatchedAgent(sc, memory_size=4096, batch_size=32, epsilon=0.8, seed=seed, graph_learner=False)

    sc, ag = fresh(1)
    for _ in range(6):
        ag.tick(train=True)
    ag.save(str(tmp_path / "policy_net.pt"), str(tmp_path / "target_net.pt"), str(tmp_path / "state.pt"))
    # the pickled class path is the reference's
    import zipfile
    z = zipfile.ZipFile(tmp_path / "policy_net.pt")
    blob = z.read([n for n in z.namelist() if n.endswith("data.pkl")][0])
    assert b"algorithm.MADQN_structure.MADQN" in blob and b"b200" not in blob
    sc2, ag2 = fresh(2)
    ag2.load(str(tmp_path / "policy_net.pt"), str(tmp_path / "target_net.pt"), str(tmp_path / "state.pt"))
    assert ag2.learn_step_counter == ag.learn_step_counter
    for a, b in zip(ag._params, ag2._params):
        assert torch.equal(a, b)
    for a, b in zip(ag._sums, ag2._sums):
        assert torch.equal(a, b)
    sc.close(); sc2.close()


@needs_ref
def test_reference_loads_our_checkpoint(pkg, tmp_path):
    """a policy_net.pt written here unpickles in a process that only has the reference on its path (use_NN)"""
    import subprocess
    import torch
    ck = importlib.import_module(PKG_NAME + ".algorithm.checkpoint")
    M = importlib.import_module(PKG_NAME + ".algorithm.MADQN_structure.MADQN")
    torch.manual_seed(4)
    net = M.Net()
    path = str(tmp_path / "policy_net.pt")
    ck.save_module(net, path)
    x = torch.arange(2 * 3 * 7 * 7, dtype=torch.float32).reshape(2, 3, 7, 7) % 2
    want = net(x).detach().numpy().tolist()
    code = ("import sys,types,torch,json\n"
            "for m in ('pygame','matplotlib','matplotlib.pyplot'): sys.modules[m]=types.ModuleType(m)\n"
            "sys.modules['matplotlib'].pyplot=sys.modules['matplotlib.pyplot']\n"
            "sys.path[:0]=[%r, %r]\n"
            "m=torch.load(%r, weights_only=False)\n"
            "x=torch.arange(2*3*7*7,dtype=torch.float32).reshape(2,3,7,7)%%2\n"
            "print(json.dumps([type(m).__module__, m(x).detach().numpy().tolist()]))\n") % (
        rh.REF_ROOT, os.path.join(rh.REF_ROOT, "src"), path)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    mod, got = json.loads(out.stdout.strip().splitlines()[-1])
    assert mod == "algorithm.MADQN_structure.MADQN"
    assert np.allclose(np.array(got), np.array(want), atol=1e-6)


@needs_ref
def test_shipped_dqn_checkpoint_q_values(pkg):
    """the reference's shipped RMFS_DQN_policy_net.pt (DQN_structure/network_picture) loads through the alias
    loader; its Q-values on full-map states encoded by the CUDA path, bf16 autocast on the GPU, stay within the
    stated tolerance |dQ| <= 2e-2 * max|Q| of the fp32 CPU forward of the same module"""
    import torch
    ck = importlib.import_module(PKG_NAME + ".algorithm.checkpoint")
    path = os.path.join(rh.REF_ROOT, "src", "algorithm", "DQN_structure", "network_picture", "RMFS_DQN_policy_net.pt")
    if not os.path.exists(path):
        pytest.skip("shipped checkpoint not vendored")
    net = ck.load_module(path)
    assert type(net).__module__ == "DQN_structure.DQN" and net.fc4.out_features == 5
    # fc3 expects 256 * (W//4) * (H//4) = 256 inputs: a 7 x 6 map, rect layout (2, 1, 2, 1, 1)
    grid = orc.rect_layout(2, 1, 2, 1, 1)
    assert grid.shape == (6, 7)
    E, N = 64, 2
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    sc = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
    sc.reset()
    states = []
    for t in range(12):
        o = sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_FULL)
        took = o["act"] >= 0
        states.append(o["obs"][took].clone())
    x = torch.cat(states)[:512]
    assert x.shape[0] >= 100 and set(x.float().unique().tolist()) <= {0.0, 1.0}
    with torch.no_grad():
        q_ref = net(x.float().cpu())
        gnet = importlib.import_module(PKG_NAME + ".algorithm.DQN_structure.DQN").Net(3, 5, 7, 6)
        gnet.load_state_dict(net.state_dict())
        gnet = gnet.cuda().to(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            q_bf16 = gnet(x.contiguous(memory_format=torch.channels_last)).float().cpu()
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False   # true fp32 convolutions for the fp32 leg (TF32 is ~1e-3 off)
        try:
            q_fp32 = gnet(x.float()).cpu()
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
    scale = q_ref.abs().max()
    assert (q_fp32 - q_ref).abs().max() <= 1e-4 * scale
    assert (q_bf16 - q_ref).abs().max() <= 2e-2 * scale
    sc.close()


def test_learner_step_equals_reference_update(pkg):
    """One optimiser step of the batched learner == the reference's update_network (MADQN.py:134-171: DDQN
    target, SmoothL1, torch.optim.Adagrad lr 0.01) on the same sampled batch and the same weights -- eagerly
    and replayed from the CUDA graph.  fp32 without TF32, so gradients compare to float tolerance; the
    Adagrad update is then compared on the SAME gradient (its first steps divide by |g|: a sign flip of a
    gradient element that is zero up to rounding would otherwise move a weight by 2 * lr)."""
    import torch
    M = importlib.import_module(PKG_NAME + ".algorithm.MADQN_structure.MADQN")
    grid = orc.README_9x9
    E, N = 64, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(e)) for e in range(E)], dtype=np.uint8)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for graph in (False, True):
            torch.manual_seed(0)
            sc = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True)
            sc.reset()
            ag = M.BatchedAgent(sc, memory_size=4096, batch_size=32, epsilon=0.8, seed=7, autocast=False,
                                graph_learner=graph, start_training_info_number=10 ** 9)  # fill the replay, no learning
            for _ in range(5):
                ag.tick(train=True)
            assert ag.learn_step_counter == 0 and ag.memory.info()[0] > 200
            pol, tgt = M.Net().cuda(), M.Net().cuda()
            opt = torch.optim.Adagrad(pol.parameters(), 0.01)
            loss_fn = torch.nn.SmoothL1Loss()
            for step in range(6):  # graph learner: 3 eager warm-up steps, the capture + first replay, 2 more replays
                # the reference learner starts every step from the batched learner's state
                pol.load_state_dict(ag.policy_network.state_dict())
                for p_ref, s_ours in zip(pol.parameters(), ag._sums):
                    opt.state[p_ref]["sum"] = s_ours.detach().clone()
                    opt.state[p_ref]["step"] = torch.tensor(float(step))
                ag.update_network()
                tgt.load_state_dict(ag.target_network.state_dict())
                b = ag._sb  # the batch the step just sampled (static buffers)
                bs, bn = b["obs"].float(), b["next_obs"].float()
                ba, br, bd = b["act"].long().unsqueeze(1), b["reward"].unsqueeze(1), b["done"].long().unsqueeze(1)
                q_sa = pol(bs).gather(1, ba)
                nxt = tgt(bn).gather(1, torch.max(pol(bn), 1)[1].unsqueeze(1))
                loss = loss_fn(q_sa, (nxt * ag.GAMMA) * (1 - bd) + br)
                opt.zero_grad()
                loss.backward()
                assert math.isclose(float(ag._loss), float(loss.detach()), rel_tol=1e-4, abs_tol=1e-6), (graph, step)
                g_ref = torch.cat([p.grad.reshape(-1) for p in pol.parameters()])
                assert (g_ref - ag._flat).abs().max() <= 1e-4 * g_ref.abs().max() + 1e-9, (graph, step)
                for p_ref, g_ours in zip(pol.parameters(), ag._views):
                    p_ref.grad.copy_(g_ours)
                opt.step()
                for p_ref, p_ours in zip(pol.parameters(), ag._params):
                    assert torch.allclose(p_ref, p_ours, rtol=1e-6, atol=1e-8), (graph, step)
                for p_ref, s_ours in zip(pol.parameters(), ag._sums):
                    assert torch.allclose(opt.state[p_ref]["sum"], s_ours, rtol=1e-6, atol=1e-12), (graph, step)
            assert (ag._graph is not None) == graph
            sc.close()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_dump_frame_and_set_state_validation(pkg):
    grid = orc.rect_layout(4, 2, 2, 2, 2)
    tasks = np.array([orc.make_tasks(grid, random.Random(1))], dtype=np.uint8)
    sc = pkg.BatchedScene(grid, 1, 4, tasks=tasks)
    sc.reset()
    for _ in range(9):
        sc.tick(pkg.PROTO_ASTAR, pkg.POLICY_ASTAR)
    env = orc.Env(grid, tasks[0], 4)
    for _ in range(9):
        env.tick(orc.PROTO_ASTAR, orc.POLICY_ASTAR)
    fr = sc.dump_frame(0)
    snap = env.snapshot()
    assert fr["running_time"] == 9 and len(fr["all_info"]) - 1 == len(snap)
    for j, row in enumerate(snap):
        name, pos, tgt, loaded = fr["all_info"][1 + j]
        assert name == "veh%d" % (j + 1) and pos + tgt + [int(loaded)] == [int(v) for v in row]
    # agv_set_state rejects snapshots that would index out of bounds
    hdr, agv = sc.get_state()
    for mutate in (lambda h, a: h.__setitem__((0, 4), 5), lambda h, a: h.__setitem__((0, 1), sc.T + 1),
                   lambda h, a: a.__setitem__((0, 0, 7), sc.T), lambda h, a: a.__setitem__((0, 0, 4), 3)):
        h2, a2 = hdr.copy(), agv.copy()
        mutate(h2, a2)
        with pytest.raises(pkg.AgvError):
            sc.set_state(h2, a2)
    sc.set_state(hdr, agv)
    sc.close()
