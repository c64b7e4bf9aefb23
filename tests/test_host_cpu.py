"""CPU-only checks of the host side: the C-ABI library exports what include/agv_b200.h
declares, Layout geometry / task RNG parity with the reference fixtures, the Expert
CSV wire format, constants, bench helpers, and the multi-rank plumbing over gloo."""
import csv
import json
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, PKG_NAME, ROOT, golden, load_pkg


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    hdr = open(os.path.join(ROOT, "include", "agv_b200.h")).read()
    declared = set(re.findall(r"\b(agv_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib_path = pkg._lib.LIB_PATH
    if not os.path.exists(lib_path):
        import __graft_entry__ as g
        g.build()
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (agv_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert declared == set(pkg._lib.SYMBOLS), declared ^ set(pkg._lib.SYMBOLS)
    lib = pkg._lib.lib()  # loads without a GPU; no compute call is made
    assert lib.agv_version().startswith(b"agv_b200")
    assert lib.agv_strerror(-4).decode().startswith("size outside")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg = load_pkg()
    with pytest.raises(pkg.AgvError):
        pkg.BatchedScene(np.zeros((4, 4), dtype=np.uint8), 1, 1)
    with pytest.raises(pkg.AgvError):
        pkg.DeviceReplay(8, (3, 7, 7))


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "reinforcement-learning-for-multi-agv-pathfinding_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_layout_matches_reference_fixtures():
    pkg = load_pkg()
    for c in golden("astar_traj.json"):
        random.seed(c["seed"])
        if c["name"].startswith("C2"):
            from oracle.oracle import README_9x9
            lay = pkg.Layout(layout_list=README_9x9.tolist())
        elif c["name"].startswith("C1"):
            lay = pkg.Layout(4, 2, 2, 2, 2)
        elif c["name"] == "rect_3_2_2_2_2_n1":
            lay = pkg.Layout(3, 2, 2, 2, 2)
        else:
            lay = pkg.Layout(2, 1, 2, 2, 1)
        assert lay.layout_original == c["grid"]
        assert [list(t) for t in lay.task_list] == c["tasks"]
        assert lay.scene_x_width == len(c["grid"][0]) and lay.scene_y_width == len(c["grid"])
    lay = pkg.Layout(4, 2, 2, 2, 2, task_list=[(2, 2, 4, 9)])
    assert lay.task_is_fixed and lay.task_list == [(2, 2, 4, 9)]
    lay.init()
    assert lay.task_list == [(2, 2, 4, 9)]
    assert lay.picking_station_list[0].x_position == 4 and lay.storage_station_list[0].y_position == 2
    # a layout_list with an obstacle code: only 1/2 are stations
    lay = pkg.Layout(layout_list=[[0, 3, 0], [1, 0, 2]])
    assert lay.layout_original == [[0, 0, 0], [1, 0, 2]]


def test_direction_and_flags():
    pkg = load_pkg()
    d = pkg.Direction()
    assert [d.action_str_value(n) for n in ("UP", "RIGHT", "DOWN", "LEFT", "STOP")] == [0, 1, 2, 3, 4]
    assert d.action_value_str(3) == "LEFT" and d.action_num() == 4
    with pytest.raises(KeyError):
        d.action_value_str(7)
    assert pkg.SuperParas.Sparse_Reward is False and pkg.SuperParas.Explorer_Always_Loaded is False


def test_expert_csv_format_parses_reference_file():
    """the reference-written CSV round-trips through our reader (and json/csv only)"""
    pkg = load_pkg()
    Expert = pkg.Expert
    for tag in ("a", "b"):
        path = os.path.join(GOLDEN, "expert_ref_%s.csv" % tag)
        rows = list(csv.reader(open(path, encoding="utf-8")))
        assert rows[0] == ["layout", "info", "veh_name", "action"]
        for row in rows[1:]:
            layout = json.loads(row[0])
            infos = Expert.data_restore(row[1], layout)
            names = json.loads(row[2].replace("'", '"'))
            actions = json.loads(row[3])
            assert len(infos) == len(names) == len(actions) > 0
            assert infos[0][0] == layout and infos[0][1][0] == "veh1" and infos[0][1][1] == [1, 1]
            assert all(a in (0, 1, 2, 3, 4) for a in actions)
    assert Expert.create_file_name([3, 2, 2, 2, 1, 1]) == "expert_data_3_2_2_2_1_1.csv"


def test_bench_helpers():
    import bench
    g = bench.rect_layout(6, 2, 9, 20, 8)
    from oracle import oracle as orc
    assert np.array_equal(g, orc.rect_layout(6, 2, 9, 20, 8)) and g.shape == (64, 64)
    t = bench.synth_tasks(g, 3, 0)
    assert t.shape == (3, 2160, 4)
    for e in range(3):
        assert len({(int(a), int(b)) for a, b in t[e, :, :2]}) == 2160  # every storage station once
        assert (g[t[e, :, 1] - 1, t[e, :, 0] - 1] == 1).all() and (g[t[e, :, 3] - 1, t[e, :, 2] - 1] == 2).all()
    g5 = bench.layout_128()
    assert g5.shape == (128, 128) and (g5 == 2).sum() == 16
    # every storage station touches a track cell (README.md:82)
    ys, xs = np.nonzero(g5 == 1)
    pad = np.pad(g5, 1, constant_values=1)
    assert all((pad[y, x + 1] == 0) or (pad[y + 2, x + 1] == 0) or (pad[y + 1, x] == 0) or (pad[y + 1, x + 2] == 0)
               for y, x in zip(ys, xs))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    du = __import__("importlib").import_module("reinforcement-learning-for-multi-agv-pathfinding_b200.dist_util")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = du.shard_range(10, rank, world)
    ms, units = du.reduce_timing(10.0 + rank, 100 * (rank + 1))
    du.barrier()
    q.put((rank, lo, hi, ms, units))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction_gloo():
    import torch.multiprocessing as mp
    pkg = load_pkg()
    du = pkg.dist_util if hasattr(pkg, "dist_util") else __import__("importlib").import_module(
        "reinforcement-learning-for-multi-agv-pathfinding_b200.dist_util")
    assert du.shard_range(10, 0, 3) == (0, 4) and du.shard_range(10, 2, 3) == (7, 10)
    assert du.reduce_timing(3.0, 5) == (3.0, 5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert [r[1:3] for r in res] == [(0, 5), (5, 10)]
    assert all(r[3] == 11.0 and r[4] == 300 for r in res)


def test_checkpoint_aliases_pickle_under_the_reference_class_path(tmp_path):
    """whole-module pickles are written under the reference's class path and load back without the
    reference on sys.path (algorithm/checkpoint.py); CPU only"""
    import importlib
    import zipfile
    import torch
    ck = importlib.import_module(PKG_NAME + ".algorithm.checkpoint")
    M = importlib.import_module(PKG_NAME + ".algorithm.MADQN_structure.MADQN")
    torch.manual_seed(1)
    net = M.Net()
    p = str(tmp_path / "policy_net.pt")
    ck.save_module(net, p)
    z = zipfile.ZipFile(p)
    blob = z.read([n for n in z.namelist() if n.endswith("data.pkl")][0])
    assert b"algorithm.MADQN_structure.MADQN" in blob and PKG_NAME.encode() not in blob
    assert "algorithm.MADQN_structure.MADQN" not in sys.modules  # the aliases are gone again
    back = ck.load_module(p)
    x = torch.rand(3, 3, 7, 7)
    assert torch.equal(net(x), back(x))
    D = importlib.import_module(PKG_NAME + ".algorithm.DQN_structure.DQN")
    dnet = D.Net(3, 5, 7, 6)
    p2 = str(tmp_path / "dqn.pt")
    ck.save_module(dnet, p2, ref_module="DQN_structure.DQN")
    assert torch.equal(ck.load_module(p2)(torch.zeros(1, 3, 6, 7)), dnet(torch.zeros(1, 3, 6, 7)))


def _grad_worker(rank, world, port, q):
    import importlib
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = importlib.import_module(PKG_NAME + ".algorithm.MADQN_structure.MADQN")
    torch.manual_seed(0)
    pol, tgt = M.Net(), M.Net()
    g = torch.Generator().manual_seed(5)
    B = 16
    obs = (torch.rand(B, 3, 7, 7, generator=g) > 0.5).float()
    nxt = (torch.rand(B, 3, 7, 7, generator=g) > 0.5).float()
    act = torch.randint(0, 4, (B, 1), generator=g)
    rew = torch.randint(-1, 2, (B, 1), generator=g).float()
    done = (torch.rand(B, 1, generator=g) > 0.8).float()
    lo, hi = rank * B // world, (rank + 1) * B // world
    loss = M.ddqn_loss(pol, tgt, obs[lo:hi], act[lo:hi], rew[lo:hi], nxt[lo:hi], done[lo:hi], 0.95, torch.nn.SmoothL1Loss())
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in pol.parameters()])
    dist.all_reduce(flat)
    flat /= world                       # BatchedAgent._adagrad_step: mean over ranks of the flat gradient
    if rank == 0:
        pol.zero_grad()
        M.ddqn_loss(pol, tgt, obs, act, rew, nxt, done, 0.95, torch.nn.SmoothL1Loss()).backward()
        full = torch.cat([p.grad.reshape(-1) for p in pol.parameters()])
        q.put(float((flat - full).abs().max() / full.abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_mean_equals_single_rank_gradient():
    """the multi-rank learner's gradient -- per-rank DDQN loss on its half of the batch, all-reduced flat
    gradient divided by the world size -- equals the single-rank gradient on the concatenated batch
    (world_size 2, gloo, CPU)"""
    import multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    err = q.get(timeout=300)
    for p in ps:
        p.join(timeout=120)
    assert err < 1e-5, err


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the driver's second arm) runs without a GPU: the unmodified Python
    reference from baseline/_ref (or /root/reference) when present, else the oracle C port, and prints one
    JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--ref-seconds", "0.5", "--workload", "c2"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "AGV-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["agvs_per_scene"] == 4 and line["config"]["scenes_per_gpu"] == 4096
