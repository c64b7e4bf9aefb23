"""exercises the kernels outside the headline rollout once each, for ncu (tuning/r2_ncu_rest.sh):
   python tuning/prof_kernels.py replay|substep|bfs|shaped"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
what = sys.argv[1]
gridf, N, E, _ = bench.WORKLOADS["c3"]
grid = gridf()
if what == "replay":
    # K4: push of one 8-tick Expert rollout (8.4 M rows offered), PER sample + gather of 4096 rows
    E = 4096
    sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1), auto_reset=True)
    sc.reset()
    rp = pkg.DeviceReplay(1 << 21, (3, 7, 7))
    for _ in range(3):
        sc.collect_expert(rp, 8, ticks_per_launch=8)
    u = torch.rand(4096, dtype=torch.float64, device="cuda")
    for _ in range(3):
        b = rp.sample(4096, u)
        rp.gather(b["slot"])
    torch.cuda.synchronize()
    print("replay entries", rp.info())
elif what == "substep":
    # sub-step protocol kernels at c3 scale: k_begin_tick, k_encode, k_bfs_action, k_step_turn, k_end_tick
    sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1), auto_reset=True)
    sc.reset()
    sc.rollout(96, pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, want_lens=False)
    minus1 = torch.full((E,), -1, dtype=torch.int8, device="cuda")
    for _ in range(2):
        sc.begin_tick()
        for k in range(N):
            sc.encode(k, pkg.OBS_CLIP)
            sc.bfs_action(k)
            sc.step_turn(k, pkg.PROTO_INTELLIGENT, minus1)
        sc.end_tick()
    torch.cuda.synchronize()
elif what == "bfs":
    # stand-alone FindPathAstar batch: 65536 problems on c3-like matrices (state matrices with AGV obstacles)
    rng = np.random.default_rng(0)
    B = 65536
    base = (grid != 2).astype(np.uint8)
    valid = np.broadcast_to(base, (B, 64, 64)).copy()
    idx = rng.integers(0, 64, size=(B, 48, 2))
    for b in range(0, B, 1):
        valid[b, idx[b, :, 0], idx[b, :, 1]] = 0
    start = rng.integers(0, 64, size=(B, 2)).astype(np.int32)
    target = rng.integers(0, 64, size=(B, 2)).astype(np.int32)
    v = torch.as_tensor(valid, device="cuda")
    for _ in range(3):
        pkg.bfs_batch(v, start, target)
        pkg.bfs_field(v[:8192], target[:8192])
    torch.cuda.synchronize()
elif what == "shaped":
    sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1), auto_reset=True, shaped_reward=True)
    sc.reset()
    for _ in range(100):
        sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP, want_lens=False)
    torch.cuda.synchronize()
print("done", what)
