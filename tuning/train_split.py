"""where does an AG-DQN training sub-step spend its time? (tuning aid)"""
import importlib, sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
M = importlib.import_module(bench.PKG + ".algorithm.MADQN_structure.MADQN")
gridf, N, E, _ = bench.WORKLOADS["c2"]
grid = gridf()
sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1), auto_reset=True)
sc.reset()
ag = M.BatchedAgent(sc, memory_size=1 << 20, batch_size=256, epsilon=0.8)
for _ in range(12): ag.tick(train=True)
def timeit(fn, n=20):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
obs, _ = sc.encode(0, pkg.OBS_CLIP)
print("tick(train)        gpu %.3f ms  wall %.3f ms" % timeit(lambda: ag.tick(train=True), 10))
print("tick(no train)     gpu %.3f ms  wall %.3f ms" % timeit(lambda: ag.tick(train=False), 10))
print("forward 4096       gpu %.3f ms  wall %.3f ms" % timeit(lambda: ag.q_values(ag.policy_network, obs)))
print("update_network     gpu %.3f ms  wall %.3f ms" % timeit(ag.update_network))
def env_only():
    sc.begin_tick()
    for k in range(sc.N):
        sc.encode(k, pkg.OBS_CLIP); sc.step_turn(k, pkg.PROTO_INTELLIGENT, torch.full((E,), -1, dtype=torch.int8, device=sc.device)); sc.encode(k, pkg.OBS_CLIP, post=True)
    sc.end_tick()
print("env sub-steps only gpu %.3f ms  wall %.3f ms" % timeit(env_only, 10))
