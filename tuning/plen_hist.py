"""distribution of BFS path lengths / STOP rate in the c3 benchmark workload (tuning aid)"""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
gridf, N, _, _ = bench.WORKLOADS["c3"]
grid = gridf(); E = 2048
sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1), auto_reset=True)
sc.reset()
for _ in range(128): sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP)
pl = []; stops = 0; turns = 0; man = []
for _ in range(40):
    hdr, agv = sc.get_state()
    o = sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP)
    act = o["act"].cpu().numpy(); p = (o["plen"].cpu().numpy().astype(np.int32) & 0xFFFF)
    m = act >= 0
    turns += m.sum(); stops += (act == 4).sum(); pl.append(p[m & (act != 4)])
    md = np.abs(agv[..., 0] - agv[..., 2]) + np.abs(agv[..., 1] - agv[..., 3])
    man.append((p - md)[m & (act != 4)])
pl = np.concatenate(pl); man = np.concatenate(man)
print("turns", turns, "STOP frac %.4f" % (stops / turns), "plen mean %.1f median %d p90 %d max %d" % (pl.mean(), np.median(pl), np.percentile(pl, 90), pl.max()))
print("plen - manhattan(at tick start): mean %.2f, ==0: %.3f, <=3: %.3f" % (man.mean(), (man == 0).mean(), (man <= 3).mean()))
