"""how much of a congested scene's BFS chain could overlap if the search of AGV k+1 were started
while AGV k's is in flight (valid iff AGV k does not move)?  (tuning aid)"""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
gridf, N, _, _ = bench.WORKLOADS["c3"]
grid = gridf(); E = 16384
sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1), auto_reset=True)
sc.reset()
for _ in range(134): sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP)
hdr, agv = sc.get_state()
o = sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP)
act = o["act"].cpu().numpy(); p = (o["plen"].cpu().numpy().astype(np.int32) & 0xFFFF)
took = act >= 0
md = np.abs(agv[..., 0] - agv[..., 2]) + np.abs(agv[..., 1] - agv[..., 3])
bfs = took & (((act == 4) & (md > 0)) | ((act != 4) & (p != md)))
moved = took & (act != 4)
nb = bfs.sum(1)
heavy = np.argsort(-nb)[:160]
tot = chain1 = chain2 = chain3 = 0
for e in heavy:
    idx = np.where(took[e])[0]
    b = bfs[e][idx]; mv = moved[e][idx]
    tot += b.sum()
    # greedy grouping: a group = maximal run of BFS turns where every member but the last did not move
    def groups(depth):
        g = 0; i = 0
        while i < len(idx):
            if not b[i]: i += 1; continue
            g += 1; d = 1
            while d < depth and i + 1 < len(idx) and b[i + 1] and not mv[i]:
                i += 1; d += 1
            i += 1
        return g
    chain1 += groups(1); chain2 += groups(2); chain3 += groups(3)
print("top-160 scenes: BFS turns %d; sequential BFS latencies with speculation depth 1/2/3: %d / %d / %d" % (tot, chain1, chain2, chain3))
print("worst scene: bfs %d, stuck (no move) among them %d" % (nb[heavy[0]], (bfs[heavy[0]] & ~moved[heavy[0]]).sum()))
