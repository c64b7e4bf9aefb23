"""per-scene BFS fallback load in the c3 benchmark workload (tuning aid): how uneven is the
number / length of level-synchronous searches across scenes?"""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
gridf, N, _, _ = bench.WORKLOADS["c3"]
grid = gridf(); E = 16384
sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1), auto_reset=True)
sc.reset()
for _ in range(128 + 6): sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP)
for t in range(3):
    hdr, agv = sc.get_state()
    o = sc.tick(pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP)
    act = o["act"].cpu().numpy(); p = (o["plen"].cpu().numpy().astype(np.int32) & 0xFFFF)
    took = act >= 0
    md = np.abs(agv[..., 0] - agv[..., 2]) + np.abs(agv[..., 1] - agv[..., 3])
    stop_search = took & (act == 4) & (md > 0)          # no path: BFS runs until the frontier dies
    detour = took & (act != 4) & (p != md)
    lv = np.where(stop_search, 150, 0) + np.where(detour, p, 0)   # rough level count
    per_scene = lv.sum(1)
    nb = (stop_search | detour).sum(1)
    print("tick", t, "turns", took.sum(), "fallback %.3f" % ((stop_search | detour).sum() / took.sum()),
          "nopath %.4f" % (stop_search.sum() / took.sum()))
    print("  BFS per scene: mean %.1f  p50 %d p90 %d p99 %d max %d" % (nb.mean(), np.median(nb), np.percentile(nb, 90), np.percentile(nb, 99), nb.max()))
    print("  levels per scene: mean %.0f p50 %d p90 %d p99 %d max %d" % (per_scene.mean(), np.median(per_scene), np.percentile(per_scene, 90), np.percentile(per_scene, 99), per_scene.max()))
    print("  scenes with >=8 no-path searches: %d ; tick counter of the worst scene %d, created %d" % ((stop_search.sum(1) >= 8).sum(), hdr[per_scene.argmax(), 0], hdr[per_scene.argmax(), 4]))
    w32 = per_scene.reshape(-1, 32).sum(1); print("  per 32-scene warp levels: mean %.0f max %d" % (w32.mean(), w32.max()))
    top = np.argsort(-per_scene)[:6]
    for e in top:
        print("   scene %5d: nopath %2d detour %2d (mean plen %.0f, mean manhattan %.0f) acting %d" % (
            e, stop_search[e].sum(), detour[e].sum(), p[e][detour[e]].mean() if detour[e].any() else 0,
            md[e][detour[e]].mean() if detour[e].any() else 0, took[e].sum()))
    print("   all scenes: nopath %d detour %d, detour levels %d" % (stop_search.sum(), detour.sum(), p[detour].sum()))
