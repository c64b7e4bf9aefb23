"""how the c3 workload drifts with scene age: ms per tick and turns per tick of successive 8-tick rollouts from reset"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
gridf, N, E, _ = bench.WORKLOADS["c3"]
grid = gridf()
sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1000), auto_reset=True)
sc.reset()
out = {}
K = 8
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    sc.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sc.rollout(K, pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, pkg.OBS_CLIP, want_lens=False, out=out)
    ev1.record(); torch.cuda.synchronize()
    t, e, _ = sc.counters()
    ms = ev0.elapsed_time(ev1)
    if s % 5 == 4 or s < 3:
        hdr, _ = sc.get_state()
        print("ticks %4d..%4d: %.3f ms/tick  %7.0f turns/tick  %.3e AGV-steps/s  episodes ended/tick %.0f  mean scene age %.0f  created mean %.1f" % (
            s * K, s * K + K, ms / K, t / K, t / ms * 1e3, e / K, hdr[:, 0].mean(), hdr[:, 4].mean()))
