"""per-scene timeline of one agv_rollout launch (needs a -DAGV_ROLL_PROF library: tuning/mkvariant.sh prof
"-DAGV_ROLL_PROF"; run with AGV_B200_LIB=$PWD/tuning/lib_prof.so AGV_ROLL_PROF=1).  Tuning aid."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module(bench.PKG)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
wl = sys.argv[2] if len(sys.argv) > 2 else "c3"
mode = sys.argv[3] if len(sys.argv) > 3 else "guided"   # "expert": A_star control mode + next_obs (collect_expert's rollout)
PROTO, POL, NEXT = ((pkg.PROTO_ASTAR, pkg.POLICY_ASTAR, True) if mode == "expert" else
                    (pkg.PROTO_INTELLIGENT, pkg.POLICY_GUIDED, False))
gridf, N, E, _ = bench.WORKLOADS[wl]
grid = gridf()
sc = pkg.BatchedScene(grid, E, N, tasks=bench.synth_tasks(grid, E, 1000), auto_reset=True)
sc.reset()
out = {}
for _ in range(2 * N // 16 + 1):
    sc.rollout(16, PROTO, POL, pkg.OBS_CLIP, want_next=NEXT, want_lens=False, out=out)
out = {}
for _ in range(2):
    sc.rollout(K, PROTO, POL, pkg.OBS_CLIP, want_next=NEXT, want_lens=False, out=out)
for rep in range(2):
    sc.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sc.rollout(K, PROTO, POL, pkg.OBS_CLIP, want_next=NEXT, want_lens=False, out=out)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if rep == 1:
        sys.stderr.write("PHASES(next line: counters[3..47]; [8+10w+k] = cycles of phase k of scene warp w, +8 iterations, +9 warps)\n")
    turns = sc.counters()[0]
    p = np.zeros((E, 8), dtype=np.uint64)
    rc = sc.lib.agv_rollout_profile(sc._h, p.ctypes.data_as(C.c_void_p), sc._stream())
    assert rc == 0, rc
p = p.astype(np.int64)
t0 = p[:, 4].min()
start, fin = (p[:, 4] - t0) / 1e6, (p[:, 0] - t0) / 1e6   # ms
dur = fin - start
print("K=%d: %.3f ms (%.3f ms/tick), %d turns, %.3e AGV-steps/s; last finish %.3f ms" % (K, ms, ms / K, turns, turns / ms * 1e3, fin.max()))
print("scene duration ms: mean %.3f p50 %.3f p90 %.3f p99 %.3f max %.3f; start max %.3f" % (
    dur.mean(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), start.max()))
for q in (0.25, 0.5, 0.75, 0.9, 0.95, 0.99):
    tq = fin.max() * q
    print("  at %.0f%% of the kernel (%.2f ms): %5d scenes still running" % (q * 100, tq, ((start <= tq) & (fin > tq)).sum()))
bfs, trn, wait, plen = p[:, 1], p[:, 2], p[:, 3], p[:, 7]
print("per scene over K ticks: bfs mean %.1f max %d; turns mean %.0f; wait cycles/bfs mean %.0f" % (bfs.mean(), bfs.max(), trn.mean(), wait.sum() / max(1, bfs.sum())))
# per-CTA finish time (its slowest scene) and how the load spread over CTAs
cta = p[:, 6]
last = np.zeros(int(cta.max()) + 1)
np.maximum.at(last, cta, fin)
print("CTA finish ms: mean %.3f p10 %.3f p50 %.3f p90 %.3f max %.3f" % (last.mean(), np.percentile(last, 10), np.median(last), np.percentile(last, 90), last.max()))
late = fin > np.percentile(fin, 99)
print("latest 1%% of scenes: turns mean %.0f (all %.0f), bfs mean %.1f (all %.1f), pickups mean %.2f, start mean %.3f ms" % (trn[late].mean(), trn.mean(), bfs[late].mean(), bfs.mean(), p[late, 5].mean(), start[late].mean()))
order = np.argsort(-fin)[:12]
for e in order:
    print("  scene %5d: dur %.3f ms (start %.3f) bfs %4d (%.1f/tick) turns %4d wait %.3f ms (%.0f cyc/bfs, mean plen %.0f) own %.0f cyc/turn  pickups %d cta %d" % (
        e, dur[e], start[e], bfs[e], bfs[e] / K, trn[e], wait[e] / 1.965e6, wait[e] / max(1, bfs[e]), plen[e] / max(1, bfs[e]),
        (dur[e] * 1.965e6 - wait[e]) / max(1, trn[e]), p[e, 5], p[e, 6]))
# regression dur ~ a*turns + b*bfs + c*plen
A = np.stack([trn, bfs, plen, np.ones(E)], 1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A, dur * 1.965e6, rcond=None)
print("fit: cycles = %.0f*turns + %.0f*bfs + %.1f*levels + %.0f" % tuple(coef))
light = bfs < 16 * K // 8
print("light scenes (bfs/tick < 2): n %d, cycles per turn %.0f" % (light.sum(), (dur[light] * 1.965e6 / np.maximum(1, trn[light])).mean()))
