"""formats the 'agv prof:' counter dump of a -DAGV_ROLL_PROF run (stdin passthrough + per-iteration phase cycles)"""
import sys
names = ["0 idle sleep", "1 slow section", "2 B finish turn", "3 C skip", "4 -", "5 load+window",
         "6 search misc+post", "7 row DP"]
for ln in sys.stdin:
    if ln.startswith("agv prof:"):
        v = [int(x) for x in ln.split()[2:]]          # v[j] = counters[3 + j]
        c = lambda i: v[i - 3]
        for w in (0, 1):
            it = max(1, c(8 + 10 * w + 8)); nw = max(1, c(8 + 10 * w + 9))
            tot = sum(c(8 + 10 * w + k) for k in range(8))
            print("scene warp %d: %d warps, %.0f iterations/warp, %.0f cycles/iteration, DP rows/iteration %.1f" % (
                w, nw, it / nw, tot / it, c(28 + w) / it))
            print("   " + "  ".join("%s: %.0f" % (names[k], c(8 + 10 * w + k) / it) for k in range(8)))
    else:
        sys.stdout.write(ln)
