"""per-source-line attribution of an ncu --set full report (CPU box):
   python tuning/lineprof.py <report.ncu-rep> <kernel symbol> [cubin name] [top N]
Joins `ncu --page source --csv` (samples / executed per SASS instruction) with `nvdisasm -g` line info of
the in-tree library's cubin (the i-th instruction of both listings is the same instruction)."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, sym = sys.argv[1], sys.argv[2]
cub = sys.argv[3] if len(sys.argv) > 3 else "agv_rollout"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("AGV_B200_LIB", os.path.join(root, "reinforcement-learning-for-multi-agv-pathfinding_b200", "libagv_b200.so"))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub + ".sm_100a.cubin")], capture_output=True, text=True).stdout
lines, cur, on = [], ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        on = sym in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
assert len(data) == len(lines), (len(data), len(lines))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for (f, l), r in zip(lines, data):
    a = agg[(f, l)]
    a[0] += int(r[ix["# Samples"]]); a[1] += int(r[ix["Instructions Executed"]]); a[2] += 1
    for s in stalls:
        a[3][s[6:]] += int(r[ix[s]] or 0)
tot = sum(a[0] for a in agg.values())
srcs = {}
def text(f, l):
    if f not in srcs:
        p = os.path.join(root, "reinforcement-learning-for-multi-agv-pathfinding_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ""
print("total samples", tot)
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = " ".join("%s=%d%%" % (k, 100 * v // max(1, a[0])) for k, v in a[3].most_common(3))
    print("%5.1f%% %-16s:%4d sass %3d exec %.2e  %-34s | %s" % (100 * a[0] / tot, f, l, a[2], a[1], st, text(f, l)))
