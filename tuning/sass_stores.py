"""lists the global stores / atomics of the headline rollout kernel and its out-of-line record expander by source
line (CPU box; cuobjdump + nvdisasm -g of the in-tree library):   python tuning/sass_stores.py > profiles/r02_sass_stores_k_rollout.txt"""
import collections, os, re, subprocess, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "reinforcement-learning-for-multi-agv-pathfinding_b200", "libagv_b200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, "agv_rollout.sm_100a.cubin")], capture_output=True, text=True).stdout
print("global stores of k_rollout<1,2,OBS_CLIP,32,3,MODE 1> and its out-of-line record expander (roll_expand), by source line")
print("(cuobjdump / nvdisasm -g of the in-tree libagv_b200.so; the observation path = agv_rollout.cuh roll_expand: STG.E.128 only,")
print(" plus the per-record fallback expand_clip7 (agv_lanes.cuh) that runs only when N % 8 != 0)\n")
src_cache = {}
def src(path, line):
    if path not in src_cache:
        try:
            src_cache[path] = open(path).read().splitlines()
        except OSError:
            src_cache[path] = []
    L = src_cache[path]
    return L[line - 1].strip()[:90] if 0 < line <= len(L) else ""
on, cur, seen, order = False, ("?", 0), collections.Counter(), []
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        on = ("k_rolloutILi1ELi2ELi1ELi32ELi3ELi1E" in ln) or ("roll_expand" in ln)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?((?:STG|ATOMG|RED\b)[A-Z0-9_.]*)", ln)
    if m:
        key = (m.group(1), cur)
        if key not in seen:
            order.append(key)
        seen[key] += 1
for op, (path, line) in order:
    print("%-12s x%d  %-16s: %d | %s" % (op, seen[(op, (path, line))], os.path.basename(path), line, src(path, line)))
