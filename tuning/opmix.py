"""executed-instruction mix of a kernel in an ncu --set full report, by opcode and by code region (CPU box):
   python tuning/opmix.py <report.ncu-rep> <kernel symbol substring>
Joins `ncu --page source --csv` (Instructions Executed per SASS instruction) with `nvdisasm -g` of the in-tree
library (same instruction order)."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, sym = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("AGV_B200_LIB", os.path.join(root, "reinforcement-learning-for-multi-agv-pathfinding_b200", "libagv_b200.so"))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, "agv_rollout.sm_100a.cubin")], capture_output=True, text=True).stdout
ins, cur, on = [], ("?", 0), False
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
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m:
        ins.append((cur, m.group(1)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
assert len(data) == len(ins), (len(data), len(ins))
ALU = {"LOP3", "SHF", "IADD3", "ISETP", "SEL", "LEA", "PLOP3", "PRMT", "IABS", "VIADD", "VIMNMX", "VIMNMX3", "FLO", "BREV", "POPC", "SGXT", "BMSK", "IMNMX", "MOV", "CS2R", "P2R", "R2P", "FSEL", "FSETP"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "IDP"}
def region(f, l, text):
    return f
byop, bypipe, byfile = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
tot = 0
for ((f, l), op), r in zip(ins, data):
    n = int(r[ix["Instructions Executed"]])
    tot += n; byop[op] += n
    pipe = "alu" if op in ALU else ("fma" if op in FMA else ("lsu" if op[:2] in ("LD", "ST", "AT", "RE") else "other"))
    bypipe[pipe] += n; byfile[(f, l // 25 * 25)][pipe] += n
print("warp instructions executed: %.3e" % tot)
print("by pipe:", {k: "%.1f%%" % (100 * v / tot) for k, v in bypipe.most_common()})
print("by opcode:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in byop.most_common(24)))
print("ALU-pipe instructions by source region (file : 25-line block):")
for k, v in sorted(byfile.items(), key=lambda kv: -kv[1]["alu"])[:28]:
    print("  %5.1f%% of all instr on alu, %5.1f%% total   %s:%d" % (100 * v["alu"] / tot, 100 * sum(v.values()) / tot, k[0], k[1]))
