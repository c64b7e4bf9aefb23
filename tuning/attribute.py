"""Attribute executed warp-instructions of one profiled kernel to source functions.

  ncu -i X.ncu-rep --page source --csv > src.csv          (SASS rows, in address order)
  nvdisasm -g -fun <symidx> kernels.cubin > k.sass         (same SASS with //## File/line marks)
  python tuning/attribute.py src.csv k.sass <text-section-name> [turns]

Lines of agv_core.cuh are mapped to the enclosing function by scanning the header for
`__device__` definitions; the innermost inlined frame is what nvdisasm reports.
"""
import csv
import os
import re
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
CORE = os.path.join(HERE, "..", "reinforcement-learning-for-multi-agv-pathfinding_b200", "csrc", "agv_core.cuh")


def func_ranges(path):
    """[(first_line, name)] of every function definition, by a crude regex"""
    out = []
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:__device__|static|inline|__forceinline__|\s)+.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$")
    for n, line in enumerate(open(path), 1):
        if "__device__" in line and "(" in line and not line.strip().startswith("//"):
            m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*\(", line.split("__forceinline__")[-1])
            if m:
                out.append((n, m.group(1)))
    return out


def main():
    src_csv, sass, sect = sys.argv[1], sys.argv[2], sys.argv[3]
    turns = float(sys.argv[4]) if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("Source")
    cst = hdr.index("Warp Stall Sampling (All Samples)")
    prof = [(r[cs].strip(), int(r[ci]), int(r[cst])) for r in rows[hdr_i + 1:] if len(r) > ci]
    # instruction -> (file, line)
    marks = []
    cur = ("?", 0)
    on = False
    for line in open(sass):
        if line.startswith(".text.") and sect in line:
            on = True
            continue
        if on and line.startswith("//-----"):
            break
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            marks.append(cur)
    assert len(marks) == len(prof), (len(marks), len(prof))
    fr = func_ranges(CORE)

    def func_of(f, ln):
        if f != "agv_core.cuh":
            return f
        name = "?"
        for first, nm in fr:
            if first <= ln:
                name = nm
            else:
                break
        return name

    by_func, by_line, st_func = defaultdict(int), defaultdict(int), defaultdict(int)
    total = 0
    tstall = 0
    for (f, ln), (_, n, st) in zip(marks, prof):
        by_func[func_of(f, ln)] += n
        st_func[func_of(f, ln)] += st
        by_line[(f, ln)] += n
        total += n
        tstall += st
    print("total warp-instructions %d%s" % (total, "  (%.0f per turn)" % (total / turns) if turns else ""))
    print("%-28s %14s %6s %8s %7s" % ("function", "instr", "%", "/turn", "stall%"))
    for k, v in sorted(by_func.items(), key=lambda kv: -kv[1]):
        print("%-28s %14d %6.2f %8.1f %7.2f" % (k, v, 100.0 * v / total, v / turns if turns else 0, 100.0 * st_func[k] / max(1, tstall)))
    print("\ntop source lines")
    for (f, ln), v in sorted(by_line.items(), key=lambda kv: -kv[1])[:40]:
        print("%-18s %5d %14d %6.2f" % (f, ln, v, 100.0 * v / total))


if __name__ == "__main__":
    main()
