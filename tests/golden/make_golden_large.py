"""Generates tests/golden/bfs_large.json by running the LIVE, unmodified Python reference's
FindPathAstar (src/utils/astar.py:53-153) at BASELINE sizes: 64x64 (configs[2]) and 128x128
(configs[4]).  One call costs ~0.5 s at 64x64 and several seconds at 128x128 (O(V^2) list scans,
astar.py:95-98,109-117), hence a separate, smaller fixture.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_large.py
Test infrastructure (see oracle/agv_oracle.c header).  Matrices are stored one string of
'0'/'1' per row.
"""
from __future__ import annotations

import json
import os
import random
import sys
import time
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ACT = {"UP": 0, "RIGHT": 1, "DOWN": 2, "LEFT": 3, "STOP": 4}


def rect_codes(w, h, nx, ny, ps):
    from oracle.oracle import rect_layout
    return rect_layout(w, h, nx, ny, ps)


def layout_128():
    g = np.zeros((128, 128), dtype=np.uint8)
    for iy in range(41):
        for dy in range(2):
            for ix in range(18):
                for dx in range(6):
                    g[iy * 3 + dy + 1, ix * 7 + dx + 1] = 1
    for i in range(16):
        g[126, int(128 * (i + 1) / 17 + 1) - 1] = 2
    return g


def warehouse_case(codes, rng, n_agv, loaded):
    """matrix (b) of StateManager.create_path_matrix (StateManager.py:43-61) for a random AGV
    population on a warehouse layout: track 1, storage 1 iff not loaded, picking 0, own cell and
    target 1, other AGVs 0"""
    H, W = codes.shape
    valid = ((codes == 0) | ((codes == 1) & (not loaded))).astype(np.uint8)
    track = np.argwhere(codes == 0)
    if rng.random() < 0.5:   # a congested pocket: AGVs concentrated in one region
        cy, cx = track[rng.randrange(len(track))]
        near = track[(np.abs(track[:, 0] - cy) + np.abs(track[:, 1] - cx)) < rng.choice([6, 10, 16])]
        pool = near if len(near) >= n_agv else track
    else:
        pool = track
    idx = rng.sample(range(len(pool)), min(n_agv, len(pool)))
    cells = [tuple(pool[i]) for i in idx]
    for (y, x) in cells[1:]:
        valid[y, x] = 0
    sy, sx = cells[0]
    stations = np.argwhere(codes == (1 if rng.random() < 0.6 else 2))
    ty, tx = stations[rng.randrange(len(stations))]
    valid[sy, sx] = 1
    valid[ty, tx] = 1
    for (y, x) in cells[1:]:
        valid[y, x] = 0
    return valid, (int(sx), int(sy)), (int(tx), int(ty))


def random_case(H, W, rng, dens):
    valid = (np.array([[0 if rng.random() < dens else 1 for _ in range(W)] for _ in range(H)], dtype=np.uint8))
    s = (rng.randrange(W), rng.randrange(H))
    t = (rng.randrange(W), rng.randrange(H))
    return valid, s, t


def solve(job):
    from oracle import ref_harness as rh
    R = rh.load()
    name, valid, s, t = job
    t0 = time.time()
    f = R.astar.FindPathAstar([[float(v) for v in row] for row in valid.tolist()], s, t)
    found, _pl, _pm, al = f.run_astar_method()
    return {"name": name, "H": int(valid.shape[0]), "W": int(valid.shape[1]),
            "valid": ["".join(str(int(v)) for v in row) for row in valid],
            "start": list(s), "target": list(t), "found": bool(found), "actions": [ACT[a] for a in al],
            "ref_seconds": round(time.time() - t0, 2)}


def main():
    rng = random.Random(20261020)
    jobs = []
    c3 = rect_codes(6, 2, 9, 20, 8)            # bench.py workload c3 (64 x 64)
    assert c3.shape == (64, 64)
    for k in range(40):
        v, s, t = warehouse_case(c3, rng, rng.choice([8, 24, 48, 64]), loaded=bool(k & 1))
        jobs.append(("c3_warehouse_%d" % k, v, s, t))
    for k in range(12):
        v, s, t = random_case(64, 64, rng, rng.choice([0.0, 0.1, 0.25, 0.35]))
        jobs.append(("rand64_%d" % k, v, s, t))
    v = np.ones((64, 64), dtype=np.uint8)
    jobs.append(("open64_corner", v, (0, 0), (63, 63)))
    jobs.append(("open64_same", v, (5, 9), (5, 9)))
    v2 = v.copy(); v2[30, :] = 0                 # a wall: unreachable
    jobs.append(("wall64_unreachable", v2, (3, 3), (60, 60)))
    v3 = v.copy(); v3[30, :] = 0; v3[30, 63] = 1  # a wall with one gap at the far side
    jobs.append(("wall64_gap", v3, (0, 29), (0, 31)))
    c5 = layout_128()
    for k in range(5):
        vv, s, t = warehouse_case(c5, rng, 64, loaded=bool(k & 1))
        jobs.append(("c5_warehouse_%d" % k, vv, s, t))
    vv, s, t = random_case(128, 128, rng, 0.2)
    jobs.append(("rand128_0", vv, s, t))
    jobs.append(("open128_corner", np.ones((128, 128), dtype=np.uint8), (127, 0), (0, 127)))
    t0 = time.time()
    with ProcessPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        cases = list(ex.map(solve, jobs))
    path = os.path.join(HERE, "bfs_large.json")
    with open(path, "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("wrote %s: %d cases, %d bytes, %.0f s; found %d, lengths %s" % (
        path, len(cases), os.path.getsize(path), time.time() - t0, sum(c["found"] for c in cases),
        sorted(len(c["actions"]) for c in cases)))


if __name__ == "__main__":
    main()
