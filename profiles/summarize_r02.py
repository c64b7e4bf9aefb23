"""Summarises the round-2 ncu raw-page CSVs (gpurun_out/r02_*.raw.csv, made by tuning/r2_ncu_rest.sh) into
profiles/r02_ncu_kernels.txt and profiles/rollout_<workload>.json (read by bench.py).  CPU box."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6545.0)
KEYS = {"dur_us": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
        "inst": "smsp__inst_executed.sum", "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1_st_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_global_op_st.sum",
        "st_inst": "smsp__inst_executed_op_global_st.sum", "grid": "launch__grid_size", "block": "launch__block_size",
        "smem": "launch__shared_mem_per_block_dynamic"}


def rows(name):
    p = os.path.join(OUT, "r02_%s.raw.csv" % name)
    if not os.path.exists(p):
        return []
    r = list(csv.reader(open(p)))
    hdr, units = r[0], r[1]
    ix = {n: i for i, n in enumerate(hdr)}
    out = []
    for line in r[2:]:
        d = {"kernel": line[ix["Kernel Name"]]}
        for k, m in KEYS.items():
            if m in ix:
                v, u = line[ix[m]], units[ix[m]]
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    continue
                if k == "dur_us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                if k in ("dram_rd", "dram_wr"):
                    v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
                d[k] = v
        out.append(d)
    return out


def short(k):
    k = k.replace("void agv::", "").replace("agv::", "")
    return k.split("(")[0][:60]


lines = ["round 2 -- ncu --set full --clock-control none, one launch each (tuning/r2_ncu_rest.sh, tuning/prof_kernels.py)",
         "HBM peak for the fractions: %.0f GB/s (MEASURED_PEAKS.json); times are ncu's serialised, cold-cache launch times" % PEAK,
         "",
         "%-58s %9s %9s %9s %7s %6s %6s %5s" % ("kernel", "time us", "DRAM MB", "GB/s", "of peak", "issue%", "warps%", "regs")]
for name, title in (("c3", "k_rollout, c3 (64x64, 16384 scenes, 16 ticks, clip obs, guided)"),
                    ("c5", "k_rollout, c5 (128x128, 16384 scenes, 8 ticks)"),
                    ("shaped", "k_tick_lanes, c3 shaped reward (one tick)"),
                    ("replay", "replay: push of an 8-tick Expert rollout of 4096 scenes (2.1 M rows offered), sample + gather 4096"),
                    ("substep", "sub-step protocol kernels, c3 (16384 scenes, one AGV index)"),
                    ("bfs", "stand-alone BFS: 65536 problems 64x64 / 8192 distance fields")):
    rs = rows(name)
    if not rs:
        continue
    lines += ["", "# " + title]
    for d in rs:
        tot = d.get("dram_rd", 0) + d.get("dram_wr", 0)
        gbs = tot / (d["dur_us"] * 1e-6) / 1e9 if d.get("dur_us") else 0
        lines.append("%-58s %9.1f %9.2f %9.1f %7.3f %6.1f %6.1f %5.0f" % (short(d["kernel"]), d.get("dur_us", 0), tot / 1e6, gbs,
                                                                          gbs / PEAK, d.get("issue_pct", 0), d.get("warps_pct", 0), d.get("regs", 0)))
open(os.path.join(ROOT, "profiles", "r02_ncu_kernels.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

# bench.py's roofline inputs for the rollout kernel
for wl, ticks in (("c3", 16), ("c5", 8)):
    rs = rows(wl)
    if not rs:
        continue
    d = rs[0]
    turns = None
    try:  # AGV-steps of the profiled launch: turns_per_step of the plain run of the same command
        txt = open(os.path.join(OUT, "ncu_plain_%s.log" % wl)).read().strip().splitlines()[-1]
        turns = json.loads(txt)["turns_per_step"]
    except Exception:
        pass
    rec = {"kernel": short(d["kernel"]), "ticks_per_launch": ticks, "dram_bytes_per_launch": d.get("dram_rd", 0) + d.get("dram_wr", 0),
           "warp_instr_per_launch": d.get("inst"), "agv_steps_per_launch": turns,
           "warp_instr_per_agv_step": (d["inst"] / turns) if (turns and d.get("inst")) else None,
           "issue_active_pct": d.get("issue_pct"), "alu_pipe_pct": d.get("alu_pct"), "warps_active_pct": d.get("warps_pct"),
           "duration_us_under_ncu": d.get("dur_us"), "registers": d.get("regs"),
           "source": "ncu --set full --clock-control none of one k_rollout launch, profiles/r02_ncu_kernels.txt (smsp__inst_executed.sum / "
                     "AGV-steps of the launch; dram__bytes_read.sum + dram__bytes_write.sum)"}
    json.dump(rec, open(os.path.join(ROOT, "profiles", "rollout_%s.json" % wl), "w"), indent=1)
    print(wl, rec)
