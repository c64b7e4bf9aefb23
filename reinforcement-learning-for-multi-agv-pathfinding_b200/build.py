"""In-tree build of the C-ABI library (libagv_b200.so) with nvcc for sm_100a.

The built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libagv_b200.so")
SOURCES = ["agv_kernels.cu", "agv_lanes.cu", "agv_rollout.cu", "agv_replay.cu"]
HEADERS = ["agv_core.cuh", "agv_lanes.cuh", "agv_async.cuh", "agv_rollout.cuh", os.path.join("..", "..", "include", "agv_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(HERE, "build")


def _nvcc():
    n = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(n):
        raise RuntimeError("nvcc not found; the agv_b200 CUDA library cannot be built")
    return n


def hash_flags(flags) -> int:
    import zlib
    return zlib.crc32(" ".join(flags).encode()) & 0xFFFFFFFF


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if (not force and not needs_build() and not os.environ.get("AGV_NVCC_EXTRA") and _last_flags() == ""
            and not os.environ.get("AGV_BUILD_OUT")):
        return LIB
    # one object per translation unit, compiled in parallel and only when stale, then one link
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    extra = os.environ.get("AGV_NVCC_EXTRA", "").split()
    # objects of different flag sets (tuning / profiling builds) live side by side
    obj_dir = os.path.join(OBJ_DIR, "default" if not extra else "x%08x" % (hash_flags(extra)))
    os.makedirs(obj_dir, exist_ok=True)
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(obj_dir, s.replace(".cu", ".o"))
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t,
                                                                              os.path.getmtime(os.path.abspath(__file__)))
        if stale:
            cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for s, p in jobs:
        out, err = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out + err)
            raise RuntimeError("nvcc failed compiling %s" % s)
        if verbose:
            sys.stderr.write(err)
    objs = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s in SOURCES]
    out = os.environ.get("AGV_BUILD_OUT") or LIB   # tuning variants are linked elsewhere (tuning/mkvariant.sh)
    r = subprocess.run([_nvcc(), "-shared", "-o", out] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking %s" % out)
    if out == LIB:
        with open(os.path.join(OBJ_DIR, "last_flags"), "w") as f:
            f.write(" ".join(extra))
    return out


def _last_flags() -> str:
    """flag set the in-tree .so was last linked from ("" = the default build)"""
    try:
        with open(os.path.join(OBJ_DIR, "last_flags")) as f:
            return f.read().strip()
    except OSError:
        return ""


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
