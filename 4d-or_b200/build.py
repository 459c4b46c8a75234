"""Builds ``libsg4d.so`` (the C-ABI CUDA library, sm_100a only) in-tree with plain nvcc.

    python 4d-or_b200/build.py [--force] [--verbose]

No torch headers are involved: the library's boundary is ``include/sg4d.h`` (raw pointers + ints).
The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libsg4d.so")
SOURCES = ["fps.cu", "ball_query.cu", "group.cu", "gnn.cu", "mlp.cu", "spatial.cu", "interpolate.cu", "frontend.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "sg4d.h"))
    return hdrs


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False, debug=False):
    """debug=True: a separate libsg4d_dbg.so with -DSG4D_DEBUG (ablation switches + timeline tracing of the MLP kernels;
    developer tools select it with SG4D_LIBRARY=<path>)."""
    global OBJ, SO
    flags = list(FLAGS)
    if debug:
        OBJ, SO = os.path.join(HERE, "build_dbg"), os.path.join(HERE, "libsg4d_dbg.so")
        flags.append("-DSG4D_DEBUG")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _deps()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(SO, objs):
        run([NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, debug="--debug" in sys.argv))
