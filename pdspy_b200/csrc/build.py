"""Build libpdsb.so (the C-ABI library, include/pdsb.h) with nvcc for sm_100a, in-tree.

    python pdspy_b200/csrc/build.py [--force]

The .so lands next to the Python package (pdspy_b200/libpdsb.so) so that it travels with
the repo snapshot to the GPU box; it is git-ignored.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libpdsb.so")
SOURCES = ["runtime.cu", "dft.cu", "dft_tc5.cu", "dft_f64.cu", "trift.cu", "vis.cu", "grid.cu", "grid_fast.cu", "fft.cu", "cube.cu", "clean.cu"]
HEADERS = ["common.cuh", "dft.cuh", "grid.cuh", "fft_r16.cuh", os.path.join("..", "..", "include", "pdsb.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("PDSB_NVCC_EXTRA", "").split()          # e.g. -DPDSB_TC5_PROBES for timing experiments
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libpdsb")
    tmp = OUT + ".tmp%d" % os.getpid()                  # link beside the target, then rename: a reader (or a snapshot
    subprocess.check_call([nvcc(), "-shared", "-cudart", "static", "-o", tmp] + objs)      # of the tree) never sees half a file
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
