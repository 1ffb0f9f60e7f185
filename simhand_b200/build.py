"""Builds libsimhand_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m simhand_b200.build [--force]

The shared library is the C-ABI boundary declared in include/simhand_b200.h; it links only the CUDA
runtime (static), so it loads without torch.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libsimhand_b200.so")

SOURCES = ["smh_api.cu", "smh_prep.cu", "smh_mpjpe.cu", "smh_sweep_fp32.cu", "smh_sweep_tc.cu", "smh_exchange.cu", "smh_shard.cu", "smh_head.cu", "smh_transform.cu",
           "smh_finalize.cu", "smh_selftest.cu"]
HEADERS = ["smh_common.cuh", "smh_internal.h", os.path.join("..", "..", "include", "simhand_b200.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, trace: bool = False, variant: str = "", defines=()) -> str:
    """trace=True: development build with -DSMH_TRACE (per-role cycle counters in the sweeps, tools/trace_sweeps.py)
    into lib/libsimhand_b200_trace.so; the product library is never built with it."""
    global OBJDIR, LIB
    flags = list(FLAGS)
    if trace:
        OBJDIR, LIB = os.path.join(HERE, "build_trace"), os.path.join(LIBDIR, "libsimhand_b200_trace.so")
        flags.append("-DSMH_TRACE")
    if variant:
        # experiment builds (python -m simhand_b200.build --variant NAME -DFLAG ...): lib/libsimhand_b200_NAME.so, picked up
        # through SMH_LIB; never the product library
        OBJDIR, LIB = os.path.join(HERE, "build_" + variant), os.path.join(LIBDIR, f"libsimhand_b200_{variant}.so")
        flags.extend(defines)
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC, *flags, "-c", s, "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        return s, r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s, log in ex.map(compile_one, jobs):
                if verbose:
                    print(f"== {os.path.basename(s)}\n{log}")
                with open(os.path.join(OBJDIR, os.path.basename(s) + ".ptxas.log"), "w") as fh:
                    fh.write(log)
    objs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or force or _newer(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xcompiler", "-fPIC"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    path = build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, trace="--trace" in sys.argv, variant=var,
                 defines=[a for a in sys.argv if a.startswith("-D")])
    print(path)
