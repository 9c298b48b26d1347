"""Build libgensmc.so (C ABI + sm_100a kernels) in-tree with nvcc.

Flags that matter for correctness, not only speed:
  -fmad=false / -ffp-contract=off   no implicit FMA contraction, so device and host arithmetic
                                    round exactly like the reference formulas written out in
                                    models.cuh (explicit fma() calls in gsmc_math.h stay fused)
  -gencode arch=compute_100a,code=sm_100a   B200 only; no other architecture is built
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgensmc.so")
SOURCES = ["gsmc.cu"]
HEADERS = ["kernels.cuh", "models.cuh", "gsmc_rng.cuh", "gsmc_math.h", "gsmc_tables.h", "gsmc_fixed.h", "plugin.h", os.path.join("..", "..", "include", "gen_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libgensmc.so cannot be built")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-shared", "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
           "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"] + os.environ.get("GSMC_NVCC_EXTRA", "").split()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libgensmc.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
