"""Builds juicer_b200/libjuicer_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjuicer_b200.so")
SOURCES = ["jgpu_engine.cu", "host_loaders.cpp", "host_mmf.cpp", "host_queue.cpp"]
DEPS = SOURCES + ["jgpu_device.cuh", "jgpu_gmm.cuh", "jgpu_search.cuh", "jgpu_err.h", "jgpu_softplus_table.h", "host_models.h", "host_queue.h",
                  os.path.join("..", "..", "include", "juicer_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",            # scores must never see a contracted FMA (parity with the CPU decoder)
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-lrt"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """`defines` / `out` build an instrumented variant next to the product library (e.g.
    defines=("JG_TRACE",), out=libjuicer_b200_trace.so, loaded through JUICER_B200_LIB)."""
    if out == LIB and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines] + ["-o", out] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build(force=True, verbose="-v" in sys.argv, defines=("JG_TRACE",),
                    out=os.path.join(HERE, "libjuicer_b200_trace.so")))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
