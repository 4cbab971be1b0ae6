"""Build libffb200.so (CUDA, sm_100a) in-tree with nvcc. Cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libffb200.so")
SOURCES = ["ffb200_api.cu", "ffb200_sort.cu", "ffb200_p2g.cu", "ffb200_g2p.cu", "ffb200_advect.cu", "ffb200_slab.cu", "ffb200_extrapolate.cu", "ffb200_remove.cu", "ffb200_liquid_sdf.cu"]

# -fmad=false: the reference is built without FMA contraction; wherever results are compared
# bit-for-bit a*b+c must round twice. Explicit fmaf()/fma() calls are still honoured.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "--use_fast_math=false", "-Xcompiler", "-fPIC,-O2,-Wall", "-shared"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ffb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """Build the library. `extra_flags`/`out` build a tuning variant (e.g. -DFFB_CELLS_MINB=4) next to
    the default one; engine.load_library() picks it up through FFB200_LIBRARY."""
    out = out or LIB
    if not force and out == LIB and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(out), exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + list(extra_flags)
    cmd = [nvcc_path()] + flags + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    env = dict(os.environ)
    env.pop("CXX", None), env.pop("CC", None)      # the image's CXX wrapper lacks libgomp specs; nvcc wants plain g++
    r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libffb200.so")
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
