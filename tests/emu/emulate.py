"""TEST INFRASTRUCTURE: host emulation of product CUDA kernels that have no barriers or warp intrinsics.

The kernel section of a product .cu file (everything before its first launcher) is compiled by g++ as plain C++
through tests/emu/fake/cuda_runtime.h and run thread by thread. This checks the kernel SOURCE -- index arithmetic,
shared-memory layout, bit packing, float evaluation order -- against the reference-generated fixtures without a GPU.
It is not a product path and says nothing about performance or about concurrency (atomics are serialised)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "blender_flip_fluids_b200", "csrc")
_libs = {}


def _kernel_section(cu_name, first_launcher):
    text = open(os.path.join(CSRC, cu_name)).read()
    cut = text.index("}  // namespace\n\n" + first_launcher)
    return text[:cut] + "}  // namespace\n}  // namespace ffb200\n"


def liquid_sdf_lib():
    if "sdf" in _libs:
        return _libs["sdf"]
    d = tempfile.mkdtemp(prefix="ffemu_")
    with open(os.path.join(d, "liquid_sdf_kernels.inc"), "w") as f:
        f.write("#include <vector>\n" + _kernel_section("ffb200_liquid_sdf.cu", "int launch_liquid_sdf_postprocess"))
    out = os.path.join(d, "libemu_sdf.so")
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + os.path.join(HERE, "fake"), "-I" + d,
           "-I" + CSRC, os.path.join(HERE, "emulate_liquid_sdf.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stderr[-3000:])
    lib = C.CDLL(out)
    _libs["sdf"] = lib
    return lib


def liquid_sdf(I, J, K, dx, pos, radius, variant=0, solid=None):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    phi = np.empty((K, J, I), np.float32)
    fp = C.POINTER(C.c_float)
    sp = None if solid is None else np.ascontiguousarray(solid, dtype=np.float32).ctypes.data_as(fp)
    rc = liquid_sdf_lib().emu_liquid_sdf(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_double(radius), C.c_int(pos.shape[0]),
                                         pos.ctypes.data_as(fp), C.c_int(variant), sp, phi.ctypes.data_as(fp))
    if rc != 0:
        raise ValueError("the variant's launch gate rejects this radius")
    return phi
