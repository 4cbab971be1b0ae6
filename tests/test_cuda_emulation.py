"""CPU check of CUDA kernel SOURCE (no GPU): tests/emu compiles the kernel section of a product .cu file as host C++ and
runs it thread by thread. Covers the liquid-SDF kernels -- the default scatter, and the per-axis variant and post-process
kernel that were written after round 1's GPU minutes were spent -- against the reference-generated fixtures and the
oracle, bit for bit. (The -m gpu suite remains the parity test proper; this guards index arithmetic, shared-memory
layout, bit packing and float evaluation order.)"""
import os
import sys

import numpy as np
import pytest

from conftest import bits_equal, load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emulate  # noqa: E402


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name", ["liquid_sdf_23x21x25_seams", "liquid_sdf_22x24x20_radius2", "liquid_sdf_post_24x20x22"])
def test_liquid_sdf_kernels_emulated(name, variant):
    meta, e = load_golden(name)
    _, src = load_golden(meta["source"])
    args = (meta["I"], meta["J"], meta["K"], meta["dx"], src[meta["key"]], meta["radius"])
    assert bits_equal(emulate.liquid_sdf(*args, variant=variant), e["out_phi"])
    if meta.get("solid_key"):
        assert bits_equal(emulate.liquid_sdf(*args, variant=variant, solid=src[meta["solid_key"]]), e["out_phi_post"])


def test_liquid_sdf_kernels_emulated_boundaries(oracle):
    I, J, K, dx = 23, 31, 12, 0.013
    rng = np.random.default_rng(77)
    pos = (rng.random((12000, 3)) * [I * dx, J * dx, K * dx]).astype(np.float32)
    pos[:1500] = (rng.random((1500, 3)) * [I * dx * 1.2, J * dx * 1.2, K * dx * 1.2] - 0.1 * I * dx).astype(np.float32)
    pos[1500:4000] = (rng.integers(0, 3, (2500, 3)) * np.float32(10 * dx) + rng.normal(0, 0.02 * dx, (2500, 3))).astype(np.float32)
    for radius in (0.5 * dx * np.sqrt(3.0), dx * np.sqrt(3.0), 0.3 * dx):
        want = oracle.liquid_sdf(I, J, K, dx, pos, radius)
        for variant in (0, 1):
            assert bits_equal(emulate.liquid_sdf(I, J, K, dx, pos, radius, variant=variant), want), (radius, variant)
    assert bits_equal(emulate.liquid_sdf(I, J, K, dx, pos[:0], 0.01), np.full((K, J, I), np.float32(3.0 * dx)))
    with pytest.raises(ValueError):                                     # the variant's launch gate (sr <= 4.5 dx)
        emulate.liquid_sdf(I, J, K, dx, pos, 2.4 * dx, variant=1)
