"""ctypes loader for the C oracle (oracle/flip_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product package never does (tests/test_abi.py::test_product_never_imports_oracle enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libflip_oracle.so")
_lib = None

FLIP, APIC = 0, 1


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, seconds). Returns the .so path."""
    src = os.path.join(_HERE, "flip_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "flip_oracle.h")))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def mac_shapes(I, J, K):
    """C-order [K,J,I] shapes of the x-fastest u, v, w arrays (macvelocityfield.cpp:46-54)."""
    return (K, J, I + 1), (K, J + 1, I), (K + 1, J, I)


def bin_sort(I, J, K, dx, pos):
    pos = _f32(pos)
    n = pos.shape[0]
    cell = np.empty(n, np.int32)
    hkey = np.empty(n, np.uint32)
    perm = np.empty(n, np.uint32)
    lib().flip_oracle_bin_sort(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_int(n), _p(pos),
                               _p(cell, C.c_int32), _p(hkey, C.c_uint32), _p(perm, C.c_uint32))
    return cell, hkey, perm


def p2g(I, J, K, dx, radius, method, pos, vel, affx=None, affy=None, affz=None, with_wsum=False):
    pos, vel, affx, affy, affz = map(_f32, (pos, vel, affx, affy, affz))
    n = pos.shape[0]
    su, sv, sw = mac_shapes(I, J, K)
    u, v, w = np.zeros(su, np.float32), np.zeros(sv, np.float32), np.zeros(sw, np.float32)
    vu, vv, vw = np.zeros(su, np.uint8), np.zeros(sv, np.uint8), np.zeros(sw, np.uint8)
    wu, wv, ww = np.zeros(su, np.float32), np.zeros(sv, np.float32), np.zeros(sw, np.float32)
    lib().flip_oracle_p2g_w(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_double(radius), C.c_int(method),
                            C.c_int(n), _p(pos), _p(vel), _p(affx), _p(affy), _p(affz), _p(u), _p(v), _p(w),
                            _p(vu, C.c_uint8), _p(vv, C.c_uint8), _p(vw, C.c_uint8), _p(wu), _p(wv), _p(ww))
    if with_wsum:
        return (u, v, w), (vu, vv, vw), (wu, wv, ww)
    return (u, v, w), (vu, vv, vw)


def g2p_flip(I, J, K, dx, pos, vel, mac, saved, ratio):
    pos = _f32(pos)
    out = _f32(vel).copy()
    u, v, w = map(_f32, mac)
    su, sv, sw = map(_f32, saved)
    lib().flip_oracle_g2p_flip(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_int(pos.shape[0]), _p(pos),
                               _p(out), _p(u), _p(v), _p(w), _p(su), _p(sv), _p(sw), C.c_double(ratio))
    return out


def g2p_apic(I, J, K, dx, pos, mac):
    pos = _f32(pos)
    n = pos.shape[0]
    u, v, w = map(_f32, mac)
    vel = np.zeros((n, 3), np.float32)
    ax, ay, az = (np.zeros((n, 3), np.float32) for _ in range(3))
    lib().flip_oracle_g2p_apic(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_int(n), _p(pos), _p(vel),
                               _p(ax), _p(ay), _p(az), _p(u), _p(v), _p(w))
    return vel, ax, ay, az


def mac_sample(I, J, K, dx, pos, mac):
    pos = _f32(pos)
    u, v, w = map(_f32, mac)
    out = np.zeros_like(pos)
    lib().flip_oracle_mac_sample(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_int(pos.shape[0]), _p(pos),
                                 _p(out), _p(u), _p(v), _p(w))
    return out


def near_dims(I, J, K, dx):
    gi, gj, gk = C.c_int(), C.c_int(), C.c_int()
    lib().flip_oracle_near_dims(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.byref(gi), C.byref(gj),
                                C.byref(gk))
    return gi.value, gj.value, gk.value


def advect(I, J, K, dx, pos, mac, phi, near, dt, cfl=5.0, collide=True):
    pos = _f32(pos)
    u, v, w = map(_f32, mac)
    phi = _f32(phi)
    near = np.ascontiguousarray(near, dtype=np.uint8)
    out = np.zeros_like(pos)
    lib().flip_oracle_advect(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_int(pos.shape[0]), _p(pos),
                             _p(out), _p(u), _p(v), _p(w), _p(phi), _p(near, C.c_uint8), C.c_double(dt),
                             C.c_double(cfl), C.c_int(1 if collide else 0))
    return out


def extrapolate(grid, valid, layers):
    """GridUtils::extrapolateGrid on one (d, h, w) float grid; returns the extrapolated copy."""
    g = np.array(grid, dtype=np.float32, order="C", copy=True)
    v = np.ascontiguousarray(valid, dtype=np.uint8)
    d, h, w = g.shape
    lib().flip_oracle_extrapolate(C.c_int(w), C.c_int(h), C.c_int(d), _p(g), _p(v, C.c_uint8), C.c_int(layers))
    return g


def extrapolation_layers(cfl=5.0):
    """fluidsimulation.cpp:6284."""
    import math
    return int(math.ceil(math.sqrt(3) * cfl)) + 3


def max_particle_speed(vel):
    vel = _f32(vel)
    f = lib().flip_oracle_max_particle_speed
    f.restype = C.c_double
    return float(f(C.c_int(vel.shape[0]), _p(vel)))


def remove_particles(I, J, K, dx, pos, vel, phi, dt, cfl=5.0, max_per_cell=250, extreme_removal=True, open_bounds=None,
                     pre_removed=None):
    """_removeMarkerParticles: (removed mask, number removed for extreme velocity). open_bounds: None or the six planes
    (x-, x+, y-, y+, z-, z+), +-inf where closed; pre_removed: None or a byte mask (the lifetime rule)."""
    pos, vel, phi = _f32(pos), _f32(vel), _f32(phi)
    removed = np.zeros(pos.shape[0], np.uint8)
    ob = None if open_bounds is None else np.ascontiguousarray(open_bounds, dtype=np.float32).reshape(6)
    pre = None if pre_removed is None else np.ascontiguousarray(pre_removed, dtype=np.uint8).reshape(pos.shape[0])
    nx = C.c_int()
    lib().flip_oracle_remove_particles(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_int(pos.shape[0]), _p(pos),
                                       _p(vel), _p(phi), C.c_double(dt), C.c_double(cfl), C.c_int(max_per_cell),
                                       C.c_int(1 if extreme_removal else 0), _p(ob) if ob is not None else None,
                                       _p(pre, C.c_uint8) if pre is not None else None, _p(removed, C.c_uint8), C.byref(nx))
    return removed, nx.value


def liquid_sdf(I, J, K, dx, pos, radius=None):
    """ParticleLevelSet::calculateSignedDistanceField: phi[K, J, I] (cell centred). radius defaults to the
    reference's _liquidSDFParticleRadius = 0.5 * dx * sqrt(3) (_liquidSDFParticleScale = 1)."""
    pos = _f32(pos)
    radius = 0.5 * dx * np.sqrt(3.0) if radius is None else radius
    phi = np.empty((K, J, I), np.float32)
    lib().flip_oracle_liquid_sdf(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_double(radius), C.c_int(pos.shape[0]),
                                 _p(pos), _p(phi))
    return phi


def liquid_sdf_axes(I, J, K, dx, pos, radius=None):
    """liquid_sdf through the per-axis decomposition of the device variant: (phi, candidates skipped by the filter)."""
    pos = _f32(pos)
    radius = 0.5 * dx * np.sqrt(3.0) if radius is None else radius
    phi = np.empty((K, J, I), np.float32)
    f = lib().flip_oracle_liquid_sdf_axes
    f.restype = C.c_longlong
    skipped = f(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_double(radius), C.c_int(pos.shape[0]), _p(pos), _p(phi))
    return phi, int(skipped)


def attribute_p2g(I, J, K, dx, pos, attr, radius):
    """AttributeToGridTransfer<float>::transfer: (grid[K, J, I], valid[K, J, I])."""
    pos = _f32(pos)
    attr = np.ascontiguousarray(attr, dtype=np.float32).reshape(pos.shape[0])
    grid, valid = np.zeros((K, J, I), np.float32), np.zeros((K, J, I), np.uint8)
    lib().flip_oracle_attribute_p2g(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_double(radius), C.c_int(pos.shape[0]),
                                    _p(pos), _p(attr), _p(grid), _p(valid, C.c_uint8))
    return grid, valid


def liquid_sdf_postprocess(I, J, K, dx, phi, solid):
    """ParticleLevelSet::postProcessSignedDistanceField: returns the processed copy of phi[K, J, I]."""
    out = np.ascontiguousarray(phi, dtype=np.float32).reshape(K, J, I).copy()
    solid = _f32(solid)
    lib().flip_oracle_liquid_sdf_postprocess(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), _p(out), _p(solid))
    return out


def attribute_p2g_vec3(I, J, K, dx, pos, attr, radius):
    """AttributeToGridTransfer<vmath::vec3>::transfer: (grid[K, J, I, 3], valid[K, J, I])."""
    pos = _f32(pos)
    attr = np.ascontiguousarray(attr, dtype=np.float32).reshape(pos.shape[0], 3)
    grid, valid = np.zeros((K, J, I, 3), np.float32), np.zeros((K, J, I), np.uint8)
    lib().flip_oracle_attribute_p2g_vec3(C.c_int(I), C.c_int(J), C.c_int(K), C.c_double(dx), C.c_double(radius),
                                         C.c_int(pos.shape[0]), _p(pos), _p(attr), _p(grid), _p(valid, C.c_uint8))
    return grid, valid
