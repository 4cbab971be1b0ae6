"""ctypes host side of libffb200.so -- mirrors how the reference's ``ffengine`` package binds
its library (src/engine/ffengine/pybindings.py:26-60: argtypes, an error flag, ``RuntimeError``
raised with the library's message).

``FlipContext`` holds one device context (grid + resident particles/fields). The three
reference-named operators take and return HOST numpy arrays in the reference's layouts and
are what the parity tests and the e2e benchmark call:

    velocity_advector_advect            <- VelocityAdvector::advect          (velocityadvector.cpp:38)
    update_marker_particle_velocities   <- _updateMarkerParticleVelocitiesThread (fluidsimulation.cpp:6845)
    advance_marker_particles            <- _advanceMarkerParticles RK3+collide   (fluidsimulation.cpp:7853)

There is no CPU fallback: if the CUDA library is missing or no device is present the
constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

FLIP, APIC = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libffb200.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)


class Timing(C.Structure):
    _fields_ = [("sort_ms", C.c_float), ("p2g_prep_ms", C.c_float), ("p2g_ms", C.c_float), ("g2p_ms", C.c_float),
                ("advect_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("sort_launches", C.c_int),
                ("p2g_prep_launches", C.c_int), ("p2g_launches", C.c_int), ("g2p_launches", C.c_int),
                ("advect_launches", C.c_int)]


class DeviceBuffers(C.Structure):
    """ffb200_device_buffers: raw device pointers of the resident arrays (include/ffb200.h)."""
    _fields_ = [("pos", C.c_void_p * 3), ("vel", C.c_void_p * 3), ("aff", C.c_void_p * 9), ("ids", C.c_void_p),
                ("n", C.c_int), ("capacity", C.c_int), ("field", C.c_void_p * 3), ("saved", C.c_void_p * 3),
                ("valid", C.c_void_p * 3), ("face_count", C.c_longlong * 3), ("face_plane", C.c_int * 3),
                ("kbase", C.c_int), ("kloc", C.c_int), ("k_own_begin", C.c_int), ("k_own_end", C.c_int),
                ("phi", C.c_void_p)]


# name -> argtypes; every entry point of include/ffb200.h (tests/test_abi.py checks the list
# against the header and the built library).
SIGNATURES = {
    "ffb200_create": [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int],
    "ffb200_create_slab": [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int],
    "ffb200_destroy": [C.c_void_p],
    "ffb200_get_error_message": [],
    "ffb200_get_version": [C.POINTER(C.c_int)] * 3,
    "ffb200_set_stream": [C.c_void_p, C.c_void_p],
    "ffb200_reset_stream": [C.c_void_p],
    "ffb200_set_fixed_batch": [C.c_void_p, C.c_int],
    "ffb200_synchronize": [C.c_void_p],
    "ffb200_get_timing": [C.c_void_p, C.POINTER(Timing)],
    "ffb200_set_valid_guard": [C.c_void_p, C.c_float, C.c_float],
    "ffb200_set_particles": [C.c_void_p, C.c_int] + [_f32p] * 5,
    "ffb200_get_particles": [C.c_void_p] + [_f32p] * 5,
    "ffb200_get_num_particles": [C.c_void_p, C.POINTER(C.c_int)],
    "ffb200_get_device_buffers": [C.c_void_p, C.POINTER(DeviceBuffers)],
    "ffb200_reserve_particles": [C.c_void_p, C.c_int, C.c_int],
    "ffb200_set_num_particles": [C.c_void_p, C.c_int, C.c_int],
    "ffb200_slab_record_floats": [C.c_void_p, C.POINTER(C.c_int)],
    "ffb200_slab_pack_layers": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int],
    "ffb200_slab_route": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)],
    "ffb200_slab_route_begin": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int],
    "ffb200_slab_route_end": [C.c_void_p, C.POINTER(C.c_int)],
    "ffb200_slab_route_end_known": [C.c_void_p, C.c_int, C.POINTER(C.c_int)],
    "ffb200_slab_route_ghosts_begin": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)],
    "ffb200_slab_append": [C.c_void_p, C.c_void_p, C.c_int, C.c_int],
    "ffb200_sort_particles": [C.c_void_p],
    "ffb200_get_binning": [C.c_void_p, _i32p, _u32p, _u32p],
    "ffb200_set_velocity_field": [C.c_void_p] + [_f32p] * 3,
    "ffb200_set_saved_velocity_field": [C.c_void_p] + [_f32p] * 3,
    "ffb200_get_velocity_field": [C.c_void_p] + [_f32p] * 3 + [_u8p] * 3,
    "ffb200_get_weight_sums": [C.c_void_p] + [_f32p] * 3,
    "ffb200_save_velocity_field": [C.c_void_p],
    "ffb200_set_valid_velocities": [C.c_void_p, _u8p, _u8p, _u8p],
    "ffb200_extrapolate_velocity_field": [C.c_void_p, C.c_int],
    "ffb200_set_solid": [C.c_void_p, _f32p, _u8p],
    "ffb200_set_solid_device": [C.c_void_p, C.c_void_p, C.c_void_p],
    "ffb200_set_precision": [C.c_void_p, C.c_int],
    "ffb200_set_particle_window": [C.c_void_p, C.c_int, C.c_int, C.c_int],
    "ffb200_get_tolerance_stats": [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int],
    "ffb200_attribute_to_grid_transfer": [C.c_void_p, C.c_int, _f32p, _f32p, C.c_int, C.c_double, C.c_int, _f32p, _u8p],
    "ffb200_p2g": [C.c_void_p, C.c_double, C.c_int],
    "ffb200_g2p": [C.c_void_p, C.c_int, C.c_double],
    "ffb200_advect": [C.c_void_p, C.c_double, C.c_double, C.c_int],
    "ffb200_velocity_advector_advect": [C.c_void_p, C.c_int] + [_f32p] * 5 + [C.c_double, C.c_int] + [_f32p] * 3 + [_u8p] * 3,
    "ffb200_declare_resident": [C.c_void_p, C.c_uint],
    "ffb200_get_maximum_particle_speed": [C.c_void_p, C.POINTER(C.c_double)],
    "ffb200_liquid_sdf": [C.c_void_p, C.c_double],
    "ffb200_get_liquid_sdf": [C.c_void_p, _f32p],
    "ffb200_postprocess_liquid_sdf": [C.c_void_p],
    "ffb200_calculate_signed_distance_field": [C.c_void_p, C.c_int, _f32p, C.c_double, _f32p],
    "ffb200_remove_marker_particles": [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "ffb200_remove_marker_particles_masked": [C.c_void_p, _f32p, _u8p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _u8p,
                                              C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "ffb200_pin_host_memory": [C.c_void_p, C.c_void_p, C.c_size_t],
    "ffb200_unpin_host_memory": [C.c_void_p, C.c_void_p],
    "ffb200_mark_removed_marker_particles": [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, _u8p, _f32p, _u8p, C.c_double, C.c_double, C.c_int,
                                             C.c_int, C.c_int, _u8p, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "ffb200_extrapolate_fluid_velocities": [C.c_void_p, _f32p, _f32p, _f32p, _u8p, _u8p, _u8p, C.c_int, C.c_int],
    "ffb200_update_marker_particle_velocities": [C.c_void_p, C.c_int] + [_f32p] * 11 + [C.c_int, C.c_double],
    "ffb200_advance_marker_particles": [C.c_void_p, C.c_int] + [_f32p] * 5 + [_u8p, C.c_double, C.c_double],
}


def load_library():
    """Load libffb200.so (built in-tree by build.py). Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FFB200_LIBRARY", LIB_PATH)      # tuning builds (build.build(out=...)); default: in-tree lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: run `python -m blender_flip_fluids_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.ffb200_destroy.restype = None
    lib.ffb200_get_error_message.restype = C.c_char_p
    _lib = lib
    return lib


def _check(ok, name):
    if ok != 1:                                           # FFB200_SUCCESS
        msg = load_library().ffb200_get_error_message().decode("utf-8", "replace")
        raise RuntimeError(msg if msg else name)


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def _ptr(a, t=_f32p):
    return None if a is None else a.ctypes.data_as(t)


def mac_shapes(I, J, K):
    return (K, J, I + 1), (K, J + 1, I), (K + 1, J, I)


class FlipContext:
    def __init__(self, isize, jsize, ksize, dx, device=0, slab=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.I, self.J, self.K, self.dx = int(isize), int(jsize), int(ksize), float(dx)
        if slab is None:
            ok = self._lib.ffb200_create(C.byref(self._h), self.I, self.J, self.K, self.dx, int(device))
        else:
            k0, k1, halo = slab
            ok = self._lib.ffb200_create_slab(C.byref(self._h), self.I, self.J, self.K, self.dx, int(device), int(k0),
                                              int(k1), int(halo))
        _check(ok, "ffb200_create")
        self.n = 0
        self.near_dims = (-(-self.K // 3), -(-self.J // 3), -(-self.I // 3))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ffb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _call(self, name, *args):
        _check(getattr(self._lib, name)(self._h, *args), name)

    # ---- resident state ------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        """Use an existing cudaStream_t handle (int); 0 is the legacy default stream."""
        self._call("ffb200_set_stream", C.c_void_p(int(cuda_stream)))

    def reset_stream(self):
        self._call("ffb200_reset_stream")

    def set_fixed_batch(self, on=True):
        self._call("ffb200_set_fixed_batch", 1 if on else 0)

    def synchronize(self):
        self._call("ffb200_synchronize")

    def timing(self):
        t = Timing()
        self._call("ffb200_get_timing", C.byref(t))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def set_valid_guard(self, abs_tol, per_contrib_tol):
        self._call("ffb200_set_valid_guard", C.c_float(abs_tol), C.c_float(per_contrib_tol))

    def set_particles(self, pos, vel, affx=None, affy=None, affz=None):
        pos = _f32(pos)
        n = pos.shape[0]
        vel, affx, affy, affz = (_f32(a, (n, 3)) for a in (vel, affx, affy, affz))
        self._keep = (pos, vel, affx, affy, affz)             # host buffers stay alive until the copy is done
        self._call("ffb200_set_particles", n, _ptr(pos), _ptr(vel), _ptr(affx), _ptr(affy), _ptr(affz))
        self.synchronize()
        self.n = n

    def get_particles(self, pos=True, vel=True, affine=False):
        n = self.n
        mk = lambda want: np.empty((n, 3), np.float32) if want else None
        p, v, ax, ay, az = mk(pos), mk(vel), mk(affine), mk(affine), mk(affine)
        self._call("ffb200_get_particles", _ptr(p), _ptr(v), _ptr(ax), _ptr(ay), _ptr(az))
        return p, v, ax, ay, az

    def device_buffers(self):
        b = DeviceBuffers()
        self._call("ffb200_get_device_buffers", C.byref(b))
        return b

    def reserve_particles(self, capacity, with_affine=False):
        self._call("ffb200_reserve_particles", int(capacity), 1 if with_affine else 0)

    def set_num_particles(self, n, has_affine=False):
        self._call("ffb200_set_num_particles", int(n), 1 if has_affine else 0)
        self.n = int(n)

    def slab_record_floats(self):
        r = C.c_int()
        self._call("ffb200_slab_record_floats", C.byref(r))
        return r.value

    def slab_pack_layers(self, lo_a, hi_a, ptr_a, lo_b, hi_b, ptr_b, capacity):
        self._call("ffb200_slab_pack_layers", int(lo_a), int(hi_a), C.c_void_p(ptr_a or 0), int(lo_b), int(hi_b),
                   C.c_void_p(ptr_b or 0), int(capacity))

    def slab_route(self, k_begin, k_end, ptr_up, ptr_down, capacity):
        counts = (C.c_int * 3)()
        self._call("ffb200_slab_route", int(k_begin), int(k_end), C.c_void_p(ptr_up or 0), C.c_void_p(ptr_down or 0),
                   int(capacity), counts)
        self.n = counts[0]
        return counts[0], counts[1], counts[2]

    def slab_route_begin(self, k_begin, k_end, ptr_up, ptr_down, capacity):
        self._call("ffb200_slab_route_begin", int(k_begin), int(k_end), C.c_void_p(ptr_up or 0), C.c_void_p(ptr_down or 0),
                   int(capacity))

    def slab_route_end(self):
        counts = (C.c_int * 3)()
        self._call("ffb200_slab_route_end", counts)
        self.n = counts[0]
        return counts[0], counts[1], counts[2]

    def slab_route_end_known(self, leaving):
        counts = (C.c_int * 3)()
        self._call("ffb200_slab_route_end_known", int(leaving), counts)
        self.n = counts[0]
        return counts[0]

    def slab_route_ghosts_begin(self, k_begin, k_end, ghost_layers, block_up, block_down, capacities):
        """capacities = (up migrants, up ghosts, down migrants, down ghosts)."""
        caps = (C.c_int * 4)(*[int(x) for x in capacities])
        self._call("ffb200_slab_route_ghosts_begin", int(k_begin), int(k_end), int(ghost_layers), C.c_void_p(block_up or 0),
                   C.c_void_p(block_down or 0), caps)

    def slab_append(self, ptr, count, as_ghost=False):
        self._call("ffb200_slab_append", C.c_void_p(ptr), int(count), 1 if as_ghost else 0)
        self.n += int(count)

    def sort_particles(self):
        self._call("ffb200_sort_particles")

    def get_binning(self):
        n = self.n
        cell, hkey, perm = np.empty(n, np.int32), np.empty(n, np.uint32), np.empty(n, np.uint32)
        self._call("ffb200_get_binning", _ptr(cell, _i32p), _ptr(hkey, _u32p), _ptr(perm, _u32p))
        return cell, hkey, perm

    def set_velocity_field(self, u, v, w, saved=False):
        su, sv, sw = mac_shapes(self.I, self.J, self.K)
        u, v, w = _f32(u, su), _f32(v, sv), _f32(w, sw)
        self._call("ffb200_set_saved_velocity_field" if saved else "ffb200_set_velocity_field", _ptr(u), _ptr(v), _ptr(w))
        self.synchronize()

    def get_velocity_field(self):
        su, sv, sw = mac_shapes(self.I, self.J, self.K)
        u, v, w = np.zeros(su, np.float32), np.zeros(sv, np.float32), np.zeros(sw, np.float32)
        vu, vv, vw = np.zeros(su, np.uint8), np.zeros(sv, np.uint8), np.zeros(sw, np.uint8)
        self._call("ffb200_get_velocity_field", _ptr(u), _ptr(v), _ptr(w), _ptr(vu, _u8p), _ptr(vv, _u8p), _ptr(vw, _u8p))
        return (u, v, w), (vu, vv, vw)

    def get_weight_sums(self):
        su, sv, sw = mac_shapes(self.I, self.J, self.K)
        u, v, w = np.zeros(su, np.float32), np.zeros(sv, np.float32), np.zeros(sw, np.float32)
        self._call("ffb200_get_weight_sums", _ptr(u), _ptr(v), _ptr(w))
        return u, v, w

    def save_velocity_field(self):
        self._call("ffb200_save_velocity_field")

    def maximum_particle_speed(self):
        """_getMaximumMarkerParticleSpeed on the resident velocities (the CFL time step's input)."""
        out = C.c_double()
        self._call("ffb200_get_maximum_particle_speed", C.byref(out))
        return out.value

    def liquid_sdf(self, radius=None, download=True):
        """ParticleLevelSet::calculateSignedDistanceField on the resident positions -> phi[K, J, I] (or None)."""
        radius = 0.5 * self.dx * np.sqrt(3.0) if radius is None else radius
        self._call("ffb200_liquid_sdf", C.c_double(radius))
        if not download:
            return None
        phi = np.empty((self.K, self.J, self.I), np.float32)
        self._call("ffb200_get_liquid_sdf", _ptr(phi))
        return phi

    def postprocess_liquid_sdf(self):
        """ParticleLevelSet::postProcessSignedDistanceField on the device field -> phi[K, J, I]."""
        self._call("ffb200_postprocess_liquid_sdf")
        phi = np.empty((self.K, self.J, self.I), np.float32)
        self._call("ffb200_get_liquid_sdf", _ptr(phi))
        return phi

    def calculate_signed_distance_field(self, pos, radius=None):
        """Host-array flavour; pos None: declared resident (declare_resident(particles=True))."""
        radius = 0.5 * self.dx * np.sqrt(3.0) if radius is None else radius
        if pos is not None:
            pos = _f32(pos)
            self.n = pos.shape[0]
        phi = np.empty((self.K, self.J, self.I), np.float32)
        self._call("ffb200_calculate_signed_distance_field", self.n, _ptr(pos), C.c_double(radius), _ptr(phi))
        return phi

    def remove_marker_particles(self, dt, cfl=5.0, max_particles_per_cell=250, max_frame_time_steps=6, extreme_velocity_removal=True,
                                open_bounds=None):
        """_removeMarkerParticles on the resident particles (solid SDF of set_solid). open_bounds: None (closed domain) or
        the six planes (x-, x+, y-, y+, z-, z+), +-inf on closed sides. Returns (remaining, extreme removed)."""
        remaining, extreme = C.c_int(), C.c_int()
        ob = None if open_bounds is None else np.ascontiguousarray(open_bounds, dtype=np.float32).reshape(6)
        self._call("ffb200_remove_marker_particles", C.c_double(dt), C.c_double(cfl), int(max_particles_per_cell),
                   int(max_frame_time_steps), 1 if extreme_velocity_removal else 0, _ptr(ob), C.byref(remaining), C.byref(extreme))
        self.n = remaining.value
        return remaining.value, extreme.value

    def remove_marker_particles_masked(self, dt, cfl=5.0, max_particles_per_cell=250, max_frame_time_steps=6,
                                       extreme_velocity_removal=True, open_bounds=None, pre_removed=None):
        """_removeMarkerParticles on the RESIDENT particles, compacting the device set and returning the mask (in the
        order the ids had before the call) for the host's own compaction -> (removed mask, extreme removed)."""
        n = self.n
        ob = None if open_bounds is None else np.ascontiguousarray(open_bounds, dtype=np.float32).reshape(6)
        pre = None if pre_removed is None else np.ascontiguousarray(pre_removed, dtype=np.uint8).reshape(n)
        removed = np.ones(max(n, 1), np.uint8)
        remaining, extreme = C.c_int(), C.c_int()
        self._call("ffb200_remove_marker_particles_masked", _ptr(ob), _ptr(pre, _u8p), C.c_double(dt), C.c_double(cfl),
                   int(max_particles_per_cell), int(max_frame_time_steps), 1 if extreme_velocity_removal else 0,
                   _ptr(removed, _u8p), C.byref(remaining), C.byref(extreme))
        removed = removed[:n]
        assert remaining.value == n - int(removed.sum())
        self.n = remaining.value
        return removed, extreme.value

    def mark_removed_marker_particles(self, pos, vel, phi, near_solid, dt, cfl=5.0, max_particles_per_cell=250, max_frame_time_steps=6,
                                      extreme_velocity_removal=True, open_bounds=None, pre_removed=None):
        """_removeMarkerParticles on host arrays -> (removed mask in the caller's order, extreme removed). pos/vel None:
        declared resident (declare_resident(particles=True)); phi/near_solid None: declare_resident(solid=True)."""
        if pos is not None:
            pos = _f32(pos)
            n = pos.shape[0]
            vel = _f32(vel, (n, 3))
            self.n = n
        n = self.n
        if phi is not None:
            phi = _f32(phi, (self.K + 1, self.J + 1, self.I + 1))
            near_solid = np.ascontiguousarray(near_solid, dtype=np.uint8)
        ob = None if open_bounds is None else np.ascontiguousarray(open_bounds, dtype=np.float32).reshape(6)
        pre = None if pre_removed is None else np.ascontiguousarray(pre_removed, dtype=np.uint8).reshape(n)
        removed = np.zeros(n, np.uint8)
        nrem, extreme = C.c_int(), C.c_int()
        self._call("ffb200_mark_removed_marker_particles", n, _ptr(pos), _ptr(vel), _ptr(phi), _ptr(near_solid, _u8p), _ptr(ob),
                   _ptr(pre, _u8p), C.c_double(dt), C.c_double(cfl), int(max_particles_per_cell), int(max_frame_time_steps),
                   1 if extreme_velocity_removal else 0, _ptr(removed, _u8p), C.byref(nrem), C.byref(extreme))
        assert nrem.value == int(removed.sum())
        return removed, extreme.value

    def declare_resident(self, particles=False, field=False, solid=False, saved=False):
        """ffb200_declare_resident: the next host-buffer call may skip uploading what the device already holds."""
        self._call("ffb200_declare_resident", C.c_uint((1 if particles else 0) | (2 if field else 0) | (4 if solid else 0) |
                                                       (8 if saved else 0)))

    def set_valid_velocities(self, validu, validv, validw):
        su, sv, sw = mac_shapes(self.I, self.J, self.K)
        m = [np.ascontiguousarray(a, dtype=np.uint8) for a in (validu, validv, validw)]
        for a, s in zip(m, (su, sv, sw)):
            if tuple(a.shape) != tuple(s):
                raise ValueError(f"expected shape {s}, got {a.shape}")
        self._call("ffb200_set_valid_velocities", *[_ptr(a, _u8p) for a in m])

    def extrapolate_velocity_field(self, num_layers=None, cfl=5.0):
        """_extrapolateFluidVelocities: num_layers defaults to ceil(sqrt(3) * cfl) + 3 (fluidsimulation.cpp:6284)."""
        if num_layers is None:
            import math
            num_layers = int(math.ceil(math.sqrt(3) * cfl)) + 3
        self._call("ffb200_extrapolate_velocity_field", C.c_int(num_layers))

    def set_solid(self, phi, near_solid):
        phi = _f32(phi, (self.K + 1, self.J + 1, self.I + 1))
        near = np.ascontiguousarray(near_solid, dtype=np.uint8)
        if near.shape != self.near_dims:
            raise ValueError(f"near-solid mask must have shape {self.near_dims}, got {near.shape}")
        self._call("ffb200_set_solid", _ptr(phi), _ptr(near, _u8p))
        self.synchronize()

    def set_particle_window(self, k_lo, k_hi, mode):
        """ffb200_set_particle_window: mode 0 off, 1 inside the planes [k_lo, k_hi), 2 outside."""
        self._call("ffb200_set_particle_window", int(k_lo), int(k_hi), int(mode))

    def set_precision(self, tolerance):
        """ffb200_set_precision: False = exact (bit-identical gathers), True = tolerance (fp32 gathers, 1e-5)."""
        self._call("ffb200_set_precision", 1 if tolerance else 0)

    def tolerance_stats(self, reset=False):
        """-> dict(g2p_exact_components, advected, advected_exact) since the last reset."""
        c = (C.c_ulonglong * 4)()
        self._call("ffb200_get_tolerance_stats", c, 1 if reset else 0)
        return {"g2p_exact": int(c[1]), "advected": int(c[2]), "advected_exact": int(c[3])}

    def set_solid_device(self, phi_ptr, near_ptr):
        """ffb200_set_solid with device pointers (stored node planes of phi, whole near-solid grid)."""
        self._call("ffb200_set_solid_device", C.c_void_p(int(phi_ptr)), C.c_void_p(int(near_ptr)))

    def attribute_to_grid_transfer(self, pos, attr, radius, normalize=True):
        """AttributeToGridTransfer<T>::transfer: attr [n] (float) or [n, 3] (vmath::vec3) -> (grid [K, J, I(, 3)], valid [K, J, I])."""
        pos = _f32(pos)
        n = pos.shape[0]
        attr = np.ascontiguousarray(attr, dtype=np.float32)
        ncomp = 3 if attr.ndim == 2 else 1
        if attr.shape != ((n, 3) if ncomp == 3 else (n,)):
            raise ValueError(f"attribute array of shape {attr.shape} for {n} particles")
        grid = np.zeros((self.K, self.J, self.I) + ((3,) if ncomp == 3 else ()), np.float32)
        valid = np.zeros((self.K, self.J, self.I), np.uint8)
        self._call("ffb200_attribute_to_grid_transfer", n, _ptr(pos), _ptr(attr), ncomp, C.c_double(radius), 1 if normalize else 0,
                   _ptr(grid), _ptr(valid, _u8p))
        self.n = n
        return grid, valid

    # ---- stages on resident data ----------------------------------------------------------------
    def p2g(self, radius, method):
        self._call("ffb200_p2g", C.c_double(radius), int(method))

    def g2p(self, method, ratio_pic_flip=0.05):
        self._call("ffb200_g2p", int(method), C.c_double(ratio_pic_flip))

    def advect(self, dt, cfl=5.0, collide=True):
        self._call("ffb200_advect", C.c_double(dt), C.c_double(cfl), 1 if collide else 0)

    # ---- reference-named host-buffer operators --------------------------------------------------------
    def velocity_advector_advect(self, pos, vel, affx=None, affy=None, affz=None, radius=None, method=FLIP, out=None):
        """VelocityAdvector::advect on host arrays -> ((u,v,w), (validU,validV,validW)).
        pos None (after declare_resident(particles=True)): the resident particles; out False: the field stays resident."""
        if pos is None:
            n = self.n
        else:
            pos = _f32(pos)
            n = pos.shape[0]
            vel, affx, affy, affz = (_f32(a, (n, 3)) for a in (vel, affx, affy, affz))
        radius = 0.5 * self.dx * np.sqrt(3.0) if radius is None else radius
        if out is False:
            out = (None,) * 6
        elif out is None:
            su, sv, sw = mac_shapes(self.I, self.J, self.K)
            out = (np.zeros(su, np.float32), np.zeros(sv, np.float32), np.zeros(sw, np.float32),
                   np.zeros(su, np.uint8), np.zeros(sv, np.uint8), np.zeros(sw, np.uint8))
        u, v, w, vu, vv, vw = out
        self._call("ffb200_velocity_advector_advect", n, _ptr(pos), _ptr(vel), _ptr(affx), _ptr(affy), _ptr(affz),
                   C.c_double(radius), int(method), _ptr(u), _ptr(v), _ptr(w), _ptr(vu, _u8p), _ptr(vv, _u8p),
                   _ptr(vw, _u8p))
        self.n = n
        return (u, v, w), (vu, vv, vw)

    def update_marker_particle_velocities(self, pos, vel, mac, saved=None, method=FLIP, ratio_pic_flip=0.05,
                                          inplace=False, aff_out=None):
        """G2P on host arrays. FLIP -> new velocities; APIC -> (velocities, affx, affy, affz).

        inplace=True updates ``vel`` itself, as the reference does (fluidsimulation.cpp:6782);
        aff_out=(ax, ay, az) supplies the APIC output buffers (e.g. pinned memory)."""
        apic = method == APIC
        if pos is None:                       # resident particles (declare_resident(particles=True)): results stay on the device
            u, v, w = (None, None, None) if mac is None else (_f32(a) for a in mac)
            su, sv, sw = (None, None, None) if saved is None else (_f32(a) for a in saved)
            self._call("ffb200_update_marker_particle_velocities", self.n, None, None, None, None, None, _ptr(u), _ptr(v), _ptr(w),
                       _ptr(su), _ptr(sv), _ptr(sw), int(method), C.c_double(ratio_pic_flip))
            return None
        pos = _f32(pos)
        n = pos.shape[0]
        vel = _f32(vel, (n, 3))
        if not inplace:
            vel = vel.copy()
        u, v, w = (_f32(a) for a in mac)
        su, sv, sw = (None, None, None) if saved is None else (_f32(a) for a in saved)
        if apic and aff_out is not None:
            ax, ay, az = (_f32(a, (n, 3)) for a in aff_out)
        else:
            ax, ay, az = (np.zeros((n, 3), np.float32) if apic else None for _ in range(3))
        self._call("ffb200_update_marker_particle_velocities", n, _ptr(pos), _ptr(vel), _ptr(ax), _ptr(ay), _ptr(az),
                   _ptr(u), _ptr(v), _ptr(w), _ptr(su), _ptr(sv), _ptr(sw), int(method), C.c_double(ratio_pic_flip))
        self.n = n
        return (vel, ax, ay, az) if apic else vel

    def advance_marker_particles(self, pos, mac, phi=None, near_solid=None, dt=1.0 / 60.0, cfl=5.0, inplace=False):
        """RK3 + collision on host arrays -> new positions (inplace=True overwrites ``pos``)."""
        if pos is None:                       # resident particles and field: the advected positions stay on the device
            phi = _f32(phi)
            near = None if near_solid is None else np.ascontiguousarray(near_solid, dtype=np.uint8)
            u, v, w = (None, None, None) if mac is None else (_f32(a) for a in mac)
            self._call("ffb200_advance_marker_particles", self.n, None, _ptr(u), _ptr(v), _ptr(w), _ptr(phi), _ptr(near, _u8p),
                       C.c_double(dt), C.c_double(cfl))
            return None
        out = _f32(pos) if inplace else _f32(pos).copy()
        n = out.shape[0]
        u, v, w = (_f32(a) for a in mac)
        phi = _f32(phi)
        near = None if near_solid is None else np.ascontiguousarray(near_solid, dtype=np.uint8)
        self._call("ffb200_advance_marker_particles", n, _ptr(out), _ptr(u), _ptr(v), _ptr(w), _ptr(phi),
                   _ptr(near, _u8p), C.c_double(dt), C.c_double(cfl))
        self.n = n
        return out


class AttributeTransfer:
    """AttributeToGridTransfer<T>::transfer (attributetogridtransfer.h:52-157) through ffb200_attribute_to_grid_transfer:
    scalar and vmath::vec3 payloads, radii of 1 to 3 dx, on the cell-centred I x J x K grid. (Round 1 composed this from
    the U-direction velocity transfer of a grid one cell narrower; the identity behind that is still checked on the CPU,
    tests/test_oracle_golden.py::test_attribute_p2g_is_a_shifted_u_transfer.)"""

    def __init__(self, I, J, K, dx, device=0):
        self.I, self.J, self.K, self.dx = I, J, K, dx
        self.ctx = FlipContext(I, J, K, dx, device)

    def close(self):
        self.ctx.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def transfer(self, pos, attr, radius, normalize=True):
        """-> (grid[K, J, I(, 3)] float32, valid[K, J, I] uint8)."""
        return self.ctx.attribute_to_grid_transfer(pos, attr, radius, normalize)
