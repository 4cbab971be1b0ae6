"""Minimal ctypes driver for the reference engine's FluidSimulation C ABI -- TEST INFRASTRUCTURE.

Binds the handful of `FluidSimulation_*` functions the drop-in test needs, with the signatures
of c_bindings/fluidsimulation_c.cpp (cited per function) and the error convention of
ffengine/pybindings.py:26-60 (trailing `int *err`, 1 = success, message from
CBindings_get_error_message). It can load either the unmodified reference library
(oracle/_ref/libffengine_ref.so) or the drop-in (blender_flip_fluids_b200/lib/libffengine_b200.so):
both export the same ABI, which is the point.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


class MarkerParticleData(C.Structure):          # fluidsimulation.h:153-157
    _fields_ = [("size", C.c_int), ("positions", C.c_char_p), ("velocities", C.c_char_p)]


class MarkerParticleAffineData(C.Structure):    # fluidsimulation.h:159-164
    _fields_ = [("size", C.c_int), ("affineX", C.c_char_p), ("affineY", C.c_char_p), ("affineZ", C.c_char_p)]


class MeshStats(C.Structure):                   # fluidsimulation.h:62-67
    _fields_ = [("enabled", C.c_int), ("vertices", C.c_int), ("triangles", C.c_int), ("bytes", C.c_uint)]


class TimingStats(C.Structure):                 # fluidsimulation.h:69-78
    _fields_ = [(k, C.c_double) for k in ("total", "mesh", "advection", "particles", "pressure", "diffuse", "viscosity", "objects")]


class FrameStats(C.Structure):                  # fluidsimulation.h:80-150: 50 mesh records, the timing block last
    _fields_ = [("frame", C.c_int), ("substeps", C.c_int), ("delta_time", C.c_double), ("fluid_particles", C.c_int),
                ("diffuse_particles", C.c_int), ("performance_score", C.c_int), ("pressure_solver_enabled", C.c_int),
                ("pressure_solver_success", C.c_int), ("pressure_solver_error", C.c_double),
                ("pressure_solver_iterations", C.c_int), ("pressure_solver_max_iterations", C.c_int),
                ("viscosity_solver_enabled", C.c_int), ("viscosity_solver_success", C.c_int),
                ("viscosity_solver_error", C.c_double), ("viscosity_solver_iterations", C.c_int),
                ("viscosity_solver_max_iterations", C.c_int), ("mesh", MeshStats * 50), ("timing", TimingStats)]


class Engine:
    def __init__(self, lib_path, isize, jsize, ksize, dx):
        self.lib = C.CDLL(lib_path)
        self.lib.CBindings_get_error_message.restype = C.c_char_p
        f = self.lib.FluidSimulation_new_from_dimensions                      # fluidsimulation_c.cpp:56
        f.restype = C.c_void_p
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int)]
        err = C.c_int(0)
        self.obj = C.c_void_p(f(isize, jsize, ksize, dx, C.byref(err)))
        self._check(err, "FluidSimulation_new_from_dimensions")

    def _check(self, err, name):
        if err.value != 1:
            raise RuntimeError(name + " - " + self.lib.CBindings_get_error_message().decode())

    def _void(self, name, *args, argtypes=()):
        f = getattr(self.lib, name)
        f.restype = None
        f.argtypes = [C.c_void_p] + list(argtypes) + [C.POINTER(C.c_int)]
        err = C.c_int(0)
        f(self.obj, *args, C.byref(err))
        self._check(err, name)

    def close(self):
        if self.obj:
            self.lib.FluidSimulation_destroy.argtypes = [C.c_void_p]          # :71
            self.lib.FluidSimulation_destroy.restype = None
            self.lib.FluidSimulation_destroy(self.obj)
            self.obj = None

    def disable_console_output(self): self._void("FluidSimulation_disable_console_output")             # :593
    def disable_surface_reconstruction(self): self._void("FluidSimulation_disable_surface_reconstruction")  # :614
    def set_apic(self): self._void("FluidSimulation_set_velocity_transfer_method_APIC")                # :2980
    def set_picflip_ratio(self, r): self._void("FluidSimulation_set_PICFLIP_ratio", r, argtypes=[C.c_double])  # :3007
    def set_max_thread_count(self, n): self._void("FluidSimulation_set_max_thread_count", n, argtypes=[C.c_int])  # :2591
    def add_body_force(self, x, y, z): self._void("FluidSimulation_add_body_force", x, y, z, argtypes=[C.c_double] * 3)  # :2599
    def initialize(self): self._void("FluidSimulation_initialize")                                     # :97
    def update(self, dt): self._void("FluidSimulation_update", dt, argtypes=[C.c_double])              # :115

    def enable_surface_velocity_attribute(self): self._void("FluidSimulation_enable_surface_velocity_attribute")      # :1070
    def enable_surface_velocity_attribute_against_obstacles(self):                                     # :1091
        self._void("FluidSimulation_enable_surface_velocity_attribute_against_obstacles")
    def enable_surface_age_attribute(self): self._void("FluidSimulation_enable_surface_age_attribute")           # :1154
    def enable_surface_color_attribute(self): self._void("FluidSimulation_enable_surface_color_attribute")       # :1282
    def enable_surface_viscosity_attribute(self): self._void("FluidSimulation_enable_surface_viscosity_attribute")   # :1389
    def enable_fluid_particle_lifetime_attribute(self): self._void("FluidSimulation_enable_fluid_particle_lifetime_attribute")  # :938

    def set_fluid_boundary_collisions(self, active6):                                                  # :226 (x-, x+, y-, y+, z-, z+)
        arr = (C.c_int * 6)(*[int(bool(a)) for a in active6])
        self._void("FluidSimulation_set_fluid_boundary_collisions", arr, argtypes=[C.POINTER(C.c_int)])

    def frame_stats(self):                                                                             # :4618 (struct by value)
        f = self.lib.FluidSimulation_get_frame_stats_data
        f.restype = FrameStats
        f.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        err = C.c_int(0)
        st = f(self.obj, C.byref(err))
        self._check(err, "FluidSimulation_get_frame_stats_data")
        return st

    def load_marker_particle_data(self, pos, vel):                                                     # :4872
        pos, vel = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(vel, np.float32)
        d = MarkerParticleData(pos.shape[0], pos.tobytes(), vel.tobytes())
        self._void("FluidSimulation_load_marker_particle_data", d, argtypes=[MarkerParticleData])

    def load_marker_particle_affine_data(self, ax, ay, az):                                            # :4880
        d = MarkerParticleAffineData(ax.shape[0], *(np.ascontiguousarray(a, np.float32).tobytes() for a in (ax, ay, az)))
        self._void("FluidSimulation_load_marker_particle_affine_data", d, argtypes=[MarkerParticleAffineData])

    def num_marker_particles(self):                                                                    # :3151
        f = self.lib.FluidSimulation_get_num_marker_particles
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        err = C.c_int(0)
        n = f(self.obj, C.byref(err))
        self._check(err, "FluidSimulation_get_num_marker_particles")
        return n

    def _range(self, name):                                                                            # :4625-4679
        n = self.num_marker_particles()
        out = np.empty((n, 3), np.float32)
        self._void(name, 0, n, out.ctypes.data_as(C.c_char_p), argtypes=[C.c_int, C.c_int, C.c_char_p])
        return out

    def positions(self): return self._range("FluidSimulation_get_marker_particle_position_data_range")
    def velocities(self): return self._range("FluidSimulation_get_marker_particle_velocity_data_range")
    def affinex(self): return self._range("FluidSimulation_get_marker_particle_affinex_data_range")
