"""CPU suite: the C-ABI library loads and exports every symbol include/ffb200.h declares, the
ctypes table covers the header, the product never touches oracle/, and without a GPU the
library fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ffb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ffb200_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from blender_flip_fluids_b200 import build
    return build.build()                    # nvcc cross-compiles without a GPU


def test_header_symbols_exported(lib_path):
    names = declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ffb200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ffb200_[a-z0-9_]+)", out))
    assert exported == set(names), f"header/library mismatch: {exported ^ set(names)}"


def test_ctypes_table_matches_header():
    from blender_flip_fluids_b200 import engine
    assert sorted(engine.SIGNATURES) == declared_symbols()


def test_no_gpu_fails_loudly(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from blender_flip_fluids_b200 import engine
    with pytest.raises(RuntimeError) as e:
        engine.FlipContext(8, 8, 8, 0.1)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
    lib = engine.load_library()
    h = C.c_void_p()
    assert lib.ffb200_create(C.byref(h), -1, 8, 8, 0.1, 0) == 0        # FFB200_FAIL, message set
    assert lib.ffb200_get_error_message()
    assert lib.ffb200_p2g(None, 0.1, 0) == 0
    assert b"null context" in lib.ffb200_get_error_message()


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use oracle/."""
    pkg = os.path.join(ROOT, "blender_flip_fluids_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "flip_oracle" not in text, f"{f} references the oracle"
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"


def test_scene_generators():
    from blender_flip_fluids_b200 import scenes
    sc = scenes.dam_break(16, apic=True)
    assert sc.pos.shape == sc.vel.shape == sc.affx.shape and sc.pos.dtype.name == "float32"
    assert sc.n == 8 * (6 - 3) * (12 - 3) * (16 - 6)
    lo, hi = sc.pos.min(axis=0), sc.pos.max(axis=0)
    assert (lo >= 3 * sc.dx - 1e-6).all() and hi[0] < 0.4 + 1e-6 and hi[1] < 0.8 + 1e-6
    phi, near = scenes.analytic_solid_sdf(12, 9, 15, 0.1, sphere=(0.6, 0.4, 0.7, 0.2))
    assert phi.shape == (16, 10, 13) and near.shape == (5, 3, 4) and near.max() == 1
    assert phi[8, 4, 6] < 0                                  # inside the sphere obstacle


def test_dropin_reexports_reference_abi():
    """libffengine_b200.so defines exactly the interposed C++ members and resolves the
    reference's whole extern "C" surface through its DT_NEEDED reference library."""
    dropin = os.path.join(ROOT, "blender_flip_fluids_b200", "lib", "libffengine_b200.so")
    ref = os.path.join(ROOT, "oracle", "_ref", "libffengine_ref.so")
    if not (os.path.exists(dropin) and os.path.exists(ref)):
        pytest.skip("drop-in not built (needs /root/reference)")
    out = subprocess.run(["nm", "-D", "--defined-only", dropin], capture_output=True, text=True, check=True).stdout
    defined = set(re.findall(r" T (\S+)", out))
    assert defined == {"_ZN16VelocityAdvector6advectE26VelocityAdvectorParameters",
                       "_ZN15FluidSimulation37_updateMarkerParticleVelocitiesThreadEv",
                       "_ZN15FluidSimulation23_advanceMarkerParticlesEd",
                       "_ZN15FluidSimulation27_extrapolateFluidVelocitiesER16MACVelocityFieldR26ValidVelocityComponentGrid",
                       "_ZN16ParticleLevelSet28calculateSignedDistanceFieldER14ParticleSystemd",
                       "_ZN15FluidSimulation30_getMaximumMarkerParticleSpeedEv",
                       # explicit specialisations of the weak template instantiations the reference calls through the PLT
                       "_ZN23AttributeToGridTransferIfE8transferE27AttributeTransferParametersIfE",
                       "_ZN23AttributeToGridTransferIN5vmath4vec3EE8transferE27AttributeTransferParametersIS1_E",
                       # bookkeeping hooks that forward to the reference's definition (dlsym RTLD_NEXT)
                       "_ZN14ParticleSystem25getAttributeValuesVector3ER23ParticleSystemAttribute",
                       "_ZN15FluidSimulation10initializeEv"}
    needed = subprocess.run(["readelf", "-d", dropin], capture_output=True, text=True, check=True).stdout
    assert "libffb200.so" in needed and "libffengine_cpu.so" in needed
    refsyms = subprocess.run(["nm", "-D", "--defined-only", ref], capture_output=True, text=True, check=True).stdout
    c_abi = set(re.findall(r" T ((?:FluidSimulation|MeshObject|MeshFluidSource|ForceField\w*|Mixbox|CBindings)_\w+)", refsyms))
    assert len(c_abi) >= 700                              # 728 extern "C" exports
    lib = C.CDLL(dropin)
    missing = [s for s in sorted(c_abi) if not hasattr(lib, s)]
    assert not missing, missing[:5]


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours): exactly one JSON line on
    stdout with the contract's keys; the other ranks of a multi-rank launch print nothing and exit 0."""
    import json
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        pytest.skip("reference harness not built (needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--grid", "128"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same scene as the GPU arm names (identical workload string), on a stated sample of it
    sys.path.insert(0, ROOT)
    import bench
    from blender_flip_fluids_b200 import scenes
    assert d["config"]["workload"] == bench.workload_string(128, scenes.dam_break_count(128, 128, 128, 8))
    assert d["cpu_baseline"]["same_config"] is True and d["cpu_baseline"]["scene_particles"] == 4637952
    assert 0 < d["cpu_baseline"]["sample_particles"] <= 4637952 and d["scaling"] == "strong"
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_dropin_without_usable_device_reports_error_flag(tmp_path):
    """No CUDA device (this CPU box) or an ordinal that does not exist (FFB200_DEVICE=99 on a GPU box): the drop-in
    has no CPU fallback, so FluidSimulation_update must fail -- as err = 0 + message (cbindings.h:48-154), with the
    process alive and the particle getters still answering, never as std::terminate."""
    dropin = os.path.join(ROOT, "blender_flip_fluids_b200", "lib", "libffengine_b200.so")
    if not os.path.exists(dropin):
        pytest.skip("drop-in not built (needs /root/reference)")
    code = '''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from blender_flip_fluids_b200 import scenes
from ffengine_mini import Engine
sc = scenes.dam_break(16, apic=False, dx=0.02, vel="swirl", v0=0.4, seed=17)
e = Engine(%r, 16, 16, 16, sc.dx)
e.disable_console_output(); e.disable_surface_reconstruction(); e.set_max_thread_count(2)
e.load_marker_particle_data(sc.pos, sc.vel)
e.initialize()
try:
    e.update(1.0 / 60.0)
    print("UPDATED")
except RuntimeError as ex:
    print("ERR", ex)
print("ALIVE", e.num_marker_particles(), e.positions().shape[0])
e.close()
''' % (ROOT, os.path.join(ROOT, "tests"), dropin)
    env = dict(os.environ, FFB200_DEVICE="99")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "ERR FluidSimulation_update - ffb200_create" in r.stdout and "UPDATED" not in r.stdout, r.stdout[-1500:]
    assert "ALIVE 2160 2160" in r.stdout


def test_header_compiles_as_plain_c_and_cxx(tmp_path):
    """include/ffb200.h is the boundary: it must be consumable by a C compiler (cgo / JNI / ctypes generators read it
    as C) and by C++ (the interposer), with no CUDA or torch types in any signature."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include")
    src = tmp_path / "use.c"
    src.write_text('#include "ffb200.h"\nint main(void) { ffb200_context *c = 0; int a, b, r; (void)c; return ffb200_get_version(&a, &b, &r) ? 0 : 1; }\n')
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", f"-I{inc}", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cxx = tmp_path / "use.cpp"
    cxx.write_text(src.read_text())
    r = subprocess.run([shutil.which("g++"), "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", f"-I{inc}", str(cxx)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, "ffb200.h")).read(), flags=re.S)      # declarations only
    for banned in ("cuda", "torch", "at::", "std::", "Tensor"):
        assert banned not in code.replace("void *cuda_stream", ""), banned
