// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Stage-level driver for the UNMODIFIED reference engine (oracle/_ref/libffengine_ref.so).
// Compiled with -fno-access-control against the reference headers where they lie, so the
// private hot-path members of FluidSimulation can be called directly:
//   P2G      VelocityAdvector::advect                       velocityadvector.cpp:38-43
//   G2P      FluidSimulation::_updateMarkerParticleVelocitiesThread   fluidsimulation.cpp:6845
//   advect   FluidSimulation::_advanceMarkerParticlesThread           fluidsimulation.cpp:7625
// It is used (a) to pin oracle/flip_oracle.c bit-for-bit, (b) to generate the golden fixtures
// in tests/golden/ (tests/golden/make_golden.py), and (c) as bench.py's "reference" CPU arm.
//
// Usage: ref_harness <mode> <workdir> key=value ...
//   modes: p2g | g2p | advect | scene
// Arrays are exchanged as little-endian NumPy .npy files in <workdir> (in_*.npy -> out_*.npy).
// One JSON line with timings is printed on stdout.

#include "fluidsimulation.h"
#include "particlelevelset.h"
#include "gridutils.h"
#include "threadutils.h"
#include "attributetogridtransfer.h"
#include "velocityadvector.h"
#include "macvelocityfield.h"
#include "particlesystem.h"
#include "meshlevelset.h"
#include "meshobject.h"
#include "trianglemesh.h"
#include "threadutils.h"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

// ------------------------------------------------------------------ tiny .npy reader/writer
struct Npy {
    std::string descr;            // "<f4", "|u1", "<i4"
    std::vector<size_t> shape;
    std::vector<char> data;
    size_t count() const { size_t n = 1; for (size_t s : shape) n *= s; return n; }
    float *f32() { return reinterpret_cast<float *>(data.data()); }
    uint8_t *u8() { return reinterpret_cast<uint8_t *>(data.data()); }
};

static size_t descr_size(const std::string &d) { return (size_t)atoi(d.c_str() + 2); }

static bool npy_exists(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

static Npy npy_load(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "ref_harness: cannot open %s\n", path.c_str()); exit(2); }
    unsigned char pre[10];
    if (fread(pre, 1, 10, f) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) {
        fprintf(stderr, "ref_harness: %s is not an npy file\n", path.c_str()); exit(2);
    }
    size_t hlen = pre[8] | (pre[9] << 8);
    if (pre[6] >= 2) {                       // v2/v3: 4-byte header length
        unsigned char extra[2];
        if (fread(extra, 1, 2, f) != 2) exit(2);
        hlen |= ((size_t)extra[0] << 16) | ((size_t)extra[1] << 24);
    }
    std::string hdr(hlen, ' ');
    if (fread(&hdr[0], 1, hlen, f) != hlen) exit(2);
    Npy a;
    size_t p = hdr.find("'descr':");
    size_t q0 = hdr.find('\'', p + 8), q1 = hdr.find('\'', q0 + 1);
    a.descr = hdr.substr(q0 + 1, q1 - q0 - 1);
    if (hdr.find("'fortran_order': True") != std::string::npos) {
        fprintf(stderr, "ref_harness: fortran order unsupported\n"); exit(2);
    }
    p = hdr.find("'shape':");
    size_t s0 = hdr.find('(', p), s1 = hdr.find(')', s0);
    std::string sh = hdr.substr(s0 + 1, s1 - s0 - 1);
    const char *c = sh.c_str();
    while (*c) {
        while (*c && (*c < '0' || *c > '9')) c++;
        if (!*c) break;
        a.shape.push_back((size_t)strtoull(c, (char **)&c, 10));
    }
    a.data.resize(a.count() * descr_size(a.descr));
    if (!a.data.empty() && fread(a.data.data(), 1, a.data.size(), f) != a.data.size()) {
        fprintf(stderr, "ref_harness: short read on %s\n", path.c_str()); exit(2);
    }
    fclose(f);
    return a;
}

static void npy_save(const std::string &path, const char *descr, const std::vector<size_t> &shape,
                     const void *data, size_t nbytes) {
    std::string sh = "(";
    for (size_t i = 0; i < shape.size(); i++) sh += std::to_string(shape[i]) + ",";
    sh += ")";
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + sh + ", }";
    size_t total = 10 + hdr.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    hdr += std::string(pad, ' ') + "\n";
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "ref_harness: cannot write %s\n", path.c_str()); exit(2); }
    unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0,
                             (unsigned char)(hdr.size() & 0xff), (unsigned char)(hdr.size() >> 8)};
    fwrite(pre, 1, 10, f);
    fwrite(hdr.data(), 1, hdr.size(), f);
    if (nbytes) fwrite(data, 1, nbytes, f);
    fclose(f);
}

// ------------------------------------------------------------------ helpers
static std::map<std::string, std::string> g_kv;
static std::string g_dir;

static std::string kv(const char *k, const char *dflt = nullptr) {
    auto it = g_kv.find(k);
    if (it != g_kv.end()) return it->second;
    if (!dflt) { fprintf(stderr, "ref_harness: missing argument %s=\n", k); exit(2); }
    return dflt;
}
static int kvi(const char *k, const char *d = nullptr) { return atoi(kv(k, d).c_str()); }
static double kvd(const char *k, const char *d = nullptr) { return strtod(kv(k, d).c_str(), nullptr); }
static std::string P(const char *name) { return g_dir + "/" + name + ".npy"; }
static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static std::vector<vmath::vec3> load_vec3(const char *name) {
    Npy a = npy_load(P(name));
    if (a.descr != "<f4" || a.count() % 3) { fprintf(stderr, "ref_harness: %s must be float32 [N,3]\n", name); exit(2); }
    std::vector<vmath::vec3> v(a.count() / 3);
    memcpy((void *)v.data(), a.data.data(), a.data.size());      // vec3 is 3 packed floats (vmath.h:37-61)
    return v;
}
static void save_vec3(const char *name, std::vector<vmath::vec3> &v) {
    npy_save(P(name), "<f4", {v.size(), 3}, v.data(), v.size() * 12);
}
static void load_grid(const char *name, Array3d<float> &g) {
    Npy a = npy_load(P(name));
    if (a.descr != "<f4" || (int)a.count() != g.width * g.height * g.depth) {
        fprintf(stderr, "ref_harness: %s has %zu elements, grid wants %d\n", name, a.count(), g.width * g.height * g.depth);
        exit(2);
    }
    memcpy(g._grid, a.data.data(), a.data.size());                // x-fastest flat layout (array3d.h:774-777)
}
static void save_grid(const char *name, Array3d<float> &g) {
    npy_save(P(name), "<f4", {(size_t)g.depth, (size_t)g.height, (size_t)g.width}, g._grid,
             (size_t)g.width * g.height * g.depth * 4);
}
static void save_mask(const char *name, Array3d<bool> &g) {
    static_assert(sizeof(bool) == 1, "bool must be one byte");
    npy_save(P(name), "|u1", {(size_t)g.depth, (size_t)g.height, (size_t)g.width}, g._grid,
             (size_t)g.width * g.height * g.depth);
}
static void load_mask(const char *name, Array3d<bool> &g) {
    Npy a = npy_load(P(name));
    if ((int)a.count() != g.width * g.height * g.depth) { fprintf(stderr, "ref_harness: %s size mismatch\n", name); exit(2); }
    for (size_t i = 0; i < a.count(); i++) g._grid[i] = a.u8()[i] != 0;
}

static void fill_particles(ParticleSystem &ps, bool apic) {
    std::vector<vmath::vec3> pos = load_vec3("in_pos"), vel = load_vec3("in_vel");
    ps.addAttributeVector3("POSITION");
    ps.addAttributeVector3("VELOCITY");
    ps.addValues("POSITION", pos);
    ps.addValues("VELOCITY", vel);
    if (apic) {
        std::vector<vmath::vec3> ax = load_vec3("in_affx"), ay = load_vec3("in_affy"), az = load_vec3("in_affz");
        ps.addAttributeVector3("AFFINEX");
        ps.addAttributeVector3("AFFINEY");
        ps.addAttributeVector3("AFFINEZ");
        ps.addValues("AFFINEX", ax);
        ps.addValues("AFFINEY", ay);
        ps.addValues("AFFINEZ", az);
    }
    ps.update();
}

static void dump_particles(ParticleSystem &ps, bool apic, const std::string &prefix) {
    std::vector<vmath::vec3> *v;
    ps.getAttributeValues("POSITION", v); save_vec3((prefix + "pos").c_str(), *v);
    ps.getAttributeValues("VELOCITY", v); save_vec3((prefix + "vel").c_str(), *v);
    if (apic) {
        ps.getAttributeValues("AFFINEX", v); save_vec3((prefix + "affx").c_str(), *v);
        ps.getAttributeValues("AFFINEY", v); save_vec3((prefix + "affy").c_str(), *v);
        ps.getAttributeValues("AFFINEZ", v); save_vec3((prefix + "affz").c_str(), *v);
    }
}

// ------------------------------------------------------------------ mode: p2g
static int mode_p2g() {
    int I = kvi("I"), J = kvi("J"), K = kvi("K");
    double dx = kvd("dx");
    bool apic = kv("method", "flip") == "apic";
    double radius = kvd("radius", "0");
    if (radius <= 0) radius = 0.5 * dx * sqrt(3.0);               // fluidsimulation.cpp:4351
    int reps = kvi("reps", "1");

    ParticleSystem ps;
    fill_particles(ps, apic);
    MACVelocityField mac(I, J, K, dx);
    ValidVelocityComponentGrid valid(I, J, K);
    VelocityAdvector va;
    VelocityAdvectorParameters prm;
    prm.particles = &ps;
    prm.vfield = &mac;
    prm.validVelocities = &valid;
    prm.particleRadius = radius;
    prm.velocityTransferMethod = apic ? VelocityAdvectorTransferMethod::APIC : VelocityAdvectorTransferMethod::FLIP;

    double best = 1e30;
    std::string times = "[";
    for (int r = 0; r < reps; r++) {
        valid.reset();                                            // fluidsimulation.cpp:5630-5631
        mac.clear();
        double t0 = now();
        va.advect(prm);
        double t = now() - t0;
        best = std::min(best, t);
        times += (r ? ", " : "") + std::to_string(t);
    }
    times += "]";
    save_grid("out_u", *mac.getArray3dU());
    save_grid("out_v", *mac.getArray3dV());
    save_grid("out_w", *mac.getArray3dW());
    save_mask("out_validu", valid.validU);
    save_mask("out_validv", valid.validV);
    save_mask("out_validw", valid.validW);
    printf("{\"mode\": \"p2g\", \"particles\": %zu, \"threads\": %d, \"t_p2g\": %.6f, \"times\": %s}\n", ps.size(),
           ThreadUtils::getMaxThreadCount(), best, times.c_str());
    return 0;
}

// A FluidSimulation shell with just the members the G2P / advect stages read.
static void shell_sim(FluidSimulation &sim, int I, int J, int K, double dx, bool apic) {
    sim._isize = I; sim._jsize = J; sim._ksize = K; sim._dx = dx;
    sim._velocityTransferMethod = apic ? FluidSimulation::VelocityTransferMethod::APIC
                                       : FluidSimulation::VelocityTransferMethod::FLIP;
    sim._MACVelocity = MACVelocityField(I, J, K, dx);
    fill_particles(sim._markerParticles, apic);
}

// ------------------------------------------------------------------ mode: g2p
static int mode_g2p() {
    int I = kvi("I"), J = kvi("J"), K = kvi("K");
    double dx = kvd("dx");
    bool apic = kv("method", "flip") == "apic";
    FluidSimulation sim(I, J, K, dx);
    sim.disableConsoleOutput();
    shell_sim(sim, I, J, K, dx, apic);
    sim._ratioPICFLIP = kvd("ratio", "0.05");
    load_grid("in_u", *sim._MACVelocity.getArray3dU());
    load_grid("in_v", *sim._MACVelocity.getArray3dV());
    load_grid("in_w", *sim._MACVelocity.getArray3dW());
    if (!apic) {
        sim._savedVelocityField = MACVelocityField(I, J, K, dx);
        load_grid("in_su", *sim._savedVelocityField.getArray3dU());
        load_grid("in_sv", *sim._savedVelocityField.getArray3dV());
        load_grid("in_sw", *sim._savedVelocityField.getArray3dW());
    }
    int reps = kvi("reps", "1");
    std::string times = "[";
    double t = 0;
    for (int r = 0; r < reps; r++) {               // reps > 1 is for timing only (FLIP updates in place)
        double t0 = now();
        sim._updateMarkerParticleVelocitiesThread();
        t = now() - t0;
        if (r == 0) dump_particles(sim._markerParticles, apic, "out_");
        times += (r ? ", " : "") + std::to_string(t);
    }
    times += "]";
    printf("{\"mode\": \"g2p\", \"particles\": %zu, \"threads\": %d, \"t_g2p\": %.6f, \"times\": %s}\n",
           sim._markerParticles.size(), ThreadUtils::getMaxThreadCount(), t, times.c_str());
    return 0;
}

static double run_advect(FluidSimulation &sim, double dt, std::vector<vmath::vec3> &out) {
    std::vector<vmath::vec3> *positions;
    sim._markerParticles.getAttributeValues("POSITION", positions);
    std::vector<vmath::vec3> copy = *positions;
    out.assign(copy.size(), vmath::vec3());
    int nt = (int)fmin(ThreadUtils::getMaxThreadCount(), copy.size());      // fluidsimulation.cpp:7867-7879
    std::vector<std::thread> th(nt);
    std::vector<int> iv = ThreadUtils::splitRangeIntoIntervals(0, copy.size(), nt);
    double t0 = now();
    for (int i = 0; i < nt; i++) {
        th[i] = std::thread(&FluidSimulation::_advanceMarkerParticlesThread, &sim, dt, iv[i], iv[i + 1], &copy, &out);
    }
    for (auto &t : th) t.join();
    return now() - t0;
}

// ------------------------------------------------------------------ mode: advect
static int mode_advect() {
    int I = kvi("I"), J = kvi("J"), K = kvi("K");
    double dx = kvd("dx"), dt = kvd("dt");
    FluidSimulation sim(I, J, K, dx);
    sim.disableConsoleOutput();
    shell_sim(sim, I, J, K, dx, false);
    sim._CFLConditionNumber = kvd("cfl", "5");
    load_grid("in_u", *sim._MACVelocity.getArray3dU());
    load_grid("in_v", *sim._MACVelocity.getArray3dV());
    load_grid("in_w", *sim._MACVelocity.getArray3dW());
    sim._solidSDF = MeshLevelSet(I, J, K, dx);
    load_grid("in_phi", sim._solidSDF._phi);
    sim._nearSolidGridCellSize = sim._nearSolidGridCellSizeFactor * dx;       // fluidsimulation.cpp:5448-5451
    int gi = (int)std::ceil(I * dx / sim._nearSolidGridCellSize);
    int gj = (int)std::ceil(J * dx / sim._nearSolidGridCellSize);
    int gk = (int)std::ceil(K * dx / sim._nearSolidGridCellSize);
    sim._nearSolidGrid = Array3d<bool>(gi, gj, gk, false);
    load_mask("in_near", sim._nearSolidGrid);
    std::vector<vmath::vec3> out;
    int reps = kvi("reps", "1");
    std::string times = "[";
    double t = 0;
    for (int r = 0; r < reps; r++) {
        t = run_advect(sim, dt, out);
        times += (r ? ", " : "") + std::to_string(t);
    }
    times += "]";
    save_vec3("out_pos", out);
    printf("{\"mode\": \"advect\", \"particles\": %zu, \"threads\": %d, \"t_advect\": %.6f, \"times\": %s}\n",
           out.size(), ThreadUtils::getMaxThreadCount(), t, times.c_str());
    return 0;
}

// ------------------------------------------------------------------ mode: scene
static TriangleMesh box_mesh(double x0, double y0, double z0, double w, double h, double d) {
    TriangleMesh m;
    for (int c = 0; c < 8; c++) {
        m.vertices.push_back(vmath::vec3(x0 + ((c & 1) ? w : 0), y0 + ((c & 2) ? h : 0), z0 + ((c & 4) ? d : 0)));
    }
    static const int q[6][4] = {{0, 2, 3, 1}, {4, 5, 7, 6}, {0, 1, 5, 4}, {2, 6, 7, 3}, {0, 4, 6, 2}, {1, 3, 7, 5}};
    for (auto &f : q) {
        m.triangles.push_back(Triangle(f[0], f[1], f[2]));
        m.triangles.push_back(Triangle(f[0], f[2], f[3]));
    }
    return m;
}

static TriangleMesh sphere_mesh(double cx, double cy, double cz, double r, int nlat = 24, int nlon = 48) {
    TriangleMesh m;
    m.vertices.push_back(vmath::vec3(cx, cy + r, cz));
    for (int a = 1; a < nlat; a++) {
        double th = M_PI * a / nlat;
        for (int b = 0; b < nlon; b++) {
            double ph = 2 * M_PI * b / nlon;
            m.vertices.push_back(vmath::vec3(cx + r * sin(th) * cos(ph), cy + r * cos(th), cz + r * sin(th) * sin(ph)));
        }
    }
    m.vertices.push_back(vmath::vec3(cx, cy - r, cz));
    int south = (int)m.vertices.size() - 1;
    auto ring = [&](int a, int b) { return 1 + (a - 1) * nlon + (b % nlon); };
    for (int b = 0; b < nlon; b++) {
        m.triangles.push_back(Triangle(0, ring(1, b + 1), ring(1, b)));
        m.triangles.push_back(Triangle(south, ring(nlat - 1, b), ring(nlat - 1, b + 1)));
    }
    for (int a = 1; a < nlat - 1; a++) {
        for (int b = 0; b < nlon; b++) {
            m.triangles.push_back(Triangle(ring(a, b), ring(a, b + 1), ring(a + 1, b)));
            m.triangles.push_back(Triangle(ring(a, b + 1), ring(a + 1, b + 1), ring(a + 1, b)));
        }
    }
    return m;
}

// Whole simulation through the reference's public API, then the three hot stages called
// one by one on the live object with every intermediate array dumped. warm=k runs k
// reference update() calls first so velocities, solid SDF and near-solid grid are the
// reference's own (SDF construction is out of scope for the GPU path).
static int mode_scene() {
    int I = kvi("I"), J = kvi("J"), K = kvi("K");
    double dx = kvd("dx");
    bool apic = kv("method", "flip") == "apic";
    double dt = kvd("dt", "0.016666666666666666");
    int warm = kvi("warm", "1");
    bool dump = kvi("dump", "1") != 0;

    FluidSimulation sim(I, J, K, dx);
    sim.disableConsoleOutput();
    sim.disableSurfaceReconstruction();
    if (apic) sim.setVelocityTransferMethodAPIC();
    sim.setPICFLIPRatio(kvd("ratio", "0.05"));
    sim.addBodyForce(0, kvd("gravity", "-9.81"), 0);

    std::vector<MeshObject *> keep;
    std::string obs = kv("obstacle", "none");
    if (obs != "none") {
        double a[6] = {0, 0, 0, 0, 0, 0};
        size_t colon = obs.find(':');
        std::string kind = obs.substr(0, colon), rest = obs.substr(colon + 1);
        int n = sscanf(rest.c_str(), "%lf,%lf,%lf,%lf,%lf,%lf", a, a + 1, a + 2, a + 3, a + 4, a + 5);
        TriangleMesh m = (kind == "sphere" && n >= 4) ? sphere_mesh(a[0], a[1], a[2], a[3])
                                                       : box_mesh(a[0], a[1], a[2], a[3], a[4], a[5]);
        MeshObject *mo = new MeshObject(I, J, K, dx);
        mo->updateMeshStatic(m);
        sim.addMeshObstacle(mo);
        keep.push_back(mo);
    }

    std::vector<vmath::vec3> pos = load_vec3("in_pos"), vel = load_vec3("in_vel");
    FluidSimulationMarkerParticleData d;
    d.size = (int)pos.size();
    d.positions = (char *)pos.data();
    d.velocities = (char *)vel.data();
    sim.loadMarkerParticleData(d);
    std::vector<vmath::vec3> ax, ay, az;
    if (apic && npy_exists(P("in_affx"))) {
        ax = load_vec3("in_affx"); ay = load_vec3("in_affy"); az = load_vec3("in_affz");
        FluidSimulationMarkerParticleAffineData ad;
        ad.size = (int)ax.size();
        ad.affineX = (char *)ax.data();
        ad.affineY = (char *)ay.data();
        ad.affineZ = (char *)az.data();
        sim.loadMarkerParticleAffineData(ad);
    }
    sim.initialize();
    double twarm0 = now();
    for (int i = 0; i < warm; i++) sim.update(dt);
    double twarm = now() - twarm0;

    size_t n = sim._markerParticles.size();
    sim._currentFrameDeltaTime = dt;
    sim._currentFrameTimeStep = dt;
    if (dump) dump_particles(sim._markerParticles, apic, "s0_");

    // --- P2G (fluidsimulation.cpp:5623-5652 without the extrapolation)
    sim._validVelocities.reset();
    sim._MACVelocity.clear();
    VelocityAdvectorParameters prm;
    prm.particles = &sim._markerParticles;
    prm.vfield = &sim._MACVelocity;
    prm.validVelocities = &sim._validVelocities;
    prm.particleRadius = sim._liquidSDFParticleRadius;
    prm.velocityTransferMethod = apic ? VelocityAdvectorTransferMethod::APIC : VelocityAdvectorTransferMethod::FLIP;
    double t0 = now();
    sim._velocityAdvector.advect(prm);
    double t_p2g = now() - t0;
    if (dump) {
        save_grid("s1_u", *sim._MACVelocity.getArray3dU());
        save_grid("s1_v", *sim._MACVelocity.getArray3dV());
        save_grid("s1_w", *sim._MACVelocity.getArray3dW());
        save_mask("s1_validu", sim._validVelocities.validU);
        save_mask("s1_validv", sim._validVelocities.validV);
        save_mask("s1_validw", sim._validVelocities.validW);
    }
    // --- CPU stages that stay on the reference path
    t0 = now();
    sim._extrapolateFluidVelocities(sim._MACVelocity, sim._validVelocities);
    double t_extrap = now() - t0;
    sim._saveVelocityField();
    sim._applyBodyForcesToVelocityField(dt);
    sim._pressureSolve(dt);
    sim._constrainVelocityFields();
    if (dump) {
        save_grid("s2_u", *sim._MACVelocity.getArray3dU());
        save_grid("s2_v", *sim._MACVelocity.getArray3dV());
        save_grid("s2_w", *sim._MACVelocity.getArray3dW());
        save_grid("s2_su", *sim._savedVelocityField.getArray3dU());
        save_grid("s2_sv", *sim._savedVelocityField.getArray3dV());
        save_grid("s2_sw", *sim._savedVelocityField.getArray3dW());
        save_grid("s2_phi", sim._solidSDF._phi);
        save_mask("s2_near", sim._nearSolidGrid);
    }
    // --- G2P
    t0 = now();
    sim._updateMarkerParticleVelocitiesThread();
    double t_g2p = now() - t0;
    if (dump) dump_particles(sim._markerParticles, apic, "s3_");
    // --- advect (fan-out of fluidsimulation.cpp:7853-7879, removal excluded)
    std::vector<vmath::vec3> out;
    double t_adv = run_advect(sim, dt, out);
    if (dump) save_vec3("s4_pos", out);

    double tot = t_p2g + t_g2p + t_adv;
    printf("{\"mode\": \"scene\", \"method\": \"%s\", \"particles\": %zu, \"threads\": %d, \"t_warm\": %.6f, "
           "\"t_p2g\": %.6f, \"t_extrapolate\": %.6f, \"t_g2p\": %.6f, \"t_advect\": %.6f, \"t_path\": %.6f, "
           "\"particle_updates_per_s\": %.1f, \"cfl\": %g, \"ratio\": %g, \"radius\": %.17g}\n",
           apic ? "apic" : "flip", n, ThreadUtils::getMaxThreadCount(), twarm, t_p2g, t_extrap, t_g2p, t_adv, tot,
           (double)n / tot, sim._CFLConditionNumber, sim._ratioPICFLIP, sim._liquidSDFParticleRadius);
    return 0;
}

// extrapolate: MACVelocityField::extrapolateVelocityField (macvelocityfield.cpp:671-677 ->
// GridUtils::extrapolateGrid, gridutils.h:94-163) on a given field + valid masks.
static int mode_extrapolate() {
    int I = kvi("I"), J = kvi("J"), K = kvi("K");
    double dx = kvd("dx");
    int layers = kvi("layers", "12");
    MACVelocityField f(I, J, K, dx);
    ValidVelocityComponentGrid valid(I, J, K);
    load_grid("in_u", *f.getArray3dU());
    load_grid("in_v", *f.getArray3dV());
    load_grid("in_w", *f.getArray3dW());
    load_mask("in_validu", valid.validU);
    load_mask("in_validv", valid.validV);
    load_mask("in_validw", valid.validW);
    int reps = kvi("reps", "1");
    double t = 0;
    for (int r = 0; r < reps; r++) {
        MACVelocityField g = f;
        double t0 = now();
        g.extrapolateVelocityField(valid, layers);
        t = now() - t0;
        if (r == 0) {
            save_grid("out_u", *g.getArray3dU());
            save_grid("out_v", *g.getArray3dV());
            save_grid("out_w", *g.getArray3dW());
        }
    }
    printf("{\"mode\": \"extrapolate\", \"layers\": %d, \"threads\": %d, \"t_extrapolate\": %.6f}\n", layers,
           ThreadUtils::getMaxThreadCount(), t);
    return 0;
}

// remove: FluidSimulation::_removeMarkerParticles (fluidsimulation.cpp:7773-7851) on given particles and solid SDF.
static int mode_remove() {
    int I = kvi("I"), J = kvi("J"), K = kvi("K");
    double dx = kvd("dx"), dt = kvd("dt");
    FluidSimulation sim(I, J, K, dx);
    sim.disableConsoleOutput();
    shell_sim(sim, I, J, K, dx, false);
    sim._CFLConditionNumber = kvd("cfl", "5");
    sim._maxMarkerParticlesPerCell = kvi("max_per_cell", "250");
    sim._isExtremeVelocityRemovalEnabled = kvi("extreme", "1") != 0;
    const int open = kvi("open", "0");                 // bit 0..5: x-, x+, y-, y+, z-, z+
    sim._openBoundaryXNeg = (open & 1) != 0;  sim._openBoundaryXPos = (open & 2) != 0;
    sim._openBoundaryYNeg = (open & 4) != 0;  sim._openBoundaryYPos = (open & 8) != 0;
    sim._openBoundaryZNeg = (open & 16) != 0; sim._openBoundaryZPos = (open & 32) != 0;
    sim._openBoundaryWidth = kvi("open_width", "2");
    sim._solidSDF = MeshLevelSet(I, J, K, dx);
    load_grid("in_phi", sim._solidSDF._phi);
    size_t before = sim._markerParticles.size();
    double t0 = now();
    sim._removeMarkerParticles(dt);
    double t = now() - t0;
    dump_particles(sim._markerParticles, false, "out_");
    // the boundary box the open-boundary planes are measured from (the planes themselves are the caller's arithmetic)
    AABB box = sim._getBoundaryAABB();
    vmath::vec3 bmin = box.getMinPoint(), bmax = box.getMaxPoint();
    printf("{\"mode\": \"remove\", \"before\": %zu, \"after\": %zu, \"extreme\": %d, \"threads\": %d, \"t_remove\": %.6f, "
           "\"box_min\": [%.9g, %.9g, %.9g], \"box_max\": [%.9g, %.9g, %.9g]}\n",
           before, sim._markerParticles.size(), sim._currentExtremeVelocityParticlesRemoved, ThreadUtils::getMaxThreadCount(), t,
           bmin.x, bmin.y, bmin.z, bmax.x, bmax.y, bmax.z);
    return 0;
}

// liquidsdf: ParticleLevelSet::_computeSignedDistanceFromParticles (particlelevelset.cpp:335-398), what
// calculateSignedDistanceField (:161-168, called from fluidsimulation.cpp:5599) runs on the marker positions.
static int mode_liquidsdf() {
    int I = kvi("I"), J = kvi("J"), K = kvi("K");
    double dx = kvd("dx"), radius = kvd("radius");
    std::vector<vmath::vec3> points = load_vec3("in_pos");
    int reps = kvi("reps", "1");
    double t = 0;
    for (int r = 0; r < reps; r++) {
        ParticleLevelSet ls(I, J, K, dx);
        double t0 = now();
        ls._computeSignedDistanceFromParticles(points, radius);
        t = now() - t0;
        if (r == 0) {
            save_grid("out_phi", ls._phi);
            if (npy_exists(P("in_phi"))) {                 // + postProcessSignedDistanceField against a solid SDF
                MeshLevelSet solid(I, J, K, dx);
                load_grid("in_phi", solid._phi);
                ls.postProcessSignedDistanceField(solid);
                save_grid("out_phi_post", ls._phi);
            }
        }
    }
    printf("{\"mode\": \"liquidsdf\", \"particles\": %zu, \"threads\": %d, \"t_sdf\": %.6f}\n", points.size(),
           ThreadUtils::getMaxThreadCount(), t);
    return 0;
}

// attribute: AttributeToGridTransfer<float>::transfer (attributetogridtransfer.h:52-157) the way
// _updateMarkerParticleAgeAttributeGrid calls it (fluidsimulation.cpp:6990-7015): cell-centred grid, offset dx/2.
static int mode_attribute() {
    int I = kvi("I"), J = kvi("J"), K = kvi("K");
    double dx = kvd("dx"), radius = kvd("radius");
    std::vector<vmath::vec3> positions = load_vec3("in_pos");
    Npy a = npy_load(P("in_attr"));
    std::vector<float> attributes((float *)a.data.data(), (float *)a.data.data() + positions.size());
    Array3d<float> grid(I, J, K, 0.0f);
    Array3d<bool> valid(I, J, K, false);
    AttributeTransferParameters<float> params;
    params.positions = &positions;
    params.attributes = &attributes;
    params.attributeGrid = &grid;
    params.validGrid = &valid;
    params.gridOffset = vmath::vec3(0.5 * dx, 0.5 * dx, 0.5 * dx);
    params.particleRadius = radius;
    params.dx = dx;
    double t0 = now();
    AttributeToGridTransfer<float> transfer;
    transfer.transfer(params);
    double t = now() - t0;
    save_grid("out_grid", grid);
    save_mask("out_valid", valid);
    if (npy_exists(P("in_attr3"))) {                       // the vec3 flavour (colour, whitewater proximity) on the same positions
        std::vector<vmath::vec3> attributes3 = load_vec3("in_attr3");
        Array3d<vmath::vec3> grid3(I, J, K, vmath::vec3());
        Array3d<bool> valid3(I, J, K, false);
        AttributeTransferParameters<vmath::vec3> p3;
        p3.positions = &positions;
        p3.attributes = &attributes3;
        p3.attributeGrid = &grid3;
        p3.validGrid = &valid3;
        p3.gridOffset = vmath::vec3(0.5 * dx, 0.5 * dx, 0.5 * dx);
        p3.particleRadius = radius;
        p3.dx = dx;
        AttributeToGridTransfer<vmath::vec3> transfer3;
        transfer3.transfer(p3);
        std::vector<vmath::vec3> flat((size_t)I * J * K);
        for (int k = 0; k < K; k++)
            for (int j = 0; j < J; j++)
                for (int i = 0; i < I; i++) flat[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)] = grid3(i, j, k);
        save_vec3("out_grid3", flat);
        save_mask("out_valid3", valid3);
    }
    printf("{\"mode\": \"attribute\", \"particles\": %zu, \"threads\": %d, \"t_transfer\": %.6f}\n", positions.size(),
           ThreadUtils::getMaxThreadCount(), t);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: ref_harness <p2g|g2p|advect|scene|extrapolate|remove|liquidsdf|attribute> <workdir> key=value ...\n");
        return 2;
    }
    std::string mode = argv[1];
    g_dir = argv[2];
    for (int i = 3; i < argc; i++) {
        std::string s = argv[i];
        size_t e = s.find('=');
        if (e == std::string::npos) continue;
        g_kv[s.substr(0, e)] = s.substr(e + 1);
    }
    int threads = kvi("threads", "0");
    if (threads > 0) ThreadUtils::setMaxThreadCount(threads);
    if (mode == "p2g") return mode_p2g();
    if (mode == "g2p") return mode_g2p();
    if (mode == "advect") return mode_advect();
    if (mode == "scene") return mode_scene();
    if (mode == "extrapolate") return mode_extrapolate();
    if (mode == "remove") return mode_remove();
    if (mode == "liquidsdf") return mode_liquidsdf();
    if (mode == "attribute") return mode_attribute();
    fprintf(stderr, "ref_harness: unknown mode %s\n", mode.c_str());
    return 2;
}
