// ffengine_b200.cpp -- drop-in libffengine for the FLIP Fluids addon with the particle<->grid
// substep on B200.
//
// The reference calls these stages through the PLT (it is built -fPIC without -Bsymbolic;
// SURVEY.md section 8b), so a library that DEFINES the C++ member symbols below and DT_NEEDEDs the
// unmodified engine takes the calls over without touching reference code:
//
//   VelocityAdvector::advect(VelocityAdvectorParameters)              velocityadvector.cpp:38
//   FluidSimulation::_extrapolateFluidVelocities(MACVelocityField&,   fluidsimulation.cpp:6282
//                                   ValidVelocityComponentGrid&)
//   FluidSimulation::_updateMarkerParticleVelocitiesThread()          fluidsimulation.cpp:6845
//   FluidSimulation::_advanceMarkerParticles(double)                  fluidsimulation.cpp:7853
//   ParticleLevelSet::calculateSignedDistanceField(ParticleSystem&, double)   particlelevelset.cpp:161
//   FluidSimulation::_getMaximumMarkerParticleSpeed()                 fluidsimulation.cpp:10188
//   AttributeToGridTransfer<float>::transfer, <vmath::vec3>::transfer attributetogridtransfer.h:78 (weak template
//                                   instantiations the reference calls through the PLT: explicit specialisations here)
//
// and two bookkeeping hooks that forward to the reference's own definition (dlsym RTLD_NEXT):
//
//   ParticleSystem::getAttributeValuesVector3(ParticleSystemAttribute&)   particlesystem.cpp (the one accessor
//                                   through which ALL host code reaches a vec3 attribute vector)
//   FluidSimulation::initialize()                                     fluidsimulation.cpp:84
//
// RESIDENCY. The marker particles live on the device across the six stages and across substeps. The host's
// std::vector copies of POSITION / VELOCITY / AFFINEX..Z go stale when a stage leaves its results on the device;
// they are brought up to date lazily, when host code outside this file asks for a vec3 attribute vector (the
// accessor hook). Such an access may also WRITE (inflows, sheet seeding, velocity constraints, outflows), so it
// ends the residency: the next stage uploads the host's arrays again. In the plain substep (liquid SDF, P2G,
// extrapolation, G2P, advection, removal, CFL speed) no host code touches the particle vectors, and per substep
// only the MAC field (host pressure solve), the valid masks, the solid SDF, the liquid SDF and -- when something
// was removed -- one byte per particle cross PCIe. Host index i and device id i always name the same particle:
// the device compacts exactly as ParticleSystem::removeParticles does.
//
// ERRORS. A failing GPU call throws std::runtime_error on the thread that called FluidSimulation::update, so the
// reference's C bindings turn it into err = 0 + CBindings_get_error_message (cbindings.h:48-154). Three of the
// members run on std::threads the reference joins at once (fluidsimulation.cpp:5663-5669, 5611-5618); an exception
// there would reach std::terminate, so it is stashed and rethrown by the next member that runs on the caller's
// thread (the G2P of the same substep at the latest).
// Compiled against the UNMODIFIED reference headers with -fno-access-control.
// There is no fallback to the CPU originals: if the GPU call fails, the substep fails.
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "fluidsimulation.h"
#include "gridutils.h"
#include "threadutils.h"
#include "attributetogridtransfer.h"
#include "particlelevelset.h"
#include "stopwatch.h"
#include "velocityadvector.h"

#include "ffb200.h"

namespace {

static_assert(sizeof(vmath::vec3) == 12, "vmath::vec3 must be three packed floats");
static_assert(sizeof(bool) == 1, "Array3d<bool> must be one byte per element");

typedef std::vector<vmath::vec3> Vec3Array;

std::recursive_mutex g_mutex;
std::map<std::tuple<int, int, int, double>, ffb200_context *> g_contexts;

void check(int ok) {
    if (ok != FFB200_SUCCESS) throw std::runtime_error(ffb200_get_error_message());
}

// One device context per grid shape, created on first use. FFB200_DEVICE selects the GPU,
// FFB200_EXACT_P2G=1 sends every face through the reference-order summation (bit-exact P2G).
ffb200_context *context_for(int I, int J, int K, double dx) {
    std::lock_guard<std::recursive_mutex> lock(g_mutex);
    auto key = std::make_tuple(I, J, K, dx);
    auto it = g_contexts.find(key);
    if (it != g_contexts.end()) return it->second;
    const char *dev = std::getenv("FFB200_DEVICE");
    ffb200_context *ctx = nullptr;
    check(ffb200_create(&ctx, I, J, K, dx, dev ? std::atoi(dev) : 0));
    const char *exact = std::getenv("FFB200_EXACT_P2G");
    if (exact && std::atoi(exact) != 0) ffb200_set_valid_guard(ctx, 1e30f, 0.0f);
    g_contexts[key] = ctx;
    return ctx;
}

float *raw(Vec3Array *v) { return (!v || v->empty()) ? nullptr : &((*v)[0].x); }

// ---- deferred errors ----------------------------------------------------------------------------------
std::thread::id g_caller_thread;            // the thread inside FluidSimulation::update / initialize
std::string g_deferred;                     // first error raised on a worker thread since the last rethrow
bool g_has_deferred = false;

void note_caller_thread() {
    std::lock_guard<std::recursive_mutex> lock(g_mutex);
    g_caller_thread = std::this_thread::get_id();
}

void rethrow_deferred() {
    std::lock_guard<std::recursive_mutex> lock(g_mutex);
    if (!g_has_deferred) return;
    g_has_deferred = false;
    throw std::runtime_error(g_deferred);
}

// Runs a stage body; on the caller's thread errors propagate, on any other thread they are stashed.
template <class F>
void run_stage(F &&body) {
    bool on_caller;
    {
        std::lock_guard<std::recursive_mutex> lock(g_mutex);
        on_caller = std::this_thread::get_id() == g_caller_thread;
    }
    if (on_caller) {
        body();
        return;
    }
    try {
        body();
    } catch (const std::exception &e) {
        std::lock_guard<std::recursive_mutex> lock(g_mutex);
        if (!g_has_deferred) g_deferred = e.what();
        g_has_deferred = true;
    }
}

// Failure injection for the error-path tests: FFB200_DROPIN_INJECT=<liquid_sdf|p2g|extrapolate|g2p|advect|max_speed>
// makes that stage throw where a failing GPU call would.
void inject(const char *stage) {
    static const char *which = std::getenv("FFB200_DROPIN_INJECT");
    if (which && std::strcmp(which, stage) == 0) throw std::runtime_error(std::string("ffengine_b200: injected failure in ") + stage);
}

// ---- residency ----------------------------------------------------------------------------------------
thread_local int tl_internal = 0;           // > 0 while this file itself fetches attribute vectors
struct Internal {
    Internal() { tl_internal++; }
    ~Internal() { tl_internal--; }
};

struct Residency {
    ffb200_context *ctx = nullptr;
    ParticleSystem *ps = nullptr;           // the particle system mirrored on the device (nullptr: none)
    size_t count = 0;
    bool affine = false;                    // the device holds the affine rows as well
    bool pos_stale = false, vel_stale = false, aff_stale = false;   // host copies older than the device's
} g_res;

bool g_lazy = true;                         // FFB200_DROPIN_LAZY=0: download every result at once (round-1 behaviour)

struct Vectors {
    Vec3Array *pos = nullptr, *vel = nullptr, *ax = nullptr, *ay = nullptr, *az = nullptr;
};

Vectors vectors_of(ParticleSystem &ps, bool affine) {
    Internal guard;
    Vectors v;
    ps.getAttributeValues("POSITION", v.pos);
    ps.getAttributeValues("VELOCITY", v.vel);
    if (affine) {
        ps.getAttributeValues("AFFINEX", v.ax);
        ps.getAttributeValues("AFFINEY", v.ay);
        ps.getAttributeValues("AFFINEZ", v.az);
    }
    return v;
}

bool has_affine(ParticleSystem &ps) {
    ParticleSystemAttribute att = ps.getAttribute("AFFINEX");
    return att.id != -1;
}

// Bring the host vectors of the mirrored particle system up to date (no-op unless something is stale).
void sync_host_locked() {
    Residency &r = g_res;
    if (!r.ps || !(r.pos_stale || r.vel_stale || r.aff_stale)) return;
    Vectors v = vectors_of(*r.ps, r.affine);
    const bool sizes_ok = v.pos->size() >= r.count && v.vel->size() >= r.count &&
                          (!r.affine || (v.ax->size() >= r.count && v.ay->size() >= r.count && v.az->size() >= r.count));
    if (!sizes_ok) {
        r.ps = nullptr;
        throw std::runtime_error("ffengine_b200: the host shrank the particle system while its results were still on the device");
    }
    if (r.count > 0)
        check(ffb200_get_particles(r.ctx, r.pos_stale ? raw(v.pos) : nullptr, r.vel_stale ? raw(v.vel) : nullptr,
                                   r.aff_stale ? raw(v.ax) : nullptr, r.aff_stale ? raw(v.ay) : nullptr,
                                   r.aff_stale ? raw(v.az) : nullptr));
    r.pos_stale = r.vel_stale = r.aff_stale = false;
}

// Host code outside this file asked for a vec3 attribute vector of `ps`: it may read anything and write
// anything. Make the host copy current, then consider the device copy gone.
void host_access(ParticleSystem *ps) {
    std::lock_guard<std::recursive_mutex> lock(g_mutex);
    if (g_res.ps != ps) return;
    try {
        sync_host_locked();
    } catch (const std::exception &e) {
        if (!g_has_deferred) g_deferred = e.what();
        g_has_deferred = true;
    }
    g_res.ps = nullptr;
}

// The device holds exactly the particles of `ps` (uploading them if it does not).
void ensure_resident(ffb200_context *ctx, ParticleSystem &ps, bool need_affine) {
    std::lock_guard<std::recursive_mutex> lock(g_mutex);
    Residency &r = g_res;
    const bool affine = has_affine(ps);
    if (need_affine && !affine) throw std::runtime_error("ffengine_b200: APIC transfer without AFFINE attributes");
    Vectors v = vectors_of(ps, affine);
    if (r.ps == &ps && r.ctx == ctx && r.count == v.pos->size() && r.affine == affine) return;
    sync_host_locked();                     // results still on the device belong to the system mirrored so far: bring its
    r.ps = nullptr;                         // host copy up to date before the device set is overwritten (or re-read)
    check(ffb200_set_particles(ctx, (int)v.pos->size(), raw(v.pos), raw(v.vel), affine ? raw(v.ax) : nullptr,
                               affine ? raw(v.ay) : nullptr, affine ? raw(v.az) : nullptr));
    r.ctx = ctx;
    r.ps = &ps;
    r.count = v.pos->size();
    r.affine = affine;
    r.pos_stale = r.vel_stale = r.aff_stale = false;
}

void after_device_write(bool pos, bool vel, bool aff) {
    std::lock_guard<std::recursive_mutex> lock(g_mutex);
    g_res.pos_stale |= pos;
    g_res.vel_stale |= vel;
    g_res.aff_stale |= aff && g_res.affine;
    if (!g_lazy) sync_host_locked();
}

// ---- field bookkeeping --------------------------------------------------------------------------------
// The field object whose faces our P2G left on the device without downloading them: the extrapolation
// that always follows (fluidsimulation.cpp:5652-5654, 6971-6973) works on them in place and downloads
// the result. Any other stage that finds it still pending downloads it first.
MACVelocityField *g_pending_p2g_field = nullptr;
ffb200_context *g_pending_ctx = nullptr;
MACVelocityField *g_resident_field = nullptr;    // the field object our G2P uploaded last (advection reuses it)

void flush_pending_field() {
    if (!g_pending_p2g_field) return;
    MACVelocityField *f = g_pending_p2g_field;
    g_pending_p2g_field = nullptr;
    check(ffb200_get_velocity_field(g_pending_ctx, f->getArray3dU()->getRawArray(), f->getArray3dV()->getRawArray(),
                                    f->getArray3dW()->getRawArray(), nullptr, nullptr, nullptr));
}

// Long-lived host arrays that cross PCIe every substep are page-locked once (FFB200_DROPIN_PIN=0: never).
std::map<void *, size_t> g_pinned;
bool g_pin = true;

void pin(ffb200_context *ctx, void *ptr, size_t bytes) {
    if (!g_pin || !ptr || bytes == 0) return;
    std::lock_guard<std::recursive_mutex> lock(g_mutex);
    auto it = g_pinned.find(ptr);
    if (it != g_pinned.end() && it->second == bytes) return;
    if (it != g_pinned.end()) ffb200_unpin_host_memory(ctx, ptr);
    g_pinned[ptr] = ffb200_pin_host_memory(ctx, ptr, bytes) == FFB200_SUCCESS ? bytes : 0;   // failure: stay pageable
}

void unpin_all() {
    std::lock_guard<std::recursive_mutex> lock(g_mutex);
    for (auto &kv : g_pinned)
        if (kv.second && !g_contexts.empty()) ffb200_unpin_host_memory(g_contexts.begin()->second, kv.first);
    g_pinned.clear();
}

void pin_field(ffb200_context *ctx, MACVelocityField &f) {
    Array3d<float> *a[3] = {f.getArray3dU(), f.getArray3dV(), f.getArray3dW()};
    for (int d = 0; d < 3; d++) pin(ctx, a[d]->getRawArray(), (size_t)a[d]->getNumElements() * sizeof(float));
}

// ---- optional stage profile (FFB200_DROPIN_PROFILE=<file>: JSON written at exit) ---------------------------
struct Profile {
    const char *path = nullptr;
    double seconds[6] = {};
    long calls[6] = {};
    long particles = 0;
    ~Profile() {
        if (!path) return;
        FILE *f = std::fopen(path, "w");
        if (!f) return;
        static const char *names[6] = {"liquid_sdf", "p2g", "extrapolate", "g2p", "advect_remove", "max_speed"};
        std::fprintf(f, "{");
        for (int i = 0; i < 6; i++) std::fprintf(f, "\"%s_s\": %.6f, \"%s_calls\": %ld, ", names[i], seconds[i], names[i], calls[i]);
        std::fprintf(f, "\"particle_substeps\": %ld}\n", particles);
        std::fclose(f);
    }
} g_profile;

struct StageClock {
    int id;
    std::chrono::steady_clock::time_point t0;
    explicit StageClock(int i) : id(i), t0(std::chrono::steady_clock::now()) {}
    ~StageClock() {
        if (!g_profile.path) return;
        std::lock_guard<std::recursive_mutex> lock(g_mutex);
        g_profile.seconds[id] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        g_profile.calls[id]++;
    }
};

struct EnvInit {
    EnvInit() {
        const char *e = std::getenv("FFB200_DROPIN_LAZY");
        if (e) g_lazy = std::atoi(e) != 0;
        e = std::getenv("FFB200_DROPIN_PIN");
        if (e) g_pin = std::atoi(e) != 0;
        g_profile.path = std::getenv("FFB200_DROPIN_PROFILE");
    }
} g_env_init;

}  // namespace

// ---- bookkeeping hooks --------------------------------------------------------------------------------
std::vector<vmath::vec3> *ParticleSystem::getAttributeValuesVector3(ParticleSystemAttribute &att) {
    typedef std::vector<vmath::vec3> *(*Fn)(ParticleSystem *, ParticleSystemAttribute &);
    static Fn next = reinterpret_cast<Fn>(dlsym(RTLD_NEXT, "_ZN14ParticleSystem25getAttributeValuesVector3ER23ParticleSystemAttribute"));
    if (!next) throw std::runtime_error("ffengine_b200: the reference's ParticleSystem::getAttributeValuesVector3 was not found");
    if (tl_internal == 0) host_access(this);
    return next(this, att);
}

void FluidSimulation::initialize() {
    typedef void (*Fn)(FluidSimulation *);
    static Fn next = reinterpret_cast<Fn>(dlsym(RTLD_NEXT, "_ZN15FluidSimulation10initializeEv"));
    if (!next) throw std::runtime_error("ffengine_b200: the reference's FluidSimulation::initialize was not found");
    {
        // a new simulation: whatever an earlier one left on the device (or pinned) is not ours any more
        std::lock_guard<std::recursive_mutex> lock(g_mutex);
        g_res = Residency();
        g_pending_p2g_field = nullptr;
        g_resident_field = nullptr;
        g_has_deferred = false;
    }
    unpin_all();
    note_caller_thread();
    next(this);
}

// ---- liquid SDF from particles ------------------------------------------------------------------------
// ParticleLevelSet::calculateSignedDistanceField (particlelevelset.cpp:161-168), called from _updateLiquidLevelSet
// (fluidsimulation.cpp:5599) on a std::thread that _stepFluid joins at once (:10082-10083).
void ParticleLevelSet::calculateSignedDistanceField(ParticleSystem &particles, double radius) {
    run_stage([&] {
        StageClock clock(0);
        inject("liquid_sdf");
        ffb200_context *ctx = context_for(_isize, _jsize, _ksize, _dx);
        flush_pending_field();
        if (particles.getAttribute("VELOCITY").id == -1) {
            // a positions-only scratch system (the upscaling loader, fluidsimulation.cpp:4769-4770): host positions in,
            // nothing kept resident
            std::lock_guard<std::recursive_mutex> lock(g_mutex);
            sync_host_locked();
            g_res.ps = nullptr;
            Internal guard;
            Vec3Array *positions;
            particles.getAttributeValues("POSITION", positions);
            check(ffb200_calculate_signed_distance_field(ctx, (int)positions->size(), raw(positions), radius, _phi.getRawArray()));
            return;
        }
        ensure_resident(ctx, particles, false);
        check(ffb200_declare_resident(ctx, FFB200_RESIDENT_PARTICLES));
        check(ffb200_calculate_signed_distance_field(ctx, (int)g_res.count, nullptr, radius, _phi.getRawArray()));
    });
}

// ---- P2G ----------------------------------------------------------------------------------------------
void VelocityAdvector::advect(VelocityAdvectorParameters params) {
    run_stage([&] {
        StageClock clock(1);
        inject("p2g");
        int I, J, K;
        params.vfield->getGridDimensions(&I, &J, &K);
        ffb200_context *ctx = context_for(I, J, K, params.vfield->getGridCellSize());
        flush_pending_field();
        const bool apic = params.velocityTransferMethod == VelocityAdvectorTransferMethod::APIC;
        ensure_resident(ctx, *params.particles, apic);
        ValidVelocityComponentGrid *valid = params.validVelocities;
        pin_field(ctx, *params.vfield);
        check(ffb200_declare_resident(ctx, FFB200_RESIDENT_PARTICLES));
        // faces stay on the device for the extrapolation that follows; the valid masks go to the host now
        check(ffb200_velocity_advector_advect(
            ctx, (int)g_res.count, nullptr, nullptr, nullptr, nullptr, nullptr, params.particleRadius,
            apic ? FFB200_TRANSFER_APIC : FFB200_TRANSFER_FLIP, nullptr, nullptr, nullptr,
            reinterpret_cast<uint8_t *>(valid->validU.getRawArray()), reinterpret_cast<uint8_t *>(valid->validV.getRawArray()),
            reinterpret_cast<uint8_t *>(valid->validW.getRawArray())));
        g_pending_p2g_field = params.vfield;
        g_pending_ctx = ctx;
        g_resident_field = nullptr;
        if (g_profile.path) g_profile.particles += (long)g_res.count;
    });
}

// ---- valid-face extrapolation -----------------------------------------------------------------------
void FluidSimulation::_extrapolateFluidVelocities(MACVelocityField &MACGrid, ValidVelocityComponentGrid &validVelocities) {
    run_stage([&] {
        StageClock clock(2);
        inject("extrapolate");
        int I, J, K;
        MACGrid.getGridDimensions(&I, &J, &K);
        ffb200_context *ctx = context_for(I, J, K, MACGrid.getGridCellSize());
        const int numLayers = (int)std::ceil(std::sqrt(3) * _CFLConditionNumber) + 3;      // fluidsimulation.cpp:6284
        const bool fresh = g_pending_p2g_field == &MACGrid && g_pending_ctx == ctx;
        if (fresh)
            g_pending_p2g_field = nullptr;      // extrapolated in place on the device, downloaded below
        else
            flush_pending_field();
        check(ffb200_extrapolate_fluid_velocities(
            ctx, MACGrid.getArray3dU()->getRawArray(), MACGrid.getArray3dV()->getRawArray(), MACGrid.getArray3dW()->getRawArray(),
            reinterpret_cast<uint8_t *>(validVelocities.validU.getRawArray()),
            reinterpret_cast<uint8_t *>(validVelocities.validV.getRawArray()),
            reinterpret_cast<uint8_t *>(validVelocities.validW.getRawArray()), numLayers, fresh ? 1 : 0));
        g_resident_field = nullptr;
        // (The device's copy is NOT kept as the FLIP "saved" field: the host constrains _savedVelocityField against
        // the solids after saving it, _constrainVelocityFields fluidsimulation.cpp:6444-6453, and this member runs a
        // second time on the projected field, :6268. The G2P uploads the host's saved field.)
    });
}

// ---- G2P ----------------------------------------------------------------------------------------------
void FluidSimulation::_updateMarkerParticleVelocitiesThread() {
    note_caller_thread();
    rethrow_deferred();
    if (_markerParticles.empty()) return;
    StageClock clock(3);
    inject("g2p");
    ffb200_context *ctx = context_for(_isize, _jsize, _ksize, _dx);
    flush_pending_field();
    const bool apic = _velocityTransferMethod == VelocityTransferMethod::APIC;
    ensure_resident(ctx, _markerParticles, apic);
    pin_field(ctx, _MACVelocity);
    check(ffb200_declare_resident(ctx, FFB200_RESIDENT_PARTICLES));
    // the projected field goes up, the results stay on the device
    check(ffb200_update_marker_particle_velocities(
        ctx, (int)g_res.count, nullptr, nullptr, nullptr, nullptr, nullptr, _MACVelocity.getArray3dU()->getRawArray(),
        _MACVelocity.getArray3dV()->getRawArray(), _MACVelocity.getArray3dW()->getRawArray(),
        apic ? nullptr : _savedVelocityField.getArray3dU()->getRawArray(),
        apic ? nullptr : _savedVelocityField.getArray3dV()->getRawArray(),
        apic ? nullptr : _savedVelocityField.getArray3dW()->getRawArray(),
        apic ? FFB200_TRANSFER_APIC : FFB200_TRANSFER_FLIP, _ratioPICFLIP));
    g_resident_field = &_MACVelocity;
    after_device_write(false, true, apic);
}

// ---- advect + removal ---------------------------------------------------------------------------------
// _removeMarkerParticles (fluidsimulation.cpp:7773-7851), the tail of the advection stage: decided and applied on the
// device, where the advection has just left the positions (and the G2P the velocities); the host applies the same mask to
// every attribute of the particle system with the reference's own ParticleSystem::removeParticles. The lifetime rule reads
// a host-only float attribute and is evaluated here; the open-boundary planes are the reference's float arithmetic on
// its own boundary box. (_constrainMarkerParticleVelocities, :6922-6929, fetches the velocity vector on the host when an
// inflow constrains velocities: the accessor hook has then ended the residency and ensure_resident uploads the
// host's -- constrained -- velocities again before the advection.)
static void remove_marker_particles_b200(FluidSimulation &sim, ffb200_context *ctx, double dt) {
    const size_t n = g_res.count;
    if (n == 0) {
        sim._currentExtremeVelocityParticlesRemoved = 0;
        return;
    }
    AABB boundaryAABB = sim._getBoundaryAABB();
    vmath::vec3 minp = boundaryAABB.getMinPoint();
    vmath::vec3 maxp = boundaryAABB.getMaxPoint();
    float buffer = sim._openBoundaryWidth * sim._dx;
    const float inf = std::numeric_limits<float>::infinity();
    const float bounds[6] = {sim._openBoundaryXNeg ? minp.x + buffer : -inf, sim._openBoundaryXPos ? maxp.x - buffer : inf,
                             sim._openBoundaryYNeg ? minp.y + buffer : -inf, sim._openBoundaryYPos ? maxp.y - buffer : inf,
                             sim._openBoundaryZNeg ? minp.z + buffer : -inf, sim._openBoundaryZPos ? maxp.z - buffer : inf};
    const bool closed = !sim._openBoundaryXNeg && !sim._openBoundaryXPos && !sim._openBoundaryYNeg && !sim._openBoundaryYPos &&
                        !sim._openBoundaryZNeg && !sim._openBoundaryZPos;
    std::vector<uint8_t> dead;
    if (sim._isSurfaceLifetimeAttributeEnabled || sim._isFluidParticleLifetimeAttributeEnabled) {
        std::vector<float> *lifetimes = nullptr;
        sim._markerParticles.getAttributeValues("LIFETIME", lifetimes);
        dead.resize(n);
        float eps = 1e-6f;
        for (size_t i = 0; i < n; i++) dead[i] = lifetimes->at(i) <= sim._surfaceLifetimeAttributeDeathTime + eps ? 1 : 0;
    }
    std::vector<uint8_t> mask(n);
    int remaining = (int)n, extreme = 0;
    check(ffb200_remove_marker_particles_masked(ctx, closed ? nullptr : bounds, dead.empty() ? nullptr : dead.data(), dt,
                                                sim._CFLConditionNumber, sim._maxMarkerParticlesPerCell, sim._maxFrameTimeSteps,
                                                sim._isExtremeVelocityRemovalEnabled ? 1 : 0, mask.data(), &remaining, &extreme));
    if ((size_t)remaining != n) {
        std::vector<bool> isRemoved(n);
        for (size_t i = 0; i < n; i++) isRemoved[i] = mask[i] != 0;
        // compacts every attribute vector of the host system (stale vec3 copies included: they keep their size
        // and order, which is all the lazy download needs). removeParticles (particlesystem.cpp:88-99) walks the
        // attribute vectors directly, so the accessor hook does not fire.
        sim._markerParticles.removeParticles(isRemoved);
        std::lock_guard<std::recursive_mutex> lock(g_mutex);
        g_res.count = (size_t)remaining;
    }
    sim._currentExtremeVelocityParticlesRemoved = extreme;
}

// Same bracket as the reference stage (log lines, the advanceMarkerParticles timer the addon's
// stats read, fluidsimulation.cpp:7896-7897) and the same tail: particle removal runs inside this stage.
void FluidSimulation::_advanceMarkerParticles(double dt) {
    note_caller_thread();
    rethrow_deferred();
    _logfile.logString(_logfile.getTime() + " BEGIN       Advect Marker Particles");
    StopWatch timer;
    timer.start();
    if (_isFluidInSimulation()) {
        StageClock clock(4);
        inject("advect");
        ffb200_context *ctx = context_for(_isize, _jsize, _ksize, _dx);
        flush_pending_field();
        ensure_resident(ctx, _markerParticles, false);
        // directly after our G2P of the same field: nothing to upload but the solid
        const bool field_resident = g_resident_field == &_MACVelocity;
        g_resident_field = nullptr;
        pin(ctx, _solidSDF._phi.getRawArray(), (size_t)_solidSDF._phi.getNumElements() * sizeof(float));
        check(ffb200_declare_resident(ctx, FFB200_RESIDENT_PARTICLES | (field_resident ? FFB200_RESIDENT_FIELD : 0u)));
        check(ffb200_advance_marker_particles(
            ctx, (int)g_res.count, nullptr, _MACVelocity.getArray3dU()->getRawArray(), _MACVelocity.getArray3dV()->getRawArray(),
            _MACVelocity.getArray3dW()->getRawArray(), _solidSDF._phi.getRawArray(),
            reinterpret_cast<uint8_t *>(_nearSolidGrid.getRawArray()), dt, _CFLConditionNumber));
        after_device_write(true, false, false);
        remove_marker_particles_b200(*this, ctx, _currentFrameDeltaTime);
    }
    timer.stop();
    _timingData.advanceMarkerParticles += timer.getTime();
    _logfile.logString(_logfile.getTime() + " COMPLETE    Advect Marker Particles");
}

// ---- CFL input ----------------------------------------------------------------------------------------
// _getMaximumMarkerParticleSpeed (fluidsimulation.cpp:10188-10202), called from update() before every substep: the
// reduction runs on the resident velocities (a maximum is order independent: the same double comes back).
double FluidSimulation::_getMaximumMarkerParticleSpeed() {
    note_caller_thread();
    rethrow_deferred();
    if (_markerParticles.empty()) return 0.0;
    StageClock clock(5);
    inject("max_speed");
    ffb200_context *ctx = context_for(_isize, _jsize, _ksize, _dx);
    flush_pending_field();
    ensure_resident(ctx, _markerParticles, false);
    double speed = 0.0;
    check(ffb200_get_maximum_particle_speed(ctx, &speed));
    return speed;
}

// ---- attribute transfers ------------------------------------------------------------------------------
// AttributeToGridTransfer<T>::transfer (attributetogridtransfer.h:78-157): the age / lifetime / viscosity / density /
// colour / whitewater-proximity grids (fluidsimulation.cpp:4636-4742, 6010, 7011-7170). The transfer uploads its own
// positions and payload and replaces the device's particle set, so the marker-particle residency ends here (its host
// copy is brought up to date first); these run once per frame, and only when an attribute is enabled.
namespace {

template <class T>
void attribute_transfer_b200(AttributeTransferParameters<T> &params, int ncomp) {
    run_stage([&] {
        Array3d<T> *grid = params.attributeGrid;
        Array3d<bool> *validGrid = params.validGrid;
        const int I = grid->width, J = grid->height, K = grid->depth;
        const float h = (float)(0.5 * params.dx);
        if (params.gridOffset.x != h || params.gridOffset.y != h || params.gridOffset.z != h)
            throw std::runtime_error("ffengine_b200: attribute transfer with a grid offset other than (dx/2, dx/2, dx/2)");
        const size_t n = params.positions->size();
        if (n == 0) return;                                     // no block holds particles: nothing is written (:96-99)
        inject("attribute");
        ffb200_context *ctx = context_for(I, J, K, params.dx);
        flush_pending_field();
        {
            std::lock_guard<std::recursive_mutex> lock(g_mutex);
            sync_host_locked();
            g_res.ps = nullptr;
        }
        const size_t cells = (size_t)I * J * K;
        std::vector<float> out(cells * ncomp);
        std::vector<uint8_t> valid(cells);
        check(ffb200_attribute_to_grid_transfer(ctx, (int)n, raw(params.positions), reinterpret_cast<const float *>(params.attributes->data()),
                                                ncomp, params.particleRadius, params.normalize ? 1 : 0, out.data(), valid.data()));
        // the reference writes the cells of the blocks it processed and only ever SETS valid flags (:128-141)
        float *dst = reinterpret_cast<float *>(grid->getRawArray());
        bool *vdst = validGrid->getRawArray();
        for (size_t i = 0; i < cells; i++) {
            bool touched = valid[i] != 0;
            for (int q = 0; q < ncomp; q++) touched = touched || out[i * ncomp + q] != 0.0f;
            if (touched)
                for (int q = 0; q < ncomp; q++) dst[i * ncomp + q] = out[i * ncomp + q];
            if (valid[i]) vdst[i] = true;
        }
    });
}

}  // namespace

static_assert(sizeof(Array3d<vmath::vec3>) > 0 && sizeof(vmath::vec3) == 3 * sizeof(float), "interleaved vec3 grid");

template <>
void AttributeToGridTransfer<float>::transfer(AttributeTransferParameters<float> params) {
    attribute_transfer_b200(params, 1);
}

template <>
void AttributeToGridTransfer<vmath::vec3>::transfer(AttributeTransferParameters<vmath::vec3> params) {
    attribute_transfer_b200(params, 3);
}
