// ffengine_b200.cpp -- drop-in libffengine for the FLIP Fluids addon with the particle<->grid
// substep on B200.
//
// The reference calls these four stages through the PLT (it is built -fPIC without
// -Bsymbolic), so a library that DEFINES those four C++ member symbols
// and DT_NEEDEDs the unmodified engine takes the calls over without touching reference code:
//
//   VelocityAdvector::advect(VelocityAdvectorParameters)              velocityadvector.cpp:38
//   FluidSimulation::_extrapolateFluidVelocities(MACVelocityField&,   fluidsimulation.cpp:6282
//                                   ValidVelocityComponentGrid&)
//   FluidSimulation::_updateMarkerParticleVelocitiesThread()          fluidsimulation.cpp:6845
//   FluidSimulation::_advanceMarkerParticles(double)                  fluidsimulation.cpp:7853
//
// Each definition marshals the reference's own host containers (std::vector<vmath::vec3>,
// Array3d<float>, Array3d<bool>) into the C ABI of libffb200.so (include/ffb200.h) and throws
// std::runtime_error on failure on the calling thread, so the reference's C bindings turn it
// into err = 0 + CBindings_get_error_message (cbindings.h:48-154). The marker-particle removal at the tail of the
// advection stage (fluidsimulation.cpp:7892, its only call site) is decided on the device as well. Every other
// stage (pressure, level sets, meshing, I/O ...) stays on the reference CPU code.
// Compiled against the UNMODIFIED reference headers with -fno-access-control.
// There is no fallback to the CPU originals: if the GPU call fails, the substep fails.
#include <cmath>
#include <cstdlib>
#include <limits>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "fluidsimulation.h"
#include "stopwatch.h"
#include "velocityadvector.h"

#include "ffb200.h"

namespace {

static_assert(sizeof(vmath::vec3) == 12, "vmath::vec3 must be three packed floats");
static_assert(sizeof(bool) == 1, "Array3d<bool> must be one byte per element");

std::mutex g_mutex;
std::map<std::tuple<int, int, int, double>, ffb200_context *> g_contexts;

// One device context per grid shape, created on first use. FFB200_DEVICE selects the GPU,
// FFB200_EXACT_P2G=1 sends every face through the reference-order summation (bit-exact P2G).
ffb200_context *context_for(int I, int J, int K, double dx) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto key = std::make_tuple(I, J, K, dx);
    auto it = g_contexts.find(key);
    if (it != g_contexts.end()) return it->second;
    const char *dev = std::getenv("FFB200_DEVICE");
    ffb200_context *ctx = nullptr;
    if (ffb200_create(&ctx, I, J, K, dx, dev ? std::atoi(dev) : 0) != FFB200_SUCCESS)
        throw std::runtime_error(ffb200_get_error_message());
    const char *exact = std::getenv("FFB200_EXACT_P2G");
    if (exact && std::atoi(exact) != 0) ffb200_set_valid_guard(ctx, 1e30f, 0.0f);
    g_contexts[key] = ctx;
    return ctx;
}

void check(int ok) {
    if (ok != FFB200_SUCCESS) throw std::runtime_error(ffb200_get_error_message());
}

float *raw(std::vector<vmath::vec3> *v) { return v->empty() ? nullptr : &((*v)[0].x); }

// The field object whose host arrays were just filled by our P2G (and are therefore identical to
// the device-resident field): lets the extrapolation that follows it (fluidsimulation.cpp:5652-5654)
// skip the upload. Cleared by every other interposed stage.
MACVelocityField *g_fresh_p2g_field = nullptr;

// Resident-input tracking (ffb200_declare_resident): the particle system whose arrays our P2G
// uploaded last and how many particles it held, and the field object our G2P uploaded last.
// fluidsimulation.cpp:10078-10121: between the P2G and the G2P only sheet seeding can add or remove
// marker particles, between the G2P and the advection nothing touches particles or _MACVelocity.
ParticleSystem *g_resident_particles = nullptr;
size_t g_resident_count = 0;
MACVelocityField *g_resident_field = nullptr;

}  // namespace

// ---- P2G ------------------------------------------------------------------------------------------
void VelocityAdvector::advect(VelocityAdvectorParameters params) {
    int I, J, K;
    params.vfield->getGridDimensions(&I, &J, &K);
    ffb200_context *ctx = context_for(I, J, K, params.vfield->getGridCellSize());

    std::vector<vmath::vec3> *pos, *vel, *ax = nullptr, *ay = nullptr, *az = nullptr;
    params.particles->getAttributeValues("POSITION", pos);
    params.particles->getAttributeValues("VELOCITY", vel);
    const bool apic = params.velocityTransferMethod == VelocityAdvectorTransferMethod::APIC;
    if (apic) {
        params.particles->getAttributeValues("AFFINEX", ax);
        params.particles->getAttributeValues("AFFINEY", ay);
        params.particles->getAttributeValues("AFFINEZ", az);
    }
    ValidVelocityComponentGrid *valid = params.validVelocities;
    check(ffb200_velocity_advector_advect(
        ctx, (int)pos->size(), raw(pos), raw(vel), apic ? raw(ax) : nullptr, apic ? raw(ay) : nullptr,
        apic ? raw(az) : nullptr, params.particleRadius, apic ? FFB200_TRANSFER_APIC : FFB200_TRANSFER_FLIP,
        params.vfield->getArray3dU()->getRawArray(), params.vfield->getArray3dV()->getRawArray(),
        params.vfield->getArray3dW()->getRawArray(), reinterpret_cast<uint8_t *>(valid->validU.getRawArray()),
        reinterpret_cast<uint8_t *>(valid->validV.getRawArray()), reinterpret_cast<uint8_t *>(valid->validW.getRawArray())));
    g_fresh_p2g_field = params.vfield;
    g_resident_particles = params.particles;
    g_resident_count = pos->size();
    g_resident_field = nullptr;
}

// ---- valid-face extrapolation -----------------------------------------------------------------------
void FluidSimulation::_extrapolateFluidVelocities(MACVelocityField &MACGrid, ValidVelocityComponentGrid &validVelocities) {
    int I, J, K;
    MACGrid.getGridDimensions(&I, &J, &K);
    ffb200_context *ctx = context_for(I, J, K, MACGrid.getGridCellSize());
    const int numLayers = (int)std::ceil(std::sqrt(3) * _CFLConditionNumber) + 3;      // fluidsimulation.cpp:6284
    const bool fresh = g_fresh_p2g_field == &MACGrid;
    g_fresh_p2g_field = nullptr;
    check(ffb200_extrapolate_fluid_velocities(
        ctx, MACGrid.getArray3dU()->getRawArray(), MACGrid.getArray3dV()->getRawArray(), MACGrid.getArray3dW()->getRawArray(),
        reinterpret_cast<uint8_t *>(validVelocities.validU.getRawArray()),
        reinterpret_cast<uint8_t *>(validVelocities.validV.getRawArray()),
        reinterpret_cast<uint8_t *>(validVelocities.validW.getRawArray()), numLayers, fresh ? 1 : 0));
}

// ---- G2P ------------------------------------------------------------------------------------------
void FluidSimulation::_updateMarkerParticleVelocitiesThread() {
    g_fresh_p2g_field = nullptr;
    if (_markerParticles.empty()) return;
    ffb200_context *ctx = context_for(_isize, _jsize, _ksize, _dx);
    std::vector<vmath::vec3> *pos, *vel, *ax = nullptr, *ay = nullptr, *az = nullptr;
    _markerParticles.getAttributeValues("POSITION", pos);
    _markerParticles.getAttributeValues("VELOCITY", vel);
    const bool apic = _velocityTransferMethod == VelocityTransferMethod::APIC;
    if (apic) {
        _markerParticles.getAttributeValues("AFFINEX", ax);
        _markerParticles.getAttributeValues("AFFINEY", ay);
        _markerParticles.getAttributeValues("AFFINEZ", az);
    }
    const bool same_particles = !_isSheetSeedingEnabled && g_resident_particles == &_markerParticles &&
                                g_resident_count == pos->size();
    if (same_particles) check(ffb200_declare_resident(ctx, FFB200_RESIDENT_PARTICLES));
    g_resident_particles = &_markerParticles;
    g_resident_count = pos->size();
    check(ffb200_update_marker_particle_velocities(
        ctx, (int)pos->size(), raw(pos), raw(vel), apic ? raw(ax) : nullptr, apic ? raw(ay) : nullptr,
        apic ? raw(az) : nullptr, _MACVelocity.getArray3dU()->getRawArray(), _MACVelocity.getArray3dV()->getRawArray(),
        _MACVelocity.getArray3dW()->getRawArray(), apic ? nullptr : _savedVelocityField.getArray3dU()->getRawArray(),
        apic ? nullptr : _savedVelocityField.getArray3dV()->getRawArray(),
        apic ? nullptr : _savedVelocityField.getArray3dW()->getRawArray(),
        apic ? FFB200_TRANSFER_APIC : FFB200_TRANSFER_FLIP, _ratioPICFLIP));
    g_resident_field = &_MACVelocity;
}

// ---- advect ---------------------------------------------------------------------------------------
// _removeMarkerParticles (fluidsimulation.cpp:7773-7851), the tail of the advection stage: the decisions are taken
// on the device, where the advection has just left the positions (and the G2P the velocities); the host applies the
// mask to every attribute of the particle system with the reference's own ParticleSystem::removeParticles. The
// lifetime rule reads a host-only attribute and is evaluated here, the open-boundary planes are the reference's
// float arithmetic on its own boundary box.
// _updateMarkerParticleVelocities (fluidsimulation.cpp:6931-6944) runs _constrainMarkerParticleVelocities on the HOST
// after the G2P: with an enabled inflow that constrains fluid velocities (:6922-6929) the host velocities are no
// longer the ones the G2P left on the device, and the removal's speed rules must see the host's.
static bool host_velocities_may_differ(FluidSimulation &sim) {
    for (size_t i = 0; i < sim._meshFluidSources.size(); i++) {
        MeshFluidSource *source = sim._meshFluidSources[i];
        if (source->isEnabled() && source->isInflow() && source->isConstrainedFluidVelocityEnabled()) return true;
    }
    return false;
}

static void remove_marker_particles_b200(FluidSimulation &sim, ffb200_context *ctx, bool particles_resident, double dt) {
    std::vector<vmath::vec3> *pos, *vel;
    sim._markerParticles.getAttributeValues("POSITION", pos);
    sim._markerParticles.getAttributeValues("VELOCITY", vel);
    const size_t n = pos->size();
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
    int removed = 0, extreme = 0;
    check(ffb200_declare_resident(ctx, FFB200_RESIDENT_SOLID | (particles_resident ? FFB200_RESIDENT_PARTICLES : 0u)));
    check(ffb200_mark_removed_marker_particles(ctx, (int)n, raw(pos), raw(vel), nullptr, nullptr, closed ? nullptr : bounds,
                                               dead.empty() ? nullptr : dead.data(), dt, sim._CFLConditionNumber,
                                               sim._maxMarkerParticlesPerCell, sim._maxFrameTimeSteps,
                                               sim._isExtremeVelocityRemovalEnabled ? 1 : 0, mask.data(), &removed, &extreme));
    std::vector<bool> isRemoved(n);
    for (size_t i = 0; i < n; i++) isRemoved[i] = mask[i] != 0;
    sim._markerParticles.removeParticles(isRemoved);
    sim._currentExtremeVelocityParticlesRemoved = extreme;
}

// Same bracket as the reference stage (log lines, the advanceMarkerParticles timer the addon's
// stats read, fluidsimulation.cpp:7896-7897) and the same tail: particle removal runs inside this stage.
void FluidSimulation::_advanceMarkerParticles(double dt) {
    _logfile.logString(_logfile.getTime() + " BEGIN       Advect Marker Particles");
    StopWatch timer;
    timer.start();
    g_fresh_p2g_field = nullptr;
    if (_isFluidInSimulation()) {
        ffb200_context *ctx = context_for(_isize, _jsize, _ksize, _dx);
        std::vector<vmath::vec3> *pos;
        _markerParticles.getAttributeValues("POSITION", pos);
        // directly after our G2P of the same particle system and field: nothing to upload but the solid
        unsigned resident = 0;
        if (g_resident_particles == &_markerParticles && g_resident_count == pos->size() && g_resident_field == &_MACVelocity)
            resident = FFB200_RESIDENT_PARTICLES | FFB200_RESIDENT_FIELD;
        g_resident_particles = nullptr;
        g_resident_field = nullptr;
        if (resident) check(ffb200_declare_resident(ctx, resident));
        check(ffb200_advance_marker_particles(
            ctx, (int)pos->size(), raw(pos), _MACVelocity.getArray3dU()->getRawArray(),
            _MACVelocity.getArray3dV()->getRawArray(), _MACVelocity.getArray3dW()->getRawArray(),
            _solidSDF._phi.getRawArray(), reinterpret_cast<uint8_t *>(_nearSolidGrid.getRawArray()), dt,
            _CFLConditionNumber));
        // device velocities are the host's only if this advection found the G2P's particles resident
        remove_marker_particles_b200(*this, ctx, resident != 0 && !host_velocities_may_differ(*this), _currentFrameDeltaTime);
    }
    timer.stop();
    _timingData.advanceMarkerParticles += timer.getTime();
    _logfile.logString(_logfile.getTime() + " COMPLETE    Advect Marker Particles");
}

// ---- liquid SDF from particles (next symbol; compiled only with -DFFB200_DROPIN_LIQUID_SDF until it has run on hardware) ----
// ParticleLevelSet::calculateSignedDistanceField (particlelevelset.cpp:161-168), called from _updateLiquidLevelSet
// (fluidsimulation.cpp:5599) on a std::thread that _stepFluid joins at once (:10082-10083), so never concurrently with
// the stages above. It leaves positions only on the device: the resident-input tracking is reset.
#ifdef FFB200_DROPIN_LIQUID_SDF
#include "particlelevelset.h"
void ParticleLevelSet::calculateSignedDistanceField(ParticleSystem &particles, double radius) {
    std::vector<vmath::vec3> *positions;
    particles.getAttributeValues("POSITION", positions);
    ffb200_context *ctx = context_for(_isize, _jsize, _ksize, _dx);
    g_fresh_p2g_field = nullptr;
    g_resident_particles = nullptr;
    g_resident_field = nullptr;
    check(ffb200_calculate_signed_distance_field(ctx, (int)positions->size(), raw(positions), radius, _phi.getRawArray()));
}
#endif
