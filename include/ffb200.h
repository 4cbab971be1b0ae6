/* ffb200.h -- C ABI of libffb200.so, the B200 (sm_100a) particle<->grid substep of the FLIP
 * Fluids engine (rlguy/Blender-FLIP-Fluids src/engine v1.8.5).
 *
 * The reference has no C-ABI seam for this path: its ctypes surface stops at
 * FluidSimulation_update (c_bindings/fluidsimulation_c.cpp:115) and the three hot stages are
 * C++ members reached from FluidSimulation::_stepFluid (fluidsimulation.cpp:10078-10121).
 * This header is the seam a maintainer binds instead (INTEGRATION.md shows the C++ interposer
 * for libffengine and the ctypes stub): one entry point per reference call site, plain
 * pointers and sizes, host buffers in the reference's own layouts.
 *
 *   reference call site                                         replaced by
 *   ---------------------------------------------------------   -------------------------------
 *   VelocityAdvector::advect(params)   velocityadvector.cpp:38  ffb200_velocity_advector_advect
 *     (called from fluidsimulation.cpp:5652 and :6971)
 *   FluidSimulation::_extrapolateFluidVelocities
 *                                             fs.cpp:6282-6286  ffb200_extrapolate_fluid_velocities
 *   FluidSimulation::_saveVelocityField       fs.cpp:5671-5679  ffb200_save_velocity_field
 *   FluidSimulation::_updateMarkerParticleVelocitiesThread
 *                                             fs.cpp:6845-6863  ffb200_update_marker_particle_velocities
 *   FluidSimulation::_advanceMarkerParticles  fs.cpp:7853-7890  ffb200_advance_marker_particles
 *     (the RK3 + _resolveCollision fan-out; _removeMarkerParticles stays on the CPU)
 *
 * Layouts (all little-endian, tightly packed):
 *   particle attributes  float[3] per particle, POSITION / VELOCITY / AFFINEX / AFFINEY / AFFINEZ
 *                        (std::vector<vmath::vec3>, particlesystem.h:303-325, vmath.h:37-61)
 *   MAC faces            x-fastest Array3d<float>, flat = i + w*(j + h*k)  (array3d.h:774-777):
 *                        u (I+1)*J*K, v I*(J+1)*K, w I*J*(K+1)       (macvelocityfield.cpp:46-54)
 *   valid masks          one byte per face, same dims (ValidVelocityComponentGrid, mac.h:35-50)
 *   solid SDF            node-centred float (I+1)*(J+1)*(K+1)          (meshlevelset.cpp:35-42)
 *   near-solid mask      one byte per 3dx cell, dims ceil(I/3) x ceil(J/3) x ceil(K/3)
 *                        (fluidsimulation.cpp:5448-5451)
 *
 * Error convention, mirroring c_bindings/cbindings.cpp:35-47: every call returns
 * FFB200_SUCCESS (1) or FFB200_FAIL (0) -- the same values the reference writes through its
 * trailing `int *err` -- and the message of the last failure is kept in one global buffer
 * read by ffb200_get_error_message() (cf. CBindings_get_error_message, cbindings.cpp:100-103).
 * There is no CPU fallback: without a CUDA device ffb200_create fails.
 */
#ifndef FFB200_H
#define FFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFB200_SUCCESS 1
#define FFB200_FAIL 0

/* VelocityAdvectorTransferMethod (velocityadvector.h:63-66) */
#define FFB200_TRANSFER_FLIP 0
#define FFB200_TRANSFER_APIC 1

typedef struct ffb200_context ffb200_context;

/* Device milliseconds (CUDA events on the context's stream) of the most recent call of each
 * stage, and how many kernels of this library that call launched. */
typedef struct ffb200_timing {
    float sort_ms;        /* cell keys + radix sort + bin table + SoA reorder */
    float p2g_prep_ms;    /* block masks + 10^3-block membership words */
    float p2g_ms;         /* the three transfer kernels (U, V, W) */
    float g2p_ms;
    float advect_ms;
    float h2d_ms, d2h_ms; /* host<->device copies inside the last host-buffer call */
    int sort_launches, p2g_prep_launches, p2g_launches, g2p_launches, advect_launches;
} ffb200_timing;

/* ---- lifetime ------------------------------------------------------------------------------- */

/* Grid of isize x jsize x ksize cells of size dx on CUDA device `device`
 * (cf. FluidSimulation_new_from_dimensions, fluidsimulation_c.cpp:56). */
int ffb200_create(ffb200_context **ctx, int isize, int jsize, int ksize, double dx, int device);

/* z-slab flavour for one rank of a multi-GPU run: the context stores cell planes
 * [k_begin - halo, k_end + halo) clipped to [0, ksize) and owns [k_begin, k_end). */
int ffb200_create_slab(ffb200_context **ctx, int isize, int jsize, int ksize, double dx, int device,
                       int k_begin, int k_end, int halo);

void ffb200_destroy(ffb200_context *ctx);
const char *ffb200_get_error_message(void);
int ffb200_get_version(int *major, int *minor, int *revision);

/* Run the context's work on an existing CUDA stream (a cudaStream_t passed as void*), e.g.
 * torch.cuda.current_stream().cuda_stream; the value is taken literally, so 0 is the legacy
 * default stream. ffb200_reset_stream goes back to the context's own non-blocking stream. */
int ffb200_set_stream(ffb200_context *ctx, void *cuda_stream);
int ffb200_reset_stream(ffb200_context *ctx);
/* Fixed-batch mode (benchmarking): every ffb200_p2g re-bins and re-sorts the resident batch,
 * and G2P / advection write their results into the spare SoA buffer instead of in place, so
 * each step sees identical inputs. Off by default (reference semantics: in-place update). */
int ffb200_set_fixed_batch(ffb200_context *ctx, int on);
int ffb200_synchronize(ffb200_context *ctx);
int ffb200_get_timing(ffb200_context *ctx, ffb200_timing *out);

/* Guard band of the valid-face test (DESIGN.md "valid masks"): faces whose fast weight sum
 * lies within abs_tol + per_contrib_tol * contributions of the 1e-6 threshold are re-summed
 * in the reference's exact order. Negative values restore the defaults; abs_tol = +inf
 * sends every face through the exact path (bit-exact P2G, slow; used by the tests). */
int ffb200_set_valid_guard(ffb200_context *ctx, float abs_tol, float per_contrib_tol);

/* ---- resident particle state ------------------------------------------------------------------ */

/* Upload n particles (host AoS, reference order). affx/affy/affz may be NULL (FLIP). */
int ffb200_set_particles(ffb200_context *ctx, int n, const float *pos, const float *vel,
                         const float *affx, const float *affy, const float *affz);
/* Download in the ORIGINAL particle order; any pointer may be NULL. */
int ffb200_get_particles(ffb200_context *ctx, float *pos, float *vel, float *affx, float *affy, float *affz);
int ffb200_get_num_particles(ffb200_context *ctx, int *n);

/* ---- device-resident access (multi-GPU plumbing: halo exchange and particle migration) ---------- */

/* Raw DEVICE pointers of the context's resident arrays, valid until the next call that sorts,
 * reallocates or destroys (a sort flips the particle double buffer). The slab driver wraps them
 * as zero-copy tensors so NCCL can send/receive grid planes and particle streams in place. */
typedef struct ffb200_device_buffers {
    float *pos[3], *vel[3], *aff[9];   /* particle SoA streams, `capacity` floats each (aff may be NULL) */
    uint32_t *ids;                     /* original / global particle id per slot */
    int n, capacity;
    float *field[3], *saved[3];        /* u, v, w components: face_count[d] floats, stored planes only */
    uint8_t *valid[3];
    long long face_count[3];
    int face_plane[3];                 /* faces per z-plane (gi*gj) */
    int kbase, kloc;                   /* first stored cell plane, number of stored cell planes */
    int k_own_begin, k_own_end;
    float *phi;
} ffb200_device_buffers;
int ffb200_get_device_buffers(ffb200_context *ctx, ffb200_device_buffers *out);
/* Make room for `capacity` particles (contents are preserved), with affine streams if asked. */
int ffb200_reserve_particles(ffb200_context *ctx, int capacity, int with_affine);
/* Declare that the streams now hold n particles written through the device pointers (ids
 * included); the next P2G re-bins and re-sorts them. */
int ffb200_set_num_particles(ffb200_context *ctx, int n, int has_affine);

/* Particles travel between slabs as packed records of `floats_per_particle` floats (6 or 15
 * attribute components + the global id bit pattern), `count` records back to back in a DEVICE
 * buffer of block_capacity records followed by a 4-int header {count, overflowed, 0, 0} that the
 * kernels fill in, so a fixed-size buffer can be sent without a host round trip for its size. */
int ffb200_slab_record_floats(ffb200_context *ctx, int *floats_per_particle);
/* Copy the particles whose cell plane k = floor(z/dx) lies in [lo_a, hi_a) into block_a and
 * those in [lo_b, hi_b) into block_b (ghost layers for the neighbours' P2G). Asynchronous. */
int ffb200_slab_pack_layers(ffb200_context *ctx, int lo_a, int hi_a, float *block_a, int lo_b, int hi_b,
                            float *block_b, int block_capacity);
/* Split the resident particles by cell plane: k in [k_begin, k_end) stay (compacted), k >= k_end
 * are moved to block_up, k < k_begin to block_down; a NULL block drops those particles. Ghost
 * copies are always dropped. counts = {stay, up, down}; synchronises the stream. */
int ffb200_slab_route(ffb200_context *ctx, int k_begin, int k_end, float *block_up, float *block_down,
                      int block_capacity, int *counts);
/* The same in two halves, so the neighbour exchange of the packed buffers can be issued while
 * the routing kernel runs: _begin is asynchronous, _end synchronises, fills the holes the
 * leavers left and returns counts = {stay, up, down}. */
int ffb200_slab_route_begin(ffb200_context *ctx, int k_begin, int k_end, float *block_up, float *block_down,
                            int block_capacity);
int ffb200_slab_route_end(ffb200_context *ctx, int *counts);
/* ffb200_slab_route_begin with the NEXT substep's ghost exchange folded in, so that one neighbour
 * exchange per substep carries migrants and ghost copies. A block holds two sections,
 * [migrant records][ghost-copy records], followed by an 8-int header
 * {migrants, overflow, ghosts, overflow, leaving, 0, 0, 0} (true counts even on overflow; `leaving` =
 * particles removed from the resident set, for ffb200_slab_route_end_known).
 * capacities = {up migrants, up ghosts, down migrants, down ghosts}: per face, so that both ranks of
 * a face can size their buffers from numbers they both know (the previous exchange's headers).
 * Owned particles that stay within ghost_layers cell planes of a face are copied into the ghost
 * section; a migrant landing within ghost_layers planes beyond the face is sent and also kept here
 * as a ghost (its new owner would return it as one); a migrant that does not fit its section stays
 * with the sender for this substep; the ghosts of the substep that ended are dropped. Finish with
 * ffb200_slab_route_end, then append the received sections (ffb200_slab_append, as_ghost = 0 / 1). */
int ffb200_slab_route_ghosts_begin(ffb200_context *ctx, int k_begin, int k_end, int ghost_layers,
                                   float *block_up, float *block_down, const int *capacities);
/* ffb200_slab_route_end for a caller that already knows how many particles leave the resident set:
 * header word 4 of either block written by ffb200_slab_route_ghosts_begin, which the caller reads
 * together with the received counts after the exchange. Saves the second host synchronisation.
 * counts[1], counts[2] (migrants sent up / down) are not filled: they are in the headers too. */
int ffb200_slab_route_end_known(ffb200_context *ctx, int leaving, int *counts);
/* Append `count` packed records (device buffer) to the resident particles. as_ghost marks them
 * (top id bit) as ghost copies: they take part in P2G and are dropped by the next
 * ffb200_slab_route wherever they have moved. */
int ffb200_slab_append(ffb200_context *ctx, const float *block, int count, int as_ghost);

/* Cell binning + stable sort (runs implicitly before P2G when positions changed). */
int ffb200_sort_particles(ffb200_context *ctx);
/* Per particle in ORIGINAL order: cell = flat reference cell index or -1 (grid3d.h:55-60,
 * 504-512), hkey = half-cell bin key; perm[j] = original index of the j-th sorted particle. */
int ffb200_get_binning(ffb200_context *ctx, int32_t *cell, uint32_t *hkey, uint32_t *perm);

/* ---- resident grids ----------------------------------------------------------------------------- */

int ffb200_set_velocity_field(ffb200_context *ctx, const float *u, const float *v, const float *w);
int ffb200_set_saved_velocity_field(ffb200_context *ctx, const float *u, const float *v, const float *w);
int ffb200_get_velocity_field(ffb200_context *ctx, float *u, float *v, float *w,
                              uint8_t *validu, uint8_t *validv, uint8_t *validw);
/* Raw weight sums of the last P2G (diagnostics / tests); any pointer may be NULL. */
int ffb200_get_weight_sums(ffb200_context *ctx, float *wu, float *wv, float *ww);
/* _saveVelocityField: device-side deep copy current -> saved. */
int ffb200_save_velocity_field(ffb200_context *ctx);
/* Valid masks of the current field (ValidVelocityComponentGrid, one byte per face, the layout of
 * ffb200_get_velocity_field). ffb200_p2g leaves them on the device; this uploads them from the host. */
int ffb200_set_valid_velocities(ffb200_context *ctx, const uint8_t *validu, const uint8_t *validv, const uint8_t *validw);
/* FluidSimulation::_extrapolateFluidVelocities (fluidsimulation.cpp:6282-6286) ->
 * MACVelocityField::extrapolateVelocityField (macvelocityfield.cpp:671-677) -> GridUtils::extrapolateGrid
 * (gridutils.h:94-163): grows the valid faces of u, v, w by num_layers 6-neighbour layers, in place, bit
 * for bit as the reference does (the reference passes ceil(sqrt(3) * CFL) + 3 = 12). Uses the valid
 * masks of the last ffb200_p2g / ffb200_set_valid_velocities. Whole-grid contexts only. */
int ffb200_extrapolate_velocity_field(ffb200_context *ctx, int num_layers);
int ffb200_set_solid(ffb200_context *ctx, const float *phi, const uint8_t *near_solid);
/* FluidSimulation::_getMaximumMarkerParticleSpeed (fluidsimulation.cpp:10188-10202), the input of the CFL
 * time step (_calculateNextTimeStep, :10229-10262), on the resident velocities: sqrt of the largest float
 * dot product v.v, bit-identical to the reference. Synchronises the stream. On a z-slab context the ghost
 * copies are ignored (reduce the per-rank values with a max). */
int ffb200_get_maximum_particle_speed(ffb200_context *ctx, double *speed);
/* FluidSimulation::_removeMarkerParticles (fluidsimulation.cpp:7773-7851) with
 * _getMarkerParticleSpeedLimit (:7723-7771) on the resident particles: drops the particles inside the solid SDF
 * of ffb200_set_solid (MeshLevelSet::trilinearInterpolateSolidPoints, phi < 0), those beyond an open domain
 * boundary, those beyond the first max_particles_per_cell (_maxMarkerParticlesPerCell = 250) of a cell in
 * particle-index order, and, when extreme_velocity_removal is set (_isExtremeVelocityRemovalEnabled), those
 * faster than the limit derived from the speed histogram over max_frame_time_steps (_maxFrameTimeSteps = 6)
 * bins of width CFL * dx / dt. open_bounds is NULL for a closed domain, else the six planes
 * {x-, x+, y-, y+, z-, z+} of :7780-7788 (boundary AABB -/+ _openBoundaryWidth * dx on open sides,
 * -INFINITY / +INFINITY on closed ones). The removed set is bit-for-bit the reference's. The survivors keep
 * their relative order and are renumbered 0..num_remaining-1 (ParticleSystem::removeParticles);
 * ffb200_get_particles then returns num_remaining rows. num_extreme_removed is
 * _currentExtremeVelocityParticlesRemoved. Synchronises the stream. Whole-grid contexts only;
 * max_particles_per_cell >= 0, dt > 0; with the extreme-velocity rule on, 1 <= max_frame_time_steps <= 4096 (the
 * addon UI allows 100; the setting sizes the speed histogram, fluidsimulation.cpp:7727). Anything else fails with a message. */
int ffb200_remove_marker_particles(ffb200_context *ctx, double dt, double cfl_condition_number, int max_particles_per_cell,
                                   int max_frame_time_steps, int extreme_velocity_removal, const float *open_bounds,
                                   int *num_remaining, int *num_extreme_removed);

/* ParticleLevelSet::calculateSignedDistanceField (particlelevelset.cpp:161-168, 335-668; called from
 * FluidSimulation::_updateLiquidLevelSet, fluidsimulation.cpp:5599) on the resident positions: the cell-centred
 * liquid SDF, phi = min(3 dx, min over particles of |cell centre - p| - particle_radius) restricted to the
 * (particle, 10^3-block) pairs and block-local float frames of the reference, bit for bit (a minimum does not
 * depend on the order). particle_radius = _liquidSDFParticleRadius = 0.5 * dx * sqrt(3) (x2 with the smooth
 * surface-tension kernel), below 5 dx. The field stays on the device; ffb200_get_liquid_sdf copies its
 * I*J*K floats (x fastest) to the host. Whole-grid contexts only. */
int ffb200_liquid_sdf(ffb200_context *ctx, double particle_radius);
int ffb200_get_liquid_sdf(ffb200_context *ctx, float *phi);
/* ParticleLevelSet::postProcessSignedDistanceField (particlelevelset.cpp:170-195, called at fluidsimulation.cpp:5616) on
 * the device field of ffb200_liquid_sdf, against the solid SDF of ffb200_set_solid: cells whose centre lies in the
 * solid go to -dx/2, magnitudes below 0.005 dx are clamped away from zero. (Experimental in round 1: pinned against
 * the reference on the CPU, its kernel has not run on hardware yet.) */
int ffb200_postprocess_liquid_sdf(ffb200_context *ctx);

/* Arithmetic of the gathers (G2P, advection). FFB200_PRECISION_EXACT (default) repeats the reference's fp64
 * index / fraction / blend operation for operation: particle velocities, APIC rows and advected positions are
 * bit-identical to the reference. FFB200_PRECISION_TOLERANCE evaluates the same trilinear interpolant in fp32
 * from float-pair cell coordinates (no fp64, no conversion instructions) and meets the north star's 1e-5
 * relative bar instead; every discrete decision of the advection -- collision gate, clearance shortcut, boundary
 * clamp -- is taken with a guard band, and particles whose decisions are not clear-cut (those that can touch a
 * solid) run the exact code in full. Binning, sort order and valid-face masks do not depend on the mode.
 * ffb200_get_tolerance_stats: counts[0] unused, counts[1] G2P particles with a component sent to the exact
 * code (gradient frame within 1e-6 cells of a plane), counts[2] particles advected in tolerance mode,
 * counts[3] those of them that took the exact path; reset != 0 clears the counters. */
#define FFB200_PRECISION_EXACT 0
#define FFB200_PRECISION_TOLERANCE 1
int ffb200_set_precision(ffb200_context *ctx, int mode);
int ffb200_get_tolerance_stats(ffb200_context *ctx, unsigned long long *counts, int reset);

/* Restricts the NEXT ffb200_g2p / ffb200_advect / ffb200_slab_route_ghosts_begin calls to the particles whose (sorted)
 * cell plane lies inside (mode 1) or outside (mode 2) the global plane range [k_lo, k_hi); mode 0 lifts the
 * restriction. Valid while the particle order is the one of the last sort (ffb200_p2g ... ffb200_advect of one
 * substep). The range is resolved on the device from the bin table, so no size crosses PCIe: the z-slab driver
 * advects the particles near its faces first, starts the neighbour exchange, and runs the interior meanwhile. */
int ffb200_set_particle_window(ffb200_context *ctx, int k_lo, int k_hi, int mode);

/* ffb200_set_solid with DEVICE pointers: d_phi holds the context's stored node planes only
 * ((I+1)(J+1)(kloc+1) floats, first plane = the context's first stored cell plane), d_near_solid the whole
 * ceil(I/3) x ceil(J/3) x ceil(K/3) byte grid. For scenes too large to stage through host arrays per rank. */
int ffb200_set_solid_device(ffb200_context *ctx, const float *d_phi, const uint8_t *d_near_solid);

/* AttributeToGridTransfer<T>::transfer (attributetogridtransfer.h:52-157, 213-520; the age / lifetime / viscosity /
 * density / colour / whitewater-proximity grids of fluidsimulation.cpp:6990-7170): the FLIP-kernel splat of one
 * float (num_components = 1) or one vmath::vec3 (num_components = 3, interleaved, normalised with vec3 /= float)
 * per particle onto the CELL-CENTRED I x J x K grid (gridOffset (dx/2, dx/2, dx/2) at every call site),
 * particle_radius = <attribute radius in voxels> * dx (1 to 3 dx in the reference). grid: I*J*K*num_components
 * floats, valid: I*J*K bytes (weight > 1e-6), both x fastest; normalize = AttributeTransferParameters::normalize
 * (0: the weighted sums are kept, the whitewater-proximity grid of fluidsimulation.cpp:7083). Valid masks bit for
 * bit, values to summation order (1e-5), the velocity transfer's contract. Uploads positions and payload (host
 * order) and REPLACES the resident particle set; whole-grid contexts only. */
int ffb200_attribute_to_grid_transfer(ffb200_context *ctx, int n, const float *pos, const float *attr, int num_components,
                                      double particle_radius, int normalize, float *grid, uint8_t *valid);

/* ---- stages on resident data ---------------------------------------------------------------------- */

int ffb200_p2g(ffb200_context *ctx, double particle_radius, int transfer_method);
int ffb200_g2p(ffb200_context *ctx, int transfer_method, double ratio_pic_flip);
int ffb200_advect(ffb200_context *ctx, double dt, double cfl_condition_number, int resolve_collisions);

/* ---- one-call host-buffer entry points (what the libffengine interposer calls) ---------------------- */

/* VelocityAdvector::advect: upload particles, sort, transfer, download faces + valid masks. Each direction's
 * faces and mask leave the device behind that direction's kernels, under the kernels of the other two (the copies
 * overlap when the host arrays are page-locked: ffb200_pin_host_memory). All outputs are complete on return. */
int ffb200_velocity_advector_advect(ffb200_context *ctx, int n, const float *pos, const float *vel,
                                    const float *affx, const float *affy, const float *affz,
                                    double particle_radius, int transfer_method,
                                    float *u, float *v, float *w,
                                    uint8_t *validu, uint8_t *validv, uint8_t *validw);

/* The reference hands the SAME particle arrays to its P2G, its G2P and its advection within one
 * substep (fluidsimulation.cpp:10078-10121: nothing moves marker particles in between unless sheet
 * seeding is on), and the same MAC field to the G2P and the advection. A caller that knows this may
 * say so: the mask applies to the NEXT host-buffer entry point on this context only and lets it
 * skip the upload of what the device already holds from the previous one (particles stay in their
 * sorted order on the device; outputs are still written to the host arrays in the caller's order).
 *   FFB200_RESIDENT_PARTICLES  positions, velocities (and affine rows) equal those of the previous call
 *   FFB200_RESIDENT_FIELD      u, v, w equal those of the previous call
 *   FFB200_RESIDENT_SAVED_FIELD the device's saved field (ffb200_save_velocity_field) is the caller's _savedVelocityField
 * Honoured by ffb200_velocity_advector_advect, ffb200_update_marker_particle_velocities,
 * ffb200_advance_marker_particles, ffb200_mark_removed_marker_particles and
 * ffb200_calculate_signed_distance_field; the particle count must match or the call fails. The declaration is
 * consumed -- or dropped -- by the very next call on the context, whichever entry point that is.
 *
 * Residency across stages AND substeps (the generation protocol the interposer runs): with
 * FFB200_RESIDENT_PARTICLES the particle pointers of those entry points may be NULL. Inputs are then taken
 * from the device and OUTPUTS STAY THERE (G2P velocities / affine rows, advected positions); the host copy is
 * stale until ffb200_get_particles fetches what host code actually reads. Null field outputs of
 * ffb200_velocity_advector_advect likewise stay resident (ffb200_get_velocity_field). */
#define FFB200_RESIDENT_PARTICLES 1u
#define FFB200_RESIDENT_FIELD 2u
#define FFB200_RESIDENT_SOLID 4u     /* ffb200_mark_removed_marker_particles: phi equals that of the previous call */
#define FFB200_RESIDENT_SAVED_FIELD 8u
int ffb200_declare_resident(ffb200_context *ctx, unsigned mask);

/* _removeMarkerParticles (fluidsimulation.cpp:7773-7851, called at :7892 right after the advection) on host
 * arrays: the decisions of ffb200_remove_marker_particles, returned as one byte per particle in the caller's
 * order (removed[i] != 0: the reference's isRemoved[i]) so that the caller can compact every attribute of its
 * particle system (ParticleSystem::removeParticles). pre_removed (NULL or n bytes) marks particles that are
 * dropped before the per-cell count whatever their position -- the lifetime rule of :7808-7814, which only the
 * caller can evaluate. phi / near_solid as in ffb200_advance_marker_particles; both may be NULL after
 * ffb200_declare_resident(FFB200_RESIDENT_SOLID), pos / vel after FFB200_RESIDENT_PARTICLES (the reference's
 * call order: the advection has just left exactly these positions, the G2P these velocities, on the device). */
int ffb200_mark_removed_marker_particles(ffb200_context *ctx, int n, const float *pos, const float *vel, const float *phi,
                                         const uint8_t *near_solid, const float *open_bounds, const uint8_t *pre_removed,
                                         double dt, double cfl_condition_number, int max_particles_per_cell,
                                         int max_frame_time_steps, int extreme_velocity_removal, uint8_t *removed,
                                         int *num_removed, int *num_extreme_removed);

/* The same decisions for RESIDENT particles and a resident solid SDF (the advection has just left both on the
 * device), applied on both sides: the device set is compacted exactly as ffb200_remove_marker_particles does
 * (survivors keep their order, ids renumbered 0..num_remaining-1) and `removed` (n bytes, the count before the
 * call) tells the host which entries of its own attribute vectors to drop (ParticleSystem::removeParticles),
 * so that host index i and device id i stay the same particle without the arrays crossing PCIe. When nothing
 * is removed -- the usual substep -- `removed` is cleared on the host and no mask is downloaded. */
int ffb200_remove_marker_particles_masked(ffb200_context *ctx, const float *open_bounds, const uint8_t *pre_removed, double dt,
                                          double cfl_condition_number, int max_particles_per_cell, int max_frame_time_steps,
                                          int extreme_velocity_removal, uint8_t *removed, int *num_remaining,
                                          int *num_extreme_removed);

/* Page-lock / release a host range the caller keeps alive (cudaHostRegister): the interposer pins the reference's
 * long-lived MAC field arrays once, so that the per-substep field transfers run at pinned-memory speed out of the
 * reference's own containers. Pinning an already pinned range succeeds. */
int ffb200_pin_host_memory(ffb200_context *ctx, void *ptr, size_t bytes);
int ffb200_unpin_host_memory(ffb200_context *ctx, void *ptr);

/* ParticleLevelSet::calculateSignedDistanceField on host positions -> phi (I*J*K floats, the layout of
 * Array3d<float>::getRawArray()). Uploads the positions unless FFB200_RESIDENT_PARTICLES was declared. */
int ffb200_calculate_signed_distance_field(ffb200_context *ctx, int n, const float *pos, double particle_radius, float *phi);

/* _extrapolateFluidVelocities (fluidsimulation.cpp:6282-6286; the reference passes
 * num_layers = ceil(sqrt(3) * CFL) + 3): u, v, w are extrapolated in place on the host arrays.
 * device_field_is_current != 0 states that u, v, w and the valid masks are exactly what the
 * immediately preceding ffb200_velocity_advector_advect on this context returned (the reference's
 * call order, fluidsimulation.cpp:5652-5654): the upload is skipped and the valid pointers may be NULL. */
int ffb200_extrapolate_fluid_velocities(ffb200_context *ctx, float *u, float *v, float *w,
                                        const uint8_t *validu, const uint8_t *validv, const uint8_t *validw,
                                        int num_layers, int device_field_is_current);

/* _updateMarkerParticleVelocitiesThread: upload particles and both fields, gather, download.
 * vel is updated in place; affx/affy/affz are outputs for APIC (ignored for FLIP);
 * su/sv/sw (the saved field) may be NULL for APIC. Large fields arrive in plane chunks on a copy stream with the
 * gather following range by range over the sorted particles, so the upload hides the kernel (page-locked host
 * arrays); the host field arrays may be reused as soon as the call returns. */
int ffb200_update_marker_particle_velocities(ffb200_context *ctx, int n, const float *pos, float *vel,
                                             float *affx, float *affy, float *affz,
                                             const float *u, const float *v, const float *w,
                                             const float *su, const float *sv, const float *sw,
                                             int transfer_method, double ratio_pic_flip);

/* _advanceMarkerParticles (RK3 + _resolveCollision): pos is updated in place. phi/near_solid
 * may be NULL to reuse the arrays of the previous call (static solids). */
int ffb200_advance_marker_particles(ffb200_context *ctx, int n, float *pos,
                                    const float *u, const float *v, const float *w,
                                    const float *phi, const uint8_t *near_solid,
                                    double dt, double cfl_condition_number);

#ifdef __cplusplus
}
#endif
#endif /* FFB200_H */
