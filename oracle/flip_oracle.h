/* oracle/flip_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C, single-threaded CPU restatement of the FLIP Fluids engine's particle<->grid hot
 * path (rlguy/Blender-FLIP-Fluids src/engine v1.8.5). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load it; the product (libffb200.so) never does.
 *
 * Pinning: the reference has no tests or golden vectors of its own, so
 * this restatement is pinned against the UNMODIFIED reference built from its own sources
 * (oracle/_ref/libffengine_ref.so through oracle/ref_harness.cpp): tests/golden/ holds the
 * reference-generated fixtures (tests/golden/make_golden.py, index in tests/golden/README.md)
 * that travel to the GPU box, and tests/test_oracle_golden.py checks every function below
 * against them bit for bit.
 *
 * Layouts are the reference's host layouts: particle attributes are packed float[3] per
 * particle (vmath::vec3, vmath.h:37-61); grids are dense x-fastest, flat = i + w*(j + h*k)
 * (array3d.h:774-777): u is (I+1)*J*K, v is I*(J+1)*K, w is I*J*(K+1) (macvelocityfield.cpp:
 * 46-54), phi is (I+1)*(J+1)*(K+1) node-centred (meshlevelset.cpp:35-42); masks are 1 byte.
 */
#ifndef FLIP_ORACLE_H
#define FLIP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLIP_ORACLE_FLIP 0
#define FLIP_ORACLE_APIC 1

/* Cell binning: Grid3d::positionToGridIndex (grid3d.h:55-60) + getFlatIndex (grid3d.h:504-512).
 * cell[n] = i + I*(j + J*k), or -1 if the index is outside [0,I)x[0,J)x[0,K).
 * hkey[n] = half-cell bin key (hi+A) + HX*((hj+A) + HY*(hk+A)) with h = floor(p * (2*(1.0/dx)))
 * (so that h>>1 is exactly the reference cell index), A = 4 apron half-cells per side,
 * HX = 2I+2A etc.; HX*HY*HZ for particles outside the apron.
 * perm = stable ascending sort of hkey (ties keep ascending particle index, which is the
 * order the reference accumulates a block's particles in, velocityadvector.cpp:383-413). */
void flip_oracle_bin_sort(int I, int J, int K, double dx, int n, const float *pos,
                          int32_t *cell, uint32_t *hkey, uint32_t *perm);

/* P2G: VelocityAdvector::advect (velocityadvector.cpp:38-623) onto cleared grids
 * (fluidsimulation.cpp:5630-5631). affx/affy/affz may be NULL for FLIP. */
void flip_oracle_p2g(int I, int J, int K, double dx, double radius, int method, int n,
                     const float *pos, const float *vel,
                     const float *affx, const float *affy, const float *affz,
                     float *u, float *v, float *w,
                     uint8_t *validu, uint8_t *validv, uint8_t *validw);

/* Same, additionally returning the raw weight sums (for guard-band tests). Any of the
 * wsum pointers may be NULL. */
void flip_oracle_p2g_w(int I, int J, int K, double dx, double radius, int method, int n,
                       const float *pos, const float *vel,
                       const float *affx, const float *affy, const float *affz,
                       float *u, float *v, float *w,
                       uint8_t *validu, uint8_t *validv, uint8_t *validw,
                       float *wsumu, float *wsumv, float *wsumw);

/* G2P FLIP: _updatePICFLIPMarkerParticleVelocitiesThread (fluidsimulation.cpp:6771-6784).
 * vel is updated in place. */
void flip_oracle_g2p_flip(int I, int J, int K, double dx, int n, const float *pos, float *vel,
                          const float *u, const float *v, const float *w,
                          const float *su, const float *sv, const float *sw, double ratio_pic_flip);

/* G2P APIC: _updatePICAPICMarkerParticleVelocitiesThread (fluidsimulation.cpp:6791-6843). */
void flip_oracle_g2p_apic(int I, int J, int K, double dx, int n, const float *pos, float *vel,
                          float *affx, float *affy, float *affz,
                          const float *u, const float *v, const float *w);

/* MAC gather: MACVelocityField::evaluateVelocityAtPositionLinear (macvelocityfield.cpp:631-645). */
void flip_oracle_mac_sample(int I, int J, int K, double dx, int n, const float *pos, float *out,
                            const float *u, const float *v, const float *w);

/* Advect: _RK3 (fluidsimulation.cpp:7616-7623) + _resolveCollision (7646-7721) with the
 * boundary of _getBoundaryAABB (5175-5180) shrunk by _solidBufferWidth*dx (7639-7640).
 * near is the 3dx-granular near-solid mask, dims ceil(I*dx/(3dx)) etc. (5448-5451).
 * collide=0 runs RK3 only. */
void flip_oracle_advect(int I, int J, int K, double dx, int n, const float *pos_in, float *pos_out,
                        const float *u, const float *v, const float *w,
                        const float *phi, const uint8_t *near, double dt, double cfl, int collide);

/* near-solid grid dimensions used by flip_oracle_advect. */
void flip_oracle_near_dims(int I, int J, int K, double dx, int *gi, int *gj, int *gk);

/* GridUtils::extrapolateGrid (gridutils.h:94-163) on one w x h x d float grid (x fastest), in place. */
void flip_oracle_extrapolate(int w, int h, int d, float *grid, const uint8_t *valid, int layers);

/* FluidSimulation::_getMaximumMarkerParticleSpeed (fluidsimulation.cpp:10188-10202). */
double flip_oracle_max_particle_speed(int n, const float *vel);

/* FluidSimulation::_removeMarkerParticles (fluidsimulation.cpp:7773-7851). open_bounds: NULL (closed domain) or the six
 * planes {x-, x+, y-, y+, z-, z+} of :7780-7788; pre_removed: NULL or the caller-evaluated lifetime rule (:7808-7814). */
void flip_oracle_remove_particles(int I, int J, int K, double dx, int n, const float *pos, const float *vel,
                                  const float *phi, double dt, double cfl, int max_per_cell, int extreme_removal,
                                  const float *open_bounds, const uint8_t *pre_removed, uint8_t *removed, int *num_extreme);

/* ParticleLevelSet::calculateSignedDistanceField (particlelevelset.cpp:161-168, 335-668): cell-centred liquid SDF
 * phi[K][J][I] from the marker positions; radius = _liquidSDFParticleRadius (fluidsimulation.cpp:4351). */
void flip_oracle_liquid_sdf(int I, int J, int K, double dx, double radius, int n, const float *pos, float *phi);

/* ParticleLevelSet::postProcessSignedDistanceField (particlelevelset.cpp:170-195), in place. */
void flip_oracle_liquid_sdf_postprocess(int I, int J, int K, double dx, float *phi, const float *solid);

/* AttributeToGridTransfer<float>::transfer (attributetogridtransfer.h:52-157): scalar attribute -> cell-centred grid
 * grid[K][J][I], normalised, valid = weight > 1e-6. */
void flip_oracle_attribute_p2g(int I, int J, int K, double dx, double radius, int n, const float *pos, const float *attr,
                               float *grid, uint8_t *valid);

/* AttributeToGridTransfer<vmath::vec3>::transfer: packed float[3] attributes -> grid[K][J][I][3], one valid mask. */
void flip_oracle_attribute_p2g_vec3(int I, int J, int K, double dx, double radius, int n, const float *pos, const float *attr,
                                    float *grid, uint8_t *valid);

/* flip_oracle_liquid_sdf evaluated through per-axis lists with a squared-distance pre-filter (the device's
 * k_sdf_scatter_axes): must be bit-identical; returns the number of candidates the filter skipped. */
long long flip_oracle_liquid_sdf_axes(int I, int J, int K, double dx, double radius, int n, const float *pos, float *phi);

#ifdef __cplusplus
}
#endif
#endif
