// ffb200_ctx.h -- host-side context and the launch functions each .cu file provides.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "../../include/ffb200.h"
#include "ffb200_common.cuh"

namespace ffb200 {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define FFB_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            throw ffb200::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +  \
                                    " (" __FILE__ ":" + std::to_string(__LINE__) + ")");          \
    } while (0)

// Particle attributes on the device, struct-of-arrays, one float stream per component.
struct ParticleSoA {
    float *p[3] = {nullptr, nullptr, nullptr};       // position
    float *v[3] = {nullptr, nullptr, nullptr};       // velocity
    float *a[9] = {};                                // AFFINEX.xyz, AFFINEY.xyz, AFFINEZ.xyz
    uint32_t *orig = nullptr;                        // original (host) index of the particle in this slot
};

// Per direction (U, V, W): face grid + the reference's 10^3-node block grid.
struct FaceGrid {
    int gi = 0, gj = 0, gk = 0;   // global face dims
    int bi = 0, bj = 0, bk = 0;   // block dims, blockarray3d.h:66-70
    int kstore = 0;               // stored face planes (kloc, or kloc+1 for w)
    size_t count = 0;             // gi*gj*kstore
    float *vel = nullptr;         // current field component
    float *saved = nullptr;       // _savedVelocityField component
    float *wsum = nullptr;        // weight sums of the last P2G
    uint8_t *valid = nullptr;
    uint8_t *home = nullptr, *active = nullptr;   // block masks (bi*bj*bk)
    uint8_t *status[2] = {nullptr, nullptr};      // extrapolation: [0] layer stamp per face, [1] tile flags (lazy)
};

struct SortScratch {
    uint32_t *key[2] = {nullptr, nullptr};
    uint32_t *val[2] = {nullptr, nullptr};
    uint32_t *tile_hist = nullptr;       // digits x tiles, digit-major
    size_t tile_hist_cap = 0;
    uint32_t *bin_start = nullptr;       // nbins + 2 entries
    uint32_t *scan_partials = nullptr;   // block sums for the scans
    size_t scan_partials_cap = 0;
    uint32_t *seam = nullptr;            // 3 words per slot [dir*cap + slot]: 10^3-block membership
    uint32_t *seam_cell = nullptr;       // membership words of the cell-centred attribute grid (lazy)
    int seam_cell_cap = 0;
    // cell-partial splat scratch, one set per MAC direction so that the three transfers can run
    // concurrently (direction 0 on the context stream, 1 and 2 on auxiliary streams)
    struct CellScratch {
        void *partial = nullptr;         // 8 float2 per shifted cell
        uint8_t *cell_flag = nullptr;
        void *cell_list = nullptr;       // occupied shifted cells (plain from the front, seam from the back)
        uint32_t *list_count = nullptr;
        void *ovf = nullptr;             // edge-particle overflow list + its counter
        int *ovf_count = nullptr;
        size_t cells = 0;
        cudaStream_t stream = nullptr;   // auxiliary stream (directions 1, 2)
        cudaEvent_t done = nullptr;
    } cell[3];
    cudaEvent_t fork = nullptr;
    uint32_t *edge_list = nullptr;       // 3 x edge_cap sorted slots of near-plane ("edge") particles
    uint32_t *edge_count = nullptr;      // 4 counters (one per direction)
    uint32_t edge_cap = 0;
};

struct Context {
    GridDesc g;
    int device = 0;
    int sm_count = 148;                  // multiprocessors of `device` (grid sizing of the persistent kernels)
    int k_own_begin = 0, k_own_end = 0, halo = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    int n = 0, cap = 0;
    bool has_affine = false;
    bool sorted = false;                 // soa[cur] is in (hkey, orig) order and bin_start is valid
    int cur = 0;
    ParticleSoA soa[2];
    float *aos_stage = nullptr;          // device staging for AoS <-> SoA transposes (cap*3 floats)
    void *h_stage = nullptr;             // pinned host staging
    size_t h_stage_bytes = 0;
    SortScratch sort;
    FaceGrid face[3];
    FaceGrid cell;                       // the cell-centred grid of the attribute transfer (block masks + weight sums; lazy)

    float *phi = nullptr;                // (I+1)(J+1)(kloc+1) node-centred solid SDF
    uint8_t *near_solid = nullptr;       // ni*nj*nk
    uint8_t *solid_clear[2] = {nullptr, nullptr};   // per stored cell: distance to the nearest cell with a non-positive SDF node
    int ni = 0, nj = 0, nk = 0;
    bool has_solid = false;

    float guard_abs = -1.f, guard_per = -1.f;
    int *slab_counters = nullptr;        // 8 device ints for the slab pack/route kernels
    uint32_t *remove_words = nullptr;    // device scalars of ffb200_remove_marker_particles (lazy)
    int *liquid_phi = nullptr;           // I*J*K cell-centred liquid SDF (ffb200_liquid_sdf; floats once decoded, lazy)
    uint8_t *liquid_blocks = nullptr;    // home + active masks of its 10^3 blocks
    // API-call epoch: bumped by every call that may change particles or fields. An APIC ffb200_g2p
    // records it; an ffb200_advect that finds it unchanged reuses the G2P samples as RK3 stage 1.
    unsigned long long epoch = 0, k1_epoch = ~0ull;
    int k1_buf = 0;                      // 0/1: soa[k1_buf].v (APIC), 2: k1s (FLIP vPIC)
    float *k1s[3] = {nullptr, nullptr, nullptr};
    int k1s_cap = 0;
    unsigned resident_next = 0;          // ffb200_declare_resident: inputs the next host-buffer call may skip uploading
    unsigned resident_arg = 0;           // ... as latched for the call in progress (cleared for every other call)
    Window window;                            // ffb200_set_particle_window: the particles the next G2P / advection / routing touch
    int precision = FFB200_PRECISION_EXACT;   // ffb200_set_precision: exact (reference arithmetic) or tolerance (fp32 gathers, 1e-5)
    unsigned long long tol_advected = 0;      // particles advected in tolerance mode since the last reset (host-side count)
    unsigned long long *tol_stats = nullptr;  // device: {particles advected on the fp32 path, particles sent to the exact code, ...}
    bool nondestructive = false;         // G2P/advect write to the spare SoA buffer (fixed-batch benchmarking)
    ffb200_timing timing = {};
};

// float-pair 1/dx and float grid bounds of the tolerance path (ffb200_common.cuh)
FastGrid make_fast_grid(const GridDesc &g);
unsigned long long *tolerance_stats(Context &c);         // 4 device counters, allocated on first use

// ---- launchers (each returns the number of kernels it launched) -------------------------------

// ffb200_sort.cu
int launch_unpack_aos(Context &c, const float *aos, float *const dst[3], int n);
int launch_pack_aos(Context &c, const float *const src[3], const uint32_t *orig, float *aos, int n);
int launch_iota(Context &c, uint32_t *dst, int n);
struct SeamParams;                        // ffb200_seam.cuh
// keys -> counting sort -> bin table -> reorder (flips c.cur). With `seam`, the reorder also writes the
// P2G membership words / home marks / edge list of every particle (what k_seam_home would do next).
int launch_sort(Context &c, const SeamParams *seam = nullptr);
int launch_binning_dump(Context &c, int32_t *cell, uint32_t *hkey, uint32_t *perm);   // device outputs
int launch_exclusive_scan(Context &c, uint32_t *data, size_t n);                      // in place, on c.stream

// ffb200_extrapolate.cu
int launch_extrapolate(Context &c, int layers);          // GridUtils::extrapolateGrid on u, v, w in place

// ffb200_p2g.cu
void p2g_seam_begin(Context &c, double radius, SeamParams &sp);   // clears marks / counters, fills sp
int launch_p2g_prepare(Context &c, double radius, bool seam_done);   // [membership words +] block masks
// host destinations of a transfer (reference layouts, whole-grid extents): each direction's faces and valid bytes are
// copied out on the stream that produced them, so the copies of one direction overlap the kernels of the others
struct HostFieldOut {
    float *vel[3] = {nullptr, nullptr, nullptr};
    uint8_t *valid[3] = {nullptr, nullptr, nullptr};
};
int launch_p2g(Context &c, double radius, int method, const HostFieldOut *host = nullptr);  // the three transfer kernels
int launch_attribute_p2g(Context &c, double radius, int ncomp, int normalize, float *d_out, uint8_t *d_valid);   // AttributeToGridTransfer<T>

// ffb200_g2p.cu
int launch_g2p(Context &c, int method, double ratio, int first = 0, int count = -1);   // [first, first + count) of the sorted particles

// ffb200_advect.cu
int launch_solid_clearance(Context &c);                  // after every change of the solid SDF
int launch_advect(Context &c, double dt, double cfl, int collide);
int launch_max_speed_sq(Context &c, uint32_t *out_bits);  // bits of max float v.v over the owned particles (device word)

// ffb200_remove.cu
struct RemoveRules {
    double dt = 0, cfl = 5.0;
    int max_per_cell = 250, max_frame_steps = 6, extreme_on = 1;
    float bounds[6];                     // open-boundary planes {x-, x+, y-, y+, z-, z+}; -inf / +inf where closed
};
// _removeMarkerParticles on the resident particles; updates c.n, leaves the survivors in host order
int launch_remove_particles(Context &c, const RemoveRules &r, int *remaining, int *extreme_removed,
                            const uint8_t *pre_removed = nullptr, uint8_t *removed_by_orig = nullptr);
// the same decisions without the compaction: one byte per ORIGINAL index (device pointers; pre_removed may be null)
int launch_remove_mask(Context &c, const RemoveRules &r, const uint8_t *pre_removed, uint8_t *removed_by_orig, int *remaining,
                       int *extreme_removed);

// ffb200_liquid_sdf.cu
int launch_liquid_sdf(Context &c, double radius);        // ParticleLevelSet::calculateSignedDistanceField -> c.liquid_phi
int launch_liquid_sdf_postprocess(Context &c);           // ParticleLevelSet::postProcessSignedDistanceField, in place

// ffb200_slab.cu
int slab_rows(Context &c);
int launch_pack_layers(Context &c, int lo_a, int hi_a, float *block_a, int lo_b, int hi_b, float *block_b, int cap);
// caps (ghost_layers > 0 only): {up migrants, up ghosts, down migrants, down ghosts} section capacities
int launch_route_begin(Context &c, int k_begin, int k_end, float *block_up, float *block_down, int cap, int ghost_layers = 0,
                       const int *caps = nullptr);
int launch_route_end(Context &c, int counts_host[3], int known_holes = -1);
int launch_append(Context &c, const float *block, int count, bool as_ghost);

}  // namespace ffb200
