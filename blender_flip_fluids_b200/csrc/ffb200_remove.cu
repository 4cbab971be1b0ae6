// ffb200_remove.cu -- marker-particle removal on the resident SoA streams.
//
//   FluidSimulation::_getMarkerParticleSpeedLimit   fluidsimulation.cpp:7723-7771
//   FluidSimulation::_removeMarkerParticles         fluidsimulation.cpp:7773-7851   (closed boundaries, no lifetimes)
//   MeshLevelSet::trilinearInterpolateSolidPoints   meshlevelset.h:203-219, 333-339 (phi(p) < 0)
//   Interpolation::trilinearInterpolate(vec3,dx,g)  interpolation.cpp:72-112
//
// The reference walks the particles once, in index order, with a per-cell counter: a particle
// inside the solid is dropped; otherwise it is dropped when its cell already holds
// `max_particles_per_cell` earlier survivors of the solid test, and counted if not; a counted
// particle is dropped when its speed exceeds the limit derived from a 6-bin speed histogram.
// Only the per-cell cap depends on the order, and only in cells that hold more candidates than
// the cap. So: histogram + maximum (one pass), outlier counts (one pass), the scalar limit (one
// thread), classification with one integer atomic per candidate, and an order-exact ranking by
// original index restricted to the candidates of over-full cells (normally none: the kernels
// exit on a flag). Survivors are then written in host order (destination = rank of the original
// index among the survivors), so the compaction needs one scan and no sort, and the ids of the
// survivors are again 0..n'-1 exactly like ParticleSystem::removeParticles leaves them.
#include "ffb200_ctx.h"

#include <cstring>

namespace ffb200 {

namespace {

constexpr int kThreads = 256;
// _maxFrameTimeSteps bins of the speed histogram (the reference sizes it from the setting, :7727). The addon's
// 'Max Substeps' goes to 100 in its normal UI; 4096 bins are 16 KB of shared memory. Beyond that the limit is
// only enforced when the extreme-velocity rule -- the histogram's only user -- is enabled.
constexpr int kMaxSteps = 4096;

// device scalars of one removal (uint32 words)
enum Word {
    W_HIST = 0,             // kMaxSteps histogram bins
    W_MAXBITS = kMaxSteps,  // float bits of the largest speed
    W_NLOWER,
    W_NTHR,
    W_LIMIT,                // float bits of the speed limit
    W_OVERFULL,             // some cell holds more candidates than the cap
    W_LIST,                 // entries of the over-full list
    W_SURVIVORS,
    W_EXTREME,
    W_COUNT
};

struct SolidView {
    GridDesc g;
    const float *phi;       // (I+1)(J+1)(kloc+1), first stored node plane = kbase
};

// Interpolation::trilinearInterpolate(vec3, dx, Array3d<float>) on the node-centred solid SDF:
// the arithmetic of sdf_sample in ffb200_advect.cu (out-of-range corners read as 0).
__device__ __forceinline__ float solid_phi(const SolidView &S, float x, float y, float z) {
    const GridDesc &g = S.g;
    const int w = g.I + 1, h = g.J + 1, d = g.K + 1;
    const int i = pos2idx(x, g.inv_dx), j = pos2idx(y, g.inv_dx), k = pos2idx(z, g.inv_dx);
    const double ix = (double)(x - idx2posf(i, g.dx)) * g.inv_dx;
    const double iy = (double)(y - idx2posf(j, g.dx)) * g.inv_dx;
    const double iz = (double)(z - idx2posf(k, g.dx)) * g.inv_dx;
    const bool i0 = (unsigned)i < (unsigned)w, i1 = (unsigned)(i + 1) < (unsigned)w;
    const bool j0 = (unsigned)j < (unsigned)h, j1 = (unsigned)(j + 1) < (unsigned)h;
    const bool k0 = (unsigned)k < (unsigned)d, k1 = (unsigned)(k + 1) < (unsigned)d;
    const long long sj = w, sk = (long long)w * h;
    const long long base = (long long)i + sj * j + sk * (long long)(k - g.kbase);
    const float *f = S.phi;
    double p[8];
    p[0] = (i0 && j0 && k0) ? (double)__ldg(f + base) : 0.0;
    p[1] = (i1 && j0 && k0) ? (double)__ldg(f + base + 1) : 0.0;
    p[2] = (i0 && j1 && k0) ? (double)__ldg(f + base + sj) : 0.0;
    p[3] = (i0 && j0 && k1) ? (double)__ldg(f + base + sk) : 0.0;
    p[4] = (i1 && j0 && k1) ? (double)__ldg(f + base + sk + 1) : 0.0;
    p[5] = (i0 && j1 && k1) ? (double)__ldg(f + base + sk + sj) : 0.0;
    p[6] = (i1 && j1 && k0) ? (double)__ldg(f + base + sj + 1) : 0.0;
    p[7] = (i1 && j1 && k1) ? (double)__ldg(f + base + sk + sj + 1) : 0.0;
    return (float)trilerp8(p, ix, iy, iz);
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// pass 1 of _getMarkerParticleSpeedLimit (:7729-7737): histogram of floor(speed / step), largest speed
__global__ void __launch_bounds__(kThreads) k_speed_hist(const float *__restrict__ vx, const float *__restrict__ vy,
                                                         const float *__restrict__ vz, int n, double step, int steps,
                                                         uint32_t *__restrict__ words) {
    __shared__ uint32_t sh[kMaxSteps + 1];
    for (int q = threadIdx.x; q <= kMaxSteps; q += blockDim.x) sh[q] = 0u;
    __syncthreads();
    // warp-uniform trip count; the lanes of a warp that fall in the same bin share one shared-memory atomic
    const int lane = threadIdx.x & 31;
    uint32_t top = 0u;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j - lane < n; j += gridDim.x * blockDim.x) {
        const bool ok = j < n;
        const unsigned live = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const float s = vlen3(vx[j], vy[j], vz[j]);
            int b = (int)fmin(floor((double)s / step), (double)(steps - 1));
            b = b < 0 ? 0 : b;
            const unsigned peers = __match_any_sync(live, b);
            if (lane == __ffs(peers) - 1) atomicAdd(&sh[b], (uint32_t)__popc(peers));
            top = max(top, __float_as_uint(s));                 // speeds are >= 0: the bit pattern orders them
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
    if (lane == 0 && top) atomicMax(&sh[kMaxSteps], top);
    __syncthreads();
    for (int q = threadIdx.x; q < steps; q += blockDim.x)
        if (sh[q]) atomicAdd(&words[W_HIST + q], sh[q]);
    if (threadIdx.x == 0 && sh[kMaxSteps]) atomicMax(&words[W_MAXBITS], sh[kMaxSteps]);
}

// pass 2 (:7753-7762): particles in [0.90, 0.99999) and [0.99999, 1] of the largest speed
__global__ void __launch_bounds__(kThreads) k_speed_outliers(const float *__restrict__ vx, const float *__restrict__ vy,
                                                             const float *__restrict__ vz, int n,
                                                             uint32_t *__restrict__ words) {
    const double top = (double)__uint_as_float(words[W_MAXBITS]);
    const double lower = 0.90 * top, thr = 0.99999 * top;
    uint32_t nl = 0, nt = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const double s = (double)vlen3(vx[j], vy[j], vz[j]);
        nl += (s >= lower && s < thr) ? 1u : 0u;
        nt += (s >= thr) ? 1u : 0u;
    }
    nl = warp_sum(nl);
    nt = warp_sum(nt);
    if ((threadIdx.x & 31) == 0) {
        if (nl) atomicAdd(&words[W_NLOWER], nl);
        if (nt) atomicAdd(&words[W_NTHR], nt);
    }
}

// the scalar tail of _getMarkerParticleSpeedLimit (:7739-7770)
__global__ void k_speed_limit(int n, double step, int steps, uint32_t *__restrict__ words) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double maxpct = 0.0005;                               // _maxExtremeVelocityRemovalPercent
    const int maxabs = 35;                                      // _maxExtremeVelocityRemovalAbsolute
    const int max_removal = (int)fmin((double)(int)((double)n * maxpct), (double)maxabs);
    double maxspeed = steps * step;
    int current = 0;
    for (int i = steps - 1; i > 0; i--) {
        const int cnt = (int)words[W_HIST + i];
        if (current + cnt > max_removal) break;
        current += cnt;
        const int s = i + 4 > steps ? i + 4 : steps;            // _minTimeStepIncreaseForRemoval
        maxspeed = s * step;
    }
    const double top = (double)__uint_as_float(words[W_MAXBITS]);
    const double thr = 0.99999 * top;
    if (words[W_NTHR] <= 6u && words[W_NLOWER] <= 6u)           // _maxExtremeVelocityOutlierRemovalAbsolute
        maxspeed = thr < maxspeed ? thr : maxspeed;
    words[W_LIMIT] = __float_as_uint((float)maxspeed);
}

// state per slot: 0 counted candidate, 1 dropped before counting (inside the solid, pre-removed, beyond an open
// boundary, outside the grid), 2 over the cell cap, 3 extreme speed
struct OpenBounds {
    float xn, xp, yn, yp, zn, zp;       // -inf / +inf on closed sides (:7780-7788)
};

__global__ void __launch_bounds__(kThreads) k_remove_classify(SolidView S, OpenBounds B, const float *__restrict__ px,
                                                              const float *__restrict__ py, const float *__restrict__ pz,
                                                              const uint32_t *__restrict__ orig,
                                                              const uint8_t *__restrict__ pre_removed,
                                                              int n, int cap, uint32_t *__restrict__ state,
                                                              int *__restrict__ cell_of, uint32_t *__restrict__ count,
                                                              uint32_t *__restrict__ words) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float x = px[j], y = py[j], z = pz[j];
    const GridDesc &g = S.g;
    uint32_t st = solid_phi(S, x, y, z) < 0.0f ? 1u : 0u;
    if (pre_removed && pre_removed[orig[j]] != 0) st = 1u;             // e.g. the lifetime rule (:7808-7814), decided by the caller
    if (x < B.xn || x > B.xp || y < B.yn || y > B.yp || z < B.zn || z > B.zp) st = 1u;   // open boundaries (:7817-7823)
    int cell = -1;
    if (st == 0u) {
        const int ci = pos2idx(x, g.inv_dx), cj = pos2idx(y, g.inv_dx), ck = pos2idx(z, g.inv_dx);
        if (in_range3(ci, cj, ck, g.I, g.J, g.K)) {
            cell = ci + g.I * (cj + g.J * ck);
            if (atomicAdd(&count[cell], 1u) >= (uint32_t)cap) words[W_OVERFULL] = 1u;
        } else {
            st = 1u;                                            // the reference would index out of the count grid
        }
    }
    state[j] = st;
    cell_of[j] = cell;
}

// candidates of over-full cells, in arbitrary order (normally there are none)
__global__ void __launch_bounds__(kThreads) k_remove_overfull_list(int n, int cap, const uint32_t *__restrict__ state,
                                                                   const int *__restrict__ cell_of,
                                                                   const uint32_t *__restrict__ count,
                                                                   uint32_t *__restrict__ list, uint32_t *__restrict__ words) {
    if (words[W_OVERFULL] == 0u) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || state[j] != 0u) return;
    if (count[cell_of[j]] > (uint32_t)cap) list[atomicAdd(&words[W_LIST], 1u)] = (uint32_t)j;
}

// rank of every listed candidate among the candidates of its cell by original index; the first
// `cap` of a cell are the ones the reference's counter lets through (:7826-7830)
__global__ void __launch_bounds__(kThreads) k_remove_overfull_rank(int cap, const uint32_t *__restrict__ list,
                                                                   const int *__restrict__ cell_of,
                                                                   const uint32_t *__restrict__ orig,
                                                                   uint32_t *__restrict__ state,
                                                                   const uint32_t *__restrict__ words) {
    __shared__ int sh_cell[kThreads];
    __shared__ uint32_t sh_orig[kThreads];
    const uint32_t L = words[W_LIST];
    if (words[W_OVERFULL] == 0u || (uint32_t)blockIdx.x * blockDim.x >= L) return;     // uniform per block
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = e < L;
    const uint32_t slot = live ? list[e] : 0u;
    const int my_cell = live ? cell_of[slot] : -2;
    const uint32_t my_orig = live ? orig[slot] : 0u;
    uint32_t rank = 0;
    for (uint32_t base = 0; base < L; base += blockDim.x) {
        const uint32_t f = base + threadIdx.x;
        if (f < L) {
            const uint32_t s = list[f];
            sh_cell[threadIdx.x] = cell_of[s];
            sh_orig[threadIdx.x] = orig[s];
        } else {
            sh_cell[threadIdx.x] = -1;
            sh_orig[threadIdx.x] = 0u;
        }
        __syncthreads();
        for (int q = 0; q < kThreads; q++) rank += (sh_cell[q] == my_cell && sh_orig[q] < my_orig) ? 1u : 0u;
        __syncthreads();
    }
    if (live && rank >= (uint32_t)cap) state[slot] = 2u;
}

// extreme-speed test of the counted candidates (:7832-7838) and the keep flag of every original index
__global__ void __launch_bounds__(kThreads) k_remove_final(const float *__restrict__ vx, const float *__restrict__ vy,
                                                           const float *__restrict__ vz, const uint32_t *__restrict__ orig,
                                                           int n, int extreme_on, uint32_t *__restrict__ state,
                                                           uint32_t *__restrict__ keep_by_orig,
                                                           uint8_t *__restrict__ removed_by_orig, uint32_t *__restrict__ words) {
    __shared__ uint32_t sh_count[2];
    if (threadIdx.x < 2) sh_count[threadIdx.x] = 0u;
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t keep = 0u, extreme = 0u;
    if (j < n) {
        uint32_t st = state[j];
        if (st == 0u && extreme_on) {
            const float limit = __uint_as_float(words[W_LIMIT]);
            const double limit_sq = (double)(limit * limit);    // `double maxspeedsq = maxspeed * maxspeed;` with a float maxspeed
            const float x = vx[j], y = vy[j], z = vz[j];
            const float d = x * x + y * y + z * z;              // vmath::dot
            if ((double)d > limit_sq) {
                st = 3u;
                extreme = 1u;
                state[j] = st;
            }
        }
        keep = st == 0u ? 1u : 0u;
        const uint32_t o = orig[j];
        keep_by_orig[o] = keep;
        if (removed_by_orig) removed_by_orig[o] = (uint8_t)(keep ^ 1u);
    }
    // one global atomic per CTA: every warp has survivors, and same-address atomics serialise in L2
    const uint32_t nk = warp_sum(keep), nx = warp_sum(extreme);
    if ((threadIdx.x & 31) == 0) {
        if (nk) atomicAdd(&sh_count[0], nk);
        if (nx) atomicAdd(&sh_count[1], nx);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sh_count[0]) atomicAdd(&words[W_SURVIVORS], sh_count[0]);
        if (sh_count[1]) atomicAdd(&words[W_EXTREME], sh_count[1]);
    }
}

struct CompactArgs {
    const float *src[15];
    float *dst[15];
    int nstreams;
};

// survivors to slot = rank of their original index among the survivors (host order), ids renumbered
__global__ void __launch_bounds__(kThreads) k_remove_compact(const __grid_constant__ CompactArgs a, int n,
                                                             const uint32_t *__restrict__ state,
                                                             const uint32_t *__restrict__ orig,
                                                             const uint32_t *__restrict__ new_index,
                                                             uint32_t *__restrict__ orig_new) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || state[j] != 0u) return;
    const uint32_t dst = new_index[orig[j]];
    float v[15];
#pragma unroll
    for (int t = 0; t < 15; t++)
        if (t < a.nstreams) v[t] = __ldg(a.src[t] + j);
#pragma unroll
    for (int t = 0; t < 15; t++)
        if (t < a.nstreams) a.dst[t][dst] = v[t];
    orig_new[dst] = dst;
}

inline int blocks_for(int n) { return (n + kThreads - 1) / kThreads; }

// everything up to the keep flags; the counts stay in c.remove_words
int remove_mark(Context &c, const RemoveRules &r, const uint8_t *pre_removed, uint8_t *removed_by_orig) {
    if (!c.has_solid) throw CudaError("ffb200_remove_marker_particles: needs the solid SDF (ffb200_set_solid) first");
    if (c.g.kbase != 0 || c.g.kloc != c.g.K)
        throw CudaError("ffb200_remove_marker_particles: not available on z-slab contexts");
    if (r.extreme_on && (r.max_frame_steps < 1 || r.max_frame_steps > kMaxSteps))
        throw CudaError("ffb200_remove_marker_particles: max_frame_time_steps must be in [1, 4096] for the extreme-velocity rule");
    if (r.max_per_cell < 0) throw CudaError("ffb200_remove_marker_particles: negative per-cell cap");
    if (!(r.dt > 0.0)) throw CudaError("ffb200_remove_marker_particles: dt must be positive");
    const int n = c.n;
    if (!c.remove_words) FFB_CUDA(cudaMalloc(&c.remove_words, W_COUNT * sizeof(uint32_t)));
    uint32_t *words = c.remove_words;
    FFB_CUDA(cudaMemsetAsync(words, 0, W_COUNT * sizeof(uint32_t), c.stream));
    if (n == 0) return 0;
    const GridDesc &g = c.g;
    ParticleSoA &src = c.soa[c.cur];
    SortScratch &s = c.sort;
    // scratch of the (now stale) sort: the bin table doubles as the per-cell counter grid
    uint32_t *state = s.key[0], *keep_by_orig = s.key[1], *list = s.val[1], *count = s.bin_start;
    int *cell_of = reinterpret_cast<int *>(s.val[0]);
    const size_t cells = (size_t)g.I * g.J * g.K;
    FFB_CUDA(cudaMemsetAsync(count, 0, cells * sizeof(uint32_t), c.stream));
    c.sorted = false;

    const double step = r.cfl * g.dx / r.dt;                    // speedLimitStep
    const int blocks = blocks_for(n);
    const int persistent = blocks < 8 * c.sm_count ? blocks : 8 * c.sm_count;
    int launches = 0;
    if (r.extreme_on) {
        k_speed_hist<<<persistent, kThreads, 0, c.stream>>>(src.v[0], src.v[1], src.v[2], n, step, r.max_frame_steps, words);
        k_speed_outliers<<<persistent, kThreads, 0, c.stream>>>(src.v[0], src.v[1], src.v[2], n, words);
        k_speed_limit<<<1, 32, 0, c.stream>>>(n, step, r.max_frame_steps, words);
        launches += 3;
    }
    SolidView S{g, c.phi};
    OpenBounds B{r.bounds[0], r.bounds[1], r.bounds[2], r.bounds[3], r.bounds[4], r.bounds[5]};
    k_remove_classify<<<blocks, kThreads, 0, c.stream>>>(S, B, src.p[0], src.p[1], src.p[2], src.orig, pre_removed, n, r.max_per_cell,
                                                          state, cell_of, count, words);
    k_remove_overfull_list<<<blocks, kThreads, 0, c.stream>>>(n, r.max_per_cell, state, cell_of, count, list, words);
    k_remove_overfull_rank<<<blocks, kThreads, 0, c.stream>>>(r.max_per_cell, list, cell_of, src.orig, state, words);
    k_remove_final<<<blocks, kThreads, 0, c.stream>>>(src.v[0], src.v[1], src.v[2], src.orig, n, r.extreme_on, state, keep_by_orig,
                                                       removed_by_orig, words);
    launches += 4;
    FFB_CUDA(cudaGetLastError());
    return launches;
}

void read_counts(Context &c, int *remaining, int *extreme_removed) {
    static_assert(W_EXTREME == W_SURVIVORS + 1, "the two counts are read with one copy");
    uint32_t host[2];
    FFB_CUDA(cudaMemcpyAsync(host, c.remove_words + W_SURVIVORS, sizeof(host), cudaMemcpyDeviceToHost, c.stream));
    FFB_CUDA(cudaStreamSynchronize(c.stream));
    *remaining = (int)host[0];
    *extreme_removed = (int)host[1];
}

}  // namespace

int launch_remove_mask(Context &c, const RemoveRules &r, const uint8_t *pre_removed, uint8_t *removed_by_orig, int *remaining,
                       int *extreme_removed) {
    const int launches = remove_mark(c, r, pre_removed, removed_by_orig);
    read_counts(c, remaining, extreme_removed);
    return launches;
}

int launch_remove_particles(Context &c, const RemoveRules &r, int *remaining, int *extreme_removed, const uint8_t *pre_removed,
                            uint8_t *removed_by_orig) {
    int launches = remove_mark(c, r, pre_removed, removed_by_orig);
    const int n = c.n;
    read_counts(c, remaining, extreme_removed);
    if (*remaining != n) {                                   // the usual substep removes nothing: no compaction then
        ParticleSoA &src = c.soa[c.cur], &dst = c.soa[c.cur ^ 1];
        uint32_t *state = c.sort.key[0], *keep_by_orig = c.sort.key[1];
        launches += launch_exclusive_scan(c, keep_by_orig, (size_t)n);
        CompactArgs a;
        int t = 0;
        for (int q = 0; q < 3; q++) { a.src[t] = src.p[q]; a.dst[t] = dst.p[q]; t++; }
        for (int q = 0; q < 3; q++) { a.src[t] = src.v[q]; a.dst[t] = dst.v[q]; t++; }
        if (c.has_affine)
            for (int q = 0; q < 9; q++) { a.src[t] = src.a[q]; a.dst[t] = dst.a[q]; t++; }
        a.nstreams = t;
        for (; t < 15; t++) { a.src[t] = nullptr; a.dst[t] = nullptr; }
        k_remove_compact<<<blocks_for(n), kThreads, 0, c.stream>>>(a, n, state, src.orig, keep_by_orig, dst.orig);
        launches++;
        FFB_CUDA(cudaGetLastError());
        c.cur ^= 1;
        c.n = *remaining;
    }
    return launches;
}

}  // namespace ffb200
