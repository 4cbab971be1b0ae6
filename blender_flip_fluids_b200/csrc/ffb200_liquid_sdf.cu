// ffb200_liquid_sdf.cu -- liquid signed distance field from the marker particles.
//
//   ParticleLevelSet::calculateSignedDistanceField          particlelevelset.cpp:161-168 (fluidsimulation.cpp:5599)
//   ParticleLevelSet::_computeSignedDistanceFromParticles   particlelevelset.cpp:335-398
//   _initializeBlockGrid / _initializeActiveBlocksThread    :400-450  (10^3 blocks holding a particle, 26-feathered)
//   _computeGridCountDataThread                             :515-569  (which blocks a particle is handed to)
//   _computeExactBandProducerThread                         :620-668  (block-local distances, running minimum)
//
// phi(cell) = min(3dx, min over the (particle, block) pairs the reference forms of |centre - p_local| - r), with the
// particle expressed in the block's own frame in float and the cell range clamped to the block. A minimum does
// not depend on the order, so the reference's sort into blocks and its producer/consumer threads decide
// nothing: one thread per particle repeats the block test and the block-local arithmetic and lowers the cells
// with an integer atomicMin on an order-preserving encoding of the float. Most candidates do not lower anything;
// they are filtered by a plain read first (a stale read is only ever too large, so it can cost an atomic but
// never lose one). A last pass decodes the grid in place.
#include "ffb200_ctx.h"

#include <cstdlib>
#include <cstring>

namespace ffb200 {

namespace {

constexpr int kThreads = 128;
constexpr int kBlockWidth = 10;        // ParticleLevelSet::_blockwidth (particlelevelset.h:147)

// float <-> int, monotone: a < b  <=>  enc(a) < enc(b) (no NaNs here); its own inverse
__device__ __forceinline__ int order_bits(int b) { return b >= 0 ? b : b ^ 0x7fffffff; }

struct SdfParams {
    GridDesc g;
    int bi, bj, bk;                    // block grid
    float blockdx;                     // float blockdx = _blockwidth * _dx   (:441, :529)
    double inv_blockdx;                // 1.0 / blockdx                       (positionToGridIndex(p, blockdx))
    double chunk;                      // _blockwidth * _dx, a double         (:632)
    double hw;                         // 0.5 * dx                            (GridIndexToCellCenter, grid3d.h:100-103)
    float r, sr;                       // (float)radius, _searchRadiusFactor * r
    const float *px, *py, *pz;
    int n;
};

__global__ void k_sdf_fill(int *__restrict__ phi, size_t cells, int value) {
    for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) phi[c] = value;
}

// _initializeActiveBlocksThread: blocks that hold a particle
__global__ void __launch_bounds__(256) k_sdf_home(SdfParams P, uint8_t *__restrict__ home) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    const int a = pos2idx(P.px[j], P.inv_blockdx), b = pos2idx(P.py[j], P.inv_blockdx), c = pos2idx(P.pz[j], P.inv_blockdx);
    if (in_range3(a, b, c, P.bi, P.bj, P.bk)) {
        uint8_t *h = home + (a + P.bi * (b + P.bj * c));
        if (*h == 0) *h = 1;                                   // thousands of particles per block: read before the contended store
    }
}

// GridUtils::featherGrid26
__global__ void k_sdf_feather(const uint8_t *__restrict__ home, uint8_t *__restrict__ active, int bi, int bj, int bk) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bi * bj * bk) return;
    const int i = t % bi, j = (t / bi) % bj, k = t / (bi * bj);
    uint8_t on = 0;
    for (int c = -1; c <= 1; c++)
        for (int b = -1; b <= 1; b++)
            for (int a = -1; a <= 1; a++)
                if (in_range3(i + a, j + b, k + c, bi, bj, bk)) on |= home[(i + a) + bi * ((j + b) + bj * (k + c))];
    active[t] = on;
}

// one particle into one block: _computeExactBandProducerThread :636-662
__device__ __forceinline__ void sdf_block(const SdfParams &P, int ci, int cj, int ck, float x, float y, float z,
                                          int *__restrict__ phi) {
    const GridDesc &g = P.g;
    const float lx = x - idx2posf(ci, P.chunk), ly = y - idx2posf(cj, P.chunk), lz = z - idx2posf(ck, P.chunk);
    const float sr = P.sr;
    int i0 = pos2idx(lx - sr, g.inv_dx), j0 = pos2idx(ly - sr, g.inv_dx), k0 = pos2idx(lz - sr, g.inv_dx);
    int i1 = pos2idx(lx + sr, g.inv_dx), j1 = pos2idx(ly + sr, g.inv_dx), k1 = pos2idx(lz + sr, g.inv_dx);
    i0 = max(i0, 0); j0 = max(j0, 0); k0 = max(k0, 0);
    i1 = min(i1, kBlockWidth - 1); j1 = min(j1, kBlockWidth - 1); k1 = min(k1, kBlockWidth - 1);
    // cells of the block outside the grid are dropped by the write-out (:377-379)
    i1 = min(i1, g.I - 1 - ci * kBlockWidth); j1 = min(j1, g.J - 1 - cj * kBlockWidth); k1 = min(k1, g.K - 1 - ck * kBlockWidth);
    for (int k = k0; k <= k1; k++) {
        const float dz = (float)((double)(float)k * g.dx + P.hw) - lz;
        for (int j = j0; j <= j1; j++) {
            const float dy = (float)((double)(float)j * g.dx + P.hw) - ly;
            int *row = phi + ((size_t)(ci * kBlockWidth) + (size_t)g.I * ((size_t)(cj * kBlockWidth + j) + (size_t)g.J * (size_t)(ck * kBlockWidth + k)));
            for (int i = i0; i <= i1; i++) {
                const float dxc = (float)((double)(float)i * g.dx + P.hw) - lx;
                const float dist = vlen3(dxc, dy, dz) - P.r;
                const int e = order_bits(__float_as_int(dist));
                if (e < row[i]) atomicMin(row + i, e);
            }
        }
    }
}

// _computeGridCountDataThread :527-567 (which blocks) + the producer above (what it does there)
__global__ void __launch_bounds__(kThreads) k_sdf_scatter(SdfParams P, const uint8_t *__restrict__ active, int *__restrict__ phi) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n) return;
    const float x = P.px[t], y = P.py[t], z = P.pz[t];
    const float sr = P.sr, blockdx = P.blockdx;
    const int b0 = pos2idx(x, P.inv_blockdx), b1 = pos2idx(y, P.inv_blockdx), b2 = pos2idx(z, P.inv_blockdx);
    const double bdx = (double)blockdx;
    const float bx = idx2posf(b0, bdx), by = idx2posf(b1, bdx), bz = idx2posf(b2, bdx);
    int lo0 = b0, lo1 = b1, lo2 = b2, hi0 = b0, hi1 = b1, hi2 = b2;
    const bool simple = x - sr > bx && y - sr > by && z - sr > bz && x + sr < bx + blockdx && y + sr < by + blockdx &&
                        z + sr < bz + blockdx;
    if (!simple) {
        lo0 = pos2idx(x - sr, P.inv_blockdx); lo1 = pos2idx(y - sr, P.inv_blockdx); lo2 = pos2idx(z - sr, P.inv_blockdx);
        hi0 = pos2idx(x + sr, P.inv_blockdx); hi1 = pos2idx(y + sr, P.inv_blockdx); hi2 = pos2idx(z + sr, P.inv_blockdx);
    }
    lo0 = max(lo0, 0); lo1 = max(lo1, 0); lo2 = max(lo2, 0);                    // getBlockID == -1 outside the block grid
    hi0 = min(hi0, P.bi - 1); hi1 = min(hi1, P.bj - 1); hi2 = min(hi2, P.bk - 1);
    for (int ck = lo2; ck <= hi2; ck++)
        for (int cj = lo1; cj <= hi1; cj++)
            for (int ci = lo0; ci <= hi0; ci++)
                if (active[ci + P.bi * (cj + P.bj * ck)]) sdf_block(P, ci, cj, ck, x, y, z, phi);
}

// ---- per-axis variant (the default since it ran on hardware: 1.51 ms vs 2.24 ms at 128^3, field bit-identical;
// FFB200_SDF_VARIANT=0 selects k_sdf_scatter; its decomposition is also proven bit-exact on the CPU by
// tests/test_oracle_golden.py:test_liquid_sdf_axes_decomposition) --
//
// The reference's block set and its block-local cell boxes are products of per-axis ranges, and each squared
// distance term depends only on (axis, block index along the axis, local cell index). A particle is therefore
// reduced to three short lists in shared memory -- (global cell index | block offset << 24, (centre - local
// coordinate)^2) -- and the 3-D work is their product, gated by a 27-bit active-block mask: no conversions, no
// block arithmetic and no per-block loop nest inside, which is what held k_sdf_scatter at 9.8 active lanes. A
// squared-distance pre-filter against the value already loaded skips the IEEE square root wherever the candidate
// provably cannot lower the cell (sqrt and the subtraction are monotone; the margins cover every rounding).
constexpr int kAxisMax = 12;           // entries per axis: 2 sr / dx + 3 <= 12, i.e. sr <= 4.5 dx

__device__ __forceinline__ int sdf_axis_list(const SdfParams &P, float x, bool simple, int nb, int ncell, int *__restrict__ sh_g,
                                             float *__restrict__ sh_sq, int &lo_out) {
    const GridDesc &g = P.g;
    const int b = pos2idx(x, P.inv_blockdx);
    int lo = b, hi = b;
    if (!simple) {
        lo = pos2idx(x - P.sr, P.inv_blockdx);
        hi = pos2idx(x + P.sr, P.inv_blockdx);
    }
    lo = max(lo, 0);
    hi = min(hi, nb - 1);
    lo_out = lo;
    int n = 0;
    for (int c = lo; c <= hi; c++) {
        const float l = x - idx2posf(c, P.chunk);
        int i0 = max(pos2idx(l - P.sr, g.inv_dx), 0);
        int i1 = min(min(pos2idx(l + P.sr, g.inv_dx), kBlockWidth - 1), ncell - 1 - c * kBlockWidth);
        for (int i = i0; i <= i1 && n < kAxisMax; i++) {
            const float d = (float)((double)(float)i * g.dx + P.hw) - l;
            sh_g[n * kThreads] = (c * kBlockWidth + i) | ((c - lo) << 24);
            sh_sq[n * kThreads] = d * d;
            n++;
        }
    }
    return n;
}

__global__ void __launch_bounds__(kThreads) k_sdf_scatter_axes(SdfParams P, const uint8_t *__restrict__ active, int *phi) {
    __shared__ int sh_g[3 * kAxisMax * kThreads];
    __shared__ float sh_sq[3 * kAxisMax * kThreads];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n) return;
    const GridDesc &g = P.g;
    const float x = P.px[t], y = P.py[t], z = P.pz[t];
    const float sr = P.sr, blockdx = P.blockdx;
    const int b0 = pos2idx(x, P.inv_blockdx), b1 = pos2idx(y, P.inv_blockdx), b2 = pos2idx(z, P.inv_blockdx);
    const double bdx = (double)blockdx;
    const float bx = idx2posf(b0, bdx), by = idx2posf(b1, bdx), bz = idx2posf(b2, bdx);
    const bool simple = x - sr > bx && y - sr > by && z - sr > bz && x + sr < bx + blockdx && y + sr < by + blockdx &&
                        z + sr < bz + blockdx;
    int *gx = sh_g + threadIdx.x, *gy = gx + kAxisMax * kThreads, *gz = gy + kAxisMax * kThreads;
    float *qx = sh_sq + threadIdx.x, *qy = qx + kAxisMax * kThreads, *qz = qy + kAxisMax * kThreads;
    int lo0, lo1, lo2;
    const int nx = sdf_axis_list(P, x, simple, P.bi, g.I, gx, qx, lo0);
    const int ny = sdf_axis_list(P, y, simple, P.bj, g.J, gy, qy, lo1);
    const int nz = sdf_axis_list(P, z, simple, P.bk, g.K, gz, qz, lo2);
    if (nx == 0 || ny == 0 || nz == 0) return;
    // active blocks among the (at most 3 x 3 x 3) blocks the lists touch: bit u0 + 3 (u1 + 3 u2)
    const int hu0 = min(gx[(nx - 1) * kThreads] >> 24, 2), hu1 = min(gy[(ny - 1) * kThreads] >> 24, 2), hu2 = min(gz[(nz - 1) * kThreads] >> 24, 2);
    unsigned mask = 0u;
    for (int w = 0; w <= hu2; w++)
        for (int v = 0; v <= hu1; v++)
            for (int u = 0; u <= hu0; u++)
                if (active[(lo0 + u) + P.bi * ((lo1 + v) + P.bj * (lo2 + w))]) mask |= 1u << (u + 3 * (v + 3 * w));
    if (mask == 0u) return;
    const float r = P.r;
    for (int c = 0; c < nz; c++) {
        const int ez = gz[c * kThreads];
        const float sz = qz[c * kThreads];
        for (int b = 0; b < ny; b++) {
            const int ey = gy[b * kThreads];
            const unsigned m3 = (mask >> (3 * ((ey >> 24) + 3 * (ez >> 24)))) & 7u;
            if (m3 == 0u) continue;
            const float sy = qy[b * kThreads];
            int *row = phi + (size_t)g.I * ((size_t)(ey & 0xffffff) + (size_t)g.J * (size_t)(ez & 0xffffff));
            for (int a = 0; a < nx; a++) {
                const int ex = gx[a * kThreads];
                if (!((m3 >> (ex >> 24)) & 1u)) continue;
                const float d2 = qx[a * kThreads] + sy + sz;                 // (x*x + y*y) + z*z, vmath::length's order
                int *cell = row + (ex & 0xffffff);
                const int seen = *cell;
                const float cur = __int_as_float(order_bits(seen));
                const float reach = (cur + r) * 1.00001f;
                if (reach > 0.0f && d2 > reach * reach * 1.00001f) continue;   // sqrt(d2) - r >= cur for certain
                const int e = order_bits(__float_as_int(sqrtf(d2) - r));
                if (e < seen) atomicMin(cell, e);
            }
        }
    }
}

__global__ void k_sdf_decode(int *phi, size_t cells) {
    float *out = reinterpret_cast<float *>(phi);
    for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x)
        out[c] = __int_as_float(order_bits(phi[c]));
}

// ParticleLevelSet::postProcessSignedDistanceField (particlelevelset.cpp:170-195) with
// MeshLevelSet::getDistanceAtCellCenter (meshlevelset.cpp:152-162); one thread per cell. Written at the end of round 1
// with k_sdf_scatter_axes: pinned on the CPU, not yet run on hardware.
__global__ void __launch_bounds__(256) k_sdf_postprocess(GridDesc g, float *__restrict__ phi, const float *__restrict__ solid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= g.I) return;
    const size_t f = (size_t)i + (size_t)g.I * ((size_t)j + (size_t)g.J * (size_t)k);
    const int w = g.I + 1;
    const size_t sj = (size_t)w, sk = (size_t)w * (size_t)(g.J + 1);
    const float *s = solid + ((size_t)i + sj * (size_t)j + sk * (size_t)k);
    const float eps = (float)(0.005 * g.dx);
    float val = phi[f];
    if ((double)val < 0.5 * g.dx) {
        const float centre = 0.125f * (s[0] + s[1] + s[sj] + s[sj + 1] + s[sk] + s[sk + 1] + s[sk + sj] + s[sk + sj + 1]);
        if (centre < 0.0f) val = (float)(-0.5 * g.dx);
    }
    if (fabsf(val) < eps) val = val > 0.0f ? eps : -eps;
    phi[f] = val;
}

}  // namespace

int launch_liquid_sdf_postprocess(Context &c) {
    const GridDesc &g = c.g;
    if (g.kbase != 0 || g.kloc != g.K) throw CudaError("ffb200_postprocess_liquid_sdf: not available on z-slab contexts");
    if (!c.liquid_phi) throw CudaError("ffb200_postprocess_liquid_sdf: no liquid SDF on the device (ffb200_liquid_sdf first)");
    if (!c.has_solid) throw CudaError("ffb200_postprocess_liquid_sdf: needs the solid SDF (ffb200_set_solid) first");
    dim3 grid((g.I + 255) / 256, g.J, g.K);
    k_sdf_postprocess<<<grid, 256, 0, c.stream>>>(g, reinterpret_cast<float *>(c.liquid_phi), c.phi);
    FFB_CUDA(cudaGetLastError());
    return 1;
}

int launch_liquid_sdf(Context &c, double radius) {
    const GridDesc &g = c.g;
    if (g.kbase != 0 || g.kloc != g.K) throw CudaError("ffb200_liquid_sdf: not available on z-slab contexts");
    if (!(radius > 0.0)) throw CudaError("ffb200_liquid_sdf: the particle radius must be positive");
    SdfParams P;
    P.g = g;
    P.bi = (g.I + kBlockWidth - 1) / kBlockWidth;
    P.bj = (g.J + kBlockWidth - 1) / kBlockWidth;
    P.bk = (g.K + kBlockWidth - 1) / kBlockWidth;
    P.blockdx = (float)(kBlockWidth * g.dx);
    P.inv_blockdx = 1.0 / (double)P.blockdx;
    P.chunk = kBlockWidth * g.dx;
    P.hw = 0.5 * g.dx;
    P.r = (float)radius;
    P.sr = 2.0f * P.r;                                           // _searchRadiusFactor = 2.0f (particlelevelset.h:149)
    if (!(P.sr < P.blockdx)) throw CudaError("ffb200_liquid_sdf: the search radius must be smaller than a block (radius < 5 dx)");
    const ParticleSoA &s = c.soa[c.cur];
    P.px = s.p[0]; P.py = s.p[1]; P.pz = s.p[2];
    P.n = c.n;
    const size_t cells = (size_t)g.I * g.J * g.K;
    const size_t blocks = (size_t)P.bi * P.bj * P.bk;
    if (!c.liquid_phi) {
        FFB_CUDA(cudaMalloc(&c.liquid_phi, cells * sizeof(int)));
        FFB_CUDA(cudaMalloc(&c.liquid_blocks, 2 * blocks));
    }
    uint8_t *home = c.liquid_blocks, *active = c.liquid_blocks + blocks;
    const float maxd = (float)(3.0 * g.dx);                      // _getMaxDistance (:331-333)
    int maxbits;
    memcpy(&maxbits, &maxd, sizeof(maxbits));                    // positive: its own encoding
    const int fill_blocks = (int)((cells + 255) / 256 < (size_t)(16 * c.sm_count) ? (cells + 255) / 256 : (size_t)(16 * c.sm_count));
    k_sdf_fill<<<fill_blocks, 256, 0, c.stream>>>(c.liquid_phi, cells, maxbits);
    int launches = 1;
    if (c.n > 0) {
        FFB_CUDA(cudaMemsetAsync(home, 0, blocks, c.stream));
        k_sdf_home<<<(c.n + 255) / 256, 256, 0, c.stream>>>(P, home);
        k_sdf_feather<<<(int)((blocks + 127) / 128), 128, 0, c.stream>>>(home, active, P.bi, P.bj, P.bk);
        static const int variant = [] { const char *e = std::getenv("FFB200_SDF_VARIANT"); return e ? std::atoi(e) : 1; }();
        if (variant == 1 && 2.0 * (double)P.sr / g.dx + 3.0 <= (double)kAxisMax && g.I < (1 << 24) && g.J < (1 << 24) && g.K < (1 << 24))
            k_sdf_scatter_axes<<<(c.n + kThreads - 1) / kThreads, kThreads, 0, c.stream>>>(P, active, c.liquid_phi);
        else
            k_sdf_scatter<<<(c.n + kThreads - 1) / kThreads, kThreads, 0, c.stream>>>(P, active, c.liquid_phi);
        launches += 3;
    }
    k_sdf_decode<<<fill_blocks, 256, 0, c.stream>>>(c.liquid_phi, cells);
    launches++;
    FFB_CUDA(cudaGetLastError());
    return launches;
}

}  // namespace ffb200
