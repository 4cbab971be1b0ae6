// ffb200_p2g.cu -- particle-to-grid velocity transfer (VelocityAdvector::advect,
// velocityadvector.cpp:38-623) as an atomics-free, deterministic GATHER over sorted particles.
//
// One thread owns one MAC face. It walks the half-cell bins that can hold particles within the
// kernel radius (4x4x4 bins for the default r = 0.866 dx), in bin order, and accumulates
// sum(w*v) and sum(w) in registers: no atomics, no inter-CTA exchange, run-to-run identical.
//
// What the reference does per face, and how it is reproduced:
//  * weights are evaluated in the frame of the face's 10^3-node block: p_local = (p - offset)
//    - blockOrigin and gpos = (float)(i_local*dx) (velocityadvector.cpp:488-493, 510-512). The
//    fast path below uses exactly these float operations, so FLIP weights are bit-identical.
//  * a particle contributes to a face only if it was sorted into the face's block
//    (_computeGridCountDataThread, :296-353). That membership ("seam word") is evaluated once
//    per particle with the reference's own mixed float/double arithmetic; for APIC it is a
//    first-order effect (the block-seam drop of :596-599).
//  * sums run in ascending particle index inside the block (:383-413). The fast path sums in
//    bin order instead; faces whose weight sum lands within a guard band of the 1e-6 validity
//    threshold (:160, :527) are re-summed by exact_face() in the reference's order with the
//    reference's exact arithmetic, so the valid mask is bit-exact.
#include "ffb200_ctx.h"
#include "ffb200_seam.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace ffb200 {

namespace {

struct OverflowEntry;
struct CellRec;

struct P2GParams {
    GridDesc g;
    int gi, gj, gk;              // global face dims of this direction
    int kstore;                  // stored face planes, first one is g.kbase
    int kw0, kw1;                // global face planes [kw0, kw1) this rank must produce (owned + 1)
    int bi, bj, bk;              // block dims
    const uint8_t *active;
    const uint32_t *bin_start;
    const float *px, *py, *pz;   // sorted positions
    const float *vel;            // sorted velocity component of this direction
    const float *ax, *ay, *az;   // sorted affine row of this direction (APIC)
    const uint32_t *seam;        // membership word of this direction per sorted slot
    const uint32_t *edge_list;   // sorted slots of this direction's "edge" particles (unordered)
    const uint32_t *edge_count;  // how many k_seam_home found (may exceed edge_cap: list overflowed)
    uint32_t edge_cap;
    const uint32_t *orig;
    float2 *partial;             // cell-partial splat: 8 (sum w, sum w*v) pairs per shifted cell
    uint8_t *cell_flag;          // 1 if the cell holds particles (its partial slot is valid)
    CellRec *cell_list;          // occupied shifted cells: plain ones from the front, seam ones from the back
    uint32_t *list_count;        // [0] plain, [1] seam
    uint32_t list_cap;
    int ccx, ccy, ccz, ck0;      // shifted-cell grid: cells b = (ix-1, iy-1, iz+ck0)
    OverflowEntry *ovf;          // contributions of edge particles outside their bin cell
    int *ovf_count;
    float *out, *wsum;
    uint8_t *valid;
    float off[3];                // _getDirectionOffset, velocityadvector.cpp:177-188
    float r, sr, rsq, c1, c2, c3;
    float inv_s;                 // (float)(1.0 / (float)dx): vec3 / _dx of :574
    double chunk;                // _chunkWidth * _dx
    int wm;                      // half-cell window half width
    float guard_abs, guard_per;
    int out_stride = 1;          // floats between consecutive outputs (one channel of an interleaved vec3 grid)
    int vec3_norm = 0;           // normalise with vmath's vec3 /= float: x * (float)(1.0 / w) (vmath.cpp:105-111)
    int no_norm = 0;             // AttributeTransferParameters::normalize == false: keep the weighted sum
};

__global__ void __launch_bounds__(256) k_seam_home(const __grid_constant__ SeamParams s) {
    __shared__ int sh[3 * 256];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int hidx[3] = {-1, -1, -1};
    if (j < s.n) seam_particle(s, s.px[j], s.py[j], s.pz[j], j, hidx);
    seam_mark_home(s, hidx, sh);
}

// featherGrid26 (gridutils.cpp:264-297) as a gather: active = OR of home over the 3x3x3 stencil.
__global__ void k_dilate26(const uint8_t *__restrict__ home, uint8_t *__restrict__ active, int bi, int bj, int bk) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bi * bj * bk) return;
    const int i = t % bi, j = (t / bi) % bj, k = t / (bi * bj);
    uint8_t a = 0;
    for (int c = -1; c <= 1; c++)
        for (int b = -1; b <= 1; b++)
            for (int d = -1; d <= 1; d++)
                if (in_range3(i + d, j + b, k + c, bi, bj, bk)) a |= home[(i + d) + bi * ((j + b) + bj * (k + c))];
    active[t] = a;
}

__device__ __forceinline__ bool seam_member(uint32_t word, int nbx, int nby, int nbz) {
    const int lx = (int)(word & 255u) - 1, sx = (int)((word >> 8) & 3u);
    const int ly = (int)((word >> 10) & 255u) - 1, sy = (int)((word >> 18) & 3u);
    const int lz = (int)((word >> 20) & 255u) - 1, sz = (int)((word >> 28) & 3u);
    return (unsigned)(nbx - lx) <= (unsigned)sx && (unsigned)(nby - ly) <= (unsigned)sy &&
           (unsigned)(nbz - lz) <= (unsigned)sz;
}

struct FaceFrame {
    int nb[3];        // block of the face
    int lo[3];        // local node index inside the block
    float bpos[3];    // block origin, GridIndexToPosition(blockIndex, _chunkWidth*_dx)
    float gpos[3];    // local node position, GridIndexToPosition(i, j, k, _dx)
    float gposm[3];   // position of the previous local node (lo - 1)
    int h0[3], h1[3]; // inclusive half-cell window (bin-grid coordinates, apron included)
};

// Contribution of sorted particle q to the face, with every test the reference makes.
// Returns false when the particle does not contribute.
template <int DIR, int METHOD>
__device__ __forceinline__ bool exact_contribution(const P2GParams &P, const FaceFrame &f, uint32_t q, float &w_out,
                                                   float &wv_out) {
    if (!seam_member(P.seam[q], f.nb[0], f.nb[1], f.nb[2])) return false;
    float xl[3];
    xl[0] = (P.px[q] - P.off[0]) - f.bpos[0];
    xl[1] = (P.py[q] - P.off[1]) - f.bpos[1];
    xl[2] = (P.pz[q] - P.off[2]) - f.bpos[2];
    const float velocity = P.vel[q];
    if (METHOD == FFB200_TRANSFER_FLIP) {
#pragma unroll
        for (int a = 0; a < 3; a++) {                        // node range of :495-506
            int gmin = pos2idx(xl[a] - P.sr, P.g.inv_dx), gmax = pos2idx(xl[a] + P.sr, P.g.inv_dx);
            gmin = gmin < 0 ? 0 : gmin;
            gmax = gmax > kChunk - 1 ? kChunk - 1 : gmax;
            if (f.lo[a] < gmin || f.lo[a] > gmax) return false;
        }
        const float vx = f.gpos[0] - xl[0], vy = f.gpos[1] - xl[1], vz = f.gpos[2] - xl[2];
        const float d2 = vx * vx + vy * vy + vz * vz;
        if (!(d2 < P.rsq)) return false;
        const float w = 1.0f - P.c1 * d2 * d2 * d2 + P.c2 * d2 * d2 - P.c3 * d2;
        w_out = w;
        wv_out = w * velocity;
        return true;
    } else {
        float fac[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {                        // :567-592
            const int gidx = pos2idx(xl[a], P.g.inv_dx);
            const float ip = (xl[a] - idx2posf(gidx, P.g.dx)) * P.inv_s;
            const int c = f.lo[a] - gidx;
            if (c != 0 && c != 1) return false;
            fac[a] = c ? ip : (1.0f - ip);
        }
        const float w = fac[0] * fac[1] * fac[2];
        const float apic = P.ax[q] * (f.gpos[0] - xl[0]) + P.ay[q] * (f.gpos[1] - xl[1]) + P.az[q] * (f.gpos[2] - xl[2]);
        w_out = w;
        wv_out = w * (velocity + apic);
        return true;
    }
}

// The reference's sum for one face: contributions in ascending original particle index.
template <int DIR, int METHOD>
__device__ __noinline__ void exact_face(const P2GParams &P, const FaceFrame &f, float &sw, float &swv) {
    sw = 0.0f;
    swv = 0.0f;
    // one bin wider than the fast window: a particle a rounding error outside it can still
    // pass the reference's tests in the block-local frame
    const int H[3] = {P.g.HX, P.g.HY, P.g.HZ};
    int e0[3], e1[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        e0[a] = f.h0[a] > 0 ? f.h0[a] - 1 : 0;
        e1[a] = f.h1[a] < H[a] - 1 ? f.h1[a] + 1 : H[a] - 1;
    }
    long long last = -1;
    for (;;) {
        long long best = 0x7fffffffffffffffLL;
        uint32_t bestq = 0;
        for (int hz = e0[2]; hz <= e1[2]; hz++)
            for (int hy = e0[1]; hy <= e1[1]; hy++) {
                const size_t row = ((size_t)hz * P.g.HY + hy) * P.g.HX;
                const uint32_t s = P.bin_start[row + e0[0]], e = P.bin_start[row + e1[0] + 1];
                for (uint32_t q = s; q < e; q++) {
                    const long long o = P.orig[q];
                    if (o > last && o < best) { best = o; bestq = q; }
                }
            }
        if (best == 0x7fffffffffffffffLL) break;
        last = best;
        float w, wv;
        if (exact_contribution<DIR, METHOD>(P, f, bestq, w, wv)) {
            swv += wv;
            sw += w;
        }
    }
}

template <int DIR, int METHOD>
__global__ void __launch_bounds__(256) k_p2g(const __grid_constant__ P2GParams P) {
    const int ni = blockIdx.x * blockDim.x + threadIdx.x;
    const int nj = blockIdx.y * blockDim.y + threadIdx.y;
    const int ks = blockIdx.z * blockDim.z + threadIdx.z;     // stored plane
    if (ni >= P.gi || nj >= P.gj || ks >= P.kstore) return;
    const int nk = ks + P.g.kbase;
    const size_t fidx = (size_t)ni + (size_t)P.gi * ((size_t)nj + (size_t)P.gj * ks);

    FaceFrame f;
    const int n[3] = {ni, nj, nk};
    const int H[3] = {P.g.HX, P.g.HY, P.g.HZ};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        f.nb[a] = n[a] / kChunk;
        f.lo[a] = n[a] - f.nb[a] * kChunk;
        f.bpos[a] = idx2posf(f.nb[a], P.chunk);
        f.gpos[a] = idx2posf(f.lo[a], P.g.dx);
        f.gposm[a] = idx2posf(f.lo[a] - 1, P.g.dx);
        const int c = 2 * n[a] + (a == DIR ? 0 : 1) + kApron - (a == 2 ? 2 * P.g.kbase : 0);
        int h0 = c - P.wm, h1 = c + P.wm - 1;
        f.h0[a] = h0 < 0 ? 0 : h0;
        f.h1[a] = h1 > H[a] - 1 ? H[a] - 1 : h1;
    }
    if (!P.active[f.nb[0] + P.bi * (f.nb[1] + P.bj * f.nb[2])]) {
        P.out[fidx * (size_t)P.out_stride] = 0.0f;
        P.wsum[fidx] = 0.0f;
        P.valid[fidx] = 0;
        return;
    }

    float sw = 0.0f, swv = 0.0f;
    int cnt = 0;
    for (int hz = f.h0[2]; hz <= f.h1[2]; hz++)
        for (int hy = f.h0[1]; hy <= f.h1[1]; hy++) {
            const size_t row = ((size_t)hz * P.g.HY + hy) * P.g.HX;
            const uint32_t s = __ldg(P.bin_start + row + f.h0[0]), e = __ldg(P.bin_start + row + f.h1[0] + 1);
            for (uint32_t q = s; q < e; q++) {
                if (!seam_member(__ldg(P.seam + q), f.nb[0], f.nb[1], f.nb[2])) continue;
                const float xl0 = (__ldg(P.px + q) - P.off[0]) - f.bpos[0];
                const float xl1 = (__ldg(P.py + q) - P.off[1]) - f.bpos[1];
                const float xl2 = (__ldg(P.pz + q) - P.off[2]) - f.bpos[2];
                const float vx = f.gpos[0] - xl0, vy = f.gpos[1] - xl1, vz = f.gpos[2] - xl2;
                if (METHOD == FFB200_TRANSFER_FLIP) {
                    // same float operations as the reference: the weight is bit-identical
                    const float d2 = vx * vx + vy * vy + vz * vz;
                    if (d2 < P.rsq) {
                        const float w = 1.0f - P.c1 * d2 * d2 * d2 + P.c2 * d2 * d2 - P.c3 * d2;
                        swv += w * __ldg(P.vel + q);
                        sw += w;
                        cnt++;
                    }
                } else {
                    // Trilinear factors exactly as the reference forms them (:567-592): a particle
                    // at or above the node is in the node's cell (factor 1 - ipos), one below it
                    // is in the previous cell (factor ipos measured from the previous node). The
                    // float comparison stands in for the reference's double floor; the two can
                    // only disagree within ~6e-7 cells of a node plane, and those pairs take the
                    // reference's exact arithmetic below, so every weight is bit-identical.
                    const bool upx = vx <= 0.0f, upy = vy <= 0.0f, upz = vz <= 0.0f;
                    const float t0 = (xl0 - (upx ? f.gpos[0] : f.gposm[0])) * P.inv_s;
                    const float t1 = (xl1 - (upy ? f.gpos[1] : f.gposm[1])) * P.inv_s;
                    const float t2 = (xl2 - (upz ? f.gpos[2] : f.gposm[2])) * P.inv_s;
                    // within a few ulps of a node plane: let the reference's own arithmetic decide
                    const float band = 4e-6f;
                    const bool edge = fabsf(t0) < band || fabsf(t0 - 1.0f) < band || fabsf(t1) < band ||
                                      fabsf(t1 - 1.0f) < band || fabsf(t2) < band || fabsf(t2 - 1.0f) < band;
                    if (edge) {
                        float w, wv;
                        if (exact_contribution<DIR, METHOD>(P, f, q, w, wv)) {
                            swv += wv;
                            sw += w;
                            cnt++;
                        }
                        continue;
                    }
                    const float fx = upx ? 1.0f - t0 : t0;
                    const float fy = upy ? 1.0f - t1 : t1;
                    const float fz = upz ? 1.0f - t2 : t2;
                    if (fx > 0.0f && fy > 0.0f && fz > 0.0f && t0 >= 0.0f && t1 >= 0.0f && t2 >= 0.0f) {
                        const float w = fx * fy * fz;
                        const float apic = __ldg(P.ax + q) * vx + __ldg(P.ay + q) * vy + __ldg(P.az + q) * vz;
                        swv += w * (__ldg(P.vel + q) + apic);
                        sw += w;
                        cnt++;
                    }
                }
            }
        }

    const float eps = 1e-6f;
    if (fabsf(sw - eps) <= P.guard_abs + P.guard_per * (float)cnt) exact_face<DIR, METHOD>(P, f, sw, swv);
    float s = swv;
    if (sw > eps && !P.no_norm) s = P.vec3_norm ? s * finv(sw) : s / sw;   // :527-531; attributetogridtransfer.h vec3 flavour
    P.out[fidx * (size_t)P.out_stride] = s;                    // write-out :155-162
    P.wsum[fidx] = sw;
    P.valid[fidx] = sw > eps ? 1 : 0;
}

// Membership word and home-block mark of the cell-centred "fourth direction" (offset (dx/2, dx/2, dx/2)) of
// AttributeToGridTransfer<T>::transfer (attributetogridtransfer.h:213-330: the same block structure as the velocity
// transfer), for its own radius. FLIP kernel only, so no edge flags.
__global__ void __launch_bounds__(256) k_seam_cell(const __grid_constant__ SeamParams s, uint32_t *__restrict__ words,
                                                   uint8_t *__restrict__ home, int bi, int bj, int bk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= s.n) return;
    const AxisSeam X = axis_seam(s.px[j] - s.h, s), Y = axis_seam(s.py[j] - s.h, s), Z = axis_seam(s.pz[j] - s.h, s);
    if (in_range3(X.home, Y.home, Z.home, bi, bj, bk)) home[X.home + bi * (Y.home + bj * Z.home)] = 1;
    const bool simple = X.simple && Y.simple && Z.simple;
    words[j] = simple ? (X.fs | (Y.fs << 10) | (Z.fs << 20)) : (X.fn | (Y.fn << 10) | (Z.fn << 20));
}

// ---- shared-memory brick kernel -------------------------------------------------------------------
//
// One CTA owns a brick of 10 x 5 x 5 faces, a quarter of ONE reference 10^3 block, so the block
// frame (blockOrigin) and the block membership test are CTA-uniform. The particles of the
// brick's bin region (22 x 12 x 12 half cells for the default radius) are staged into shared
// memory once, already transformed: (p - offset) - blockOrigin as the reference forms it,
// non-members parked at +inf so they fail every support test, near-plane ("edge") particles
// diverted to a short list that takes the exact arithmetic. The face loop then costs one
// LDS.128 (two for APIC) and ~15-35 ALU instructions per candidate instead of 5-8 global loads
// and a membership decode. Summation order per face is unchanged (bin order), so results are
// bit-identical to k_p2g except where edge particles are involved (they are added last).
constexpr int kBrickX = 10, kBrickY = 5, kBrickZ = 5;
constexpr int kBrickThreads = 256;
constexpr int kMaxRows = 256;        // (2*5 - 2 + 2*wm)^2 with wm <= 4
constexpr int kMaxRX = 26;           // 2*10 - 2 + 2*wm
constexpr int kMaxFlagged = 96;

template <int METHOD>
struct BrickCap {
    static constexpr int value = METHOD == FFB200_TRANSFER_APIC ? 1536 : 2048;
};

struct BrickShared {
    uint32_t row_gstart[kMaxRows];
    uint32_t row_off[kMaxRows + 1];
    uint16_t binoff[kMaxRows][kMaxRX + 2];
    int chunk_row[66];
    int nchunks;
    int nflag;
    int oversize;
    uint32_t flagged[kMaxFlagged];
    uint32_t scan[33];
};

template <int DIR, int METHOD>
__device__ __forceinline__ void accumulate_global(const P2GParams &P, const FaceFrame &f, float &sw, float &swv, int &cnt);

template <int DIR, int METHOD>
__global__ void __launch_bounds__(kBrickThreads) k_p2g_brick(const __grid_constant__ P2GParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CAP = BrickCap<METHOD>::value;
    float4 *rec = reinterpret_cast<float4 *>(smem_raw);
    float4 *rec2 = rec + (METHOD == FFB200_TRANSFER_APIC ? CAP : 0);
    BrickShared &S = *reinterpret_cast<BrickShared *>(smem_raw + sizeof(float4) * CAP * (METHOD == FFB200_TRANSFER_APIC ? 2 : 1));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // brick -> block and face origin
    const int nbx = (int)blockIdx.x, nby = (int)(blockIdx.y >> 1), nbz = (int)(blockIdx.z >> 1) + P.g.kbase / kChunk;
    const int o[3] = {nbx * kChunk, nby * kChunk + (int)(blockIdx.y & 1) * kBrickY,
                      nbz * kChunk + (int)(blockIdx.z & 1) * kBrickZ};
    const int nbv[3] = {nbx, nby, nbz};
    const int H[3] = {P.g.HX, P.g.HY, P.g.HZ};
    const int dims[3] = {P.gi, P.gj, P.gk};
    const int bsz[3] = {kBrickX, kBrickY, kBrickZ};

    // this thread's face
    const int fxl = tid % kBrickX, fyl = (tid / kBrickX) % kBrickY, fzl = tid / (kBrickX * kBrickY);
    const int n[3] = {o[0] + fxl, o[1] + fyl, o[2] + fzl};
    const int ks = n[2] - P.g.kbase;
    const bool live = tid < kBrickX * kBrickY * kBrickZ && n[0] < P.gi && n[1] < P.gj && n[2] < P.gk && ks >= 0 &&
                      ks < P.kstore;
    const size_t fidx = live ? (size_t)n[0] + (size_t)P.gi * ((size_t)n[1] + (size_t)P.gj * ks) : 0;

    if (nbx >= P.bi || nby >= P.bj || nbz >= P.bk) return;
    if (!P.active[nbx + P.bi * (nby + P.bj * nbz)]) {
        if (live) { P.out[fidx] = 0.0f; P.wsum[fidx] = 0.0f; P.valid[fidx] = 0; }
        return;
    }

    // region of bins the brick's faces can see, per axis (bin-grid coordinates)
    int r0[3], rn[3];
    float bpos[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int last = min(o[a] + bsz[a] - 1, dims[a] - 1);
        const int c0 = 2 * o[a] + (a == DIR ? 0 : 1) + kApron - (a == 2 ? 2 * P.g.kbase : 0);
        const int c1 = 2 * last + (a == DIR ? 0 : 1) + kApron - (a == 2 ? 2 * P.g.kbase : 0);
        const int lo = max(c0 - P.wm, 0), hi = min(c1 + P.wm - 1, H[a] - 1);
        r0[a] = lo;
        rn[a] = max(hi - lo + 1, 0);
        bpos[a] = idx2posf(nbv[a], P.chunk);
    }
    const int nrows = rn[1] * rn[2];

    FaceFrame f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        f.nb[a] = nbv[a];
        f.lo[a] = n[a] - nbv[a] * kChunk;
        f.bpos[a] = bpos[a];
        f.gpos[a] = idx2posf(f.lo[a], P.g.dx);
        f.gposm[a] = idx2posf(f.lo[a] - 1, P.g.dx);
        const int c = 2 * n[a] + (a == DIR ? 0 : 1) + kApron - (a == 2 ? 2 * P.g.kbase : 0);
        f.h0[a] = max(c - P.wm, 0);
        f.h1[a] = min(c + P.wm - 1, H[a] - 1);
    }

    // ---- row table: global start and length of every (hy, hz) row of the region --------------------
    if (tid == 0) { S.nflag = 0; S.oversize = 0; }
    uint32_t mylen = 0;
    if (tid < nrows) {
        const int ry = tid % rn[1], rz = tid / rn[1];
        const size_t row = ((size_t)(r0[2] + rz) * P.g.HY + (r0[1] + ry)) * P.g.HX + r0[0];
        const uint32_t s = __ldg(P.bin_start + row), e = __ldg(P.bin_start + row + rn[0]);
        S.row_gstart[tid] = s;
        mylen = e - s;
    }
    {   // exclusive scan of the row lengths (nrows <= 256 = blockDim)
        uint32_t inc = mylen;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) S.scan[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = lane < kBrickThreads / 32 ? S.scan[lane] : 0u, winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += t;
            }
            S.scan[lane] = winc - w;
        }
        __syncthreads();
        const uint32_t ex = S.scan[warp] + inc - mylen;
        if (tid < nrows) {
            S.row_off[tid] = ex;
            if (mylen > (uint32_t)CAP || mylen > 65535u) S.oversize = 1;
        }
        if (tid == nrows - 1 || (nrows == 0 && tid == 0)) S.row_off[nrows] = nrows ? ex + mylen : 0u;
    }
    // per-bin offsets inside each row
    for (int t = tid; t < nrows * (rn[0] + 1); t += kBrickThreads) {
        const int r = t / (rn[0] + 1), b = t - r * (rn[0] + 1);
        const int ry = r % rn[1], rz = r / rn[1];
        const size_t row = ((size_t)(r0[2] + rz) * P.g.HY + (r0[1] + ry)) * P.g.HX + r0[0];
        S.binoff[r][b] = (uint16_t)(__ldg(P.bin_start + row + b) - __ldg(P.bin_start + row));
    }
    __syncthreads();
    const uint32_t total = S.row_off[nrows];

    float sw = 0.0f, swv = 0.0f;
    int cnt = 0;
    if (tid == 0 && !S.oversize) {   // greedy row chunks of at most CAP particles
        int nc = 0, start = 0;
        S.chunk_row[0] = 0;
        for (int r = 0; r < nrows; r++) {
            if (S.row_off[r + 1] - S.row_off[start] > (uint32_t)CAP) {
                if (nc >= 63) { S.oversize = 1; break; }
                S.chunk_row[++nc] = r;
                start = r;
            }
        }
        S.chunk_row[++nc] = nrows;
        S.nchunks = nc;
    }
    __syncthreads();
    if (S.oversize) {
        // the region does not fit the staging scheme (hundreds of particles per half cell):
        // this brick reads its candidates from global memory instead
        if (live) accumulate_global<DIR, METHOD>(P, f, sw, swv, cnt);
    } else if (total > 0) {
        const int nchunks = S.nchunks;
        for (int ch = 0; ch < nchunks; ch++) {
            const int rc0 = S.chunk_row[ch], rc1 = S.chunk_row[ch + 1];
            const uint32_t base = S.row_off[rc0];
            // ---- stage: one warp per row, coalesced reads of the sorted streams -----------------------
            for (int r = rc0 + warp; r < rc1; r += kBrickThreads / 32) {
                const uint32_t g0 = S.row_gstart[r], len = S.row_off[r + 1] - S.row_off[r];
                const uint32_t dst0 = S.row_off[r] - base;
                for (uint32_t i = lane; i < len; i += 32) {
                    const uint32_t q = g0 + i;
                    const uint32_t word = __ldg(P.seam + q);
                    const bool member = seam_member(word, nbv[0], nbv[1], nbv[2]);
                    const bool edge = METHOD == FFB200_TRANSFER_APIC && (word & kEdgeBit) != 0;
                    float4 v4;
                    v4.x = (__ldg(P.px + q) - P.off[0]) - bpos[0];
                    v4.y = (__ldg(P.py + q) - P.off[1]) - bpos[1];
                    v4.z = (__ldg(P.pz + q) - P.off[2]) - bpos[2];
                    v4.w = __ldg(P.vel + q);
                    if (!member || edge) v4.x = __int_as_float(0x7f800000);     // +inf: fails every support test
                    rec[dst0 + i] = v4;
                    if (METHOD == FFB200_TRANSFER_APIC)
                        rec2[dst0 + i] = make_float4(__ldg(P.ax + q), __ldg(P.ay + q), __ldg(P.az + q), 0.0f);
                    if (member && edge) {
                        const int slot = atomicAdd(&S.nflag, 1);
                        if (slot < kMaxFlagged) S.flagged[slot] = q;
                    }
                }
            }
            __syncthreads();
            // ---- face loop over the staged rows --------------------------------------------------------
            if (live) {
                const int xb0 = f.h0[0] - r0[0], xb1 = f.h1[0] + 1 - r0[0];
                for (int hz = f.h0[2]; hz <= f.h1[2]; hz++) {
                    const int rz = hz - r0[2];
                    for (int hy = f.h0[1]; hy <= f.h1[1]; hy++) {
                        const int r = rz * rn[1] + (hy - r0[1]);
                        if (r < rc0 || r >= rc1) continue;
                        const uint32_t rb = S.row_off[r] - base;
                        const uint32_t s = rb + S.binoff[r][xb0], e = rb + S.binoff[r][xb1];
                        for (uint32_t q = s; q < e; q++) {
                            const float4 pr = rec[q];
                            const float vx = f.gpos[0] - pr.x, vy = f.gpos[1] - pr.y, vz = f.gpos[2] - pr.z;
                            if (METHOD == FFB200_TRANSFER_FLIP) {
                                const float d2 = vx * vx + vy * vy + vz * vz;
                                if (d2 < P.rsq) {
                                    const float w = 1.0f - P.c1 * d2 * d2 * d2 + P.c2 * d2 * d2 - P.c3 * d2;
                                    swv += w * pr.w;
                                    sw += w;
                                    cnt++;
                                }
                            } else {
                                const bool upx = vx <= 0.0f, upy = vy <= 0.0f, upz = vz <= 0.0f;
                                const float t0 = (pr.x - (upx ? f.gpos[0] : f.gposm[0])) * P.inv_s;
                                const float t1 = (pr.y - (upy ? f.gpos[1] : f.gposm[1])) * P.inv_s;
                                const float t2 = (pr.z - (upz ? f.gpos[2] : f.gposm[2])) * P.inv_s;
                                const float fx = upx ? 1.0f - t0 : t0;
                                const float fy = upy ? 1.0f - t1 : t1;
                                const float fz = upz ? 1.0f - t2 : t2;
                                if (fx > 0.0f && fy > 0.0f && fz > 0.0f) {
                                    const float4 af = rec2[q];
                                    const float w = fx * fy * fz;
                                    const float apic = af.x * vx + af.y * vy + af.z * vz;
                                    swv += w * (pr.w + apic);
                                    sw += w;
                                    cnt++;
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        // ---- edge particles: the reference's exact arithmetic, in a fixed (sorted) order ----------------
        const int nflag = S.nflag;
        if (nflag > kMaxFlagged) {
            // pathological: redo this brick from global memory with the per-pair exact test
            sw = 0.0f; swv = 0.0f; cnt = 0;
            if (live) accumulate_global<DIR, METHOD>(P, f, sw, swv, cnt);
        } else if (nflag > 0) {
            if (tid == 0) {   // slots were claimed in arbitrary order: sort the few entries
                for (int i = 1; i < nflag; i++) {
                    const uint32_t key = S.flagged[i];
                    int j = i - 1;
                    while (j >= 0 && S.flagged[j] > key) { S.flagged[j + 1] = S.flagged[j]; j--; }
                    S.flagged[j + 1] = key;
                }
            }
            __syncthreads();
            if (live) {
                for (int i = 0; i < nflag; i++) {
                    float w, wv;
                    if (exact_contribution<DIR, METHOD>(P, f, S.flagged[i], w, wv)) {
                        swv += wv;
                        sw += w;
                        cnt++;
                    }
                }
            }
        }
    }

    if (!live) return;
    const float eps = 1e-6f;
    if (fabsf(sw - eps) <= P.guard_abs + P.guard_per * (float)cnt) exact_face<DIR, METHOD>(P, f, sw, swv);
    float s = swv;
    if (sw > eps) s /= sw;                                     // :527-531
    P.out[fidx] = s;                                           // write-out :155-162
    P.wsum[fidx] = sw;
    P.valid[fidx] = sw > eps ? 1 : 0;
}

// The candidate walk of k_p2g as a device function (fallback of the brick kernel).
template <int DIR, int METHOD>
__device__ __forceinline__ void accumulate_global(const P2GParams &P, const FaceFrame &f, float &sw, float &swv, int &cnt) {
    for (int hz = f.h0[2]; hz <= f.h1[2]; hz++)
        for (int hy = f.h0[1]; hy <= f.h1[1]; hy++) {
            const size_t row = ((size_t)hz * P.g.HY + hy) * P.g.HX;
            const uint32_t s = __ldg(P.bin_start + row + f.h0[0]), e = __ldg(P.bin_start + row + f.h1[0] + 1);
            for (uint32_t q = s; q < e; q++) {
                float w, wv;
                if (exact_contribution<DIR, METHOD>(P, f, q, w, wv)) {
                    swv += wv;
                    sw += w;
                    cnt++;
                }
            }
        }
}

// ---- cell-partial splat (default) -------------------------------------------------------------------
//
// The splat again, but with no CTA structure at all, so every particle is visited exactly once
// per direction and all lanes stay busy:
//
//   k_p2g_cell_list  lists the occupied SHIFTED CELLS (the staggered-frame cells whose corner nodes
//                 are base + {0,1}^3; a cell's particles are four runs of the sorted streams) in two
//                 classes: plain (one 10^3 block holds all 8 corners) and seam / border.
//   k_p2g_cells   one thread per listed cell, so warps are full whatever the fill pattern. The
//                 thread splats its particles onto its own 8 corner nodes -- each in the frame of
//                 THAT node's 10^3 block, with that block's membership test, exactly as the
//                 reference would when it processes the block -- and stores the 8 (sum w, sum w*v)
//                 pairs in its private slot of `partial`. No two threads share a slot: no atomics,
//                 no colouring, no barriers. Plain cells (73 %) share the per-axis factors between
//                 the corners: ~80 float operations per particle and direction.
//   k_p2g_nodes   one thread per face: adds the 8 partial sums of the cells around it in a fixed
//                 order, then the guard-band / exact_face / normalise / valid-byte epilogue.
//
// Summation order per face: cell order (fixed), sorted particle order inside a cell: bitwise
// deterministic. APIC "edge" particles are skipped by the splat and added by k_p2g_edge with
// exact_contribution(): to their own cell's partial sums, and -- the rare case that the
// reference's double floor puts such a particle in a neighbouring cell, so that it also reaches
// nodes outside its bin cell -- to a short overflow list that k_p2g_nodes adds in sorted order.
struct OverflowEntry {
    uint32_t node;      // flat stored face index
    uint32_t q;         // sorted particle slot
    float w, wv;
};
constexpr int kOverflowCap = 4096;

struct AxisNodes {
    int nb[2];          // block of node 0 / node 1 (-1: node outside the face grid)
    float bpos[2];      // blockOrigin of that block
    float gpos[2];      // local node position in that block's frame
    float gposm1;       // position of local node (lo1 - 1) in node 1's frame
    int lo[2];
};

__device__ __forceinline__ AxisNodes axis_nodes(int b, int dim, double chunk, double dx) {
    AxisNodes a;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int n = b + j;
        const bool ok = n >= 0 && n < dim;
        const int nb = ok ? n / kChunk : 0;
        a.nb[j] = ok ? nb : -1;
        a.lo[j] = n - nb * kChunk;
        a.bpos[j] = idx2posf(nb, chunk);
        a.gpos[j] = idx2posf(a.lo[j], dx);
    }
    a.gposm1 = idx2posf(a.lo[1] - 1, dx);
    return a;
}

// One listed (occupied) shifted cell: its flat index and its four runs of sorted slots (two
// half-cell rows in y times two in z; the two x half-cells of a row are adjacent bins).
struct __align__(16) CellRec {
    uint32_t t;         // flat shifted-cell index
    uint32_t rs[4];     // first slot of each run
    uint32_t n01, n23;  // run lengths, 16 bits each (0xffffffff in n01: lengths do not fit, re-read the bin table)
    uint32_t seam;      // class: 0 plain, 1 seam / border
};

// Loads are unconditional on clamped indices so that all eight are in flight together. The bin
// table has fewer than 2^32 entries (32-bit keys), so 32-bit index arithmetic is exact.
template <int DIR>
__device__ __forceinline__ uint32_t cell_runs(const P2GParams &P, const int b[3], uint32_t rs[4], uint32_t re[4]) {
    const int H[3] = {P.g.HX, P.g.HY, P.g.HZ};
    int hb[3];
#pragma unroll
    for (int a = 0; a < 3; a++) hb[a] = 2 * b[a] + (a == DIR ? 0 : 1) + kApron - (a == 2 ? 2 * P.g.kbase : 0);
    const int hx0 = max(hb[0], 0), hx1 = min(hb[0] + 1, H[0] - 1);
    uint32_t i0[4], i1[4];
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int hz = hb[2] + (rr >> 1), hy = hb[1] + (rr & 1);
        const bool ok = hx0 <= hx1 && (unsigned)hz < (unsigned)H[2] && (unsigned)hy < (unsigned)H[1];
        const uint32_t row = ((uint32_t)hz * (uint32_t)P.g.HY + (uint32_t)hy) * (uint32_t)P.g.HX;
        i0[rr] = ok ? row + (uint32_t)hx0 : 0u;
        i1[rr] = ok ? row + (uint32_t)hx1 + 1u : 0u;
    }
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        rs[rr] = __ldg(P.bin_start + i0[rr]);
        re[rr] = __ldg(P.bin_start + i1[rr]);
    }
    return (re[0] - rs[0]) + (re[1] - rs[1]) + (re[2] - rs[2]) + (re[3] - rs[3]);
}

// Pass 1: classify every shifted cell of this direction and list the occupied ones, so that the
// splat kernel below runs on full warps. Class "plain": both corner nodes of every axis lie inside
// the face grid and in the same 10^3 block -- one block frame, one membership test per particle.
// Class "seam": everything else (block seams, grid border). A CTA covers a region of 1024
// consecutive cells and appends [its plain cells][its seam cells], each group padded to a multiple
// of 32 entries, with one atomic: every warp of the splat kernel sees one class only, and the
// plain and seam cells of a region -- whose particles share 32-byte sectors -- are processed at
// the same time by neighbouring warps. The list order has no influence on the result (every cell
// owns its slot of `partial`).
constexpr int kListThreads = 256;
constexpr uint32_t kEmptyCell = 0xffffffffu;              // padding entry of the cell list
constexpr int kListRounds = 4;                             // cells per thread: a CTA lists a region of 1024 cells
template <int DIR>
__global__ void __launch_bounds__(kListThreads) k_p2g_cell_list(const __grid_constant__ P2GParams P) {
    constexpr int kWarps = kListThreads / 32;
    __shared__ uint32_t warp_cnt[2][kListRounds][kWarps];
    __shared__ uint32_t cta_cnt[2];
    __shared__ uint32_t cta_base;
    const uint32_t nxy = (uint32_t)P.ccx * (uint32_t)P.ccy;
    const int iz = (int)blockIdx.y;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, below = (1u << lane) - 1u;
    int cls[kListRounds];                                     // 0 empty, 1 plain, 2 seam
    uint4 ra[kListRounds], rb[kListRounds];                   // the CellRec of this thread's cell of round r
    uint32_t rank[kListRounds];                               // rank inside its warp's class group
#pragma unroll
    for (int r = 0; r < kListRounds; r++) {
        const uint32_t xy = (blockIdx.x * (uint32_t)kListRounds + r) * (uint32_t)kListThreads + threadIdx.x;
        cls[r] = 0;
        ra[r] = make_uint4(0u, 0u, 0u, 0u);
        rb[r] = make_uint4(0u, 0u, 0u, 0u);
        if (xy < nxy) {
            const int iy = (int)(xy / (uint32_t)P.ccx), ix = (int)(xy - (uint32_t)iy * (uint32_t)P.ccx);
            const uint32_t t = xy + nxy * (uint32_t)iz;
            const int b[3] = {ix - 1, iy - 1, iz + P.ck0};
            // a cell can only hold particles if its 10^3 block is in the (dilated) particle block mask
            const int kb = min(max(b[2], 0) / kChunk, P.bk - 1);
            const bool maybe = P.active[max(b[0], 0) / kChunk + P.bi * (max(b[1], 0) / kChunk + P.bj * kb)] != 0;
            uint32_t rs[4] = {0u, 0u, 0u, 0u}, re[4] = {0u, 0u, 0u, 0u};
            const uint32_t total = maybe ? cell_runs<DIR>(P, b, rs, re) : 0u;
            const uint32_t n0 = re[0] - rs[0], n1 = re[1] - rs[1], n2 = re[2] - rs[2], n3 = re[3] - rs[3];
            uint32_t n01 = n0 | (n1 << 16);
            if ((n0 | n1 | n2 | n3) > 0xffffu) n01 = 0xffffffffu;
            const int dims[3] = {P.gi, P.gj, P.gk};
            bool plain = true;
#pragma unroll
            for (int a = 0; a < 3; a++) plain = plain && b[a] >= 0 && b[a] + 1 < dims[a] && (b[a] % kChunk) != kChunk - 1;
            cls[r] = total == 0 ? 0 : (plain ? 1 : 2);
            P.cell_flag[t] = total != 0;
            ra[r] = make_uint4(t, rs[0], rs[1], rs[2]);
            rb[r] = make_uint4(rs[3], n01, n2 | (n3 << 16), plain ? 0u : 1u);
        }
        const unsigned mp = __ballot_sync(0xffffffffu, cls[r] == 1), ms = __ballot_sync(0xffffffffu, cls[r] == 2);
        if (lane == 0) { warp_cnt[0][r][warp] = __popc(mp); warp_cnt[1][r][warp] = __popc(ms); }
        rank[r] = __popc((cls[r] == 2 ? ms : mp) & below);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t np = 0, ns = 0;
        for (int r = 0; r < kListRounds; r++)
            for (int w = 0; w < kWarps; w++) { np += warp_cnt[0][r][w]; ns += warp_cnt[1][r][w]; }
        cta_cnt[0] = np; cta_cnt[1] = ns;
        const uint32_t padded = ((np + 31u) & ~31u) + ((ns + 31u) & ~31u);
        cta_base = padded ? atomicAdd(P.list_count, padded) : 0u;
    }
    __syncthreads();
    // region layout in the list: [plain cells, padded to a warp][seam cells, padded to a warp]
    const uint32_t np = cta_cnt[0], ns = cta_cnt[1], np32 = (np + 31u) & ~31u, ns32 = (ns + 31u) & ~31u;
    const uint32_t seam_base = cta_base + np32;
#pragma unroll
    for (int r = 0; r < kListRounds; r++) {
        if (!cls[r]) continue;
        const int k = cls[r] - 1;
        uint32_t pos = (k ? seam_base : cta_base) + rank[r];
        for (int rr = 0; rr <= r; rr++)
            for (unsigned w = 0; w < (rr == r ? warp : (unsigned)kWarps); w++) pos += warp_cnt[k][rr][w];
        uint4 *dst = reinterpret_cast<uint4 *>(P.cell_list + pos);
        dst[0] = ra[r];
        dst[1] = rb[r];
    }
    if (threadIdx.x < np32 - np)
        reinterpret_cast<uint4 *>(P.cell_list + cta_base + np + threadIdx.x)[0] = make_uint4(kEmptyCell, 0u, 0u, 0u);
    if (threadIdx.x < ns32 - ns)
        reinterpret_cast<uint4 *>(P.cell_list + seam_base + ns + threadIdx.x)[0] = make_uint4(kEmptyCell, 0u, 0u, 0u);
}

// APIC "edge" particle (within a few ulps of a cell plane) of shifted cell b: the reference's own
// arithmetic, corner by corner. Out of line: it is rare, and it would otherwise dominate the
// register budget of the splat loop. Returns the mask of corners that received a contribution.
template <int DIR, int METHOD>
__device__ __noinline__ unsigned edge_cell_contrib(const P2GParams &P, uint32_t q, int bx, int by, int bz, float2 *out) {
    const AxisNodes X = axis_nodes(bx, P.gi, P.chunk, P.g.dx), Y = axis_nodes(by, P.gj, P.chunk, P.g.dx),
                    Z = axis_nodes(bz, P.gk, P.chunk, P.g.dx);
    const uint32_t word = __ldg(P.seam + q);
    const int lx = (int)(word & 255u) - 1, sx = (int)((word >> 8) & 3u);
    const int ly = (int)((word >> 10) & 255u) - 1, sy = (int)((word >> 18) & 3u);
    const int lz = (int)((word >> 20) & 255u) - 1, sz = (int)((word >> 28) & 3u);
    unsigned mask = 0;
    for (int c = 0; c < 8; c++) {
        const int cx = c & 1, cy = (c >> 1) & 1, cz = c >> 2;
        out[c] = make_float2(0.0f, 0.0f);
        if (X.nb[cx] < 0 || Y.nb[cy] < 0 || Z.nb[cz] < 0) continue;
        if ((unsigned)(X.nb[cx] - lx) > (unsigned)sx || (unsigned)(Y.nb[cy] - ly) > (unsigned)sy ||
            (unsigned)(Z.nb[cz] - lz) > (unsigned)sz)
            continue;
        FaceFrame f;
        f.nb[0] = X.nb[cx]; f.nb[1] = Y.nb[cy]; f.nb[2] = Z.nb[cz];
        f.lo[0] = X.lo[cx]; f.lo[1] = Y.lo[cy]; f.lo[2] = Z.lo[cz];
        f.bpos[0] = X.bpos[cx]; f.bpos[1] = Y.bpos[cy]; f.bpos[2] = Z.bpos[cz];
        f.gpos[0] = X.gpos[cx]; f.gpos[1] = Y.gpos[cy]; f.gpos[2] = Z.gpos[cz];
        float w, wv;
        if (exact_contribution<DIR, METHOD>(P, f, q, w, wv)) {
            out[c] = make_float2(w, wv);
            mask |= 1u << c;
        }
    }
    return mask;
}

// Particle staging of the splat kernel: every thread copies the records of (up to) kStage of ITS
// cell's particles into its private shared-memory column with 4-byte cp.async, all at once, and
// only then starts the arithmetic. A cell-per-thread walk has no other memory parallelism: with
// plain loads a thread has one particle in flight. No barrier is needed -- a thread only reads
// what it copied itself (cp.async.wait_group makes that visible to it).
constexpr int kCellThreads = 128;
#ifndef FFB_CELLS_STAGE
#define FFB_CELLS_STAGE 4
#endif
constexpr int kStage = FFB_CELLS_STAGE;

template <int METHOD>
struct CellStage {
    float4 a[kStage][kCellThreads];                                           // px, py, pz, vel
    float4 b[METHOD == FFB200_TRANSFER_APIC ? kStage : 1][kCellThreads];      // affine row, seam word (APIC)
    uint32_t w[METHOD == FFB200_TRANSFER_APIC ? 1 : kStage][kCellThreads];    // seam word (FLIP)
};

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Pass 2, one listed cell per thread: splat the cell's particles onto its 8 corner nodes.
template <int DIR, int METHOD, bool SEAM>
__device__ __forceinline__ void splat_cell(const P2GParams &P, const uint4 r0, const uint4 r1, CellStage<METHOD> &S) {
    const int tid = threadIdx.x;
    const uint32_t t = r0.x;
    const uint32_t nxy = (uint32_t)P.ccx * (uint32_t)P.ccy;
    const int iz = (int)(t / nxy);
    const uint32_t xy = t - (uint32_t)iz * nxy;
    const int iy = (int)(xy / (uint32_t)P.ccx), ix = (int)(xy - (uint32_t)iy * (uint32_t)P.ccx);
    const int b[3] = {ix - 1, iy - 1, iz + P.ck0};
    uint32_t rs[4] = {r0.y, r0.z, r0.w, r1.x};
    uint32_t n[4] = {r1.y & 0xffffu, r1.y >> 16, r1.z & 0xffffu, r1.z >> 16};
    if (r1.y == 0xffffffffu) {                                 // run lengths above 65535: re-read the bin table
        uint32_t re[4];
        cell_runs<DIR>(P, b, rs, re);
#pragma unroll
        for (int rr = 0; rr < 4; rr++) n[rr] = re[rr] - rs[rr];
    }
    // flattened walk over the four runs: slot of the i-th particle = i + o(i)
    const uint32_t c1 = n[0], c2 = c1 + n[1], c3 = c2 + n[2], total = c3 + n[3];
    const uint32_t o0 = rs[0], o1 = rs[1] - c1, o2 = rs[2] - c2, o3 = rs[3] - c3;
#define FFB_SLOT(i) ((i) + ((i) < c1 ? o0 : ((i) < c2 ? o1 : ((i) < c3 ? o2 : o3))))
#define FFB_STAGE_FILL(i0)                                                                               \
    {                                                                                                    \
        _Pragma("unroll") for (int k = 0; k < kStage; k++) {                                            \
            const uint32_t i = (i0) + k;                                                                 \
            if (i < total) {                                                                             \
                const uint32_t q = FFB_SLOT(i);                                                          \
                cp_async4(&S.a[k][tid].x, P.px + q);                                                     \
                cp_async4(&S.a[k][tid].y, P.py + q);                                                     \
                cp_async4(&S.a[k][tid].z, P.pz + q);                                                     \
                cp_async4(&S.a[k][tid].w, P.vel + q);                                                    \
                if (METHOD == FFB200_TRANSFER_APIC) {                                                    \
                    cp_async4(&S.b[k][tid].x, P.ax + q);                                                 \
                    cp_async4(&S.b[k][tid].y, P.ay + q);                                                 \
                    cp_async4(&S.b[k][tid].z, P.az + q);                                                 \
                    cp_async4(&S.b[k][tid].w, P.seam + q);                                               \
                } else {                                                                                 \
                    cp_async4(&S.w[k][tid], P.seam + q);                                                 \
                }                                                                                        \
            }                                                                                            \
        }                                                                                                \
        cp_async_wait_all();                                                                             \
    }

    FFB_STAGE_FILL(0u);

    float aw[8], awv[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { aw[c] = 0.0f; awv[c] = 0.0f; }

    if (!SEAM) {
        // ---- plain cell: one block frame ---------------------------------------------------------------
        const int nbx = b[0] / kChunk, nby = b[1] / kChunk, nbz = b[2] / kChunk;
        const float bpx = idx2posf(nbx, P.chunk), bpy = idx2posf(nby, P.chunk), bpz = idx2posf(nbz, P.chunk);
        const int lox = b[0] - nbx * kChunk, loy = b[1] - nby * kChunk, loz = b[2] - nbz * kChunk;
        const float gx[2] = {idx2posf(lox, P.g.dx), idx2posf(lox + 1, P.g.dx)};
        const float gy[2] = {idx2posf(loy, P.g.dx), idx2posf(loy + 1, P.g.dx)};
        const float gz[2] = {idx2posf(loz, P.g.dx), idx2posf(loz + 1, P.g.dx)};
        for (uint32_t i0 = 0; i0 < total; i0 += kStage) {
            if (i0) FFB_STAGE_FILL(i0);
            const int cnt = (int)min(total - i0, (uint32_t)kStage);
            for (int k = 0; k < cnt; k++) {
                const float4 pa = S.a[k][tid];
                float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t word;
                if (METHOD == FFB200_TRANSFER_APIC) { pb = S.b[k][tid]; word = __float_as_uint(pb.w); } else word = S.w[k][tid];
                const int lx = (int)(word & 255u) - 1, sx = (int)((word >> 8) & 3u);
                const int ly = (int)((word >> 10) & 255u) - 1, sy = (int)((word >> 18) & 3u);
                const int lz = (int)((word >> 20) & 255u) - 1, sz = (int)((word >> 28) & 3u);
                const bool member = (unsigned)(nbx - lx) <= (unsigned)sx && (unsigned)(nby - ly) <= (unsigned)sy &&
                                    (unsigned)(nbz - lz) <= (unsigned)sz;
                // APIC edge particles (near a cell plane) are added by k_p2g_edge with the exact arithmetic
                if (member && !(METHOD == FFB200_TRANSFER_APIC && (word & kEdgeBit))) {
                    const float xl = (pa.x - P.off[0]) - bpx, yl = (pa.y - P.off[1]) - bpy, zl = (pa.z - P.off[2]) - bpz;
                    const float vx[2] = {gx[0] - xl, gx[1] - xl};
                    const float vy[2] = {gy[0] - yl, gy[1] - yl};
                    const float vz[2] = {gz[0] - zl, gz[1] - zl};
                    if (METHOD == FFB200_TRANSFER_FLIP) {
                        const float xx[2] = {vx[0] * vx[0], vx[1] * vx[1]};
                        const float yy[2] = {vy[0] * vy[0], vy[1] * vy[1]};
                        const float zz[2] = {vz[0] * vz[0], vz[1] * vz[1]};
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const int cx = c & 1, cy = (c >> 1) & 1, cz = c >> 2;
                            const float d2 = xx[cx] + yy[cy] + zz[cz];
                            if (d2 < P.rsq) {
                                const float w = 1.0f - P.c1 * d2 * d2 * d2 + P.c2 * d2 * d2 - P.c3 * d2;
                                awv[c] += w * pa.w;
                                aw[c] += w;
                            }
                        }
                    } else {
                        // ipos = (xl - gpos(cell)) * inv_s (:567-592); xl - g0 == -(g0 - xl) exactly
                        const float tx = (-vx[0]) * P.inv_s, ty = (-vy[0]) * P.inv_s, tz = (-vz[0]) * P.inv_s;
                        const float fx[2] = {1.0f - tx, tx}, fy[2] = {1.0f - ty, ty}, fz[2] = {1.0f - tz, tz};
                        const float ax[2] = {pb.x * vx[0], pb.x * vx[1]};
                        const float ay[2] = {pb.y * vy[0], pb.y * vy[1]};
                        const float az[2] = {pb.z * vz[0], pb.z * vz[1]};
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const int cx = c & 1, cy = (c >> 1) & 1, cz = c >> 2;
                            const float w = fx[cx] * fy[cy] * fz[cz];
                            const float apic = ax[cx] + ay[cy] + az[cz];
                            awv[c] += w * (pa.w + apic);
                            aw[c] += w;
                        }
                    }
                }
            }
        }
    } else {
        // ---- seam / border cell: node 0 and node 1 of an axis may sit in different blocks ---------------
        const AxisNodes X = axis_nodes(b[0], P.gi, P.chunk, P.g.dx), Y = axis_nodes(b[1], P.gj, P.chunk, P.g.dx),
                        Z = axis_nodes(b[2], P.gk, P.chunk, P.g.dx);
        for (uint32_t i0 = 0; i0 < total; i0 += kStage) {
            if (i0) FFB_STAGE_FILL(i0);
            const int cnt = (int)min(total - i0, (uint32_t)kStage);
            for (int k = 0; k < cnt; k++) {
                const float4 pa = S.a[k][tid];
                float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t word;
                if (METHOD == FFB200_TRANSFER_APIC) { pb = S.b[k][tid]; word = __float_as_uint(pb.w); } else word = S.w[k][tid];
                if (!(METHOD == FFB200_TRANSFER_APIC && (word & kEdgeBit))) {
                    // membership of the particle in the block of node 0 / node 1, per axis
                    const int lx = (int)(word & 255u) - 1, sx = (int)((word >> 8) & 3u);
                    const int ly = (int)((word >> 10) & 255u) - 1, sy = (int)((word >> 18) & 3u);
                    const int lz = (int)((word >> 20) & 255u) - 1, sz = (int)((word >> 28) & 3u);
                    const bool mx[2] = {X.nb[0] >= 0 && (unsigned)(X.nb[0] - lx) <= (unsigned)sx,
                                        X.nb[1] >= 0 && (unsigned)(X.nb[1] - lx) <= (unsigned)sx};
                    const bool my[2] = {Y.nb[0] >= 0 && (unsigned)(Y.nb[0] - ly) <= (unsigned)sy,
                                        Y.nb[1] >= 0 && (unsigned)(Y.nb[1] - ly) <= (unsigned)sy};
                    const bool mz[2] = {Z.nb[0] >= 0 && (unsigned)(Z.nb[0] - lz) <= (unsigned)sz,
                                        Z.nb[1] >= 0 && (unsigned)(Z.nb[1] - lz) <= (unsigned)sz};
                    const float xs = pa.x - P.off[0], ys = pa.y - P.off[1], zs = pa.z - P.off[2];
                    // block-local coordinates in the frame of node 0's and node 1's block
                    const float xl[2] = {xs - X.bpos[0], xs - X.bpos[1]};
                    const float yl[2] = {ys - Y.bpos[0], ys - Y.bpos[1]};
                    const float zl[2] = {zs - Z.bpos[0], zs - Z.bpos[1]};
                    const float vx[2] = {X.gpos[0] - xl[0], X.gpos[1] - xl[1]};
                    const float vy[2] = {Y.gpos[0] - yl[0], Y.gpos[1] - yl[1]};
                    const float vz[2] = {Z.gpos[0] - zl[0], Z.gpos[1] - zl[1]};
                    if (METHOD == FFB200_TRANSFER_FLIP) {
                        const float xx[2] = {vx[0] * vx[0], vx[1] * vx[1]};
                        const float yy[2] = {vy[0] * vy[0], vy[1] * vy[1]};
                        const float zz[2] = {vz[0] * vz[0], vz[1] * vz[1]};
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const int cx = c & 1, cy = (c >> 1) & 1, cz = c >> 2;
                            const float d2 = xx[cx] + yy[cy] + zz[cz];
                            if (mx[cx] && my[cy] && mz[cz] && d2 < P.rsq) {
                                const float w = 1.0f - P.c1 * d2 * d2 * d2 + P.c2 * d2 * d2 - P.c3 * d2;
                                awv[c] += w * pa.w;
                                aw[c] += w;
                            }
                        }
                    } else {
                        // node 0: the particle is in node 0's cell, factor 1 - ipos; node 1: it is in the cell
                        // below node 1, factor ipos measured from the node before node 1 (:567-592)
                        const float fx[2] = {1.0f - (xl[0] - X.gpos[0]) * P.inv_s, (xl[1] - X.gposm1) * P.inv_s};
                        const float fy[2] = {1.0f - (yl[0] - Y.gpos[0]) * P.inv_s, (yl[1] - Y.gposm1) * P.inv_s};
                        const float fz[2] = {1.0f - (zl[0] - Z.gpos[0]) * P.inv_s, (zl[1] - Z.gposm1) * P.inv_s};
                        const float ax[2] = {pb.x * vx[0], pb.x * vx[1]};
                        const float ay[2] = {pb.y * vy[0], pb.y * vy[1]};
                        const float az[2] = {pb.z * vz[0], pb.z * vz[1]};
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const int cx = c & 1, cy = (c >> 1) & 1, cz = c >> 2;
                            if (mx[cx] && my[cy] && mz[cz]) {
                                const float w = fx[cx] * fy[cy] * fz[cz];
                                const float apic = ax[cx] + ay[cy] + az[cz];
                                awv[c] += w * (pa.w + apic);
                                aw[c] += w;
                            }
                        }
                    }
                }
            }
        }
    }
#undef FFB_STAGE_FILL
#undef FFB_SLOT
    // corner-major layout: the faces of one x-row read consecutive cells of one corner plane
    const size_t plane = (size_t)P.ccx * P.ccy * P.ccz;
#pragma unroll
    for (int c = 0; c < 8; c++) P.partial[(size_t)c * plane + t] = make_float2(aw[c], awv[c]);
}

#ifndef FFB_CELLS_MINB
#define FFB_CELLS_MINB 6
#endif
template <int DIR, int METHOD>
__global__ void __launch_bounds__(kCellThreads, FFB_CELLS_MINB) k_p2g_cells(const __grid_constant__ P2GParams P) {
    __shared__ CellStage<METHOD> S;
    const uint32_t stride = gridDim.x * (uint32_t)kCellThreads;       // a multiple of 32: warps stay aligned to the list
    const uint32_t count = __ldg(P.list_count);
    for (uint32_t e = blockIdx.x * (uint32_t)kCellThreads + threadIdx.x; e < count; e += stride) {
        const uint4 *rec = reinterpret_cast<const uint4 *>(P.cell_list + e);
        const uint4 r0 = __ldg(rec);
        if (r0.x == kEmptyCell) continue;                      // padding
        const uint4 r1 = __ldg(rec + 1);
        if (r1.w)                                              // uniform per warp by construction of the list
            splat_cell<DIR, METHOD, true>(P, r0, r1, S);
        else
            splat_cell<DIR, METHOD, false>(P, r0, r1, S);
    }
}

// APIC edge particles, one thread each. (1) An edge particle whose exact cell (the reference's
// double floor in some block frame) differs from its bin cell also reaches nodes outside the 8
// corners of its cell: list those contributions (almost always none). (2) Its contributions to its
// own cell's corners, see below.
template <int DIR, int METHOD>
__global__ void k_p2g_edge(const __grid_constant__ P2GParams P) {
    const uint32_t nedge = min(__ldg(P.edge_count), P.edge_cap);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nedge) return;
    const uint32_t q = __ldg(P.edge_list + t);
    const float p[3] = {__ldg(P.px + q), __ldg(P.py + q), __ldg(P.pz + q)};
    const int dims[3] = {P.gi, P.gj, P.gk};
    const int H[3] = {P.g.HX, P.g.HY, P.g.HZ};
    int b[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int h = __double2int_rd((double)p[a] * P.g.inv_2dx);
        b[a] = (h - (a == DIR ? 0 : 1)) >> 1;                 // the shifted cell whose thread owns this particle
    }
    for (int dz = -1; dz <= 2; dz++)
        for (int dy = -1; dy <= 2; dy++)
            for (int dx = -1; dx <= 2; dx++) {
                if ((unsigned)dx <= 1u && (unsigned)dy <= 1u && (unsigned)dz <= 1u) continue;   // the cell's own corners
                const int n[3] = {b[0] + dx, b[1] + dy, b[2] + dz};
                if (n[0] < 0 || n[1] < 0 || n[2] < 0 || n[0] >= dims[0] || n[1] >= dims[1] || n[2] >= dims[2]) continue;
                if (n[2] < P.kw0 || n[2] >= P.kw1) continue;
                FaceFrame f;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    f.nb[a] = n[a] / kChunk;
                    f.lo[a] = n[a] - f.nb[a] * kChunk;
                    f.bpos[a] = idx2posf(f.nb[a], P.chunk);
                    f.gpos[a] = idx2posf(f.lo[a], P.g.dx);
                    f.gposm[a] = idx2posf(f.lo[a] - 1, P.g.dx);
                    f.h0[a] = 0; f.h1[a] = H[a] - 1;
                }
                float w, wv;
                if (exact_contribution<DIR, METHOD>(P, f, q, w, wv)) {
                    const int slot = atomicAdd(P.ovf_count, 1);
                    if (slot < kOverflowCap) {
                        OverflowEntry e;
                        e.node = (uint32_t)((size_t)n[0] + (size_t)P.gi * ((size_t)n[1] + (size_t)P.gj * (n[2] - P.g.kbase)));
                        e.q = q; e.w = w; e.wv = wv;
                        P.ovf[slot] = e;
                    }
                }
            }
    // ---- own cell: k_p2g_cells left the edge particles out. The thread of the cell's first edge
    // particle (in the splat's walk order) adds all of them to the cell's 8 partial sums, with the
    // reference's arithmetic; one thread per cell, after the splat: no race, fixed order.
    const int ix = b[0] + 1, iy = b[1] + 1, iz = b[2] - P.ck0;
    if ((unsigned)ix >= (unsigned)P.ccx || (unsigned)iy >= (unsigned)P.ccy || (unsigned)iz >= (unsigned)P.ccz) return;
    uint32_t rs[4], re[4];
    cell_runs<DIR>(P, b, rs, re);
    uint32_t firstq = 0xffffffffu;
    for (int rr = 0; rr < 4 && firstq == 0xffffffffu; rr++)
        for (uint32_t qq = rs[rr]; qq < re[rr]; qq++)
            if (__ldg(P.seam + qq) & kEdgeBit) { firstq = qq; break; }
    if (firstq != q) return;
    const size_t cell = (size_t)ix + (size_t)P.ccx * ((size_t)iy + (size_t)P.ccy * iz);
    const size_t plane = (size_t)P.ccx * P.ccy * P.ccz;
    float2 acc[8];
    for (int c = 0; c < 8; c++) acc[c] = P.partial[(size_t)c * plane + cell];
    for (int rr = 0; rr < 4; rr++)
        for (uint32_t qq = rs[rr]; qq < re[rr]; qq++) {
            if (!(__ldg(P.seam + qq) & kEdgeBit)) continue;
            float2 tmp[8];
            const unsigned mask = edge_cell_contrib<DIR, METHOD>(P, qq, b[0], b[1], b[2], tmp);
            for (int c = 0; c < 8; c++)
                if (mask & (1u << c)) { acc[c].y += tmp[c].y; acc[c].x += tmp[c].x; }
        }
    for (int c = 0; c < 8; c++) P.partial[(size_t)c * plane + cell] = acc[c];
}

#ifndef FFB_NODES_BY
#define FFB_NODES_BY 4
#endif
template <int DIR, int METHOD>
__global__ void __launch_bounds__(32 * FFB_NODES_BY) k_p2g_nodes(const __grid_constant__ P2GParams P) {
    const int ni = blockIdx.x * blockDim.x + threadIdx.x;
    const int nj = blockIdx.y * blockDim.y + threadIdx.y;
    const int nk = blockIdx.z * blockDim.z + threadIdx.z + P.kw0;
    if (ni >= P.gi || nj >= P.gj || nk >= P.kw1) return;
    // face and cell counts fit 32 bits (checked by the launcher): 32-bit index arithmetic
    const uint32_t fidx = (uint32_t)ni + (uint32_t)P.gi * ((uint32_t)nj + (uint32_t)P.gj * (uint32_t)(nk - P.g.kbase));
    const int n[3] = {ni, nj, nk};
    const int nbv[3] = {ni / kChunk, nj / kChunk, nk / kChunk};
    if (!P.active[nbv[0] + P.bi * (nbv[1] + P.bj * nbv[2])]) {
        P.out[fidx] = 0.0f;
        P.wsum[fidx] = 0.0f;
        P.valid[fidx] = 0;
        return;
    }
    float sw = 0.0f, swv = 0.0f;
    const size_t plane = (size_t)P.ccx * P.ccy * P.ccz;
    // node n is corner c = (cx, cy, cz) of the shifted cell n - (cx, cy, cz). Flags, then partial
    // sums, as two batches of independent loads; the additions keep the fixed corner order.
    const uint32_t sy = (uint32_t)P.ccx, sz = (uint32_t)P.ccx * (uint32_t)P.ccy;
    const int iz1 = nk - P.ck0;                                // cell layer of the corners with cz = 0
    const uint32_t base = (uint32_t)(ni + 1) + sy * (uint32_t)(nj + 1) + sz * (uint32_t)max(iz1, 0);
    const bool zok[2] = {iz1 >= 0 && iz1 < P.ccz, iz1 - 1 >= 0 && iz1 - 1 < P.ccz};   // cells outside this rank's range hold nothing
    uint32_t cell[8];
    bool have[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int cx = c & 1, cy = (c >> 1) & 1, cz = c >> 2;
        have[c] = zok[cz];
        cell[c] = have[c] ? base - (uint32_t)cx - (cy ? sy : 0u) - (cz && iz1 > 0 ? sz : 0u) : 0u;
    }
#pragma unroll
    for (int c = 0; c < 8; c++) have[c] = have[c] && __ldg(P.cell_flag + cell[c]) != 0;
    float2 pr[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        pr[c] = make_float2(0.0f, 0.0f);
        if (have[c]) pr[c] = __ldg(P.partial + (size_t)c * plane + cell[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; c++)
        if (have[c]) {
            sw += pr[c].x;
            swv += pr[c].y;
        }
    const int novf = *P.ovf_count;
    bool redo = novf > kOverflowCap || __ldg(P.edge_count) > P.edge_cap;
    if (!redo && novf > 0) {                                   // rare: add the matching entries in ascending slot order
        long long last = -1;
        for (;;) {
            long long best = 0x7fffffffffffffffLL;
            int bi = -1;
            for (int i = 0; i < novf; i++) {
                const OverflowEntry e = P.ovf[i];
                if (e.node == (uint32_t)fidx && (long long)e.q > last && (long long)e.q < best) { best = e.q; bi = i; }
            }
            if (bi < 0) break;
            last = best;
            sw += P.ovf[bi].w;
            swv += P.ovf[bi].wv;
        }
    }
    const float eps = 1e-6f;
    if (redo || fabsf(sw - eps) <= P.guard_abs + P.guard_per * 512.0f) {
        FaceFrame f;
        const int H[3] = {P.g.HX, P.g.HY, P.g.HZ};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            f.nb[a] = nbv[a];
            f.lo[a] = n[a] - nbv[a] * kChunk;
            f.bpos[a] = idx2posf(nbv[a], P.chunk);
            f.gpos[a] = idx2posf(f.lo[a], P.g.dx);
            f.gposm[a] = idx2posf(f.lo[a] - 1, P.g.dx);
            const int c = 2 * n[a] + (a == DIR ? 0 : 1) + kApron - (a == 2 ? 2 * P.g.kbase : 0);
            f.h0[a] = max(c - P.wm, 0);
            f.h1[a] = min(c + P.wm - 1, H[a] - 1);
        }
        exact_face<DIR, METHOD>(P, f, sw, swv);
    }
    float s = swv;
    if (sw > eps) s /= sw;                                     // :527-531
    P.out[fidx] = s;                                           // write-out :155-162
    P.wsum[fidx] = sw;
    P.valid[fidx] = sw > eps ? 1 : 0;
}

template <int DIR, int METHOD>
int launch_cells(Context &c, P2GParams &P, cudaStream_t st) {
    int launches = 0;
    FFB_CUDA(cudaMemsetAsync(P.ovf_count, 0, sizeof(int), st));
    FFB_CUDA(cudaMemsetAsync(P.list_count, 0, 2 * sizeof(uint32_t), st));
    const long long ncell = (long long)P.ccx * P.ccy * P.ccz;
    const unsigned nxy = (unsigned)P.ccx * (unsigned)P.ccy;
    k_p2g_cell_list<DIR><<<dim3((nxy + kListThreads * kListRounds - 1) / (kListThreads * kListRounds), P.ccz), kListThreads, 0, st>>>(P);
    launches++;
    // grid-stride over the device-side list: enough CTAs to cover every cell once, capped at a few waves
    const long long want = (ncell + kCellThreads - 1) / kCellThreads;
    const unsigned ctas = (unsigned)std::min<long long>(want, (long long)c.sm_count * FFB_CELLS_MINB * 8);
    k_p2g_cells<DIR, METHOD><<<ctas, kCellThreads, 0, st>>>(P);
    launches++;
    if (METHOD == FFB200_TRANSFER_APIC) {
        k_p2g_edge<DIR, METHOD><<<(P.edge_cap + 127) / 128, 128, 0, st>>>(P);
        launches++;
    }
    dim3 block(32, FFB_NODES_BY, 1);
    dim3 grid((P.gi + 31) / 32, (P.gj + FFB_NODES_BY - 1) / FFB_NODES_BY, P.kw1 - P.kw0);
    k_p2g_nodes<DIR, METHOD><<<grid, block, 0, st>>>(P);
    launches++;
    return launches;
}

template <int DIR, int METHOD>
void launch_brick(Context &c, P2GParams &P) {
    constexpr int CAP = BrickCap<METHOD>::value;
    const size_t smem = sizeof(float4) * CAP * (METHOD == FFB200_TRANSFER_APIC ? 2 : 1) + sizeof(BrickShared);
    static bool configured = false;
    if (!configured) {
        FFB_CUDA(cudaFuncSetAttribute(k_p2g_brick<DIR, METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int zb0 = P.g.kbase / kChunk, zb1 = (P.g.kbase + P.kstore - 1) / kChunk;
    dim3 grid(P.bi, P.bj * 2, (zb1 - zb0 + 1) * 2);
    k_p2g_brick<DIR, METHOD><<<grid, kBrickThreads, smem, c.stream>>>(P);
}

template <int DIR>
int launch_dir(Context &c, P2GParams &P, int method, int variant, cudaStream_t st) {
    // variant 0: cell-partial splat (support of one cell: default radius and APIC); 1: brick gather (any
    // radius up to 2 dx); 2: whole-grid gather (also the attribute transfer's kernel). FFB200_P2G_VARIANT
    // overrides; radii above dx always take the brick gather.
    if (variant == 0 && P.wm == 2) {
        if (method == FFB200_TRANSFER_APIC) return launch_cells<DIR, FFB200_TRANSFER_APIC>(c, P, st);
        return launch_cells<DIR, FFB200_TRANSFER_FLIP>(c, P, st);
    }
    if (variant <= 1) {
        if (method == FFB200_TRANSFER_APIC)
            launch_brick<DIR, FFB200_TRANSFER_APIC>(c, P);
        else
            launch_brick<DIR, FFB200_TRANSFER_FLIP>(c, P);
        return 1;
    }
    dim3 block(32, 4, 2);
    dim3 grid((P.gi + block.x - 1) / block.x, (P.gj + block.y - 1) / block.y, (P.kstore + block.z - 1) / block.z);
    if (method == FFB200_TRANSFER_APIC)
        k_p2g<DIR, FFB200_TRANSFER_APIC><<<grid, block, 0, c.stream>>>(P);
    else
        k_p2g<DIR, FFB200_TRANSFER_FLIP><<<grid, block, 0, c.stream>>>(P);
    return 1;
}

}  // namespace

void p2g_seam_begin(Context &c, double radius, SeamParams &sp) {
    const GridDesc &g = c.g;
    const double chunkdx = g.dx * kChunk;                     // velocityadvector.cpp:53
    const float eps = 1e-6f;
    FFB_CUDA(cudaMemsetAsync(c.sort.edge_count, 0, 4 * sizeof(uint32_t), c.stream));
    sp.g = g;
    for (int d = 0; d < 3; d++) {
        FaceGrid &f = c.face[d];
        FFB_CUDA(cudaMemsetAsync(f.home, 0, (size_t)f.bi * f.bj * f.bk, c.stream));
        sp.bdim[d][0] = f.bi; sp.bdim[d][1] = f.bj; sp.bdim[d][2] = f.bk;
        sp.home[d] = f.home;
    }
    sp.seam = c.sort.seam;
    sp.edge_list = c.sort.edge_list;
    sp.edge_count = c.sort.edge_count;
    sp.edge_cap = c.sort.edge_cap;
    sp.cap = c.cap;
    ParticleSoA &s = c.soa[c.cur];
    sp.px = s.p[0]; sp.py = s.p[1]; sp.pz = s.p[2];
    sp.h = (float)(0.5 * g.dx);
    sp.sr = (float)(radius + (double)eps);                     // float sr = _particleRadius + eps;
    sp.blockdx = (float)chunkdx;
    sp.inv_blockdx = 1.0 / (double)sp.blockdx;
    sp.inv_chunkdx = 1.0 / chunkdx;
    sp.n = c.n;
}

int launch_p2g_prepare(Context &c, double radius, bool seam_done) {
    int launches = 0;
    if (!seam_done) {                                          // particles were already sorted: stand-alone pass
        SeamParams sp;
        p2g_seam_begin(c, radius, sp);
        if (c.n > 0) {
            k_seam_home<<<(c.n + 255) / 256, 256, 0, c.stream>>>(sp);
            launches++;
        }
    }
    for (int d = 0; d < 3; d++) {
        FaceGrid &f = c.face[d];
        int nb = f.bi * f.bj * f.bk;
        k_dilate26<<<(nb + 127) / 128, 128, 0, c.stream>>>(f.home, f.active, f.bi, f.bj, f.bk);
        launches++;
    }
    FFB_CUDA(cudaGetLastError());
    return launches;
}

int launch_p2g(Context &c, double radius, int method, const HostFieldOut *host) {
    int launches = 0;
    const GridDesc &g = c.g;
    ParticleSoA &s = c.soa[c.cur];
    const float eps = 1e-6f;
    const float sr = (float)(radius + (double)eps);            // float sr = _particleRadius + eps;
    // FFB200_P2G_VARIANT: 0 cell-partial splat (default), 1 brick gather, 2 whole-grid gather
    static const int variant = [] { const char *e = std::getenv("FFB200_P2G_VARIANT"); return e ? std::atoi(e) : 0; }();
    static const bool multi_stream = [] { const char *e = std::getenv("FFB200_P2G_STREAMS"); return e ? std::atoi(e) != 0 : true; }();
    static const bool prioritised = [] { const char *e = std::getenv("FFB200_P2G_PRIORITY"); return e ? std::atoi(e) != 0 : true; }();
    auto copy_out = [&](int d, cudaStream_t st) {
        if (!host) return;
        FaceGrid &f = c.face[d];
        const size_t off = (size_t)f.gi * f.gj * g.kbase;
        if (host->vel[d]) FFB_CUDA(cudaMemcpyAsync(host->vel[d] + off, f.vel, f.count * 4, cudaMemcpyDeviceToHost, st));
        if (host->valid[d]) FFB_CUDA(cudaMemcpyAsync(host->valid[d] + off, f.valid, f.count, cudaMemcpyDeviceToHost, st));
    };
    bool forked = false;
    P2GParams deferred;
    for (int d = 0; d < 3; d++) {
        FaceGrid &f = c.face[d];
        P2GParams P;
        P.g = g;
        P.gi = f.gi; P.gj = f.gj; P.gk = f.gk;
        P.kstore = f.kstore;
        // a slab rank only produces its owned planes (+1: the shared w plane); the halo planes
        // are filled by the face-halo exchange
        P.kw0 = std::max(g.kbase, c.k_own_begin);
        P.kw1 = std::min(g.kbase + f.kstore, c.k_own_end + 1);
        P.bi = f.bi; P.bj = f.bj; P.bk = f.bk;
        P.active = f.active;
        P.bin_start = c.sort.bin_start;
        P.px = s.p[0]; P.py = s.p[1]; P.pz = s.p[2];
        P.vel = s.v[d];
        P.ax = s.a[3 * d + 0]; P.ay = s.a[3 * d + 1]; P.az = s.a[3 * d + 2];
        P.seam = c.sort.seam + (size_t)d * c.cap;
        P.edge_list = c.sort.edge_list + (size_t)d * c.sort.edge_cap;
        P.edge_count = c.sort.edge_count + d;
        P.edge_cap = c.sort.edge_cap;
        P.orig = s.orig;
        P.out = f.vel; P.wsum = f.wsum; P.valid = f.valid;
        const float h = (float)(0.5 * g.dx);
        P.off[0] = P.off[1] = P.off[2] = h;
        P.off[d] = 0.0f;
        P.r = (float)radius;                                   // float r = _particleRadius; :472
        P.sr = sr;
        P.rsq = P.r * P.r;
        P.c1 = (4.0f / 9.0f) * (1.0f / (P.r * P.r * P.r * P.r * P.r * P.r));
        P.c2 = (17.0f / 9.0f) * (1.0f / (P.r * P.r * P.r * P.r));
        P.c3 = (22.0f / 9.0f) * (1.0f / (P.r * P.r));
        P.inv_s = (float)(1.0 / (double)(float)g.dx);
        P.chunk = kChunk * g.dx;
        // half-cell window: bins within sr of the face along each axis
        // FLIP: |x_p - x_face| < r  <=>  bins [c - wm, c + wm - 1], wm = ceil(2r/dx);
        // APIC: the trilinear tent spans one cell either side: wm = 2.
        const bool apic = method == FFB200_TRANSFER_APIC;
        P.wm = apic ? 2 : (int)std::floor(2.0 * (double)sr / g.dx + 1e-3) + 1;
        if (P.wm > kApron) throw CudaError("ffb200_p2g: particle radius above 2*dx is not supported");
        P.guard_abs = c.guard_abs >= 0.f ? c.guard_abs : 1e-9f;
        P.guard_per = c.guard_per >= 0.f ? c.guard_per : 1e-12f;
        P.ccx = f.gi + 1; P.ccy = f.gj + 1; P.ccz = P.kw1 - P.kw0 + 1; P.ck0 = P.kw0 - 1;
        P.partial = nullptr; P.cell_flag = nullptr; P.ovf = nullptr; P.ovf_count = nullptr;
        P.cell_list = nullptr; P.list_count = nullptr; P.list_cap = 0;
        cudaStream_t st = c.stream;
        if (variant == 0 && P.wm == 2) {
            SortScratch::CellScratch &cs = c.sort.cell[d];
            const size_t cells = (size_t)P.ccx * P.ccy * P.ccz;
            if (cells > cs.cells) {                            // grow-only scratch
                FFB_CUDA(cudaDeviceSynchronize());
                if (cs.partial) FFB_CUDA(cudaFree(cs.partial));
                if (cs.cell_flag) FFB_CUDA(cudaFree(cs.cell_flag));
                if (cs.cell_list) FFB_CUDA(cudaFree(cs.cell_list));
                cs.cells = cells + cells / 16;
                if (cs.cells >= 0xffffffffull || (unsigned long long)f.count >= 0xffffffffull)
                    throw CudaError("ffb200_p2g: more than 2^32 shifted cells / faces per rank");
                FFB_CUDA(cudaMalloc(&cs.partial, cs.cells * 8 * sizeof(float2)));
                FFB_CUDA(cudaMalloc(&cs.cell_flag, cs.cells));
                // occupied cells + the per-region padding of the two classes (at most 62 entries per 256 cells)
                FFB_CUDA(cudaMalloc(&cs.cell_list, (cs.cells + cs.cells / 4 + 64 * (size_t)(P.ccz + 16)) * sizeof(CellRec)));
            }
            if (!cs.ovf) {
                FFB_CUDA(cudaMalloc(&cs.ovf, (size_t)kOverflowCap * sizeof(OverflowEntry)));
                FFB_CUDA(cudaMalloc(&cs.ovf_count, sizeof(int)));
                FFB_CUDA(cudaMalloc(&cs.list_count, 2 * sizeof(uint32_t)));
            }
            P.partial = reinterpret_cast<float2 *>(cs.partial);
            P.cell_flag = cs.cell_flag;
            P.ovf = reinterpret_cast<OverflowEntry *>(cs.ovf);
            P.ovf_count = cs.ovf_count;
            P.cell_list = reinterpret_cast<CellRec *>(cs.cell_list);
            P.list_count = cs.list_count;
            P.list_cap = (uint32_t)cells;
            // the three directions are independent: 1 and 2 run on auxiliary streams, forked from and
            // joined back into the context stream with events (FFB200_P2G_STREAMS=0: one stream)
            if (d > 0 && multi_stream) {
                if (!cs.stream) {
                    // direction 1 above direction 2 above the context stream: the directions then FINISH one after
                    // the other instead of together, and a finished direction's device-to-host copy (host outputs)
                    // runs under the kernels of the next
                    int least = 0, greatest = 0;
                    FFB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
                    const int prio = prioritised ? std::max(greatest, std::min(least, greatest + (d - 1))) : least;
                    FFB_CUDA(cudaStreamCreateWithPriority(&cs.stream, cudaStreamNonBlocking, prio));
                    FFB_CUDA(cudaEventCreateWithFlags(&cs.done, cudaEventDisableTiming));
                }
                if (!c.sort.fork) FFB_CUDA(cudaEventCreateWithFlags(&c.sort.fork, cudaEventDisableTiming));
                if (!forked) {
                    FFB_CUDA(cudaEventRecord(c.sort.fork, c.stream));
                    forked = true;
                }
                FFB_CUDA(cudaStreamWaitEvent(cs.stream, c.sort.fork, 0));
                st = cs.stream;
            }
        }
        if (d == 0) deferred = P;                              // direction 0 is enqueued last, after the forks
        if (d == 1) launches += launch_dir<1>(c, P, method, variant, st);
        if (d == 2) launches += launch_dir<2>(c, P, method, variant, st);
        if (d > 0) copy_out(d, st);
        if (st != c.stream) FFB_CUDA(cudaEventRecord(c.sort.cell[d].done, st));
    }
    launches += launch_dir<0>(c, deferred, method, variant, c.stream);
    copy_out(0, c.stream);
    if (forked)
        for (int d = 1; d < 3; d++) FFB_CUDA(cudaStreamWaitEvent(c.stream, c.sort.cell[d].done, 0));
    FFB_CUDA(cudaGetLastError());
    return launches;
}

// AttributeToGridTransfer<T>::transfer on the sorted resident particles: the payload channels are the velocity
// streams v[0 .. ncomp), the result goes to d_out (ncomp interleaved floats per cell) and d_valid. Any radius: the
// gather's half-cell window is clamped to the bin grid.
int launch_attribute_p2g(Context &c, double radius, int ncomp, int normalize, float *d_out, uint8_t *d_valid) {
    const GridDesc &g = c.g;
    if (g.kbase != 0 || g.kloc != g.K) throw CudaError("ffb200_attribute_to_grid_transfer: not available on z-slab contexts");
    ParticleSoA &s = c.soa[c.cur];
    FaceGrid &f = c.cell;
    const size_t cells = (size_t)g.I * g.J * g.K;
    if (!f.wsum) {
        f.gi = g.I; f.gj = g.J; f.gk = g.K;
        f.bi = (g.I + kChunk - 1) / kChunk; f.bj = (g.J + kChunk - 1) / kChunk; f.bk = (g.K + kChunk - 1) / kChunk;
        f.kstore = g.K;
        f.count = cells;
        FFB_CUDA(cudaMalloc(&f.wsum, cells * sizeof(float)));
        FFB_CUDA(cudaMalloc(&f.home, (size_t)f.bi * f.bj * f.bk));
        FFB_CUDA(cudaMalloc(&f.active, (size_t)f.bi * f.bj * f.bk));
    }
    if (c.sort.seam_cell_cap < c.cap) {
        if (c.sort.seam_cell) FFB_CUDA(cudaFree(c.sort.seam_cell));
        FFB_CUDA(cudaMalloc(&c.sort.seam_cell, (size_t)c.cap * sizeof(uint32_t)));
        c.sort.seam_cell_cap = c.cap;
    }
    int launches = 0;
    const float eps = 1e-6f;
    const int nb = f.bi * f.bj * f.bk;
    FFB_CUDA(cudaMemsetAsync(f.home, 0, (size_t)nb, c.stream));
    SeamParams sp;
    sp.g = g;
    sp.px = s.p[0]; sp.py = s.p[1]; sp.pz = s.p[2];
    sp.h = (float)(0.5 * g.dx);
    sp.sr = (float)(radius + (double)eps);
    const double chunkdx = g.dx * kChunk;
    sp.blockdx = (float)chunkdx;
    sp.inv_blockdx = 1.0 / (double)sp.blockdx;
    sp.inv_chunkdx = 1.0 / chunkdx;
    sp.n = c.n;
    if (c.n > 0) {
        k_seam_cell<<<(c.n + 255) / 256, 256, 0, c.stream>>>(sp, c.sort.seam_cell, f.home, f.bi, f.bj, f.bk);
        launches++;
    }
    k_dilate26<<<(nb + 127) / 128, 128, 0, c.stream>>>(f.home, f.active, f.bi, f.bj, f.bk);
    launches++;
    P2GParams P;
    P.g = g;
    P.gi = f.gi; P.gj = f.gj; P.gk = f.gk;
    P.kstore = f.kstore;
    P.kw0 = 0; P.kw1 = g.K;
    P.bi = f.bi; P.bj = f.bj; P.bk = f.bk;
    P.active = f.active;
    P.bin_start = c.sort.bin_start;
    P.px = s.p[0]; P.py = s.p[1]; P.pz = s.p[2];
    P.ax = P.ay = P.az = nullptr;
    P.seam = c.sort.seam_cell;
    P.edge_list = nullptr; P.edge_count = nullptr; P.edge_cap = 0;
    P.orig = s.orig;
    P.partial = nullptr; P.cell_flag = nullptr; P.cell_list = nullptr; P.list_count = nullptr; P.list_cap = 0;
    P.ccx = P.ccy = P.ccz = P.ck0 = 0;
    P.ovf = nullptr; P.ovf_count = nullptr;
    P.wsum = f.wsum; P.valid = d_valid;
    const float h = (float)(0.5 * g.dx);
    P.off[0] = P.off[1] = P.off[2] = h;                        // gridOffset (dx/2, dx/2, dx/2), e.g. fluidsimulation.cpp:7033
    P.r = (float)radius;
    P.sr = sp.sr;
    P.rsq = P.r * P.r;
    P.c1 = (4.0f / 9.0f) * (1.0f / (P.r * P.r * P.r * P.r * P.r * P.r));
    P.c2 = (17.0f / 9.0f) * (1.0f / (P.r * P.r * P.r * P.r));
    P.c3 = (22.0f / 9.0f) * (1.0f / (P.r * P.r));
    P.inv_s = (float)(1.0 / (double)(float)g.dx);
    P.chunk = kChunk * g.dx;
    P.wm = (int)std::floor(2.0 * (double)P.sr / g.dx + 1e-3) + 1;
    P.guard_abs = c.guard_abs >= 0.f ? c.guard_abs : 1e-9f;
    P.guard_per = c.guard_per >= 0.f ? c.guard_per : 1e-12f;
    P.out_stride = ncomp;
    P.vec3_norm = ncomp == 3 ? 1 : 0;
    P.no_norm = normalize ? 0 : 1;
    dim3 block(32, 4, 2);
    dim3 grid((P.gi + block.x - 1) / block.x, (P.gj + block.y - 1) / block.y, (P.kstore + block.z - 1) / block.z);
    for (int ch = 0; ch < ncomp; ch++) {
        P.vel = s.v[ch];
        P.out = d_out + ch;
        k_p2g<3, FFB200_TRANSFER_FLIP><<<grid, block, 0, c.stream>>>(P);
        launches++;
    }
    FFB_CUDA(cudaGetLastError());
    return launches;
}

}  // namespace ffb200
