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

#include <cmath>

namespace ffb200 {

namespace {

struct P2GParams {
    GridDesc g;
    int gi, gj, gk;              // global face dims of this direction
    int kstore;                  // stored face planes, first one is g.kbase
    int bi, bj, bk;              // block dims
    const uint8_t *active;
    const uint32_t *bin_start;
    const float *px, *py, *pz;   // sorted positions
    const float *vel;            // sorted velocity component of this direction
    const float *ax, *ay, *az;   // sorted affine row of this direction (APIC)
    const uint32_t *seam;        // membership word of this direction per sorted slot
    const uint32_t *orig;
    float *out, *wsum;
    uint8_t *valid;
    float off[3];                // _getDirectionOffset, velocityadvector.cpp:177-188
    float r, sr, rsq, c1, c2, c3;
    float inv_s;                 // (float)(1.0 / (float)dx): vec3 / _dx of :574
    double chunk;                // _chunkWidth * _dx
    int wm;                      // half-cell window half width
    float guard_abs, guard_per;
};

struct SeamParams {
    GridDesc g;
    int bdim[3][3];              // block dims per direction
    uint8_t *home[3];
    uint32_t *seam;              // [dir*cap + slot]
    int cap;
    const float *px, *py, *pz;
    float h;                     // (float)(0.5*dx)
    float sr;                    // (float)(radius + 1e-6f)
    float blockdx;               // (float)_chunkdx
    double inv_blockdx;          // 1.0 / (double)blockdx
    double inv_chunkdx;          // 1.0 / _chunkdx
    int n;
};

// Per particle and direction: (a) mark the home block exactly as _initializeActiveBlocksThread
// does (double _chunkdx, :240-251); (b) the inclusive block range the particle is sorted into,
// exactly as _computeGridCountDataThread does (float blockdx, float sr, :306-352).
// Word layout per axis a (10 bits at 10a): (lo+1) in 8 bits, (hi-lo) in 2 bits.
__global__ void k_seam_home(SeamParams s) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= s.n) return;
    const float p[3] = {s.px[j], s.py[j], s.pz[j]};
#pragma unroll
    for (int dir = 0; dir < 3; dir++) {
        float x[3];
#pragma unroll
        for (int a = 0; a < 3; a++) x[a] = p[a] - (a == dir ? 0.0f : s.h);
        // (a) home block
        const int hbx = pos2idx(x[0], s.inv_chunkdx), hby = pos2idx(x[1], s.inv_chunkdx), hbz = pos2idx(x[2], s.inv_chunkdx);
        if (in_range3(hbx, hby, hbz, s.bdim[dir][0], s.bdim[dir][1], s.bdim[dir][2]))
            s.home[dir][hbx + s.bdim[dir][0] * (hby + s.bdim[dir][1] * hbz)] = 1;
        // (b) membership range
        int b[3];
        bool simple = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            b[a] = pos2idx(x[a], s.inv_blockdx);
            const float bp = idx2posf(b[a], (double)s.blockdx);
            simple = simple && (x[a] - s.sr > bp) && (x[a] + s.sr < bp + s.blockdx);
        }
        uint32_t word = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            int lo = b[a], hi = b[a];
            if (!simple) {
                lo = pos2idx(x[a] - s.sr, s.inv_blockdx);
                hi = pos2idx(x[a] + s.sr, s.inv_blockdx);
            }
            int span = hi - lo;
            span = span < 0 ? 0 : (span > 3 ? 3 : span);
            int lo1 = lo + 1;
            // out-of-range block indices can never match a face's block: park them at 255
            if (lo1 < 0 || lo1 > 254) { lo1 = 255; span = 0; }
            word |= ((uint32_t)lo1 | ((uint32_t)span << 8)) << (10 * a);
        }
        s.seam[(size_t)dir * s.cap + j] = word;
    }
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
__global__ void __launch_bounds__(256) k_p2g(P2GParams P) {
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
        P.out[fidx] = 0.0f;
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
    if (sw > eps) s /= sw;                                     // :527-531
    P.out[fidx] = s;                                           // write-out :155-162
    P.wsum[fidx] = sw;
    P.valid[fidx] = sw > eps ? 1 : 0;
}

template <int DIR>
void launch_dir(Context &c, P2GParams &P, int method) {
    dim3 block(32, 4, 2);
    dim3 grid((P.gi + block.x - 1) / block.x, (P.gj + block.y - 1) / block.y, (P.kstore + block.z - 1) / block.z);
    if (method == FFB200_TRANSFER_APIC)
        k_p2g<DIR, FFB200_TRANSFER_APIC><<<grid, block, 0, c.stream>>>(P);
    else
        k_p2g<DIR, FFB200_TRANSFER_FLIP><<<grid, block, 0, c.stream>>>(P);
}

}  // namespace

int launch_p2g_prepare(Context &c, double radius) {
    int launches = 0;
    const GridDesc &g = c.g;
    ParticleSoA &s = c.soa[c.cur];
    const double chunkdx = g.dx * kChunk;                     // velocityadvector.cpp:53
    const float eps = 1e-6f;
    const float sr = (float)(radius + (double)eps);            // float sr = _particleRadius + eps;

    // block masks + membership words
    for (int d = 0; d < 3; d++) {
        FaceGrid &f = c.face[d];
        FFB_CUDA(cudaMemsetAsync(f.home, 0, (size_t)f.bi * f.bj * f.bk, c.stream));
    }
    if (c.n > 0) {
        SeamParams sp;
        sp.g = g;
        for (int d = 0; d < 3; d++) {
            sp.bdim[d][0] = c.face[d].bi; sp.bdim[d][1] = c.face[d].bj; sp.bdim[d][2] = c.face[d].bk;
            sp.home[d] = c.face[d].home;
        }
        sp.seam = c.sort.seam;
        sp.cap = c.cap;
        sp.px = s.p[0]; sp.py = s.p[1]; sp.pz = s.p[2];
        sp.h = (float)(0.5 * g.dx);
        sp.sr = sr;
        sp.blockdx = (float)chunkdx;
        sp.inv_blockdx = 1.0 / (double)sp.blockdx;
        sp.inv_chunkdx = 1.0 / chunkdx;
        sp.n = c.n;
        k_seam_home<<<(c.n + 255) / 256, 256, 0, c.stream>>>(sp);
        launches++;
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

int launch_p2g(Context &c, double radius, int method) {
    int launches = 0;
    const GridDesc &g = c.g;
    ParticleSoA &s = c.soa[c.cur];
    const float eps = 1e-6f;
    const float sr = (float)(radius + (double)eps);            // float sr = _particleRadius + eps;
    for (int d = 0; d < 3; d++) {
        FaceGrid &f = c.face[d];
        P2GParams P;
        P.g = g;
        P.gi = f.gi; P.gj = f.gj; P.gk = f.gk;
        P.kstore = f.kstore;
        P.bi = f.bi; P.bj = f.bj; P.bk = f.bk;
        P.active = f.active;
        P.bin_start = c.sort.bin_start;
        P.px = s.p[0]; P.py = s.p[1]; P.pz = s.p[2];
        P.vel = s.v[d];
        P.ax = s.a[3 * d + 0]; P.ay = s.a[3 * d + 1]; P.az = s.a[3 * d + 2];
        P.seam = c.sort.seam + (size_t)d * c.cap;
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
        if (d == 0) launch_dir<0>(c, P, method);
        if (d == 1) launch_dir<1>(c, P, method);
        if (d == 2) launch_dir<2>(c, P, method);
        launches++;
    }
    FFB_CUDA(cudaGetLastError());
    return launches;
}

}  // namespace ffb200
