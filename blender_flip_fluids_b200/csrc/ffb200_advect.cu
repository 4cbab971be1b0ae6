// ffb200_advect.cu -- marker-particle advection on the sorted SoA streams.
//
//   _RK3                         fluidsimulation.cpp:7616-7623  (Ralston RK3 through the MAC field)
//   _resolveCollision            fluidsimulation.cpp:7646-7721  (march + SDF projection)
//   _getBoundaryAABB             fluidsimulation.cpp:5175-5180
//   AABB::{expand,isPointInside,getNearestPointInsideAABB}   aabb.cpp:122-133, 493-518
//   Interpolation::trilinearInterpolate(vec3, dx, grid)      interpolation.cpp:72-112
//   Interpolation::trilinearInterpolateGradient              interpolation.cpp:197-259
//
// One thread per particle. The arithmetic repeats the reference's float/double mix operation
// for operation, so positions are bit-identical on identical inputs. The collision march is
// the data-dependent slow path: sorted particles share their 3dx near-solid flag with their
// neighbours, which keeps warps mostly convergent on the gate.
#include "ffb200_ctx.h"

#include <algorithm>
#include <cstdlib>

namespace ffb200 {

namespace {

struct Box {
    float px, py, pz;       // AABB::position (floats)
    double w, h, d;         // AABB::width/height/depth (doubles)
};

#ifndef FFB_ADV_THREADS
#define FFB_ADV_THREADS 256
#endif
// resident CTAs per SM the exact gathers are compiled for. They are latency-bound (L1/L2-hit gathers feeding long fp64
// chains), so occupancy beats register comfort: measured at 512^3, G2P + advection take 21.5 ms unbounded / at 3 CTAs
// (68-80 registers, no spills), 19.6 ms at 4 (64 registers), 18.9 ms at 5 (48 registers, 72-160 B of spills), 19.4 ms at 6
#ifndef FFB_ADV_MINB
#define FFB_ADV_MINB 5
#endif
#define FFB_ADV_BOUNDS __launch_bounds__(FFB_ADV_THREADS, FFB_ADV_MINB)

struct AdvectParams {
    GridDesc g;
    MacView mac;
    const float *phi;       // (I+1)(J+1)(kloc+1), first stored node plane = kbase
    const uint8_t *near_solid;
    const uint8_t *clear;   // per stored cell: Chebyshev distance (cells, capped) to the nearest cell with a node phi <= margin
    int ni, nj, nk;
    double inv_near;        // 1.0 / (3*dx)
    Box box;
    const float *px, *py, *pz;   // position in
    const float *k1x, *k1y, *k1z;   // RK3 stage-1 samples left by an APIC G2P on the same field (or null)
    float *opx, *opy, *opz;      // position out
    float c2, c3, c9;       // (float)(0.5dt), (float)(0.75dt), (float)(dt/9.0f)
    float step;             // _markerParticleStepDistanceFactor * (float)_dx
    float maxdist;          // (float)(_CFLConditionNumber * _dx)
    double buffer;          // (double)_solidBufferWidth * _dx
    int collide;
    int n;
    Window win;
};

__device__ __forceinline__ bool box_inside(const Box &b, float x, float y, float z) {
    return x >= b.px && y >= b.py && z >= b.pz && (double)x < (double)b.px + b.w && (double)y < (double)b.py + b.h &&
           (double)z < (double)b.pz + b.d;
}

// inside with room to spare (1e-4 of the box size): rounding of interpolated points cannot leave the box
__device__ __forceinline__ bool box_inside_margin(const Box &b, float x, float y, float z) {
    const double mx = 1e-4 * b.w, my = 1e-4 * b.h, mz = 1e-4 * b.d;
    return (double)x > (double)b.px + mx && (double)y > (double)b.py + my && (double)z > (double)b.pz + mz &&
           (double)x < (double)b.px + b.w - mx && (double)y < (double)b.py + b.h - my && (double)z < (double)b.pz + b.d - mz;
}

__device__ __forceinline__ void box_nearest_inside(const Box &b, float &x, float &y, float &z) {
    if (box_inside(b, x, y, z)) return;
    const double eps = 1e-6;
    const float mx = b.px + (float)b.w, my = b.py + (float)b.h, mz = b.pz + (float)b.d;
    x = fmaxf(x, b.px); y = fmaxf(y, b.py); z = fmaxf(z, b.pz);
    x = (float)fmin((double)x, (double)mx - eps);
    y = (float)fmin((double)y, (double)my - eps);
    z = (float)fmin((double)z, (double)mz - eps);
}

struct SdfCell {
    double ix, iy, iz;
    float v[8];             // 000,100,010,001,101,011,110,111 (the reference's corner order)
};

__device__ __forceinline__ void sdf_cell(const AdvectParams &P, float x, float y, float z, SdfCell &c) {
    const GridDesc &g = P.g;
    const int w = g.I + 1, h = g.J + 1, d = g.K + 1;
    const int i = pos2idx(x, g.inv_dx), j = pos2idx(y, g.inv_dx), k = pos2idx(z, g.inv_dx);
    c.ix = (double)(x - idx2posf(i, g.dx)) * g.inv_dx;
    c.iy = (double)(y - idx2posf(j, g.dx)) * g.inv_dx;
    c.iz = (double)(z - idx2posf(k, g.dx)) * g.inv_dx;
    const bool i0 = (unsigned)i < (unsigned)w, i1 = (unsigned)(i + 1) < (unsigned)w;
    const bool j0 = (unsigned)j < (unsigned)h, j1 = (unsigned)(j + 1) < (unsigned)h;
    const bool k0 = (unsigned)k < (unsigned)d, k1 = (unsigned)(k + 1) < (unsigned)d;
    const long long sj = w, sk = (long long)w * h;
    const long long base = (long long)i + sj * j + sk * (long long)(k - g.kbase);
    const float *f = P.phi;
    c.v[0] = (i0 && j0 && k0) ? __ldg(f + base) : 0.0f;
    c.v[1] = (i1 && j0 && k0) ? __ldg(f + base + 1) : 0.0f;
    c.v[2] = (i0 && j1 && k0) ? __ldg(f + base + sj) : 0.0f;
    c.v[3] = (i0 && j0 && k1) ? __ldg(f + base + sk) : 0.0f;
    c.v[4] = (i1 && j0 && k1) ? __ldg(f + base + sk + 1) : 0.0f;
    c.v[5] = (i0 && j1 && k1) ? __ldg(f + base + sk + sj) : 0.0f;
    c.v[6] = (i1 && j1 && k0) ? __ldg(f + base + sj + 1) : 0.0f;
    c.v[7] = (i1 && j1 && k1) ? __ldg(f + base + sk + sj + 1) : 0.0f;
}

__device__ __forceinline__ float sdf_sample(const AdvectParams &P, float x, float y, float z) {
    SdfCell c;
    sdf_cell(P, x, y, z, c);
    double p[8];
#pragma unroll
    for (int q = 0; q < 8; q++) p[q] = (double)c.v[q];
    return (float)trilerp8(p, c.ix, c.iy, c.iz);
}

__device__ __forceinline__ double bilerp(double v00, double v10, double v01, double v11, double ix, double iy) {
    const double l1 = (1 - ix) * v00 + ix * v10;
    const double l2 = (1 - ix) * v01 + ix * v11;
    return (1 - iy) * l1 + iy * l2;
}

__device__ __forceinline__ void sdf_gradient(const AdvectParams &P, float x, float y, float z, float &gx, float &gy, float &gz) {
    SdfCell c;
    sdf_cell(P, x, y, z, c);
    const float v000 = c.v[0], v100 = c.v[1], v010 = c.v[2], v001 = c.v[3], v101 = c.v[4], v011 = c.v[5], v110 = c.v[6],
                v111 = c.v[7];
    const float ddx00 = v100 - v000, ddx10 = v110 - v010, ddx01 = v101 - v001, ddx11 = v111 - v011;
    gx = (float)bilerp(ddx00, ddx10, ddx01, ddx11, c.iy, c.iz);
    const float ddy00 = v010 - v000, ddy10 = v110 - v100, ddy01 = v011 - v001, ddy11 = v111 - v101;
    gy = (float)bilerp(ddy00, ddy10, ddy01, ddy11, c.ix, c.iz);
    const float ddz00 = v001 - v000, ddz10 = v101 - v100, ddz01 = v011 - v010, ddz11 = v111 - v110;
    gz = (float)bilerp(ddz00, ddz10, ddz01, ddz11, c.ix, c.iy);
}

__device__ __forceinline__ bool near_solid(const AdvectParams &P, float x, float y, float z) {
    const int i = pos2idx(x, P.inv_near), j = pos2idx(y, P.inv_near), k = pos2idx(z, P.inv_near);
    // unchecked Array3d<bool> read in the reference (:7654-7656); out of range -> "near"
    if (!in_range3(i, j, k, P.ni, P.nj, P.nk)) return true;
    return P.near_solid[i + P.ni * (j + P.nj * k)] != 0;
}

// The early exits of _resolveCollision (:7646-7658) and the clearance shortcut: true when the march must run.
// n is clamped into the boundary box first, as the reference does, when its cell is outside the grid.
__device__ __noinline__ bool collision_gate(const AdvectParams &P, float ox, float oy, float oz, float &nx, float &ny, float &nz) {
    const GridDesc &g = P.g;
    const int gi = pos2idx(nx, g.inv_dx), gj = pos2idx(ny, g.inv_dx), gk = pos2idx(nz, g.inv_dx);
    if (!in_range3(gi, gj, gk, g.I, g.J, g.K)) box_nearest_inside(P.box, nx, ny, nz);
    if (!near_solid(P, ox, oy, oz) && !near_solid(P, nx, ny, nz)) return false;
    const int oi = pos2idx(ox, g.inv_dx), oj = pos2idx(oy, g.inv_dx), ok = pos2idx(oz, g.inv_dx);
    const int ni2 = pos2idx(nx, g.inv_dx), nj2 = pos2idx(ny, g.inv_dx), nk2 = pos2idx(nz, g.inv_dx);
    if (in_range3(oi, oj, ok - g.kbase, g.I, g.J, g.kloc)) {
        const int reach = max(max(abs(ni2 - oi), abs(nj2 - oj)), abs(nk2 - ok)) + 1;
        const int c = P.clear[(size_t)oi + (size_t)g.I * ((size_t)oj + (size_t)g.J * (ok - g.kbase))];
        if (c > reach && box_inside_margin(P.box, ox, oy, oz) && box_inside_margin(P.box, nx, ny, nz)) return false;
    }
    return true;
}

__device__ __noinline__ void resolve_collision(const AdvectParams &P, float ox, float oy, float oz, float &nx, float &ny,
                                               float &nz) {
    const GridDesc &g = P.g;
    const int gi = pos2idx(nx, g.inv_dx), gj = pos2idx(ny, g.inv_dx), gk = pos2idx(nz, g.inv_dx);
    if (!in_range3(gi, gj, gk, g.I, g.J, g.K)) box_nearest_inside(P.box, nx, ny, nz);
    if (!near_solid(P, ox, oy, oz) && !near_solid(P, nx, ny, nz)) return;
    // Exact shortcut of the march below. Every sample lies on the segment o -> n. If all SDF nodes of
    // all cells within the segment's cell bounding box (+1 cell for rounding of the sample positions)
    // are positive by a margin, every trilinear sample -- a convex combination of 8 such nodes -- is
    // positive, and if both end points are inside the (convex) boundary box by a margin so is every
    // sample: the march finds nothing and returns n unchanged. `clear` holds, per cell, the distance to
    // the nearest cell with a non-positive node (k_solid_clearance).
    {
        const int oi = pos2idx(ox, g.inv_dx), oj = pos2idx(oy, g.inv_dx), ok = pos2idx(oz, g.inv_dx);
        const int ni2 = pos2idx(nx, g.inv_dx), nj2 = pos2idx(ny, g.inv_dx), nk2 = pos2idx(nz, g.inv_dx);
        if (in_range3(oi, oj, ok - g.kbase, g.I, g.J, g.kloc)) {
            const int reach = max(max(abs(ni2 - oi), abs(nj2 - oj)), abs(nk2 - ok)) + 1;
            const int c = P.clear[(size_t)oi + (size_t)g.I * ((size_t)oj + (size_t)g.J * (ok - g.kbase))];
            if (c > reach && box_inside_margin(P.box, ox, oy, oz) && box_inside_margin(P.box, nx, ny, nz)) return;
        }
    }

    const float eps = 1e-6f;
    const float dxx = nx - ox, dyy = ny - oy, dzz = nz - oz;
    const float travel = vlen3(dxx, dyy, dzz);
    if (travel < eps) return;
    const int nsteps = (int)ceilf(travel / P.step);
    const float invlen = finv(travel);
    const float dirx = dxx * invlen, diry = dyy * invlen, dirz = dzz * invlen;

    float lx = ox, ly = oy, lz = oz;
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;
    bool found = false;
    float cphi = 0.0f;
    for (int st = 0; st < nsteps; st++) {
        if (st == nsteps - 1) {
            cx = nx; cy = ny; cz = nz;
        } else {
            const float t = (float)(st + 1) * P.step;
            cx = ox + dirx * t; cy = oy + diry * t; cz = oz + dirz * t;
        }
        const float phi = sdf_sample(P, cx, cy, cz);
        if (phi < 0.0f || !box_inside(P.box, cx, cy, cz)) {
            cphi = phi;
            found = true;
            break;
        }
        lx = cx; ly = cy; lz = cz;
    }
    if (!found) return;

    float rx, ry, rz;
    float gx, gy, gz;
    sdf_gradient(P, cx, cy, cz, gx, gy, gz);
    const float glen = vlen3(gx, gy, gz);
    if (glen > eps) {
        const float ginv = finv(glen);
        gx *= ginv; gy *= ginv; gz *= ginv;
        const float push = (float)((double)cphi - P.buffer);
        rx = cx - gx * push; ry = cy - gy * push; rz = cz - gz * push;
        const float rphi = sdf_sample(P, rx, ry, rz);
        const float rdist = vlen3(rx - cx, ry - cy, rz - cz);
        if (rphi < 0 || rdist > P.maxdist) { rx = lx; ry = ly; rz = lz; }
    } else {
        rx = lx; ry = ly; rz = lz;
    }
    if (!box_inside(P.box, rx, ry, rz)) {
        const float qx = rx, qy = ry, qz = rz;
        box_nearest_inside(P.box, rx, ry, rz);
        const float rphi = sdf_sample(P, rx, ry, rz);
        const float rdist = vlen3(rx - qx, ry - qy, rz - qz);
        if (rphi < 0.0f || rdist > P.maxdist) { rx = lx; ry = ly; rz = lz; }
    }
    nx = rx; ny = ry; nz = rz;
}

// With `list` the particles whose collision march must run are appended to it instead (k_advect_exact_list, pass 2, runs
// them in full warps): the march is 10-50 SDF samples for a few percent of the particles, and inline it leaves every
// warp that holds one of them idling on a lane or two.
__global__ void FFB_ADV_BOUNDS k_advect(const __grid_constant__ AdvectParams P, uint32_t *__restrict__ list,
                                        unsigned long long *__restrict__ stats) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n || window_skip(P.win, j)) return;
    const float x0 = P.px[j], y0 = P.py[j], z0 = P.pz[j];
    float k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z;
    if (P.k1x) {
        k1x = P.k1x[j]; k1y = P.k1y[j]; k1z = P.k1z[j];
    } else {
        mac_eval(P.g, P.mac, x0, y0, z0, k1x, k1y, k1z);
    }
    mac_eval(P.g, P.mac, x0 + k1x * P.c2, y0 + k1y * P.c2, z0 + k1z * P.c2, k2x, k2y, k2z);
    mac_eval(P.g, P.mac, x0 + k2x * P.c3, y0 + k2y * P.c3, z0 + k2z * P.c3, k3x, k3y, k3z);
    float x1 = x0 + ((k1x * 2.0f + k2x * 3.0f) + k3x * 4.0f) * P.c9;
    float y1 = y0 + ((k1y * 2.0f + k2y * 3.0f) + k3y * 4.0f) * P.c9;
    float z1 = z0 + ((k1z * 2.0f + k2z * 3.0f) + k3z * 4.0f) * P.c9;
    bool defer = false;
    if (P.collide) {
        if (list) defer = collision_gate(P, x0, y0, z0, x1, y1, z1);
        else resolve_collision(P, x0, y0, z0, x1, y1, z1);
    }
    if (!defer) {
        P.opx[j] = x1; P.opy[j] = y1; P.opz[j] = z1;
    }
    if (list) {
        const unsigned active = __activemask();
        const unsigned rej = __ballot_sync(active, defer);
        if (rej) {
            const int lane = threadIdx.x & 31, leader = __ffs(rej) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(stats + 0, (unsigned long long)__popc(rej));
            base = __shfl_sync(active, base, leader);
            if (defer) list[base + __popc(rej & ((1u << lane) - 1u))] = (uint32_t)j;
        }
    }
}

// ---- tolerance mode ---------------------------------------------------------------------------------------------
//
// RK3 through the fp32 evaluation of ffb200_common.cuh. The approximate end point differs from the exact one by a few
// float ulps of the coordinate plus ~1e-6 of the displacement; that is far inside the 1e-5 the north star allows for
// positions, but the collision code takes DISCRETE decisions on it (cell range, 3dx near-solid gate, clearance
// shortcut, SDF sign along the march), and a flipped decision moves a particle by up to 0.1 dx. So the approximate
// end point is only accepted where every decision is clear-cut with a guard band around it:
//   * its cell and its 3dx gate cell are at least kBand cells away from their planes and inside the grid,
//   * and either no gate cell (start or end) is marked near-solid, or the per-cell clearance proves, with one cell
//     of slack more than the exact shortcut needs, that the march would find nothing, with both end points inside
//     the boundary box by its 1e-4 margin.
// Everything else -- the particles that can actually touch a solid -- takes the exact path in full: the reference's
// fp64 RK3 and _resolveCollision, bit for bit. `stats` counts both populations.
// floor of a gate-cell coordinate t (units of 3 dx) with a band: false when t is so close to an integer that the few
// ulps between the approximate and the exact end point (8 ulps of the coordinate, ~5e-7 |t| relative, plus the RK3
// error, far below 1e-5 cells) could change the floor.
__device__ __forceinline__ bool gate_cell(float t, int &cell) {
    const float mm = t + kMagic;
    float q = t - (mm - kMagic);                           // in [-0.5, 0.5]
    int i = __float_as_int(mm) - kMagicBits;
    if (q < 0.0f) { q += 1.0f; i -= 1; }
    cell = i;
    const float band = fmaf(fabsf(t), 1e-6f, 2e-5f);
    return q > band && q < 1.0f - band;
}

__device__ __noinline__ void exact_advect(const AdvectParams &P, float x0, float y0, float z0, float &x1, float &y1, float &z1,
                                          bool have_k1 = false, float k1x = 0.0f, float k1y = 0.0f, float k1z = 0.0f) {
    float k2x, k2y, k2z, k3x, k3y, k3z;
    if (!have_k1) mac_eval(P.g, P.mac, x0, y0, z0, k1x, k1y, k1z);
    mac_eval(P.g, P.mac, x0 + k1x * P.c2, y0 + k1y * P.c2, z0 + k1z * P.c2, k2x, k2y, k2z);
    mac_eval(P.g, P.mac, x0 + k2x * P.c3, y0 + k2y * P.c3, z0 + k2z * P.c3, k3x, k3y, k3z);
    x1 = x0 + ((k1x * 2.0f + k2x * 3.0f) + k3x * 4.0f) * P.c9;
    y1 = y0 + ((k1y * 2.0f + k2y * 3.0f) + k3y * 4.0f) * P.c9;
    z1 = z0 + ((k1z * 2.0f + k2z * 3.0f) + k3z * 4.0f) * P.c9;
    if (P.collide) resolve_collision(P, x0, y0, z0, x1, y1, z1);
}

// Pass 1: every particle on the fp32 path. Accepted end points are written; the others are appended to a compact
// list (warp-aggregated slot allocation) for pass 2 -- running the exact code inline instead leaves the warps of a
// near-wall region with a handful of active lanes each (13.8 of 32 measured), which cost more than the whole exact kernel.
#ifndef FFB_ADV_FAST_MINB
#define FFB_ADV_FAST_MINB 5
#endif
__global__ void __launch_bounds__(FFB_ADV_THREADS, FFB_ADV_FAST_MINB) k_advect_fast(const __grid_constant__ AdvectParams P, const __grid_constant__ FastGrid fg,
                                             float inv_near, uint32_t *__restrict__ list, unsigned long long *__restrict__ stats) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n || window_skip(P.win, j)) return;
    const float x0 = P.px[j], y0 = P.py[j], z0 = P.pz[j];
    float k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z;
    if (P.k1x) {
        k1x = P.k1x[j]; k1y = P.k1y[j]; k1z = P.k1z[j];
    } else {
        fast_mac_eval(P.g, fg, P.mac, x0, y0, z0, k1x, k1y, k1z);
    }
    fast_mac_eval(P.g, fg, P.mac, x0 + k1x * P.c2, y0 + k1y * P.c2, z0 + k1z * P.c2, k2x, k2y, k2z);
    fast_mac_eval(P.g, fg, P.mac, x0 + k2x * P.c3, y0 + k2y * P.c3, z0 + k2z * P.c3, k3x, k3y, k3z);
    const float x1 = x0 + ((k1x * 2.0f + k2x * 3.0f) + k3x * 4.0f) * P.c9;
    const float y1 = y0 + ((k1y * 2.0f + k2y * 3.0f) + k3y * 4.0f) * P.c9;
    const float z1 = z0 + ((k1z * 2.0f + k2z * 3.0f) + k3z * 4.0f) * P.c9;
    bool accept = !P.collide;
    if (P.collide) {
        const GridDesc &g = P.g;
        // 3dx gate cells of the start and the end point, magic-number floors with a band around the planes: the start is
        // an exact input, but its reference floor is taken in double; the end differs from the exact one by a few ulps
        int s3[3], e3[3];
        bool sure = gate_cell(x0 * inv_near, s3[0]) & gate_cell(y0 * inv_near, s3[1]) & gate_cell(z0 * inv_near, s3[2]) &
                    gate_cell(x1 * inv_near, e3[0]) & gate_cell(y1 * inv_near, e3[1]) & gate_cell(z1 * inv_near, e3[2]);
        // the end point's CELL must be inside the grid (the reference clamps it into the boundary box otherwise)
        const float m = 1e-2f * P.step;                        // 1e-3 dx: above 8 ulps of a coordinate up to 4096 cells
        sure = sure && x1 > m && y1 > m && z1 > m && x1 < fg.xmax - m && y1 < fg.ymax - m && z1 < fg.zmax - m;
        if (sure) {
            const bool near_s = !in_range3(s3[0], s3[1], s3[2], P.ni, P.nj, P.nk) || P.near_solid[s3[0] + P.ni * (s3[1] + P.nj * s3[2])] != 0;
            const bool near_e = !in_range3(e3[0], e3[1], e3[2], P.ni, P.nj, P.nk) || P.near_solid[e3[0] + P.ni * (e3[1] + P.nj * e3[2])] != 0;
            if (!near_s && !near_e) {
                accept = true;                                   // the reference returns before looking at the SDF (:7654-7658)
            } else {
                // clearance shortcut of resolve_collision with one more cell of slack: the cells come from float-pair
                // floors (start: within 2^-44 of the double floor; end: a few ulps from the exact end point), so either
                // may be off by one near a plane
                const FastAxis sx = fast_axis(x0, fg), sy = fast_axis(y0, fg), sz = fast_axis(z0, fg);
                const FastAxis ex = fast_axis(x1, fg), ey = fast_axis(y1, fg), ez = fast_axis(z1, fg);
                if (in_range3(sx.i, sy.i, sz.i - g.kbase, g.I, g.J, g.kloc)) {
                    const int reach = max(max(abs(ex.i - sx.i), abs(ey.i - sy.i)), abs(ez.i - sz.i)) + 3;
                    int c = P.clear[(size_t)sx.i + (size_t)g.I * ((size_t)sy.i + (size_t)g.J * (sz.i - g.kbase))];
                    accept = c > reach && box_inside_margin(P.box, x0, y0, z0) && box_inside_margin(P.box, x1, y1, z1);
                }
            }
        }
    }
    if (accept) {
        P.opx[j] = x1; P.opy[j] = y1; P.opz[j] = z1;
    }
    const unsigned active = __activemask();
    const unsigned rej = __ballot_sync(active, !accept);
    if (rej) {
        const int lane = threadIdx.x & 31, leader = __ffs(rej) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(stats + 0, (unsigned long long)__popc(rej));
        base = __shfl_sync(active, base, leader);
        if (!accept) list[base + __popc(rej & ((1u << lane) - 1u))] = (uint32_t)j;
    }
}

// Pass 2: the listed particles through the reference's own arithmetic, one thread each, full warps.
// exact_k1: the k1 streams hold the exact G2P samples (exact mode); in tolerance mode they are fp32 values and stage 1 is
// evaluated again.
__global__ void FFB_ADV_BOUNDS k_advect_exact_list(const __grid_constant__ AdvectParams P, const uint32_t *__restrict__ list,
                                                   unsigned long long *__restrict__ stats, int exact_k1) {
    const unsigned long long count = stats[0];
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < count;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t j = list[t];
        float x1, y1, z1;
        if (exact_k1 && P.k1x)
            exact_advect(P, P.px[j], P.py[j], P.pz[j], x1, y1, z1, true, P.k1x[j], P.k1y[j], P.k1z[j]);
        else
            exact_advect(P, P.px[j], P.py[j], P.pz[j], x1, y1, z1);
        P.opx[j] = x1; P.opy[j] = y1; P.opz[j] = z1;
    }
    if (!exact_k1 && blockIdx.x == 0 && threadIdx.x == 0) stats[3] += count;   // tolerance mode: fallback total (single writer)
}

// clearance pass 1: 0 for a cell with any of its 8 SDF nodes <= margin (or outside the stored slab), else the cap
__global__ void k_solid_unsafe(const float *__restrict__ phi, uint8_t *__restrict__ clear, int I, int J, int kloc, float margin,
                               int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
    if (i >= I || j >= J) return;
    const size_t sj = I + 1, sk = (size_t)(I + 1) * (J + 1);
    const float *b = phi + i + sj * j + sk * k;
    float m = b[0];
    m = fminf(m, b[1]); m = fminf(m, b[sj]); m = fminf(m, b[sj + 1]);
    m = fminf(m, b[sk]); m = fminf(m, b[sk + 1]); m = fminf(m, b[sk + sj]); m = fminf(m, b[sk + sj + 1]);
    clear[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)] = (m > margin) ? (uint8_t)cap : (uint8_t)0;   // NaN -> 0
}

// clearance pass 2 (repeated cap times): Chebyshev distance relaxation over the 26 neighbours; cells
// beyond the stored range count as unsafe
__global__ void k_solid_relax(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int I, int J, int kloc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
    if (i >= I || j >= J) return;
    int m = in[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)];
    for (int c = -1; c <= 1; c++)
        for (int b = -1; b <= 1; b++)
            for (int a = -1; a <= 1; a++) {
                const int v = in_range3(i + a, j + b, k + c, I, J, kloc)
                                  ? in[(size_t)(i + a) + (size_t)I * ((size_t)(j + b) + (size_t)J * (k + c))] : 0;
                m = min(m, v + 1);
            }
    out[(size_t)i + (size_t)I * ((size_t)j + (size_t)J * k)] = (uint8_t)m;
}

void box_expand(Box &b, double v) {                     // aabb.cpp:122-128
    const double hh = 0.5 * v;
    const float hf = (float)hh;
    b.px -= hf; b.py -= hf; b.pz -= hf;
    b.w += v; b.h += v; b.d += v;
}

}  // namespace

// Called whenever the solid SDF changes (ffb200_set_solid): per-cell clearance used by the march shortcut.
int launch_solid_clearance(Context &c) {
    const GridDesc &g = c.g;
    const size_t cells = (size_t)g.I * g.J * g.kloc;
    if (!c.solid_clear[0]) {
        FFB_CUDA(cudaMalloc(&c.solid_clear[0], cells));
        FFB_CUDA(cudaMalloc(&c.solid_clear[1], cells));
    }
    constexpr int cap = 8;                                 // displacements are bounded by CFL * dx (5 cells)
    dim3 block(32, 4, 1), grid((g.I + 31) / 32, (g.J + 3) / 4, g.kloc);
    k_solid_unsafe<<<grid, block, 0, c.stream>>>(c.phi, c.solid_clear[0], g.I, g.J, g.kloc, (float)(1e-3 * g.dx), cap);
    for (int it = 0; it < cap; it++)
        k_solid_relax<<<grid, block, 0, c.stream>>>(c.solid_clear[it & 1], c.solid_clear[(it & 1) ^ 1], g.I, g.J, g.kloc);
    FFB_CUDA(cudaGetLastError());
    return cap + 1;                                        // result in solid_clear[0] (cap is even)
}

int launch_advect(Context &c, double dt, double cfl, int collide) {
    if (c.n == 0) return 0;
    if (collide && !c.has_solid) throw CudaError("ffb200_advect: collision resolution needs ffb200_set_solid first");
    ParticleSoA &s = c.soa[c.cur];
    const GridDesc &g = c.g;
    AdvectParams P;
    P.g = g;
    P.mac = MacView{c.face[0].vel, c.face[1].vel, c.face[2].vel};
    P.phi = c.phi;
    P.near_solid = c.near_solid;
    P.clear = c.solid_clear[0];
    P.ni = c.ni; P.nj = c.nj; P.nk = c.nk;
    P.inv_near = 1.0 / (3 * g.dx);                      // _nearSolidGridCellSizeFactor * _dx
    P.box.px = P.box.py = P.box.pz = 0.0f;              // _getBoundaryAABB
    P.box.w = g.I * g.dx; P.box.h = g.J * g.dx; P.box.d = g.K * g.dx;
    box_expand(P.box, -3 * g.dx - 1e-4);
    box_expand(P.box, -0.2f * g.dx);                    // boundary.expand(-_solidBufferWidth * _dx)
    ParticleSoA &o = c.nondestructive ? c.soa[c.cur ^ 1] : s;
    P.px = s.p[0]; P.py = s.p[1]; P.pz = s.p[2];
    P.opx = o.p[0]; P.opy = o.p[1]; P.opz = o.p[2];
    P.k1x = P.k1y = P.k1z = nullptr;
    static const bool reuse = [] { const char *e = std::getenv("FFB200_REUSE_G2P"); return e ? std::atoi(e) != 0 : true; }();
    if (reuse && c.k1_epoch == c.epoch) {                  // nothing touched particles or field since the G2P
        if (c.k1_buf == 2) {
            P.k1x = c.k1s[0]; P.k1y = c.k1s[1]; P.k1z = c.k1s[2];
        } else {
            ParticleSoA &k = c.soa[c.k1_buf];
            P.k1x = k.v[0]; P.k1y = k.v[1]; P.k1z = k.v[2];
        }
    }
    P.c2 = (float)(0.5 * dt);
    P.c3 = (float)(0.75 * dt);
    P.c9 = (float)(dt / 9.0f);
    P.step = 0.1f * (float)g.dx;
    P.maxdist = (float)(cfl * g.dx);
    P.buffer = (double)0.2f * g.dx;
    P.collide = collide;
    P.n = c.n;
    P.win = c.window;
    if (c.precision == FFB200_PRECISION_TOLERANCE) c.tol_advected += (unsigned long long)c.n;
    if (c.precision == FFB200_PRECISION_TOLERANCE) {
        unsigned long long *stats = tolerance_stats(c);
        uint32_t *list = c.sort.val[1];                        // sort scratch, free between the P2G and the next sort
        FFB_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long), c.stream));
        k_advect_fast<<<(c.n + FFB_ADV_THREADS - 1) / FFB_ADV_THREADS, FFB_ADV_THREADS, 0, c.stream>>>(
            P, make_fast_grid(g), (float)P.inv_near, list, stats);
        if (collide) {
            const int want = (c.n + FFB_ADV_THREADS - 1) / FFB_ADV_THREADS;
            k_advect_exact_list<<<std::min(want, c.sm_count * 8), FFB_ADV_THREADS, 0, c.stream>>>(P, list, stats, 0);
        }
        FFB_CUDA(cudaGetLastError());
        if (!c.nondestructive) c.sorted = false;
        return collide ? 2 : 1;
    }
    // exact mode: the march of the few particles that need one runs as a second pass over a compact list
    // (FFB200_ADVECT_TWO_PASS=0: inline, the round-1 kernel)
    static const bool two_pass = [] { const char *e = std::getenv("FFB200_ADVECT_TWO_PASS"); return e ? std::atoi(e) != 0 : true; }();
    if (two_pass && collide) {
        unsigned long long *stats = tolerance_stats(c);
        uint32_t *list = c.sort.val[1];
        FFB_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long), c.stream));
        k_advect<<<(c.n + FFB_ADV_THREADS - 1) / FFB_ADV_THREADS, FFB_ADV_THREADS, 0, c.stream>>>(P, list, stats);
        const int want = (c.n + FFB_ADV_THREADS - 1) / FFB_ADV_THREADS;
        k_advect_exact_list<<<std::min(want, c.sm_count * 8), FFB_ADV_THREADS, 0, c.stream>>>(P, list, stats, 1);
        FFB_CUDA(cudaGetLastError());
        if (!c.nondestructive) c.sorted = false;
        return 2;
    }
    k_advect<<<(c.n + FFB_ADV_THREADS - 1) / FFB_ADV_THREADS, FFB_ADV_THREADS, 0, c.stream>>>(P, nullptr, nullptr);
    FFB_CUDA(cudaGetLastError());
    if (!c.nondestructive) c.sorted = false;            // positions moved: bins are stale
    return 1;
}

}  // namespace ffb200
