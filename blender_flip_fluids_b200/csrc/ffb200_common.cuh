// ffb200_common.cuh -- shared descriptors and exact-arithmetic device helpers.
//
// Everything here is compiled with -fmad=false: the reference is built for baseline x86-64
// (no FMA contraction, scalar SSE2), so a*b+c must round twice here too wherever a result is
// compared bit-for-bit with it. Where a fused multiply-add is wanted for speed it is written
// explicitly with fmaf()/fma().
//
// Reference paths are relative to rlguy/Blender-FLIP-Fluids src/engine (v1.8.5).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ffb200 {

constexpr int kChunk = 10;          // VelocityAdvector::_chunkWidth (velocityadvector.h:187)
constexpr int kApron = 4;           // half-cell apron around the bin grid (2 cells per side)

// Face/cell grid description. Arrays on the device cover the global k range
// [kbase, kbase + kloc (+1 for w / phi)); i and j are never split. A single-GPU context has
// kbase = 0 and kloc = K. All index maths is done with GLOBAL indices so a z-slab rank
// produces the same bits as a single-GPU run.
struct GridDesc {
    int I, J, K;                    // global cell dimensions
    int kbase, kloc;                // first stored cell plane and number of stored cell planes
    double dx;                      // cell size
    double inv_dx;                  // 1.0 / dx                    (grid3d.h:34)
    double inv_2dx;                 // 2.0 * (1.0 / dx)            (half-cell bins; h >> 1 == cell)
    // bin grid (half cells + apron), x-fastest
    int HX, HY, HZ;                 // 2I+2A, 2J+2A, 2*kloc+2A
    uint32_t nbins;                 // HX*HY*HZ ; key nbins = "outside the bin grid"
};

// ---- Grid3d index maths ---------------------------------------------------------------------

// Grid3d::positionToGridIndex (grid3d.h:32-60): (int)floor(x * (1.0/dx)) in double.
__device__ __forceinline__ int pos2idx(double x, double inv_dx) { return __double2int_rd(x * inv_dx); }

// Grid3d::GridIndexToPosition, vec3 flavour (grid3d.h:80-82): (float)i * dx in double, narrowed
// to float by the vec3 constructor.
__device__ __forceinline__ float idx2posf(int i, double dx) { return (float)((double)(float)i * dx); }

__device__ __forceinline__ bool in_range3(int i, int j, int k, int w, int h, int d) {
    return (unsigned)i < (unsigned)w && (unsigned)j < (unsigned)h && (unsigned)k < (unsigned)d;
}

// vmath::length (vmath.h:85-91): sqrt of a left-to-right float dot product.
__device__ __forceinline__ float vlen3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

// 1.0 / s evaluated in double and narrowed: the `float inv = 1.0 / s;` of vmath.cpp:100-103.
__device__ __forceinline__ float finv(float s) { return (float)(1.0 / (double)s); }

// ---- MAC field view -------------------------------------------------------------------------

struct MacView {
    const float *u, *v, *w;         // (I+1)*J*kloc, I*(J+1)*kloc, I*J*(kloc+1), x-fastest
};

// Interpolation::trilinearInterpolate(double p[8], x, y, z) (interpolation.cpp:61-70): the
// reference's corner order {000,100,010,001,101,011,110,111} and left-to-right term order.
__device__ __forceinline__ double trilerp8(const double p[8], double x, double y, double z) {
    return p[0] * (1 - x) * (1 - y) * (1 - z) +
           p[1] * x * (1 - y) * (1 - z) +
           p[2] * (1 - x) * y * (1 - z) +
           p[3] * (1 - x) * (1 - y) * z +
           p[4] * x * (1 - y) * z +
           p[5] * (1 - x) * y * z +
           p[6] * x * y * (1 - z) +
           p[7] * x * y * z;
}

// Grid3d::isPositionInGrid (grid3d.h:134-136).
__device__ __forceinline__ bool pos_in_grid(double x, double y, double z, const GridDesc &g) {
    return x >= 0 && y >= 0 && z >= 0 && x < g.dx * g.I && y < g.dx * g.J && z < g.dx * g.K;
}

// One axis of the reference's index/fraction maths (macvelocityfield.cpp:527-535):
// i = floor(x * (1/dx)), ix = (x - i*dx) * (1/dx), all in double.
struct AxisCoord {
    int i;
    double f;
};
__device__ __forceinline__ AxisCoord axis_coord(double x, const GridDesc &g) {
    AxisCoord c;
    c.i = pos2idx(x, g.inv_dx);
    c.f = (x - (double)c.i * g.dx) * g.inv_dx;
    return c;
}

// MACVelocityField::_interpolateLinear{U,V,W} (macvelocityfield.cpp:519-613) for one component,
// given the per-axis index/fraction of that component's staggered frame. Out-of-range corners
// read as 0 (_outOfRangeVector default, macvelocityfield.cpp:537-546); the common interior case
// skips the eight range tests. The blend is trilerp8: the reference's corner and term order.
template <int COMP>
__device__ __forceinline__ double mac_lerp(const GridDesc &g, const float *__restrict__ f, const AxisCoord &cx,
                                           const AxisCoord &cy, const AxisCoord &cz) {
    const int gw = g.I + (COMP == 0), gh = g.J + (COMP == 1), gd = g.K + (COMP == 2);
    const int i = cx.i, j = cy.i, k = cz.i;
    // stored planes start at kbase; callers guarantee the halo covers every sampled plane
    const long long sj = gw, sk = (long long)gw * gh;
    const float *b = f + ((long long)i + sj * j + sk * (long long)(k - g.kbase));
    double p[8];
    if ((unsigned)i < (unsigned)(gw - 1) && (unsigned)j < (unsigned)(gh - 1) && (unsigned)k < (unsigned)(gd - 1)) {
        p[0] = (double)__ldg(b);
        p[1] = (double)__ldg(b + 1);
        p[2] = (double)__ldg(b + sj);
        p[3] = (double)__ldg(b + sk);
        p[4] = (double)__ldg(b + sk + 1);
        p[5] = (double)__ldg(b + sk + sj);
        p[6] = (double)__ldg(b + sj + 1);
        p[7] = (double)__ldg(b + sk + sj + 1);
    } else {
        const bool i0 = (unsigned)i < (unsigned)gw, i1 = (unsigned)(i + 1) < (unsigned)gw;
        const bool j0 = (unsigned)j < (unsigned)gh, j1 = (unsigned)(j + 1) < (unsigned)gh;
        const bool k0 = (unsigned)k < (unsigned)gd, k1 = (unsigned)(k + 1) < (unsigned)gd;
        p[0] = (i0 && j0 && k0) ? (double)__ldg(b) : 0.0;
        p[1] = (i1 && j0 && k0) ? (double)__ldg(b + 1) : 0.0;
        p[2] = (i0 && j1 && k0) ? (double)__ldg(b + sj) : 0.0;
        p[3] = (i0 && j0 && k1) ? (double)__ldg(b + sk) : 0.0;
        p[4] = (i1 && j0 && k1) ? (double)__ldg(b + sk + 1) : 0.0;
        p[5] = (i0 && j1 && k1) ? (double)__ldg(b + sk + sj) : 0.0;
        p[6] = (i1 && j1 && k0) ? (double)__ldg(b + sj + 1) : 0.0;
        p[7] = (i1 && j1 && k1) ? (double)__ldg(b + sk + sj + 1) : 0.0;
    }
    return trilerp8(p, cx.f, cy.f, cz.f);
}

// mac_lerp of two fields of the same component at the same point (FLIP: new and saved field):
// one index / range computation, two sets of loads.
template <int COMP>
__device__ __forceinline__ void mac_lerp_pair(const GridDesc &g, const float *__restrict__ fa, const float *__restrict__ fb,
                                              const AxisCoord &cx, const AxisCoord &cy, const AxisCoord &cz, double &ra,
                                              double &rb) {
    const int gw = g.I + (COMP == 0), gh = g.J + (COMP == 1), gd = g.K + (COMP == 2);
    const int i = cx.i, j = cy.i, k = cz.i;
    const long long sj = gw, sk = (long long)gw * gh;
    const long long base = (long long)i + sj * j + sk * (long long)(k - g.kbase);
    const long long off[8] = {0, 1, sj, sk, sk + 1, sk + sj, sj + 1, sk + sj + 1};
    double pa[8], pb[8];
    if ((unsigned)i < (unsigned)(gw - 1) && (unsigned)j < (unsigned)(gh - 1) && (unsigned)k < (unsigned)(gd - 1)) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            pa[q] = (double)__ldg(fa + base + off[q]);
            pb[q] = (double)__ldg(fb + base + off[q]);
        }
    } else {
        const bool i0 = (unsigned)i < (unsigned)gw, i1 = (unsigned)(i + 1) < (unsigned)gw;
        const bool j0 = (unsigned)j < (unsigned)gh, j1 = (unsigned)(j + 1) < (unsigned)gh;
        const bool k0 = (unsigned)k < (unsigned)gd, k1 = (unsigned)(k + 1) < (unsigned)gd;
        const bool ok[8] = {i0 && j0 && k0, i1 && j0 && k0, i0 && j1 && k0, i0 && j0 && k1,
                            i1 && j0 && k1, i0 && j1 && k1, i1 && j1 && k0, i1 && j1 && k1};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            pa[q] = ok[q] ? (double)__ldg(fa + base + off[q]) : 0.0;
            pb[q] = ok[q] ? (double)__ldg(fb + base + off[q]) : 0.0;
        }
    }
    ra = trilerp8(pa, cx.f, cy.f, cz.f);
    rb = trilerp8(pb, cx.f, cy.f, cz.f);
}

// MACVelocityField::evaluateVelocityAtPositionLinear(vec3) (macvelocityfield.cpp:631-645):
// float position widened to double, zero outside the grid, components narrowed to float.
// Each axis needs its index/fraction twice only: in the unshifted frame (the component
// normal to it) and shifted by half a cell (the two other components).
__device__ __forceinline__ void mac_eval(const GridDesc &g, const MacView &m, float px, float py, float pz,
                                         float &ox, float &oy, float &oz) {
    const double x = px, y = py, z = pz;
    if (!pos_in_grid(x, y, z, g)) {
        ox = oy = oz = 0.0f;
        return;
    }
    const double hdx = 0.5 * g.dx;
    const AxisCoord xu = axis_coord(x, g), yu = axis_coord(y, g), zu = axis_coord(z, g);
    const AxisCoord xs = axis_coord(x - hdx, g), ys = axis_coord(y - hdx, g), zs = axis_coord(z - hdx, g);
    ox = (float)mac_lerp<0>(g, m.u, xu, ys, zs);
    oy = (float)mac_lerp<1>(g, m.v, xs, yu, zs);
    oz = (float)mac_lerp<2>(g, m.w, xs, ys, zu);
}

// ---- error handling -------------------------------------------------------------------------

}  // namespace ffb200
