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

// ---- particle window ------------------------------------------------------------------------
// A contiguous range of the SORTED particle streams, given by two entries of the bin table (bins are z-major, so
// "all particles of the cell planes [k_lo, k_hi)" is such a range). The per-particle kernels take it as a predicate,
// so the host can split a stage into "near the slab faces" and "interior" launches without knowing the indices
// (no device-to-host read): the z-slab driver overlaps its neighbour exchange with the interior part.
struct Window {
    const uint32_t *bin_start = nullptr;
    uint32_t lo_bin = 0, hi_bin = 0;
    int mode = 0;                   // 0: every particle; 1: inside [bin_start[lo_bin], bin_start[hi_bin]); 2: outside it
};

__device__ __forceinline__ bool window_skip(const Window &w, int j) {
    if (w.mode == 0) return false;
    const uint32_t r0 = __ldg(w.bin_start + w.lo_bin), r1 = __ldg(w.bin_start + w.hi_bin);
    const bool inside = (uint32_t)j >= r0 && (uint32_t)j < r1;
    return w.mode == 1 ? !inside : inside;
}

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

// eight faces at p + {0,1} + {0,sj} + {0,sk} in trilerp8's corner order {000,100,010,001,101,011,110,111};
// unconditional (interior cells only)
__device__ __forceinline__ void load8t(const float *__restrict__ p, int sj, int sk, float v[8]) {
    v[0] = __ldg(p);           v[1] = __ldg(p + 1);
    v[2] = __ldg(p + sj);      v[3] = __ldg(p + sk);
    v[4] = __ldg(p + sk + 1);  v[5] = __ldg(p + sk + sj);
    v[6] = __ldg(p + sj + 1);  v[7] = __ldg(p + sk + sj + 1);
}

__device__ __forceinline__ double trilerp8f(const float v[8], double x, double y, double z) {
    const double p[8] = {(double)v[0], (double)v[1], (double)v[2], (double)v[3], (double)v[4], (double)v[5], (double)v[6], (double)v[7]};
    return trilerp8(p, x, y, z);
}

// MACVelocityField::evaluateVelocityAtPositionLinear(vec3) (macvelocityfield.cpp:631-645):
// float position widened to double, zero outside the grid, components narrowed to float.
// Each axis needs its index/fraction twice only: in the unshifted frame (the component
// normal to it) and shifted by half a cell (the two other components). A point whose (unshifted) cell has its whole
// 3 x 3 x 3 neighbourhood inside the grid touches only existing faces in every frame: its 24 loads need no range test,
// use 32-bit offsets (face counts below 2^31, checked by the host) and are issued back to back; the arithmetic is the
// same either way.
__device__ __forceinline__ void mac_eval(const GridDesc &g, const MacView &m, float px, float py, float pz,
                                         float &ox, float &oy, float &oz) {
    const double x = px, y = py, z = pz;
    const double hdx = 0.5 * g.dx;
    const AxisCoord xu = axis_coord(x, g), yu = axis_coord(y, g), zu = axis_coord(z, g);
    if ((unsigned)(xu.i - 1) < (unsigned)(g.I - 2) && (unsigned)(yu.i - 1) < (unsigned)(g.J - 2) &&
        (unsigned)(zu.i - 1) < (unsigned)(g.K - 2)) {
        const AxisCoord xs = axis_coord(x - hdx, g), ys = axis_coord(y - hdx, g), zs = axis_coord(z - hdx, g);
        const int sju = g.I + 1, sku = (g.I + 1) * g.J, sjv = g.I, skv = g.I * (g.J + 1), sjw = g.I, skw = g.I * g.J;
        const int ks = zs.i - g.kbase, kk = zu.i - g.kbase;
        float a[8], b[8], c[8];                                  // all 24 loads in flight, then the three fp64 blends
        load8t(m.u + (xu.i + sju * ys.i + sku * ks), sju, sku, a);
        load8t(m.v + (xs.i + sjv * yu.i + skv * ks), sjv, skv, b);
        load8t(m.w + (xs.i + sjw * ys.i + skw * kk), sjw, skw, c);
        ox = (float)trilerp8f(a, xu.f, ys.f, zs.f);
        oy = (float)trilerp8f(b, xs.f, yu.f, zs.f);
        oz = (float)trilerp8f(c, xs.f, ys.f, zu.f);
        return;
    }
    if (!pos_in_grid(x, y, z, g)) {
        ox = oy = oz = 0.0f;
        return;
    }
    const AxisCoord xs = axis_coord(x - hdx, g), ys = axis_coord(y - hdx, g), zs = axis_coord(z - hdx, g);
    ox = (float)mac_lerp<0>(g, m.u, xu, ys, zs);
    oy = (float)mac_lerp<1>(g, m.v, xs, yu, zs);
    oz = (float)mac_lerp<2>(g, m.w, xs, ys, zu);
}

// ---- tolerance mode (ffb200_set_precision(FFB200_PRECISION_TOLERANCE)) ------------------------------------
//
// The north star asks for 1e-5 relative on particle velocities, APIC matrices and advected positions; the
// reference's fp64 gathers cost ~230 instructions per evaluation, 42 of them on the 16-lane conversion pipe.
// The tolerance path evaluates the same trilinear interpolant in fp32 with NO conversion and NO fp64
// instruction: the cell coordinate t = x / dx is carried as an unevaluated float pair (error ~2^-45 t), its
// floor is taken with the magic-number add, and the fraction -- the only place where fp32 would lose the
// 1e-5 at non-dyadic dx (SURVEY hard part 3) -- comes out with an absolute error near 2^-24. Every discrete
// decision of the advection (collision gate, clearance shortcut, boundary clamp) is taken with a guard band
// and falls back to the exact code when it is not clear-cut.

struct FastGrid {
    float inv_hi, inv_lo;           // 1/dx as a float pair: inv_hi + inv_lo == 1/dx to ~2^-48
    float xmax, ymax, zmax;         // largest floats strictly below dx*I, dx*J, dx*K (Grid3d::isPositionInGrid in float)
    int I, J, K, kbase;             // the interior fast path uses 32-bit face offsets (face counts below 2^31: checked by the host)
    int sju, sku, sjv, skv, sjw, skw;   // row / plane strides of u, v, w
};

struct FastAxis {
    int i;                          // floor(t)
    float f;                        // t - floor(t), in [0, 1]
};

constexpr float kMagic = 12582912.0f;               // 1.5 * 2^23: (t + kMagic) - kMagic rounds to nearest for |t| < 2^22
constexpr int kMagicBits = 0x4B400000;

// floor and fraction of t = x * (1/dx), unshifted frame.
__device__ __forceinline__ FastAxis fast_axis(float x, const FastGrid &g) {
    const float th = x * g.inv_hi;
    const float tl = fmaf(x, g.inv_lo, fmaf(x, g.inv_hi, -th));      // exact residual of the product + the low word
    const float m = th + kMagic;
    float f = (th - (m - kMagic)) + tl;                              // th - round(th) is exact
    int i = __float_as_int(m) - kMagicBits;
    if (f < 0.0f) { f += 1.0f; i -= 1; }
    if (f >= 1.0f) { f -= 1.0f; i += 1; }
    FastAxis r;
    r.i = i;
    r.f = f;
    return r;
}

// the same coordinate in the frame shifted by half a cell: t - 0.5
__device__ __forceinline__ FastAxis fast_shift(const FastAxis &a) {
    FastAxis r;
    const bool up = a.f >= 0.5f;
    r.i = up ? a.i : a.i - 1;
    r.f = up ? a.f - 0.5f : a.f + 0.5f;
    return r;
}

__device__ __forceinline__ float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }

// the eight faces around (cx, cy, cz) of component COMP, index c = di + 2 dj + 4 dk; out-of-range faces read 0
template <int COMP>
__device__ __forceinline__ void fast_faces(const GridDesc &g, const float *__restrict__ f, int i, int j, int k, float v[8]) {
    const int gw = g.I + (COMP == 0), gh = g.J + (COMP == 1), gd = g.K + (COMP == 2);
    const long long sj = gw, sk = (long long)gw * gh;
    const float *b = f + ((long long)i + sj * j + sk * (long long)(k - g.kbase));
    if ((unsigned)i < (unsigned)(gw - 1) && (unsigned)j < (unsigned)(gh - 1) && (unsigned)k < (unsigned)(gd - 1)) {
        v[0] = __ldg(b);           v[1] = __ldg(b + 1);
        v[2] = __ldg(b + sj);      v[3] = __ldg(b + sj + 1);
        v[4] = __ldg(b + sk);      v[5] = __ldg(b + sk + 1);
        v[6] = __ldg(b + sk + sj); v[7] = __ldg(b + sk + sj + 1);
    } else {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int di = c & 1, dj = (c >> 1) & 1, dk = c >> 2;
            v[c] = in_range3(i + di, j + dj, k + dk, gw, gh, gd) ? __ldg(b + di + sj * dj + sk * dk) : 0.0f;
        }
    }
}

__device__ __forceinline__ float fast_trilerp(const float v[8], float fx, float fy, float fz) {
    const float a0 = lerpf(v[0], v[1], fx), a1 = lerpf(v[2], v[3], fx), a2 = lerpf(v[4], v[5], fx), a3 = lerpf(v[6], v[7], fx);
    return lerpf(lerpf(a0, a1, fy), lerpf(a2, a3, fy), fz);
}

// gradient of the trilinear interpolant (the APIC affine row, fluidsimulation.cpp:6709-6769, up to rounding)
__device__ __forceinline__ void fast_gradient(const float v[8], float fx, float fy, float fz, float invdx, float &gx, float &gy,
                                              float &gz) {
    gx = invdx * lerpf(lerpf(v[1] - v[0], v[3] - v[2], fy), lerpf(v[5] - v[4], v[7] - v[6], fy), fz);
    gy = invdx * lerpf(lerpf(v[2] - v[0], v[3] - v[1], fx), lerpf(v[6] - v[4], v[7] - v[5], fx), fz);
    gz = invdx * lerpf(lerpf(v[4] - v[0], v[5] - v[1], fx), lerpf(v[6] - v[2], v[7] - v[3], fx), fy);
}

struct FastFrames {
    FastAxis xu, yu, zu, xs, ys, zs;      // unshifted / shifted by half a cell
};

__device__ __forceinline__ FastFrames fast_frames(float x, float y, float z, const FastGrid &fg) {
    FastFrames F;
    F.xu = fast_axis(x, fg); F.yu = fast_axis(y, fg); F.zu = fast_axis(z, fg);
    F.xs = fast_shift(F.xu); F.ys = fast_shift(F.yu); F.zs = fast_shift(F.zu);
    return F;
}

__device__ __forceinline__ bool fast_in_grid(float x, float y, float z, const FastGrid &fg) {
    return x >= 0.0f && y >= 0.0f && z >= 0.0f && x <= fg.xmax && y <= fg.ymax && z <= fg.zmax;
}

// eight faces at p + {0,1} + {0,sj} + {0,sk}, index c = di + 2 dj + 4 dk; unconditional (interior cells only)
__device__ __forceinline__ void load8(const float *__restrict__ p, int sj, int sk, float v[8]) {
    v[0] = __ldg(p);           v[1] = __ldg(p + 1);
    v[2] = __ldg(p + sj);      v[3] = __ldg(p + sj + 1);
    v[4] = __ldg(p + sk);      v[5] = __ldg(p + sk + 1);
    v[6] = __ldg(p + sk + sj); v[7] = __ldg(p + sk + sj + 1);
}

// A cell whose 3 x 3 x 3 neighbourhood lies inside the grid: every face any staggered frame of a point in this cell
// can touch exists, so the 24 loads of an evaluation need no range test and are issued back to back.
__device__ __forceinline__ bool fast_interior(int i, int j, int k, const FastGrid &fg) {
    return (unsigned)(i - 1) < (unsigned)(fg.I - 2) && (unsigned)(j - 1) < (unsigned)(fg.J - 2) && (unsigned)(k - 1) < (unsigned)(fg.K - 2);
}

// generic path (grid border, outside): per-face range tests, zero outside the grid
static __device__ __noinline__ void fast_mac_eval_border(const GridDesc &g, const FastGrid &fg, const MacView &m, float x, float y, float z,
                                                  float &ox, float &oy, float &oz) {
    if (!fast_in_grid(x, y, z, fg)) {
        ox = oy = oz = 0.0f;
        return;
    }
    const FastFrames F = fast_frames(x, y, z, fg);
    float v[8];
    fast_faces<0>(g, m.u, F.xu.i, F.ys.i, F.zs.i, v);
    ox = fast_trilerp(v, F.xu.f, F.ys.f, F.zs.f);
    fast_faces<1>(g, m.v, F.xs.i, F.yu.i, F.zs.i, v);
    oy = fast_trilerp(v, F.xs.f, F.yu.f, F.zs.f);
    fast_faces<2>(g, m.w, F.xs.i, F.ys.i, F.zu.i, v);
    oz = fast_trilerp(v, F.xs.f, F.ys.f, F.zu.f);
}

// MACVelocityField::evaluateVelocityAtPositionLinear in fp32 (zero outside the grid)
__device__ __forceinline__ void fast_mac_eval(const GridDesc &g, const FastGrid &fg, const MacView &m, float x, float y, float z,
                                              float &ox, float &oy, float &oz) {
    const FastAxis xu = fast_axis(x, fg), yu = fast_axis(y, fg), zu = fast_axis(z, fg);
    if (fast_interior(xu.i, yu.i, zu.i, fg)) {
        const FastAxis xs = fast_shift(xu), ys = fast_shift(yu), zs = fast_shift(zu);
        const int ks = zs.i - fg.kbase, kk = zu.i - fg.kbase;
        float a[8], b[8], c[8];
        load8(m.u + (xu.i + fg.sju * ys.i + fg.sku * ks), fg.sju, fg.sku, a);
        load8(m.v + (xs.i + fg.sjv * yu.i + fg.skv * ks), fg.sjv, fg.skv, b);
        load8(m.w + (xs.i + fg.sjw * ys.i + fg.skw * kk), fg.sjw, fg.skw, c);
        ox = fast_trilerp(a, xu.f, ys.f, zs.f);
        oy = fast_trilerp(b, xs.f, yu.f, zs.f);
        oz = fast_trilerp(c, xs.f, ys.f, zu.f);
        return;
    }
    fast_mac_eval_border(g, fg, m, x, y, z, ox, oy, oz);
}

// ---- error handling -------------------------------------------------------------------------

}  // namespace ffb200
