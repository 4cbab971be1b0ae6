// ffb200_g2p.cu -- grid-to-particle velocity update on the sorted SoA particle streams.
//
//   FLIP  _updatePICFLIPMarkerParticleVelocitiesThread   fluidsimulation.cpp:6771-6784
//   APIC  _updatePICAPICMarkerParticleVelocitiesThread   fluidsimulation.cpp:6791-6843
//         + _getIndicesAndGradientWeights                fluidsimulation.cpp:6709-6769
//
// One thread per particle; consecutive threads hold spatially adjacent particles (sorted by
// half-cell), so the 8-corner face loads of a warp fall into a few cache lines. The gathers
// repeat the reference's fp64 index/fraction/blend arithmetic operation for operation
// (mac_eval in ffb200_common.cuh), so on identical inputs the results are bit-identical.
#include "ffb200_ctx.h"

namespace ffb200 {

namespace {

#ifndef FFB_G2P_THREADS
#define FFB_G2P_THREADS 256
#endif
// resident CTAs per SM the exact gathers are compiled for. They are latency-bound (L1/L2-hit gathers feeding long fp64
// chains), so occupancy beats register comfort: measured at 512^3, G2P + advection take 21.5 ms unbounded / at 3 CTAs
// (68-80 registers, no spills), 19.6 ms at 4 (64 registers), 18.9 ms at 5 (48 registers, 72-160 B of spills), 19.4 ms at 6
#ifndef FFB_G2P_MINB
#define FFB_G2P_MINB 5
#endif
#define FFB_G2P_BOUNDS __launch_bounds__(FFB_G2P_THREADS, FFB_G2P_MINB)
// the FLIP gather samples two fields per component and keeps more state live: 4 CTAs (64 registers, no spills)
#ifndef FFB_G2P_FLIP_MINB
#define FFB_G2P_FLIP_MINB 4
#endif

struct G2PParams {
    GridDesc g;
    MacView cur, saved;
    const float *px, *py, *pz;
    const float *vx, *vy, *vz;   // velocity in
    float *ovx, *ovy, *ovz;      // velocity out (same arrays unless the context is non-destructive)
    float *a[9];
    float *k1[3];       // FLIP: vPIC per particle, kept for the advection
    float rp, rf;       // (float)_ratioPICFLIP, (float)(1 - _ratioPICFLIP)
    float h;            // 0.5f * _dx
    float inv_s;        // (float)(1.0 / (float)_dx)   (vec3 / _dx)
    float invdx;        // 1.0f / _dx
    int n;              // end of the particle range of this launch ...
    int first = 0;      // ... and its start (a pipelined upload launches the sorted particles range by range)
    Window win;
};

__global__ void __launch_bounds__(FFB_G2P_THREADS, FFB_G2P_FLIP_MINB) k_g2p_flip(const __grid_constant__ G2PParams P) {
    const int j = P.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n || window_skip(P.win, j)) return;
    const float px = P.px[j], py = P.py[j], pz = P.pz[j];
    const GridDesc &g = P.g;
    const double x = px, y = py, z = pz;
    float pic[3] = {0.0f, 0.0f, 0.0f}, old[3] = {0.0f, 0.0f, 0.0f};
    if (pos_in_grid(x, y, z, g)) {                             // both fields sampled with one index computation
        const double hdx = 0.5 * g.dx;
        const AxisCoord xu = axis_coord(x, g), yu = axis_coord(y, g), zu = axis_coord(z, g);
        const AxisCoord xs = axis_coord(x - hdx, g), ys = axis_coord(y - hdx, g), zs = axis_coord(z - hdx, g);
        double a, b;
        mac_lerp_pair<0>(g, P.cur.u, P.saved.u, xu, ys, zs, a, b);
        pic[0] = (float)a; old[0] = (float)b;
        mac_lerp_pair<1>(g, P.cur.v, P.saved.v, xs, yu, zs, a, b);
        pic[1] = (float)a; old[1] = (float)b;
        mac_lerp_pair<2>(g, P.cur.w, P.saved.w, xs, ys, zu, a, b);
        pic[2] = (float)a; old[2] = (float)b;
    }
    const float v0 = P.vx[j], v1 = P.vy[j], v2 = P.vz[j];
    // vFLIP = vel + vPIC - saved(p); v = r*vPIC + (1-r)*vFLIP   (:6779-6781)
    const float f0 = (v0 + pic[0]) - old[0], f1 = (v1 + pic[1]) - old[1], f2 = (v2 + pic[2]) - old[2];
    P.ovx[j] = pic[0] * P.rp + f0 * P.rf;
    P.ovy[j] = pic[1] * P.rp + f1 * P.rf;
    P.ovz[j] = pic[2] * P.rp + f2 * P.rf;
    // vPIC is also the first RK3 stage of the advection that follows
    P.k1[0][j] = pic[0]; P.k1[1][j] = pic[1]; P.k1[2][j] = pic[2];
}

// One MAC component of the APIC update: the affine row (sum over the 8 faces around the particle
// in the component's staggered frame of gradWeight * face, :6709-6769, float arithmetic) and the
// interpolated velocity (mac_lerp, double arithmetic). Both read the same 8 faces whenever the
// float and the double index arithmetic agree on the cell -- always, except within an ulp of a
// cell plane -- so they are loaded once.
__device__ __forceinline__ void apic_math(const G2PParams &P, const float v[8], float ix, float iy, float iz, float &ox, float &oy,
                                          float &oz);

template <int DIR>
__device__ __forceinline__ void apic_component(const G2PParams &P, const float *__restrict__ f, float px, float py, float pz,
                                               bool in_grid, const AxisCoord &cx, const AxisCoord &cy, const AxisCoord &cz,
                                               float &vel, float &ox, float &oy, float &oz) {
    const GridDesc &g = P.g;
    const int gw = g.I + (DIR == 0), gh = g.J + (DIR == 1), gd = g.K + (DIR == 2);
    const float x = px - (DIR == 0 ? 0.0f : P.h), y = py - (DIR == 1 ? 0.0f : P.h), z = pz - (DIR == 2 ? 0.0f : P.h);
    const int gi = pos2idx(x, g.inv_dx), gj = pos2idx(y, g.inv_dx), gk = pos2idx(z, g.inv_dx);
    const float ix = (x - idx2posf(gi, g.dx)) * P.inv_s;
    const float iy = (y - idx2posf(gj, g.dx)) * P.inv_s;
    const float iz = (z - idx2posf(gk, g.dx)) * P.inv_s;
    // faces in the order c = di + 2 dj + 4 dk; out-of-range faces read 0 (skipping a term and adding
    // w * 0 give the same sum)
    const long long sj = gw, sk = (long long)gw * gh;
    const float *b = f + ((long long)gi + sj * gj + sk * (long long)(gk - g.kbase));
    float v[8];
    if ((unsigned)gi < (unsigned)(gw - 1) && (unsigned)gj < (unsigned)(gh - 1) && (unsigned)gk < (unsigned)(gd - 1)) {
        v[0] = __ldg(b);          v[1] = __ldg(b + 1);
        v[2] = __ldg(b + sj);     v[3] = __ldg(b + sj + 1);
        v[4] = __ldg(b + sk);     v[5] = __ldg(b + sk + 1);
        v[6] = __ldg(b + sk + sj); v[7] = __ldg(b + sk + sj + 1);
    } else {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int di = c & 1, dj = (c >> 1) & 1, dk = c >> 2;
            v[c] = in_range3(gi + di, gj + dj, gk + dk, gw, gh, gd) ? __ldg(b + di + sj * dj + sk * dk) : 0.0f;
        }
    }

    apic_math(P, v, ix, iy, iz, ox, oy, oz);

    // velocity component (zero outside the grid, macvelocityfield.cpp:631-645)
    if (!in_grid) {
        vel = 0.0f;
    } else if (gi == cx.i && gj == cy.i && gk == cz.i) {
        // trilerp8 corner order {000,100,010,001,101,011,110,111}
        const double p[8] = {(double)v[0], (double)v[1], (double)v[2], (double)v[4],
                             (double)v[5], (double)v[6], (double)v[3], (double)v[7]};
        vel = (float)trilerp8(p, cx.f, cy.f, cz.f);
    } else {
        vel = (float)mac_lerp<DIR>(g, f, cx, cy, cz);
    }
}

// The affine row from the eight faces v (index c = di + 2 dj + 4 dk) and the float fractions of the gradient frame.
__device__ __forceinline__ void apic_math(const G2PParams &P, const float v[8], float ix, float iy, float iz, float &ox, float &oy,
                                          float &oz) {
    const float invdx = P.invdx;
    const float mx = 1.0f - ix, my = 1.0f - iy, mz = 1.0f - iz;
    // gradient weights (fluidsimulation.cpp:6737-6768). The reference's 24 three-factor products
    // reduce to 21 multiplications: (-a)*b == -(a*b) exactly, and factors are shared where the
    // reference associates them the same way. Signs are applied in the sums below.
    const float C = invdx * my, D = invdx * iy;               // x row: ((+-invdx) * {my|iy}) * {mz|iz}
    const float xw[4] = {C * mz, D * mz, C * iz, D * iz};
    const float A = mx * invdx, B = ix * invdx;               // y row: ({mx|ix} * (+-invdx)) * {mz|iz}
    const float yw[4] = {A * mz, B * mz, A * iz, B * iz};
    const float z0 = A * my;                                   // z row: (-invdx*mx)*my, then ({mx|ix}*{my|iy}) * (+-invdx)
    const float z1 = (ix * my) * invdx, z2 = (mx * iy) * invdx, z3 = (ix * iy) * invdx, z4 = (mx * my) * invdx;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    sx -= xw[0] * v[0]; sy -= yw[0] * v[0]; sz -= z0 * v[0];
    sx += xw[0] * v[1]; sy -= yw[1] * v[1]; sz -= z1 * v[1];
    sx -= xw[1] * v[2]; sy += yw[0] * v[2]; sz -= z2 * v[2];
    sx += xw[1] * v[3]; sy += yw[1] * v[3]; sz -= z3 * v[3];
    sx -= xw[2] * v[4]; sy -= yw[2] * v[4]; sz += z4 * v[4];
    sx += xw[2] * v[5]; sy -= yw[3] * v[5]; sz += z1 * v[5];
    sx -= xw[3] * v[6]; sy += yw[2] * v[6]; sz += z2 * v[6];
    sx += xw[3] * v[7]; sy += yw[3] * v[7]; sz += z3 * v[7];
    ox = sx; oy = sy; oz = sz;
}

// trilerp8 of faces given in the c = di + 2 dj + 4 dk order (reordered to the reference's corner order)
__device__ __forceinline__ float trilerp_c(const float v[8], double fx, double fy, double fz) {
    const double p[8] = {(double)v[0], (double)v[1], (double)v[2], (double)v[4], (double)v[5], (double)v[6], (double)v[3], (double)v[7]};
    return (float)trilerp8(p, fx, fy, fz);
}

__global__ void FFB_G2P_BOUNDS k_g2p_apic(const __grid_constant__ G2PParams P) {
    const int j = P.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n || window_skip(P.win, j)) return;
    const float px = P.px[j], py = P.py[j], pz = P.pz[j];
    const GridDesc &g = P.g;
    const double x = px, y = py, z = pz;
    const double hdx = 0.5 * g.dx;
    const AxisCoord xu = axis_coord(x, g), yu = axis_coord(y, g), zu = axis_coord(z, g);
    const AxisCoord xs = axis_coord(x - hdx, g), ys = axis_coord(y - hdx, g), zs = axis_coord(z - hdx, g);
    // Interior cell (its 3 x 3 x 3 neighbourhood is inside the grid: every face of every frame exists) whose gradient
    // frames -- the reference's FLOAT-shifted coordinates, floored in double (:6717-6735) -- pick the same cells as
    // the velocity frames: the 24 faces are loaded once, back to back, with 32-bit offsets; same arithmetic as below.
    if ((unsigned)(xu.i - 1) < (unsigned)(g.I - 2) && (unsigned)(yu.i - 1) < (unsigned)(g.J - 2) &&
        (unsigned)(zu.i - 1) < (unsigned)(g.K - 2)) {
        const float fxs = px - P.h, fys = py - P.h, fzs = pz - P.h;
        const int gxs = pos2idx(fxs, g.inv_dx), gys = pos2idx(fys, g.inv_dx), gzs = pos2idx(fzs, g.inv_dx);
        if (gxs == xs.i && gys == ys.i && gzs == zs.i) {
            const int sju = g.I + 1, sku = (g.I + 1) * g.J, sjv = g.I, skv = g.I * (g.J + 1), sjw = g.I, skw = g.I * g.J;
            const int ks = zs.i - g.kbase, kk = zu.i - g.kbase;
            float a[8], b[8], c[8];
            load8(P.cur.u + (xu.i + sju * ys.i + sku * ks), sju, sku, a);
            load8(P.cur.v + (xs.i + sjv * yu.i + skv * ks), sjv, skv, b);
            load8(P.cur.w + (xs.i + sjw * ys.i + skw * kk), sjw, skw, c);
            // float fractions of the gradient frames: (x - GridIndexToPosition(g)) * (float)(1 / (float)dx)
            const float ixu = (px - idx2posf(xu.i, g.dx)) * P.inv_s, iyu = (py - idx2posf(yu.i, g.dx)) * P.inv_s,
                        izu = (pz - idx2posf(zu.i, g.dx)) * P.inv_s;
            const float ixs = (fxs - idx2posf(gxs, g.dx)) * P.inv_s, iys = (fys - idx2posf(gys, g.dx)) * P.inv_s,
                        izs = (fzs - idx2posf(gzs, g.dx)) * P.inv_s;
            float g0, g1, g2;
            apic_math(P, a, ixu, iys, izs, g0, g1, g2);
            P.a[0][j] = g0; P.a[1][j] = g1; P.a[2][j] = g2;
            P.ovx[j] = trilerp_c(a, xu.f, ys.f, zs.f);
            apic_math(P, b, ixs, iyu, izs, g0, g1, g2);
            P.a[3][j] = g0; P.a[4][j] = g1; P.a[5][j] = g2;
            P.ovy[j] = trilerp_c(b, xs.f, yu.f, zs.f);
            apic_math(P, c, ixs, iys, izu, g0, g1, g2);
            P.a[6][j] = g0; P.a[7][j] = g1; P.a[8][j] = g2;
            P.ovz[j] = trilerp_c(c, xs.f, ys.f, zu.f);
            return;
        }
    }
    const bool in_grid = pos_in_grid(x, y, z, g);
    float v0, v1, v2, ax, ay, az;
    apic_component<0>(P, P.cur.u, px, py, pz, in_grid, xu, ys, zs, v0, ax, ay, az);
    P.a[0][j] = ax; P.a[1][j] = ay; P.a[2][j] = az;
    apic_component<1>(P, P.cur.v, px, py, pz, in_grid, xs, yu, zs, v1, ax, ay, az);
    P.a[3][j] = ax; P.a[4][j] = ay; P.a[5][j] = az;
    apic_component<2>(P, P.cur.w, px, py, pz, in_grid, xs, ys, zu, v2, ax, ay, az);
    P.a[6][j] = ax; P.a[7][j] = ay; P.a[8][j] = az;
    P.ovx[j] = v0; P.ovy[j] = v1; P.ovz[j] = v2;
}

// ---- tolerance mode (ffb200_common.cuh, "tolerance mode") ---------------------------------------------------
//
// The same updates with the trilinear interpolant evaluated in fp32 from float-pair cell coordinates: no fp64 and
// no conversion instruction on the common path. Velocities and affine rows agree with the exact kernels within a
// few fp32 ulps of the field's magnitude (the 1e-5 of the north star leaves two orders of margin). The affine rows
// are the gradient of the interpolant and jump across cell planes, so a component whose gradient frame sits within
// 1e-6 cells of a plane -- where the float-pair floor could disagree with the reference's double floor -- is
// evaluated by the exact apic_component instead.
struct FastCount {
    unsigned long long fast, exact;
};

__device__ __forceinline__ void fast_count(unsigned long long *counter, bool slow) {
    const unsigned m = __ballot_sync(__activemask(), slow);
    if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1) && m) atomicAdd(counter, (unsigned long long)__popc(m));
}

__device__ __forceinline__ bool near_plane(float f) { return f < 1e-6f || f > 1.0f - 1e-6f; }

template <int DIR>
__device__ __noinline__ void exact_apic_fallback(const G2PParams &P, const float *__restrict__ f, float px, float py, float pz,
                                                 bool in_grid, float &vel, float &ox, float &oy, float &oz) {
    const GridDesc &g = P.g;
    const double x = px, y = py, z = pz, hdx = 0.5 * g.dx;
    const AxisCoord cx = axis_coord(DIR == 0 ? x : x - hdx, g), cy = axis_coord(DIR == 1 ? y : y - hdx, g),
                    cz = axis_coord(DIR == 2 ? z : z - hdx, g);
    apic_component<DIR>(P, f, px, py, pz, in_grid, cx, cy, cz, vel, ox, oy, oz);
}

template <int DIR>
__device__ __forceinline__ void fast_apic_component(const G2PParams &P, const float *__restrict__ f, float px, float py, float pz,
                                                    bool in_grid, const FastAxis &vx, const FastAxis &vy, const FastAxis &vz,
                                                    const FastAxis &gx, const FastAxis &gy, const FastAxis &gz, float &vel,
                                                    float &ox, float &oy, float &oz, bool &slow) {
    if (near_plane(gx.f) || near_plane(gy.f) || near_plane(gz.f)) {          // rare: the reference's own arithmetic
        exact_apic_fallback<DIR>(P, f, px, py, pz, in_grid, vel, ox, oy, oz);
        slow = true;
        return;
    }
    float v[8];
    fast_faces<DIR>(P.g, f, gx.i, gy.i, gz.i, v);
    fast_gradient(v, gx.f, gy.f, gz.f, P.invdx, ox, oy, oz);
    if (!in_grid) {
        vel = 0.0f;
        return;
    }
    if (vx.i != gx.i || vy.i != gy.i || vz.i != gz.i) fast_faces<DIR>(P.g, f, vx.i, vy.i, vz.i, v);   // an ulp from a plane
    vel = fast_trilerp(v, vx.f, vy.f, vz.f);
}

// border cells, points outside the grid, gradient frames within 1e-6 cells of a plane: the per-component code above
__device__ __noinline__ bool g2p_apic_fast_generic(const G2PParams &P, const FastGrid &fg, float px, float py, float pz, int j) {
    float vel[3], aff[9];
    bool slow = false;
    const bool in_grid = fast_in_grid(px, py, pz, fg);
    const FastFrames F = fast_frames(px, py, pz, fg);
    const FastAxis gxs = fast_axis(px - P.h, fg), gys = fast_axis(py - P.h, fg), gzs = fast_axis(pz - P.h, fg);
    fast_apic_component<0>(P, P.cur.u, px, py, pz, in_grid, F.xu, F.ys, F.zs, F.xu, gys, gzs, vel[0], aff[0], aff[1], aff[2], slow);
    fast_apic_component<1>(P, P.cur.v, px, py, pz, in_grid, F.xs, F.yu, F.zs, gxs, F.yu, gzs, vel[1], aff[3], aff[4], aff[5], slow);
    fast_apic_component<2>(P, P.cur.w, px, py, pz, in_grid, F.xs, F.ys, F.zu, gxs, gys, F.zu, vel[2], aff[6], aff[7], aff[8], slow);
#pragma unroll
    for (int q = 0; q < 9; q++) P.a[q][j] = aff[q];
    P.ovx[j] = vel[0]; P.ovy[j] = vel[1]; P.ovz[j] = vel[2];
    return slow;
}

#ifndef FFB_G2P_FAST_MINB
#define FFB_G2P_FAST_MINB 5
#endif
__global__ void __launch_bounds__(FFB_G2P_THREADS, FFB_G2P_FAST_MINB)
    k_g2p_apic_fast(const __grid_constant__ G2PParams P, const __grid_constant__ FastGrid fg, unsigned long long *__restrict__ stats) {
    const int j = P.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n || window_skip(P.win, j)) return;
    const float px = P.px[j], py = P.py[j], pz = P.pz[j];
    const FastAxis xu = fast_axis(px, fg), yu = fast_axis(py, fg), zu = fast_axis(pz, fg);
    const FastAxis xs = fast_shift(xu), ys = fast_shift(yu), zs = fast_shift(zu);
    // the affine rows use the reference's FLOAT-shifted coordinates (p - (float)(0.5f * dx), fluidsimulation.cpp:6717)
    const FastAxis gxs = fast_axis(px - P.h, fg), gys = fast_axis(py - P.h, fg), gzs = fast_axis(pz - P.h, fg);
    bool slow = false;
    // interior cell, both shifts agree on the cell, every gradient frame clear of its planes: 24 loads back to back
    const bool quick = fast_interior(xu.i, yu.i, zu.i, fg) && gxs.i == xs.i && gys.i == ys.i && gzs.i == zs.i &&
                       !(near_plane(xu.f) || near_plane(yu.f) || near_plane(zu.f) || near_plane(gxs.f) || near_plane(gys.f) ||
                         near_plane(gzs.f));
    if (quick) {
        const int ks = zs.i - fg.kbase, kk = zu.i - fg.kbase;
        float a[8], b[8], c[8];
        load8(P.cur.u + (xu.i + fg.sju * ys.i + fg.sku * ks), fg.sju, fg.sku, a);
        load8(P.cur.v + (xs.i + fg.sjv * yu.i + fg.skv * ks), fg.sjv, fg.skv, b);
        load8(P.cur.w + (xs.i + fg.sjw * ys.i + fg.skw * kk), fg.sjw, fg.skw, c);
        float g0, g1, g2;
        P.ovx[j] = fast_trilerp(a, xu.f, ys.f, zs.f);
        fast_gradient(a, xu.f, gys.f, gzs.f, P.invdx, g0, g1, g2);
        P.a[0][j] = g0; P.a[1][j] = g1; P.a[2][j] = g2;
        P.ovy[j] = fast_trilerp(b, xs.f, yu.f, zs.f);
        fast_gradient(b, gxs.f, yu.f, gzs.f, P.invdx, g0, g1, g2);
        P.a[3][j] = g0; P.a[4][j] = g1; P.a[5][j] = g2;
        P.ovz[j] = fast_trilerp(c, xs.f, ys.f, zu.f);
        fast_gradient(c, gxs.f, gys.f, zu.f, P.invdx, g0, g1, g2);
        P.a[6][j] = g0; P.a[7][j] = g1; P.a[8][j] = g2;
    } else {
        slow = g2p_apic_fast_generic(P, fg, px, py, pz, j);
    }
    fast_count(stats + 1, slow);
}

__device__ __noinline__ void g2p_flip_fast_generic(const G2PParams &P, const FastGrid &fg, float px, float py, float pz, float pic[3],
                                                    float old[3]) {
    pic[0] = pic[1] = pic[2] = old[0] = old[1] = old[2] = 0.0f;
    if (!fast_in_grid(px, py, pz, fg)) return;
    const FastFrames F = fast_frames(px, py, pz, fg);
    float v[8];
    fast_faces<0>(P.g, P.cur.u, F.xu.i, F.ys.i, F.zs.i, v);
    pic[0] = fast_trilerp(v, F.xu.f, F.ys.f, F.zs.f);
    fast_faces<0>(P.g, P.saved.u, F.xu.i, F.ys.i, F.zs.i, v);
    old[0] = fast_trilerp(v, F.xu.f, F.ys.f, F.zs.f);
    fast_faces<1>(P.g, P.cur.v, F.xs.i, F.yu.i, F.zs.i, v);
    pic[1] = fast_trilerp(v, F.xs.f, F.yu.f, F.zs.f);
    fast_faces<1>(P.g, P.saved.v, F.xs.i, F.yu.i, F.zs.i, v);
    old[1] = fast_trilerp(v, F.xs.f, F.yu.f, F.zs.f);
    fast_faces<2>(P.g, P.cur.w, F.xs.i, F.ys.i, F.zu.i, v);
    pic[2] = fast_trilerp(v, F.xs.f, F.ys.f, F.zu.f);
    fast_faces<2>(P.g, P.saved.w, F.xs.i, F.ys.i, F.zu.i, v);
    old[2] = fast_trilerp(v, F.xs.f, F.ys.f, F.zu.f);
}

__global__ void __launch_bounds__(FFB_G2P_THREADS, FFB_G2P_FAST_MINB)
    k_g2p_flip_fast(const __grid_constant__ G2PParams P, const __grid_constant__ FastGrid fg) {
    const int j = P.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n || window_skip(P.win, j)) return;
    const float px = P.px[j], py = P.py[j], pz = P.pz[j];
    const float v0 = P.vx[j], v1 = P.vy[j], v2 = P.vz[j];
    const FastAxis xu = fast_axis(px, fg), yu = fast_axis(py, fg), zu = fast_axis(pz, fg);
    float pic[3], old[3];
    if (fast_interior(xu.i, yu.i, zu.i, fg)) {
        const FastAxis xs = fast_shift(xu), ys = fast_shift(yu), zs = fast_shift(zu);
        const int ks = zs.i - fg.kbase, kk = zu.i - fg.kbase;
        const int ou = xu.i + fg.sju * ys.i + fg.sku * ks, ov = xs.i + fg.sjv * yu.i + fg.skv * ks, ow = xs.i + fg.sjw * ys.i + fg.skw * kk;
        float a[8], b[8], c[8];
        load8(P.cur.u + ou, fg.sju, fg.sku, a);
        load8(P.cur.v + ov, fg.sjv, fg.skv, b);
        load8(P.cur.w + ow, fg.sjw, fg.skw, c);
        pic[0] = fast_trilerp(a, xu.f, ys.f, zs.f);
        pic[1] = fast_trilerp(b, xs.f, yu.f, zs.f);
        pic[2] = fast_trilerp(c, xs.f, ys.f, zu.f);
        load8(P.saved.u + ou, fg.sju, fg.sku, a);
        load8(P.saved.v + ov, fg.sjv, fg.skv, b);
        load8(P.saved.w + ow, fg.sjw, fg.skw, c);
        old[0] = fast_trilerp(a, xu.f, ys.f, zs.f);
        old[1] = fast_trilerp(b, xs.f, yu.f, zs.f);
        old[2] = fast_trilerp(c, xs.f, ys.f, zu.f);
    } else {
        g2p_flip_fast_generic(P, fg, px, py, pz, pic, old);
    }
    const float f0 = (v0 + pic[0]) - old[0], f1 = (v1 + pic[1]) - old[1], f2 = (v2 + pic[2]) - old[2];
    P.ovx[j] = pic[0] * P.rp + f0 * P.rf;
    P.ovy[j] = pic[1] * P.rp + f1 * P.rf;
    P.ovz[j] = pic[2] * P.rp + f2 * P.rf;
    P.k1[0][j] = pic[0]; P.k1[1][j] = pic[1]; P.k1[2][j] = pic[2];
}

// FluidSimulation::_getMaximumMarkerParticleSpeed (fluidsimulation.cpp:10188-10202): max over the particles of
// the float dot product v.v (left to right); a max is order independent, so the reduction is bit-exact.
// Non-negative floats order like their bit patterns: integer atomicMax.
__global__ void k_max_speed_sq(const float *__restrict__ vx, const float *__restrict__ vy, const float *__restrict__ vz,
                               const uint32_t *__restrict__ ids, int n, uint32_t *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    float d = 0.0f;
    if (j < n && !(ids[j] & 0x80000000u)) {                    // ghost copies of a z-slab rank belong to the neighbour
        const float x = vx[j], y = vy[j], z = vz[j];
        d = x * x + y * y + z * z;
        if (!(d >= 0.0f)) d = 0.0f;                            // NaN never wins the reference's `distsq > maxsq` either
    }
    uint32_t b = __float_as_uint(d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b) atomicMax(out, b);
}

}  // namespace

int launch_max_speed_sq(Context &c, uint32_t *out_bits) {
    FFB_CUDA(cudaMemsetAsync(out_bits, 0, sizeof(uint32_t), c.stream));
    if (c.n == 0) return 0;
    ParticleSoA &s = c.soa[c.cur];
    k_max_speed_sq<<<(c.n + 255) / 256, 256, 0, c.stream>>>(s.v[0], s.v[1], s.v[2], s.orig, c.n, out_bits);
    FFB_CUDA(cudaGetLastError());
    return 1;
}

int launch_g2p(Context &c, int method, double ratio, int first, int count) {
    if (count < 0) count = c.n - first;
    if (count <= 0) return 0;
    ParticleSoA &s = c.soa[c.cur];
    G2PParams P;
    P.g = c.g;
    P.cur = MacView{c.face[0].vel, c.face[1].vel, c.face[2].vel};
    P.saved = MacView{c.face[0].saved, c.face[1].saved, c.face[2].saved};
    P.px = s.p[0]; P.py = s.p[1]; P.pz = s.p[2];
    ParticleSoA &o = c.nondestructive ? c.soa[c.cur ^ 1] : s;   // spare SoA buffer keeps the inputs pristine
    P.vx = s.v[0]; P.vy = s.v[1]; P.vz = s.v[2];
    P.ovx = o.v[0]; P.ovy = o.v[1]; P.ovz = o.v[2];
    for (int q = 0; q < 9; q++) P.a[q] = o.a[q];
    for (int q = 0; q < 3; q++) P.k1[q] = c.k1s[q];
    P.rp = (float)ratio;
    P.rf = (float)(1 - ratio);
    P.h = (float)(0.5f * c.g.dx);
    P.inv_s = (float)(1.0 / (double)(float)c.g.dx);
    P.invdx = (float)(1.0f / c.g.dx);
    P.first = first;
    P.n = first + count;
    P.win = c.window;
    const int blocks = (count + FFB_G2P_THREADS - 1) / FFB_G2P_THREADS;
    if (c.precision == FFB200_PRECISION_TOLERANCE) {
        const FastGrid fg = make_fast_grid(c.g);
        if (method == FFB200_TRANSFER_APIC)
            k_g2p_apic_fast<<<blocks, FFB_G2P_THREADS, 0, c.stream>>>(P, fg, tolerance_stats(c));
        else
            k_g2p_flip_fast<<<blocks, FFB_G2P_THREADS, 0, c.stream>>>(P, fg);
    } else if (method == FFB200_TRANSFER_APIC)
        k_g2p_apic<<<blocks, FFB_G2P_THREADS, 0, c.stream>>>(P);
    else
        k_g2p_flip<<<blocks, FFB_G2P_THREADS, 0, c.stream>>>(P);
    FFB_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace ffb200
