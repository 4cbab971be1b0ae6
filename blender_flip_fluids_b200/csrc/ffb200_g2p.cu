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

struct G2PParams {
    GridDesc g;
    MacView cur, saved;
    const float *px, *py, *pz;
    const float *vx, *vy, *vz;   // velocity in
    float *ovx, *ovy, *ovz;      // velocity out (same arrays unless the context is non-destructive)
    float *a[9];
    float rp, rf;       // (float)_ratioPICFLIP, (float)(1 - _ratioPICFLIP)
    float h;            // 0.5f * _dx
    float inv_s;        // (float)(1.0 / (float)_dx)   (vec3 / _dx)
    float invdx;        // 1.0f / _dx
    int n;
};

__global__ void __launch_bounds__(256) k_g2p_flip(G2PParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    const float x = P.px[j], y = P.py[j], z = P.pz[j];
    float pic[3], old[3];
    mac_eval(P.g, P.cur, x, y, z, pic[0], pic[1], pic[2]);
    mac_eval(P.g, P.saved, x, y, z, old[0], old[1], old[2]);
    const float v0 = P.vx[j], v1 = P.vy[j], v2 = P.vz[j];
    // vFLIP = vel + vPIC - saved(p); v = r*vPIC + (1-r)*vFLIP   (:6779-6781)
    const float f0 = (v0 + pic[0]) - old[0], f1 = (v1 + pic[1]) - old[1], f2 = (v2 + pic[2]) - old[2];
    P.ovx[j] = pic[0] * P.rp + f0 * P.rf;
    P.ovy[j] = pic[1] * P.rp + f1 * P.rf;
    P.ovz[j] = pic[2] * P.rp + f2 * P.rf;
}

// affineDir = sum over the 8 faces around the (staggered) particle of gradWeight * Face(g).
template <int DIR>
__device__ __forceinline__ void apic_affine(const G2PParams &P, const float *__restrict__ f, float px, float py, float pz,
                                            float &ox, float &oy, float &oz) {
    const GridDesc &g = P.g;
    const int gw = g.I + (DIR == 0), gh = g.J + (DIR == 1), gd = g.K + (DIR == 2);
    const float x = px - (DIR == 0 ? 0.0f : P.h), y = py - (DIR == 1 ? 0.0f : P.h), z = pz - (DIR == 2 ? 0.0f : P.h);
    const int gi = pos2idx(x, g.inv_dx), gj = pos2idx(y, g.inv_dx), gk = pos2idx(z, g.inv_dx);
    const float ix = (x - idx2posf(gi, g.dx)) * P.inv_s;
    const float iy = (y - idx2posf(gj, g.dx)) * P.inv_s;
    const float iz = (z - idx2posf(gk, g.dx)) * P.inv_s;
    const float invdx = P.invdx;
    const float mx = 1.0f - ix, my = 1.0f - iy, mz = 1.0f - iz;
    // gradient weights in the reference's operand order (fluidsimulation.cpp:6737-6768)
    float w[8][3];
    w[0][0] = -invdx * my * mz;      w[0][1] = -invdx * mx * mz;      w[0][2] = -invdx * mx * my;
    w[1][0] = invdx * my * mz;       w[1][1] = ix * (-invdx) * mz;    w[1][2] = ix * my * (-invdx);
    w[2][0] = (-invdx) * iy * mz;    w[2][1] = mx * invdx * mz;       w[2][2] = mx * iy * (-invdx);
    w[3][0] = invdx * iy * mz;       w[3][1] = ix * invdx * mz;       w[3][2] = ix * iy * (-invdx);
    w[4][0] = (-invdx) * my * iz;    w[4][1] = mx * (-invdx) * iz;    w[4][2] = mx * my * invdx;
    w[5][0] = invdx * my * iz;       w[5][1] = ix * (-invdx) * iz;    w[5][2] = ix * my * invdx;
    w[6][0] = (-invdx) * iy * iz;    w[6][1] = mx * invdx * iz;       w[6][2] = mx * iy * invdx;
    w[7][0] = invdx * iy * iz;       w[7][1] = ix * invdx * iz;       w[7][2] = ix * iy * invdx;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    const long long sj = gw, sk = (long long)gw * gh;
    const long long base = (long long)gi + sj * gj + sk * (long long)(gk - g.kbase);
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int di = c & 1, dj = (c >> 1) & 1, dk = (c >> 2) & 1;
        if (!in_range3(gi + di, gj + dj, gk + dk, gw, gh, gd)) continue;
        const float fv = __ldg(f + base + di + sj * dj + sk * dk);
        sx += w[c][0] * fv;
        sy += w[c][1] * fv;
        sz += w[c][2] * fv;
    }
    ox = sx; oy = sy; oz = sz;
}

__global__ void __launch_bounds__(256) k_g2p_apic(G2PParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    const float x = P.px[j], y = P.py[j], z = P.pz[j];
    float ax, ay, az;
    apic_affine<0>(P, P.cur.u, x, y, z, ax, ay, az);
    P.a[0][j] = ax; P.a[1][j] = ay; P.a[2][j] = az;
    apic_affine<1>(P, P.cur.v, x, y, z, ax, ay, az);
    P.a[3][j] = ax; P.a[4][j] = ay; P.a[5][j] = az;
    apic_affine<2>(P, P.cur.w, x, y, z, ax, ay, az);
    P.a[6][j] = ax; P.a[7][j] = ay; P.a[8][j] = az;
    float v0, v1, v2;
    mac_eval(P.g, P.cur, x, y, z, v0, v1, v2);
    P.ovx[j] = v0; P.ovy[j] = v1; P.ovz[j] = v2;
}

}  // namespace

int launch_g2p(Context &c, int method, double ratio) {
    if (c.n == 0) return 0;
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
    P.rp = (float)ratio;
    P.rf = (float)(1 - ratio);
    P.h = (float)(0.5f * c.g.dx);
    P.inv_s = (float)(1.0 / (double)(float)c.g.dx);
    P.invdx = (float)(1.0f / c.g.dx);
    P.n = c.n;
    const int blocks = (c.n + 255) / 256;
    if (method == FFB200_TRANSFER_APIC)
        k_g2p_apic<<<blocks, 256, 0, c.stream>>>(P);
    else
        k_g2p_flip<<<blocks, 256, 0, c.stream>>>(P);
    FFB_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace ffb200
