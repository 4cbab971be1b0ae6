/* oracle/flip_oracle.c -- TEST INFRASTRUCTURE, not product code (see flip_oracle.h).
 *
 * Single-threaded restatement of the reference hot path, written so that every float and
 * double operation happens in the same order and precision as in the reference sources
 * (cited per function, paths relative to /root/reference/src/engine). Build with
 * -ffp-contract=off on baseline x86-64 (oracle/Makefile): scalar SSE2, no FMA, exactly like
 * the reference build. The reference sums a face's contributions in ascending particle
 * index inside the face's 10^3 block (velocityadvector.cpp:383-413, probe-verified to be
 * thread-count independent), which is what the particle-major loops below reproduce.
 */
#include "flip_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define CHUNK 10 /* VelocityAdvector::_chunkWidth, velocityadvector.h:187 */

/* ---- Grid3d index maths (grid3d.h:32-82) ---------------------------------------------- */

/* positionToGridIndex: (int)floor(x * (1.0/dx)), all in double. */
static inline int pos2idx(double x, double dx) {
    double invdx = 1.0 / dx;
    return (int)floor(x * invdx);
}

/* GridIndexToPosition(vec3 flavour): (float)i*dx is evaluated in double, the vec3
 * constructor narrows it to float (grid3d.h:80-82, vmath.cpp:35). */
static inline float idx2posf(int i, double dx) { return (float)((double)(float)i * dx); }

static inline int in_range(int i, int j, int k, int w, int h, int d) {
    return i >= 0 && j >= 0 && k >= 0 && i < w && j < h && k < d;
}

static inline size_t flat(int i, int j, int k, int w, int h) {
    return (size_t)i + (size_t)w * ((size_t)j + (size_t)h * (size_t)k);
}

/* ---- cell binning + stable sort -------------------------------------------------------- */

typedef struct { uint32_t key, idx; } keyidx;

static int cmp_keyidx(const void *a, const void *b) {
    const keyidx *x = (const keyidx *)a, *y = (const keyidx *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

void flip_oracle_bin_sort(int I, int J, int K, double dx, int n, const float *pos,
                          int32_t *cell, uint32_t *hkey, uint32_t *perm) {
    const int A = 4;                                          /* half-cell apron (2 cells per side) */
    double inv2 = 2.0 * (1.0 / dx);
    uint32_t HX = 2u * (uint32_t)I + 2u * A, HY = 2u * (uint32_t)J + 2u * A, HZ = 2u * (uint32_t)K + 2u * A;
    uint32_t sentinel = HX * HY * HZ;
    keyidx *ki = (keyidx *)malloc(sizeof(keyidx) * (size_t)(n > 0 ? n : 1));
    for (int p = 0; p < n; p++) {
        int ci = pos2idx(pos[3 * p + 0], dx), cj = pos2idx(pos[3 * p + 1], dx), ck = pos2idx(pos[3 * p + 2], dx);
        int ok = in_range(ci, cj, ck, I, J, K);
        if (cell) cell[p] = ok ? (int32_t)(ci + I * (cj + J * ck)) : -1;
        int hi = (int)floor((double)pos[3 * p + 0] * inv2) + A;
        int hj = (int)floor((double)pos[3 * p + 1] * inv2) + A;
        int hk = (int)floor((double)pos[3 * p + 2] * inv2) + A;
        uint32_t key = sentinel;
        if (in_range(hi, hj, hk, (int)HX, (int)HY, (int)HZ))
            key = (uint32_t)hi + HX * ((uint32_t)hj + HY * (uint32_t)hk);
        if (hkey) hkey[p] = key;
        ki[p].key = key;
        ki[p].idx = (uint32_t)p;
    }
    if (perm) {
        qsort(ki, (size_t)n, sizeof(keyidx), cmp_keyidx);
        for (int p = 0; p < n; p++) perm[p] = ki[p].idx;
    }
    free(ki);
}

/* ---- P2G -------------------------------------------------------------------------------- */

typedef struct {
    int gi, gj, gk;      /* face-grid dimensions for this direction */
    int bi, bj, bk;      /* block-grid dimensions, blockarray3d.h:66-70 */
    float off[3];        /* _getDirectionOffset, velocityadvector.cpp:177-188 */
    float *scalar;       /* running sum(w*v) per face */
    float *weight;       /* running sum(w) per face */
    uint8_t *active;     /* active 10^3 blocks, velocityadvector.cpp:190-251 */
} p2g_dir;

static void p2g_active_blocks(p2g_dir *d, double dx, int n, const float *pos) {
    double chunkdx = dx * CHUNK;                              /* velocityadvector.cpp:53 (double) */
    size_t nb = (size_t)d->bi * d->bj * d->bk;
    uint8_t *home = (uint8_t *)calloc(nb, 1);
    for (int p = 0; p < n; p++) {                             /* velocityadvector.cpp:240-251 */
        float x = pos[3 * p + 0] - d->off[0], y = pos[3 * p + 1] - d->off[1], z = pos[3 * p + 2] - d->off[2];
        int gi = pos2idx(x, chunkdx), gj = pos2idx(y, chunkdx), gk = pos2idx(z, chunkdx);
        if (in_range(gi, gj, gk, d->bi, d->bj, d->bk)) home[flat(gi, gj, gk, d->bi, d->bj)] = 1;
    }
    memcpy(d->active, home, nb);
    for (int k = 0; k < d->bk; k++)                           /* featherGrid26, gridutils.cpp:264-297 */
        for (int j = 0; j < d->bj; j++)
            for (int i = 0; i < d->bi; i++) {
                if (!home[flat(i, j, k, d->bi, d->bj)]) continue;
                for (int c = -1; c <= 1; c++)
                    for (int b = -1; b <= 1; b++)
                        for (int a = -1; a <= 1; a++)
                            if (in_range(i + a, j + b, k + c, d->bi, d->bj, d->bk))
                                d->active[flat(i + a, j + b, k + c, d->bi, d->bj)] = 1;
            }
    free(home);
}

typedef struct { float r, sr, rsq, coef1, coef2, coef3; } flip_kernel;

/* One particle splatted into one block with the spherical kernel,
 * _advectionFLIPProducerThread, velocityadvector.cpp:467-536. (px,py,pz) is the offset
 * particle position, still in the global frame. */
static void splat_flip(p2g_dir *d, double dx, const flip_kernel *fk, int bi, int bj, int bk,
                       float px, float py, float pz, float velocity) {
    double chunk = CHUNK * dx;                                /* _chunkWidth * _dx: int * double */
    float x = px - idx2posf(bi, chunk), y = py - idx2posf(bj, chunk), z = pz - idx2posf(bk, chunk);
    float sr = fk->sr;
    int i0 = pos2idx(x - sr, dx), j0 = pos2idx(y - sr, dx), k0 = pos2idx(z - sr, dx);
    int i1 = pos2idx(x + sr, dx), j1 = pos2idx(y + sr, dx), k1 = pos2idx(z + sr, dx);
    if (i0 < 0) i0 = 0;
    if (j0 < 0) j0 = 0;
    if (k0 < 0) k0 = 0;
    if (i1 > CHUNK - 1) i1 = CHUNK - 1;
    if (j1 > CHUNK - 1) j1 = CHUNK - 1;
    if (k1 > CHUNK - 1) k1 = CHUNK - 1;
    for (int k = k0; k <= k1; k++)
        for (int j = j0; j <= j1; j++)
            for (int i = i0; i <= i1; i++) {
                float vx = idx2posf(i, dx) - x, vy = idx2posf(j, dx) - y, vz = idx2posf(k, dx) - z;
                float d2 = vx * vx + vy * vy + vz * vz;
                if (d2 < fk->rsq) {
                    float wgt = 1.0f - fk->coef1 * d2 * d2 * d2 + fk->coef2 * d2 * d2 - fk->coef3 * d2;
                    int ni = bi * CHUNK + i, nj = bj * CHUNK + j, nk = bk * CHUNK + k;
                    if (!in_range(ni, nj, nk, d->gi, d->gj, d->gk)) continue;   /* write-out drops these, :155 */
                    size_t f = flat(ni, nj, nk, d->gi, d->gj);
                    d->scalar[f] += wgt * velocity;
                    d->weight[f] += wgt;
                }
            }
}

/* One particle splatted into one block with trilinear weights + affine term,
 * _advectionAPICProducerThread, velocityadvector.cpp:543-623. Nodes outside the block are
 * skipped (:596-599): this is the block-seam drop of SURVEY.md section 0.5(iii). */
static void splat_apic(p2g_dir *d, double dx, int bi, int bj, int bk,
                       float px, float py, float pz, float velocity, const float *aff) {
    double chunk = CHUNK * dx;
    float x = px - idx2posf(bi, chunk), y = py - idx2posf(bj, chunk), z = pz - idx2posf(bk, chunk);
    int gi = pos2idx(x, dx), gj = pos2idx(y, dx), gk = pos2idx(z, dx);
    float s = (float)dx;                                     /* vec3 / _dx: vmath.cpp:100-103 */
    float inv = (float)(1.0 / (double)s);
    float ix = (x - idx2posf(gi, dx)) * inv, iy = (y - idx2posf(gj, dx)) * inv, iz = (z - idx2posf(gk, dx)) * inv;
    float wts[8];
    wts[0] = (1.0f - ix) * (1.0f - iy) * (1.0f - iz);
    wts[1] = ix * (1.0f - iy) * (1.0f - iz);
    wts[2] = (1.0f - ix) * iy * (1.0f - iz);
    wts[3] = ix * iy * (1.0f - iz);
    wts[4] = (1.0f - ix) * (1.0f - iy) * iz;
    wts[5] = ix * (1.0f - iy) * iz;
    wts[6] = (1.0f - ix) * iy * iz;
    wts[7] = ix * iy * iz;
    for (int c = 0; c < 8; c++) {
        int i = gi + (c & 1), j = gj + ((c >> 1) & 1), k = gk + ((c >> 2) & 1);
        if (i < 0 || j < 0 || k < 0 || i >= CHUNK || j >= CHUNK || k >= CHUNK) continue;
        float dxn = idx2posf(i, dx) - x, dyn = idx2posf(j, dx) - y, dzn = idx2posf(k, dx) - z;
        float apic = aff[0] * dxn + aff[1] * dyn + aff[2] * dzn;
        float wgt = wts[c];
        int ni = bi * CHUNK + i, nj = bj * CHUNK + j, nk = bk * CHUNK + k;
        if (!in_range(ni, nj, nk, d->gi, d->gj, d->gk)) continue;
        size_t f = flat(ni, nj, nk, d->gi, d->gj);
        d->scalar[f] += wgt * (velocity + apic);
        d->weight[f] += wgt;
    }
}

static void p2g_direction_n(int dir, int I, int J, int K, double dx, double radius, int method, int n,
                            const float *pos, const float *vel, const float *aff,
                            float *out, uint8_t *valid, float *wsum, int vec3_norm, int stride);

static void p2g_direction(int dir, int I, int J, int K, double dx, double radius, int method, int n,
                          const float *pos, const float *vel, const float *aff,
                          float *out, uint8_t *valid, float *wsum) {
    p2g_direction_n(dir, I, J, K, dx, radius, method, n, pos, vel, aff, out, valid, wsum, 0, 1);
}

/* vec3_norm: normalise with vmath's vec3 /= float (multiply by float(1.0 / w), vmath.cpp:105-111) instead of a float
 * division; stride: distance in floats between the scalar payloads of consecutive particles (dir 3 only) and between
 * consecutive outputs (one channel of an interleaved vec3 grid). */
static void p2g_direction_n(int dir, int I, int J, int K, double dx, double radius, int method, int n,
                            const float *pos, const float *vel, const float *aff,
                            float *out, uint8_t *valid, float *wsum, int vec3_norm, int stride) {
    p2g_dir d;
    d.gi = I + (dir == 0);
    d.gj = J + (dir == 1);
    d.gk = K + (dir == 2);
    d.bi = (d.gi + CHUNK - 1) / CHUNK;
    d.bj = (d.gj + CHUNK - 1) / CHUNK;
    d.bk = (d.gk + CHUNK - 1) / CHUNK;
    float h = (float)(0.5 * dx);                              /* vec3(0.0, 0.5*_dx, 0.5*_dx) narrows */
    d.off[0] = d.off[1] = d.off[2] = h;
    if (dir < 3) d.off[dir] = 0.0f;                           /* dir 3: cell-centred scalar grid, offset (h, h, h) */
    size_t nf = (size_t)d.gi * d.gj * d.gk;
    d.scalar = (float *)calloc(nf, sizeof(float));
    d.weight = (float *)calloc(nf, sizeof(float));
    d.active = (uint8_t *)calloc((size_t)d.bi * d.bj * d.bk, 1);
    p2g_active_blocks(&d, dx, n, pos);

    float eps = 1e-6;                                         /* float eps = 1e-6; */
    flip_kernel fk;
    fk.r = (float)radius;                                     /* float r = _particleRadius; :472 */
    fk.sr = (float)(radius + (double)eps);                    /* float sr = _particleRadius + eps; */
    fk.rsq = fk.r * fk.r;
    fk.coef1 = (4.0f / 9.0f) * (1.0f / (fk.r * fk.r * fk.r * fk.r * fk.r * fk.r));
    fk.coef2 = (17.0f / 9.0f) * (1.0f / (fk.r * fk.r * fk.r * fk.r));
    fk.coef3 = (22.0f / 9.0f) * (1.0f / (fk.r * fk.r));

    double chunkdx = dx * CHUNK;
    float sr = fk.sr;
    float blockdx = (float)chunkdx;                           /* float blockdx = _chunkdx; :307 */
    for (int p = 0; p < n; p++) {                             /* _computeGridCountDataThread :309-352 */
        float x = pos[3 * p + 0] - d.off[0], y = pos[3 * p + 1] - d.off[1], z = pos[3 * p + 2] - d.off[2];
        int b0 = pos2idx(x, blockdx), b1 = pos2idx(y, blockdx), b2 = pos2idx(z, blockdx);
        float bx = idx2posf(b0, blockdx), by = idx2posf(b1, blockdx), bz = idx2posf(b2, blockdx);
        int lo[3], hi[3];
        if (x - sr > bx && y - sr > by && z - sr > bz &&
            x + sr < bx + blockdx && y + sr < by + blockdx && z + sr < bz + blockdx) {
            lo[0] = hi[0] = b0;
            lo[1] = hi[1] = b1;
            lo[2] = hi[2] = b2;
        } else {
            lo[0] = pos2idx(x - sr, blockdx); lo[1] = pos2idx(y - sr, blockdx); lo[2] = pos2idx(z - sr, blockdx);
            hi[0] = pos2idx(x + sr, blockdx); hi[1] = pos2idx(y + sr, blockdx); hi[2] = pos2idx(z + sr, blockdx);
        }
        float velocity = dir < 3 ? vel[3 * p + dir] : vel[(size_t)stride * p]; /* dir 3: one scalar attribute per particle */
        for (int bk = lo[2]; bk <= hi[2]; bk++)
            for (int bj = lo[1]; bj <= hi[1]; bj++)
                for (int bi = lo[0]; bi <= hi[0]; bi++) {
                    if (!in_range(bi, bj, bk, d.bi, d.bj, d.bk)) continue;      /* getBlockID == -1 */
                    if (!d.active[flat(bi, bj, bk, d.bi, d.bj)]) continue;
                    if (method == FLIP_ORACLE_APIC) splat_apic(&d, dx, bi, bj, bk, x, y, z, velocity, aff + 3 * (size_t)p);
                    else splat_flip(&d, dx, &fk, bi, bj, bk, x, y, z, velocity);
                }
    }
    for (size_t f = 0; f < nf; f++) {                         /* normalise :527-531, write-out :140-168 */
        float s = d.scalar[f], wt = d.weight[f];
        if (wt > eps) {
            if (vec3_norm) s *= (float)(1.0 / (double)wt);
            else s /= wt;
        }
        out[(size_t)stride * f] = s;
        valid[f] = wt > eps ? 1 : 0;
        if (wsum) wsum[f] = wt;
    }
    free(d.scalar);
    free(d.weight);
    free(d.active);
}

void flip_oracle_p2g_w(int I, int J, int K, double dx, double radius, int method, int n,
                       const float *pos, const float *vel,
                       const float *affx, const float *affy, const float *affz,
                       float *u, float *v, float *w,
                       uint8_t *validu, uint8_t *validv, uint8_t *validw,
                       float *wsumu, float *wsumv, float *wsumw) {
    p2g_direction(0, I, J, K, dx, radius, method, n, pos, vel, affx, u, validu, wsumu);
    p2g_direction(1, I, J, K, dx, radius, method, n, pos, vel, affy, v, validv, wsumv);
    p2g_direction(2, I, J, K, dx, radius, method, n, pos, vel, affz, w, validw, wsumw);
}

void flip_oracle_p2g(int I, int J, int K, double dx, double radius, int method, int n,
                     const float *pos, const float *vel,
                     const float *affx, const float *affy, const float *affz,
                     float *u, float *v, float *w,
                     uint8_t *validu, uint8_t *validv, uint8_t *validw) {
    flip_oracle_p2g_w(I, J, K, dx, radius, method, n, pos, vel, affx, affy, affz, u, v, w,
                      validu, validv, validw, NULL, NULL, NULL);
}

/* ---- MAC gather (macvelocityfield.cpp:519-645, interpolation.cpp:61-70) ------------------ */

typedef struct {
    int I, J, K;
    double dx;
    const float *u, *v, *w;
} macfield;

static inline int pos_in_grid(double x, double y, double z, double dx, int i, int j, int k) {
    return x >= 0 && y >= 0 && z >= 0 && x < dx * i && y < dx * j && z < dx * k;   /* grid3d.h:134-136 */
}

static inline double trilerp(const double p[8], double x, double y, double z) {
    return p[0] * (1 - x) * (1 - y) * (1 - z) +
           p[1] * x * (1 - y) * (1 - z) +
           p[2] * (1 - x) * y * (1 - z) +
           p[3] * (1 - x) * (1 - y) * z +
           p[4] * x * (1 - y) * z +
           p[5] * (1 - x) * y * z +
           p[6] * x * y * (1 - z) +
           p[7] * x * y * z;
}

/* comp 0/1/2 = _interpolateLinearU/V/W. */
static double mac_lerp(const macfield *m, int comp, double x, double y, double z) {
    if (!pos_in_grid(x, y, z, m->dx, m->I, m->J, m->K)) return 0.0;
    const float *g = comp == 0 ? m->u : (comp == 1 ? m->v : m->w);
    int gw = m->I + (comp == 0), gh = m->J + (comp == 1), gd = m->K + (comp == 2);
    if (comp != 0) x -= 0.5 * m->dx;
    if (comp != 1) y -= 0.5 * m->dx;
    if (comp != 2) z -= 0.5 * m->dx;
    int i = pos2idx(x, m->dx), j = pos2idx(y, m->dx), k = pos2idx(z, m->dx);
    double gx = (double)i * m->dx, gy = (double)j * m->dx, gz = (double)k * m->dx;
    double inv_dx = 1 / m->dx;
    double ix = (x - gx) * inv_dx, iy = (y - gy) * inv_dx, iz = (z - gz) * inv_dx;
    double pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};                  /* _outOfRangeVector defaults to 0 */
    if (in_range(i, j, k, gw, gh, gd)) pt[0] = g[flat(i, j, k, gw, gh)];
    if (in_range(i + 1, j, k, gw, gh, gd)) pt[1] = g[flat(i + 1, j, k, gw, gh)];
    if (in_range(i, j + 1, k, gw, gh, gd)) pt[2] = g[flat(i, j + 1, k, gw, gh)];
    if (in_range(i, j, k + 1, gw, gh, gd)) pt[3] = g[flat(i, j, k + 1, gw, gh)];
    if (in_range(i + 1, j, k + 1, gw, gh, gd)) pt[4] = g[flat(i + 1, j, k + 1, gw, gh)];
    if (in_range(i, j + 1, k + 1, gw, gh, gd)) pt[5] = g[flat(i, j + 1, k + 1, gw, gh)];
    if (in_range(i + 1, j + 1, k, gw, gh, gd)) pt[6] = g[flat(i + 1, j + 1, k, gw, gh)];
    if (in_range(i + 1, j + 1, k + 1, gw, gh, gd)) pt[7] = g[flat(i + 1, j + 1, k + 1, gw, gh)];
    return trilerp(pt, ix, iy, iz);
}

/* evaluateVelocityAtPositionLinear(vec3): float position widened to double, result narrowed. */
static void mac_eval(const macfield *m, const float p[3], float out[3]) {
    double x = p[0], y = p[1], z = p[2];
    if (!pos_in_grid(x, y, z, m->dx, m->I, m->J, m->K)) {
        out[0] = out[1] = out[2] = 0.0f;
        return;
    }
    out[0] = (float)mac_lerp(m, 0, x, y, z);
    out[1] = (float)mac_lerp(m, 1, x, y, z);
    out[2] = (float)mac_lerp(m, 2, x, y, z);
}

void flip_oracle_mac_sample(int I, int J, int K, double dx, int n, const float *pos, float *out,
                            const float *u, const float *v, const float *w) {
    macfield m = {I, J, K, dx, u, v, w};
    for (int p = 0; p < n; p++) mac_eval(&m, pos + 3 * (size_t)p, out + 3 * (size_t)p);
}

/* ---- G2P ------------------------------------------------------------------------------------ */

void flip_oracle_g2p_flip(int I, int J, int K, double dx, int n, const float *pos, float *vel,
                          const float *u, const float *v, const float *w,
                          const float *su, const float *sv, const float *sw, double ratio) {
    macfield cur = {I, J, K, dx, u, v, w}, saved = {I, J, K, dx, su, sv, sw};
    float rp = (float)ratio, rf = (float)(1 - ratio);         /* fluidsimulation.cpp:6781 */
    for (int p = 0; p < n; p++) {
        float pic[3], old[3];
        mac_eval(&cur, pos + 3 * (size_t)p, pic);
        mac_eval(&saved, pos + 3 * (size_t)p, old);
        for (int c = 0; c < 3; c++) {
            float flipv = (vel[3 * (size_t)p + c] + pic[c]) - old[c];
            vel[3 * (size_t)p + c] = pic[c] * rp + flipv * rf;   /* operator*(float, vec3): v.x*s */
        }
    }
}

/* _getIndicesAndGradientWeights (fluidsimulation.cpp:6709-6769) + the accumulation loop of
 * :6813-6837 for one direction. */
static void apic_affine_dir(int dir, int I, int J, int K, double dx, const float *p, const float *g, float out[3]) {
    int gw = I + (dir == 0), gh = J + (dir == 1), gd = K + (dir == 2);
    float h = 0.5f * dx;                                      /* float h = 0.5f * _dx; */
    float off[3] = {h, h, h};
    off[dir] = 0.0f;
    float x = p[0] - off[0], y = p[1] - off[1], z = p[2] - off[2];
    int gi = pos2idx(x, dx), gj = pos2idx(y, dx), gk = pos2idx(z, dx);
    float s = (float)dx;
    float inv = (float)(1.0 / (double)s);                     /* vec3 / _dx */
    float ix = (x - idx2posf(gi, dx)) * inv, iy = (y - idx2posf(gj, dx)) * inv, iz = (z - idx2posf(gk, dx)) * inv;
    float invdx = 1.0f / dx;                                  /* float invdx = 1.0f / _dx; */
    float wt[8][3];
    wt[0][0] = -invdx * (1.0f - iy) * (1.0f - iz);
    wt[0][1] = -invdx * (1.0f - ix) * (1.0f - iz);
    wt[0][2] = -invdx * (1.0f - ix) * (1.0f - iy);
    wt[1][0] = invdx * (1.0f - iy) * (1.0f - iz);
    wt[1][1] = ix * (-invdx) * (1.0f - iz);
    wt[1][2] = ix * (1.0f - iy) * (-invdx);
    wt[2][0] = (-invdx) * iy * (1.0f - iz);
    wt[2][1] = (1.0f - ix) * invdx * (1.0f - iz);
    wt[2][2] = (1.0f - ix) * iy * (-invdx);
    wt[3][0] = invdx * iy * (1.0f - iz);
    wt[3][1] = ix * invdx * (1.0f - iz);
    wt[3][2] = ix * iy * (-invdx);
    wt[4][0] = (-invdx) * (1.0f - iy) * iz;
    wt[4][1] = (1.0f - ix) * (-invdx) * iz;
    wt[4][2] = (1.0f - ix) * (1.0f - iy) * invdx;
    wt[5][0] = invdx * (1.0f - iy) * iz;
    wt[5][1] = ix * (-invdx) * iz;
    wt[5][2] = ix * (1.0f - iy) * invdx;
    wt[6][0] = (-invdx) * iy * iz;
    wt[6][1] = (1.0f - ix) * invdx * iz;
    wt[6][2] = (1.0f - ix) * iy * invdx;
    wt[7][0] = invdx * iy * iz;
    wt[7][1] = ix * invdx * iz;
    wt[7][2] = ix * iy * invdx;
    out[0] = out[1] = out[2] = 0.0f;
    for (int c = 0; c < 8; c++) {
        int i = gi + (c & 1), j = gj + ((c >> 1) & 1), k = gk + ((c >> 2) & 1);
        if (!in_range(i, j, k, gw, gh, gd)) continue;
        float f = g[flat(i, j, k, gw, gh)];
        out[0] += wt[c][0] * f;
        out[1] += wt[c][1] * f;
        out[2] += wt[c][2] * f;
    }
}

void flip_oracle_g2p_apic(int I, int J, int K, double dx, int n, const float *pos, float *vel,
                          float *affx, float *affy, float *affz,
                          const float *u, const float *v, const float *w) {
    macfield cur = {I, J, K, dx, u, v, w};
    for (int p = 0; p < n; p++) {
        const float *pp = pos + 3 * (size_t)p;
        apic_affine_dir(0, I, J, K, dx, pp, u, affx + 3 * (size_t)p);
        apic_affine_dir(1, I, J, K, dx, pp, v, affy + 3 * (size_t)p);
        apic_affine_dir(2, I, J, K, dx, pp, w, affz + 3 * (size_t)p);
        mac_eval(&cur, pp, vel + 3 * (size_t)p);
    }
}

/* ---- advection ------------------------------------------------------------------------------ */

typedef struct {
    float px, py, pz;         /* AABB::position (vec3 of floats) */
    double w, h, d;           /* AABB::width/height/depth (doubles), aabb.h:60-63 */
} aabb;

static void aabb_expand(aabb *b, double v) {                 /* aabb.cpp:122-128 */
    double hh = 0.5 * v;
    float hf = (float)hh;                                     /* vec3(h,h,h) narrows */
    b->px -= hf; b->py -= hf; b->pz -= hf;
    b->w += v; b->h += v; b->d += v;
}

static int aabb_inside(const aabb *b, const float p[3]) {    /* aabb.cpp:130-133 */
    return p[0] >= b->px && p[1] >= b->py && p[2] >= b->pz &&
           p[0] < b->px + b->w && p[1] < b->py + b->h && p[2] < b->pz + b->d;
}

static void aabb_nearest_inside(const aabb *b, float p[3]) { /* aabb.cpp:493-518, eps = 1e-6 */
    if (aabb_inside(b, p)) return;
    double eps = 1e-6;
    float mx = b->px + (float)b->w, my = b->py + (float)b->h, mz = b->pz + (float)b->d;   /* getMaxPoint */
    p[0] = fmax(p[0], b->px); p[1] = fmax(p[1], b->py); p[2] = fmax(p[2], b->pz);
    p[0] = fmin(p[0], mx - eps); p[1] = fmin(p[1], my - eps); p[2] = fmin(p[2], mz - eps);
}

typedef struct {
    int I, J, K;
    double dx;
    const float *phi;         /* (I+1)(J+1)(K+1) */
    const uint8_t *near;
    int ni, nj, nk;
    double near_dx;
    double cfl;
    aabb boundary;
} solid;

/* Interpolation::trilinearInterpolate(vec3, dx, Array3d<float>&), interpolation.cpp:72-112:
 * cell origin is a float vec3, the offset is taken in float and scaled in double. */
static float sdf_sample(const solid *s, const float p[3]) {
    int w = s->I + 1, h = s->J + 1, d = s->K + 1;
    int i = pos2idx(p[0], s->dx), j = pos2idx(p[1], s->dx), k = pos2idx(p[2], s->dx);
    double inv_dx = 1.0 / s->dx;
    double ix = (p[0] - idx2posf(i, s->dx)) * inv_dx;
    double iy = (p[1] - idx2posf(j, s->dx)) * inv_dx;
    double iz = (p[2] - idx2posf(k, s->dx)) * inv_dx;
    double pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float *g = s->phi;
    if (in_range(i, j, k, w, h, d)) pt[0] = g[flat(i, j, k, w, h)];
    if (in_range(i + 1, j, k, w, h, d)) pt[1] = g[flat(i + 1, j, k, w, h)];
    if (in_range(i, j + 1, k, w, h, d)) pt[2] = g[flat(i, j + 1, k, w, h)];
    if (in_range(i, j, k + 1, w, h, d)) pt[3] = g[flat(i, j, k + 1, w, h)];
    if (in_range(i + 1, j, k + 1, w, h, d)) pt[4] = g[flat(i + 1, j, k + 1, w, h)];
    if (in_range(i, j + 1, k + 1, w, h, d)) pt[5] = g[flat(i, j + 1, k + 1, w, h)];
    if (in_range(i + 1, j + 1, k, w, h, d)) pt[6] = g[flat(i + 1, j + 1, k, w, h)];
    if (in_range(i + 1, j + 1, k + 1, w, h, d)) pt[7] = g[flat(i + 1, j + 1, k + 1, w, h)];
    return (float)trilerp(pt, ix, iy, iz);
}

static inline double bilerp(double v00, double v10, double v01, double v11, double ix, double iy) {
    double l1 = (1 - ix) * v00 + ix * v10;                    /* interpolation.cpp:189-195 */
    double l2 = (1 - ix) * v01 + ix * v11;
    return (1 - iy) * l1 + iy * l2;
}

/* Interpolation::trilinearInterpolateGradient, interpolation.cpp:197-259: corner differences
 * in float, bilinear blend in double, narrowed to float. Not divided by dx. */
static void sdf_gradient(const solid *s, const float p[3], float grad[3]) {
    int w = s->I + 1, h = s->J + 1, d = s->K + 1;
    int i = pos2idx(p[0], s->dx), j = pos2idx(p[1], s->dx), k = pos2idx(p[2], s->dx);
    double inv_dx = 1.0 / s->dx;
    double ix = (p[0] - idx2posf(i, s->dx)) * inv_dx;
    double iy = (p[1] - idx2posf(j, s->dx)) * inv_dx;
    double iz = (p[2] - idx2posf(k, s->dx)) * inv_dx;
    const float *g = s->phi;
    float v000 = 0, v001 = 0, v010 = 0, v011 = 0, v100 = 0, v101 = 0, v110 = 0, v111 = 0;
    if (in_range(i, j, k, w, h, d)) v000 = g[flat(i, j, k, w, h)];
    if (in_range(i + 1, j, k, w, h, d)) v100 = g[flat(i + 1, j, k, w, h)];
    if (in_range(i, j + 1, k, w, h, d)) v010 = g[flat(i, j + 1, k, w, h)];
    if (in_range(i, j, k + 1, w, h, d)) v001 = g[flat(i, j, k + 1, w, h)];
    if (in_range(i + 1, j, k + 1, w, h, d)) v101 = g[flat(i + 1, j, k + 1, w, h)];
    if (in_range(i, j + 1, k + 1, w, h, d)) v011 = g[flat(i, j + 1, k + 1, w, h)];
    if (in_range(i + 1, j + 1, k, w, h, d)) v110 = g[flat(i + 1, j + 1, k, w, h)];
    if (in_range(i + 1, j + 1, k + 1, w, h, d)) v111 = g[flat(i + 1, j + 1, k + 1, w, h)];
    float ddx00 = v100 - v000, ddx10 = v110 - v010, ddx01 = v101 - v001, ddx11 = v111 - v011;
    grad[0] = (float)bilerp(ddx00, ddx10, ddx01, ddx11, iy, iz);
    float ddy00 = v010 - v000, ddy10 = v110 - v100, ddy01 = v011 - v001, ddy11 = v111 - v101;
    grad[1] = (float)bilerp(ddy00, ddy10, ddy01, ddy11, ix, iz);
    float ddz00 = v001 - v000, ddz10 = v101 - v100, ddz01 = v011 - v010, ddz11 = v111 - v110;
    grad[2] = (float)bilerp(ddz00, ddz10, ddz01, ddz11, ix, iy);
}

static inline float vlen(float x, float y, float z) {        /* vmath::length, vmath.h:89-91 */
    return sqrtf(x * x + y * y + z * z);
}

static int near_solid(const solid *s, const float p[3]) {
    int i = pos2idx(p[0], s->near_dx), j = pos2idx(p[1], s->near_dx), k = pos2idx(p[2], s->near_dx);
    /* The reference reads Array3d<bool> unchecked here (fluidsimulation.cpp:7654-7656); marker
     * particles are always inside the boundary so the index is always in range. Out-of-range
     * (undefined in the reference) is treated as "near". */
    if (!in_range(i, j, k, s->ni, s->nj, s->nk)) return 1;
    return s->near[flat(i, j, k, s->ni, s->nj)] != 0;
}

/* FluidSimulation::_resolveCollision, fluidsimulation.cpp:7646-7721. */
static void resolve_collision(const solid *s, const float oldp[3], float newp[3]) {
    int gi = pos2idx(newp[0], s->dx), gj = pos2idx(newp[1], s->dx), gk = pos2idx(newp[2], s->dx);
    if (!in_range(gi, gj, gk, s->I, s->J, s->K)) aabb_nearest_inside(&s->boundary, newp);
    if (!near_solid(s, oldp) && !near_solid(s, newp)) return;

    float eps = 1e-6;
    float step = 0.1f * (float)s->dx;                         /* _markerParticleStepDistanceFactor * (float)_dx */
    float dxx = newp[0] - oldp[0], dyy = newp[1] - oldp[1], dzz = newp[2] - oldp[2];
    float travel = vlen(dxx, dyy, dzz);
    if (travel < eps) return;
    int nsteps = (int)ceilf(travel / step);
    float invlen = (float)(1.0 / (double)travel);             /* normalize: v / len, vmath.cpp:100-103 */
    float dir[3] = {dxx * invlen, dyy * invlen, dzz * invlen};

    float last[3] = {oldp[0], oldp[1], oldp[2]};
    float cur[3] = {0, 0, 0};
    int found = 0;
    float cphi = 0.0f;
    for (int st = 0; st < nsteps; st++) {
        if (st == nsteps - 1) {
            cur[0] = newp[0]; cur[1] = newp[1]; cur[2] = newp[2];
        } else {
            float t = (float)(st + 1) * step;
            cur[0] = oldp[0] + dir[0] * t; cur[1] = oldp[1] + dir[1] * t; cur[2] = oldp[2] + dir[2] * t;
        }
        float phi = sdf_sample(s, cur);
        if (phi < 0.0f || !aabb_inside(&s->boundary, cur)) {
            cphi = phi;
            found = 1;
            break;
        }
        last[0] = cur[0]; last[1] = cur[1]; last[2] = cur[2];
    }
    if (!found) return;

    float res[3];
    float maxdist = (float)(s->cfl * s->dx);                  /* float maxResolvedDistance = _CFLConditionNumber * _dx; */
    float grad[3];
    sdf_gradient(s, cur, grad);
    float glen = vlen(grad[0], grad[1], grad[2]);
    if (glen > eps) {
        float ginv = (float)(1.0 / (double)glen);
        grad[0] *= ginv; grad[1] *= ginv; grad[2] *= ginv;
        float push = (float)((double)cphi - (double)0.2f * s->dx);   /* (collisionPhi - _solidBufferWidth*_dx) -> float s */
        res[0] = cur[0] - grad[0] * push; res[1] = cur[1] - grad[1] * push; res[2] = cur[2] - grad[2] * push;
        float rphi = sdf_sample(s, res);
        float rdist = vlen(res[0] - cur[0], res[1] - cur[1], res[2] - cur[2]);
        if (rphi < 0 || rdist > maxdist) { res[0] = last[0]; res[1] = last[1]; res[2] = last[2]; }
    } else {
        res[0] = last[0]; res[1] = last[1]; res[2] = last[2];
    }
    if (!aabb_inside(&s->boundary, res)) {
        float orig[3] = {res[0], res[1], res[2]};
        aabb_nearest_inside(&s->boundary, res);
        float rphi = sdf_sample(s, res);
        float rdist = vlen(res[0] - orig[0], res[1] - orig[1], res[2] - orig[2]);
        if (rphi < 0.0f || rdist > maxdist) { res[0] = last[0]; res[1] = last[1]; res[2] = last[2]; }
    }
    newp[0] = res[0]; newp[1] = res[1]; newp[2] = res[2];
}

void flip_oracle_near_dims(int I, int J, int K, double dx, int *gi, int *gj, int *gk) {
    double cell = 3 * dx;                                     /* _nearSolidGridCellSizeFactor * _dx, :5448 */
    *gi = (int)ceil(I * dx / cell);
    *gj = (int)ceil(J * dx / cell);
    *gk = (int)ceil(K * dx / cell);
}

void flip_oracle_advect(int I, int J, int K, double dx, int n, const float *pos_in, float *pos_out,
                        const float *u, const float *v, const float *w,
                        const float *phi, const uint8_t *near, double dt, double cfl, int collide) {
    macfield m = {I, J, K, dx, u, v, w};
    solid s;
    s.I = I; s.J = J; s.K = K; s.dx = dx; s.phi = phi; s.near = near; s.cfl = cfl;
    s.near_dx = 3 * dx;
    flip_oracle_near_dims(I, J, K, dx, &s.ni, &s.nj, &s.nk);
    s.boundary.px = s.boundary.py = s.boundary.pz = 0.0f;    /* _getBoundaryAABB, :5175-5180 */
    s.boundary.w = I * dx; s.boundary.h = J * dx; s.boundary.d = K * dx;
    aabb_expand(&s.boundary, -3 * dx - 1e-4);
    aabb_expand(&s.boundary, -0.2f * dx);                     /* boundary.expand(-_solidBufferWidth * _dx), :7640 */

    float c2 = (float)(0.5 * dt), c3 = (float)(0.75 * dt), c9 = (float)(dt / 9.0f);
    for (int p = 0; p < n; p++) {
        const float *p0 = pos_in + 3 * (size_t)p;
        float k1[3], k2[3], k3[3], q[3];
        mac_eval(&m, p0, k1);                                 /* _RK3, :7616-7623 */
        for (int c = 0; c < 3; c++) q[c] = p0[c] + k1[c] * c2;
        mac_eval(&m, q, k2);
        for (int c = 0; c < 3; c++) q[c] = p0[c] + k2[c] * c3;
        mac_eval(&m, q, k3);
        float p1[3];
        for (int c = 0; c < 3; c++) {
            float sum = (k1[c] * 2.0f + k2[c] * 3.0f) + k3[c] * 4.0f;
            p1[c] = p0[c] + sum * c9;
        }
        if (collide) resolve_collision(&s, p0, p1);
        pos_out[3 * (size_t)p + 0] = p1[0];
        pos_out[3 * (size_t)p + 1] = p1[1];
        pos_out[3 * (size_t)p + 2] = p1[2];
    }
}

/* ---- valid-face extrapolation (next-row f1) -------------------------------------------------------
 * GridUtils::extrapolateGrid (gridutils.h:94-163) with _initializeStatusGridThread /
 * _findExtrapolationCells (gridutils.cpp:99-173) and _extrapolateCellsThread (gridutils.h:42-91),
 * called per MAC component by MACVelocityField::extrapolateVelocityField (macvelocityfield.cpp:671-677)
 * with numLayers = ceil(sqrt(3) * CFL) + 3 (fluidsimulation.cpp:6282-6286).
 * The reference splits both passes over threads; the set of cells found per layer and every cell's
 * neighbour sum (fixed +i,-i,+j,-j,+k,-k order, DONE neighbours only) do not depend on that split,
 * so this sequential restatement is bit-identical to the threaded original. */
void flip_oracle_extrapolate(int w, int h, int d, float *grid, const uint8_t *valid, int layers) {
    enum { UNKNOWN = 0, WAITING = 1, KNOWN = 2, DONE = 3 };
    size_t n = (size_t)w * h * d;
    unsigned char *status = (unsigned char *)calloc(n ? n : 1, 1);
    size_t *cells = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
    const long long off[6] = {1, -1, w, -(long long)w, (long long)w * h, -(long long)w * h};
    for (int k = 0; k < d; k++)
        for (int j = 0; j < h; j++)
            for (int i = 0; i < w; i++) {
                size_t idx = (size_t)i + (size_t)w * ((size_t)j + (size_t)h * k);
                int border = i == 0 || j == 0 || k == 0 || i == w - 1 || j == h - 1 || k == d - 1;
                status[idx] = border ? DONE : (valid[idx] ? KNOWN : UNKNOWN);
            }
    for (int layer = 0; layer < layers; layer++) {
        size_t ncells = 0;
        for (size_t idx = 0; idx < n; idx++) {                 /* _findExtrapolationCells */
            if (status[idx] != KNOWN) continue;
            for (int q = 0; q < 6; q++) {
                size_t nb = (size_t)((long long)idx + off[q]);
                if (status[nb] == UNKNOWN) {
                    status[nb] = WAITING;
                    cells[ncells++] = nb;
                }
            }
            status[idx] = DONE;
        }
        for (size_t c = 0; c < ncells; c++) {                  /* _extrapolateCellsThread */
            size_t idx = cells[c];
            float sum = 0.0f;
            int count = 0;
            for (int q = 0; q < 6; q++) {
                size_t nb = (size_t)((long long)idx + off[q]);
                if (status[nb] == DONE) {
                    sum += grid[nb];
                    count++;
                }
            }
            grid[idx] = sum / (float)count;
        }
        if (layer != layers - 1)
            for (size_t c = 0; c < ncells; c++) status[cells[c]] = KNOWN;
    }
    free(cells);
    free(status);
}

/* FluidSimulation::_getMaximumMarkerParticleSpeed (fluidsimulation.cpp:10188-10202). */
double flip_oracle_max_particle_speed(int n, const float *vel) {
    double maxsq = 0.0;
    for (int i = 0; i < n; i++) {
        const float *v = vel + 3 * (size_t)i;
        float d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];    /* vmath::dot, float, left to right */
        double distsq = d;
        if (distsq > maxsq) maxsq = distsq;
    }
    return sqrt(maxsq);
}

/* ---- marker-particle removal (next-row f2; oracle groundwork, no device path yet) -------------------
 * FluidSimulation::_removeMarkerParticles (fluidsimulation.cpp:7773-7851) with
 * _getMarkerParticleSpeedLimit (:7723-7771) and MeshLevelSet::trilinearInterpolateSolidPoints
 * (meshlevelset.h:203-216, 333-339), for the default configuration: all domain boundaries closed, no
 * lifetime attribute. removed[i] = 1 for the particles ParticleSystem::removeParticles would drop;
 * *num_extreme = _currentExtremeVelocityParticlesRemoved. The per-cell cap runs in particle-index
 * order and a particle is counted BEFORE its extreme-velocity test, exactly as in the reference. */
static float speed_limit(int n, const float *vel, double dx, double dt, double cfl, int max_frame_steps) {
    double maxParticleSpeed = 0.0;
    double speedLimitStep = cfl * dx / dt;
    int counts[64];
    for (int i = 0; i < max_frame_steps; i++) counts[i] = 0;
    for (int i = 0; i < n; i++) {
        const float *v = vel + 3 * (size_t)i;
        double speed = (double)sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);      /* vmath::length */
        int idx = (int)fmin(floor(speed / speedLimitStep), max_frame_steps - 1);
        counts[idx]++;
        maxParticleSpeed = speed > maxParticleSpeed ? speed : maxParticleSpeed;
    }
    double maxpct = 0.0005;                                   /* _maxExtremeVelocityRemovalPercent */
    int maxabs = 35;                                          /* _maxExtremeVelocityRemovalAbsolute */
    int maxRemovalCount = (int)fmin((int)((double)n * maxpct), maxabs);
    double maxspeed = max_frame_steps * speedLimitStep;
    int currentRemovalCount = 0;
    for (int i = max_frame_steps - 1; i > 0; i--) {
        if (currentRemovalCount + counts[i] > maxRemovalCount) break;
        currentRemovalCount += counts[i];
        int steps = i + 4 > max_frame_steps ? i + 4 : max_frame_steps;      /* _minTimeStepIncreaseForRemoval = 4 */
        maxspeed = steps * speedLimitStep;
    }
    double lower = 0.90 * maxParticleSpeed, thr = 0.99999 * maxParticleSpeed;
    int nlower = 0, nthr = 0;
    for (int i = 0; i < n; i++) {
        const float *v = vel + 3 * (size_t)i;
        double speed = (double)sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (speed >= lower && speed < thr) nlower++;
        if (speed >= thr) nthr++;
    }
    if (nthr <= 6 && nlower <= 6) maxspeed = thr < maxspeed ? thr : maxspeed;   /* _maxExtremeVelocityOutlierRemovalAbsolute */
    return (float)maxspeed;
}

void flip_oracle_remove_particles(int I, int J, int K, double dx, int n, const float *pos, const float *vel,
                                  const float *phi, double dt, double cfl, int max_per_cell, int extreme_removal,
                                  const float *open_bounds, const uint8_t *pre_removed, uint8_t *removed, int *num_extreme) {
    solid s;
    s.I = I; s.J = J; s.K = K; s.dx = dx; s.phi = phi; s.near = 0; s.cfl = cfl;
    int *count = (int *)calloc((size_t)I * J * K, sizeof(int));
    float maxspeed = speed_limit(n, vel, dx, dt, cfl, 6);     /* _maxFrameTimeSteps = 6 */
    double maxspeedsq = maxspeed * maxspeed;                  /* float product, widened */
    int extreme = 0;
    for (int i = 0; i < n; i++) {
        const float *p = pos + 3 * (size_t)i, *v = vel + 3 * (size_t)i;
        removed[i] = sdf_sample(&s, p) < 0.0f;
        if (removed[i]) continue;
        if (pre_removed && pre_removed[i]) { removed[i] = 1; continue; }          /* lifetime rule (:7808-7814), caller-evaluated */
        if (open_bounds && (p[0] < open_bounds[0] || p[0] > open_bounds[1] || p[1] < open_bounds[2] || p[1] > open_bounds[3] ||
                            p[2] < open_bounds[4] || p[2] > open_bounds[5])) {   /* :7817-7823 */
            removed[i] = 1;
            continue;
        }
        int gi = pos2idx(p[0], dx), gj = pos2idx(p[1], dx), gk = pos2idx(p[2], dx);
        if (!in_range(gi, gj, gk, I, J, K)) { removed[i] = 1; continue; }   /* the reference would throw here */
        size_t c = flat(gi, gj, gk, I, J);
        if (count[c] >= max_per_cell) { removed[i] = 1; continue; }
        count[c]++;
        float d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];    /* vmath::dot */
        if (extreme_removal && d > maxspeedsq) { removed[i] = 1; extreme++; continue; }
    }
    *num_extreme = extreme;
    free(count);
}

/* ---- liquid SDF from particles ------------------------------------------
 * ParticleLevelSet::_computeSignedDistanceFromParticles, particlelevelset.cpp:335-398, with
 * _initializeBlockGrid (:400-437), _computeGridCountDataThread (:515-569: which 10^3 blocks a particle is
 * handed to) and _computeExactBandProducerThread (:620-668: block-local distances). phi is the
 * cell-centred Array3d<float>(I, J, K), initial value _getMaxDistance() = (float)(3.0 * dx). The result
 * is a minimum, so neither the particle order nor the thread count matters. */
void flip_oracle_liquid_sdf(int I, int J, int K, double dx, double radius, int n, const float *pos, float *phi) {
    const int W = 10;                                         /* _blockwidth */
    const float maxd = (float)(3.0 * dx);                     /* :331-333 */
    size_t cells = (size_t)I * J * K;
    for (size_t c = 0; c < cells; c++) phi[c] = maxd;
    if (n == 0) return;
    int bi = (I + W - 1) / W, bj = (J + W - 1) / W, bk = (K + W - 1) / W;
    size_t nb = (size_t)bi * bj * bk;
    uint8_t *home = (uint8_t *)calloc(nb, 1), *active = (uint8_t *)calloc(nb, 1);
    float blockdx = (float)(W * dx);                          /* float blockdx = _blockwidth * _dx; */
    for (int p = 0; p < n; p++) {                             /* _initializeActiveBlocksThread :439-450 */
        int a = pos2idx(pos[3 * p], blockdx), b = pos2idx(pos[3 * p + 1], blockdx), c = pos2idx(pos[3 * p + 2], blockdx);
        if (in_range(a, b, c, bi, bj, bk)) home[flat(a, b, c, bi, bj)] = 1;
    }
    for (int k = 0; k < bk; k++)                              /* GridUtils::featherGrid26 */
        for (int j = 0; j < bj; j++)
            for (int i = 0; i < bi; i++) {
                if (!home[flat(i, j, k, bi, bj)]) continue;
                for (int c = -1; c <= 1; c++)
                    for (int b = -1; b <= 1; b++)
                        for (int a = -1; a <= 1; a++)
                            if (in_range(i + a, j + b, k + c, bi, bj, bk)) active[flat(i + a, j + b, k + c, bi, bj)] = 1;
            }
    float r = (float)radius;                                  /* float r = block.radius; */
    float sr = 2.0f * r;                                      /* _searchRadiusFactor * (float)radius */
    double chunk = W * dx;                                    /* _blockwidth * _dx, a double, in the producer */
    for (int p = 0; p < n; p++) {
        float x = pos[3 * p], y = pos[3 * p + 1], z = pos[3 * p + 2];
        int b0 = pos2idx(x, blockdx), b1 = pos2idx(y, blockdx), b2 = pos2idx(z, blockdx);
        float bx = idx2posf(b0, blockdx), by = idx2posf(b1, blockdx), bz = idx2posf(b2, blockdx);
        int lo[3], hi[3];
        if (x - sr > bx && y - sr > by && z - sr > bz && x + sr < bx + blockdx && y + sr < by + blockdx && z + sr < bz + blockdx) {
            lo[0] = hi[0] = b0; lo[1] = hi[1] = b1; lo[2] = hi[2] = b2;
        } else {
            lo[0] = pos2idx(x - sr, blockdx); lo[1] = pos2idx(y - sr, blockdx); lo[2] = pos2idx(z - sr, blockdx);
            hi[0] = pos2idx(x + sr, blockdx); hi[1] = pos2idx(y + sr, blockdx); hi[2] = pos2idx(z + sr, blockdx);
        }
        for (int ck = lo[2]; ck <= hi[2]; ck++)
            for (int cj = lo[1]; cj <= hi[1]; cj++)
                for (int ci = lo[0]; ci <= hi[0]; ci++) {
                    if (!in_range(ci, cj, ck, bi, bj, bk) || !active[flat(ci, cj, ck, bi, bj)]) continue;   /* getBlockID == -1 */
                    float lx = x - idx2posf(ci, chunk), ly = y - idx2posf(cj, chunk), lz = z - idx2posf(ck, chunk);
                    int i0 = pos2idx(lx - sr, dx), j0 = pos2idx(ly - sr, dx), k0 = pos2idx(lz - sr, dx);
                    int i1 = pos2idx(lx + sr, dx), j1 = pos2idx(ly + sr, dx), k1 = pos2idx(lz + sr, dx);
                    if (i0 < 0) i0 = 0;
                    if (j0 < 0) j0 = 0;
                    if (k0 < 0) k0 = 0;
                    if (i1 > W - 1) i1 = W - 1;
                    if (j1 > W - 1) j1 = W - 1;
                    if (k1 > W - 1) k1 = W - 1;
                    double hw = 0.5 * dx;
                    for (int k = k0; k <= k1; k++)
                        for (int j = j0; j <= j1; j++)
                            for (int i = i0; i <= i1; i++) {
                                int gi = ci * W + i, gj = cj * W + j, gk = ck * W + k;
                                if (!in_range(gi, gj, gk, I, J, K)) continue;                  /* write-out :377-379 */
                                /* GridIndexToCellCenter(i, j, k, dx): (float)i*dx + hw in double, narrowed by vec3 */
                                float cx = (float)((double)(float)i * dx + hw), cy = (float)((double)(float)j * dx + hw),
                                      cz = (float)((double)(float)k * dx + hw);
                                float dist = vlen(cx - lx, cy - ly, cz - lz) - r;
                                size_t f = flat(gi, gj, gk, I, J);
                                if (dist < phi[f]) phi[f] = dist;
                            }
                }
    }
    free(home);
    free(active);
}

/* The same field evaluated the way the device's per-axis variant does (k_sdf_scatter_axes): the reference's
 * block set and its block-local cell boxes are both products of per-axis ranges, and each distance term depends
 * only on (axis, block index along that axis, local cell index). So a particle is reduced to three short lists
 * of (global cell index, centre - local coordinate, block offset) and the 3-D work is their product, gated by
 * the 3x3x3 active-block mask; a squared-distance pre-filter skips the square root where the candidate cannot
 * lower the cell. This function exists to prove that decomposition (and the filter) bit-exact against
 * flip_oracle_liquid_sdf on the CPU. Returns the number of candidates the filter skipped. */
#define SDF_AXIS_MAX 24
typedef struct { int n; int g[SDF_AXIS_MAX]; float d[SDF_AXIS_MAX]; int u[SDF_AXIS_MAX]; int lo; } sdf_axis;

static void sdf_axis_list(sdf_axis *a, float x, int simple, int W, int nb, int ncell, float blockdx, double chunk, double dx,
                          float sr) {
    int b = pos2idx(x, blockdx);
    int lo = b, hi = b;
    if (!simple) { lo = pos2idx(x - sr, blockdx); hi = pos2idx(x + sr, blockdx); }
    if (lo < 0) lo = 0;
    if (hi > nb - 1) hi = nb - 1;
    a->n = 0;
    a->lo = lo;
    double hw = 0.5 * dx;
    for (int c = lo; c <= hi; c++) {
        float l = x - idx2posf(c, chunk);
        int i0 = pos2idx(l - sr, dx), i1 = pos2idx(l + sr, dx);
        if (i0 < 0) i0 = 0;
        if (i1 > W - 1) i1 = W - 1;
        if (i1 > ncell - 1 - c * W) i1 = ncell - 1 - c * W;
        for (int i = i0; i <= i1 && a->n < SDF_AXIS_MAX; i++) {
            a->g[a->n] = c * W + i;
            a->d[a->n] = (float)((double)(float)i * dx + hw) - l;
            a->u[a->n] = c - lo;
            a->n++;
        }
    }
}

long long flip_oracle_liquid_sdf_axes(int I, int J, int K, double dx, double radius, int n, const float *pos, float *phi) {
    const int W = 10;
    const float maxd = (float)(3.0 * dx);
    size_t cells = (size_t)I * J * K;
    long long skipped = 0;
    for (size_t c = 0; c < cells; c++) phi[c] = maxd;
    if (n == 0) return 0;
    int bi = (I + W - 1) / W, bj = (J + W - 1) / W, bk = (K + W - 1) / W;
    size_t nb = (size_t)bi * bj * bk;
    uint8_t *home = (uint8_t *)calloc(nb, 1), *active = (uint8_t *)calloc(nb, 1);
    float blockdx = (float)(W * dx);
    for (int p = 0; p < n; p++) {
        int a = pos2idx(pos[3 * p], blockdx), b = pos2idx(pos[3 * p + 1], blockdx), c = pos2idx(pos[3 * p + 2], blockdx);
        if (in_range(a, b, c, bi, bj, bk)) home[flat(a, b, c, bi, bj)] = 1;
    }
    for (int k = 0; k < bk; k++)
        for (int j = 0; j < bj; j++)
            for (int i = 0; i < bi; i++) {
                if (!home[flat(i, j, k, bi, bj)]) continue;
                for (int c = -1; c <= 1; c++)
                    for (int b = -1; b <= 1; b++)
                        for (int a = -1; a <= 1; a++)
                            if (in_range(i + a, j + b, k + c, bi, bj, bk)) active[flat(i + a, j + b, k + c, bi, bj)] = 1;
            }
    float r = (float)radius, sr = 2.0f * r;
    double chunk = W * dx;
    for (int p = 0; p < n; p++) {
        float x = pos[3 * p], y = pos[3 * p + 1], z = pos[3 * p + 2];
        int b0 = pos2idx(x, blockdx), b1 = pos2idx(y, blockdx), b2 = pos2idx(z, blockdx);
        float bx = idx2posf(b0, blockdx), by = idx2posf(b1, blockdx), bz = idx2posf(b2, blockdx);
        int simple = x - sr > bx && y - sr > by && z - sr > bz && x + sr < bx + blockdx && y + sr < by + blockdx && z + sr < bz + blockdx;
        sdf_axis ax, ay, az;
        sdf_axis_list(&ax, x, simple, W, bi, I, blockdx, chunk, dx, sr);
        sdf_axis_list(&ay, y, simple, W, bj, J, blockdx, chunk, dx, sr);
        sdf_axis_list(&az, z, simple, W, bk, K, blockdx, chunk, dx, sr);
        for (int c = 0; c < az.n; c++)
            for (int b = 0; b < ay.n; b++) {
                float sy = ay.d[b] * ay.d[b], sz = az.d[c] * az.d[c];
                for (int a = 0; a < ax.n; a++) {
                    int ci = ax.lo + ax.u[a], cj = ay.lo + ay.u[b], ck = az.lo + az.u[c];
                    if (!active[flat(ci, cj, ck, bi, bj)]) continue;
                    size_t f = flat(ax.g[a], ay.g[b], az.g[c], I, J);
                    float d2 = ax.d[a] * ax.d[a] + sy + sz;
                    float cur = phi[f];
                    float A = (cur + r) * 1.00001f;
                    if (A > 0.0f && d2 > A * A * 1.00001f) { skipped++; continue; }     /* cannot lower the cell */
                    float dist = sqrtf(d2) - r;
                    if (dist < cur) phi[f] = dist;
                }
            }
    }
    free(home);
    free(active);
    return skipped;
}

/* ---- scalar attribute P2G ------------------------------------------------
 * AttributeToGridTransfer<float>::transfer (attributetogridtransfer.h:52-157, 213-520): the FLIP-kernel splat
 * of VelocityAdvector with one scalar per particle, onto the cell-centred I x J x K grid (gridOffset =
 * (dx/2, dx/2, dx/2) at every call site, e.g. fluidsimulation.cpp:7001-7012), normalised, valid = weight > 1e-6.
 * Same block structure, same block-local float frames, same ascending-index summation order as the velocity
 * transfer, so it is p2g_direction with a fourth "direction". The caller pre-fills grid and mask with 0. */
void flip_oracle_attribute_p2g(int I, int J, int K, double dx, double radius, int n, const float *pos, const float *attr,
                               float *grid, uint8_t *valid) {
    p2g_direction(3, I, J, K, dx, radius, FLIP_ORACLE_FLIP, n, pos, attr, NULL, grid, valid, NULL);
}

/* ParticleLevelSet::postProcessSignedDistanceField (particlelevelset.cpp:170-195): liquid cells whose centre is inside
 * the solid (MeshLevelSet::getDistanceAtCellCenter, meshlevelset.cpp:152-162: 0.125f * the 8 corner nodes summed in
 * i-fastest order) are pushed to -dx/2; magnitudes below 0.005 dx are clamped away from zero. In place on phi[K][J][I];
 * solid is the node-centred (K+1)(J+1)(I+1) SDF. */
void flip_oracle_liquid_sdf_postprocess(int I, int J, int K, double dx, float *phi, const float *solid) {
    float eps = (float)(0.005 * dx);
    int w = I + 1, h = J + 1;
    for (int k = 0; k < K; k++)
        for (int j = 0; j < J; j++)
            for (int i = 0; i < I; i++) {
                size_t f = flat(i, j, k, I, J);
                if ((double)phi[f] < 0.5 * dx) {
                    float c = 0.125f * (solid[flat(i, j, k, w, h)] + solid[flat(i + 1, j, k, w, h)] + solid[flat(i, j + 1, k, w, h)] +
                                        solid[flat(i + 1, j + 1, k, w, h)] + solid[flat(i, j, k + 1, w, h)] +
                                        solid[flat(i + 1, j, k + 1, w, h)] + solid[flat(i, j + 1, k + 1, w, h)] +
                                        solid[flat(i + 1, j + 1, k + 1, w, h)]);
                    if (c < 0) phi[f] = (float)(-0.5f * dx);
                }
                float val = phi[f];
                if (fabsf(val) < eps) phi[f] = val > 0 ? eps : -eps;
            }
}

/* AttributeToGridTransfer<vmath::vec3>::transfer (colour, whitewater proximity): three channels accumulated like
 * scalars (vec3 * float and vec3 += are per component), normalised with vec3 /= float. attr and grid are packed
 * float[3] per particle / cell. */
void flip_oracle_attribute_p2g_vec3(int I, int J, int K, double dx, double radius, int n, const float *pos, const float *attr,
                                    float *grid, uint8_t *valid) {
    for (int ch = 0; ch < 3; ch++)
        p2g_direction_n(3, I, J, K, dx, radius, FLIP_ORACLE_FLIP, n, pos, attr + ch, NULL, grid + ch, valid, NULL, 1, 3);
}
