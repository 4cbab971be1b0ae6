// ffb200_seam.cuh -- per-particle block membership ("seam words"), home-block marks and edge flags
// of the P2G transfer. Shared by the stand-alone pass (k_seam_home, ffb200_p2g.cu) and the particle
// reorder of the sort (ffb200_sort.cu), which computes them on the fly while it is waiting for
// memory anyway.
#pragma once
#include "ffb200_ctx.h"

namespace ffb200 {

struct SeamParams {
    GridDesc g;
    int bdim[3][3];              // block dims per direction
    uint8_t *home[3];
    uint32_t *seam;              // [dir*cap + slot]
    uint32_t *edge_list;         // [dir*edge_cap + i]
    uint32_t *edge_count;        // [dir]
    uint32_t edge_cap;
    int cap;
    const float *px, *py, *pz;
    float h;                     // (float)(0.5*dx)
    float sr;                    // (float)(radius + 1e-6f)
    float blockdx;               // (float)_chunkdx
    double inv_blockdx;          // 1.0 / (double)blockdx
    double inv_chunkdx;          // 1.0 / _chunkdx
    int n;
};

constexpr uint32_t kEdgeBit = 1u << 30;

// Per particle and direction: (a) mark the home block exactly as _initializeActiveBlocksThread
// does (double _chunkdx, :240-251); (b) the inclusive block range the particle is sorted into,
// exactly as _computeGridCountDataThread does (float blockdx, float sr, :306-352).
// Word layout per axis a (10 bits at 10a): (lo+1) in 8 bits, (hi-lo) in 2 bits.
struct AxisSeam {
    int home;           // block by the double _chunkdx (:240-251)
    uint32_t fs, fn;    // packed 10-bit field when the direction is "simple" (range b..b) / not (range of x -+ sr, :330-345)
    bool simple;
    bool near_plane;    // float screen of the edge test failed: needs the exact test
    float x;
};

__device__ __forceinline__ uint32_t seam_field(int lo, int hi) {
    int span = hi - lo;
    span = span < 0 ? 0 : (span > 3 ? 3 : span);
    int lo1 = lo + 1;
    // out-of-range block indices can never match a face's block: park them at 255
    if (lo1 < 0 || lo1 > 254) { lo1 = 255; span = 0; }
    return (uint32_t)lo1 | ((uint32_t)span << 8);
}

// Everything the membership pass needs from one coordinate in one frame (unshifted or shifted by
// dx/2). A direction combines three of these, and only six distinct ones exist per particle.
// (Screening the block floors below with float tests that decide the common case was tried: the
// compiler if-converts them and the kernel gets slower, so only the edge test keeps its screen.)
__device__ __forceinline__ AxisSeam axis_seam(float x, const SeamParams &s) {
    AxisSeam r;
    r.x = x;
    const int b = pos2idx(x, s.inv_blockdx);
    // GridIndexToPosition(b, blockdx) with a float blockdx: the double product of two floats is
    // exact, so narrowing it rounds once -- the float product
    const float bp = __fmul_rn((float)b, s.blockdx);
    const float xm = x - s.sr, xp = x + s.sr, top = bp + s.blockdx;
    r.simple = (xm > bp) && (xp < top);
    r.home = pos2idx(x, s.inv_chunkdx);
    r.fs = seam_field(b, b);
    r.fn = seam_field(pos2idx(xm, s.inv_blockdx), pos2idx(xp, s.inv_blockdx));
    // float screen of the "edge" test: distance of t = x/dx to the nearest integer (magic-number
    // rounding, exact for |t| < 2^22) above 2e-3 -- the float rounding of t is below 1e-4 for |t| < 1000
    const float tf = x * (float)s.g.inv_dx;
    const float rn = (tf + 12582912.0f) - 12582912.0f;
    r.near_plane = !(fabsf(tf - rn) > 2e-3f) || !(fabsf(tf) < 1000.0f);
    return r;
}

// "edge": within a few float ulps of a cell plane in this frame. There a float compare of
// block-local coordinates may disagree with the reference's double floor, so the transfer
// kernels give such particles the exact arithmetic.
__device__ __forceinline__ bool seam_edge_exact(float x, const SeamParams &s) {
    const double t = (double)x * s.g.inv_dx;
    const double fr = t - floor(t);
    const double band = 4e-6 + fabs(t) * 2.5e-7;
    return fr < band || fr > 1.0 - band;
}

// The three membership words, home-block marks and edge-list entries of the particle in sorted
// slot j at position (px, py, pz). Returns the home block per direction in hidx (-1: outside);
// the caller marks them (seam_mark_home).
__device__ __forceinline__ void seam_particle(const SeamParams &s, float px, float py, float pz, int j, int hidx[3]) {
    const float p[3] = {px, py, pz};
    hidx[0] = hidx[1] = hidx[2] = -1;
    AxisSeam un[3], sh[3];                                    // per axis: unshifted frame, frame shifted by dx/2
    bool any_near = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        un[a] = axis_seam(p[a] - 0.0f, s);
        sh[a] = axis_seam(p[a] - s.h, s);
        any_near = any_near || un[a].near_plane || sh[a].near_plane;
    }
    bool eun[3] = {false, false, false}, esh[3] = {false, false, false};
    if (any_near) {                                           // rare: one branch for all six frames
#pragma unroll
        for (int a = 0; a < 3; a++) {
            eun[a] = un[a].near_plane && seam_edge_exact(un[a].x, s);
            esh[a] = sh[a].near_plane && seam_edge_exact(sh[a].x, s);
        }
    }
#pragma unroll
    for (int dir = 0; dir < 3; dir++) {
        const AxisSeam &X = dir == 0 ? un[0] : sh[0], &Y = dir == 1 ? un[1] : sh[1], &Z = dir == 2 ? un[2] : sh[2];
        // (a) home block
        if (in_range3(X.home, Y.home, Z.home, s.bdim[dir][0], s.bdim[dir][1], s.bdim[dir][2]))
            hidx[dir] = X.home + s.bdim[dir][0] * (Y.home + s.bdim[dir][1] * Z.home);
        // (b) membership range: b..b on every axis if all three are "simple", else the x -+ sr ranges
        const bool simple = X.simple && Y.simple && Z.simple;
        uint32_t word = simple ? (X.fs | (Y.fs << 10) | (Z.fs << 20)) : (X.fn | (Y.fn << 10) | (Z.fn << 20));
        const bool edge = (dir == 0 ? eun[0] : esh[0]) || (dir == 1 ? eun[1] : esh[1]) || (dir == 2 ? eun[2] : esh[2]);
        if (edge) word |= kEdgeBit;
        s.seam[(size_t)dir * s.cap + j] = word;
        if (edge) {
            const uint32_t slot = atomicAdd(s.edge_count + dir, 1u);
            if (slot < s.edge_cap) s.edge_list[(size_t)dir * s.edge_cap + slot] = (uint32_t)j;
        }
    }
}

// Marks the home blocks of a CTA's particles. Sorted neighbours share their home block, and
// millions of same-address stores queue up at L2: a thread only stores where its block differs
// from its left neighbour's (block-wide, through shared memory). Every thread of the CTA must call
// this (pass -1 for "nothing to mark"); `sh` holds 3 * blockDim.x ints.
__device__ __forceinline__ void seam_mark_home(const SeamParams &s, const int hidx[3], int *sh) {
    const int t = threadIdx.x, nt = blockDim.x;
#pragma unroll
    for (int dir = 0; dir < 3; dir++) sh[dir * nt + t] = hidx[dir];
    __syncthreads();
#pragma unroll
    for (int dir = 0; dir < 3; dir++)
        if (hidx[dir] >= 0 && (t == 0 || sh[dir * nt + t - 1] != hidx[dir])) s.home[dir][hidx[dir]] = 1;
}

}  // namespace ffb200
