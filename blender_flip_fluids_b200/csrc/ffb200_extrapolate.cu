// ffb200_extrapolate.cu -- valid-face extrapolation of the MAC field on the device.
//
//   GridUtils::extrapolateGrid                     gridutils.h:94-163
//     _initializeStatusGridThread                  gridutils.cpp:99-118
//     _findExtrapolationCells                      gridutils.cpp:120-173
//     _extrapolateCellsThread                      gridutils.h:42-91
//   MACVelocityField::extrapolateVelocityField     macvelocityfield.cpp:671-677  (u, v, w in turn)
//   FluidSimulation::_extrapolateFluidVelocities   fluidsimulation.cpp:6282-6286 (layers = ceil(sqrt(3) CFL) + 3)
//
// It runs right after the P2G inside the reference's "Advect Velocity Field" stage
// (fluidsimulation.cpp:5652-5654); keeping it on the CPU would force a device->host->device round
// trip of the whole field before _saveVelocityField.
//
// The reference grows the valid region one 6-neighbour layer at a time: cells next to a KNOWN cell
// get the mean of their DONE neighbours, summed in the fixed order +i, -i, +j, -j, +k, -k. Its
// threaded frontier lists only decide WHICH thread handles a cell, never the value, so a gather
// -- one thread per face, one launch per layer (u, v and w together) -- produces the same bits.
// The status grid is a layer STAMP, updated in place:
//     255  border cell: DONE from the start             0    valid face: KNOWN in layer 0
//     254  UNKNOWN                                      s    filled during layer s-1: KNOWN in layer s
// so that in layer l "KNOWN" is stamp == l and "DONE when the reference averages" is stamp <= l or
// 255. A thread only ever writes its own UNKNOWN cell (value, then stamp l+1); concurrent readers
// treat both 254 and l+1 as "not there yet", and values are only read from cells with stamp <= l,
// which nobody writes: race-free without a second buffer.
// Work follows the frontier: the grid is cut into 32x4x4 tiles, and a tile is processed in layer l
// only if it or one of its 26 neighbours holds a cell with stamp l (a byte per tile and layer parity,
// written by the tile's own warp); everything else exits after reading 27 bytes. One warp per tile.
#include "ffb200_ctx.h"

#include <algorithm>

namespace ffb200 {

namespace {

constexpr uint8_t kBorder = 255, kUnknown = 254;
constexpr int kTX = 32, kTY = 4, kTZ = 4;

struct ExtrapArgs {
    const uint8_t *valid[3];
    uint8_t *stamp[3];
    float *grid[3];
    uint8_t *tile_has[2][3];     // [layer parity][component]: the tile holds cells with that layer's stamp
    int w[3], h[3], d[3];
    int tx[3], ty[3], tz[3];     // tiles per axis
    int layer;
};

// One WARP per tile (lanes along x, 16 (y, z) rows in turn): an inactive tile costs one warp that
// reads 27 flags. Warps are numbered over the tiles of u, then v, then w.
__device__ __forceinline__ bool extrap_tile(const ExtrapArgs &a, int &c, int &tix, int &tiy, int &tiz, int &tile) {
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (c = 0; c < 3; c++) {
        const long long nt = (long long)a.tx[c] * a.ty[c] * a.tz[c];
        if (wid < nt) break;
        wid -= nt;
    }
    if (c == 3) return false;
    tile = (int)wid;
    tix = tile % a.tx[c];
    tiy = (tile / a.tx[c]) % a.ty[c];
    tiz = tile / (a.tx[c] * a.ty[c]);
    return true;
}

__global__ void __launch_bounds__(256) k_extrap_init(const __grid_constant__ ExtrapArgs a) {
    int c, tix, tiy, tiz, tile;
    if (!extrap_tile(a, c, tix, tiy, tiz, tile)) return;
    const int w = a.w[c], h = a.h[c], d = a.d[c];
    const int i = tix * kTX + (threadIdx.x & 31);
    bool known_any = false;
    for (int r = 0; r < kTY * kTZ; r++) {
        const int j = tiy * kTY + (r % kTY), k = tiz * kTZ + r / kTY;
        if (i < w && j < h && k < d) {
            const size_t idx = (size_t)i + (size_t)w * ((size_t)j + (size_t)h * k);
            const bool border = i == 0 || j == 0 || k == 0 || i == w - 1 || j == h - 1 || k == d - 1;   // isGridIndexOnBorder
            const bool known = !border && a.valid[c][idx] != 0;
            a.stamp[c][idx] = border ? kBorder : (known ? (uint8_t)0 : kUnknown);
            known_any = known_any || known;
        }
    }
    const bool any = __any_sync(0xffffffffu, known_any);
    if ((threadIdx.x & 31) == 0) a.tile_has[0][c][tile] = any ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_extrap_layer(const __grid_constant__ ExtrapArgs a) {
    int c, tix, tiy, tiz, tile;
    if (!extrap_tile(a, c, tix, tiy, tiz, tile)) return;
    const int l = a.layer;
    const uint8_t *has = a.tile_has[l & 1][c];
    uint8_t *has_next = a.tile_has[(l + 1) & 1][c];
    const int tx = a.tx[c], ty = a.ty[c], tz = a.tz[c];
    const int lane = threadIdx.x & 31;
    // frontier test: does this tile or a neighbour hold KNOWN cells of this layer?
    bool near = false;
    if (lane < 27) {
        const int ni = tix + lane % 3 - 1, nj = tiy + (lane / 3) % 3 - 1, nk = tiz + lane / 9 - 1;
        near = ni >= 0 && ni < tx && nj >= 0 && nj < ty && nk >= 0 && nk < tz && has[ni + tx * (nj + ty * nk)] != 0;
    }
    if (!__any_sync(0xffffffffu, near)) {
        if (lane == 0) has_next[tile] = 0;
        return;
    }
    const int w = a.w[c], h = a.h[c], d = a.d[c];
    uint8_t *stamp = a.stamp[c];
    float *grid = a.grid[c];
    const long long sj = w, sk = (long long)w * h;
    const int i = tix * kTX + lane;
    bool filled = false;
    for (int r = 0; r < kTY * kTZ; r++) {
        const int j = tiy * kTY + (r % kTY), k = tiz * kTZ + r / kTY;
        if (!(i < w && j < h && k < d)) continue;
        const long long idx = (long long)i + sj * j + sk * k;
        if (stamp[idx] != kUnknown) continue;                  // an UNKNOWN cell is never a border cell: all six neighbours exist
        const long long nb[6] = {idx + 1, idx - 1, idx + sj, idx - sj, idx + sk, idx - sk};
        uint8_t ns[6];
        bool found = false;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            ns[q] = stamp[nb[q]];
            found = found || ns[q] == (uint8_t)l;
        }
        if (!found) continue;
        float sum = 0.0f;
        int count = 0;
#pragma unroll
        for (int q = 0; q < 6; q++)
            if (ns[q] <= (uint8_t)l || ns[q] == kBorder) {
                sum += grid[nb[q]];
                count++;
            }
        grid[idx] = sum / (float)count;
        stamp[idx] = (uint8_t)(l + 1);
        filled = true;
    }
    const bool any = __any_sync(0xffffffffu, filled);
    if (lane == 0) has_next[tile] = any ? 1 : 0;
}

}  // namespace

int launch_extrapolate(Context &c, int layers) {
    if (c.g.kbase != 0 || c.g.kloc != c.g.K)
        throw CudaError("ffb200_extrapolate_velocity_field: not available on a z-slab context (the layers cross slab planes)");
    if (layers > 250) throw CudaError("ffb200_extrapolate_velocity_field: at most 250 layers");
    ExtrapArgs a;
    long long tiles_total = 0;
    for (int dir = 0; dir < 3; dir++) {
        FaceGrid &f = c.face[dir];
        a.w[dir] = f.gi; a.h[dir] = f.gj; a.d[dir] = f.gk;
        a.tx[dir] = (f.gi + kTX - 1) / kTX; a.ty[dir] = (f.gj + kTY - 1) / kTY; a.tz[dir] = (f.gk + kTZ - 1) / kTZ;
        const size_t tiles = (size_t)a.tx[dir] * a.ty[dir] * a.tz[dir];
        if (!f.status[0]) {
            FFB_CUDA(cudaMalloc(&f.status[0], f.count));
            FFB_CUDA(cudaMalloc(&f.status[1], 2 * tiles));    // the two parity planes of the tile flags
        }
        a.valid[dir] = f.valid;
        a.grid[dir] = f.vel;
        a.stamp[dir] = f.status[0];
        a.tile_has[0][dir] = f.status[1];
        a.tile_has[1][dir] = f.status[1] + tiles;
        tiles_total += (long long)tiles;
    }
    const unsigned block = 256, grid = (unsigned)((tiles_total * 32 + block - 1) / block);
    int launches = 0;
    a.layer = 0;
    k_extrap_init<<<grid, block, 0, c.stream>>>(a);
    launches++;
    for (int l = 0; l < layers; l++) {
        a.layer = l;
        k_extrap_layer<<<grid, block, 0, c.stream>>>(a);
        launches++;
    }
    FFB_CUDA(cudaGetLastError());
    return launches;
}

}  // namespace ffb200
