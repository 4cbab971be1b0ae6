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
// over the whole grid -- one thread per face, one launch per layer (u, v and w together), status ping-ponged between two
// byte arrays -- produces the same bits:
//     UNKNOWN with a KNOWN neighbour -> value = mean of neighbours that are KNOWN or DONE in the
//                                       incoming status (KNOWN cells turn DONE before the
//                                       reference averages), status KNOWN
//     KNOWN                          -> DONE
// Only UNKNOWN cells are written and only KNOWN/DONE cells are read, so the field is updated in place.
#include "ffb200_ctx.h"

#include <algorithm>

namespace ffb200 {

namespace {

enum : uint8_t { kUnknown = 0, kWaiting = 1, kKnown = 2, kDone = 3 };

// u, v and w advance together: one launch per layer covers the planes of all three components.
struct ExtrapArgs {
    const uint8_t *valid[3];
    const uint8_t *sin[3];
    uint8_t *sout[3];
    float *grid[3];
    int w[3], h[3], d[3];
    int last;
};

__device__ __forceinline__ bool extrap_locate(const ExtrapArgs &a, int &c, int &i, int &j, int &k) {
    int plane = blockIdx.z;
    c = 0;
    if (plane >= a.d[0]) { plane -= a.d[0]; c = 1; }
    if (c == 1 && plane >= a.d[1]) { plane -= a.d[1]; c = 2; }
    k = plane;
    i = blockIdx.x * blockDim.x + threadIdx.x;
    j = blockIdx.y * blockDim.y + threadIdx.y;
    return i < a.w[c] && j < a.h[c];
}

__global__ void k_extrap_init(const __grid_constant__ ExtrapArgs a) {
    int c, i, j, k;
    if (!extrap_locate(a, c, i, j, k)) return;
    const int w = a.w[c], h = a.h[c], d = a.d[c];
    const size_t idx = (size_t)i + (size_t)w * ((size_t)j + (size_t)h * k);
    const bool border = i == 0 || j == 0 || k == 0 || i == w - 1 || j == h - 1 || k == d - 1;   // isGridIndexOnBorder
    a.sout[c][idx] = border ? kDone : (a.valid[c][idx] ? kKnown : kUnknown);
}

__global__ void k_extrap_layer(const __grid_constant__ ExtrapArgs a) {
    int c, i, j, k;
    if (!extrap_locate(a, c, i, j, k)) return;
    const uint8_t *__restrict__ sin = a.sin[c];
    float *grid = a.grid[c];
    const long long sj = a.w[c], sk = (long long)a.w[c] * a.h[c];
    const long long idx = (long long)i + sj * j + sk * k;
    const uint8_t s = sin[idx];
    uint8_t o = s;
    if (s == kKnown) {
        o = kDone;
    } else if (s == kUnknown) {                                // never a border cell: all six neighbours exist
        const long long nb[6] = {idx + 1, idx - 1, idx + sj, idx - sj, idx + sk, idx - sk};
        uint8_t ns[6];
        bool found = false;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            ns[q] = sin[nb[q]];
            found = found || ns[q] == kKnown;
        }
        if (found) {
            float sum = 0.0f;
            int count = 0;
#pragma unroll
            for (int q = 0; q < 6; q++)
                if (ns[q] >= kKnown) {
                    sum += grid[nb[q]];
                    count++;
                }
            grid[idx] = sum / (float)count;
            o = a.last ? kWaiting : kKnown;                   // status.set(cells, KNOWN) except after the last layer
        }
    }
    a.sout[c][idx] = o;
}

}  // namespace

int launch_extrapolate(Context &c, int layers) {
    if (c.g.kbase != 0 || c.g.kloc != c.g.K)
        throw CudaError("ffb200_extrapolate_velocity_field: not available on a z-slab context (the layers cross slab planes)");
    ExtrapArgs a;
    int wmax = 0, hmax = 0, planes = 0;
    for (int dir = 0; dir < 3; dir++) {
        FaceGrid &f = c.face[dir];
        if (!f.status[0]) {
            FFB_CUDA(cudaMalloc(&f.status[0], f.count));
            FFB_CUDA(cudaMalloc(&f.status[1], f.count));
        }
        a.valid[dir] = f.valid;
        a.grid[dir] = f.vel;
        a.w[dir] = f.gi; a.h[dir] = f.gj; a.d[dir] = f.gk;
        wmax = std::max(wmax, f.gi); hmax = std::max(hmax, f.gj);
        planes += f.gk;
    }
    dim3 block(64, 4, 1), grid((wmax + 63) / 64, (hmax + 3) / 4, planes);
    int launches = 0;
    a.last = 0;
    for (int dir = 0; dir < 3; dir++) { a.sin[dir] = nullptr; a.sout[dir] = c.face[dir].status[0]; }
    k_extrap_init<<<grid, block, 0, c.stream>>>(a);
    launches++;
    for (int l = 0; l < layers; l++) {
        for (int dir = 0; dir < 3; dir++) { a.sin[dir] = c.face[dir].status[l & 1]; a.sout[dir] = c.face[dir].status[(l & 1) ^ 1]; }
        a.last = l == layers - 1;
        k_extrap_layer<<<grid, block, 0, c.stream>>>(a);
        launches++;
    }
    FFB_CUDA(cudaGetLastError());
    return launches;
}

}  // namespace ffb200
