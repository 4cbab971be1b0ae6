// tests/emu/fake/cuda_runtime.h -- TEST INFRASTRUCTURE. A stand-in for <cuda_runtime.h> that lets g++ compile the
// KERNEL SECTION of a product .cu file as plain host C++ (tests/emu/emulate.py cuts the launchers off), so that
// kernels without barriers or warp intrinsics can be run thread by thread on the CPU against the golden fixtures.
// Host float arithmetic with -ffp-contract=off is IEEE like the device's with -fmad=false, so results must be
// bit-identical. Only what the emulated kernels use is provided.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static dim3 blockIdx, threadIdx, blockDim, gridDim;

typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }

using std::max;
using std::min;

static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __double2int_rd(double x) { return (int)floor(x); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicMin(int *p, int v) { int old = *p; if (v < old) *p = v; return old; }

// one "launch": every thread of every block in turn (valid for kernels without barriers / warp intrinsics)
template <class F>
static void emu_launch(dim3 grid, dim3 block, F body) {
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++)
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++) {
                            blockIdx = dim3(bx, by, bz);
                            threadIdx = dim3(tx, ty, tz);
                            body();
                        }
}
