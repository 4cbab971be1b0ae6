// tests/emu/emulate_liquid_sdf.cpp -- TEST INFRASTRUCTURE. Runs the kernels of
// blender_flip_fluids_b200/csrc/ffb200_liquid_sdf.cu (compiled as host code through fake/cuda_runtime.h; the file
// "liquid_sdf_kernels.inc" is that source with its launchers cut off, written by emulate.py) thread by thread and
// returns the field. Launch geometry and order repeat launch_liquid_sdf / launch_liquid_sdf_postprocess.
#include "liquid_sdf_kernels.inc"

using namespace ffb200;

extern "C" int emu_liquid_sdf(int I, int J, int K, double dx, double radius, int n, const float *pos, int variant,
                              const float *solid, float *phi_out) {
    GridDesc g;
    memset(&g, 0, sizeof(g));
    g.I = I; g.J = J; g.K = K; g.kbase = 0; g.kloc = K; g.dx = dx; g.inv_dx = 1.0 / dx; g.inv_2dx = 2.0 * (1.0 / dx);
    std::vector<float> px(n), py(n), pz(n);
    for (int p = 0; p < n; p++) { px[p] = pos[3 * p]; py[p] = pos[3 * p + 1]; pz[p] = pos[3 * p + 2]; }
    SdfParams P;
    P.g = g;
    P.bi = (I + kBlockWidth - 1) / kBlockWidth;
    P.bj = (J + kBlockWidth - 1) / kBlockWidth;
    P.bk = (K + kBlockWidth - 1) / kBlockWidth;
    P.blockdx = (float)(kBlockWidth * dx);
    P.inv_blockdx = 1.0 / (double)P.blockdx;
    P.chunk = kBlockWidth * dx;
    P.hw = 0.5 * dx;
    P.r = (float)radius;
    P.sr = 2.0f * P.r;
    P.px = px.data(); P.py = py.data(); P.pz = pz.data();
    P.n = n;
    if (variant == 1 && !(2.0 * (double)P.sr / dx + 3.0 <= (double)kAxisMax)) return 1;     // the launcher's own gate
    const size_t cells = (size_t)I * J * K, blocks = (size_t)P.bi * P.bj * P.bk;
    std::vector<int> phi(cells);
    std::vector<uint8_t> home(blocks, 0), active(blocks, 0);
    const float maxd = (float)(3.0 * dx);
    int maxbits;
    memcpy(&maxbits, &maxd, 4);
    emu_launch(dim3(7), dim3(256), [&] { k_sdf_fill(phi.data(), cells, maxbits); });
    if (n > 0) {
        emu_launch(dim3((n + 255) / 256), dim3(256), [&] { k_sdf_home(P, home.data()); });
        emu_launch(dim3((unsigned)((blocks + 127) / 128)), dim3(128), [&] { k_sdf_feather(home.data(), active.data(), P.bi, P.bj, P.bk); });
        if (variant == 1)
            emu_launch(dim3((n + kThreads - 1) / kThreads), dim3(kThreads), [&] { k_sdf_scatter_axes(P, active.data(), phi.data()); });
        else
            emu_launch(dim3((n + kThreads - 1) / kThreads), dim3(kThreads), [&] { k_sdf_scatter(P, active.data(), phi.data()); });
    }
    emu_launch(dim3(5), dim3(256), [&] { k_sdf_decode(phi.data(), cells); });
    if (solid)
        emu_launch(dim3((I + 255) / 256, J, K), dim3(256), [&] { k_sdf_postprocess(g, reinterpret_cast<float *>(phi.data()), solid); });
    memcpy(phi_out, phi.data(), cells * 4);
    return 0;
}
