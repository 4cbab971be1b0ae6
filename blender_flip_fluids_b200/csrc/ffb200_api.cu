// ffb200_api.cu -- the extern "C" layer of libffb200.so (include/ffb200.h) and device memory
// management. No kernels here; see ffb200_{sort,p2g,g2p,advect}.cu.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <limits>
#include <new>

#include "ffb200_ctx.h"

#include <cstdlib>
#include "ffb200_seam.cuh"

using namespace ffb200;

namespace {

// One global message buffer, like the reference's c_bindings/cbindings.cpp:37-47.
char g_error[4096] = "";

void set_error(const char *fn, const char *what) { snprintf(g_error, sizeof(g_error), "%s - %s", fn, what); }

template <class F>
int guarded(const char *fn, ffb200_context *ctx, F &&body, bool mutates = true, bool declares = false) {
    try {
        if (!ctx) throw std::invalid_argument("null context");
        Context &c = *reinterpret_cast<Context *>(ctx);
        // an ffb200_declare_resident holds for the IMMEDIATELY following call only, whatever that call is and
        // however it ends: latched here, before any argument check of the body can throw
        c.resident_arg = declares ? 0u : c.resident_next;
        if (!declares) c.resident_next = 0;
        FFB_CUDA(cudaSetDevice(c.device));
        if (mutates) c.epoch++;                               // anything cached about particles / field is stale
        body(c);
        return FFB200_SUCCESS;
    } catch (const std::exception &e) {
        set_error(fn, e.what());
    } catch (...) {
        set_error(fn, "unknown exception");
    }
    return FFB200_FAIL;
}

template <class T>
void dev_alloc(T *&p, size_t count) {
    FFB_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), (count ? count : 1) * sizeof(T)));
}

template <class T>
void dev_free(T *&p) {
    if (p) cudaFree(p);
    p = nullptr;
}

enum Stage { kSort = 0, kP2GPrep, kP2G, kG2P, kAdvect, kH2D, kD2H, kNumStages };

struct StageEvents {
    cudaEvent_t start[kNumStages] = {}, stop[kNumStages] = {};
    bool used[kNumStages] = {};
};

// Context plus the bookkeeping that only this file needs.
constexpr int kUploadChunks = 8;

struct ContextImpl : Context {
    StageEvents evs;
    int launches[kNumStages] = {};
    // pipelined field upload (g2p_upload_pipelined): copy stream, its fork event, one event per plane chunk
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_fork = nullptr, chunk_done[kUploadChunks] = {};
};

// Stage timing events are skipped while the stream is being captured into a CUDA graph (a caller may
// capture whole substeps: every launch of the resident stages is capturable once the scratch buffers
// exist, i.e. after one eager substep; event timing inside a graph is not).
struct StageTimer {
    ContextImpl &c;
    Stage s;
    bool live;
    StageTimer(ContextImpl &ctx, Stage st) : c(ctx), s(st), live(true) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(c.stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) live = false;
        if (live) FFB_CUDA(cudaEventRecord(c.evs.start[s], c.stream));
    }
    void done(int launched) {
        if (live) {
            FFB_CUDA(cudaEventRecord(c.evs.stop[s], c.stream));
            c.evs.used[s] = true;
        }
        c.launches[s] = launched;
    }
};

void free_particles(ContextImpl &c) {
    for (int b = 0; b < 2; b++) {
        for (int q = 0; q < 3; q++) { dev_free(c.soa[b].p[q]); dev_free(c.soa[b].v[q]); }
        for (int q = 0; q < 9; q++) dev_free(c.soa[b].a[q]);
        dev_free(c.soa[b].orig);
        dev_free(c.sort.key[b]);
        dev_free(c.sort.val[b]);
    }
    dev_free(c.sort.seam);
    dev_free(c.sort.edge_list);
    dev_free(c.aos_stage);
    c.cap = 0;
}

void ensure_affine(ContextImpl &c) {
    if (c.soa[0].a[0]) return;
    for (int b = 0; b < 2; b++)
        for (int q = 0; q < 9; q++) {
            dev_alloc(c.soa[b].a[q], (size_t)c.cap);
            FFB_CUDA(cudaMemsetAsync(c.soa[b].a[q], 0, (size_t)c.cap * sizeof(float), c.stream));
        }
}

// Grow-only. With `preserve` the first c.n entries of the current SoA buffer (and their ids)
// survive the reallocation (device-side migration appends to live data).
void ensure_capacity(ContextImpl &c, int n, bool affine, bool preserve = false) {
    if (n > c.cap) {
        const bool had_affine = c.soa[0].a[0] != nullptr;
        FFB_CUDA(cudaStreamSynchronize(c.stream));
        ParticleSoA old = c.soa[c.cur];
        const int keep = preserve ? c.n : 0;
        if (keep > 0) {                                      // detach the live buffer from free_particles
            for (int q = 0; q < 3; q++) { c.soa[c.cur].p[q] = nullptr; c.soa[c.cur].v[q] = nullptr; }
            for (int q = 0; q < 9; q++) c.soa[c.cur].a[q] = nullptr;
            c.soa[c.cur].orig = nullptr;
        }
        free_particles(c);
        c.cap = n + n / 8 + 1024;
        for (int b = 0; b < 2; b++) {
            for (int q = 0; q < 3; q++) { dev_alloc(c.soa[b].p[q], (size_t)c.cap); dev_alloc(c.soa[b].v[q], (size_t)c.cap); }
            dev_alloc(c.soa[b].orig, (size_t)c.cap);
            dev_alloc(c.sort.key[b], (size_t)c.cap);
            dev_alloc(c.sort.val[b], (size_t)c.cap);
        }
        dev_alloc(c.sort.seam, (size_t)c.cap * 3);
        c.sort.edge_cap = (uint32_t)(c.cap / 64 + 4096);
        dev_alloc(c.sort.edge_list, (size_t)c.sort.edge_cap * 3);
        dev_alloc(c.aos_stage, (size_t)c.cap * 3);
        if (had_affine || affine) ensure_affine(c);
        if (keep > 0) {
            ParticleSoA &dst = c.soa[c.cur];
            const size_t bytes = (size_t)keep * sizeof(float);
            for (int q = 0; q < 3; q++) {
                FFB_CUDA(cudaMemcpyAsync(dst.p[q], old.p[q], bytes, cudaMemcpyDeviceToDevice, c.stream));
                FFB_CUDA(cudaMemcpyAsync(dst.v[q], old.v[q], bytes, cudaMemcpyDeviceToDevice, c.stream));
            }
            if (old.a[0])
                for (int q = 0; q < 9; q++)
                    FFB_CUDA(cudaMemcpyAsync(dst.a[q], old.a[q], bytes, cudaMemcpyDeviceToDevice, c.stream));
            FFB_CUDA(cudaMemcpyAsync(dst.orig, old.orig, bytes, cudaMemcpyDeviceToDevice, c.stream));
            FFB_CUDA(cudaStreamSynchronize(c.stream));
            for (int q = 0; q < 3; q++) { dev_free(old.p[q]); dev_free(old.v[q]); }
            for (int q = 0; q < 9; q++) dev_free(old.a[q]);
            dev_free(old.orig);
        }
    }
    if (affine) ensure_affine(c);
}

void destroy_impl(ContextImpl *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_particles(*c);
    dev_free(c->sort.tile_hist);
    dev_free(c->sort.bin_start);
    dev_free(c->sort.scan_partials);
    for (int d = 0; d < 3; d++) {
        FaceGrid &f = c->face[d];
        dev_free(f.vel); dev_free(f.saved); dev_free(f.wsum); dev_free(f.valid); dev_free(f.home); dev_free(f.active); dev_free(f.status[0]); dev_free(f.status[1]);
    }
    dev_free(c->phi);
    dev_free(c->near_solid);
    dev_free(c->solid_clear[0]);
    dev_free(c->solid_clear[1]);
    dev_free(c->slab_counters);
    dev_free(c->remove_words);
    dev_free(c->liquid_phi);
    dev_free(c->liquid_blocks);
    dev_free(c->tol_stats);
    dev_free(c->cell.wsum); dev_free(c->cell.home); dev_free(c->cell.active);
    dev_free(c->sort.seam_cell);
    dev_free(c->sort.edge_count);
    for (int q = 0; q < 3; q++) dev_free(c->k1s[q]);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->copy_fork) cudaEventDestroy(c->copy_fork);
    for (auto &e : c->chunk_done)
        if (e) cudaEventDestroy(e);
    for (auto &cs : c->sort.cell) {
        if (cs.partial) cudaFree(cs.partial);
        if (cs.cell_list) cudaFree(cs.cell_list);
        if (cs.ovf) cudaFree(cs.ovf);
        dev_free(cs.cell_flag);
        dev_free(cs.list_count);
        dev_free(cs.ovf_count);
        if (cs.done) cudaEventDestroy(cs.done);
        if (cs.stream) cudaStreamDestroy(cs.stream);
    }
    if (c->sort.fork) cudaEventDestroy(c->sort.fork);
    for (int s = 0; s < kNumStages; s++) {
        if (c->evs.start[s]) cudaEventDestroy(c->evs.start[s]);
        if (c->evs.stop[s]) cudaEventDestroy(c->evs.stop[s]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int create_impl(ffb200_context **out, int I, int J, int K, double dx, int device, int k_begin, int k_end, int halo) {
    ContextImpl *c = nullptr;
    try {
        if (!out) throw std::invalid_argument("null output pointer");
        *out = nullptr;
        if (I <= 0 || J <= 0 || K <= 0 || !(dx > 0.0)) throw std::domain_error("grid dimensions and dx must be positive");
        if (k_begin < 0 || k_end > K || k_begin >= k_end || halo < 0) throw std::domain_error("invalid z-slab range");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw CudaError(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                            "); libffb200 has no CPU fallback");
        if (device < 0 || device >= ndev) throw std::domain_error("CUDA device ordinal out of range");
        FFB_CUDA(cudaSetDevice(device));
        c = new ContextImpl();
        c->device = device;
        FFB_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
        c->k_own_begin = k_begin; c->k_own_end = k_end; c->halo = halo;
        GridDesc &g = c->g;
        g.I = I; g.J = J; g.K = K;
        g.kbase = k_begin - halo < 0 ? 0 : k_begin - halo;
        const int ktop = k_end + halo > K ? K : k_end + halo;
        g.kloc = ktop - g.kbase;
        g.dx = dx;
        g.inv_dx = 1.0 / dx;
        g.inv_2dx = 2.0 * g.inv_dx;
        g.HX = 2 * I + 2 * kApron; g.HY = 2 * J + 2 * kApron; g.HZ = 2 * g.kloc + 2 * kApron;
        const unsigned long long nb = (unsigned long long)g.HX * g.HY * g.HZ;
        if (nb >= 0xfffffff0ull) throw std::domain_error("grid too large for 32-bit bin keys");
        g.nbins = (uint32_t)nb;
        FFB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        for (int s = 0; s < kNumStages; s++) {
            FFB_CUDA(cudaEventCreate(&c->evs.start[s]));
            FFB_CUDA(cudaEventCreate(&c->evs.stop[s]));
        }
        dev_alloc(c->sort.bin_start, (size_t)g.nbins + 2);
        for (int d = 0; d < 3; d++) {
            FaceGrid &f = c->face[d];
            f.gi = I + (d == 0); f.gj = J + (d == 1); f.gk = K + (d == 2);
            f.bi = (f.gi + kChunk - 1) / kChunk; f.bj = (f.gj + kChunk - 1) / kChunk; f.bk = (f.gk + kChunk - 1) / kChunk;
            f.kstore = g.kloc + (d == 2);
            f.count = (size_t)f.gi * f.gj * f.kstore;
            dev_alloc(f.vel, f.count); dev_alloc(f.saved, f.count); dev_alloc(f.wsum, f.count); dev_alloc(f.valid, f.count);
            dev_alloc(f.home, (size_t)f.bi * f.bj * f.bk); dev_alloc(f.active, (size_t)f.bi * f.bj * f.bk);
            FFB_CUDA(cudaMemsetAsync(f.vel, 0, f.count * 4, c->stream));
            FFB_CUDA(cudaMemsetAsync(f.saved, 0, f.count * 4, c->stream));
            FFB_CUDA(cudaMemsetAsync(f.wsum, 0, f.count * 4, c->stream));
            FFB_CUDA(cudaMemsetAsync(f.valid, 0, f.count, c->stream));
        }
        const double cell = 3 * dx;                         // fluidsimulation.cpp:5448-5451
        c->ni = (int)std::ceil(I * dx / cell); c->nj = (int)std::ceil(J * dx / cell); c->nk = (int)std::ceil(K * dx / cell);
        dev_alloc(c->phi, (size_t)(I + 1) * (J + 1) * (g.kloc + 1));
        dev_alloc(c->near_solid, (size_t)c->ni * c->nj * c->nk);
        dev_alloc(c->slab_counters, 8);
        dev_alloc(c->sort.edge_count, 4);
        FFB_CUDA(cudaStreamSynchronize(c->stream));
        *out = reinterpret_cast<ffb200_context *>(static_cast<Context *>(c));
        return FFB200_SUCCESS;
    } catch (const std::exception &e) {
        set_error("ffb200_create", e.what());
    } catch (...) {
        set_error("ffb200_create", "unknown exception");
    }
    destroy_impl(c);
    return FFB200_FAIL;
}

ContextImpl &impl(Context &c) { return static_cast<ContextImpl &>(c); }

float below(double lim) {                                     // largest float f with (double)f < lim
    float f = (float)lim;
    if ((double)f >= lim) f = std::nextafterf(f, -INFINITY);
    return f;
}

void upload_attr(ContextImpl &c, const float *host, float *const dst[3], int n) {
    FFB_CUDA(cudaMemcpyAsync(c.aos_stage, host, (size_t)n * 12, cudaMemcpyHostToDevice, c.stream));
    launch_unpack_aos(c, c.aos_stage, dst, n);
}

void download_attr(ContextImpl &c, const float *const src[3], float *host, int n) {
    launch_pack_aos(c, src, c.soa[c.cur].orig, c.aos_stage, n);
    FFB_CUDA(cudaMemcpyAsync(host, c.aos_stage, (size_t)n * 12, cudaMemcpyDeviceToHost, c.stream));
}

void set_particles_impl(ContextImpl &c, int n, const float *pos, const float *vel, const float *affx, const float *affy,
                        const float *affz) {
    if (n < 0) throw std::domain_error("negative particle count");
    if (n > 0 && (!pos || !vel)) throw std::invalid_argument("positions and velocities are required");
    const bool affine = affx && affy && affz;
    ensure_capacity(c, n, affine);
    c.n = n;
    c.has_affine = affine;
    c.sorted = false;
    if (n == 0) return;
    StageTimer t(c, kH2D);
    ParticleSoA &s = c.soa[c.cur];
    upload_attr(c, pos, s.p, n);
    upload_attr(c, vel, s.v, n);
    if (affine) {
        upload_attr(c, affx, s.a + 0, n);
        upload_attr(c, affy, s.a + 3, n);
        upload_attr(c, affz, s.a + 6, n);
    }
    launch_iota(c, s.orig, n);
    t.done(0);
}

void get_particles_impl(ContextImpl &c, float *pos, float *vel, float *affx, float *affy, float *affz) {
    const int n = c.n;
    if (n > 0) {
        StageTimer t(c, kD2H);
        ParticleSoA &s = c.soa[c.cur];
        if (pos) download_attr(c, s.p, pos, n);
        if (vel) download_attr(c, s.v, vel, n);
        if (affx || affy || affz) {
            if (!s.a[0]) throw std::logic_error("no affine data on the device");
            if (affx) download_attr(c, s.a + 0, affx, n);
            if (affy) download_attr(c, s.a + 3, affy, n);
            if (affz) download_attr(c, s.a + 6, affz, n);
        }
        t.done(0);
    }
    FFB_CUDA(cudaStreamSynchronize(c.stream));
}

void sort_impl(ContextImpl &c) {
    if (c.sorted) return;
    StageTimer t(c, kSort);
    int l = launch_sort(c);
    t.done(l);
}

// Host grids are full-size reference arrays; a slab context copies its stored planes.
void upload_field(ContextImpl &c, bool saved, const float *u, const float *v, const float *w) {
    const float *h[3] = {u, v, w};
    for (int d = 0; d < 3; d++) {
        if (!h[d]) throw std::invalid_argument("null velocity field pointer");
        FaceGrid &f = c.face[d];
        const size_t off = (size_t)f.gi * f.gj * c.g.kbase;
        FFB_CUDA(cudaMemcpyAsync(saved ? f.saved : f.vel, h[d] + off, f.count * 4, cudaMemcpyHostToDevice, c.stream));
    }
}

void set_solid_impl(ContextImpl &c, const float *phi, const uint8_t *near_solid) {
    if (!phi || !near_solid) throw std::invalid_argument("null solid SDF / near-solid pointer");
    const GridDesc &g = c.g;
    const size_t plane = (size_t)(g.I + 1) * (g.J + 1);
    FFB_CUDA(cudaMemcpyAsync(c.phi, phi + plane * g.kbase, plane * (g.kloc + 1) * 4, cudaMemcpyHostToDevice, c.stream));
    FFB_CUDA(cudaMemcpyAsync(c.near_solid, near_solid, (size_t)c.ni * c.nj * c.nk, cudaMemcpyHostToDevice, c.stream));
    launch_solid_clearance(c);
    c.has_solid = true;
}

void p2g_impl(ContextImpl &c, double radius, int method, const HostFieldOut *host = nullptr) {
    if (method != FFB200_TRANSFER_FLIP && method != FFB200_TRANSFER_APIC) throw std::domain_error("unknown transfer method");
    if (!(radius > 0.0)) throw std::domain_error("particle radius must be positive");
    if (method == FFB200_TRANSFER_APIC && !c.has_affine) throw std::logic_error("APIC transfer needs affine particle data");
    if (c.nondestructive) c.sorted = false;                  // fixed-batch mode: always re-bin and re-sort
    bool seam_done = false;
    static const bool fuse_seam = [] { const char *e = std::getenv("FFB200_FUSE_SEAM"); return e ? std::atoi(e) != 0 : true; }();
    if (fuse_seam && !c.sorted && c.n > 0) {
        // the reorder pass of the sort also computes the membership words of this transfer
        StageTimer ts(c, kSort);
        SeamParams sp;
        p2g_seam_begin(c, radius, sp);
        int ls = launch_sort(c, &sp);
        ts.done(ls);
        seam_done = true;
    } else {
        sort_impl(c);
    }
    StageTimer tp(c, kP2GPrep);
    int lp = launch_p2g_prepare(c, radius, seam_done);
    tp.done(lp);
    StageTimer t(c, kP2G);
    int l = launch_p2g(c, radius, method, host);
    t.done(l);
    if (host) FFB_CUDA(cudaStreamSynchronize(c.stream));      // the copies were enqueued behind each direction's kernels
}

bool g2p_upload_pipelined(ContextImpl &c, int method, double ratio, const float *const cur[3], const float *const saved[3]);

// upload_cur / upload_saved: host fields to bring in first (null: the device copies are current)
void g2p_impl(ContextImpl &c, int method, double ratio, const float *const *upload_cur = nullptr,
              const float *const *upload_saved = nullptr) {
    if (method != FFB200_TRANSFER_FLIP && method != FFB200_TRANSFER_APIC) throw std::domain_error("unknown transfer method");
    if (method == FFB200_TRANSFER_APIC) {
        ensure_capacity(c, c.n, true);
        c.has_affine = true;
    }
    if (method == FFB200_TRANSFER_FLIP && c.k1s_cap < c.cap) {   // vPIC streams, grow-only
        FFB_CUDA(cudaStreamSynchronize(c.stream));
        for (int q = 0; q < 3; q++) {
            dev_free(c.k1s[q]);
            dev_alloc(c.k1s[q], (size_t)c.cap);
        }
        c.k1s_cap = c.cap;
    }
    if (!upload_cur || !g2p_upload_pipelined(c, method, ratio, upload_cur, upload_saved)) {
        if (upload_cur) {
            StageTimer th(c, kH2D);
            upload_field(c, false, upload_cur[0], upload_cur[1], upload_cur[2]);
            th.done(0);
            if (upload_saved) upload_field(c, true, upload_saved[0], upload_saved[1], upload_saved[2]);
        }
        StageTimer t(c, kG2P);
        int l = launch_g2p(c, method, ratio);
        t.done(l);
    }
    if (method == FFB200_TRANSFER_FLIP) {
        c.k1_epoch = c.epoch;
        c.k1_buf = 2;
    }
    if (method == FFB200_TRANSFER_APIC) {
        // the APIC particle velocity IS the field sampled at the particle (fluidsimulation.cpp:6839-6841),
        // i.e. the first RK3 stage of the advection that follows: remember where it lives
        c.k1_epoch = c.epoch;
        c.k1_buf = c.nondestructive ? (c.cur ^ 1) : c.cur;
    }
}

// G2P with the field(s) coming from the host: the stored planes travel in kUploadChunks chunks on a copy stream, and
// behind each chunk the gather runs over the sorted particles (z-major bins) whose support lies wholly in the planes
// that have arrived -- the upload of a 512^3 field (1.6 GB, ~32 ms over PCIe) then hides the gather instead of
// preceding it. A particle of cell plane k reads planes k-1 .. k+1 of u and v (staggered half a cell in z) and k, k+1
// of w, in every kernel variant; ranges end two planes short of the uploaded extent. Needs sorted resident particles
// (their bin table gives the range bounds: one small read) and no active particle window; false = not applicable.
bool g2p_upload_pipelined(ContextImpl &c, int method, double ratio, const float *const cur[3], const float *const saved[3]) {
    static const bool on = [] { const char *e = std::getenv("FFB200_PIPELINED_UPLOAD"); return e ? std::atoi(e) != 0 : true; }();
    const GridDesc &g = c.g;
    if (!on || !c.sorted || c.n == 0 || c.window.mode != 0 || g.kloc < 8 * kUploadChunks) return false;
    if ((size_t)c.face[0].count * 4 < ((size_t)8 << 20)) return false;   // small fields: one copy, one launch
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(c.stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return false;
    if (!c.copy_stream) {
        FFB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        FFB_CUDA(cudaEventCreateWithFlags(&c.copy_fork, cudaEventDisableTiming));
        for (int q = 0; q < kUploadChunks; q++) FFB_CUDA(cudaEventCreateWithFlags(&c.chunk_done[q], cudaEventDisableTiming));
    }
    // particle index where each range ends: the first particle of cell plane kbase + z_end - 2
    int zend[kUploadChunks];
    uint32_t bound[kUploadChunks];
    for (int q = 0; q < kUploadChunks; q++) {
        zend[q] = (int)((long long)g.kloc * (q + 1) / kUploadChunks);
        if (q + 1 < kUploadChunks) {
            const unsigned long long hz = (unsigned long long)(2 * (zend[q] - 2) + kApron);
            const unsigned long long bin = hz * (unsigned long long)g.HY * (unsigned long long)g.HX;
            FFB_CUDA(cudaMemcpyAsync(&bound[q], c.sort.bin_start + bin, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
        }
    }
    FFB_CUDA(cudaEventRecord(c.copy_fork, c.stream));          // the copies overwrite fields earlier work may still read
    FFB_CUDA(cudaStreamWaitEvent(c.copy_stream, c.copy_fork, 0));
    FFB_CUDA(cudaStreamSynchronize(c.stream));
    bound[kUploadChunks - 1] = (uint32_t)c.n;
    StageTimer t(c, kG2P);
    int launches = 0, z0 = 0;
    uint32_t first = 0;
    for (int q = 0; q < kUploadChunks; q++) {
        for (int pass = 0; pass < 2; pass++) {
            const float *const *h = pass == 0 ? saved : cur;
            if (!h) continue;
            for (int d = 0; d < 3; d++) {
                FaceGrid &f = c.face[d];
                const size_t plane = (size_t)f.gi * f.gj;
                // w stores kloc + 1 planes: the extra one goes with the last chunk
                const int z1 = (q + 1 == kUploadChunks) ? f.kstore : zend[q];
                float *dst = pass == 0 ? f.saved : f.vel;
                FFB_CUDA(cudaMemcpyAsync(dst + plane * z0, h[d] + plane * ((size_t)g.kbase + z0), plane * (size_t)(z1 - z0) * 4,
                                         cudaMemcpyHostToDevice, c.copy_stream));
            }
        }
        FFB_CUDA(cudaEventRecord(c.chunk_done[q], c.copy_stream));
        FFB_CUDA(cudaStreamWaitEvent(c.stream, c.chunk_done[q], 0));
        const uint32_t end = std::min(std::max(bound[q], first), (uint32_t)c.n);
        launches += launch_g2p(c, method, ratio, (int)first, (int)(end - first));
        first = end;
        z0 = zend[q];
    }
    t.done(launches);
    // the caller's buffers are free again when this returns, whether or not a download follows
    FFB_CUDA(cudaEventSynchronize(c.chunk_done[kUploadChunks - 1]));
    return true;
}

void advect_impl(ContextImpl &c, double dt, double cfl, int collide) {
    StageTimer t(c, kAdvect);
    int l = launch_advect(c, dt, cfl, collide);
    t.done(l);
}

void get_field_impl(ContextImpl &c, float *u, float *v, float *w, uint8_t *vu, uint8_t *vv, uint8_t *vw) {
    float *h[3] = {u, v, w};
    uint8_t *hv[3] = {vu, vv, vw};
    StageTimer t(c, kD2H);
    for (int d = 0; d < 3; d++) {
        FaceGrid &f = c.face[d];
        const size_t off = (size_t)f.gi * f.gj * c.g.kbase;
        if (h[d]) FFB_CUDA(cudaMemcpyAsync(h[d] + off, f.vel, f.count * 4, cudaMemcpyDeviceToHost, c.stream));
        if (hv[d]) FFB_CUDA(cudaMemcpyAsync(hv[d] + off, f.valid, f.count, cudaMemcpyDeviceToHost, c.stream));
    }
    t.done(0);
    FFB_CUDA(cudaStreamSynchronize(c.stream));
}


RemoveRules remove_rules(double dt, double cfl, int max_per_cell, int max_frame_steps, int extreme_on, const float *open_bounds) {
    RemoveRules r;
    r.dt = dt;
    r.cfl = cfl;
    r.max_per_cell = max_per_cell;
    r.max_frame_steps = max_frame_steps;
    r.extreme_on = extreme_on != 0;
    for (int q = 0; q < 6; q++) r.bounds[q] = open_bounds ? open_bounds[q] : ((q & 1) ? INFINITY : -INFINITY);
    return r;
}

}  // namespace

namespace ffb200 {

FastGrid make_fast_grid(const GridDesc &g) {
    FastGrid f;
    f.inv_hi = (float)g.inv_dx;
    f.inv_lo = (float)(g.inv_dx - (double)f.inv_hi);
    f.xmax = below(g.dx * g.I);                               // Grid3d::isPositionInGrid: x < dx * I in double (grid3d.h:134-136)
    f.ymax = below(g.dx * g.J);
    f.zmax = below(g.dx * g.K);
    f.I = g.I; f.J = g.J; f.K = g.K; f.kbase = g.kbase;
    f.sju = g.I + 1; f.sku = (g.I + 1) * g.J;
    f.sjv = g.I;     f.skv = g.I * (g.J + 1);
    f.sjw = g.I;     f.skw = g.I * g.J;
    return f;
}

unsigned long long *tolerance_stats(Context &c) {
    if (!c.tol_stats) {
        FFB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c.tol_stats), 4 * sizeof(unsigned long long)));
        FFB_CUDA(cudaMemsetAsync(c.tol_stats, 0, 4 * sizeof(unsigned long long), c.stream));
    }
    return c.tol_stats;
}

}  // namespace ffb200

extern "C" {

int ffb200_set_precision(ffb200_context *ctx, int mode) {
    return guarded("ffb200_set_precision", ctx, [&](Context &c) {
        if (mode != FFB200_PRECISION_EXACT && mode != FFB200_PRECISION_TOLERANCE) throw std::domain_error("unknown precision mode");
        c.precision = mode;
    });
}

int ffb200_set_particle_window(ffb200_context *ctx, int k_lo, int k_hi, int mode) {
    return guarded("ffb200_set_particle_window", ctx, [&](Context &c) {
        if (mode < 0 || mode > 2) throw std::domain_error("window mode must be 0 (off), 1 (inside) or 2 (outside)");
        const GridDesc &g = c.g;
        auto bin_of_plane = [&](long long k) -> uint32_t {          // first bin of cell plane k (clamped to the bin grid)
            long long hz = 2 * (k - g.kbase) + kApron;
            hz = hz < 0 ? 0 : (hz > g.HZ ? g.HZ : hz);
            return (uint32_t)((unsigned long long)hz * (unsigned long long)g.HY * (unsigned long long)g.HX);
        };
        c.window.bin_start = c.sort.bin_start;
        c.window.lo_bin = bin_of_plane(k_lo);
        c.window.hi_bin = bin_of_plane(k_hi);
        c.window.mode = mode;
    }, false);
}

int ffb200_get_tolerance_stats(ffb200_context *ctx, unsigned long long *counts, int reset) {
    return guarded("ffb200_get_tolerance_stats", ctx, [&](Context &c) {
        if (!counts) throw std::invalid_argument("null output pointer");
        unsigned long long *d = tolerance_stats(c);
        FFB_CUDA(cudaMemcpyAsync(counts, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
        FFB_CUDA(cudaStreamSynchronize(c.stream));
        counts[2] = c.tol_advected;
        if (reset) {
            FFB_CUDA(cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long), c.stream));
            c.tol_advected = 0;
        }
    }, false);
}

int ffb200_create(ffb200_context **ctx, int isize, int jsize, int ksize, double dx, int device) {
    return create_impl(ctx, isize, jsize, ksize, dx, device, 0, ksize, 0);
}

int ffb200_create_slab(ffb200_context **ctx, int isize, int jsize, int ksize, double dx, int device, int k_begin,
                       int k_end, int halo) {
    return create_impl(ctx, isize, jsize, ksize, dx, device, k_begin, k_end, halo);
}

void ffb200_destroy(ffb200_context *ctx) {
    if (ctx) destroy_impl(static_cast<ContextImpl *>(reinterpret_cast<Context *>(ctx)));
}

const char *ffb200_get_error_message(void) { return g_error; }

int ffb200_get_version(int *major, int *minor, int *revision) {
    if (major) *major = 0;
    if (minor) *minor = 1;
    if (revision) *revision = 0;
    return FFB200_SUCCESS;
}

int ffb200_set_stream(ffb200_context *ctx, void *cuda_stream) {
    return guarded("ffb200_set_stream", ctx, [&](Context &c) {
        FFB_CUDA(cudaStreamSynchronize(c.stream));
        c.stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    });
}

int ffb200_reset_stream(ffb200_context *ctx) {
    return guarded("ffb200_reset_stream", ctx, [&](Context &c) {
        FFB_CUDA(cudaStreamSynchronize(c.stream));
        c.stream = c.own_stream;
    });
}

int ffb200_set_fixed_batch(ffb200_context *ctx, int on) {
    return guarded("ffb200_set_fixed_batch", ctx, [&](Context &c) { c.nondestructive = on != 0; });
}

int ffb200_synchronize(ffb200_context *ctx) {
    return guarded("ffb200_synchronize", ctx, [&](Context &c) { FFB_CUDA(cudaStreamSynchronize(c.stream)); }, false);
}

int ffb200_get_timing(ffb200_context *ctx, ffb200_timing *out) {
    return guarded("ffb200_get_timing", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        if (!out) throw std::invalid_argument("null output pointer");
        FFB_CUDA(cudaStreamSynchronize(c.stream));
        float ms[kNumStages] = {};
        for (int s = 0; s < kNumStages; s++)
            if (c.evs.used[s]) FFB_CUDA(cudaEventElapsedTime(&ms[s], c.evs.start[s], c.evs.stop[s]));
        out->sort_ms = ms[kSort]; out->p2g_prep_ms = ms[kP2GPrep]; out->p2g_ms = ms[kP2G]; out->g2p_ms = ms[kG2P];
        out->advect_ms = ms[kAdvect];
        out->h2d_ms = ms[kH2D]; out->d2h_ms = ms[kD2H];
        out->sort_launches = c.launches[kSort]; out->p2g_prep_launches = c.launches[kP2GPrep];
        out->p2g_launches = c.launches[kP2G];
        out->g2p_launches = c.launches[kG2P]; out->advect_launches = c.launches[kAdvect];
    }, false);
}

int ffb200_set_valid_guard(ffb200_context *ctx, float abs_tol, float per_contrib_tol) {
    return guarded("ffb200_set_valid_guard", ctx, [&](Context &c) {
        c.guard_abs = abs_tol;
        c.guard_per = per_contrib_tol;
    }, false);
}

int ffb200_set_particles(ffb200_context *ctx, int n, const float *pos, const float *vel, const float *affx,
                         const float *affy, const float *affz) {
    return guarded("ffb200_set_particles", ctx, [&](Context &c) { set_particles_impl(impl(c), n, pos, vel, affx, affy, affz); });
}

int ffb200_get_particles(ffb200_context *ctx, float *pos, float *vel, float *affx, float *affy, float *affz) {
    return guarded("ffb200_get_particles", ctx, [&](Context &c) { get_particles_impl(impl(c), pos, vel, affx, affy, affz); });
}

int ffb200_get_num_particles(ffb200_context *ctx, int *n) {
    return guarded("ffb200_get_num_particles", ctx, [&](Context &c) {
        if (!n) throw std::invalid_argument("null output pointer");
        *n = c.n;
    }, false);
}

int ffb200_get_device_buffers(ffb200_context *ctx, ffb200_device_buffers *out) {
    return guarded("ffb200_get_device_buffers", ctx, [&](Context &c) {
        if (!out) throw std::invalid_argument("null output pointer");
        ParticleSoA &s = c.soa[c.cur];
        for (int q = 0; q < 3; q++) { out->pos[q] = s.p[q]; out->vel[q] = s.v[q]; }
        for (int q = 0; q < 9; q++) out->aff[q] = s.a[q];
        out->ids = s.orig;
        out->n = c.n;
        out->capacity = c.cap;
        for (int d = 0; d < 3; d++) {
            out->field[d] = c.face[d].vel;
            out->saved[d] = c.face[d].saved;
            out->valid[d] = c.face[d].valid;
            out->face_count[d] = (long long)c.face[d].count;
            out->face_plane[d] = c.face[d].gi * c.face[d].gj;
        }
        out->kbase = c.g.kbase;
        out->kloc = c.g.kloc;
        out->k_own_begin = c.k_own_begin;
        out->k_own_end = c.k_own_end;
        out->phi = c.phi;
    });
}

int ffb200_reserve_particles(ffb200_context *ctx, int capacity, int with_affine) {
    return guarded("ffb200_reserve_particles", ctx, [&](Context &c) {
        if (capacity < 0) throw std::domain_error("negative capacity");
        ensure_capacity(impl(c), capacity, with_affine != 0, true);
        FFB_CUDA(cudaStreamSynchronize(c.stream));
    });
}

int ffb200_set_num_particles(ffb200_context *ctx, int n, int has_affine) {
    return guarded("ffb200_set_num_particles", ctx, [&](Context &c) {
        if (n < 0 || n > c.cap) throw std::domain_error("particle count exceeds the reserved capacity");
        if (has_affine && !c.soa[c.cur].a[0]) throw std::logic_error("affine streams were not reserved");
        c.n = n;
        c.has_affine = has_affine != 0;
        c.sorted = false;
    });
}

int ffb200_slab_record_floats(ffb200_context *ctx, int *floats_per_particle) {
    return guarded("ffb200_slab_record_floats", ctx, [&](Context &c) {
        if (!floats_per_particle) throw std::invalid_argument("null output pointer");
        *floats_per_particle = slab_rows(c);
    }, false);
}

int ffb200_slab_pack_layers(ffb200_context *ctx, int lo_a, int hi_a, float *block_a, int lo_b, int hi_b, float *block_b,
                            int block_capacity) {
    return guarded("ffb200_slab_pack_layers", ctx, [&](Context &c) {
        if (block_capacity <= 0) throw std::domain_error("block capacity must be positive");
        launch_pack_layers(c, lo_a, hi_a, block_a, lo_b, hi_b, block_b, block_capacity);
    });
}

int ffb200_slab_route(ffb200_context *ctx, int k_begin, int k_end, float *block_up, float *block_down, int block_capacity,
                      int *counts) {
    return guarded("ffb200_slab_route", ctx, [&](Context &c) {
        if (!counts) throw std::invalid_argument("null counts pointer");
        if ((block_up || block_down) && block_capacity <= 0) throw std::domain_error("block capacity must be positive");
        launch_route_begin(c, k_begin, k_end, block_up, block_down, block_capacity);
        launch_route_end(c, counts);
    });
}

int ffb200_slab_route_begin(ffb200_context *ctx, int k_begin, int k_end, float *block_up, float *block_down,
                            int block_capacity) {
    return guarded("ffb200_slab_route_begin", ctx, [&](Context &c) {
        if ((block_up || block_down) && block_capacity <= 0) throw std::domain_error("block capacity must be positive");
        launch_route_begin(c, k_begin, k_end, block_up, block_down, block_capacity);
    });
}

int ffb200_slab_route_ghosts_begin(ffb200_context *ctx, int k_begin, int k_end, int ghost_layers, float *block_up,
                                   float *block_down, const int *capacities) {
    return guarded("ffb200_slab_route_ghosts_begin", ctx, [&](Context &c) {
        if (ghost_layers <= 0) throw std::domain_error("ghost layer count must be positive");
        if (!capacities) throw std::invalid_argument("null capacities pointer");
        if ((block_up && (capacities[0] <= 0 || capacities[1] <= 0)) || (block_down && (capacities[2] <= 0 || capacities[3] <= 0)))
            throw std::domain_error("section capacities must be positive");
        launch_route_begin(c, k_begin, k_end, block_up, block_down, 0, ghost_layers, capacities);
    });
}

int ffb200_slab_route_end_known(ffb200_context *ctx, int leaving, int *counts) {
    return guarded("ffb200_slab_route_end_known", ctx, [&](Context &c) {
        if (!counts) throw std::invalid_argument("null counts pointer");
        if (leaving < 0) throw std::domain_error("negative count of leaving particles");
        launch_route_end(c, counts, leaving);
    });
}

int ffb200_slab_route_end(ffb200_context *ctx, int *counts) {
    return guarded("ffb200_slab_route_end", ctx, [&](Context &c) {
        if (!counts) throw std::invalid_argument("null counts pointer");
        launch_route_end(c, counts);
    });
}

int ffb200_slab_append(ffb200_context *ctx, const float *block, int count, int as_ghost) {
    return guarded("ffb200_slab_append", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        if (count < 0 || (count > 0 && !block)) throw std::invalid_argument("bad packed block");
        if (c.n + count > c.cap) ensure_capacity(c, c.n + count + (c.n + count) / 8, c.has_affine, true);
        launch_append(c, block, count, as_ghost != 0);
    });
}

int ffb200_sort_particles(ffb200_context *ctx) {
    return guarded("ffb200_sort_particles", ctx, [&](Context &c) { sort_impl(impl(c)); });
}

int ffb200_get_binning(ffb200_context *ctx, int32_t *cell, uint32_t *hkey, uint32_t *perm) {
    return guarded("ffb200_get_binning", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        if (!cell || !hkey || !perm) throw std::invalid_argument("null output pointer");
        sort_impl(c);
        const int n = c.n;
        if (n == 0) return;
        // key/val scratch is free after the sort: reuse it for the dump
        int32_t *d_cell = reinterpret_cast<int32_t *>(c.sort.key[1]);
        uint32_t *d_hkey = c.sort.val[1], *d_perm = reinterpret_cast<uint32_t *>(c.aos_stage);
        launch_binning_dump(c, d_cell, d_hkey, d_perm);
        FFB_CUDA(cudaMemcpyAsync(cell, d_cell, (size_t)n * 4, cudaMemcpyDeviceToHost, c.stream));
        FFB_CUDA(cudaMemcpyAsync(hkey, d_hkey, (size_t)n * 4, cudaMemcpyDeviceToHost, c.stream));
        FFB_CUDA(cudaMemcpyAsync(perm, d_perm, (size_t)n * 4, cudaMemcpyDeviceToHost, c.stream));
        FFB_CUDA(cudaStreamSynchronize(c.stream));
    });
}

int ffb200_set_velocity_field(ffb200_context *ctx, const float *u, const float *v, const float *w) {
    return guarded("ffb200_set_velocity_field", ctx, [&](Context &c) { upload_field(impl(c), false, u, v, w); });
}

int ffb200_set_saved_velocity_field(ffb200_context *ctx, const float *u, const float *v, const float *w) {
    return guarded("ffb200_set_saved_velocity_field", ctx, [&](Context &c) { upload_field(impl(c), true, u, v, w); });
}

int ffb200_get_velocity_field(ffb200_context *ctx, float *u, float *v, float *w, uint8_t *validu, uint8_t *validv,
                              uint8_t *validw) {
    return guarded("ffb200_get_velocity_field", ctx,
                   [&](Context &c) { get_field_impl(impl(c), u, v, w, validu, validv, validw); }, false);
}

int ffb200_get_weight_sums(ffb200_context *ctx, float *wu, float *wv, float *ww) {
    return guarded("ffb200_get_weight_sums", ctx, [&](Context &c) {
        float *h[3] = {wu, wv, ww};
        for (int d = 0; d < 3; d++) {
            FaceGrid &f = c.face[d];
            const size_t off = (size_t)f.gi * f.gj * c.g.kbase;
            if (h[d]) FFB_CUDA(cudaMemcpyAsync(h[d] + off, f.wsum, f.count * 4, cudaMemcpyDeviceToHost, c.stream));
        }
        FFB_CUDA(cudaStreamSynchronize(c.stream));
    }, false);
}

int ffb200_save_velocity_field(ffb200_context *ctx) {
    return guarded("ffb200_save_velocity_field", ctx, [&](Context &c) {
        for (int d = 0; d < 3; d++)
            FFB_CUDA(cudaMemcpyAsync(c.face[d].saved, c.face[d].vel, c.face[d].count * 4, cudaMemcpyDeviceToDevice, c.stream));
    }, false);
}

int ffb200_set_valid_velocities(ffb200_context *ctx, const uint8_t *validu, const uint8_t *validv, const uint8_t *validw) {
    return guarded("ffb200_set_valid_velocities", ctx, [&](Context &c) {
        const uint8_t *h[3] = {validu, validv, validw};
        for (int d = 0; d < 3; d++) {
            if (!h[d]) throw std::invalid_argument("null valid-mask pointer");
            FaceGrid &f = c.face[d];
            const size_t off = (size_t)f.gi * f.gj * c.g.kbase;
            FFB_CUDA(cudaMemcpyAsync(f.valid, h[d] + off, f.count, cudaMemcpyHostToDevice, c.stream));
        }
    });
}

int ffb200_extrapolate_velocity_field(ffb200_context *ctx, int num_layers) {
    return guarded("ffb200_extrapolate_velocity_field", ctx, [&](Context &c) {
        if (num_layers < 0) throw std::domain_error("negative layer count");
        launch_extrapolate(c, num_layers);
    });
}

int ffb200_declare_resident(ffb200_context *ctx, unsigned mask) {
    return guarded("ffb200_declare_resident", ctx, [&](Context &c) { c.resident_next = mask; }, false, true);
}

int ffb200_get_maximum_particle_speed(ffb200_context *ctx, double *speed) {
    return guarded("ffb200_get_maximum_particle_speed", ctx, [&](Context &c) {
        if (!speed) throw std::invalid_argument("null output pointer");
        uint32_t *dev = reinterpret_cast<uint32_t *>(c.slab_counters) + 7;     // a spare device word
        launch_max_speed_sq(c, dev);
        uint32_t bits = 0;
        FFB_CUDA(cudaMemcpyAsync(&bits, dev, sizeof(bits), cudaMemcpyDeviceToHost, c.stream));
        FFB_CUDA(cudaStreamSynchronize(c.stream));
        float f;
        std::memcpy(&f, &bits, sizeof(f));
        *speed = std::sqrt((double)f);                         // sqrt(maxsq), maxsq the double of a float dot product
    }, false);
}

int ffb200_remove_marker_particles(ffb200_context *ctx, double dt, double cfl_condition_number, int max_particles_per_cell,
                                   int max_frame_time_steps, int extreme_velocity_removal, const float *open_bounds,
                                   int *num_remaining, int *num_extreme_removed) {
    return guarded("ffb200_remove_marker_particles", ctx, [&](Context &c) {
        int remaining = 0, extreme = 0;
        launch_remove_particles(c, remove_rules(dt, cfl_condition_number, max_particles_per_cell, max_frame_time_steps,
                                                extreme_velocity_removal, open_bounds),
                                &remaining, &extreme);
        if (num_remaining) *num_remaining = remaining;
        if (num_extreme_removed) *num_extreme_removed = extreme;
    });
}

int ffb200_set_solid(ffb200_context *ctx, const float *phi, const uint8_t *near_solid) {
    return guarded("ffb200_set_solid", ctx, [&](Context &c) { set_solid_impl(impl(c), phi, near_solid); });
}

int ffb200_set_solid_device(ffb200_context *ctx, const float *d_phi, const uint8_t *d_near_solid) {
    return guarded("ffb200_set_solid_device", ctx, [&](Context &c) {
        if (!d_phi || !d_near_solid) throw std::invalid_argument("null solid SDF / near-solid pointer");
        const GridDesc &g = c.g;
        const size_t plane = (size_t)(g.I + 1) * (g.J + 1);
        FFB_CUDA(cudaMemcpyAsync(c.phi, d_phi, plane * (g.kloc + 1) * 4, cudaMemcpyDeviceToDevice, c.stream));
        FFB_CUDA(cudaMemcpyAsync(c.near_solid, d_near_solid, (size_t)c.ni * c.nj * c.nk, cudaMemcpyDeviceToDevice, c.stream));
        launch_solid_clearance(c);
        c.has_solid = true;
    });
}

int ffb200_p2g(ffb200_context *ctx, double particle_radius, int transfer_method) {
    return guarded("ffb200_p2g", ctx, [&](Context &c) { p2g_impl(impl(c), particle_radius, transfer_method); });
}

int ffb200_g2p(ffb200_context *ctx, int transfer_method, double ratio_pic_flip) {
    return guarded("ffb200_g2p", ctx, [&](Context &c) { g2p_impl(impl(c), transfer_method, ratio_pic_flip); });
}

int ffb200_advect(ffb200_context *ctx, double dt, double cfl_condition_number, int resolve_collisions) {
    // not counted as a mutation at entry: an advection that directly follows ffb200_g2p reuses its samples
    return guarded(
        "ffb200_advect", ctx,
        [&](Context &c) {
            advect_impl(impl(c), dt, cfl_condition_number, resolve_collisions);
            c.epoch++;
        },
        false);
}

int ffb200_velocity_advector_advect(ffb200_context *ctx, int n, const float *pos, const float *vel, const float *affx,
                                    const float *affy, const float *affz, double particle_radius, int transfer_method,
                                    float *u, float *v, float *w, uint8_t *validu, uint8_t *validv, uint8_t *validw) {
    return guarded("ffb200_velocity_advector_advect", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        const unsigned res = c.resident_arg;
        if (res & FFB200_RESIDENT_PARTICLES) {                 // the caller vouches: the device holds exactly these particles
            if (n != c.n) throw std::logic_error("FFB200_RESIDENT_PARTICLES: the particle count differs from the resident set");
        } else {
            set_particles_impl(c, n, pos, vel, affx, affy, affz);
        }
        // null outputs stay on the device (ffb200_get_velocity_field fetches them later); the others are copied out
        // direction by direction behind the transfer kernels
        HostFieldOut host;
        host.vel[0] = u; host.vel[1] = v; host.vel[2] = w;
        host.valid[0] = validu; host.valid[1] = validv; host.valid[2] = validw;
        const bool any = u || v || w || validu || validv || validw;
        p2g_impl(c, particle_radius, transfer_method, any ? &host : nullptr);
    });
}

int ffb200_extrapolate_fluid_velocities(ffb200_context *ctx, float *u, float *v, float *w, const uint8_t *validu,
                                        const uint8_t *validv, const uint8_t *validw, int num_layers,
                                        int device_field_is_current) {
    return guarded("ffb200_extrapolate_fluid_velocities", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        if (num_layers < 0) throw std::domain_error("negative layer count");
        if (!u || !v || !w) throw std::invalid_argument("null velocity field pointer");
        if (!device_field_is_current) {
            if (!validu || !validv || !validw) throw std::invalid_argument("null valid-mask pointer");
            StageTimer t(c, kH2D);
            upload_field(c, false, u, v, w);
            const uint8_t *h[3] = {validu, validv, validw};
            for (int d = 0; d < 3; d++) {
                FaceGrid &f = c.face[d];
                FFB_CUDA(cudaMemcpyAsync(f.valid, h[d] + (size_t)f.gi * f.gj * c.g.kbase, f.count, cudaMemcpyHostToDevice, c.stream));
            }
            t.done(0);
        }
        launch_extrapolate(c, num_layers);
        get_field_impl(c, u, v, w, nullptr, nullptr, nullptr);
    });
}

int ffb200_update_marker_particle_velocities(ffb200_context *ctx, int n, const float *pos, float *vel, float *affx,
                                             float *affy, float *affz, const float *u, const float *v, const float *w,
                                             const float *su, const float *sv, const float *sw, int transfer_method,
                                             double ratio_pic_flip) {
    return guarded("ffb200_update_marker_particle_velocities", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        const bool apic = transfer_method == FFB200_TRANSFER_APIC;
        const unsigned res = c.resident_arg;
        const bool lazy = (res & FFB200_RESIDENT_PARTICLES) && !vel;      // outputs stay on the device
        if (apic && !lazy && (!affx || !affy || !affz)) throw std::invalid_argument("APIC needs affine output buffers");
        if (!apic && !(res & FFB200_RESIDENT_SAVED_FIELD) && (!su || !sv || !sw))
            throw std::invalid_argument("FLIP needs the saved velocity field");
        if (res & FFB200_RESIDENT_PARTICLES) {                 // the caller vouches: same particles as the previous call
            if (n != c.n) throw std::logic_error("FFB200_RESIDENT_PARTICLES: the particle count differs from the resident set");
        } else {
            set_particles_impl(c, n, pos, vel, nullptr, nullptr, nullptr);
        }
        sort_impl(c);                                          // spatial order for the gathers (no-op if still sorted)
        const float *const h_cur[3] = {u, v, w}, *const h_saved[3] = {su, sv, sw};
        const bool up_cur = !(res & FFB200_RESIDENT_FIELD), up_saved = !apic && !(res & FFB200_RESIDENT_SAVED_FIELD);
        if (up_cur && (!u || !v || !w)) throw std::invalid_argument("null velocity field pointer");
        if (up_saved && !up_cur) upload_field(c, true, su, sv, sw);
        g2p_impl(c, transfer_method, ratio_pic_flip, up_cur ? h_cur : nullptr, up_cur && up_saved ? h_saved : nullptr);
        // resident particles with null outputs: the results stay on the device until ffb200_get_particles asks
        if (vel || !(res & FFB200_RESIDENT_PARTICLES))
            get_particles_impl(c, nullptr, vel, apic ? affx : nullptr, apic ? affy : nullptr, apic ? affz : nullptr);
    });
}

int ffb200_advance_marker_particles(ffb200_context *ctx, int n, float *pos, const float *u, const float *v,
                                    const float *w, const float *phi, const uint8_t *near_solid, double dt,
                                    double cfl_condition_number) {
    // epoch handling by hand: with resident particles and field this advection directly follows the G2P
    // of the same state and may reuse its samples (see ffb200_advect)
    return guarded(
        "ffb200_advance_marker_particles", ctx,
        [&](Context &cc) {
            ContextImpl &c = impl(cc);
            const unsigned res = c.resident_arg;
            if (n > 0 && !pos && !(res & FFB200_RESIDENT_PARTICLES)) throw std::invalid_argument("null position pointer");
            const unsigned both = FFB200_RESIDENT_PARTICLES | FFB200_RESIDENT_FIELD;
            if ((res & both) != both) c.epoch++;
            if (res & FFB200_RESIDENT_PARTICLES) {
                if (n != c.n) throw std::logic_error("FFB200_RESIDENT_PARTICLES: the particle count differs from the resident set");
            } else {
                // velocities are not needed by advection
                ensure_capacity(c, n, false);
                c.n = n;
                c.has_affine = false;
                c.sorted = false;
                if (n > 0) {
                    StageTimer t(c, kH2D);
                    upload_attr(c, pos, c.soa[c.cur].p, n);
                    launch_iota(c, c.soa[c.cur].orig, n);
                    t.done(0);
                }
            }
            if (!(res & FFB200_RESIDENT_FIELD)) upload_field(c, false, u, v, w);
            if (phi && near_solid) set_solid_impl(c, phi, near_solid);
            sort_impl(c);
            advect_impl(c, dt, cfl_condition_number, 1);
            c.epoch++;
            if (pos) get_particles_impl(c, pos, nullptr, nullptr, nullptr, nullptr);   // null + resident: stays on the device
        },
        false);
}

int ffb200_mark_removed_marker_particles(ffb200_context *ctx, int n, const float *pos, const float *vel, const float *phi,
                                         const uint8_t *near_solid, const float *open_bounds, const uint8_t *pre_removed,
                                         double dt, double cfl_condition_number, int max_particles_per_cell,
                                         int max_frame_time_steps, int extreme_velocity_removal, uint8_t *removed,
                                         int *num_removed, int *num_extreme_removed) {
    return guarded("ffb200_mark_removed_marker_particles", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        if (n < 0) throw std::domain_error("negative particle count");
        if (n > 0 && !removed) throw std::invalid_argument("null output mask");
        const unsigned res = c.resident_arg;
        if (n == 0 && !(res & FFB200_RESIDENT_PARTICLES)) {     // nothing to decide; the resident set is left alone
            if (num_removed) *num_removed = 0;
            if (num_extreme_removed) *num_extreme_removed = 0;
            return;
        }
        if (res & FFB200_RESIDENT_PARTICLES) {
            if (n != c.n) throw std::logic_error("FFB200_RESIDENT_PARTICLES: the particle count differs from the resident set");
        } else {
            if (n > 0 && (!pos || !vel)) throw std::invalid_argument("positions and velocities are required");
            ensure_capacity(c, n, false);
            c.n = n;
            c.has_affine = false;
            c.sorted = false;
            if (n > 0) {
                StageTimer t(c, kH2D);
                upload_attr(c, pos, c.soa[c.cur].p, n);
                upload_attr(c, vel, c.soa[c.cur].v, n);
                launch_iota(c, c.soa[c.cur].orig, n);
                t.done(0);
            }
        }
        if (!(res & FFB200_RESIDENT_SOLID)) set_solid_impl(c, phi, near_solid);
        // byte scratch in the AoS staging buffer (12 bytes per particle of capacity): [pre-removed | removed]
        uint8_t *bytes = reinterpret_cast<uint8_t *>(c.aos_stage);
        uint8_t *d_pre = nullptr, *d_removed = bytes + (size_t)c.cap * 4;
        if (pre_removed && n > 0) {
            d_pre = bytes;
            FFB_CUDA(cudaMemcpyAsync(d_pre, pre_removed, (size_t)n, cudaMemcpyHostToDevice, c.stream));
        }
        int remaining = 0, extreme = 0;
        if (n > 0) {
            launch_remove_mask(c, remove_rules(dt, cfl_condition_number, max_particles_per_cell, max_frame_time_steps,
                                               extreme_velocity_removal, open_bounds),
                               d_pre, d_removed, &remaining, &extreme);
            StageTimer t(c, kD2H);
            FFB_CUDA(cudaMemcpyAsync(removed, d_removed, (size_t)n, cudaMemcpyDeviceToHost, c.stream));
            t.done(0);
            FFB_CUDA(cudaStreamSynchronize(c.stream));
        }
        if (num_removed) *num_removed = n - remaining;
        if (num_extreme_removed) *num_extreme_removed = extreme;
    });
}

int ffb200_remove_marker_particles_masked(ffb200_context *ctx, const float *open_bounds, const uint8_t *pre_removed, double dt,
                                          double cfl_condition_number, int max_particles_per_cell, int max_frame_time_steps,
                                          int extreme_velocity_removal, uint8_t *removed, int *num_remaining,
                                          int *num_extreme_removed) {
    return guarded("ffb200_remove_marker_particles_masked", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        const int n = c.n;
        if (n > 0 && !removed) throw std::invalid_argument("null output mask");
        uint8_t *bytes = reinterpret_cast<uint8_t *>(c.aos_stage);     // 12 bytes per particle of capacity: [pre-removed | removed]
        uint8_t *d_pre = nullptr, *d_removed = bytes + (size_t)c.cap * 4;
        if (pre_removed && n > 0) {
            d_pre = bytes;
            FFB_CUDA(cudaMemcpyAsync(d_pre, pre_removed, (size_t)n, cudaMemcpyHostToDevice, c.stream));
        }
        int remaining = n, extreme = 0;
        if (n > 0) {
            launch_remove_particles(c, remove_rules(dt, cfl_condition_number, max_particles_per_cell, max_frame_time_steps,
                                                    extreme_velocity_removal, open_bounds),
                                    &remaining, &extreme, d_pre, d_removed);
            if (remaining != n) {                                  // the usual substep removes nothing: no mask traffic then
                StageTimer t(c, kD2H);
                FFB_CUDA(cudaMemcpyAsync(removed, d_removed, (size_t)n, cudaMemcpyDeviceToHost, c.stream));
                t.done(0);
                FFB_CUDA(cudaStreamSynchronize(c.stream));
            } else {
                std::memset(removed, 0, (size_t)n);
            }
        }
        if (num_remaining) *num_remaining = remaining;
        if (num_extreme_removed) *num_extreme_removed = extreme;
    });
}

int ffb200_pin_host_memory(ffb200_context *ctx, void *ptr, size_t bytes) {
    return guarded("ffb200_pin_host_memory", ctx, [&](Context &) {
        if (!ptr || bytes == 0) throw std::invalid_argument("null host range");
        cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
        if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return; }
        if (e != cudaSuccess) { cudaGetLastError(); throw CudaError(std::string("cudaHostRegister failed: ") + cudaGetErrorString(e)); }
    }, false);
}

int ffb200_unpin_host_memory(ffb200_context *ctx, void *ptr) {
    return guarded("ffb200_unpin_host_memory", ctx, [&](Context &) {
        if (!ptr) throw std::invalid_argument("null host pointer");
        cudaError_t e = cudaHostUnregister(ptr);
        if (e != cudaSuccess) cudaGetLastError();                  // not registered (any more): nothing to undo
    }, false);
}

int ffb200_attribute_to_grid_transfer(ffb200_context *ctx, int n, const float *pos, const float *attr, int num_components,
                                      double particle_radius, int normalize, float *grid, uint8_t *valid) {
    return guarded("ffb200_attribute_to_grid_transfer", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        if (num_components != 1 && num_components != 3) throw std::domain_error("1 (scalar) or 3 (vmath::vec3) attribute components");
        if (!(particle_radius > 0.0)) throw std::domain_error("particle radius must be positive");
        if (n < 0) throw std::domain_error("negative particle count");
        if (!grid || !valid) throw std::invalid_argument("null output pointer");
        if (n > 0 && (!pos || !attr)) throw std::invalid_argument("null position / attribute pointer");
        ensure_capacity(c, n, false);
        c.n = n;
        c.has_affine = false;
        c.sorted = false;
        const size_t cells = (size_t)c.g.I * c.g.J * c.g.K;
        if (n > 0) {
            // the payload rides in the velocity streams through the sort
            ParticleSoA &s = c.soa[c.cur];
            StageTimer t(c, kH2D);
            upload_attr(c, pos, s.p, n);
            if (num_components == 3)
                upload_attr(c, attr, s.v, n);
            else
                FFB_CUDA(cudaMemcpyAsync(s.v[0], attr, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c.stream));
            launch_iota(c, s.orig, n);
            t.done(0);
        }
        sort_impl(c);
        float *d_out = nullptr;
        uint8_t *d_valid = nullptr;
        FFB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_out), cells * num_components * sizeof(float)));
        FFB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_valid), cells));
        try {
            launch_attribute_p2g(c, particle_radius, num_components, normalize, d_out, d_valid);
            FFB_CUDA(cudaMemcpyAsync(grid, d_out, cells * num_components * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
            FFB_CUDA(cudaMemcpyAsync(valid, d_valid, cells, cudaMemcpyDeviceToHost, c.stream));
            FFB_CUDA(cudaStreamSynchronize(c.stream));
        } catch (...) {
            cudaFree(d_out);
            cudaFree(d_valid);
            throw;
        }
        cudaFree(d_out);
        cudaFree(d_valid);
    });
}

int ffb200_liquid_sdf(ffb200_context *ctx, double particle_radius) {
    return guarded("ffb200_liquid_sdf", ctx, [&](Context &c) { launch_liquid_sdf(c, particle_radius); }, false);
}

int ffb200_postprocess_liquid_sdf(ffb200_context *ctx) {
    return guarded("ffb200_postprocess_liquid_sdf", ctx, [&](Context &c) { launch_liquid_sdf_postprocess(c); }, false);
}

int ffb200_get_liquid_sdf(ffb200_context *ctx, float *phi) {
    return guarded("ffb200_get_liquid_sdf", ctx, [&](Context &c) {
        if (!phi) throw std::invalid_argument("null output pointer");
        if (!c.liquid_phi) throw std::logic_error("no liquid SDF on the device (ffb200_liquid_sdf first)");
        const size_t cells = (size_t)c.g.I * c.g.J * c.g.K;
        FFB_CUDA(cudaMemcpyAsync(phi, c.liquid_phi, cells * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
        FFB_CUDA(cudaStreamSynchronize(c.stream));
    }, false);
}

int ffb200_calculate_signed_distance_field(ffb200_context *ctx, int n, const float *pos, double particle_radius, float *phi) {
    return guarded("ffb200_calculate_signed_distance_field", ctx, [&](Context &cc) {
        ContextImpl &c = impl(cc);
        if (n < 0) throw std::domain_error("negative particle count");
        if (!phi) throw std::invalid_argument("null output pointer");
        const unsigned res = c.resident_arg;
        if (res & FFB200_RESIDENT_PARTICLES) {
            if (n != c.n) throw std::logic_error("FFB200_RESIDENT_PARTICLES: the particle count differs from the resident set");
        } else {
            if (n > 0 && !pos) throw std::invalid_argument("null position pointer");
            ensure_capacity(c, n, false);                      // positions only, like the advection entry point
            c.n = n;
            c.has_affine = false;
            c.sorted = false;
            if (n > 0) {
                StageTimer t(c, kH2D);
                upload_attr(c, pos, c.soa[c.cur].p, n);
                launch_iota(c, c.soa[c.cur].orig, n);
                t.done(0);
            }
        }
        launch_liquid_sdf(c, particle_radius);
        const size_t cells = (size_t)c.g.I * c.g.J * c.g.K;
        FFB_CUDA(cudaMemcpyAsync(phi, c.liquid_phi, cells * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
        FFB_CUDA(cudaStreamSynchronize(c.stream));
    });
}

}  // extern "C"
