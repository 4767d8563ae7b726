// api.cu -- C-ABI entry points of libcuda_render.so (see include/xray_cuda_render.h).
//
// Host orchestration around the kernels: scene upload (cached per device), fp64 sample-lattice
// tables, camera preparation, view batching with pinned double-buffered D2H, modulo view
// sharding over several GPUs (one host thread per device, volume replicated over NVLink P2P),
// and the three legacy symbols of the reference plugin (cuda_backend.h).
//
// Contract kept from the reference plugin (cuda_backend.cu:82-203): synchronous unless the
// *Device* variant is used, 0 on success, small positive codes on failure, never exit/abort,
// nothing printed to stdout, callable from any OS thread (device set explicitly per call).
#include <cuda_runtime.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define XRAY_STREAM_COPY 1
#endif

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/xray_cuda_render.h"
#include "device_types.h"
#include "scene.h"

namespace xr {
cudaError_t launch_render_scene(const RenderParams& P, int precision, int integrator, cudaStream_t stream);
cudaError_t launch_voxelize_scene(const RenderParams& P, int res, float* d_out, cudaStream_t stream);
cudaError_t launch_clip_probe(const RenderParams& P, cudaStream_t stream);
cudaError_t launch_render_volume_fast(const float* d_vol, int nx, int ny, int nz, const RenderParams& P, cudaStream_t stream);
cudaError_t launch_render_fast(const RenderParams& P, int shape, int integrator, bool count, bool list, int prim,
                               const unsigned char* d_nfine, int i_coll, int i_tess, cudaStream_t stream);
size_t fast_kernel_smem_bytes(const RenderParams& P);
cudaError_t launch_render_async(const RenderParams& P, int shape, int integrator, bool count, int prim, const unsigned char* d_nfine,
                                int i_coll, int i_tess, cudaStream_t stream);
cudaError_t launch_render_span(const RenderParams& P, int integrator, bool count, const unsigned char* d_nfine, const unsigned char* d_section,
                               unsigned int section_bytes, unsigned int* d_tile_list, unsigned int* d_tile_count, unsigned int* d_bins,
                               unsigned int bin_cap, unsigned int n_instances, cudaStream_t stream);
size_t span_kernel_smem_bytes(unsigned int section_bytes);
cudaError_t launch_render_volume_tex(unsigned long long tex, const float* d_vol, int nx, int ny, int nz, const RenderParams& P,
                                     int warp_shape, const unsigned char* occ, int integrator, const unsigned char* d_nfine,
                                     cudaStream_t stream);
cudaError_t launch_render_volume_f64(const void* d_vol, int dtype, int nx, int ny, int nz, const RenderParams& P, int integrator,
                                     const unsigned char* d_nfine, cudaStream_t stream);
size_t volume_brick_count(int nx, int ny, int nz);
cudaError_t build_volume_occupancy(const float* d_vol, int nx, int ny, int nz, unsigned char* occ_a, unsigned char* occ_b,
                                   const unsigned char** result, unsigned long long surf, cudaStream_t stream);
cudaError_t launch_voxelize_cylinders(const CylinderParams* d_cyl, int n, int res, float dm, const int* d_off,
                                      const int* d_idx, int grid_dim, float* d_out, cudaStream_t stream);
cudaError_t measure_fp32_peak(double* tflops);
}  // namespace xr

using namespace xr;

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
static int fail_cuda(int code, const char* what, cudaError_t e) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();  // clear sticky-less errors
    return code;
}
#define CU(code, call)                                          \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return fail_cuda(code, #call, e__); \
    } while (0)

static const double kCubeHalfDiagonal = 1.74;  // main.go:46

// ---------------------------------------------------------------------------------------
// Device-side scene cache
// ---------------------------------------------------------------------------------------
struct DevScene {
    int dev = -1;
    unsigned char* d_blob = nullptr;
    VoxelDev* d_vox_table = nullptr;
    void* d_vox[kMaxVoxelSlots] = {nullptr};
    size_t vox_bytes[kMaxVoxelSlots] = {0};
    uint64_t vox_version[kMaxVoxelSlots] = {0};
    bool vox_borrowed[kMaxVoxelSlots] = {false};
};
struct SceneCache {
    std::mutex mu;
    std::map<int, DevScene> per_dev;
};

static SceneCache* cache_of(XRayScene* sc) {
    if (!sc->device_cache) sc->device_cache = new SceneCache();
    return (SceneCache*)sc->device_cache;
}

static void free_dev_scene(DevScene& ds) {
    if (ds.dev < 0) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(ds.dev);
    if (ds.d_blob) cudaFree(ds.d_blob);
    if (ds.d_vox_table) cudaFree(ds.d_vox_table);
    for (int s = 0; s < kMaxVoxelSlots; ++s)
        if (ds.d_vox[s] && !ds.vox_borrowed[s]) cudaFree(ds.d_vox[s]);
    cudaSetDevice(cur);
    ds = DevScene();
}

// ---------------------------------------------------------------------------------------
// Host -> device upload of a big pageable buffer (the legacy ABI hands over Go heap memory, cuda_backend.go:316):
// cudaMemcpy from pageable memory stages through one driver thread at ~10 GB/s.  Here T host threads each own a
// contiguous slice, copy it piecewise into their own pinned double buffer and push it on their own stream, so the
// CPU copies and the DMA overlap and the link is the limit.  Pinned sources go out in one async copy.
// ---------------------------------------------------------------------------------------
static std::mutex g_stage_mu;
static const int kStageThreads = 16;
static const size_t kStageChunk = (size_t)8 << 20;
static void* g_stage_buf[kStageThreads][2] = {};
static int g_stage_dev = -1;

static void stage_release_locked() {
    for (int t = 0; t < kStageThreads; ++t)
        for (int b = 0; b < 2; ++b) {
            if (g_stage_buf[t][b]) cudaFreeHost(g_stage_buf[t][b]);
            g_stage_buf[t][b] = nullptr;
        }
    g_stage_dev = -1;
}

// Large host-to-host copy with streaming (non-temporal) stores. The staging copies move tens of megabytes that nobody
// reads back soon; a cached store first reads the destination line for ownership, so memcpy's per-thread pieces (below
// glibc's own non-temporal threshold) cost 3 bytes of DRAM traffic per byte copied where this costs 2. SSE2 is part of
// the x86-64 baseline, so there is no dispatch; other hosts take memcpy. Measured with 8 threads on 64 MiB blocks:
// 21-29 GB/s (memcpy) -> 29-40 GB/s.
static void stream_copy(void* dst, const void* src, size_t n) {
#ifdef XRAY_STREAM_COPY
    if (n < ((size_t)256 << 10) || getenv("XRAY_NO_STREAM_COPY")) {
        memcpy(dst, src, n);
        return;
    }
    unsigned char* d = (unsigned char*)dst;
    const unsigned char* s = (const unsigned char*)src;
    const size_t head = (size_t)((16 - ((uintptr_t)d & 15)) & 15);
    memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
    const size_t blocks = n >> 7;
    for (size_t i = 0; i < blocks; ++i, s += 128, d += 128) {
        const __m128i v0 = _mm_loadu_si128((const __m128i*)(s)), v1 = _mm_loadu_si128((const __m128i*)(s + 16));
        const __m128i v2 = _mm_loadu_si128((const __m128i*)(s + 32)), v3 = _mm_loadu_si128((const __m128i*)(s + 48));
        const __m128i v4 = _mm_loadu_si128((const __m128i*)(s + 64)), v5 = _mm_loadu_si128((const __m128i*)(s + 80));
        const __m128i v6 = _mm_loadu_si128((const __m128i*)(s + 96)), v7 = _mm_loadu_si128((const __m128i*)(s + 112));
        _mm_stream_si128((__m128i*)(d), v0);
        _mm_stream_si128((__m128i*)(d + 16), v1);
        _mm_stream_si128((__m128i*)(d + 32), v2);
        _mm_stream_si128((__m128i*)(d + 48), v3);
        _mm_stream_si128((__m128i*)(d + 64), v4);
        _mm_stream_si128((__m128i*)(d + 80), v5);
        _mm_stream_si128((__m128i*)(d + 96), v6);
        _mm_stream_si128((__m128i*)(d + 112), v7);
    }
    _mm_sfence();  // the streamed lines are globally visible before anyone (a DMA engine, the caller) is told so
    memcpy(d, s, n & 127);
#else
    memcpy(dst, src, n);
#endif
}

// Host-side copy out of pinned staging into the caller's pageable buffer, split over the host's cores: one core moves
// ~8-10 GB/s, the PCIe link delivers ~55, and a cgo / ctypes caller always hands over pageable memory.
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    int T = (int)std::min<size_t>(16, std::max<size_t>(1, bytes / ((size_t)4 << 20)));
    unsigned hc = std::thread::hardware_concurrency();
    bool shared_host = false;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {  // one process per GPU (torchrun): the ranks share the host's cores
        const int lw = atoi(e);
        if (lw > 1) {
            hc = std::max(1u, hc / (unsigned)lw);
            shared_host = true;
        }
    }
    // The caller is blocked in this call, so every core of its share may copy (16-core host, 64 MiB blocks: 8 threads
    // 20.3 ms, 12 threads 18.8, 16 threads 17.6 per 512 MB of images; two ranks on that host: 4 / 8 / 16 threads per rank
    // give 4895 / 5080 / 5140 Gsamples/s end to end on the lattice, 1649 / 1707 / 1804 on the pillar array).
    if (hc > 0) T = std::min<int>(T, std::max(shared_host ? 2u : 1u, hc));
    if (const char* e = getenv("XRAY_DRAIN_THREADS")) T = std::max(1, std::min(32, atoi(e)));
    if (T <= 1) {
        stream_copy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / T) + 4095) & ~(size_t)4095;
    for (int t = 1; t < T; ++t) {
        const size_t off = std::min(bytes, per * t), end = std::min(bytes, per * (t + 1));
        if (end > off) th.emplace_back([=] { stream_copy((unsigned char*)dst + off, (const unsigned char*)src + off, end - off); });
    }
    stream_copy(dst, src, std::min(bytes, per));
    for (auto& t : th) t.join();
}

static cudaError_t upload_h2d(void* d_dst, const void* h_src, size_t bytes, int dev, cudaStream_t stream) {
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, h_src) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    const unsigned int hw = std::thread::hardware_concurrency();
    int T = (int)std::min<unsigned int>(8u, std::max(1u, hw / 2));
    if (const char* e = getenv("XRAY_STAGE_THREADS")) T = std::max(1, std::min(kStageThreads, atoi(e)));
    if (pinned || bytes < ((size_t)64 << 20) || T < 2 || getenv("XRAY_NO_STAGED_UPLOAD"))
        return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream);
    std::lock_guard<std::mutex> lk(g_stage_mu);
    const auto t_begin = std::chrono::steady_clock::now();
    if (g_stage_dev != dev) stage_release_locked();
    for (int t = 0; t < T; ++t)
        for (int b = 0; b < 2; ++b)
            if (!g_stage_buf[t][b]) {
                cudaError_t e = cudaMallocHost(&g_stage_buf[t][b], kStageChunk);
                if (e != cudaSuccess) {
                    stage_release_locked();
                    cudaGetLastError();
                    return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream);
                }
            }
    g_stage_dev = dev;
    cudaError_t e0 = cudaStreamSynchronize(stream);  // whatever was queued before must not race the side streams
    if (e0 != cudaSuccess) return e0;
    std::vector<cudaError_t> errs((size_t)T, cudaSuccess);
    std::vector<std::thread> th;
    const size_t per = ((bytes / T) + 255) & ~(size_t)255;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            cudaError_t e = cudaSetDevice(dev);
            cudaStream_t st = nullptr;
            cudaEvent_t ev[2] = {nullptr, nullptr};
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
            const size_t lo = std::min(bytes, per * t), hi = std::min(bytes, per * (t + 1));
            int b = 0;
            for (size_t off = lo; off < hi && e == cudaSuccess; off += kStageChunk, b ^= 1) {
                const size_t n = std::min(kStageChunk, hi - off);
                e = cudaEventSynchronize(ev[b]);  // the previous copy out of this buffer is done (no-op the first time)
                if (e != cudaSuccess) break;
                stream_copy(g_stage_buf[t][b], (const unsigned char*)h_src + off, n);
                e = cudaMemcpyAsync((unsigned char*)d_dst + off, g_stage_buf[t][b], n, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
            }
            if (st) {
                cudaError_t e2 = cudaStreamSynchronize(st);
                if (e == cudaSuccess) e = e2;
                cudaStreamDestroy(st);
            }
            for (int q = 0; q < 2; ++q)
                if (ev[q]) cudaEventDestroy(ev[q]);
            errs[(size_t)t] = e;
        });
    for (auto& x : th) x.join();
    for (cudaError_t e : errs)
        if (e != cudaSuccess) return e;
    if (getenv("XRAY_DEBUG_TIMING")) {
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        fprintf(stderr, "[xray] staged upload: %.1f MB in %.1f ms (%.1f GB/s, %d threads)\n", bytes / 1e6, dt * 1e3, bytes / dt / 1e9, T);
    }
    return cudaSuccess;
}

// Upload (or reuse) the program and voxel data on the current device.
// borrowed_vox: optional device pointer to use for slot 0 instead of host data.
// peer_dev >= 0: voxel data that needs (re)loading is pulled from that device's copy over NVLink
// (cudaMemcpyPeerAsync) instead of another H2D of the caller's buffer.
static int ensure_dev_scene(XRayScene* sc, int dev, cudaStream_t stream, const void* borrowed_vox, DevScene** out,
                            int peer_dev = -1) {
    SceneCache* c = cache_of(sc);
    // The lock covers the map only (its nodes are stable): one thread works on one device's entry, and a peer's entry is
    // read-only while others pull from it, so the copies of different devices run side by side (tree replication below).
    DevScene* dsp = nullptr;
    const DevScene* peer_ds = nullptr;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        dsp = &c->per_dev[dev];
        if (peer_dev >= 0 && peer_dev != dev) {
            auto it = c->per_dev.find(peer_dev);
            if (it != c->per_dev.end()) peer_ds = &it->second;
        }
    }
    DevScene& ds = *dsp;
    const Header* h = (const Header*)sc->blob.data();
    if (ds.dev < 0) {
        ds.dev = dev;
        CU(3, cudaMalloc(&ds.d_blob, sc->blob.size()));
        CU(4, cudaMemcpyAsync(ds.d_blob, sc->blob.data(), sc->blob.size(), cudaMemcpyHostToDevice, stream));
        CU(3, cudaMalloc(&ds.d_vox_table, sizeof(VoxelDev) * kMaxVoxelSlots));
    }
    VoxelDev table[kMaxVoxelSlots] = {};
    for (int s = 0; s < (int)h->n_voxel_slots; ++s) {
        VoxelHost& vh = sc->vox[s];
        table[s].nx = vh.nx;
        table[s].ny = vh.ny;
        table[s].nz = vh.nz;
        table[s].dtype = vh.dtype;
        if (s == 0 && borrowed_vox) {
            if (ds.d_vox[0] && !ds.vox_borrowed[0]) cudaFree(ds.d_vox[0]);
            ds.d_vox[0] = const_cast<void*>(borrowed_vox);
            ds.vox_borrowed[0] = true;
            table[0].dtype = 0;
        } else if (vh.data) {
            size_t bytes = (size_t)vh.nx * vh.ny * vh.nz * (vh.dtype == 0 ? 4 : 8);
            if (ds.vox_version[s] != vh.version || ds.vox_bytes[s] != bytes || ds.vox_borrowed[s]) {
                if (ds.d_vox[s] && !ds.vox_borrowed[s] && ds.vox_bytes[s] != bytes) {
                    cudaFree(ds.d_vox[s]);
                    ds.d_vox[s] = nullptr;
                }
                if (ds.vox_borrowed[s]) ds.d_vox[s] = nullptr;
                ds.vox_borrowed[s] = false;
                if (!ds.d_vox[s]) CU(3, cudaMalloc(&ds.d_vox[s], bytes));
                const DevScene* src = nullptr;
                if (peer_ds && peer_ds->d_vox[s] && peer_ds->vox_version[s] == vh.version && peer_ds->vox_bytes[s] == bytes) src = peer_ds;
                if (src) CU(4, cudaMemcpyPeerAsync(ds.d_vox[s], dev, src->d_vox[s], peer_dev, bytes, stream));
                else CU(4, upload_h2d(ds.d_vox[s], vh.data, bytes, dev, stream));
                ds.vox_bytes[s] = bytes;
                ds.vox_version[s] = vh.version;
            }
        }
        table[s].data = ds.d_vox[s];
    }
    CU(4, cudaMemcpyAsync(ds.d_vox_table, table, sizeof(table), cudaMemcpyHostToDevice, stream));
    CU(4, cudaStreamSynchronize(stream));  // table[] is a stack temporary
    *out = &ds;
    return 0;
}

static void fill_scene_dev(const XRayScene* sc, const DevScene& ds, SceneDev& sd) {
    const Header* h = (const Header*)sc->blob.data();
    sd.instr = (const Instr*)(ds.d_blob + h->instr_off);
    sd.f32 = (const float4*)(ds.d_blob + h->f32_off);
    sd.f64 = (const double*)(ds.d_blob + h->f64_off);
    sd.grids = (const unsigned long long*)(ds.d_blob + h->grid_off);
    sd.deform = (const DeformRec*)(ds.d_blob + h->deform_off);
    sd.n_instr = (int)h->n_instr;
    sd.n_deform = (int)h->n_deform;
    sd.f32_count = (int)h->f32_count;
    sd.save_depth = (int)h->save_depth;
    sd.vox = ds.d_vox_table;
}

// ---------------------------------------------------------------------------------------
// fp64 sample lattice (main.go:147 `s += ds`, :176-196 `right += DS`), built by repeated addition.
// ---------------------------------------------------------------------------------------
static void build_lattice(int integrator, double ds, double smin, double smax, std::vector<double>& s_tab) {
    s_tab.clear();
    if (integrator == XRAY_INTEGRATE_SIMPLE) {
        for (double s = smin; s < smax; s += ds) s_tab.push_back(s);
    } else {
        s_tab.push_back(smin);  // left of the first interval
        for (double right = smin + ds; right <= smax; right += ds) s_tab.push_back(right);
    }
}

static double focal_from_fov(double fov_deg) { return 1 / std::tan((fov_deg / 2) * M_PI / 180.0); }  // main.go:457

// Number of fine sub-steps the Go loop takes in coarse interval k (main.go:183-189):
//   left = s_tab[k] + ds; for left < right { ...; left += ds }
static void build_nfine(const std::vector<double>& s_tab, double ds_fine, std::vector<unsigned char>& nfine) {
    nfine.assign(s_tab.size(), 0);
    for (size_t k = 0; k + 1 < s_tab.size(); ++k) {
        double left = s_tab[k], right = s_tab[k + 1];
        int n = 0;
        left += ds_fine;
        while (left < right && n < 255) {
            ++n;
            left += ds_fine;
        }
        nfine[k] = (unsigned char)n;
    }
}

// Scene shapes with a dedicated fp32 kernel (render_fast.cu): 1 = a primitive or one collection of
// primitives, 2 = tessellation of one collection of primitives, 0 = generic interpreter.
static int detect_shape(const XRayScene* sc, int& i_coll, int& i_tess) {
    const Header* h = (const Header*)sc->blob.data();
    const Instr* I = (const Instr*)(sc->blob.data() + h->instr_off);
    const int n = (int)h->n_instr;
    auto is_prim = [](uint32_t op) { return op >= OP_SPHERE && op <= OP_GYROID; };
    auto coll_of_prims = [&](int c) -> bool {
        if (I[c].op != OP_COLL_BEGIN) return false;
        const int e = (int)I[c].skip_to;
        if (e <= c || e >= n || I[e].op != OP_COLL_END) return false;
        for (int k = c + 1; k < e; ++k)
            if (!is_prim(I[k].op)) return false;
        return true;
    };
    i_coll = i_tess = 0;
    if (n >= 2 && is_prim(I[0].op) && I[0].n == 1 && I[1].op == OP_END) return 1;
    if (coll_of_prims(0) && I[I[0].skip_to + 1].op == OP_END) return 1;
    if (n >= 5 && I[0].op == OP_TESS_BEGIN && coll_of_prims(1)) {
        const int e = (int)I[1].skip_to;
        if (I[e + 1].op == OP_TESS_END && I[e + 2].op == OP_END) {
            i_tess = 0;
            i_coll = 1;
            return 2;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// One device's share of a render
// ---------------------------------------------------------------------------------------
struct Job {
    XRayScene* scene;
    const XRayCameraParams64* cams;
    std::vector<int> views;  // indices into cams / output, all with the same R
    int res;
    XRayRenderOpts opts;
    double ds;
    void* out;          // host or device base pointer of the full output
    bool out_on_device;
    int dev;
    cudaStream_t user_stream;
    bool use_user_stream;
    const void* borrowed_vox;
    bool fast_volume;  // dedicated voxel kernel (single voxel_grid root, fp32, simple, no warp)
    int peer_dev;      // device already holding the voxel data (multi-GPU replication), or -1
    unsigned long long stats[XRAY_NUM_STATS];
    int rc;
    std::string err;
};

// Pixel footprint of a warp in the voxel kernel (render_volume.cu launch_render_volume_tex).  Texture layers are
// z slices, so lanes should sit side by side along the image axis that moves least in z: camera x (matrix column 0)
// for the usual CT geometry (polar = 90 deg), camera y when the camera is rolled; the square-ish 4 x 8 otherwise.
static int volume_warp_shape(const Job& J, int v0, int n) {
    if (const char* e = getenv("XRAY_VOLUME_TILE")) return atoi(e);
    double zi = 0.0, zj = 0.0;
    for (int v = v0; v < v0 + n; ++v) {
        const XRayCameraParams64& c = J.cams[J.views[v]];
        zi += std::fabs(c.view[2 * 4 + 0]);
        zj += std::fabs(c.view[2 * 4 + 1]);
    }
    if (zi <= 0.5 * zj) return 1;  // 32 x 1 (i x j)
    if (zj <= 0.5 * zi) return 2;  // 1 x 32
    return 0;
}

// Persistent per-device scratch: grow-only buffers, cached sample-lattice tables, one stream.
// Keeps the steady-state render path free of cudaMalloc/cudaFree (both synchronise the device).
struct DevCtx {
    std::mutex mu;
    int dev = -1;
    cudaStream_t stream = nullptr;
    CamDev* d_cams = nullptr;
    size_t cams_cap = 0;
    double* d_s = nullptr;
    float* d_t = nullptr;
    unsigned char* d_nfine = nullptr;
    size_t tab_cap = 0;
    int tab_integ = -1;
    double tab_ds = 0, tab_smin = 0, tab_smax = 0;
    int tab_steps = 0;
    unsigned long long* d_stats = nullptr;
    unsigned char* d_img[2] = {nullptr, nullptr};
    size_t img_cap = 0;
    unsigned char* h_pin[2] = {nullptr, nullptr};
    size_t pin_cap = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};   // batch b has reached host memory (recorded on copy_stream)
    cudaEvent_t evk[2] = {nullptr, nullptr};  // batch b's kernel is done (recorded on the compute stream)
    cudaStream_t copy_stream = nullptr;       // D2H of batch b overlaps the kernel of batch b + 1
    CamDev* h_cams = nullptr;  // pinned staging for the camera upload
    size_t h_cams_cap = 0;
    // voxel volume as a layered 2D array (layers = z, width = y, height = x) for texture gather
    cudaArray_t vol_arr = nullptr;
    cudaTextureObject_t vol_tex = 0;
    cudaSurfaceObject_t vol_surf = 0;  // same array, for the fused copy + empty-space pass (0 if unavailable)
    int vol_dims[3] = {0, 0, 0};
    // empty-space map of the volume (render_volume.cu build_volume_occupancy), two ping-pong buffers
    unsigned char* vol_occ[2] = {nullptr, nullptr};
    size_t vol_occ_cap = 0;
    // tiles the span renderer hands to the marching kernels: compacted ids (one slot per CTA of a launch) + their count
    unsigned int* d_tile_list[2] = {nullptr, nullptr};  // one per output buffer of the double-buffered host path
    size_t tile_list_cap = 0;
    // The scratch above (cameras, lattice tables, stats, bins ...) is shared by every call on this device, and the device-output
    // entry points return with their kernels still in flight on the caller's stream.  ev_busy marks the end of the last call's
    // work; the next call -- whatever stream it uses -- waits for it before it touches the scratch.
    cudaEvent_t ev_busy = nullptr;
    bool busy = false;
    cudaStream_t aux_stream = nullptr;        // hand-over pass of batch b runs here, beside the span kernel of batch b + 1
    cudaEvent_t ev_span[2] = {nullptr, nullptr};
    // screen-space bins of the span renderer: per (view, tile) a count and bin_cap instance codes
    unsigned int* d_bins = nullptr;
    size_t bins_cap_words = 0;
};
static std::mutex g_ctx_mu;
static std::map<int, DevCtx*> g_ctx;

static DevCtx* ctx_for(int dev) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(dev);
    if (it != g_ctx.end()) return it->second;
    DevCtx* c = new DevCtx();
    c->dev = dev;
    g_ctx[dev] = c;
    return c;
}

static void ctx_release(DevCtx* c) {
    cudaSetDevice(c->dev);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->d_cams) cudaFree(c->d_cams);
    if (c->d_s) cudaFree(c->d_s);
    if (c->d_t) cudaFree(c->d_t);
    if (c->d_nfine) cudaFree(c->d_nfine);
    if (c->d_stats) cudaFree(c->d_stats);
    if (c->h_cams) cudaFreeHost(c->h_cams);
    if (c->vol_tex) cudaDestroyTextureObject(c->vol_tex);
    if (c->vol_surf) cudaDestroySurfaceObject(c->vol_surf);
    if (c->vol_arr) cudaFreeArray(c->vol_arr);
    c->vol_tex = 0;
    c->vol_surf = 0;
    c->vol_arr = nullptr;
    c->vol_dims[0] = c->vol_dims[1] = c->vol_dims[2] = 0;
    for (int b = 0; b < 2; ++b) {
        if (c->vol_occ[b]) cudaFree(c->vol_occ[b]);
        c->vol_occ[b] = nullptr;
    }
    c->vol_occ_cap = 0;
    for (int b = 0; b < 2; ++b) {
        if (c->d_tile_list[b]) cudaFree(c->d_tile_list[b]);
        c->d_tile_list[b] = nullptr;
        if (c->ev_span[b]) cudaEventDestroy(c->ev_span[b]);
        c->ev_span[b] = nullptr;
    }
    c->tile_list_cap = 0;
    if (c->ev_busy) cudaEventDestroy(c->ev_busy);
    c->ev_busy = nullptr;
    c->busy = false;
    if (c->aux_stream) {
        cudaStreamSynchronize(c->aux_stream);
        cudaStreamDestroy(c->aux_stream);
    }
    c->aux_stream = nullptr;
    if (c->d_bins) cudaFree(c->d_bins);
    c->d_bins = nullptr;
    c->bins_cap_words = 0;
    for (int b = 0; b < 2; ++b) {
        if (c->d_img[b]) cudaFree(c->d_img[b]);
        if (c->h_pin[b]) cudaFreeHost(c->h_pin[b]);
        if (c->ev[b]) cudaEventDestroy(c->ev[b]);
        if (c->evk[b]) cudaEventDestroy(c->evk[b]);
    }
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
    }
    c->copy_stream = nullptr;
    if (c->stream) cudaStreamDestroy(c->stream);
    c->stream = nullptr;
    c->d_cams = nullptr; c->cams_cap = 0;
    c->d_s = nullptr; c->d_t = nullptr; c->d_nfine = nullptr; c->tab_cap = 0; c->tab_integ = -1;
    c->d_stats = nullptr;
    c->h_cams = nullptr; c->h_cams_cap = 0;
    for (int b = 0; b < 2; ++b) { c->d_img[b] = nullptr; c->h_pin[b] = nullptr; c->ev[b] = nullptr; c->evk[b] = nullptr; }
    c->img_cap = 0; c->pin_cap = 0;
}

static int run_job(Job& J) {
    cudaError_t e;
    int rc = 0;
    CU(3, cudaSetDevice(J.dev));
    DevCtx* C = ctx_for(J.dev);
    std::lock_guard<std::mutex> ctx_lock(C->mu);
#define CUJ(code, call)                               \
    do {                                              \
        e = (call);                                   \
        if (e != cudaSuccess) {                       \
            rc = fail_cuda(code, #call, e);           \
            J.err = g_last_error;                     \
            return rc;                                \
        }                                             \
    } while (0)
    if (!C->stream) CUJ(3, cudaStreamCreateWithFlags(&C->stream, cudaStreamNonBlocking));
    cudaStream_t stream = J.use_user_stream ? J.user_stream : C->stream;
    if (C->busy) {  // an earlier call (possibly on another stream) may still be reading the per-device scratch
        CUJ(7, cudaEventSynchronize(C->ev_busy));
        C->busy = false;
    }
    const int nv = (int)J.views.size();
    if (nv == 0) return 0;
    const int res = J.res;
    const size_t esz = J.opts.out_dtype == XRAY_OUT_F64 ? 8 : 4;
    const size_t img_bytes = (size_t)res * res * esz;
    RenderParams P = {};

    DevScene* ds = nullptr;
    rc = ensure_dev_scene(J.scene, J.dev, stream, J.borrowed_vox, &ds, J.peer_dev);
    if (rc) {
        J.err = g_last_error;
        return rc;
    }
    const Header* h = (const Header*)J.scene->blob.data();
    fill_scene_dev(J.scene, *ds, P.scene);

    const double R = J.cams[J.views[0]].R;
    P.smin = R - kCubeHalfDiagonal;  // main.go:465
    P.smax = R + kCubeHalfDiagonal;
    P.s_center = R;
    P.ds = J.ds;
    P.ds_fine = J.ds / 10.0;  // main.go:178
    int i_coll = 0, i_tess = 0;
    int shape = detect_shape(J.scene, i_coll, i_tess);
    if (getenv("XRAY_GENERIC_KERNEL")) shape = 0;

    // sample lattice: cached on the device while (integrator, ds, window) stay the same
    if (C->tab_integ != J.opts.integration || C->tab_ds != J.ds || C->tab_smin != P.smin || C->tab_smax != P.smax) {
        std::vector<double> s_tab;
        std::vector<float> t_tab;
        std::vector<unsigned char> nfine;
        build_lattice(J.opts.integration, J.ds, P.smin, P.smax, s_tab);
        t_tab.resize(s_tab.size());
        for (size_t k = 0; k < s_tab.size(); ++k) t_tab[k] = (float)(s_tab[k] - P.s_center);
        if (J.opts.integration == XRAY_INTEGRATE_HIERARCHICAL) build_nfine(s_tab, P.ds_fine, nfine);
        else nfine.assign(s_tab.size(), 0);
        const size_t need = std::max<size_t>(16, s_tab.size());
        CUJ(7, cudaStreamSynchronize(stream));  // earlier launches may still read the old tables
        if (need > C->tab_cap) {
            if (C->d_s) cudaFree(C->d_s);
            if (C->d_t) cudaFree(C->d_t);
            if (C->d_nfine) cudaFree(C->d_nfine);
            C->d_s = nullptr; C->d_t = nullptr; C->d_nfine = nullptr; C->tab_cap = 0;
            CUJ(3, cudaMalloc(&C->d_s, sizeof(double) * need));
            CUJ(3, cudaMalloc(&C->d_t, sizeof(float) * need));
            CUJ(3, cudaMalloc(&C->d_nfine, need));
            C->tab_cap = need;
        }
        if (!s_tab.empty()) {
            CUJ(4, cudaMemcpy(C->d_s, s_tab.data(), sizeof(double) * s_tab.size(), cudaMemcpyHostToDevice));
            CUJ(4, cudaMemcpy(C->d_t, t_tab.data(), sizeof(float) * t_tab.size(), cudaMemcpyHostToDevice));
            CUJ(4, cudaMemcpy(C->d_nfine, nfine.data(), nfine.size(), cudaMemcpyHostToDevice));
        }
        C->tab_integ = J.opts.integration;
        C->tab_ds = J.ds;
        C->tab_smin = P.smin;
        C->tab_smax = P.smax;
        C->tab_steps = J.opts.integration == XRAY_INTEGRATE_SIMPLE ? (int)s_tab.size() : (int)s_tab.size() - 1;
    }
    P.n_steps = C->tab_steps;
    P.s_tab = C->d_s;
    P.t_tab = C->d_t;

    // cameras: pinned staging -> device, ordered on the stream
    if ((size_t)nv > C->cams_cap) {
        CUJ(7, cudaStreamSynchronize(stream));
        if (C->d_cams) cudaFree(C->d_cams);
        if (C->h_cams) cudaFreeHost(C->h_cams);
        C->d_cams = nullptr; C->h_cams = nullptr; C->cams_cap = 0;
        const size_t cap = std::max<size_t>(64, (size_t)nv * 2);
        CUJ(3, cudaMalloc(&C->d_cams, sizeof(CamDev) * cap));
        CUJ(3, cudaMallocHost(&C->h_cams, sizeof(CamDev) * cap));
        C->cams_cap = cap;
    } else {
        CUJ(7, cudaStreamSynchronize(stream));  // the staging buffer may still be in flight from the previous call
    }
    for (int v = 0; v < nv; ++v) {
        const XRayCameraParams64& c = J.cams[J.views[v]];
        memcpy(C->h_cams[v].eye, c.eye, sizeof(c.eye));
        memcpy(C->h_cams[v].view, c.view, sizeof(c.view));
        C->h_cams[v].f = focal_from_fov(c.fov_y);
    }
    CUJ(4, cudaMemcpyAsync(C->d_cams, C->h_cams, sizeof(CamDev) * nv, cudaMemcpyHostToDevice, stream));
    if (J.opts.stats) {
        if (!C->d_stats) CUJ(3, cudaMalloc(&C->d_stats, sizeof(unsigned long long) * XRAY_NUM_STATS));
        CUJ(4, cudaMemsetAsync(C->d_stats, 0, sizeof(unsigned long long) * XRAY_NUM_STATS, stream));
    }
    P.res = res;
    P.tiles_i = (res + kTileI - 1) / kTileI;
    P.tiles_j = (res + kTileJ - 1) / kTileJ;
    P.flat_field = J.opts.flat_field;
    P.dm = J.opts.density_multiplier;
    P.dm_f = (float)P.dm;
    P.ds_f = (float)P.ds;
    P.ds_fine_f = (float)P.ds_fine;
    for (int a = 0; a < 3; ++a) {
        P.aabb_lo[a] = h->aabb_lo[a];
        P.aabb_hi[a] = h->aabb_hi[a];
    }
    P.out_f64 = J.opts.out_dtype == XRAY_OUT_F64;
    P.stats = J.opts.stats ? C->d_stats : nullptr;
    P.skip_m2s = getenv("XRAY_NO_SKIP") ? 0.0f : (float)(0.999 / (J.ds * std::fmax(1.0, h->warp_lipschitz)));
    P.dbg_cause = 0;
#ifdef XRAY_DEV_KNOBS
    if (const char* e = getenv("XRAY_DEBUG_FB_CAUSE")) P.dbg_cause = atoi(e);  // which fallbacks stats[2] counts (development builds)
#endif
    {
        size_t prog = ((size_t)h->n_instr * 2 + h->f32_count) * 16;
        size_t stack = (size_t)h->save_depth * kBlockThreads * (5 * sizeof(double) + 5 * sizeof(float) + 8 * sizeof(unsigned int));
        size_t queue = (size_t)kQueueCap * kBlockThreads * sizeof(int);
        P.prog_in_smem = (prog + stack + queue) <= 160 * 1024 ? 1 : 0;
        P.smem_prog_bytes = P.prog_in_smem ? (unsigned int)prog : 0u;
    }

    // Dedicated voxel kernel: (re)fill the layered array from the linear device volume.  Falls back to the
    // __ldg kernel when the array cannot be had (dimension limits: 2048 layers, 32768 x 32768 texels).
    bool use_tex = false;
    const unsigned char* vol_occ = nullptr;
    if (J.fast_volume && J.opts.precision == XRAY_PRECISION_FP32 && !getenv("XRAY_VOLUME_LDG")) {
        const int vx = h->voxel_dims[0][0], vy = h->voxel_dims[0][1], vz = h->voxel_dims[0][2];
        if (vz <= 2048 && vx <= 32768 && vy <= 32768) {
            if (C->vol_dims[0] != vx || C->vol_dims[1] != vy || C->vol_dims[2] != vz) {
                CUJ(7, cudaStreamSynchronize(stream));
                if (C->vol_tex) cudaDestroyTextureObject(C->vol_tex);
                if (C->vol_surf) cudaDestroySurfaceObject(C->vol_surf);
                if (C->vol_arr) cudaFreeArray(C->vol_arr);
                C->vol_tex = 0;
                C->vol_surf = 0;
                C->vol_arr = nullptr;
                C->vol_dims[0] = C->vol_dims[1] = C->vol_dims[2] = 0;
                cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
                cudaError_t ea = cudaMalloc3DArray(&C->vol_arr, &cd, make_cudaExtent((size_t)vy, (size_t)vx, (size_t)vz),
                                                   cudaArrayLayered | cudaArraySurfaceLoadStore);
                bool with_surface = ea == cudaSuccess;
                if (ea != cudaSuccess) {
                    cudaGetLastError();
                    C->vol_arr = nullptr;
                    ea = cudaMalloc3DArray(&C->vol_arr, &cd, make_cudaExtent((size_t)vy, (size_t)vx, (size_t)vz), cudaArrayLayered);
                }
                if (ea == cudaSuccess) {
                    cudaResourceDesc rd = {};
                    rd.resType = cudaResourceTypeArray;
                    rd.res.array.array = C->vol_arr;
                    cudaTextureDesc td = {};
                    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
                    td.filterMode = cudaFilterModePoint;
                    td.readMode = cudaReadModeElementType;
                    td.normalizedCoords = 0;
                    if (cudaCreateTextureObject(&C->vol_tex, &rd, &td, nullptr) == cudaSuccess) {
                        C->vol_dims[0] = vx; C->vol_dims[1] = vy; C->vol_dims[2] = vz;
                        if (with_surface && cudaCreateSurfaceObject(&C->vol_surf, &rd) != cudaSuccess) {
                            C->vol_surf = 0;
                            cudaGetLastError();
                        }
                    } else {
                        cudaFreeArray(C->vol_arr);
                        C->vol_arr = nullptr;
                        C->vol_tex = 0;
                    }
                }
                cudaGetLastError();
            }
            if (C->vol_arr) {
                use_tex = true;
                // empty-space map (and, with a surface, the array fill in the same pass over the volume)
                bool filled = false;
                if (!getenv("XRAY_VOLUME_NO_SKIP")) {
                    const size_t nb = volume_brick_count(vx, vy, vz);
                    if (nb > C->vol_occ_cap) {
                        CUJ(7, cudaStreamSynchronize(stream));
                        for (int b = 0; b < 2; ++b) {
                            if (C->vol_occ[b]) cudaFree(C->vol_occ[b]);
                            C->vol_occ[b] = nullptr;
                        }
                        C->vol_occ_cap = 0;
                        if (cudaMalloc(&C->vol_occ[0], nb) == cudaSuccess && cudaMalloc(&C->vol_occ[1], nb) == cudaSuccess) C->vol_occ_cap = nb;
                        cudaGetLastError();
                    }
                    if (C->vol_occ_cap >= nb) {
                        const unsigned long long surf = getenv("XRAY_VOLUME_MEMCPY3D") ? 0ull : (unsigned long long)C->vol_surf;
                        CUJ(5, build_volume_occupancy((const float*)ds->d_vox[0], vx, vy, vz, C->vol_occ[0], C->vol_occ[1], &vol_occ, surf, stream));
                        filled = surf != 0;
                    }
                }
                if (!filled) {
                    cudaMemcpy3DParms cp = {};
                    cp.srcPtr = make_cudaPitchedPtr(ds->d_vox[0], (size_t)vy * sizeof(float), (size_t)vy, (size_t)vx);
                    cp.dstArray = C->vol_arr;
                    cp.extent = make_cudaExtent((size_t)vy, (size_t)vx, (size_t)vz);
                    cp.kind = cudaMemcpyDeviceToDevice;
                    CUJ(4, cudaMemcpy3DAsync(&cp, stream));
                }
            }
        }
    }

    // big collections carry a cell-list grid: their fp32 pool stays in global memory (only instructions are staged)
    bool use_list = false;
    bool single_async = false;
    int single_prim = 0;  // the op when the (collection of the) scene is exactly one run of cylinders / gyroids / spheres / boxes
    if (shape != 0) {
        const Instr* I = (const Instr*)(J.scene->blob.data() + h->instr_off);
        use_list = I[i_coll].op == OP_COLL_BEGIN && (I[i_coll].flags & F_HAS_LIST);
        const int rb = I[i_coll].op == OP_COLL_BEGIN ? i_coll + 1 : i_coll;
        const int re = I[i_coll].op == OP_COLL_BEGIN ? (int)I[i_coll].skip_to : i_coll + 1;
        if (!use_list && re - rb == 1 && I[rb].child_bit == 0 && (I[rb].op == OP_CYL || I[rb].op == OP_GYROID || I[rb].op == OP_SPHERE || I[rb].op == OP_BOX) &&
            !getenv("XRAY_NO_SINGLE_PRIM"))
            single_prim = (int)I[rb].op;
        // exactly one primitive: nothing is shared between the lanes of a warp, so each lane marches on its own
        single_async = single_prim != 0 && I[rb].n == 1 && !getenv("XRAY_NO_ASYNC");
        if (use_list && J.opts.precision == XRAY_PRECISION_FP32) {
            P.prog_in_smem = 0;
            P.smem_prog_bytes = (unsigned int)((size_t)h->n_instr * sizeof(Instr));
        }
    }

    // host output: small batches, the un-overlapped tail is one batch of D2H; device output: fewer launches
    const size_t max_batch_bytes = J.out_on_device ? (size_t)1 << 30 : (size_t)64 << 20;
    int max_batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)nv, max_batch_bytes / img_bytes));
    unsigned long long n_launches = 0;
    // Span renderer (render_span.cu): scenes of convex primitives under no / an affine warp, both precisions.  It flags
    // the tiles it will not vouch for; the marching kernels below re-render exactly those.
    bool use_span = h->span_bytes > 0 && !J.fast_volume && !getenv("XRAY_NO_SPAN") && (size_t)P.n_steps * 16 < ((size_t)1 << 26) &&
                    span_kernel_smem_bytes(h->span_bytes) <= 200 * 1024;
    if (use_span) {
        const size_t need = (size_t)max_batch * P.tiles_i * P.tiles_j * 4;  // warp tiles
        if (need > C->tile_list_cap) {
            CUJ(7, cudaStreamSynchronize(stream));
            if (C->aux_stream) CUJ(7, cudaStreamSynchronize(C->aux_stream));
            for (int b = 0; b < 2; ++b) {
                if (C->d_tile_list[b]) cudaFree(C->d_tile_list[b]);
                C->d_tile_list[b] = nullptr;
            }
            C->tile_list_cap = 0;
            for (int b = 0; b < 2; ++b)
                CUJ(3, cudaMalloc(&C->d_tile_list[b], (2 * need + 2) * sizeof(unsigned int)));  // [0], [1] = counts; hand-over ids; settle-pass ids
            C->tile_list_cap = need;
        }
    }
    // Screen-space bins: un-warped scenes seen by ordinary pinhole cameras (orthonormal axes, the eye in the translation column)
    // that stand clear of the scene; anything else walks the candidate grid per ray.
    unsigned int bin_cap = 0, n_instances = 0;
    if (use_span && !getenv("XRAY_SPAN_NO_BINS")) {
        const SpanHeader* sh = (const SpanHeader*)(J.scene->blob.data() + h->span_off);
        const bool warped = (sh->flags & SPAN_HAS_WARP) != 0;
        bool ok = !warped || sh->f_wscale > 0.0f;  // (an affine warp keeps the rays straight and through one point: bins in world space)
        for (int v = 0; v < nv && ok; ++v) {
            const XRayCameraParams64& c = J.cams[J.views[v]];
            const double* m = c.view;
            for (int a = 0; a < 3 && ok; ++a) {
                for (int b = a; b < 3 && ok; ++b) {
                    const double d = m[0 * 4 + a] * m[0 * 4 + b] + m[1 * 4 + a] * m[1 * 4 + b] + m[2 * 4 + a] * m[2 * 4 + b];
                    ok = std::fabs(d - (a == b ? 1.0 : 0.0)) < 1e-6;
                }
                ok = ok && std::fabs(m[a * 4 + 3] - c.eye[a]) < 1e-9 && m[12 + a] == 0.0;
            }
            ok = ok && m[15] == 1.0;
            // depth of the nearest corner of the scene's box along the optical axis (camera looks down -z)
            double dmin = 1e300;
            for (int q = 0; q < 8 && ok; ++q) {
                double X[3];
                for (int a = 0; a < 3; ++a) X[a] = sh->outer[(q >> a & 1) ? 3 + a : a];
                if (warped) {  // the box is in warped space: its corner in the world
                    const double u[3] = {X[0] - sh->warp_b[0], X[1] - sh->warp_b[1], X[2] - sh->warp_b[2]};
                    for (int a = 0; a < 3; ++a) X[a] = sh->f_winv[a * 3 + 0] * u[0] + sh->f_winv[a * 3 + 1] * u[1] + sh->f_winv[a * 3 + 2] * u[2];
                }
                double d = 0.0;
                for (int a = 0; a < 3; ++a) d -= m[a * 4 + 2] * (X[a] - m[a * 4 + 3]);
                dmin = std::fmin(dmin, d);
            }
            ok = ok && dmin > 0.25;
        }
        if (ok) {
            n_instances = sh->n_periods * sh->n_children;
            // 128 entries per tile when that stays under 1 GiB (a tile on a row of lattice nodes, seen edge-on, lists ~100), else 64
            bin_cap = std::min<unsigned int>(128u, n_instances);
            size_t words = (size_t)max_batch * P.tiles_i * P.tiles_j * (1 + (size_t)bin_cap);
            if (words * sizeof(unsigned int) > ((size_t)1 << 30)) {
                bin_cap = std::min<unsigned int>(64u, n_instances);
                words = (size_t)max_batch * P.tiles_i * P.tiles_j * (1 + (size_t)bin_cap);
            }
            if (words * sizeof(unsigned int) > ((size_t)1 << 30)) bin_cap = 0;  // never more than 1 GiB of bins
            else if (words > C->bins_cap_words) {
                CUJ(7, cudaStreamSynchronize(stream));
                if (C->d_bins) cudaFree(C->d_bins);
                C->d_bins = nullptr;
                C->bins_cap_words = 0;
                if (cudaMalloc(&C->d_bins, words * sizeof(unsigned int)) == cudaSuccess) C->bins_cap_words = words;
                else {
                    cudaGetLastError();
                    bin_cap = 0;
                }
            }
        }
    }
    auto launch_march = [&](cudaStream_t stream) -> cudaError_t {  // (shadows the job's stream on purpose)
        if (J.fast_volume && J.opts.precision == XRAY_PRECISION_FP64)
            return launch_render_volume_f64(ds->d_vox[0], J.scene->vox[0].dtype == XRAY_VOXEL_F64 ? 1 : 0, h->voxel_dims[0][0], h->voxel_dims[0][1],
                                            h->voxel_dims[0][2], P, J.opts.integration == XRAY_INTEGRATE_HIERARCHICAL ? 1 : 0, C->d_nfine, stream);
        if (J.fast_volume && use_tex)
            return launch_render_volume_tex((unsigned long long)C->vol_tex, (const float*)ds->d_vox[0], h->voxel_dims[0][0],
                                            h->voxel_dims[0][1], h->voxel_dims[0][2], P, volume_warp_shape(J, (int)(P.cams - C->d_cams), P.n_views), vol_occ,
                                            J.opts.integration == XRAY_INTEGRATE_HIERARCHICAL ? 1 : 0, C->d_nfine, stream);
        if (J.fast_volume && J.opts.integration == XRAY_INTEGRATE_SIMPLE && J.scene->vox[0].dtype == XRAY_VOXEL_F32)
            return launch_render_volume_fast((const float*)ds->d_vox[0], h->voxel_dims[0][0], h->voxel_dims[0][1],
                                             h->voxel_dims[0][2], P, stream);
        if (J.opts.precision == XRAY_PRECISION_FP32 && shape != 0 && single_async && P.prog_in_smem && fast_kernel_smem_bytes(P) <= 200 * 1024)
            return launch_render_async(P, shape, J.opts.integration, P.stats != nullptr, single_prim, C->d_nfine, i_coll, i_tess, stream);
        if (J.opts.precision == XRAY_PRECISION_FP32 && shape != 0 && (P.prog_in_smem || use_list) && fast_kernel_smem_bytes(P) <= 200 * 1024)
            return launch_render_fast(P, shape, J.opts.integration, P.stats != nullptr, use_list, single_prim, C->d_nfine, i_coll, i_tess,
                                      stream);
        return launch_render_scene(P, J.opts.precision, J.opts.integration, stream);
    };
    // parity: which tile list / event to use; march_stream: where the hand-over pass goes (the job's stream, or the auxiliary
    // one, so that it overlaps the next batch's span kernel -- a handful of hard warp tiles take ~1 ms on their own)
    // (n_launches counts the kernels of this library that are launched: stats[5], bench.py's gpu_launches)
    auto launch = [&](int v0, int n, void* d_dst, int parity, cudaStream_t march_stream) -> cudaError_t {
        P.cams = C->d_cams + v0;
        P.n_views = n;
        P.out = d_dst;
        P.out_vec4 = ((uintptr_t)d_dst % 16 == 0 && img_bytes % 16 == 0) ? 1 : 0;  // 128-bit stores need a 16-byte aligned image base
        P.tile_list = nullptr;
        P.tile_count = nullptr;
        // the grid is one CTA per (view, tile); keep it below 2^31
        if ((size_t)n * P.tiles_i * P.tiles_j > 0x7fffffffull) return cudaErrorInvalidValue;
        if (P.stats && J.opts.integration == XRAY_INTEGRATE_HIERARCHICAL) {  // main.go:162-169 clipping warnings, as stats[7] bits 17 / 18
            const int smem_flag = P.prog_in_smem;
            const unsigned int smem_bytes = P.smem_prog_bytes;
            if (use_list) {  // (list scenes stage only the instructions for the fast kernels; the probe reads everything from global)
                P.prog_in_smem = 0;
                P.smem_prog_bytes = 0;
            }
            cudaError_t ec = launch_clip_probe(P, stream);
            ++n_launches;
            P.prog_in_smem = smem_flag;
            P.smem_prog_bytes = smem_bytes;
            if (ec != cudaSuccess) return ec;
        }
        if (!use_span) {
            ++n_launches;
            return launch_march(stream);
        }
        unsigned int* tl = C->d_tile_list[parity];
        cudaError_t es = launch_render_span(P, J.opts.integration, P.stats != nullptr, C->d_nfine, ds->d_blob + h->span_off, h->span_bytes,
                                            tl + 2, tl, bin_cap ? C->d_bins : nullptr, bin_cap, n_instances, stream);
        if (es != cudaSuccess) return es;
        n_launches += bin_cap ? 3 : 2;  // (binning,) fast pass, settle pass
#ifdef XRAY_DEV_KNOBS
        if (P.dbg_cause == 77) return es;  // leave the interval renderer's hand-over codes in the image
#endif
        ++n_launches;  // the marching kernels' pass over the hand-over list
        if (march_stream != stream) {
            es = cudaEventRecord(C->ev_span[parity], stream);
            if (es == cudaSuccess) es = cudaStreamWaitEvent(march_stream, C->ev_span[parity], 0);
            if (es != cudaSuccess) return es;
        }
        P.tile_list = tl + 2;
        P.tile_count = tl;
        es = launch_march(march_stream);
        P.tile_list = nullptr;
        P.tile_count = nullptr;
        return es;
    };

    if (J.out_on_device) {
        // views of this job that are adjacent in the output share one launch
        int v = 0;
        while (v < nv) {
            int n = 1;
            while (v + n < nv && n < max_batch && J.views[v + n] == J.views[v + n - 1] + 1) ++n;
            CUJ(5, launch(v, n, (unsigned char*)J.out + (size_t)J.views[v] * img_bytes, 0, stream));
            v += n;
        }
    } else {
        cudaPointerAttributes attr;
        bool out_pinned = cudaPointerGetAttributes(&attr, J.out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const size_t need = (size_t)max_batch * img_bytes;
        if (need > C->img_cap) {
            CUJ(7, cudaStreamSynchronize(stream));
            for (int b = 0; b < 2; ++b) {
                if (C->d_img[b]) cudaFree(C->d_img[b]);
                C->d_img[b] = nullptr;
            }
            C->img_cap = 0;
            for (int b = 0; b < 2; ++b) CUJ(3, cudaMalloc(&C->d_img[b], need));
            C->img_cap = need;
        }
        if (!out_pinned && need > C->pin_cap) {
            CUJ(7, cudaStreamSynchronize(stream));
            for (int b = 0; b < 2; ++b) {
                if (C->h_pin[b]) cudaFreeHost(C->h_pin[b]);
                C->h_pin[b] = nullptr;
            }
            C->pin_cap = 0;
            for (int b = 0; b < 2; ++b) CUJ(3, cudaMallocHost(&C->h_pin[b], need));
            C->pin_cap = need;
        }
        for (int b = 0; b < 2; ++b) {
            if (!C->ev[b]) CUJ(3, cudaEventCreateWithFlags(&C->ev[b], cudaEventDisableTiming));
            if (!C->evk[b]) CUJ(3, cudaEventCreateWithFlags(&C->evk[b], cudaEventDisableTiming));
        }
        if (!C->copy_stream) CUJ(3, cudaStreamCreateWithFlags(&C->copy_stream, cudaStreamNonBlocking));
        cudaStream_t cstream = C->copy_stream;
        cudaStream_t mstream = stream;
        if (use_span) {
            if (!C->aux_stream) CUJ(3, cudaStreamCreateWithFlags(&C->aux_stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b)
                if (!C->ev_span[b]) CUJ(3, cudaEventCreateWithFlags(&C->ev_span[b], cudaEventDisableTiming));
            mstream = C->aux_stream;
        }
        int pend_v0[2] = {0, 0}, pend_n[2] = {0, 0};
        auto drain = [&](int b) -> cudaError_t {  // copy batch b from pinned staging into the caller's buffer
            if (pend_n[b] == 0) return cudaSuccess;
            cudaError_t ee = cudaEventSynchronize(C->ev[b]);
            if (ee != cudaSuccess) return ee;
            if (!out_pinned) {
                int k = 0;  // contiguous runs of views move in one (multi-threaded) copy each
                while (k < pend_n[b]) {
                    int m = 1;
                    while (k + m < pend_n[b] && J.views[pend_v0[b] + k + m] == J.views[pend_v0[b] + k + m - 1] + 1) ++m;
                    parallel_memcpy((unsigned char*)J.out + (size_t)J.views[pend_v0[b] + k] * img_bytes, C->h_pin[b] + (size_t)k * img_bytes,
                                    (size_t)m * img_bytes);
                    k += m;
                }
            }
            pend_n[b] = 0;
            return cudaSuccess;
        };
        int v = 0, b = 0;
        while (v < nv) {
            int n = std::min(max_batch, nv - v);
            CUJ(7, drain(b));  // buffer b is free again once its previous batch reached the caller
            CUJ(5, launch(v, n, C->d_img[b], b, mstream));
            CUJ(7, cudaEventRecord(C->evk[b], mstream));  // the batch is complete once its hand-over pass is
            CUJ(7, cudaStreamWaitEvent(cstream, C->evk[b], 0));
            if (out_pinned) {
                // contiguous runs of views go out in one copy each
                int k = 0;
                while (k < n) {
                    int m = 1;
                    while (k + m < n && J.views[v + k + m] == J.views[v + k + m - 1] + 1) ++m;
                    CUJ(8, cudaMemcpyAsync((unsigned char*)J.out + (size_t)J.views[v + k] * img_bytes,
                                           C->d_img[b] + (size_t)k * img_bytes, (size_t)m * img_bytes, cudaMemcpyDeviceToHost, cstream));
                    k += m;
                }
            } else {
                CUJ(8, cudaMemcpyAsync(C->h_pin[b], C->d_img[b], (size_t)n * img_bytes, cudaMemcpyDeviceToHost, cstream));
            }
            CUJ(7, cudaEventRecord(C->ev[b], cstream));
            pend_v0[b] = v;
            pend_n[b] = n;
            v += n;
            b ^= 1;
        }
        CUJ(7, drain(b));
        CUJ(7, drain(b ^ 1));
        if (mstream != stream) CUJ(7, cudaStreamSynchronize(mstream));  // (already idle: both drains waited on its batches)
    }
    if (J.opts.stats) {
        CUJ(8, cudaMemcpyAsync(J.stats, C->d_stats, sizeof(unsigned long long) * XRAY_NUM_STATS, cudaMemcpyDeviceToHost, stream));
        CUJ(7, cudaStreamSynchronize(stream));
        J.stats[5] = n_launches;
    } else if (!J.out_on_device) {
        CUJ(7, cudaStreamSynchronize(stream));
    } else {
        if (!C->ev_busy) CUJ(3, cudaEventCreateWithFlags(&C->ev_busy, cudaEventDisableTiming));
        CUJ(7, cudaEventRecord(C->ev_busy, stream));
        C->busy = true;
    }
    return 0;
#undef CUJ
}

static int resolve_ds(const XRayScene* sc, double ds_in, double& ds) {
    const Header* h = (const Header*)sc->blob.data();
    ds = ds_in > 0 ? ds_in : h->min_feature_size / 5.0;  // main.go:350-353
    if (!(ds > 0) || !std::isfinite(ds)) return fail(2, "step size ds is not positive/finite (scene has no finite MinFeatureSize?)");
    if (3.48 / ds > 5.0e7) return fail(2, "step size ds too small for the sample-lattice table");
    return 0;
}

// Largest |coordinate| of pc = o + d*R over the corner rays of every camera (the extremes of a pinhole fan),
// rays built like eval.cuh make_ray (main.go:457-465).
static bool fp32_position_bound_ok(const Header* h, const XRayCameraParams64* cams, int n, int res) {
    const double kU32 = 5.9604644775390625e-08, tmax = 1.75;
    double pcmax = 0.0;
    for (int v = 0; v < n; ++v) {
        const XRayCameraParams64& c = cams[v];
        const double f = 1.0 / std::tan(c.fov_y * 3.14159265358979323846 / 360.0);
        const double ext[2] = {-1.0, 1.0 - 2.0 / (double)res};
        for (int q = 0; q < 4; ++q) {
            const double px = ext[q & 1], py = ext[q >> 1], pz = -f;
            double t[4];
            for (int r = 0; r < 4; ++r) t[r] = c.view[r * 4 + 0] * px + c.view[r * 4 + 1] * py + c.view[r * 4 + 2] * pz + c.view[r * 4 + 3];
            double dv[3], l = 0.0;
            for (int a = 0; a < 3; ++a) {
                dv[a] = t[a] / t[3] - c.eye[a];
                l += dv[a] * dv[a];
            }
            l = std::sqrt(l);
            if (!(l > 0.0)) return false;
            for (int a = 0; a < 3; ++a) pcmax = std::fmax(pcmax, std::fabs(c.eye[a] + dv[a] / l * c.R));
        }
    }
    if (!std::isfinite(pcmax)) return false;
    double amax = 0.0;
    for (int a = 0; a < 3; ++a) amax = std::fmax(amax, std::fmax(std::fabs(h->aabb_lo[a]), std::fabs(h->aabb_hi[a])));
    const double xmax = std::fmin(pcmax + tmax, std::isfinite(amax) ? amax + 0.01 : 1e300);  // only positions inside the scene matter
    const double base = 1.0e-6;  // kEpsPosBase (scene_compile.cpp): the budget for the un-warped position
    // roundings: fl(d)*t, fl(t) (coarse: the table entry; fine: t_tab[k] + j*ds_fine is rounded twice), fl(pc), the fma
    if (kU32 * (3.0 * tmax + pcmax + xmax) * 1.25 > base) return false;
    if (h->n_deform > 0 && pcmax + tmax > 4.0) return false;  // warp error terms assume |x| <= 4 (scene_compile.cpp deform_eps_pos)
    return true;
}

static bool volume_fast_path_ok(const XRayRenderOpts& o, int dtype);

static int render_common(XRayScene* scene, const XRayCameraParams64* cams, int n, int res, const XRayRenderOpts* opts_in,
                         void* out, bool out_on_device, const void* borrowed_vox, bool fast_volume) {
    if (!scene || !cams || !out) return fail(1, "null pointer argument");
    if (n <= 0 || res <= 0) return fail(2, "num_cameras and image_res must be positive");
    XRayRenderOpts opts;
    if (opts_in) {
        if (opts_in->struct_size != sizeof(XRayRenderOpts)) return fail(2, "XRayRenderOpts.struct_size mismatch");
        opts = *opts_in;
    } else {
        XRayRenderOptsInit(&opts);
    }
    if (opts.integration != XRAY_INTEGRATE_SIMPLE && opts.integration != XRAY_INTEGRATE_HIERARCHICAL)
        return fail(2, "unknown integration mode");
    if (opts.precision != XRAY_PRECISION_FP32 && opts.precision != XRAY_PRECISION_FP64) return fail(2, "unknown precision");
    if (opts.out_dtype != XRAY_OUT_F32 && opts.out_dtype != XRAY_OUT_F64) return fail(2, "unknown out_dtype");
    const Header* h = (const Header*)scene->blob.data();
    for (int s = 0; s < (int)h->n_voxel_slots; ++s)
        if (!scene->vox[s].data && !(s == 0 && borrowed_vox)) return fail(2, "voxel_grid slot has no data (XRaySceneSetVoxelData)");
    // Degenerate cameras (polar = 0 or 180 deg makes LookAtV singular -> NaN matrix, main.go:236-237) would turn
    // into NaN rays; the reference then writes NaN images.  Refuse them instead of marching garbage.
    for (int v = 0; v < n; ++v) {
        bool ok = std::isfinite(cams[v].R) && std::isfinite(cams[v].fov_y) && cams[v].fov_y > 0.0 && cams[v].fov_y < 180.0;
        for (int a = 0; a < 3 && ok; ++a) ok = std::isfinite(cams[v].eye[a]);
        for (int a = 0; a < 16 && ok; ++a) ok = std::isfinite(cams[v].view[a]);
        if (ok && !(std::fabs(cams[v].view[15]) > 1e-12)) ok = false;  // TransformCoordinate divides by w
        if (!ok) return fail(2, "camera " + std::to_string(v) + " is degenerate (non-finite matrix, fov outside (0,180) or w = 0)");
    }
    if (!std::isfinite(opts.flat_field) || !std::isfinite(opts.density_multiplier)) return fail(2, "flat_field / density_multiplier must be finite");
    double ds;
    if (int rc = resolve_ds(scene, opts.ds, ds)) return rc;

    // fp32 mode is only sound while every sample position x = fma(fl(d), fl(s - R), fl(o + d*R)) stays within
    // the position error the guard bands were derived for (Header.eps_pos = 1e-6 x warp Lipschitz):
    // |dx| <= u32 * (3*|t|max + |pc| + |x|).  A distant camera or a wide field of view puts pc = o + d*R far
    // from the origin; such calls are promoted to the fp64 kernels (slower, always exact) instead of risking a
    // misclassified sample.
    // A scene that is nothing but one fp32 voxel grid (a voxel_grid object file, no deformation) is what the volume entry
    // points build for themselves: it takes the dedicated voxel kernel from the scene entry points as well.
    if (!fast_volume && scene->root.type == N_VOXEL && scene->deforms.empty() && h->n_voxel_slots == 1 && !borrowed_vox && scene->vox[0].data)
        fast_volume = volume_fast_path_ok(opts, scene->vox[0].dtype);
    if (opts.precision == XRAY_PRECISION_FP32 && !fp32_position_bound_ok(h, cams, n, res)) {
        opts.precision = XRAY_PRECISION_FP64;  // (a voxel scene keeps its dedicated kernel: the fp64 one)
    }

    std::vector<int> devs;
    if (out_on_device || opts.num_devices <= 0) {
        int cur = 0;
        CU(3, cudaGetDevice(&cur));
        devs.push_back(cur);
    } else {
        int count = 0;
        CU(3, cudaGetDeviceCount(&count));
        if (opts.num_devices > XRAY_MAX_DEVICES) return fail(2, "too many devices");
        for (int i = 0; i < opts.num_devices; ++i) {
            if (opts.devices[i] < 0 || opts.devices[i] >= count) return fail(2, "device index out of range");
            devs.push_back(opts.devices[i]);
        }
    }
    const int G = (int)devs.size();

    // Multi-GPU with host voxel data: upload once to the first device, the others pull it over NVLink.
    int replicated_from = -1;
    if (G > 1 && h->n_voxel_slots > 0 && !borrowed_vox) {
        int cur = 0;
        cudaGetDevice(&cur);
        for (int a = 0; a < G; ++a)
            for (int b = 0; b < G; ++b)
                if (a != b) {
                    int can = 0;
                    if (cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) == cudaSuccess && can) {
                        cudaSetDevice(devs[a]);
                        cudaDeviceEnablePeerAccess(devs[b], 0);  // "already enabled" is fine
                        cudaGetLastError();
                    }
                }
        CU(3, cudaSetDevice(devs[0]));
        DevScene* ds0 = nullptr;
        if (int rc = ensure_dev_scene(scene, devs[0], 0, nullptr, &ds0)) {
            cudaSetDevice(cur);
            return rc;
        }
        // Binomial tree over NVLink / NVSwitch: in round r the 2^r devices that hold the volume each feed one that does not,
        // so G devices are served in ceil(log2 G) copy times instead of G - 1 pulls from the one root.
        std::vector<int> tree_rc(G, 0);
        std::vector<std::string> tree_err(G);
        for (int span = 1; span < G; span <<= 1) {
            std::vector<std::thread> th;
            for (int i = span; i < std::min(2 * span, G); ++i)
                th.emplace_back([&, i, span]() {
                    cudaSetDevice(devs[i]);
                    DevScene* dsi = nullptr;
                    tree_rc[i] = ensure_dev_scene(scene, devs[i], 0, nullptr, &dsi, devs[i - span]);
                    if (tree_rc[i]) tree_err[i] = g_last_error;
                });
            for (auto& t : th) t.join();
        }
        cudaSetDevice(cur);
        for (int i = 1; i < G; ++i)
            if (tree_rc[i]) return fail(tree_rc[i], tree_err[i]);
        replicated_from = devs[0];  // (every device already holds the current version: the jobs find nothing left to copy)
    }

    // group views by R (the sample lattice depends on R), then shard each group modulo G
    std::vector<double> Rs;
    for (int v = 0; v < n; ++v)
        if (std::find(Rs.begin(), Rs.end(), cams[v].R) == Rs.end()) Rs.push_back(cams[v].R);
    unsigned long long stats_total[XRAY_NUM_STATS] = {0};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    for (double R : Rs) {
        std::vector<Job> jobs(G);
        for (int g = 0; g < G; ++g) {
            Job& J = jobs[g];
            J.scene = scene;
            J.cams = cams;
            J.res = res;
            J.opts = opts;
            J.ds = ds;
            J.out = out;
            J.out_on_device = out_on_device;
            J.dev = devs[g];
            J.user_stream = (cudaStream_t)(uintptr_t)opts.stream;
            J.use_user_stream = out_on_device;
            J.borrowed_vox = borrowed_vox;
            J.fast_volume = fast_volume;
            J.peer_dev = replicated_from;
            memset(J.stats, 0, sizeof(J.stats));
            J.rc = 0;
        }
        int k = 0;
        for (int v = 0; v < n; ++v)
            if (cams[v].R == R) jobs[(k++) % G].views.push_back(v);  // v -> devices[v mod G], main.go:244
        if (G == 1) {
            jobs[0].rc = run_job(jobs[0]);
        } else {
            std::vector<std::thread> th;
            for (int g = 0; g < G; ++g) th.emplace_back([&jobs, g]() { jobs[g].rc = run_job(jobs[g]); });
            for (auto& t : th) t.join();
        }
        cudaSetDevice(cur_dev);
        for (int g = 0; g < G; ++g) {
            if (jobs[g].rc) return fail(jobs[g].rc, jobs[g].err);
            for (int s = 0; s < XRAY_NUM_STATS; ++s) {
                if (s == 7) stats_total[s] |= jobs[g].stats[s];  // a bit set, not a counter
                else stats_total[s] += jobs[g].stats[s];
            }
        }
    }
    if (opts.stats)
        for (int s = 0; s < XRAY_NUM_STATS; ++s) {
            if (s == 7) opts.stats[s] |= stats_total[s];
            else opts.stats[s] += stats_total[s];
        }
    return 0;
}

// ---------------------------------------------------------------------------------------
// Extended C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

const char* XRayLastError(void) { return g_last_error.c_str(); }

void XRayReleaseCaches(void) {
    {
        std::lock_guard<std::mutex> ls(g_stage_mu);
        stage_release_locked();
    }
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : g_ctx) {
        std::lock_guard<std::mutex> l2(kv.second->mu);
        ctx_release(kv.second);
    }
    cudaSetDevice(cur);
}

#define XR_STR2(x) #x
#define XR_STR(x) XR_STR2(x)
const char* XRayBuildInfo(void) {
    return "libcuda_render.so (xray_projection_render_b200), sm_100a only, nvcc " XR_STR(__CUDACC_VER_MAJOR__) "." XR_STR(
        __CUDACC_VER_MINOR__) "." XR_STR(__CUDACC_VER_BUILD__) ", api.cu compiled " __DATE__ " " __TIME__
#ifdef XRAY_DEV_KNOBS
           ", DEVELOPMENT build (XRAY_DEV_KNOBS)"
#endif
        ;
}

int XRayDeviceCount(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void XRayRenderOptsInit(XRayRenderOpts* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->struct_size = sizeof(XRayRenderOpts);
    o->integration = XRAY_INTEGRATE_HIERARCHICAL;  // reference default, main.go:39
    o->precision = XRAY_PRECISION_FP32;
    o->out_dtype = XRAY_OUT_F32;
    o->ds = -1.0;
    o->flat_field = 0.0;
    o->density_multiplier = 1.0;
}

// No C++ exception may cross the C boundary (a cgo / ctypes caller would be terminated): allocation failures become error 9.
#define XR_NOTHROW(expr)                                                        \
    try {                                                                       \
        return (expr);                                                          \
    } catch (const std::bad_alloc&) {                                           \
        return fail(9, "out of host memory");                                   \
    } catch (const std::exception& ex) {                                        \
        return fail(9, std::string("internal error: ") + ex.what());            \
    } catch (...) {                                                             \
        return fail(9, "internal error");                                       \
    }

static int compile_json_impl(const char* object_json, const char* deformation_json, XRayScene** out_scene) {
    std::string err;
    int rc = compile_scene_json(object_json, deformation_json, out_scene, err);
    if (rc) return fail(rc, err);
    return 0;
}
int XRaySceneCompileJSON(const char* object_json, const char* deformation_json, XRayScene** out_scene) {
    XR_NOTHROW(compile_json_impl(object_json, deformation_json, out_scene))
}

void XRaySceneFree(XRayScene* scene) {
    if (!scene) return;
    if (scene->device_cache) {
        SceneCache* c = (SceneCache*)scene->device_cache;
        for (auto& kv : c->per_dev) free_dev_scene(kv.second);
        delete c;
    }
    delete scene;
}

double XRaySceneMinFeatureSize(const XRayScene* scene) {
    return scene ? ((const Header*)scene->blob.data())->min_feature_size : 0.0;
}

const void* XRaySceneProgram(const XRayScene* scene, size_t* num_bytes) {
    if (!scene) return nullptr;
    if (num_bytes) *num_bytes = scene->blob.size();
    return scene->blob.data();
}

void XRaySceneBounds(const XRayScene* scene, double* lo, double* hi) {
    if (!scene) return;
    const Header* h = (const Header*)scene->blob.data();
    for (int a = 0; a < 3; ++a) {
        if (lo) lo[a] = h->aabb_lo[a];
        if (hi) hi[a] = h->aabb_hi[a];
    }
}

int XRaySceneNumVoxelSlots(const XRayScene* scene) { return scene ? scene->n_vox : 0; }

int XRaySceneVoxelDims(const XRayScene* scene, int slot, int* nx, int* ny, int* nz) {
    if (!scene || slot < 0 || slot >= scene->n_vox) return fail(2, "voxel slot out of range");
    if (nx) *nx = scene->vox[slot].nx;
    if (ny) *ny = scene->vox[slot].ny;
    if (nz) *nz = scene->vox[slot].nz;
    return 0;
}

int XRaySceneSetVoxelData(XRayScene* scene, int slot, const void* data, int nx, int ny, int nz, int dtype) {
    if (!scene || !data) return fail(1, "null pointer argument");
    if (slot < 0 || slot >= scene->n_vox) return fail(2, "voxel slot out of range");
    VoxelHost& v = scene->vox[slot];
    if (nx != v.nx || ny != v.ny || nz != v.nz) return fail(2, "voxel data dimensions do not match the scene's resolution");
    if (dtype != XRAY_VOXEL_F32 && dtype != XRAY_VOXEL_F64) return fail(2, "unknown voxel dtype");
    v.data = data;
    v.dtype = dtype;
    v.version++;
    return 0;
}

double XRaySceneDensityHost(const XRayScene* scene, double x, double y, double z, double density_multiplier) {
    return scene ? host_density(*scene, x, y, z, density_multiplier) : 0.0;
}

int XRayCameraFromAngles(double azimuthal_deg, double polar_deg, double R, double fov_deg, XRayCameraParams64* out) {
    if (!out) return fail(1, "null pointer argument");
    camera_from_angles(azimuthal_deg, polar_deg, R, out->eye, out->view);
    out->fov_y = fov_deg;
    out->R = R;
    return 0;
}

int XRayRenderSceneCUDA(XRayScene* scene, const XRayCameraParams64* cameras, int num_cameras, int image_res,
                        const XRayRenderOpts* opts, void* out_images) {
    XR_NOTHROW(render_common(scene, cameras, num_cameras, image_res, opts, out_images, false, nullptr, false))
}

int XRayRenderSceneDeviceCUDA(XRayScene* scene, const XRayCameraParams64* cameras, int num_cameras, int image_res,
                              const XRayRenderOpts* opts, void* d_out_images) {
    XR_NOTHROW(render_common(scene, cameras, num_cameras, image_res, opts, d_out_images, true, nullptr, false))
}

}  // extern "C"

// A voxel_grid-rooted scene for the volume entry points.
static int make_volume_scene(int nx, int ny, int nz, XRayScene** out) {
    char json[160];
    snprintf(json, sizeof(json), "{\"type\":\"voxel_grid\",\"resolution\":[%d,%d,%d]}", nx, ny, nz);
    std::string err;
    int rc = compile_scene_json(json, nullptr, out, err);
    if (rc) return fail(rc, err);
    return 0;
}

static bool volume_fast_path_ok(const XRayRenderOpts& o, int dtype) {
    if (getenv("XRAY_VOLUME_GENERIC")) return false;
    // fp32 mode: fp32 voxels, both integrators (hierarchical: the texture kernel only); fp64 mode: either voxel type
    return o.precision == XRAY_PRECISION_FP64 || dtype == XRAY_VOXEL_F32;
}

extern "C" {

int XRayRenderVolumeExCUDA(const void* volume, int volume_dtype, int nx, int ny, int nz, const XRayCameraParams64* cameras,
                           int num_cameras, int image_res, const XRayRenderOpts* opts, void* out_images) try {
    if (!volume || !cameras || !out_images) return fail(1, "null pointer argument");
    if (nx <= 0 || ny <= 0 || nz <= 0 || image_res <= 0 || num_cameras <= 0) return fail(2, "dimensions must be positive");
    XRayScene* sc = nullptr;
    if (int rc = make_volume_scene(nx, ny, nz, &sc)) return rc;
    int rc = XRaySceneSetVoxelData(sc, 0, volume, nx, ny, nz, volume_dtype);
    XRayRenderOpts o;
    if (opts) o = *opts;
    else XRayRenderOptsInit(&o);
    if (!rc) rc = render_common(sc, cameras, num_cameras, image_res, opts, out_images, false, nullptr,
                                volume_fast_path_ok(o, volume_dtype));
    std::string keep = g_last_error;
    XRaySceneFree(sc);
    g_last_error = keep;
    return rc;
} catch (const std::bad_alloc&) {
    return fail(9, "out of host memory");
} catch (...) {
    return fail(9, "internal error");
}

int XRayRenderVolumeDeviceCUDA(const float* d_volume, int nx, int ny, int nz, const XRayCameraParams64* cameras,
                               int num_cameras, int image_res, const XRayRenderOpts* opts, void* d_out_images) try {
    if (!d_volume || !cameras || !d_out_images) return fail(1, "null pointer argument");
    if (nx <= 0 || ny <= 0 || nz <= 0 || image_res <= 0 || num_cameras <= 0) return fail(2, "dimensions must be positive");
    XRayScene* sc = nullptr;
    if (int rc = make_volume_scene(nx, ny, nz, &sc)) return rc;
    XRayRenderOpts o;
    if (opts) o = *opts;
    else XRayRenderOptsInit(&o);
    int rc = render_common(sc, cameras, num_cameras, image_res, opts, d_out_images, true, d_volume,
                           volume_fast_path_ok(o, XRAY_VOXEL_F32));
    std::string keep = g_last_error;
    XRaySceneFree(sc);
    g_last_error = keep;
    return rc;
} catch (const std::bad_alloc&) {
    return fail(9, "out of host memory");
} catch (...) {
    return fail(9, "internal error");
}

int XRayRenderVolumeDeviceToHostCUDA(const float* d_volume, int nx, int ny, int nz, const XRayCameraParams64* cameras,
                                     int num_cameras, int image_res, const XRayRenderOpts* opts, void* out_images) try {
    if (!d_volume || !cameras || !out_images) return fail(1, "null pointer argument");
    if (nx <= 0 || ny <= 0 || nz <= 0 || image_res <= 0 || num_cameras <= 0) return fail(2, "dimensions must be positive");
    XRayScene* sc = nullptr;
    if (int rc = make_volume_scene(nx, ny, nz, &sc)) return rc;
    XRayRenderOpts o;
    if (opts) o = *opts;
    else XRayRenderOptsInit(&o);
    o.num_devices = 0;  // the volume lives on the current device
    int rc = render_common(sc, cameras, num_cameras, image_res, &o, out_images, false, d_volume, volume_fast_path_ok(o, XRAY_VOXEL_F32));
    std::string keep = g_last_error;
    XRaySceneFree(sc);
    g_last_error = keep;
    return rc;
} catch (const std::bad_alloc&) {
    return fail(9, "out of host memory");
} catch (...) {
    return fail(9, "internal error");
}

int XRayVoxelizeSceneCUDA(XRayScene* scene, int res, double density_multiplier, float* out_volume) try {
    if (!scene || !out_volume) return fail(1, "null pointer argument");
    if (res <= 0) return fail(2, "res must be positive");
    int dev = 0;
    CU(3, cudaGetDevice(&dev));
    DevScene* ds = nullptr;
    if (int rc = ensure_dev_scene(scene, dev, 0, nullptr, &ds)) return rc;
    const Header* h = (const Header*)scene->blob.data();
    RenderParams P = {};
    fill_scene_dev(scene, *ds, P.scene);
    P.dm = density_multiplier;
    size_t prog = ((size_t)h->n_instr * 2 + h->f32_count) * 16;
    P.prog_in_smem = prog <= 160 * 1024;
    P.smem_prog_bytes = P.prog_in_smem ? (unsigned int)prog : 0u;
    const size_t total = (size_t)res * res * res;
    float* d_out = nullptr;
    CU(3, cudaMalloc(&d_out, total * sizeof(float)));
    cudaError_t e = launch_voxelize_scene(P, res, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out_volume, d_out, total * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail_cuda(5, "voxelize", e);
    return 0;
} catch (const std::bad_alloc&) {
    return fail(9, "out of host memory");
} catch (...) {
    return fail(9, "internal error");
}

int XRayMeasureFp32Peak(double* tflops) {
    if (!tflops) return fail(1, "null pointer argument");
    cudaError_t e = measure_fp32_peak(tflops);
    if (e != cudaSuccess) return fail_cuda(5, "measure_fp32_peak", e);
    return 0;
}

// ---------------------------------------------------------------------------------------
// Legacy plugin surface (reference cuda_backend.h)
// ---------------------------------------------------------------------------------------

// Devices for the legacy symbols: XRAY_CUDA_DEVICES="all" | "0,1,2" ; default = device 0
// (the reference plugin never selects a device, i.e. implicit device 0).
static void legacy_devices(XRayRenderOpts& o) {
    o.num_devices = 1;
    o.devices[0] = 0;
    const char* e = getenv("XRAY_CUDA_DEVICES");
    if (!e || !*e) return;
    int count = XRayDeviceCount();
    if (!strcmp(e, "all")) {
        o.num_devices = std::max(1, std::min(count, XRAY_MAX_DEVICES));
        for (int i = 0; i < o.num_devices; ++i) o.devices[i] = i;
        return;
    }
    int n = 0;
    const char* p = e;
    while (*p && n < XRAY_MAX_DEVICES) {
        char* end = nullptr;
        long v = strtol(p, &end, 10);
        if (end == p) break;
        if (v >= 0 && v < count) o.devices[n++] = (int)v;
        p = (*end == ',') ? end + 1 : end;
    }
    if (n > 0) o.num_devices = n;
}

int RenderVolumeProjectionsCUDA(const float* volume, int nx, int ny, int nz, const XRayCameraParams* cameras, int num_cameras,
                                int image_res, float ds, float flat_field, float* out_images) try {
    if (!volume || !cameras || !out_images) return fail(1, "null pointer argument");  // cuda_backend.cu:95-97
    if (nx <= 0 || ny <= 0 || nz <= 0 || image_res <= 0 || num_cameras <= 0) return fail(2, "dimensions must be positive");
    if (!(ds > 0.0f)) return fail(2, "ds must be positive");
    // Everything crossing this boundary was narrowed to fp32 by the Go caller
    // (cuda_backend.go:124-134,329-330); widen it back and integrate those exact values.
    std::vector<XRayCameraParams64> cams(num_cameras);
    for (int c = 0; c < num_cameras; ++c) {
        for (int a = 0; a < 3; ++a) cams[c].eye[a] = (double)cameras[c].eye[a];
        for (int a = 0; a < 16; ++a) cams[c].view[a] = (double)cameras[c].view[a];
        cams[c].fov_y = (double)cameras[c].fov_y;
        cams[c].R = (double)cameras[c].R;
    }
    XRayRenderOpts o;
    XRayRenderOptsInit(&o);
    o.integration = XRAY_INTEGRATE_SIMPLE;  // cuda_path.go:99 always takes the fixed-step integrator
    o.precision = XRAY_PRECISION_FP32;
    o.out_dtype = XRAY_OUT_F32;
    o.ds = (double)ds;
    o.flat_field = (double)flat_field;
    o.density_multiplier = 1.0;
    legacy_devices(o);
    return XRayRenderVolumeExCUDA(volume, XRAY_VOXEL_F32, nx, ny, nz, cams.data(), num_cameras, image_res, &o, out_images);
} catch (const std::bad_alloc&) {
    return fail(9, "out of host memory");
} catch (...) {
    return fail(9, "internal error");
}

static int assemble_common(const CylinderParams* cylinders, int num_cylinders, int res, float density_multiplier, int grid_dim,
                           const int* cell_offsets, const int* cyl_indices, int num_cyl_indices, float* out_volume) {
    if (!cylinders || !out_volume) return fail(1, "null pointer argument");
    if (num_cylinders <= 0 || res <= 0) return fail(2, "num_cylinders and res must be positive");
    const size_t total = (size_t)res * res * res;
    // runs on the calling thread's current device, like the reference plugin (which never selects one)
    CylinderParams* d_cyl = nullptr;
    int *d_off = nullptr, *d_idx = nullptr;
    float* d_out = nullptr;
    cudaError_t e = cudaSuccess;
    int rc = 0;
    auto done = [&](int code, const char* what) {
        if (e != cudaSuccess && !rc) rc = fail_cuda(code, what, e);
        if (d_cyl) cudaFree(d_cyl);
        if (d_off) cudaFree(d_off);
        if (d_idx) cudaFree(d_idx);
        if (d_out) cudaFree(d_out);
        return rc;
    };
    if ((e = cudaMalloc(&d_cyl, sizeof(CylinderParams) * num_cylinders)) != cudaSuccess) return done(3, "cudaMalloc");
    if ((e = cudaMalloc(&d_out, total * sizeof(float))) != cudaSuccess) return done(3, "cudaMalloc");
    if ((e = cudaMemcpy(d_cyl, cylinders, sizeof(CylinderParams) * num_cylinders, cudaMemcpyHostToDevice)) != cudaSuccess)
        return done(4, "cudaMemcpy");
    if (grid_dim > 0) {
        const size_t ncell = (size_t)grid_dim * grid_dim * grid_dim;
        if ((e = cudaMalloc(&d_off, sizeof(int) * (ncell + 1))) != cudaSuccess) return done(3, "cudaMalloc");
        if ((e = cudaMalloc(&d_idx, sizeof(int) * std::max(1, num_cyl_indices))) != cudaSuccess) return done(3, "cudaMalloc");
        if ((e = cudaMemcpy(d_off, cell_offsets, sizeof(int) * (ncell + 1), cudaMemcpyHostToDevice)) != cudaSuccess)
            return done(4, "cudaMemcpy");
        if (num_cyl_indices > 0 &&
            (e = cudaMemcpy(d_idx, cyl_indices, sizeof(int) * num_cyl_indices, cudaMemcpyHostToDevice)) != cudaSuccess)
            return done(4, "cudaMemcpy");
    }
    e = launch_voxelize_cylinders(d_cyl, num_cylinders, res, density_multiplier, d_off, d_idx, grid_dim, d_out, 0);
    if (e != cudaSuccess) return done(5, "voxelize kernel");
    if ((e = cudaMemcpy(out_volume, d_out, total * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess)
        return done(6, "cudaMemcpy D2H");
    return done(0, "");
}

int AssembleVoxelGridCUDA(const CylinderParams* cylinders, int num_cylinders, int res, float density_multiplier,
                          float* out_volume) try {
    return assemble_common(cylinders, num_cylinders, res, density_multiplier, 0, nullptr, nullptr, 0, out_volume);
} catch (const std::bad_alloc&) {
    return fail(9, "out of host memory");
} catch (...) {
    return fail(9, "internal error");
}

int AssembleVoxelGridSpatialCUDA(const CylinderParams* cylinders, int num_cylinders, int res, float density_multiplier,
                                 int grid_dim, const int* cell_offsets, const int* cyl_indices, int num_cyl_indices,
                                 float* out_volume) try {
    if (!cell_offsets || (!cyl_indices && num_cyl_indices > 0)) return fail(1, "null pointer argument");
    if (grid_dim <= 0 || num_cyl_indices < 0) return fail(2, "grid_dim must be positive");
    // validate the caller's CSR before trusting it on the device
    const size_t ncell = (size_t)grid_dim * grid_dim * grid_dim;
    if (cell_offsets[0] != 0 || cell_offsets[ncell] != num_cyl_indices) return fail(2, "cell_offsets do not span cyl_indices");
    for (size_t c = 0; c < ncell; ++c)
        if (cell_offsets[c] > cell_offsets[c + 1]) return fail(2, "cell_offsets not monotonic");
    for (int k = 0; k < num_cyl_indices; ++k)
        if (cyl_indices[k] < 0 || cyl_indices[k] >= num_cylinders) return fail(2, "cyl_indices entry out of range");
    return assemble_common(cylinders, num_cylinders, res, density_multiplier, grid_dim, cell_offsets, cyl_indices,
                           num_cyl_indices, out_volume);
} catch (const std::bad_alloc&) {
    return fail(9, "out of host memory");
} catch (...) {
    return fail(9, "internal error");
}

}  // extern "C"
