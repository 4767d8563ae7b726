// render_scene.cu -- analytic-scene projection kernels for sm_100a.
//
// Replaces the reference CPU hot loop (main.go:123-199,456-470 + objects.go Density tree +
// deformations.go Apply) with one CUDA thread per ray:
//   * the object tree is a flattened, warp-uniform instruction stream (program.h) staged in
//     shared memory; all lanes of a warp execute the same opcode, greedy early-out and bounds
//     rejection are lane predicates plus warp votes;
//   * collections carry a child-mask grid: the warp ORs the masks of its lanes' cells and only
//     tests children in the union (exact: skipped children are provably 0 there);
//   * rays are clipped against the conservative scene bounds, the fp64 sample lattice
//     (repeated addition, main.go:147,176-196) is read from a host-built table so step
//     counts and positions match the Go loops exactly;
//   * fp32 mode evaluates in fp32 with guard bands around every discontinuous predicate and
//     re-evaluates a sample in fp64 (reference operation order, no FMA contraction) when any
//     predicate is within its error bound of flipping; hierarchical refinements are deferred
//     into a per-lane queue so a warp refines together instead of one lane at a time;
//   * fp64 mode runs the same instruction stream with the reference's exact operation order.
#include "eval.cuh"

namespace xr {

// ---------------------------------------------------------------------------------------
// fp64 ("exact") kernel: reference loops verbatim, minus steps proven to add exactly 0.
// ---------------------------------------------------------------------------------------
template <int INTEG>
__global__ void __launch_bounds__(kBlockThreads) render_scene_exact_kernel(const RenderParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneView S = stage_program(P, smem);
    SaveStack<Exact> st;
    st.r = reinterpret_cast<double*>(smem + P.smem_prog_bytes);
    st.u = reinterpret_cast<unsigned int*>(st.r + (size_t)P.scene.save_depth * 5 * blockDim.x);

    XR_TILE_LOOP_BEGIN(P)
    int view, i, j;
    pixel_of_thread(P, xr_block_, xr_sub_, view, i, j);
    const bool valid = xr_live_ && i < P.res && j < P.res;
    const CamDev& cam = P.cams[view];
    const Ray64 ray = make_ray(cam, valid ? i : 0, valid ? j : 0, P.res);
    double s_in, s_out;
    const bool hit = valid && clip_ray(ray, P.aabb_lo, P.aabb_hi, s_in, s_out);
    int k0, k1;
    step_range(P, hit, s_in, s_out, INTEG == 1 ? 1 : 0, k0, k1);
    const int wk0 = __reduce_min_sync(FULL_MASK, hit ? k0 : 0x7fffffff);
    const int wk1 = __reduce_max_sync(FULL_MASK, hit ? k1 : 0);

    Counters cnt = {0};
    unsigned long long n_eval = 0, n_fine = 0;
    double T = P.flat_field;
    double prev = 0.0;
    bool unc_dummy = false;
    for (int k = wk0; k < wk1; ++k) {
        const bool act = hit && k >= k0 && k < k1;
        if (INTEG == 0) {  // main.go:144-154
            const double s = P.s_tab[k];
            const double x = dadd(ray.o[0], dmul(ray.d[0], s));
            const double y = dadd(ray.o[1], dmul(ray.d[1], s));
            const double z = dadd(ray.o[2], dmul(ray.d[2], s));
            const double rho = dmul(eval_scene<Exact>(S, x, y, z, act, unc_dummy, st, cnt), P.dm);
            if (act) {
                T = dadd(T, dmul(rho, P.ds));
                ++n_eval;
            }
        } else {  // main.go:159-199
            const double right = P.s_tab[k + 1];
            const double x = dadd(ray.o[0], dmul(ray.d[0], right));
            const double y = dadd(ray.o[1], dmul(ray.d[1], right));
            const double z = dadd(ray.o[2], dmul(ray.d[2], right));
            const double rho = dmul(eval_scene<Exact>(S, x, y, z, act, unc_dummy, st, cnt), P.dm);
            const bool trans = act && ((rho == 0.0) != (prev == 0.0));
            if (act) ++n_eval;
            if (__any_sync(FULL_MASK, trans)) {
                double left = dadd(P.s_tab[k], P.ds_fine);
                for (;;) {
                    const bool a = trans && left < right;
                    if (!__any_sync(FULL_MASK, a)) break;
                    const double fx = dadd(ray.o[0], dmul(ray.d[0], left));
                    const double fy = dadd(ray.o[1], dmul(ray.d[1], left));
                    const double fz = dadd(ray.o[2], dmul(ray.d[2], left));
                    const double fr = dmul(eval_scene<Exact>(S, fx, fy, fz, a, unc_dummy, st, cnt), P.dm);
                    if (a) {
                        T = dadd(T, dmul(fr, P.ds_fine));
                        ++n_eval;
                        ++n_fine;
                    }
                    left = dadd(left, P.ds_fine);
                }
            }
            if (act) {
                if (trans) T = dadd(T, dmul(rho, P.ds_fine));
                else T = dadd(T, dmul(rho, P.ds));
                prev = rho;
            }
        }
    }
    store_pixel(P, view, i, j, valid, exp(-T));
    add_stats(P, valid ? (unsigned long long)P.n_steps + n_fine : 0ull, n_eval, 0ull, cnt.prim_tests, valid ? 1ull : 0ull);
    XR_TILE_LOOP_END
}

// ---------------------------------------------------------------------------------------
// fp32 ("fast") kernel
// ---------------------------------------------------------------------------------------
struct FastRay {
    float cx, cy, cz;  // o + d * s_center, rounded once
    float dx, dy, dz;
};

// Density at lattice position s (exact fp64 value) for every lane: fp32 evaluation, then an fp64
// re-evaluation of the lanes whose predicates were inside a guard band.
__device__ __forceinline__ float density_fast(const SceneView& S, const RenderParams& P, const FastRay& fr, float t,
                                              double s_exact, bool act, int view, int i, int j,
                                              SaveStack<Fast> stf, SaveStack<Exact> ste, Counters& cnt,
                                              unsigned long long& n_fallback) {
    bool unc = false;
    const float x = fmaf(fr.dx, t, fr.cx), y = fmaf(fr.dy, t, fr.cy), z = fmaf(fr.dz, t, fr.cz);
    float rho = eval_scene<Fast>(S, x, y, z, act, unc, stf, cnt);
    unc = unc && act;
    if (__any_sync(FULL_MASK, unc)) {
        // rare: rebuild the exact ray instead of keeping 12 fp64 registers alive per lane
        const Ray64 ray = make_ray(P.cams[view], i, j, P.res);
        const double ex = dadd(ray.o[0], dmul(ray.d[0], s_exact));
        const double ey = dadd(ray.o[1], dmul(ray.d[1], s_exact));
        const double ez = dadd(ray.o[2], dmul(ray.d[2], s_exact));
        bool dummy = false;
        const double r64 = eval_scene<Exact>(S, ex, ey, ez, unc, dummy, ste, cnt);
        if (unc) {
            rho = (float)r64;
            if (r64 != 0.0 && rho == 0.0f) rho = r64 > 0 ? 1e-30f : -1e-30f;  // keep zero-ness
            ++n_fallback;
        }
    }
    return rho;
}

template <int INTEG>
__global__ void __launch_bounds__(kBlockThreads) render_scene_fast_kernel(const RenderParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneView S = stage_program(P, smem);
    // save stacks (only when the scene nests) + refinement queue
    SaveStack<Exact> ste;
    ste.r = reinterpret_cast<double*>(smem + P.smem_prog_bytes);
    ste.u = reinterpret_cast<unsigned int*>(ste.r + (size_t)P.scene.save_depth * 5 * blockDim.x);
    // The fp32 walk needs its OWN frames: another warp of this CTA may be inside the fp64 fallback
    // while this warp still has live fp32 frames, and the [slot][thread] layouts of the two element
    // sizes interleave.
    SaveStack<Fast> stf;
    stf.r = reinterpret_cast<float*>(ste.u + (size_t)P.scene.save_depth * 4 * blockDim.x);
    stf.u = reinterpret_cast<unsigned int*>(stf.r + (size_t)P.scene.save_depth * 5 * blockDim.x);
    int* queue = reinterpret_cast<int*>(stf.u + (size_t)P.scene.save_depth * 4 * blockDim.x);

    XR_TILE_LOOP_BEGIN(P)
    int view, i, j;
    pixel_of_thread(P, xr_block_, xr_sub_, view, i, j);
    const bool valid = xr_live_ && i < P.res && j < P.res;
    if (!valid) { i = 0; j = 0; }
    int k0, k1;
    bool hit;
    FastRay fr;
    {
        const Ray64 ray = make_ray(P.cams[view], i, j, P.res);
        double s_in, s_out;
        hit = valid && clip_ray(ray, P.aabb_lo, P.aabb_hi, s_in, s_out);
        step_range(P, hit, s_in, s_out, INTEG == 1 ? 1 : 0, k0, k1);
        fr.cx = (float)(ray.o[0] + ray.d[0] * P.s_center);
        fr.cy = (float)(ray.o[1] + ray.d[1] * P.s_center);
        fr.cz = (float)(ray.o[2] + ray.d[2] * P.s_center);
        fr.dx = (float)ray.d[0];
        fr.dy = (float)ray.d[1];
        fr.dz = (float)ray.d[2];
    }
    const int wk0 = __reduce_min_sync(FULL_MASK, hit ? k0 : 0x7fffffff);
    const int wk1 = __reduce_max_sync(FULL_MASK, hit ? k1 : 0);

    Counters cnt = {0};
    unsigned long long n_eval = 0, n_fine = 0, n_fallback = 0;
    const float dmf = (float)P.dm;
    float accC = 0.0f, accF = 0.0f;  // sum of rho over coarse / fine weights
    double totC = 0.0, totF = 0.0;
    float prev = 0.0f;
    int qn = 0;
    const int tid = threadIdx.x;

    auto flush = [&]() {
        while (__any_sync(FULL_MASK, qn > 0)) {
            const bool has = qn > 0;
            const int k = has ? queue[(--qn) * kBlockThreads + tid] : 0;
            const double right = P.s_tab[k + 1];
            double left = dadd(P.s_tab[k], P.ds_fine);
            for (;;) {
                const bool a = has && left < right;
                if (!__any_sync(FULL_MASK, a)) break;
                const float t = (float)(left - P.s_center);
                const float rho = density_fast(S, P, fr, t, left, a, view, i, j, stf, ste, cnt, n_fallback) * dmf;
                if (a) {
                    accF += rho;
                    ++n_eval;
                    ++n_fine;
                }
                left = dadd(left, P.ds_fine);
            }
        }
    };

    for (int k = wk0; k < wk1; ++k) {
        const bool act = hit && k >= k0 && k < k1;
        if (INTEG == 0) {
            const float rho = density_fast(S, P, fr, P.t_tab[k], P.s_tab[k], act, view, i, j, stf, ste, cnt, n_fallback) * dmf;
            if (act) {
                accC += rho;
                ++n_eval;
            }
        } else {
            const float rho =
                density_fast(S, P, fr, P.t_tab[k + 1], P.s_tab[k + 1], act, view, i, j, stf, ste, cnt, n_fallback) * dmf;
            if (act) {
                ++n_eval;
                if ((rho == 0.0f) != (prev == 0.0f)) {
                    queue[qn * kBlockThreads + tid] = k;
                    ++qn;
                    accF += rho;  // T += rho*ds, main.go:188
                } else {
                    accC += rho;  // T += rho*DS, main.go:190
                }
                prev = rho;
            }
            if (__any_sync(FULL_MASK, qn == kQueueCap)) flush();
        }
        if ((k & 31) == 31) {  // bound fp32 accumulation error: fold into fp64 every 32 steps
            totC += (double)accC;
            totF += (double)accF;
            accC = accF = 0.0f;
        }
    }
    if (INTEG == 1) flush();
    totC += (double)accC;
    totF += (double)accF;
    const double T = P.flat_field + P.ds * totC + P.ds_fine * totF;
    store_pixel(P, view, i, j, valid, exp(-T));
    add_stats(P, valid ? (unsigned long long)P.n_steps + n_fine : 0ull, n_eval, n_fallback, cnt.prim_tests,
              valid ? 1ull : 0ull);
    XR_TILE_LOOP_END
}

// ---------------------------------------------------------------------------------------
// Voxeliser over the same instruction stream (main.go computeVoxel:208-214 semantics):
// out[k*res*res + i*res + j] = density(i/res*2-1, j/res*2-1, k/res*2-1), exact fp64 evaluation.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads) voxelize_scene_kernel(const RenderParams P, int res, float* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneView S = stage_program(P, smem);
    SaveStack<Exact> st;
    st.r = reinterpret_cast<double*>(smem + P.smem_prog_bytes);
    st.u = reinterpret_cast<unsigned int*>(st.r + (size_t)P.scene.save_depth * 5 * blockDim.x);
    const size_t total = (size_t)res * res * res;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < total;
    const size_t id = valid ? idx : 0;
    const int k = (int)(id / ((size_t)res * res));
    const int i = (int)((id / res) % res);
    const int j = (int)(id % res);
    const double x = dsub(dmul(ddiv((double)i, (double)res), 2.0), 1.0);
    const double y = dsub(dmul(ddiv((double)j, (double)res), 2.0), 1.0);
    const double z = dsub(dmul(ddiv((double)k, (double)res), 2.0), 1.0);
    Counters cnt = {0};
    bool dummy = false;
    const double rho = dmul(eval_scene<Exact>(S, x, y, z, valid, dummy, st, cnt), P.dm);
    if (valid) out[idx] = (float)rho;
}

// ---------------------------------------------------------------------------------------
// Clipping probe (main.go:162-169): the hierarchical integrator evaluates density() at smin and smax of every ray
// and warns once when the object sticks out of the sample window.  Here: the same two exact evaluations per ray, OR-ed
// into stats[7] (bit 17: density > 0 at smin for some ray, bit 18: at smax).  Runs only when the caller asked for stats.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads) clip_probe_kernel(const RenderParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneView S = stage_program(P, smem);
    SaveStack<Exact> st;
    st.r = reinterpret_cast<double*>(smem + P.smem_prog_bytes);
    st.u = reinterpret_cast<unsigned int*>(st.r + (size_t)P.scene.save_depth * 5 * blockDim.x);
    int view, i, j;
    pixel_of_thread(P, blockIdx.x, threadIdx.x >> 5, view, i, j);
    const bool valid = i < P.res && j < P.res;
    const Ray64 ray = make_ray(P.cams[view], valid ? i : 0, valid ? j : 0, P.res);
    unsigned int bits = 0u;
#pragma unroll 1
    for (int e = 0; e < 2; ++e) {
        const double s = e ? P.smax : P.smin;
        const double x = dadd(ray.o[0], dmul(ray.d[0], s)), y = dadd(ray.o[1], dmul(ray.d[1], s)), z = dadd(ray.o[2], dmul(ray.d[2], s));
        Counters cnt = {0};
        bool dummy = false;
        const double rho = dmul(eval_scene<Exact>(S, x, y, z, valid, dummy, st, cnt), P.dm);
        if (valid && rho > 0.0) bits |= 1u << (17 + e);
    }
    bits = __reduce_or_sync(FULL_MASK, bits);
    if ((threadIdx.x & 31) == 0 && bits && P.stats) atomicOr(P.stats + 7, (unsigned long long)bits);
}

// ---------------------------------------------------------------------------------------
// Launchers (called from api.cu)
// ---------------------------------------------------------------------------------------
cudaError_t launch_clip_probe(const RenderParams& P, cudaStream_t stream) {
    size_t smem = P.smem_prog_bytes + (size_t)P.scene.save_depth * kBlockThreads * (5 * sizeof(double) + 4 * sizeof(unsigned int));
    const unsigned int grid = (unsigned int)((size_t)P.n_views * P.tiles_i * P.tiles_j);
    if (grid == 0 || !P.stats) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(clip_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    clip_probe_kernel<<<grid, kBlockThreads, smem, stream>>>(P);
    return cudaGetLastError();
}

size_t scene_kernel_smem_bytes(const RenderParams& P, bool with_queue) {
    size_t b = P.smem_prog_bytes;
    b += (size_t)P.scene.save_depth * kBlockThreads * (5 * sizeof(double) + 4 * sizeof(unsigned int));
    if (with_queue) b += (size_t)P.scene.save_depth * kBlockThreads * (5 * sizeof(float) + 4 * sizeof(unsigned int));
    if (with_queue) b += (size_t)kQueueCap * kBlockThreads * sizeof(int);
    return b;
}

cudaError_t launch_render_scene(const RenderParams& P, int precision, int integrator, cudaStream_t stream) {
    const size_t smem = scene_kernel_smem_bytes(P, precision == 0);
    unsigned int grid = (unsigned int)((size_t)P.n_views * P.tiles_i * P.tiles_j);
    if (grid == 0) return cudaSuccess;
    if (P.tile_list) grid = grid < kTileListGrid ? grid : kTileListGrid;
#define XR_LAUNCH(KERNEL)                                                                             \
    do {                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                               \
        KERNEL<<<grid, kBlockThreads, smem, stream>>>(P);                                             \
    } while (0)
    if (precision == 0) {
        if (integrator == 0) XR_LAUNCH(render_scene_fast_kernel<0>);
        else XR_LAUNCH(render_scene_fast_kernel<1>);
    } else {
        if (integrator == 0) XR_LAUNCH(render_scene_exact_kernel<0>);
        else XR_LAUNCH(render_scene_exact_kernel<1>);
    }
#undef XR_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_voxelize_scene(const RenderParams& P, int res, float* d_out, cudaStream_t stream) {
    const size_t smem = scene_kernel_smem_bytes(P, false);
    const size_t total = (size_t)res * res * res;
    const size_t blocks = (total + kBlockThreads - 1) / kBlockThreads;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(voxelize_scene_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    voxelize_scene_kernel<<<(unsigned int)blocks, kBlockThreads, smem, stream>>>(P, res, d_out);
    return cudaGetLastError();
}

}  // namespace xr
