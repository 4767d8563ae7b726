// render_volume.cu -- dedicated voxel-volume projection kernel (fp32 mode, fixed-step integrator),
// the path behind RenderVolumeProjectionsCUDA / XRayRenderVolume*CUDA.
//
// Replaces reference cuda_backend.cu:19-80 (render_kernel: hardware 9-bit trilinear texture, fp32
// `s += ds`, half-voxel texture convention) with the Go CPU semantics the north star asks for:
// VoxelGrid.Density (objects.go:789-855: corner aligned (N-1) scaling, zero outside [-1,1]^3,
// x1 = min(x0+1, N-1)) sampled on the fp64 repeated-addition lattice of integrate_along_ray
// (main.go:144-154).
//
// One thread per ray.  Per ray, in fp64: camera ray, the [s_in, s_out] interval inside the cube,
// and the ray in voxel-index space u(t) = uc + du * t (t = s - R).  The march itself is fp32:
// 3 FMA for the index-space position, manual trilinear from 8 taps.  Samples within the guard
// band of a cube face (where Density() jumps to 0) are evaluated by the exact fp64 routine, so the
// discontinuity is classified exactly as the reference does; everywhere else the integrand is
// continuous and fp32 rounding stays ~1e-6 of T.
#include "eval.cuh"

namespace xr {

struct VolRay {
    float ucx, ucy, ucz;  // index-space position at s = s_center
    float dux, duy, duz;  // index-space direction per unit s
};

__device__ __forceinline__ float trilinear_fast(const float* __restrict__ vol, int nx, int ny, int nz, float ux, float uy,
                                                float uz) {
    // magic-number floor: (u - 0.5 + 1.5*2^23) - 1.5*2^23 rounds to nearest; w in [-eps, 1+eps]
    // is harmless because the interpolant is continuous across cell faces.
    float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
    int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    x0 = max(0, min(x0, nx - 1));
    y0 = max(0, min(y0, ny - 1));
    z0 = max(0, min(z0, nz - 1));
    const int x1 = min(x0 + 1, nx - 1), y1 = min(y0 + 1, ny - 1), z1 = min(z0 + 1, nz - 1);
    const float wx = ux - fx, wy = uy - fy, wz = uz - fz;
    const size_t sz = (size_t)nx * ny;
    const float* p00 = vol + (size_t)z0 * sz + (size_t)x0 * ny;
    const float* p01 = vol + (size_t)z1 * sz + (size_t)x0 * ny;
    const float* p10 = vol + (size_t)z0 * sz + (size_t)x1 * ny;
    const float* p11 = vol + (size_t)z1 * sz + (size_t)x1 * ny;
    const float v000 = __ldg(p00 + y0), v010 = __ldg(p00 + y1);
    const float v001 = __ldg(p01 + y0), v011 = __ldg(p01 + y1);
    const float v100 = __ldg(p10 + y0), v110 = __ldg(p10 + y1);
    const float v101 = __ldg(p11 + y0), v111 = __ldg(p11 + y1);
    const float v00 = fmaf(wz, v001 - v000, v000), v01 = fmaf(wz, v011 - v010, v010);
    const float v10 = fmaf(wz, v101 - v100, v100), v11 = fmaf(wz, v111 - v110, v110);
    const float v0 = fmaf(wy, v01 - v00, v00), v1 = fmaf(wy, v11 - v10, v10);
    return fmaf(wx, v1 - v0, v0);
}

__global__ void __launch_bounds__(kBlockThreads) render_volume_fast_kernel(const RenderParams P,
                                                                          const float* __restrict__ vol, int nx, int ny,
                                                                          int nz, float tol) {
    int view, i, j;
    pixel_of_thread(P, blockIdx.x, threadIdx.x >> 5, view, i, j);
    const bool valid = i < P.res && j < P.res;
    if (!valid) { i = 0; j = 0; }
    const Ray64 ray = make_ray(P.cams[view], i, j, P.res);

    // outer interval: cube grown by tol (outside it Density() == 0 exactly);
    // inner interval: cube shrunk by tol (inside it no bounds test can flip).
    double s_in, s_out, q_in, q_out;
    const double g = (double)tol;
    const double lo_o[3] = {-1.0 - g, -1.0 - g, -1.0 - g}, hi_o[3] = {1.0 + g, 1.0 + g, 1.0 + g};
    const double lo_i[3] = {-1.0 + g, -1.0 + g, -1.0 + g}, hi_i[3] = {1.0 - g, 1.0 - g, 1.0 - g};
    const bool hit = valid && clip_ray(ray, lo_o, hi_o, s_in, s_out);
    const bool hit_inner = hit && clip_ray(ray, lo_i, hi_i, q_in, q_out);
    int k0, k1, m0 = 0, m1 = 0;
    step_range(P, hit, s_in, s_out, 0, k0, k1);
    if (hit_inner) {
        // strictly inside: first index with s_k >= q_in (+1 for lattice drift), last with s_k <= q_out (-1)
        double a = ceil((q_in - P.smin) / P.ds) + 1.0, b = floor((q_out - P.smin) / P.ds) - 1.0;
        a = fmin(fmax(a, (double)k0), (double)k1);
        b = fmin(fmax(b + 1.0, a), (double)k1);  // exclusive end
        m0 = (int)a;
        m1 = (int)b;
    } else {
        m0 = m1 = k0;
    }
    VolRay vr;
    {
        const double hx = 0.5 * (double)(nx - 1), hy = 0.5 * (double)(ny - 1), hz = 0.5 * (double)(nz - 1);
        vr.ucx = (float)((ray.o[0] + ray.d[0] * P.s_center + 1.0) * hx);
        vr.ucy = (float)((ray.o[1] + ray.d[1] * P.s_center + 1.0) * hy);
        vr.ucz = (float)((ray.o[2] + ray.d[2] * P.s_center + 1.0) * hz);
        vr.dux = (float)(ray.d[0] * hx);
        vr.duy = (float)(ray.d[1] * hy);
        vr.duz = (float)(ray.d[2] * hz);
    }
    const int wk0 = __reduce_min_sync(FULL_MASK, hit ? k0 : 0x7fffffff);
    const int wk1 = __reduce_max_sync(FULL_MASK, hit ? k1 : 0);

    VoxelDev vd;
    vd.data = vol;
    vd.nx = nx;
    vd.ny = ny;
    vd.nz = nz;
    vd.dtype = 0;
    unsigned long long n_eval = 0, n_fallback = 0;
    float acc = 0.0f;
    double tot = 0.0;
    for (int k = wk0; k < wk1; ++k) {
        const bool act = hit && k >= k0 && k < k1;
        const bool inner = act && k >= m0 && k < m1;
        const bool band = act && !inner;
        if (__any_sync(FULL_MASK, inner)) {
            const float t = P.t_tab[k];
            const float rho = trilinear_fast(vol, nx, ny, nz, fmaf(vr.dux, t, vr.ucx), fmaf(vr.duy, t, vr.ucy),
                                             fmaf(vr.duz, t, vr.ucz));
            if (inner) {
                acc += rho;
                ++n_eval;
            }
        }
        if (__any_sync(FULL_MASK, band)) {
            if (band) {
                const double s = P.s_tab[k];
                const double x = dadd(ray.o[0], dmul(ray.d[0], s));
                const double y = dadd(ray.o[1], dmul(ray.d[1], s));
                const double z = dadd(ray.o[2], dmul(ray.d[2], s));
                acc += (float)voxel_exact(vd, x, y, z);
                ++n_eval;
                ++n_fallback;
            }
        }
        if ((k & 15) == 15) {
            tot += (double)acc;
            acc = 0.0f;
        }
    }
    tot += (double)acc;
    const double T = P.flat_field + P.ds * (tot * P.dm);
    store_pixel(P, view, i, j, valid, exp(-T));
    add_stats(P, valid ? (unsigned long long)P.n_steps : 0ull, n_eval, n_fallback, n_eval, valid ? 1ull : 0ull);
}

// ---------------------------------------------------------------------------------------
// Cell-synchronous variant on a layered 2D array (texture gather).
//
// ncu on the __ldg kernel above: l1tex wavefronts at 96 % of peak, 15.6 sectors per request -- the
// eight scattered taps per sample saturate the L1 tag stage long before the ALUs.  Here the volume
// lives in a cudaArray with layers = z and (width, height) = (y, x); `tld4.r.a2d` returns the four
// exact texels of a 2x2 (y, x) footprint, so a cell's eight corners cost two texture instructions
// and the texture unit does the address arithmetic and the x1 = min(x0+1, N-1) clamp
// (objects.go:815-823) for free.  The march is organised by CELL, not by sample: the warp fetches
// each lane's current cell together, then every lane consumes the ~3 lattice samples that fall inside
// its cell from registers (5 samples per voxel edge at the reference's auto step).  Trilinear
// interpolation is continuous across cell faces, so which of two adjacent cells a face sample is
// charged to does not matter; samples within the guard band of the cube faces still take the exact
// fp64 routine.
// ---- empty-space map: bricks of kBrick^3 cells whose every corner voxel is +-0, with the Chebyshev distance (in
// bricks) to the nearest brick that is not.  A sample inside such a region is exactly 0 in the reference
// (objects.go:826-853 interpolates zeros), so stepping over it changes nothing.  Rebuilt per call: the
// volume's contents may have changed.
constexpr int kBrick = 8;
constexpr int kOccPasses = 8;  // even: the result lands back in the first ping-pong buffer

// one CTA per brick row (fixed brick X, Z): stream the (kBrick+1)^2 voxel rows it touches, coalesced along y.
// With a surface (surf != 0) the same pass also fills the layered array the render kernel gathers from
// (rows x in [8X, 8X+8), z in [8Z, 8Z+8) are this CTA's to write; the +1 halo rows are only inspected), which
// replaces a separate cudaMemcpy3D of the whole volume.
__global__ void __launch_bounds__(256) brick_occupancy_kernel(const float* __restrict__ vol, int nx, int ny, int nz, int bnx, int bny,
                                                              unsigned char* __restrict__ occ, cudaSurfaceObject_t surf) {
    extern __shared__ unsigned int s_flag[];  // one word per 32 voxels along y
    const int bX = blockIdx.x % bnx, bZ = blockIdx.x / bnx;
    const int nwords = (ny + 31) >> 5;
    for (int w = threadIdx.x; w < nwords; w += blockDim.x) s_flag[w] = 0u;
    __syncthreads();
    const int x0 = bX * kBrick, z0 = bZ * kBrick;
    const int x1 = min(x0 + kBrick, nx - 1), z1 = min(z0 + kBrick, nz - 1);
    for (int z = z0; z <= z1; ++z)
        for (int x = x0; x <= x1; ++x) {
            const float* __restrict__ row = vol + ((size_t)z * nx + x) * ny;
            const bool mine = surf != 0 && x < x0 + kBrick && z < z0 + kBrick;
            for (int y = threadIdx.x; y < ny; y += blockDim.x) {
                const float v = row[y];
                if (mine) surf2DLayeredwrite(v, surf, y * (int)sizeof(float), x, z);
                const bool nz_ = (__float_as_uint(v) << 1) != 0u;  // anything but +-0 (NaN and Inf included)
                const unsigned int m = __ballot_sync(__activemask(), nz_);
                // lanes of a warp cover 32 consecutive y starting at a multiple of 32 (blockDim = 256, y = tid + k*256)
                if ((threadIdx.x & 31) == 0 && m) atomicOr(&s_flag[y >> 5], m);
            }
        }
    __syncthreads();
    for (int bY = threadIdx.x; bY < bny; bY += blockDim.x) {
        const int y0 = bY * kBrick, y1 = min(y0 + kBrick, ny - 1);
        bool any = false;
        for (int y = y0; y <= y1; ++y) any = any || ((s_flag[y >> 5] >> (y & 31)) & 1u);
        occ[((size_t)bZ * bnx + bX) * bny + bY] = any ? 0 : 255;  // 0 = occupied, 255 = empty, distance unknown yet
    }
}

// one relaxation step of the Chebyshev distance transform: empty bricks take 1 + min over the 26 neighbours
__global__ void brick_distance_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int bnx, int bny, int bnz) {
    const size_t n = (size_t)bnx * bny * bnz;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int bY = (int)(idx % bny), bX = (int)((idx / bny) % bnx), bZ = (int)(idx / ((size_t)bny * bnx));
    unsigned int v = in[idx];
    if (v != 0u) {
        unsigned int m = 254u;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy) {
                    const int z = bZ + dz, x = bX + dx, y = bY + dy;
                    if (z < 0 || z >= bnz || x < 0 || x >= bnx || y < 0 || y >= bny) continue;  // outside the cube is empty too
                    m = min(m, (unsigned int)in[((size_t)z * bnx + x) * bny + y]);
                }
        v = min(255u, m + 1u);
    }
    out[idx] = (unsigned char)v;
}

__device__ __forceinline__ float4 gather_layer(cudaTextureObject_t tex, int layer, float tx, float ty) {
    float4 r;
    asm volatile("tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%7}];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(tex), "r"(layer), "f"(tx), "f"(ty));
    return r;  // .w=(y0,x0) .z=(y1,x0) .x=(y0,x1) .y=(y1,x1)
}

// WI x WJ = the pixel footprint of one warp (i = camera x, j = camera y; j is the fast output index).  The best
// shape is the one whose lanes share texture layers (z): the launcher picks it from the camera orientation.
//
// INTEG = 1: integrate_hierarchical (main.go:159-199).  Its coarse samples are the lattice positions s_tab[k + 1], k in
// [0, n): the same closed-form sum, over the lattice shifted by one.  On top of it every coarse interval whose sample
// flips (rho == 0) != (prev_rho == 0) is refined: T += ds_fine * (sum of its fine samples + rho) instead of DS * rho.  Zero-ness
// is exact by construction: a sample safely inside a cell whose 8 corners are all 0 is 0; safely inside a cell whose
// corners are not all 0 and share a sign it is not 0; every other sample (closer than tolw to a cell face in fp32, or in
// a mixed-sign cell with a small value) is evaluated with the reference's own fp64 expression.  The refined intervals (a
// handful per ray: where it enters or leaves the support of the volume) are evaluated sample by sample in fp64.
template <int WI, int WJ, int INTEG>
__global__ void __launch_bounds__(kBlockThreads, INTEG == 1 ? 6 : 8) render_volume_tex_kernel(const RenderParams P, cudaTextureObject_t tex,
                                                                         const float* __restrict__ vol, int nx, int ny, int nz,
                                                                         float tol, const unsigned char* __restrict__ occ, int bnx,
                                                                         int bny, const unsigned char* __restrict__ nfine) {
    static_assert(WI * WJ == 32, "one warp");
    int view, i, j;
    {  // block = 2 x 2 warps
        const int tj_n = (P.res + 2 * WJ - 1) / (2 * WJ), ti_n = (P.res + 2 * WI - 1) / (2 * WI);
        const int tiles = ti_n * tj_n;
        const int b = blockIdx.x;
        view = b / tiles;
        const int t = b - view * tiles;
        const int ti = t / tj_n, tj = t - ti * tj_n;
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        i = ti * (2 * WI) + (w >> 1) * WI + lane / WJ;
        j = tj * (2 * WJ) + (w & 1) * WJ + lane % WJ;
    }
    const bool valid = i < P.res && j < P.res;
    if (!valid) { i = 0; j = 0; }
    const Ray64 ray = make_ray(P.cams[view], i, j, P.res);
    double s_in, s_out, q_in, q_out;
    const double g = (double)tol;
    const double lo_o[3] = {-1.0 - g, -1.0 - g, -1.0 - g}, hi_o[3] = {1.0 + g, 1.0 + g, 1.0 + g};
    const double lo_i[3] = {-1.0 + g, -1.0 + g, -1.0 + g}, hi_i[3] = {1.0 - g, 1.0 - g, 1.0 - g};
    const bool hit = valid && clip_ray(ray, lo_o, hi_o, s_in, s_out);
    const bool hit_inner = hit && clip_ray(ray, lo_i, hi_i, q_in, q_out);
    int k0, k1, m0 = 0, m1 = 0;
    constexpr int OFF = INTEG == 1 ? 1 : 0;  // sample k sits at s_tab[k + OFF]
    step_range(P, hit, s_in, s_out, OFF, k0, k1);
    if (hit_inner) {
        double a = ceil((q_in - P.smin) / P.ds) + 1.0 - (double)OFF, b = floor((q_out - P.smin) / P.ds) - 1.0 - (double)OFF;
        a = fmin(fmax(a, (double)k0), (double)k1);
        b = fmin(fmax(b + 1.0, a), (double)k1);
        m0 = (int)a;
        m1 = (int)b;
    } else {
        m0 = m1 = k0;
    }
    float ucx, ucy, ucz, dux, duy, duz;
    {
        const double hx = 0.5 * (double)(nx - 1), hy = 0.5 * (double)(ny - 1), hz = 0.5 * (double)(nz - 1);
        ucx = (float)((ray.o[0] + ray.d[0] * P.s_center + 1.0) * hx);
        ucy = (float)((ray.o[1] + ray.d[1] * P.s_center + 1.0) * hy);
        ucz = (float)((ray.o[2] + ray.d[2] * P.s_center + 1.0) * hz);
        dux = (float)(ray.d[0] * hx);
        duy = (float)(ray.d[1] * hy);
        duz = (float)(ray.d[2] * hz);
    }
    unsigned int n_eval = 0, n_fallback = 0;
    double tot = 0.0;
    VoxelDev vd;
    vd.data = vol;
    vd.nx = nx;
    vd.ny = ny;
    vd.nz = nz;
    vd.dtype = 0;
    // the reference's density at lattice parameter s (main.go:147-149 position, objects.go:789-855)
    auto exact_at = [&](double s) -> double {
        const double x = dadd(ray.o[0], dmul(ray.d[0], s));
        const double y = dadd(ray.o[1], dmul(ray.d[1], s));
        const double z = dadd(ray.o[2], dmul(ray.d[2], s));
        return voxel_exact(vd, x, y, z);
    };
    // hierarchical integrator: zero-ness of the last coarse sample, the refinements' sum, their fine-sample count
    bool prev_z = true;  // prev_rho := 0.0 (main.go:179)
    double corr = 0.0;
    unsigned int n_fine = 0;
    // coarse sample k (interval [s_tab[k], s_tab[k + 1]]) is zero (z) or not: refine the interval if that flips
    // The VALUE of a sample safely inside the cube, fp32 (two texture gathers): what the refined intervals' fine samples need
    // (their zero-ness decides nothing); inside the guard band of the cube's faces the exact routine takes over.
    auto value_at = [&](double s) -> double {
        if (!(hit_inner && s > q_in + 1.0e-6 && s < q_out - 1.0e-6)) return exact_at(s);
        const float t = (float)(s - P.s_center);
        const float ux = fmaf(dux, t, ucx), uy = fmaf(duy, t, ucy), uz = fmaf(duz, t, ucz);
        const float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
        const float wx = ux - fx, wy = uy - fy, wz = uz - fz;
        const int z0 = max(0, min(nz - 1, (int)fz)), z1 = min(z0 + 1, nz - 1);
        const float4 a = gather_layer(tex, z0, fy + 1.0f, fx + 1.0f);
        const float4 b = gather_layer(tex, z1, fy + 1.0f, fx + 1.0f);
        // .w = (y0, x0), .z = (y1, x0), .x = (y0, x1), .y = (y1, x1)
        const float v00 = fmaf(wz, b.w - a.w, a.w), v01 = fmaf(wz, b.z - a.z, a.z);
        const float v10 = fmaf(wz, b.x - a.x, a.x), v11 = fmaf(wz, b.y - a.y, a.y);
        const float v0 = fmaf(wy, v01 - v00, v00), v1 = fmaf(wy, v11 - v10, v10);
        return (double)fmaf(wx, v1 - v0, v0);
    };
    auto see = [&](int k, bool z) {
        if (z != prev_z) {
            const int nf = (int)__ldg(nfine + k);
            double s = P.s_tab[k], fsum = 0.0;
            for (int q = 0; q < nf; ++q) {
                s = dadd(s, P.ds_fine);  // main.go:183,189: left += ds
                fsum += value_at(s);
            }
            const double rk = z ? 0.0 : value_at(P.s_tab[k + 1]);
            corr += P.ds_fine * (fsum + rk) - P.ds * rk;
            n_fine += (unsigned int)nf;
            n_fallback += (unsigned int)nf + 1u;
        }
        prev_z = z;
    };
    const bool dm_zero = P.dm == 0.0;  // (rho * 0 == 0 everywhere: nothing ever flips)

    // (1) guard-band samples at the entry and exit faces: exact fp64 (a handful per ray)
    if (INTEG == 0) {
        const int nb_in = m0 - k0, nb = nb_in + (k1 - m1);
        const int wmax = __reduce_max_sync(FULL_MASK, hit ? nb : 0);
        for (int r = 0; r < wmax; ++r) {
            if (hit && r < nb) {
                const int k = r < nb_in ? k0 + r : m1 + (r - nb_in);
                tot += exact_at(P.s_tab[k]);
                ++n_eval;
                ++n_fallback;
            }
        }
    } else if (hit) {  // in lattice order: the entry face now, the exit face after the interior
        for (int k = k0; k < m0; ++k) {
            const double r = exact_at(P.s_tab[k + OFF]);
            tot += r;
            see(k, dm_zero || r == 0.0);
            ++n_eval;
            ++n_fallback;
        }
    }

    // (2) interior samples, cell-synchronous.
    // Cell index by magic-number rounding on the FMA pipe instead of FRND/F2I on the (4x slower) XU pipe:
    // r = (u - 0.5) + 1.5*2^23 rounds u - 0.5 to the nearest integer.  Where that differs from floor(u)
    // (u an exact integer, ties-to-even) the weight comes out as exactly 1 or 0 on the neighbouring cell,
    // which is the same value because the interpolant is continuous across cell faces.
    const float kMagic = 12582912.0f;  // 1.5 * 2^23, bit pattern 0x4B400000 (ulp = 1 there)
    // work with v = u - 0.5 so that v + kMagic rounds to floor(u) (nearest integer of u - 0.5)
    const float vcx = ucx - 0.5f, vcy = ucy - 0.5f, vcz = ucz - 0.5f;
    // Per-cell run length by DDA: the lattice is uniform to ~1e-12, so inside a cell the weights advance
    // by dw = du*ds per sample and the number of samples before the ray leaves the cell is
    // ceil(min_i dist_i/|dw_i|), dist_i = distance to the exit face = 0.5 - (w_i - 0.5)*sign(dw_i).
    // A face sample charged to the "wrong" side is evaluated 1e-6 outside its cell: same value to 1e-6.
    // The interpolant restricted to a straight ray is a cubic in the step index q inside one cell, so the n samples
    // of a cell sum in closed form (c0*n + c1*S1 + c2*S2 + c3*S3, S_p = sum q^p): no per-sample loop, no divergence.
    const float dwx = dux * P.ds_f, dwy = duy * P.ds_f, dwz = duz * P.ds_f;
    const float dxy = dwx * dwy, dxz = dwx * dwz, dyz = dwy * dwz, dxyz = dwx * dwy * dwz;
    // sample parameter t_k = s_k - R without the table: t0 + (k - kref)*ds, ds split hi + lo (half an ulp, like the table)
    const int kref = (P.n_steps >> 1) - OFF;  // (k - kref = lattice index - n_steps / 2)
    const float t0 = P.t_tab[P.n_steps >> 1];
    // fp32 error of an index-space position v = fma(du, t, vc) with |v| <= N, |du| <= N / 2, |t| < 2: half an ulp each for vc,
    // the fma and t0, du's rounding times |t|, one ulp of t times |du|: 2.0e-7 N in all; with margin
    const float tolw = 3.0e-7f * (float)max(nx, max(ny, nz)) + 1.0e-6f;
    const float ds_lo = (float)(P.ds - (double)P.ds_f);
    const float ivx = fabsf(dwx) > 1e-12f ? 1.0f / dwx : 0.0f, hvx = fabsf(dwx) > 1e-12f ? 0.5f / fabsf(dwx) : 1e30f;
    const float ivy = fabsf(dwy) > 1e-12f ? 1.0f / dwy : 0.0f, hvy = fabsf(dwy) > 1e-12f ? 0.5f / fabsf(dwy) : 1e30f;
    const float ivz = fabsf(dwz) > 1e-12f ? 1.0f / dwz : 0.0f, hvz = fabsf(dwz) > 1e-12f ? 0.5f / fabsf(dwz) : 1e30f;
    float acc = 0.0f, cmp = 0.0f;  // Kahan over per-cell partial sums
    unsigned int n_skip = 0;       // samples stepped over inside empty regions
    int k = m0;
    const int kend = hit ? m1 : m0;
    while (__any_sync(FULL_MASK, k < kend)) {
        const bool live = k < kend;
        const float kf = (float)(k - kref);
        const float t = fmaf(kf, P.ds_f, fmaf(kf, ds_lo, t0));
        const float vx = fmaf(dux, t, vcx), vy = fmaf(duy, t, vcy), vz = fmaf(duz, t, vcz);
        const float rx = vx + kMagic, ry = vy + kMagic, rz = vz + kMagic;  // cell id (as magic floats)
        const float fx = rx - kMagic, fy = ry - kMagic, fz = rz - kMagic;  // cell origin (integers)
        float wx = vx - (fx - 0.5f), wy = vy - (fy - 0.5f), wz = vz - (fz - 0.5f);  // = u - f in [0, 1]
        const int z0 = max(0, min(nz - 1, __float_as_int(rz) - 0x4B400000));
        const int z1 = min(z0 + 1, nz - 1);
        const float4 a = gather_layer(tex, z0, fy + 1.0f, fx + 1.0f);
        const float4 b = gather_layer(tex, z1, fy + 1.0f, fx + 1.0f);
        if (live) {
            const float steps = fminf(fmaf(0.5f - wx, ivx, hvx), fminf(fmaf(0.5f - wy, ivy, hvy), fmaf(0.5f - wz, ivz, hvz)));
            int n = max(1, min(kend - k, __float2int_ru(fminf(steps, 1.0e6f))));
            const unsigned int any_bits = (__float_as_uint(a.x) | __float_as_uint(a.y) | __float_as_uint(a.z)) |
                                          (__float_as_uint(a.w) | __float_as_uint(b.x) | __float_as_uint(b.y)) |
                                          (__float_as_uint(b.z) | __float_as_uint(b.w));
            float part = 0.0f;
            float hc0 = 0.0f, hc1 = 0.0f, hc2 = 0.0f, hc3 = 0.0f;  // (INTEG = 1) the cell's cubic along the ray
            bool stretched = false;                                 // the run was stretched over an empty region
            if ((any_bits << 1) == 0u) {
                // all eight corners are +-0: this cell adds nothing.  If the whole brick is empty, run to the edge of
                // the empty region around it: cells [8(b-(d-1)), 8(b+d)) per axis, d = Chebyshev distance in bricks.
                if (occ) {
                    const int ix = min(nx - 1, max(0, __float_as_int(rx) - 0x4B400000));
                    const int iy = min(ny - 1, max(0, __float_as_int(ry) - 0x4B400000));
                    // 255 = still unknown after the kOccPasses relaxations of build_volume_occupancy: at least one more
                    const unsigned int d = min((unsigned int)(kOccPasses + 1), (unsigned int)occ[((size_t)(z0 >> 3) * bnx + (ix >> 3)) * bny + (iy >> 3)]);
                    if (d != 0u) {
                        const float reach = (float)(8u * d - 4u) * 2.0f;  // half-width of the region, times 2 (hv = 0.5/|dw|)
                        const float cx = (float)(4 - (ix & 7)), cy = (float)(4 - (iy & 7)), cz = (float)(4 - (z0 & 7));
                        const float run = fminf(fmaf(cx - wx, ivx, reach * hvx),
                                                fminf(fmaf(cy - wy, ivy, reach * hvy), fmaf(cz - wz, ivz, reach * hvz)));
                        // one step of slack: the run length is fp32 arithmetic on up to ~2000 steps
                        const int m = min(kend - k, __float2int_rd(fminf(run, 1.0e6f)) - 1);
                        if (m > n) {
                            n_skip += (unsigned int)(m - n);
                            n = m;
                            stretched = true;
                        }
                    }
                }
            } else {
            // f = v000 + ax X + ay Y + az Z + axy XY + axz XZ + ayz YZ + axyz XYZ on the cell's corners
            const float v000 = a.w, v010 = a.z, v100 = a.x, v110 = a.y, v001 = b.w, v011 = b.z, v101 = b.x, v111 = b.y;
            const float ax = v100 - v000, ay = v010 - v000, az = v001 - v000;
            const float bx = v101 - v001;
            const float axy = (v110 - v010) - ax, axz = bx - ax, ayz = (v011 - v001) - ay;
            const float axyz = ((v111 - v011) - bx) - axy;
            const float tx = fmaf(wy, axyz, axz), ty = fmaf(wx, axyz, ayz), tz = fmaf(wz, axyz, axy);  // mixed partials
            const float gx = fmaf(wz, tx, fmaf(wy, axy, ax));                                           // gradient at the
            const float gy = fmaf(wz, ty, fmaf(wx, axy, ay));                                           // first sample
            const float gz = fmaf(wy, ty, fmaf(wx, axz, az));
            const float c0 = fmaf(wx, gx, fmaf(wy, fmaf(wz, ayz, ay), fmaf(wz, az, v000)));
            const float c1 = fmaf(dwx, gx, fmaf(dwy, gy, dwz * gz));
            const float c2 = fmaf(dxy, tz, fmaf(dxz, tx, dyz * ty));
            const float c3 = axyz * dxyz;
            const float nf = (float)n;
            const float S1 = 0.5f * nf * (nf - 1.0f);
            const float S2 = S1 * fmaf(2.0f, nf, -1.0f) * 0.333333343f;
            const float S3 = S1 * S1;
            part = fmaf(c3, S3, fmaf(c2, S2, fmaf(c1, S1, c0 * nf)));
            if (INTEG == 1) { hc0 = c0; hc1 = c1; hc2 = c2; hc3 = c3; }
            }
            if (INTEG == 1 && !dm_zero) {
                // Zero-ness of the n coarse samples of this run, in lattice order.  [lead, trail) = the samples that are surely
                // all alike (zero in a cell of zeros, non-zero in a cell whose corners share a sign); the others -- closer than
                // tolw to a face of their cell, or anywhere in a mixed-sign cell unless the value is clearly not 0 -- ask fp64.
                const bool zcell = (any_bits << 1) == 0u;
                // some corner carries a sign bit, some does not (a cell of negatives and +0 counts as mixed: only slower)
                const unsigned int all_bits = (__float_as_uint(a.x) & __float_as_uint(a.y) & __float_as_uint(a.z)) &
                                              (__float_as_uint(a.w) & __float_as_uint(b.x) & __float_as_uint(b.y)) &
                                              (__float_as_uint(b.z) & __float_as_uint(b.w));
                const bool mixed = !zcell && (any_bits >> 31) != 0u && (all_bits >> 31) == 0u;
                auto near = [&](int q) -> bool {  // sample q of the run sits at w + q dw (possibly several cells on)
                    const float qf = (float)q;
                    float ax = fmaf(qf, dwx, wx), ay = fmaf(qf, dwy, wy), az = fmaf(qf, dwz, wz);
                    ax = fabsf(ax - rintf(ax));
                    ay = fabsf(ay - rintf(ay));
                    az = fabsf(az - rintf(az));
                    return fminf(ax, fminf(ay, az)) < tolw;
                };
                int lead = 0, trail = n;
                if (mixed) {
                    lead = n;  // sample by sample
                } else {
                    // the usual case first: neither end of the run is near a face (within one cell w stays in [0, 1]; a run
                    // that was stretched over an empty region ends at least one step inside it)
                    const float m_first = fminf(fminf(fminf(wx, 1.0f - wx), fminf(wy, 1.0f - wy)), fminf(wz, 1.0f - wz));
                    float m_last = 1.0f;
                    if (n > 1 && !stretched) {
                        const float qf = (float)(n - 1);
                        const float lx = fmaf(qf, dwx, wx), ly = fmaf(qf, dwy, wy), lz = fmaf(qf, dwz, wz);
                        m_last = fminf(fminf(fminf(lx, 1.0f - lx), fminf(ly, 1.0f - ly)), fminf(lz, 1.0f - lz));
                    }
                    if (m_first < tolw || m_last < tolw) {
                        while (lead < n && near(lead)) ++lead;
                        while (trail > lead + 1 && near(trail - 1)) --trail;
                    }
                }
                const float vmax = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                                         fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
                for (int q = 0; q < n;) {
                    bool z;
                    int adv = 1;
                    if (q >= lead && q < trail) {
                        z = zcell;
                        adv = trail - q;  // only the first of them can flip
                    } else {
                        bool sure = false;
                        if (mixed) {
                            const float qf = (float)q;
                            const float val = fmaf(qf, fmaf(qf, fmaf(qf, hc3, hc2), hc1), hc0);
                            sure = !near(q) && fabsf(val) > 1.0e-3f * vmax;
                        }
                        if (sure) {
                            z = false;
                        } else {
                            ++n_fallback;
                            z = dmul(exact_at(P.s_tab[k + q + OFF]), P.dm) == 0.0;
                        }
                    }
                    see(k + q, z);
                    q += adv;
                }
            }
            k += n;
            const float y_ = part - cmp;
            const float t_ = acc + y_;
            cmp = (t_ - acc) - y_;
            acc = t_;
        }
    }
    if (hit) n_eval += (unsigned int)(m1 - m0) - n_skip;
    tot += (double)acc - (double)cmp;
    if (INTEG == 1 && hit) {
        for (int k = m1; k < k1; ++k) {  // the exit face
            const double r = exact_at(P.s_tab[k + OFF]);
            tot += r;
            see(k, dm_zero || r == 0.0);
            ++n_eval;
            ++n_fallback;
        }
        if (k1 < P.n_steps) see(k1, true);  // the first sample beyond the cube is 0 (objects.go:795)
    }
    const double T = P.flat_field + (P.ds * tot + corr) * P.dm;
    if (WJ % 4 == 0) {
        store_pixel(P, view, i, j, valid, exp(-T));
    } else if (valid) {  // narrow warp footprints: lanes 4q..4q+3 are not 4 consecutive j
        const size_t idx = ((size_t)view * P.res + i) * P.res + j;
        if (P.out_f64) reinterpret_cast<double*>(P.out)[idx] = exp(-T);
        else reinterpret_cast<float*>(P.out)[idx] = (float)exp(-T);
    }
    add_stats(P, valid ? (unsigned long long)P.n_steps + n_fine : 0ull, n_eval, n_fallback, n_eval, valid ? 1ull : 0ull);
}

// ---------------------------------------------------------------------------------------------------------
// fp64 mode (the <= 1e-9 verification mode) for a scene that is one voxel grid: the same cell-synchronous walk in double.
// Within one cell the interpolant along the ray is a cubic in the ray parameter; its coefficients come from the eight
// corners once per cell and every lattice sample of the cell costs a Horner step at its own parameter s_tab[k] - s_tab[k0]
// (the lattice is NOT assumed uniform here: its repeated-addition drift of ~1e-12 would show at 1e-9).  Against the
// reference's staged lerp the cubic differs by rounding only (~1e-15 relative), far inside the gate.  A sample the fp64
// position puts within 1e-9 voxel of a cell face may be charged to the neighbouring cell: same value, the interpolant is
// continuous.  Guard band at the cube's faces, the hierarchical integrator's zero tests and refinements: as in
// render_volume_tex_kernel, with the reference's own expression (voxel_exact) wherever a decision is taken.
// DT: 0 = fp32 voxels, 1 = fp64 voxels (objects.go:789-855 holds float64; the extended entry points accept either).
// ---------------------------------------------------------------------------------------------------------
template <int INTEG, int DT>
__global__ void __launch_bounds__(kBlockThreads, 3) render_volume_f64_kernel(const RenderParams P, const void* __restrict__ vol, int nx, int ny,
                                                                            int nz, const unsigned char* __restrict__ nfine) {
    int view, i, j;
    pixel_of_thread(P, blockIdx.x, threadIdx.x >> 5, view, i, j);
    const bool valid = i < P.res && j < P.res;
    if (!valid) { i = 0; j = 0; }
    const Ray64 ray = make_ray(P.cams[view], i, j, P.res);
    constexpr int OFF = INTEG == 1 ? 1 : 0;  // sample k sits at s_tab[k + OFF]
    const double g = 1.0e-9;                 // guard band at the cube's faces (fp64 positions are good to ~1e-15)
    const double lo_o[3] = {-1.0 - g, -1.0 - g, -1.0 - g}, hi_o[3] = {1.0 + g, 1.0 + g, 1.0 + g};
    const double lo_i[3] = {-1.0 + g, -1.0 + g, -1.0 + g}, hi_i[3] = {1.0 - g, 1.0 - g, 1.0 - g};
    double s_in, s_out, q_in, q_out;
    const bool hit = valid && clip_ray(ray, lo_o, hi_o, s_in, s_out);
    const bool hit_inner = hit && clip_ray(ray, lo_i, hi_i, q_in, q_out);
    int k0, k1, m0 = 0, m1 = 0;
    step_range(P, hit, s_in, s_out, OFF, k0, k1);
    if (hit_inner) {
        double a = ceil((q_in - P.smin) / P.ds) + 1.0 - (double)OFF, b = floor((q_out - P.smin) / P.ds) - 1.0 - (double)OFF;
        a = fmin(fmax(a, (double)k0), (double)k1);
        b = fmin(fmax(b + 1.0, a), (double)k1);
        m0 = (int)a;
        m1 = (int)b;
    } else {
        m0 = m1 = k0;
    }
    VoxelDev vd;
    vd.data = vol;
    vd.nx = nx;
    vd.ny = ny;
    vd.nz = nz;
    vd.dtype = DT;
    auto exact_at = [&](double s) -> double {
        const double x = dadd(ray.o[0], dmul(ray.d[0], s));
        const double y = dadd(ray.o[1], dmul(ray.d[1], s));
        const double z = dadd(ray.o[2], dmul(ray.d[2], s));
        return voxel_exact(vd, x, y, z);
    };
    unsigned int n_eval = 0, n_fallback = 0, n_fine = 0;
    double tot = 0.0, corr = 0.0;
    bool prev_z = true;  // prev_rho := 0.0 (main.go:179)
    const bool dm_zero = P.dm == 0.0;
    auto see = [&](int k, bool z) {  // coarse sample k of the hierarchical integrator is zero / is not: refine the interval on a flip
        if (z != prev_z) {
            const int nf = (int)__ldg(nfine + k);
            double s = P.s_tab[k], fsum = 0.0;
            for (int q = 0; q < nf; ++q) {
                s = dadd(s, P.ds_fine);  // main.go:183,189
                fsum += exact_at(s);
            }
            const double rk = z ? 0.0 : exact_at(P.s_tab[k + 1]);
            corr += P.ds_fine * (fsum + rk) - P.ds * rk;
            n_fine += (unsigned int)nf;
            n_fallback += (unsigned int)nf + 1u;
        }
        prev_z = z;
    };
    if (hit) {
        for (int k = k0; k < m0; ++k) {  // the entry face
            const double r = exact_at(P.s_tab[k + OFF]);
            tot += r;
            if (INTEG == 1) see(k, dm_zero || dmul(r, P.dm) == 0.0);
            ++n_eval;
            ++n_fallback;
        }
    }
    // the ray in index space, u(s) = ub + du * s (objects.go:796-801: (x + 1) / 2 * (N - 1))
    const double hx = 0.5 * (double)(nx - 1), hy = 0.5 * (double)(ny - 1), hz = 0.5 * (double)(nz - 1);
    const double ubx = (ray.o[0] + 1.0) * hx, uby = (ray.o[1] + 1.0) * hy, ubz = (ray.o[2] + 1.0) * hz;
    const double dux = ray.d[0] * hx, duy = ray.d[1] * hy, duz = ray.d[2] * hz;
    const double tolw = 1.0e-9;  // index-space distance to a cell face below which a sample's cell (hence its zero-ness) is in doubt
    const size_t NY = (size_t)ny, NXY = (size_t)nx * ny;
    const int kend = hit ? m1 : m0;
    int k = m0;
    while (k < kend) {
        const double s0 = P.s_tab[k + OFF];
        const double ux = fma(dux, s0, ubx), uy = fma(duy, s0, uby), uz = fma(duz, s0, ubz);
        int x0 = max(0, min(nx - 1, (int)floor(ux))), y0 = max(0, min(ny - 1, (int)floor(uy))), z0 = max(0, min(nz - 1, (int)floor(uz)));
        const double wx = ux - (double)x0, wy = uy - (double)y0, wz = uz - (double)z0;
        // parameter length until the ray leaves the cell [x0, x0 + 1] x ... (an axis it does not move along never ends it)
        double len = 1.0e300;
        if (dux > 0.0) len = fmin(len, (1.0 - wx) / dux); else if (dux < 0.0) len = fmin(len, -wx / dux);
        if (duy > 0.0) len = fmin(len, (1.0 - wy) / duy); else if (duy < 0.0) len = fmin(len, -wy / duy);
        if (duz > 0.0) len = fmin(len, (1.0 - wz) / duz); else if (duz < 0.0) len = fmin(len, -wz / duz);
        // samples k .. k + n - 1 lie before that point (at least this one; the count is settled on the lattice table itself)
        int n = (int)fmin(fmax(len / P.ds, 0.0), 1.0e6) + 1;
        n = max(1, min(n, kend - k));
        while (n < kend - k && P.s_tab[k + n + OFF] - s0 < len) ++n;
        while (n > 1 && P.s_tab[k + n - 1 + OFF] - s0 >= len) --n;
        const int x1 = min(x0 + 1, nx - 1), y1 = min(y0 + 1, ny - 1), z1 = min(z0 + 1, nz - 1);
#define XR_V(zz, xx, yy) (DT == 0 ? (double)__ldg((const float*)vol + ((size_t)(zz) * NXY + (size_t)(xx) * NY + (yy))) \
                                  : __ldg((const double*)vol + ((size_t)(zz) * NXY + (size_t)(xx) * NY + (yy))))
        const double v000 = XR_V(z0, x0, y0), v001 = XR_V(z1, x0, y0), v010 = XR_V(z0, x0, y1), v011 = XR_V(z1, x0, y1);
        const double v100 = XR_V(z0, x1, y0), v101 = XR_V(z1, x1, y0), v110 = XR_V(z0, x1, y1), v111 = XR_V(z1, x1, y1);
#undef XR_V
        const bool zcell = v000 == 0.0 && v001 == 0.0 && v010 == 0.0 && v011 == 0.0 && v100 == 0.0 && v101 == 0.0 && v110 == 0.0 && v111 == 0.0;
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
        if (!zcell) {
            // f = v000 + ax X + ay Y + az Z + axy XY + axz XZ + ayz YZ + axyz XYZ; along the ray X = wx + dux tau, ...
            const double ax = v100 - v000, ay = v010 - v000, az = v001 - v000;
            const double bx = v101 - v001;
            const double axy = (v110 - v010) - ax, axz = bx - ax, ayz = (v011 - v001) - ay;
            const double axyz = ((v111 - v011) - bx) - axy;
            const double tx = fma(wy, axyz, axz), ty = fma(wx, axyz, ayz), tz = fma(wz, axyz, axy);
            const double gx = fma(wz, tx, fma(wy, axy, ax)), gy = fma(wz, ty, fma(wx, axy, ay)), gz = fma(wy, ty, fma(wx, axz, az));
            c0 = fma(wx, gx, fma(wy, fma(wz, ayz, ay), fma(wz, az, v000)));
            c1 = fma(dux, gx, fma(duy, gy, duz * gz));
            c2 = fma(dux * duy, tz, fma(dux * duz, tx, duy * duz * ty));
            c3 = axyz * dux * duy * duz;
            double part = c0;  // sample 0 sits at tau = 0
            for (int q = 1; q < n; ++q) {
                const double tau = P.s_tab[k + q + OFF] - s0;
                part += fma(tau, fma(tau, fma(tau, c3, c2), c1), c0);
            }
            tot += part;
        }
        if (INTEG == 1 && !dm_zero) {
            // zero-ness of the n samples, in lattice order: sure inside a cell of zeros / of same-sign corners, away from the faces
            const double vmin = fmin(fmin(fmin(v000, v001), fmin(v010, v011)), fmin(fmin(v100, v101), fmin(v110, v111)));
            const double vmax = fmax(fmax(fmax(v000, v001), fmax(v010, v011)), fmax(fmax(v100, v101), fmax(v110, v111)));
            const bool mixed = vmin < 0.0 && vmax > 0.0;
            const double amax = fmax(-vmin, vmax);
            for (int q = 0; q < n;) {
                const double tau = P.s_tab[k + q + OFF] - s0;
                const double px = fma(dux, tau, wx), py = fma(duy, tau, wy), pz = fma(duz, tau, wz);
                const double dface = fmin(fmin(fmin(px, 1.0 - px), fmin(py, 1.0 - py)), fmin(pz, 1.0 - pz));
                bool z, sure = dface > tolw;
                if (sure && mixed) sure = fabs(fma(tau, fma(tau, fma(tau, c3, c2), c1), c0)) > 1.0e-9 * amax;
                int adv = 1;
                if (sure) {
                    z = zcell;
                    if (!mixed) {  // every later sample up to the last is as sure as this one and the last (w moves monotonically)
                        const double tl = P.s_tab[k + n - 1 + OFF] - s0;
                        const double lx = fma(dux, tl, wx), ly = fma(duy, tl, wy), lz = fma(duz, tl, wz);
                        const double dl = fmin(fmin(fmin(lx, 1.0 - lx), fmin(ly, 1.0 - ly)), fmin(lz, 1.0 - lz));
                        adv = dl > tolw ? n - q : 1;
                    }
                } else {
                    ++n_fallback;
                    z = dmul(exact_at(P.s_tab[k + q + OFF]), P.dm) == 0.0;
                }
                see(k + q, z);
                q += adv;
            }
        }
        k += n;
    }
    if (hit) n_eval += (unsigned int)(m1 - m0);
    if (hit) {
        for (int kk = m1; kk < k1; ++kk) {  // the exit face
            const double r = exact_at(P.s_tab[kk + OFF]);
            tot += r;
            if (INTEG == 1) see(kk, dm_zero || dmul(r, P.dm) == 0.0);
            ++n_eval;
            ++n_fallback;
        }
        if (INTEG == 1 && k1 < P.n_steps) see(k1, true);  // the first sample beyond the cube is 0 (objects.go:795)
    }
    const double T = P.flat_field + (P.ds * tot + corr) * P.dm;
    store_pixel(P, view, i, j, valid, exp(-T));
    add_stats(P, valid ? (unsigned long long)P.n_steps + n_fine : 0ull, n_eval, n_fallback, n_eval, valid ? 1ull : 0ull);
}

cudaError_t launch_render_volume_f64(const void* d_vol, int dtype, int nx, int ny, int nz, const RenderParams& P, int integrator,
                                     const unsigned char* d_nfine, cudaStream_t stream) {
    const size_t grid = (size_t)P.n_views * P.tiles_i * P.tiles_j;
    if (grid == 0) return cudaSuccess;
    if (grid > 0x7fffffffull) return cudaErrorInvalidValue;
    const unsigned int gsz = P.tile_list ? 0u : (unsigned int)grid;
    if (gsz == 0u) return cudaErrorInvalidValue;  // (never used for the interval renderer's hand-over lists)
    if (integrator == 0) {
        if (dtype == 0) render_volume_f64_kernel<0, 0><<<gsz, kBlockThreads, 0, stream>>>(P, d_vol, nx, ny, nz, d_nfine);
        else render_volume_f64_kernel<0, 1><<<gsz, kBlockThreads, 0, stream>>>(P, d_vol, nx, ny, nz, d_nfine);
    } else {
        if (dtype == 0) render_volume_f64_kernel<1, 0><<<gsz, kBlockThreads, 0, stream>>>(P, d_vol, nx, ny, nz, d_nfine);
        else render_volume_f64_kernel<1, 1><<<gsz, kBlockThreads, 0, stream>>>(P, d_vol, nx, ny, nz, d_nfine);
    }
    return cudaGetLastError();
}

static const float kVolumeGuardTol = 4.0e-6f;  // fp32 position error (1e-6, scene_compile.cpp) with margin

cudaError_t launch_render_volume_fast(const float* d_vol, int nx, int ny, int nz, const RenderParams& P, cudaStream_t stream) {
    const unsigned int grid = (unsigned int)((size_t)P.n_views * P.tiles_i * P.tiles_j);
    if (grid == 0) return cudaSuccess;
    render_volume_fast_kernel<<<grid, kBlockThreads, 0, stream>>>(P, d_vol, nx, ny, nz, kVolumeGuardTol);
    return cudaGetLastError();
}

template <int WI, int WJ>
static cudaError_t launch_tex_shape(unsigned long long tex, const float* d_vol, int nx, int ny, int nz, const RenderParams& P,
                                    const unsigned char* occ, int integrator, const unsigned char* d_nfine, cudaStream_t stream) {
    const size_t ti = (size_t)(P.res + 2 * WI - 1) / (2 * WI), tj = (size_t)(P.res + 2 * WJ - 1) / (2 * WJ);
    const size_t grid = (size_t)P.n_views * ti * tj;
    if (grid == 0) return cudaSuccess;
    if (grid > 0x7fffffffull) return cudaErrorInvalidValue;
    if (integrator == 0)
        render_volume_tex_kernel<WI, WJ, 0><<<(unsigned int)grid, kBlockThreads, 0, stream>>>(
            P, (cudaTextureObject_t)tex, d_vol, nx, ny, nz, kVolumeGuardTol, occ, (nx + kBrick - 1) / kBrick, (ny + kBrick - 1) / kBrick, d_nfine);
    else
        render_volume_tex_kernel<WI, WJ, 1><<<(unsigned int)grid, kBlockThreads, 0, stream>>>(
            P, (cudaTextureObject_t)tex, d_vol, nx, ny, nz, kVolumeGuardTol, occ, (nx + kBrick - 1) / kBrick, (ny + kBrick - 1) / kBrick, d_nfine);
    return cudaGetLastError();
}

// warp_shape: 0 = 4 x 8 pixels (i x j), 1 = 32 x 1, 2 = 1 x 32, 3 = 16 x 2, 4 = 2 x 16, 5 = 8 x 4
cudaError_t launch_render_volume_tex(unsigned long long tex, const float* d_vol, int nx, int ny, int nz, const RenderParams& P,
                                     int warp_shape, const unsigned char* occ, int integrator, const unsigned char* d_nfine,
                                     cudaStream_t stream) {
    switch (warp_shape) {
        case 1: return launch_tex_shape<32, 1>(tex, d_vol, nx, ny, nz, P, occ, integrator, d_nfine, stream);
        case 2: return launch_tex_shape<1, 32>(tex, d_vol, nx, ny, nz, P, occ, integrator, d_nfine, stream);
        case 3: return launch_tex_shape<16, 2>(tex, d_vol, nx, ny, nz, P, occ, integrator, d_nfine, stream);
        case 4: return launch_tex_shape<2, 16>(tex, d_vol, nx, ny, nz, P, occ, integrator, d_nfine, stream);
        case 5: return launch_tex_shape<8, 4>(tex, d_vol, nx, ny, nz, P, occ, integrator, d_nfine, stream);
        default: return launch_tex_shape<4, 8>(tex, d_vol, nx, ny, nz, P, occ, integrator, d_nfine, stream);
    }
}

size_t volume_brick_count(int nx, int ny, int nz) {
    return (size_t)((nx + kBrick - 1) / kBrick) * ((ny + kBrick - 1) / kBrick) * ((nz + kBrick - 1) / kBrick);
}

// Fill occ_a with the empty-space map of d_vol (occ_b is scratch of the same size); returns the buffer holding the result.
// surf != 0: also copy the volume into the layered array behind that surface (see brick_occupancy_kernel).
cudaError_t build_volume_occupancy(const float* d_vol, int nx, int ny, int nz, unsigned char* occ_a, unsigned char* occ_b,
                                   const unsigned char** result, unsigned long long surf, cudaStream_t stream) {
    const int bnx = (nx + kBrick - 1) / kBrick, bny = (ny + kBrick - 1) / kBrick, bnz = (nz + kBrick - 1) / kBrick;
    const size_t smem = (size_t)((ny + 31) / 32) * sizeof(unsigned int);
    brick_occupancy_kernel<<<(unsigned int)(bnx * bnz), 256, smem, stream>>>(d_vol, nx, ny, nz, bnx, bny, occ_a, (cudaSurfaceObject_t)surf);
    const size_t n = (size_t)bnx * bny * bnz;
    const unsigned int grid = (unsigned int)((n + 255) / 256);
    unsigned char *src = occ_a, *dst = occ_b;
    // kOccPasses relaxations: distances up to that many bricks (a 64-cell run per look-up) are exact, larger ones are
    // reported as kOccPasses + 1 by the reader; a ray crossing a big void simply looks up a few more times
    for (int it = 0; it < kOccPasses; ++it) {
        brick_distance_kernel<<<grid, 256, 0, stream>>>(src, dst, bnx, bny, bnz);
        unsigned char* t = src;
        src = dst;
        dst = t;
    }
    *result = src;
    return cudaGetLastError();
}

}  // namespace xr
