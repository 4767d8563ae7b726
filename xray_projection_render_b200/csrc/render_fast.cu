// render_fast.cu -- the fp32 hot kernels for the two scene shapes every bundled / benchmark scene
// has: FLAT (a primitive, or one collection of primitives) and TESS (tessellated_obj_coll whose
// unit cell is one collection of primitives).  Anything else takes the generic interpreter in
// render_scene.cu.
//
// What makes these lean (see profiles/ for the ncu evidence that drove it):
//   * the flattened program is addressed as true shared memory (LDS broadcasts), never through a
//     generic pointer;
//   * no per-op dispatch: the shape fixes the prologue (bounds, fold, child-mask lookup) and the
//     epilogue; only the loop over primitive runs is data driven;
//   * ONE evaluation site serves coarse steps, fine (refinement) steps and both integrators: the
//     march is a warp-uniform two-mode loop, so the evaluator is instantiated once and the whole
//     hot loop stays inside the instruction cache;
//   * the fp64 re-evaluation of guard-band samples is a cold, out-of-line function; the exact ray
//     is parked in shared memory instead of being rebuilt or held in registers;
//   * hierarchical refinements are queued per lane and replayed warp-wide (fine step positions and
//     counts come from host tables, so the fp32 path carries no fp64 state).
#include "eval.cuh"

namespace xr {

struct FastArgs {  // uniform per-launch data the hot loop needs, staged in shared memory
    SceneView gsv;  // global-memory view of the program for the exact path
    DeformRec d0;   // the first deformation record: one strain field is the common case, read it with LDS broadcasts
};

// Exact lattice position: coarse sample `base` (nsub = 0), or fine sample j of coarse interval `base`
// (nsub = j + 1): s_tab[base] with ds_fine added nsub times, as main.go:181-186 does.
__device__ __forceinline__ double exact_position(const double* __restrict__ s_tab, int base, int nsub, double ds_fine) {
    double left = s_tab[base];
    for (int q = 0; q < nsub; ++q) left = dadd(left, ds_fine);
    return left;
}

// Cold path: exact fp64 density at lattice position s for the lanes in `need`.
// Out of line on purpose: it must not dilute the hot loop's instruction footprint.
__device__ __noinline__ float exact_density_cold(const SceneView* sv, const double* ray, const double* __restrict__ s_tab, int base,
                                                 int nsub, double ds_fine, double dm, bool need) {
    const int tid = threadIdx.x;
    const double s = exact_position(s_tab, base, nsub, ds_fine);
    // ray[] is [6][blockDim]: o.x o.y o.z d.x d.y d.z
    const double x = dadd(ray[0 * kBlockThreads + tid], dmul(ray[3 * kBlockThreads + tid], s));
    const double y = dadd(ray[1 * kBlockThreads + tid], dmul(ray[4 * kBlockThreads + tid], s));
    const double z = dadd(ray[2 * kBlockThreads + tid], dmul(ray[5 * kBlockThreads + tid], s));
    bool dummy = false;
    Counters cnt = {0};
    SaveStack<Exact> st = {nullptr, nullptr};  // FLAT / TESS programs never need save frames
    const double r64 = dmul(eval_scene<Exact>(*sv, x, y, z, need, dummy, st, cnt), dm);
    float rho = (float)r64;
    if (r64 != 0.0 && rho == 0.0f) rho = r64 > 0 ? 1e-30f : -1e-30f;  // keep zero-ness for the transition test
    return rho;
}

// Cold path of the single-primitive variants (PRIM, n = 1): the same fp64 re-evaluation without the interpreter --
// position, warp chain, tessellation bounds + fold (objects.go:568-582, :458-464), the one primitive, and the
// collection rule for one child (objects.go:422-438).  Pointers are the kernel's global-memory arguments, so the
// loads are plain LDG instead of generic loads through the shared-memory SceneView.  About half the instructions
// of exact_density_cold, which matters for a gyroid unit cell (one crossing in three lands in the guard band).
template <int PRIM>
__device__ __noinline__ float exact_single_cold(const double* __restrict__ f64, const DeformRec* __restrict__ deform, int n_deform,
                                                const double* ray, const double* __restrict__ s_tab, int base, int nsub, double ds_fine,
                                                double dm, int tess_f64_idx, int prim_f64_idx, unsigned int cflags) {
    const int tid = threadIdx.x;
    const double s = exact_position(s_tab, base, nsub, ds_fine);
    double x = dadd(ray[0 * kBlockThreads + tid], dmul(ray[3 * kBlockThreads + tid], s));
    double y = dadd(ray[1 * kBlockThreads + tid], dmul(ray[4 * kBlockThreads + tid], s));
    double z = dadd(ray[2 * kBlockThreads + tid], dmul(ray[5 * kBlockThreads + tid], s));
    for (int i = 0; i < n_deform; ++i) Exact::deform(deform[i], x, y, z);
    SceneView S = {};
    S.f64 = f64;
    bool dummy = false;
    bool inside = true;
    if (tess_f64_idx >= 0) {
        Instr I = {};
        I.f64_idx = (unsigned int)tess_f64_idx;
        inside = Exact::tess(S, I, x, y, z, dummy);
    }
    Instr J = {};
    J.f64_idx = (unsigned int)prim_f64_idx;
    double rho = 0.0;
    bool in;
    if (PRIM == OP_GYROID) in = Exact::gyroid(S, J, 0, x, y, z, rho, dummy);
    else if (PRIM == OP_SPHERE) in = Exact::sphere(S, J, 0, x, y, z, rho, dummy);
    else if (PRIM == OP_BOX) in = Exact::box(S, J, 0, x, y, z, rho, dummy);
    else in = Exact::cyl(S, J, 0, x, y, z, rho, dummy);
    double val = (inside && in) ? rho : 0.0;
    if ((cflags & 0x100u) && !((cflags & F_GREEDY) && val > 0.0)) {  // a collection: sum of one child, clamped to [0,1]
        if (val < 0.0) val = 0.0;
        else if (val > 1.0) val = 1.0;
    }
    const double r64 = dmul(val, dm);
    float r = (float)r64;
    if (r64 != 0.0 && r == 0.0f) r = r64 > 0 ? 1e-30f : -1e-30f;  // keep zero-ness for the transition test
    return r;
}

// Cold path for samples within the guard band of a tessellation's outer box or of a unit-cell face: only the
// period (and the inclusive bounds tests of objects.go:569,459) is in doubt, not the primitives.  Redo position,
// warp and fold in fp64 exactly as the reference does and hand the folded point back to the fp32 pipeline
// (.w = 1 when the sample survives both bounds tests).  ~5x cheaper than a full exact evaluation, and it is
// the whole cost of rays that run inside a cell-face plane (the central pixel row at polar = 90 deg).
__device__ __noinline__ float4 exact_fold_cold(const SceneView* sv, const double* ray, const double* __restrict__ s_tab, int base,
                                               int nsub, double ds_fine, int tess_f64_idx) {
    const int tid = threadIdx.x;
    const double s = exact_position(s_tab, base, nsub, ds_fine);
    double x = dadd(ray[0 * kBlockThreads + tid], dmul(ray[3 * kBlockThreads + tid], s));
    double y = dadd(ray[1 * kBlockThreads + tid], dmul(ray[4 * kBlockThreads + tid], s));
    double z = dadd(ray[2 * kBlockThreads + tid], dmul(ray[5 * kBlockThreads + tid], s));
    for (int i = 0; i < sv->n_deform; ++i) Exact::deform(sv->deform[i], x, y, z);
    Instr I = {};
    I.f64_idx = (unsigned int)tess_f64_idx;
    bool dummy = false;
    const bool ok = Exact::tess(*sv, I, x, y, z, dummy);
    return make_float4((float)x, (float)y, (float)z, ok ? 1.0f : 0.0f);
}


enum Shape { SHAPE_FLAT = 1, SHAPE_TESS = 2 };

// Lane state of a collection evaluation, kept in two floats so no bool has to live in a register:
//   res == 0  : lane still needs a value ("active")      res > 0 : greedy first hit (objects.go:425-427)
//   res  < 0  : lane takes no part (outside bounds / ray inactive)
__device__ __forceinline__ void emit_child(bool in, bool near, float rho, bool greedy, float& res, float& acc, bool& unc,
                                           int& nhit) {
    const bool act = res == 0.0f;
    unc = unc || (near && act);
    if (in && act) {
        if (greedy && rho > 0.0f) {
            res = rho;
        } else {
            acc += rho;
            ++nhit;
        }
    }
}

// ---- primitive tests: `in` = surely inside, `near` = within the guard band (fp64 decides) ----
__device__ __forceinline__ void prim_cyl(const float4* __restrict__ q, float x, float y, float z, bool& in, bool& near, float& rho) {
    const float4 a = q[0], v = q[1], t = q[2];
    const float wx = x - a.x, wy = y - a.y, wz = z - a.z;
    const float cc = (wx * v.x + wy * v.y + wz * v.z) * v.w;
    const float ex = fmaf(-v.x, cc, wx), ey = fmaf(-v.y, cc, wy), ez = fmaf(-v.z, cc, wz);
    const float ar = fmaf(ex, ex, fmaf(ey, ey, ez * ez)) - t.x;  // < 0 inside the radius
    const float ac = fabsf(cc - 0.5f) - 0.5f;                     // <= 0 between the caps
    const float worst = fmaxf(ar * t.w, ac);                      // t.w = tolc/tolr: common scale
    in = worst < -t.z;
    near = !in && worst < t.z;
    rho = a.w;
}
__device__ __forceinline__ void prim_sphere(const float4* __restrict__ q, float x, float y, float z, bool& in, bool& near, float& rho) {
    const float4 a = q[0], b = q[1];
    const float dx = x - a.x, dy = y - a.y, dz = z - a.z;
    const float m = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) - b.x;
    in = m < 0.0f;
    near = fabsf(m) < b.y;
    rho = a.w;
}
__device__ __forceinline__ void prim_box(const float4* __restrict__ q, float x, float y, float z, bool& in, bool& near, float& rho) {
    const float4 a = q[0], b = q[1];
    const float m = fmaxf(fabsf(x - a.x) - b.x, fmaxf(fabsf(y - a.y) - b.y, fabsf(z - a.z) - b.z));
    in = m < 0.0f;
    near = fabsf(m) < b.w;
    rho = a.w;
}
__device__ __forceinline__ void prim_pped(const float4* __restrict__ q, float x, float y, float z, bool& in, bool& near, float& rho) {
    const float4 a = q[0], r0 = q[1], r1 = q[2], r2 = q[3];
    const float dx = x - a.x, dy = y - a.y, dz = z - a.z;
    const float qx = r0.x * dx + r0.y * dy + r0.z * dz;
    const float qy = r1.x * dx + r1.y * dy + r1.z * dz;
    const float qz = r2.x * dx + r2.y * dy + r2.z * dz;
    const float m = fmaxf(fabsf(qx - 0.5f), fmaxf(fabsf(qy - 0.5f), fabsf(qz - 0.5f))) - 0.5f;
    in = m < 0.0f;
    near = fabsf(m) < r0.w;
    rho = a.w;
}
// margin: object-space max-norm distance over which this gyroid's answer cannot change
__device__ __forceinline__ void prim_gyroid(const float4* __restrict__ q, float x, float y, float z, bool& in, bool& near, float& rho,
                                            float& margin) {
    const float4 a = q[0], b = q[1];
    float sx, cx, sy, cy, sz, cz;
    fast_sincos((x - a.x) * b.x, &sx, &cx);
    fast_sincos((y - a.y) * b.x, &sy, &cy);
    fast_sincos((z - a.z) * b.x, &sz, &cz);
    const float t = fabsf(sx * cy + sy * cz + sz * cx) - b.y;
    in = t < 0.0f;
    near = fabsf(t) < b.z;
    rho = a.w;
    margin = fmaxf(fabsf(t) - b.z, 0.0f) * b.w;
}

// Visit the set bits of a run's 64-bit child mask as two 32-bit words (cheap FLO/LOP3 per child).
#define XR_FOR_EACH_CHILD(LO, HI, ...)                                      \
    for (int half_ = 0; half_ < 2; ++half_) {                                \
        unsigned int w_ = half_ ? (HI) : (LO);                               \
        const int base_ = half_ * 32;                                        \
        while (w_) {                                                         \
            const int c = base_ + __ffs((int)w_) - 1;                        \
            w_ &= w_ - 1;                                                    \
            if (greedy && !__any_sync(FULL_MASK, res == 0.0f)) goto finish;  \
            if (COUNT) prim_tests += (res == 0.0f) ? 1u : 0u;                \
            bool in, near;                                                   \
            float rho;                                                       \
            __VA_ARGS__                                                      \
            emit_child(in, near, rho, greedy, res, acc, unc, nhit);          \
        }                                                                    \
    }

// Evaluate the primitive runs [rb, re) of one collection at (x,y,z).
template <bool COUNT>
__device__ __forceinline__ float eval_runs(const Instr* __restrict__ sI, const float4* __restrict__ sF, int rb, int re,
                                           unsigned int cflags, bool has_mask, unsigned int umask_lo, unsigned int umask_hi,
                                           float x, float y, float z, bool alive, bool& unc, unsigned int& prim_tests,
                                           float& clr) {
    // clr: object-space max-norm distance within which no evaluated child can change its answer
    // (only gyroids report one; any other child tested sets it to 0)
    const bool greedy = (cflags & F_GREEDY) != 0;
    float acc = 0.0f, res = alive ? 0.0f : -1.0f;
    int nhit = 0;
    for (int r = rb; r < re; ++r) {
        const uint4 w0 = reinterpret_cast<const uint4*>(sI + r)[0];  // op, n, flags, child_bit
        const unsigned int f32_idx = reinterpret_cast<const uint4*>(sI + r)[1].x;
        unsigned long long bits = w0.y >= 64u ? ~0ull : ((1ull << w0.y) - 1ull);
        if (has_mask) bits &= ((((unsigned long long)umask_hi << 32) | umask_lo) >> w0.w);
        const unsigned int lo = (unsigned int)bits, hi = (unsigned int)(bits >> 32);
        const float4* __restrict__ q = sF + f32_idx;
        if (w0.x == OP_CYL) {
            XR_FOR_EACH_CHILD(lo, hi, { prim_cyl(q + c * kF32Cyl, x, y, z, in, near, rho); clr = 0.0f; })
        } else if (w0.x == OP_SPHERE) {
            XR_FOR_EACH_CHILD(lo, hi, { prim_sphere(q + c * kF32Sphere, x, y, z, in, near, rho); clr = 0.0f; })
        } else if (w0.x == OP_BOX) {
            XR_FOR_EACH_CHILD(lo, hi, { prim_box(q + c * kF32Box, x, y, z, in, near, rho); clr = 0.0f; })
        } else if (w0.x == OP_PPED) {
            XR_FOR_EACH_CHILD(lo, hi, { prim_pped(q + c * kF32Pped, x, y, z, in, near, rho); clr = 0.0f; })
        } else {  // OP_GYROID
            XR_FOR_EACH_CHILD(lo, hi, {
                float margin;
                prim_gyroid(q + c * kF32Gyroid, x, y, z, in, near, rho, margin);
                if (res == 0.0f) clr = fminf(clr, margin);
            })
        }
    }
finish:
    // objects.go:422-438: greedy returns the first positive child unclamped, else sum clamped to [0,1]
    float val = acc;
    if (cflags & 0x100u) val = __saturatef(acc);                // bit 8: a collection (clamps), not a bare primitive
    if (nhit >= 2 && fabsf(acc) < 1e-5f && res == 0.0f) unc = true;  // near-cancelling sum: zero-ness needs fp64
    if (res > 0.0f) val = res;
    return val;
}

// Collections (or bare scenes) that are ONE primitive -- a gyroid unit cell, a pillar array -- skip the run decoding,
// the candidate loop and its votes altogether (objects.go:422-438 for a single child: greedy returns a positive
// rho unclamped, otherwise the "sum" is that one rho, clamped to [0,1] when the parent is a collection).
template <int PRIM, bool COUNT>
__device__ __forceinline__ float eval_single(const float4* __restrict__ q, unsigned int n, unsigned int cflags, bool has_mask,
                                             unsigned int umask_lo, unsigned int umask_hi, float x, float y, float z, bool alive,
                                             bool& unc, unsigned int& prim_tests, float& clr) {
    if (n != 1u) {
        // one run of n like primitives (a strut lattice): the candidate loop without run decoding or type dispatch
        const bool greedy = (cflags & F_GREEDY) != 0;
        float acc = 0.0f, res = alive ? 0.0f : -1.0f;
        int nhit = 0;
        unsigned long long bits = n >= 64u ? ~0ull : ((1ull << n) - 1ull);
        if (has_mask) bits &= ((unsigned long long)umask_hi << 32) | umask_lo;
        const unsigned int lo = (unsigned int)bits, hi = (unsigned int)(bits >> 32);
        if (PRIM == OP_GYROID) {
            XR_FOR_EACH_CHILD(lo, hi, {
                float margin;
                prim_gyroid(q + c * kF32Gyroid, x, y, z, in, near, rho, margin);
                if (res == 0.0f) clr = fminf(clr, margin);
            })
        } else if (PRIM == OP_SPHERE) {
            XR_FOR_EACH_CHILD(lo, hi, { prim_sphere(q + c * kF32Sphere, x, y, z, in, near, rho); clr = 0.0f; })
        } else if (PRIM == OP_BOX) {
            XR_FOR_EACH_CHILD(lo, hi, { prim_box(q + c * kF32Box, x, y, z, in, near, rho); clr = 0.0f; })
        } else {
            XR_FOR_EACH_CHILD(lo, hi, { prim_cyl(q + c * kF32Cyl, x, y, z, in, near, rho); clr = 0.0f; })
        }
    finish:
        float val = acc;
        if (cflags & 0x100u) val = __saturatef(acc);
        if (nhit >= 2 && fabsf(acc) < 1e-5f && res == 0.0f) unc = true;
        if (res > 0.0f) val = res;
        return val;
    }
    bool in, near;
    float rho;
    if (PRIM == OP_GYROID) {
        float margin;
        prim_gyroid(q, x, y, z, in, near, rho, margin);
        if (alive) clr = fminf(clr, margin);
    } else {
        if (PRIM == OP_SPHERE) prim_sphere(q, x, y, z, in, near, rho);
        else if (PRIM == OP_BOX) prim_box(q, x, y, z, in, near, rho);
        else prim_cyl(q, x, y, z, in, near, rho);
        clr = 0.0f;
    }
    if (COUNT) prim_tests += alive ? 1u : 0u;
    unc = unc || (near && alive);
    float val = (in && alive) ? rho : 0.0f;
    if ((cflags & 0x100u) && !((cflags & F_GREEDY) && rho > 0.0f)) val = __saturatef(val);
    return val;
}

// Big collections (> 63 primitive children): every lane owns the ascending child list [lp, le) of its grid
// cell; a warp min-reduction merges the lists, so the union is visited in child order and each candidate
// is tested once by the whole warp (parameters through __ldg: uniform address, one L1 sector).
template <bool COUNT>
__device__ __forceinline__ float eval_list(const float4* __restrict__ F, const unsigned int* __restrict__ tab,
                                           const unsigned short* __restrict__ idx, unsigned int lp, unsigned int le,
                                           unsigned int cflags, float x, float y, float z, bool alive, bool& unc,
                                           unsigned int& prim_tests) {
    const bool greedy = (cflags & F_GREEDY) != 0;
    float acc = 0.0f, res = alive ? 0.0f : -1.0f;
    int nhit = 0;
    for (;;) {
        const unsigned int mine = lp < le ? (unsigned int)__ldg(idx + lp) : 0xFFFFu;
        const unsigned int c = __reduce_min_sync(FULL_MASK, mine);
        if (c == 0xFFFFu) break;
        if (greedy && !__any_sync(FULL_MASK, res == 0.0f)) break;
        if (mine == c) ++lp;
        const unsigned int t = __ldg(tab + c);
        const float4* __restrict__ q = F + (t & 0xFFFFFFu);
        if (COUNT) prim_tests += (res == 0.0f) ? 1u : 0u;
        bool in, near;
        float rho, margin;
        switch (t >> 24) {
            case OP_CYL: prim_cyl(q, x, y, z, in, near, rho); break;
            case OP_SPHERE: prim_sphere(q, x, y, z, in, near, rho); break;
            case OP_BOX: prim_box(q, x, y, z, in, near, rho); break;
            case OP_PPED: prim_pped(q, x, y, z, in, near, rho); break;
            default: prim_gyroid(q, x, y, z, in, near, rho, margin); break;
        }
        emit_child(in, near, rho, greedy, res, acc, unc, nhit);
    }
    float val = __saturatef(acc);
    if (nhit >= 2 && fabsf(acc) < 1e-5f && res == 0.0f) unc = true;
    if (res > 0.0f) val = res;
    return val;
}

template <int SHAPE, int INTEG, bool COUNT, bool LIST, int PRIM>
#ifndef XR_FAST_MINBLOCKS
#define XR_FAST_MINBLOCKS 7  // 72 registers: the best of the 6 / 7 / 8 sweep on every bench scene
#endif
__global__ void __launch_bounds__(kBlockThreads, XR_FAST_MINBLOCKS) render_fast_kernel(const RenderParams P, const unsigned char* __restrict__ nfine_tab,
                                                                       int i_coll, int i_tess) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: [instr | f32 pool] [FastArgs] [ray64 6 x nt doubles] [queue kQueueCap x nt ints]
    const Instr* sI = reinterpret_cast<const Instr*>(smem);
    // LIST scenes (big collections) keep the fp32 pool in global memory: it would crowd out occupancy
    const float4* sF = LIST ? P.scene.f32 : reinterpret_cast<const float4*>(smem + (size_t)P.scene.n_instr * sizeof(Instr));
    FastArgs* sA = reinterpret_cast<FastArgs*>(smem + P.smem_prog_bytes);
    double* sRay = reinterpret_cast<double*>(sA + 1);
    int* queue = reinterpret_cast<int*>(sRay + 6 * kBlockThreads);
    const int tid = threadIdx.x;
    {
        uint4* dst = reinterpret_cast<uint4*>(smem);
        const uint4* srcI = reinterpret_cast<const uint4*>(P.scene.instr);
        const int nI = P.scene.n_instr * 2;
        for (int k = tid; k < nI; k += kBlockThreads) dst[k] = srcI[k];
        const uint4* srcF = reinterpret_cast<const uint4*>(P.scene.f32);
        if (!LIST)
            for (int k = tid; k < P.scene.f32_count; k += kBlockThreads) dst[nI + k] = srcF[k];
        if (P.scene.n_deform > 0 && tid < (int)(sizeof(DeformRec) / sizeof(unsigned int)))
            reinterpret_cast<unsigned int*>(&sA->d0)[tid] = reinterpret_cast<const unsigned int*>(P.scene.deform)[tid];
        if (tid == 0) {
            SceneView g;
            g.instr = P.scene.instr;
            g.f32 = P.scene.f32;
            g.f64 = P.scene.f64;
            g.grids = P.scene.grids;
            g.deform = P.scene.deform;
            g.n_instr = P.scene.n_instr;
            g.n_deform = P.scene.n_deform;
            g.vox = P.scene.vox;
            sA->gsv = g;
        }
    }

    XR_TILE_LOOP_BEGIN(P)
    int view, i, j;
    pixel_of_thread(P, xr_block_, xr_sub_, view, i, j);
    const bool valid = xr_live_ && i < P.res && j < P.res;
    if (!valid) { i = 0; j = 0; }
    int k0, k1;
    bool hit;
    float pcx, pcy, pcz, pdx, pdy, pdz;
    {
        const Ray64 ray = make_ray(P.cams[view], i, j, P.res);
        double s_in, s_out;
        hit = valid && clip_ray(ray, P.aabb_lo, P.aabb_hi, s_in, s_out);
        step_range(P, hit, s_in, s_out, INTEG == 1 ? 1 : 0, k0, k1);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            sRay[a * kBlockThreads + tid] = ray.o[a];
            sRay[(3 + a) * kBlockThreads + tid] = ray.d[a];
        }
        pcx = (float)(ray.o[0] + ray.d[0] * P.s_center);
        pcy = (float)(ray.o[1] + ray.d[1] * P.s_center);
        pcz = (float)(ray.o[2] + ray.d[2] * P.s_center);
        pdx = (float)ray.d[0];
        pdy = (float)ray.d[1];
        pdz = (float)ray.d[2];
    }
    __syncthreads();
    const int wk0 = __reduce_min_sync(FULL_MASK, hit ? k0 : 0x7fffffff);
    const int wk1 = __reduce_max_sync(FULL_MASK, hit ? k1 : 0);

    // shape constants (uniform)
    const uint4 cw0 = reinterpret_cast<const uint4*>(sI + i_coll)[0];
    const uint4 cw1 = reinterpret_cast<const uint4*>(sI + i_coll)[1];
    const bool is_coll = cw0.x == OP_COLL_BEGIN;
    const int rb = is_coll ? i_coll + 1 : i_coll;
    const int re = is_coll ? (int)cw1.z : i_coll + 1;  // skip_to = matching COLL_END
    const unsigned int cflags = is_coll ? (cw0.z | 0x100u) : 0u;
    const bool has_grid = LIST || (is_coll && (cw0.z & F_HAS_GRID));
    const float4* gF = sF + cw1.x;  // grid record (valid when has_grid)
    const float4* q1 = sF + reinterpret_cast<const uint4*>(sI + rb)[1].x;  // PRIM != 0: the one run's records
    const unsigned int n1 = reinterpret_cast<const uint4*>(sI + rb)[0].y;   //            and its child count
    const int prim_f64_idx = (int)reinterpret_cast<const uint4*>(sI + rb)[1].y;
    const unsigned long long* __restrict__ grids = P.scene.grids + cw1.w;
    // cell-list grid sub-arrays (LIST): offsets are in the 5th record word, relative to `grids`
    const uint4 lb = LIST ? *reinterpret_cast<const uint4*>(gF + 4) : make_uint4(0u, 0u, 0u, 0u);
    const unsigned int* __restrict__ l_off = reinterpret_cast<const unsigned int*>(grids + lb.x);
    const unsigned char* __restrict__ l_dist = reinterpret_cast<const unsigned char*>(grids + lb.y);
    const unsigned short* __restrict__ l_idx = reinterpret_cast<const unsigned short*>(grids + lb.z);
    const unsigned int* __restrict__ l_tab = reinterpret_cast<const unsigned int*>(grids + lb.w);
    const float4* tF = sF + (SHAPE == SHAPE_TESS ? reinterpret_cast<const uint4*>(sI + i_tess)[1].x : 0u);
    const int tess_f64_idx = SHAPE == SHAPE_TESS ? (int)reinterpret_cast<const uint4*>(sI + i_tess)[1].y : 0;
    const int n_deform = P.scene.n_deform;
    const DeformRec* __restrict__ deform = P.scene.deform;

    const float dmf = P.dm_f;
    const float dsf = P.ds_fine_f;
    // T - flat_field = sum rho*w, w = DS for plain coarse steps and ds for refined ones; Kahan
    // compensated so the fp32 sum stays ~1e-7 relative however many steps a ray takes
    const float wC = P.ds_f, wF = P.ds_fine_f;
    float accT = 0.0f, cmpT = 0.0f;
#define XR_KADD(V)                         \
    do {                                   \
        const float y_ = (V)-cmpT;         \
        const float t_ = accT + y_;        \
        cmpT = (t_ - accT) - y_;           \
        accT = t_;                         \
    } while (0)
    float prev = 0.0f;
    int qn = 0;
    unsigned int n_eval = 0, n_fine = 0, n_fallback = 0, prim_tests = 0;

    // two-mode warp-uniform march: mode 0 walks the coarse lattice, mode 1 replays queued refinements
    int k = wk0;
    int mode = 0;
    int kf = 0, jf = 0, nf = 0;  // fine state: interval, sub-step, sub-step count
    for (;;) {
        float t;
        bool act;
        if (mode == 0) {
            if (k >= wk1) {
                if (INTEG == 0 || !__any_sync(FULL_MASK, qn > 0)) break;
                mode = 1;
                nf = 0;
                jf = 0;
                continue;
            }
            act = hit && k >= k0 && k < k1;
            t = P.t_tab[k + (INTEG == 1 ? 1 : 0)];
        } else {
            act = jf < nf;
            if (!__any_sync(FULL_MASK, act)) {  // this round of intervals is done: pop the next one per lane
                const bool has = qn > 0;
                if (!__any_sync(FULL_MASK, has)) {
                    mode = 0;
                    continue;
                }
                if (has) {
                    kf = queue[(--qn) * kBlockThreads + tid];
                    nf = nfine_tab[kf];
                } else {
                    nf = 0;
                }
                jf = 0;
                continue;
            }
            t = fmaf((float)(jf + 1), dsf, P.t_tab[act ? kf : 0]);
        }

        // ---- the single evaluation site ----
        float x = fmaf(pdx, t, pcx), y = fmaf(pdy, t, pcy), z = fmaf(pdz, t, pcz);
        bool unc = false;
        if (n_deform != 0) {  // uniform
            if (n_deform == 1) Fast::deform(sA->d0, x, y, z);  // parameters come from shared memory
            else
                for (int d = 0; d < n_deform; ++d) Fast::deform(deform[d], x, y, z);
        }
        bool alive = act;
        unsigned int um_lo = ~0u, um_hi = ~0u;
        // clearance of this lane, in lattice steps, inside which density() is provably 0 (skipping):
        // 0 unless the lane sits in an empty grid cell / outside the tessellation's outer box
        float clear = 0.0f;
        unsigned int lp = 0u, le = 0u;  // LIST: this lane's child list
        float tess_limit = 3.0e38f;  // object-space distance to the nearest unit-cell face / outer-box exit
        if (SHAPE == SHAPE_TESS) {
            const float4 oc = tF[0], oh = tF[1], um = tF[2], dd = tF[3], id = tF[4];
            const float m = fmaxf(fabsf(x - oc.x) - oh.x, fmaxf(fabsf(y - oc.y) - oh.y, fabsf(z - oc.z) - oh.z));
            const float qx = (x - um.x) * id.x, qy = (y - um.y) * id.y, qz = (z - um.z) * id.z;
            const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
            float rx = qx - fx, ry = qy - fy, rz = qz - fz;
            const float lo = fminf(rx, fminf(ry, rz)), hi = fmaxf(rx, fmaxf(ry, rz));
            bool inside = m <= 0.0f;  // outer bounds inclusive (objects.go:569)
            const bool edge = alive && ((fabsf(m) < oc.w) || (inside && (lo < id.w || hi > 1.0f - id.w)));
            x = fmaf(-dd.x, fx, x);
            y = fmaf(-dd.y, fy, y);
            z = fmaf(-dd.z, fz, z);
            if (__any_sync(FULL_MASK, edge)) {  // period / bounds in doubt: exact fold, then carry on in fp32
                const float4 e = exact_fold_cold(&sA->gsv, sRay, P.s_tab, mode == 0 ? k + (INTEG == 1 ? 1 : 0) : kf,
                                                 mode == 0 ? 0 : jf + 1, P.ds_fine, tess_f64_idx);
                if (edge) {
                    x = e.x;
                    y = e.y;
                    z = e.z;
                    inside = e.w != 0.0f;
                    rx = (x - um.x) * id.x;
                    ry = (y - um.y) * id.y;
                    rz = (z - um.z) * id.z;
                    if (COUNT && (P.dbg_cause == 0 || (P.dbg_cause & (fabsf(m) < oc.w ? 1 : 2)))) ++n_fallback;
                }
            }
            alive = alive && inside;
            if (has_grid) {  // the grid spans exactly the unit cell: cell = floor(fraction * g), fraction in [0,1]
                const float4 gd = gF[2], gf = gF[3];
                const int gx = __float_as_int(gd.x), gy = __float_as_int(gd.y), gz = __float_as_int(gd.z);
                const int ix = min(gx - 1, (int)(rx * gf.x));
                const int iy = min(gy - 1, (int)(ry * gf.y));
                const int iz = min(gz - 1, (int)(rz * gf.z));
                const unsigned int cell = (unsigned int)((iz * gy + iy) * gx + ix);
                if (LIST) {
                    if (alive) {
                        lp = __ldg(l_off + cell);
                        le = __ldg(l_off + cell + 1);
                        if (lp == le) clear = (float)((unsigned int)__ldg(l_dist + cell) - 1u) * (gF[0].w * P.skip_m2s);
                    }
                    um_lo = __any_sync(FULL_MASK, lp < le) ? 1u : 0u;
                    um_hi = 0u;
                } else {
                    uint2 mk = make_uint2(0u, 0u);
                    if (alive) mk = __ldg(reinterpret_cast<const uint2*>(grids) + cell);
                    if (mk.y >> 31) {  // empty cell: low byte = Chebyshev distance (cells) to the nearest occupied cell
                        clear = (float)((mk.x & 255u) - 1u) * (gF[0].w * P.skip_m2s);
                        mk = make_uint2(0u, 0u);
                    }
                    um_lo = __reduce_or_sync(FULL_MASK, mk.x);
                    um_hi = __reduce_or_sync(FULL_MASK, mk.y);
                }
            }
            // outside the outer box (Chebyshev distance m > 0) nothing can be hit for m / (ds * lip) steps
            if (act && !inside) clear = fmaxf(m - 4.0f * oc.w, 0.0f) * P.skip_m2s;
            // a skip justified by a child's own margin must stay inside this period and inside the outer box
            if (!has_grid)
                tess_limit = fminf(fminf(fminf(rx, 1.0f - rx) * fabsf(dd.x), fminf(ry, 1.0f - ry) * fabsf(dd.y)),
                                   fminf(fminf(rz, 1.0f - rz) * fabsf(dd.z), -m));
        } else if (has_grid) {
            const float4 g0 = gF[0], g1 = gF[1], gd = gF[2];
            const int gx = __float_as_int(gd.x), gy = __float_as_int(gd.y), gz = __float_as_int(gd.z);
            const int ix = min(gx - 1, max(0, __float2int_rd((x - g0.x) * g1.x)));
            const int iy = min(gy - 1, max(0, __float2int_rd((y - g0.y) * g1.y)));
            const int iz = min(gz - 1, max(0, __float2int_rd((z - g0.z) * g1.z)));
            const unsigned int cell = (unsigned int)((iz * gy + iy) * gx + ix);
            if (LIST) {
                if (alive) {
                    lp = __ldg(l_off + cell);
                    le = __ldg(l_off + cell + 1);
                    if (lp == le) clear = (float)((unsigned int)__ldg(l_dist + cell) - 1u) * (g0.w * P.skip_m2s);
                }
                um_lo = __any_sync(FULL_MASK, lp < le) ? 1u : 0u;
                um_hi = 0u;
            } else {
                uint2 mk = make_uint2(0u, 0u);
                if (alive) mk = __ldg(reinterpret_cast<const uint2*>(grids) + cell);
                if (mk.y >> 31) {
                    clear = (float)((mk.x & 255u) - 1u) * (g0.w * P.skip_m2s);
                    mk = make_uint2(0u, 0u);
                }
                um_lo = __reduce_or_sync(FULL_MASK, mk.x);
                um_hi = __reduce_or_sync(FULL_MASK, mk.y);
            }
        }
        float rho = 0.0f;
        float clr = 3.0e38f;
        const bool evaluated = (um_lo | um_hi) != 0u && __any_sync(FULL_MASK, alive);
        if (evaluated) {
            if (LIST) rho = eval_list<COUNT>(sF, l_tab, l_idx, lp, le, cflags, x, y, z, alive, unc, prim_tests);
            else if (PRIM != 0) rho = eval_single<PRIM, COUNT>(q1, n1, cflags, has_grid, um_lo, um_hi, x, y, z, alive, unc, prim_tests, clr);
            else rho = eval_runs<COUNT>(sI, sF, rb, re, cflags, has_grid, um_lo, um_hi, x, y, z, alive, unc, prim_tests, clr);
        }
        if (!has_grid) {  // uniform: scenes with a grid never pay for the margin arithmetic
            if (evaluated && alive) clear = fmaxf(fminf(clr, tess_limit) - 1.0e-5f, 0.0f) * P.skip_m2s;
        }
        rho *= dmf;
        unc = unc && act;
        if (__any_sync(FULL_MASK, unc)) {
            float r;
            if (PRIM != 0 && n1 == 1u)
                r = exact_single_cold<PRIM>(P.scene.f64, P.scene.deform, n_deform, sRay, P.s_tab, mode == 0 ? k + (INTEG == 1 ? 1 : 0) : kf,
                                            mode == 0 ? 0 : jf + 1, P.ds_fine, P.dm, SHAPE == SHAPE_TESS ? tess_f64_idx : -1, prim_f64_idx,
                                            cflags);
            else
                r = exact_density_cold(&sA->gsv, sRay, P.s_tab, mode == 0 ? k + (INTEG == 1 ? 1 : 0) : kf, mode == 0 ? 0 : jf + 1,
                                       P.ds_fine, P.dm, unc);
            if (unc) {
                rho = r;
                if (COUNT && (P.dbg_cause == 0 || (P.dbg_cause & 4))) ++n_fallback;
            }
        }

        // ---- bookkeeping ----
        if (mode == 0) {
            if (act) {
                if (COUNT) ++n_eval;
                float w = wC;
                if (INTEG == 1 && ((rho == 0.0f) != (prev == 0.0f))) {
                    queue[qn * kBlockThreads + tid] = k;
                    ++qn;
                    w = wF;  // T += rho*ds (main.go:188) instead of rho*DS (main.go:190)
                }
                XR_KADD(rho * w);
                prev = rho;
            }
            // Exact empty-space skipping: when no lane of the warp can meet a non-zero density within the
            // next n steps (empty grid cells with clearance / still outside the outer box / not yet inside
            // its clipped range), those steps would add rho*w = +0 and cannot flip (rho==0)!=(prev==0).
            // The same holds inside a solid whose margin is known (gyroid |g|-t with its gradient bound):
            // the skipped steps all return this step's rho, so they add (n-1)*rho*DS and flip nothing.
            int adv = 1;
            if ((um_lo | um_hi) == 0u || !has_grid) {
                float a = (((um_lo | um_hi) == 0u && rho != 0.0f) || unc) ? 0.0f : clear;
                if (!act) a = (!hit || k >= k1) ? 1.0e6f : (float)(k0 - k);
                else a = fminf(a, (float)(k1 - k));  // never run past the lane's last lattice sample (unbounded solids)
                const int n = __reduce_min_sync(FULL_MASK, (int)fminf(a, 1.0e6f));
                if (n >= 2) {  // steps k+1 .. k+n-1 are skipped, k+n is evaluated again
                    adv = n;
                    if (act && rho != 0.0f) XR_KADD(rho * wC * (float)(n - 1));
                }
            }
            k += adv;
            if (INTEG == 1 && __any_sync(FULL_MASK, qn == kQueueCap)) {
                mode = 1;
                nf = 0;
                jf = 0;
            }
        } else {
            if (act) {
                XR_KADD(rho * wF);
                if (COUNT) {
                    ++n_eval;
                    ++n_fine;
                }
            }
            ++jf;
        }
    }
    const double T = P.flat_field + ((double)accT - (double)cmpT);
    store_pixel(P, view, i, j, valid, exp(-T));
    if (COUNT) {
        // the count of fine steps is ray independent given the refined intervals; recount exactly
        add_stats(P, valid ? (unsigned long long)P.n_steps + n_fine : 0ull, n_eval, n_fallback, prim_tests, valid ? 1ull : 0ull);
    }
    XR_TILE_LOOP_END
}


// =========================================================================================
// Lane-asynchronous march for scenes that are ONE primitive (a gyroid unit cell, a pillar array):
// every lane walks its own lattice index, so a lane never waits at another lane's surface crossing.
// With the warp-synchronous loop above a warp evaluates the UNION of its lanes' critical samples
// (gyroid + sigmoid config: 340 iterations per warp); here it evaluates the MAXIMUM over lanes
// (~70, simulation in DESIGN.md section 5).  Nothing is shared between lanes of a one-primitive scene
// (no candidate masks to merge), so lockstep buys nothing there.  Refinements are not queued: a lane
// that sees (rho==0) != (prev==0) replays the fine sub-steps of that interval at once, then resumes
// its coarse walk (same terms as main.go:176-196, summed in a different order; Kahan fp32).
// =========================================================================================

// Warp stage 0 (the only one these scenes may have for the second-order rule) applied to the sample, plus
// e = J d, the image of the ray direction under the stage's Jacobian (rigid / linear / affine / sigmoid).
__device__ __forceinline__ void deform_dir(const DeformRec& r, float& x, float& y, float& z, float dx, float dy, float dz, float& ex,
                                           float& ey, float& ez) {
    const float* p = r.f;
    ex = dx;
    ey = dy;
    ez = dz;
    switch (r.type) {  // uniform
        case D_AFFINE: {
            const float nx = p[0] * x + p[1] * y + p[2] * z, ny = p[3] * x + p[4] * y + p[5] * z, nz = p[6] * x + p[7] * y + p[8] * z;
            ex = p[0] * dx + p[1] * dy + p[2] * dz;
            ey = p[3] * dx + p[4] * dy + p[5] * dz;
            ez = p[6] * dx + p[7] * dy + p[8] * dz;
            x = nx; y = ny; z = nz;
            break;
        }
        case D_LINEAR: {
            const float nx = x + p[0] * x + p[5] * y + p[4] * z, ny = y + p[5] * x + p[1] * y + p[3] * z,
                        nz = z + p[4] * x + p[3] * y + p[2] * z;
            ex = dx + p[0] * dx + p[5] * dy + p[4] * dz;
            ey = dy + p[5] * dx + p[1] * dy + p[3] * dz;
            ez = dz + p[4] * dx + p[3] * dy + p[2] * dz;
            x = nx; y = ny; z = nz;
            break;
        }
        case D_SIGMOID: {
            float q = r.axis == 0 ? x : (r.axis == 1 ? y : z);
            const float E = __expf((q - p[1]) * p[2]);  // p[2] = -1/L
            const float sg = __fdividef(1.0f, 1.0f + E);
            q += __fdividef(p[0], 1.0f + E);            // same expression as Fast::deform (eps_pos budget)
            // 1 + (A/L) sigma (1 - sigma); written so that E = inf (sigma = 0) gives 1, not inf * 0
            const float jac = fmaf(-p[0] * p[2], sg * (1.0f - sg), 1.0f);
            if (r.axis == 0) { x = q; ex = dx * jac; }
            else if (r.axis == 1) { y = q; ey = dy * jac; }
            else { z = q; ez = dz * jac; }
            break;
        }
        default: Fast::deform(r, x, y, z); break;  // rigid: e = d; anything else never uses e (M2 = 0)
    }
}

// Gyroid test with the second-order skip bound: the largest delta (distance along the ray) with
// G1 delta + M2 delta^2 / 2 <= m, G1 >= |h'(t)| from the local gradient, m = distance of |g| - thickness from the
// guard band.  Inside [t, t + delta] the exact classification cannot change.  margin is returned in the units the
// march converts with skip_m2s (object-space max-norm distance = delta * Lipschitz factor).
__device__ __forceinline__ void prim_gyroid_so(const float4* __restrict__ q, float x, float y, float z, float ex, float ey, float ez,
                                               bool& in, bool& near, float& rho, float& margin) {
    const float4 a = q[0], b = q[1], c = q[2];
    float sx, cx, sy, cy, sz, cz;
    fast_sincos((x - a.x) * b.x, &sx, &cx);
    fast_sincos((y - a.y) * b.x, &sy, &cy);
    fast_sincos((z - a.z) * b.x, &sz, &cz);
    const float t = fabsf(sx * cy + sy * cz + sz * cx) - b.y;
    in = t < 0.0f;
    near = fabsf(t) < b.z;
    rho = a.w;
    const float m = fmaxf(fabsf(t) - b.z, 0.0f);
    if (c.x > 0.0f) {  // uniform
        const float gx = cx * cy - sz * sx, gy = cy * cz - sx * sy, gz = cz * cx - sy * sz;
        const float G1 = (fabsf(gx * ex + gy * ey + gz * ez) + c.y) * fabsf(b.x);
        const float delta = __fdividef(2.0f * m, G1 + sqrtf(fmaf(2.0f * c.x, m, G1 * G1)));
        margin = delta * c.z;
    } else {
        margin = m * b.w;
    }
}

constexpr int kPendCap = 8;  // deferred guard-band sub-steps per lane; a ninth is settled at once

template <int SHAPE, int INTEG, bool COUNT, int PRIM>
#ifndef XR_ASYNC_MINBLOCKS
#define XR_ASYNC_MINBLOCKS 7
#endif
__global__ void __launch_bounds__(kBlockThreads, XR_ASYNC_MINBLOCKS) render_async_kernel(const RenderParams P, const unsigned char* __restrict__ nfine_tab,
                                                                         int i_coll, int i_tess) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: [instr | f32 pool] [FastArgs] [ray64 6 x nt doubles] [deferred guard-band sub-steps kPendCap x nt ints]
    const Instr* sI = reinterpret_cast<const Instr*>(smem);
    const float4* sF = reinterpret_cast<const float4*>(smem + (size_t)P.scene.n_instr * sizeof(Instr));
    FastArgs* sA = reinterpret_cast<FastArgs*>(smem + P.smem_prog_bytes);
    double* sRay = reinterpret_cast<double*>(sA + 1);
    int* pend = reinterpret_cast<int*>(sRay + 6 * kBlockThreads);
    const int tid = threadIdx.x;
    {
        uint4* dst = reinterpret_cast<uint4*>(smem);
        const uint4* srcI = reinterpret_cast<const uint4*>(P.scene.instr);
        const int nI = P.scene.n_instr * 2;
        for (int k = tid; k < nI; k += kBlockThreads) dst[k] = srcI[k];
        const uint4* srcF = reinterpret_cast<const uint4*>(P.scene.f32);
        for (int k = tid; k < P.scene.f32_count; k += kBlockThreads) dst[nI + k] = srcF[k];
        if (P.scene.n_deform > 0 && tid < (int)(sizeof(DeformRec) / sizeof(unsigned int)))
            reinterpret_cast<unsigned int*>(&sA->d0)[tid] = reinterpret_cast<const unsigned int*>(P.scene.deform)[tid];
        if (tid == 0) {
            SceneView g;
            g.instr = P.scene.instr;
            g.f32 = P.scene.f32;
            g.f64 = P.scene.f64;
            g.grids = P.scene.grids;
            g.deform = P.scene.deform;
            g.n_instr = P.scene.n_instr;
            g.n_deform = P.scene.n_deform;
            g.vox = P.scene.vox;
            sA->gsv = g;
        }
    }

    XR_TILE_LOOP_BEGIN(P)
    int view, i, j;
    pixel_of_thread(P, xr_block_, xr_sub_, view, i, j);
    const bool valid = xr_live_ && i < P.res && j < P.res;
    if (!valid) { i = 0; j = 0; }
    int k, k1;
    float pcx, pcy, pcz, pdx, pdy, pdz;
    {
        const Ray64 ray = make_ray(P.cams[view], i, j, P.res);
        double s_in, s_out;
        const bool hit = valid && clip_ray(ray, P.aabb_lo, P.aabb_hi, s_in, s_out);
        step_range(P, hit, s_in, s_out, INTEG == 1 ? 1 : 0, k, k1);  // !hit: k = k1 = 0, the lane never becomes active
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            sRay[a * kBlockThreads + tid] = ray.o[a];
            sRay[(3 + a) * kBlockThreads + tid] = ray.d[a];
        }
        pcx = (float)(ray.o[0] + ray.d[0] * P.s_center);
        pcy = (float)(ray.o[1] + ray.d[1] * P.s_center);
        pcz = (float)(ray.o[2] + ray.d[2] * P.s_center);
        pdx = (float)ray.d[0];
        pdy = (float)ray.d[1];
        pdz = (float)ray.d[2];
    }
    __syncthreads();

    // shape constants (uniform)
    const uint4 cw0 = reinterpret_cast<const uint4*>(sI + i_coll)[0];
    const uint4 cw1 = reinterpret_cast<const uint4*>(sI + i_coll)[1];
    const bool is_coll = cw0.x == OP_COLL_BEGIN;
    const int rb = is_coll ? i_coll + 1 : i_coll;
    const unsigned int cflags = is_coll ? (cw0.z | 0x100u) : 0u;
    const bool has_grid = is_coll && (cw0.z & F_HAS_GRID);
    const float4* gF = sF + cw1.x;  // grid record (valid when has_grid)
    const float4* q1 = sF + reinterpret_cast<const uint4*>(sI + rb)[1].x;  // the primitive's records
    const int prim_f64_idx = (int)reinterpret_cast<const uint4*>(sI + rb)[1].y;
    const unsigned long long* __restrict__ grids = P.scene.grids + cw1.w;
    const float4* tF = sF + (SHAPE == SHAPE_TESS ? reinterpret_cast<const uint4*>(sI + i_tess)[1].x : 0u);
    const int tess_f64_idx = SHAPE == SHAPE_TESS ? (int)reinterpret_cast<const uint4*>(sI + i_tess)[1].y : 0;
    const int n_deform = P.scene.n_deform;
    const DeformRec* __restrict__ deform = P.scene.deform;
    const bool clamps = (cflags & 0x100u) != 0u;  // a collection clamps its sum (objects.go:431-436) ...
    const bool greedy = (cflags & F_GREEDY) != 0u;  // ... unless greedy returned the first positive child (:425-427)

    const float dmf = P.dm_f, dsf = P.ds_fine_f, wC = P.ds_f, wF = P.ds_fine_f;
    float accT = 0.0f, cmpT = 0.0f;
    float prev = 0.0f;
    int jf = 0, nf = 0, kf = 0;  // fine replay of interval kf: sub-step jf of nf
    // Gyroids only (compile time): their band decides 1.4 samples per ray and their clearance is known everywhere; the
    // other primitives hardly ever meet a guard band, and the bookkeeping costs the pillar array 5 %.
    constexpr bool kDefer = INTEG == 1 && PRIM == (int)OP_GYROID;
    int np = 0;                  // guard-band fine sub-steps whose fp64 re-evaluation is deferred to the end
    const float fine_per_coarse = (float)(P.ds / P.ds_fine);  // 10 (main.go:178), a hair less if the division rounds down
    unsigned int n_eval = 0, n_fine = 0, n_fallback = 0, prim_tests = 0;

    while (__any_sync(FULL_MASK, jf < nf || k < k1)) {
        const bool fine = jf < nf;
        const bool act = fine || k < k1;
        const int base = act ? (fine ? kf : k + (INTEG == 1 ? 1 : 0)) : 0;
        const int nsub = fine ? jf + 1 : 0;
        float t = __ldg(P.t_tab + base);
        if (fine) t = fmaf((float)nsub, dsf, t);

        // ---- the evaluation site (same arithmetic as the warp-synchronous kernel) ----
        float x = fmaf(pdx, t, pcx), y = fmaf(pdy, t, pcy), z = fmaf(pdz, t, pcz);
        float ex = pdx, ey = pdy, ez = pdz;
        if (n_deform != 0) {  // uniform
            if (n_deform == 1) deform_dir(sA->d0, x, y, z, pdx, pdy, pdz, ex, ey, ez);
            else
                for (int d = 0; d < n_deform; ++d) Fast::deform(deform[d], x, y, z);
        }
        bool alive = act;
        bool occupied = true;   // the lane's grid cell lists the primitive (always true without a grid)
        float clear = 0.0f;     // lattice steps inside which density() provably keeps this sample's value
        float tess_limit = 3.0e38f;  // object-space distance to the nearest unit-cell face / outer-box exit
        if (SHAPE == SHAPE_TESS) {
            const float4 oc = tF[0], oh = tF[1], um = tF[2], dd = tF[3], id = tF[4];
            const float m = fmaxf(fabsf(x - oc.x) - oh.x, fmaxf(fabsf(y - oc.y) - oh.y, fabsf(z - oc.z) - oh.z));
            const float qx = (x - um.x) * id.x, qy = (y - um.y) * id.y, qz = (z - um.z) * id.z;
            const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
            float rx = qx - fx, ry = qy - fy, rz = qz - fz;
            const float lo = fminf(rx, fminf(ry, rz)), hi = fmaxf(rx, fmaxf(ry, rz));
            bool inside = m <= 0.0f;  // outer bounds inclusive (objects.go:569)
            const bool edge = alive && ((fabsf(m) < oc.w) || (inside && (lo < id.w || hi > 1.0f - id.w)));
            x = fmaf(-dd.x, fx, x);
            y = fmaf(-dd.y, fy, y);
            z = fmaf(-dd.z, fz, z);
            if (__any_sync(FULL_MASK, edge)) {  // period / bounds in doubt: exact fold, then carry on in fp32
                const float4 e = exact_fold_cold(&sA->gsv, sRay, P.s_tab, base, nsub, P.ds_fine, tess_f64_idx);
                if (edge) {
                    x = e.x;
                    y = e.y;
                    z = e.z;
                    inside = e.w != 0.0f;
                    rx = (x - um.x) * id.x;
                    ry = (y - um.y) * id.y;
                    rz = (z - um.z) * id.z;
                    if (COUNT && (P.dbg_cause == 0 || (P.dbg_cause & (fabsf(m) < oc.w ? 1 : 2)))) ++n_fallback;
                }
            }
            alive = alive && inside;
            if (has_grid) {  // the grid spans exactly the unit cell: cell = floor(fraction * g), fraction in [0,1]
                const float4 gd = gF[2], gf = gF[3];
                const int gx = __float_as_int(gd.x), gy = __float_as_int(gd.y), gz = __float_as_int(gd.z);
                const int ix = min(gx - 1, (int)(rx * gf.x));
                const int iy = min(gy - 1, (int)(ry * gf.y));
                const int iz = min(gz - 1, (int)(rz * gf.z));
                const unsigned int cell = (unsigned int)((iz * gy + iy) * gx + ix);
                uint2 mk = make_uint2(0u, 0u);
                if (alive) mk = __ldg(reinterpret_cast<const uint2*>(grids) + cell);
                if (mk.y >> 31) {  // empty cell: low byte = Chebyshev distance (cells) to the nearest occupied cell
                    clear = (float)((mk.x & 255u) - 1u) * (gF[0].w * P.skip_m2s);
                    mk = make_uint2(0u, 0u);
                }
                occupied = (mk.x | mk.y) != 0u;
            }
            // outside the outer box (Chebyshev distance m > 0) nothing can be hit for m / (ds * lip) steps
            if (act && !inside) clear = fmaxf(m - 4.0f * oc.w, 0.0f) * P.skip_m2s;
            // a skip justified by the primitive's own margin must stay inside this period and inside the outer box
            if (!has_grid)
                tess_limit = fminf(fminf(fminf(rx, 1.0f - rx) * fabsf(dd.x), fminf(ry, 1.0f - ry) * fabsf(dd.y)),
                                   fminf(fminf(rz, 1.0f - rz) * fabsf(dd.z), -m));
        } else if (has_grid) {
            const float4 g0 = gF[0], g1 = gF[1], gd = gF[2];
            const int gx = __float_as_int(gd.x), gy = __float_as_int(gd.y), gz = __float_as_int(gd.z);
            const int ix = min(gx - 1, max(0, __float2int_rd((x - g0.x) * g1.x)));
            const int iy = min(gy - 1, max(0, __float2int_rd((y - g0.y) * g1.y)));
            const int iz = min(gz - 1, max(0, __float2int_rd((z - g0.z) * g1.z)));
            const unsigned int cell = (unsigned int)((iz * gy + iy) * gx + ix);
            uint2 mk = make_uint2(0u, 0u);
            if (alive) mk = __ldg(reinterpret_cast<const uint2*>(grids) + cell);
            if (mk.y >> 31) {
                clear = (float)((mk.x & 255u) - 1u) * (g0.w * P.skip_m2s);
                mk = make_uint2(0u, 0u);
            }
            occupied = (mk.x | mk.y) != 0u;
        }
        const bool test = alive && occupied;
        float rho = 0.0f;
        bool unc = false;
        if (__any_sync(FULL_MASK, test)) {
            bool in, near;
            float pr, clr = 0.0f;
            if (PRIM == OP_GYROID) prim_gyroid_so(q1, x, y, z, ex, ey, ez, in, near, pr, clr);
            else if (PRIM == OP_SPHERE) prim_sphere(q1, x, y, z, in, near, pr);
            else if (PRIM == OP_BOX) prim_box(q1, x, y, z, in, near, pr);
            else prim_cyl(q1, x, y, z, in, near, pr);
            if (COUNT) prim_tests += test ? 1u : 0u;
            unc = near && test;
            rho = (in && test) ? pr : 0.0f;
            if (clamps && !(greedy && pr > 0.0f)) rho = __saturatef(rho);
            // the gyroid's own clearance, never across a unit-cell face or out of the outer box.  (Distance-to-surface
            // clearances for spheres / boxes / cylinders were measured and rejected: on the pillar array they remove 3 % of
            // the evaluations and cost 30 % in extra arithmetic per sample; grid scenes skip through empty cells only.)
            if (!has_grid && test) clear = fmaxf(fminf(clr, tess_limit) - 1.0e-5f, 0.0f) * P.skip_m2s;
        }
        rho *= dmf;
        // A guard-band sample of a REFINED sub-step feeds nothing but the sum (main.go:186: T += rho*ds), so its fp64
        // re-evaluation can wait: note (interval, sub-step), add nothing now, and settle all of them after the march,
        // where one call of the cold function serves every lane of the warp that has one left (max over lanes ~5 calls
        // instead of one call per band sample of any lane: 22 per warp on the gyroid + sigmoid config, each for 2.7 lanes).
        if (kDefer && unc && fine && np < kPendCap) {
            pend[np * kBlockThreads + tid] = (kf << 4) | nsub;
            ++np;
            rho = 0.0f;
            unc = false;
        }
        // Guard-band samples of the coarse lattice decide refinements: the fp64 reference answers at once.  (Parking such lanes until several of a warp wait and
        // serving them with one call was measured: 4x fewer calls, but the parked lanes stretch the warp's critical path --
        // gyroid +1 %, pillar array -24 % at a threshold of 4 lanes.)
        if (__any_sync(FULL_MASK, unc)) {
            const float r = exact_single_cold<PRIM>(P.scene.f64, P.scene.deform, n_deform, sRay, P.s_tab, base, nsub, P.ds_fine, P.dm,
                                                    SHAPE == SHAPE_TESS ? tess_f64_idx : -1, prim_f64_idx, cflags);
            if (unc) {
                rho = r;
                if (COUNT && (P.dbg_cause == 0 || (P.dbg_cause & 4))) ++n_fallback;
            }
        }

        // ---- per-lane bookkeeping ----
        if (fine) {
            // the clearance holds for refined sub-steps as well: `clear` coarse steps are clear * (DS / ds) sub-steps, and
            // the skipped ones return this sub-step's rho (a gyroid wall is crossed in 3-4 evaluations instead of 9-10)
            int adv = 1;
            if (kDefer && !has_grid) {  // (grid scenes have no clearance next to a surface)
                const float a = fminf(unc ? 0.0f : clear * fine_per_coarse, (float)(nf - jf));
                const int n = (int)a;
                if (n >= 2) {
                    adv = n;
                    if (rho != 0.0f) XR_KADD(rho * wF * (float)(n - 1));
                }
            }
            XR_KADD(rho * wF);
            jf += adv;
            if (COUNT) {
                ++n_eval;
                n_fine += (unsigned int)adv;
            }
        } else if (act) {
            if (COUNT) ++n_eval;
            float w = wC;
            if (INTEG == 1 && ((rho == 0.0f) != (prev == 0.0f))) {  // main.go:181: refine (left, right], right = this sample
                kf = k;
                nf = nfine_tab[k];
                jf = 0;
                w = wF;  // T += rho*ds (main.go:188) instead of rho*DS (main.go:190)
            }
            XR_KADD(rho * w);
            prev = rho;
            // Exact skipping, per lane: the next n - 1 lattice samples provably return this sample's rho (empty grid
            // cells with clearance, outside the outer box, or the primitive's own margin), so they add
            // (n-1)*rho*DS and flip nothing.
            int adv = 1;
            if (!occupied || !has_grid) {
                float a = unc ? 0.0f : clear;
                a = fminf(a, (float)(k1 - k));  // never past the lane's last lattice sample
                const int n = (int)a;
                if (n >= 2) {
                    adv = n;
                    if (rho != 0.0f) XR_KADD(rho * wC * (float)(n - 1));
                }
            }
            k += adv;
        }
    }
    if (kDefer) {
        while (__any_sync(FULL_MASK, np > 0)) {  // settle the deferred sub-steps, newest first
            const bool has = np > 0;
            const int ent = has ? pend[(np - 1) * kBlockThreads + tid] : 0;
            const float r = exact_single_cold<PRIM>(P.scene.f64, P.scene.deform, n_deform, sRay, P.s_tab, ent >> 4, ent & 15, P.ds_fine, P.dm,
                                                    SHAPE == SHAPE_TESS ? tess_f64_idx : -1, prim_f64_idx, cflags);
            if (has) {
                --np;
                XR_KADD(r * wF);
                if (COUNT && (P.dbg_cause == 0 || (P.dbg_cause & 4))) ++n_fallback;
            }
        }
    }
    const double T = P.flat_field + ((double)accT - (double)cmpT);
    store_pixel(P, view, i, j, valid, exp(-T));
    if (COUNT) add_stats(P, valid ? (unsigned long long)P.n_steps + n_fine : 0ull, n_eval, n_fallback, prim_tests, valid ? 1ull : 0ull);
    XR_TILE_LOOP_END
}

size_t async_kernel_smem_bytes(const RenderParams& P) {
    return (size_t)P.smem_prog_bytes + sizeof(FastArgs) + 6 * kBlockThreads * sizeof(double) + (size_t)kPendCap * kBlockThreads * sizeof(int);
}

template <int SHAPE, int INTEG, bool COUNT, int PRIM>
static cudaError_t launch_async_one(const RenderParams& P, const unsigned char* nfine, int i_coll, int i_tess, unsigned int grid,
                                    cudaStream_t stream) {
    auto kern = render_async_kernel<SHAPE, INTEG, COUNT, PRIM>;
    const size_t smem = async_kernel_smem_bytes(P);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kBlockThreads, smem, stream>>>(P, nfine, i_coll, i_tess);
    return cudaGetLastError();
}

// One-primitive scenes (prim = OP_CYL / OP_GYROID / OP_SPHERE / OP_BOX, run length 1, no cell-list grid).
cudaError_t launch_render_async(const RenderParams& P, int shape, int integrator, bool count, int prim, const unsigned char* d_nfine,
                                int i_coll, int i_tess, cudaStream_t stream) {
    unsigned int grid = (unsigned int)((size_t)P.n_views * P.tiles_i * P.tiles_j);
    if (grid == 0) return cudaSuccess;
    if (P.tile_list) grid = grid < kTileListGrid ? grid : kTileListGrid;
#define XR_AGO(S, I, C, Q) return launch_async_one<S, I, C, Q>(P, d_nfine, i_coll, i_tess, grid, stream)
#define XR_APRIM(S, I, C)                                          \
    do {                                                           \
        if (prim == (int)OP_CYL) XR_AGO(S, I, C, (int)OP_CYL);     \
        if (prim == (int)OP_GYROID) XR_AGO(S, I, C, (int)OP_GYROID); \
        if (prim == (int)OP_SPHERE) XR_AGO(S, I, C, (int)OP_SPHERE); \
        if (prim == (int)OP_BOX) XR_AGO(S, I, C, (int)OP_BOX);     \
    } while (0)
#define XR_ACOUNT(S, I)                  \
    do {                                 \
        if (count) XR_APRIM(S, I, true); \
        else XR_APRIM(S, I, false);      \
    } while (0)
    if (shape == SHAPE_FLAT) {
        if (integrator == 0) XR_ACOUNT(SHAPE_FLAT, 0);
        else XR_ACOUNT(SHAPE_FLAT, 1);
    } else {
        if (integrator == 0) XR_ACOUNT(SHAPE_TESS, 0);
        else XR_ACOUNT(SHAPE_TESS, 1);
    }
#undef XR_ACOUNT
#undef XR_APRIM
#undef XR_AGO
    return cudaErrorInvalidValue;
}

size_t fast_kernel_smem_bytes(const RenderParams& P) {
    return (size_t)P.smem_prog_bytes + sizeof(FastArgs) + 6 * kBlockThreads * sizeof(double) + (size_t)kQueueCap * kBlockThreads * sizeof(int);
}

template <int SHAPE, int INTEG, bool COUNT, bool LIST, int PRIM>
static cudaError_t launch_one(const RenderParams& P, const unsigned char* nfine, int i_coll, int i_tess, size_t smem, unsigned int grid,
                              cudaStream_t stream) {
    auto kern = render_fast_kernel<SHAPE, INTEG, COUNT, LIST, PRIM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kBlockThreads, smem, stream>>>(P, nfine, i_coll, i_tess);
    return cudaGetLastError();
}

// shape: SHAPE_FLAT / SHAPE_TESS; i_coll: index of the COLL_BEGIN (or of the lone primitive run);
// i_tess: index of the TESS_BEGIN.
// prim: OP_CYL / OP_GYROID / OP_SPHERE / OP_BOX when the collection is exactly one run of that type (specialised variants), else 0.
cudaError_t launch_render_fast(const RenderParams& P, int shape, int integrator, bool count, bool list, int prim,
                               const unsigned char* d_nfine, int i_coll, int i_tess, cudaStream_t stream) {
    const size_t smem = fast_kernel_smem_bytes(P);
    unsigned int grid = (unsigned int)((size_t)P.n_views * P.tiles_i * P.tiles_j);
    if (grid == 0) return cudaSuccess;
    if (P.tile_list) grid = grid < kTileListGrid ? grid : kTileListGrid;
#define XR_GO(S, I, C, L, Q) return launch_one<S, I, C, L, Q>(P, d_nfine, i_coll, i_tess, smem, grid, stream)
#define XR_PICK2(S, I, C)                                              \
    do {                                                               \
        if (list) XR_GO(S, I, C, true, 0);                             \
        else if (prim == (int)OP_CYL) XR_GO(S, I, C, false, (int)OP_CYL);       \
        else if (prim == (int)OP_GYROID) XR_GO(S, I, C, false, (int)OP_GYROID); \
        else if (prim == (int)OP_SPHERE) XR_GO(S, I, C, false, (int)OP_SPHERE); \
        else if (prim == (int)OP_BOX) XR_GO(S, I, C, false, (int)OP_BOX);       \
        else XR_GO(S, I, C, false, 0);                                 \
    } while (0)
#define XR_PICK(S, I)                      \
    do {                                   \
        if (count) XR_PICK2(S, I, true);   \
        else XR_PICK2(S, I, false);        \
    } while (0)
    if (shape == SHAPE_FLAT) {
        if (integrator == 0) XR_PICK(SHAPE_FLAT, 0);
        else XR_PICK(SHAPE_FLAT, 1);
    } else {
        if (integrator == 0) XR_PICK(SHAPE_TESS, 0);
        else XR_PICK(SHAPE_TESS, 1);
    }
#undef XR_PICK
#undef XR_PICK2
#undef XR_GO
    return cudaErrorInvalidValue;
}

}  // namespace xr
