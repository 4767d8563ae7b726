// render_span.cu -- the interval ("span") renderer for scenes made of CONVEX primitives.
//
// The reference marches every ray sample by sample (main.go:144-199) and asks the object tree at each one.  For a
// sphere, box, parallelepiped or cylinder (objects.go:63-72, 171-179, 247-255, 334-350) the set of ray parameters
// inside the primitive is ONE interval, also under an affine warp (deformations.go:87-92, 136-141, 173-175) and also per
// period of a tessellation (objects.go:568-582: inside one period the fold is a translation).  So the whole march
// collapses to
//   1. walking the ray through a coarse candidate grid (fp32 DDA; per cell a 64-bit set of children) and
//      pre-filtering the candidates with a conservative fp32 test,
//   2. intersecting the ray with each surviving (period, child) in fp64 and mapping the interval end points to ORDINALS
//      of the reference's sample lattice (the same host-built fp64 tables the marching kernels index: s_tab, nfine),
//   3. one integer sweep over the sorted end points that reproduces integrate_along_ray / integrate_hierarchical term
//      by term: which coarse samples see which density, which coarse intervals flip (rho == 0) != (prev_rho == 0) and
//      are refined, and what their fine sub-steps add.  Densities of a piece come from the set of children covering
//      it, combined in child order in fp64 exactly as ObjectCollection.Density does (objects.go:422-438).
// No sample is ever classified in fp32: the only inexact step is the fp64 root of a quadratic, accurate to ~1e-15, and
// a lattice sample closer than 1e-11 (plus the conditioning of the root) to ANY end point, cell face or bound raises
// `doubt` for the ray.  Tiles with a doubtful ray (and rays that overflow the small per-ray lists) are flagged and
// re-rendered by the marching kernels, which settle such samples with the reference's own operation order.  The
// result therefore equals the marching kernels' to ~1e-15 in T, in both precision modes.
#include "eval.cuh"

namespace xr {

#ifndef XR_SPAN_MINBLOCKS
#define XR_SPAN_MINBLOCKS 4
#endif
#ifndef XR_SPAN_PERSISTENT
#define XR_SPAN_PERSISTENT 0
#endif
#if XR_SPAN_PERSISTENT
#define XR_SPAN_NEXT continue
#else
#define XR_SPAN_NEXT return
#endif
#ifndef XR_SPAN_CAPMAX
#define XR_SPAN_CAPMAX 64
#endif
// Per-ray list capacity (candidates that survive the pre-filter = intervals at most): whatever fits the shared memory a CTA may
// use at XR_SPAN_MINBLOCKS CTAs per SM, at most 64.
constexpr int kSpanCapMax = XR_SPAN_CAPMAX;
constexpr double kZone = 1.0e-11; // half-width (in s) of the doubt zone around every end point

struct SpanArgs {
    const unsigned char* section;  // device copy of the SpanHeader section
    unsigned int section_bytes;
    unsigned int* tile_list;       // out: warp tiles ((view, tile) id * 4 + warp position) with a ray that needs the marching kernels
    unsigned int* tile_count;      // out: their number (zeroed before the launch); tile_count[1] = the work counter of the launch
    unsigned int total_items;      // warp tiles of the launch = views * tiles * 4
    int cap;                       // per-ray list capacity
    // screen-space bins (span_bin_kernel): per (view, tile) the (period, child) instances whose projection meets the tile
    const unsigned int* bin_counts;  // null = no bins: every ray walks the candidate grid
    const unsigned int* bin_lists;
    unsigned int bin_cap;
};

static int span_list_cap(unsigned int section_bytes) {
    const size_t budget = (size_t)216 * 1024 / XR_SPAN_MINBLOCKS - 1024;  // 228 KB per SM, 1 KB per CTA reserved, some slack
    const size_t sec = (section_bytes + 15u) & ~15u;
    if (budget < sec + (size_t)8 * kBlockThreads * 8) return 0;
    const size_t cap = (budget - sec) / ((size_t)kBlockThreads * 8);
    return (int)(cap < (size_t)kSpanCapMax ? cap : (size_t)kSpanCapMax);
}

size_t span_kernel_smem_bytes(unsigned int section_bytes) {
    const int cap = span_list_cap(section_bytes);
    if (cap <= 0) return (size_t)1 << 30;  // does not fit: the caller keeps to the marching kernels
    return (size_t)((section_bytes + 15u) & ~15u) + (size_t)cap * kBlockThreads * 8;
}

// The set of ray parameters that satisfy every constraint seen so far is an interval whose ends are known up to a
// doubt zone each:  t > L1 surely passes every lower bound, t < L0 surely fails one;  t < U0 surely passes every upper
// bound, t > U1 surely fails one.  A constraint with lower bound a (error z) gives L0 = max(L0, a - z), L1 = max(L1, a + z).
struct SpanRange {
    double L0, L1, U0, U1;
};

// lo <= c0 + cd * t <= hi.  A coordinate that does not move over the window (|cd| < 1e-9, |t| <= 1.75) is decided from
// c0 alone, with doubt when it sits within 4e-9 of a bound (a ray running inside a face plane).  eps bounds the
// rounding error of the coordinate as the reference computes it.
__device__ __forceinline__ bool span_slab(double c0, double cd, double inv_cd, double lo, double hi, double eps, SpanRange& r, unsigned int& doubt) {
    if (fabs(cd) < 1.0e-9) {
        const double dl = c0 - lo, dh = hi - c0;
        if (fabs(dl) < 4.0e-9 + eps || fabs(dh) < 4.0e-9 + eps) doubt |= 1u;  // a ray inside a bounding plane
        return dl >= 0.0 && dh >= 0.0;
    }
    double a = (lo - c0) * inv_cd, b = (hi - c0) * inv_cd;
    if (a > b) {
        const double tmp = a;
        a = b;
        b = tmp;
    }
    const double z = kZone + eps * fabs(inv_cd);
    r.L0 = fmax(r.L0, a - z);
    r.L1 = fmax(r.L1, a + z);
    r.U0 = fmin(r.U0, b - z);
    r.U1 = fmin(r.U1, b + z);
    return r.L0 <= r.U1;
}

// A t^2 + 2 B t + C < 0 with A >= 0 (the radial test of a sphere / an infinite cylinder along the ray).
__device__ __forceinline__ bool span_quadratic(double A, double B, double C, SpanRange& r, unsigned int& doubt) {
    if (A < 1.0e-19) {  // ray (anti)parallel to the axis: the radial distance does not change over the window
        if (fabs(C) < 1.0e-9) doubt |= 2u;
        return C < 0.0;
    }
    const double delta = 1.0e-13 * (2.0 * fabs(B) + fabs(C) + A);  // >> rounding error of the discriminant
    const double disc = B * B - A * C;
    if (disc < -delta) return false;
    const double iA = 1.0 / A;
    const double s_out = sqrt(disc + delta), s_in = sqrt(fmax(disc - delta, 0.0));  // roots of the widest / narrowest possible parabola
    r.L0 = fmax(r.L0, (-B - s_out) * iA - kZone);
    r.L1 = fmax(r.L1, (-B - s_in) * iA + kZone);
    r.U0 = fmin(r.U0, (-B + s_in) * iA - kZone);
    r.U1 = fmin(r.U1, (-B + s_out) * iA + kZone);
    return r.L0 <= r.U1;
}

// Smallest lattice ordinal whose position is > s; doubt when a lattice position lies within zw of s.
//   simple integrator:       ordinal k        <-> s_tab[k],                 k in [0, n)
//   hierarchical integrator: ordinal 16 k + j <-> s_tab[k] + j * ds_fine    (j = 1 .. nfine[k], main.go:183-189)
//                            ordinal 16 k + 15 <-> s_tab[k + 1]              (the coarse sample `right`, main.go:176-180)
template <int INTEG>
__device__ __forceinline__ unsigned int span_ordinal_after(const RenderParams& P, const unsigned char* __restrict__ nfine, double s,
                                                           double zw, double inv_ds, double inv_dsf, unsigned int& doubt) {
    const int n = P.n_steps;
    const double* __restrict__ S = P.s_tab;
    if (INTEG == 0) {
        if (!(s >= P.smin)) {
            if (P.smin - s < zw) doubt |= 4u;
            return 0u;
        }
        const double r = (s - P.smin) * inv_ds;
        if (r >= (double)n) {
            if (n > 0 && s - __ldg(S + n - 1) < zw) doubt |= 4u;
            return (unsigned int)n;
        }
        int k = (int)r;
        k = max(0, min(n - 1, k));
        while (k < n && __ldg(S + k) <= s) ++k;
        while (k > 0 && __ldg(S + k - 1) > s) --k;
        if (k > 0 && s - __ldg(S + k - 1) < zw) doubt |= 4u;
        if (k < n && __ldg(S + k) - s < zw) doubt |= 4u;
        return (unsigned int)k;
    }
    if (!(s >= P.smin)) {
        if (P.smin + P.ds_fine - s < zw) doubt |= 4u;
        return 1u;
    }
    const double s_end = __ldg(S + n);
    if (s >= s_end) {
        if (s - s_end < zw) doubt |= 4u;
        return 16u * (unsigned int)n;
    }
    int k = (int)((s - P.smin) * inv_ds);
    k = max(0, min(n - 1, k));
    while (k + 1 < n && __ldg(S + k + 1) <= s) ++k;
    while (k > 0 && __ldg(S + k) > s) --k;
    const double left = __ldg(S + k), right = __ldg(S + k + 1);
    const int nf = (int)__ldg(nfine + k);
    const double r = (s - left) * inv_dsf;
    if (fabs(r - rint(r)) * P.ds_fine < zw || right - s < zw) doubt |= 4u;
    const int j0 = (int)r + 1;
    return 16u * (unsigned int)k + (j0 > nf ? 15u : (unsigned int)j0);
}

// A ray that runs INSIDE a cell-face plane of a tessellation (the central pixel row at polar = 90 deg lies in z = 0): its
// coordinate along that axis stays within 1e-8 of the face, and which period a sample folds into is decided by the sign
// of a 1e-16 quantity.  The reference's expressions (main.go:147-149 position, objects.go:570-572 fold, :459 bounds)
// are monotone in s, so the period is a step function of the sample ordinal with at most one step: evaluate those
// expressions exactly at the ends of the ray's ordinal range and bisect for the step.
struct SpanDegenerate {
    int nA, nB;            // period of the first / last sample
    unsigned int eA, eB;   // the ray's ordinal range [eA, eB) inside the outer box
    unsigned int e_sw;     // first ordinal whose period is nB
};

template <int INTEG>
__device__ __forceinline__ bool span_degenerate_axis(const RenderParams& P, const unsigned char* __restrict__ nfine, double o_a, double d_a,
                                                  double lo, double hi, double D, SpanDegenerate& g) {
    auto period = [&](unsigned int e, bool& inside) -> int {
        double s;
        if (INTEG == 0) {
            s = __ldg(P.s_tab + e);
        } else {
            const unsigned int k = e >> 4, j = e & 15u;
            const unsigned int nf = (unsigned int)__ldg(nfine + k);
            if (j == 0u || j > nf) {
                s = __ldg(P.s_tab + k + (j == 0u ? 0u : 1u));  // not a fine sample: the neighbouring coarse one
            } else {
                s = __ldg(P.s_tab + k);
                for (unsigned int q = 0; q < j; ++q) s = dadd(s, P.ds_fine);  // main.go:183,189: left += ds
            }
        }
        const double x = dadd(o_a, dmul(d_a, s));
        const double n = floor(ddiv(dsub(x, lo), D));
        const double xf = dsub(x, dmul(D, n));
        inside = !(xf < lo || xf > hi);
        return (int)n;
    };
    if (g.eA >= g.eB) return true;
    bool inA, inB, in0, in1;
    g.nA = period(g.eA, inA);
    g.nB = period(g.eB - 1u, inB);
    g.e_sw = g.eB;
    if (g.nA == g.nB) return inA && inB;
    if (abs(g.nA - g.nB) != 1) return false;
    unsigned int a = g.eA, b = g.eB - 1u;
    while (b - a > 1u) {
        const unsigned int m = a + ((b - a) >> 1);
        bool dummy;
        const int n = period(m, dummy);
        if (n == g.nA) a = m;
        else if (n == g.nB) b = m;
        else return false;
    }
    period(a, in0);
    period(b, in1);
    g.e_sw = b;
    return inA && inB && in0 && in1;
}

// ObjectCollection.Density (objects.go:422-438) for the set of children that contain the point, times the multiplier
// (main.go:139).  Children outside return 0.0 and leave the sum untouched, so only the members matter; order = child order.
__device__ __forceinline__ double span_combine(const SpanChild* __restrict__ ch, unsigned long long active, unsigned int flags, double dm) {
    if (active == 0ull) return 0.0;
    if (!(flags & SPAN_CLAMPS)) return dmul(ch[0].rho, dm);  // bare primitive
    double sum = 0.0;
    unsigned int lo = (unsigned int)active, hi = (unsigned int)(active >> 32);
    for (int half = 0; half < 2; ++half) {
        unsigned int w = half ? hi : lo;
        while (w) {
            const int c = half * 32 + __ffs((int)w) - 1;
            w &= w - 1;
            const double rho = ch[c].rho;
            if ((flags & SPAN_GREEDY) && rho > 0.0) return dmul(rho, dm);
            sum = dadd(sum, rho);
        }
    }
    if (sum < 0.0) sum = 0.0;
    else if (sum > 1.0) sum = 1.0;
    return dmul(sum, dm);
}

template <int INTEG, bool COUNT>
__global__ void __launch_bounds__(kBlockThreads, XR_SPAN_MINBLOCKS) render_span_kernel(const RenderParams P, const unsigned char* __restrict__ nfine,
                                                                       const SpanArgs SA) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    {
        uint4* dst = reinterpret_cast<uint4*>(smem);
        const uint4* src = reinterpret_cast<const uint4*>(SA.section);
        const int n16 = (int)((SA.section_bytes + 15u) >> 4);
        for (int k = tid; k < n16; k += kBlockThreads) dst[k] = __ldg(src + k);
    }
    if (COUNT && P.stats && blockIdx.x == 0 && tid == 0) atomicOr(P.stats + 7, 0x10000ull);  // "the interval renderer ran"
    const SpanHeader& H = *reinterpret_cast<const SpanHeader*>(smem);
    unsigned int* cand = reinterpret_cast<unsigned int*>(smem + ((SA.section_bytes + 15u) & ~15u));
    // interval q is written after candidate q has been read and there are never more intervals than candidates read, so
    // the intervals' first words live in the candidate slots
    unsigned int* iv_in = cand;
    const int cap = SA.cap;
    unsigned int* iv_out = cand + cap * kBlockThreads;
    __syncthreads();
    const SpanChild* __restrict__ ch = reinterpret_cast<const SpanChild*>(smem + H.child_off);
    const unsigned long long* __restrict__ masks = reinterpret_cast<const unsigned long long*>(smem + H.mask_off);
    const unsigned int flags = H.flags;
    const bool tess = (flags & SPAN_TESS) != 0u;

    // Persistent warps: every warp of the grid draws warp tiles (32 pixels, the same 4 x 8 footprint the marching kernels
    // use) from a global counter, four at a time (= one CTA tile of the marching kernels), until none are left; rays differ
    // a lot in cost (most miss the object), so a fixed assignment would leave warps idle.  Nothing below synchronises beyond
    // the warp.
#if XR_SPAN_PERSISTENT
    constexpr unsigned int kChunk = 4;
    for (;;) {
    unsigned int base = 0u;
    if ((tid & 31) == 0) base = atomicAdd(SA.tile_count + 1, kChunk);
    base = __shfl_sync(FULL_MASK, base, 0);
    if (base >= SA.total_items) break;
    for (unsigned int item = base; item < min(base + kChunk, SA.total_items); ++item) {
#else
    {
    {
    const unsigned int item = blockIdx.x * (kBlockThreads / 32) + (tid >> 5);
#endif
    int view, i, j;
    pixel_of_thread(P, item >> 2, item & 3u, view, i, j);
    const bool valid = i < P.res && j < P.res;
    if (!valid) { i = 0; j = 0; }

    unsigned int doubt = 0u;  // reasons why this ray is left to the marching kernels (bit codes, reported in stats[7])
    // ---- the ray in object space, centred on the window: x(t) = c + d t, t = s - R ----
    double cx, cy, cz, dx, dy, dz;
    {
        const Ray64 ray = make_ray(P.cams[view], i, j, P.res);
        double ox = ray.o[0], oy = ray.o[1], oz = ray.o[2];
        dx = ray.d[0]; dy = ray.d[1]; dz = ray.d[2];
        if (flags & SPAN_HAS_WARP) {
            const double* m = H.warp_m;
            const double nx = m[0] * ox + m[1] * oy + m[2] * oz + H.warp_b[0], ny = m[3] * ox + m[4] * oy + m[5] * oz + H.warp_b[1],
                         nz = m[6] * ox + m[7] * oy + m[8] * oz + H.warp_b[2];
            const double ex = m[0] * dx + m[1] * dy + m[2] * dz, ey = m[3] * dx + m[4] * dy + m[5] * dz, ez = m[6] * dx + m[7] * dy + m[8] * dz;
            ox = nx; oy = ny; oz = nz;
            dx = ex; dy = ey; dz = ez;
        }
        cx = ox + dx * P.s_center;
        cy = oy + dy * P.s_center;
        cz = oz + dz * P.s_center;
    }
    const double idx = 1.0 / dx, idy = 1.0 / dy, idz = 1.0 / dz;  // inf for an axis-parallel ray: never used then (|d| < 1e-9 branch)
    const double epsx = 1.0e-14 * (8.0 + fabs(cx) + fabs(cy) + fabs(cz));  // bound of the rounding error of a reference-side coordinate

    // ---- ray versus the outer box (TESS: objects.go:569, inclusive) / the region (FLAT) ----
    const double half_win = 0.5 * (P.smax - P.smin) + 1.0;
    SpanRange R0 = {-half_win, -half_win, half_win, half_win};
    bool hit = valid;
    hit = hit && span_slab(cx, dx, idx, H.outer[0], H.outer[3], epsx, R0, doubt);
    hit = hit && span_slab(cy, dy, idy, H.outer[1], H.outer[4], epsx, R0, doubt);
    hit = hit && span_slab(cz, dz, idz, H.outer[2], H.outer[5], epsx, R0, doubt);
    if (!valid) doubt = 0u;
    const double ta = R0.L0, tb = R0.U1;
    const double inv_ds = 1.0 / P.ds, inv_dsf = 1.0 / P.ds_fine;

    // ---- a ray inside a cell-face plane of the tessellation: exact period per sample along that axis ----
    int deg_axis = -1, deg_m = 0;
    SpanDegenerate dg = {0, 0, 0u, 0u, 0u};
    if (tess) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double da = a == 0 ? dx : (a == 1 ? dy : dz), ca = a == 0 ? cx : (a == 1 ? cy : cz);
            if (hit && fabs(da) < 1.0e-9) {
                const double q = (ca - H.uc_lo[a]) / H.uc_d[a], m = rint(q);
                if (fabs(q - m) * fabs(H.uc_d[a]) < 6.0e-9 + epsx) {
                    if (deg_axis >= 0 || (flags & SPAN_HAS_WARP)) doubt |= 8u;  // two such axes, or a warp between ray and fold
                    else {
                        deg_axis = a;
                        deg_m = (int)m;
                    }
                }
            }
        }
        if (__any_sync(FULL_MASK, deg_axis >= 0)) {
            if (deg_axis >= 0) {
                dg.eA = span_ordinal_after<INTEG>(P, nfine, 0.5 * (R0.L0 + R0.L1) + P.s_center, 0.5 * (R0.L1 - R0.L0), inv_ds, inv_dsf, doubt);
                dg.eB = span_ordinal_after<INTEG>(P, nfine, 0.5 * (R0.U0 + R0.U1) + P.s_center, 0.5 * (R0.U1 - R0.U0), inv_ds, inv_dsf, doubt);
            }
            const int a = max(deg_axis, 0);
            const double da = a == 0 ? dx : (a == 1 ? dy : dz);
            const bool ok = span_degenerate_axis<INTEG>(P, nfine, P.cams[view].eye[a], da, H.uc_lo[a], H.uc_hi[a], H.uc_d[a], dg);
            if (deg_axis >= 0 && !ok) doubt |= 16u;
        }
    }

    // ---- phase 1: candidates.  Either the tile's screen-space bin (one list per 8 x 16 pixel tile, built per launch by
    // span_bin_kernel) or, without one, an fp32 walk of this ray through the candidate grid.  Either way every candidate
    // goes through a conservative fp32 pre-filter and the survivors land in cand[]. ----
    int ncand = 0;
    bool overflow = false;
    const float fdx = (float)dx, fdy = (float)dy, fdz = (float)dz;
    const float fcx = (float)(cx - H.uc_lo[0]), fcy = (float)(cy - H.uc_lo[1]), fcz = (float)(cz - H.uc_lo[2]);  // relative to the cell origin
    const float fta = (float)ta, ftb = (float)tb;
    const float fA = fdx * fdx + fdy * fdy + fdz * fdz;
    // child c of the period coded in pcode; (qx, qy, qz) = the ray centre relative to that period's copy of the cell
    auto test_and_push = [&](int c, int pcode, float qx, float qy, float qz) {
        const SpanChild& K = ch[c];
        bool pass = true;
        const float wx = qx - K.f[0], wy = qy - K.f[1], wz = qz - K.f[2];
        float A = fA, B = wx * fdx + wy * fdy + wz * fdz, C = wx * wx + wy * wy + wz * wz;
        float dv = 0.0f, wv = 0.0f, ivv = 0.0f;
        if (K.type == OP_CYL) {  // the infinite cylinder, inflated
            dv = fdx * K.f[3] + fdy * K.f[4] + fdz * K.f[5];
            wv = wx * K.f[3] + wy * K.f[4] + wz * K.f[5];
            ivv = K.f[6];
            A -= dv * dv * ivv;
            B -= wv * dv * ivv;
            C -= wv * wv * ivv;
            C -= K.f[7];
        } else {  // the sphere itself / the bounding sphere of a box or parallelepiped, inflated
            C -= K.f[3];
        }
        const float disc = B * B - A * C;
        if (disc < -2.0e-6f * (1.0f + B * B)) pass = false;
        else {
            const float iA = 1.0f / fmaxf(A, 1.0e-12f);
            const float th = sqrtf(fmaxf(disc, 0.0f)) * iA + 1.0e-3f, tm = -B * iA;
            if (tm + th < fta || tm - th > ftb) pass = false;
            if (K.type == OP_CYL) {
                const float cm = (wv + tm * dv) * ivv, hc = th * fabsf(dv) * ivv + 1.0e-3f;
                if (cm + hc < 0.0f || cm - hc > 1.0f) pass = false;
            }
        }
        if (pass) {
            if (ncand < cap) cand[ncand * kBlockThreads + tid] = (unsigned int)c | (unsigned int)pcode;
            else overflow = true;
            ++ncand;
        }
    };
    bool walk = hit;
    if (SA.bin_counts) {  // uniform
        const unsigned int nb = __ldg(SA.bin_counts + (item >> 2));
        if (nb <= SA.bin_cap) {  // (a bin that overflowed is not used: those tiles walk)
            walk = false;
            const unsigned int* __restrict__ bl = SA.bin_lists + (size_t)(item >> 2) * SA.bin_cap;
            const float ucdx = (float)H.uc_d[0], ucdy = (float)H.uc_d[1], ucdz = (float)H.uc_d[2];
            const float ulx = H.f_uc_lo[0], uly = H.f_uc_lo[1], ulz = H.f_uc_lo[2];
            for (unsigned int q = 0; q < nb; ++q) {
                const unsigned int code = __ldg(bl + q);
                const int px = (int)((code >> 6) & 31u) - 16, py = (int)((code >> 11) & 31u) - 16, pz = (int)((code >> 16) & 31u) - 16;
                if (deg_axis >= 0) {  // a ray inside a face plane only ever folds into the two periods either side of it
                    const int pa = deg_axis == 0 ? px : (deg_axis == 1 ? py : pz);
                    if (pa != dg.nA && pa != dg.nB) continue;
                }
                if (hit)
                    test_and_push((int)(code & 63u), (int)(code & ~63u), fcx - (float)px * ucdx + ulx, fcy - (float)py * ucdy + uly,
                                  fcz - (float)pz * ucdz + ulz);
            }
        }
    }
    const int n_pass = walk ? (deg_axis >= 0 ? 2 : 1) : 0;  // a ray inside a face plane walks the cells on either side of it
    const int max_pass = __reduce_max_sync(FULL_MASK, n_pass);
    for (int pass = 0; pass < max_pass; ++pass) {
        if (pass >= n_pass) continue;
        const int gx = (int)H.g[0], gy = (int)H.g[1], gz = (int)H.g[2];
        const float icx = H.f_inv_cell[0], icy = H.f_inv_cell[1], icz = H.f_inv_cell[2];
        const float csx = H.f_cell[0], csy = H.f_cell[1], csz = H.f_cell[2];
        const float ucdx = (float)H.uc_d[0], ucdy = (float)H.uc_d[1], ucdz = (float)H.uc_d[2];
        const float ulx = H.f_uc_lo[0], uly = H.f_uc_lo[1], ulz = H.f_uc_lo[2];
        const float slack = 2.0e-4f * (1.0f + fabsf(fcx) + fabsf(fcy) + fabsf(fcz));
        const float t_begin = fta - slack, t_end = ftb + slack;
        const float ifx = 1.0f / fdx, ify = 1.0f / fdy, ifz = 1.0f / fdz;
        int ix = __float2int_rd(fmaf(fdx, t_begin, fcx) * icx), iy = __float2int_rd(fmaf(fdy, t_begin, fcy) * icy),
            iz = __float2int_rd(fmaf(fdz, t_begin, fcz) * icz);
        const int sx = fdx > 0.0f ? 1 : -1, sy = fdy > 0.0f ? 1 : -1, sz = fdz > 0.0f ? 1 : -1;
        if (deg_axis == 0) ix = deg_m * gx - 1 + pass;  // last cell of period m - 1, then first cell of period m
        if (deg_axis == 1) iy = deg_m * gy - 1 + pass;
        if (deg_axis == 2) iz = deg_m * gz - 1 + pass;
        const float tie_d = 0.25f * slack;
        unsigned long long done = 0ull;

        // visit one grid cell (local index l*, period p*): new children of its mask are pre-filtered and pushed
        // (`done` = the children already seen in the current period; the walk clears it when it enters another period.  The
        // rare extra cells of a near-tie are visited with `side` set: they may belong to a neighbouring period, so they
        // neither read nor update it -- a child pushed twice just yields the same interval twice.)
        auto visit = [&](int lx, int ly, int lz, int px, int py, int pz, bool side) {
            if (!tess && (px | py | pz) != 0) return;  // outside the region
            unsigned long long nw = masks[(lz * gy + ly) * gx + lx];
            if (!side) {
                nw &= ~done;
                done |= nw;
            }
            if (nw == 0ull) return;
            const int pcode = ((px + 16) << 6) | ((py + 16) << 11) | ((pz + 16) << 16);  // |p| <= 15: scene_compile.cpp build_span
            // ray centre relative to this period's copy of the cell
            const float qx = fcx - (float)px * ucdx + ulx, qy = fcy - (float)py * ucdy + uly, qz = fcz - (float)pz * ucdz + ulz;
            while (nw) {
                const int c = __ffsll((long long)nw) - 1;
                nw &= nw - 1ull;
                test_and_push(c, pcode, qx, qy, qz);
            }
        };

        // local cell index and period per axis, kept incrementally (one division each, here)
        int px = __float2int_rd(((float)ix + 0.5f) / (float)gx), py = __float2int_rd(((float)iy + 0.5f) / (float)gy),
            pz = __float2int_rd(((float)iz + 0.5f) / (float)gz);
        int lx = ix - px * gx, ly = iy - py * gy, lz = iz - pz * gz;
        // parameter at which the ray leaves the cell along each axis, from the absolute cell index (no drift)
        // (an axis the ray does not move along -- exactly, or to 1e-9 inside a face plane -- is never stepped)
        float tx = (fdx != 0.0f && deg_axis != 0) ? (((float)(ix + (sx > 0 ? 1 : 0))) * csx - fcx) * ifx : 3.0e38f;
        float ty = (fdy != 0.0f && deg_axis != 1) ? (((float)(iy + (sy > 0 ? 1 : 0))) * csy - fcy) * ify : 3.0e38f;
        float tz = (fdz != 0.0f && deg_axis != 2) ? (((float)(iz + (sz > 0 ? 1 : 0))) * csz - fcz) * ifz : 3.0e38f;
        // two faces count as reached together when the second is closer than tie_d (in space) at that moment
        const float hx = (fdx != 0.0f && deg_axis != 0) ? tie_d * fabsf(ifx) : 0.0f, hy = (fdy != 0.0f && deg_axis != 1) ? tie_d * fabsf(ify) : 0.0f,
                    hz = (fdz != 0.0f && deg_axis != 2) ? tie_d * fabsf(ifz) : 0.0f;
        const float hmax = fmaxf(hx, fmaxf(hy, hz));
        for (int guard = 0; guard < 4096; ++guard) {
            visit(lx, ly, lz, px, py, pz, false);
            const float tm = fminf(tx, fminf(ty, tz));
            if (!(tm <= t_end)) break;
            const bool nx = tx - tm < hx, ny = ty - tm < hy, nz = tz - tm < hz;
            if (fmaxf(fminf(tx, ty), fminf(fmaxf(tx, ty), tz)) - tm < hmax) {  // the second face is near: look closer
              if ((nx ? 1 : 0) + (ny ? 1 : 0) + (nz ? 1 : 0) > 1) {
                // the exact ray may cross these faces in another order: visit every cell of the little block between here
                // and the far corner (the far corner itself is the next regular visit)
                for (int m = 1; m < 7; ++m) {
                    const bool bx = m & 1, by = m & 2, bz = m & 4;
                    if ((bx && !nx) || (by && !ny) || (bz && !nz)) continue;
                    if (bx == nx && by == ny && bz == nz) continue;
                    int vx = lx + (bx ? sx : 0), vy = ly + (by ? sy : 0), vz = lz + (bz ? sz : 0), qx = px, qy = py, qz = pz;
                    if (vx == gx) { vx = 0; ++qx; } else if (vx < 0) { vx = gx - 1; --qx; }
                    if (vy == gy) { vy = 0; ++qy; } else if (vy < 0) { vy = gy - 1; --qy; }
                    if (vz == gz) { vz = 0; ++qz; } else if (vz < 0) { vz = gz - 1; --qz; }
                    visit(vx, vy, vz, qx, qy, qz, true);
                }
              }
            }
            if (nx) {
                ix += sx;
                lx += sx;
                if (lx == gx) { lx = 0; ++px; done = 0ull; } else if (lx < 0) { lx = gx - 1; --px; done = 0ull; }
                tx = (((float)(ix + (sx > 0 ? 1 : 0))) * csx - fcx) * ifx;
            }
            if (ny) {
                iy += sy;
                ly += sy;
                if (ly == gy) { ly = 0; ++py; done = 0ull; } else if (ly < 0) { ly = gy - 1; --py; done = 0ull; }
                ty = (((float)(iy + (sy > 0 ? 1 : 0))) * csy - fcy) * ify;
            }
            if (nz) {
                iz += sz;
                lz += sz;
                if (lz == gz) { lz = 0; ++pz; done = 0ull; } else if (lz < 0) { lz = gz - 1; --pz; done = 0ull; }
                tz = (((float)(iz + (sz > 0 ? 1 : 0))) * csz - fcz) * ifz;
            }
            if (guard == 4095) overflow = true;
        }
    }
    if (overflow) ncand = 0;

    // ---- phase 2: exact interval of every surviving (period, child), mapped to lattice ordinals ----
    // (the reciprocals are recomputed here rather than kept alive across the walk: 6 registers for 3 divisions per ray)
    const double jdx = 1.0 / dx, jdy = 1.0 / dy, jdz = 1.0 / dz;
    const unsigned int e_first = INTEG == 1 ? 1u : 0u;
    const unsigned int e_last = INTEG == 1 ? 16u * (unsigned int)P.n_steps : (unsigned int)P.n_steps;  // one past the last ordinal
    int niv = 0;
    unsigned int prim_tests = 0;
    const int max_cand = __reduce_max_sync(FULL_MASK, ncand);
    for (int q = 0; q < max_cand; ++q) {
        if (q >= ncand) continue;
        const unsigned int code = cand[q * kBlockThreads + tid];
        const int c = (int)(code & 63u);
        const int px = (int)((code >> 6) & 31u) - 16, py = (int)((code >> 11) & 31u) - 16, pz = (int)((code >> 16) & 31u) - 16;
        if (COUNT) ++prim_tests;
        // ray centre in the coordinates of this period: x' = x - dx * n (objects.go:571)
        const double ux = cx - H.uc_d[0] * (double)px, uy = cy - H.uc_d[1] * (double)py, uz = cz - H.uc_d[2] * (double)pz;
        SpanRange r = R0;
        bool ok = true;
        unsigned int w_lo = e_first, w_hi = e_last;  // ordinal window of this period along a degenerate axis
        if (deg_axis >= 0) {
            const int pa = deg_axis == 0 ? px : (deg_axis == 1 ? py : pz);
            if (pa == dg.nA) {
                w_lo = dg.eA;
                w_hi = dg.nA == dg.nB ? dg.eB : dg.e_sw;
            } else if (pa == dg.nB) {
                w_lo = dg.e_sw;
                w_hi = dg.eB;
            } else {
                continue;
            }
        }
        if (tess) {  // the period's own cell: floor((x - min) / d) == n  <=>  min <= x' < min + d
            if (deg_axis != 0) ok = ok && span_slab(ux, dx, jdx, H.uc_lo[0], H.uc_lo[0] + H.uc_d[0], epsx, r, doubt);
            if (deg_axis != 1) ok = ok && span_slab(uy, dy, jdy, H.uc_lo[1], H.uc_lo[1] + H.uc_d[1], epsx, r, doubt);
            if (deg_axis != 2) ok = ok && span_slab(uz, dz, jdz, H.uc_lo[2], H.uc_lo[2] + H.uc_d[2], epsx, r, doubt);
        }
        const SpanChild& K = ch[c];
        if (ok) {
            const double* p = K.p;
            if (K.type == OP_CYL || K.type == OP_SPHERE) {
                const double wx = ux - p[0], wy = uy - p[1], wz = uz - p[2];
                double A = dx * dx + dy * dy + dz * dz, B = wx * dx + wy * dy + wz * dz, C = wx * wx + wy * wy + wz * wz;
                if (K.type == OP_CYL) {
                    const double dv = dx * p[3] + dy * p[4] + dz * p[5], wv = wx * p[3] + wy * p[4] + wz * p[5], ivv = p[6];
                    A -= dv * dv * ivv;
                    B -= wv * dv * ivv;
                    C -= wv * wv * ivv;
                    C -= p[7];
                    // caps: 0 <= (w.v + t d.v) / v.v <= 1, inclusive (objects.go:339-341)
                    const double cd = dv * ivv;
                    ok = span_slab(wv * ivv, cd, 1.0 / cd, 0.0, 1.0, 1.0e-13, r, doubt);
                } else {
                    C -= p[3];
                }
                ok = ok && span_quadratic(fmax(A, 0.0), B, C, r, doubt);
            } else if (K.type == OP_BOX) {
                ok = ok && span_slab(ux, dx, jdx, p[0] - p[3], p[0] + p[3], epsx, r, doubt);
                ok = ok && span_slab(uy, dy, jdy, p[1] - p[4], p[1] + p[4], epsx, r, doubt);
                ok = ok && span_slab(uz, dz, jdz, p[2] - p[5], p[2] + p[5], epsx, r, doubt);
            } else {  // parallelepiped: 0 < Minv (x - o) < 1 per component (objects.go:249-254)
                const double wx = ux - p[0], wy = uy - p[1], wz = uz - p[2];
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const double q0 = p[3 + 3 * rr] * wx + p[4 + 3 * rr] * wy + p[5 + 3 * rr] * wz;
                    const double qd = p[3 + 3 * rr] * dx + p[4 + 3 * rr] * dy + p[5 + 3 * rr] * dz;
                    ok = ok && span_slab(q0, qd, 1.0 / qd, 0.0, 1.0, epsx * (1.0 + p[12 + rr]), r, doubt);
                }
            }
        }
        if (!ok) continue;
        if (r.L1 > r.U0) {  // no parameter is surely inside: any lattice sample between L0 and U1 is undecided
            span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.L0 + r.U1) + P.s_center, 0.5 * (r.U1 - r.L0), inv_ds, inv_dsf, doubt);
            continue;
        }
        const unsigned int e_in = max(w_lo, span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.L0 + r.L1) + P.s_center, 0.5 * (r.L1 - r.L0), inv_ds, inv_dsf, doubt));
        const unsigned int e_out = min(w_hi, span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.U0 + r.U1) + P.s_center, 0.5 * (r.U1 - r.U0), inv_ds, inv_dsf, doubt));
        if (e_in >= e_out) continue;  // no lattice sample inside
        if (niv < cap) {
            iv_in[niv * kBlockThreads + tid] = e_in | ((unsigned int)c << 26);
            iv_out[niv * kBlockThreads + tid] = e_out;
        } else {
            overflow = true;
        }
        ++niv;
    }
    if (overflow) niv = 0;

    // ---- phase 3: integer sweep over the end points = the reference's loop, piece by piece ----
    double T = 0.0;
    unsigned int n_fine = 0;
    if (niv > 0) {
        const double dm = P.dm, DS = P.ds, dsf = P.ds_fine;
        bool prev_z = false;  // prev_rho := 0.0 (main.go:179)
        double facc = 0.0;    // sum of the fine samples seen so far in the coarse interval under way
        // nothing before the first interval and nothing after the last one: both stretches add 0 and flip nothing
        unsigned int e = e_last, e_stop = 0u;
        for (int q = 0; q < niv; ++q) {
            e = min(e, iv_in[q * kBlockThreads + tid] & 0x3ffffffu);
            e_stop = max(e_stop, iv_out[q * kBlockThreads + tid]);
        }
        e = max(e, e_first);
        e_stop = min(e_stop + (INTEG == 1 ? 32u : 0u), e_last);  // (+2 coarse intervals: the exit transition's refinement)
        while (e < e_stop) {
            unsigned int next = e_stop;
            unsigned long long active = 0ull;
            for (int q = 0; q < niv; ++q) {
                const unsigned int a = iv_in[q * kBlockThreads + tid], b = iv_out[q * kBlockThreads + tid];
                const unsigned int ein = a & 0x3ffffffu;
                if (ein <= e && e < b) {
                    active |= 1ull << (a >> 26);
                    next = min(next, b);
                } else if (ein > e) {
                    next = min(next, ein);
                }
            }
            const double v = span_combine(ch, active, flags, dm);
            if (INTEG == 0) {
                T += v * (P.ds * (double)(next - e));  // every sample of the piece adds v * ds (main.go:147-152)
            } else if (v != 0.0 || prev_z) {
                const int k0 = (int)(e >> 4), j0 = (int)(e & 15u), k1 = (int)((next - 1u) >> 4), j1 = (int)((next - 1u) & 15u);
                // one coarse sample `right` of interval k with density v (main.go:180-192)
#define XR_SPAN_COARSE(KK)                                          \
    do {                                                            \
        const bool z_ = v != 0.0;                                   \
        if (z_ != prev_z) {                                         \
            T += dsf * (v + facc);                                  \
            n_fine += (unsigned int)__ldg(nfine + (KK));            \
        } else {                                                    \
            T += DS * v;                                            \
        }                                                           \
        prev_z = z_;                                                \
        facc = 0.0;                                                 \
    } while (0)
                if (k0 == k1) {
                    const int nf = (int)__ldg(nfine + k0);
                    const int cf = max(0, min(j1, nf) - j0 + 1);
                    facc += v * (double)cf;
                    if (j1 == 15) XR_SPAN_COARSE(k0);
                } else {
                    const int nf0 = (int)__ldg(nfine + k0);
                    facc += v * (double)max(0, nf0 - j0 + 1);
                    XR_SPAN_COARSE(k0);
                    const int mid = k1 - k0 - 1;
                    if (mid > 0) T += DS * v * (double)mid;  // whole coarse intervals inside the piece: nothing flips
                    const int nf1 = (int)__ldg(nfine + k1);
                    facc = v * (double)min(j1, nf1);
                    if (j1 == 15) XR_SPAN_COARSE(k1);
                }
#undef XR_SPAN_COARSE
            } else {
                // v == 0 and the last coarse sample was 0: adds nothing, flips nothing
                if ((e >> 4) != ((next - 1u) >> 4) || ((next - 1u) & 15u) == 15u) facc = 0.0;
            }
            e = next;
        }
    }
    if (overflow) doubt |= 32u;
    const bool bad = valid && doubt != 0u;
    const bool tile_bad = __any_sync(FULL_MASK, bad);
    if ((tid & 31) == 0 && tile_bad) {
        SA.tile_list[atomicAdd(SA.tile_count, 1u)] = item;
        if (COUNT && P.stats) atomicAdd(P.stats + 6, 1ull);  // warp tiles handed to the marching kernels
    }
    if (COUNT && P.stats && bad) atomicOr(P.stats + 7, (unsigned long long)doubt);
#ifdef XRAY_DEV_KNOBS
    if (P.dbg_cause == 77) {  // development builds: show which rays are handed over, and why
        store_pixel(P, view, i, j, valid, bad ? -(double)doubt : exp(-(P.flat_field + T)));
        XR_SPAN_NEXT;
    }
#endif
    store_pixel(P, view, i, j, valid, exp(-(P.flat_field + T)));
    if (COUNT && !tile_bad)
        add_stats(P, valid ? (unsigned long long)P.n_steps + n_fine : 0ull, (unsigned long long)niv, 0ull, prim_tests, valid ? 1ull : 0ull);
    }  // warp tiles of the chunk
    }  // chunks
}

// ---------------------------------------------------------------------------------------------------------
// Screen-space binning.  Which children can a ray meet?  All rays of a view leave one point, so the answer is a 2-D
// question on the detector: project every (period, child) instance of the scene -- its bounding sphere, or for a
// cylinder the capsule around its axis -- and list it in the 8 x 16 pixel tiles its projection touches.  One warp per
// instance and view; lanes stride over the tiles of the projection's bounding box.  Conservative by construction (a
// sphere of radius R at depth D around the image point p projects inside a disc of radius f R / (D - R) * sqrt(1 + |p|^2 / f^2);
// the bins use the larger (1 + |p|^2 / f^2), 2 % more and two pixel pitches of padding); a bin that overflows its
// capacity is ignored by the renderer, which then walks the candidate grid for that tile instead.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads) span_bin_kernel(const RenderParams P, const unsigned char* __restrict__ section,
                                                                 unsigned int* __restrict__ counts, unsigned int* __restrict__ lists,
                                                                 unsigned int cap) {
    const SpanHeader& H = *reinterpret_cast<const SpanHeader*>(section);
    const SpanChild* __restrict__ ch = reinterpret_cast<const SpanChild*>(section + H.child_off);
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int inst = blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5);
    if (inst >= H.n_periods * H.n_children) return;
    const int c = (int)(inst % H.n_children);
    const unsigned int pidx = inst / H.n_children;
    const int nx = H.n_hi[0] - H.n_lo[0] + 1, ny = H.n_hi[1] - H.n_lo[1] + 1;
    const int px = H.n_lo[0] + (int)(pidx % (unsigned int)nx), py = H.n_lo[1] + (int)((pidx / (unsigned int)nx) % (unsigned int)ny),
              pz = H.n_lo[2] + (int)(pidx / (unsigned int)(nx * ny));
    const int view = blockIdx.y;
    const CamDev& cam = P.cams[view];
    const SpanChild& K = ch[c];
    const float sx = (float)px * (float)H.uc_d[0], sy = (float)py * (float)H.uc_d[1], sz = (float)pz * (float)H.uc_d[2];
    float ax, ay, az, bx, by, bz;
    if (K.type == OP_CYL) {
        ax = K.f[0] + sx; ay = K.f[1] + sy; az = K.f[2] + sz;
        bx = ax + K.f[3]; by = ay + K.f[4]; bz = az + K.f[5];
    } else {
        ax = bx = K.bs[0] + sx; ay = by = K.bs[1] + sy; az = bz = K.bs[2] + sz;
    }
    const float R = K.bs[3];
    // camera coordinates: X = V3 * cam + t  =>  cam = V3^T (X - t), V3 orthonormal (checked by the caller)
    const float tx = (float)cam.view[3], ty = (float)cam.view[7], tz = (float)cam.view[11];
    float ca[3], cb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v0 = (float)cam.view[0 * 4 + k], v1 = (float)cam.view[1 * 4 + k], v2 = (float)cam.view[2 * 4 + k];
        ca[k] = v0 * (ax - tx) + v1 * (ay - ty) + v2 * (az - tz);
        cb[k] = v0 * (bx - tx) + v1 * (by - ty) + v2 * (bz - tz);
    }
    const float f = (float)cam.f, h = 0.5f * (float)P.res, pitch = 1.0f / h;
    const float Da = -ca[2], Db = -cb[2], Dmin = fminf(Da, Db);
    int ti_lo = 0, ti_hi = P.tiles_i - 1, tj_lo = 0, tj_hi = P.tiles_j - 1;
    float pax = 0.0f, pay = 0.0f, pbx = 0.0f, pby = 0.0f, Rn = 3.0e38f;
    if (Dmin - R > 0.05f) {  // otherwise (an instance at or behind the eye) every tile gets it
        pax = f * ca[0] / Da; pay = f * ca[1] / Da;
        pbx = f * cb[0] / Db; pby = f * cb[1] / Db;
        const float pm2 = fmaxf(pax * pax + pay * pay, pbx * pbx + pby * pby);
        Rn = 1.02f * f * R / (Dmin - R) * (1.0f + pm2 / (f * f)) + 2.0f * pitch;
        const float xmin = fminf(pax, pbx) - Rn, xmax = fmaxf(pax, pbx) + Rn, ymin = fminf(pay, pby) - Rn, ymax = fmaxf(pay, pby) + Rn;
        // tile ti holds the pixels i in [ti * kTileI, ti * kTileI + kTileI - 1], pixel i sits at i * pitch - 1 (main.go:463)
        ti_lo = max(0, (int)ceilf(((xmin + 1.0f) * h - (float)(kTileI - 1)) / (float)kTileI));
        ti_hi = min(P.tiles_i - 1, (int)floorf((xmax + 1.0f) * h / (float)kTileI));
        tj_lo = max(0, (int)ceilf(((ymin + 1.0f) * h - (float)(kTileJ - 1)) / (float)kTileJ));
        tj_hi = min(P.tiles_j - 1, (int)floorf((ymax + 1.0f) * h / (float)kTileJ));
    }
    if (ti_hi < ti_lo || tj_hi < tj_lo) return;
    const int wj = tj_hi - tj_lo + 1, n = (ti_hi - ti_lo + 1) * wj;
    const unsigned int code = (unsigned int)c | (unsigned int)(((px + 16) << 6) | ((py + 16) << 11) | ((pz + 16) << 16));
    const float hd = 0.5f * pitch * sqrtf((float)((kTileI - 1) * (kTileI - 1) + (kTileJ - 1) * (kTileJ - 1)));
    const float ex = pbx - pax, ey = pby - pay, ee = ex * ex + ey * ey;
    const size_t tile0 = (size_t)view * P.tiles_i * P.tiles_j;
    for (int t = (int)lane; t < n; t += 32) {
        const int ti = ti_lo + t / wj, tj = tj_lo + t % wj;
        if (Rn < 1.0e38f) {  // distance from the tile centre to the projected axis against Rn + the tile's half diagonal
            const float mx = ((float)(ti * kTileI) + 0.5f * (float)(kTileI - 1)) * pitch - 1.0f - pax;
            const float my = ((float)(tj * kTileJ) + 0.5f * (float)(kTileJ - 1)) * pitch - 1.0f - pay;
            const float u = ee > 0.0f ? fminf(fmaxf((mx * ex + my * ey) / ee, 0.0f), 1.0f) : 0.0f;
            const float qx = mx - u * ex, qy = my - u * ey, lim = Rn + hd;
            if (qx * qx + qy * qy > lim * lim) continue;
        }
        const size_t tile = tile0 + (size_t)ti * P.tiles_j + tj;
        const unsigned int slot = atomicAdd(counts + tile, 1u);
        if (slot < cap) lists[tile * cap + slot] = code;
    }
}

// d_bins: null, or room for tiles * (1 + bin_cap) words (counts, then lists); n_instances = periods * children of the scene.
cudaError_t launch_render_span(const RenderParams& P, int integrator, bool count, const unsigned char* d_nfine, const unsigned char* d_section,
                               unsigned int section_bytes, unsigned int* d_tile_list, unsigned int* d_tile_count, unsigned int* d_bins,
                               unsigned int bin_cap, unsigned int n_instances, cudaStream_t stream) {
    const size_t tiles = (size_t)P.n_views * P.tiles_i * P.tiles_j;
    if (tiles == 0) return cudaSuccess;
    if (tiles * 4 > 0xffffffffull) return cudaErrorInvalidValue;
    const size_t smem = span_kernel_smem_bytes(section_bytes);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    SpanArgs SA = {d_section, section_bytes, d_tile_list, d_tile_count, (unsigned int)(tiles * 4), span_list_cap(section_bytes), nullptr, nullptr, 0u};
    if (d_bins && bin_cap > 0 && n_instances > 0) {
        cudaError_t eb = cudaMemsetAsync(d_bins, 0, tiles * sizeof(unsigned int), stream);
        if (eb != cudaSuccess) return eb;
        const dim3 bgrid((n_instances + kBlockThreads / 32 - 1) / (kBlockThreads / 32), (unsigned int)P.n_views);
        span_bin_kernel<<<bgrid, kBlockThreads, 0, stream>>>(P, d_section, d_bins, d_bins + tiles, bin_cap);
        eb = cudaGetLastError();
        if (eb != cudaSuccess) return eb;
        SA.bin_counts = d_bins;
        SA.bin_lists = d_bins + tiles;
        SA.bin_cap = bin_cap;
    }
    cudaError_t e0 = cudaMemsetAsync(d_tile_count, 0, 2 * sizeof(unsigned int), stream);  // hand-over count, work counter
    if (e0 != cudaSuccess) return e0;
#define XR_SGO(I, C)                                                                                       \
    do {                                                                                                   \
        auto kern = render_span_kernel<I, C>;                                                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                    \
        /* persistent warps: exactly as many CTAs as are resident at once */                               \
        int occ = 0;                                                                                       \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlockThreads, smem);                \
        if (e != cudaSuccess) return e;                                                                    \
        const size_t resident = XR_SPAN_PERSISTENT ? (size_t)sms * (size_t)(occ > 0 ? occ : 1) : tiles;    \
        const unsigned int grid = (unsigned int)(tiles < resident ? tiles : resident);                     \
        kern<<<grid, kBlockThreads, smem, stream>>>(P, d_nfine, SA);                                       \
        return cudaGetLastError();                                                                         \
    } while (0)
    if (integrator == 0) {
        if (count) XR_SGO(0, true);
        else XR_SGO(0, false);
    } else {
        if (count) XR_SGO(1, true);
        else XR_SGO(1, false);
    }
#undef XR_SGO
    return cudaErrorInvalidValue;
}

}  // namespace xr
