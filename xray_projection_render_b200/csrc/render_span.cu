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
// `doubt` for the ray.  Three passes share the work:
//   fast pass    every warp tile; a ray with any doubt only flags its warp tile (no exact arithmetic compiled in);
//   settle pass  the flagged warp tiles again, with the reference's own expressions in its own operation order for
//                exactly the samples in doubt: samples inside a doubt zone (span_settle), rays INSIDE a bounding plane of
//                a child (a monotone step per plane: bisected), rays inside one or two cell-face planes of the tessellation
//                (span_degenerate_axis: the period step bisected per axis), rays with more intervals than the lists
//                hold (rendered in windows of the lattice), tiles whose screen-space bin overflowed (grid walk);
//   hand-over    what neither can vouch for -- two undecided planes at once, a plane constraint whose rounded expression
//                is not monotone, a ray parallel to and ON a cylinder's surface, a ray inside a face of the outer box, any
//                of the above under a warp -- is compacted into a list and rendered by the marching kernels.
// The result therefore equals the marching kernels' to ~1e-15 in T, in both precision modes.
//
// Doubt bits (stats[7] reports those of the rays that were handed over): 1 a ray inside a bounding plane (settled unless
// one of the cases above); 2 not settled here by construction (see hand-over); 4 a lattice sample inside a doubt zone
// that span_settle could not settle (guards, more than kSpanFuzzCap such candidates); 8 a second cell-face plane (settled
// when the tile has a bin); 16 the period along a degenerate axis is not a single step; 32 more intervals than the
// lists hold (settled in windows when the tile has a bin); 64 (internal) bin overflow seen by a fast pass without the walk.
#include <cstdio>
#include "eval.cuh"

namespace xr {

#ifndef XR_SPAN_MINBLOCKS
#define XR_SPAN_MINBLOCKS 4
#endif
#ifndef XR_SPAN_CAPMAX
#define XR_SPAN_CAPMAX 64
#endif
// Per-ray list capacity (candidates that survive the pre-filter = intervals at most): whatever fits the shared memory a CTA may
// use at XR_SPAN_MINBLOCKS CTAs per SM, at most 64.
constexpr int kSpanCapMax = XR_SPAN_CAPMAX;
constexpr int kSpanCapDense = 24;  // smallest list capacity worth running 5 CTAs per SM for
constexpr int kSpanFuzzCap = 4;  // candidates per ray set aside for exact settling
constexpr double kZone = 1.0e-11; // half-width (in s) of the doubt zone around every end point

struct SpanArgs {
    const unsigned char* section;  // device copy of the SpanHeader section
    unsigned int section_bytes;
    unsigned int* tile_list;       // out: warp tiles ((view, tile) id * 4 + warp position) with a ray that needs the marching kernels
    unsigned int* tile_count;      // out: [0] their number, [1] the number of settle_list entries (both zeroed before the launch)
    unsigned int* settle_list;     // out (fast pass) / in (settle pass): warp tiles whose only doubt is a lattice sample inside a doubt zone
    unsigned int total_items;      // warp tiles of the launch = views * tiles * 4
    int cap;                       // per-ray list capacity
    // screen-space bins (span_bin_kernel): per (view, tile) the (period, child) instances whose projection meets the tile
    const unsigned int* bin_counts;  // null = no bins: every ray walks the candidate grid
    const unsigned int* bin_lists;
    unsigned int bin_cap;
};

static int span_list_cap(unsigned int section_bytes, bool with_fuzz, int min_blocks = XR_SPAN_MINBLOCKS) {
    const size_t budget = (size_t)216 * 1024 / (size_t)min_blocks - 1024;  // 228 KB per SM, 1 KB per CTA reserved, some slack
    const size_t sec = (section_bytes + 15u) & ~15u;
    const size_t fz = with_fuzz ? (size_t)kSpanFuzzCap * kBlockThreads * 4 : 0;
    if (budget < sec + fz + (size_t)8 * kBlockThreads * 8) return 0;
    const size_t cap = (budget - sec - fz) / ((size_t)kBlockThreads * 8);
    return (int)(cap < (size_t)kSpanCapMax ? cap : (size_t)kSpanCapMax);
}

static size_t span_smem_bytes(unsigned int section_bytes, bool with_fuzz, int min_blocks = XR_SPAN_MINBLOCKS) {
    const int cap = span_list_cap(section_bytes, with_fuzz, min_blocks);
    if (cap <= 0) return (size_t)1 << 30;  // does not fit: the caller keeps to the marching kernels
    return (size_t)((section_bytes + 15u) & ~15u) + (size_t)cap * kBlockThreads * 8 + (with_fuzz ? (size_t)kSpanFuzzCap * kBlockThreads * 4 : 0);
}
size_t span_kernel_smem_bytes(unsigned int section_bytes) { return span_smem_bytes(section_bytes, true); }

// The set of ray parameters that satisfy every constraint seen so far is an interval whose ends are known up to a
// doubt zone each:  t > L1 surely passes every lower bound, t < L0 surely fails one;  t < U0 surely passes every upper
// bound, t > U1 surely fails one.  A constraint with lower bound a (error z) gives L0 = max(L0, a - z), L1 = max(L1, a + z).
struct SpanRange {
    double L0, L1, U0, U1;
};

// lo <= c0 + cd * t <= hi.  A coordinate that does not move over the window (|cd| < 1e-9, |t| <= 1.75) is decided from
// c0 alone, with doubt when it sits within 4e-9 of a bound (a ray running inside a face plane).  eps bounds the
// rounding error of the coordinate as the reference computes it.  Doubt bits: see render_span_kernel.
__device__ __forceinline__ bool span_slab(double c0, double cd, double inv_cd, double lo, double hi, double eps, SpanRange& r, unsigned int& doubt) {
    if (fabs(cd) < 1.0e-9) {
        const double dl = c0 - lo, dh = hi - c0;
        const bool nl = fabs(dl) < 4.0e-9 + eps, nh = fabs(dh) < 4.0e-9 + eps;
        if (nl || nh) {  // a ray inside a bounding plane: undecided here, so the constraint passes and the settle pass (or the
            doubt = (doubt | (nl && nh ? 3u : 1u)) + 0x100u;  // marching kernels) decide sample by sample.  Bits 8.. count these
            return true;                                      // constraints (span_settle needs to know whether there is just one)
        }
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

// A sum of products whose terms move in opposite directions along the ray need not be monotone once rounded: such a
// constraint cannot be bisected (span_settle) when the ray runs inside its plane.
__device__ __forceinline__ bool span_mixed(double a, double b, double c) { return (a > 0.0 || b > 0.0 || c > 0.0) && (a < 0.0 || b < 0.0 || c < 0.0); }

// The ray in object space, centred on the window (x(t) = c + d t, t = s - R), with what the slabs need.
struct SpanRay {
    double cx, cy, cz, dx, dy, dz, jdx, jdy, jdz, epsx;
};

// Parameter range of the ray inside child K of period (px, py, pz), intersected with R0 (the outer box): three period
// slabs (not along the axes of skip_axes, bit a = axis a: a degenerate axis, whose period is settled per ordinal), then the primitive.
template <bool MIXED>  // MIXED: also flag the plane constraints that span_settle must not bisect (span_mixed); the fast pass does not care
__device__ __forceinline__ bool span_candidate_range(const SpanHeader& H, const SpanChild& K, const SpanRay& y, const SpanRange& R0, int px, int py,
                                                     int pz, int skip_axes, SpanRange& r, unsigned int& doubt) {
    const double dx = y.dx, dy = y.dy, dz = y.dz, epsx = y.epsx;
    // ray centre in the coordinates of this period: x' = x - dx * n (objects.go:571)
    const double ux = y.cx - H.uc_d[0] * (double)px, uy = y.cy - H.uc_d[1] * (double)py, uz = y.cz - H.uc_d[2] * (double)pz;
    r = R0;
    bool ok = true;
    if (H.flags & SPAN_TESS) {  // the period's own cell: floor((x - min) / d) == n  <=>  min <= x' < min + d
        if (!(skip_axes & 1)) ok = ok && span_slab(ux, dx, y.jdx, H.uc_lo[0], H.uc_lo[0] + H.uc_d[0], epsx, r, doubt);
        if (!(skip_axes & 2)) ok = ok && span_slab(uy, dy, y.jdy, H.uc_lo[1], H.uc_lo[1] + H.uc_d[1], epsx, r, doubt);
        if (!(skip_axes & 4)) ok = ok && span_slab(uz, dz, y.jdz, H.uc_lo[2], H.uc_lo[2] + H.uc_d[2], epsx, r, doubt);
    }
    if (!ok) return false;
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
            const unsigned int d0 = doubt;
            ok = span_slab(wv * ivv, cd, 1.0 / cd, 0.0, 1.0, 1.0e-13, r, doubt);
            if (MIXED && doubt != d0 && span_mixed(dx * p[3], dy * p[4], dz * p[5])) doubt |= 2u;
        } else {
            C -= p[3];
        }
        ok = ok && span_quadratic(fmax(A, 0.0), B, C, r, doubt);
    } else if (K.type == OP_BOX) {
        ok = ok && span_slab(ux, dx, y.jdx, p[0] - p[3], p[0] + p[3], epsx, r, doubt);
        ok = ok && span_slab(uy, dy, y.jdy, p[1] - p[4], p[1] + p[4], epsx, r, doubt);
        ok = ok && span_slab(uz, dz, y.jdz, p[2] - p[5], p[2] + p[5], epsx, r, doubt);
    } else {  // parallelepiped: 0 < Minv (x - o) < 1 per component (objects.go:249-254)
        const double wx = ux - p[0], wy = uy - p[1], wz = uz - p[2];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            const double q0 = p[3 + 3 * rr] * wx + p[4 + 3 * rr] * wy + p[5 + 3 * rr] * wz;
            const double qd = p[3 + 3 * rr] * dx + p[4 + 3 * rr] * dy + p[5 + 3 * rr] * dz;
            const unsigned int d0 = doubt;
            ok = ok && span_slab(q0, qd, 1.0 / qd, 0.0, 1.0, epsx * (1.0 + p[12 + rr]), r, doubt);
            if (MIXED && doubt != d0 && span_mixed(dx * p[3 + 3 * rr], dy * p[4 + 3 * rr], dz * p[5 + 3 * rr])) doubt |= 2u;
        }
    }
    return ok;
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

// Exact position of a lattice ordinal, as the reference's loops produce it (repeated addition, main.go:147,183-189).
template <int INTEG>
__device__ __forceinline__ double span_exact_pos(const RenderParams& P, const unsigned char* __restrict__ nfine, unsigned int e) {
    if (INTEG == 0) return __ldg(P.s_tab + e);
    const unsigned int k = e >> 4, j = e & 15u;
    const unsigned int nf = (unsigned int)__ldg(nfine + k);
    if (j == 0u || j > nf) return __ldg(P.s_tab + k + (j == 0u ? 0u : 1u));  // not a fine sample: the neighbouring coarse one
    double s = __ldg(P.s_tab + k);
    for (unsigned int q = 0; q < j; ++q) s = dadd(s, P.ds_fine);
    return s;
}

template <int INTEG>
__device__ __forceinline__ unsigned int span_next_ordinal(const unsigned char* __restrict__ nfine, unsigned int e) {
    if (INTEG == 0) return e + 1u;
    const unsigned int k = e >> 4, j = e & 15u;
    if (j == 15u) return 16u * (k + 1u) + 1u;
    return j < (unsigned int)__ldg(nfine + k) ? e + 1u : 16u * k + 15u;
}

template <int INTEG>
__device__ __forceinline__ unsigned int span_prev_ordinal(const unsigned char* __restrict__ nfine, unsigned int e) {
    if (INTEG == 0) return e - 1u;
    const unsigned int k = e >> 4, j = e & 15u;
    if (j == 15u) {
        const unsigned int nf = (unsigned int)__ldg(nfine + k);
        return nf > 0u ? 16u * k + nf : 16u * (k - 1u) + 15u;
    }
    return j <= 1u ? 16u * (k - 1u) + 15u : e - 1u;
}

// Does the reference see lattice sample e of this ray inside child K of period (px, py, pz)?  The reference's own
// expressions in its own operation order (no FMA): position main.go:147-149, outer bounds / fold / unit-cell bounds
// objects.go:568-582, 458-464, primitive tests objects.go:63-72, 171-179, 247-255, 334-350.  Cold: only lattice samples
// that fall inside the doubt zone of an interval end point come here (un-warped scenes).
template <int INTEG>
__device__ __forceinline__ bool span_exact_member(const RenderParams& P, const unsigned char* __restrict__ nfine, const SpanHeader* H,
                                               const SpanChild* K, const double* __restrict__ eye, double dx, double dy, double dz,
                                               unsigned int e, int px, int py, int pz) {
    const double s = span_exact_pos<INTEG>(P, nfine, e);
    double x = dadd(eye[0], dmul(dx, s)), y = dadd(eye[1], dmul(dy, s)), z = dadd(eye[2], dmul(dz, s));
    if (H->flags & SPAN_TESS) {
        const double* o = H->outer;
        if (x < o[0] || x > o[3] || y < o[1] || y > o[4] || z < o[2] || z > o[5]) return false;
        const double nx = floor(ddiv(dsub(x, H->uc_lo[0]), H->uc_d[0])), ny = floor(ddiv(dsub(y, H->uc_lo[1]), H->uc_d[1])),
                     nz = floor(ddiv(dsub(z, H->uc_lo[2]), H->uc_d[2]));
        if ((int)nx != px || (int)ny != py || (int)nz != pz) return false;
        x = dsub(x, dmul(H->uc_d[0], nx));
        y = dsub(y, dmul(H->uc_d[1], ny));
        z = dsub(z, dmul(H->uc_d[2], nz));
        if (x < H->uc_lo[0] || x > H->uc_hi[0] || y < H->uc_lo[1] || y > H->uc_hi[1] || z < H->uc_lo[2] || z > H->uc_hi[2]) return false;
    }
    const double* p = K->p;
    if (K->type == OP_SPHERE) {
        const double ax = dsub(x, p[0]), ay = dsub(y, p[1]), az = dsub(z, p[2]);
        return dadd(dadd(dmul(ax, ax), dmul(ay, ay)), dmul(az, az)) < p[3];
    }
    if (K->type == OP_BOX) return fabs(dsub(x, p[0])) < p[3] && fabs(dsub(y, p[1])) < p[4] && fabs(dsub(z, p[2])) < p[5];
    if (K->type == OP_CYL) {
        const double wx = dsub(x, p[0]), wy = dsub(y, p[1]), wz = dsub(z, p[2]);
        const double wv = dadd(dadd(dmul(wx, p[3]), dmul(wy, p[4])), dmul(wz, p[5]));
        const double cc = ddiv(wv, p[8]);
        if (cc < 0.0 || cc > 1.0) return false;
        const double ex = dsub(wx, dmul(p[3], cc)), ey = dsub(wy, dmul(p[4], cc)), ez = dsub(wz, dmul(p[5], cc));
        return __dsqrt_rn(dadd(dadd(dmul(ex, ex), dmul(ey, ey)), dmul(ez, ez))) < p[9];
    }
    const double ax = dsub(x, p[0]), ay = dsub(y, p[1]), az = dsub(z, p[2]);
    const double qx = dadd(dadd(dmul(p[3], ax), dmul(p[4], ay)), dmul(p[5], az));
    const double qy = dadd(dadd(dmul(p[6], ax), dmul(p[7], ay)), dmul(p[8], az));
    const double qz = dadd(dadd(dmul(p[9], ax), dmul(p[10], ay)), dmul(p[11], az));
    return qx > 0.0 && qx < 1.0 && qy > 0.0 && qy < 1.0 && qz > 0.0 && qz < 1.0;
}

// Settle the lattice samples that lie inside the doubt zone of an interval's end point(s) one by one with the reference's own
// expressions: by convexity the first member from the lower side and the first non-member after it fix [e_in, e_out).
// Cold and out of line (called after the hot loop for the few candidates it set aside): recomputes the candidate's range.
// [e_first, e_last) = the ordinals to look at: the lattice window, cut to where the ray folds into this candidate's period
// along degenerate axes.  Returns doubt bits (0 = settled; e_in >= e_out = no lattice sample inside).
template <int INTEG>
__device__ __noinline__ unsigned int span_settle(const RenderParams& P, const unsigned char* __restrict__ nfine, const SpanHeader* H,
                                                 const SpanChild* ch, const double* __restrict__ eye, const SpanRay* ray, const SpanRange* R0,
                                                 unsigned int code, int skip_axes, unsigned int e_first, unsigned int e_last, unsigned int& e_in,
                                                 unsigned int& e_out) {
    const bool fuzzy = (code >> 31) != 0u;  // some constraint of this candidate is a ray inside its bounding plane (span_slab)
    const int c = (int)(code & 63u);
    const int px = (int)((code >> 6) & 31u) - 16, py = (int)((code >> 11) & 31u) - 16, pz = (int)((code >> 16) & 31u) - 16;
    const SpanChild* K = ch + c;
    const double inv_ds = 1.0 / P.ds, inv_dsf = 1.0 / P.ds_fine;
    SpanRange r;
    unsigned int bits = 0u, near_lo = 0u, near_hi = 0u, dummy = 0u;
    e_in = e_out = 0u;
    if (!span_candidate_range<true>(*H, *K, *ray, *R0, px, py, pz, skip_axes, r, bits)) return 0u;  // (surely outside some constraint)
    if (bits & 0xfeu) return bits & 0xffu;  // a case that is not settled here (a thin slab, a non-monotone plane constraint, ...)
    // bit 0 is flagged again (that is why we are here) and bits 8.. count the undecided constraints
    if (!fuzzy) bits = 0u;
    const double sc = P.s_center;
    const double dx = ray->dx, dy = ray->dy, dz = ray->dz;
    const bool sure = r.L1 <= r.U0;
    auto member = [&](unsigned int e) { return span_exact_member<INTEG>(P, nfine, H, K, eye, dx, dy, dz, e, px, py, pz); };
    if (fuzzy) {
        bits &= ~0xffu;
        // The ray runs inside a bounding plane of this child (span_slab): that constraint was left out of r.  Along the ray it
        // is a step function of the ordinal (monotone expressions: span_mixed cases never get here), so the members are the
        // ordinals of one interval still.  Core = the ordinals surely inside every other constraint: its members are found
        // from its two ends and a bisection; the doubt zones either side are looked at sample by sample.
        unsigned int e0 = max(span_ordinal_after<INTEG>(P, nfine, r.L0 + sc, 0.0, inv_ds, inv_dsf, dummy), e_first);
        unsigned int e1 = min(span_ordinal_after<INTEG>(P, nfine, r.U1 + sc, 0.0, inv_ds, inv_dsf, dummy), e_last);
        if (e0 >= e1) return bits;
        // first member in [from, to) and the first non-member after it (to if none); false when the guard trips
        auto scan_run = [&](unsigned int from, unsigned int to, unsigned int& lo, unsigned int& hi) -> bool {
            int guard = 0;
            unsigned int e = from;
            while (e < to && !member(e)) {
                e = span_next_ordinal<INTEG>(nfine, e);
                if (++guard > 24) return false;
            }
            lo = e;
            while (e < to && member(e)) {
                e = span_next_ordinal<INTEG>(nfine, e);
                if (++guard > 48) return false;
            }
            hi = e;
            return true;
        };
        unsigned int a = e0, b = e0;  // the core [a, b)
        if (sure) {
            a = min(max(span_ordinal_after<INTEG>(P, nfine, r.L1 + sc + 1.0e-13, 0.0, inv_ds, inv_dsf, dummy), e0), e1);
            // ordinals at positions < U0: up to (not including) the first ordinal at a position > U0 - tiny
            b = min(max(span_ordinal_after<INTEG>(P, nfine, r.U0 + sc - 1.0e-13, 0.0, inv_ds, inv_dsf, dummy), a), e1);
            while (b > a && span_exact_pos<INTEG>(P, nfine, span_prev_ordinal<INTEG>(nfine, b)) >= r.U0 + sc - 1.0e-13) b = span_prev_ordinal<INTEG>(nfine, b);
        }
        unsigned int lo = 0u, hi = 0u;
        if (a >= b) {  // no lattice sample in the core: the zones hold a handful at most
            if (!scan_run(e0, e1, lo, hi)) return bits | 4u;
            e_in = lo;
            e_out = hi;
            return bits;
        }
        const unsigned int last = span_prev_ordinal<INTEG>(nfine, b);
        const bool mA = member(a), mB = member(last);
        if (!mA && !mB) {
            // One undecided constraint: a step function that is false at both ends of the core is false all over it.  Two
            // (a ray along an edge) can leave members in the middle: not settled here.
            if ((bits >> 8) != 1u) return (bits & 0xffu) | 2u;
            if (!scan_run(e0, a, lo, hi)) return bits | 4u;  // members, if any, sit in one of the zones
            if (lo >= hi && !scan_run(b, e1, lo, hi)) return bits | 4u;
            e_in = lo;
            e_out = hi;
            return bits;
        }
        lo = a;
        hi = b;
        if (mA != mB) {
            unsigned int u = a, v = last;  // member(u) != member(v): bisect for the step
            while (v - u > 1u) {
                const unsigned int m = u + ((v - u) >> 1);
                if (member(m) == mA) u = m;
                else v = m;
            }
            // (ordinals 16 k + j with j beyond the interval's fine samples evaluate as its coarse sample, 16 k + 15)
            if (INTEG == 1 && (v & 15u) != 15u && ((v & 15u) == 0u || (v & 15u) > (unsigned int)__ldg(nfine + (v >> 4)))) {
                v = (v & 15u) == 0u ? v + 1u : (v & ~15u) + 15u;
                if ((v & 15u) == 1u && __ldg(nfine + (v >> 4)) == 0) v = (v & ~15u) + 15u;
            }
            if (mA) hi = v;
            else lo = v;
        }
        if (lo == a && e0 < a) {  // the lower zone: first member there, if any
            unsigned int zl = 0u, zh = 0u;
            if (!scan_run(e0, a, zl, zh)) return bits | 4u;
            if (zl < a) lo = zl;
        }
        if (hi == b && b < e1) {  // the upper zone: members directly after the core
            unsigned int e = b;
            int guard = 0;
            while (e < e1 && member(e)) {
                e = span_next_ordinal<INTEG>(nfine, e);
                if (++guard > 24) return bits | 4u;
            }
            hi = e;
        }
        e_in = lo;
        e_out = hi;
        return bits;
    }
    if (sure) {
        e_in = span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.L0 + r.L1) + sc, 0.5 * (r.L1 - r.L0), inv_ds, inv_dsf, near_lo);
        e_out = span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.U0 + r.U1) + sc, 0.5 * (r.U1 - r.U0), inv_ds, inv_dsf, near_hi);
    } else {
        e_in = e_out = span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.L0 + r.U1) + sc, 0.5 * (r.U1 - r.L0), inv_ds, inv_dsf, near_lo);
        near_hi = near_lo;
    }
    if (near_lo) {
        unsigned int e = max(span_ordinal_after<INTEG>(P, nfine, r.L0 + sc, 0.0, inv_ds, inv_dsf, dummy), e_first);  // first sample beyond L0
        const double stop = (sure ? r.L1 : r.U1) + sc + 1.0e-13;
        int guard = 0;
        while (e < e_last && span_exact_pos<INTEG>(P, nfine, e) <= stop) {
            if (member(e)) break;
            e = span_next_ordinal<INTEG>(nfine, e);
            if (++guard > 12) {
                bits |= 4u;
                break;
            }
        }
        e_in = e;
        if (!sure) e_out = e;  // (the upper loop starts here)
    }
    if (near_hi) {
        unsigned int e = sure ? span_ordinal_after<INTEG>(P, nfine, r.U0 + sc, 0.0, inv_ds, inv_dsf, dummy) : e_out;
        e = max(max(e, e_in), e_first);  // (e_first: samples before it fold into another period along a degenerate axis)
        const double stop = r.U1 + sc + 1.0e-13;
        int guard = 0;
        while (e < e_last && span_exact_pos<INTEG>(P, nfine, e) <= stop) {
            if (!member(e)) break;
            e = span_next_ordinal<INTEG>(nfine, e);
            if (++guard > 12) {
                bits |= 4u;
                break;
            }
        }
        e_out = e;
    }
    return bits;
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

// SETTLE = false: the fast pass over every warp tile of the launch; a ray with a lattice sample inside a doubt zone only
// flags its warp tile for the settle pass.  SETTLE = true: the same code over the flagged warp tiles, with the exact settling
// of those samples compiled in (span_settle; it costs the hot loop 15 % in registers and stack traffic, so the fast pass
// does not carry it).  Whatever neither pass can vouch for goes to the marching kernels.
// WALK = false (fast pass of a launch with screen-space bins): the grid walk is not compiled in -- it costs the binned hot
// loop 3-7 % in registers -- and the few tiles whose bin overflowed are left to the settle pass, which always has it.
// MINB = CTAs per SM the registers are budgeted for: 4 (128 registers), or 5 (96) for the binned fast pass when the scene's
// section leaves room for lists of kSpanCapDense entries in a fifth of the shared memory (+10 % on configs 2 and 5).
template <int INTEG, bool COUNT, bool SETTLE, bool WALK, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB) render_span_kernel(const RenderParams P, const unsigned char* __restrict__ nfine,
                                                                       const SpanArgs SA) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    if (SETTLE && blockIdx.x * (kBlockThreads / 32) >= SA.tile_count[1]) return;  // (usually a handful of warp tiles: most CTAs leave here)
    {
        uint4* dst = reinterpret_cast<uint4*>(smem);
        const uint4* src = reinterpret_cast<const uint4*>(SA.section);
        const int n16 = (int)((SA.section_bytes + 15u) >> 4);
        for (int k = tid; k < n16; k += kBlockThreads) dst[k] = __ldg(src + k);
    }
    if (!SETTLE && COUNT && P.stats && blockIdx.x == 0 && tid == 0) atomicOr(P.stats + 7, 0x10000ull);  // "the interval renderer ran"
    const SpanHeader& H = *reinterpret_cast<const SpanHeader*>(smem);
    unsigned int* cand = reinterpret_cast<unsigned int*>(smem + ((SA.section_bytes + 15u) & ~15u));
    // interval q is written after candidate q has been read and there are never more intervals than candidates read, so
    // the intervals' first words live in the candidate slots
    unsigned int* iv_in = cand;
    const int cap = SA.cap;
    unsigned int* iv_out = cand + cap * kBlockThreads;
    unsigned int* fuzz = iv_out + cap * kBlockThreads;  // settle pass only
    __syncthreads();
    const SpanChild* __restrict__ ch = reinterpret_cast<const SpanChild*>(smem + H.child_off);
    const unsigned long long* __restrict__ masks = reinterpret_cast<const unsigned long long*>(smem + H.mask_off);
    const unsigned int flags = H.flags;
    const bool tess = (flags & SPAN_TESS) != 0u;

    // One warp = one warp tile (32 pixels, the same 4 x 8 footprint the marching kernels use); nothing below synchronises
    // beyond the warp.  The settle pass strides over its list.
    const unsigned int n_items = SETTLE ? SA.tile_count[1] : SA.total_items;
    for (unsigned int wi = blockIdx.x * (kBlockThreads / 32) + (tid >> 5); wi < n_items; wi = SETTLE ? wi + gridDim.x * (kBlockThreads / 32) : n_items) {
    const unsigned int item = SETTLE ? SA.settle_list[wi] : wi;
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
    // A ray inside a face of the outer box: every sample's outer test is undecided (TESS: left to the marching kernels);
    // the region of a flat collection is only a bound of ours, the children decide for themselves.
    if (doubt) doubt = tess ? 2u : 0u;
    const double ta = R0.L0, tb = R0.U1;
    const double inv_ds = 1.0 / P.ds, inv_dsf = 1.0 / P.ds_fine;

    // ---- a ray inside a cell-face plane of the tessellation: exact period per sample along that axis ----
    int deg_axis = -1, deg_axis2 = -1, deg_m = 0, skip_axes = 0;
    SpanDegenerate dg = {0, 0, 0u, 0u, 0u}, dg2 = {0, 0, 0u, 0u, 0u};
    if (tess) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double da = a == 0 ? dx : (a == 1 ? dy : dz), ca = a == 0 ? cx : (a == 1 ? cy : cz);
            if (hit && fabs(da) < 1.0e-9) {
                const double q = (ca - H.uc_lo[a]) / H.uc_d[a], m = rint(q);
                if (fabs(q - m) * fabs(H.uc_d[a]) < 6.0e-9 + epsx) {
                    if (flags & SPAN_HAS_WARP) doubt |= 8u;  // a warp between the ray and the fold: not settled here
                    else if (deg_axis >= 0) {  // a second such axis (a ray along a cell edge): the settle pass gives it the same
                        if (SETTLE && SA.bin_counts) {  // treatment; it takes its candidates from the tile's bin
                            deg_axis2 = a;
                            skip_axes |= 1 << a;
                        } else {
                            doubt |= 8u;
                        }
                    } else {
                        deg_axis = a;
                        deg_m = (int)m;
                        skip_axes |= 1 << a;
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
        if (SETTLE && __any_sync(FULL_MASK, deg_axis2 >= 0)) {
            dg2.eA = dg.eA;
            dg2.eB = dg.eB;
            const int a = max(deg_axis2, 0);
            const double da = a == 0 ? dx : (a == 1 ? dy : dz);
            const bool ok = span_degenerate_axis<INTEG>(P, nfine, P.cams[view].eye[a], da, H.uc_lo[a], H.uc_hi[a], H.uc_d[a], dg2);
            if (deg_axis2 >= 0 && !ok) doubt |= 16u;
        }
    }
#ifdef XRAY_DEV_KNOBS
    const bool dbg_px = P.dbg_cause >= 1000000 && i == (P.dbg_cause - 1000000) / 1000 && j == (P.dbg_cause % 1000) && valid;
    if (dbg_px)
        printf("[span %d] px (%d,%d) c (%.17g %.17g %.17g) d (%.17g %.17g %.17g) hit %d doubt %u R0 [%g %g %g %g] deg %d %d skip %d dg n %d %d e %u %u sw %u dg2 n %d %d sw %u\n",
               (int)SETTLE, i, j, cx, cy, cz, dx, dy, dz, (int)hit, doubt, R0.L0, R0.L1, R0.U0, R0.U1, deg_axis, deg_axis2, skip_axes, dg.nA, dg.nB, dg.eA,
               dg.eB, dg.e_sw, dg2.nA, dg2.nB, dg2.e_sw);
#endif
    // ordinal window of a period along the degenerate axes; false = the ray never folds into that period
    auto deg_window = [&](int px, int py, int pz, unsigned int& w_lo, unsigned int& w_hi) -> bool {
        if (deg_axis >= 0) {
            const int pa = deg_axis == 0 ? px : (deg_axis == 1 ? py : pz);
            if (pa == dg.nA) {
                w_lo = max(w_lo, dg.eA);
                w_hi = min(w_hi, dg.nA == dg.nB ? dg.eB : dg.e_sw);
            } else if (pa == dg.nB) {
                w_lo = max(w_lo, dg.e_sw);
                w_hi = min(w_hi, dg.eB);
            } else {
                return false;
            }
        }
        if (SETTLE && deg_axis2 >= 0) {
            const int pa = deg_axis2 == 0 ? px : (deg_axis2 == 1 ? py : pz);
            if (pa == dg2.nA) {
                w_lo = max(w_lo, dg2.eA);
                w_hi = min(w_hi, dg2.nA == dg2.nB ? dg2.eB : dg2.e_sw);
            } else if (pa == dg2.nB) {
                w_lo = max(w_lo, dg2.e_sw);
                w_hi = min(w_hi, dg2.eB);
            } else {
                return false;
            }
        }
        return true;
    };

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
        const float disc = B * B - A * C, tol = 2.0e-6f * (1.0f + B * B);
        if (disc < -tol) pass = false;
        else if (A > 1.0e-3f * fA) {  // (a ray within 2 degrees of a cylinder's axis: A and B are rounding noise in fp32 -- no range test)
            const float iA = 1.0f / A;
            const float th = sqrtf(disc + tol) * iA + 1.0e-3f, tm = -B * iA;
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
    unsigned int nb = 0u;                        // entries of this tile's bin (binned mode)
    const unsigned int* __restrict__ bl = nullptr;
    if (SA.bin_counts) {  // uniform
        nb = __ldg(SA.bin_counts + (item >> 2));
        if (!WALK && nb > SA.bin_cap) {  // (a bin that overflowed is not used: those tiles walk -- in the settle pass)
            walk = false;
            if (hit) doubt |= 64u;
        }
        if (nb <= SA.bin_cap) {
            walk = false;
            bl = SA.bin_lists + (size_t)(item >> 2) * SA.bin_cap;
            const float ucdx = (float)H.uc_d[0], ucdy = (float)H.uc_d[1], ucdz = (float)H.uc_d[2];
            const float ulx = H.f_uc_lo[0], uly = H.f_uc_lo[1], ulz = H.f_uc_lo[2];
            for (unsigned int q = 0; q < nb; ++q) {
                const unsigned int code = __ldg(bl + q);
                const int px = (int)((code >> 6) & 31u) - 16, py = (int)((code >> 11) & 31u) - 16, pz = (int)((code >> 16) & 31u) - 16;
                if (deg_axis >= 0) {  // a ray inside a face plane only ever folds into the two periods either side of it
                    const int pa = deg_axis == 0 ? px : (deg_axis == 1 ? py : pz);
                    if (pa != dg.nA && pa != dg.nB) continue;
                }
                if (SETTLE && deg_axis2 >= 0) {
                    const int pa = deg_axis2 == 0 ? px : (deg_axis2 == 1 ? py : pz);
                    if (pa != dg2.nA && pa != dg2.nB) continue;
                }
                if (hit)
                    test_and_push((int)(code & 63u), (int)(code & ~63u), fcx - (float)px * ucdx + ulx, fcy - (float)py * ucdy + uly,
                                  fcz - (float)pz * ucdz + ulz);
            }
        }
    }
    if (SETTLE && deg_axis2 >= 0 && bl == nullptr) doubt |= 8u;  // (the walk knows one such axis only)
#ifdef XRAY_DEV_KNOBS
    if (dbg_px) printf("[span %d] bins: nb %u bl %d ncand %d overflow %d\n", (int)SETTLE, nb, (int)(bl != nullptr), ncand, (int)overflow);
#endif
    // The walk runs in fp32 (position error a few 1e-6) while the exact ray may sit on the other side of a cell face -- for good,
    // when it runs inside or alongside one.  So every stretch of the walk also visits the neighbours whose face the fp32
    // position comes closer to than tie_d (side cells; across a period boundary that is the neighbouring period's cell,
    // whose children the dilated masks of this side do not list).
    const int max_pass = (WALK && __any_sync(FULL_MASK, walk)) ? 1 : 0;
    for (int pass = 0; pass < max_pass; ++pass) {
        if (!walk) continue;
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
        if (deg_axis == 0) ix = deg_m * gx;  // first cell of period m: the ray runs inside its lower face (the side visits take the other side)
        if (deg_axis == 1) iy = deg_m * gy;
        if (deg_axis == 2) iz = deg_m * gz;
        const float tie_d = 0.25f * slack;
        // children already pushed, per period: the period the walk is in, and up to three neighbouring ones seen by side visits
        // (a child pushed twice would only yield the same interval twice, but a ray that runs along a face would do that at every step)
        unsigned long long done = 0ull, done1 = 0ull, done2 = 0ull, done3 = 0ull;
        int main_code = -1, key1 = -1, key2 = -1, key3 = -1, slot = 0;

        auto visit = [&](int lx, int ly, int lz, int px, int py, int pz) {
            if (!tess && (px | py | pz) != 0) return;  // outside the region
            unsigned long long nw = masks[(lz * gy + ly) * gx + lx];
            if (nw == 0ull) return;
            const int pcode = ((px + 16) << 6) | ((py + 16) << 11) | ((pz + 16) << 16);  // |p| <= 15: scene_compile.cpp build_span
            if (pcode == main_code) { nw &= ~done; done |= nw; }
            else if (pcode == key1) { nw &= ~done1; done1 |= nw; }
            else if (pcode == key2) { nw &= ~done2; done2 |= nw; }
            else if (pcode == key3) { nw &= ~done3; done3 |= nw; }
            else {
                if (slot == 0) { key1 = pcode; done1 = nw; }
                else if (slot == 1) { key2 = pcode; done2 = nw; }
                else { key3 = pcode; done3 = nw; }
                slot = slot == 2 ? 0 : slot + 1;
            }
            if (nw == 0ull) return;
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
        main_code = ((px + 16) << 6) | ((py + 16) << 11) | ((pz + 16) << 16);
        // parameter at which the ray leaves the cell along each axis, from the absolute cell index (no drift)
        // (an axis the ray does not move along -- exactly, or to 1e-9 inside a face plane -- is never stepped)
        float tx = (fdx != 0.0f && deg_axis != 0) ? (((float)(ix + (sx > 0 ? 1 : 0))) * csx - fcx) * ifx : 3.0e38f;
        float ty = (fdy != 0.0f && deg_axis != 1) ? (((float)(iy + (sy > 0 ? 1 : 0))) * csy - fcy) * ify : 3.0e38f;
        float tz = (fdz != 0.0f && deg_axis != 2) ? (((float)(iz + (sz > 0 ? 1 : 0))) * csz - fcz) * ifz : 3.0e38f;
        // two faces count as reached together when the second is closer than tie_d (in space) at that moment
        const float hx = (fdx != 0.0f && deg_axis != 0) ? tie_d * fabsf(ifx) : 0.0f, hy = (fdy != 0.0f && deg_axis != 1) ? tie_d * fabsf(ify) : 0.0f,
                    hz = (fdz != 0.0f && deg_axis != 2) ? tie_d * fabsf(ifz) : 0.0f;
        float t_prev = t_begin;
        for (int guard = 0; guard < 4096; ++guard) {
            visit(lx, ly, lz, px, py, pz);
            const float tm = fminf(tx, fminf(ty, tz));
            // the stretch [t_prev, tm] of the ray against the faces of this cell
            {
                const float te = fminf(tm, t_end);
                const float x0 = fmaf(fdx, t_prev, fcx), x1 = fmaf(fdx, te, fcx), y0 = fmaf(fdy, t_prev, fcy), y1 = fmaf(fdy, te, fcy),
                            z0 = fmaf(fdz, t_prev, fcz), z1 = fmaf(fdz, te, fcz);
                const float lox = (float)ix * csx, loy = (float)iy * csy, loz = (float)iz * csz;
                const int mx = fminf(x0, x1) - lox < tie_d ? -1 : 0, Mx = lox + csx - fmaxf(x0, x1) < tie_d ? 1 : 0;
                const int my = fminf(y0, y1) - loy < tie_d ? -1 : 0, My = loy + csy - fmaxf(y0, y1) < tie_d ? 1 : 0;
                const int mz = fminf(z0, z1) - loz < tie_d ? -1 : 0, Mz = loz + csz - fmaxf(z0, z1) < tie_d ? 1 : 0;
                if ((mx | Mx | my | My | mz | Mz) != 0) {
                    for (int oz = mz; oz <= Mz; ++oz)
                        for (int oy = my; oy <= My; ++oy)
                            for (int ox = mx; ox <= Mx; ++ox) {
                                if ((ox | oy | oz) == 0) continue;
                                int vx = lx + ox, vy = ly + oy, vz = lz + oz, qx = px, qy = py, qz = pz;
                                if (vx == gx) { vx = 0; ++qx; } else if (vx < 0) { vx = gx - 1; --qx; }
                                if (vy == gy) { vy = 0; ++qy; } else if (vy < 0) { vy = gy - 1; --qy; }
                                if (vz == gz) { vz = 0; ++qz; } else if (vz < 0) { vz = gz - 1; --qz; }
                                visit(vx, vy, vz, qx, qy, qz);
                            }
                }
            }
            if (!(tm <= t_end)) break;
            const bool nx = tx - tm < hx, ny = ty - tm < hy, nz = tz - tm < hz;
            bool moved = false;  // into another period
            if (nx) {
                ix += sx;
                lx += sx;
                if (lx == gx) { lx = 0; ++px; moved = true; } else if (lx < 0) { lx = gx - 1; --px; moved = true; }
                tx = (((float)(ix + (sx > 0 ? 1 : 0))) * csx - fcx) * ifx;
            }
            if (ny) {
                iy += sy;
                ly += sy;
                if (ly == gy) { ly = 0; ++py; moved = true; } else if (ly < 0) { ly = gy - 1; --py; moved = true; }
                ty = (((float)(iy + (sy > 0 ? 1 : 0))) * csy - fcy) * ify;
            }
            if (nz) {
                iz += sz;
                lz += sz;
                if (lz == gz) { lz = 0; ++pz; moved = true; } else if (lz < 0) { lz = gz - 1; --pz; moved = true; }
                tz = (((float)(iz + (sz > 0 ? 1 : 0))) * csz - fcz) * ifz;
            }
            if (moved) {  // keep what is known about the period walked into, if a side visit has seen it
                const int nc = ((px + 16) << 6) | ((py + 16) << 11) | ((pz + 16) << 16);
                const unsigned long long keep = nc == key1 ? done1 : (nc == key2 ? done2 : (nc == key3 ? done3 : 0ull));
                // the period left behind takes that slot (or the next one to be replaced)
                if (nc == key1) { key1 = main_code; done1 = done; }
                else if (nc == key2) { key2 = main_code; done2 = done; }
                else if (nc == key3) { key3 = main_code; done3 = done; }
                else {
                    if (slot == 0) { key1 = main_code; done1 = done; }
                    else if (slot == 1) { key2 = main_code; done2 = done; }
                    else { key3 = main_code; done3 = done; }
                    slot = slot == 2 ? 0 : slot + 1;
                }
                main_code = nc;
                done = keep;
            }
            t_prev = tm;
            if (guard == 4095) overflow = true;
        }
    }
    // More survivors than the list holds: with a bin at hand the ray simply takes every entry of the bin through the exact
    // stage (no list needed); a walking ray is handed over.
    // More survivors than the list holds: with a bin at hand the ray takes every entry of the bin through the exact stage
    // instead (no list needed); a walking ray is handed over.
    bool redo = overflow && bl != nullptr;
    if (redo) overflow = false;
    if (overflow || redo) ncand = 0;

    // ---- phase 2: exact interval of every surviving (period, child), mapped to lattice ordinals ----
    // (the reciprocals are recomputed here rather than kept alive across the walk: 6 registers for 3 divisions per ray)
    const SpanRay ray = {cx, cy, cz, dx, dy, dz, 1.0 / dx, 1.0 / dy, 1.0 / dz, epsx};  // (its address is taken in the settle pass only)
    const unsigned int e_first = INTEG == 1 ? 1u : 0u;
    const unsigned int e_last = INTEG == 1 ? 16u * (unsigned int)P.n_steps : (unsigned int)P.n_steps;  // one past the last ordinal
    unsigned int prim_tests = 0;
    double T = 0.0;
    unsigned int n_fine = 0;
    int niv = 0;
    // A ray that crosses more primitives than its interval list holds (rays along a row of lattice nodes) is rendered in
    // WINDOWS of the lattice: the settle pass retries with 2, 4, ... 16 windows (aligned to coarse intervals; the sweep's
    // state -- zero-ness of the last coarse sample -- carries over), candidates straight from the tile's bin.  The fast
    // pass makes one attempt with one window and flags the warp tile otherwise.
    unsigned int n_win = 1u;
    for (;;) {
    T = 0.0;
    n_fine = 0;
    prim_tests = 0;
    bool prev_z = false;  // prev_rho := 0.0 (main.go:179)
    double facc = 0.0;    // sum of the fine samples seen so far in the coarse interval under way
    bool ovf = overflow;
    for (unsigned int win = 0; win < n_win; ++win) {
    unsigned int win_lo = e_first, win_hi = e_last;
    if (SETTLE && n_win > 1u) {
        const unsigned int span_e = e_last - e_first;
        win_lo = e_first + (unsigned int)(((unsigned long long)span_e * win) / n_win);
        win_hi = e_first + (unsigned int)(((unsigned long long)span_e * (win + 1u)) / n_win);
        if (INTEG == 1) {  // windows start on the first fine sample of a coarse interval: ordinals 16 k + 1
            win_lo = ((win_lo - 1u) & ~15u) + 1u;
            win_hi = win + 1u == n_win ? e_last : ((win_hi - 1u) & ~15u) + 1u;
        }
    }
    niv = 0;
    int nfz = 0;
    const int my_cand = redo ? (int)nb : ncand;
    const int max_cand = __reduce_max_sync(FULL_MASK, my_cand);
    for (int q = 0; q < max_cand; ++q) {
        if (q >= my_cand) continue;
        const unsigned int code = redo ? __ldg(bl + q) : cand[q * kBlockThreads + tid];
        const int c = (int)(code & 63u);
        const int px = (int)((code >> 6) & 31u) - 16, py = (int)((code >> 11) & 31u) - 16, pz = (int)((code >> 16) & 31u) - 16;
        if (COUNT) ++prim_tests;
        unsigned int w_lo = e_first, w_hi = e_last;  // ordinal window of this period along a degenerate axis
        if (!deg_window(px, py, pz, w_lo, w_hi)) continue;
        const SpanChild& K = ch[c];
        SpanRange r;
        if (SETTLE) {
            unsigned int cb = 0u;  // this candidate's doubts
            const bool some = span_candidate_range<true>(H, K, ray, R0, px, py, pz, skip_axes, r, cb);
            if (some && (cb & 0xffu) == 1u && nfz < kSpanFuzzCap && !(flags & SPAN_HAS_WARP)) {  // the ray runs inside a plane of this child: settled below
                fuzz[(nfz++) * kBlockThreads + tid] = code | 0x80000000u;
                continue;
            }
            doubt |= cb & 0xffu;
            if (!some) continue;
        } else {  // (the fast pass only flags; bits 8.. of doubt are masked off at the end)
            if (!span_candidate_range<false>(H, K, ray, R0, px, py, pz, skip_axes, r, doubt)) continue;
        }
#ifdef XRAY_DEV_KNOBS
        if (dbg_px) printf("[span %d]   cand c %d p (%d %d %d) w [%u %u) r [%.12g %.12g %.12g %.12g] doubt %u\n", (int)SETTLE, c, px, py, pz, w_lo, w_hi, r.L0, r.L1, r.U0, r.U1, doubt);
#endif
        // End points -> lattice ordinals.  Common case: no lattice sample lies inside either doubt zone.  A candidate with one
        // that does is set aside (fuzz list) and settled after this loop, sample by sample, with the reference's own expressions.
        const double sc = P.s_center;
        unsigned int near = 0u;
        unsigned int e_in, e_out;
        if (r.L1 <= r.U0) {  // some parameter is surely inside
            e_in = span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.L0 + r.L1) + sc, 0.5 * (r.L1 - r.L0), inv_ds, inv_dsf, near);
            e_out = span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.U0 + r.U1) + sc, 0.5 * (r.U1 - r.U0), inv_ds, inv_dsf, near);
        } else {
            e_in = e_out = span_ordinal_after<INTEG>(P, nfine, 0.5 * (r.L0 + r.U1) + sc, 0.5 * (r.U1 - r.L0), inv_ds, inv_dsf, near);
        }
        if (near) {
            if (!SETTLE || (flags & SPAN_HAS_WARP) || nfz >= kSpanFuzzCap) doubt |= 4u;  // (fast pass: flag it; a warp: cannot be settled here)
            else fuzz[(nfz++) * kBlockThreads + tid] = code;
            continue;
        }
        e_in = max(max(w_lo, win_lo), e_in);
        e_out = min(min(w_hi, win_hi), e_out);
        if (e_in >= e_out) continue;  // no lattice sample inside (this window)
        if (niv < cap) {
            iv_in[niv * kBlockThreads + tid] = e_in | ((unsigned int)c << 26);
            iv_out[niv * kBlockThreads + tid] = e_out;
        } else {
            ovf = true;
        }
        ++niv;
    }
    if (SETTLE && __any_sync(FULL_MASK, nfz > 0)) {
        for (int f = 0; f < nfz; ++f) {
            const unsigned int code = fuzz[f * kBlockThreads + tid];
            unsigned int e_in = 0u, e_out = 0u;
            // the ordinals in which the ray folds into this candidate's period along the degenerate axes (span_settle's
            // bisection needs that settled: inside this range only the plane constraint can step), and the lattice window
            unsigned int w_lo = win_lo, w_hi = win_hi;
            deg_window((int)((code >> 6) & 31u) - 16, (int)((code >> 11) & 31u) - 16, (int)((code >> 16) & 31u) - 16, w_lo, w_hi);
            if (w_lo >= w_hi) continue;
            doubt |= span_settle<INTEG>(P, nfine, &H, ch, P.cams[view].eye, &ray, &R0, code, skip_axes, w_lo, w_hi, e_in, e_out) & 0xffu;
#ifdef XRAY_DEV_KNOBS
            if (dbg_px) printf("[span]   settled code %x w [%u %u) -> [%u %u) doubt %u\n", code, w_lo, w_hi, e_in, e_out, doubt);
#endif
            e_in = max(w_lo, e_in);
            e_out = min(w_hi, e_out);
            if (e_in >= e_out) continue;
            if (niv < cap) {
                iv_in[niv * kBlockThreads + tid] = e_in | ((code & 63u) << 26);
                iv_out[niv * kBlockThreads + tid] = e_out;
            } else {
                ovf = true;
            }
            ++niv;
        }
    }
    if (ovf) niv = 0;

    // ---- phase 3: integer sweep over the end points = the reference's loop, piece by piece ----
    if (!ovf && (niv > 0 || prev_z)) {
        const double dm = P.dm, DS = P.ds, dsf = P.ds_fine;
        // nothing before the first interval (unless the previous window ended inside material) and nothing after the last
        // one: both stretches add 0 and flip nothing
        unsigned int e = win_hi, e_stop = win_lo;
        for (int q = 0; q < niv; ++q) {
            e = min(e, iv_in[q * kBlockThreads + tid] & 0x3ffffffu);
            e_stop = max(e_stop, iv_out[q * kBlockThreads + tid]);
        }
        e = prev_z ? win_lo : max(e, win_lo);
        e_stop = min(max(e_stop, e) + (INTEG == 1 ? 32u : 0u), win_hi);  // (+2 coarse intervals: the exit transition's refinement)
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
    }  // windows
    overflow = ovf;
    if (!SETTLE) break;
    if (!__any_sync(FULL_MASK, ovf && bl != nullptr && hit) || n_win >= 16u) break;
    n_win *= 2u;   // (warp-uniform: every lane of the warp goes again, lanes that were fine just repeat their work)
    redo = true;   // candidates straight from the bin: the candidate list shares its slots with the intervals and is gone
    ncand = 0;
    overflow = false;
    }  // attempts
    if (overflow) doubt |= 32u;
    doubt &= 0xffu;
    const bool bad = valid && doubt != 0u;
    const bool tile_bad = __any_sync(FULL_MASK, bad);
    // fast pass: a warp tile whose only trouble is samples inside doubt zones goes to the settle pass, not to the marching kernels
    // (samples inside doubt zones; more intervals than the list holds, when the tile has a bin to take the candidates from)
    // (under a warp nothing is settled exactly, but a tile whose bin overflowed still gets its grid walk there)
    const bool settleable = (flags & SPAN_HAS_WARP) ? (doubt & ~64u) == 0u
                                                    : ((doubt & ~(1u | 4u | 8u | 32u | 64u)) == 0u && (!(doubt & (8u | 32u)) || bl != nullptr));
    const bool to_settle = !SETTLE && tile_bad && !__any_sync(FULL_MASK, bad && !settleable);
    if ((tid & 31) == 0 && tile_bad) {
        if (to_settle) {
            SA.settle_list[atomicAdd(SA.tile_count + 1, 1u)] = item;
        } else {
            SA.tile_list[atomicAdd(SA.tile_count, 1u)] = item;
            if (COUNT && P.stats) atomicAdd(P.stats + 6, 1ull);  // warp tiles handed to the marching kernels
        }
    }
    if (COUNT && P.stats && bad && !to_settle) atomicOr(P.stats + 7, (unsigned long long)(doubt & 0x3fu));
#ifdef XRAY_DEV_KNOBS
    if (P.dbg_cause == 77) {  // development builds: show which rays are handed over, and why
        store_pixel(P, view, i, j, valid, bad && !to_settle ? -(double)doubt : exp(-(P.flat_field + T)));
        continue;
    }
#endif
    store_pixel(P, view, i, j, valid, exp(-(P.flat_field + T)));
    if (COUNT && !tile_bad)
        add_stats(P, valid ? (unsigned long long)P.n_steps + n_fine : 0ull, (unsigned long long)niv, 0ull, prim_tests, valid ? 1ull : 0ull);
    }  // warp tiles
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
    float R = K.bs[3];
    if (H.flags & SPAN_HAS_WARP) {  // the children live in warped space: back to the world the camera sees (conservative radius)
        const float* w = H.f_winv;
        const float b0 = (float)H.warp_b[0], b1 = (float)H.warp_b[1], b2 = (float)H.warp_b[2];
        const float ux = ax - b0, uy = ay - b1, uz = az - b2, vx = bx - b0, vy = by - b1, vz = bz - b2;
        ax = w[0] * ux + w[1] * uy + w[2] * uz; ay = w[3] * ux + w[4] * uy + w[5] * uz; az = w[6] * ux + w[7] * uy + w[8] * uz;
        bx = w[0] * vx + w[1] * vy + w[2] * vz; by = w[3] * vx + w[4] * vy + w[5] * vz; bz = w[6] * vx + w[7] * vy + w[8] * vz;
        R = R * H.f_wscale + 1.0e-5f;
    }
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
    const bool dense = d_bins && bin_cap > 0 && n_instances > 0 && span_list_cap(section_bytes, false, XR_SPAN_MINBLOCKS + 1) >= kSpanCapDense;
    const int fast_blocks = dense ? XR_SPAN_MINBLOCKS + 1 : XR_SPAN_MINBLOCKS;
    const size_t smem_fast = span_smem_bytes(section_bytes, false, fast_blocks), smem_settle = span_smem_bytes(section_bytes, true);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // work lists: d_tile_list[0 .. items) = hand-over to the marching kernels, [items .. 2 items) = settle pass
    SpanArgs SA = {d_section, section_bytes, d_tile_list, d_tile_count, d_tile_list + tiles * 4, (unsigned int)(tiles * 4),
                   span_list_cap(section_bytes, false, fast_blocks), nullptr, nullptr, 0u};
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
    cudaError_t e0 = cudaMemsetAsync(d_tile_count, 0, 2 * sizeof(unsigned int), stream);  // hand-over count, settle count
    if (e0 != cudaSuccess) return e0;
    SpanArgs SB = SA;
    SB.cap = span_list_cap(section_bytes, true);
    const unsigned int grid_settle = (unsigned int)(tiles < (size_t)sms ? tiles : (size_t)sms);
#define XR_SGO(I, C)                                                                                                        \
    do {                                                                                                                    \
        auto fast = dense ? render_span_kernel<I, C, false, false, XR_SPAN_MINBLOCKS + 1>                                   \
                          : (SA.bin_counts ? render_span_kernel<I, C, false, false, XR_SPAN_MINBLOCKS>                      \
                                           : render_span_kernel<I, C, false, true, XR_SPAN_MINBLOCKS>);                     \
        auto settle = render_span_kernel<I, C, true, true, XR_SPAN_MINBLOCKS>;                                              \
        cudaError_t e = cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast);            \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(settle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_settle); \
        if (e != cudaSuccess) return e;                                                                                     \
        fast<<<(unsigned int)tiles, kBlockThreads, smem_fast, stream>>>(P, d_nfine, SA);                                    \
        settle<<<grid_settle, kBlockThreads, smem_settle, stream>>>(P, d_nfine, SB);                                        \
        return cudaGetLastError();                                                                                          \
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
