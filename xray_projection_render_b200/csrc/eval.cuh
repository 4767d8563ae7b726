// eval.cuh -- device-side scene interpreter shared by the render / voxelise kernels.
// (see render_scene.cu for the design notes)
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "device_types.h"

namespace xr {

#define FULL_MASK 0xffffffffu

// ---------------------------------------------------------------------------------------
// Contraction-free fp64 helpers: Go/amd64 rounds every * and + separately.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

struct SceneView {  // where the kernel reads the program from (shared memory or global)
    const Instr* instr;
    const float4* f32;
    const double* f64;
    const unsigned long long* grids;
    const DeformRec* deform;
    int n_instr, n_deform;
    const VoxelDev* vox;
};

// objects.go:789-855 VoxelGrid.Density, reference operation order.
__device__ __forceinline__ double voxel_exact(const VoxelDev& v, double x, double y, double z) {
        if (!v.data) return 0.0;
        if (x < -1 || x > 1 || y < -1 || y > 1 || z < -1 || z > 1) return 0.0;
        x = ddiv(dadd(x, 1.0), 2.0);
        y = ddiv(dadd(y, 1.0), 2.0);
        z = ddiv(dadd(z, 1.0), 2.0);
        x = dmul(x, (double)(v.nx - 1));
        y = dmul(y, (double)(v.ny - 1));
        z = dmul(z, (double)(v.nz - 1));
        long long x0 = (long long)floor(x), y0 = (long long)floor(y), z0 = (long long)floor(z);
        long long x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
        if (x0 < 0) x0 = 0;
        if (y0 < 0) y0 = 0;
        if (z0 < 0) z0 = 0;
        if (x1 >= v.nx) x1 = v.nx - 1;
        if (y1 >= v.ny) y1 = v.ny - 1;
        if (z1 >= v.nz) z1 = v.nz - 1;
        double wx = dsub(x, (double)x0), wy = dsub(y, (double)y0), wz = dsub(z, (double)z0);
        const long long NX = v.nx, NY = v.ny;
#define XR_AT(zz, xx, yy)                                                              \
    (v.dtype == 0 ? (double)__ldg((const float*)v.data + ((zz) * NX * NY + (xx) * NY + (yy))) \
                  : __ldg((const double*)v.data + ((zz) * NX * NY + (xx) * NY + (yy))))
        double v000 = XR_AT(z0, x0, y0), v001 = XR_AT(z1, x0, y0), v010 = XR_AT(z0, x0, y1), v011 = XR_AT(z1, x0, y1);
        double v100 = XR_AT(z0, x1, y0), v101 = XR_AT(z1, x1, y0), v110 = XR_AT(z0, x1, y1), v111 = XR_AT(z1, x1, y1);
#undef XR_AT
        double omz = dsub(1.0, wz), omy = dsub(1.0, wy), omx = dsub(1.0, wx);
        double v00 = dadd(dmul(v000, omz), dmul(v001, wz));
        double v01 = dadd(dmul(v010, omz), dmul(v011, wz));
        double v10 = dadd(dmul(v100, omz), dmul(v101, wz));
        double v11 = dadd(dmul(v110, omz), dmul(v111, wz));
        double v0 = dadd(dmul(v00, omy), dmul(v01, wy));
        double v1 = dadd(dmul(v10, omy), dmul(v11, wy));
        return dadd(dmul(v0, omx), dmul(v1, wx));
    }

// ---------------------------------------------------------------------------------------
// Exact fp64 evaluators (reference operation order)
// ---------------------------------------------------------------------------------------
struct Exact {
    typedef double real;
    static constexpr bool kFast = false;

    // deformations.go Apply methods
    static __device__ void deform(const DeformRec& r, double& x, double& y, double& z) {
        const double* p = r.d;
        switch (r.type) {
            case D_GAUSSIAN: {  // :29-38
                double x0 = dsub(x, p[6]), y0 = dsub(y, p[7]), z0 = dsub(z, p[8]);
                double r2 = dadd(dadd(dmul(x0, x0), dmul(y0, y0)), dmul(z0, z0));
                double ex = dmul(p[0], exp(ddiv(-r2, dmul(dmul(2.0, p[3]), p[3]))));
                double ey = dmul(p[1], exp(ddiv(-r2, dmul(dmul(2.0, p[4]), p[4]))));
                double ez = dmul(p[2], exp(ddiv(-r2, dmul(dmul(2.0, p[5]), p[5]))));
                x = dadd(x, ex);
                y = dadd(y, ey);
                z = dadd(z, ez);
                break;
            }
            case D_AFFINE: {  // :87-92
                double nx = dadd(dadd(dmul(p[0], x), dmul(p[1], y)), dmul(p[2], z));
                double ny = dadd(dadd(dmul(p[3], x), dmul(p[4], y)), dmul(p[5], z));
                double nz = dadd(dadd(dmul(p[6], x), dmul(p[7], y)), dmul(p[8], z));
                x = nx; y = ny; z = nz;
                break;
            }
            case D_LINEAR: {  // :136-141
                double nx = dadd(dadd(dadd(x, dmul(p[0], x)), dmul(p[5], y)), dmul(p[4], z));
                double ny = dadd(dadd(dadd(y, dmul(p[5], x)), dmul(p[1], y)), dmul(p[3], z));
                double nz = dadd(dadd(dadd(z, dmul(p[4], x)), dmul(p[3], y)), dmul(p[2], z));
                x = nx; y = ny; z = nz;
                break;
            }
            case D_RIGID:  // :173-175
                x = dadd(x, p[0]);
                y = dadd(y, p[1]);
                z = dadd(z, p[2]);
                break;
            case D_SIGMOID: {  // :210-222
                double q = r.axis == 0 ? x : (r.axis == 1 ? y : z);
                q = dadd(q, ddiv(p[0], dadd(1.0, exp(ddiv(-dsub(q, p[1]), p[2])))));
                if (r.axis == 0) x = q;
                else if (r.axis == 1) y = q;
                else z = q;
                break;
            }
        }
    }

    static __device__ __forceinline__ bool sphere(const SceneView& S, const Instr& I, int c, double x, double y, double z,
                                                  double& rho, bool&) {
        const double* p = S.f64 + I.f64_idx + c * kF64Sphere;
        x = dsub(x, p[0]);
        y = dsub(y, p[1]);
        z = dsub(z, p[2]);
        double r2 = dadd(dadd(dmul(x, x), dmul(y, y)), dmul(z, z));
        rho = p[4];
        return r2 < dmul(p[3], p[3]);
    }
    static __device__ __forceinline__ bool box(const SceneView& S, const Instr& I, int c, double x, double y, double z,
                                               double& rho, bool&) {
        const double* p = S.f64 + I.f64_idx + c * kF64Box;
        x = fabs(dsub(x, p[0]));
        y = fabs(dsub(y, p[1]));
        z = fabs(dsub(z, p[2]));
        rho = p[6];
        return x < dmul(0.5, p[3]) && y < dmul(0.5, p[4]) && z < dmul(0.5, p[5]);
    }
    static __device__ __forceinline__ bool pped(const SceneView& S, const Instr& I, int c, double x, double y, double z,
                                                double& rho, bool&) {
        const double* p = S.f64 + I.f64_idx + c * kF64Pped;
        const double* m = p + 3;
        double dx = dsub(x, p[0]), dy = dsub(y, p[1]), dz = dsub(z, p[2]);
        double qx = dadd(dadd(dmul(m[0], dx), dmul(m[3], dy)), dmul(m[6], dz));
        double qy = dadd(dadd(dmul(m[1], dx), dmul(m[4], dy)), dmul(m[7], dz));
        double qz = dadd(dadd(dmul(m[2], dx), dmul(m[5], dy)), dmul(m[8], dz));
        rho = p[12];
        return qx > 0.0 && qx < 1.0 && qy > 0.0 && qy < 1.0 && qz > 0.0 && qz < 1.0;
    }
    static __device__ __forceinline__ bool cyl(const SceneView& S, const Instr& I, int c, double x, double y, double z,
                                               double& rho, bool&) {
        const double* p = S.f64 + I.f64_idx + c * kF64Cyl;
        double vx = dsub(p[3], p[0]), vy = dsub(p[4], p[1]), vz = dsub(p[5], p[2]);
        double wx = dsub(x, p[0]), wy = dsub(y, p[1]), wz = dsub(z, p[2]);
        double wv = dadd(dadd(dmul(wx, vx), dmul(wy, vy)), dmul(wz, vz));
        double vv = dadd(dadd(dmul(vx, vx), dmul(vy, vy)), dmul(vz, vz));
        double cc = ddiv(wv, vv);
        rho = p[7];
        if (cc < 0.0 || cc > 1.0) return false;  // NaN (degenerate cylinder) falls through like Go
        double ex = dsub(wx, dmul(vx, cc)), ey = dsub(wy, dmul(vy, cc)), ez = dsub(wz, dmul(vz, cc));
        double d = __dsqrt_rn(dadd(dadd(dmul(ex, ex), dmul(ey, ey)), dmul(ez, ez)));
        return d < p[6];
    }
    static __device__ __forceinline__ bool gyroid(const SceneView& S, const Instr& I, int c, double x, double y, double z,
                                                  double& rho, bool&) {
        const double* p = S.f64 + I.f64_idx + c * kF64Gyroid;
        x = ddiv(dsub(x, p[0]), p[3]);
        y = ddiv(dsub(y, p[1]), p[3]);
        z = ddiv(dsub(z, p[2]), p[3]);
        double sx, cx, sy, cy, sz, cz;
        sincos(x, &sx, &cx);
        sincos(y, &sy, &cy);
        sincos(z, &sz, &cz);
        double g = dadd(dadd(dmul(sx, cy), dmul(sy, cz)), dmul(sz, cx));
        rho = p[5];
        return fabs(g) < p[4];
    }
    // objects.go:568-582 + :458-464
    static __device__ __forceinline__ bool tess(const SceneView& S, const Instr& I, double& x, double& y, double& z, bool&) {
        const double* p = S.f64 + I.f64_idx;
        if (x < p[0] || x > p[1] || y < p[2] || y > p[3] || z < p[4] || z > p[5]) return false;
        const double* u = p + 6;
        double dx = dsub(u[1], u[0]);
        x = dsub(x, dmul(dx, floor(ddiv(dsub(x, u[0]), dx))));
        double dy = dsub(u[3], u[2]);
        y = dsub(y, dmul(dy, floor(ddiv(dsub(y, u[2]), dy))));
        double dz = dsub(u[5], u[4]);
        z = dsub(z, dmul(dz, floor(ddiv(dsub(z, u[4]), dz))));
        if (x < u[0] || x > u[1] || y < u[2] || y > u[3] || z < u[4] || z > u[5]) return false;
        return true;
    }
    static __device__ double voxel(const SceneView& S, const Instr& I, double x, double y, double z, bool&) {
        const VoxelDev v = S.vox[I.aux];
        return voxel_exact(v, x, y, z);
    }
};

// ---------------------------------------------------------------------------------------
// Fast fp32 evaluators with guard bands.  `near` is set when the predicate is within its
// error bound of flipping; the caller then re-evaluates the sample with Exact.
// ---------------------------------------------------------------------------------------
// sin and cos through the SFU: Cody-Waite reduction to [-pi, pi] (2*pi = hi - 1.7484555e-7), then
// MUFU.SIN / MUFU.COS (abs error 2^-21.4 / 2^-21.2 on that range).  Total error < 1e-6 for |a| < 1e4;
// the gyroid guard band (scene_compile.cpp) budgets for it.
__device__ __forceinline__ void fast_sincos(float a, float* s, float* c) {
    const float k = rintf(a * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, a);
    r = fmaf(k, 1.7484555e-7f, r);
    *s = __sinf(r);
    *c = __cosf(r);
}

struct Fast {
    typedef float real;
    static constexpr bool kFast = true;

    static __device__ void deform(const DeformRec& r, float& x, float& y, float& z) {
        const float* p = r.f;
        switch (r.type) {
            case D_GAUSSIAN: {
                float x0 = x - p[6], y0 = y - p[7], z0 = z - p[8];
                float r2 = x0 * x0 + y0 * y0 + z0 * z0;
                x += p[0] * expf(r2 * p[3]);
                y += p[1] * expf(r2 * p[4]);
                z += p[2] * expf(r2 * p[5]);
                break;
            }
            case D_AFFINE: {
                float nx = p[0] * x + p[1] * y + p[2] * z;
                float ny = p[3] * x + p[4] * y + p[5] * z;
                float nz = p[6] * x + p[7] * y + p[8] * z;
                x = nx; y = ny; z = nz;
                break;
            }
            case D_LINEAR: {
                float nx = x + p[0] * x + p[5] * y + p[4] * z;
                float ny = y + p[5] * x + p[1] * y + p[3] * z;
                float nz = z + p[4] * x + p[3] * y + p[2] * z;
                x = nx; y = ny; z = nz;
                break;
            }
            case D_RIGID:
                x += p[0];
                y += p[1];
                z += p[2];
                break;
            case D_SIGMOID: {
                float q = r.axis == 0 ? x : (r.axis == 1 ? y : z);
                // SFU exp and reciprocal: relative error ~1e-6 on a term bounded by |A| (budgeted in eps_pos)
                q += __fdividef(p[0], 1.0f + __expf((q - p[1]) * p[2]));  // p[2] = -1/L
                if (r.axis == 0) x = q;
                else if (r.axis == 1) y = q;
                else z = q;
                break;
            }
        }
    }

    static __device__ __forceinline__ bool sphere(const SceneView& S, const Instr& I, int c, float x, float y, float z,
                                                  float& rho, bool& near) {
        const float4* q = S.f32 + I.f32_idx + c * kF32Sphere;
        const float4 a = q[0], b = q[1];
        float dx = x - a.x, dy = y - a.y, dz = z - a.z;
        float d2 = dx * dx + dy * dy + dz * dz;
        rho = a.w;
        near = fabsf(d2 - b.x) < b.y;
        return d2 < b.x;
    }
    static __device__ __forceinline__ bool box(const SceneView& S, const Instr& I, int c, float x, float y, float z,
                                               float& rho, bool& near) {
        const float4* q = S.f32 + I.f32_idx + c * kF32Box;
        const float4 a = q[0], b = q[1];
        float m = fmaxf(fabsf(x - a.x) - b.x, fmaxf(fabsf(y - a.y) - b.y, fabsf(z - a.z) - b.z));
        rho = a.w;
        near = fabsf(m) < b.w;
        return m < 0.0f;
    }
    static __device__ __forceinline__ bool pped(const SceneView& S, const Instr& I, int c, float x, float y, float z,
                                                float& rho, bool& near) {
        const float4* q = S.f32 + I.f32_idx + c * kF32Pped;
        const float4 a = q[0], r0 = q[1], r1 = q[2], r2 = q[3];
        float dx = x - a.x, dy = y - a.y, dz = z - a.z;
        float qx = r0.x * dx + r0.y * dy + r0.z * dz;
        float qy = r1.x * dx + r1.y * dy + r1.z * dz;
        float qz = r2.x * dx + r2.y * dy + r2.z * dz;
        float m = fmaxf(fabsf(qx - 0.5f), fmaxf(fabsf(qy - 0.5f), fabsf(qz - 0.5f))) - 0.5f;
        rho = a.w;
        near = fabsf(m) < r0.w;
        return m < 0.0f;
    }
    static __device__ __forceinline__ bool cyl(const SceneView& S, const Instr& I, int c, float x, float y, float z,
                                               float& rho, bool& near) {
        const float4* q = S.f32 + I.f32_idx + c * kF32Cyl;
        const float4 a = q[0], v = q[1], t = q[2];
        float wx = x - a.x, wy = y - a.y, wz = z - a.z;
        float cc = (wx * v.x + wy * v.y + wz * v.z) * v.w;
        float ex = wx - v.x * cc, ey = wy - v.y * cc, ez = wz - v.z * cc;
        float d2 = ex * ex + ey * ey + ez * ez;
        rho = a.w;
        // |cc - 0.5| <= 0.5 -+ tolc  <=>  cc in [0,1] shrunk / grown by tolc
        float ac = fabsf(cc - 0.5f) - 0.5f;  // <= 0 inside the axial range (caps inclusive)
        float ar = d2 - t.x;                 // < 0 inside the radius
        bool sure = (ac < -t.z) && (ar < -t.y);
        bool maybe = (ac < t.z) && (ar < t.y);
        near = maybe && !sure;
        return sure;
    }
    static __device__ __forceinline__ bool gyroid(const SceneView& S, const Instr& I, int c, float x, float y, float z,
                                                  float& rho, bool& near) {
        const float4* q = S.f32 + I.f32_idx + c * kF32Gyroid;
        const float4 a = q[0], b = q[1];
        float ax = (x - a.x) * b.x, ay = (y - a.y) * b.x, az = (z - a.z) * b.x;
        float sx, cx, sy, cy, sz, cz;
        sincosf(ax, &sx, &cx);
        sincosf(ay, &sy, &cy);
        sincosf(az, &sz, &cz);
        float g = sx * cy + sy * cz + sz * cx;
        float t = fabsf(g) - b.y;
        rho = a.w;
        near = fabsf(t) < b.z;
        return t < 0.0f;
    }
    static __device__ __forceinline__ bool tess(const SceneView& S, const Instr& I, float& x, float& y, float& z, bool& near) {
        const float4* q = S.f32 + I.f32_idx;
        const float4 oc = q[0], oh = q[1], um = q[2], d = q[3], id = q[4];
        float m = fmaxf(fabsf(x - oc.x) - oh.x, fmaxf(fabsf(y - oc.y) - oh.y, fabsf(z - oc.z) - oh.z));
        bool inside = m <= 0.0f;  // outer bounds are inclusive
        near = fabsf(m) < oc.w;
        float qx = (x - um.x) * id.x, qy = (y - um.y) * id.y, qz = (z - um.z) * id.z;
        float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
        float rx = qx - fx, ry = qy - fy, rz = qz - fz;
        float lo = fminf(rx, fminf(ry, rz)), hi = fmaxf(rx, fmaxf(ry, rz));
        // a fold within tolq of a cell face may pick the other period (and then the inclusive
        // unit-cell bounds test of UnitCell.Density can fire): let fp64 decide
        near = near || (inside && (lo < id.w || hi > 1.0f - id.w));
        x = fmaf(-d.x, fx, x);
        y = fmaf(-d.y, fy, y);
        z = fmaf(-d.z, fz, z);
        return inside;
    }
    static __device__ float voxel(const SceneView& S, const Instr& I, float x, float y, float z, bool& near) {
        const VoxelDev v = S.vox[I.aux];
        near = false;
        if (!v.data) return 0.0f;
        const float tol = S.f32[I.f32_idx].x;
        float m = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))) - 1.0f;
        near = fabsf(m) < tol;
        if (m > 0.0f) return 0.0f;
        float ux = (x + 1.0f) * 0.5f * (float)(v.nx - 1);
        float uy = (y + 1.0f) * 0.5f * (float)(v.ny - 1);
        float uz = (z + 1.0f) * 0.5f * (float)(v.nz - 1);
        int x0 = max(0, min(v.nx - 1, __float2int_rd(ux)));
        int y0 = max(0, min(v.ny - 1, __float2int_rd(uy)));
        int z0 = max(0, min(v.nz - 1, __float2int_rd(uz)));
        int x1 = min(x0 + 1, v.nx - 1), y1 = min(y0 + 1, v.ny - 1), z1 = min(z0 + 1, v.nz - 1);
        float wx = ux - (float)x0, wy = uy - (float)y0, wz = uz - (float)z0;
        const long long NX = v.nx, NY = v.ny;
#define XR_AT(zz, xx, yy)                                                       \
    (v.dtype == 0 ? __ldg((const float*)v.data + ((zz) * NX * NY + (xx) * NY + (yy))) \
                  : (float)__ldg((const double*)v.data + ((zz) * NX * NY + (xx) * NY + (yy))))
        float v000 = XR_AT(z0, x0, y0), v001 = XR_AT(z1, x0, y0), v010 = XR_AT(z0, x0, y1), v011 = XR_AT(z1, x0, y1);
        float v100 = XR_AT(z0, x1, y0), v101 = XR_AT(z1, x1, y0), v110 = XR_AT(z0, x1, y1), v111 = XR_AT(z1, x1, y1);
#undef XR_AT
        float v00 = fmaf(wz, v001 - v000, v000), v01 = fmaf(wz, v011 - v010, v010);
        float v10 = fmaf(wz, v101 - v100, v100), v11 = fmaf(wz, v111 - v110, v110);
        float v0 = fmaf(wy, v01 - v00, v00), v1 = fmaf(wy, v11 - v10, v10);
        const float val = fmaf(wx, v1 - v0, v0);
        // The VALUE is continuous, so fp32 is good to ~1e-6 everywhere inside the cube.  Whether it is exactly ZERO is not
        // (integrate_hierarchical refines where (rho == 0) flips, main.go:180): safely inside a cell the reference's lerp is 0
        // iff all eight corners are (same-sign corners cannot cancel); within the position error of a cell face the sample
        // may belong to the neighbouring cell, and in a mixed-sign cell the interpolant may cross zero: fp64 decides those.
        {
            const float tix = fmaf(tol, 0.5f * (float)(v.nx - 1), 4.0e-7f * (float)v.nx);
            const float tiy = fmaf(tol, 0.5f * (float)(v.ny - 1), 4.0e-7f * (float)v.ny);
            const float tiz = fmaf(tol, 0.5f * (float)(v.nz - 1), 4.0e-7f * (float)v.nz);
            // (an axis with a single layer has x1 == x0: its weight does not matter)
            const bool face = (v.nx > 1 && fminf(wx, 1.0f - wx) < tix) || (v.ny > 1 && fminf(wy, 1.0f - wy) < tiy) ||
                              (v.nz > 1 && fminf(wz, 1.0f - wz) < tiz);
            const float lo = fminf(fminf(fminf(v000, v001), fminf(v010, v011)), fminf(fminf(v100, v101), fminf(v110, v111)));
            const float hi = fmaxf(fmaxf(fmaxf(v000, v001), fmaxf(v010, v011)), fmaxf(fmaxf(v100, v101), fmaxf(v110, v111)));
            const bool mixed = lo < 0.0f && hi > 0.0f;
            near = near || face || (mixed && fabsf(val) <= 1.0e-3f * fmaxf(-lo, hi));
        }
        return val;
    }
};

// ---------------------------------------------------------------------------------------
// The interpreter.  All 32 lanes of the warp must call it together (warp votes inside);
// lanes with alive == false take part in the control flow only.
// ---------------------------------------------------------------------------------------
struct Counters {
    unsigned int prim_tests;  // primitive tests executed by this lane
};

template <class P>
struct SaveStack {
    typename P::real* r;  // [depth][5][blockDim]
    unsigned int* u;      // [depth][4][blockDim]
};

template <class P>
__device__ __forceinline__ unsigned long long grid_lookup(const SceneView& S, const Instr& I, typename P::real x,
                                                          typename P::real y, typename P::real z, bool alive) {
    const float4* g = S.f32 + I.f32_idx;
    const float4 gmin = g[0], ic = g[1], dims = g[2];
    const int gx = __float_as_int(dims.x), gy = __float_as_int(dims.y), gz = __float_as_int(dims.z);
    int ix = __float2int_rd(((float)x - gmin.x) * ic.x);
    int iy = __float2int_rd(((float)y - gmin.y) * ic.y);
    int iz = __float2int_rd(((float)z - gmin.z) * ic.z);
    ix = max(0, min(gx - 1, ix));
    iy = max(0, min(gy - 1, iy));
    iz = max(0, min(gz - 1, iz));
    unsigned long long m = 0ull;
    if (alive) m = __ldg(S.grids + I.aux + ((size_t)iz * gy + iy) * gx + ix);
    if (m >> 63) m = 0ull;  // empty cell: the word carries a skip distance, not a mask
    unsigned int lo = __reduce_or_sync(FULL_MASK, (unsigned int)m);
    unsigned int hi = __reduce_or_sync(FULL_MASK, (unsigned int)(m >> 32));
    return ((unsigned long long)hi << 32) | lo;
}

template <class P>
__device__ typename P::real eval_scene(const SceneView& S, typename P::real x, typename P::real y, typename P::real z,
                                       bool alive, bool& unc, SaveStack<P> st, Counters& cnt) {
    typedef typename P::real real;
    for (int i = 0; i < S.n_deform; ++i) P::deform(S.deform[i], x, y, z);

    real acc = 0, res = 0;
    bool done = false, greedy = false, had = false, multi = false, has_mask = false;
    unsigned long long umask = ~0ull;
    int coll_end = S.n_instr - 1;
    int sp = 0;
    int pc = 0;
    const int tid = threadIdx.x, nt = blockDim.x;

    for (;;) {
        const Instr I = S.instr[pc];
        switch (I.op) {
            case OP_END: return acc;

            case OP_SPHERE:
            case OP_BOX:
            case OP_CYL:
            case OP_PPED:
            case OP_GYROID: {
                unsigned long long bits = I.n >= 64 ? ~0ull : ((1ull << I.n) - 1ull);
                if (has_mask) bits &= (umask >> I.child_bit);
                bool jumped = false;
                while (bits) {
                    const int c = __ffsll((long long)bits) - 1;
                    bits &= bits - 1;
                    const bool act = alive && !done;
                    if (greedy && !__any_sync(FULL_MASK, act)) {
                        jumped = true;
                        break;
                    }
                    real rho;
                    bool near = false, in;
                    switch (I.op) {
                        case OP_SPHERE: in = P::sphere(S, I, c, x, y, z, rho, near); break;
                        case OP_BOX: in = P::box(S, I, c, x, y, z, rho, near); break;
                        case OP_CYL: in = P::cyl(S, I, c, x, y, z, rho, near); break;
                        case OP_PPED: in = P::pped(S, I, c, x, y, z, rho, near); break;
                        default: in = P::gyroid(S, I, c, x, y, z, rho, near); break;
                    }
                    if (act) {
                        cnt.prim_tests++;
                        if (P::kFast && near) unc = true;
                        if (in) {
                            if (greedy && rho > (real)0) {
                                res = rho;
                                done = true;
                            } else {
                                multi = multi || had;
                                had = true;
                                acc += rho;
                            }
                        }
                    }
                }
                pc = jumped ? coll_end : pc + 1;
                break;
            }

            case OP_VOXEL: {
                if (!has_mask || ((umask >> I.child_bit) & 1ull)) {
                    const bool act = alive && !done;
                    bool near = false;
                    real rho = act ? P::voxel(S, I, x, y, z, near) : (real)0;
                    if (act) {
                        cnt.prim_tests++;
                        if (P::kFast && near) unc = true;
                        if (greedy && rho > (real)0) {
                            res = rho;
                            done = true;
                        } else {
                            if (rho != (real)0) {
                                multi = multi || had;
                                had = true;
                            }
                            acc += rho;
                        }
                    }
                }
                ++pc;
                break;
            }

            case OP_COLL_BEGIN: {
                if (!(I.flags & F_NOSAVE)) {
                    st.r[(sp * 5 + 0) * nt + tid] = acc;
                    st.r[(sp * 5 + 1) * nt + tid] = res;
                    st.u[(sp * 4 + 0) * nt + tid] =
                        (done ? 1u : 0u) | (greedy ? 2u : 0u) | (had ? 4u : 0u) | (multi ? 8u : 0u) | (has_mask ? 16u : 0u);
                    st.u[(sp * 4 + 1) * nt + tid] = (unsigned int)umask;
                    st.u[(sp * 4 + 2) * nt + tid] = (unsigned int)(umask >> 32);
                    st.u[(sp * 4 + 3) * nt + tid] = (unsigned int)coll_end;
                    ++sp;
                }
                const bool act = alive && !done;  // lanes the parent still needs a value from
                acc = 0;
                res = 0;
                done = !act;  // lanes that need nothing behave as already finished
                had = multi = false;
                greedy = (I.flags & F_GREEDY) != 0;
                coll_end = (int)I.skip_to;
                has_mask = false;
                umask = ~0ull;
                bool skip = !__any_sync(FULL_MASK, act);
                if (!skip && (I.flags & F_HAS_LIST)) {
                    // big collection: merge the lanes' ascending cell lists (see render_fast.cu eval_list);
                    // the runs that follow are skipped
                    const float4* g = S.f32 + I.f32_idx;
                    const float4 gmin = g[0], ic = g[1], dims = g[2];
                    const uint4 lb = *reinterpret_cast<const uint4*>(g + 4);
                    const unsigned int lb64 = reinterpret_cast<const uint4*>(g + 5)->x;
                    const unsigned long long* gbase = S.grids + I.aux;
                    const unsigned int* l_off = reinterpret_cast<const unsigned int*>(gbase + lb.x);
                    const unsigned short* l_idx = reinterpret_cast<const unsigned short*>(gbase + lb.z);
                    const unsigned int* l_tab = reinterpret_cast<const unsigned int*>(gbase + lb.w);
                    const unsigned int* l_tab64 = reinterpret_cast<const unsigned int*>(gbase + lb64);
                    const int gx = __float_as_int(dims.x), gy = __float_as_int(dims.y), gz = __float_as_int(dims.z);
                    const int ix = max(0, min(gx - 1, __float2int_rd(((float)x - gmin.x) * ic.x)));
                    const int iy = max(0, min(gy - 1, __float2int_rd(((float)y - gmin.y) * ic.y)));
                    const int iz = max(0, min(gz - 1, __float2int_rd(((float)z - gmin.z) * ic.z)));
                    const unsigned int cell = (unsigned int)((iz * gy + iy) * gx + ix);
                    unsigned int lp = 0u, le = 0u;
                    if (act) {
                        lp = __ldg(l_off + cell);
                        le = __ldg(l_off + cell + 1);
                    }
                    for (;;) {
                        const unsigned int mine = lp < le ? (unsigned int)__ldg(l_idx + lp) : 0xFFFFu;
                        const unsigned int c = __reduce_min_sync(FULL_MASK, mine);
                        if (c == 0xFFFFu) break;
                        if (greedy && !__any_sync(FULL_MASK, alive && !done)) break;
                        if (mine == c) ++lp;
                        const unsigned int t = __ldg(l_tab + c);
                        Instr J = I;
                        J.f32_idx = t & 0xFFFFFFu;
                        J.f64_idx = __ldg(l_tab64 + c);
                        real rho;
                        bool near = false, in;
                        switch (t >> 24) {
                            case OP_SPHERE: in = P::sphere(S, J, 0, x, y, z, rho, near); break;
                            case OP_BOX: in = P::box(S, J, 0, x, y, z, rho, near); break;
                            case OP_CYL: in = P::cyl(S, J, 0, x, y, z, rho, near); break;
                            case OP_PPED: in = P::pped(S, J, 0, x, y, z, rho, near); break;
                            default: in = P::gyroid(S, J, 0, x, y, z, rho, near); break;
                        }
                        if (alive && !done) {
                            cnt.prim_tests++;
                            if (P::kFast && near) unc = true;
                            if (in) {
                                if (greedy && rho > (real)0) {
                                    res = rho;
                                    done = true;
                                } else {
                                    multi = multi || had;
                                    had = true;
                                    acc += rho;
                                }
                            }
                        }
                    }
                    skip = true;  // children handled: go straight to COLL_END
                }
                if (!skip && (I.flags & F_HAS_GRID)) {
                    umask = grid_lookup<P>(S, I, x, y, z, act);
                    has_mask = true;
                    skip = (umask == 0ull);
                }
                pc = skip ? coll_end : pc + 1;
                break;
            }

            case OP_COLL_END: {
                // objects.go:431-436: sum then clamp to [0,1]; greedy returns the first hit unclamped
                real val;
                if (greedy && res > (real)0) val = res;
                else {
                    val = acc;
                    if (val < (real)0) val = (real)0;
                    else if (val > (real)1) val = (real)1;
                    if (P::kFast && multi && fabsf((float)acc) < 1e-5f && alive) unc = true;
                }
                if (I.flags & F_NOSAVE) {
                    acc = val;
                    res = 0;
                    done = false;
                    greedy = false;
                    had = multi = false;
                    has_mask = false;
                    umask = ~0ull;
                    coll_end = S.n_instr - 1;
                } else {
                    --sp;
                    acc = st.r[(sp * 5 + 0) * nt + tid];
                    res = st.r[(sp * 5 + 1) * nt + tid];
                    const unsigned int f = st.u[(sp * 4 + 0) * nt + tid];
                    done = f & 1u;
                    greedy = f & 2u;
                    had = f & 4u;
                    multi = f & 8u;
                    has_mask = f & 16u;
                    umask = (unsigned long long)st.u[(sp * 4 + 1) * nt + tid] |
                            ((unsigned long long)st.u[(sp * 4 + 2) * nt + tid] << 32);
                    coll_end = (int)st.u[(sp * 4 + 3) * nt + tid];
                    if (alive && !done) {
                        if (greedy && val > (real)0) {
                            res = val;
                            done = true;
                        } else {
                            if (val != (real)0) {
                                multi = multi || had;
                                had = true;
                            }
                            acc += val;
                        }
                    }
                }
                ++pc;
                break;
            }

            case OP_TESS_BEGIN: {
                if (has_mask && !((umask >> I.child_bit) & 1ull)) {  // no lane's cell can see this child
                    pc = (int)I.skip_to + 1;
                    break;
                }
                if (!(I.flags & F_NOSAVE)) {
                    st.r[(sp * 5 + 0) * nt + tid] = x;
                    st.r[(sp * 5 + 1) * nt + tid] = y;
                    st.r[(sp * 5 + 2) * nt + tid] = z;
                    st.u[(sp * 4 + 0) * nt + tid] = alive ? 1u : 0u;
                    ++sp;
                }
                const bool act = alive && !done;
                bool near = false;
                const bool inside = P::tess(S, I, x, y, z, near);
                if (P::kFast && act && near) unc = true;
                alive = act && inside;
                ++pc;
                break;
            }

            case OP_TESS_END: {
                if (!(I.flags & F_NOSAVE)) {
                    --sp;
                    x = st.r[(sp * 5 + 0) * nt + tid];
                    y = st.r[(sp * 5 + 1) * nt + tid];
                    z = st.r[(sp * 5 + 2) * nt + tid];
                    alive = st.u[(sp * 4 + 0) * nt + tid] != 0u;
                }
                ++pc;
                break;
            }

            default: return acc;  // unreachable for a validated program
        }
    }
}

// ---------------------------------------------------------------------------------------
// Ray set-up (fp64, reference order): main.go:457-465 + mgl64.TransformCoordinate/Normalize
// ---------------------------------------------------------------------------------------
struct Ray64 {
    double o[3], d[3];
};

__device__ __forceinline__ Ray64 make_ray(const CamDev& c, int i, int j, int res) {
    const double half = ddiv((double)res, 2.0);
    const double px = dsub(ddiv((double)i, half), 1.0);
    const double py = dsub(ddiv((double)j, half), 1.0);
    const double pz = -c.f;
    const double* m = c.view;
    double t[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        t[r] = dadd(dadd(dadd(dmul(m[r * 4 + 0], px), dmul(m[r * 4 + 1], py)), dmul(m[r * 4 + 2], pz)), dmul(m[r * 4 + 3], 1.0));
    const double iw = ddiv(1.0, t[3]);
    Ray64 ry;
    double v[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        ry.o[a] = c.eye[a];
        v[a] = dsub(dmul(t[a], iw), c.eye[a]);
    }
    const double l = ddiv(1.0, __dsqrt_rn(dadd(dadd(dmul(v[0], v[0]), dmul(v[1], v[1])), dmul(v[2], v[2]))));
#pragma unroll
    for (int a = 0; a < 3; ++a) ry.d[a] = dmul(v[a], l);
    return ry;
}

// Ray / scene-bounds slab test -> [s_in, s_out]; false when the ray misses.
__device__ __forceinline__ bool clip_ray(const Ray64& r, const double* lo, const double* hi, double& s_in, double& s_out) {
    double t0 = -CUDART_INF, t1 = CUDART_INF;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (r.d[a] == 0.0) {
            if (r.o[a] < lo[a] || r.o[a] > hi[a]) return false;
        } else {
            double inv = 1.0 / r.d[a];
            double a0 = (lo[a] - r.o[a]) * inv, a1 = (hi[a] - r.o[a]) * inv;
            if (a0 > a1) {
                double tmp = a0;
                a0 = a1;
                a1 = tmp;
            }
            if (a0 > t0) t0 = a0;  // NaN (inf*0) compares false and leaves the bound untouched
            if (a1 < t1) t1 = a1;
        }
    }
    s_in = t0;
    s_out = t1;
    return t0 <= t1;
}

// The kernels render one (view, tile) per CTA, or -- second pass after render_span.cu -- loop over a compacted list of
// WARP tiles: entry = (view, tile) id * 4 + position of the warp's 32 pixels inside the tile; every warp of a CTA takes
// its own entry (xr_live_ is false for a warp without one: it walks along with valid == false).
#define XR_TILE_LOOP_BEGIN(P)                                                                              \
    const unsigned int xr_n_ = (P).tile_list ? *(P).tile_count : 0u;                                       \
    const unsigned int xr_end_ = (P).tile_list ? (xr_n_ + 3u) / 4u : blockIdx.x + 1u;                      \
    const unsigned int xr_stride_ = (P).tile_list ? gridDim.x : 0x40000000u;                               \
    for (unsigned int xr_vb_ = blockIdx.x; xr_vb_ < xr_end_; xr_vb_ += xr_stride_) {                       \
        unsigned int xr_block_ = xr_vb_, xr_sub_ = threadIdx.x >> 5;                                       \
        bool xr_live_ = true;                                                                              \
        if ((P).tile_list) {                                                                               \
            const unsigned int q_ = xr_vb_ * 4u + (threadIdx.x >> 5);                                      \
            xr_live_ = q_ < xr_n_;                                                                         \
            const unsigned int e_ = xr_live_ ? (P).tile_list[q_] : 0u;                                     \
            xr_block_ = e_ >> 2;                                                                           \
            xr_sub_ = e_ & 3u;                                                                             \
        }
#define XR_TILE_LOOP_END }

// block: (view, tile) id; sub: which of the tile's 2 x 2 warp positions this warp renders
__device__ __forceinline__ void pixel_of_thread(const RenderParams& P, unsigned int block, unsigned int sub, int& view, int& i, int& j) {
    const int tiles = P.tiles_i * P.tiles_j;
    const int b = (int)block;
    view = b / tiles;
    const int t = b - view * tiles;
    const int ti = t / P.tiles_j, tj = t - ti * P.tiles_j;
    const int w = (int)sub, lane = threadIdx.x & 31;
    i = ti * kTileI + (w >> 1) * kWarpI + lane / kWarpJ;
    j = tj * kTileJ + (w & 1) * kWarpJ + lane % kWarpJ;
}

// Stage instruction stream + fp32 pool in shared memory (warp-uniform LDS broadcasts).
__device__ __forceinline__ SceneView stage_program(const RenderParams& P, unsigned char* smem) {
    SceneView S;
    const SceneDev& D = P.scene;
    S.instr = D.instr;
    S.f32 = D.f32;
    S.f64 = D.f64;
    S.grids = D.grids;
    S.deform = D.deform;
    S.n_instr = D.n_instr;
    S.n_deform = D.n_deform;
    S.vox = D.vox;
    if (P.prog_in_smem) {
        uint4* dst = reinterpret_cast<uint4*>(smem);
        const uint4* srcI = reinterpret_cast<const uint4*>(D.instr);
        const int nI = D.n_instr * 2;
        for (int k = threadIdx.x; k < nI; k += blockDim.x) dst[k] = srcI[k];
        const uint4* srcF = reinterpret_cast<const uint4*>(D.f32);
        for (int k = threadIdx.x; k < D.f32_count; k += blockDim.x) dst[nI + k] = srcF[k];
        __syncthreads();
        S.instr = reinterpret_cast<const Instr*>(dst);
        S.f32 = reinterpret_cast<const float4*>(dst + nI);
    }
    return S;
}

__device__ __forceinline__ void store_pixel(const RenderParams& P, int view, int i, int j, bool valid, double value) {
    const size_t idx = ((size_t)view * P.res + i) * P.res + j;
    if (P.out_f64) {
        if (valid) reinterpret_cast<double*>(P.out)[idx] = value;
        return;
    }
    // 128-bit stores: lanes 4q..4q+3 hold 4 consecutive j of one image row
    float v = (float)value;
    float v1 = __shfl_down_sync(FULL_MASK, v, 1);
    float v2 = __shfl_down_sync(FULL_MASK, v, 2);
    float v3 = __shfl_down_sync(FULL_MASK, v, 3);
    float* out = reinterpret_cast<float*>(P.out);
    if (kWarpJ % 4 == 0 && (P.res & 3) == 0 && P.out_vec4) {
        if ((threadIdx.x & 3) == 0 && valid) *reinterpret_cast<float4*>(out + idx) = make_float4(v, v1, v2, v3);
    } else if (valid) {
        out[idx] = v;
    }
}

__device__ __forceinline__ void add_stats(const RenderParams& P, unsigned long long ref_samples,
                                          unsigned long long eval_samples, unsigned long long fallbacks,
                                          unsigned long long prim_tests, unsigned long long rays) {
    if (!P.stats) return;
    unsigned long long v[5] = {ref_samples, eval_samples, fallbacks, prim_tests, rays};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        unsigned long long s = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(P.stats + k, s);
    }
}

// Lane step range [k0,k1) covering every lattice position s_tab[k + off] inside [s_in, s_out].
__device__ __forceinline__ void step_range(const RenderParams& P, bool hit, double s_in, double s_out, int first_off,
                                           int& k0, int& k1) {
    if (!hit) {
        k0 = k1 = 0;
        return;
    }
    // position of step k is ~ smin + (k + first_off) * ds; repeated-addition drift << ds
    double a = (s_in - P.smin) / P.ds - first_off, b = (s_out - P.smin) / P.ds - first_off;
    a = fmin(fmax(a - 2.0, 0.0), (double)P.n_steps);
    b = fmin(fmax(b + 3.0, 0.0), (double)P.n_steps);
    k0 = (int)a;
    k1 = (int)b;
    if (k1 < k0) k1 = k0;
}


}  // namespace xr
