// xray_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU fp64 restatement of the per-pixel projection hot path of
// igrega348/xray_projection_render (the Go CPU path).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (libcuda_render.so) never links or calls it.
//
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
// -ffp-contract=off matters: Go/amd64 never fuses a*b+c, so neither may we.
//
// PARITY STATUS: PINNED AGAINST OUTPUT OF THE REFERENCE'S GO BINARY.  No Go toolchain
// exists in this image, but the reference repository carries one rendered image: the
// cell output stored in examples/demo.ipynb (three 300 x 300 projections of
// cube_w_hole.yaml).  This restatement reproduces all 270 000 of its pixels (52 067
// attenuated, 154 grey levels) exactly -- camera from angles, pixel -> ray, the
// hierarchical integrator with its refinements, collection / cube / sphere / cylinder
// densities with negative rho, exp(-T), 8-bit quantisation and y flip -- and no nearby
// parameter or convention does (tests/test_reference_go_output.py; fixture and extraction
// script under tests/golden/).  It is further pinned against every known-answer case of
// the reference's own unit tests (main_test.go, objects/objects_test.go,
// deformations/deformations_test.go; tests/test_oracle_known_answers.py), against an
// independent pure-Python restatement (tests/test_oracle_vs_python.py), and -- camera use,
// pixel mapping, layouts, voxeliser -- against the reference's own CUDA plugin built from
// /root/reference (tests/test_gpu_reference_pin.py).
// What the stored image does not exercise stays pinned only by the reference's unit-test
// answers and invariants: polar angles other than 90 degrees, Parallelepiped
// (mgl64.Mat3.Inv, from the un-vendored github.com/go-gl/mathgl v1.1.0, go.mod:6;
// M*Minv=I asserted), gyroid, tessellation, voxel grids and the deformations.
//
// Every function cites the reference file:line it follows.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------
// mgl64 (github.com/go-gl/mathgl v1.1.0, package mgl64) helpers, restated.
// ---------------------------------------------------------------------------
struct Vec3 {
    double v[3];
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
};

// mgl64 Vec3.Sub
static inline Vec3 vsub(const Vec3& a, const Vec3& b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
// mgl64 Vec3.Mul (scalar)
static inline Vec3 vmul(const Vec3& a, double c) { return {{a[0] * c, a[1] * c, a[2] * c}}; }
// mgl64 Vec3.Dot: a0*b0 + a1*b1 + a2*b2, left to right
static inline double vdot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// mgl64 Vec3.Len: sqrt(v0*v0 + v1*v1 + v2*v2)
static inline double vlen(const Vec3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
// mgl64 Vec3.Normalize: l := 1.0 / v.Len(); v * l
static inline Vec3 vnormalize(const Vec3& a) {
    double l = 1.0 / vlen(a);
    return {{a[0] * l, a[1] * l, a[2] * l}};
}
// mgl64 Vec3.Cross
static inline Vec3 vcross(const Vec3& a, const Vec3& b) {
    return {{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}

// mgl64 Mat4 is column-major: m[col*4+row].
struct Mat4 {
    double m[16];
    double at(int r, int c) const { return m[c * 4 + r]; }  // mgl64 Mat4.At
};

// mgl64 Mat4.Mul4: standard product, each element a left-to-right 4-term sum.
static Mat4 mat4_mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr)
            r.m[c * 4 + rr] = a.m[0 * 4 + rr] * b.m[c * 4 + 0] + a.m[1 * 4 + rr] * b.m[c * 4 + 1] +
                              a.m[2 * 4 + rr] * b.m[c * 4 + 2] + a.m[3 * 4 + rr] * b.m[c * 4 + 3];
    return r;
}

// mgl64 Mat4.Det: 24-term Leibniz expansion in lexicographic order.
static double mat4_det(const Mat4& M) {
    const double* m = M.m;
    return m[0] * m[5] * m[10] * m[15] - m[0] * m[5] * m[11] * m[14] - m[0] * m[6] * m[9] * m[15] +
           m[0] * m[6] * m[11] * m[13] + m[0] * m[7] * m[9] * m[14] - m[0] * m[7] * m[10] * m[13] -
           m[1] * m[4] * m[10] * m[15] + m[1] * m[4] * m[11] * m[14] + m[1] * m[6] * m[8] * m[15] -
           m[1] * m[6] * m[11] * m[12] - m[1] * m[7] * m[8] * m[14] + m[1] * m[7] * m[10] * m[12] +
           m[2] * m[4] * m[9] * m[15] - m[2] * m[4] * m[11] * m[13] - m[2] * m[5] * m[8] * m[15] +
           m[2] * m[5] * m[11] * m[12] + m[2] * m[7] * m[8] * m[13] - m[2] * m[7] * m[9] * m[12] -
           m[3] * m[4] * m[9] * m[14] + m[3] * m[4] * m[10] * m[13] + m[3] * m[5] * m[8] * m[14] -
           m[3] * m[5] * m[10] * m[12] - m[3] * m[6] * m[8] * m[13] + m[3] * m[6] * m[9] * m[12];
}

// mgl64 FloatEqual(det, 0): with b == 0 the threshold test reduces to |det| < 1e-20.
static inline bool float_equal_zero(double a) {
    if (a == 0.0) return true;
    return std::fabs(a) < 1e-10 * 1e-10;
}

// mgl64 Mat4.Inv: adjugate (explicit cofactors) times 1/det; zero matrix if det ~ 0.
static Mat4 mat4_inv(const Mat4& M) {
    const double* m = M.m;
    double det = mat4_det(M);
    Mat4 r;
    if (float_equal_zero(det)) {
        std::memset(r.m, 0, sizeof(r.m));
        return r;
    }
    double a[16] = {
        -m[7] * m[10] * m[13] + m[6] * m[11] * m[13] + m[7] * m[9] * m[14] - m[5] * m[11] * m[14] - m[6] * m[9] * m[15] + m[5] * m[10] * m[15],
        m[3] * m[10] * m[13] - m[2] * m[11] * m[13] - m[3] * m[9] * m[14] + m[1] * m[11] * m[14] + m[2] * m[9] * m[15] - m[1] * m[10] * m[15],
        -m[3] * m[6] * m[13] + m[2] * m[7] * m[13] + m[3] * m[5] * m[14] - m[1] * m[7] * m[14] - m[2] * m[5] * m[15] + m[1] * m[6] * m[15],
        m[3] * m[6] * m[9] - m[2] * m[7] * m[9] - m[3] * m[5] * m[10] + m[1] * m[7] * m[10] + m[2] * m[5] * m[11] - m[1] * m[6] * m[11],
        m[7] * m[10] * m[12] - m[6] * m[11] * m[12] - m[7] * m[8] * m[14] + m[4] * m[11] * m[14] + m[6] * m[8] * m[15] - m[4] * m[10] * m[15],
        -m[3] * m[10] * m[12] + m[2] * m[11] * m[12] + m[3] * m[8] * m[14] - m[0] * m[11] * m[14] - m[2] * m[8] * m[15] + m[0] * m[10] * m[15],
        m[3] * m[6] * m[12] - m[2] * m[7] * m[12] - m[3] * m[4] * m[14] + m[0] * m[7] * m[14] + m[2] * m[4] * m[15] - m[0] * m[6] * m[15],
        -m[3] * m[6] * m[8] + m[2] * m[7] * m[8] + m[3] * m[4] * m[10] - m[0] * m[7] * m[10] - m[2] * m[4] * m[11] + m[0] * m[6] * m[11],
        -m[7] * m[9] * m[12] + m[5] * m[11] * m[12] + m[7] * m[8] * m[13] - m[4] * m[11] * m[13] - m[5] * m[8] * m[15] + m[4] * m[9] * m[15],
        m[3] * m[9] * m[12] - m[1] * m[11] * m[12] - m[3] * m[8] * m[13] + m[0] * m[11] * m[13] + m[1] * m[8] * m[15] - m[0] * m[9] * m[15],
        -m[3] * m[5] * m[12] + m[1] * m[7] * m[12] + m[3] * m[4] * m[13] - m[0] * m[7] * m[13] - m[1] * m[4] * m[15] + m[0] * m[5] * m[15],
        m[3] * m[5] * m[8] - m[1] * m[7] * m[8] - m[3] * m[4] * m[9] + m[0] * m[7] * m[9] + m[1] * m[4] * m[11] - m[0] * m[5] * m[11],
        m[6] * m[9] * m[12] - m[5] * m[10] * m[12] - m[6] * m[8] * m[13] + m[4] * m[10] * m[13] + m[5] * m[8] * m[14] - m[4] * m[9] * m[14],
        -m[2] * m[9] * m[12] + m[1] * m[10] * m[12] + m[2] * m[8] * m[13] - m[0] * m[10] * m[13] - m[1] * m[8] * m[14] + m[0] * m[9] * m[14],
        m[2] * m[5] * m[12] - m[1] * m[6] * m[12] - m[2] * m[4] * m[13] + m[0] * m[6] * m[13] + m[1] * m[4] * m[14] - m[0] * m[5] * m[14],
        -m[2] * m[5] * m[8] + m[1] * m[6] * m[8] + m[2] * m[4] * m[9] - m[0] * m[6] * m[9] - m[1] * m[4] * m[10] + m[0] * m[5] * m[10],
    };
    double inv = 1 / det;  // retMat.Mul(1 / det)
    for (int i = 0; i < 16; ++i) r.m[i] = a[i] * inv;
    return r;
}

// mgl64 LookAtV(eye, center, up).
static Mat4 look_at(const Vec3& eye, const Vec3& center, const Vec3& up) {
    Vec3 f = vnormalize(vsub(center, eye));
    Vec3 s = vnormalize(vcross(f, vnormalize(up)));
    Vec3 u = vcross(s, f);
    Mat4 M = {{s[0], u[0], -f[0], 0, s[1], u[1], -f[1], 0, s[2], u[2], -f[2], 0, 0, 0, 0, 1}};
    // Translate3D(-eye)
    Mat4 T = {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -eye[0], -eye[1], -eye[2], 1}};
    return mat4_mul(M, T);
}

// mgl64 TransformCoordinate(v, m): t = m.Mul4x1(v.Vec4(1)); t.Vec3().Mul(1 / t[3])
static Vec3 transform_coordinate(const Vec3& v, const Mat4& M) {
    const double* m = M.m;
    double t0 = m[0] * v[0] + m[4] * v[1] + m[8] * v[2] + m[12] * 1.0;
    double t1 = m[1] * v[0] + m[5] * v[1] + m[9] * v[2] + m[13] * 1.0;
    double t2 = m[2] * v[0] + m[6] * v[1] + m[10] * v[2] + m[14] * 1.0;
    double t3 = m[3] * v[0] + m[7] * v[1] + m[11] * v[2] + m[15] * 1.0;
    double iw = 1 / t3;
    return {{t0 * iw, t1 * iw, t2 * iw}};
}

// mgl64 Mat3 (column-major m[col*3+row]) Det / Inv, used by Parallelepiped (objects.go:243).
static void mat3_inv(const double* m, double* out) {
    double det = m[0] * m[4] * m[8] + m[3] * m[7] * m[2] + m[6] * m[1] * m[5] - m[6] * m[4] * m[2] - m[3] * m[1] * m[8] -
                 m[0] * m[7] * m[5];
    if (float_equal_zero(det)) {
        for (int i = 0; i < 9; ++i) out[i] = 0;
        return;
    }
    double a[9] = {
        m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[2] * m[4],
        m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
        m[3] * m[7] - m[4] * m[6], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3],
    };
    double inv = 1 / det;
    for (int i = 0; i < 9; ++i) out[i] = a[i] * inv;
}

// mgl64.DegToRad: angle * Pi / 180
static inline double deg_to_rad(double a) { return a * M_PI / 180.0; }

// ---------------------------------------------------------------------------
// objects/objects.go -- the Object tree.
// ---------------------------------------------------------------------------
struct Object {
    virtual ~Object() {}
    virtual double Density(double x, double y, double z) const = 0;
    virtual double MinFeatureSize() const = 0;
};

// objects.go:63-76
struct Sphere : Object {
    Vec3 Center;
    double Radius, Rho;
    double Density(double x, double y, double z) const override {
        x = x - Center[0];
        y = y - Center[1];
        z = z - Center[2];
        double r_2 = x * x + y * y + z * z;
        if (r_2 < Radius * Radius) return Rho;
        return 0.0;
    }
    double MinFeatureSize() const override { return Radius; }
};

// objects.go:171-183 (Cube delegates to an embedded Box, objects.go:115-125)
struct Box : Object {
    Vec3 Center, Sides;
    double Rho;
    double Density(double x, double y, double z) const override {
        x = std::fabs(x - Center[0]);
        y = std::fabs(y - Center[1]);
        z = std::fabs(z - Center[2]);
        if (x < 0.5 * Sides[0] && y < 0.5 * Sides[1] && z < 0.5 * Sides[2]) return Rho;
        return 0.0;
    }
    double MinFeatureSize() const override { return 0.1 * std::fmin(Sides[0], std::fmin(Sides[1], Sides[2])); }
};

// objects.go:243-259
struct Parallelepiped : Object {
    Vec3 Origin, V0, V1, V2;
    double Rho;
    double mat[9];  // column-major inverse of [V0 V1 V2]
    void finish() {
        double m[9] = {V0[0], V0[1], V0[2], V1[0], V1[1], V1[2], V2[0], V2[1], V2[2]};  // Mat3FromCols
        mat3_inv(m, mat);
    }
    double Density(double x, double y, double z) const override {
        Vec3 d = vsub(Vec3{{x, y, z}}, Origin);
        // Mat3.Mul3x1: m[0]*v0 + m[3]*v1 + m[6]*v2, ...
        double qx = mat[0] * d[0] + mat[3] * d[1] + mat[6] * d[2];
        double qy = mat[1] * d[0] + mat[4] * d[1] + mat[7] * d[2];
        double qz = mat[2] * d[0] + mat[5] * d[1] + mat[8] * d[2];
        if (qx > 0.0 && qx < 1.0 && qy > 0.0 && qy < 1.0 && qz > 0.0 && qz < 1.0) return Rho;
        return 0.0;
    }
    double MinFeatureSize() const override { return 0.2 * std::fmin(vlen(V0), std::fmin(vlen(V1), vlen(V2))); }
};

// objects.go:334-354
struct Cylinder : Object {
    Vec3 P0, P1;
    double Radius, Rho;
    double Density(double x, double y, double z) const override {
        Vec3 v = vsub(P1, P0);
        Vec3 w = vsub(Vec3{{x, y, z}}, P0);
        double c = vdot(w, v) / vdot(v, v);
        if (c < 0.0 || c > 1.0) return 0.0;
        double d = vlen(vsub(w, vmul(v, c)));
        if (d < Radius) return Rho;
        return 0.0;
    }
    double MinFeatureSize() const override { return Radius; }
};

// objects.go:1014-1037
struct Gyroid : Object {
    Vec3 Center;
    double Scale, Thickness, Rho;
    double Density(double x, double y, double z) const override {
        x = (x - Center[0]) / Scale;
        y = (y - Center[1]) / Scale;
        z = (z - Center[2]) / Scale;
        double g = std::sin(x) * std::cos(y) + std::sin(y) * std::cos(z) + std::sin(z) * std::cos(x);
        if (std::fabs(g) < Thickness) return Rho;
        return 0.0;
    }
    double MinFeatureSize() const override { return Scale * Thickness * 0.1; }
};

// objects.go:422-446
struct ObjectCollection : Object {
    std::vector<std::unique_ptr<Object>> Objects;
    bool GreedyDensEval = false;
    double Density(double x, double y, double z) const override {
        double density = 0.0;
        for (const auto& o : Objects) {
            double rho = o->Density(x, y, z);
            if (GreedyDensEval && rho > 0.0) return rho;
            density += rho;
        }
        if (density < 0.0)
            density = 0.0;
        else if (density > 1.0)
            density = 1.0;
        return density;
    }
    double MinFeatureSize() const override {
        double out = INFINITY;
        for (const auto& o : Objects) out = std::fmin(out, o->MinFeatureSize());
        return out;
    }
};

// objects.go:458-464 (UnitCell) and 568-586 (TessellatedObjColl)
struct TessellatedObjColl : Object {
    ObjectCollection UC;  // UnitCell.Objects (greedy forced on at load, objects.go:487)
    double ucXmin, ucXmax, ucYmin, ucYmax, ucZmin, ucZmax;
    double Xmin, Xmax, Ymin, Ymax, Zmin, Zmax;
    double ucDensity(double x, double y, double z) const {
        if (x < ucXmin || x > ucXmax || y < ucYmin || y > ucYmax || z < ucZmin || z > ucZmax) return 0.0;
        return UC.Density(x, y, z);
    }
    double Density(double x, double y, double z) const override {
        if (x < Xmin || x > Xmax || y < Ymin || y > Ymax || z < Zmin || z > Zmax) return 0.0;
        double dx = ucXmax - ucXmin;
        x = x - dx * std::floor((x - ucXmin) / dx);
        double dy = ucYmax - ucYmin;
        y = y - dy * std::floor((y - ucYmin) / dy);
        double dz = ucZmax - ucZmin;
        z = z - dz * std::floor((z - ucZmin) / dz);
        return ucDensity(x, y, z);
    }
    double MinFeatureSize() const override { return UC.MinFeatureSize(); }
};

// objects.go:789-860
struct VoxelGrid : Object {
    std::vector<double> Rho;
    // Test hook for full-size volumes (1024^3): borrow the caller's fp32 array instead of copying it into 8 GiB of
    // doubles.  Widening fp32 -> fp64 is exact, so Density() returns what it would for the same values held as float64
    // (which is what the Go host has after loading a float32 .raw, objects.go:935-941).
    const float* RhoF = nullptr;
    long NX, NY, NZ;
    inline double at(long idx) const { return RhoF ? (double)RhoF[idx] : Rho[idx]; }
    double Density(double x, double y, double z) const override {
        if (x < -1 || x > 1 || y < -1 || y > 1 || z < -1 || z > 1) return 0.0;
        x = (x + 1) / 2;
        y = (y + 1) / 2;
        z = (z + 1) / 2;
        x = x * double(NX - 1);
        y = y * double(NY - 1);
        z = z * double(NZ - 1);
        long x0 = (long)std::floor(x), y0 = (long)std::floor(y), z0 = (long)std::floor(z);
        long x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
        if (x0 < 0) x0 = 0;
        if (y0 < 0) y0 = 0;
        if (z0 < 0) z0 = 0;
        if (x1 >= NX) x1 = NX - 1;
        if (y1 >= NY) y1 = NY - 1;
        if (z1 >= NZ) z1 = NZ - 1;
        double wx = x - double(x0), wy = y - double(y0), wz = z - double(z0);
        double v000 = at(z0 * NX * NY + x0 * NY + y0);
        double v001 = at(z1 * NX * NY + x0 * NY + y0);
        double v010 = at(z0 * NX * NY + x0 * NY + y1);
        double v011 = at(z1 * NX * NY + x0 * NY + y1);
        double v100 = at(z0 * NX * NY + x1 * NY + y0);
        double v101 = at(z1 * NX * NY + x1 * NY + y0);
        double v110 = at(z0 * NX * NY + x1 * NY + y1);
        double v111 = at(z1 * NX * NY + x1 * NY + y1);
        double v00 = v000 * (1 - wz) + v001 * wz;
        double v01 = v010 * (1 - wz) + v011 * wz;
        double v10 = v100 * (1 - wz) + v101 * wz;
        double v11 = v110 * (1 - wz) + v111 * wz;
        double v0 = v00 * (1 - wy) + v01 * wy;
        double v1 = v10 * (1 - wy) + v11 * wy;
        return v0 * (1 - wx) + v1 * wx;
    }
    double MinFeatureSize() const override { return 2.0 / double(std::max(NX, std::max(NY, NZ))); }
};

// ---------------------------------------------------------------------------
// deformations/deformations.go
// ---------------------------------------------------------------------------
struct Deformation {
    virtual ~Deformation() {}
    virtual void Apply(double& x, double& y, double& z) const = 0;
};
// deformations.go:29-38
struct GaussianDeformation : Deformation {
    double A[3], S[3], C[3];
    void Apply(double& x, double& y, double& z) const override {
        double x0 = x - C[0], y0 = y - C[1], z0 = z - C[2];
        double r2 = x0 * x0 + y0 * y0 + z0 * z0;
        double dx = A[0] * std::exp(-r2 / (2 * S[0] * S[0]));
        double dy = A[1] * std::exp(-r2 / (2 * S[1] * S[1]));
        double dz = A[2] * std::exp(-r2 / (2 * S[2] * S[2]));
        x = x + dx;
        y = y + dy;
        z = z + dz;
    }
};
// deformations.go:87-92
struct AffineDeformation : Deformation {
    double M[3][3];
    void Apply(double& x, double& y, double& z) const override {
        double _x = M[0][0] * x + M[0][1] * y + M[0][2] * z;
        double _y = M[1][0] * x + M[1][1] * y + M[1][2] * z;
        double _z = M[2][0] * x + M[2][1] * y + M[2][2] * z;
        x = _x;
        y = _y;
        z = _z;
    }
};
// deformations.go:136-141
struct LinearDeformation : Deformation {
    double S[6];
    void Apply(double& x, double& y, double& z) const override {
        double _x = x + S[0] * x + S[5] * y + S[4] * z;
        double _y = y + S[5] * x + S[1] * y + S[3] * z;
        double _z = z + S[4] * x + S[3] * y + S[2] * z;
        x = _x;
        y = _y;
        z = _z;
    }
};
// deformations.go:173-175
struct RigidDeformation : Deformation {
    double D[3];
    void Apply(double& x, double& y, double& z) const override {
        x = x + D[0];
        y = y + D[1];
        z = z + D[2];
    }
};
// deformations.go:210-222
struct SigmoidDeformation : Deformation {
    double Amplitude, Center, Lengthscale;
    int dir;  // 0 x, 1 y, 2 z
    void Apply(double& x, double& y, double& z) const override {
        switch (dir) {
            case 0: x = x + Amplitude / (1 + std::exp(-(x - Center) / Lengthscale)); break;
            case 1: y = y + Amplitude / (1 + std::exp(-(y - Center) / Lengthscale)); break;
            default: z = z + Amplitude / (1 + std::exp(-(z - Center) / Lengthscale)); break;
        }
    }
};
// deformations.go:269-274
struct ComposedDeformation : Deformation {
    std::vector<std::unique_ptr<Deformation>> list;
    void Apply(double& x, double& y, double& z) const override {
        for (const auto& d : list) d->Apply(x, y, z);
    }
};

// ---------------------------------------------------------------------------
// Token-stream parser for the test-side scene description (see oracle/oracle.py).
// Numbers are C99 hex floats, so the doubles arrive bit-exact.
// ---------------------------------------------------------------------------
struct Tok {
    const char* p;
    std::string err;
    const double* const* vox_data;
    int n_vox;
    bool next(std::string& out) {
        while (*p == ' ' || *p == '\n' || *p == '\t') ++p;
        if (!*p) return false;
        const char* s = p;
        while (*p && *p != ' ' && *p != '\n' && *p != '\t') ++p;
        out.assign(s, p - s);
        return true;
    }
    double num() {
        std::string t;
        if (!next(t)) {
            err = "unexpected end";
            return 0;
        }
        char* e = nullptr;
        double v = std::strtod(t.c_str(), &e);
        if (e == t.c_str() || *e) err = "bad number '" + t + "'";
        return v;
    }
    Vec3 vec3() {
        Vec3 v;
        v[0] = num();
        v[1] = num();
        v[2] = num();
        return v;
    }
};

static std::unique_ptr<Object> parse_object(Tok& t);

static bool parse_collection_body(Tok& t, ObjectCollection& oc) {
    oc.GreedyDensEval = t.num() != 0.0;
    int n = (int)t.num();
    for (int i = 0; i < n && t.err.empty(); ++i) {
        auto o = parse_object(t);
        if (!o) return false;
        oc.Objects.push_back(std::move(o));
    }
    return t.err.empty();
}

static std::unique_ptr<Object> parse_object(Tok& t) {
    std::string k;
    if (!t.next(k)) {
        t.err = "missing object";
        return nullptr;
    }
    if (k == "sphere") {
        auto o = std::make_unique<Sphere>();
        o->Center = t.vec3();
        o->Radius = t.num();
        o->Rho = t.num();
        return o;
    }
    if (k == "box") {
        auto o = std::make_unique<Box>();
        o->Center = t.vec3();
        o->Sides = t.vec3();
        o->Rho = t.num();
        return o;
    }
    if (k == "cube") {  // Cube -> embedded Box with Sides = (Side, Side, Side), objects.go:115
        auto o = std::make_unique<Box>();
        o->Center = t.vec3();
        double s = t.num();
        o->Sides = {{s, s, s}};
        o->Rho = t.num();
        return o;
    }
    if (k == "pped") {
        auto o = std::make_unique<Parallelepiped>();
        o->Origin = t.vec3();
        o->V0 = t.vec3();
        o->V1 = t.vec3();
        o->V2 = t.vec3();
        o->Rho = t.num();
        o->finish();
        return o;
    }
    if (k == "cylinder") {
        auto o = std::make_unique<Cylinder>();
        o->P0 = t.vec3();
        o->P1 = t.vec3();
        o->Radius = t.num();
        o->Rho = t.num();
        return o;
    }
    if (k == "gyroid") {
        auto o = std::make_unique<Gyroid>();
        o->Center = t.vec3();
        o->Scale = t.num();
        o->Thickness = t.num();
        o->Rho = t.num();
        return o;
    }
    if (k == "collection") {
        auto o = std::make_unique<ObjectCollection>();
        if (!parse_collection_body(t, *o)) return nullptr;
        return o;
    }
    if (k == "tess") {
        auto o = std::make_unique<TessellatedObjColl>();
        o->Xmin = t.num();
        o->Xmax = t.num();
        o->Ymin = t.num();
        o->Ymax = t.num();
        o->Zmin = t.num();
        o->Zmax = t.num();
        o->ucXmin = t.num();
        o->ucXmax = t.num();
        o->ucYmin = t.num();
        o->ucYmax = t.num();
        o->ucZmin = t.num();
        o->ucZmax = t.num();
        std::string c;
        if (!t.next(c) || c != "collection") {
            t.err = "tess needs a collection";
            return nullptr;
        }
        if (!parse_collection_body(t, o->UC)) return nullptr;
        o->UC.GreedyDensEval = true;  // objects.go:487
        return o;
    }
    if (k == "voxel" || k == "voxelf") {
        auto o = std::make_unique<VoxelGrid>();
        o->NX = (long)t.num();
        o->NY = (long)t.num();
        o->NZ = (long)t.num();
        int idx = (int)t.num();
        if (idx < 0 || idx >= t.n_vox || !t.vox_data[idx]) {
            t.err = "voxel data index out of range";
            return nullptr;
        }
        size_t n = (size_t)o->NX * o->NY * o->NZ;
        if (k == "voxelf") o->RhoF = reinterpret_cast<const float*>(t.vox_data[idx]);  // borrowed: the caller keeps it alive
        else o->Rho.assign(t.vox_data[idx], t.vox_data[idx] + n);
        return o;
    }
    t.err = "unknown object '" + k + "'";
    return nullptr;
}

static std::unique_ptr<Deformation> parse_deformation(Tok& t) {
    std::string k;
    if (!t.next(k)) {
        t.err = "missing deformation";
        return nullptr;
    }
    if (k == "gaussian") {
        auto d = std::make_unique<GaussianDeformation>();
        for (int i = 0; i < 3; ++i) d->A[i] = t.num();
        for (int i = 0; i < 3; ++i) d->S[i] = t.num();
        for (int i = 0; i < 3; ++i) d->C[i] = t.num();
        return d;
    }
    if (k == "affine") {
        auto d = std::make_unique<AffineDeformation>();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) d->M[i][j] = t.num();
        return d;
    }
    if (k == "linear") {
        auto d = std::make_unique<LinearDeformation>();
        for (int i = 0; i < 6; ++i) d->S[i] = t.num();
        return d;
    }
    if (k == "rigid") {
        auto d = std::make_unique<RigidDeformation>();
        for (int i = 0; i < 3; ++i) d->D[i] = t.num();
        return d;
    }
    if (k == "sigmoid") {
        auto d = std::make_unique<SigmoidDeformation>();
        d->Amplitude = t.num();
        d->Center = t.num();
        d->Lengthscale = t.num();
        d->dir = (int)t.num();
        return d;
    }
    if (k == "composed") {
        auto d = std::make_unique<ComposedDeformation>();
        int n = (int)t.num();
        for (int i = 0; i < n && t.err.empty(); ++i) {
            auto c = parse_deformation(t);
            if (!c) return nullptr;
            d->list.push_back(std::move(c));
        }
        return d;
    }
    t.err = "unknown deformation '" + k + "'";
    return nullptr;
}

// ---------------------------------------------------------------------------
// main.go globals (lat, df, density_multiplier, flat_field) live in a Scene.
// ---------------------------------------------------------------------------
struct Scene {
    std::unique_ptr<Object> lat;
    std::unique_ptr<Deformation> df;  // 0 or 1 deformation (main.go:123-133)
    double density_multiplier = 1.0;
    double flat_field = 0.0;

    // main.go:137-140 (+ deform, main.go:123-133)
    inline double density(double x, double y, double z) const {
        if (df) df->Apply(x, y, z);
        return lat->Density(x, y, z) * density_multiplier;
    }
};

// main.go:144-154.  *nsamples counts density() calls (reference-equivalent samples).
static double integrate_along_ray(const Scene& sc, Vec3 origin, Vec3 direction, double ds, double smin, double smax,
                                  long* nsamples) {
    direction = vnormalize(direction);
    double T = sc.flat_field;
    long n = 0;
    for (double s = smin; s < smax; s += ds) {
        double x = origin[0] + direction[0] * s;
        double y = origin[1] + direction[1] * s;
        double z = origin[2] + direction[2] * s;
        T += sc.density(x, y, z) * ds;
        ++n;
    }
    if (nsamples) *nsamples += n;
    return std::exp(-T);
}

// main.go:159-199.  The two clipping probes (main.go:162-169) only emit warnings and
// do not change the result; they are skipped and not counted as samples.
static double integrate_hierarchical(const Scene& sc, Vec3 origin, Vec3 direction, double DS, double smin, double smax,
                                     long* nsamples) {
    direction = vnormalize(direction);
    double right = smin + DS;
    double left = smin;
    double ds = DS / 10.0;
    double prev_rho = 0.0;
    double T = sc.flat_field;
    long n = 0;
    while (right <= smax) {
        double x = origin[0] + direction[0] * right;
        double y = origin[1] + direction[1] * right;
        double z = origin[2] + direction[2] * right;
        double rho = sc.density(x, y, z);
        ++n;
        if ((rho == 0) != (prev_rho == 0)) {
            left += ds;
            while (left < right) {
                double x2 = origin[0] + direction[0] * left;
                double y2 = origin[1] + direction[1] * left;
                double z2 = origin[2] + direction[2] * left;
                T += sc.density(x2, y2, z2) * ds;
                ++n;
                left += ds;
            }
            T += rho * ds;
        } else {
            T += rho * DS;
        }
        prev_rho = rho;
        left = right;
        right += DS;
    }
    if (nsamples) *nsamples += n;
    return std::exp(-T);
}

// main.go:226-239
static void compute_camera_from_angles(double az_deg, double polar_deg, double R, Vec3& eye, Mat4& camera) {
    double th = deg_to_rad(az_deg);
    double phi = deg_to_rad(polar_deg);
    eye = {{R * std::cos(th) * std::sin(phi), R * std::sin(th) * std::sin(phi), std::cos(phi) * R}};
    Vec3 center = {{0, 0, 0}};
    Vec3 up = {{0, 0, 1}};
    camera = mat4_inv(look_at(eye, center, up));
}

const double cube_half_diagonal = 1.74;  // main.go:46

}  // namespace

// ---------------------------------------------------------------------------
// C ABI for ctypes.
// ---------------------------------------------------------------------------
extern "C" {

static thread_local std::string g_err;
const char* oracle_last_error() { return g_err.c_str(); }

// desc: object token stream; deform_desc: deformation token stream or NULL/"".
// vox: array of n_vox pointers to fp64 volumes in the reference layout idx = z*NX*NY + x*NY + y.
void* oracle_scene_create(const char* desc, const char* deform_desc, const double* const* vox, int n_vox) {
    auto sc = std::make_unique<Scene>();
    Tok t{desc, "", vox, n_vox};
    sc->lat = parse_object(t);
    if (!sc->lat || !t.err.empty()) {
        g_err = t.err.empty() ? "parse failed" : t.err;
        return nullptr;
    }
    if (deform_desc && *deform_desc) {
        Tok td{deform_desc, "", nullptr, 0};
        sc->df = parse_deformation(td);
        if (!sc->df || !td.err.empty()) {
            g_err = td.err.empty() ? "deformation parse failed" : td.err;
            return nullptr;
        }
    }
    return sc.release();
}
void oracle_scene_destroy(void* h) { delete (Scene*)h; }
void oracle_scene_set_globals(void* h, double flat_field, double density_multiplier) {
    ((Scene*)h)->flat_field = flat_field;
    ((Scene*)h)->density_multiplier = density_multiplier;
}
double oracle_min_feature_size(void* h) { return ((Scene*)h)->lat->MinFeatureSize(); }
// objects: lat[0].Density (no warp, no multiplier)
double oracle_object_density(void* h, double x, double y, double z) { return ((Scene*)h)->lat->Density(x, y, z); }
// main.go density(): warp + multiplier
double oracle_density(void* h, double x, double y, double z) { return ((Scene*)h)->density(x, y, z); }
void oracle_deform(void* h, double* xyz) {
    Scene* s = (Scene*)h;
    if (s->df) s->df->Apply(xyz[0], xyz[1], xyz[2]);
}

// integrator: 0 = integrate_along_ray, 1 = integrate_hierarchical
double oracle_integrate(void* h, int integrator, const double* origin, const double* dir, double ds, double smin,
                        double smax, long* nsamples) {
    const Scene& sc = *(Scene*)h;
    Vec3 o = {{origin[0], origin[1], origin[2]}}, d = {{dir[0], dir[1], dir[2]}};
    return integrator == 0 ? integrate_along_ray(sc, o, d, ds, smin, smax, nsamples)
                           : integrate_hierarchical(sc, o, d, ds, smin, smax, nsamples);
}

// main.go:226-239.  eye[3]; camera_rowmajor[16] with [r*4+c] = camera.At(r,c)
// (the layout cuda_path.go:66-71 hands to the plugin and main.go:449-453 writes to transforms.json).
void oracle_camera_from_angles(double az_deg, double polar_deg, double R, double* eye, double* camera_rowmajor) {
    Vec3 e;
    Mat4 cam;
    compute_camera_from_angles(az_deg, polar_deg, R, e, cam);
    for (int i = 0; i < 3; ++i) eye[i] = e[i];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) camera_rowmajor[r * 4 + c] = cam.at(r, c);
}

// Render one view given an explicit camera (eye + row-major camera->world matrix), main.go:456-471.
// out[i*res + j] (the GPU ABI layout, cuda_backend.h:96-97).  rows [i0,i1) only (for bounded samples).
// Returns the number of reference-equivalent samples evaluated.
long oracle_render_view(void* h, const double* eye_in, const double* cam_rowmajor, int res, double fov_deg, double R,
                        double ds, int integrator, int i0, int i1, int jstride, double* out, int nthreads) {
    const Scene& sc = *(Scene*)h;
    Vec3 eye = {{eye_in[0], eye_in[1], eye_in[2]}};
    Mat4 cam;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) cam.m[c * 4 + r] = cam_rowmajor[r * 4 + c];
    double res_f = double(res);
    double f = 1 / std::tan(deg_to_rad(fov_deg / 2));
    double smin = R - cube_half_diagonal, smax = R + cube_half_diagonal;
    long total = 0;
    if (jstride < 1) jstride = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int i = i0; i < i1; ++i) {
        for (int j = 0; j < res; j += jstride) {
            Vec3 vx = {{double(i) / (res_f / 2) - 1, double(j) / (res_f / 2) - 1, -f}};
            vx = transform_coordinate(vx, cam);
            Vec3 dir = vsub(vx, eye);
            long n = 0;
            double v = integrator == 0 ? integrate_along_ray(sc, eye, dir, ds, smin, smax, &n)
                                       : integrate_hierarchical(sc, eye, dir, ds, smin, smax, &n);
            out[(size_t)i * res + j] = v;
            total += n;
        }
    }
    return total;
}

// Evaluate a list of pixels (i,j) of one view -- used for random-pixel parity at large sizes.
long oracle_render_pixels(void* h, const double* eye_in, const double* cam_rowmajor, int res, double fov_deg, double R,
                          double ds, int integrator, const int* ij, int npix, double* out, int nthreads) {
    const Scene& sc = *(Scene*)h;
    Vec3 eye = {{eye_in[0], eye_in[1], eye_in[2]}};
    Mat4 cam;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) cam.m[c * 4 + r] = cam_rowmajor[r * 4 + c];
    double res_f = double(res);
    double f = 1 / std::tan(deg_to_rad(fov_deg / 2));
    double smin = R - cube_half_diagonal, smax = R + cube_half_diagonal;
    long total = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int p = 0; p < npix; ++p) {
        int i = ij[2 * p], j = ij[2 * p + 1];
        Vec3 vx = {{double(i) / (res_f / 2) - 1, double(j) / (res_f / 2) - 1, -f}};
        vx = transform_coordinate(vx, cam);
        Vec3 dir = vsub(vx, eye);
        long n = 0;
        out[p] = integrator == 0 ? integrate_along_ray(sc, eye, dir, ds, smin, smax, &n)
                                 : integrate_hierarchical(sc, eye, dir, ds, smin, smax, &n);
        total += n;
    }
    return total;
}

// Number of iterations of `for s := smin; s < smax; s += ds` (simple) or of
// `for right <= smax` (hierarchical coarse loop) in fp64 repeated addition.
long oracle_step_count(int integrator, double ds, double smin, double smax) {
    long n = 0;
    if (integrator == 0) {
        for (double s = smin; s < smax; s += ds) ++n;
    } else {
        for (double right = smin + ds; right <= smax; right += ds) ++n;
    }
    return n;
}

// Mat3 inverse exactly as Parallelepiped.FromMap builds it (column-major in/out).
void oracle_mat3_inv(const double* m, double* out) { mat3_inv(m, out); }

int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
