// scene_compile.cpp -- host-side scene front end of libcuda_render.so.
//
// JSON (schema of reference objects/objects.go FromMap + deformations/deformations.go
// NewDeformation) -> object tree -> flattened warp-uniform instruction buffer (program.h),
// plus MinFeatureSize (auto ds, main.go:350-353), the conservative world-space bounds used
// for ray clipping, child-mask grids for collections, and the fp32 guard-band tolerances.
//
// Compiled with -ffp-contract=off: the fp64 host helpers here (camera, Mat3 inverse, host
// density) follow the reference's operation order (Go/amd64 never fuses multiply-add).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include "json_min.h"
#include "scene.h"

namespace xr {

static const double kInf = std::numeric_limits<double>::infinity();
static const double kU32 = 5.9604644775390625e-08;  // 2^-24, fp32 unit roundoff
// Error bound of an fp32 sample position computed as pc + d*t around the window centre
// (DESIGN.md "fp32 guard bands"): <= 4.5e-7 for |coords| <= 2; 1e-6 leaves a 2x margin.
static const double kEpsPosBase = 1.0e-6;

// ---------------------------------------------------------------------------------------
// mgl64 v1.1.0 helpers restated (go.mod:6; not vendored in the reference tree).
// ---------------------------------------------------------------------------------------
static double len3(const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// Mat3FromCols(v0,v1,v2).Inv(), column-major in and out (objects.go:243).
static void mat3_inv_colmajor(const double* m, double* out) {
    double det = m[0] * m[4] * m[8] + m[3] * m[7] * m[2] + m[6] * m[1] * m[5] - m[6] * m[4] * m[2] - m[3] * m[1] * m[8] -
                 m[0] * m[7] * m[5];
    if (det == 0.0 || std::fabs(det) < 1e-20) {  // mgl64.FloatEqual(det, 0)
        for (int i = 0; i < 9; ++i) out[i] = 0.0;
        return;
    }
    const double a[9] = {
        m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[2] * m[4],
        m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
        m[3] * m[7] - m[4] * m[6], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3],
    };
    double inv = 1 / det;
    for (int i = 0; i < 9; ++i) out[i] = a[i] * inv;
}

static void mat4_mul_cm(const double* a, const double* b, double* r) {  // column-major, mgl64 Mat4.Mul4
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row)
            r[c * 4 + row] = a[0 + row] * b[c * 4 + 0] + a[4 + row] * b[c * 4 + 1] + a[8 + row] * b[c * 4 + 2] +
                             a[12 + row] * b[c * 4 + 3];
}

static double mat4_det_cm(const double* m) {  // mgl64 Mat4.Det
    return m[0] * m[5] * m[10] * m[15] - m[0] * m[5] * m[11] * m[14] - m[0] * m[6] * m[9] * m[15] +
           m[0] * m[6] * m[11] * m[13] + m[0] * m[7] * m[9] * m[14] - m[0] * m[7] * m[10] * m[13] -
           m[1] * m[4] * m[10] * m[15] + m[1] * m[4] * m[11] * m[14] + m[1] * m[6] * m[8] * m[15] -
           m[1] * m[6] * m[11] * m[12] - m[1] * m[7] * m[8] * m[14] + m[1] * m[7] * m[10] * m[12] +
           m[2] * m[4] * m[9] * m[15] - m[2] * m[4] * m[11] * m[13] - m[2] * m[5] * m[8] * m[15] +
           m[2] * m[5] * m[11] * m[12] + m[2] * m[7] * m[8] * m[13] - m[2] * m[7] * m[9] * m[12] -
           m[3] * m[4] * m[9] * m[14] + m[3] * m[4] * m[10] * m[13] + m[3] * m[5] * m[8] * m[14] -
           m[3] * m[5] * m[10] * m[12] - m[3] * m[6] * m[8] * m[13] + m[3] * m[6] * m[9] * m[12];
}

static void mat4_inv_cm(const double* m, double* r) {  // mgl64 Mat4.Inv: cofactors * (1/det)
    double det = mat4_det_cm(m);
    if (det == 0.0 || std::fabs(det) < 1e-20) {
        for (int i = 0; i < 16; ++i) r[i] = 0.0;
        return;
    }
    const double a[16] = {
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
    double inv = 1 / det;
    for (int i = 0; i < 16; ++i) r[i] = a[i] * inv;
}

// main.go:226-239 computeCameraFromAngles.
void camera_from_angles(double az_deg, double polar_deg, double R, double* eye, double* view_rowmajor) {
    double th = az_deg * M_PI / 180.0;  // mgl64.DegToRad
    double phi = polar_deg * M_PI / 180.0;
    eye[0] = R * std::cos(th) * std::sin(phi);
    eye[1] = R * std::sin(th) * std::sin(phi);
    eye[2] = std::cos(phi) * R;
    // LookAtV(eye, 0, (0,0,1))
    double f[3] = {0 - eye[0], 0 - eye[1], 0 - eye[2]};
    double l = 1.0 / len3(f);
    f[0] *= l; f[1] *= l; f[2] *= l;
    double up[3] = {0, 0, 1};
    double lu = 1.0 / len3(up);
    up[0] *= lu; up[1] *= lu; up[2] *= lu;
    double s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    double ls = 1.0 / len3(s);
    s[0] *= ls; s[1] *= ls; s[2] *= ls;
    double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    const double M[16] = {s[0], u[0], -f[0], 0, s[1], u[1], -f[1], 0, s[2], u[2], -f[2], 0, 0, 0, 0, 1};
    const double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -eye[0], -eye[1], -eye[2], 1};
    double look[16], cam[16];
    mat4_mul_cm(M, T, look);
    mat4_inv_cm(look, cam);
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) view_rowmajor[r * 4 + c] = cam[c * 4 + r];  // Mat4.At(r,c)
}

// ---------------------------------------------------------------------------------------
// JSON -> tree (objects.go FromMap semantics; JSON numbers are always float64)
// ---------------------------------------------------------------------------------------
struct ParseCtx {
    std::string err;
    XRayScene* sc;
};

static bool get_num(const JValue& o, const char* key, double& out, ParseCtx& c, const char* what) {
    const JValue* v = o.get(key);
    if (!v || !v->is_num()) {
        c.err = std::string(key) + " is not a " + what;
        return false;
    }
    out = v->num;
    return true;
}
static bool get_vec3(const JValue& o, const char* key, double* out, ParseCtx& c) {
    const JValue* v = o.get(key);
    if (!v || v->kind != JValue::Arr) {
        c.err = std::string(key) + " is not a Vec3";
        return false;
    }
    if (v->arr.size() > 3) {
        c.err = std::string(key) + " has more than 3 elements";
        return false;
    }
    out[0] = out[1] = out[2] = 0.0;
    for (size_t i = 0; i < v->arr.size(); ++i) {
        if (!v->arr[i].is_num()) {
            c.err = std::string(key) + "[" + std::to_string(i) + "] is not a float64";
            return false;
        }
        out[i] = v->arr[i].num;
    }
    return true;
}

static bool parse_object(const JValue& o, Node& n, ParseCtx& c, bool inside_collection);

static bool parse_collection(const JValue& o, Node& n, ParseCtx& c) {
    n.type = N_COLL;
    const JValue* g = o.get("greedy_dens_eval");
    n.greedy = g && g->kind == JValue::Bool && g->b;
    const JValue* objs = o.get("objects");
    if (!objs || objs->kind != JValue::Arr) {
        c.err = "objects is not a list";
        return false;
    }
    n.kids.resize(objs->arr.size());
    for (size_t i = 0; i < objs->arr.size(); ++i)
        if (!parse_object(objs->arr[i], n.kids[i], c, true)) return false;
    return true;
}

static bool parse_bounds(const JValue& o, double* b, ParseCtx& c) {
    static const char* keys[6] = {"xmin", "xmax", "ymin", "ymax", "zmin", "zmax"};
    for (int i = 0; i < 6; ++i)
        if (!get_num(o, keys[i], b[i], c, "float64")) return false;
    return true;
}

static bool parse_object(const JValue& o, Node& n, ParseCtx& c, bool inside_collection) {
    if (o.kind != JValue::Obj) {
        c.err = "object description is not a map";
        return false;
    }
    const JValue* t = o.get("type");
    std::string type = (t && t->kind == JValue::Str) ? t->str : "";
    if (type == "sphere") {
        n.type = N_SPHERE;
        return get_vec3(o, "center", n.p, c) && get_num(o, "radius", n.p[3], c, "float64") &&
               get_num(o, "rho", n.p[4], c, "float64");
    }
    if (type == "cube") {  // objects.go:115: Box{Center, Sides: {Side,Side,Side}, Rho}
        n.type = N_BOX;
        double side;
        if (!get_vec3(o, "center", n.p, c) || !get_num(o, "side", side, c, "float64") ||
            !get_num(o, "rho", n.p[6], c, "float64"))
            return false;
        n.p[3] = n.p[4] = n.p[5] = side;
        return true;
    }
    if (type == "box") {
        n.type = N_BOX;
        return get_vec3(o, "center", n.p, c) && get_vec3(o, "sides", n.p + 3, c) &&
               get_num(o, "rho", n.p[6], c, "float64");
    }
    if (type == "parallelepiped") {
        n.type = N_PPED;
        if (!get_vec3(o, "origin", n.p, c) || !get_vec3(o, "v0", n.p + 3, c) || !get_vec3(o, "v1", n.p + 6, c) ||
            !get_vec3(o, "v2", n.p + 9, c) || !get_num(o, "rho", n.p[12], c, "float64"))
            return false;
        mat3_inv_colmajor(n.p + 3, n.p + 13);  // columns v0,v1,v2 are already column-major
        return true;
    }
    if (type == "cylinder") {
        n.type = N_CYL;
        if (!get_vec3(o, "p0", n.p, c) || !get_vec3(o, "p1", n.p + 3, c) ||
            !get_num(o, "radius", n.p[6], c, "float64"))
            return false;
        if (!o.get("rho"))
            n.p[7] = 1.0;  // objects.go:326-330
        else if (!get_num(o, "rho", n.p[7], c, "float64"))
            return false;
        return true;
    }
    if (type == "gyroid") {
        n.type = N_GYROID;
        return get_vec3(o, "center", n.p, c) && get_num(o, "scale", n.p[3], c, "float64") &&
               get_num(o, "thickness", n.p[4], c, "float64") && get_num(o, "rho", n.p[5], c, "float64");
    }
    if (type == "object_collection") {
        if (inside_collection) {  // objects.go:391-410 has no case for nested collections
            c.err = "unknown object type";
            return false;
        }
        return parse_collection(o, n, c);
    }
    if (type == "tessellated_obj_coll") {
        n.type = N_TESS;
        const JValue* uc = o.get("uc");
        if (!uc || uc->kind != JValue::Obj) {
            c.err = "uc is not a map";
            return false;
        }
        const JValue* objs = uc->get("objects");
        if (!objs || objs->kind != JValue::Obj) {
            c.err = "objects is not a map";
            return false;
        }
        n.kids.resize(1);
        if (!parse_collection(*objs, n.kids[0], c)) return false;
        n.kids[0].greedy = true;  // objects.go:487
        return parse_bounds(*uc, n.p + 6, c) && parse_bounds(o, n.p, c);
    }
    if (type == "voxel_grid") {
        n.type = N_VOXEL;
        int dims[3] = {0, 0, 0};
        const JValue* res = o.get("resolution");
        if (res && res->kind == JValue::Arr && res->arr.size() == 3) {
            for (int i = 0; i < 3; ++i) {
                if (!res->arr[i].is_num()) {
                    c.err = "resolution must be a list of 3 integers";
                    return false;
                }
                dims[i] = (int)res->arr[i].num;
            }
        } else {
            double v[3];
            if (!get_num(o, "nx", v[0], c, "int") || !get_num(o, "ny", v[1], c, "int") ||
                !get_num(o, "nz", v[2], c, "int"))
                return false;
            for (int i = 0; i < 3; ++i) dims[i] = (int)v[i];
        }
        if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1) {
            c.err = "voxel_grid resolution must be positive";
            return false;
        }
        if (c.sc->n_vox >= kMaxVoxelSlots) {
            c.err = "too many voxel_grid nodes (max " + std::to_string(kMaxVoxelSlots) + ")";
            return false;
        }
        n.voxel_slot = c.sc->n_vox++;
        VoxelHost& vh = c.sc->vox[n.voxel_slot];
        vh.nx = dims[0];
        vh.ny = dims[1];
        vh.nz = dims[2];
        return true;
    }
    c.err = inside_collection ? "unknown object type" : "unknown object type `" + type + "`";
    return false;
}

// deformations.go:308-343 NewDeformation; "composed" is flattened into a sequence.
static bool parse_deformation(const JValue& o, std::vector<Deform>& out, ParseCtx& c, int depth) {
    if (o.kind != JValue::Obj) {
        c.err = "deformation description is not a map";
        return false;
    }
    if (depth > 16) {
        c.err = "composed deformations nested too deep";
        return false;
    }
    const JValue* t = o.get("type");
    if (!t || t->kind != JValue::Str) {
        c.err = "deformation type is nil";
        return false;
    }
    auto num_list = [&](const char* key, size_t n, double* dst, const char* msg) -> bool {
        const JValue* v = o.get(key);
        if (!v || v->kind != JValue::Arr) {
            c.err = msg;
            return false;
        }
        if (v->arr.size() < n) {
            c.err = std::string(key) + " needs " + std::to_string(n) + " elements";
            return false;
        }
        for (size_t i = 0; i < n; ++i) {
            if (!v->arr[i].is_num()) {
                c.err = std::string(key) + " elements must be float64";
                return false;
            }
            dst[i] = v->arr[i].num;
        }
        return true;
    };
    Deform d;
    const std::string& type = t->str;
    if (type == "gaussian") {
        d.type = D_GAUSSIAN;
        if (!num_list("amplitudes", 3, d.d, "amplitudes must be a list") ||
            !num_list("sigmas", 3, d.d + 3, "sigmas must be a list") ||
            !num_list("centers", 3, d.d + 6, "centers must be a list"))
            return false;
    } else if (type == "linear") {
        d.type = D_LINEAR;
        if (!num_list("strains", 6, d.d, "strains must be a list")) return false;
    } else if (type == "rigid") {
        d.type = D_RIGID;
        if (!num_list("displacements", 3, d.d, "displacements must be a list")) return false;
    } else if (type == "sigmoid") {
        d.type = D_SIGMOID;
        if (!get_num(o, "amplitude", d.d[0], c, "float")) { c.err = "amplitude must be a float"; return false; }
        if (!get_num(o, "center", d.d[1], c, "float")) { c.err = "center must be a float"; return false; }
        if (!get_num(o, "lengthscale", d.d[2], c, "float")) { c.err = "lengthscale must be a float"; return false; }
        const JValue* dir = o.get("direction");
        if (!dir || dir->kind != JValue::Str) {
            c.err = "direction must be a string";
            return false;
        }
        if (dir->str == "x") d.axis = 0;
        else if (dir->str == "y") d.axis = 1;
        else if (dir->str == "z") d.axis = 2;
        else {
            c.err = "Invalid direction";  // deformations.go:219 (log.Fatal at Apply time)
            return false;
        }
    } else if (type == "affine") {
        d.type = D_AFFINE;
        const JValue* m = o.get("matrix");
        if (!m || m->kind != JValue::Arr) { c.err = "matrix must be a list"; return false; }
        if (m->arr.size() != 3) { c.err = "matrix must have 3 rows"; return false; }
        for (int i = 0; i < 3; ++i) {
            const JValue& row = m->arr[i];
            if (row.kind != JValue::Arr) { c.err = "matrix row must be a list"; return false; }
            if (row.arr.size() != 3) { c.err = "matrix row must have 3 elements"; return false; }
            for (int j = 0; j < 3; ++j) {
                if (!row.arr[j].is_num()) { c.err = "matrix elements must be float64"; return false; }
                d.d[i * 3 + j] = row.arr[j].num;
            }
        }
    } else if (type == "composed") {
        const JValue* l = o.get("deformations");
        if (!l || l->kind != JValue::Arr) { c.err = "deformations must be a list"; return false; }
        for (const auto& sub : l->arr)
            if (!parse_deformation(sub, out, c, depth + 1)) return false;
        return true;
    } else {
        c.err = "unknown deformation type " + type;
        return false;
    }
    out.push_back(d);
    return true;
}

// ---------------------------------------------------------------------------------------
// MinFeatureSize (objects.go:74,181-183,257-259,352-354,440-446,584-586,857-860,1034-1037)
// ---------------------------------------------------------------------------------------
double node_min_feature_size(const Node& n, const XRayScene& sc) {
    switch (n.type) {
        case N_SPHERE: return n.p[3];
        case N_BOX: return 0.1 * std::fmin(n.p[3], std::fmin(n.p[4], n.p[5]));
        case N_PPED: return 0.2 * std::fmin(len3(n.p + 3), std::fmin(len3(n.p + 6), len3(n.p + 9)));
        case N_CYL: return n.p[6];
        case N_GYROID: return n.p[3] * n.p[4] * 0.1;
        case N_COLL: {
            double out = kInf;
            for (const auto& k : n.kids) out = std::fmin(out, node_min_feature_size(k, sc));
            return out;
        }
        case N_TESS: return node_min_feature_size(n.kids[0], sc);
        case N_VOXEL: {
            const VoxelHost& v = sc.vox[n.voxel_slot];
            return 2.0 / double(std::max(v.nx, std::max(v.ny, v.nz)));
        }
    }
    return kInf;
}

// ---------------------------------------------------------------------------------------
// Host fp64 density in reference operation order (tests, volume export of tiny grids).
// ---------------------------------------------------------------------------------------
static double voxel_density_host(const VoxelHost& v, double x, double y, double z) {
    if (!v.data) return 0.0;
    if (x < -1 || x > 1 || y < -1 || y > 1 || z < -1 || z > 1) return 0.0;
    x = (x + 1) / 2; y = (y + 1) / 2; z = (z + 1) / 2;
    x = x * double(v.nx - 1); y = y * double(v.ny - 1); z = z * double(v.nz - 1);
    long x0 = (long)std::floor(x), y0 = (long)std::floor(y), z0 = (long)std::floor(z);
    long x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (z0 < 0) z0 = 0;
    if (x1 >= v.nx) x1 = v.nx - 1;
    if (y1 >= v.ny) y1 = v.ny - 1;
    if (z1 >= v.nz) z1 = v.nz - 1;
    double wx = x - double(x0), wy = y - double(y0), wz = z - double(z0);
    long NX = v.nx, NY = v.ny;
    auto at = [&](long zz, long xx, long yy) -> double {
        size_t idx = (size_t)zz * NX * NY + (size_t)xx * NY + yy;
        return v.dtype == 0 ? (double)((const float*)v.data)[idx] : ((const double*)v.data)[idx];
    };
    double v00 = at(z0, x0, y0) * (1 - wz) + at(z1, x0, y0) * wz;
    double v01 = at(z0, x0, y1) * (1 - wz) + at(z1, x0, y1) * wz;
    double v10 = at(z0, x1, y0) * (1 - wz) + at(z1, x1, y0) * wz;
    double v11 = at(z0, x1, y1) * (1 - wz) + at(z1, x1, y1) * wz;
    double v0 = v00 * (1 - wy) + v01 * wy;
    double v1 = v10 * (1 - wy) + v11 * wy;
    return v0 * (1 - wx) + v1 * wx;
}

static double node_density(const Node& n, const XRayScene& sc, double x, double y, double z) {
    const double* p = n.p;
    switch (n.type) {
        case N_SPHERE: {
            x = x - p[0]; y = y - p[1]; z = z - p[2];
            double r2 = x * x + y * y + z * z;
            return r2 < p[3] * p[3] ? p[4] : 0.0;
        }
        case N_BOX: {
            x = std::fabs(x - p[0]); y = std::fabs(y - p[1]); z = std::fabs(z - p[2]);
            return (x < 0.5 * p[3] && y < 0.5 * p[4] && z < 0.5 * p[5]) ? p[6] : 0.0;
        }
        case N_PPED: {
            double dx = x - p[0], dy = y - p[1], dz = z - p[2];
            const double* m = p + 13;
            double qx = m[0] * dx + m[3] * dy + m[6] * dz;
            double qy = m[1] * dx + m[4] * dy + m[7] * dz;
            double qz = m[2] * dx + m[5] * dy + m[8] * dz;
            return (qx > 0.0 && qx < 1.0 && qy > 0.0 && qy < 1.0 && qz > 0.0 && qz < 1.0) ? p[12] : 0.0;
        }
        case N_CYL: {
            double v[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]};
            double w[3] = {x - p[0], y - p[1], z - p[2]};
            double c = (w[0] * v[0] + w[1] * v[1] + w[2] * v[2]) / (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            if (c < 0.0 || c > 1.0) return 0.0;
            double e[3] = {w[0] - v[0] * c, w[1] - v[1] * c, w[2] - v[2] * c};
            return len3(e) < p[6] ? p[7] : 0.0;
        }
        case N_GYROID: {
            x = (x - p[0]) / p[3]; y = (y - p[1]) / p[3]; z = (z - p[2]) / p[3];
            double g = std::sin(x) * std::cos(y) + std::sin(y) * std::cos(z) + std::sin(z) * std::cos(x);
            return std::fabs(g) < p[4] ? p[5] : 0.0;
        }
        case N_COLL: {
            double density = 0.0;
            for (const auto& k : n.kids) {
                double rho = node_density(k, sc, x, y, z);
                if (n.greedy && rho > 0.0) return rho;
                density += rho;
            }
            if (density < 0.0) density = 0.0;
            else if (density > 1.0) density = 1.0;
            return density;
        }
        case N_TESS: {
            if (x < p[0] || x > p[1] || y < p[2] || y > p[3] || z < p[4] || z > p[5]) return 0.0;
            const double* u = p + 6;
            double dx = u[1] - u[0];
            x = x - dx * std::floor((x - u[0]) / dx);
            double dy = u[3] - u[2];
            y = y - dy * std::floor((y - u[2]) / dy);
            double dz = u[5] - u[4];
            z = z - dz * std::floor((z - u[4]) / dz);
            if (x < u[0] || x > u[1] || y < u[2] || y > u[3] || z < u[4] || z > u[5]) return 0.0;
            return node_density(n.kids[0], sc, x, y, z);
        }
        case N_VOXEL: return voxel_density_host(sc.vox[n.voxel_slot], x, y, z);
    }
    return 0.0;
}

static void apply_deform_host(const Deform& d, double& x, double& y, double& z) {
    const double* p = d.d;
    switch (d.type) {
        case D_GAUSSIAN: {
            double x0 = x - p[6], y0 = y - p[7], z0 = z - p[8];
            double r2 = x0 * x0 + y0 * y0 + z0 * z0;
            double dx = p[0] * std::exp(-r2 / (2 * p[3] * p[3]));
            double dy = p[1] * std::exp(-r2 / (2 * p[4] * p[4]));
            double dz = p[2] * std::exp(-r2 / (2 * p[5] * p[5]));
            x = x + dx; y = y + dy; z = z + dz;
            break;
        }
        case D_AFFINE: {
            double _x = p[0] * x + p[1] * y + p[2] * z;
            double _y = p[3] * x + p[4] * y + p[5] * z;
            double _z = p[6] * x + p[7] * y + p[8] * z;
            x = _x; y = _y; z = _z;
            break;
        }
        case D_LINEAR: {
            double _x = x + p[0] * x + p[5] * y + p[4] * z;
            double _y = y + p[5] * x + p[1] * y + p[3] * z;
            double _z = z + p[4] * x + p[3] * y + p[2] * z;
            x = _x; y = _y; z = _z;
            break;
        }
        case D_RIGID: x = x + p[0]; y = y + p[1]; z = z + p[2]; break;
        case D_SIGMOID: {
            double& q = d.axis == 0 ? x : (d.axis == 1 ? y : z);
            q = q + p[0] / (1 + std::exp(-(q - p[1]) / p[2]));
            break;
        }
    }
}

double host_density(const XRayScene& sc, double x, double y, double z, double dm) {
    for (const auto& d : sc.deforms) apply_deform_host(d, x, y, z);
    return node_density(sc.root, sc, x, y, z) * dm;
}

// ---------------------------------------------------------------------------------------
// Bounds
// ---------------------------------------------------------------------------------------
static void box_add_point(Box3& b, const double* p) {
    if (b.empty) {
        for (int i = 0; i < 3; ++i) b.lo[i] = b.hi[i] = p[i];
        b.empty = false;
        return;
    }
    for (int i = 0; i < 3; ++i) {
        b.lo[i] = std::fmin(b.lo[i], p[i]);
        b.hi[i] = std::fmax(b.hi[i], p[i]);
    }
}
static void box_union(Box3& b, const Box3& o) {
    if (o.empty) return;
    box_add_point(b, o.lo);
    box_add_point(b, o.hi);
}
static Box3 box_inf() {
    Box3 b;
    b.empty = false;
    for (int i = 0; i < 3; ++i) {
        b.lo[i] = -kInf;
        b.hi[i] = kInf;
    }
    return b;
}

// Geometric extent of one node (where it can return non-zero), ignoring rho sign.
static Box3 node_extent(const Node& n) {
    Box3 b;
    const double* p = n.p;
    switch (n.type) {
        case N_SPHERE: {
            double r = std::fabs(p[3]);
            double lo[3] = {p[0] - r, p[1] - r, p[2] - r}, hi[3] = {p[0] + r, p[1] + r, p[2] + r};
            box_add_point(b, lo);
            box_add_point(b, hi);
            break;
        }
        case N_BOX: {
            double lo[3], hi[3];
            for (int i = 0; i < 3; ++i) {
                double h = std::fabs(0.5 * p[3 + i]);
                lo[i] = p[i] - h;
                hi[i] = p[i] + h;
            }
            box_add_point(b, lo);
            box_add_point(b, hi);
            break;
        }
        case N_PPED:
            for (int m = 0; m < 8; ++m) {
                double q[3];
                for (int i = 0; i < 3; ++i)
                    q[i] = p[i] + ((m & 1) ? p[3 + i] : 0) + ((m & 2) ? p[6 + i] : 0) + ((m & 4) ? p[9 + i] : 0);
                box_add_point(b, q);
            }
            break;
        case N_CYL: {
            double r = std::fabs(p[6]);
            for (int e = 0; e < 2; ++e) {
                double lo[3], hi[3];
                for (int i = 0; i < 3; ++i) {
                    lo[i] = p[3 * e + i] - r;
                    hi[i] = p[3 * e + i] + r;
                }
                box_add_point(b, lo);
                box_add_point(b, hi);
            }
            break;
        }
        case N_GYROID: return box_inf();
        case N_VOXEL: {
            double lo[3] = {-1, -1, -1}, hi[3] = {1, 1, 1};
            box_add_point(b, lo);
            box_add_point(b, hi);
            break;
        }
        case N_TESS: {
            double lo[3] = {p[0], p[2], p[4]}, hi[3] = {p[1], p[3], p[5]};
            if (lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]) {
                box_add_point(b, lo);
                box_add_point(b, hi);
            }
            break;
        }
        case N_COLL:
            for (const auto& k : n.kids) box_union(b, node_extent(k));
            break;
    }
    return b;
}

static double node_rho(const Node& n) {
    switch (n.type) {
        case N_SPHERE: return n.p[4];
        case N_BOX: return n.p[6];
        case N_PPED: return n.p[12];
        case N_CYL: return n.p[7];
        case N_GYROID: return n.p[5];
        default: return 1.0;
    }
}

// Region where Density() can be non-zero.  For a collection only children that can return a
// positive value matter: sum-and-clamp and greedy both yield 0 where no child is positive.
static Box3 node_nonzero_region(const Node& n, bool as_child) {
    if (n.type == N_COLL) {
        Box3 b;
        for (const auto& k : n.kids) box_union(b, node_nonzero_region(k, true));
        return b;
    }
    if (n.type == N_TESS) {
        Box3 inner = node_nonzero_region(n.kids[0], true);
        if (inner.empty) return Box3();
        return node_extent(n);
    }
    double rho = node_rho(n);
    if (as_child ? !(rho > 0.0) : (rho == 0.0)) return Box3();
    return node_extent(n);
}

// Pull a box back through the deformation chain: returns W with p not in W => warp(p) not in B.
static Box3 pull_back(Box3 b, const std::vector<Deform>& ds) {
    if (b.empty) return b;
    for (int s = (int)ds.size() - 1; s >= 0; --s) {
        const Deform& d = ds[s];
        switch (d.type) {
            case D_RIGID:
                for (int i = 0; i < 3; ++i) {
                    b.lo[i] -= d.d[i];
                    b.hi[i] -= d.d[i];
                }
                break;
            case D_SIGMOID: {
                double A = d.d[0];
                b.lo[d.axis] -= std::fmax(A, 0.0);
                b.hi[d.axis] -= std::fmin(A, 0.0);
                break;
            }
            case D_GAUSSIAN:
                for (int i = 0; i < 3; ++i) {
                    b.lo[i] -= std::fmax(d.d[i], 0.0);
                    b.hi[i] -= std::fmin(d.d[i], 0.0);
                }
                break;
            case D_AFFINE:
            case D_LINEAR: {
                double M[9];
                if (d.type == D_AFFINE) {
                    for (int i = 0; i < 9; ++i) M[i] = d.d[i];
                } else {
                    const double* e = d.d;
                    double L[9] = {1 + e[0], e[5], e[4], e[5], 1 + e[1], e[3], e[4], e[3], 1 + e[2]};
                    for (int i = 0; i < 9; ++i) M[i] = L[i];
                }
                // row-major M -> column-major for the inverse helper
                double cm[9] = {M[0], M[3], M[6], M[1], M[4], M[7], M[2], M[5], M[8]}, icm[9];
                double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
                             M[2] * (M[3] * M[7] - M[4] * M[6]);
                bool finite = true;
                for (int i = 0; i < 3; ++i) finite = finite && std::isfinite(b.lo[i]) && std::isfinite(b.hi[i]);
                if (!finite || std::fabs(det) < 1e-9) return box_inf();
                mat3_inv_colmajor(cm, icm);
                Box3 nb;
                for (int m = 0; m < 8; ++m) {
                    double q[3] = {(m & 1) ? b.hi[0] : b.lo[0], (m & 2) ? b.hi[1] : b.lo[1], (m & 4) ? b.hi[2] : b.lo[2]};
                    double r[3] = {icm[0] * q[0] + icm[3] * q[1] + icm[6] * q[2], icm[1] * q[0] + icm[4] * q[1] + icm[7] * q[2],
                                   icm[2] * q[0] + icm[5] * q[1] + icm[8] * q[2]};
                    box_add_point(nb, r);
                }
                for (int i = 0; i < 3; ++i) {  // rounding slack of the inverse
                    double ext = std::fmax(std::fabs(nb.lo[i]), std::fabs(nb.hi[i]));
                    nb.lo[i] -= 1e-9 * (1 + ext);
                    nb.hi[i] += 1e-9 * (1 + ext);
                }
                b = nb;
                break;
            }
        }
    }
    return b;
}

// Lipschitz factor / evaluation error of the fp32 warp chain -> position error bound.
static double deform_eps_pos(const std::vector<Deform>& ds, double* lip_out = nullptr) {
    double ep = kEpsPosBase;
    double lip_total = 1.0;
    for (const auto& d : ds) {
        double lip = 1.0, add = 0.0;
        switch (d.type) {
            case D_RIGID: add = 4 * kU32 * 4.0; break;
            case D_SIGMOID:
                lip = 1.0 + std::fabs(d.d[0] / (4.0 * d.d[2]));
                // A*sigma through __expf (2 + 1.173|a| ulp on e, damped by sigma(1-sigma) <= e^-|a|) and __fdividef (2 ulp):
                // <= 4e-7 |A|; budget 1e-6 |A|.  Plus the rounding of the final add at |x| <= 4 (api.cu checks the cameras).
                add = std::fabs(d.d[0]) * 1e-6 + 2 * kU32 * 4.0;
                break;
            case D_GAUSSIAN: {
                double g = 0.0, a = 0.0;
                for (int i = 0; i < 3; ++i) {
                    g = std::fmax(g, std::fabs(d.d[i]) * 0.8578 / std::fabs(d.d[3 + i]));
                    a = std::fmax(a, std::fabs(d.d[i]));
                }
                lip = 1.0 + 1.7321 * g;
                add = a * 4e-6 + 4 * kU32 * 4.0;
                break;
            }
            case D_AFFINE:
            case D_LINEAR: {
                double rows = 0.0;
                for (int r = 0; r < 3; ++r) {
                    double s;
                    if (d.type == D_AFFINE) s = std::fabs(d.d[3 * r]) + std::fabs(d.d[3 * r + 1]) + std::fabs(d.d[3 * r + 2]);
                    else {
                        const double* e = d.d;
                        const double L[9] = {1 + e[0], e[5], e[4], e[5], 1 + e[1], e[3], e[4], e[3], 1 + e[2]};
                        s = std::fabs(L[3 * r]) + std::fabs(L[3 * r + 1]) + std::fabs(L[3 * r + 2]);
                    }
                    rows = std::fmax(rows, s);
                }
                lip = rows;
                add = 8 * kU32 * 4.0 * rows;
                break;
            }
        }
        ep = ep * lip + add;
        lip_total *= std::fmax(lip, 1e-3);
    }
    if (lip_out) *lip_out = lip_total;
    return ep;
}

// Second-order data of the warp chain W for the gyroid skip bound (render_fast.cu prim_gyroid_so): an upper
// bound jac2 of the Jacobian's 2-norm and curv of |D^2 W[d,d]| for unit d.  Returns false when the chain
// has no such bound here (gaussian bumps, composed chains): those keep the first-order Lipschitz rule.
static bool deform_second_order(const std::vector<Deform>& ds, double* jac2, double* curv) {
    *jac2 = 1.0;
    *curv = 0.0;
    if (ds.empty()) return true;
    if (ds.size() != 1) return false;
    const Deform& d = ds[0];
    switch (d.type) {
        case D_RIGID: return true;
        case D_SIGMOID:
            if (!(std::fabs(d.d[2]) > 0.0)) return false;
            *jac2 = 1.0 + std::fabs(d.d[0] / (4.0 * d.d[2]));          // sigma' <= 1/4
            *curv = std::fabs(d.d[0]) * 0.0962251 / (d.d[2] * d.d[2]);  // |sigma''| <= 1/(6 sqrt 3)
            return true;
        case D_AFFINE:
        case D_LINEAR: {
            double fro = 0.0;
            if (d.type == D_AFFINE) {
                for (int i = 0; i < 9; ++i) fro += d.d[i] * d.d[i];
            } else {
                const double* e = d.d;
                const double L[9] = {1 + e[0], e[5], e[4], e[5], 1 + e[1], e[3], e[4], e[3], 1 + e[2]};
                for (int i = 0; i < 9; ++i) fro += L[i] * L[i];
            }
            *jac2 = std::sqrt(fro);  // Frobenius >= 2-norm
            return std::isfinite(*jac2);
        }
        default: return false;
    }
}

// ---------------------------------------------------------------------------------------
// Flattening
// ---------------------------------------------------------------------------------------
struct Builder {
    std::vector<Instr> instr;
    std::vector<float> f32;     // multiple of 4
    std::vector<double> f64;
    std::vector<uint64_t> grids;
    int save_depth = 0, save_depth_max = 0;
    double ep = kEpsPosBase;
    double warp_lip = 1.0;                    // inf-norm Lipschitz factor of the warp chain (deform_eps_pos)
    bool so_ok = false;                       // deform_second_order
    double so_jac2 = 1.0, so_curv = 0.0;
    const XRayScene* sc = nullptr;
    std::string err;

    uint32_t f32_idx() const { return (uint32_t)(f32.size() / 4); }
    uint32_t f64_idx() const { return (uint32_t)f64.size(); }
    void f4(double a, double b, double c, double d) {
        f32.push_back((float)a);
        f32.push_back((float)b);
        f32.push_back((float)c);
        f32.push_back((float)d);
    }
    void f4bits(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        uint32_t v[4] = {a, b, c, d};
        for (int i = 0; i < 4; ++i) {
            float f;
            memcpy(&f, &v[i], 4);
            f32.push_back(f);
        }
    }
    void d64(const double* p, int n, int padded) {
        for (int i = 0; i < padded; ++i) f64.push_back(i < n ? p[i] : 0.0);
    }
};

// Upward-rounded float: tolerances must never shrink when narrowed.
static double up32(double v) { return (double)std::nextafterf((float)v, INFINITY); }

static void emit_prim_params(Builder& B, const Node& n) {
    const double* p = n.p;
    const double ep = B.ep, u = kU32;
    switch (n.type) {
        case N_SPHERE: {
            double r = p[3];
            double tol = 2.0 * (2.0 * 1.7321 * std::fabs(r) * ep + 3.0 * ep * ep + 4.0 * u * r * r);
            B.f4(p[0], p[1], p[2], p[4]);
            B.f4(r * r, up32(tol), 0, 0);
            B.d64(p, 5, kF64Sphere);
            break;
        }
        case N_BOX: {
            double cmax = std::fmax(std::fabs(p[0]), std::fmax(std::fabs(p[1]), std::fabs(p[2])));
            double hmax = 0.5 * std::fmax(std::fabs(p[3]), std::fmax(std::fabs(p[4]), std::fabs(p[5])));
            double tol = 2.0 * (ep + 2.0 * u * (cmax + hmax));
            B.f4(p[0], p[1], p[2], p[6]);
            B.f4(0.5 * p[3], 0.5 * p[4], 0.5 * p[5], up32(tol));
            B.d64(p, 7, kF64Box);
            break;
        }
        case N_CYL: {
            double v[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]};
            double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            double lv = std::sqrt(vv), r = std::fabs(p[6]);
            double inv_vv = vv > 0 ? 1.0 / vv : 0.0;
            double wmax = lv + r + 8 * ep;
            double tolc_raw = vv > 0 ? (1.7321 * ep * lv + 6.0 * u * wmax * lv) / vv + 4.0 * u : 1.0;
            double tolc = 2.0 * tolc_raw;
            // d^2 = |e|^2, e = w - v*c.  The exact e is orthogonal to v, so the error of c moves the computed e
            // along the axis and enters d^2 only at second order; first order is 2*r*|dw| (position + rounding).
            double e2 = 1.7321 * ep + lv * tolc_raw + 6.0 * u * wmax;
            double tolr = 2.0 * (2.0 * r * (1.7321 * ep + 6.0 * u * wmax) + e2 * e2 + 6.0 * u * r * r);
            B.f4(p[0], p[1], p[2], p[7]);
            B.f4(v[0], v[1], v[2], inv_vv);
            B.f4(p[6] * p[6], up32(tolr), up32(tolc), up32(tolc) / up32(tolr));  // .w rescales the radial slack onto tolc
            B.d64(p, 8, kF64Cyl);
            break;
        }
        case N_PPED: {
            const double* m = p + 13;  // column-major inverse
            double ext = len3(p + 3) + len3(p + 6) + len3(p + 9) + 8 * ep;
            double tol = 0.0;
            for (int r = 0; r < 3; ++r) {
                double s = std::fabs(m[r]) + std::fabs(m[3 + r]) + std::fabs(m[6 + r]);
                tol = std::fmax(tol, 2.0 * (s * (ep + 6.0 * u * ext) + 4.0 * u));
            }
            B.f4(p[0], p[1], p[2], p[12]);
            B.f4(m[0], m[3], m[6], up32(tol));
            B.f4(m[1], m[4], m[7], 0);
            B.f4(m[2], m[5], m[8], 0);
            double rec[13];
            for (int i = 0; i < 3; ++i) rec[i] = p[i];
            for (int i = 0; i < 9; ++i) rec[3 + i] = m[i];
            rec[12] = p[12];
            B.d64(rec, 13, kF64Pped);
            break;
        }
        case N_GYROID: {
            double scale = p[3];
            double cmax = std::fmax(std::fabs(p[0]), std::fmax(std::fabs(p[1]), std::fabs(p[2])));
            double amax = (3.0 + cmax) / std::fabs(scale);  // |argument| bound inside the render window
            double argerr = ep / std::fabs(scale) + 3.0 * u * amax;
            // + 1e-6 per sin/cos: SFU evaluation after Cody-Waite reduction (eval.cuh fast_sincos)
            // argument errors: d(sin a cos b) <= da|cos a cos b| + db|sin a sin b| <= max(da,db) per term (3 terms);
            // function errors: 6 SFU evaluations, each multiplied by a factor <= 1
            // The sum below is a worst-case bound of |g_fp32 - g| (the largest error observed on the gyroid + sigmoid
            // config is 40x smaller); 1.1 covers the rounding of the bound itself.  The other primitives keep a factor 2:
            // their bands cost nothing measurable, this one decides how many refined sub-steps (g moves 4e-4 per
            // sub-step there) need the fp64 reference.
            double tol = 1.1 * (3.0 * argerr + 6.0 * (1.0e-6 + 4.0 * u * (1.0 + 1e-2 * amax)) + 12.0 * u);
#ifdef XRAY_DEV_KNOBS  // development builds only (make EXTRA=-DXRAY_DEV_KNOBS): unsound below 1, never in the release library
            if (const char* e = getenv("XRAY_DEBUG_GYROID_TOL_SCALE")) tol *= atof(e);
#endif
            B.f4(p[0], p[1], p[2], p[5]);
            // .w: object-space (max-norm) distance per unit of |g| margin: sum_i |dg/dq_i| <= 3 (max at q = 0)
            B.f4(1.0 / scale, p[4], up32(tol), std::fabs(scale) / 3.03);
            // Along a ray, h(t) = g(q(t)), q = (W(o + d t) - c)/scale:  |h''| <= 2 |q'|^2 + |grad g . q''|
            // (each term sin a cos b has a Hessian of 2-norm <= 1 on its coordinate pair and every coordinate
            // sits in two terms; |dg/dq_i| <= 2).  So |h(t + tau) - h(t)| <= |h'(t)| tau + M2 tau^2 / 2.
            {
                const double as = std::fabs(scale);
                const double M2 = B.so_ok ? 1.01 * (2.0 * (B.so_jac2 / as) * (B.so_jac2 / as) + 3.0 * B.so_curv / as) : 0.0;
                // g1eps: absolute slack added to |grad g . J d| (fp32 evaluation of gradient and Jacobian)
                B.f4(M2 > 0.0 && std::isfinite(M2) ? up32(M2) : 0.0, up32(1e-3 * B.so_jac2), up32(std::fmax(1.0, B.warp_lip)), 0);  // exactly 0 = rule off
            }
            B.d64(p, 6, kF64Gyroid);
            break;
        }
        default: break;
    }
}

static bool is_prim(NodeType t) { return t == N_SPHERE || t == N_BOX || t == N_PPED || t == N_CYL || t == N_GYROID; }
static Op prim_op(NodeType t) {
    switch (t) {
        case N_SPHERE: return OP_SPHERE;
        case N_BOX: return OP_BOX;
        case N_PPED: return OP_PPED;
        case N_CYL: return OP_CYL;
        default: return OP_GYROID;
    }
}

static double seg_point_dist(const double* a, const double* b, const double* q) {
    double v[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, w[3] = {q[0] - a[0], q[1] - a[1], q[2] - a[2]};
    double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double t = vv > 0 ? (w[0] * v[0] + w[1] * v[1] + w[2] * v[2]) / vv : 0.0;
    t = std::fmin(1.0, std::fmax(0.0, t));
    double e[3] = {w[0] - v[0] * t, w[1] - v[1] * t, w[2] - v[2] * t};
    return len3(e);
}

// Does child k possibly return non-zero somewhere in the (inflated) cell [lo,hi]?
static bool child_touches_cell(const Node& k, const double* lo, const double* hi) {
    switch (k.type) {
        case N_SPHERE: {
            double d2 = 0;
            for (int i = 0; i < 3; ++i) {
                double c = k.p[i];
                double d = c < lo[i] ? lo[i] - c : (c > hi[i] ? c - hi[i] : 0.0);
                d2 += d * d;
            }
            return d2 <= k.p[3] * k.p[3];
        }
        case N_CYL: {
            double c[3], hd = 0;
            for (int i = 0; i < 3; ++i) {
                c[i] = 0.5 * (lo[i] + hi[i]);
                double h = 0.5 * (hi[i] - lo[i]);
                hd += h * h;
            }
            return seg_point_dist(k.p, k.p + 3, c) <= std::fabs(k.p[6]) + std::sqrt(hd);
        }
        case N_GYROID: return true;
        default: {
            Box3 e = node_extent(k);
            if (e.empty) return false;
            for (int i = 0; i < 3; ++i)
                if (e.hi[i] < lo[i] || e.lo[i] > hi[i]) return false;
            return true;
        }
    }
}

static int g_grid_max = 32;  // the grid holds at most g_grid_max^3 cells, at most 4*g_grid_max per axis (env XRAY_GRID_MAX)
static int g_grid_min_children = 4;
static double g_grid_feat_scale = 0.25;  // cell size in units of the smallest child feature (env XRAY_GRID_FEAT_SCALE)

// Child-mask grid for a collection: cell -> 64-bit set of children that may be non-zero there.
// region: where the collection is evaluated (unit-cell bounds) or null (use children extents).
static bool build_grid(Builder& B, const Node& coll, const Box3* region, uint32_t& f32_idx, uint32_t& grid_idx) {
    size_t n = coll.kids.size();
    // a unit-cell collection is worth a grid even for one child: the grid also drives empty-space skipping
    if ((int)n < (region ? 1 : g_grid_min_children) || n > 63) return false;  // (bit 63 of a mask marks an empty cell)
    Box3 reg;
    if (region) reg = *region;
    else
        for (const auto& k : coll.kids) {
            Box3 e = node_extent(k);
            bool finite = !e.empty;
            for (int i = 0; i < 3 && finite; ++i) finite = std::isfinite(e.lo[i]) && std::isfinite(e.hi[i]);
            if (finite) box_union(reg, e);
        }
    if (reg.empty) return false;
    double ext[3], feat = kInf, emax = 0;
    for (int i = 0; i < 3; ++i) {
        if (!std::isfinite(reg.lo[i]) || !std::isfinite(reg.hi[i])) return false;
        ext[i] = reg.hi[i] - reg.lo[i];
        if (!(ext[i] > 0)) return false;
        emax = std::fmax(emax, ext[i]);
    }
    for (const auto& k : coll.kids) switch (k.type) {
            case N_SPHERE: feat = std::fmin(feat, std::fabs(k.p[3])); break;
            case N_CYL: feat = std::fmin(feat, std::fabs(k.p[6])); break;
            case N_BOX: feat = std::fmin(feat, 0.5 * std::fmin(std::fabs(k.p[3]), std::fmin(std::fabs(k.p[4]), std::fabs(k.p[5])))); break;
            case N_PPED: feat = std::fmin(feat, 0.5 * std::fmin(len3(k.p + 3), std::fmin(len3(k.p + 6), len3(k.p + 9)))); break;
            default: break;
        }
    if (!std::isfinite(feat)) return false;  // no bounded child: every cell would list every child
    if (!(feat > 0)) feat = emax / 8;
    // cubic cells: as fine as the smallest feature asks for, within a budget of g_grid_max^3 cells
    double cell = std::fmax(std::fmax(feat * g_grid_feat_scale, emax / (4.0 * g_grid_max)),
                            std::cbrt(ext[0] * ext[1] * ext[2] / ((double)g_grid_max * g_grid_max * g_grid_max)));
    int g[3];
    for (int i = 0; i < 3; ++i) g[i] = std::max(1, std::min(4 * g_grid_max, (int)std::ceil(ext[i] / cell - 1e-9)));
    if (g[0] * g[1] * g[2] <= 1) return false;
    double margin = 1e-5 + 8.0 * B.ep;
    grid_idx = (uint32_t)B.grids.size();
    B.grids.resize(B.grids.size() + (size_t)g[0] * g[1] * g[2], 0);
    double cs[3] = {ext[0] / g[0], ext[1] / g[1], ext[2] / g[2]};
    for (int iz = 0; iz < g[2]; ++iz)
        for (int iy = 0; iy < g[1]; ++iy)
            for (int ix = 0; ix < g[0]; ++ix) {
                double lo[3] = {reg.lo[0] + ix * cs[0] - margin, reg.lo[1] + iy * cs[1] - margin, reg.lo[2] + iz * cs[2] - margin};
                double hi[3] = {reg.lo[0] + (ix + 1) * cs[0] + margin, reg.lo[1] + (iy + 1) * cs[1] + margin,
                                reg.lo[2] + (iz + 1) * cs[2] + margin};
                uint64_t m = 0;
                for (size_t c = 0; c < n; ++c)
                    if (child_touches_cell(coll.kids[c], lo, hi)) m |= (uint64_t)1 << c;
                B.grids[grid_idx + ((size_t)iz * g[1] + iy) * g[0] + ix] = m;
            }
    // Empty cells carry a skip distance instead of a mask: bit 63 set, low byte = Chebyshev distance (in
    // cells, >= 1, capped) to the nearest cell any child can touch; periodic when the grid tiles a unit cell.
    // Needs bit 63 free, i.e. at most 63 children.
    if (n <= 63) {
        const int gx = g[0], gy = g[1], gz = g[2];
        const size_t nc = (size_t)gx * gy * gz;
        std::vector<int> dist(nc, -1);
        std::vector<size_t> frontier, next;
        for (size_t c = 0; c < nc; ++c)
            if (B.grids[grid_idx + c] != 0) {
                dist[c] = 0;
                frontier.push_back(c);
            }
        const bool periodic = region != nullptr;
        int d = 0;
        while (!frontier.empty()) {
            ++d;
            next.clear();
            for (size_t c : frontier) {
                const int ix = (int)(c % gx), iy = (int)((c / gx) % gy), iz = (int)(c / ((size_t)gx * gy));
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            int jx = ix + dx, jy = iy + dy, jz = iz + dz;
                            if (periodic) {
                                jx = (jx + gx) % gx;
                                jy = (jy + gy) % gy;
                                jz = (jz + gz) % gz;
                            } else if (jx < 0 || jy < 0 || jz < 0 || jx >= gx || jy >= gy || jz >= gz) {
                                continue;
                            }
                            const size_t q = ((size_t)jz * gy + jy) * gx + jx;
                            if (dist[q] < 0) {
                                dist[q] = d;
                                next.push_back(q);
                            }
                        }
            }
            frontier.swap(next);
        }
        for (size_t c = 0; c < nc; ++c)
            if (B.grids[grid_idx + c] == 0) {
                const int dd = dist[c] < 0 ? 255 : std::min(dist[c], 255);
                B.grids[grid_idx + c] = (1ull << 63) | (uint64_t)dd;
            }
    }
    f32_idx = B.f32_idx();
    B.f4(reg.lo[0], reg.lo[1], reg.lo[2], std::fmin(cs[0], std::fmin(cs[1], cs[2])));  // .w = smallest cell edge
    B.f4(1.0 / cs[0], 1.0 / cs[1], 1.0 / cs[2], 0);
    B.f4bits((uint32_t)g[0], (uint32_t)g[1], (uint32_t)g[2], 0);
    B.f4(g[0], g[1], g[2], 0);  // the same dims as floats (saves three I2F per sample)
    // fp64 copy for the exact path: gmin(3), inv_cell(3)
    return true;
}

// Cell-list grid for big collections of primitives (> 63 children, so no 64-bit mask): per cell the
// ascending list of children that may be non-zero there.  The kernel merges the lists of a warp's lanes
// with a min-reduction, which visits the union in child order (greedy / summation order preserved).
static bool build_list_grid(Builder& B, const Node& coll, const Box3* region, const std::vector<uint32_t>& child_f32,
                            const std::vector<uint32_t>& child_f64, const std::vector<uint32_t>& child_op, uint32_t& f32_idx,
                            uint32_t& grid_idx) {
    const size_t n = coll.kids.size();
    if (n > 65000) return false;
    Box3 reg;
    double feat = kInf;
    std::vector<Box3> ext(n);
    for (size_t c = 0; c < n; ++c) {
        const Node& k = coll.kids[c];
        ext[c] = node_extent(k);
        if (ext[c].empty) continue;
        for (int a = 0; a < 3; ++a)
            if (!std::isfinite(ext[c].lo[a]) || !std::isfinite(ext[c].hi[a])) return false;  // unbounded child (gyroid)
        if (!region) box_union(reg, ext[c]);
        switch (k.type) {
            case N_SPHERE: feat = std::fmin(feat, std::fabs(k.p[3])); break;
            case N_CYL: feat = std::fmin(feat, std::fabs(k.p[6])); break;
            case N_BOX: feat = std::fmin(feat, 0.5 * std::fmin(std::fabs(k.p[3]), std::fmin(std::fabs(k.p[4]), std::fabs(k.p[5])))); break;
            case N_PPED: feat = std::fmin(feat, 0.5 * std::fmin(len3(k.p + 3), std::fmin(len3(k.p + 6), len3(k.p + 9)))); break;
            default: break;
        }
    }
    if (region) reg = *region;
    if (reg.empty || !std::isfinite(feat)) return false;
    double e3[3], emax = 0;
    for (int a = 0; a < 3; ++a) {
        e3[a] = reg.hi[a] - reg.lo[a];
        if (!(e3[a] > 0) || !std::isfinite(e3[a])) return false;
        emax = std::fmax(emax, e3[a]);
    }
    if (!(feat > 0)) feat = emax / 8;
    const double cell = std::fmax(std::fmax(feat * g_grid_feat_scale, emax / (4.0 * g_grid_max)),
                                  std::cbrt(e3[0] * e3[1] * e3[2] / ((double)g_grid_max * g_grid_max * g_grid_max)));
    int g[3];
    for (int a = 0; a < 3; ++a) g[a] = std::max(1, std::min(4 * g_grid_max, (int)std::ceil(e3[a] / cell - 1e-9)));
    const size_t ncell = (size_t)g[0] * g[1] * g[2];
    if (ncell <= 1) return false;
    const double cs[3] = {e3[0] / g[0], e3[1] / g[1], e3[2] / g[2]};
    const double margin = 1e-5 + 8.0 * B.ep;
    std::vector<std::vector<uint16_t>> lists(ncell);
    for (size_t c = 0; c < n; ++c) {
        if (ext[c].empty) continue;
        int lo_i[3], hi_i[3];
        for (int a = 0; a < 3; ++a) {
            lo_i[a] = std::max(0, std::min(g[a] - 1, (int)std::floor((ext[c].lo[a] - margin - reg.lo[a]) / cs[a])));
            hi_i[a] = std::max(0, std::min(g[a] - 1, (int)std::floor((ext[c].hi[a] + margin - reg.lo[a]) / cs[a])));
            // children hanging out of the region still have to be listed in the border cells (positions are clamped)
        }
        for (int iz = lo_i[2]; iz <= hi_i[2]; ++iz)
            for (int iy = lo_i[1]; iy <= hi_i[1]; ++iy)
                for (int ix = lo_i[0]; ix <= hi_i[0]; ++ix) {
                    double lo[3] = {reg.lo[0] + ix * cs[0] - margin, reg.lo[1] + iy * cs[1] - margin, reg.lo[2] + iz * cs[2] - margin};
                    double hi[3] = {reg.lo[0] + (ix + 1) * cs[0] + margin, reg.lo[1] + (iy + 1) * cs[1] + margin,
                                    reg.lo[2] + (iz + 1) * cs[2] + margin};
                    // border cells also stand in for everything beyond them (clamped lookups)
                    if (ix == 0) lo[0] = -kInf;
                    if (iy == 0) lo[1] = -kInf;
                    if (iz == 0) lo[2] = -kInf;
                    if (ix == g[0] - 1) hi[0] = kInf;
                    if (iy == g[1] - 1) hi[1] = kInf;
                    if (iz == g[2] - 1) hi[2] = kInf;
                    bool touch;
                    if (coll.kids[c].type == N_CYL && std::isfinite(lo[0] + lo[1] + lo[2] + hi[0] + hi[1] + hi[2]))
                        touch = child_touches_cell(coll.kids[c], lo, hi);
                    else if (coll.kids[c].type == N_SPHERE)
                        touch = child_touches_cell(coll.kids[c], lo, hi);
                    else {
                        touch = true;
                        for (int a = 0; a < 3; ++a)
                            if (ext[c].hi[a] < lo[a] || ext[c].lo[a] > hi[a]) touch = false;
                    }
                    if (touch) lists[((size_t)iz * g[1] + iy) * g[0] + ix].push_back((uint16_t)c);
                }
    }
    // Chebyshev distance (cells) from each empty cell to the nearest non-empty one
    std::vector<int> dist(ncell, -1);
    std::vector<size_t> frontier, next;
    for (size_t c = 0; c < ncell; ++c)
        if (!lists[c].empty()) {
            dist[c] = 0;
            frontier.push_back(c);
        }
    const bool periodic = region != nullptr;
    int dd = 0;
    while (!frontier.empty()) {
        ++dd;
        next.clear();
        for (size_t c : frontier) {
            const int ix = (int)(c % g[0]), iy = (int)((c / g[0]) % g[1]), iz = (int)(c / ((size_t)g[0] * g[1]));
            for (int dz = -1; dz <= 1; ++dz)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        int jx = ix + dx, jy = iy + dy, jz = iz + dz;
                        if (periodic) {
                            jx = (jx + g[0]) % g[0];
                            jy = (jy + g[1]) % g[1];
                            jz = (jz + g[2]) % g[2];
                        } else if (jx < 0 || jy < 0 || jz < 0 || jx >= g[0] || jy >= g[1] || jz >= g[2]) {
                            continue;
                        }
                        const size_t q = ((size_t)jz * g[1] + jy) * g[0] + jx;
                        if (dist[q] < 0) {
                            dist[q] = dd;
                            next.push_back(q);
                        }
                    }
        }
        frontier.swap(next);
    }
    // pack: cell_off (u32), cell_dist (u8), idx (u16), child_tab (u32) into the u64 pool
    std::vector<uint32_t> off(ncell + 1, 0);
    size_t total = 0;
    for (size_t c = 0; c < ncell; ++c) {
        off[c] = (uint32_t)total;
        total += lists[c].size();
    }
    off[ncell] = (uint32_t)total;
    grid_idx = (uint32_t)B.grids.size();
    auto words = [](size_t bytes) { return (bytes + 7) / 8; };
    const size_t w_off = words((ncell + 1) * 4), w_dist = words(ncell), w_idx = words(std::max<size_t>(1, total) * 2), w_tab = words(n * 4);
    const size_t base_off = 0, base_dist = w_off, base_idx = w_off + w_dist, base_tab = w_off + w_dist + w_idx, base_tab64 = base_tab + w_tab;
    B.grids.resize(B.grids.size() + w_off + w_dist + w_idx + 2 * w_tab, 0);
    uint8_t* raw = reinterpret_cast<uint8_t*>(B.grids.data() + grid_idx);
    memcpy(raw + base_off * 8, off.data(), (ncell + 1) * 4);
    for (size_t c = 0; c < ncell; ++c) raw[base_dist * 8 + c] = (uint8_t)(dist[c] < 0 ? 255 : std::min(dist[c], 255));
    uint16_t* idx = reinterpret_cast<uint16_t*>(raw + base_idx * 8);
    size_t w = 0;
    for (size_t c = 0; c < ncell; ++c)
        for (uint16_t v : lists[c]) idx[w++] = v;
    uint32_t* tab = reinterpret_cast<uint32_t*>(raw + base_tab * 8);
    for (size_t c = 0; c < n; ++c) tab[c] = (child_op[c] << 24) | (child_f32[c] & 0xFFFFFFu);
    uint32_t* tab64 = reinterpret_cast<uint32_t*>(raw + base_tab64 * 8);
    for (size_t c = 0; c < n; ++c) tab64[c] = child_f64[c];
    f32_idx = B.f32_idx();
    B.f4(reg.lo[0], reg.lo[1], reg.lo[2], std::fmin(cs[0], std::fmin(cs[1], cs[2])));
    B.f4(1.0 / cs[0], 1.0 / cs[1], 1.0 / cs[2], 0);
    B.f4bits((uint32_t)g[0], (uint32_t)g[1], (uint32_t)g[2], 0);
    B.f4(g[0], g[1], g[2], 0);
    B.f4bits((uint32_t)base_off, (uint32_t)base_dist, (uint32_t)base_idx, (uint32_t)base_tab);
    B.f4bits((uint32_t)base_tab64, 0, 0, 0);
    return true;
}

static bool emit_node(Builder& B, const Node& n, bool nosave, uint32_t child_bit);

static bool emit_collection(Builder& B, const Node& coll, bool nosave, uint32_t child_bit, const Box3* region) {
    size_t begin = B.instr.size();
    Instr ib = {};
    ib.op = OP_COLL_BEGIN;
    ib.n = (uint32_t)coll.kids.size();
    ib.flags = (coll.greedy ? F_GREEDY : 0) | (nosave ? F_NOSAVE : 0);
    ib.child_bit = child_bit;
    uint32_t gf = 0, gi = 0;
    if (build_grid(B, coll, region, gf, gi)) {
        ib.flags |= F_HAS_GRID;
        ib.f32_idx = gf;
        ib.aux = gi;
    }
    B.instr.push_back(ib);
    if (!nosave) {
        if (++B.save_depth > kMaxSaveDepth) {
            B.err = "scene nests too deeply for the device evaluator";
            return false;
        }
        B.save_depth_max = std::max(B.save_depth_max, B.save_depth);
    }
    size_t i = 0, nk = coll.kids.size();
    std::vector<uint32_t> child_f32(nk, 0), child_f64(nk, 0), child_op(nk, 0);
    bool all_prims = true;
    while (i < nk) {
        const Node& k = coll.kids[i];
        if (is_prim(k.type)) {
            size_t j = i;
            Instr ir = {};
            ir.op = prim_op(k.type);
            ir.child_bit = (uint32_t)i;
            ir.f32_idx = B.f32_idx();
            ir.f64_idx = B.f64_idx();
            while (j < nk && coll.kids[j].type == k.type && (j - i) < 64 && (j / 64 == i / 64)) {
                child_f32[j] = B.f32_idx();
                child_f64[j] = B.f64_idx();
                child_op[j] = (uint32_t)ir.op;
                emit_prim_params(B, coll.kids[j]);
                ++j;
            }
            ir.n = (uint32_t)(j - i);
            B.instr.push_back(ir);
            i = j;
        } else {
            all_prims = false;
            if (!emit_node(B, k, false, (uint32_t)i)) return false;
            ++i;
        }
    }
    if (!(B.instr[begin].flags & F_HAS_GRID) && all_prims && nk > 63) {
        uint32_t lf = 0, li = 0;
        if (build_list_grid(B, coll, region, child_f32, child_f64, child_op, lf, li)) {
            B.instr[begin].flags |= F_HAS_LIST;
            B.instr[begin].f32_idx = lf;
            B.instr[begin].aux = li;
        }
    }
    Instr ie = {};
    ie.op = OP_COLL_END;
    ie.flags = ib.flags;
    B.instr.push_back(ie);
    B.instr[begin].skip_to = (uint32_t)(B.instr.size() - 1);
    if (!nosave) --B.save_depth;
    return true;
}

static bool emit_node(Builder& B, const Node& n, bool nosave, uint32_t child_bit) {
    if (is_prim(n.type)) {
        Instr ir = {};
        ir.op = prim_op(n.type);
        ir.n = 1;
        ir.child_bit = child_bit;
        ir.f32_idx = B.f32_idx();
        ir.f64_idx = B.f64_idx();
        emit_prim_params(B, n);
        B.instr.push_back(ir);
        return true;
    }
    if (n.type == N_VOXEL) {
        Instr ir = {};
        ir.op = OP_VOXEL;
        ir.n = 1;
        ir.child_bit = child_bit;
        ir.aux = (uint32_t)n.voxel_slot;
        ir.f32_idx = B.f32_idx();
        ir.f64_idx = B.f64_idx();
        B.f4(up32(2.0 * B.ep + 8.0 * kU32), 0, 0, 0);
        B.instr.push_back(ir);
        return true;
    }
    if (n.type == N_COLL) return emit_collection(B, n, nosave, child_bit, nullptr);
    // N_TESS
    const double* p = n.p;
    const double* uc = p + 6;
    size_t begin = B.instr.size();
    Instr ib = {};
    ib.op = OP_TESS_BEGIN;
    ib.n = 1;
    ib.flags = nosave ? F_NOSAVE : 0;
    ib.child_bit = child_bit;
    ib.f32_idx = B.f32_idx();
    ib.f64_idx = B.f64_idx();
    double oc[3], oh[3], d[3], xmax = 0, qmax = 0, dmin = kInf;
    for (int i = 0; i < 3; ++i) {
        oc[i] = 0.5 * (p[2 * i] + p[2 * i + 1]);
        oh[i] = 0.5 * (p[2 * i + 1] - p[2 * i]);
        d[i] = uc[2 * i + 1] - uc[2 * i];
        xmax = std::fmax(xmax, std::fmax(std::fabs(p[2 * i]), std::fabs(p[2 * i + 1])) + std::fabs(uc[2 * i]));
        if (d[i] != 0) qmax = std::fmax(qmax, (std::fabs(p[2 * i + 1] - uc[2 * i]) + std::fabs(p[2 * i] - uc[2 * i])) / std::fabs(d[i]) + 1);
        dmin = std::fmin(dmin, std::fabs(d[i]));
    }
    if (!(dmin > 0)) {
        B.err = "tessellated_obj_coll unit cell has zero extent";
        return false;
    }
    double tol_o = 2.0 * (B.ep + 4.0 * kU32 * xmax);
    double tolq = 2.0 * ((B.ep + 4.0 * kU32 * xmax) / dmin + 4.0 * kU32 * qmax);
    B.f4(oc[0], oc[1], oc[2], up32(tol_o));
    B.f4(oh[0], oh[1], oh[2], 0);
    B.f4(uc[0], uc[2], uc[4], 0);
    B.f4(d[0], d[1], d[2], 0);
    B.f4(1.0 / d[0], 1.0 / d[1], 1.0 / d[2], up32(tolq));
    B.d64(p, 12, kF64Tess);
    B.instr.push_back(ib);
    if (!nosave) {
        if (++B.save_depth > kMaxSaveDepth) {
            B.err = "scene nests too deeply for the device evaluator";
            return false;
        }
        B.save_depth_max = std::max(B.save_depth_max, B.save_depth);
    }
    Box3 region;
    double lo[3] = {uc[0], uc[2], uc[4]}, hi[3] = {uc[1], uc[3], uc[5]};
    box_add_point(region, lo);
    box_add_point(region, hi);
    if (!emit_collection(B, n.kids[0], nosave, 0, &region)) return false;
    Instr ie = {};
    ie.op = OP_TESS_END;
    ie.flags = ib.flags;
    B.instr.push_back(ie);
    B.instr[begin].skip_to = (uint32_t)(B.instr.size() - 1);
    if (!nosave) --B.save_depth;
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// Span section (program.h SpanHeader): the scene restated for the interval renderer (render_span.cu).
// Eligible: a bare convex primitive, one collection of <= 64 convex primitives, or a tessellation of such a
// collection; no warp, or a chain of affine warps (rigid / linear / affine), which keeps rays straight.
// ---------------------------------------------------------------------------------------------------------
static double g_span_feat_scale = 2.0;  // candidate-grid cell edge in units of the smallest child feature
static int g_span_grid_max = 16;
static double g_span_cells_per_cbrt = 2.5;

static bool span_convex(NodeType t) { return t == N_SPHERE || t == N_BOX || t == N_PPED || t == N_CYL; }

static bool build_span(const XRayScene& sc, std::vector<uint8_t>& out) {
    out.clear();
    const Node* coll = nullptr;
    const Node* tess = nullptr;
    const Node* bare = nullptr;
    const Node& root = sc.root;
    if (span_convex(root.type)) bare = &root;
    else if (root.type == N_COLL) coll = &root;
    else if (root.type == N_TESS && root.kids.size() == 1 && root.kids[0].type == N_COLL) {
        tess = &root;
        coll = &root.kids[0];
    } else return false;
    std::vector<const Node*> kids;
    if (bare) kids.push_back(bare);
    else
        for (const auto& k : coll->kids) {
            if (!span_convex(k.type)) return false;
            kids.push_back(&k);
        }
    if (kids.empty() || kids.size() > 64) return false;

    SpanHeader H = {};
    H.n_children = (uint32_t)kids.size();
    H.flags = (coll && coll->greedy ? SPAN_GREEDY : 0) | (coll ? SPAN_CLAMPS : 0) | (tess ? SPAN_TESS : 0);
    // the warp chain as one affine map
    double M[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, b[3] = {0, 0, 0};
    for (const auto& d : sc.deforms) {
        double A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
        const double* q = d.d;
        if (d.type == D_AFFINE) {
            for (int i = 0; i < 9; ++i) A[i] = q[i];
        } else if (d.type == D_LINEAR) {  // deformations.go:136-141, strains xx,yy,zz,yz,xz,xy
            const double E[9] = {q[0], q[5], q[4], q[5], q[1], q[3], q[4], q[3], q[2]};
            for (int i = 0; i < 9; ++i) A[i] += E[i];
        } else if (d.type == D_RIGID) {
            for (int i = 0; i < 3; ++i) t[i] = q[i];
        } else return false;
        double M2[9], b2[3];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) M2[r * 3 + c] = A[r * 3 + 0] * M[0 * 3 + c] + A[r * 3 + 1] * M[1 * 3 + c] + A[r * 3 + 2] * M[2 * 3 + c];
            b2[r] = A[r * 3 + 0] * b[0] + A[r * 3 + 1] * b[1] + A[r * 3 + 2] * b[2] + t[r];
        }
        memcpy(M, M2, sizeof(M));
        memcpy(b, b2, sizeof(b));
        H.flags |= SPAN_HAS_WARP;
    }
    for (int i = 0; i < 9; ++i) {
        if (!std::isfinite(M[i])) return false;
        H.warp_m[i] = M[i];
    }
    for (int i = 0; i < 3; ++i) {
        if (!std::isfinite(b[i])) return false;
        H.warp_b[i] = b[i];
    }
    {  // inverse of the warp, for projecting the children onto the detector (span_bin_kernel)
        const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
        H.f_wscale = 0.0f;
        if (std::fabs(det) > 1e-6) {
            const double inv[9] = {(M[4] * M[8] - M[5] * M[7]) / det, (M[2] * M[7] - M[1] * M[8]) / det, (M[1] * M[5] - M[2] * M[4]) / det,
                                   (M[5] * M[6] - M[3] * M[8]) / det, (M[0] * M[8] - M[2] * M[6]) / det, (M[2] * M[3] - M[0] * M[5]) / det,
                                   (M[3] * M[7] - M[4] * M[6]) / det, (M[1] * M[6] - M[0] * M[7]) / det, (M[0] * M[4] - M[1] * M[3]) / det};
            double fro = 0.0;
            for (int i = 0; i < 9; ++i) {
                H.f_winv[i] = (float)inv[i];
                fro += inv[i] * inv[i];
            }
            fro = std::sqrt(fro);
            if (std::isfinite(fro) && fro < 1e3) H.f_wscale = (float)(fro * 1.0001);
        }
    }

    // region the candidate grid spans
    double lo[3], hi[3];
    if (tess) {
        const double* p = tess->p;
        const double* uc = p + 6;
        for (int a = 0; a < 3; ++a) {
            H.outer[a] = p[2 * a];
            H.outer[3 + a] = p[2 * a + 1];
            lo[a] = uc[2 * a];
            hi[a] = uc[2 * a + 1];
            if (!(p[2 * a] <= p[2 * a + 1])) return false;
        }
    } else {
        Box3 reg = bare ? node_extent(*bare) : node_nonzero_region(*coll, false);
        if (reg.empty) return false;
        for (int a = 0; a < 3; ++a) {
            // a hair of slack: the region only has to CONTAIN every point of positive density
            const double m = 1e-9 * (1.0 + std::fmax(std::fabs(reg.lo[a]), std::fabs(reg.hi[a])));
            lo[a] = reg.lo[a] - m;
            hi[a] = reg.hi[a] + m;
            H.outer[a] = lo[a];
            H.outer[3 + a] = hi[a];
        }
    }
    double ext[3], feat = kInf, cmax = 0;
    for (int a = 0; a < 3; ++a) {
        if (!std::isfinite(lo[a]) || !std::isfinite(hi[a])) return false;
        ext[a] = hi[a] - lo[a];  // objects.go:570: dx := l.UC.Xmax - l.UC.Xmin
        if (!(ext[a] > 0)) return false;
        H.uc_lo[a] = lo[a];
        H.uc_d[a] = ext[a];
        H.uc_hi[a] = hi[a];
        cmax = std::fmax(cmax, std::fmax(std::fabs(H.outer[a]), std::fabs(H.outer[3 + a])) + std::fabs(lo[a]) + ext[a]);
    }
    H.n_periods = 1;
    if (tess)  // number of periods a ray can cross stays small enough for the 5-bit period codes of the kernel
        for (int a = 0; a < 3; ++a) {
            const double n_lo = std::floor((H.outer[a] - lo[a]) / ext[a]), n_hi = std::floor((H.outer[3 + a] - lo[a]) / ext[a]);
            if (!(n_lo >= -14.0 && n_hi <= 14.0)) return false;  // one period of slack: the walk starts a hair outside the box
            H.n_lo[a] = (int32_t)n_lo;
            H.n_hi[a] = (int32_t)n_hi;
            H.n_periods *= (uint32_t)(H.n_hi[a] - H.n_lo[a] + 1);
        }
    for (const Node* k : kids) switch (k->type) {
            case N_SPHERE: feat = std::fmin(feat, std::fabs(k->p[3])); break;
            case N_CYL: feat = std::fmin(feat, std::fabs(k->p[6])); break;
            case N_BOX: feat = std::fmin(feat, 0.5 * std::fmin(std::fabs(k->p[3]), std::fmin(std::fabs(k->p[4]), std::fabs(k->p[5])))); break;
            default: feat = std::fmin(feat, 0.5 * std::fmin(len3(k->p + 3), std::fmin(len3(k->p + 6), len3(k->p + 9)))); break;
        }
    if (!std::isfinite(feat) || !(feat > 0)) feat = std::fmax(ext[0], std::fmax(ext[1], ext[2]));
    // Cell edge: a walk step costs about two pre-filter tests, so the grid is only as fine as the child count pays for
    // (one child: a cell per period; the 36 struts of lattice.yaml: 8 cells per axis), and never finer than the features.
    int g[3];
    const double emax = std::fmax(ext[0], std::fmax(ext[1], ext[2]));
    const double cell = std::fmax(g_span_feat_scale * feat, emax / (g_span_cells_per_cbrt * std::cbrt((double)kids.size())));
    for (int a = 0; a < 3; ++a) g[a] = std::max(1, std::min(g_span_grid_max, (int)std::ceil(ext[a] / cell - 1e-9)));
    if (kids.size() == 1 && !tess) g[0] = g[1] = g[2] = 1;
    while ((size_t)g[0] * g[1] * g[2] > 2048) {
        int a = g[0] >= g[1] && g[0] >= g[2] ? 0 : (g[1] >= g[2] ? 1 : 2);
        g[a] = (g[a] + 1) / 2;
    }
    H.n_cells = (uint32_t)(g[0] * g[1] * g[2]);
    double cs[3];
    for (int a = 0; a < 3; ++a) {
        H.g[a] = (uint32_t)g[a];
        cs[a] = ext[a] / g[a];
        H.f_uc_lo[a] = (float)lo[a];
        H.f_inv_cell[a] = (float)(1.0 / cs[a]);
        H.f_cell[a] = (float)cs[a];
    }
    // A child is listed in every cell within `margin` of it: the kernel walks the grid in fp32 (position error a few
    // 1e-6 for |x| <= 8) and must never miss a cell the exact ray touches.
    const double margin = 2e-4 * (1.0 + cmax);
    std::vector<uint64_t> masks(H.n_cells, 0);
    for (int iz = 0; iz < g[2]; ++iz)
        for (int iy = 0; iy < g[1]; ++iy)
            for (int ix = 0; ix < g[0]; ++ix) {
                double clo[3] = {lo[0] + ix * cs[0] - margin, lo[1] + iy * cs[1] - margin, lo[2] + iz * cs[2] - margin};
                double chi[3] = {lo[0] + (ix + 1) * cs[0] + margin, lo[1] + (iy + 1) * cs[1] + margin, lo[2] + (iz + 1) * cs[2] + margin};
                uint64_t m = 0;
                for (size_t c = 0; c < kids.size(); ++c)
                    if (child_touches_cell(*kids[c], clo, chi)) m |= (uint64_t)1 << c;
                masks[((size_t)iz * g[1] + iy) * g[0] + ix] = m;
            }
    std::vector<SpanChild> recs(kids.size());
    const double fm = margin;  // pre-filter inflation (fp32 evaluation of a quadratic with O(1) coefficients)
    for (size_t c = 0; c < kids.size(); ++c) {
        const Node& k = *kids[c];
        SpanChild& r = recs[c];
        memset(&r, 0, sizeof(r));
        const double* p = k.p;
        switch (k.type) {
            case N_SPHERE:
                r.type = OP_SPHERE;
                r.rho = p[4];
                for (int i = 0; i < 3; ++i) r.p[i] = p[i], r.f[i] = (float)p[i];
                r.p[3] = p[3] * p[3];  // objects.go:67: dist2 < Radius*Radius
                r.f[3] = (float)((std::fabs(p[3]) + fm) * (std::fabs(p[3]) + fm));
                for (int i = 0; i < 3; ++i) r.bs[i] = (float)p[i];
                r.bs[3] = (float)(std::fabs(p[3]) + fm);
                break;
            case N_BOX:
                r.type = OP_BOX;
                r.rho = p[6];
                for (int i = 0; i < 3; ++i) r.p[i] = p[i], r.p[3 + i] = 0.5 * p[3 + i];  // objects.go:175: 0.5*Sides[i]
                {
                    const double R = 0.5 * std::sqrt(p[3] * p[3] + p[4] * p[4] + p[5] * p[5]) + fm;
                    for (int i = 0; i < 3; ++i) r.bs[i] = r.f[i] = (float)p[i];
                    r.bs[3] = (float)R;
                    r.f[3] = (float)(R * R);
                }
                break;
            case N_CYL: {
                r.type = OP_CYL;
                r.rho = p[7];
                double v[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]};
                const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
                if (!(vv > 0) || !std::isfinite(vv)) return false;  // degenerate cylinder: NaN semantics stay with the point evaluators
                for (int i = 0; i < 3; ++i) r.p[i] = p[i], r.p[3 + i] = v[i], r.f[i] = (float)p[i], r.f[3 + i] = (float)v[i];
                r.p[6] = 1.0 / vv;
                r.p[7] = p[6] * p[6];
                r.p[8] = vv;
                r.p[9] = p[6];
                r.f[6] = (float)(1.0 / vv);
                r.f[7] = (float)((std::fabs(p[6]) + fm) * (std::fabs(p[6]) + fm));
                if (p[6] < 0) return false;  // d < r never holds: leave such scenes to the point evaluators
                r.bs[3] = (float)(std::fabs(p[6]) + fm);  // capsule radius around the axis p0 .. p0 + v (f[0..5])
                break;
            }
            default: {
                r.type = OP_PPED;
                r.rho = p[12];
                const double* m = p + 13;  // column-major inverse: row i = (m[i], m[3+i], m[6+i])
                for (int i = 0; i < 3; ++i) {
                    r.p[i] = p[i];
                    r.p[3 + 3 * i + 0] = m[i];
                    r.p[3 + 3 * i + 1] = m[3 + i];
                    r.p[3 + 3 * i + 2] = m[6 + i];
                    r.p[12 + i] = std::fabs(m[i]) + std::fabs(m[3 + i]) + std::fabs(m[6 + i]);
                    if (!std::isfinite(r.p[12 + i])) return false;
                }
                {  // bounding sphere about the centre o + (v0 + v1 + v2) / 2
                    double c[3], R = 0;
                    for (int i = 0; i < 3; ++i) c[i] = p[i] + 0.5 * (p[3 + i] + p[6 + i] + p[9 + i]);
                    for (int m8 = 0; m8 < 8; ++m8) {
                        double q[3], d2 = 0;
                        for (int i = 0; i < 3; ++i) {
                            q[i] = p[i] + ((m8 & 1) ? p[3 + i] : 0) + ((m8 & 2) ? p[6 + i] : 0) + ((m8 & 4) ? p[9 + i] : 0);
                            d2 += (q[i] - c[i]) * (q[i] - c[i]);
                        }
                        R = std::fmax(R, std::sqrt(d2));
                    }
                    R += fm;
                    if (!std::isfinite(R)) return false;
                    for (int i = 0; i < 3; ++i) r.bs[i] = r.f[i] = (float)c[i];
                    r.bs[3] = (float)R;
                    r.f[3] = (float)(R * R);
                }
                break;
            }
        }
        if (!std::isfinite(r.rho)) return false;
    }
    size_t off = (sizeof(SpanHeader) + 15) / 16 * 16;
    H.child_off = (uint32_t)off;
    off += recs.size() * sizeof(SpanChild);
    off = (off + 15) / 16 * 16;
    H.mask_off = (uint32_t)off;
    off += masks.size() * sizeof(uint64_t);
    off = (off + 15) / 16 * 16;
    H.total_bytes = (uint32_t)off;
    out.assign(off, 0);
    memcpy(out.data(), &H, sizeof(H));
    memcpy(out.data() + H.child_off, recs.data(), recs.size() * sizeof(SpanChild));
    memcpy(out.data() + H.mask_off, masks.data(), masks.size() * sizeof(uint64_t));
    return true;
}

static std::atomic<uint64_t> g_scene_counter{1};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static bool build_blob(XRayScene& sc, std::string& err) {
#ifdef XRAY_DEV_KNOBS  // grid tuning knobs of development builds; the release library has fixed, tested values
    if (const char* e = getenv("XRAY_GRID_MAX")) g_grid_max = std::max(1, std::min(64, atoi(e)));
    if (const char* e = getenv("XRAY_GRID_MIN_CHILDREN")) g_grid_min_children = std::max(1, atoi(e));
    if (const char* e = getenv("XRAY_GRID_FEAT_SCALE")) g_grid_feat_scale = std::max(0.05, atof(e));
#endif
    Builder B;
    B.sc = &sc;
    double warp_lip = 1.0;
    B.ep = deform_eps_pos(sc.deforms, &warp_lip);
    B.warp_lip = warp_lip;
    B.so_ok = deform_second_order(sc.deforms, &B.so_jac2, &B.so_curv);
    if (!emit_node(B, sc.root, true, 0)) {
        err = B.err;
        return false;
    }
    Instr end = {};
    end.op = OP_END;
    B.instr.push_back(end);

    std::vector<DeformRec> drecs;
    for (const auto& d : sc.deforms) {
        DeformRec r = {};
        r.type = d.type;
        r.axis = (uint32_t)d.axis;
        for (int i = 0; i < 12; ++i) r.d[i] = d.d[i];
        for (int i = 0; i < 12; ++i) r.f[i] = (float)d.d[i];
        if (d.type == D_GAUSSIAN)
            for (int i = 0; i < 3; ++i) r.f[3 + i] = (float)(-1.0 / (2.0 * d.d[3 + i] * d.d[3 + i]));
        if (d.type == D_SIGMOID) r.f[2] = (float)(-1.0 / d.d[2]);
        drecs.push_back(r);
    }

    Header h = {};
    h.magic = kMagic;
    h.version = kVersion;
    size_t off = align_up(sizeof(Header), 16);
    h.n_instr = (uint32_t)B.instr.size();
    h.instr_off = (uint32_t)off;
    off = align_up(off + B.instr.size() * sizeof(Instr), 16);
    h.f32_off = (uint32_t)off;
    h.f32_count = (uint32_t)(B.f32.size() / 4);
    off = align_up(off + B.f32.size() * sizeof(float), 16);
    h.f64_off = (uint32_t)off;
    h.f64_count = (uint32_t)B.f64.size();
    off = align_up(off + B.f64.size() * sizeof(double), 16);
    h.grid_off = (uint32_t)off;
    h.grid_count = (uint32_t)B.grids.size();
    off = align_up(off + B.grids.size() * sizeof(uint64_t), 16);
    h.deform_off = (uint32_t)off;
    h.n_deform = (uint32_t)drecs.size();
    off = align_up(off + drecs.size() * sizeof(DeformRec), 16);
    std::vector<uint8_t> span;
#ifdef XRAY_DEV_KNOBS
    if (const char* e = getenv("XRAY_SPAN_FEAT_SCALE")) g_span_feat_scale = std::max(0.25, atof(e));
    if (const char* e = getenv("XRAY_SPAN_GRID_MAX")) g_span_grid_max = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("XRAY_SPAN_CELLS_PER_CBRT")) g_span_cells_per_cbrt = std::max(0.1, atof(e));
#endif
    if (build_span(sc, span)) {
        h.span_off = (uint32_t)off;
        h.span_bytes = (uint32_t)span.size();
        off = align_up(off + span.size(), 16);
    }
    h.total_bytes = (uint32_t)off;
    h.save_depth = (uint32_t)B.save_depth_max;
    h.n_voxel_slots = (uint32_t)sc.n_vox;
    h.min_feature_size = node_min_feature_size(sc.root, sc);
    h.eps_pos = B.ep;
    h.warp_lipschitz = warp_lip;
    for (int s = 0; s < sc.n_vox; ++s) {
        h.voxel_dims[s][0] = sc.vox[s].nx;
        h.voxel_dims[s][1] = sc.vox[s].ny;
        h.voxel_dims[s][2] = sc.vox[s].nz;
    }
    Box3 nz = pull_back(node_nonzero_region(sc.root, false), sc.deforms);
    if (nz.empty) {
        for (int i = 0; i < 3; ++i) {  // nothing can be non-zero: an inverted box clips every ray
            h.aabb_lo[i] = 1.0;
            h.aabb_hi[i] = -1.0;
        }
    } else {
        for (int i = 0; i < 3; ++i) {
            double m = 1e-5 + 1e-6 * std::fmax(std::fabs(nz.lo[i]), std::fabs(nz.hi[i]));
            h.aabb_lo[i] = std::isfinite(nz.lo[i]) ? nz.lo[i] - m : nz.lo[i];
            h.aabb_hi[i] = std::isfinite(nz.hi[i]) ? nz.hi[i] + m : nz.hi[i];
        }
    }
    sc.blob.assign(off, 0);
    memcpy(sc.blob.data(), &h, sizeof(h));
    memcpy(sc.blob.data() + h.instr_off, B.instr.data(), B.instr.size() * sizeof(Instr));
    if (!B.f32.empty()) memcpy(sc.blob.data() + h.f32_off, B.f32.data(), B.f32.size() * sizeof(float));
    if (!B.f64.empty()) memcpy(sc.blob.data() + h.f64_off, B.f64.data(), B.f64.size() * sizeof(double));
    if (!B.grids.empty()) memcpy(sc.blob.data() + h.grid_off, B.grids.data(), B.grids.size() * sizeof(uint64_t));
    if (!drecs.empty()) memcpy(sc.blob.data() + h.deform_off, drecs.data(), drecs.size() * sizeof(DeformRec));
    if (!span.empty()) memcpy(sc.blob.data() + h.span_off, span.data(), span.size());
    return true;
}

int compile_scene_json(const char* object_json, const char* deform_json, XRayScene** out, std::string& err) {
    if (!object_json || !out) {
        err = "null argument";
        return 1;
    }
    JValue jo;
    {
        JParser jp(object_json);
        if (!jp.parse(jo, err)) {
            err = "object JSON: " + err;
            return 2;
        }
    }
    auto sc = new XRayScene();
    ParseCtx c{"", sc};
    if (!parse_object(jo, sc->root, c, false)) {
        err = c.err;
        delete sc;
        return 3;
    }
    if (deform_json && *deform_json) {
        JValue jd;
        JParser jp(deform_json);
        if (!jp.parse(jd, err)) {
            err = "deformation JSON: " + err;
            delete sc;
            return 2;
        }
        if (!parse_deformation(jd, sc->deforms, c, 0)) {
            err = c.err;
            delete sc;
            return 4;
        }
    }
    if (!build_blob(*sc, err)) {
        delete sc;
        return 5;
    }
    sc->id = g_scene_counter.fetch_add(1);
    *out = sc;
    return 0;
}

}  // namespace xr
