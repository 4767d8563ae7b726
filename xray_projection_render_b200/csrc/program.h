// program.h -- layout of the flattened scene ("instruction buffer") shared by the host
// compiler (scene_compile.cpp) and the device interpreter (render_scene.cu).
//
// The reference evaluates density() by walking a tree of Go interface values
// (objects/objects.go Object.Density, deformations/deformations.go Deformation.Apply).
// Here the tree is flattened once per scene into a linear, warp-uniform instruction
// stream plus two parameter pools: fp32 (pre-digested constants and guard-band
// tolerances for the fast path) and fp64 (the raw reference parameters for the exact path).
#pragma once
#include <stdint.h>

namespace xr {

constexpr uint32_t kMagic = 0x58524159u;  // "XRAY"
constexpr uint32_t kVersion = 8;
constexpr int kMaxVoxelSlots = 4;
constexpr int kMaxSaveDepth = 6;  // nested save frames (collections/tessellations inside collections)
constexpr int kFrameWords = 8;    // real-typed words per save frame

enum Op : uint32_t {
    OP_END = 0,
    OP_SPHERE = 1,   // objects.go:63-72
    OP_BOX = 2,      // objects.go:171-179 (Cube -> Box, :115-121)
    OP_CYL = 3,      // objects.go:334-350
    OP_PPED = 4,     // objects.go:247-255
    OP_GYROID = 5,   // objects.go:1014-1032
    OP_VOXEL = 6,    // objects.go:789-855
    OP_COLL_BEGIN = 8,  // objects.go:422-438
    OP_COLL_END = 9,
    OP_TESS_BEGIN = 10,  // objects.go:568-582 + UnitCell bounds :458-464
    OP_TESS_END = 11,
};

enum InstrFlags : uint32_t {
    F_GREEDY = 1,    // ObjectCollection.GreedyDensEval
    F_NOSAVE = 2,    // frame needs no save/restore (it is the only thing its parent evaluates)
    F_HAS_GRID = 4,  // collection has a child-mask grid
    F_HAS_LIST = 8,  // collection (> 63 primitive children) has a cell-list grid: per cell an ascending child list
};

// 32 bytes.  Primitive ops are "runs": n consecutive children of the same type.
struct Instr {
    uint32_t op;
    uint32_t n;          // run length (primitives) / number of children (COLL_BEGIN)
    uint32_t flags;
    uint32_t child_bit;  // index of (first) child within the enclosing collection
    uint32_t f32_idx;    // float4 index into the fp32 pool
    uint32_t f64_idx;    // double index into the fp64 pool
    uint32_t skip_to;    // COLL_BEGIN/TESS_BEGIN: index of the matching END instruction
    uint32_t aux;        // OP_VOXEL: slot; COLL_BEGIN: uint64 index of the mask grid
};

// fp32 pool record sizes in float4 units, fp64 pool record sizes in doubles.
//   sphere   f32: {cx,cy,cz,rho} {r2,tol,-,-}                            f64: cx,cy,cz,r,rho,-
//   box      f32: {cx,cy,cz,rho} {hx,hy,hz,tol}                          f64: cx,cy,cz,sx,sy,sz,rho,-
//   cylinder f32: {p0,rho} {v,inv_vv} {r2,tolr,tolc,tolc/tolr}            f64: p0(3),p1(3),r,rho
//   pped     f32: {o,rho} {row0,tol} {row1,-} {row2,-}                   f64: o(3),minv colmajor(9),rho,-
//   gyroid   f32: {c,rho} {inv_scale,thickness,tol,m2d} {M2,g1eps,lip,-}     f64: c(3),scale,thickness,rho
//            (third record: second-order skip bound along the ray, M2 = 0 when the warp chain has no curvature bound)
//   voxel    f32: {tol,-,-,-}                                            f64: (none)
//   tess     f32: {oc,tol_o} {oh,-} {ucmin,-} {d,-} {inv_d,tolq}         f64: outer(6: xmin,xmax,..), uc(6)
//   grid     f32: {gmin,cs_min} {inv_cell,-} {gx,gy,gz (int bits), -} {gx,gy,gz as floats, -}
//   list grid (F_HAS_LIST) adds a 5th float4 of int bits: u64 offsets (relative to Instr.aux) of
//            cell_off[ncell+1] (u32), cell_dist[ncell] (u8), idx[] (u16), child_tab[n] (u32 = op<<24 | float4 index)
constexpr int kF32Sphere = 2, kF32Box = 2, kF32Cyl = 3, kF32Pped = 4, kF32Gyroid = 3, kF32Voxel = 1, kF32Tess = 5,
              kF32Grid = 4;
constexpr int kF64Sphere = 6, kF64Box = 8, kF64Cyl = 8, kF64Pped = 14, kF64Gyroid = 6, kF64Tess = 12;

enum DeformType : uint32_t { D_GAUSSIAN = 1, D_AFFINE = 2, D_LINEAR = 3, D_RIGID = 4, D_SIGMOID = 5 };

// One deformation stage (a "composed" deformation is flattened to a sequence).  96+64 bytes.
//   gaussian d: A(3), S(3), C(3)        f: A(3), -1/(2 S^2)(3), C(3)      deformations.go:29-38
//   affine   d: M row-major(9)          f: same                             deformations.go:87-92
//   linear   d: strains(6)              f: same                             deformations.go:136-141
//   rigid    d: D(3)                    f: same                             deformations.go:173-175
//   sigmoid  d: A, c, L ; axis          f: A, c, -1/L                       deformations.go:210-222
struct DeformRec {
    uint32_t type;
    uint32_t axis;
    uint32_t pad[2];
    double d[12];
    float f[12];
};

struct Header {
    uint32_t magic, version;
    uint32_t total_bytes;
    uint32_t n_instr, instr_off;
    uint32_t f32_off, f32_count;  // count in float4
    uint32_t f64_off, f64_count;  // count in doubles
    uint32_t grid_off, grid_count;  // count in uint64
    uint32_t deform_off, n_deform;
    uint32_t save_depth;  // max nesting of frames that need the save stack
    uint32_t n_voxel_slots;
    uint32_t flags;
    double min_feature_size;
    double aabb_lo[3], aabb_hi[3];  // world-space region outside of which density()==0
    double eps_pos;                 // position error bound assumed by the fp32 tolerances
    double warp_lipschitz;          // object-space displacement per unit world displacement (deformation chain)
    int32_t voxel_dims[kMaxVoxelSlots][4];
    uint32_t span_off, span_bytes;  // SpanHeader section for the interval ("span") renderer, 0 bytes when the scene has none
};

// ---------------------------------------------------------------------------------------------------------
// Span section: what render_span.cu needs, in fp64.  Built for scenes that are one collection of CONVEX
// primitives (sphere, box / cube, parallelepiped, cylinder; <= 64 children), bare or tessellated, under no warp or
// an affine one (rigid / linear / affine / compositions).  Along a straight ray every such child is ONE interval, so
// the reference's per-sample loop (main.go:144-199) collapses to a sweep over interval end points.
// ---------------------------------------------------------------------------------------------------------
enum SpanFlags : uint32_t { SPAN_GREEDY = 1, SPAN_CLAMPS = 2, SPAN_TESS = 4, SPAN_HAS_WARP = 8 };

struct SpanHeader {
    uint32_t n_children, flags;
    uint32_t g[3];                 // candidate grid: cells per period (TESS) / over the region (FLAT)
    uint32_t n_cells;
    uint32_t child_off, mask_off;  // byte offsets from the start of the section: SpanChild[n], uint64 mask[n_cells]
    uint32_t total_bytes, pad;
    int32_t n_lo[3], n_hi[3];      // periods that meet the outer box, per axis (FLAT: 0..0)
    uint32_t n_periods, pad2;      // their number, (n_hi - n_lo + 1) multiplied out
    double outer[6];               // lo xyz, hi xyz; inclusive bounds test of objects.go:569 (TESS) / the region (FLAT)
    double uc_lo[3], uc_d[3];      // unit-cell origin and period (objects.go:570-580); FLAT: region origin and extent
    double uc_hi[3];               // the unit cell's upper bounds as given (UnitCell.Density's inclusive test, objects.go:459)
    double warp_m[9], warp_b[3];   // object-space point = warp_m * world + warp_b (row-major), the composed affine warp
    float f_uc_lo[3], f_inv_cell[3], f_cell[3], f_pad[3];
    // for the screen-space bins of a warped scene: world = f_winv * (object - warp_b) (row-major), and an upper bound of how
    // much the inverse warp stretches lengths (Frobenius norm); f_wscale = 0: not invertible, no bins
    float f_winv[9], f_wscale, f_pad2[2];
};

// 184 bytes per child.  p[]: sphere c(3) r2 | box c(3) h(3) | cylinder p0(3) v(3) 1/(v.v) r2 v.v |
// parallelepiped o(3) Minv rows (9) rownorm(3).  f[]: fp32 copy for the conservative pre-filter
// (cylinder p0(3) v(3) 1/(v.v) (r+margin)^2; sphere / box / parallelepiped: bounding sphere c(3) (R+margin)^2).
// bs[]: bounding sphere c(3) R+margin (cylinder: R = r + margin around each end point, i.e. a capsule) for screen-space binning.
struct SpanChild {
    uint32_t type, pad;
    double rho;
    double p[15];
    float f[8];
    float bs[4];
};

}  // namespace xr
