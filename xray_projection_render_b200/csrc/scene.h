// scene.h -- host-side compiled scene shared by scene_compile.cpp and api.cu.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "program.h"

namespace xr {

enum NodeType { N_SPHERE, N_BOX, N_PPED, N_CYL, N_GYROID, N_COLL, N_TESS, N_VOXEL };

// Host copy of the object tree (reference objects.go types), raw fp64 parameters.
struct Node {
    NodeType type;
    // sphere: c(3), r, rho | box: c(3), sides(3), rho | pped: o(3), v0, v1, v2 (9), rho, minv colmajor (9)
    // cyl: p0(3), p1(3), r, rho | gyroid: c(3), scale, thickness, rho
    // tess: outer xmin,xmax,ymin,ymax,zmin,zmax (6), uc (6); kids[0] = unit-cell collection
    double p[24] = {0};
    bool greedy = false;
    std::vector<Node> kids;
    int voxel_slot = -1;
};

struct Deform {
    DeformType type;
    int axis = 0;
    double d[12] = {0};
};

struct VoxelHost {
    int nx = 0, ny = 0, nz = 0;
    int dtype = 0;              // XRAY_VOXEL_*
    const void* data = nullptr;  // caller-owned, valid until the render that consumes it returns
    uint64_t version = 0;
};

struct Box3 {
    double lo[3], hi[3];
    bool empty = true;
};

}  // namespace xr

// The opaque handle of the C ABI.
struct XRayScene {
    xr::Node root;
    std::vector<xr::Deform> deforms;
    std::vector<uint8_t> blob;  // xr::Header + pools
    xr::VoxelHost vox[xr::kMaxVoxelSlots];
    int n_vox = 0;
    uint64_t id = 0;  // unique per compile, keys the device-side cache
    void* device_cache = nullptr;
};

namespace xr {
// scene_compile.cpp
int compile_scene_json(const char* object_json, const char* deform_json, XRayScene** out, std::string& err);
double node_min_feature_size(const Node& n, const XRayScene& sc);
double host_density(const XRayScene& sc, double x, double y, double z, double dm);
void camera_from_angles(double az_deg, double polar_deg, double R, double* eye, double* view_rowmajor);
}  // namespace xr
