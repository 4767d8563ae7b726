// device_types.h -- kernel parameter blocks shared by api.cu and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

namespace xr {

struct VoxelDev {
    const void* data;  // device, layout idx = z*NX*NY + x*NY + y
    int nx, ny, nz;
    int dtype;  // 0 f32, 1 f64
};

struct SceneDev {
    const Instr* instr;
    const float4* f32;
    const double* f64;
    const unsigned long long* grids;
    const DeformRec* deform;
    int n_instr, n_deform, f32_count, save_depth;
    const VoxelDev* vox;  // device table of kMaxVoxelSlots entries
};

// Per-view camera, fp64, prepared on the host.
struct CamDev {
    double eye[3];
    double view[16];  // row-major camera->world
    double f;         // 1/tan(fov/2), main.go:457
};

struct RenderParams {
    SceneDev scene;
    const CamDev* cams;
    int n_views, res;
    int tiles_i, tiles_j;
    // Sample lattice (ray independent, built on the host by fp64 repeated addition):
    //   simple:       s_tab[k] = k-th value of `for s := smin; s < smax; s += ds`, k in [0, n_steps)
    //   hierarchical: s_tab[0] = smin, s_tab[k+1] = k-th `right`, k in [0, n_steps)
    // t_tab[k] = float(s_tab[k] - s_center).
    const double* s_tab;
    const float* t_tab;
    int n_steps;
    double ds;         // step (coarse DS for hierarchical)
    double ds_fine;    // DS / 10.0
    double smin, smax, s_center;
    double flat_field, dm;
    double aabb_lo[3], aabb_hi[3];
    void* out;
    int out_f64;
    int out_vec4;               // the output pointer of this launch allows 128-bit stores (16-byte aligned image rows)
    int prog_in_smem;           // stage instr + fp32 pool in shared memory
    unsigned int smem_prog_bytes;
    unsigned long long* stats;  // device, XRAY_NUM_STATS counters or null
    float dm_f, ds_f, ds_fine_f;  // fp32 copies for the hot loop (no F2F per iteration)
    float skip_m2s;             // object-space clearance -> number of lattice steps that stay inside it (0 disables skipping)
    int dbg_cause;              // COUNT variants only: count fp64 fallbacks of this cause mask (0 = all)
    // Second pass after the interval renderer (render_span.cu): when tile_list is set, the grid is small and fixed and the CTAs
    // loop over the compacted warp-tile entries tile_list[0 .. *tile_count) (entry = (view, tile) id * 4 + warp position).
    const unsigned int* tile_list;
    const unsigned int* tile_count;
};

constexpr int kBlockThreads = 128;
constexpr unsigned int kTileListGrid = 148 * 8;  // CTAs of a tile-list launch (a few per SM)
#ifndef XR_WARP_I
#define XR_WARP_I 4
#endif
constexpr int kWarpI = XR_WARP_I, kWarpJ = 32 / XR_WARP_I;  // warp tile in pixels (i x j); j is the fast output index
constexpr int kTileI = 2 * kWarpI, kTileJ = 2 * kWarpJ;     // CTA tile: 2 x 2 warps
constexpr int kQueueCap = 8;            // deferred refinements per lane before a flush

}  // namespace xr
