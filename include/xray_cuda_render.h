/* xray_cuda_render.h -- C ABI of libcuda_render.so (B200 / sm_100a build).
 *
 * Part 1 is byte-for-byte the plugin surface the unmodified reference host binds with
 * dlopen/dlsym (reference cuda_backend.h:22-112, loader cuda_backend.go:27-68): the same
 * two structs and the same three symbols, all of which must resolve or the Go loader
 * reports -2.  Part 2 extends it, with the same conventions (plain pointers and sizes,
 * 0 = success, non-zero = error, synchronous unless stated, caller owns every buffer),
 * to what the legacy surface cannot express: analytic scenes, the hierarchical
 * integrator, density multiplier, deformations, an fp64 mode, device selection and
 * device-resident buffers.
 *
 * No CUDA or torch types appear in any signature.
 */
#ifndef XRAY_CUDA_RENDER_H
#define XRAY_CUDA_RENDER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ======================================================================================
 * Part 1 -- legacy plugin surface (drop-in for reference cuda_backend.h)
 * ==================================================================================== */

/* reference cuda_backend.h:22-27.  32 bytes, align 4. */
typedef struct {
    float p0[3];
    float p1[3];
    float radius;
    float rho;
} CylinderParams;

/* reference cuda_backend.h:78-83.  84 bytes, align 4.
 * view = row-major camera->world matrix (view[r*4+c] = camMat.At(r,c), cuda_path.go:66-71),
 * fov_y in degrees, R = camera distance. */
typedef struct {
    float eye[3];
    float view[16];
    float fov_y;
    float R;
} XRayCameraParams;

/* Replaces reference cuda_backend.cu:254-295 (declared cuda_backend.h:33-39).
 * out_volume[k*res*res + i*res + j], world x = i/res*2-1 (y <- j, z <- k). */
int AssembleVoxelGridCUDA(const CylinderParams* cylinders, int num_cylinders, int res, float density_multiplier,
                          float* out_volume);

/* Replaces reference cuda_backend.cu:359-417 (declared cuda_backend.h:58-68).  The CSR arrays
 * built by cuda_backend.go:182-237 are accepted and validated; the voxeliser builds its own
 * exact culling structure, so results equal the brute-force symbol bit for bit. */
int AssembleVoxelGridSpatialCUDA(const CylinderParams* cylinders, int num_cylinders, int res,
                                 float density_multiplier, int grid_dim, const int* cell_offsets,
                                 const int* cyl_indices, int num_cyl_indices, float* out_volume);

/* Replaces reference cuda_backend.cu:82-203 (declared cuda_backend.h:101-112).
 * volume[k*nx*ny + i*ny + j]; out_images[cam*res*res + i*res + j].
 * Semantics: the Go CPU path (integrate_along_ray over VoxelGrid.Density: corner aligned,
 * zero outside [-1,1]^3, fp64 sample lattice), NOT the reference kernel's half-voxel
 * texture convention.  All host pointers, synchronous. */
int RenderVolumeProjectionsCUDA(const float* volume, int nx, int ny, int nz, const XRayCameraParams* cameras,
                                int num_cameras, int image_res, float ds, float flat_field, float* out_images);

/* ======================================================================================
 * Part 2 -- extended surface
 * ==================================================================================== */

/* fp64 camera: what computeCameraFromAngles (main.go:226-239) returns, un-narrowed. */
typedef struct {
    double eye[3];
    double view[16]; /* row-major camera->world, view[r*4+c] = camera.At(r,c) */
    double fov_y;    /* degrees */
    double R;
} XRayCameraParams64;

enum { XRAY_INTEGRATE_SIMPLE = 0, XRAY_INTEGRATE_HIERARCHICAL = 1 }; /* main.go:144-154 / 159-199 */
enum { XRAY_PRECISION_FP32 = 0, XRAY_PRECISION_FP64 = 1 };
enum { XRAY_OUT_F32 = 0, XRAY_OUT_F64 = 1 };
enum { XRAY_VOXEL_F32 = 0, XRAY_VOXEL_F64 = 1 };

#define XRAY_MAX_DEVICES 16
#define XRAY_NUM_STATS 8

typedef struct {
    uint32_t struct_size;      /* = sizeof(XRayRenderOpts); guards against ABI drift */
    int32_t integration;       /* XRAY_INTEGRATE_*  (reference global `integrate`, main.go:39) */
    int32_t precision;         /* XRAY_PRECISION_*: fp32 = guard-banded fp32 (|dI| <= 1e-4),
                                  fp64 = reference operation order (|dI| <= 1e-9).  An fp32 request is promoted to
                                  fp64 when the cameras put o + d*R so far from the origin (distant camera with a
                                  wide field of view) that fp32 positions would exceed the guard bands' budget. */
    int32_t out_dtype;         /* XRAY_OUT_* element type of the image buffer */
    double ds;                 /* step; <= 0 selects MinFeatureSize()/5 (main.go:350-353) */
    double flat_field;         /* reference global flat_field (main.go:40) */
    double density_multiplier; /* reference global density_multiplier (main.go:38) */
    int32_t num_devices;       /* 0 = the calling thread's current device only */
    int32_t devices[XRAY_MAX_DEVICES]; /* views are sharded v -> devices[v % num_devices]
                                          (= --jobs_modulo/--job, main.go:244) */
    uint64_t stream;           /* device-resident entry points: cudaStream_t to enqueue on (0 = default) */
    uint64_t* stats;           /* optional host array of XRAY_NUM_STATS counters, accumulated:
                                  [0] reference-equivalent samples (density() calls the Go integrator would make)
                                  [1] samples evaluated on the GPU   [2] fp64 re-evaluations in fp32 mode
                                  [3] primitive tests executed       [4] rays
                                  [5] kernel launches
                                  [6] image tiles the interval renderer handed to the marching kernels
                                  [7] why (OR of reason bits, render_span.cu); bit 16 = the interval renderer ran;
                                      bits 17 / 18 = density() > 0 at smin / smax for some ray (hierarchical integrator: the
                                      reference's "Clipping at smin/smax detected" warnings, main.go:162-169) */
    int32_t view_begin;        /* first camera's global view index (for sharding bookkeeping only) */
    int32_t reserved[7];
} XRayRenderOpts;

typedef struct XRayScene XRayScene; /* opaque: compiled scene (flattened instruction buffer) */

/* Thread-local text of the last error returned by any entry point on this thread. */
const char* XRayLastError(void);
int XRayDeviceCount(void);
/* Static text identifying this build: target architecture, nvcc version, compile date and time (so that a run's record
 * shows which binary it loaded and whether it was compiled on that machine). */
const char* XRayBuildInfo(void);
/* Free the per-device scratch (streams, staging buffers, lattice tables) the library keeps between calls. */
void XRayReleaseCaches(void);
void XRayRenderOptsInit(XRayRenderOpts* opts);

/* Compile a scene.  object_json is the object file content as JSON (the schema of
 * objects.go FromMap, i.e. what `lat[0].ToMap()` marshals to); deformation_json is the
 * deformation file content as JSON, or NULL/"" (deformations.go NewDeformation).  Host only:
 * works without a GPU.  voxel_grid nodes name a slot filled with XRaySceneSetVoxelData. */
int XRaySceneCompileJSON(const char* object_json, const char* deformation_json, XRayScene** out_scene);
void XRaySceneFree(XRayScene* scene);
/* lat[0].MinFeatureSize() (objects.go); auto ds = this / 5. */
double XRaySceneMinFeatureSize(const XRayScene* scene);
/* The flattened program (for inspection / tests). */
const void* XRaySceneProgram(const XRayScene* scene, size_t* num_bytes);
/* Conservative world-space box outside of which density() == 0: lo[3], hi[3] (may be +-inf). */
void XRaySceneBounds(const XRayScene* scene, double* lo, double* hi);
/* Number of voxel_grid nodes in the scene, and their (nx, ny, nz). */
int XRaySceneNumVoxelSlots(const XRayScene* scene);
int XRaySceneVoxelDims(const XRayScene* scene, int slot, int* nx, int* ny, int* nz);
/* Attach host data for one voxel slot: layout idx = z*NX*NY + x*NY + y (objects.go:836-843),
 * dtype XRAY_VOXEL_*.  Copied to the device(s) at the next render. */
int XRaySceneSetVoxelData(XRayScene* scene, int slot, const void* data, int nx, int ny, int nz, int dtype);
/* Host-side density() of main.go:137-140 in fp64 reference order (for tests / volume export). */
double XRaySceneDensityHost(const XRayScene* scene, double x, double y, double z, double density_multiplier);

/* main.go:226-239 computeCameraFromAngles, in fp64 with the reference operation order. */
int XRayCameraFromAngles(double azimuthal_deg, double polar_deg, double R, double fov_deg, XRayCameraParams64* out);

/* Render num_cameras views of a compiled scene into a HOST buffer
 * out_images[cam*res*res + i*res + j] of opts->out_dtype.  Synchronous. */
int XRayRenderSceneCUDA(XRayScene* scene, const XRayCameraParams64* cameras, int num_cameras, int image_res,
                        const XRayRenderOpts* opts, void* out_images);
/* Same, into a DEVICE buffer on the current device.  The kernels are enqueued on opts->stream and the call returns
 * without waiting for THEM; the caller orders later work after them on that stream (or synchronises it).  The call may
 * block the host on earlier work: the library keeps one set of per-device scratch (cameras, sample-lattice tables),
 * so it first waits for the previous call on this device -- whatever stream that one used -- to finish reading it, and a
 * scene's first use on a device uploads it synchronously.  cameras is a host array (copied before return).  With
 * opts->stats set the call synchronises the stream to read the counters back. */
int XRayRenderSceneDeviceCUDA(XRayScene* scene, const XRayCameraParams64* cameras, int num_cameras, int image_res,
                              const XRayRenderOpts* opts, void* d_out_images);

/* Voxel volume with the full option set (host volume / host images, synchronous).
 * volume_dtype XRAY_VOXEL_*.  With num_devices > 1 the volume is uploaded once and
 * broadcast to the other devices over NVLink. */
int XRayRenderVolumeExCUDA(const void* volume, int volume_dtype, int nx, int ny, int nz,
                           const XRayCameraParams64* cameras, int num_cameras, int image_res,
                           const XRayRenderOpts* opts, void* out_images);
/* Device-resident volume (fp32, reference layout) and device images, enqueued on opts->stream.  The volume is wrapped in
 * a temporary scene that is released before the call returns, which waits for the enqueued kernels: this entry point is
 * synchronous with respect to its own work (the images are complete on return). */
int XRayRenderVolumeDeviceCUDA(const float* d_volume, int nx, int ny, int nz, const XRayCameraParams64* cameras,
                               int num_cameras, int image_res, const XRayRenderOpts* opts, void* d_out_images);

/* Device-resident volume, HOST images (synchronous): what a process uses whose copy of the volume arrived over NVLink
 * (one rank uploads, ncclBroadcast replicates) but whose images go back to the host -- kernels, the D2H of each batch
 * and the copy into the caller's (pageable or pinned) buffer are overlapped like in XRayRenderVolumeExCUDA. */
int XRayRenderVolumeDeviceToHostCUDA(const float* d_volume, int nx, int ny, int nz, const XRayCameraParams64* cameras,
                                     int num_cameras, int image_res, const XRayRenderOpts* opts, void* out_images);

/* Voxelise a compiled scene (density() semantics of main.go computeVoxel:208-214):
 * out_volume[k*res*res + i*res + j] = density(i/res*2-1, j/res*2-1, k/res*2-1), fp32, host buffer. */
int XRayVoxelizeSceneCUDA(XRayScene* scene, int res, double density_multiplier, float* out_volume);

/* Measured peak of the FP32 FMA pipe on the current device in TFLOP/s (FFMA chain
 * microbenchmark, FMA = 2 flop); used as the roofline denominator for analytic scenes. */
int XRayMeasureFp32Peak(double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* XRAY_CUDA_RENDER_H */
