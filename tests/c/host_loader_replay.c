/* host_loader_replay.c -- plays the part of the reference's Go host against libcuda_render.so from plain C:
 * the load sequence of cuda_backend.go:27-43 (dlopen RTLD_LAZY|RTLD_LOCAL, then the three legacy symbols, every
 * one of them required), the library search order of cuda_backend.go:103-114 (XRAY_CUDA_LIB first), and the
 * calls the host makes through them (cuda_test.go:15-28 probe; cuda_backend.go:294-337 render; :155-270 voxelise)
 * with caller-owned pageable buffers.  It includes only the public header, as a maintainer's cgo preamble would.
 *
 *   host_loader_replay <lib> symbols   -> load + resolve, argument validation (no GPU needed)
 *   host_loader_replay <lib> probe     -> the 1x1x1 probe render and a small voxelisation (needs a GPU)
 *   host_loader_replay <lib> scene     -> the call sequence integration/go_host.patch adds to the host (cuda_dl_RenderSceneCUDA):
 *                                         six optional symbols, compile object + deformation JSON, render, free (needs a GPU)
 * Exit code 0 = everything behaved as the Go host expects.  Test infrastructure, not product code. */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "xray_cuda_render.h"

typedef int (*assemble_fn)(const CylinderParams*, int, int, float, float*);
typedef int (*assemble_spatial_fn)(const CylinderParams*, int, int, float, int, const int*, const int*, int, float*);
typedef int (*render_fn)(const float*, int, int, int, const XRayCameraParams*, int, int, float, float, float*);

struct plugin {
    void* handle;
    assemble_fn assemble;
    assemble_spatial_fn assemble_spatial;
    render_fn render;
};

/* -1: dlopen failed, -2: a symbol is missing (the Go loader's codes) */
static int plugin_load(struct plugin* p, const char* path) {
    memset(p, 0, sizeof(*p));
    p->handle = dlopen(path, RTLD_LAZY | RTLD_LOCAL);
    if (!p->handle) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return -1;
    }
    p->assemble = (assemble_fn)dlsym(p->handle, "AssembleVoxelGridCUDA");
    p->assemble_spatial = (assemble_spatial_fn)dlsym(p->handle, "AssembleVoxelGridSpatialCUDA");
    p->render = (render_fn)dlsym(p->handle, "RenderVolumeProjectionsCUDA");
    if (!p->assemble || !p->assemble_spatial || !p->render) {
        fprintf(stderr, "dlsym: %s\n", dlerror());
        return -2;
    }
    return 0;
}

#define EXPECT(cond)                                                     \
    do {                                                                 \
        if (!(cond)) {                                                   \
            fprintf(stderr, "line %d: expected %s\n", __LINE__, #cond); \
            return 1;                                                    \
        }                                                                \
    } while (0)

static int check_symbols(struct plugin* p) {
    /* struct layouts the Go side marshals into (cuda_backend.go:116-153) */
    EXPECT(sizeof(CylinderParams) == 32);
    EXPECT(sizeof(XRayCameraParams) == 84);
    /* argument validation happens before any CUDA call, so these hold on a machine without a GPU:
     * non-zero return, no crash, no exit (cuda_backend.cu:95-101 convention: 1 null pointer, 2 bad dimension) */
    float vol = 0.0f, out = 0.0f;
    XRayCameraParams cam;
    memset(&cam, 0, sizeof(cam));
    EXPECT(p->render(NULL, 1, 1, 1, &cam, 1, 1, 0.1f, 0.0f, &out) != 0);
    EXPECT(p->render(&vol, 0, 1, 1, &cam, 1, 1, 0.1f, 0.0f, &out) != 0);
    EXPECT(p->render(&vol, 1, 1, 1, &cam, 0, 1, 0.1f, 0.0f, &out) != 0);
    EXPECT(p->render(&vol, 1, 1, 1, &cam, 1, 1, 0.0f, 0.0f, &out) != 0);
    EXPECT(p->assemble(NULL, 1, 4, 1.0f, &out) != 0);
    EXPECT(p->assemble_spatial(NULL, 1, 4, 1.0f, 16, NULL, NULL, 0, &out) != 0);
    /* the extended surface is optional for the unmodified host, but this build must carry it */
    EXPECT(dlsym(p->handle, "XRayRenderSceneCUDA") != NULL);
    EXPECT(dlsym(p->handle, "XRaySceneCompileJSON") != NULL);
    return 0;
}

static int check_probe(struct plugin* p) {
    /* cuda_test.go:15-28: a 1x1x1 zero volume, identity view, one pixel -> must succeed, transmission exp(-0) = 1 */
    float vol[1] = {0.0f}, out[1] = {-1.0f};
    XRayCameraParams cam;
    memset(&cam, 0, sizeof(cam));
    cam.eye[0] = 4.0f;
    for (int d = 0; d < 4; ++d) cam.view[d * 4 + d] = 1.0f;
    cam.fov_y = 40.0f;
    cam.R = 4.0f;
    EXPECT(p->render(vol, 1, 1, 1, &cam, 1, 1, 0.1f, 0.0f, out) == 0);
    EXPECT(out[0] == 1.0f);
    /* flat field only: exp(-0.5) whatever the volume holds outside the ray */
    EXPECT(p->render(vol, 1, 1, 1, &cam, 1, 1, 0.1f, 0.5f, out) == 0);
    EXPECT(fabsf(out[0] - expf(-0.5f)) < 1e-6f);

    /* cuda_voxel.go:42-50 -> AssembleVoxelGridSpatialCUDA with a host-built CSR grid (cuda_backend.go:182-237),
     * here the trivial one: every cell lists the one cylinder.  Voxel (i,j,k) sits at i/res*2-1 (cuda_backend.cu:222-226). */
    enum { RES = 8, G = 2 };
    CylinderParams cyl = {{0.0f, 0.0f, -0.9f}, {0.0f, 0.0f, 0.9f}, 0.3f, 0.5f};
    int offs[G * G * G + 1], idx[G * G * G];
    for (int c = 0; c <= G * G * G; ++c) offs[c] = c;
    for (int c = 0; c < G * G * G; ++c) idx[c] = 0;
    float a[RES * RES * RES], b[RES * RES * RES];
    EXPECT(p->assemble_spatial(&cyl, 1, RES, 2.0f, G, offs, idx, G * G * G, a) == 0);
    EXPECT(p->assemble(&cyl, 1, RES, 2.0f, b) == 0);
    for (int k = 0; k < RES; ++k)
        for (int i = 0; i < RES; ++i)
            for (int j = 0; j < RES; ++j) {
                const float x = (float)i / RES * 2 - 1, y = (float)j / RES * 2 - 1, z = (float)k / RES * 2 - 1;
                const float want = (x * x + y * y < 0.09f && z >= -0.9f && z <= 0.9f) ? 1.0f : 0.0f; /* 0.5 * 2, clamped to 1 */
                const size_t o = ((size_t)k * RES + i) * RES + j;
                EXPECT(a[o] == want);
                EXPECT(b[o] == want);
            }
    return 0;
}

/* integration/go_host.patch, cuda_backend.go cuda_dl_load + cuda_dl_RenderSceneCUDA: the JSON is what json.Marshal of
 * lat[0].ToMap() / df[0].ToMap() produces (objects.go / deformations.go ToMap methods). */
static int check_scene(struct plugin* p) {
    typedef int (*compile_fn)(const char*, const char*, XRayScene**);
    typedef void (*free_fn)(XRayScene*);
    typedef int (*slots_fn)(const XRayScene*);
    typedef void (*optsinit_fn)(XRayRenderOpts*);
    typedef int (*render_scene_fn)(XRayScene*, const XRayCameraParams64*, int, int, const XRayRenderOpts*, void*);
    typedef const char* (*lasterr_fn)(void);
    compile_fn compile = (compile_fn)dlsym(p->handle, "XRaySceneCompileJSON");
    free_fn sfree = (free_fn)dlsym(p->handle, "XRaySceneFree");
    slots_fn slots = (slots_fn)dlsym(p->handle, "XRaySceneNumVoxelSlots");
    optsinit_fn optsinit = (optsinit_fn)dlsym(p->handle, "XRayRenderOptsInit");
    render_scene_fn render_scene = (render_scene_fn)dlsym(p->handle, "XRayRenderSceneCUDA");
    lasterr_fn lasterr = (lasterr_fn)dlsym(p->handle, "XRayLastError");
    EXPECT(compile && sfree && slots && optsinit && render_scene && lasterr);
    const char* obj = "{\"type\":\"object_collection\",\"greedy_dens_eval\":false,\"objects\":["
                      "{\"type\":\"sphere\",\"center\":[0,0,0],\"radius\":0.5,\"rho\":1}]}";
    const char* def = "{\"type\":\"rigid\",\"displacements\":[0,0,0]}";
    XRayScene* sc = NULL;
    EXPECT(compile(obj, def, &sc) == 0 && sc != NULL);
    EXPECT(slots(sc) == 0);
    /* camera on the +x axis looking at the origin: columns of view = camera axes (s, u, -f), translation = eye */
    XRayCameraParams64 cam;
    memset(&cam, 0, sizeof(cam));
    cam.eye[0] = 4.0;
    cam.view[0 * 4 + 2] = 1.0; cam.view[0 * 4 + 3] = 4.0; /* world x = cam z + 4  (camera looks down -z = -x) */
    cam.view[1 * 4 + 0] = 1.0;                            /* world y = cam x */
    cam.view[2 * 4 + 1] = 1.0;                            /* world z = cam y */
    cam.view[15] = 1.0;
    cam.fov_y = 40.0;
    cam.R = 4.0;
    XRayRenderOpts o;
    optsinit(&o);
    EXPECT(o.struct_size == sizeof(XRayRenderOpts));
    o.integration = XRAY_INTEGRATE_SIMPLE;
    o.ds = 1.0e-3;
    o.flat_field = 0.0;
    o.density_multiplier = 1.0;
    enum { RES = 4 };
    float out[RES * RES];
    int rc = render_scene(sc, &cam, 1, RES, &o, out);
    if (rc != 0) fprintf(stderr, "XRayRenderSceneCUDA: %d %s\n", rc, lasterr());
    EXPECT(rc == 0);
    /* pixel (RES/2, RES/2) is the central ray: a chord of length 1 through the unit-density sphere (main_test.go:326-361) */
    EXPECT(fabsf(out[(RES / 2) * RES + RES / 2] - expf(-1.0f)) < 2.0e-3f);
    EXPECT(out[0] == 1.0f); /* the corner ray misses */
    sfree(sc);
    /* errors come back as codes plus text, never as a crash */
    sc = NULL;
    EXPECT(compile("{\"type\":\"nonsense\"}", NULL, &sc) != 0);
    EXPECT(lasterr()[0] != 0);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <libcuda_render.so | -> <symbols|probe>\n", argv[0]);
        return 64;
    }
    /* "-" = resolve like the Go host: $XRAY_CUDA_LIB, else the bare name on the loader path */
    const char* path = argv[1];
    if (strcmp(path, "-") == 0) {
        path = getenv("XRAY_CUDA_LIB");
        if (!path || !*path) path = "libcuda_render.so";
    }
    struct plugin p;
    int rc = plugin_load(&p, path);
    if (rc != 0) return -rc; /* 1 / 2 */
    rc = strcmp(argv[2], "probe") == 0 ? check_probe(&p) : (strcmp(argv[2], "scene") == 0 ? check_scene(&p) : check_symbols(&p));
    if (rc == 0) printf("ok %s\n", argv[2]);
    return rc ? 10 : 0;
}
