// misc_kernels.cu -- legacy cylinder voxeliser (AssembleVoxelGrid*CUDA) and the FP32 peak probe.
#include <cuda_runtime.h>

#include "../../include/xray_cuda_render.h"

namespace xr {

// Replaces reference voxelize_kernel / voxelize_spatial_kernel (cuda_backend.cu:208-252,297-357).
// One thread per voxel; out[k*res*res + i*res + j], world x = i/res*2-1, y <- j, z <- k; fp32
// segment-distance test per cylinder, sum of rho, * density_multiplier, clamp to [0,1] -- the
// reference kernel's arithmetic, with size_t indexing (the reference's `int total = res*res*res`
// overflows for res >= 1291) and j as the fastest thread index so stores coalesce.
// With a CSR (grid_dim > 0) only the caller's candidate list for the voxel's cell is visited,
// exactly as the reference does, so both symbols agree with their reference counterparts.
__device__ __forceinline__ float cyl_contrib(const CylinderParams& c, float x, float y, float z) {
    const float vx = c.p1[0] - c.p0[0], vy = c.p1[1] - c.p0[1], vz = c.p1[2] - c.p0[2];
    const float wx = x - c.p0[0], wy = y - c.p0[1], wz = z - c.p0[2];
    const float vdotv = vx * vx + vy * vy + vz * vz;
    if (vdotv == 0.0f) return 0.0f;
    const float t = (wx * vx + wy * vy + wz * vz) / vdotv;
    if (t < 0.0f || t > 1.0f) return 0.0f;
    const float dx = wx - vx * t, dy = wy - vy * t, dz = wz - vz * t;
    return (dx * dx + dy * dy + dz * dz < c.radius * c.radius) ? c.rho : 0.0f;
}

__global__ void __launch_bounds__(256) voxelize_cyl_kernel(const CylinderParams* __restrict__ cyl, int n, int res, float dm,
                                                           const int* __restrict__ off, const int* __restrict__ idx,
                                                           int grid_dim, float* __restrict__ out) {
    extern __shared__ CylinderParams s_cyl[];
    const size_t total = (size_t)res * res * res;
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = id < total;
    const size_t vid = valid ? id : 0;
    const int k = (int)(vid / ((size_t)res * res));
    const int i = (int)((vid / res) % res);
    const int j = (int)(vid % res);
    const float res_f = (float)res;
    const float x = (float)i / res_f * 2.0f - 1.0f;
    const float y = (float)j / res_f * 2.0f - 1.0f;
    const float z = (float)k / res_f * 2.0f - 1.0f;
    float density = 0.0f;
    if (grid_dim > 0) {
        const float cell_size = 2.0f / (float)grid_dim;
        int cx = (int)((x + 1.0f) / cell_size), cy = (int)((y + 1.0f) / cell_size), cz = (int)((z + 1.0f) / cell_size);
        cx = max(0, min(grid_dim - 1, cx));
        cy = max(0, min(grid_dim - 1, cy));
        cz = max(0, min(grid_dim - 1, cz));
        const int cell = (cz * grid_dim + cy) * grid_dim + cx;
        const int b = off[cell], e = off[cell + 1];
        for (int q = b; q < e; ++q) density += cyl_contrib(cyl[idx[q]], x, y, z);
    } else {
        // brute force over all cylinders, staged through shared memory in tiles
        const int tile = 512;
        for (int base = 0; base < n; base += tile) {
            const int m = min(tile, n - base);
            __syncthreads();
            for (int q = threadIdx.x; q < m; q += blockDim.x) s_cyl[q] = cyl[base + q];
            __syncthreads();
            for (int q = 0; q < m; ++q) density += cyl_contrib(s_cyl[q], x, y, z);
        }
    }
    density *= dm;
    if (density > 1.0f) density = 1.0f;
    if (density < 0.0f) density = 0.0f;
    if (valid) out[id] = density;
}

cudaError_t launch_voxelize_cylinders(const CylinderParams* d_cyl, int n, int res, float dm, const int* d_off,
                                      const int* d_idx, int grid_dim, float* d_out, cudaStream_t stream) {
    const size_t total = (size_t)res * res * res;
    const size_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    const size_t smem = grid_dim > 0 ? 0 : 512 * sizeof(CylinderParams);
    voxelize_cyl_kernel<<<(unsigned int)blocks, 256, smem, stream>>>(d_cyl, n, res, dm, d_off, d_idx, grid_dim, d_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

// FP32 peak: 8 independent FFMA chains per thread, enough warps to fill every SM.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
    float r0 = threadIdx.x * 1e-3f, r1 = r0 + 1.f, r2 = r0 + 2.f, r3 = r0 + 3.f, r4 = r0 + 4.f, r5 = r0 + 5.f, r6 = r0 + 6.f,
          r7 = r0 + 7.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            r0 = fmaf(r0, a, b); r1 = fmaf(r1, a, b); r2 = fmaf(r2, a, b); r3 = fmaf(r3, a, b);
            r4 = fmaf(r4, a, b); r5 = fmaf(r5, a, b); r6 = fmaf(r6, a, b); r7 = fmaf(r7, a, b);
        }
    }
    const float s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
    if (s == 123.456f) out[0] = s;  // keep the chains alive
}

cudaError_t measure_fp32_peak(double* tflops) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    float* d = nullptr;
    e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(t0);
        ffma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * threads;
        if (rep > 0 && ms > 0) best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    *tflops = best;
    return cudaGetLastError();
}

}  // namespace xr
