// Probe: does a layered 2D cudaArray accept texture gather (tld4.a2d) on this GPU, what is the
// component order, and how fast is it compared with 8 scattered __ldg taps?
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float4 gather_a2d(cudaTextureObject_t tex, int layer, float x, float y) {
    float4 r;
    asm volatile("tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%7}];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(tex), "r"(layer), "f"(x), "f"(y));
    return r;
}
__global__ void probe(cudaTextureObject_t tex, float4* out, int layer, float x, float y) { out[0] = gather_a2d(tex, layer, x, y); }

__global__ void bench_tex(cudaTextureObject_t tex, int nx, int ny, int nz, float* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    // rays roughly along texture-y (volume x) with small lane offsets: scattered 2x2x2 footprints
    float ux = 3.0f + (t & 7) * 0.73f + (blockIdx.x % 97) * 5.1f, uy = 5.0f + ((t >> 3) & 3) * 0.73f + (blockIdx.x % 89) * 7.3f, uz = 7.0f + (blockIdx.x % 61) * 9.7f;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        ux += 0.11f; uy += 0.13f; uz += 0.09f;
        if (ux > nx - 2) ux -= nx - 4; if (uy > ny - 2) uy -= ny - 4; if (uz > nz - 2) uz -= nz - 4;
        float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
        int z0 = (int)fz;
        float4 a = gather_a2d(tex, z0, fy + 1.0f, fx + 1.0f), b = gather_a2d(tex, z0 + 1, fy + 1.0f, fx + 1.0f);
        float wx = ux - fx, wy = uy - fy, wz = uz - fz;
        float v00 = a.w + wz * (b.w - a.w), v01 = a.z + wz * (b.z - a.z), v10 = a.x + wz * (b.x - a.x), v11 = a.y + wz * (b.y - a.y);
        float v0 = v00 + wy * (v01 - v00), v1 = v10 + wy * (v11 - v10);
        acc += v0 + wx * (v1 - v0);
    }
    out[t] = acc;
}
__global__ void bench_ldg(const float* __restrict__ vol, int nx, int ny, int nz, float* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    float ux = 3.0f + (t & 7) * 0.73f + (blockIdx.x % 97) * 5.1f, uy = 5.0f + ((t >> 3) & 3) * 0.73f + (blockIdx.x % 89) * 7.3f, uz = 7.0f + (blockIdx.x % 61) * 9.7f;
    float acc = 0.f;
    size_t sz = (size_t)nx * ny;
    for (int it = 0; it < iters; ++it) {
        ux += 0.11f; uy += 0.13f; uz += 0.09f;
        if (ux > nx - 2) ux -= nx - 4; if (uy > ny - 2) uy -= ny - 4; if (uz > nz - 2) uz -= nz - 4;
        float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
        int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
        const float* p = vol + z0 * sz + (size_t)x0 * ny + y0;
        float a000 = __ldg(p), a010 = __ldg(p + 1), a100 = __ldg(p + ny), a110 = __ldg(p + ny + 1);
        float a001 = __ldg(p + sz), a011 = __ldg(p + sz + 1), a101 = __ldg(p + sz + ny), a111 = __ldg(p + sz + ny + 1);
        float wx = ux - fx, wy = uy - fy, wz = uz - fz;
        float v00 = a000 + wz * (a001 - a000), v01 = a010 + wz * (a011 - a010), v10 = a100 + wz * (a101 - a100), v11 = a110 + wz * (a111 - a110);
        float v0 = v00 + wy * (v01 - v00), v1 = v10 + wy * (v11 - v10);
        acc += v0 + wx * (v1 - v0);
    }
    out[t] = acc;
}
int main() {
    const int nx = 256, ny = 256, nz = 256;  // volume idx = z*nx*ny + x*ny + y  -> texture (width=ny, height=nx, layers=nz)
    std::vector<float> h((size_t)nx * ny * nz);
    for (int z = 0; z < nz; ++z) for (int x = 0; x < nx; ++x) for (int y = 0; y < ny; ++y) h[((size_t)z * nx + x) * ny + y] = z * 1e6f + x * 1e3f + y;
    float* d_vol; CK(cudaMalloc(&d_vol, h.size() * 4)); CK(cudaMemcpy(d_vol, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaArray_t arr;
    cudaError_t e = cudaMalloc3DArray(&arr, &cd, make_cudaExtent(ny, nx, nz), cudaArrayLayered | cudaArrayTextureGather);
    printf("layered|gather alloc: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) { cudaGetLastError(); CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(ny, nx, nz), cudaArrayLayered)); printf("plain layered alloc ok\n"); }
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr(d_vol, ny * sizeof(float), ny, nx); cp.dstArray = arr; cp.extent = make_cudaExtent(ny, nx, nz); cp.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3D(&cp));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    float4* d_o; CK(cudaMalloc(&d_o, 16)); float4 r;
    probe<<<1, 1>>>(tex, d_o, 5, 10.0f + 1.0f, 20.0f + 1.0f); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&r, d_o, 16, cudaMemcpyDeviceToHost));
    printf("gather layer5 texX(y)=10..11 texY(x)=20..21: x=%.0f y=%.0f z=%.0f w=%.0f  (value = z*1e6 + x*1e3 + y)\n", r.x, r.y, r.z, r.w);
    probe<<<1, 1>>>(tex, d_o, 255, 255.0f + 1.0f, 255.0f + 1.0f); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&r, d_o, 16, cudaMemcpyDeviceToHost));
    printf("gather at the far corner (clamp): x=%.0f y=%.0f z=%.0f w=%.0f\n", r.x, r.y, r.z, r.w);
    int sms = 148, blocks = sms * 16, threads = 128, iters = 2000; float* d_out; CK(cudaMalloc(&d_out, blocks * threads * 4));
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1); float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(t0); bench_tex<<<blocks, threads>>>(tex, nx, ny, nz, d_out, iters); cudaEventRecord(t1); CK(cudaEventSynchronize(t1)); cudaEventElapsedTime(&ms, t0, t1);
        printf("tex gather : %.3f ms -> %.1f Gsamples/s\n", ms, (double)blocks * threads * iters / ms / 1e6);
        cudaEventRecord(t0); bench_ldg<<<blocks, threads>>>(d_vol, nx, ny, nz, d_out, iters); cudaEventRecord(t1); CK(cudaEventSynchronize(t1)); cudaEventElapsedTime(&ms, t0, t1);
        printf("8x __ldg   : %.3f ms -> %.1f Gsamples/s\n", ms, (double)blocks * threads * iters / ms / 1e6);
    }
    return 0;
}
