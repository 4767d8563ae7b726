// tma_brick_probe.cu -- A/B evidence for the voxel path: what would a TMA-staged shared-memory brick sampler deliver?
//
// The production voxel kernel (csrc/render_volume.cu) gathers the 2x2x2 corners of each cell with two tld4 texture
// instructions and sums the samples of a cell in closed form.  The alternative the survey proposed stages axis-aligned
// bricks of the volume in shared memory with TMA (cp.async.bulk.tensor.3d, out-of-bounds zero fill, mbarrier double
// buffer) and lets the rays of an image tile take their trilinear taps from there.  This probe measures the two things
// that bound such a renderer from above, with the renderer's own access pattern and nothing else in the way (no ray
// set-up, no brick selection, no empty-space logic, no image write):
//   (1) the rate at which one CTA per SM slot can stage bricks (TMA bytes/s, the volume streamed through L2), and
//   (2) the rate at which the 256 rays of a 16 x 16 pixel tile take 8-tap trilinear samples from the staged brick
//       (fp32 weights, 0.2 voxel steps along an oblique direction, 0.73 voxel between neighbouring rays -- the cfg4 geometry).
// Build and run on the GPU box:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
//                                     -o /tmp/tma_probe tools/experiments/tma_brick_probe.cu && /tmp/tma_probe
// Development aid, not part of the library.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// BY (fastest, y), BX, BZ: brick extents in voxels.  One CTA = one 16 x 16 pixel tile of rays marching along +x.
template <int BY, int BX, int BZ, bool SAMPLE>
__global__ void __launch_bounds__(256) probe_kernel(const __grid_constant__ CUtensorMap map, int nx, int ny, int nz, int bricks_per_cta,
                                                    float* __restrict__ out, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kBrick = BY * BX * BZ;
    float* buf0 = reinterpret_cast<float*>(smem);
    float* buf1 = buf0 + kBrick;
    uint64_t* bars = reinterpret_cast<uint64_t*>(buf1 + kBrick);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // every CTA streams its own lane of bricks along x through the volume (y, z start depend on the CTA)
    const int tiles_y = (ny + BY - 1) / BY, tiles_z = (nz + BZ - 1) / BZ;
    const int ty = blockIdx.x % tiles_y, tz = (blockIdx.x / tiles_y) % tiles_z;
    const int y0 = ty * (BY / 2), z0 = tz * (BZ / 2);  // neighbouring tiles overlap by half a brick (halo + obliquity)
    const int nbx = (nx + BX - 1) / BX;
    // the ray of this thread inside the brick: pixel (pi, pj) of the 16 x 16 tile, 0.73 voxel pitch, oblique direction
    const int pi = tid & 15, pj = tid >> 4;
    const float ry0 = 1.0f + 0.73f * (float)pi, rz0 = 1.0f + 0.73f * (float)pj;
    const float dy = (float)(BY - 14) / (float)BX * 0.9f, dz = 0.02f;  // crosses most of the brick's spare width
    float acc = 0.0f;
    unsigned long long nsamp = 0;
    if (tid == 0) {
        mbar_expect_tx(&bars[0], kBrick * 4);
        tma_load_3d(buf0, &map, y0, 0, z0, &bars[0]);
    }
    for (int b = 0; b < bricks_per_cta; ++b) {
        const int cur = b & 1;
        if (tid == 0 && b + 1 < bricks_per_cta) {  // prefetch the next brick into the other buffer
            mbar_expect_tx(&bars[cur ^ 1], kBrick * 4);
            tma_load_3d(cur ? buf0 : buf1, &map, y0, ((b + 1) % nbx) * BX, z0 + ((b + 1) / nbx) * BZ, &bars[cur ^ 1]);
        }
        mbar_wait(&bars[cur], (b >> 1) & 1);
        const float* __restrict__ B = cur ? buf1 : buf0;
        if (SAMPLE) {
            // 5 samples per voxel along x over the brick's depth (BX - 1 cells), manual trilinear, CPU convention weights
#pragma unroll 2
            for (int s = 0; s < (BX - 1) * 5; ++s) {
                const float fx = 0.2f * (float)s, fy = ry0 + dy * fx, fz = rz0 + dz * fx;
                const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
                const float wx = fx - (float)ix, wy = fy - (float)iy, wz = fz - (float)iz;
                const float* p = B + (iz * BX + ix) * BY + iy;
                const float v000 = p[0], v010 = p[1], v100 = p[BY], v110 = p[BY + 1];
                const float v001 = p[BX * BY], v011 = p[BX * BY + 1], v101 = p[BX * BY + BY], v111 = p[BX * BY + BY + 1];
                const float v00 = fmaf(wz, v001 - v000, v000), v01 = fmaf(wz, v011 - v010, v010);
                const float v10 = fmaf(wz, v101 - v100, v100), v11 = fmaf(wz, v111 - v110, v110);
                const float v0 = fmaf(wy, v01 - v00, v00), v1 = fmaf(wy, v11 - v10, v10);
                acc += fmaf(wx, v1 - v0, v0);
            }
            nsamp += (BX - 1) * 5;
        } else {
            acc += B[tid % kBrick];
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled two iterations on
    }
    out[blockIdx.x * 256 + tid] = acc;
    if (SAMPLE) atomicAdd(counters, nsamp);
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BY, int BX, int BZ>
static void run(encode_fn encode, float* d_vol, int n, int sms, const char* label) {
    CUtensorMap map;
    cuuint64_t gdim[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)n};              // y fastest, then x, then z
    cuuint64_t gstr[2] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4};
    cuuint32_t box[3] = {BY, BX, BZ}, estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_vol, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
        exit(1);
    }
    const size_t smem = (size_t)2 * BY * BX * BZ * 4 + 64;
    const int bricks = 512;
    float* d_out;
    unsigned long long* d_cnt;
    CK(cudaMalloc(&d_cnt, 8));
    for (int sample = 0; sample < 2; ++sample) {
        auto kern = sample ? probe_kernel<BY, BX, BZ, true> : probe_kernel<BY, BX, BZ, false>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        const int grid = sms * occ;
        CK(cudaMalloc(&d_out, (size_t)grid * 256 * 4));
        CK(cudaMemset(d_cnt, 0, 8));
        kern<<<grid, 256, smem>>>(map, n, n, n, bricks, d_out, d_cnt);  // warm-up
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(d_cnt, 0, 8));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        kern<<<grid, 256, smem>>>(map, n, n, n, bricks, d_out, d_cnt);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long cnt = 0;
        CK(cudaMemcpy(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost));
        const double bytes = (double)grid * bricks * BY * BX * BZ * 4;
        printf("%-28s box %2dx%2dx%2d (y,x,z) %6.1f KB  %d CTA/SM  %s: %7.3f ms  TMA %7.1f GB/s", label, BY, BX, BZ, BY * BX * BZ * 4 / 1024.0, occ,
               sample ? "stage + sample" : "stage only    ", ms, bytes / ms / 1e6);
        if (sample) printf("  %8.1f Gsamples/s (8-tap trilinear from shared memory)", (double)cnt / ms / 1e6);
        printf("\n");
        CK(cudaFree(d_out));
    }
    CK(cudaFree(d_cnt));
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n = 1024;
    float* d_vol;
    CK(cudaMalloc(&d_vol, (size_t)n * n * n * 4));
    CK(cudaMemset(d_vol, 0, (size_t)n * n * n * 4));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
        fprintf(stderr, "cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    encode_fn encode = (encode_fn)fn;
    printf("TMA brick probe, %d SMs, volume %d^3 fp32 (4 GiB, streamed through L2)\n", sms, n);
    run<32, 8, 24>(encode, d_vol, n, sms, "oblique 45 deg, depth 8");
    run<40, 16, 28>(encode, d_vol, n, sms, "oblique 45 deg, depth 16");
    run<16, 16, 16>(encode, d_vol, n, sms, "axis aligned, depth 16");
    run<24, 32, 16>(encode, d_vol, n, sms, "mildly oblique, depth 32");
    CK(cudaFree(d_vol));
    return 0;
}
