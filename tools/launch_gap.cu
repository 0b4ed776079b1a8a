// Where do the ~10 us between consecutive tcgen05 GEMM launches go?  Back-to-back timing of kernels that add one ingredient at a time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/launch_gap.cu -o /tmp/lg && /tmp/lg
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) k(float* out, int ncols) {
    extern __shared__ uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar[4];
    if (VARIANT >= 3 && threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (VARIANT >= 2 && (threadIdx.x >> 5) == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (VARIANT >= 4) {      // write a 128 x 128 f32 tile like an epilogue would
        float* o = out + (size_t)blockIdx.x * 128 * 128;
        for (int i = threadIdx.x; i < 128 * 128; i += 256) o[i] = (float)i;
    }
    if (threadIdx.x == 0 && out == nullptr) sm[0] = 1;
    __syncthreads();
    if (VARIANT >= 2 && (threadIdx.x >> 5) == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(ncols) : "memory");
    }
}

template <int V>
static int run(const char* name, int grid, int smem, float* out) {
    CK(cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 10; ++i) k<V><<<grid, 256, smem>>>(out, 128);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 200; ++i) k<V><<<grid, 256, smem>>>(out, 128);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s grid %4d smem %6d : %6.2f us / launch\n", name, grid, smem, ms * 1e3 / 200);
    return 0;
}

int main() {
    float* out; CK(cudaMalloc(&out, (size_t)1024 * 128 * 128 * 4));
    for (int grid : {128, 384}) {
        run<0>("empty", grid, 0, out);
        run<1>("empty + 198 KB dynamic smem", grid, 198 * 1024, out);
        run<2>("+ tcgen05.alloc/dealloc 128 cols", grid, 198 * 1024, out);
        run<3>("+ mbarrier init", grid, 198 * 1024, out);
        run<4>("+ 64 KB of f32 stores per CTA", grid, 198 * 1024, out);
        run<4>("same with 96 KB smem (2 CTAs/SM)", grid, 96 * 1024, out);
    }
    return 0;
}
