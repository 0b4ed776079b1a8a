// How long does tcgen05.dealloc take, and is it serialised across the chip?  (Round-2 finding behind the GEMM's tile policy.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_dealloc_probe tools/tmem_dealloc_probe.cu && ./tmem_dealloc_probe
// Each CTA allocates `cols` TMEM columns, optionally runs `n_mma` 128 x N x 16 bf16 MMAs into the first `used` columns (operands:
// zeroed shared memory), waits for them, optionally reads the accumulator back, then times its dealloc with %globaltimer.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__global__ void __launch_bounds__(128) probe(int cols, int used, int n_mma, int do_ld, int smem_pad, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) ((uint32_t*)sm)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = slot;
    if (threadIdx.x == 32 && n_mma > 0) {
        const uint64_t ad = desc_sw128(smem_u32(sm)), bd = desc_sw128(smem_u32(sm + 16384));
        for (int c0 = 0; c0 < used; c0 += 256) {
            const int n = min(256, used - c0);
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int i = 0; i < n_mma; ++i)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(base + c0), "l"(ad), "l"(bd), "r"(idesc), "r"(i) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (n_mma > 0) {
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        } while (!ok);
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (do_ld) {
        uint32_t r0, r1, r2, r3, acc = 0;
        for (int c = 0; c < used; c += 4) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + ((uint32_t)(warp * 32) << 16) + c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += r0 + r1 + r2 + r3;
        }
        if (acc == 0x12345678u) out[0] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    unsigned long long t0 = 0, t1 = 0;
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        t0 = gtime();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
        t1 = gtime();
        if (threadIdx.x == 0) { out[2 * blockIdx.x + 8] = t0; out[2 * blockIdx.x + 9] = t1; }
    }
}
int main() {
    unsigned long long* d;
    cudaMalloc(&d, (8 + 2 * 1024) * 8);
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct C { int G, cols, used, n_mma, ld, pad; } cs[] = {
        {148, 512, 0, 0, 0, 0}, {148, 512, 512, 1, 0, 0}, {148, 512, 512, 64, 0, 0}, {148, 512, 512, 64, 1, 0}, {148, 512, 256, 64, 1, 0},
        {148, 256, 256, 64, 1, 0}, {148, 128, 128, 64, 1, 0}, {148, 64, 64, 64, 1, 0}, {74, 512, 512, 64, 1, 0}, {32, 512, 512, 64, 1, 0},
        {1, 512, 512, 64, 1, 0}, {148, 512, 512, 64, 1, 150 * 1024}, {296, 256, 256, 64, 1, 0}, {592, 128, 128, 64, 1, 0}};
    for (const C& c : cs) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(d, 0, (8 + 2 * 1024) * 8);
            probe<<<c.G, 128, smem + c.pad>>>(c.cols, c.used, c.n_mma, c.ld, c.pad, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        }
        std::vector<unsigned long long> h(8 + 2 * 1024);
        cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
        double mean = 0, mx = 0; unsigned long long t0min = ~0ull, t1max = 0;
        for (int i = 0; i < c.G; ++i) {
            const double dt = (double)(h[9 + 2 * i] - h[8 + 2 * i]);
            mean += dt; mx = std::max(mx, dt); t0min = std::min(t0min, h[8 + 2 * i]); t1max = std::max(t1max, h[9 + 2 * i]);
        }
        printf("G=%3d cols=%3d used=%3d mma=%2d ld=%d smem=%3dK: dealloc mean %.2f us, max %.2f us; first dealloc start -> last dealloc end %.2f us\n",
               c.G, c.cols, c.used, c.n_mma, c.ld, (smem + c.pad) / 1024, mean / c.G / 1e3, mx / 1e3, (t1max - t0min) / 1e3);
    }
    return 0;
}
