// How many thread-block clusters of a given shape can be resident at once?  (run on the GPU box)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k512(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
__global__ void __launch_bounds__(256, 1) k256(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
template <typename K> static void q(K kern, const char* name, int threads, int cs, int smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs, 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    printf("%s threads %d cluster %2d smem %6d : max active clusters %d (%s)\n", name, threads, cs, smem, n, cudaGetErrorString(e));
}
int main() {
    for (int cs : {2, 4, 8, 16})
        for (int smem : {32 * 1024, 100 * 1024, 203 * 1024}) { q(k512, "k512", 512, cs, smem); q(k256, "k256", 256, cs, smem); }
    return 0;
}
