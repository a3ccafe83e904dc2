// How many thread-block clusters of size 8 / 16 can be co-resident on this GPU for a kernel shaped
// like the cluster LSTM (512 threads, ~64-200 KB dynamic shared memory)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(int *p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s SMs %d\n", pr.name, pr.multiProcessorCount);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int smem : {32 * 1024, 64 * 1024, 100 * 1024, 160 * 1024, 200 * 1024}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int cs : {2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf("smem %3d KB cluster %2d -> max active clusters %d (%s)\n", smem / 1024, cs, n, cudaGetErrorString(e));
        }
    }
    return 0;
}
