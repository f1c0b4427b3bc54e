// Why does a (1,2,1)-cluster launch of the GEMM fail with cudaErrorInvalidClusterSize?  Same launch shape with a
// trivial kernel: cluster dims x dynamic smem x threads x programmatic-serialization attribute.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

__global__ void __launch_bounds__(576, 1) k(int* out) {
  extern __shared__ unsigned char smem[];
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  if (threadIdx.x == 0) { smem[0] = (unsigned char)r; atomicAdd(out, 1 + (int)r * 1000); }
}

static void attempt(int gx, int gy, int cy, int cx, int threads, int smem, bool pdl) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(gx, gy, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl) { attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  attr[n].id = cudaLaunchAttributeClusterDimension;
  attr[n].val.clusterDim.x = cx; attr[n].val.clusterDim.y = cy; attr[n].val.clusterDim.z = 1; ++n;
  cfg.attrs = attr; cfg.numAttrs = n;
  int clusters = -1;
  cudaError_t eo = cudaOccupancyMaxActiveClusters(&clusters, k, &cfg);
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
  cudaError_t es = cudaDeviceSynchronize();
  int h = 0; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  printf("grid %dx%d cluster %dx%d threads %d smem %d pdl %d -> occupancy(%s) %d clusters, launch %s, sync %s, counter %d\n", gx, gy, cx, cy,
         threads, smem, (int)pdl, cudaGetErrorName(eo), clusters, cudaGetErrorName(e), cudaGetErrorName(es), h);
  cudaGetLastError();
  cudaFree(d);
}

int main() {
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrClusterLaunch, 0);
  printf("cudaDevAttrClusterLaunch %d\n", v);
  attempt(4, 32, 2, 1, 576, 198144, true);
  attempt(4, 32, 2, 1, 576, 198144, false);
  attempt(4, 32, 1, 2, 576, 198144, false);      // pairs along x instead
  attempt(4, 32, 2, 1, 576, 100000, false);
  attempt(4, 32, 2, 1, 320, 198144, false);
  attempt(4, 32, 2, 1, 128, 1024, false);
  attempt(2, 32, 2, 1, 576, 198144, true);
  return 0;
}
