// Which hardware warp slot (and so which of the SM's four schedulers: slot % 4) the warps of two co-resident 10-warp CTAs
// get.  nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_warpslots tools/ubench_warpslots.cu && /tmp/ubench_warpslots
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 2) k(unsigned *out)
{
  extern __shared__ char sm[];
  unsigned smid, warpid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
  if ((threadIdx.x & 31) == 0) {
    out[(blockIdx.x * 10 + (threadIdx.x >> 5)) * 2] = smid;
    out[(blockIdx.x * 10 + (threadIdx.x >> 5)) * 2 + 1] = warpid;
  }
  // stay resident long enough for the whole grid to be co-resident
  long long t0 = clock64();
  while (clock64() - t0 < 200000) { }
  if (sm[threadIdx.x] == 1) out[0] = 0;
}
int main()
{
  unsigned *d, *h = new unsigned[296 * 20];
  cudaMalloc(&d, 296 * 20 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<296, 320, 100 * 1024>>>(d);
  cudaMemcpy(h, d, 296 * 20 * 4, cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  int per[148][4] = {};
  for (int b = 0; b < 296; b++)
    for (int w = 0; w < 10; w++) per[h[(b * 10 + w) * 2]][h[(b * 10 + w) * 2 + 1] % 4]++;
  int hist[16] = {};
  for (int s = 0; s < 148; s++) for (int q = 0; q < 4; q++) hist[per[s][q] < 15 ? per[s][q] : 15]++;
  printf("warps per scheduler (slot %% 4), histogram over 148 SMs x 4:");
  for (int i = 0; i < 16; i++) if (hist[i]) printf("  %d warps: %d", i, hist[i]);
  printf("\n");
  for (int b : {0, 1, 148, 149, 295}) {
    printf("CTA %3d on SM %3u: slots", b, h[b * 20]);
    for (int w = 0; w < 10; w++) printf(" %u", h[(b * 10 + w) * 2 + 1]);
    printf("\n");
  }
  for (int s : {0, 1, 77}) {
    printf("SM %d:", s);
    for (int b = 0; b < 296; b++) if (h[b * 20] == (unsigned)s) printf(" CTA %d", b);
    printf("\n");
  }
  return 0;
}
