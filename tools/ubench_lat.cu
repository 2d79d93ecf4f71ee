// Tuning probe (not product code): dependent-issue latency and per-scheduler throughput of the instructions the MPPI
// rollout loop is made of, on the GPU at hand.  nvcc -arch=sm_100a -O3 tools/ubench_lat.cu -o /tmp/ubench && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, float *outf, int iters, double a0, double b0)
{
  double x = a0 + threadIdx.x, y = b0, z = 0.5;
  float f = (float)a0;
  __shared__ double sm[256];
  sm[threadIdx.x & 255] = threadIdx.x & 31;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (OP == 0) x = fma(x, y, z);
      if (OP == 1) x = x + y;
      if (OP == 2) x = x * y;
      if (OP == 3) f = fmaf(f, (float)b0, 0.5f);
      if (OP == 4) { int lo = __double2loint(x), hi = __double2hiint(x); lo = __shfl_up_sync(0xffffffffu, lo, 1, 16); hi = __shfl_up_sync(0xffffffffu, hi, 1, 16); x = __hiloint2double(hi, lo); }
      if (OP == 5) { f = (float)x; x = (double)f + y; }
      if (OP == 6) { x = sm[(int)x & 31] ; }
      if (OP == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); }
      if (OP == 8) { float q = (float)x; x = (double)q; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (double)(t1 - t0) / (iters * 16.0); }
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = x + f;
}

int main()
{
  double *d; float *df;
  cudaMalloc(&d, 1 << 24); cudaMalloc(&df, 1 << 20);
  const char *names[] = {"DFMA", "DADD", "DMUL", "FFMA", "SHFL x2 (64-bit)", "F2F f64->f32 + f32->f64 + DADD", "LDS.64 dependent", "MUFU.EX2", "F2F round trip"};
  for (int warps = 1; warps <= 8; warps *= 2) {
    printf("---- %d warp(s) per scheduler (block of %d threads, one block per SM) ----\n", warps, warps * 128);
    for (int op = 0; op < 9; op++) {
      double h = 0;
      auto run = [&](auto k) { k<<<148, warps * 128>>>(d, df, 2000, 1.0000001, 0.9999999); cudaDeviceSynchronize(); k<<<148, warps * 128>>>(d, df, 2000, 1.0000001, 0.9999999); cudaDeviceSynchronize(); };
      switch (op) {
        case 0: run(chain<0>); break; case 1: run(chain<1>); break; case 2: run(chain<2>); break; case 3: run(chain<3>); break;
        case 4: run(chain<4>); break; case 5: run(chain<5>); break; case 6: run(chain<6>); break; case 7: run(chain<7>); break; case 8: run(chain<8>); break;
      }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%-34s %.1f cycles per dependent op (per warp)\n", names[op], h);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
