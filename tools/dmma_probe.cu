// dmma_probe.cu -- fp64 tensor-core (mma.sync m8n8k4 / m16n8k8 .f64) issue rate per SM on B200, against DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
template <int MODE>
__global__ void probe(double* out, int iters, double seed, long long* cyc) {
  double c[8][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = seed + i + j;
  double a[4] = {seed, seed * 0.5, seed + 1, seed - 1}, b[2] = {seed * 0.25 + threadIdx.x, 1.0};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { dmma884(c[i][0], c[i][1], a[0], b[0]); dmma884(c[i][2], c[i][3], a[1], b[1]); }
      if (MODE == 1) dmma1688(c[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMallocManaged(&cyc, 8);
  const int iters = 4000;
  for (int threads : {128, 256, 512}) {
    for (int mode = 0; mode < 2; mode++) {
      for (int rep = 0; rep < 2; rep++) {
        if (mode == 0) probe<0><<<148, threads>>>(out, iters, 1.5, cyc);
        else probe<1><<<148, threads>>>(out, iters, 1.5, cyc);
        cudaDeviceSynchronize();
      }
      // FMAs per warp per iteration: mode0: 16 x m8n8k4 (256) ; mode1: 8 x m16n8k8 (1024)
      double fma = (double)(threads / 32) * iters * (mode == 0 ? 16 * 256.0 : 8 * 1024.0);
      printf("threads/SM %4d  %-10s  %8.1f fp64 FMA / clk / SM  (%lld cycles)\n", threads, mode ? "m16n8k8" : "m8n8k4", fma / (double)*cyc, *cyc);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
