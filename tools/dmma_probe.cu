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
// Are the fp64 tensor pipe (DMMA) and the fp64 FMA pipe (DFMA) the same execution resource?  Half of the warps of a
// CTA run DMMAs, the other half independent DFMAs; if the two pipes were separate the combined rate would approach the
// sum of the two alone.
template <int MIX>   // 0: all warps DMMA, 1: all warps DFMA, 2: even warps DMMA / odd warps DFMA
__global__ void mix_probe(double* out, int iters, double seed, long long* cyc) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = MIX == 0 || (MIX == 2 && (warp & 1) == 0);
  double c[16][2];
  for (int i = 0; i < 16; i++) { c[i][0] = seed + i; c[i][1] = seed - i; }
  const double a = seed * 0.5 + threadIdx.x, b = 1.0 + 1e-9 * threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  if (do_mma) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);          // 16 x 256 FMA per warp
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) { c[i][0] = fma(c[i][0], b, a); c[i][1] = fma(c[i][1], b, a); }   // 128 x 32 FMA per warp
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
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
  for (int threads : {256, 512}) {
    for (int mix = 0; mix < 3; mix++) {
      for (int rep = 0; rep < 2; rep++) {
        if (mix == 0) mix_probe<0><<<148, threads>>>(out, iters, 1.5, cyc);
        if (mix == 1) mix_probe<1><<<148, threads>>>(out, iters, 1.5, cyc);
        if (mix == 2) mix_probe<2><<<148, threads>>>(out, iters, 1.5, cyc);
        cudaDeviceSynchronize();
      }
      const double warps = threads / 32;
      const double fma = mix == 2 ? warps / 2 * iters * 4096.0 * 2 : warps * iters * 4096.0;   // both kinds do 4096 FMA per warp-iteration
      printf("threads/SM %4d  %-28s %8.1f fp64 FMA / clk / SM  (%lld cycles)\n", threads,
             mix == 0 ? "DMMA m8n8k4 only" : mix == 1 ? "DFMA only" : "half DMMA + half DFMA warps", fma / (double)*cyc, *cyc);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
