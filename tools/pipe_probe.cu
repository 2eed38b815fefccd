// pipe_probe.cu -- micro-benchmark of the issue rates the Gram epilogue depends on (B200, per SM):
// fp64 FMA / MUL+ADD, int32 -> fp64 conversion, 32x32->64 integer multiply-add.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(double* out, int iters, double seed, long long* cyc) {
  double a[8];
  long long acc[8];
  int iv[8];
  for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; acc[i] = threadIdx.x + i; iv[i] = threadIdx.x * 7 + i; }
  const double m = seed * 0.5 + 1.0;
  const int im = (int)seed + 3;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = fma(a[i], m, 1.0);                                   // DFMA
      if (MODE == 1) a[i] = __dadd_rn(a[i], __dmul_rn(m, (double)iv[i])), iv[i] += im;   // I2F + DMUL + DADD + IADD
      if (MODE == 2) acc[i] += (long long)iv[i] * (long long)im, iv[i] += 1;     // IMAD.WIDE + IADD
      if (MODE == 3) a[i] = __dadd_rn(a[i], (double)iv[i]), iv[i] += im;         // I2F + DADD
      if (MODE == 4) a[i] = __dadd_rn(a[i], __hiloint2double(0x43300000, iv[i] ^ 0x80000000) - 4503601774854144.0), iv[i] += im;  // magic convert
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i] + (double)acc[i] + iv[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMallocManaged(&cyc, 8);
  const int iters = 20000;
  const char* names[] = {"DFMA", "I2F+DMUL+DADD", "IMAD.WIDE", "I2F+DADD", "magic-cvt(LOP+DADD)+DADD"};
  for (int threads : {128, 256, 512}) {
    for (int mode = 0; mode < 5; mode++) {
      for (int rep = 0; rep < 2; rep++) {
        if (mode == 0) probe<0><<<148, threads>>>(out, iters, 1.5, cyc);
        if (mode == 1) probe<1><<<148, threads>>>(out, iters, 1.5, cyc);
        if (mode == 2) probe<2><<<148, threads>>>(out, iters, 1.5, cyc);
        if (mode == 3) probe<3><<<148, threads>>>(out, iters, 1.5, cyc);
        if (mode == 4) probe<4><<<148, threads>>>(out, iters, 1.5, cyc);
        cudaDeviceSynchronize();
      }
      double per_sm_elem = (double)threads * 8 * iters / (double)*cyc;
      printf("threads/SM %4d  %-26s  %8.2f element-ops / clk / SM   (%lld cycles)\n", threads, names[mode], per_sm_elem, *cyc);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
