// epi_probe.cu -- how fast can 8 warps fold one 128x128 int32 accumulator segment into fp64?
// Mimics the Gram epilogue's per-element work without TMEM / barriers: d = m*S - sa*sb; acc += coef*double(d).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double i2d(int v) { return __dsub_rn(__hiloint2double(0x43300000, v ^ 0x80000000), 4503601774854144.0); }
template <int MODE>
__global__ void __launch_bounds__(384, 1) probe(const int* __restrict__ src, double* out, int nseg, long long* cyc) {
  __shared__ int sB[30 * 128];
  for (int i = threadIdx.x; i < 30 * 128; i += blockDim.x) sB[i] = src[i] & 1023;
  __syncthreads();
  if (threadIdx.x < 128) { if (threadIdx.x == 0 && MODE < 0) out[0] = 0; return; }   // control warps idle
  const int t = threadIdx.x - 128;
  double acc[64];
#pragma unroll
  for (int e = 0; e < 64; e++) acc[e] = 0.0;
  long long t0 = clock64();
  for (int s = 0; s < nseg; s++) {
    const int m = 100 + (s % 21) * 37, sa = src[(s * 256 + t) & 4095] & 1023;
    const double coef = 1.0 + s * 0.001;
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
      int v[16];
#pragma unroll
      for (int e = 0; e < 16; e++) v[e] = src[((s * 4 + ch) * 16 + e) * 256 + t] ;   // stands in for tcgen05.ld
      const int4* sBv = reinterpret_cast<const int4*>(sB + (s % 21) * 128 + (t >> 7) * 64 + ch * 16);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int4 b4 = sBv[q];
        const int bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int d = m * v[q * 4 + k] - sa * bb[k];
          if (MODE == 0) acc[ch * 16 + q * 4 + k] = __dadd_rn(acc[ch * 16 + q * 4 + k], __dmul_rn(coef, i2d(d)));
          if (MODE == 1) acc[ch * 16 + q * 4 + k] = fma(coef, i2d(d), acc[ch * 16 + q * 4 + k]);
          if (MODE == 2) acc[ch * 16 + q * 4 + k] = fma(coef, (double)d, acc[ch * 16 + q * 4 + k]);
        }
      }
    }
  }
  long long t1 = clock64();
  double sum = 0;
#pragma unroll
  for (int e = 0; e < 64; e++) sum += acc[e];
  out[blockIdx.x * 256 + t] = sum;
  if (t == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  int* src; double* out; long long* cyc;
  const int nseg = 210;
  cudaMalloc(&src, (size_t)(nseg * 64 + 64) * 256 * 4 + 65536);
  cudaMemset(src, 1, (size_t)(nseg * 64 + 64) * 256 * 4 + 65536);
  cudaMalloc(&out, 148 * 256 * 8);
  cudaMallocManaged(&cyc, 8);
  const char* names[] = {"magic cvt + DMUL + DADD", "magic cvt + DFMA", "I2F + DFMA"};
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      if (mode == 0) probe<0><<<148, 384>>>(src, out, nseg, cyc);
      if (mode == 1) probe<1><<<148, 384>>>(src, out, nseg, cyc);
      if (mode == 2) probe<2><<<148, 384>>>(src, out, nseg, cyc);
      cudaDeviceSynchronize();
    }
    printf("%-26s %8.0f cycles per 128x128 segment (8 warps)\n", names[mode], (double)*cyc / nseg);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
