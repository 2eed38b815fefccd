// mma_probe.cu -- what one B200 SM's tensor core sustains per tcgen05.mma shape, with NO TMA and NO epilogue:
// a single thread per CTA (CTA pair for cta_group::2) issues a long chain of SS MMAs over zeroed shared memory.
// Tells the Gram kernel (gb_gram.cu) which instruction shape can reach the pipe's rate at all.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../gauss_b200/csrc -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "probe_ptx.cuh"

using namespace gb;

enum { K_F16 = 0, K_I8 = 1, K_F8F6F4 = 2, K_MXF4 = 3 };

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int KIND, int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t sf, uint32_t acc) {
  if (CG == 1) {
    if (KIND == K_F16)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    if (KIND == K_I8) ptx::mma_i8_ss(d, da, db, idesc, acc);
    if (KIND == K_F8F6F4) ptx::mma_f8f6f4_ss(d, da, db, idesc, acc);
    if (KIND == K_MXF4) ptx::mma_mxf4_ss(d, da, db, idesc, sf, sf, acc);
  } else {
    if (KIND == K_F16)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    if (KIND == K_I8)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    if (KIND == K_F8F6F4)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    if (KIND == K_MXF4)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(sf), "r"(sf) : "memory");
  }
}

constexpr int A_BYTES = 128 * 128;
constexpr int B_BYTES = 256 * 128;
constexpr int STAGE = A_BYTES + B_BYTES;
constexpr int NSTAGE = 4;

template <int KIND, int CG, int N>
__global__ void __launch_bounds__(384, 1) probe(int iters, long long* cyc, int commit_every, int fence_every, int ld_segs, int ld_shape, int data, int dswitch = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[NSTAGE];
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < NSTAGE * STAGE / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    uint32_t w = 0;
    if (data == 1) {  // E2M1 dosages 0 / 1 / 2 (nibbles 0x0, 0x2, 0x4), allele frequency ~0.3
      for (int n = 0; n < 8; n++) { const uint32_t r = (h >> (4 * n)) & 15u; w |= (r < 8 ? 0u : r < 14 ? 2u : 4u) << (4 * n); }
    } else if (data == 2) w = h & 0x77777777u;   // random non-negative E2M1
    reinterpret_cast<uint32_t*>(smem)[i] = w;
  }
  const int warp = threadIdx.x >> 5;
  const uint32_t crank = CG == 2 ? ptx::cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    for (int i = 0; i < NSTAGE; i++) ptx::mbar_init(&bar2[i], 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  if (warp == 0) {
    if (CG == 1) {
      ptx::tmem_alloc(&tmem_ptr, 512);
      ptx::tmem_relinquish();
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0 && crank == 0) {
    constexpr int M = CG == 2 ? 256 : 128;
    const uint32_t idesc = KIND == K_F16 ? idesc_bf16(M, N) : KIND == K_I8 ? ptx::make_idesc_i8(M, N)
                           : KIND == K_F8F6F4 ? ptx::make_idesc_f8f6f4(5, M, N) : ptx::make_idesc_mxf4(M, N);
    const uint64_t desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
    const uint32_t sf = tmem + 384;   // accumulators: columns [0, 256); scale factors (zeros) further up
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 4 * NSTAGE) {
#pragma unroll
      for (int s = 0; s < NSTAGE; s++) {
        const uint64_t da = desc0 + (uint64_t)(s * (STAGE >> 4));
        const uint64_t db = da + (A_BYTES >> 4);
#pragma unroll
        // dswitch bit 0: every group of 4 MMAs goes to the next of three accumulators; bit 1: its first MMA overwrites
        const uint32_t dd = tmem + ((dswitch & 1) ? (uint32_t)(((it >> 2) + s) % 3) * 128u : 0u);
        for (int k = 0; k < 4; k++)
          mma<KIND, CG>(dd, da + 2 * k, db + 2 * k, idesc, sf, (dswitch & 2) ? (k ? 1u : 0u) : ((it | s | k) ? 1u : 0u));
        if (commit_every) {   // the Gram kernel's per-K-block release of a shared-memory stage
          if (CG == 1) ptx::mma_commit(&bar2[s]);
          else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(&bar2[s])), "h"((uint16_t)3) : "memory");
        }
        if (fence_every) ptx::tc_fence_after();
      }
    }
    if (CG == 1) ptx::mma_commit(&bar);
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    ptx::mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) *cyc = t1 - t0;
  }
  if (warp >= 4 && ld_segs > 0) {
    // the Gram epilogue's TMEM traffic: 8 warps read a 128 x 128 fp32 accumulator (another buffer than the MMA's)
    const uint32_t taddr = tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 64;
    uint32_t sink = 0;
    long long t0 = clock64();
    for (int sgm = 0; sgm < ld_segs; sgm++) {
      if (ld_shape == 0) {
        uint32_t v[16];
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          ptx::tmem_ld_32x32b_x16(taddr + ch * 16, v);
          ptx::tmem_ld_wait();
          sink ^= v[ch];
        }
      } else {
        uint32_t v[32];
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          ptx::tmem_ld_32x32b_x32(taddr + ch * 32, v);
          ptx::tmem_ld_wait();
          sink ^= v[ch];
        }
      }
    }
    long long t1 = clock64();
    if (sink == 0x12345u) cyc[2] = 1;
    if (blockIdx.x == 0 && threadIdx.x == 128) cyc[1] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();
  if (warp == 0) {
    ptx::tc_fence_after();
    if (CG == 1) ptx::tmem_dealloc(tmem, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}


// The Gram kernel's accumulator hand-off in isolation: the MMA thread switches to the next of NB TMEM buffers every
// seg_mmas instructions (commit -> tfull), 8 epilogue warps read the finished buffer (4 x tcgen05.ld x16) and hand
// it back (tempty).  No TMA, no fold arithmetic.
template <int NB>
__global__ void __launch_bounds__(384, 1) seg_probe(int n_segs, int seg_mmas, long long* cyc, int extra_alu) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t tfull[NB], tempty[NB], stage_bar[NSTAGE];
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < NSTAGE * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NB; i++) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
    for (int i = 0; i < NSTAGE; i++) ptx::mbar_init(&stage_bar[i], 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  if (warp == 0) { ptx::tmem_alloc(&tmem_ptr, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_mxf4(128, 128);
      const uint64_t desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
      const uint32_t sf = tmem + 3 * 128;
      int acc = 0, stage = 0; uint32_t ph = 0;
      long long t0 = clock64();
      for (int sg = 0; sg < n_segs; sg++) {
        ptx::mbar_wait(&tempty[acc], ph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d = tmem + acc * 128;
        uint32_t accu = 0;
        for (int m = 0; m < seg_mmas; m += 4) {
          const uint64_t da = desc0 + (uint64_t)(stage * (STAGE >> 4));
          const uint64_t db = da + (A_BYTES >> 4);
          ptx::mma_mxf4_ss(d, da, db, idesc, sf, sf, accu);
          ptx::mma_mxf4_ss(d, da + 2, db + 2, idesc, sf, sf, 1);
          ptx::mma_mxf4_ss(d, da + 4, db + 4, idesc, sf, sf, 1);
          ptx::mma_mxf4_ss(d, da + 6, db + 6, idesc, sf, sf, 1);
          accu = 1;
          ptx::mma_commit(&stage_bar[stage]);
          if (++stage == NSTAGE) stage = 0;
        }
        ptx::mma_commit(&tfull[acc]);
        if (++acc == NB) { acc = 0; ph ^= 1; }
      }
      long long t1 = clock64();
      if (blockIdx.x == 0) cyc[0] = t1 - t0;
    }
  } else if (warp >= 4) {
    int acc = 0; uint32_t ph = 0;
    uint32_t sink = 0;
    for (int sg = 0; sg < n_segs; sg++) {
      ptx::mbar_wait(&tfull[acc], ph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem + acc * 128 + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 64;
      uint32_t v[2][16];
      ptx::tmem_ld_32x32b_x16(taddr, v[0]);
#pragma unroll
      for (int ch = 0; ch < 4; ch++) {
        ptx::tmem_ld_wait();
        if (ch < 3) ptx::tmem_ld_32x32b_x16(taddr + (ch + 1) * 16, v[(ch + 1) & 1]);
        else {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
        }
#pragma unroll
        for (int e = 0; e < 16; e++) {
          uint32_t x = v[ch & 1][e];
          for (int a = 0; a < extra_alu; a++) x = x * 0x9E3779B1u + 0x7F4A7C15u;   // stand-in for the fold's issue slots
          sink ^= x;
        }
      }
      if (++acc == NB) { acc = 0; ph ^= 1; }
    }
    if (sink == 0x12345u) cyc[2] = 1;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int NB>
void run_seg(int seg_mmas, int extra_alu, long long* cyc) {
  auto kern = seg_probe<NB>;
  const int smem = NSTAGE * STAGE + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int total_mmas = 16384;
  const int n_segs = total_mmas / seg_mmas;
  for (int rep = 0; rep < 2; rep++) {
    kern<<<148, 384, smem>>>(n_segs, seg_mmas, cyc, extra_alu);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("seg_probe FAILED: %s\n", cudaGetErrorString(err)); return; }
  }
  printf("segmented NB=%d  %4d MMAs/segment (%5d clk of tensor work)  alu %2d: %8.1f clk/segment  %6.1f clk/MMA  tensor %.0f %%\n", NB, seg_mmas,
         seg_mmas * 64, extra_alu, (double)cyc[0] / n_segs, (double)cyc[0] / (n_segs * seg_mmas), 6400.0 * n_segs * seg_mmas / (double)cyc[0]);
}

// How asynchronous is tcgen05.mma issue?  One thread issues groups of G MMAs (same accumulator) separated by a
// dependent chain of X integer multiply-adds (~4 clk each).  If the pipe has a queue, time/group = max(64 G, ...);
// if issue blocks until the pipe is free, the chain adds to every group.
template <int G>
__global__ void __launch_bounds__(128, 1) queue_probe(int groups, int chain, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < NSTAGE * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_ptr, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_mxf4(128, 128);
    const uint64_t da = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
    const uint64_t db = da + (A_BYTES >> 4);
    uint32_t x = (uint32_t)groups;
    long long t0 = clock64();
    for (int g = 0; g < groups; g++) {
#pragma unroll
      for (int k = 0; k < G; k++) ptx::mma_mxf4_ss(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, tmem + 384, tmem + 384, 1);
      for (int c = 0; c < chain; c++) x = x * 0x9E3779B1u + 12345u;
    }
    ptx::mma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[3] = x; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}
template <int G>
void run_queue(long long* cyc) {
  auto kern = queue_probe<G>;
  const int smem = NSTAGE * STAGE + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int chain : {0, 8, 16, 32, 64, 128}) {
    const int groups = 8192 / G;
    for (int rep = 0; rep < 2; rep++) { kern<<<148, 128, smem>>>(groups, chain, cyc); cudaDeviceSynchronize(); }
    printf("queue: %d MMAs/group (%4d clk of tensor work), chain %3d IMADs: %7.1f clk/group\n", G, 64 * G, chain, (double)cyc[0] / groups);
  }
}

template <int KIND, int CG, int N>
void run(const char* name, int kper, long long* cyc, int commit_every = 0, int fence_every = 0, int ld_segs = 0, int ld_shape = 0, int data = 0, int iters = 8192, int dswitch = 0) {
  auto kern = probe<KIND, CG, N>;
  const int smem = NSTAGE * STAGE + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(148 / CG * CG);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cyc[1] = 0;
    cudaLaunchKernelEx(&cfg, kern, iters, cyc, commit_every, fence_every, ld_segs, ld_shape, data, dswitch);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%-10s cg%d N=%3d  FAILED: %s\n", name, CG, N, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double clk_per_mma = (double)*cyc / iters;
  const double mac_per_sm = 128.0 * N * kper;   // per SM and instruction (a pair's M = 256 is two SMs)
  const double smem_bytes = 128.0 * 32 + (double)N / CG * 32;  // operand bytes each SM's shared memory supplies
  printf("%-10s data%d c%d f%d cg%d M=%3d N=%3d K=%2d  %7.1f clk/MMA  %8.0f MAC/clk/SM  %6.1f smem B/clk/SM  (%.3f ms, chip %.0f TOP/s)\n", name, data, commit_every, fence_every, CG,
         CG * 128, N, kper, clk_per_mma, mac_per_sm / clk_per_mma, smem_bytes / clk_per_mma, best,
         2.0 * mac_per_sm * iters * (148 / CG * CG) / (best * 1e-3) / 1e12);
  if (ld_segs) printf("           + 8 warps reading %d accumulators (shape %d): %.0f clk per 64 KB accumulator = %.1f B/clk\n", ld_segs, ld_shape,
                      (double)cyc[1] / ld_segs, 65536.0 * ld_segs / (double)cyc[1]);
}

int main(int argc, char** argv) {
  long long* cyc;
  cudaMallocManaged(&cyc, 32);
  if (argc > 1 && argv[1][0] == 'q') {
    run_queue<1>(cyc); run_queue<2>(cyc); run_queue<4>(cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'w') {   // cost of switching accumulators / overwriting, no barriers involved
    for (int dsw : {0, 1, 2, 3}) {
      printf("dswitch %d: ", dsw);
      run<K_MXF4, 1, 128>("mxf4", 64, cyc, 0, 0, 0, 0, 0, 8192, dsw);
      printf("dswitch %d: ", dsw);
      run<K_I8, 1, 128>("i8", 32, cyc, 0, 0, 0, 0, 0, 8192, dsw);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'd') {   // operand data dependence (power): zeros vs dosages vs random nibbles
    for (int rep = 0; rep < 2; rep++)
      for (int data : {0, 1, 2}) {
        run<K_MXF4, 1, 128>("mxf4", 64, cyc, 1, 1, 0, 0, data, 65536);
        run<K_MXF4, 2, 256>("mxf4", 64, cyc, 1, 1, 0, 0, data, 32768);
        run<K_F8F6F4, 1, 128>("f8f6f4", 32, cyc, 1, 1, 0, 0, data, 65536);
        run<K_I8, 1, 128>("i8", 32, cyc, 1, 1, 0, 0, data, 65536);
      }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (argc > 1) {
    for (int alu : {0, 2, 4, 8})
      for (int sm : {4, 8, 16, 32, 128}) {
        run_seg<3>(sm, alu, cyc);
        if (alu == 0 || alu == 4) { run_seg<2>(sm, alu, cyc); run_seg<4>(sm, alu, cyc); }
      }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  run<K_F16, 1, 128>("bf16", 16, cyc);
  run<K_F16, 1, 256>("bf16", 16, cyc);
  run<K_F16, 2, 128>("bf16", 16, cyc);
  run<K_F16, 2, 256>("bf16", 16, cyc);
  run<K_I8, 1, 128>("i8", 32, cyc);
  run<K_I8, 1, 256>("i8", 32, cyc);
  run<K_I8, 2, 128>("i8", 32, cyc);
  run<K_I8, 2, 256>("i8", 32, cyc);
  run<K_F8F6F4, 1, 128>("f8f6f4", 32, cyc);
  run<K_F8F6F4, 1, 256>("f8f6f4", 32, cyc);
  run<K_F8F6F4, 2, 256>("f8f6f4", 32, cyc);
  run<K_MXF4, 1, 64>("mxf4", 64, cyc);
  run<K_MXF4, 1, 128>("mxf4", 64, cyc);
  run<K_MXF4, 1, 256>("mxf4", 64, cyc);
  run<K_MXF4, 2, 128>("mxf4", 64, cyc);
  run<K_MXF4, 2, 256>("mxf4", 64, cyc);
  run<K_MXF4, 1, 128>("mxf4", 64, cyc, 1, 1, 512, 0);
  run<K_MXF4, 1, 128>("mxf4", 64, cyc, 1, 1, 512, 1);
  run<K_MXF4, 1, 128>("mxf4", 64, cyc, 1, 1, 2048, 0);
  run<K_MXF4, 1, 128>("mxf4", 64, cyc, 1, 1, 2048, 1);
  run<K_I8, 1, 128>("i8", 32, cyc, 1, 1, 2048, 0);
  run<K_MXF4, 1, 128>("mxf4", 64, cyc, 1, 0);
  run<K_MXF4, 1, 128>("mxf4", 64, cyc, 0, 1);
  run<K_MXF4, 1, 128>("mxf4", 64, cyc, 1, 1);
  run<K_MXF4, 2, 128>("mxf4", 64, cyc, 1, 1);
  run<K_I8, 1, 128>("i8", 32, cyc, 1, 1);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
