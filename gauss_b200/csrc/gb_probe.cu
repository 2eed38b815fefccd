// gb_probe.cu -- pipe-peak probes run on the bench box at bench time (roofline denominators).
//
// MEASURED_PEAKS.json holds HBM GB/s and dense bf16 TF/s only; the Gram kernel runs tcgen05.mma kind::i8 / kind::mxf4
// and the solve runs fp64 mma.sync m8n8k4, so their rooflines need figures of their own.  Each probe keeps ONE pipe
// busy with nothing else going on (no TMA, no epilogue, operands resident in shared memory / registers) on all SMs and
// is timed with CUDA events: what it reports is the ceiling a kernel using that instruction shape can approach on this
// GPU at its clocks under load.  Measurement helpers, not a compute path.
#include "gb_batch.cuh"
#include "gb_ptx.cuh"

namespace gb {
namespace {

constexpr int PA_BYTES = 128 * 128;           // A: 128 rows x 128 B (one swizzled K block)
constexpr int PSTAGE = 2 * PA_BYTES;          // A + B (N = 128)
constexpr int PNSTAGE = 4;

// One thread per CTA issues `iters` back-to-back 128 x 128 MMAs (K = 32 int8 / 64 E2M1 per instruction) over four
// shared-memory stages holding dosage-like data, committing to an mbarrier after every four like the Gram kernel does.
template <int MXF4>
__global__ void __launch_bounds__(128, 1) tensor_probe_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[PNSTAGE];
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < PNSTAGE * PSTAGE / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    uint32_t w = 0;
    if (MXF4) {   // E2M1 dosages 0 / 1 / 2 (nibbles 0x0, 0x2, 0x4)
      for (int n = 0; n < 8; n++) {
        const uint32_t r = (h >> (4 * n)) & 15u;
        w |= (r < 8 ? 0u : r < 14 ? 2u : 4u) << (4 * n);
      }
    } else {      // int8 dosages 0 / 1 / 2
      for (int n = 0; n < 4; n++) {
        const uint32_t r = (h >> (8 * n)) & 15u;
        w |= (r < 8 ? 0u : r < 14 ? 1u : 2u) << (8 * n);
      }
    }
    reinterpret_cast<uint32_t*>(smem)[i] = w;
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    for (int i = 0; i < PNSTAGE; i++) ptx::mbar_init(&bar2[i], 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (MXF4 && warp < 4) {   // unit E8M0 block scales in the last 128 columns
    const uint32_t sf_addr = tmem + 384 + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 128; c += 16) ptx::tmem_st_fill_32x32b_x16(sf_addr + c, 0x7F7F7F7Fu);
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
  }
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = MXF4 ? ptx::make_idesc_mxf4(128, 128) : ptx::make_idesc_i8(128, 128);
    const uint64_t desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
    const uint32_t sf = tmem + 384;
    for (int it = 0; it < iters; it += 4 * PNSTAGE) {
#pragma unroll
      for (int s = 0; s < PNSTAGE; s++) {
        const uint64_t da = desc0 + (uint64_t)(s * (PSTAGE >> 4));
        const uint64_t db = da + (PA_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t acc = (it | s | k) ? 1u : 0u;
          if (MXF4) ptx::mma_mxf4_ss(tmem, da + 2 * k, db + 2 * k, idesc, sf, sf, acc);
          else ptx::mma_i8_ss(tmem, da + 2 * k, db + 2 * k, idesc, acc);
        }
        ptx::mma_commit(&bar2[s]);
      }
    }
    ptx::mma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// fp64 tensor core: every warp issues independent mma.sync m8n8k4 (256 FMA per warp instruction) from registers.
__global__ void __launch_bounds__(512) dmma_probe_kernel(double* out, int iters, double seed) {
  double c[16][2];
  for (int i = 0; i < 16; i++) {
    c[i][0] = seed + i;
    c[i][1] = seed - i;
  }
  const double a = seed * 0.5 + threadIdx.x, b = 1.0 + 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * (long long)blockDim.x + threadIdx.x] = s;
}

// Plain device copy (read + write bytes), as MEASURED_PEAKS.json's hbm_gbs is defined.
__global__ void __launch_bounds__(256) copy_probe_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace
}  // namespace gb

using namespace gb;

// which: 0 = tcgen05.mma kind::i8 (TOP/s), 1 = tcgen05.mma kind::mxf4 (TOP/s), 2 = fp64 mma.sync m8n8k4 (TFLOP/s),
//        3 = device copy bandwidth (GB/s, read + write).  Best of `reps` launches, each ~1 ms, event-timed.
extern "C" int gb_probe_peak(gb_ctx* ctx, int which, int reps, double* value) {
  if (!ctx || !value || which < 0 || which > 3) return GB_ERR_BAD_ARG;
  GB_CUDA(cudaSetDevice(ctx->device));
  if (reps < 1) reps = 3;
  cudaEvent_t e0, e1;
  GB_CUDA(cudaEventCreate(&e0));
  GB_CUDA(cudaEventCreate(&e1));
  double best_ms = 1e30, work = 0;
  void* buf = nullptr;
  const int n_sm = ctx->sm_count;
  const size_t copy_bytes = 1ull << 30;
  if (which >= 2) GB_CUDA(cudaMalloc(&buf, which == 2 ? sizeof(double) * 512 * (size_t)n_sm : 2 * copy_bytes));
  for (int r = 0; r < reps + 1; r++) {    // first launch = warm-up
    cudaEventRecord(e0, ctx->stream);
    if (which <= 1) {
      const int iters = 32768, smem = PNSTAGE * PSTAGE + 1024;
      auto k0 = tensor_probe_kernel<0>;
      auto k1 = tensor_probe_kernel<1>;
      cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (which == 0) k0<<<n_sm, 128, smem, ctx->stream>>>(iters);
      else k1<<<n_sm, 128, smem, ctx->stream>>>(iters);
      work = 2.0 * 128 * 128 * (which ? 64 : 32) * (double)iters * n_sm;        // ops
    } else if (which == 2) {
      const int iters = 6000;
      dmma_probe_kernel<<<n_sm, 512, 0, ctx->stream>>>(static_cast<double*>(buf), iters, 1.5);
      work = 2.0 * 16 * 256.0 * iters * 16.0 * n_sm;                            // flops: 16 warps x 16 MMAs x 256 FMA
    } else {
      copy_probe_kernel<<<n_sm * 16, 256, 0, ctx->stream>>>(static_cast<const uint4*>(buf),
                                                            reinterpret_cast<uint4*>(static_cast<uint8_t*>(buf) + copy_bytes),
                                                            (long long)(copy_bytes / 16));
      work = 2.0 * (double)copy_bytes;
    }
    cudaEventRecord(e1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
      ctx->err = std::string("probe kernel: ") + cudaGetErrorString(e);
      if (buf) cudaFree(buf);
      return GB_ERR_CUDA;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best_ms) best_ms = ms;
    ctx->launches++;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (buf) cudaFree(buf);
  *value = work / (best_ms * 1e-3) / (which == 3 ? 1e9 : 1e12);
  return GB_OK;
}
