// pair_probe.cu -- TMA-fed tcgen05 pipeline skeleton WITHOUT an epilogue: does a CTA pair (cta_group::2, M = 256,
// each CTA ingesting its own 128 A rows but only HALF of the 128 B rows) feed the tensor pipe better than two
// independent CTAs (each ingesting 128 A + 128 B rows)?  mma_probe.cu shows the pipe itself sustains a 128x128x64
// kind::mxf4 MMA per 64 clocks; the Gram kernel gets half of that, and this probe isolates the L2 -> SM feed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../gauss_b200/csrc -o pair_probe pair_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "probe_ptx.cuh"

using namespace gb;

constexpr int TILE = 128;
constexpr int ROW_BYTES = 128;                   // one swizzled smem row = 256 E2M1 dosages
constexpr int A_BYTES = TILE * ROW_BYTES;        // 16 KiB

__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

template <int CG, int STAGES>
__global__ void __launch_bounds__(128, 1)
pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int n_tiles, int n_kb,
            int n_rows, long long* cyc) {
  constexpr int B_ROWS = TILE / CG;                       // B rows each CTA ingests
  constexpr int STAGE_BYTES = A_BYTES + B_ROWS * ROW_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  uint8_t* stages = smem + 1024;
  const int warp = threadIdx.x >> 5;
  const uint32_t crank = CG == 2 ? ptx::cluster_ctarank() : 0;
  const int pair = blockIdx.x / CG, n_pairs = gridDim.x / CG;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 1) {
      ptx::tmem_alloc(tmem_ptr, 512);
      ptx::tmem_relinquish();
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 0 && ptx::elect_one()) {
    // producer (both CTAs of a pair): own A rows + own half of the B rows; completion on the LEADER's barrier
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < n_tiles; t += n_pairs) {
      const int a_row = ((t * 2 + (int)crank) * TILE) % (n_rows - TILE);
      const int b_row = ((t * 7 + 3) * TILE + (int)crank * B_ROWS) % (n_rows - TILE);
      for (int kb = 0; kb < n_kb; kb++) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = stages + stage * STAGE_BYTES;
        if (CG == 1) {
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          ptx::tma_load_2d(sa, &tm_a, &full_bar[stage], kb * 256, a_row);
          ptx::tma_load_2d(sa + A_BYTES, &tm_b, &full_bar[stage], kb * 256, b_row);
        } else {
          uint32_t leader_bar;   // the same barrier in CTA rank 0 of the pair
          asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(leader_bar) : "r"(ptx::smem_u32(&full_bar[stage])));
          if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          tma_load_2d_cg2(sa, &tm_a, leader_bar, kb * 256, a_row);
          tma_load_2d_cg2(sa + A_BYTES, &tm_b, leader_bar, kb * 256, b_row);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && crank == 0 && ptx::elect_one()) {
    const uint32_t idesc = ptx::make_idesc_mxf4(CG * TILE, TILE);
    const uint64_t desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(stages));
    const uint32_t sf = tmem + 384;
    int stage = 0;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int t = pair; t < n_tiles; t += n_pairs) {
      for (int kb = 0; kb < n_kb; kb++) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint64_t da = desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
        const uint64_t db = da + (A_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (CG == 1) ptx::mma_mxf4_ss(tmem, da + 2 * k, db + 2 * k, idesc, sf, sf, (kb | k) ? 1u : 0u);
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n"
                         ::"r"(tmem), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"((kb | k) ? 1u : 0u), "r"(sf), "r"(sf) : "memory");
        }
        if (CG == 1) ptx::mma_commit(&empty_bar[stage]);
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(ptx::smem_u32(&empty_bar[stage])), "h"((uint16_t)3) : "memory");
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (CG == 1) ptx::mma_commit(done_bar);
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(ptx::smem_u32(done_bar)), "h"((uint16_t)1) : "memory");
    ptx::mbar_wait(done_bar, 0);
    if (blockIdx.x == 0) cyc[0] = clock64() - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CG == 1) ptx::tmem_dealloc(tmem, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* base, long long n_rows, long long k_elems, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)k_elems, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)(k_elems / 2)};
  cuuint32_t box[2] = {256, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN8B, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return m;
}

template <int CG, int STAGES>
void run(EncodeTiledFn enc, void* base, int n_rows, int k_elems, long long* cyc) {
  constexpr int B_ROWS = TILE / CG;
  constexpr int STAGE_BYTES = A_BYTES + B_ROWS * ROW_BYTES;
  const int smem = 1024 + STAGES * STAGE_BYTES + 1024;
  auto kern = pair_kernel<CG, STAGES>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  CUtensorMap ma = make_map(enc, base, n_rows, k_elems, TILE), mb = make_map(enc, base, n_rows, k_elems, B_ROWS);
  const int n_kb = k_elems / 256;              // K blocks per tile
  const int n_tiles = 148 / CG * 24;           // pair tiles (a pair tile = 2 output tiles for CG = 2)
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(148 / CG * CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, kern, ma, mb, n_tiles, n_kb, n_rows, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("cg%d FAILED: %s\n", CG, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double out_tiles = (double)n_tiles * CG;                       // 128 x 128 output tiles
  const double macs = out_tiles * 128.0 * 128.0 * k_elems;
  const double kb_per_sm = (double)n_tiles / (148 / CG) * n_kb;        // K blocks each SM went through
  printf("cta_group::%d stages %d (%3d KB in flight)  %5.3f ms  %7.0f TOP/s  | per SM: %6.1f clk per K block (256 = tensor peak), ingest %5.1f B/clk\n", CG, STAGES, STAGES * STAGE_BYTES / 1024, best,
         2.0 * macs / (best * 1e-3) / 1e12, (double)cyc[0] / kb_per_sm, STAGE_BYTES / ((double)cyc[0] / kb_per_sm));
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  const int n_rows = 8192, k_elems = 32768;          // 8192 rows x 16 KiB = 128 MiB of nibbles (about the L2 size)
  uint8_t* base;
  cudaMalloc(&base, (size_t)n_rows * k_elems / 2);
  cudaMemset(base, 0x22, (size_t)n_rows * k_elems / 2);
  long long* cyc;
  cudaMallocManaged(&cyc, 64);
  run<1, 6>((EncodeTiledFn)fn, base, n_rows, k_elems, cyc);
  run<1, 3>((EncodeTiledFn)fn, base, n_rows, k_elems, cyc);
  run<1, 4>((EncodeTiledFn)fn, base, n_rows, k_elems, cyc);
  run<1, 6>((EncodeTiledFn)fn, base, n_rows, k_elems, cyc);
  run<2, 4>((EncodeTiledFn)fn, base, n_rows, k_elems, cyc);
  run<2, 6>((EncodeTiledFn)fn, base, n_rows, k_elems, cyc);
  run<2, 8>((EncodeTiledFn)fn, base, n_rows, k_elems, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
