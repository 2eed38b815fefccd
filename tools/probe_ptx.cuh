// probe_ptx.cuh -- thread-block-cluster / multicast wrappers used only by the probes in tools/ (the product kernels run
// plain CTAs: clusters with TMA multicast were measured slower in round 1).
#pragma once
#include "gb_ptx.cuh"

namespace gb {
namespace ptx {

// 2-D tiled load multicast to every CTA of the cluster named in cta_mask: the box lands at the same
// CTA-relative shared-memory offset in each destination and completes bytes on the mbarrier at the
// same offset there.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                               int32_t c0, int32_t c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "h"(cta_mask)
      : "memory");
}

// ---- thread-block clusters --------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Same, but the arrive is multicast to the mbarrier at this offset in every CTA of cta_mask.
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

}  // namespace ptx
}  // namespace gb
