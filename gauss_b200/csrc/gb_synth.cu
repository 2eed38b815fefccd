// gb_synth.cu -- synthetic reference-panel rows generated ON THE DEVICE (benchmark / test data, not a reference path).
//
// No reference panel is downloadable here (the 33KG panel is a Drive link, docs/articles/ref_33KG.md:7), and the
// genome-wide workload of BASELINE.json config 4 is ~10 M SNPs x 32,953 individuals: 330 GB as chars, far more than a
// host generator can feed.  SURVEY.md section 7 ("HBM residency") therefore asks for on-device generation.  The rows
// come out directly in the ternary host format ("pack5", five dosages per byte, gb_pack5_rows_host), i.e. as the bytes
// a cached packed panel would hold, so the same expand5 -> Gram -> solve path runs on them.
//
// A dosage is a pure function of (seed, chromosome, site, population, individual): any shard of any GPU regenerates
// exactly the rows another one would, which is what makes 1- vs N-GPU results comparable byte for byte.  Model (the
// one gauss_b200/synth.py uses on the host): per site an allele frequency f ~ U(0.01, 0.5), per population
// f_p = clip(f + 0.05 n, 0.005, 0.995); each of an individual's two haplotypes copies the allele of the previous site
// with probability rho ~ U(0.7, 0.95) and draws a fresh one otherwise (haplotype-copy LD, so B11 is realistically
// ill-conditioned); LD chains restart every 64 sites so that a row depends on at most 63 predecessors.
#include "gb_batch.cuh"

namespace gb {

namespace {

constexpr int LD_BLOCK = 64;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ float u01(uint32_t v) { return (float)(v >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {   // murmur3 finaliser: the per-haplotype draw, 2 IMULs
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

// One CTA per row.  Shared tables for the up to 64 sites of the row's LD block: a 32-bit site key, the copy threshold and
// per population the allele-frequency threshold (16-bit fixed point, compared against halves of one 32-bit hash).
__global__ void __launch_bounds__(256)
synth_pack5_rows_kernel(uint8_t* __restrict__ dst, long long dst_stride, const int64_t* __restrict__ sites,
                        long long first_site, int n_pops, const int* __restrict__ pop_sizes,
                        const int* __restrict__ boff5, int row_bytes, uint64_t seed, int chrom) {
  extern __shared__ uint32_t sh[];
  uint32_t* key = sh;                                       // [64] per-site hash key
  uint32_t* thr_copy = sh + LD_BLOCK;                       // [64] copy the previous site's allele if (bits & 0xFFFF) < thr
  uint32_t* thr_f = thr_copy + LD_BLOCK;                    // [64][n_pops] allele 1 if (bits >> 16) < thr
  int* woff = reinterpret_cast<int*>(thr_f + LD_BLOCK * n_pops);   // [n_pops + 1] first 32-bit word of each block
  const long long row = blockIdx.x;
  const long long site = sites ? sites[row] : first_site + row;
  const long long blk0 = site - (site % LD_BLOCK);
  const int depth = (int)(site - blk0);                     // sites blk0 .. site are needed
  const int tid = threadIdx.x;
  const uint64_t base = mix64(seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(chrom + 1)));
  for (int k = tid; k <= depth; k += 256) {
    const uint64_t hk = mix64(base + 0xD1B54A32D192ED03ull * (uint64_t)(blk0 + k + 1));
    key[k] = (uint32_t)(hk >> 32);
    const uint64_t a = mix64(hk ^ 0xA5A5A5A5A5A5A5A5ull);
    const float rho = 0.7f + 0.25f * u01((uint32_t)a);
    thr_copy[k] = k == 0 ? 0u : (uint32_t)(rho * 65536.0f);   // the block's first site always draws fresh
  }
  for (int i = tid; i < (depth + 1) * n_pops; i += 256) {
    const int k = i / n_pops, p = i % n_pops;
    const uint64_t hk = mix64(base + 0xD1B54A32D192ED03ull * (uint64_t)(blk0 + k + 1));
    const uint64_t a = mix64(hk ^ 0xA5A5A5A5A5A5A5A5ull);
    const float f = 0.01f + 0.49f * u01((uint32_t)(a >> 32));
    const uint64_t g = mix64(hk + 0x632BE59BD9B4E019ull * (uint64_t)(p + 1));
    // Irwin-Hall(4) stand-in for a normal deviate: exact integer -> float arithmetic, no transcendental
    const float n = (u01((uint32_t)g) + u01((uint32_t)(g >> 32)) + u01((uint32_t)(g >> 16)) + u01((uint32_t)(g >> 40)) - 2.0f) *
                    1.7320508f;
    const float fp = fminf(fmaxf(f + 0.05f * n, 0.005f), 0.995f);
    thr_f[k * n_pops + p] = (uint32_t)(fp * 65536.0f);
  }
  if (tid <= n_pops) woff[tid] = tid < n_pops ? boff5[tid] >> 2 : row_bytes >> 2;
  __syncthreads();
  uint32_t* out = reinterpret_cast<uint32_t*>(dst + row * dst_stride);
  const int n_words = row_bytes >> 2;
  for (int w = tid; w < n_words; w += 256) {
    int p = 0;
    while (p + 1 < n_pops && woff[p + 1] <= w) p++;
    const int m = pop_sizes[p];
    const int i0 = (w - woff[p]) * 20;                      // 20 dosages per 32-bit word
    uint32_t word = 0;
#pragma unroll 1
    for (int b = 0; b < 4; b++) {
      uint32_t byte = 0, mul = 1;
#pragma unroll 1
      for (int q = 0; q < 5; q++, mul *= 3) {
        const int ind = i0 + 5 * b + q;
        if (ind >= m) break;
        uint32_t dose = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint32_t gid = ((uint32_t)p * 0x01000193u + (uint32_t)ind) * 2u + (uint32_t)h + 1u;
          const uint32_t gmul = gid * 0x9E3779B1u;
          int k = depth;
          for (;;) {
            const uint32_t u = fmix32(key[k] ^ gmul);
            if ((u & 0xFFFFu) >= thr_copy[k]) {             // fresh draw at site k (always at k == 0)
              dose += (u >> 16) < thr_f[k * n_pops + p] ? 1u : 0u;
              break;
            }
            k--;
          }
        }
        byte += mul * dose;
      }
      word |= byte << (8 * b);
    }
    out[w] = word;
  }
}

}  // namespace

// dst: DEVICE memory, n_rows rows of `dst_stride` bytes (>= pack5 row bytes); d_sites: device array of site indices or
// nullptr (site = first_site + r).
int launch_synth_pack5(Ctx* ctx, uint8_t* dst, int64_t dst_stride, int64_t n_rows, const int64_t* d_sites,
                       int64_t first_site, int n_pops, const int* d_pop_sizes, const int* d_boff5, int row_bytes,
                       uint64_t seed, int chrom) {
  if (n_rows <= 0) return GB_OK;
  const size_t smem = sizeof(uint32_t) * (size_t)(2 * LD_BLOCK + LD_BLOCK * n_pops + n_pops + 2);
  for (int64_t r0 = 0; r0 < n_rows; r0 += (1ll << 30)) {
    const int64_t n = std::min<int64_t>(n_rows - r0, 1ll << 30);
    synth_pack5_rows_kernel<<<(unsigned)n, 256, smem, ctx->stream>>>(dst + r0 * dst_stride, dst_stride,
                                                                      d_sites ? d_sites + r0 : nullptr, first_site + r0,
                                                                      n_pops, d_pop_sizes, d_boff5, row_bytes, seed, chrom);
    ctx->launches++;
  }
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

}  // namespace gb

using namespace gb;

// C-ABI: synthetic pack5 rows into HOST (out_is_device == 0) or DEVICE memory.
extern "C" int gb_synth_pack5_rows(gb_ctx* ctx, uint64_t seed, int chrom, int64_t n_rows, const int64_t* sites,
                                   int64_t first_site, int n_pops, const int* pop_sizes, void* out, int64_t out_stride,
                                   int out_is_device) {
  if (!ctx || n_rows < 0 || n_pops < 1 || n_pops > P_MAX || !pop_sizes || (n_rows && !out)) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  std::vector<int> boff;
  const int rb = pack5_layout(n_pops, pop_sizes, &boff);
  if (rb < 0 || out_stride < rb) {
    ctx->err = "row stride below gb_pack5_row_bytes()";
    return GB_ERR_BAD_ARG;
  }
  if (n_rows == 0) return GB_OK;
  GB_CUDA(cudaSetDevice(ctx->device));
  int *d_ps = nullptr, *d_bo = nullptr;
  int64_t* d_sites = nullptr;
  uint8_t* d_out = nullptr;
  int rc = GB_OK;
  auto done = [&](int code) {
    if (d_ps) cudaFreeAsync(d_ps, ctx->stream);
    if (d_bo) cudaFreeAsync(d_bo, ctx->stream);
    if (d_sites) cudaFreeAsync(d_sites, ctx->stream);
    if (d_out && !out_is_device) cudaFreeAsync(d_out, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return code;
  };
  if (cudaMallocAsync(reinterpret_cast<void**>(&d_ps), sizeof(int) * (size_t)n_pops, ctx->stream) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void**>(&d_bo), sizeof(int) * (size_t)n_pops, ctx->stream) != cudaSuccess ||
      (sites && cudaMallocAsync(reinterpret_cast<void**>(&d_sites), sizeof(int64_t) * (size_t)n_rows, ctx->stream) != cudaSuccess)) {
    ctx->err = "cudaMallocAsync(synth) failed";
    cudaGetLastError();
    return done(GB_ERR_OOM);
  }
  cudaMemcpyAsync(d_ps, pop_sizes, sizeof(int) * (size_t)n_pops, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_bo, boff.data(), sizeof(int) * (size_t)n_pops, cudaMemcpyHostToDevice, ctx->stream);
  if (sites) cudaMemcpyAsync(d_sites, sites, sizeof(int64_t) * (size_t)n_rows, cudaMemcpyHostToDevice, ctx->stream);
  if (out_is_device) {
    d_out = static_cast<uint8_t*>(out);
  } else if (cudaMallocAsync(reinterpret_cast<void**>(&d_out), (size_t)n_rows * (size_t)out_stride, ctx->stream) != cudaSuccess) {
    ctx->err = "cudaMallocAsync(synth rows) failed";
    cudaGetLastError();
    d_out = nullptr;
    return done(GB_ERR_OOM);
  }
  if (out_stride > rb) cudaMemsetAsync(d_out, 0, (size_t)n_rows * (size_t)out_stride, ctx->stream);
  rc = launch_synth_pack5(ctx, d_out, out_stride, n_rows, d_sites, first_site, n_pops, d_ps, d_bo, rb, seed, chrom);
  if (!rc && !out_is_device &&
      cudaMemcpyAsync(out, d_out, (size_t)n_rows * (size_t)out_stride, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
    ctx->err = "synthetic rows: device -> host copy failed";
    rc = GB_ERR_CUDA;
  }
  return done(rc);
}
