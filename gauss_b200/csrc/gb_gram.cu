// gb_gram.cu -- K1: segmented int8 Gram on tcgen05 tensor cores with a fused fp64 epilogue.
//
// Replaces the reference's SNP-pair loops (dist.cpp:171-179,187-191; distmix.cpp:190-200,209-217;
// computeLD.cpp:106-116) and the per-pair string walks CalCor / CalWgtCov (util.cpp:49-70,103-124).
//
// For a 128 x 128 tile of SNP pairs (A rows x B rows of the packed panel) the kernel streams K
// (individuals) population by population.  Per population p ("segment") a chain of
// tcgen05.mma.kind::i8 instructions accumulates the exact int32 counts S^p = sum_k x_ik x_jk in
// TMEM; the epilogue warps pull the finished accumulator into registers and fold it into one
// fp64 accumulator per matrix entry in the reference's operation order:
//     wsumcov += (w_p * m_p/(m_p-1)) * (m_p*S^p - s^p_i*s^p_j)            (util.cpp:117-118)
// while the tensor core already works on population p+1 (4 TMEM accumulator buffers).  After the
// last population the low-rank mean terms, the division by the standard deviations and the
// forced diagonal are applied and the tile is written out.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over a static tile list):
//   warp 0      TMA producer (one elected lane): 128-byte-swizzled [128 rows x 128 B] boxes of A and B
//   warp 1      MMA issuer (one elected lane)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: warp w owns TMEM lanes 32*(w%4).. and columns 64*((w-4)/4)..
#include "gb_common.cuh"
#include "gb_ptx.cuh"

namespace gb {

namespace {

constexpr int TILE = 128;
constexpr int STAGES = 4;
constexpr int STAGE_OPERAND_BYTES = TILE * K_BLOCK;  // 16 KiB
constexpr int STAGE_BYTES = 2 * STAGE_OPERAND_BYTES; // A + B
constexpr int ACC_BUFS = 4;                          // 4 x 128 TMEM columns
constexpr int TMEM_COLS = ACC_BUFS * TILE;           // 512
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 128 + EPI_THREADS;           // 384
constexpr int EPI_COLS = 64;                         // columns per epilogue thread

// dynamic shared memory carve-up (byte offsets from a 1024-aligned base)
constexpr int OFF_STAGES = 0;
constexpr int OFF_SA = OFF_STAGES + STAGES * STAGE_BYTES;     // int32 [P_MAX][128]
constexpr int OFF_SB = OFF_SA + P_MAX * TILE * 4;             // int32 [P_MAX][128]
constexpr int OFF_GA = OFF_SB + P_MAX * TILE * 4;             // double [P_MAX][128]  w_p*(s^p_i/m_p)
constexpr int OFF_HB = OFF_GA + P_MAX * TILE * 8;             // double [P_MAX][128]  s^p_j/m_p
constexpr int OFF_AI = OFF_HB + P_MAX * TILE * 8;             // double [128] sum_p w_p s^p_i/m_p (A rows)
constexpr int OFF_BJ = OFF_AI + TILE * 8;                     // double [128] same for B rows
constexpr int OFF_SDA = OFF_BJ + TILE * 8;                    // double [128]
constexpr int OFF_SDB = OFF_SDA + TILE * 8;                   // double [128]
constexpr int OFF_BARS = OFF_SDB + TILE * 8;                  // mbarriers
constexpr int N_BARS = 2 * STAGES + 2 * ACC_BUFS;
constexpr int OFF_TMEM_PTR = OFF_BARS + N_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;  // slack for manual 1024-byte alignment
static_assert(SMEM_ALLOC <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
gram_seg_i8_kernel(const __grid_constant__ CUtensorMap tm_panel,
                   const __grid_constant__ CUtensorMap tm_scratch,
                   const __grid_constant__ GramParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + ACC_BUFS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tm_panel);
    ptx::prefetch_tmap(&tm_scratch);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < STAGES; s++) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < ACC_BUFS; b++) {
      ptx::mbar_init(&tfull_bar[b], 1);
      ptx::mbar_init(&tempty_bar[b], EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int n_seg = prm.n_seg;

  // Register budget: the epilogue threads carry 64 fp64 accumulators each; the control warps
  // need almost nothing.  64 K regs >= 128 x 56 + 256 x 224.  Each role's code sits entirely
  // inside its own branch so ptxas allocates the two regions against their own budgets.
  if (warp < 4) {
  ptx::setmaxnreg_dec<56>();
  if (warp == 0) {
    // ===================================================================== TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
        const GramTile t = prm.tiles[tile];
        const CUtensorMap* map_a = t.a_src ? &tm_scratch : &tm_panel;
        const CUtensorMap* map_b = t.b_src ? &tm_scratch : &tm_panel;
        for (int s = 0; s < n_seg; s++) {
          const int koff = prm.seg[s].koff;
          const int nblk = (prm.seg[s].natoms + 3) >> 2;
          for (int b = 0; b < nblk; b++) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + OFF_STAGES + stage * STAGE_BYTES;
            uint8_t* sb = sa + STAGE_OPERAND_BYTES;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            ptx::tma_load_2d(sa, map_a, &full_bar[stage], koff + b * K_BLOCK, t.a_row0);
            ptx::tma_load_2d(sb, map_b, &full_bar[stage], koff + b * K_BLOCK, t.b_row0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_i8(TILE, TILE);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t stage_base = ptx::smem_u32(smem + OFF_STAGES);
      for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
        for (int s = 0; s < n_seg; s++) {
          const int natoms = prm.seg[s].natoms;
          const int nblk = (natoms + 3) >> 2;
          ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TILE;
          uint32_t accumulate = 0;
          for (int b = 0; b < nblk; b++) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t a_addr = stage_base + stage * STAGE_BYTES;
            const uint32_t b_addr = a_addr + STAGE_OPERAND_BYTES;
            const int na = min(4, natoms - b * 4);
            for (int k = 0; k < na; k++) {
              const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * K_ATOM);
              const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * K_ATOM);
              ptx::mma_i8_ss(d_tmem, da, db, idesc, accumulate);
              accumulate = 1;
            }
            ptx::mma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          ptx::mma_commit(&tfull_bar[acc]);      // accumulator of segment s is complete
          if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  }
  } else {
    ptx::setmaxnreg_inc<224>();
    // ===================================================================== epilogue warps
    const int ew = warp - 4;
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;         // tile row owned by this thread
    const int c0 = (ew >> 2) * EPI_COLS;    // first tile column owned by this thread
    const int etid = threadIdx.x - 128;
    int32_t* sA = reinterpret_cast<int32_t*>(smem + OFF_SA);
    int32_t* sB = reinterpret_cast<int32_t*>(smem + OFF_SB);
    double* gA = reinterpret_cast<double*>(smem + OFF_GA);
    double* hB = reinterpret_cast<double*>(smem + OFF_HB);
    double* aiS = reinterpret_cast<double*>(smem + OFF_AI);
    double* bjS = reinterpret_cast<double*>(smem + OFF_BJ);
    double* sdA = reinterpret_cast<double*>(smem + OFF_SDA);
    double* sdB = reinterpret_cast<double*>(smem + OFF_SDB);
    const int mode = prm.mode;
    int acc_buf = 0;
    uint32_t acc_phase = 0;

    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      const GramTile t = prm.tiles[tile];
      // ---- per-tile row statistics into shared memory
      epi_bar_sync();  // previous tile's readers are done
      if (mode != GRAM_COUNTS) {
        const int side = etid >> 7;  // 0: A rows, 1: B rows
        const int idx = etid & 127;
        const int valid = side ? t.b_valid : t.a_valid;
        const int li = (side ? t.b_list0 : t.a_list0) + min(idx, valid - 1);
        const bool from_u = (side == 0) && t.a_is_u;
        if (mode == GRAM_MIX) {
          const int prow = from_u ? prm.rows_u[li] : prm.rows_t[li];
          int32_t* sdst = side ? sB : sA;
          double wsum = 0.0;
          for (int p = 0; p < n_seg; p++) {
            const int32_t sx = prm.sx[(long long)p * prm.stat_ld + prow];
            sdst[p * TILE + idx] = sx;
            const double mean = (double)sx / prm.seg[p].m;       // sumx/m       (util.cpp:119)
            const double wm = __dmul_rn(prm.wgt[p], mean);       // wgt*(sumx/m)
            // CalWgtCov(x, y): wsum_mi_mj += (wgt*(sumx/m))*(sumy/m) with x the FIRST argument.
            // B21 rows call it with x = unmeasured (our A side); B11 / LD call it with x = the
            // smaller SNP index, which in a lower-triangle tile is our B side.  Store the weighted
            // mean on the x side and the plain mean on the y side so the product rounds identically.
            const bool x_side = t.a_is_u ? (side == 0) : (side == 1);
            if (side) hB[p * TILE + idx] = x_side ? wm : mean;
            else gA[p * TILE + idx] = x_side ? wm : mean;
            wsum = __dadd_rn(wsum, wm);                          // wsum_mi += ... (util.cpp:120-121)
          }
          (side ? bjS : aiS)[idx] = wsum;
        } else {  // pooled
          (side ? sB : sA)[idx] = from_u ? prm.pool_u[li] : prm.pool_t[li];
        }
        (side ? sdB : sdA)[idx] = from_u ? prm.sd_u[li] : prm.sd_t[li];
      }
      epi_bar_sync();

      double acc[EPI_COLS];
#pragma unroll
      for (int e = 0; e < EPI_COLS; e++) acc[e] = 0.0;

      for (int s = 0; s < n_seg; s++) {
        ptx::mbar_wait(&tfull_bar[acc_buf], acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + acc_buf * TILE + ((uint32_t)(quad * 32) << 16) + c0;
        const int m = prm.seg[s].m;
        const double coef = prm.coef[s];
        const int sAr = (mode == GRAM_MIX) ? sA[s * TILE + r] : 0;
        // four 16-column chunks, one after the other, so only 16 staging registers are live
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(taddr + ch * 16, v);
          ptx::tmem_ld_wait();
          if (ch == 3) {
            // whole accumulator is in registers: hand the TMEM buffer back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc_buf]);
          }
          if (mode == GRAM_MIX) {
            const int4* sBv = reinterpret_cast<const int4*>(sB + s * TILE + c0 + ch * 16);
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const int4 b4 = sBv[q];
              const int bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int d = m * (int)v[q * 4 + k] - sAr * bb[k];  // m*sumxy - sumx*sumy, exact
                acc[ch * 16 + q * 4 + k] =
                    __dadd_rn(acc[ch * 16 + q * 4 + k], __dmul_rn(coef, (double)d));
              }
            }
          } else if (mode == GRAM_POOLED) {
#pragma unroll
            for (int e = 0; e < 16; e++) acc[ch * 16 + e] = (double)(int)v[e];
          } else {  // GRAM_COUNTS
            if (r < t.a_valid) {
              int32_t* dst = prm.out_counts + (long long)s * prm.counts_seg_stride + t.out_off +
                             (long long)(t.i0 + r) * t.ld_out + t.j0 + c0 + ch * 16;
#pragma unroll
              for (int e = 0; e < 16; e++)
                if (c0 + ch * 16 + e < t.b_valid) dst[e] = (int)v[e];
            }
          }
        }
        if (++acc_buf == ACC_BUFS) { acc_buf = 0; acc_phase ^= 1; }
      }

      if (mode == GRAM_COUNTS) continue;

      // ---- finish the entries and store
      const bool row_ok = r < t.a_valid;
      const double sd_r = sdA[r];
      double* out = (t.a_is_u ? prm.out_ut : prm.out_tt) + t.out_off;
      const long long gi = t.i0 + r;
      const bool diag_tile = (!t.a_is_u) && (t.i0 == t.j0);
      if (mode == GRAM_MIX) {
        const double ai = aiS[r];
#pragma unroll
        for (int ch = 0; ch < EPI_COLS / 8; ch++) {
          double x[8];
#pragma unroll
          for (int k = 0; k < 8; k++) x[k] = 0.0;
          for (int p = 0; p < n_seg; p++) {  // wsum_mi_mj += wgt*(sumx/m)*(sumy/m)  (util.cpp:119)
            const double g = gA[p * TILE + r];
            const double2* hv = reinterpret_cast<const double2*>(hB + p * TILE + c0 + ch * 8);
#pragma unroll
            for (int k2 = 0; k2 < 4; k2++) {
              const double2 h2 = hv[k2];
              x[2 * k2] = __dadd_rn(x[2 * k2], __dmul_rn(g, h2.x));
              x[2 * k2 + 1] = __dadd_rn(x[2 * k2 + 1], __dmul_rn(g, h2.y));
            }
          }
          // Finish 8 entries.  Deliberately a rolled loop over a small local array: unrolling
          // would put 8 IEEE divisions in flight on top of the 64 live accumulators and spill.
          double num8[8];
#pragma unroll
          for (int k = 0; k < 8; k++) num8[k] = __dadd_rn(acc[ch * 8 + k], x[k]);  // wsumcov + wsum_mi_mj
#pragma unroll 1
          for (int k = 0; k < 8; k++) {
            const int c = c0 + ch * 8 + k;
            // (wsumcov + wsum_mi_mj - wsum_mi*wsum_mj) / (stdi*stdj)   (util.cpp:123, distmix.cpp:196)
            const double cov = __dsub_rn(num8[k], __dmul_rn(ai, bjS[c]));
            double cor = __ddiv_rn(cov, __dmul_rn(sd_r, sdB[c]));
            const long long gj = t.j0 + c;
            if (diag_tile && gi == gj) cor = prm.diag;
            if (row_ok && c < t.b_valid && !(diag_tile && gi < gj)) {  // diagonal tiles: lower part only
              out[gj * t.ld_out + gi] = cor;
              if (prm.mirror) out[gi * t.ld_out + gj] = cor;
            }
          }
        }
      } else {  // pooled Pearson r (util.cpp:66-69)
        const double n = prm.n_pooled;
        const double sx = (double)sA[r];
#pragma unroll
        for (int ch = 0; ch < EPI_COLS / 8; ch++) {
          double num8[8];
#pragma unroll
          for (int k = 0; k < 8; k++) num8[k] = __dmul_rn(n, acc[ch * 8 + k]);  // num_samples*sumxy
#pragma unroll 1
          for (int k = 0; k < 8; k++) {
            const int c = c0 + ch * 8 + k;
            const double numer = __dsub_rn(num8[k], __dmul_rn(sx, (double)sB[c]));
            double cor = __ddiv_rn(numer, __dmul_rn(sd_r, sdB[c]));
            const long long gj = t.j0 + c;
            if (diag_tile && gi == gj) cor = prm.diag;
            if (row_ok && c < t.b_valid && !(diag_tile && gi < gj)) {
              out[gj * t.ld_out + gi] = cor;
              if (prm.mirror) out[gi * t.ld_out + gj] = cor;
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int make_row_tensor_map(Ctx* ctx, CUtensorMap* out, const void* base, int64_t n_rows, int64_t k_stride) {
  if (!ctx->fn_encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      ctx->err = "cuTensorMapEncodeTiled is not available from this driver";
      return GB_ERR_CUDA;
    }
    ctx->fn_encode_tiled = fn;
  }
  if (n_rows < 1) n_rows = 1;
  cuuint64_t dims[2] = {(cuuint64_t)k_stride, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_stride};
  cuuint32_t box[2] = {(cuuint32_t)K_BLOCK, (cuuint32_t)TILE};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->fn_encode_tiled)(
      out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return GB_ERR_CUDA;
  }
  return GB_OK;
}

int launch_gram(Ctx* ctx, const CUtensorMap& tmap_panel, const CUtensorMap& tmap_scratch,
                const GramParams& prm) {
  if (prm.n_tiles <= 0) return GB_OK;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(gram_seg_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    attr_set = true;
  }
  int grid = prm.n_tiles < ctx->sm_count ? prm.n_tiles : ctx->sm_count;
  gram_seg_i8_kernel<<<grid, THREADS, SMEM_ALLOC, ctx->stream>>>(tmap_panel, tmap_scratch, prm);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

}  // namespace gb
