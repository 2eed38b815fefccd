// gb_gram.cu -- K1: segmented Gram on tcgen05 tensor cores with a fused fp64 fold.
//
// Replaces the reference's SNP-pair loops (dist.cpp:171-179,187-191; distmix.cpp:190-200,209-217;
// computeLD.cpp:106-116) and the per-pair string walks CalCor / CalWgtCov (util.cpp:49-70,103-124).
//
// For a 128 x 128 tile of SNP pairs (A rows x B rows of the packed panel) the kernel streams K
// (individuals) population by population.  Per population p ("segment") a chain of tcgen05.mma
// instructions (kind::i8 for int8 panels, kind::mxf4 with unit block scales for E2M1 panels)
// accumulates the exact counts S^p = sum_k x_ik x_jk in TMEM; the epilogue warps pull the finished
// accumulator into registers and fold it into one fp64 accumulator per matrix entry while the tensor
// core already works on population p+1 (3-4 TMEM accumulator buffers).
//
// Warp roles (384 threads, 1 CTA/SM, persistent over a static tile list):
//   warp 0      TMA producer (one elected lane): 128-byte-swizzled [128 rows x 128 B] boxes of A and B; it also picks
//               the next tile (round-robin, or from a device counter shared with other launches) and publishes the
//               id to the other roles through a 16-slot shared-memory ring
//   warp 1      MMA issuer (one elected lane)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: warp w owns TMEM lanes 32*(w%4).. and columns 64*((w-4)/4)..
//
// The MMA issue loop is the kernel's pacemaker: tools/mma_probe.cu shows that the tensor pipe runs a
// 128x128 MMA in 64 clocks only while tcgen05.mma instructions follow each other back to back --
// anything the issuing thread does between them (barrier polls, constant loads, descriptor
// arithmetic, branches on the operand kind) is NOT hidden behind the previous MMA.  The loop is
// therefore templated on the operand kind, reads its per-segment table from shared memory, keeps the
// descriptor as a running sum and polls the NEXT stage's barrier before issuing the current MMAs.
// (Thread-block clusters with TMA multicast were built and measured in round 1 -- 2x1 ... 4x2, 8x1 --
// and were slower than plain CTAs; see DESIGN.md section 7.  They are gone from the code.)
#include <algorithm>

#include "gb_common.cuh"
#include "gb_ptx.cuh"

namespace gb {

namespace {

constexpr int TILE = 128;
constexpr int STAGES_FUSED = 4;   // smem pipeline depth when the row-statistics tables share shared memory
constexpr int STAGES_RAW = 7;     // ... and when they do not (E2M1 panels: finish pass in gram_finalize_kernel)
constexpr int MAX_STAGES = STAGES_RAW;
constexpr int STAGE_OPERAND_BYTES = TILE * K_BLOCK;  // 16 KiB
constexpr int STAGE_BYTES = 2 * STAGE_OPERAND_BYTES; // A + B
constexpr int ACC_BUFS = 4;                          // 4 x 128 TMEM columns
constexpr int TMEM_COLS = ACC_BUFS * TILE;           // 512
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 128 + EPI_THREADS;           // 384
constexpr int EPI_COLS = 64;                         // columns per epilogue thread
// GramParams::fkind: tensor-core instruction kind the panel format maps to
constexpr int FKIND_I8 = 0;       // kind::i8, int8 rows, K = 32 per instruction
constexpr int FKIND_F8F6F4 = 6;   // kind::f8f6f4 with E2M1 operands (TMA expands nibbles to bytes), K = 32
constexpr int FKIND_MXF4 = 7;     // kind::mxf4 with unit block scales, nibbles stay packed, K = 64

// dynamic shared memory carve-up (byte offsets from a 1024-aligned base)
constexpr int OFF_BARS = 0;                                   // mbarriers
constexpr int N_BARS = 2 * MAX_STAGES + 2 * ACC_BUFS;
constexpr int OFF_TMEM_PTR = OFF_BARS + N_BARS * 8;
constexpr int OFF_SEGTAB = 256;                               // int2 [P_MAX] {first K column, K atoms} per segment
constexpr int OFF_COEFM = OFF_SEGTAB + P_MAX * 8;             // double [P_MAX] coef_p * m_p
constexpr int TILE_RING = 16;                                 // tile ids in flight between the producer and the epilogue
constexpr int OFF_TILE_BAR = OFF_COEFM + P_MAX * 8;           // uint64 [TILE_RING] mbarriers
constexpr int OFF_TILE_ID = OFF_TILE_BAR + TILE_RING * 8;     // int [TILE_RING]
constexpr int OFF_STAGES = 1024;                              // [stages][A 16 KiB | B 16 KiB]
// fused-finish tables sit behind STAGES_FUSED stages
constexpr int OFF_SA = OFF_STAGES + STAGES_FUSED * STAGE_BYTES; // int32 [P_MAX][128]
constexpr int OFF_SB = OFF_SA + P_MAX * TILE * 4;             // int32 [P_MAX][128]
constexpr int OFF_GA = OFF_SB + P_MAX * TILE * 4;             // double [P_MAX][128]  w_p*(s^p_i/m_p)
constexpr int OFF_HB = OFF_GA + P_MAX * TILE * 8;             // double [P_MAX][128]  s^p_j/m_p
constexpr int OFF_AI = OFF_HB + P_MAX * TILE * 8;             // double [128] sum_p w_p s^p_i/m_p (A rows)
constexpr int OFF_BJ = OFF_AI + TILE * 8;                     // double [128] same for B rows
constexpr int OFF_SDA = OFF_BJ + TILE * 8;                    // double [128]
constexpr int OFF_SDB = OFF_SDA + TILE * 8;                   // double [128]
constexpr int SMEM_BYTES_FUSED = OFF_SDB + TILE * 8;
constexpr int SMEM_BYTES_RAW = OFF_STAGES + STAGES_RAW * STAGE_BYTES;
constexpr int SMEM_BYTES = SMEM_BYTES_FUSED > SMEM_BYTES_RAW ? SMEM_BYTES_FUSED : SMEM_BYTES_RAW;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;  // slack for manual 1024-byte alignment
static_assert(SMEM_ALLOC <= 232448, "shared memory budget exceeded");
static_assert(OFF_TMEM_PTR + 16 <= OFF_SEGTAB, "barrier block overlaps the segment table");
static_assert(OFF_TILE_ID + TILE_RING * 4 <= OFF_STAGES, "header tables overlap the stages");

// int32 -> double through the 2^52 trick: one LOP and one exact DADD on the fp64 pipe (64 / clk / SM)
// instead of I2F.F64, which issues at 16 / clk / SM and was the epilogue's limiter.
__device__ __forceinline__ double int_to_double(int v) {
  return __dsub_rn(__hiloint2double(0x43300000, v ^ 0x80000000), 4503601774854144.0);  // 2^52 + 2^31
}
// fp32 accumulator of kind::f8f6f4 holding an exact integer |n| < 2^22 -> int, without F2I:
// the bit pattern of n + 1.5 * 2^23 is that of 1.5 * 2^23 (0x4B400000) plus n.
__device__ __forceinline__ int f32_count_to_int(uint32_t bits) {
  const float f = __uint_as_float(bits) + 12582912.0f;
  return (int)__float_as_uint(f) - 0x4B400000;
}

// fp32 accumulator holding an exact NON-NEGATIVE integer n < 2^21 -> the same value as a double, by
// re-encoding the bits (one LEA.HI; no conversion instruction, no fp64-pipe op): exponent rebias
// 127 -> 1023 in the high word.  An integer below 2^21 has its three lowest fp32 mantissa bits clear, so
// the low word of the double is always zero (per-population counts are <= 4 * 6,360, pooled ones
// 4 * 32,147).  n = 0 (bits 0) maps to 2^-127 instead of 0: absorbed by any nonzero sum, and flushed to
// zero at the store when every count of the entry was zero.
__device__ __forceinline__ double f32_count_to_double(uint32_t bits) {
  return __hiloint2double((int)((bits >> 3) + 0x38000000u), 0);
}
constexpr double RAW_FLUSH = 1e-30;   // |raw| below this can only be a sum of the 2^-127 stand-ins

__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
}

template <int FKIND>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t sf_tmem, uint32_t accumulate) {
  if (FKIND == FKIND_MXF4)
    ptx::mma_mxf4_ss(d_tmem, da, db, ptx::make_idesc_mxf4(TILE, TILE), sf_tmem, sf_tmem, accumulate);
  else if (FKIND == FKIND_F8F6F4)
    ptx::mma_f8f6f4_ss(d_tmem, da, db, ptx::make_idesc_f8f6f4(5, TILE, TILE), accumulate);
  else
    ptx::mma_i8_ss(d_tmem, da, db, ptx::make_idesc_i8(TILE, TILE), accumulate);
}

// tm_*_panel reads the packed panel, tm_*_scratch the gathered rows of non-contiguous windows
// (boxes of {128 B, 128 rows}).
template <int FKIND>
__global__ void __launch_bounds__(THREADS, 1)
gram_seg_kernel(const __grid_constant__ CUtensorMap tm_a_panel,
                const __grid_constant__ CUtensorMap tm_a_scratch,
                const __grid_constant__ CUtensorMap tm_b_panel,
                const __grid_constant__ CUtensorMap tm_b_scratch,
                const __grid_constant__ GramParams prm) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not by integer round trip) so the compiler keeps emitting
  // shared-space loads/stores for everything derived from it
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;
  uint64_t* tempty_bar = bars + 2 * MAX_STAGES + ACC_BUFS;
  // E2M1 panels in mixture mode store the raw weighted Gram sum and leave the finish to
  // gram_finalize_kernel: no statistics tables here, so two more pipeline stages fit
  const bool raw_out = prm.raw_out != 0;
  const int n_stages = raw_out ? STAGES_RAW : STAGES_FUSED;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_TMEM_PTR);
  int2* segtab = reinterpret_cast<int2*>(smem + OFF_SEGTAB);
  double* coefm_s = reinterpret_cast<double*>(smem + OFF_COEFM);
  // The producer decides which tile comes next -- round-robin, or from a global counter when several launches share one
  // tile range (prm.tile_counter) -- and hands the id to the MMA thread and the epilogue warps through this ring.  It
  // can be at most MAX_STAGES tiles ahead of the MMA thread and that at most ACC_BUFS tiles ahead of the epilogue.
  uint64_t* tile_bar = reinterpret_cast<uint64_t*>(smem + OFF_TILE_BAR);
  volatile int* tile_id = reinterpret_cast<volatile int*>(smem + OFF_TILE_ID);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_seg = prm.n_seg;
  const int n_ctas = gridDim.x;

  if (threadIdx.x < n_seg) {
    segtab[threadIdx.x] = make_int2(prm.seg[threadIdx.x].koff, prm.seg[threadIdx.x].natoms);
    coefm_s[threadIdx.x] = prm.coefm[threadIdx.x];
  }
  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tm_a_panel);
    ptx::prefetch_tmap(&tm_a_scratch);
    ptx::prefetch_tmap(&tm_b_panel);
    ptx::prefetch_tmap(&tm_b_scratch);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < MAX_STAGES; s++) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < ACC_BUFS; b++) {
      ptx::mbar_init(&tfull_bar[b], 1);
      ptx::mbar_init(&tempty_bar[b], EPI_WARPS);
    }
    for (int i = 0; i < TILE_RING; i++) ptx::mbar_init(&tile_bar[i], 1);
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

  // kind::mxf4 keeps its (all-ones) block scale factors in the last 128 TMEM columns: 3 accumulator buffers
  constexpr int acc_bufs = FKIND == FKIND_MXF4 ? ACC_BUFS - 1 : ACC_BUFS;
  if (FKIND == FKIND_MXF4) {
    if (warp >= 4 && warp < 8) {
      const uint32_t sf_addr = tmem_base + (uint32_t)((ACC_BUFS - 1) * TILE) + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
      for (int c = 0; c < TILE; c += 16) ptx::tmem_st_fill_32x32b_x16(sf_addr + c, 0x7F7F7F7Fu);  // E8M0 2^0
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
    }
    __syncthreads();
    ptx::tc_fence_after();
  }

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
      // the mbarrier counts bytes as they sit in global memory (nibble-packed E2M1 rows expanded by the
      // TMA unit complete half of what they occupy in shared memory)
      constexpr uint32_t stage_tx = FKIND == FKIND_F8F6F4 ? STAGE_BYTES / 2 : STAGE_BYTES;
      constexpr int kblk = FKIND == FKIND_MXF4 ? 2 * K_BLOCK : K_BLOCK;  // K columns per 128-byte stage row
      int next_static = blockIdx.x;
      for (uint32_t seq = 0;; seq++) {
        int ct;
        if (prm.tile_counter) {
          ct = atomicAdd(prm.tile_counter, 1);
        } else {
          ct = next_static;
          next_static += n_ctas;
        }
        if (ct >= prm.n_tiles) ct = -1;
        tile_id[seq & (TILE_RING - 1)] = ct;
        ptx::mbar_arrive(&tile_bar[seq & (TILE_RING - 1)]);   // release: the id is visible to whoever sees the phase flip
        if (ct < 0) break;
        int2 rows = *reinterpret_cast<const int2*>(&prm.tiles[ct].a_row0);   // a_row0, b_row0
        if (prm.feed_test == 1) rows = make_int2(0, 128);
        if (prm.feed_test == 2) rows = make_int2((int)blockIdx.x * 256, (int)blockIdx.x * 256 + 128);
        const int2 srcs = *reinterpret_cast<const int2*>(&prm.tiles[ct].a_src);    // a_src, b_src
        const CUtensorMap* map_a = srcs.x ? &tm_a_scratch : &tm_a_panel;
        const CUtensorMap* map_b = srcs.y ? &tm_b_scratch : &tm_b_panel;
        for (int s = 0; s < n_seg; s++) {
          const int2 sg = segtab[s];
          const int kend = sg.x + ((sg.y + 3) >> 2) * kblk;
          for (int kcol = sg.x; kcol < kend; kcol += kblk) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);  // the MMAs that read this stage have retired
            uint8_t* sa = smem + OFF_STAGES + stage * STAGE_BYTES;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
            ptx::tma_load_2d(sa, map_a, &full_bar[stage], kcol, rows.x);
            ptx::tma_load_2d(sa + STAGE_OPERAND_BYTES, map_b, &full_bar[stage], kcol, rows.y);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // One thread feeds the tensor core.  Per K block (4 instructions, 256 clocks of tensor work):
    // one non-blocking poll of the NEXT stage's barrier, four tcgen05.mma, one commit.  The smem
    // descriptor of K atom k is the stage descriptor + 2*k (32 bytes >> 4) in its low word.
    if (ptx::elect_one()) {
      const uint32_t sf_tmem = tmem_base + (uint32_t)((ACC_BUFS - 1) * TILE);
      const uint64_t desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + OFF_STAGES));
      const uint64_t desc_end = desc0 + (uint64_t)(n_stages * (STAGE_BYTES >> 4));
      uint64_t da = desc0;
      uint64_t* fullp = full_bar;             // barrier of the current stage
      uint64_t* const full_end = full_bar + n_stages;
      uint32_t phase = 0;
      bool ready = false;                     // the current stage's barrier was already seen complete
      int acc = 0;
      uint32_t acc_phase = 0;
      long long* const dbg = prm.dbg;
      long long w_tempty = 0, w_full = 0, n_tempty_miss = 0, n_full_miss = 0;
      const long long t_begin = dbg ? clock64() : 0;
      bool acc_ready = false;                 // the next accumulator buffer was already seen handed back
      for (uint32_t seq = 0;; seq++) {
        ptx::mbar_wait(&tile_bar[seq & (TILE_RING - 1)], (seq / TILE_RING) & 1);
        if (tile_id[seq & (TILE_RING - 1)] < 0) break;
        for (int s = 0; s < n_seg; s++) {
          int atoms = segtab[s].y;
          if (!acc_ready) {
            if (dbg && !ptx::mbar_test_wait(&tempty_bar[acc], acc_phase ^ 1)) {
              const long long c0 = clock64();
              ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
              w_tempty += clock64() - c0;
              n_tempty_miss++;
            }
            ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          }
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TILE;
          int next_acc = acc + 1;
          uint32_t next_acc_phase = acc_phase;
          if (next_acc == acc_bufs) { next_acc = 0; next_acc_phase ^= 1; }
          uint32_t accumulate = 0;
          for (; atoms > 0; atoms -= 4) {
            if (!ready) {
              if (dbg && !ptx::mbar_test_wait(fullp, phase)) {
                const long long c0 = clock64();
                ptx::mbar_wait(fullp, phase);
                w_full += clock64() - c0;
                n_full_miss++;
              }
              ptx::mbar_wait(fullp, phase);
            }
            ptx::tc_fence_after();
            uint64_t* nextp = fullp + 1;
            uint32_t next_phase = phase;
            if (nextp == full_end) { nextp = full_bar; next_phase ^= 1; }
            // polls whose results are needed only after the MMAs below: the next stage, and -- in the last K block
            // of a segment -- the accumulator buffer the next segment will write
            ready = ptx::mbar_test_wait(nextp, next_phase);
            if (atoms <= 4) acc_ready = ptx::mbar_test_wait(&tempty_bar[next_acc], next_acc_phase ^ 1);
            const uint64_t db = da + (STAGE_OPERAND_BYTES >> 4);
            if (atoms >= 4) {
              mma_ss<FKIND>(d_tmem, da, db, sf_tmem, accumulate);
              mma_ss<FKIND>(d_tmem, da + 2, db + 2, sf_tmem, 1);
              mma_ss<FKIND>(d_tmem, da + 4, db + 4, sf_tmem, 1);
              mma_ss<FKIND>(d_tmem, da + 6, db + 6, sf_tmem, 1);
            } else {
              for (int k = 0; k < atoms; k++) mma_ss<FKIND>(d_tmem, da + 2 * k, db + 2 * k, sf_tmem, accumulate | (uint32_t)k);
            }
            accumulate = 1;
            ptx::mma_commit(fullp + MAX_STAGES);   // empty_bar[stage]: frees the smem stage once these MMAs retire
            fullp = nextp;
            phase = next_phase;
            da += (uint64_t)(STAGE_BYTES >> 4);
            if (da == desc_end) da = desc0;
          }
          ptx::mma_commit(&tfull_bar[acc]);      // accumulator of segment s is complete
          acc = next_acc;
          acc_phase = next_acc_phase;
        }
      }
      if (dbg) {
        dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;
        dbg[blockIdx.x * 8 + 1] = w_tempty;
        dbg[blockIdx.x * 8 + 2] = n_tempty_miss;
        dbg[blockIdx.x * 8 + 3] = w_full;
        dbg[blockIdx.x * 8 + 4] = n_full_miss;
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
    constexpr bool f32acc = FKIND != FKIND_I8;
    // E2M1 panels (non-negative counts in fp32 accumulators) use the regrouped fold
    //   cov_ij = sum_p (coef_p m_p) S^p_ij + sum_p kappa_p s^p_i s^p_j - (sum_p w_p mu^p_i)(sum_p w_p mu^p_j),
    //   kappa_p = w_p / m_p^2 - coef_p,
    // i.e. one DFMA per entry and population while the tile streams and one more in the finish,
    // instead of CalWgtCov's literal term order (kept for int8 panels: 2 IMAD + 3 fp64 ops + the
    // mean term).  Same value up to ~1e-13 of the variance; the test bar is 1e-6.
    const bool fast = raw_out;
    int acc_buf = 0;
    uint32_t acc_phase = 0;

    for (uint32_t seq = 0;; seq++) {
      ptx::mbar_wait(&tile_bar[seq & (TILE_RING - 1)], (seq / TILE_RING) & 1);
      const int ct = tile_id[seq & (TILE_RING - 1)];
      if (ct < 0) break;
      const GramTile t = prm.tiles[ct];
      if (t.a_valid <= 0 || t.b_valid <= 0) {
        // empty descriptor: the MMAs ran, nothing is stored; just hand the accumulators back
        for (int s = 0; s < n_seg; s++) {
          ptx::mbar_wait(&tfull_bar[acc_buf], acc_phase);
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc_buf]);
          if (++acc_buf == acc_bufs) { acc_buf = 0; acc_phase ^= 1; }
        }
        continue;
      }
      // ---- per-tile row statistics into shared memory: plain copies of what row_prep_kernel
      // precomputed per listed row (sum x and sum x / m per population) -- no division here
      if (!fast) epi_bar_sync();  // previous tile's readers are done
      if (mode != GRAM_COUNTS && !fast) {
        const int side = etid >> 7;  // 0: A rows, 1: B rows
        const int idx = etid & 127;
        const int valid = side ? t.b_valid : t.a_valid;
        const long long li = (side ? t.b_list0 : t.a_list0) + min(idx, valid - 1);
        const bool from_u = (side == 0) && t.a_is_u;
        if (mode == GRAM_MIX) {
          const int32_t* st_sx = from_u ? prm.st_sx_u : prm.st_sx_t;
          const double* st_mean = from_u ? prm.st_mean_u : prm.st_mean_t;
          const long long ld = from_u ? prm.st_ld_u : prm.st_ld_t;
          int32_t* sdst = side ? sB : sA;
          double* mdst = side ? hB : gA;
          // CalWgtCov(x, y): wsum_mi_mj += (wgt*(sumx/m))*(sumy/m) with x the FIRST argument.
          // B21 rows call it with x = unmeasured (our A side); B11 / LD call it with x = the
          // smaller SNP index, which in a lower-triangle tile is our B side.  Store the weighted
          // mean on the x side and the plain mean on the y side so the product rounds identically.
          const bool x_side = t.a_is_u ? (side == 0) : (side == 1);
          double wsum = 0.0;
#pragma unroll 4
          for (int p = 0; p < n_seg; p++) {
            const double mean = st_mean[p * ld + li];            // sumx/m       (util.cpp:119)
            sdst[p * TILE + idx] = st_sx[p * ld + li];
            const double wm = __dmul_rn(prm.wgt[p], mean);       // wgt*(sumx/m)
            mdst[p * TILE + idx] = x_side ? wm : mean;
            wsum = __dadd_rn(wsum, wm);                          // wsum_mi += ... (util.cpp:120-121)
          }
          (side ? bjS : aiS)[idx] = wsum;
        } else {  // pooled
          (side ? sB : sA)[idx] = from_u ? prm.pool_u[li] : prm.pool_t[li];
        }
        (side ? sdB : sdA)[idx] = from_u ? prm.sd_u[li] : prm.sd_t[li];
      }
      if (!fast) epi_bar_sync();

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
        // Four 16-column chunks, software pipelined: the tcgen05.ld of chunk ch+1 is in flight
        // while chunk ch is folded (two 16-register staging buffers).
        uint32_t vbuf[2][16];
        ptx::tmem_ld_32x32b_x16(taddr, vbuf[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          uint32_t (&v)[16] = vbuf[ch & 1];
          ptx::tmem_ld_wait();
          if (ch < 3) {
            ptx::tmem_ld_32x32b_x16(taddr + (ch + 1) * 16, vbuf[(ch + 1) & 1]);
          } else {
            // whole accumulator is in registers: hand the TMEM buffer back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc_buf]);
          }
          if (fast) {
            const double coefm = coefm_s[s];
#pragma unroll
            for (int e = 0; e < 16; e++) acc[ch * 16 + e] = fma(coefm, f32_count_to_double(v[e]), acc[ch * 16 + e]);
            continue;
          }
          int cnt[16];
#pragma unroll
          for (int e = 0; e < 16; e++) cnt[e] = f32acc ? f32_count_to_int(v[e]) : (int)v[e];
          if (mode == GRAM_MIX) {
            const int4* sBv = reinterpret_cast<const int4*>(sB + s * TILE + c0 + ch * 16);
            if (prm.wide_fold) {
              // population blocks large enough that m*sumxy - sumx*sumy can leave int32 (m > 23,170 for dosages,
              // m > 364 for arbitrary bytes): both products are exact in fp64 (< 2^49) and so is their fused difference
              const double dm = int_to_double(m), dsA = int_to_double(sAr);
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const int4 b4 = sBv[q];
                const int bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const double d = __fma_rn(dm, int_to_double(cnt[q * 4 + k]), -__dmul_rn(dsA, int_to_double(bb[k])));
                  acc[ch * 16 + q * 4 + k] = __dadd_rn(acc[ch * 16 + q * 4 + k], __dmul_rn(coef, d));
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const int4 b4 = sBv[q];
                const int bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const int d = m * cnt[q * 4 + k] - sAr * bb[k];  // m*sumxy - sumx*sumy, exact (plan-time guard: fits int32)
                  acc[ch * 16 + q * 4 + k] =
                      __dadd_rn(acc[ch * 16 + q * 4 + k], __dmul_rn(coef, int_to_double(d)));
                }
              }
            }
          } else if (mode == GRAM_POOLED) {
#pragma unroll
            for (int e = 0; e < 16; e++) acc[ch * 16 + e] = int_to_double(cnt[e]);
          } else {  // GRAM_COUNTS
            if (r < t.a_valid) {
              int32_t* dst = prm.out_counts + (long long)s * prm.counts_seg_stride + t.out_off +
                             (long long)(t.i0 + r) * t.ld_out + t.j0 + c0 + ch * 16;
#pragma unroll
              for (int e = 0; e < 16; e++)
                if (c0 + ch * 16 + e < t.b_valid) dst[e] = cnt[e];
            }
          }
        }
        if (++acc_buf == acc_bufs) { acc_buf = 0; acc_phase ^= 1; }
      }

      if (mode == GRAM_COUNTS) continue;

      if (fast) {
        // raw sum_p (coef_p m_p) S^p_ij; gram_finalize_kernel adds the mean terms and normalises in place
        double* out = (t.a_is_u ? prm.out_ut : prm.out_tt) + t.out_off;
        const long long gi = t.i0 + r;
        const bool diag_tile = (!t.a_is_u) && (t.i0 == t.j0);
        if (r < t.a_valid) {
#pragma unroll
          for (int e = 0; e < EPI_COLS; e++) {
            const int c = c0 + e;
            const long long gj = t.j0 + c;
            if (c < t.b_valid && !(diag_tile && gi < gj)) out[gj * t.ld_out + gi] = fabs(acc[e]) < RAW_FLUSH ? 0.0 : acc[e];
          }
        }
        continue;
      }

      // ---- finish the entries and store
      const bool row_ok = r < t.a_valid;
      const double sd_r = sdA[r];
      double* out = (t.a_is_u ? prm.out_ut : prm.out_tt) + t.out_off;
      const long long gi = t.i0 + r;
      const bool diag_tile = (!t.a_is_u) && (t.i0 == t.j0);
      const double pooled_n = prm.n_pooled;
      const double pooled_sx = (mode == GRAM_POOLED) ? int_to_double(sA[r]) : 0.0;
      const double ai = (mode == GRAM_MIX) ? aiS[r] : 0.0;
#pragma unroll
      for (int ch = 0; ch < EPI_COLS / 8; ch++) {
        double num8[8];
        if (mode == GRAM_MIX) {
          double x[8];
#pragma unroll
          for (int k = 0; k < 8; k++) x[k] = 0.0;
#pragma unroll 3
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
#pragma unroll
          for (int k = 0; k < 8; k++) {
            // wsumcov + wsum_mi_mj - wsum_mi*wsum_mj   (util.cpp:123)
            num8[k] = __dsub_rn(__dadd_rn(acc[ch * 8 + k], x[k]), __dmul_rn(ai, bjS[c0 + ch * 8 + k]));
          }
        } else {  // pooled Pearson r: num_samples*sumxy - sumx*sumy   (util.cpp:66)
#pragma unroll
          for (int k = 0; k < 8; k++)
            num8[k] = __dsub_rn(__dmul_rn(pooled_n, acc[ch * 8 + k]),
                                __dmul_rn(pooled_sx, int_to_double(sB[c0 + ch * 8 + k])));
        }
        // 8 IEEE divisions, four at a time (a rolled loop would expose the whole ~30-instruction
        // latency chain of each division; eight in flight would spill next to the 64 accumulators)
#pragma unroll
        for (int h4 = 0; h4 < 2; h4++) {
          double cor4[4];
#pragma unroll
          for (int k = 0; k < 4; k++)   // numerator / (stdi*stdj)   (distmix.cpp:196, util.cpp:69)
            cor4[k] = __ddiv_rn(num8[h4 * 4 + k], __dmul_rn(sd_r, sdB[c0 + ch * 8 + h4 * 4 + k]));
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int c = c0 + ch * 8 + h4 * 4 + k;
            const long long gj = t.j0 + c;
            double cor = cor4[k];
            if (diag_tile && gi == gj) cor = prm.diag;
            if (row_ok && c < t.b_valid && !(diag_tile && gi < gj)) {  // diagonal tiles: lower part only
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

// ---- int8-split solve: a B21 tile leaves the finish pass as digit planes ---------------------------------------------
// The solve multiplies B21 by L^-T on the int8 tensor cores (gb_ozaki.cu), so this pass writes the OZ_NDIG = 6 signed
// 8-bit digit planes of a tile's correlations instead of its doubles:
//   q = rint(cor 2^46),  q = sum_p d_p 256^p,  d_p in [-128, 127]     (|cor| <= 1.95; 48-bit fixed point)
// Digits without a carry chain: q + BIAS with BIAS = sum_p 128 256^p has the unsigned base-256 digits d_p + 128, so
// the six low bytes of (q + BIAS) ^ BIAS are the digits.  q comes out of the mantissa of cor 2^46 + 1.5 2^52 (one FMA,
// no 64-bit conversion); the planes of 4 consecutive k are a 4 x 4 byte transpose (PRMT).  The correlation is formed as
// cov (1/sd_i 1/sd_j): a couple of ulps from the reference's division, three orders of magnitude below the 2^-46 the
// digits resolve.  A value the digits cannot carry (NaN: a monomorphic unmeasured SNP has sd = 0) marks its row in
// oz_nan; the solve returns NaN for it, as the doubles would.
// A thread takes 4 rows (unmeasured SNPs) x 8 consecutive k per pass; the planes want k contiguous per row, so 64
// columns at a time go through shared memory ([plane][row][64 B]) and leave as whole 32-byte sectors.
constexpr int OZ_PASS_COLS = 64;  // columns per trip through the staging buffer
constexpr int OZ_STG_ROW = OZ_PASS_COLS + 8;   // bytes per staged row: 8 of padding (8-byte accesses, bank-conflict-free per half-warp)
constexpr size_t OZ_STAGE_BYTES = (size_t)OZ_NDIG * TILE * OZ_STG_ROW;

__device__ __forceinline__ void finalize_oz_tile(const GramParams& prm, const GramTile& t, const double* gA, const double* hB,
                                                 const double* aiS, const double* bjS, const double* isdA, const double* isdB,
                                                 uint8_t* stage) {
  constexpr unsigned long long MAGIC_BITS = 0x4338000000000000ull;       // bits of 1.5 * 2^52
  const int tid = threadIdx.x;
  const int lane = tid & 31, cg = tid >> 5;     // rows lane + 32 j (j < 4); column group cg: 8 consecutive k per pass
  const int n_seg = prm.n_seg;
  const double* out = prm.out_ut + t.out_off + t.i0;
  unsigned row_bad = 0;
#pragma unroll 1
  for (int pass = 0; pass < TILE / OZ_PASS_COLS; pass++) {
    const int cb = pass * OZ_PASS_COLS + cg * 8;    // this thread's 8 columns (measured SNPs k)
    // 4 rows x 8 columns per thread: the rank-P loop below reads 4 + 4 shared-memory vectors per 32 DFMAs (one row per
    // thread took 5 per 8, and the shared-memory pipe, not fp64 or HBM, set this pass's time)
    double x[4][8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const bool col_ok = cb + k < t.b_valid;
      const double* col = out + (long long)(t.j0 + cb + k) * t.ld_out;
#pragma unroll
      for (int j = 0; j < 4; j++) x[j][k] = (col_ok && lane + 32 * j < t.a_valid) ? col[lane + 32 * j] : 0.0;
    }
#pragma unroll 1
    for (int p = 0; p < n_seg; p++) {   // + kappa_p s^p_i s^p_j
      double g[4];
#pragma unroll
      for (int j = 0; j < 4; j++) g[j] = gA[p * TILE + lane + 32 * j];
      const double2* hv = reinterpret_cast<const double2*>(hB + p * TILE + cb);
#pragma unroll
      for (int k2 = 0; k2 < 4; k2++) {
        const double2 h2 = hv[k2];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          x[j][2 * k2] = fma(g[j], h2.x, x[j][2 * k2]);
          x[j][2 * k2 + 1] = fma(g[j], h2.y, x[j][2 * k2 + 1]);
        }
      }
    }
    double bj[8], isd_c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      bj[k] = bjS[cb + k];
      isd_c[k] = isdB[cb + k];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = lane + 32 * j;
      const double ai = aiS[r], isd_r = isdA[r];
      uint32_t lo[8], hi[8];   // bytes of lo: digits 0..3; low bytes of hi: digits 4, 5
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const bool ok = r < t.a_valid && cb + k < t.b_valid;
        double cor = ok ? fma(-ai, bj[k], x[j][k]) * (isd_r * isd_c[k]) : 0.0;
        if (!(fabs(cor) <= 1.95)) {
          row_bad |= 1u << j;
          cor = 0.0;
        }
        const unsigned long long bits = (unsigned long long)__double_as_longlong(fma(cor, 70368744177664.0, 6755399441055744.0));
        const unsigned long long q = (bits + (OZ_BIAS - MAGIC_BITS)) ^ OZ_BIAS;
        lo[k] = (uint32_t)q;
        hi[k] = (uint32_t)(q >> 32);
      }
      uint8_t* srow = stage + (size_t)r * OZ_STG_ROW + cg * 8;
      uint32_t pl[OZ_NDIG][2];
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {   // 4 x 4 byte transposes: plane p of k = 4 hf .. 4 hf + 3
        const uint32_t a = __byte_perm(lo[4 * hf], lo[4 * hf + 1], 0x5140), b = __byte_perm(lo[4 * hf + 2], lo[4 * hf + 3], 0x5140);
        const uint32_t c = __byte_perm(lo[4 * hf], lo[4 * hf + 1], 0x7362), d = __byte_perm(lo[4 * hf + 2], lo[4 * hf + 3], 0x7362);
        const uint32_t e = __byte_perm(hi[4 * hf], hi[4 * hf + 1], 0x5140), f = __byte_perm(hi[4 * hf + 2], hi[4 * hf + 3], 0x5140);
        pl[0][hf] = __byte_perm(a, b, 0x5410);
        pl[1][hf] = __byte_perm(a, b, 0x7632);
        pl[2][hf] = __byte_perm(c, d, 0x5410);
        pl[3][hf] = __byte_perm(c, d, 0x7632);
        pl[4][hf] = __byte_perm(e, f, 0x5410);
        pl[5][hf] = __byte_perm(e, f, 0x7632);
      }
#pragma unroll
      for (int p = 0; p < OZ_NDIG; p++)
        *reinterpret_cast<uint2*>(srow + (size_t)p * TILE * OZ_STG_ROW) = make_uint2(pl[p][0], pl[p][1]);
    }
    __syncthreads();
    // 6 planes x 128 rows x 64 bytes: eight lanes per row, a warp stores eight whole sectors per instruction
#pragma unroll 2
    for (int idx = tid; idx < OZ_NDIG * TILE * (OZ_PASS_COLS / 8); idx += 256) {
      const int q8 = idx & 7, row = (idx >> 3) & (TILE - 1), p = idx >> 10;
      if (row < t.a_valid) {
        const uint2 v = *reinterpret_cast<const uint2*>(stage + ((size_t)p * TILE + row) * OZ_STG_ROW + q8 * 8);
        *reinterpret_cast<uint2*>(prm.oz_pa + ((t.oz_row0 + row) + (long long)p * t.oz_ra) * prm.oz_kpad + t.j0 +
                                  pass * OZ_PASS_COLS + q8 * 8) = v;
      }
    }
    __syncthreads();
  }
  if (row_bad && prm.oz_nan) {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (row_bad >> j & 1) prm.oz_nan[t.a_list0 + lane + 32 * j] = 1;
  }
}

// Finish pass of the regrouped fold (E2M1 panels, mixture mode).  One CTA per 128 x 128 tile of
// the same tile list; reads the raw sum_p (coef_p m_p) S^p_ij the Gram kernel stored and writes
//   cor_ij = (raw + sum_p kappa_p s^p_i s^p_j - a_i b_j) / (sd_i sd_j),   a_i = sum_p w_p s^p_i / m_p,
// in place (diagonal forced, symmetric mirror for computeLD).  A separate, fully occupied kernel
// instead of a tail on the 8 epilogue warps of the tensor-core kernel: there it serialised with
// the next tile's MMAs (3 TMEM buffers ahead at most) and cost 27 % of the kernel.
__global__ void __launch_bounds__(256, 2)
gram_finalize_kernel(const __grid_constant__ GramParams prm) {
  const GramTile t = prm.tiles[blockIdx.x];
  if (t.a_valid <= 0 || t.b_valid <= 0) return;
  extern __shared__ __align__(16) double fs[];
  const int n_seg = prm.n_seg;
  double* gA = fs;                        // [n_seg][128] kappa_p * s^p_i
  double* hB = gA + n_seg * TILE;         // [n_seg][128] s^p_j
  double* aiS = hB + n_seg * TILE;        // [128]
  double* bjS = aiS + TILE;
  double* sdA = bjS + TILE;
  double* sdB = sdA + TILE;
  const int tid = threadIdx.x;
  const bool oz = prm.oz_pa != nullptr && t.a_is_u && t.oz_ra > 0;
  {
    const int side = tid >> 7, idx = tid & 127;
    const int valid = side ? t.b_valid : t.a_valid;
    const long long li = (side ? t.b_list0 : t.a_list0) + min(idx, valid - 1);
    const bool from_u = (side == 0) && t.a_is_u;
    const int32_t* st_sx = from_u ? prm.st_sx_u : prm.st_sx_t;
    const double* st_mean = from_u ? prm.st_mean_u : prm.st_mean_t;
    const long long ld = from_u ? prm.st_ld_u : prm.st_ld_t;
    double* mdst = side ? hB : gA;
    double wsum = 0.0;
#pragma unroll 4
    for (int p = 0; p < n_seg; p++) {
      const double sxd = int_to_double(st_sx[p * ld + li]);
      mdst[p * TILE + idx] = side ? sxd : __dmul_rn(prm.kappa[p], sxd);
      wsum = __dadd_rn(wsum, __dmul_rn(prm.wgt[p], st_mean[p * ld + li]));   // wsum_mi  (util.cpp:120-121)
    }
    (side ? bjS : aiS)[idx] = wsum;
    const double sd = from_u ? prm.sd_u[li] : prm.sd_t[li];
    (side ? sdB : sdA)[idx] = oz ? 1.0 / sd : sd;   // the digit planes take cov * (1/sd_i * 1/sd_j): see finalize_oz_tile
  }
  __syncthreads();
  if (oz) {
    finalize_oz_tile(prm, t, gA, hB, aiS, bjS, sdA, sdB, reinterpret_cast<uint8_t*>(sdB + TILE));
    return;
  }
  const int r = tid & 127;
  const int c0 = (tid >> 7) * EPI_COLS;
  if (r >= t.a_valid) return;
  double* out = (t.a_is_u ? prm.out_ut : prm.out_tt) + t.out_off;
  const long long gi = t.i0 + r;
  const bool diag_tile = (!t.a_is_u) && (t.i0 == t.j0);
  const double ai = aiS[r], sd_r = sdA[r];
#pragma unroll 1
  for (int ch = 0; ch < EPI_COLS / 8; ch++) {
    double x[8];
    bool ok[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int c = c0 + ch * 8 + k;
      const long long gj = t.j0 + c;
      ok[k] = c < t.b_valid && !(diag_tile && gi < gj);
      x[k] = ok[k] ? out[gj * t.ld_out + gi] : 0.0;
    }
#pragma unroll 3
    for (int p = 0; p < n_seg; p++) {   // + kappa_p s^p_i s^p_j
      const double g = gA[p * TILE + r];
      const double2* hv = reinterpret_cast<const double2*>(hB + p * TILE + c0 + ch * 8);
#pragma unroll
      for (int k2 = 0; k2 < 4; k2++) {
        const double2 h2 = hv[k2];
        x[2 * k2] = fma(g, h2.x, x[2 * k2]);
        x[2 * k2 + 1] = fma(g, h2.y, x[2 * k2 + 1]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int c = c0 + ch * 8 + k;
      const long long gj = t.j0 + c;
      double cor = __ddiv_rn(fma(-ai, bjS[c], x[k]), __dmul_rn(sd_r, sdB[c]));   // cov / (stdi*stdj)  (distmix.cpp:196)
      if (diag_tile && gi == gj) cor = prm.diag;
      if (ok[k]) {
        out[gj * t.ld_out + gi] = cor;
        if (prm.mirror) out[gi * t.ld_out + gj] = cor;
      }
    }
  }
}

// prep_zmix5's pair loop (zmix.cpp:151-170): per SNP pair i < j and population p the Pearson r of
// CalCor(std::string&, std::string&) (util.cpp:153-169) from the exact per-population counts the Gram kernel
// wrote in GRAM_COUNTS mode, in the reference's operation order (n*sumxy - sumx*sumy over
// sqrt(n*sumxsq - sumx^2) * sqrt(n*sumysq - sumy^2); every sum is an exact integer held in a double).
// Output column-major [n(n-1)/2][1 + P] like the reference's NumericMatrix: column 0 = z_i z_j.
// Thread (i, j): consecutive j are consecutive output rows, so every column is written coalesced.
__global__ void __launch_bounds__(256)
zmix_pair_kernel(const int32_t* __restrict__ counts, int n, int n_pops, const int* __restrict__ pop_sizes,
                 const int32_t* __restrict__ rows, const int32_t* __restrict__ sx, const int32_t* __restrict__ sxx,
                 long long stat_ld, const double* __restrict__ z, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n || j <= i) return;
  const long long total = (long long)n * (n - 1) / 2;
  const long long row = (long long)i * n - (long long)i * (i + 1) / 2 + (j - i - 1);
  out[row] = __dmul_rn(z[i], z[j]);
  const long long ri = rows[i], rj = rows[j];
  for (int p = 0; p < n_pops; p++) {
    const double m = (double)pop_sizes[p];
    const double sumxy = (double)counts[((long long)p * n + i) * n + j];
    const double sumx = (double)sx[p * stat_ld + ri], sumy = (double)sx[p * stat_ld + rj];
    const double sumxsq = (double)sxx[p * stat_ld + ri], sumysq = (double)sxx[p * stat_ld + rj];
    const double numer = __dsub_rn(__dmul_rn(m, sumxy), __dmul_rn(sumx, sumy));
    const double denor = __dmul_rn(__dsqrt_rn(__dsub_rn(__dmul_rn(m, sumxsq), __dmul_rn(sumx, sumx))),
                                   __dsqrt_rn(__dsub_rn(__dmul_rn(m, sumysq), __dmul_rn(sumy, sumy))));
    out[(long long)(p + 1) * total + row] = __ddiv_rn(numer, denor);
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

int make_row_tensor_map(Ctx* ctx, CUtensorMap* out, const void* base, int64_t n_rows, int64_t k_elems,
                        int64_t k_stride_bytes, int format, int box_rows) {
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
  cuuint64_t dims[2] = {(cuuint64_t)k_elems, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_stride_bytes};
  // every box row is 128 bytes of shared memory: 128 int8 / expanded-E2M1 columns, or 256 packed nibbles
  cuuint32_t box[2] = {(cuuint32_t)(format == MAP_E2M1_PACKED ? 2 * K_BLOCK : K_BLOCK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  // E2M1 rows are nibble-packed in HBM; the TMA unit expands every 16 nibbles (8 B) into a 16-byte
  // shared-memory slot, which is the operand layout kind::f8f6f4 reads, so a K block of 128 dosages
  // occupies the same 128-byte swizzled row as 128 int8 dosages but crosses L2->SM as 64 bytes.
  // (kind::mxf4 reads the nibbles packed: same rows, no expansion, 256 dosages per 128-byte row.)
  const CUtensorMapDataType dt = format == MAP_E2M1_EXPAND   ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B
                                 : format == MAP_E2M1_PACKED ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN8B
                                                             : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->fn_encode_tiled)(
      out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return GB_ERR_CUDA;
  }
  return GB_OK;
}

int make_row_tensor_maps(Ctx* ctx, RowMaps* out, const void* base, int64_t n_rows, int64_t k_elems,
                         int64_t k_stride_bytes, int format) {
  for (int i = 0; i < RowMaps::N; i++) {
    int rc = make_row_tensor_map(ctx, &out->m[i], base, n_rows, k_elems, k_stride_bytes, format, TILE >> i);
    if (rc) return rc;
  }
  return GB_OK;
}

namespace {

template <int FKIND>
int launch_gram_t(Ctx* ctx, const RowMaps& panel, const RowMaps& scratch, const GramParams& prm, int max_ctas) {
  auto kern = gram_seg_kernel<FKIND>;
  static bool attr_set_dev[64] = {};   // function attributes are per device
  bool& attr_set = attr_set_dev[ctx->device & 63];
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
    attr_set = true;
  }
  // persistent: one CTA per SM (229 KB of shared memory each), tiles dealt round-robin
  const int cap = max_ctas > 0 && max_ctas < ctx->sm_count ? max_ctas : ctx->sm_count;
  const int n_ctas = prm.n_tiles < cap ? prm.n_tiles : cap;
  if (getenv("GB_GRAM_TRACE")) {   // diagnostics: where the MMA thread waits (per launch, to stderr)
    static long long* dbg = nullptr;
    if (!dbg) GB_CUDA(cudaMallocManaged(&dbg, 1024 * 8 * sizeof(long long)));
    GramParams p2 = prm;
    p2.dbg = dbg;
    kern<<<(unsigned)n_ctas, THREADS, SMEM_ALLOC, ctx->stream>>>(panel.m[0], scratch.m[0], panel.m[0], scratch.m[0], p2);
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    double sum[8] = {};
    for (int i = 0; i < n_ctas; i++) for (int k = 0; k < 8; k++) sum[k] += (double)dbg[i * 8 + k] / n_ctas;
    fprintf(stderr, "[gram trace] CTAs %d tiles %d segs %d | mma thread total %.0f clk | tempty misses %.0f (%.0f clk) | full misses %.0f (%.0f clk)\n",
            n_ctas, prm.n_tiles, prm.n_seg, sum[0], sum[2], sum[1], sum[4], sum[3]);
    ctx->launches++;
    return GB_OK;
  }
  kern<<<(unsigned)n_ctas, THREADS, SMEM_ALLOC, ctx->stream>>>(panel.m[0], scratch.m[0], panel.m[0], scratch.m[0], prm);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

}  // namespace

int launch_zmix_pairs(Ctx* ctx, const Panel* panel, const int32_t* d_counts, int n, const int32_t* d_rows,
                      const double* d_z, double* d_out) {
  if (n < 2) return GB_OK;
  zmix_pair_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)(n - 1)), 256, 0, ctx->stream>>>(
      d_counts, n, panel->n_pops, panel->d_pop_sizes, d_rows, panel->d_sx, panel->d_sxx, panel->capacity, d_z, d_out);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_gram_finalize(Ctx* ctx, const GramParams& prm, int n_descriptors) {
  if (n_descriptors <= 0) return GB_OK;
  const size_t smem = sizeof(double) * (size_t)(2 * prm.n_seg * TILE + 4 * TILE) + (prm.oz_pa ? OZ_STAGE_BYTES : 0);
  static bool attr_set_dev[64] = {};
  if (!attr_set_dev[ctx->device & 63]) {
    GB_CUDA(cudaFuncSetAttribute(gram_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(double) * (2 * P_MAX * TILE + 4 * TILE) + OZ_STAGE_BYTES)));
    attr_set_dev[ctx->device & 63] = true;
  }
  gram_finalize_kernel<<<(unsigned)n_descriptors, 256, smem, ctx->stream>>>(prm);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_gram(Ctx* ctx, const RowMaps& panel, const RowMaps& scratch, const GramParams& prm, int max_ctas) {
  if (prm.n_tiles <= 0) return GB_OK;
  if (prm.n_seg > P_MAX) {
    ctx->err = "too many population segments for the Gram kernel";
    return GB_ERR_UNSUPPORTED;
  }
  if (prm.raw_out && ctx->seg_order != 0 && prm.n_seg > 2) {
    // The regrouped fold is a plain sum over populations, so its order is free (the int8 fold keeps the
    // reference's).  Largest population first: its long MMA chain covers the epilogue's store phase of the
    // previous tile; then big and small ones alternate so the epilogue catches up behind every long chain.
    GramParams q = prm;
    if (const char* e = getenv("GB_GRAM_FEEDTEST")) q.feed_test = atoi(e);   // diagnostics: timing only
    int idx[P_MAX];
    for (int i = 0; i < prm.n_seg; i++) idx[i] = i;
    std::sort(idx, idx + prm.n_seg, [&](int a, int b) { return prm.seg[a].natoms > prm.seg[b].natoms; });
    int order[P_MAX];
    if (ctx->seg_order == 1) {
      for (int i = 0; i < prm.n_seg; i++) order[i] = idx[i];
    } else {
      int lo = 0, hi = prm.n_seg - 1;
      for (int i = 0; i < prm.n_seg; i++) order[i] = (i & 1) ? idx[hi--] : idx[lo++];
    }
    for (int i = 0; i < prm.n_seg; i++) {
      q.seg[i] = prm.seg[order[i]];
      q.coefm[i] = prm.coefm[order[i]];
    }
    switch (prm.fkind) {
      case FKIND_F8F6F4: return launch_gram_t<FKIND_F8F6F4>(ctx, panel, scratch, q, max_ctas);
      case FKIND_MXF4: return launch_gram_t<FKIND_MXF4>(ctx, panel, scratch, q, max_ctas);
    }
  }
  switch (prm.fkind) {
    case FKIND_I8: return launch_gram_t<FKIND_I8>(ctx, panel, scratch, prm, max_ctas);
    case FKIND_F8F6F4: return launch_gram_t<FKIND_F8F6F4>(ctx, panel, scratch, prm, max_ctas);
    case FKIND_MXF4: return launch_gram_t<FKIND_MXF4>(ctx, panel, scratch, prm, max_ctas);
  }
  ctx->err = "unknown Gram operand kind";
  return GB_ERR_BAD_ARG;
}

}  // namespace gb
