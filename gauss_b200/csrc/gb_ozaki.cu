// gb_ozaki.cu -- the solve's n_t^2 n_u term on the int8 tensor cores (tcgen05.mma kind::i8), exact-integer split.
//
// fp64 is the weak pipe of this chip (36.9 TFLOP/s measured against 4,518 int8 TOP/s), and the triangular solve
// W = L^-1 B21^T is the longest kernel of a step.  This path replaces it by a GEMM without the trsm's sequential
// dependency and moves the GEMM to the tensor cores:
//   X  = L^-1                 (explicit: the blocked trsm kernel on identity columns, zero blocks skipped; + n_t^3 / 3)
//   W^T = B21 X^T             (n_u x n_t, only ||W_col||^2 and W_col . y are needed: W is never stored)
// Both operands are split into NDIG = 6 signed 8-bit digits of a 48-bit fixed-point value (x = 2^e sum_s d_s 2^(8s - 46),
// d_s in [-128, 127]); every digit-pair product sum_k dB_s[u, k] dX_t[r, k] is an exact int32 (|.| <= 128^2 K per pair,
// 6 pairs and K <= 2048 at most: 2^27.6); pairs of equal weight s + t share one TMEM accumulator; the NG = 7 heaviest
// weight groups (s + t >= 4: 26 of the 36 pairs) are recombined in fp64 (2^(8 (s + t) - 92 + e_X) each).  Quantisation
// (2^-46 of the scale per entry) and the dropped groups (< 2^-52 of a product term) leave |dz| ~ 2e-12 on a 750-SNP
// window (simulated against extended precision): the 1e-6 bar with five orders to spare and tighter than the fp64 path's
// own distance to the LU-based CPU restatement.  (Seven 7-bit digits with 28 pairs, the first version, measured 4.6e-11 and cost a
// plane and two pairs more; s + t >= 5, 21 pairs, gives 7.5e-11.)  Balanced base-256 digits need no carry chain: with
// BIAS = sum_p 128 256^p the bytes of (q + BIAS) ^ 0x80..80 ARE the digits.  L^-1 is lower triangular, so the K range of
// the 128-row tile J of X ends at 128 (J + 1).
//
// Kernel = the Gram kernel's skeleton: warp 0 TMA producer (128-byte-swizzled [128 rows x 128 B] boxes, 6 stages),
// warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue; work item = (window, 128 unmeasured SNPs), inner loop
// over the X row tiles J, so info_u and z_u accumulate in two registers per thread and are written once.
#include <algorithm>
#include <cstring>

#include <queue>

#include "gb_batch.cuh"
#include "gb_ptx.cuh"

namespace gb {

namespace {

constexpr int NDIG = OZ_NDIG;       // signed 8-bit digits per operand (6: a 48-bit fixed-point value)
constexpr int DBITS = 8;
constexpr int GMIN = 4;             // lightest digit-weight group kept (s + t >= GMIN)
constexpr int NG = 2 * (NDIG - 1) - GMIN + 1;   // 7 groups: s + t = 10 .. 4, 26 digit pairs
constexpr int QBITS = OZ_QBITS;     // value = 2^e sum_s d_s 2^(8 s - QBITS), |value| 2^-e < 1
constexpr int OZ_STAGES = 6;
constexpr int OZ_TILE = 128;
constexpr int OZ_STAGE_OPERAND = OZ_TILE * 128;      // 16 KiB
constexpr int OZ_STAGE_BYTES = 2 * OZ_STAGE_OPERAND;
constexpr int OZ_ACC_BUFS = 4;
constexpr int OZ_TMEM_COLS = OZ_ACC_BUFS * OZ_TILE;
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_EPI_THREADS = OZ_EPI_WARPS * 32;
constexpr int OZ_THREADS = 128 + OZ_EPI_THREADS;
constexpr int OZ_OFF_BARS = 0;
constexpr int OZ_N_BARS = 2 * OZ_STAGES + 2 * OZ_ACC_BUFS;
constexpr int OZ_OFF_TMEM_PTR = OZ_OFF_BARS + OZ_N_BARS * 8;
constexpr int OZ_OFF_RED = 256;                              // double [2][128] cross-half reduction
constexpr int OZ_OFF_STAGES = 4096;
constexpr int OZ_Y_MAX = 2048;                               // measured SNPs per window the smem copy of y holds
constexpr int OZ_OFF_Y = OZ_OFF_STAGES + OZ_STAGES * OZ_STAGE_BYTES;
constexpr int OZ_SMEM = OZ_OFF_Y + OZ_Y_MAX * 8;
constexpr int OZ_SMEM_ALLOC = OZ_SMEM + 1024;
static_assert(OZ_SMEM_ALLOC <= 232448, "shared memory budget exceeded");
static_assert(NDIG == 6 && QBITS == 46, "oz_digits and the finish pass read the digits as the six low bytes of a 64-bit word");
static_assert(OZ_OFF_RED + 2 * 128 * 8 <= OZ_OFF_STAGES, "header overlaps the stages");

struct OzWin {
  int n_t, n_u;
  int nbt, nbu;            // 128-row tiles of X rows / of unmeasured rows
  long long a_row0;        // first row of the window's B21 digit planes in the A tensor map (plane p: + p * ra)
  long long b_row0;        // ... of its X digit planes in the B tensor map (plane p: + p * rb)
  int ra, rb;              // rows per plane
  long long off_t, off_u;  // into y / into z_u, info_u
};
struct OzTile {
  int win, ut;
};

__device__ __forceinline__ double oz_int_to_double(int v) {
  return __dsub_rn(__hiloint2double(0x43300000, v ^ 0x80000000), 4503601774854144.0);  // 2^52 + 2^31
}
__device__ __forceinline__ void oz_epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_THREADS) : "memory"); }

// digit pairs of group gi (weight g' = 2 (NDIG - 1) - gi): s from max(0, g' - (NDIG - 1)) to min(NDIG - 1, g')
__device__ __forceinline__ void oz_group_range(int gi, int& gw, int& s_lo, int& s_hi) {
  gw = 2 * (NDIG - 1) - gi;
  s_lo = gw - (NDIG - 1) > 0 ? gw - (NDIG - 1) : 0;
  s_hi = gw < NDIG - 1 ? gw : NDIG - 1;
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_solve_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const OzWin* __restrict__ wins, const OzTile* __restrict__ tiles, int n_tiles,
                   const int* __restrict__ ex, const double* __restrict__ y, const uint8_t* __restrict__ nanflag,
                   double* __restrict__ zu, double* __restrict__ info) {
  extern __shared__ uint8_t oz_smem_raw[];
  uint8_t* smem = oz_smem_raw + ((1024u - (ptx::smem_u32(oz_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OZ_OFF_BARS);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + OZ_STAGES;
  uint64_t* tfull_bar = bars + 2 * OZ_STAGES;
  uint64_t* tempty_bar = bars + 2 * OZ_STAGES + OZ_ACC_BUFS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OZ_OFF_TMEM_PTR);
  double* red = reinterpret_cast<double*>(smem + OZ_OFF_RED);
  double* ys = reinterpret_cast<double*>(smem + OZ_OFF_Y);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_b);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < OZ_STAGES; s++) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < OZ_ACC_BUFS; b++) {
      ptx::mbar_init(&tfull_bar[b], 1);
      ptx::mbar_init(&tempty_bar[b], OZ_EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_smem, OZ_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4) {
    ptx::setmaxnreg_dec<56>();
    if (warp == 0) {
      // ===================================================================== TMA producer
      if (ptx::elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int ct = blockIdx.x; ct < n_tiles; ct += gridDim.x) {
          const OzTile t = tiles[ct];
          if (t.win < 0) continue;   // unused slot of the longest-first deal
          const OzWin w = wins[t.win];
          const int a_row = (int)(w.a_row0 + (long long)t.ut * OZ_TILE);
          for (int J = 0; J < w.nbt; J++) {
            const int b_row = (int)(w.b_row0 + (long long)J * OZ_TILE);
            for (int gi = 0; gi < NG; gi++) {
              int gw, s_lo, s_hi;
              oz_group_range(gi, gw, s_lo, s_hi);
              for (int s = s_lo; s <= s_hi; s++) {
                const int ra = a_row + s * w.ra, rb = b_row + (gw - s) * w.rb;
                for (int kb = 0; kb <= J; kb++) {
                  ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                  uint8_t* sa = smem + OZ_OFF_STAGES + stage * OZ_STAGE_BYTES;
                  ptx::mbar_arrive_expect_tx(&full_bar[stage], OZ_STAGE_BYTES);
                  ptx::tma_load_2d(sa, &tm_a, &full_bar[stage], kb * 128, ra);
                  ptx::tma_load_2d(sa + OZ_STAGE_OPERAND, &tm_b, &full_bar[stage], kb * 128, rb);
                  if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
                }
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================================================================== MMA issuer
      if (ptx::elect_one()) {
        const uint32_t idesc = ptx::make_idesc_i8(OZ_TILE, OZ_TILE);
        const uint64_t desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + OZ_OFF_STAGES));
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int ct = blockIdx.x; ct < n_tiles; ct += gridDim.x) {
          if (tiles[ct].win < 0) continue;
          const OzWin w = wins[tiles[ct].win];
          for (int J = 0; J < w.nbt; J++) {
            for (int gi = 0; gi < NG; gi++) {
              int gw, s_lo, s_hi;
              oz_group_range(gi, gw, s_lo, s_hi);
              ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
              ptx::tc_fence_after();
              const uint32_t d_tmem = tmem_base + acc * OZ_TILE;
              uint32_t accumulate = 0;
              const int n_blocks = (s_hi - s_lo + 1) * (J + 1);
              for (int b = 0; b < n_blocks; b++) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint64_t da = desc0 + (uint64_t)(stage * (OZ_STAGE_BYTES >> 4));
                const uint64_t db = da + (OZ_STAGE_OPERAND >> 4);
                ptx::mma_i8_ss(d_tmem, da, db, idesc, accumulate);
                ptx::mma_i8_ss(d_tmem, da + 2, db + 2, idesc, 1);
                ptx::mma_i8_ss(d_tmem, da + 4, db + 4, idesc, 1);
                ptx::mma_i8_ss(d_tmem, da + 6, db + 6, idesc, 1);
                accumulate = 1;
                ptx::mma_commit(&empty_bar[stage]);
                if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
              }
              ptx::mma_commit(&tfull_bar[acc]);
              if (++acc == OZ_ACC_BUFS) { acc = 0; acc_phase ^= 1; }
            }
          }
        }
      }
    }
  } else {
    ptx::setmaxnreg_inc<224>();
    // ===================================================================== epilogue warps
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // row of the tile = unmeasured SNP
    const int half = ew >> 2;                // which 64 columns (X rows) this thread owns
    const int c0 = half * 64;
    const int etid = threadIdx.x - 128;
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    for (int ct = blockIdx.x; ct < n_tiles; ct += gridDim.x) {
      const OzTile t = tiles[ct];
      if (t.win < 0) continue;
      const OzWin w = wins[t.win];
      const int e_x = ex[t.win];
      oz_epi_bar_sync();   // the previous tile's readers of ys / red are done
      for (int i = etid; i < w.nbt * OZ_TILE; i += OZ_EPI_THREADS) ys[i] = i < w.n_t ? y[w.off_t + i] : 0.0;
      oz_epi_bar_sync();
      double p_info = 0.0, p_z = 0.0;
      for (int J = 0; J < w.nbt; J++) {
        double a[64];
#pragma unroll
        for (int e = 0; e < 64; e++) a[e] = 0.0;
        for (int gi = 0; gi < NG; gi++) {
          const int gw = 2 * (NDIG - 1) - gi;
          const double sc = __hiloint2double((1023 + DBITS * gw - 2 * QBITS + e_x) << 20, 0);   // 2^(8 g' - 2 QBITS + e_X)
          ptx::mbar_wait(&tfull_bar[acc_buf], acc_phase);
          ptx::tc_fence_after();
          const uint32_t taddr = tmem_base + acc_buf * OZ_TILE + ((uint32_t)(quad * 32) << 16) + c0;
          uint32_t vbuf[2][16];
          ptx::tmem_ld_32x32b_x16(taddr, vbuf[0]);
#pragma unroll
          for (int ch = 0; ch < 4; ch++) {
            uint32_t (&v)[16] = vbuf[ch & 1];
            ptx::tmem_ld_wait();
            if (ch < 3) {
              ptx::tmem_ld_32x32b_x16(taddr + (ch + 1) * 16, vbuf[(ch + 1) & 1]);
            } else {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc_buf]);
            }
#pragma unroll
            for (int e = 0; e < 16; e++) a[ch * 16 + e] = fma(sc, oz_int_to_double((int)v[e]), a[ch * 16 + e]);
          }
          if (++acc_buf == OZ_ACC_BUFS) { acc_buf = 0; acc_phase ^= 1; }
        }
        const double* yj = ys + J * OZ_TILE + c0;
#pragma unroll
        for (int e = 0; e < 64; e++) {
          p_info = fma(a[e], a[e], p_info);
          p_z = fma(a[e], yj[e], p_z);
        }
      }
      // the two column halves of a row, in a fixed order
      if (half == 1) {
        red[r] = p_info;
        red[128 + r] = p_z;
      }
      oz_epi_bar_sync();
      if (half == 0) {
        const int u = t.ut * OZ_TILE + r;
        if (u < w.n_u) {
          const double s_info = p_info + red[r], s_z = p_z + red[128 + r];
          // a row of B21 the digit planes could not carry (NaN: sd = 0) gives NaN, as its doubles would
          const double inf = nanflag[w.off_u + u] ? __longlong_as_double(0x7ff8000000000000ll) : fabs(s_info);   // info = |b21 B11^-1 b12|  (dist.cpp:198)
          zu[w.off_u + u] = s_z / sqrt(inf);           // z / sqrt(info)               (dist.cpp:200)
          info[w.off_u + u] = inf;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, OZ_TMEM_COLS);
  }
}

// ---- version 2: the A planes of a K block stay in shared memory across the digit pairs ----------------------------
// Version 1 above streams a fresh [A plane | X plane] pair of tiles (32 KB) for every 128 x 128 x 128 MMA block, and the
// L2 -> SM path, not the tensor pipe, sets its time (imma pipe 70 %, L2 79 %).  Here a work unit is (window, 128 unmeasured
// SNPs, 64 rows of X): for one K block the six A-plane tiles (16 KB each) and the six X-plane tiles (64 rows: 8 KB each)
// feed all 26 digit pairs -- 144 KB for 13 MMA-block equivalents instead of 416 KB -- because the seven weight groups
// now accumulate at the same time: seven 64-column TMEM accumulators (+ one spare of the 512 columns).
//   producer: two FIFO rings (8 A slots, 8 X slots); order per K block  A0 X5 A1..A5 X4..X0
//   MMA     : bursts t = 5..0 over the X planes, inner s (A planes) ascending; A_s is released after its last burst
//             (s = 0 after t = 4, ... s = 4, 5 after t = 0: the order they were loaded in), X_t after its burst
//   groups  : weight 10 - gi; in the last K block gi = 0, 1, 2, 3, 4, 6, 5 complete in that order, and the next unit
//             first needs gi = 5, 4, 3, 2, 1, 0, 6: it takes the spare and then the slots in their order of completion,
//             so the epilogue drains an accumulator while the MMAs of the next unit already run in another
// MEASURED (chr22 batch, results equal to fp64 rounding): 1.23 ms against 1.07 ms for version 1 -- a loss, kept as the record
// (GB_OZ_KERNEL=2).  With N = 64 a tcgen05.mma reads 128 A rows from shared memory for half the work: 6 KB per 32-clock
// instruction is 192 B/clk against the ~128 B/clk shared memory delivers, so the pipe is paced by operand READS at ~48 clk
// per instruction (1.5x), and one thread has to issue an MMA every 32 clocks (the first version of the loop, with modulo
// slot arithmetic, ran 1.65 ms).  The seven groups cannot be live at N = 128 (896 of 512 TMEM columns).
constexpr int V2_NX = 64;                 // X rows per unit = MMA N
constexpr int V2_A_SLOTS = 8, V2_A_BYTES = OZ_TILE * 128;    // powers of two: the issuing thread's slot arithmetic is masks
constexpr int V2_B_SLOTS = 8, V2_B_BYTES = V2_NX * 128;
constexpr int V2_ACC = 8;                 // 64-column accumulators
constexpr int V2_OFF_AFULL = 0, V2_OFF_AEMPTY = V2_OFF_AFULL + 8 * V2_A_SLOTS, V2_OFF_BFULL = V2_OFF_AEMPTY + 8 * V2_A_SLOTS;
constexpr int V2_OFF_BEMPTY = V2_OFF_BFULL + 8 * V2_B_SLOTS, V2_OFF_TFULL = V2_OFF_BEMPTY + 8 * V2_B_SLOTS;
constexpr int V2_OFF_TEMPTY = V2_OFF_TFULL + 8 * V2_ACC, V2_OFF_TMEM_PTR = V2_OFF_TEMPTY + 8 * V2_ACC;
constexpr int V2_OFF_RED = 512;           // double [2][128]
constexpr int V2_OFF_Y = V2_OFF_RED + 2 * 128 * 8;   // double [2][64]
constexpr int V2_OFF_A = 4096;
constexpr int V2_OFF_B = V2_OFF_A + V2_A_SLOTS * V2_A_BYTES;
constexpr int V2_SMEM_ALLOC = V2_OFF_B + V2_B_SLOTS * V2_B_BYTES + 1024;
static_assert(V2_OFF_TMEM_PTR + 8 <= V2_OFF_RED && V2_OFF_Y + 2 * 64 * 8 <= V2_OFF_A, "v2 header layout");
static_assert(V2_SMEM_ALLOC <= 232448, "shared memory budget exceeded");
static_assert(NDIG == 6 && GMIN == 4, "the burst schedule below is written for six planes and s + t >= 4");

// accumulator slots of the seven groups, 4 bits each (bits 28..31: the spare)
__device__ __forceinline__ uint32_t v2_rotate(uint32_t m) {
  auto get = [&](int i) { return (m >> (4 * i)) & 15u; };
  // need order gi = 5, 4, 3, 2, 1, 0, 6  <-  spare, old[0], old[1], old[2], old[3], old[4], old[6]; new spare = old[5]
  return (get(4) << 0) | (get(3) << 4) | (get(2) << 8) | (get(1) << 12) | (get(0) << 16) | (get(7) << 20) | (get(6) << 24) |
         (get(5) << 28);
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_solve_kernel_v2(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const OzWin* __restrict__ wins, const OzTile* __restrict__ tiles, int n_tiles,
                      const int* __restrict__ ex, const double* __restrict__ y, const uint8_t* __restrict__ nanflag,
                      double* __restrict__ zu, double* __restrict__ info) {
  extern __shared__ uint8_t oz_smem_raw[];
  uint8_t* smem = oz_smem_raw + ((1024u - (ptx::smem_u32(oz_smem_raw) & 1023u)) & 1023u);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + V2_OFF_AFULL);
  uint64_t* a_empty = reinterpret_cast<uint64_t*>(smem + V2_OFF_AEMPTY);
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + V2_OFF_BFULL);
  uint64_t* b_empty = reinterpret_cast<uint64_t*>(smem + V2_OFF_BEMPTY);
  uint64_t* tfull = reinterpret_cast<uint64_t*>(smem + V2_OFF_TFULL);
  uint64_t* tempty = reinterpret_cast<uint64_t*>(smem + V2_OFF_TEMPTY);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + V2_OFF_TMEM_PTR);
  double* red = reinterpret_cast<double*>(smem + V2_OFF_RED);
  double* ys = reinterpret_cast<double*>(smem + V2_OFF_Y);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_b);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int i = 0; i < V2_A_SLOTS; i++) {
      ptx::mbar_init(&a_full[i], 1);
      ptx::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < V2_B_SLOTS; i++) {
      ptx::mbar_init(&b_full[i], 1);
      ptx::mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < V2_ACC; i++) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], OZ_EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_smem, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t SLOTS0 = 0x76543210u;   // group gi -> slot gi, spare 7

  if (warp < 4) {   // (no setmaxnreg here: 384 threads x 168 registers fit, and the unrolled issue loop wants more than 56)
    if (warp == 0) {
      // ===================================================================== TMA producer
      if (ptx::elect_one()) {
        uint32_t a_cnt = 0, b_cnt = 0;
        for (int ct = blockIdx.x; ct < n_tiles; ct += gridDim.x) {
          const OzTile t = tiles[ct];
          if (t.win < 0) continue;
          const OzWin w = wins[t.win];
          const int a_row = (int)(w.a_row0 + (long long)t.ut * OZ_TILE);
          const int nbh = (w.n_t + V2_NX - 1) / V2_NX;
          for (int jh = 0; jh < nbh; jh++) {
            const int b_row = (int)(w.b_row0 + (long long)jh * V2_NX);
            const int n_kb = jh / 2 + 1;
            for (int kb = 0; kb < n_kb; kb++) {
              auto load_a = [&](int s) {
                const uint32_t slot = a_cnt & (V2_A_SLOTS - 1), par = (a_cnt / V2_A_SLOTS) & 1u;
                ptx::mbar_wait(&a_empty[slot], par ^ 1);
                ptx::mbar_arrive_expect_tx(&a_full[slot], V2_A_BYTES);
                ptx::tma_load_2d(smem + V2_OFF_A + slot * V2_A_BYTES, &tm_a, &a_full[slot], kb * 128, a_row + s * w.ra);
                a_cnt++;
              };
              auto load_b = [&](int tt) {
                const uint32_t slot = b_cnt & (V2_B_SLOTS - 1), par = (b_cnt / V2_B_SLOTS) & 1u;
                ptx::mbar_wait(&b_empty[slot], par ^ 1);
                ptx::mbar_arrive_expect_tx(&b_full[slot], V2_B_BYTES);
                ptx::tma_load_2d(smem + V2_OFF_B + slot * V2_B_BYTES, &tm_b, &b_full[slot], kb * 128, b_row + tt * w.rb);
                b_cnt++;
              };
              load_a(0);
              load_b(5);
              for (int s = 1; s < NDIG; s++) load_a(s);
              for (int tt = 4; tt >= 0; tt--) load_b(tt);
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================================================================== MMA issuer
      if (ptx::elect_one()) {
        const uint32_t idesc = ptx::make_idesc_i8(OZ_TILE, V2_NX);
        const uint64_t desc_a0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + V2_OFF_A));
        const uint64_t desc_b0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + V2_OFF_B));
        uint32_t a_cnt = 0, b_cnt = 0, slots = SLOTS0, acc_par = 0;   // acc_par bit = parity of the slot's CURRENT use
        bool first = true;
        for (int ct = blockIdx.x; ct < n_tiles; ct += gridDim.x) {
          if (tiles[ct].win < 0) continue;
          const OzWin w = wins[tiles[ct].win];
          const int nbh = (w.n_t + V2_NX - 1) / V2_NX;
          for (int jh = 0; jh < nbh; jh++) {
            if (!first) slots = v2_rotate(slots);
            first = false;
            const int n_kb = jh / 2 + 1;
            uint32_t started = 0;
            for (int kb = 0; kb < n_kb; kb++) {
              const bool last_kb = kb == n_kb - 1;
              // the 26 pairs, fully unrolled: group, release and completion points are compile-time constants, the
              // single issuing thread is left with masks and shifts (its instruction stream paces the small N = 64 MMAs)
#pragma unroll
              for (int tt = NDIG - 1; tt >= 0; tt--) {
                const uint32_t bslot = b_cnt & (V2_B_SLOTS - 1), bpar = (b_cnt / V2_B_SLOTS) & 1u;
                ptx::mbar_wait(&b_full[bslot], bpar);
                ptx::tc_fence_after();
                const uint64_t db = desc_b0 + (uint64_t)(bslot * (V2_B_BYTES >> 4));
#pragma unroll
                for (int s = (GMIN - tt > 0 ? GMIN - tt : 0); s < NDIG; s++) {
                  const uint32_t ai = a_cnt + (uint32_t)s, aslot = ai & (V2_A_SLOTS - 1), apar = (ai / V2_A_SLOTS) & 1u;
                  if (tt == NDIG - 1) {   // first use of this plane's tile in this K block
                    ptx::mbar_wait(&a_full[aslot], apar);
                    ptx::tc_fence_after();
                  }
                  const int gi = 2 * (NDIG - 1) - (s + tt);
                  const uint32_t slot = (slots >> (4 * gi)) & 15u;
                  const uint32_t acc = (started >> gi) & 1u;
                  if (!acc) {             // first MMA of the group in this unit: the slot's previous owner must be drained
                    ptx::mbar_wait(&tempty[slot], ((acc_par >> slot) & 1u) ^ 1u);
                    ptx::tc_fence_after();
                    started |= 1u << gi;
                  }
                  const uint32_t d_tmem = tmem_base + slot * V2_NX;
                  const uint64_t da = desc_a0 + (uint64_t)(aslot * (V2_A_BYTES >> 4));
                  ptx::mma_i8_ss(d_tmem, da, db, idesc, acc);
                  ptx::mma_i8_ss(d_tmem, da + 2, db + 2, idesc, 1);
                  ptx::mma_i8_ss(d_tmem, da + 4, db + 4, idesc, 1);
                  ptx::mma_i8_ss(d_tmem, da + 6, db + 6, idesc, 1);
                  if (tt == (GMIN - s > 0 ? GMIN - s : 0)) ptx::mma_commit(&a_empty[aslot]);   // last burst that uses plane s
                  if (last_kb && tt == (s + tt - (NDIG - 1) > 0 ? s + tt - (NDIG - 1) : 0)) {    // the group's last pair
                    ptx::mma_commit(&tfull[slot]);
                    acc_par ^= 1u << slot;
                  }
                }
                ptx::mma_commit(&b_empty[bslot]);
                b_cnt++;
              }
              a_cnt += NDIG;
            }
          }
        }
      }
    }
  } else {
    // ===================================================================== epilogue warps
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // row of the tile = unmeasured SNP
    const int half = ew >> 2;                // which 32 of the unit's 64 X rows this thread owns
    const int etid = threadIdx.x - 128;
    uint32_t slots = SLOTS0, full_par = 0;
    bool first = true;
    int unit = 0;
    for (int ct = blockIdx.x; ct < n_tiles; ct += gridDim.x) {
      const OzTile t = tiles[ct];
      if (t.win < 0) continue;
      const OzWin w = wins[t.win];
      const int e_x = ex[t.win];
      const int nbh = (w.n_t + V2_NX - 1) / V2_NX;
      double p_info = 0.0, p_z = 0.0;
      for (int jh = 0; jh < nbh; jh++, unit++) {
        if (!first) slots = v2_rotate(slots);
        first = false;
        double* yb = ys + (unit & 1) * V2_NX;
        if (etid < V2_NX) {
          const int k = jh * V2_NX + etid;
          yb[etid] = k < w.n_t ? y[w.off_t + k] : 0.0;
        }
        oz_epi_bar_sync();   // (the readers of this buffer two units ago passed the previous unit's barrier)
        double a[32];
#pragma unroll
        for (int e = 0; e < 32; e++) a[e] = 0.0;
#pragma unroll 1
        for (int idx = 0; idx < NG; idx++) {
          const int gi = idx == 5 ? 6 : idx == 6 ? 5 : idx;       // order of completion
          const int gw = 2 * (NDIG - 1) - gi;
          const double sc = __hiloint2double((1023 + DBITS * gw - 2 * QBITS + e_x) << 20, 0);   // 2^(8 g' - 2 QBITS + e_X)
          const uint32_t slot = (slots >> (4 * gi)) & 15u;
          ptx::mbar_wait(&tfull[slot], (full_par >> slot) & 1u);
          full_par ^= 1u << slot;
          ptx::tc_fence_after();
          const uint32_t taddr = tmem_base + slot * V2_NX + ((uint32_t)(quad * 32) << 16) + half * 32;
          uint32_t v0[16], v1[16];
          ptx::tmem_ld_32x32b_x16(taddr, v0);
          ptx::tmem_ld_32x32b_x16(taddr + 16, v1);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty[slot]);
#pragma unroll
          for (int e = 0; e < 16; e++) {
            a[e] = fma(sc, oz_int_to_double((int)v0[e]), a[e]);
            a[16 + e] = fma(sc, oz_int_to_double((int)v1[e]), a[16 + e]);
          }
        }
        const double* yj = yb + half * 32;
#pragma unroll
        for (int e = 0; e < 32; e++) {
          p_info = fma(a[e], a[e], p_info);
          p_z = fma(a[e], yj[e], p_z);
        }
      }
      // the two halves of a row, in a fixed order
      oz_epi_bar_sync();   // the previous tile's readers of red are done
      if (half == 1) {
        red[r] = p_info;
        red[128 + r] = p_z;
      }
      oz_epi_bar_sync();
      if (half == 0) {
        const int u = t.ut * OZ_TILE + r;
        if (u < w.n_u) {
          const double s_info = p_info + red[r], s_z = p_z + red[128 + r];
          const double inf = nanflag[w.off_u + u] ? __longlong_as_double(0x7ff8000000000000ll) : fabs(s_info);   // info = |b21 B11^-1 b12|  (dist.cpp:198)
          zu[w.off_u + u] = s_z / sqrt(inf);           // z / sqrt(info)               (dist.cpp:200)
          info[w.off_u + u] = inf;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- operand preparation -------------------------------------------------------------------------------------------
__global__ void oz_exponent_kernel(const unsigned long long* __restrict__ amax, int n, int* ex) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = __longlong_as_double((long long)amax[i]);
  int e = 0;
  if (m > 0.0 && m < 1e300) frexp(m, &e);    // m = f 2^e, f in [0.5, 1)  ->  |v| 2^-e < 1
  ex[i] = e;
}

// q = rint(v 2^(QBITS - e)) = sum_p d_p 256^p, d_p in [-128, 127]: q + BIAS has the unsigned base-256 digits d_p + 128,
// so the bytes of (q + BIAS) ^ BIAS are the signed digits themselves -- no carry chain (|v| 2^-e < 1)
__device__ __forceinline__ void oz_digits(double v, int e, int8_t (&d)[NDIG]) {
  const unsigned long long q = ((unsigned long long)__double2ll_rn(ldexp(v, QBITS - e)) + OZ_BIAS) ^ OZ_BIAS;
#pragma unroll
  for (int p = 0; p < NDIG; p++) d[p] = (int8_t)(uint8_t)(q >> (8 * p));
}

// X = L^-1 (row-major n_t x ld_t, lower triangular) -> digit planes: one CTA per (row, window), k contiguous
__global__ void __launch_bounds__(256)
oz_slice_x_kernel(const SolveWin* __restrict__ wins, const OzWin* __restrict__ ow, const double* __restrict__ X,
                  const int* __restrict__ ex, int8_t* __restrict__ planes, int kpad) {
  const SolveWin w = wins[blockIdx.y];
  const OzWin o = ow[blockIdx.y];
  const int r = blockIdx.x;
  if (r >= o.rb) return;
  const int e = ex[blockIdx.y];
  const double* row = X + w.off_tt + (long long)r * w.ld_t;
  for (int k4 = 4 * threadIdx.x; k4 < kpad; k4 += 4 * 256) {   // 4 consecutive k per thread: one 32-bit store per plane
    int8_t d[4][NDIG];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = k4 + q;
      oz_digits((r < w.n_t && k <= r) ? row[k] : 0.0, e, d[q]);
    }
#pragma unroll
    for (int p = 0; p < NDIG; p++)
      *reinterpret_cast<uint32_t*>(planes + (o.b_row0 + (long long)p * o.rb + r) * kpad + k4) =
          (uint32_t)(uint8_t)d[0][p] | ((uint32_t)(uint8_t)d[1][p] << 8) | ((uint32_t)(uint8_t)d[2][p] << 16) |
          ((uint32_t)(uint8_t)d[3][p] << 24);
  }
}

// B21^T (row-major n_t x ld_u: k rows, u contiguous) -> digit planes with k contiguous per unmeasured SNP u (transpose
// through shared memory: a CTA takes 32 u x 128 k)
__global__ void __launch_bounds__(256)
oz_slice_b_kernel(const SolveWin* __restrict__ wins, const OzWin* __restrict__ ow, const double* __restrict__ ut,
                  int8_t* __restrict__ planes, int kpad, uint8_t* __restrict__ nanflag) {
  __shared__ double tile[128][33];
  const SolveWin w = wins[blockIdx.z];
  const OzWin o = ow[blockIdx.z];
  const int u0 = blockIdx.x * 32, k0 = blockIdx.y * 128;
  if (u0 >= o.ra || k0 >= kpad) return;
  const double* B = ut + w.off_ut;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int kk = wid; kk < 128; kk += 8) {
    const int k = k0 + kk, u = u0 + lane;
    tile[kk][lane] = (k < w.n_t && u < w.n_u) ? B[(long long)k * w.ld_u + u] : 0.0;
  }
  __syncthreads();
  // warp -> one u at a time, lane -> 4 consecutive k: a warp writes 128 contiguous bytes per plane
  for (int uu = wid; uu < 32; uu += 8) {
    int8_t d[4][NDIG];
    bool bad = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      double v = tile[4 * lane + q][uu];
      if (!(fabs(v) <= 1.95)) {   // NaN (sd = 0) or beyond what the digits carry at scale 2^0
        bad = true;
        v = 0.0;
      }
      oz_digits(v, 0, d[q]);
    }
    if (bad) nanflag[w.off_u + u0 + uu] = 1;
#pragma unroll
    for (int p = 0; p < NDIG; p++) {
      const uint32_t word = (uint32_t)(uint8_t)d[0][p] | ((uint32_t)(uint8_t)d[1][p] << 8) | ((uint32_t)(uint8_t)d[2][p] << 16) |
                            ((uint32_t)(uint8_t)d[3][p] << 24);
      *reinterpret_cast<uint32_t*>(planes + (o.a_row0 + (long long)p * o.ra + u0 + uu) * kpad + k0 + 4 * lane) = word;
    }
  }
}

}  // namespace

int make_row_tensor_map(Ctx* ctx, CUtensorMap* out, const void* base, int64_t n_rows, int64_t k_elems,
                        int64_t k_stride_bytes, int format, int box_rows);

// GB_OZ_KERNEL=2 selects the all-groups-live kernel (measured slower: see its header); default = the streaming kernel
static bool oz_use_v2() {   // read per call (plan and launch of a batch must see the same value)
  const char* e = getenv("GB_OZ_KERNEL");
  return e && atoi(e) == 2;
}

size_t ozaki_win_bytes() { return sizeof(OzWin); }
size_t ozaki_tile_bytes() { return sizeof(OzTile); }

// Host-side plan of the int8-split solve for a batch: fills the window / tile descriptors (in the order of `wins`,
// heaviest first) and returns the plane sizes.  kpad = K bytes per plane row (multiple of 128, >= the largest n_t).
void ozaki_plan(const SolveWin* wins, int n_wins, int kpad, int n_ctas, void* ow_out, std::vector<uint8_t>* tiles_out,
                long long* a_rows, long long* b_rows) {
  OzWin* ow = static_cast<OzWin*>(ow_out);
  long long ar = 0, br = 0;
  std::vector<OzTile> tiles;
  for (int i = 0; i < n_wins; i++) {
    OzWin o{};
    o.n_t = wins[i].n_t;
    o.n_u = wins[i].n_u;
    o.nbt = (o.n_t + OZ_TILE - 1) / OZ_TILE;
    o.nbu = (o.n_u + OZ_TILE - 1) / OZ_TILE;
    o.ra = o.nbu * OZ_TILE;
    o.rb = o.nbt * OZ_TILE;
    o.a_row0 = ar;
    o.b_row0 = br;
    o.off_t = wins[i].off_t;
    o.off_u = wins[i].off_u;
    ar += (long long)NDIG * o.ra;
    br += (long long)NDIG * o.rb;
    ow[i] = o;
    for (int ut = 0; ut < o.nbu; ut++) tiles.push_back(OzTile{i, ut});
  }
  (void)kpad;
  // A tile of a window with nbt column blocks costs nbt (nbt + 1) / 2 K-block sweeps, known here: deal the tiles to the
  // persistent CTAs longest-processing-time-first (the windows come heaviest first) instead of round-robin.  CTA c runs
  // slots c, c + n_ctas, ...; unused slots hold win = -1.
  n_ctas = std::max(1, std::min<int>(n_ctas, (int)tiles.size()));
  std::vector<std::vector<OzTile>> per_cta((size_t)n_ctas);
  std::priority_queue<std::pair<long long, int>, std::vector<std::pair<long long, int>>, std::greater<>> load;
  for (int c = 0; c < n_ctas; c++) load.push({0, c});
  std::stable_sort(tiles.begin(), tiles.end(), [&](const OzTile& a, const OzTile& c) { return ow[a.win].nbt > ow[c.win].nbt; });
  size_t rounds = 0;
  for (const OzTile& t : tiles) {
    auto [l, c] = load.top();
    load.pop();
    per_cta[(size_t)c].push_back(t);
    rounds = std::max(rounds, per_cta[(size_t)c].size());
    long long cost = (long long)ow[t.win].nbt * (ow[t.win].nbt + 1) / 2;
    if (oz_use_v2()) {   // 64-row blocks of X, block jh sweeps jh / 2 + 1 K blocks
      cost = 0;
      for (long long jh = 0; jh < (ow[t.win].n_t + 63) / 64; jh++) cost += jh / 2 + 1;
    }
    load.push({l + cost, c});
  }
  std::vector<OzTile> slots(rounds * (size_t)n_ctas, OzTile{-1, 0});
  for (int c = 0; c < n_ctas; c++)
    for (size_t r = 0; r < per_cta[(size_t)c].size(); r++) slots[r * (size_t)n_ctas + (size_t)c] = per_cta[(size_t)c][r];
  tiles_out->resize(slots.size() * sizeof(OzTile));
  if (!slots.empty()) std::memcpy(tiles_out->data(), slots.data(), tiles_out->size());
  *a_rows = ar;
  *b_rows = br;
}

// Device side of the solve: X = L^-1 and y come from the factorisation stage (linv_row_kernel); slicing + the GEMM here.
void ozaki_tile_rows(const void* h_ow, int win, long long* a_row0, int* ra) {
  const OzWin& o = static_cast<const OzWin*>(h_ow)[win];
  *a_row0 = o.a_row0;
  *ra = o.ra;
}

// slice_b21 != 0: B21 is in d_ut as doubles (int8 panels: the Gram kernel finishes its tiles itself) and is sliced here;
// otherwise the finish pass of the Gram stage has already written the digit planes.
int launch_ozaki_solve(Ctx* ctx, const SolveWin* d_wins, const void* d_ow, const void* h_ow, int n_wins, const void* d_tiles,
                       int n_tiles, int kpad, const double* d_x, const double* d_ut, int slice_b21, int8_t* d_planes_a,
                       long long a_rows, int8_t* d_planes_b, long long b_rows, unsigned long long* d_amax, int* d_ex,
                       const double* d_y, uint8_t* d_nan, double* d_zu, double* d_info) {
  if (n_wins == 0 || n_tiles == 0) return GB_OK;
  const OzWin* how = static_cast<const OzWin*>(h_ow);
  int max_ra = 0, max_rb = 0, max_nt = 0;
  for (int i = 0; i < n_wins; i++) {
    max_ra = std::max(max_ra, how[i].ra);
    max_rb = std::max(max_rb, how[i].rb);
    max_nt = std::max(max_nt, how[i].n_t);
  }
  if (max_rb > OZ_Y_MAX) {
    ctx->err = "window has too many measured SNPs for the int8-split solve";
    return GB_ERR_UNSUPPORTED;
  }
  const bool trace = getenv("GB_OZ_TRACE") != nullptr;   // diagnostics: event-timed pieces, printed per call
  cudaEvent_t ev[6];
  int n_ev = 0;
  auto mark = [&]() {
    if (!trace) return;
    cudaEventCreate(&ev[n_ev]);
    cudaEventRecord(ev[n_ev++], ctx->stream);
  };
  mark();
  // d_amax: max |L^-1| per window, left by the triangular solve that formed X
  oz_exponent_kernel<<<(unsigned)((n_wins + 127) / 128), 128, 0, ctx->stream>>>(d_amax, n_wins, d_ex);
  mark();
  oz_slice_x_kernel<<<dim3((unsigned)max_rb, (unsigned)n_wins), 256, 0, ctx->stream>>>(d_wins, static_cast<const OzWin*>(d_ow), d_x,
                                                                                    d_ex, d_planes_b, kpad);
  mark();
  if (slice_b21)
    oz_slice_b_kernel<<<dim3((unsigned)(max_ra / 32), (unsigned)(kpad / 128), (unsigned)n_wins), 256, 0, ctx->stream>>>(
        d_wins, static_cast<const OzWin*>(d_ow), d_ut, d_planes_a, kpad, d_nan);
  mark();
  GB_CUDA(cudaGetLastError());
  ctx->launches += 3;
  const bool v1 = !oz_use_v2();
  CUtensorMap tm_a, tm_b;
  int rc;
  if ((rc = make_row_tensor_map(ctx, &tm_a, d_planes_a, a_rows, kpad, kpad, MAP_INT8, OZ_TILE))) return rc;
  if ((rc = make_row_tensor_map(ctx, &tm_b, d_planes_b, b_rows, kpad, kpad, MAP_INT8, v1 ? OZ_TILE : V2_NX))) return rc;
  static bool attr_set_dev[64] = {};
  if (!attr_set_dev[ctx->device & 63]) {
    GB_CUDA(cudaFuncSetAttribute(ozaki_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_ALLOC));
    GB_CUDA(cudaFuncSetAttribute(ozaki_solve_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM_ALLOC));
    attr_set_dev[ctx->device & 63] = true;
  }
  const int n_ctas = std::min(n_tiles, ctx->heavy_sms > 0 ? ctx->heavy_sms : ctx->sm_count);   // ozaki_plan dealt the slots to this many CTAs
  if (v1)
    ozaki_solve_kernel<<<(unsigned)n_ctas, OZ_THREADS, OZ_SMEM_ALLOC, ctx->stream>>>(
        tm_a, tm_b, static_cast<const OzWin*>(d_ow), static_cast<const OzTile*>(d_tiles), n_tiles, d_ex, d_y, d_nan, d_zu, d_info);
  else
    ozaki_solve_kernel_v2<<<(unsigned)n_ctas, OZ_THREADS, V2_SMEM_ALLOC, ctx->stream>>>(
        tm_a, tm_b, static_cast<const OzWin*>(d_ow), static_cast<const OzTile*>(d_tiles), n_tiles, d_ex, d_y, d_nan, d_zu, d_info);
  mark();
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  if (trace) {
    cudaStreamSynchronize(ctx->stream);
    static const char* const nm[] = {"exponent", "slice_x", "slice_b21", "gemm"};
    fprintf(stderr, "[oz trace]");
    for (int i = 0; i + 1 < n_ev; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      fprintf(stderr, " %s %.3f ms |", nm[i], ms);
    }
    fprintf(stderr, " tiles %d kpad %d\n", n_tiles, kpad);
    for (int i = 0; i < n_ev; i++) cudaEventDestroy(ev[i]);
  }
  return GB_OK;
}

}  // namespace gb
