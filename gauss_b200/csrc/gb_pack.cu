// gb_pack.cu -- K0: panel packing and per-row statistics (HBM-bound byte streams).
//
// The reference keeps genotypes as one std::string of '0'/'1'/'2' per population on every Snp
// (snp.h:109, filled by ReadGenotype gauss.cpp:720-785) and re-derives sum x, sum x^2 for every
// SNP pair.  Here a SNP is packed ONCE into a row of int8 bytes or E2M1 nibbles whose population
// blocks start on 32-column boundaries (zero padded: zeros are neutral for sum xy, sum x,
// sum x^2), and the per-population integer sums are computed once per SNP.
#include "gb_common.cuh"

namespace gb {

namespace {

// 4-bit E2M1 code of a dosage (sign | 2-bit exponent | 1-bit mantissa); 0xFF when the panel format
// does not take it: values E2M1 cannot hold exactly, and negative ones (the Gram epilogue of E2M1
// panels relies on non-negative counts).
__device__ __forceinline__ uint32_t e2m1_code(int v) {
  switch (v) {
    case 0: return 0x0;
    case 1: return 0x2;
    case 2: return 0x4;
    case 3: return 0x5;
    case 4: return 0x6;
    case 6: return 0x7;
    default: return 0xFFu;
  }
}

// One CTA per SNP row.  The row is cut into items of 512 consecutive dosages of one population (a warp-step: 16 dosages
// per lane); the 8 warps take items round-robin, so a 6,360-individual population and an 86-individual one no longer
// decide which warp finishes last (populations used to be dealt whole).  Per lane: sixteen source bytes out of five
// ALIGNED 32-bit loads (population offsets are not aligned), byte loads only where a block ends or the buffer itself is
// unaligned; when all sixteen are dosages 0 / 1 / 2 -- every real panel -- the work is word-wide: SIMD byte subtract,
// DP4A for sum x and sum x^2, and for E2M1 a bit compress of bytes to nibbles (code = dosage << 1); anything else
// takes the per-byte path.  One 128-bit (int8) or 64-bit (E2M1 nibbles, low nibble = lower K index) store per lane.
template <int FORMAT>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const uint8_t* __restrict__ src, long long src_stride, int is_ascii,
                 int8_t* __restrict__ dst, int k_elems, int k_stride, long long row0, int n_pops,
                 const int* __restrict__ pop_sizes, const int* __restrict__ koff,
                 int32_t* __restrict__ sx, int32_t* __restrict__ sxx, long long stat_ld, int* flags,
                 int seg_align) {
  __shared__ int s_sum[P_MAX], s_sq[P_MAX], s_soff[P_MAX + 1], s_item0[P_MAX + 1];
  const long long row = blockIdx.x;
  const uint8_t* s = src + row * src_stride;
  int8_t* d = dst + (row0 + row) * (long long)k_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < P_MAX) s_sum[threadIdx.x] = s_sq[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    int so = 0, it = 0;
    for (int p = 0; p < n_pops; p++) {
      s_soff[p] = so;
      s_item0[p] = it;
      so += pop_sizes[p];
      it += ((pop_sizes[p] + seg_align - 1) / seg_align * seg_align + 511) >> 9;
    }
    s_soff[n_pops] = so;
    s_item0[n_pops] = it;
  }
  __syncthreads();
  const uint32_t sub4 = is_ascii ? 0x30303030u : 0u;
  const int sub = is_ascii ? 48 : 0;
  bool bad = false, big = false;
  const bool aligned_src = (reinterpret_cast<uintptr_t>(src) & 3) == 0;   // else the first word of the buffer may not be touched
  const int n_items = s_item0[n_pops];
  int p = 0;
  for (int it = warp; it < n_items; it += 8) {
    while (it >= s_item0[p + 1]) p++;
    const int m = pop_sizes[p];
    const int kp = (m + seg_align - 1) / seg_align * seg_align;   // a multiple of 32: 16-dosage lanes never straddle it
    const int src_off = s_soff[p];
    const uint8_t* sp = s + src_off;
    const int j = ((it - s_item0[p]) << 9) + lane * 16;
    int sum = 0, sq = 0;
    if (j < kp) {
      uint32_t raw[4] = {0u, 0u, 0u, 0u};
      // the five words cover [a & ~3, a & ~3 + 20): inside the row's population block, and not before the buffer
      const bool fast = j + 20 <= m && (aligned_src || row * src_stride + src_off + j >= 4);
      if (fast) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(sp + j);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3) * 8;
        uint32_t w[5];
#pragma unroll
        for (int q = 0; q < 5; q++) w[q] = __ldg(wp + q);
#pragma unroll
        for (int q = 0; q < 4; q++) raw[q] = __vsub4(__funnelshift_r(w[q], w[q + 1], sh), sub4);
      }
      const uint32_t any = raw[0] | raw[1] | raw[2] | raw[3];
      const uint32_t three = ((raw[0] & (raw[0] >> 1)) | (raw[1] & (raw[1] >> 1)) | (raw[2] & (raw[2] >> 1)) | (raw[3] & (raw[3] >> 1))) &
                             0x01010101u;   // a byte with bits 0 and 1 set
      uint32_t out[4] = {0u, 0u, 0u, 0u};   // int8: the 16 bytes; E2M1: out[0], out[1] = 16 nibbles
      if (fast && (any & 0xFCFCFCFCu) == 0u && three == 0u) {
        // all sixteen are 0 / 1 / 2
        uint32_t su = 0u, sq_u = 0u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          su = __dp4a(raw[q], 0x01010101u, su);
          sq_u = __dp4a(raw[q], raw[q], sq_u);
        }
        sum = (int)su;
        sq = (int)sq_u;
        if (FORMAT == GB_PANEL_E2M1) {
          uint32_t c4[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint32_t t = (raw[q] | (raw[q] >> 4)) & 0x00FF00FFu;
            c4[q] = (t | (t >> 8)) & 0x0000FFFFu;
          }
          out[0] = (c4[0] | (c4[1] << 16)) << 1;   // E2M1 code of dosage 0 / 1 / 2 = dosage << 1
          out[1] = (c4[2] | (c4[3] << 16)) << 1;
        } else {
#pragma unroll
          for (int q = 0; q < 4; q++) out[q] = raw[q];
        }
      } else {
#pragma unroll
        for (int b = 0; b < 16; b++) {
          int v = 0;
          if (fast) v = (int)(signed char)(raw[b >> 2] >> (8 * (b & 3)));
          else if (j + b < m) v = (int)(signed char)((int)sp[j + b] - sub);
          sum += v;
          sq += v * v;
          big |= (unsigned)v > 2u;
          if (FORMAT == GB_PANEL_E2M1) {
            const uint32_t c = e2m1_code(v);
            bad |= c == 0xFFu;
            out[b >> 3] |= (c & 0xFu) << (4 * (b & 7));
          } else {
            out[b >> 2] |= (uint32_t)(v & 0xff) << (8 * (b & 3));
          }
        }
      }
      if (FORMAT == GB_PANEL_E2M1) *reinterpret_cast<uint2*>(d + ((koff[p] + j) >> 1)) = make_uint2(out[0], out[1]);
      else *reinterpret_cast<uint4*>(d + koff[p] + j) = make_uint4(out[0], out[1], out[2], out[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
      atomicAdd(&s_sum[p], sum);
      atomicAdd(&s_sq[p], sq);
    }
  }
  __syncthreads();
  if (threadIdx.x < n_pops) {
    sx[(long long)threadIdx.x * stat_ld + row0 + row] = s_sum[threadIdx.x];
    sxx[(long long)threadIdx.x * stat_ld + row0 + row] = s_sq[threadIdx.x];
  }
  if (bad) atomicOr(flags, 1);
  if (big) atomicOr(flags, 2);   // a byte outside {0, 1, 2}: the int8 fold may not assume |d| <= 4 m^2
  // zero the tail between the last population block and the row stride
  const int k_end = koff[n_pops - 1] + (pop_sizes[n_pops - 1] + seg_align - 1) / seg_align * seg_align;
  const int b_end = FORMAT == GB_PANEL_E2M1 ? k_end >> 1 : k_end;
  for (int j = b_end + threadIdx.x * 4; j < k_stride; j += 256 * 4)
    *reinterpret_cast<uint32_t*>(d + j) = 0u;
}


// 2-bit host rows ("pack2": 4 dosages per byte, low bits = lower K index, the same column positions as the
// E2M1 panel row) -> E2M1 nibbles plus per-population sum x, sum x^2.  One CTA per row; warp w expands
// populations w, w+8, ...; a lane turns one 32-bit word (16 dosages) into one 64-bit word of nibbles.
// Code c in {0,1,2} is dosage c and E2M1 nibble c << 1; code 3 is not a dosage the format holds (flagged).
// PCIe carries a quarter of the bytes a char/int8 row needs.
__global__ void __launch_bounds__(256)
expand2_rows_kernel(const uint8_t* __restrict__ src, long long src_stride, int8_t* __restrict__ dst, int k_stride,
                    long long row0, int n_pops, const int* __restrict__ pop_sizes, const int* __restrict__ koff,
                    int32_t* __restrict__ sx, int32_t* __restrict__ sxx, long long stat_ld, int* flags, int seg_align) {
  const long long row = blockIdx.x;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src + row * src_stride);
  uint2* d = reinterpret_cast<uint2*>(dst + (row0 + row) * (long long)k_stride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t bad = 0;
  for (int p = warp; p < n_pops; p += 8) {
    const int w0 = koff[p] >> 4;                                               // first 16-dosage word of the block
    const int nw = ((pop_sizes[p] + seg_align - 1) / seg_align * seg_align) >> 4;
    int ones = 0, twos = 0;
    for (int j = lane; j < nw; j += 32) {
      const uint32_t w = s[w0 + j];
      const uint32_t b0 = w & 0x55555555u, b1 = (w >> 1) & 0x55555555u;
      bad |= b0 & b1;
      ones += __popc(b0);
      twos += __popc(b1);
      uint32_t h[2] = {w & 0xFFFFu, w >> 16};
#pragma unroll
      for (int q = 0; q < 2; q++) {   // spread eight 2-bit codes over eight nibbles, then code -> code << 1
        uint32_t t = (h[q] | (h[q] << 8)) & 0x00FF00FFu;
        t = (t | (t << 4)) & 0x0F0F0F0Fu;
        t = (t | (t << 2)) & 0x33333333u;
        h[q] = t << 1;
      }
      d[w0 + j] = make_uint2(h[0], h[1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ones += __shfl_xor_sync(0xffffffffu, ones, o);
      twos += __shfl_xor_sync(0xffffffffu, twos, o);
    }
    if (lane == 0) {
      sx[(long long)p * stat_ld + row0 + row] = ones + 2 * twos;
      sxx[(long long)p * stat_ld + row0 + row] = ones + 4 * twos;
    }
  }
  if (bad) atomicOr(flags, 1);
  const int k_end = koff[n_pops - 1] + (pop_sizes[n_pops - 1] + seg_align - 1) / seg_align * seg_align;
  for (int j = (k_end >> 1) + threadIdx.x * 4; j < k_stride; j += 256 * 4)
    *reinterpret_cast<uint32_t*>(dst + (row0 + row) * (long long)k_stride + j) = 0u;
}

// Ternary host rows ("pack5": five dosages per byte, byte = d0 + 3 d1 + 9 d2 + 27 d3 + 81 d4, 1.6 bits per dosage;
// population p starts at byte boff5[p], blocks padded to 4 bytes) -> E2M1 nibbles plus per-population sum x,
// sum x^2.  A dosage in {0,1,2} has log2(3) = 1.58 bits of information, so this is within 1 % of the densest
// fixed-width code and takes another fifth off what crosses PCIe after pack2.  It is also the format panel rows are
// kept RESIDENT in when a whole genome sits on one GPU (gb_genome), so this kernel runs once per batch of windows
// and has to stream: 6.5 KB in and 16.7 KB out per row of the 33KG shape.
//
// Work item = 8 input words (32 bytes = 160 dosages) of one population -> 20 output words (80 bytes = 5 x 16 B,
// 16-byte aligned because population blocks start on 128-dosage boundaries), all in registers: one shared-memory
// table lookup per byte (five nibbles = 20 bits, the number of ones and twos, an invalid flag) and static shifts.
// A CTA takes ROWS_PER_CTA consecutive rows so the table is built once per ~100 KB of traffic.  Bytes 243..255 are
// not codes, and digits past a population's size must be zero (the host packer writes them so): both are flagged.
constexpr int EXP5_ROWS_PER_CTA = 8;
__global__ void __launch_bounds__(256)
expand5_rows_kernel(const uint8_t* __restrict__ src, long long src_stride, int8_t* __restrict__ dst, int k_elems,
                    int k_stride, long long row0, long long n_rows, int n_pops, const int* __restrict__ pop_sizes,
                    const int* __restrict__ koff, const int* __restrict__ boff5, int32_t* __restrict__ sx,
                    int32_t* __restrict__ sxx, long long stat_ld, int* flags, int seg_align) {
  __shared__ uint32_t lut[256];          // nibbles (digit << 1) of the five digits | ones << 20 | twos << 24 | invalid << 31
  __shared__ int item0[P_MAX + 1];       // first work item of each population
  __shared__ int cnt[2 * P_MAX];         // per population: ones, twos of the current row
  const int tid = threadIdx.x;
  {
    uint32_t v = 0x80000000u;
    if (tid < 243) {
      int b = tid, ones = 0, twos = 0;
      v = 0;
#pragma unroll
      for (int q = 0; q < 5; q++) {
        const int d = b % 3;
        b /= 3;
        ones += d == 1;
        twos += d == 2;
        v |= (uint32_t)(d << 1) << (4 * q);
      }
      v |= ((uint32_t)ones << 20) | ((uint32_t)twos << 24);
    }
    lut[tid] = v;
  }
  if (tid == 0) {
    int it = 0;
    for (int p = 0; p < n_pops; p++) {
      item0[p] = it;
      const int kp = (pop_sizes[p] + seg_align - 1) / seg_align * seg_align;
      it += (kp + 159) / 160;            // items cover the whole padded block (the padding is written as zeros)
    }
    item0[n_pops] = it;
  }
  __syncthreads();
  const int n_items = item0[n_pops];
  uint32_t bad = 0;
  const long long r_begin = (long long)blockIdx.x * EXP5_ROWS_PER_CTA;
  const long long r_end = r_begin + EXP5_ROWS_PER_CTA < n_rows ? r_begin + EXP5_ROWS_PER_CTA : n_rows;
  for (long long row = r_begin; row < r_end; row++) {
    if (tid < 2 * n_pops) cnt[tid] = 0;
    __syncthreads();
    const uint8_t* s = src + row * src_stride;
    uint32_t* drow = reinterpret_cast<uint32_t*>(dst + (row0 + row) * (long long)k_stride);
    for (int it = tid; it < n_items; it += 256) {
      int p = 0;
      while (item0[p + 1] <= it) p++;
      const int j = it - item0[p];
      const int m = pop_sizes[p];
      const int nbytes = (m + 4) / 5;                      // code bytes of the block
      const int nwords = (nbytes + 3) >> 2;
      const int kp = (m + seg_align - 1) / seg_align * seg_align;
      const uint32_t* sp = reinterpret_cast<const uint32_t*>(s + boff5[p]) + 8 * j;
      uint32_t w[8];
#pragma unroll
      for (int q = 0; q < 8; q++) w[q] = (8 * j + q < nwords) ? __ldg(sp + q) : 0u;
      // 32 bytes -> 32 x 20 bits of nibbles = 20 output words
      uint32_t out[20];
#pragma unroll
      for (int q = 0; q < 20; q++) out[q] = 0;
      int ones = 0, twos = 0;
#pragma unroll
      for (int bq = 0; bq < 32; bq++) {
        const uint32_t e = lut[(w[bq >> 2] >> (8 * (bq & 3))) & 255u];
        bad |= e & 0x80000000u;
        ones += (e >> 20) & 7;
        twos += (e >> 24) & 7;
        const uint32_t nib = e & 0xFFFFFu;
        const int bit = 20 * bq;                           // static: the loops are fully unrolled
        out[bit >> 5] |= nib << (bit & 31);
        if ((bit & 31) > 12) out[(bit >> 5) + 1] |= nib >> (32 - (bit & 31));
      }
      // digits at or past the population's size must be zero
      const int d0 = 160 * j;
      if (d0 + 160 > m) {
#pragma unroll
        for (int q = 0; q < 20; q++) {
          const int first = d0 + 8 * q;                    // dosage index of the word's lowest nibble
          if (first + 8 > m) {
            const uint32_t keep = first >= m ? 0u : (0xFFFFFFFFu >> (4 * (first + 8 - m)));
            bad |= out[q] & ~keep;
            out[q] &= keep;
          }
        }
      }
      uint4* o = reinterpret_cast<uint4*>(drow + (koff[p] >> 3) + 20 * j);
      const int words_left = (kp >> 3) - 20 * j;           // output words of the padded block from here on (multiple of 4)
#pragma unroll
      for (int q = 0; q < 5; q++)
        if (4 * q < words_left) o[q] = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
      if (ones | twos) {
        atomicAdd(&cnt[2 * p], ones);
        atomicAdd(&cnt[2 * p + 1], twos);
      }
    }
    // tail between the last population block and the row stride
    {
      const int k_end = koff[n_pops - 1] + (pop_sizes[n_pops - 1] + seg_align - 1) / seg_align * seg_align;
      for (int jb = (k_end >> 1) + tid * 4; jb < k_stride; jb += 256 * 4)
        *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(drow) + jb) = 0u;
    }
    __syncthreads();
    if (tid < n_pops) {
      const int ones = cnt[2 * tid], twos = cnt[2 * tid + 1];
      sx[(long long)tid * stat_ld + row0 + row] = ones + 2 * twos;
      sxx[(long long)tid * stat_ld + row0 + row] = ones + 4 * twos;
    }
    __syncthreads();
  }
  if (bad) atomicOr(flags, 1);
}

// dst[i] = panel row rows[i]; 16-byte vectors, one CTA per row.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const int8_t* __restrict__ panel, int k_stride, const int32_t* __restrict__ rows,
                   int8_t* __restrict__ dst) {
  const long long i = blockIdx.x;
  const uint4* s = reinterpret_cast<const uint4*>(panel + (long long)rows[i] * k_stride);
  uint4* d = reinterpret_cast<uint4*>(dst + i * (long long)k_stride);
  const int n16 = k_stride >> 4;
  for (int j = threadIdx.x; j < n16; j += 256) d[j] = s[j];
}

// Per listed row: the standard deviation the reference derives from CalWgtCov(g,g)
// (distmix.cpp:180-187, computeLD.cpp:100-103) or the pooled denominator factor of CalCor
// (util.cpp:67).  One thread per row, populations in panel order so the fp64 operation order
// is the reference's.
__global__ void row_prep_kernel(const int32_t* __restrict__ rows, long long n, int mode, int n_pops,
                                const int* __restrict__ pop_sizes, const double* __restrict__ coef,
                                const double* __restrict__ wgt, const int32_t* __restrict__ sx,
                                const int32_t* __restrict__ sxx, long long stat_ld,
                                double* __restrict__ sd, int32_t* __restrict__ pool,
                                double* __restrict__ rq, int32_t* __restrict__ st_sx,
                                double* __restrict__ st_mean) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long prow = rows[i];
  if (mode == GRAM_MIX) {
    double wsumcov = 0.0, wsum_mi_mj = 0.0, wsum_mi = 0.0, musq = 0.0;
    for (int p = 0; p < n_pops; p++) {
      const int m = pop_sizes[p];
      const long long s = sx[(long long)p * stat_ld + prow];
      const long long q = sxx[(long long)p * stat_ld + prow];
      const double d = (double)((long long)m * q - s * s);          // m*sumxy - sumx*sumy (exact)
      wsumcov = __dadd_rn(wsumcov, __dmul_rn(coef[p], d));          // util.cpp:118
      const double mean = __ddiv_rn((double)s, (double)m);
      st_sx[(long long)p * n + i] = (int32_t)s;       // list-ordered copies the Gram epilogue streams per tile
      st_mean[(long long)p * n + i] = mean;
      const double wm = __dmul_rn(wgt[p], mean);
      wsum_mi_mj = __dadd_rn(wsum_mi_mj, __dmul_rn(wm, mean));      // util.cpp:119
      wsum_mi = __dadd_rn(wsum_mi, wm);                             // util.cpp:120-121
      musq = fma(mean, mean, musq);
    }
    const double cov = __dsub_rn(__dadd_rn(wsumcov, wsum_mi_mj), __dmul_rn(wsum_mi, wsum_mi));
    sd[i] = __dsqrt_rn(cov);
    if (rq) rq[i] = musq / cov;   // feeds the analytic PD certificate (pd_bound_kernel)
  } else {
    long long s = 0, q = 0, n_ind = 0;
    for (int p = 0; p < n_pops; p++) {
      s += sx[(long long)p * stat_ld + prow];
      q += sxx[(long long)p * stat_ld + prow];
      n_ind += pop_sizes[p];
    }
    pool[i] = (int32_t)s;
    // sqrt(num_samples*sumxsq - sumx*sumx)   (util.cpp:67)
    sd[i] = __dsqrt_rn(__dsub_rn(__dmul_rn((double)n_ind, (double)q), __dmul_rn((double)s, (double)s)));
    if (rq) rq[i] = 0.0;
  }
}

}  // namespace

int launch_pack(Ctx* ctx, Panel* panel, const void* dev_src, int64_t src_stride, int is_ascii,
                int64_t row0, int64_t n_rows) {
  if (n_rows <= 0) return GB_OK;
  if (panel->format == GB_PANEL_E2M1)
    pack_rows_kernel<GB_PANEL_E2M1><<<(unsigned)n_rows, 256, 0, ctx->stream>>>(
        static_cast<const uint8_t*>(dev_src), src_stride, is_ascii, panel->d_rows, panel->k_elems, panel->k_stride,
        row0, panel->n_pops, panel->d_pop_sizes, panel->d_koff, panel->d_sx, panel->d_sxx, panel->capacity,
        panel->d_flags, panel->seg_align);
  else
    pack_rows_kernel<GB_PANEL_INT8><<<(unsigned)n_rows, 256, 0, ctx->stream>>>(
        static_cast<const uint8_t*>(dev_src), src_stride, is_ascii, panel->d_rows, panel->k_elems, panel->k_stride,
        row0, panel->n_pops, panel->d_pop_sizes, panel->d_koff, panel->d_sx, panel->d_sxx, panel->capacity,
        panel->d_flags, panel->seg_align);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_expand2(Ctx* ctx, Panel* panel, const void* dev_src, int64_t src_stride, int64_t row0, int64_t n_rows) {
  if (n_rows <= 0) return GB_OK;
  expand2_rows_kernel<<<(unsigned)n_rows, 256, 0, ctx->stream>>>(
      static_cast<const uint8_t*>(dev_src), src_stride, panel->d_rows, panel->k_stride, row0, panel->n_pops,
      panel->d_pop_sizes, panel->d_koff, panel->d_sx, panel->d_sxx, panel->capacity, panel->d_flags, panel->seg_align);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_expand5(Ctx* ctx, Panel* panel, const void* dev_src, int64_t src_stride, int64_t row0, int64_t n_rows) {
  if (n_rows <= 0) return GB_OK;
  if (src_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(dev_src) & 3)) {
    ctx->err = "pack5 rows must be 4-byte aligned (use gb_pack5_row_bytes() as the row stride)";
    return GB_ERR_BAD_ARG;
  }
  const long long n_ctas = (n_rows + EXP5_ROWS_PER_CTA - 1) / EXP5_ROWS_PER_CTA;
  expand5_rows_kernel<<<(unsigned)n_ctas, 256, 0, ctx->stream>>>(
      static_cast<const uint8_t*>(dev_src), src_stride, panel->d_rows, panel->k_elems, panel->k_stride, row0, n_rows,
      panel->n_pops, panel->d_pop_sizes, panel->d_koff, panel->d_boff5, panel->d_sx, panel->d_sxx, panel->capacity,
      panel->d_flags, panel->seg_align);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_gather_rows(Ctx* ctx, const Panel* panel, const int32_t* d_rows, int64_t n, int8_t* dst) {
  if (n <= 0) return GB_OK;
  gather_rows_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(panel->d_rows, panel->k_stride, d_rows, dst);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_row_prep(Ctx* ctx, const Panel* panel, const int32_t* d_rows, int64_t n, int mode,
                    const double* d_coef, const double* d_wgt, double* d_sd, int32_t* d_pool, double* d_rq,
                    int32_t* d_st_sx, double* d_st_mean) {
  if (n <= 0) return GB_OK;
  const int bs = 128;
  row_prep_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, ctx->stream>>>(
      d_rows, n, mode, panel->n_pops, panel->d_pop_sizes, d_coef, d_wgt, panel->d_sx, panel->d_sxx,
      panel->capacity, d_sd, d_pool, d_rq, d_st_sx, d_st_mean);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

}  // namespace gb
