// gb_solve.cu -- K2: batched fp64 Cholesky of B11 and blocked triangular solves giving z_u, info_u.
//
// Replaces MakePosDef + InvMat + the three MpMatMat per unmeasured SNP of the reference
// (dist.cpp:181-202, distmix.cpp:202-228; util.cpp:262-264,298-318).  With B11 = L L^T:
//     W = L^-1 B21^T,  y = L^-1 Z1   ->   z = W^T y,  info = |colsumsq(W)|,  z_out = z / sqrt(info)
// which equals b21 B11^-1 Z1 and |b21 B11^-1 b12| of the reference.
// MakePosDef is a no-op whenever lambda_min(B11) >= min_abs_eig; that is certified by a second
// Cholesky of B11 - min_abs_eig*I (succeeds  <=>  lambda_min > min_abs_eig), run in the SAME
// launches as the real factorisation (the shifted copies are just extra "windows"); otherwise
// the window is flagged GB_ERR_NOT_PD -- there is no eigen-clip path and no CPU fallback.
//
// Storage: B11 / L column-major n_t x ld_t (lower triangle significant); B21^T / W row-major
// n_t x ld_u (unmeasured SNPs contiguous).  Everything is blocked by NB = 64.
//
// Cholesky = right-looking, three kernels per block column k, batched over all windows:
//   chol_diag_kernel    one CTA per window turns the 64x64 diagonal block into inv(L_kk) and publishes it
//   chol_panel_kernel   blocks below it: L_ik = A_ik inv(L_kk)^T
//   chol_update_kernel  trailing tiles A_ij -= L_ik L_jk^T
// trsm_finalize_kernel  one CTA per (window, 128 unmeasured SNPs): forward substitution by row
//                       blocks with cp.async double-buffered operand chunks, W written in place,
//                       y = L^-1 Z1 carried along as one more right-hand-side column, column sums
//                       of squares and W^T y reduced in a fixed order.
#include "gb_common.cuh"

namespace gb {

namespace {

constexpr int NB = 64;
constexpr int STATUS_BREAKDOWN = 1;

__device__ __forceinline__ int win_nb(int n) { return (n + NB - 1) / NB; }

// C(8x8) += A(8x4) B(4x8) on the fp64 tensor cores: lane holds A(row = lane/4, k = lane%4),
// B(k = lane%4, col = lane/4), C(row = lane/4, cols 2(lane%4), 2(lane%4)+1).  Same 64 FMA/clk/SM peak
// as DFMA on B200 (tools/dmma_probe.cu) at 1/8 of the instructions and operand fetches.
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 64 x 64 x 64 product on 8 warps: C(m, n) += sign * sum_k As[k*TS + m] * Bs[k*TS + n], k < kmax (multiple of 4).
// Warp (wr = warp/4, wc = warp%4) owns rows 32 wr .. +31 and columns 16 wc .. +15 as 4 x 2 MMA tiles;
// C[mt][nt][e] is element (32 wr + 8 mt + lane/4, 16 wc + 8 nt + 2 (lane%4) + e).  TS = 68 doubles
// (8 words mod 32) keeps the fragment loads conflict-free.
constexpr int TS = NB + 4;
__device__ __forceinline__ void tile64_dmma(const double* As, const double* Bs, double (&C)[4][2][2], int kmax,
                                            double sign, int lane, int warp) {
  const int gid = lane >> 2, tig = lane & 3;
  const int row0 = 32 * (warp >> 2) + gid, col0 = 16 * (warp & 3) + gid;
#pragma unroll 4
  for (int k = 0; k < kmax; k += 4) {
    double a[4], b[2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++) a[mt] = sign * As[(k + tig) * TS + row0 + 8 * mt];
#pragma unroll
    for (int nt = 0; nt < 2; nt++) b[nt] = Bs[(k + tig) * TS + col0 + 8 * nt];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) dmma_m8n8k4(C[mt][nt][0], C[mt][nt][1], a[mt], b[nt]);
  }
}

constexpr int MP = NB + 2;  // column stride of the 64 x 64 shared-memory blocks of the factorisation

// Warp-synchronous B x B version of the same transform (B = 16): lane r < B keeps row r of A (lower)
// and row r of Y = inv(L) in registers; column j of L and row j of Y travel by register shuffles (the
// shared-memory round trips they used to take, with a __syncwarp each, were on the critical path of the 64
// sequential pivots of a diagonal block: ncu put ~520 clocks on a pivot).  No block barrier inside.  A at Ms[(o+c)*MP + o+r]; inv(L) is
// written (lower triangle, zeros above) to Xs at the same coordinates.  Kept out of line: the fully
// unrolled pivot loop is straight-line code (a 32 x 32 version was 140 KB and bound by instruction
// fetch); one 16 x 16 copy is called four times per diagonal block.
template <int B>
__device__ __noinline__ bool invchol_warp(const double* Ms, double* Xs, int o, double* colbuf, double* rowbuf,
                                          int lane) {
  double a[B], y[B];
  const bool act = lane < B;
#pragma unroll
  for (int c = 0; c < B; c++) {
    a[c] = (act && c <= lane) ? Ms[(o + c) * MP + o + lane] : 0.0;
    y[c] = 0.0;
  }
  bool bad_any = false;
  (void)colbuf;
  (void)rowbuf;
#pragma unroll
  for (int j = 0; j < B; j++) {
    const double piv = __shfl_sync(0xffffffffu, a[j], j);
    const bool bad = !(piv > 0.0);
    bad_any |= bad;
    const double rs = rsqrt(bad ? 1.0 : piv);
    const double l = a[j] * rs;                 // lanes r > j: L(r, j)
    const bool is_j = lane == j;
#pragma unroll
    for (int c = 0; c < j; c++) y[c] = is_j ? y[c] * rs : y[c];   // Y(j, c) *= rs
    if (is_j) y[j] = rs;                                          // Y(j, j) = rs
    // column j of L and row j of Y travel by register shuffles (no shared-memory round trip, no __syncwarp on the
    // critical path of the 64 sequential pivots of a diagonal block)
    const double lm = (act && lane > j) ? -l : 0.0;
#pragma unroll
    for (int c = j + 1; c < B; c++) a[c] = fma(lm, __shfl_sync(0xffffffffu, l, c), a[c]);       // A(r, c) -= L(r, j) L(c, j)
#pragma unroll
    for (int c = 0; c <= j; c++) y[c] = fma(lm, __shfl_sync(0xffffffffu, y[c], j), y[c]);       // Y(r, c) -= L(r, j) Y(j, c)
  }
  if (act) {
#pragma unroll
    for (int c = 0; c < B; c++) Xs[(o + c) * MP + o + lane] = (c <= lane) ? y[c] : 0.0;
  }
  return !bad_any;
}

// Step k, part 1: one CTA per window turns the 64x64 diagonal block into inv(L_kk) and publishes it
// (the only form of L_kk anything downstream uses: panel blocks, y and the solve).  This is the
// sequential critical path of the factorisation, so it runs exactly once per window and step and
// is organised for latency: a right-looking factorisation over 4 x 4 blocks of 16 inside shared
// memory -- warp 0 inverts the 16 x 16 diagonal block in registers, all 8 warps form the panel blocks
// L_ik = A_ik inv(L_kk)^T and the trailing update -- followed by the block forward substitution
//   X_ij = -X_ii sum_{q=j..i-1} L_iq X_qj   (i > j),   X = inv(L_kk).
__device__ __forceinline__ void chol_diag_body(const SolveWin& w, int win, const double* __restrict__ tt, double* dinv,
                                               int* status, int k, double* sm) {
  const int n = w.n_t;
  if (k >= win_nb(n)) return;
  const double* A = tt + w.off_tt;
  const int ld = w.ld_t;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int B = 16, NBLK = NB / B;
  double* Ms = sm;                      // [64][MP] A_kk (lower); off-diagonal blocks become L
  double* Xs = Ms + NB * MP;            // [64][MP] inv(L_kk)
  double* Tb = Xs + NB * MP;            // [3][16][17] block temporaries of the substitution
  double* colbuf = Tb + 3 * B * (B + 1);
  double* rowbuf = colbuf + B;
  __shared__ int ok_s;
  const int k0 = k * NB;
  // rows/cols past n are padded with the identity
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int c = idx >> 6, r = idx & 63;
    double v = (r == c) ? 1.0 : 0.0;
    if (k0 + r < n && k0 + c < n && r >= c) v = A[(long long)(k0 + c) * ld + k0 + r];
    Ms[c * MP + r] = v;
    Xs[c * MP + r] = 0.0;
  }
  if (tid == 0) ok_s = 1;
  __syncthreads();
  const int er = tid & 15, ec = (tid >> 4) & 15;   // element (er, ec) of a 16 x 16 block; one block per 256 threads
  for (int kb = 0; kb < NBLK; kb++) {
    const int ko = kb * B;
    if (warp == 0 && !invchol_warp<B>(Ms, Xs, ko, colbuf, rowbuf, lane) && lane == 0) ok_s = 0;
    __syncthreads();
    // panel: L_ik(r, c) = sum_{j<=c} A_ik(r, j) X_kk(c, j), blocks i = kb+1 .. 3 (one element per thread and block)
    double pv[NBLK - 1];
#pragma unroll
    for (int bi = 0; bi < NBLK - 1; bi++) {
      const int io = (kb + 1 + bi) * B;
      double v = 0.0;
      if (kb + 1 + bi < NBLK)
        for (int j = 0; j <= ec; j++) v = fma(Ms[(ko + j) * MP + io + er], Xs[(ko + j) * MP + ko + ec], v);
      pv[bi] = v;
    }
    __syncthreads();   // every A_ik entry has been read before it is overwritten by L_ik
#pragma unroll
    for (int bi = 0; bi < NBLK - 1; bi++)
      if (kb + 1 + bi < NBLK) Ms[(ko + ec) * MP + (kb + 1 + bi) * B + er] = pv[bi];
    __syncthreads();
    // trailing update: A_ij(r, c) -= sum_q L_ik(r, q) L_jk(c, q), kb < j <= i
    for (int i = kb + 1; i < NBLK; i++)
      for (int j = kb + 1; j <= i; j++) {
        double v = Ms[(j * B + ec) * MP + i * B + er];
#pragma unroll
        for (int q = 0; q < B; q++) v = fma(-Ms[(ko + q) * MP + i * B + er], Ms[(ko + q) * MP + j * B + ec], v);
        Ms[(j * B + ec) * MP + i * B + er] = v;
      }
    __syncthreads();
  }
  // block forward substitution for the off-diagonal blocks of X = inv(L), by block distance d = i - j
  for (int d = 1; d < NBLK; d++) {
    double tv[NBLK - 1];
#pragma unroll
    for (int bj = 0; bj < NBLK - 1; bj++) {      // T_ij(r, c) = sum_{q=j..i-1} sum_t L_iq(r, t) X_qj(t, c)
      double v = 0.0;
      if (bj + d < NBLK) {
        const int i = bj + d;
        for (int q = bj; q < i; q++)
          for (int t = 0; t < B; t++) v = fma(Ms[(q * B + t) * MP + i * B + er], Xs[(bj * B + ec) * MP + q * B + t], v);
      }
      tv[bj] = v;
    }
#pragma unroll
    for (int bj = 0; bj < NBLK - 1; bj++)
      if (bj + d < NBLK) Tb[bj * B * (B + 1) + ec * (B + 1) + er] = tv[bj];
    __syncthreads();
#pragma unroll
    for (int bj = 0; bj < NBLK - 1; bj++)        // X_ij(r, c) = -sum_{q<=r} X_ii(r, q) T_ij(q, c)
      if (bj + d < NBLK) {
        const int i = bj + d;
        double v = 0.0;
        for (int q = 0; q <= er; q++) v = fma(Xs[(i * B + q) * MP + i * B + er], Tb[bj * B * (B + 1) + ec * (B + 1) + q], v);
        Xs[(bj * B + ec) * MP + i * B + er] = -v;
      }
    __syncthreads();
  }
  if (!ok_s && tid == 0) atomicOr(&status[win], STATUS_BREAKDOWN);
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int c = idx >> 6, r = idx & 63;
    dinv[w.off_dinv + (long long)k * NB * NB + c * NB + r] = Xs[c * MP + r];  // zero above the diagonal
  }
}

__global__ void __launch_bounds__(256)
chol_diag_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ tt, double* dinv, int* status,
                 const int* __restrict__ skip, int k) {
  if (skip && skip[blockIdx.x]) return;  // certificate copy whose analytic bound already holds
  extern __shared__ __align__(16) double sm[];
  const SolveWin w = wins[blockIdx.x];
  chol_diag_body(w, blockIdx.x, tt, dinv, status, k, sm);
}

// Step k, part 2: panel blocks below the diagonal, L_ik = A_ik inv(L_kk)^T, one CTA per block.
__global__ void __launch_bounds__(256)
chol_panel_kernel(const SolveWin* __restrict__ wins, double* tt, const double* __restrict__ dinv,
                  const int* __restrict__ skip, int k) {
  if (skip && skip[blockIdx.y]) return;
  const SolveWin w = wins[blockIdx.y];
  const int n = w.n_t;
  const int nb = win_nb(n);
  const int ib = k + 1 + blockIdx.x;
  if (ib >= nb) return;
  double* A = tt + w.off_tt;
  const int ld = w.ld_t;
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) double sm[];
  double* X = sm;                 // [64][TS] inv(L_kk), X[j*TS + c] = inv(L_kk)(c, j)
  double* T = X + NB * TS;        // [64][TS] A_ik tile, T[j*TS + r] = A(ib*64+r, k*64+j)

  const int k0 = k * NB;
  const int i0 = ib * NB;
  const double* D = dinv + w.off_dinv + (long long)k * NB * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int c = idx >> 6, r = idx & 63;
    X[c * TS + r] = D[c * NB + r];
    T[c * TS + r] = (i0 + r < n && k0 + c < n) ? A[(long long)(k0 + c) * ld + i0 + r] : 0.0;
  }
  __syncthreads();

  // L(r, c) = sum_{j<=c} A(r, j) X(c, j) = sum_j T[j][r] X[j][c] on the fp64 tensor cores
  // (X[j][c] = 0 for j > c, so the k range of column quarter wc ends at 16 wc + 16)
  {
    const int lane = tid & 31, warp = tid >> 5;
    double C[4][2][2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) C[mt][nt][0] = C[mt][nt][1] = 0.0;
    tile64_dmma(T, X, C, 16 * ((warp & 3) + 1), 1.0, lane, warp);
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
      const int r = i0 + 32 * (warp >> 2) + 8 * mt + gid;
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = k0 + 16 * (warp & 3) + 8 * nt + 2 * tig + e;
          if (r < n && c < n) A[(long long)c * ld + r] = C[mt][nt][e];
        }
    }
  }
}

// trailing update of step k: A_ij -= L_ik L_jk^T for k < j <= i.  One CTA per block ROW i: L_ik stays in shared memory
// for the whole strip j = k+1 .. i, and the next tile's L_jk and A_ij travel in registers while the current product
// runs (a CTA per tile spent its 9 us on three exposed loads for half a microsecond of tensor work, and a step of the
// largest window launched 171 of them).
__global__ void __launch_bounds__(256)
chol_update_kernel(const SolveWin* __restrict__ wins, double* tt, double* dinv, int* status,
                   const int* __restrict__ skip, int k) {
  if (skip && skip[blockIdx.y]) return;
  const SolveWin w = wins[blockIdx.y];
  const int n = w.n_t;
  const int nb = win_nb(n);
  const int tb = nb - k - 1;
  const int ii = blockIdx.x;
  if (tb <= 0 || ii >= tb) return;
  const int i0 = (k + 1 + ii) * NB, k0 = k * NB;
  double* A = tt + w.off_tt;
  const int ld = w.ld_t;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;

  extern __shared__ __align__(16) double sm[];
  double* Ls = sm;            // Ls[kk*TS + r] = L(i0+r, k0+kk)
  double* Rs = sm + NB * TS;  // Rs[kk*TS + c] = L(j0+c, k0+kk)
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int kk = idx >> 6, r = idx & 63;
    Ls[kk * TS + r] = (i0 + r < n) ? A[(long long)(k0 + kk) * ld + i0 + r] : 0.0;
  }
  double pr[16];          // the next tile's L_jk, element idx = tid + 256 q
  double Cn[4][2][2];     // ... and its A_ij fragment
  auto fetch = [&](int jj) {
    const int j0 = (k + 1 + jj) * NB;
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int idx = tid + 256 * q, kk = idx >> 6, r = idx & 63;
      pr[q] = (j0 + r < n) ? A[(long long)(k0 + kk) * ld + j0 + r] : 0.0;
    }
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
      const int r = i0 + 32 * (warp >> 2) + 8 * mt + gid;
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = j0 + 16 * (warp & 3) + 8 * nt + 2 * tig + e;
          Cn[mt][nt][e] = (r < n && c < n && r >= c) ? A[(long long)c * ld + r] : 0.0;
        }
    }
  };
  fetch(0);
  for (int jj = 0; jj <= ii; jj++) {
    const int j0 = (k + 1 + jj) * NB;
    __syncthreads();   // the previous product has read Rs (first trip: orders nothing that matters)
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int idx = tid + 256 * q;
      Rs[(idx >> 6) * TS + (idx & 63)] = pr[q];
    }
    double C[4][2][2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        C[mt][nt][0] = Cn[mt][nt][0];
        C[mt][nt][1] = Cn[mt][nt][1];
      }
    __syncthreads();   // Ls (first trip) and Rs are complete
    if (jj < ii) fetch(jj + 1);
    tile64_dmma(Ls, Rs, C, NB, -1.0, lane, warp);   // C -= L_ik L_jk^T
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
      const int r = i0 + 32 * (warp >> 2) + 8 * mt + gid;
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = j0 + 16 * (warp & 3) + 8 * nt + 2 * tig + e;
          if (r < n && c < n && r >= c) A[(long long)c * ld + r] = C[mt][nt][e];
        }
    }
  }
  // Look-ahead: the CTA of block row k + 1 has just finished the next diagonal tile (k+1, k+1): it factors and inverts it
  // right away, while the other CTAs of this launch are still updating the rest of the trailing matrix.  The 64 x 64
  // factorisation is the sequential part of a block step; as a separate launch it sat on the critical path.
  if (ii == 0) {
    __syncthreads();   // the tile is complete in global memory (visible to this CTA) and the operand buffers are free
    chol_diag_body(w, blockIdx.y, tt, dinv, status, k + 1, sm);
  }
}

// dst = src with the diagonal lowered by `shift`, per window (dst is the certificate copy)
__global__ void copy_shift_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ src, double* dst,
                                  double shift, const int* __restrict__ skip, int nreal) {
  if (skip[nreal + blockIdx.y]) return;
  const SolveWin w = wins[blockIdx.y];
  const int n = w.n_t, ld = w.ld_t;
  const long long total = (long long)n * ld;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx / ld), r = (int)(idx % ld);
    double v = src[w.off_tt + idx];
    if (r == c) v -= shift;
    dst[w.off_tt + idx] = v;
  }
}

// Analytic positive-definiteness certificate (DESIGN.md "MakePosDef").  B11 = R + lambda*I with
// R = D^-1/2 (Wn + E) D^-1/2: Wn = sum_p w_p m_p/(m_p-1) (m_p X_p X_p^T - s_p s_p^T) is a positively
// weighted sum of centred Gram matrices, hence PSD for w_p >= 0; E = M (diag(w) - w w^T) M^T with
// M = [mu_ip] satisfies  x^T E x >= -(sum(w)-1)_+ max(w) |M^T x|^2.  So
//     lambda_min(B11) >= lambda - gneg * sum_i (sum_p mu_ip^2) / cov_ii - rounding slack,
// gneg = (sum(w)-1)_+ * max(w) (0 for the pooled r of dist(): a Pearson matrix is PSD).  When that
// bound exceeds min_abs_eig the reference's MakePosDef is provably a no-op and the shifted
// factorisation is skipped; otherwise (lambda = 0, negative weights, NaN, ...) it runs.
__global__ void __launch_bounds__(256)
pd_bound_kernel(const SolveWin* __restrict__ wins, int nreal, const double* __restrict__ rq_t, double lambda,
                double gneg, double min_abs_eig, int* skip) {
  const SolveWin w = wins[blockIdx.x];
  __shared__ double red[256];
  double s = 0.0;
  if (rq_t)
    for (int i = threadIdx.x; i < w.n_t; i += 256) s += rq_t[w.off_t + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double lb = lambda - ((gneg > 0.0) ? gneg * red[0] : 0.0) - 1e-9 * (1.0 + w.n_t / 1000.0);
    skip[blockIdx.x] = 0;
    skip[nreal + blockIdx.x] = (lb > min_abs_eig) ? 1 : 0;  // NaN compares false -> exact certificate runs
  }
}

// ---------------------------------------------------------------------------------------------
// Blocked forward substitution W = L^-1 B21^T for 128 unmeasured SNPs, fused with the reductions.
//
// The two products of every 64-row block -- acc = B_i - sum_k L_ik W_k and W_i = inv(L_ii) acc --
// run on the fp64 tensor cores (mma.sync m8n8k4 .f64: same 64 FMA/clk/SM peak as DFMA on B200, but
// 256 FMAs per warp instruction with 2 operand registers, so the kernel is no longer bound by
// instruction issue and shared-memory operand fetch; measured tools/dmma_probe.cu).  8 warps =
// 2 row halves x 4 column quarters, 32 x 32 per warp = 4 x 4 MMA tiles, C fragment (row = lane / 4,
// columns 2 (lane % 4) + {0, 1}).  Operand chunks (32 k-steps of L and of W) are double-buffered
// with cp.async.cg; their shared-memory row strides (68 and 132 doubles = 8 words mod 32) make the
// A / B fragment loads conflict-free.  inv(L_ii) and the accumulator tile alias the chunk buffers
// once the k-loop of a row block is done; two CTAs share an SM.
constexpr int UB = 128;   // unmeasured SNPs (columns of W) per CTA; a 64-column variant serves launches of about one wave
constexpr int KC = 32;    // k-steps per staged chunk
constexpr int LSS = NB + 4;   // row stride of staged L / inv(L_ii)   [k][row]
constexpr int tr_smem_bytes(int ub) { return 2 * (KC * LSS + KC * (ub + 4)) * 8; }   // 102,400 for 128 columns
static_assert(2 * KC >= NB, "aliased tiles must fit the chunk buffers");

// 16-byte asynchronous copy of which only the first `n_valid` doubles (0, 1 or 2) are read; the rest is zero-filled
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int n_valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = n_valid >= 2 ? 16 : n_valid == 1 ? 8 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int UB_>
__global__ void __launch_bounds__(256, 2)
trsm_finalize_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ tt,
                     const double* __restrict__ dinv, double* ut, const double* __restrict__ zt,
                     double* zu, double* info, double* y_out, int tri, unsigned long long* amax_out) {
  const SolveWin w = wins[blockIdx.y];
  const int n = w.n_t, nu = w.n_u;
  constexpr int NTW = UB_ / 32;      // 8-column MMA tiles per warp
  constexpr int WSS = UB_ + 4;       // row stride of staged W / accumulator [k][col]
  const int u0 = blockIdx.x * UB_;
  if (u0 >= nu) return;
  const int nb = win_nb(n);
  const double* L = tt + w.off_tt;
  const int ld = w.ld_t;
  double* W = ut + w.off_ut;
  const int ldu = w.ld_u;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wr = warp >> 2, wc = warp & 3;       // row half / column quarter of this warp
  const int row0 = 32 * wr + gid;                // + 8 mt: rows of this thread's A / C fragments
  const int col0 = 8 * NTW * wc;                      // + 8 nt (+ gid for B fragments, + 2 tig for C)

  extern __shared__ __align__(16) double sm[];
  double* LsBuf = sm;                        // 2 x [KC][LSS]
  double* WsBuf = sm + 2 * KC * LSS;         // 2 x [KC][WSS]
  double* Ds = LsBuf;                        // [64][LSS] inv(L_ii)(r, kk) at [kk][r], aliases both L chunks
  double* Ts = WsBuf;                        // [64][WSS] accumulator tile, aliases both W chunks
  double* red = WsBuf;                       // [2][2][UB_] final reductions
  double* ys = sm + 2 * (KC * LSS + KC * WSS);  // [nb*64] y = L^-1 Z1 solved so far (every CTA carries this extra
  double* rys = ys + nb * NB;                   // [64]     right-hand side column itself: ~1/128 more work, no
                                                //          separate latency-bound kernel and no y round trip)
  double* ypart = rys + NB;                     // [4][64]  partial dot products of the y column (4 threads per row)
  const int yr_ = tid & 63, yq = tid >> 6;
  double p_info[NTW][2], p_z[NTW][2];            // per owned column (nt, e): partial sums over this thread's rows
#pragma unroll
  for (int nt = 0; nt < NTW; nt++) p_info[nt][0] = p_info[nt][1] = p_z[nt][0] = p_z[nt][1] = 0.0;
  const int cvalid = ldu - u0;               // columns that exist in the row (ldu is a multiple of 8)

  // tri: the right-hand side is the identity (explicit L^-1 for the int8-split solve): block rows above this CTA's
  // first column hold zeros and stay zero, and the block columns of W left of it contribute nothing
  const int ib0 = tri ? u0 / NB : 0;
  double vmax = 0.0;   // tri: max |L^-1| of this CTA's columns (the int8-split solve scales its digit planes by it)
  for (int ib = ib0; ib < nb; ib++) {
    const int i0 = ib * NB;
    double C[4][NTW][2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
      const int r = i0 + row0 + 8 * mt;
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) {
        const int c = col0 + 8 * nt + 2 * tig;
        double2 v = make_double2(0.0, 0.0);
        if (r < n && c < cvalid) v = *reinterpret_cast<const double2*>(&W[(long long)r * ldu + u0 + c]);
        C[mt][nt][0] = v.x;
        C[mt][nt][1] = v.y;
      }
    }
    // C -= L(ib, 0:i0) * W(0:i0) in chunks of KC k-steps
    const int nchunk = ib * (NB / KC);
    auto issue = [&](int ch, int buf) {
      const int kbase = ch * KC;
      double* ls = LsBuf + buf * KC * LSS;
      double* ws = WsBuf + buf * KC * WSS;
      // L chunk: KC columns x 64 rows in 16-byte pieces, 32 per column
#pragma unroll
      for (int it = 0; it < (KC * NB / 2) / 256; it++) {
        const int idx = tid + it * 256;
        const int kk = idx >> 5, r2 = (idx & 31) * 2;
        // row n of an odd-sized window is padding nothing ever wrote: it must read as zero, not as stale memory (a NaN
        // there would reach the valid rows through 0 * NaN in the inv(L_ii) product)
        cp_async16(ls + kk * LSS + r2, L + (long long)(kbase + kk) * ld + i0 + r2, n - (i0 + r2));
      }
      // W chunk: KC rows x 128 columns, 64 pieces per row
#pragma unroll
      for (int it = 0; it < (KC * UB_ / 2) / 256; it++) {
        const int idx = tid + it * 256;
        const int kk = idx / (UB_ / 2), c2 = (idx % (UB_ / 2)) * 2;
        cp_async16(ws + kk * WSS + c2, W + (long long)(kbase + kk) * ldu + u0 + c2, c2 < cvalid ? 2 : 0);
      }
      cp_async_commit();
    };
    double acc_y = 0.0;  // thread (row yr_, quarter yq): sum over k = yq mod 4 of L(i0 + yr_, k) y_k
    __syncthreads();  // previous row block finished with the aliased buffers (Ts / Ds)
    // (the y column of a tri launch is exact only in the CTA of column block 0, the one that writes y_out)
    const int ch0 = ib0 * (NB / KC);
    if (nchunk > ch0) issue(ch0, ch0 & 1);
    for (int ch = ch0; ch < nchunk; ch++) {
      const int buf = ch & 1;
      if (ch + 1 < nchunk) {
        issue(ch + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const double* ls = LsBuf + buf * KC * LSS;
      const double* ws = WsBuf + buf * KC * WSS;
#pragma unroll 2
      for (int ks = 0; ks < KC / 4; ks++) {
        double a[4], b[4];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = -ls[(4 * ks + tig) * LSS + row0 + 8 * mt];
#pragma unroll
        for (int nt = 0; nt < NTW; nt++) b[nt] = ws[(4 * ks + tig) * WSS + col0 + 8 * nt + gid];
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
          for (int nt = 0; nt < NTW; nt++) dmma_m8n8k4(C[mt][nt][0], C[mt][nt][1], a[mt], b[nt]);
      }
      {
        const double* yk = ys + ch * KC;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) acc_y = fma(ls[(kk + yq) * LSS + yr_], yk[kk + yq], acc_y);
      }
      __syncthreads();  // buffer `buf` may be refilled by the next issue
    }
    // W_i = inv(L_ii) * C
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < NTW; nt++)
        *reinterpret_cast<double2*>(&Ts[(row0 + 8 * mt) * WSS + col0 + 8 * nt + 2 * tig]) =
            make_double2(C[mt][nt][0], C[mt][nt][1]);
    for (int idx = tid; idx < NB * NB / 2; idx += 256) {  // Ds[kk*LSS + r] = inv(L_ii)(r, kk), zero above diag
      const int kk = idx >> 5, r2 = (idx & 31) * 2;
      *reinterpret_cast<double2*>(&Ds[kk * LSS + r2]) =
          reinterpret_cast<const double2*>(dinv + w.off_dinv + (long long)ib * NB * NB)[idx];
    }
    ypart[yq * NB + yr_] = acc_y;
    __syncthreads();
    if (tid < NB)
      rys[tid] = (i0 + tid < n)
                     ? zt[w.off_t + i0 + tid] - ((ypart[tid] + ypart[NB + tid]) + (ypart[2 * NB + tid] + ypart[3 * NB + tid]))
                     : 0.0;
    __syncthreads();
    {  // y_i = inv(L_ii) (z_i - sum_k L_ik y_k): 4 threads per row, combined after the tile product below
      double v = 0.0;
      for (int j = yq; j <= yr_; j += 4) v = fma(Ds[j * LSS + yr_], rys[j], v);
      ypart[yq * NB + yr_] = v;
    }
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) C[mt][nt][0] = C[mt][nt][1] = 0.0;
    const int ksmax = 8 * (wr + 1);  // inv(L_ii)(r, kk) == 0 for kk > r, and this warp's rows end at 32 wr + 31
#pragma unroll 2
    for (int ks = 0; ks < ksmax; ks++) {
      double a[4], b[4];
#pragma unroll
      for (int mt = 0; mt < 4; mt++) a[mt] = Ds[(4 * ks + tig) * LSS + row0 + 8 * mt];
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) b[nt] = Ts[(4 * ks + tig) * WSS + col0 + 8 * nt + gid];
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++) dmma_m8n8k4(C[mt][nt][0], C[mt][nt][1], a[mt], b[nt]);
    }
    __syncthreads();
    if (tid < NB) ys[i0 + tid] = (ypart[tid] + ypart[NB + tid]) + (ypart[2 * NB + tid] + ypart[3 * NB + tid]);
    __syncthreads();  // ys[i0 .. i0+63] is complete
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
      const int r = i0 + row0 + 8 * mt;
      if (r >= n) continue;
      const double yr = ys[r];
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) {
        const int c = col0 + 8 * nt + 2 * tig;
        if (c < cvalid)
          *reinterpret_cast<double2*>(&W[(long long)r * ldu + u0 + c]) = make_double2(C[mt][nt][0], C[mt][nt][1]);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const double v = C[mt][nt][e];
          if (c + e < cvalid) vmax = fmax(vmax, fabs(v));
          p_info[nt][e] = fma(v, v, p_info[nt][e]);
          p_z[nt][e] = fma(yr, v, p_z[nt][e]);
        }
      }
    }
  }
  // column reductions in a fixed order: over the 8 row groups of a warp (shuffles), then over the
  // two row-half warps through shared memory
#pragma unroll
  for (int nt = 0; nt < NTW; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        p_info[nt][e] += __shfl_xor_sync(0xffffffffu, p_info[nt][e], o);
        p_z[nt][e] += __shfl_xor_sync(0xffffffffu, p_z[nt][e], o);
      }
  __syncthreads();
  if (gid == 0) {
#pragma unroll
    for (int nt = 0; nt < NTW; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = col0 + 8 * nt + 2 * tig + e;
        red[(0 * 2 + wr) * UB_ + c] = p_info[nt][e];
        red[(1 * 2 + wr) * UB_ + c] = p_z[nt][e];
      }
  }
  __syncthreads();
  if (amax_out) {   // bits of a non-negative double order like unsigned integers
    for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && vmax > 0.0) atomicMax(&amax_out[blockIdx.y], (unsigned long long)__double_as_longlong(vmax));
  }
  if (y_out && blockIdx.x == 0)   // qcat: y = L^-1 Z1 is an output too (every CTA carries the same column)
    for (int i = tid; i < n; i += 256) y_out[w.off_t + i] = ys[i];
  if (tid < UB_ && u0 + tid < nu) {
    const double s_info = red[0 * UB_ + tid] + red[1 * UB_ + tid];
    const double s_z = red[2 * UB_ + tid] + red[3 * UB_ + tid];
    const double inf = fabs(s_info);                 // info = |b21 B11^-1 b12|      (dist.cpp:198)
    zu[w.off_u + u0 + tid] = s_z / sqrt(inf);        // z / sqrt(info)               (dist.cpp:200)
    info[w.off_u + u0 + tid] = inf;
  }
}

// ---------------------------------------------------------------------------------------------
// Row block j of X = L^-1 (the right operand of the int8-split solve), launched right after diagonal block j has been
// factored, on a stream of its own, so that the explicit inverse grows in the shadow of the factorisation's own steps:
//   X_jj = inv(L_jj)  (published by the factorisation as dinv),
//   X_ji = -inv(L_jj) sum_{k=i}^{j-1} L_jk X_ki   (i < j),         y_j = inv(L_jj) (z_j - sum_{k<j} L_jk y_k).
// Row j of L is final once step j - 1 has run; rows < j of X come from the earlier launches of this kernel.  One CTA
// per 64 x 64 tile: grid (j + 1, windows); CTA i < j forms X_ji, CTA j copies X_jj and solves the y block.  X is
// row-major n_t x ld_t like B11's buffer; blocks above the diagonal are never written (nothing reads them).  Compared
// with the block-sequential triangular solve on identity columns (one long CTA per 64 columns) this is nb^2 / 2 short
// uniform CTAs per window: a quarter of the SM-time, and no latency of its own.
__global__ void __launch_bounds__(256)
linv_row_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ tt, const double* __restrict__ dinv, double* xmat,
                const double* __restrict__ zt, double* y, unsigned long long* amax, int j) {
  const SolveWin w = wins[blockIdx.y];
  const int n = w.n_t;
  const int i = blockIdx.x;
  if (j >= win_nb(n) || i > j) return;
  const double* L = tt + w.off_tt;
  double* X = xmat + w.off_tt;
  const int ld = w.ld_t;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = j * NB, i0 = i * NB;
  const double* D = dinv + w.off_dinv + (long long)j * NB * NB;   // D[c * NB + r] = inv(L_jj)(r, c), zero above the diagonal
  extern __shared__ __align__(16) double sm[];
  double vmax = 0.0;
  if (i == j) {
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int c = idx >> 6, r = idx & 63;     // consecutive threads: consecutive r (dinv) ...
      sm[r * TS + c] = D[idx];
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int r = idx >> 6, c = idx & 63;     // ... consecutive c (X is row-major)
      if (j0 + r < n && j0 + c < n) {
        const double v = sm[r * TS + c];
        X[(long long)(j0 + r) * ld + j0 + c] = v;
        vmax = fmax(vmax, fabs(v));
      }
    }
    // y_j: 4 threads per row gather sum_{col < j0} L(j0 + r, col) y_col, then the 64 x 64 triangular product
    double* part = sm + NB * TS;     // [4][64]
    double* rhs = part + 4 * NB;     // [64]
    {
      const int r = tid & 63, q = tid >> 6;
      double acc = 0.0;
      if (j0 + r < n)
        for (int col = q; col < j0; col += 4) acc = fma(L[(long long)col * ld + j0 + r], y[w.off_t + col], acc);
      part[q * NB + r] = acc;
    }
    __syncthreads();
    if (tid < NB)
      rhs[tid] = (j0 + tid < n) ? zt[w.off_t + j0 + tid] - ((part[tid] + part[NB + tid]) + (part[2 * NB + tid] + part[3 * NB + tid])) : 0.0;
    __syncthreads();
    if (tid < NB && j0 + tid < n) {
      double v = 0.0;
      for (int q = 0; q <= tid; q++) v = fma(sm[tid * TS + q], rhs[q], v);
      y[w.off_t + j0 + tid] = v;
    }
  } else {
    double* Ls = sm;             // Ls[kk * TS + r] = L(j0 + r, k0 + kk)
    double* Xs = sm + NB * TS;   // Xs[kk * TS + c] = X(k0 + kk, i0 + c)
    double C[4][2][2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) C[mt][nt][0] = C[mt][nt][1] = 0.0;
    // the tiles of step k + 1 travel in registers while step k is multiplied (the launch is a chain of dependent
    // launches: its loads' latency is all it has to hide)
    double pl[16], px[16];
    auto fetch = [&](int k) {
      const int k0 = k * NB;
#pragma unroll
      for (int q = 0; q < 16; q++) {
        const int idx = tid + 256 * q, kk = idx >> 6, r = idx & 63;
        pl[q] = (j0 + r < n) ? L[(long long)(k0 + kk) * ld + j0 + r] : 0.0;
        // X_ii is lower triangular and its upper part was never written: read it as the zeros it stands for
        px[q] = (k > i || r <= kk) ? X[(long long)(k0 + kk) * ld + i0 + r] : 0.0;
      }
    };
    fetch(i);
    for (int k = i; k < j; k++) {
#pragma unroll
      for (int q = 0; q < 16; q++) {
        const int idx = tid + 256 * q, kk = idx >> 6, r = idx & 63;
        Ls[kk * TS + r] = pl[q];
        Xs[kk * TS + r] = px[q];
      }
      __syncthreads();
      if (k + 1 < j) fetch(k + 1);
      tile64_dmma(Ls, Xs, C, NB, 1.0, lane, warp);
      __syncthreads();
    }
    // X_ji = -inv(L_jj) C
    const int gid = lane >> 2, tig = lane & 3;
    double* Cs = Xs;   // Cs[q * TS + c] = C(q, c)
    double* Ds = Ls;   // Ds[q * TS + r] = inv(L_jj)(r, q)
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++)
          Cs[(32 * (warp >> 2) + 8 * mt + gid) * TS + 16 * (warp & 3) + 8 * nt + 2 * tig + e] = C[mt][nt][e];
    for (int idx = tid; idx < NB * NB; idx += 256) Ds[(idx >> 6) * TS + (idx & 63)] = D[idx];
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) C[mt][nt][0] = C[mt][nt][1] = 0.0;
    tile64_dmma(Ds, Cs, C, NB, -1.0, lane, warp);
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
      const int r = j0 + 32 * (warp >> 2) + 8 * mt + gid;
      if (r >= n) continue;
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        const int c = i0 + 16 * (warp & 3) + 8 * nt + 2 * tig;
        *reinterpret_cast<double2*>(&X[(long long)r * ld + c]) = make_double2(C[mt][nt][0], C[mt][nt][1]);
        vmax = fmax(vmax, fmax(fabs(C[mt][nt][0]), fabs(C[mt][nt][1])));
      }
    }
  }
  if (amax) {   // bits of a non-negative double order like unsigned integers
    for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && vmax > 0.0) atomicMax(&amax[blockIdx.y], (unsigned long long)__double_as_longlong(vmax));
  }
}

// ---------------------------------------------------------------------------------------------
// MakePosDef / CountPC for windows the Cholesky certificate cannot vouch for (util.cpp:302-318, 355-388): a symmetric
// eigendecomposition of B11 on the device.  Rare path (lambda = 0 with duplicated SNPs, an eigenvalue cut-off above the
// ridge, negative weights), so it is built for robustness, not speed: cyclic two-sided Jacobi, one CTA per window, the
// rotations of one round-robin step (n / 2 disjoint pairs) applied as independent 2 x 2 block updates
//   A'[k][l] = J_k^T A[k][l] J_l,   V'[:, l] = V[:, l] J_l
// so every element is read and written exactly once per step.  G (n x ld, a full symmetric copy of B11) is
// diagonalised in place, V accumulates the eigenvectors, B11 itself stays untouched until the clip:
//   B11 <- B11 + sum_{lambda_i < min_abs_eig} (min_abs_eig - lambda_i) v_i v_i^T      (= V max(Lambda, min) V^T)
// which is a no-op when nothing lies below the threshold, exactly like the reference's branch.
__global__ void __launch_bounds__(512)
eig_jacobi_kernel(const SolveWin* __restrict__ wins, double* tt, double* G_all, double* V_all, double* evals,
                  double min_abs_eig, int clip, int* n_clipped) {
  extern __shared__ __align__(16) unsigned char eig_smem[];
  const SolveWin w = wins[blockIdx.x];
  const int n = w.n_t, ld = w.ld_t;
  const int m = (n + 1) & ~1, half = m >> 1;
  int* pp = reinterpret_cast<int*>(eig_smem);          // [half] pair members of the current step
  int* qq = pp + half;
  double* cs = reinterpret_cast<double*>(qq + half + (half & 1));   // [half] cosines, [half] sines
  double* sn = cs + half;
  __shared__ double red[512];
  __shared__ int stop;
  double* A = tt + w.off_tt;
  double* G = G_all + w.off_tt;
  double* V = V_all + w.off_tt;
  const int tid = threadIdx.x, nth = blockDim.x;
  for (long long idx = tid; idx < (long long)n * n; idx += nth) {
    const int j = (int)(idx / n), i = (int)(idx % n);
    G[(long long)j * ld + i] = i >= j ? A[(long long)j * ld + i] : A[(long long)i * ld + j];   // lower triangle is the valid one
    V[(long long)j * ld + i] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int sweep = 0; sweep < 40; sweep++) {
    double off = 0.0, tot = 0.0;
    for (long long idx = tid; idx < (long long)n * n; idx += nth) {
      const int j = (int)(idx / n), i = (int)(idx % n);
      const double v = G[(long long)j * ld + i];
      tot += v * v;
      if (i != j) off += v * v;
    }
    red[tid] = off;
    __syncthreads();
    for (int o = nth >> 1; o > 0; o >>= 1) {
      if (tid < o) red[tid] += red[tid + o];
      __syncthreads();
    }
    off = red[0];
    __syncthreads();
    red[tid] = tot;
    __syncthreads();
    for (int o = nth >> 1; o > 0; o >>= 1) {
      if (tid < o) red[tid] += red[tid + o];
      __syncthreads();
    }
    tot = red[0];
    if (tid == 0) stop = !(off > 1e-30 * tot);   // also stops on NaN
    __syncthreads();
    if (stop) break;
    for (int step = 0; step < m - 1; step++) {
      // round-robin pairing: player 0 stays, the others rotate
      for (int k = tid; k < half; k += nth) {
        const int a = k == 0 ? 0 : 1 + (k - 1 + step) % (m - 1);
        const int bidx = m - 1 - k;
        const int b = 1 + (bidx - 1 + step) % (m - 1);
        const int p = min(a, b), q = max(a, b);
        pp[k] = p;
        qq[k] = q;
        double c = 1.0, sgn = 0.0;
        if (q < n) {
          const double apq = G[(long long)q * ld + p];
          if (apq != 0.0) {
            const double app = G[(long long)p * ld + p], aqq = G[(long long)q * ld + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            sgn = t * c;
          }
        }
        cs[k] = c;
        sn[k] = sgn;
      }
      __syncthreads();
      for (int blk = tid; blk < half * half; blk += nth) {
        const int k = blk % half, l = blk / half;            // rows of pair k, columns of pair l
        const int pk = pp[k], qk = qq[k], pl = pp[l], ql = qq[l];
        const bool rq = qk < n, cq = ql < n;                 // the dummy player of an odd n
        const double ck = cs[k], sk = sn[k], cl = cs[l], sl = sn[l];
        double a00 = G[(long long)pl * ld + pk];
        double a01 = cq ? G[(long long)ql * ld + pk] : 0.0;
        double a10 = rq ? G[(long long)pl * ld + qk] : 0.0;
        double a11 = (rq && cq) ? G[(long long)ql * ld + qk] : 0.0;
        // columns: (x_p, x_q) <- (c x_p - s x_q, s x_p + c x_q)
        double b00 = cl * a00 - sl * a01, b01 = sl * a00 + cl * a01;
        double b10 = cl * a10 - sl * a11, b11 = sl * a10 + cl * a11;
        // rows, same rotation of pair k
        a00 = ck * b00 - sk * b10;
        a10 = sk * b00 + ck * b10;
        a01 = ck * b01 - sk * b11;
        a11 = sk * b01 + ck * b11;
        if (k == l && rq) a01 = a10 = 0.0;                   // the annihilated element, exactly
        G[(long long)pl * ld + pk] = a00;
        if (cq) G[(long long)ql * ld + pk] = a01;
        if (rq) G[(long long)pl * ld + qk] = a10;
        if (rq && cq) G[(long long)ql * ld + qk] = a11;
      }
      for (long long idx = tid; idx < (long long)n * half; idx += nth) {
        const int l = (int)(idx / n), r = (int)(idx % n);
        const int pl = pp[l], ql = qq[l];
        if (ql >= n) continue;
        const double cl = cs[l], sl = sn[l];
        const double vp = V[(long long)pl * ld + r], vq = V[(long long)ql * ld + r];
        V[(long long)pl * ld + r] = cl * vp - sl * vq;
        V[(long long)ql * ld + r] = sl * vp + cl * vq;
      }
      __syncthreads();
    }
  }
  int clipped = 0;
  for (int i = tid; i < n; i += nth) {
    const double ev = G[(long long)i * ld + i];
    evals[w.off_t + i] = ev;
    clipped += ev < min_abs_eig;
  }
  if (clip) {
    // low-rank correction of the ORIGINAL matrix, eigenvector by eigenvector (usually none or a handful)
    for (int i = 0; i < n; i++) {
      const double ev = G[(long long)i * ld + i];
      if (!(ev < min_abs_eig)) continue;
      const double d = min_abs_eig - ev;
      const double* v = V + (long long)i * ld;
      for (long long idx = tid; idx < (long long)n * n; idx += nth) {
        const int c = (int)(idx / n), r = (int)(idx % n);
        if (r >= c) A[(long long)c * ld + r] += d * v[r] * v[c];
      }
      __syncthreads();
    }
  }
  if (n_clipped && clipped) atomicAdd(&n_clipped[blockIdx.x], clipped);
}

// qcat: the tested measured SNPs ride along as extra right-hand-side columns; their correlation row is a row
// of B11, whose own entry is the forced diagonal 1 + lambda (qcat.cpp:186), not the computed self-correlation.
__global__ void qcat_patch_kernel(const SolveWin* __restrict__ wins, double* ut, int n_u, int core_first, int n_core,
                                  double diag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_core) return;
  const SolveWin w = wins[0];
  ut[w.off_ut + (long long)(core_first + c) * w.ld_u + n_u + c] = diag;
}

// qcat: per tested SNP u, Pearson correlation (util.cpp:193-202, two passes like Eigen's mean-then-centre) of
// y = L^-1 Z1 and W(:, u) = L^-1 b_u over the measured SNPs; qcat_t = sqrt(num_eig - 3) r, qcat_chisq =
// (num_eig - 3) r^2 (qcat.cpp:224-229).  One thread per column: W rows are read coalesced.
__global__ void __launch_bounds__(128)
qcat_finalize_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ ut, const double* __restrict__ y,
                     int num_eig, double* __restrict__ qt, double* __restrict__ qchisq) {
  const SolveWin w = wins[0];
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= w.n_u) return;
  const int n = w.n_t;
  const double* W = ut + w.off_ut + u;
  const double* yy = y + w.off_t;
  double sy = 0.0, sw = 0.0;
  for (int k = 0; k < n; k++) {
    sy += yy[k];
    sw += W[(long long)k * w.ld_u];
  }
  const double my = sy / n, mw = sw / n;
  double sxx = 0.0, syy = 0.0, sxy = 0.0;
  for (int k = 0; k < n; k++) {
    const double dy = yy[k] - my, dw = W[(long long)k * w.ld_u] - mw;
    sxx = fma(dy, dy, sxx);
    syy = fma(dw, dw, syy);
    sxy = fma(dy, dw, sxy);
  }
  const double r = sxy / sqrt(sxx * syy);
  const double dof = (double)(num_eig - 3);
  qt[u] = sqrt(dof) * r;
  qchisq[u] = dof * r * r;
}

}  // namespace

static int linv_attr(Ctx* ctx) {
  static bool set_dev[64] = {};
  if (!set_dev[ctx->device & 63]) {
    GB_CUDA(cudaFuncSetAttribute(linv_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 2 * NB * TS)));
    set_dev[ctx->device & 63] = true;
  }
  return GB_OK;
}

// Rows [j_lo, j_hi) of X = L^-1 (and of y) on `stream`, one launch per row block.
static int launch_linv_rows_on(Ctx* ctx, cudaStream_t stream, const LinvArgs& a, int j_lo, int j_hi) {
  int rc = linv_attr(ctx);
  if (rc) return rc;
  for (int j = j_lo; j < j_hi; j++) {
    linv_row_kernel<<<dim3((unsigned)(j + 1), (unsigned)a.n_real), 256, sizeof(double) * 2 * NB * TS, stream>>>(
        a.d_wins, a.d_tt, a.d_dinv, a.d_x, a.d_zt, a.d_y, a.d_amax, j);
    ctx->launches++;
  }
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

int launch_linv_rows(Ctx* ctx, const LinvArgs& a, int max_nt) {
  if (a.n_real == 0) return GB_OK;
  return launch_linv_rows_on(ctx, ctx->stream, a, 0, (max_nt + NB - 1) / NB);
}

int launch_cholesky(Ctx* ctx, const SolveWin* d_wins, int n_wins, int max_nt, double* d_tt, double* d_dinv,
                    int* d_status, const int* d_skip, const LinvArgs* linv) {
  if (n_wins == 0) return GB_OK;
  const int nb_max = (max_nt + NB - 1) / NB;
  // the explicit inverse of the int8-split solve: row block j on an auxiliary stream as soon as diagonal block j exists
  if (linv && linv->n_real > 0) {
    if (!ctx->aux_stream) {
      int lo = 0, hi = 0;
      GB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      GB_CUDA(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi));
      GB_CUDA(cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
    }
  } else {
    linv = nullptr;
  }
  auto linv_row_after = [&](int j) -> int {   // diagonal block j has just been enqueued on ctx->stream
    if (!linv) return GB_OK;
    GB_CUDA(cudaEventRecord(ctx->ev_aux, ctx->stream));
    GB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_aux, 0));
    return launch_linv_rows_on(ctx, ctx->aux_stream, *linv, j, j + 1);
  };
  const size_t smem_panel = sizeof(double) * 2 * NB * TS;
  const size_t smem_diag = sizeof(double) * (2 * NB * MP + 3 * 16 * 17 + 32);
  const size_t smem_update = std::max(sizeof(double) * 2 * NB * TS, smem_diag);   // the update CTA of tile (k+1, k+1) also factors it
  static bool attr_set_dev[64] = {};
  bool& attr_set = attr_set_dev[ctx->device & 63];
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_diag));
    GB_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel));
    GB_CUDA(cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_update));
    attr_set = true;
  }
  // diagonal block 0 by its own launch; every later one is factored inside the update launch that completes it
  const bool trace = getenv("GB_CHOL_TRACE") != nullptr;   // diagnostics: event-timed kernels, printed per call
  std::vector<cudaEvent_t> ev;
  auto mark = [&]() {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, ctx->stream);
    ev.push_back(e);
  };
  mark();
  chol_diag_kernel<<<n_wins, 256, smem_diag, ctx->stream>>>(d_wins, d_tt, d_dinv, d_status, d_skip, 0);
  ctx->launches++;
  mark();
  int rc_l = linv_row_after(0);
  if (rc_l) return rc_l;
  for (int k = 0; k + 1 < nb_max; k++) {
    const int tb = nb_max - k - 1;
    chol_panel_kernel<<<dim3(tb, n_wins), 256, smem_panel, ctx->stream>>>(d_wins, d_tt, d_dinv, d_skip, k);
    mark();
    chol_update_kernel<<<dim3(tb, n_wins), 256, smem_update, ctx->stream>>>(d_wins, d_tt, d_dinv, d_status,
                                                                                         d_skip, k);
    mark();
    ctx->launches += 2;
    if ((rc_l = linv_row_after(k + 1))) return rc_l;   // the update launch has factored diagonal block k + 1
  }
  if (linv) {   // join: whatever follows on ctx->stream needs all of X and y
    GB_CUDA(cudaEventRecord(ctx->ev_aux, ctx->aux_stream));
    GB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_aux, 0));
  }
  if (trace) {
    cudaStreamSynchronize(ctx->stream);
    fprintf(stderr, "[chol trace] diag0 ");
    for (size_t i = 0; i + 1 < ev.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      fprintf(stderr, "%s%.1f", i == 0 ? "" : (i % 2 ? " | P " : " U "), ms * 1e3);
    }
    fprintf(stderr, " (us)\n");
    for (auto e : ev) cudaEventDestroy(e);
  }
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

int launch_trsm_finalize(Ctx* ctx, const SolveWin* d_wins, int n_wins, int max_nt, int max_nu, const double* d_tt,
                         const double* d_dinv, double* d_ut, const double* d_zt, double* d_zu, double* d_info,
                         double* d_y_out, int tri, unsigned long long* d_amax) {
  if (n_wins == 0 || max_nu == 0) return GB_OK;
  const int nb_max = (max_nt + NB - 1) / NB;
  // 128 columns per CTA when the launch has waves to spare, 64 when it is about one wave (a single window, the
  // chromosome driver's quarter batches): the longest CTA then sets the kernel's duration, and it is half as long
  const long long ctas128 = (long long)((max_nu + UB - 1) / UB) * n_wins;
  bool narrow = ctas128 < 3LL * ctx->sm_count;     // 2 CTAs per SM: below 1.5 waves
  if (tri) narrow = true;   // L^-1 runs in the latency-bound factorisation lane: shorter CTAs, and the zero-block skipping is finer
  if (const char* e = getenv("GB_TRSM_NARROW")) narrow = atoi(e) != 0;   // tuning knob
  const int ub = narrow ? 64 : UB;
  const size_t smem = tr_smem_bytes(ub) + sizeof(double) * (size_t)(nb_max * NB + 5 * NB);
  if (smem > 227 * 1024) {
    ctx->err = "window has too many measured SNPs for trsm_finalize_kernel";
    return GB_ERR_UNSUPPORTED;
  }
  const dim3 grid((unsigned)((max_nu + ub - 1) / ub), (unsigned)n_wins);
  // the opt-in shared-memory limit is a per-device function attribute: raise it only when a launch needs more than
  // any earlier one did
  static size_t attr_smem[64][2] = {};
  size_t& have = attr_smem[ctx->device & 63][narrow ? 1 : 0];
  if (smem > have) {
    if (narrow) GB_CUDA(cudaFuncSetAttribute(trsm_finalize_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else GB_CUDA(cudaFuncSetAttribute(trsm_finalize_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
  if (narrow) trsm_finalize_kernel<64><<<grid, 256, smem, ctx->stream>>>(d_wins, d_tt, d_dinv, d_ut, d_zt, d_zu, d_info, d_y_out, tri, d_amax);
  else trsm_finalize_kernel<128><<<grid, 256, smem, ctx->stream>>>(d_wins, d_tt, d_dinv, d_ut, d_zt, d_zu, d_info, d_y_out, tri, d_amax);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

// G / V: workspaces laid out like the B11 buffer (tt_elems doubles each); evals per measured SNP (off_t); n_clipped
// (optional) zero-initialised ints, one per window.
int launch_eig_jacobi(Ctx* ctx, const SolveWin* d_wins, int n_wins, int max_nt, double* d_tt, double* d_G, double* d_V,
                      double* d_evals, double min_abs_eig, int clip, int* d_n_clipped) {
  if (n_wins == 0) return GB_OK;
  const int half = (max_nt + 1) / 2;
  const size_t smem = sizeof(int) * (size_t)(2 * half + 2) + sizeof(double) * (size_t)(2 * half) + 16;
  if (smem > 200 * 1024) {
    ctx->err = "window too large for the eigen-clip path";
    return GB_ERR_UNSUPPORTED;
  }
  GB_CUDA(cudaFuncSetAttribute(eig_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eig_jacobi_kernel<<<(unsigned)n_wins, 512, smem, ctx->stream>>>(d_wins, d_tt, d_G, d_V, d_evals, min_abs_eig, clip, d_n_clipped);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_qcat_patch(Ctx* ctx, const SolveWin* d_wins, double* d_ut, int n_u, int core_first, int n_core, double diag) {
  if (n_core <= 0) return GB_OK;
  qcat_patch_kernel<<<(unsigned)((n_core + 127) / 128), 128, 0, ctx->stream>>>(d_wins, d_ut, n_u, core_first, n_core, diag);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_qcat_finalize(Ctx* ctx, const SolveWin* d_wins, const double* d_ut, const double* d_y, int n_tested,
                         int num_eig, double* d_qt, double* d_qchisq) {
  if (n_tested <= 0) return GB_OK;
  qcat_finalize_kernel<<<(unsigned)((n_tested + 127) / 128), 128, 0, ctx->stream>>>(d_wins, d_ut, d_y, num_eig, d_qt, d_qchisq);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_pd_bound(Ctx* ctx, const SolveWin* d_wins, int n_real, const double* d_rq_t, double lambda,
                    double gneg, double min_abs_eig, int* d_skip) {
  if (n_real == 0) return GB_OK;
  pd_bound_kernel<<<n_real, 256, 0, ctx->stream>>>(d_wins, n_real, d_rq_t, lambda, gneg, min_abs_eig, d_skip);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_copy_shift(Ctx* ctx, const SolveWin* d_wins, int n_wins, const double* d_src, double* d_dst,
                      double shift, const int* d_skip) {
  if (n_wins == 0) return GB_OK;
  copy_shift_kernel<<<dim3(64, n_wins), 256, 0, ctx->stream>>>(d_wins, d_src, d_dst, shift, d_skip, n_wins);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

}  // namespace gb
