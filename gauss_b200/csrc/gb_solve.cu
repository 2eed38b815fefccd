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

// ---------------------------------------------------------------------------------------------
// In-place transform of a 64x64 SPD block M (lower triangle, M[c*PAD + r] = element (r, c)) into
// inv(L), L = chol(M); L itself is never formed (nothing downstream needs it).  256 threads:
// thread (r = tid & 63, q = tid >> 6) owns row r, columns c = q, q+4, ...
// Step j of the right-looking factorisation and step j of the forward substitution L Y = I share
// one pass: once column j of A has been consumed its slot holds column j of Y.
//   rs = 1/sqrt(a_jj);  l_rj = a_rj rs
//   row j:   Y(j,c) *= rs (c<j),  Y(j,j) = rs
//   rows r>j: Y(r,c) -= l_rj Y(j,c) (c<j),  Y(r,j) = -l_rj rs,  A(r,c) -= l_rj l_cj (j<c<=r)
// Two barriers per step (read phase / write phase); ~250 cycles per pivot.
// Returns false on a non-positive / NaN pivot (flag only; a substitute pivot keeps it finite).
constexpr int MP = NB + 2;  // even stride: double2-aligned rows for the GEMM that follows

__device__ bool spd_block_to_inv_chol(double* M) {
  const int tid = threadIdx.x;
  const int r = tid & 63;
  const int q = tid >> 6;
  bool bad_any = false;
  for (int j = 0; j < NB; j++) {
    __syncthreads();
    const double piv = M[j * MP + j];
    const double a_rj = M[j * MP + r];
    double m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int c = q + 4 * i;
      m[i] = (c < j) ? M[c * MP + j] : M[j * MP + c];  // Y(j,c) for c<j, a_cj for c>j
    }
    __syncthreads();
    const bool bad = !(piv > 0.0);
    bad_any |= bad;
    const double rs = rsqrt(bad ? 1.0 : piv);
    if (r > j) {
      const double l = a_rj * rs;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int c = q + 4 * i;
        if (c == j) M[j * MP + r] = -l * rs;
        else if (c < j || c <= r) M[c * MP + r] = fma(-l, m[i] * rs, M[c * MP + r]);
      }
    } else if (r == j) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int c = q + 4 * i;
        if (c < j) M[c * MP + j] = m[i] * rs;
        else if (c == j) M[j * MP + j] = rs;
      }
    }
  }
  __syncthreads();
  return !bad_any;
}

// Warp-synchronous 32x32 version of the same transform: lane r keeps row r of A (lower) and row r
// of Y = inv(L) in registers; column j of L and row j of Y travel through two 32-double shared
// buffers.  No block barrier inside, ~250 clocks per pivot.  A at Ms[(o+c)*MP + o+r]; inv(L) is
// written (lower triangle, zeros above) to Xs at the same coordinates.
__device__ bool invchol32_warp(const double* Ms, double* Xs, int o, double* colbuf, double* rowbuf, int lane) {
  double a[32], y[32];
#pragma unroll
  for (int c = 0; c < 32; c++) {
    a[c] = (c <= lane) ? Ms[(o + c) * MP + o + lane] : 0.0;
    y[c] = 0.0;
  }
  bool bad_any = false;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const double piv = __shfl_sync(0xffffffffu, a[j], j);
    const bool bad = !(piv > 0.0);
    bad_any |= bad;
    const double rs = rsqrt(bad ? 1.0 : piv);
    const double l = a[j] * rs;                 // lanes r > j: L(r, j)
    const bool is_j = lane == j;
#pragma unroll
    for (int c = 0; c < j; c++) y[c] = is_j ? y[c] * rs : y[c];   // Y(j, c) *= rs
    if (is_j) y[j] = rs;                                          // Y(j, j) = rs
    colbuf[lane] = l;
    if (is_j) {
#pragma unroll
      for (int c = 0; c <= j; c++) rowbuf[c] = y[c];
    }
    __syncwarp();
    const double lm = (lane > j) ? -l : 0.0;
#pragma unroll
    for (int c = j + 1; c < 32; c++) a[c] = fma(lm, colbuf[c], a[c]);   // A(r, c) -= L(r, j) L(c, j)
#pragma unroll
    for (int c = 0; c <= j; c++) y[c] = fma(lm, rowbuf[c], y[c]);       // Y(r, c) -= L(r, j) Y(j, c)
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < 32; c++) Xs[(o + c) * MP + o + lane] = (c <= lane) ? y[c] : 0.0;
  return !bad_any;
}

// Step k, part 1: one CTA per window turns the 64x64 diagonal block into inv(L_kk) and publishes it
// (the only form of L_kk anything downstream uses: panel blocks, y and the solve).  This is the
// sequential critical path of the factorisation, so it runs exactly once per window and step and
// is organised for latency: two warp-synchronous 32x32 inverse-Cholesky passes joined by three
// small products, with A = [A11 .; A21 A22]:
//   X11 = inv(chol(A11));  L21 = A21 X11^T;  S = A22 - L21 L21^T;  X22 = inv(chol(S));
//   X21 = -X22 L21 X11;    inv(L_kk) = [X11 0; X21 X22]
__global__ void __launch_bounds__(256)
chol_diag_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ tt, double* dinv, int* status,
                 const int* __restrict__ skip, int k) {
  if (skip && skip[blockIdx.x]) return;  // certificate copy whose analytic bound already holds
  const SolveWin w = wins[blockIdx.x];
  const int n = w.n_t;
  if (k >= win_nb(n)) return;
  const double* A = tt + w.off_tt;
  const int ld = w.ld_t;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int H = 32;
  extern __shared__ __align__(16) double sm[];
  double* Ms = sm;                      // [64][MP] A_kk (lower), later S in its (1,1) quadrant
  double* Xs = Ms + NB * MP;            // [64][MP] inv(L_kk)
  double* Ls = Xs + NB * MP;            // [32][34] L21, then T = L21 X11: [j*(H+2) + r]
  double* colbuf = Ls + H * (H + 2);    // [32]
  double* rowbuf = colbuf + H;          // [32]
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
  if (warp == 0 && !invchol32_warp(Ms, Xs, 0, colbuf, rowbuf, lane) && lane == 0) ok_s = 0;
  __syncthreads();
  // L21(r, c) = sum_{j<=c} A21(r, j) X11(c, j): thread -> (r = tid & 31, c = (tid >> 5) + 8 i)
  {
    const int r = tid & 31;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = (tid >> 5) + 8 * i;
      double v = 0.0;
      for (int j = 0; j <= c; j++) v = fma(Ms[j * MP + H + r], Xs[j * MP + c], v);
      Ls[c * (H + 2) + r] = v;
    }
  }
  __syncthreads();
  // S(r, c) = A22(r, c) - sum_j L21(r, j) L21(c, j), lower triangle
  {
    const int r = tid & 31;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = (tid >> 5) + 8 * i;
      if (c <= r) {
        double v = Ms[(H + c) * MP + H + r];
#pragma unroll 8
        for (int j = 0; j < H; j++) v = fma(-Ls[j * (H + 2) + r], Ls[j * (H + 2) + c], v);
        Ms[(H + c) * MP + H + r] = v;
      }
    }
  }
  __syncthreads();
  if (warp == 0 && !invchol32_warp(Ms, Xs, H, colbuf, rowbuf, lane) && lane == 0) ok_s = 0;
  __syncthreads();
  // T(r, c) = sum_{j>=c} L21(r, j) X11(j, c)   (kept in registers across the barrier, then stored over L21)
  double tv[4];
  {
    const int r = tid & 31;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = (tid >> 5) + 8 * i;
      double v = 0.0;
      for (int j = c; j < H; j++) v = fma(Ls[j * (H + 2) + r], Xs[c * MP + j], v);
      tv[i] = v;
    }
  }
  __syncthreads();
  {
    const int r = tid & 31;
#pragma unroll
    for (int i = 0; i < 4; i++) Ls[((tid >> 5) + 8 * i) * (H + 2) + r] = tv[i];
  }
  __syncthreads();
  // X21(r, c) = -sum_{kk<=r} X22(r, kk) T(kk, c)
  {
    const int r = tid & 31;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = (tid >> 5) + 8 * i;
      double v = 0.0;
      for (int kk = 0; kk <= r; kk++) v = fma(Xs[(H + kk) * MP + H + r], Ls[c * (H + 2) + kk], v);
      Xs[c * MP + H + r] = -v;
    }
  }
  __syncthreads();
  if (!ok_s && tid == 0) atomicOr(&status[blockIdx.x], STATUS_BREAKDOWN);
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int c = idx >> 6, r = idx & 63;
    dinv[w.off_dinv + (long long)k * NB * NB + c * NB + r] = Xs[c * MP + r];  // zero above the diagonal
  }
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
  double* X = sm;                 // [64][MP] inv(L_kk), X[j*MP + c] = inv(L_kk)(c, j)
  double* T = X + NB * MP;        // [64][MP] A_ik tile, T[j*MP + r] = A(ib*64+r, k*64+j)

  const int k0 = k * NB;
  const int i0 = ib * NB;
  const double* D = dinv + w.off_dinv + (long long)k * NB * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int c = idx >> 6, r = idx & 63;
    X[c * MP + r] = D[c * NB + r];
    T[c * MP + r] = (i0 + r < n && k0 + c < n) ? A[(long long)(k0 + c) * ld + i0 + r] : 0.0;
  }
  __syncthreads();

  // L(r, c) = sum_{j<=c} A(r, j) X(c, j); 4x4 register tile per thread,
  // rows {2a,2a+1,32+2a,32+2a+1}, columns 4b..4b+3
  {
    const int a2 = (tid & 15) * 2, cb = (tid >> 4) * 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
    const int jmax = cb + 3;  // X(c, j) = 0 for j > c
    for (int j = 0; j <= jmax; j++) {
      const double2 t01 = *reinterpret_cast<const double2*>(&T[j * MP + a2]);
      const double2 t23 = *reinterpret_cast<const double2*>(&T[j * MP + 32 + a2]);
      const double2 x01 = *reinterpret_cast<const double2*>(&X[j * MP + cb]);
      const double2 x23 = *reinterpret_cast<const double2*>(&X[j * MP + cb + 2]);
      const double tv[4] = {t01.x, t01.y, t23.x, t23.y};
      const double xv[4] = {x01.x, x01.y, x23.x, x23.y};
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = fma(tv[a], xv[b], acc[a][b]);
    }
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int c = k0 + cb + b;
      if (c >= n) continue;
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int r = i0 + a2 + (a & 1) + (a >> 1) * 32;
        if (r < n) A[(long long)c * ld + r] = acc[a][b];
      }
    }
  }
}

// trailing update of step k: A_ij -= L_ik L_jk^T for k < j <= i
__global__ void __launch_bounds__(256)
chol_update_kernel(const SolveWin* __restrict__ wins, double* tt, const int* __restrict__ skip, int k) {
  if (skip && skip[blockIdx.y]) return;
  const SolveWin w = wins[blockIdx.y];
  const int n = w.n_t;
  const int nb = win_nb(n);
  const int tb = nb - k - 1;
  if (tb <= 0) return;
  const int t = blockIdx.x;
  if (t >= tb * (tb + 1) / 2) return;
  int ii = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ii + 1) * (ii + 2) / 2 <= t) ii++;
  while (ii * (ii + 1) / 2 > t) ii--;
  const int jj = t - ii * (ii + 1) / 2;
  const int i0 = (k + 1 + ii) * NB, j0 = (k + 1 + jj) * NB, k0 = k * NB;
  double* A = tt + w.off_tt;
  const int ld = w.ld_t;
  const int tid = threadIdx.x;

  extern __shared__ double sm[];
  double (*Ls)[NB] = reinterpret_cast<double (*)[NB]>(sm);            // Ls[kk][r] = L(i0+r, k0+kk)
  double (*Rs)[NB] = reinterpret_cast<double (*)[NB]>(sm + NB * NB);  // Rs[kk][c] = L(j0+c, k0+kk)
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int kk = idx >> 6, r = idx & 63;
    Ls[kk][r] = (i0 + r < n) ? A[(long long)(k0 + kk) * ld + i0 + r] : 0.0;
    Rs[kk][r] = (j0 + r < n) ? A[(long long)(k0 + kk) * ld + j0 + r] : 0.0;
  }
  __syncthreads();
  const int a2 = (tid & 15) * 2;   // rows {a2, a2+1, 32+a2, 32+a2+1}: conflict-free 16-byte smem reads
  const int tc = (tid >> 4) * 4;   // cols tc..tc+3
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 8
  for (int kk = 0; kk < NB; kk++) {
    const double2 l01 = *reinterpret_cast<const double2*>(&Ls[kk][a2]);
    const double2 l23 = *reinterpret_cast<const double2*>(&Ls[kk][32 + a2]);
    const double2 r01 = *reinterpret_cast<const double2*>(&Rs[kk][tc]);
    const double2 r23 = *reinterpret_cast<const double2*>(&Rs[kk][tc + 2]);
    const double lv[4] = {l01.x, l01.y, l23.x, l23.y};
    const double rv[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = fma(lv[a], rv[b], acc[a][b]);
  }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int c = j0 + tc + b;
    if (c >= n) continue;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int r = i0 + a2 + (a & 1) + (a >> 1) * 32;
      if (r < n && r >= c) A[(long long)c * ld + r] -= acc[a][b];
    }
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
// Thread tile 8 rows x 4 columns: warp g owns rows 8g..8g+7 of the 64-row block, lane l owns
// columns {2l, 2l+1, 64+2l, 64+2l+1}.  Operand chunks (32 k-steps of L: 16 KiB, of W: 32 KiB) are
// double-buffered with cp.async.cg (L2 path: W rows were written by this CTA earlier).  The
// accumulator tile and inv(L_ii) alias the chunk buffers once the k-loop of a row block is done,
// so a CTA needs 96 KiB and two CTAs share an SM.
constexpr int UB = 128;   // unmeasured SNPs (columns of W) per CTA
constexpr int KC = 32;    // k-steps per staged chunk
constexpr int TR_SMEM_BYTES = 2 * (KC * NB + KC * UB) * 8;  // 98,304

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256, 2)
trsm_finalize_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ tt,
                     const double* __restrict__ dinv, double* ut, const double* __restrict__ zt,
                     double* zu, double* info) {
  const SolveWin w = wins[blockIdx.y];
  const int n = w.n_t, nu = w.n_u;
  const int u0 = blockIdx.x * UB;
  if (u0 >= nu) return;
  const int nb = win_nb(n);
  const double* L = tt + w.off_tt;
  const int ld = w.ld_t;
  double* W = ut + w.off_ut;
  const int ldu = w.ld_u;
  const int tid = threadIdx.x;
  const int lane = tid & 31, g = tid >> 5;
  const int tr = g * 8;

  extern __shared__ __align__(16) double sm[];
  double* LsBuf = sm;                       // 2 x [KC][64]
  double* WsBuf = sm + 2 * KC * NB;         // 2 x [KC][128]
  double* Ds = LsBuf;                       // [64][64] inv(L_ii), aliases both L chunks
  double* Ts = WsBuf;                       // [64][128] accumulator tile, aliases both W chunks
  double* red = WsBuf;                      // [8][128] final reductions
  double* ys = sm + 2 * (KC * NB + KC * UB); // [nb*64] y = L^-1 Z1 solved so far (every CTA carries this extra
  double* rys = ys + nb * NB;                // [64]     right-hand side column itself: ~1/128 more work, no
                                             //          separate latency-bound kernel and no y round trip)

  double p_info[4] = {0.0, 0.0, 0.0, 0.0}, p_z[4] = {0.0, 0.0, 0.0, 0.0};
  const int ccol[2] = {2 * lane, 64 + 2 * lane};   // first column of each 2-wide strip of this thread
  const int cvalid = ldu - u0;                     // columns that exist in the row (ldu is a multiple of 8)

  for (int ib = 0; ib < nb; ib++) {
    const int i0 = ib * NB;
    double acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int r = i0 + tr + a;
#pragma unroll
      for (int s = 0; s < 2; s++) {
        double2 v = make_double2(0.0, 0.0);
        if (r < n && ccol[s] < cvalid) v = *reinterpret_cast<const double2*>(&W[(long long)r * ldu + u0 + ccol[s]]);
        acc[a][2 * s] = v.x;
        acc[a][2 * s + 1] = v.y;
      }
    }
    // acc -= L(ib, 0:i0) * W(0:i0) in chunks of KC k-steps
    const int nchunk = ib * (NB / KC);
    auto issue = [&](int ch, int buf) {
      const int kbase = ch * KC;
      double* ls = LsBuf + buf * KC * NB;
      double* ws = WsBuf + buf * KC * UB;
      // L chunk: KC columns x 64 rows in 16-byte pieces, 32 per column
#pragma unroll
      for (int it = 0; it < (KC * NB / 2) / 256; it++) {
        const int idx = tid + it * 256;
        const int kk = idx >> 5, r2 = (idx & 31) * 2;
        cp_async16(ls + kk * NB + r2, L + (long long)(kbase + kk) * ld + i0 + r2, i0 + r2 < n);
      }
      // W chunk: KC rows x 128 columns, 64 pieces per row
#pragma unroll
      for (int it = 0; it < (KC * UB / 2) / 256; it++) {
        const int idx = tid + it * 256;
        const int kk = idx >> 6, c2 = (idx & 63) * 2;
        cp_async16(ws + kk * UB + c2, W + (long long)(kbase + kk) * ldu + u0 + c2, c2 < cvalid);
      }
      cp_async_commit();
    };
    double acc_y = 0.0;  // threads 0..63: sum_k L(i0 + tid, k) y_k
    __syncthreads();  // previous row block finished with the aliased buffers (Ts / Ds)
    if (nchunk > 0) issue(0, 0);
    for (int ch = 0; ch < nchunk; ch++) {
      const int buf = ch & 1;
      if (ch + 1 < nchunk) {
        issue(ch + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const double* ls = LsBuf + buf * KC * NB;
      const double* ws = WsBuf + buf * KC * UB;
#pragma unroll 2
      for (int kk = 0; kk < KC; kk++) {
        double lv[8], wv[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double2 t2 = *reinterpret_cast<const double2*>(&ls[kk * NB + tr + 2 * q]);
          lv[2 * q] = t2.x;
          lv[2 * q + 1] = t2.y;
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const double2 t2 = *reinterpret_cast<const double2*>(&ws[kk * UB + ccol[s]]);
          wv[2 * s] = t2.x;
          wv[2 * s + 1] = t2.y;
        }
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) acc[a][b] = fma(-lv[a], wv[b], acc[a][b]);
      }
      if (tid < NB) {
        const double* yk = ys + ch * KC;
#pragma unroll 8
        for (int kk = 0; kk < KC; kk++) acc_y = fma(ls[kk * NB + tid], yk[kk], acc_y);
      }
      __syncthreads();  // buffer `buf` may be refilled by the next issue
    }
    // W_i = inv(L_ii) * acc
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
      for (int s = 0; s < 2; s++)
        *reinterpret_cast<double2*>(&Ts[(tr + a) * UB + ccol[s]]) = make_double2(acc[a][2 * s], acc[a][2 * s + 1]);
    for (int idx = tid; idx < NB * NB / 2; idx += 256)  // Ds[kk*64 + r] = inv(L_ii)(r, kk), zero above diag
      reinterpret_cast<double2*>(Ds)[idx] =
          reinterpret_cast<const double2*>(dinv + w.off_dinv + (long long)ib * NB * NB)[idx];
    if (tid < NB) rys[tid] = (i0 + tid < n) ? zt[w.off_t + i0 + tid] - acc_y : 0.0;
    __syncthreads();
    if (tid < NB) {  // y_i = inv(L_ii) (z_i - sum_k L_ik y_k)
      double v = 0.0;
      for (int j = 0; j <= tid; j++) v = fma(Ds[j * NB + tid], rys[j], v);
      ys[i0 + tid] = v;
    }
    double out[8][4];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) out[a][b] = 0.0;
    const int kmax = tr + 8;  // inv(L_ii)(r, kk) == 0 for kk > r
#pragma unroll 1
    for (int kk = 0; kk < kmax; kk++) {
      double lv[8], tv[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double2 t2 = *reinterpret_cast<const double2*>(&Ds[kk * NB + tr + 2 * q]);
        lv[2 * q] = t2.x;
        lv[2 * q + 1] = t2.y;
      }
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const double2 t2 = *reinterpret_cast<const double2*>(&Ts[kk * UB + ccol[s]]);
        tv[2 * s] = t2.x;
        tv[2 * s + 1] = t2.y;
      }
#pragma unroll
      for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) out[a][b] = fma(lv[a], tv[b], out[a][b]);
    }
    __syncthreads();  // ys[i0 .. i0+63] is complete
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int r = i0 + tr + a;
      if (r >= n) continue;
      const double yr = ys[r];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (ccol[s] < cvalid)
          *reinterpret_cast<double2*>(&W[(long long)r * ldu + u0 + ccol[s]]) =
              make_double2(out[a][2 * s], out[a][2 * s + 1]);
      }
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const double v = out[a][b];
        p_info[b] = fma(v, v, p_info[b]);
        p_z[b] = fma(yr, v, p_z[b]);
      }
    }
  }
  // column reductions over the 8 row groups, fixed order
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 2; s++) {
    red[g * UB + ccol[s]] = p_info[2 * s];
    red[g * UB + ccol[s] + 1] = p_info[2 * s + 1];
  }
  __syncthreads();
  double s_info = 0.0;
  if (tid < UB)
    for (int gg = 0; gg < 8; gg++) s_info += red[gg * UB + tid];
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 2; s++) {
    red[g * UB + ccol[s]] = p_z[2 * s];
    red[g * UB + ccol[s] + 1] = p_z[2 * s + 1];
  }
  __syncthreads();
  if (tid < UB && u0 + tid < nu) {
    double s_z = 0.0;
    for (int gg = 0; gg < 8; gg++) s_z += red[gg * UB + tid];
    const double inf = fabs(s_info);                 // info = |b21 B11^-1 b12|      (dist.cpp:198)
    zu[w.off_u + u0 + tid] = s_z / sqrt(inf);        // z / sqrt(info)               (dist.cpp:200)
    info[w.off_u + u0 + tid] = inf;
  }
}

}  // namespace

int launch_cholesky(Ctx* ctx, const SolveWin* d_wins, int n_wins, int max_nt, double* d_tt, double* d_dinv,
                    int* d_status, const int* d_skip) {
  if (n_wins == 0) return GB_OK;
  const int nb_max = (max_nt + NB - 1) / NB;
  const size_t smem_panel = sizeof(double) * 2 * NB * MP;
  const size_t smem_update = sizeof(double) * 2 * NB * NB;
  const size_t smem_diag = sizeof(double) * (2 * NB * MP + 32 * 34 + 64);
  static bool attr_set_dev[64] = {};
  bool& attr_set = attr_set_dev[ctx->device & 63];
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_diag));
    GB_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel));
    GB_CUDA(cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_update));
    attr_set = true;
  }
  for (int k = 0; k < nb_max; k++) {
    chol_diag_kernel<<<n_wins, 256, smem_diag, ctx->stream>>>(d_wins, d_tt, d_dinv, d_status, d_skip, k);
    ctx->launches++;
    const int tb = nb_max - k - 1;
    if (tb > 0) {
      chol_panel_kernel<<<dim3(tb, n_wins), 256, smem_panel, ctx->stream>>>(d_wins, d_tt, d_dinv, d_skip, k);
      chol_update_kernel<<<dim3(tb * (tb + 1) / 2, n_wins), 256, smem_update, ctx->stream>>>(d_wins, d_tt, d_skip, k);
      ctx->launches += 2;
    }
  }
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

int launch_trsm_finalize(Ctx* ctx, const SolveWin* d_wins, int n_wins, int max_nt, int max_nu, const double* d_tt,
                         const double* d_dinv, double* d_ut, const double* d_zt, double* d_zu, double* d_info) {
  if (n_wins == 0 || max_nu == 0) return GB_OK;
  const int nb_max = (max_nt + NB - 1) / NB;
  const size_t smem = TR_SMEM_BYTES + sizeof(double) * (size_t)(nb_max * NB + NB);
  if (smem > 227 * 1024) {
    ctx->err = "window has too many measured SNPs for trsm_finalize_kernel";
    return GB_ERR_UNSUPPORTED;
  }
  GB_CUDA(cudaFuncSetAttribute(trsm_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  trsm_finalize_kernel<<<dim3((max_nu + UB - 1) / UB, n_wins), 256, smem, ctx->stream>>>(
      d_wins, d_tt, d_dinv, d_ut, d_zt, d_zu, d_info);
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
