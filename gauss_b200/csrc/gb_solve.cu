// gb_solve.cu -- K2: batched fp64 Cholesky of B11 and blocked triangular solves giving z_u, info_u.
//
// Replaces MakePosDef + InvMat + the three MpMatMat per unmeasured SNP of the reference
// (dist.cpp:181-202, distmix.cpp:202-228; util.cpp:262-264,298-318).  With B11 = L L^T:
//     W = L^-1 B21^T,  y = L^-1 Z1   ->   z = W^T y,  info = |colsumsq(W)|,  z_out = z / sqrt(info)
// which equals b21 B11^-1 Z1 and |b21 B11^-1 b12| of the reference.
// MakePosDef is a no-op whenever lambda_min(B11) >= min_abs_eig; that is certified by a second
// Cholesky of B11 - min_abs_eig*I (succeeds  <=>  lambda_min > min_abs_eig); otherwise the window
// is flagged GB_ERR_NOT_PD -- there is no eigen-clip path and no CPU fallback.
//
// Storage: B11 / L column-major n_t x ld_t (lower triangle significant); B21^T / W row-major
// n_t x ld_u (unmeasured SNPs contiguous).  Everything is blocked by NB = 64.
//
// Cholesky = right-looking, two kernels per block column k, batched over all windows:
//   chol_panel_kernel   every CTA re-factors the 64x64 diagonal block in shared memory and inverts
//                       it; CTA ib==k publishes inv(L_kk) and y_k; CTAs ib>k form
//                       L_ik = A_ik inv(L_kk)^T
//   chol_update_kernel  trailing tiles A_ij -= L_ik L_jk^T
// Triangular solve = one CTA per (window, 128 unmeasured SNPs): forward substitution by row
// blocks, W written in place, column sums of squares and W^T y reduced in a fixed order.
#include "gb_common.cuh"

namespace gb {

namespace {

constexpr int NB = 64;
constexpr int STATUS_BREAKDOWN = 1;

__device__ __forceinline__ int win_nb(int n) { return (n + NB - 1) / NB; }

// ---------------------------------------------------------------------------------------------
// Factor the 64x64 SPD block D (lower triangle, D[c][r] = element (r, c), padded stride) in place
// and put inv(L) into X (same layout, strictly-upper part zero).  256 threads.  Returns false on
// a non-positive / NaN pivot (flag only; execution continues with a substituted pivot).
constexpr int LDS_PAD = NB + 1;

__device__ bool factor_and_invert_block(double* D, double* X) {
  const int tid = threadIdx.x;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (int j = 0; j < NB; j++) {
    __syncthreads();
    const double piv = D[j * LDS_PAD + j];
    const bool bad = !(piv > 0.0);
    const double dj = sqrt(bad ? 1.0 : piv);
    __syncthreads();
    if (tid == 0) {
      D[j * LDS_PAD + j] = dj;
      if (bad) s_bad = 1;
    }
    if (tid > j && tid < NB) D[j * LDS_PAD + tid] = D[j * LDS_PAD + tid] / dj;  // column j below diag
    __syncthreads();
    // trailing update: D(r, c) -= L(r, j) * L(c, j) for j < c <= r
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int c = idx >> 6, r = idx & 63;
      if (c > j && r >= c) D[c * LDS_PAD + r] = fma(-D[j * LDS_PAD + r], D[j * LDS_PAD + c], D[c * LDS_PAD + r]);
    }
  }
  __syncthreads();
  // inverse of the lower-triangular factor, one column per thread (forward substitution on e_c)
  if (tid < NB) {
    const int c = tid;
    for (int r = 0; r < NB; r++) {
      double v = 0.0;
      if (r >= c) {
        double acc = (r == c) ? 1.0 : 0.0;
        for (int j = c; j < r; j++) acc = fma(-D[j * LDS_PAD + r], X[c * LDS_PAD + j], acc);
        v = acc / D[r * LDS_PAD + r];
      }
      X[c * LDS_PAD + r] = v;
    }
  }
  __syncthreads();
  return s_bad == 0;
}

__global__ void __launch_bounds__(256)
chol_panel_kernel(const SolveWin* __restrict__ wins, double* tt, double* dinv, const double* __restrict__ zt,
                  double* y, int* status, int k, int want_y) {
  const SolveWin w = wins[blockIdx.y];
  const int n = w.n_t;
  const int nb = win_nb(n);
  const int ib = k + blockIdx.x;
  if (k >= nb || ib >= nb) return;
  double* A = tt + w.off_tt;
  const int ld = w.ld_t;
  const int tid = threadIdx.x;

  extern __shared__ double sm[];
  double* D = sm;                      // [64][65]
  double* X = D + NB * LDS_PAD;        // [64][65]
  double* T = X + NB * LDS_PAD;        // [64][65]  A_ik tile, T[j][r] = A(ib*64+r, k*64+j)

  const int k0 = k * NB;
  // diagonal block; rows/cols past n are padded with the identity
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int c = idx >> 6, r = idx & 63;
    double v = (r == c) ? 1.0 : 0.0;
    if (k0 + r < n && k0 + c < n && r >= c) v = A[(long long)(k0 + c) * ld + k0 + r];
    D[c * LDS_PAD + r] = v;
  }
  const bool ok = factor_and_invert_block(D, X);

  if (ib == k) {
    if (!ok && tid == 0) atomicOr(&status[blockIdx.y], STATUS_BREAKDOWN);
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int c = idx >> 6, r = idx & 63;
      // L_kk itself is never needed again (panel blocks, y and the solve all use inv(L_kk)), and
      // writing it here would race with the sibling CTAs still loading A_kk.
      dinv[w.off_dinv + (long long)k * NB * NB + c * NB + r] = X[c * LDS_PAD + r];
    }
    if (want_y) {
      // y_k = inv(L_kk) (z_k - sum_{j<k} L_kj y_j)
      double* rhs = T;  // reuse
      if (tid < NB) {
        const int r = k0 + tid;
        double acc = 0.0;
        if (r < n) {
          acc = zt[w.off_t + r];
          for (int c = 0; c < k0; c++) acc = fma(-A[(long long)c * ld + r], y[w.off_t + c], acc);
        }
        rhs[tid] = acc;
      }
      __syncthreads();
      if (tid < NB) {
        double acc = 0.0;
        for (int j = 0; j <= tid; j++) acc = fma(X[j * LDS_PAD + tid], rhs[j], acc);
        if (k0 + tid < n) y[w.off_t + k0 + tid] = acc;
      }
    }
    return;
  }

  // off-diagonal panel block: L_ik = A_ik inv(L_kk)^T, i.e. L(r, c) = sum_{j<=c} A(r, j) X(c, j)
  const int i0 = ib * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int j = idx >> 6, r = idx & 63;
    T[j * LDS_PAD + r] = (i0 + r < n && k0 + j < n) ? A[(long long)(k0 + j) * ld + i0 + r] : 0.0;
  }
  __syncthreads();
  {
    const int r = tid & 63;
    const int cq = tid >> 6;  // 0..3 -> columns cq, cq+4, ...
    for (int c = cq; c < NB; c += 4) {
      double acc = 0.0;
      for (int j = 0; j <= c; j++) acc = fma(T[j * LDS_PAD + r], X[j * LDS_PAD + c], acc);
      if (i0 + r < n && k0 + c < n) A[(long long)(k0 + c) * ld + i0 + r] = acc;
    }
  }
}

// trailing update of step k: A_ij -= L_ik L_jk^T for k < j <= i
__global__ void __launch_bounds__(256)
chol_update_kernel(const SolveWin* __restrict__ wins, double* tt, int k) {
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
  const int tr = (tid & 15) * 4;   // rows tr..tr+3   (fastest across threads -> coalesced stores)
  const int tc = (tid >> 4) * 4;   // cols tc..tc+3
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 8
  for (int kk = 0; kk < NB; kk++) {
    const double2 l01 = *reinterpret_cast<const double2*>(&Ls[kk][tr]);
    const double2 l23 = *reinterpret_cast<const double2*>(&Ls[kk][tr + 2]);
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
      const int r = i0 + tr + a;
      if (r < n && r >= c) A[(long long)c * ld + r] -= acc[a][b];
    }
  }
}

// dst = src with the diagonal lowered by `shift` (lower triangle only), per window
__global__ void copy_shift_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ src, double* dst,
                                  double shift) {
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

// ---------------------------------------------------------------------------------------------
// Blocked forward substitution W = L^-1 B21^T for 128 unmeasured SNPs, fused with the reductions.
constexpr int UB = 128;            // unmeasured SNPs (columns of W) per CTA
constexpr int TR_SMEM_DOUBLES = NB * NB /*Ls*/ + NB * UB /*Ws*/ + NB * UB /*Ts*/ + 16 * UB /*red*/;

__global__ void __launch_bounds__(256)
trsm_finalize_kernel(const SolveWin* __restrict__ wins, const double* __restrict__ tt,
                     const double* __restrict__ dinv, double* ut, const double* __restrict__ y,
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

  extern __shared__ double sm[];
  double* Ls = sm;                 // [64 kk][64 r]
  double* Ws = Ls + NB * NB;       // [64 kk][128 c]
  double* Ts = Ws + NB * UB;       // [64 r ][128 c]
  double* red = Ts + NB * UB;      // [16][128]

  const int tr = (tid >> 4) * 4;   // rows tr..tr+3 of the 64-row block
  const int tc = (tid & 15) * 8;   // cols tc..tc+7 of the 128-column block
  double p_info[8], p_z[8];
#pragma unroll
  for (int b = 0; b < 8; b++) p_info[b] = 0.0, p_z[b] = 0.0;

  for (int ib = 0; ib < nb; ib++) {
    const int i0 = ib * NB;
    double acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int r = i0 + tr + a;
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const int c = u0 + tc + b;
        acc[a][b] = (r < n && c < nu) ? W[(long long)r * ldu + c] : 0.0;
      }
    }
    // acc -= L(ib, jb) * W(jb) for jb < ib
    for (int jb = 0; jb < ib; jb++) {
      const int j0 = jb * NB;
      __syncthreads();
      for (int idx = tid; idx < NB * NB; idx += 256) {
        const int kk = idx >> 6, r = idx & 63;
        Ls[kk * NB + r] = (i0 + r < n) ? L[(long long)(j0 + kk) * ld + i0 + r] : 0.0;
      }
      for (int idx = tid; idx < NB * UB; idx += 256) {
        const int kk = idx >> 7, c = idx & 127;
        Ws[kk * UB + c] = (u0 + c < nu) ? W[(long long)(j0 + kk) * ldu + u0 + c] : 0.0;
      }
      __syncthreads();
#pragma unroll 4
      for (int kk = 0; kk < NB; kk++) {
        const double2 l01 = *reinterpret_cast<const double2*>(&Ls[kk * NB + tr]);
        const double2 l23 = *reinterpret_cast<const double2*>(&Ls[kk * NB + tr + 2]);
        const double lv[4] = {l01.x, l01.y, l23.x, l23.y};
        double wv[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double2 t2 = *reinterpret_cast<const double2*>(&Ws[kk * UB + tc + 2 * q]);
          wv[2 * q] = t2.x;
          wv[2 * q + 1] = t2.y;
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 8; b++) acc[a][b] = fma(-lv[a], wv[b], acc[a][b]);
      }
    }
    // W_i = inv(L_ii) * acc
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 8; b++) Ts[(tr + a) * UB + tc + b] = acc[a][b];
    for (int idx = tid; idx < NB * NB; idx += 256)  // Ls[kk][r] = inv(L_ii)(r, kk) (zero above diag)
      Ls[idx] = dinv[w.off_dinv + (long long)ib * NB * NB + idx];
    __syncthreads();
    double out[4][8];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 8; b++) out[a][b] = 0.0;
    const int kmax = tr + 4;  // inv(L_ii)(r, kk) == 0 for kk > r
    for (int kk = 0; kk < kmax; kk++) {
      const double2 l01 = *reinterpret_cast<const double2*>(&Ls[kk * NB + tr]);
      const double2 l23 = *reinterpret_cast<const double2*>(&Ls[kk * NB + tr + 2]);
      const double lv[4] = {l01.x, l01.y, l23.x, l23.y};
      double tv[8];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double2 t2 = *reinterpret_cast<const double2*>(&Ts[kk * UB + tc + 2 * q]);
        tv[2 * q] = t2.x;
        tv[2 * q + 1] = t2.y;
      }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) out[a][b] = fma(lv[a], tv[b], out[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int r = i0 + tr + a;
      if (r >= n) continue;
      const double yr = y[w.off_t + r];
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const int c = u0 + tc + b;
        const double v = out[a][b];
        if (c < nu) W[(long long)r * ldu + c] = v;
        p_info[b] = fma(v, v, p_info[b]);
        p_z[b] = fma(yr, v, p_z[b]);
      }
    }
  }
  // column reductions over the 16 row groups, fixed order
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 8; b++) red[(tid >> 4) * UB + tc + b] = p_info[b];
  __syncthreads();
  double s_info = 0.0;
  if (tid < UB)
    for (int g = 0; g < 16; g++) s_info += red[g * UB + tid];
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 8; b++) red[(tid >> 4) * UB + tc + b] = p_z[b];
  __syncthreads();
  if (tid < UB && u0 + tid < nu) {
    double s_z = 0.0;
    for (int g = 0; g < 16; g++) s_z += red[g * UB + tid];
    const double inf = fabs(s_info);                 // info = |b21 B11^-1 b12|      (dist.cpp:198)
    zu[w.off_u + u0 + tid] = s_z / sqrt(inf);        // z / sqrt(info)               (dist.cpp:200)
    info[w.off_u + u0 + tid] = inf;
  }
}

}  // namespace

static int max_nt(const std::vector<SolveWin>& h) {
  int m = 0;
  for (const auto& w : h) m = w.n_t > m ? w.n_t : m;
  return m;
}

int launch_cholesky(Ctx* ctx, const SolveWin* d_wins, const std::vector<SolveWin>& h_wins, double* d_tt,
                    double* d_dinv, const double* d_zt, double* d_y, int* d_status, double /*shift*/,
                    int want_y) {
  const int nw = (int)h_wins.size();
  if (nw == 0) return GB_OK;
  const int nb_max = (max_nt(h_wins) + NB - 1) / NB;
  const size_t smem_panel = sizeof(double) * 3 * NB * LDS_PAD;
  const size_t smem_update = sizeof(double) * 2 * NB * NB;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel));
    GB_CUDA(cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_update));
    attr_set = true;
  }
  for (int k = 0; k < nb_max; k++) {
    chol_panel_kernel<<<dim3(nb_max - k, nw), 256, smem_panel, ctx->stream>>>(d_wins, d_tt, d_dinv, d_zt, d_y,
                                                                             d_status, k, want_y);
    ctx->launches++;
    const int tb = nb_max - k - 1;
    if (tb > 0) {
      chol_update_kernel<<<dim3(tb * (tb + 1) / 2, nw), 256, smem_update, ctx->stream>>>(d_wins, d_tt, k);
      ctx->launches++;
    }
  }
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

int launch_trsm_finalize(Ctx* ctx, const SolveWin* d_wins, const std::vector<SolveWin>& h_wins,
                         const double* d_tt, const double* d_dinv, double* d_ut, const double* d_y,
                         double* d_zu, double* d_info) {
  const int nw = (int)h_wins.size();
  if (nw == 0) return GB_OK;
  int nu_max = 0;
  for (const auto& w : h_wins) nu_max = w.n_u > nu_max ? w.n_u : nu_max;
  if (nu_max == 0) return GB_OK;
  const size_t smem = sizeof(double) * TR_SMEM_DOUBLES;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(trsm_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  trsm_finalize_kernel<<<dim3((nu_max + UB - 1) / UB, nw), 256, smem, ctx->stream>>>(d_wins, d_tt, d_dinv, d_ut,
                                                                                    d_y, d_zu, d_info);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

int launch_copy_shift(Ctx* ctx, const SolveWin* d_wins, const std::vector<SolveWin>& h_wins,
                      const double* d_src, double* d_dst, double shift) {
  const int nw = (int)h_wins.size();
  if (nw == 0) return GB_OK;
  copy_shift_kernel<<<dim3(64, nw), 256, 0, ctx->stream>>>(d_wins, d_src, d_dst, shift);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

}  // namespace gb
