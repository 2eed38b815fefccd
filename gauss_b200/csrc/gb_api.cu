// gb_api.cu -- C-ABI of gauss_b200 (include/gauss_b200.h): contexts, packed panels, the window
// batch engine and the host-side mirror of run_dist / run_distmix (dist.cpp:129-227,
// distmix.cpp:138-253).  All compute is CUDA; there is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <type_traits>

#include "gb_common.cuh"
#include <atomic>
#include <chrono>
#include <thread>

using namespace gb;

#include <nvtx3/nvToolsExt.h>

#include "gb_batch.cuh"

extern "C" void gb_pipe_destroy(gb_pipe* pp);

static thread_local std::string g_create_err;

namespace gb {

namespace {

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Persistent buffers (descriptors, row lists, z_t): stream-ordered pool memory owned by the batch -- the per-window
// entry points create and destroy a batch per call, and cudaMalloc/cudaFree would serialise the device every time.
template <class T>
int dev_alloc(gb_batch* b, T** p, size_t n) {
  Ctx* ctx = b->ctx;
  *p = nullptr;
  if (n == 0) n = 1;
  GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(p), n * sizeof(T), ctx->stream));
  b->owned.push_back(*p);
  return GB_OK;
}
// Transient buffers (rewritten by every run): from the shared arena when the batch has one.
template <class T>
int dev_alloc_transient(gb_batch* b, T** p, size_t n) {
  if (!b->arena.base) return dev_alloc(b, p, n);
  if (n == 0) n = 1;
  const size_t bytes = align256(n * sizeof(T));
  if (b->arena_used + bytes > b->arena.cap) {
    b->ctx->err = "batch arena too small";
    return GB_ERR_OOM;
  }
  *p = reinterpret_cast<T*>(b->arena.base + b->arena_used);
  b->arena_used += bytes;
  return GB_OK;
}
template <class T>
int dev_upload(gb_batch* b, T** p, const std::vector<T>& h) {
  Ctx* ctx = b->ctx;
  int rc = dev_alloc(b, p, h.size());
  if (rc) return rc;
  if (!h.empty()) GB_CUDA(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return GB_OK;
}

int unrepresentable(Ctx* ctx) {
  ctx->err = "the panel holds dosages that the E2M1 operand format cannot represent exactly; "
             "repack it with gb_panel_create_fmt(..., GB_PANEL_INT8, ...)";
  return GB_ERR_UNSUPPORTED;
}

}  // namespace

void batch_free_device(gb_batch* b) {
  for (void* p : b->owned)
    if (p) cudaFreeAsync(p, b->ctx->stream);
  b->owned.clear();
  if (b->ev_fork) cudaEventDestroy(b->ev_fork);
  if (b->ev_join) cudaEventDestroy(b->ev_join);
  b->ev_fork = b->ev_join = nullptr;
}

// Shared planner, phase 1 (host only).  rows_u may be empty (ld_mode).  In counts_mode rows_u plays the A side and
// rows_t the B side of one full rectangle.
int batch_plan_host(gb_batch* b, const int64_t* t_off, const int64_t* rows_t, const int64_t* u_off,
                    const int64_t* rows_u, const double* z_t, const double* pop_wgt) {
  Ctx* ctx = b->ctx;
  Panel* pn = b->panel;
  const int64_t nw = b->n_windows;
  b->t_off.assign(t_off, t_off + nw + 1);
  if (u_off) b->u_off.assign(u_off, u_off + nw + 1);
  else b->u_off.assign((size_t)nw + 1, 0);
  b->n_t_total = b->t_off[nw];
  b->n_u_total = b->u_off[nw];
  if (b->t_off[0] != 0 || b->u_off[0] != 0) {
    ctx->err = "window offsets must start at 0";
    return GB_ERR_BAD_ARG;
  }
  std::vector<int32_t>& h_rows_t = b->h_rows_t;
  std::vector<int32_t>& h_rows_u = b->h_rows_u;
  h_rows_t.resize((size_t)b->n_t_total);
  h_rows_u.resize((size_t)b->n_u_total);
  for (int64_t i = 0; i < b->n_t_total; i++) {
    if (rows_t[i] < 0 || rows_t[i] >= pn->n_rows) {
      ctx->err = "measured row index out of range";
      return GB_ERR_BAD_ARG;
    }
    h_rows_t[(size_t)i] = (int32_t)rows_t[i];
  }
  for (int64_t i = 0; i < b->n_u_total; i++) {
    if (rows_u[i] < 0 || rows_u[i] >= pn->n_rows) {
      ctx->err = "unmeasured row index out of range";
      return GB_ERR_BAD_ARG;
    }
    h_rows_u[(size_t)i] = (int32_t)rows_u[i];
  }
  if (z_t && !b->ld_mode && !b->counts_mode) b->h_zt.assign(z_t, z_t + b->n_t_total);

  // ---- segments / coefficients
  GramParams& gp = b->gp;
  std::memset(&gp, 0, sizeof(gp));
  // int8 rows -> kind::i8; E2M1 rows -> kind::mxf4 (nibbles stay packed, K = 64 per instruction) or kind::f8f6f4
  b->fkind = pn->format == GB_PANEL_E2M1 ? (ctx->e2m1_mxf4 ? 7 : 6) : 0;
  gp.fkind = b->fkind;
  const int atom = b->fkind == 7 ? 2 * K_ATOM : K_ATOM;  // K columns one MMA instruction consumes
  {
    const long long pop_max = (long long)*std::max_element(pn->pop_sizes.begin(), pn->pop_sizes.end());
    const long long seg_max = b->mode == GRAM_POOLED ? (long long)pn->n_samples : pop_max;
    if (b->fkind != 0) {
      // E2M1 panels count in fp32: a segment's largest possible sum (dosage 6 on both sides, 36 per individual) must stay
      // below 2^21 for the one-instruction count -> fp64 re-encoding of the mixture fold and below 2^22 for the magic-add
      // count -> int32 of the pooled / counts paths (both far inside fp32's exact-integer range)
      const long long limit = (b->mode == GRAM_MIX && !b->counts_mode) ? (1ll << 21) : (1ll << 22);
      if (36 * seg_max >= limit) {
        ctx->err = "a population block of " + std::to_string(seg_max) + " individuals is too large for the exact fp32 counts of an "
                   "E2M1 panel (limit " + std::to_string((limit - 1) / 36) + "); an int8 panel (GB_PANEL_INT8) takes blocks of up to 133,143 individuals";
        return GB_ERR_UNSUPPORTED;
      }
    } else if (127ll * 127ll * seg_max > 2147483647ll) {
      // int8 panels count in int32: |sum x_i x_j| <= 127^2 * m must fit (the fold itself runs in 64-bit / fp64)
      ctx->err = "a population block of " + std::to_string(seg_max) + " individuals overflows the int32 counts of an int8 panel "
                 "(limit 133,143)";
      return GB_ERR_UNSUPPORTED;
    }
  }
  b->h_coef.assign((size_t)pn->n_pops, 0.0);
  b->h_wgt.assign((size_t)pn->n_pops, 0.0);
  std::vector<double>&h_coef = b->h_coef, &h_wgt = b->h_wgt;
  if (b->mode == GRAM_POOLED) {
    gp.n_seg = 1;
    const int last = pn->n_pops - 1;
    const int k_used = pn->koff[last] + round_up(pn->pop_sizes[last], K_ATOM);
    gp.seg[0] = Seg{0, (k_used + atom - 1) / atom, pn->n_samples, 0};
    gp.n_pooled = (double)pn->n_samples;
  } else {
    gp.n_seg = pn->n_pops;
    for (int p = 0; p < pn->n_pops; p++) {
      const int m = pn->pop_sizes[p];
      gp.seg[p] = Seg{pn->koff[p], (m + atom - 1) / atom, m, 0};
      if (pop_wgt) {
        const double factor = ((double)m) / (m - 1);   // util.cpp:117
        h_wgt[(size_t)p] = pop_wgt[p];
        h_coef[(size_t)p] = pop_wgt[p] * factor;       // wgt_val*factor, left-assoc in util.cpp:118
        gp.coef[p] = h_coef[(size_t)p];
        gp.wgt[p] = h_wgt[(size_t)p];
        gp.coefm[p] = h_coef[(size_t)p] * m;
        gp.kappa[p] = pop_wgt[p] / ((double)m * m) - h_coef[(size_t)p];
      }
    }
  }
  if (b->mode == GRAM_MIX && pop_wgt) {
    double sw = 0.0, wmax = 0.0;
    bool applies = true;
    for (int p = 0; p < pn->n_pops; p++) {
      sw += pop_wgt[p];
      wmax = std::max(wmax, pop_wgt[p]);
      if (!(pop_wgt[p] >= 0.0) || pn->pop_sizes[p] < 2) applies = false;
    }
    b->gneg = applies ? std::max(0.0, sw - 1.0) * wmax : std::numeric_limits<double>::infinity();
  }
  gp.mode = b->counts_mode ? GRAM_COUNTS : b->mode;
  {
    // int8 mixture fold: d = m*sumxy - sumx*sumy is formed in int32 when it provably fits and in exact fp64 otherwise
    // (any byte the reference's (c - '0') can see: |d| <= 127^2 m^2; refined in phase 2 when the panel holds only 0/1/2)
    const long long m = (long long)*std::max_element(pn->pop_sizes.begin(), pn->pop_sizes.end());
    gp.wide_fold = (b->fkind == 0 && 16129ll * m * m > 2147483647ll) ? 1 : 0;
  }
  gp.mirror = b->ld_mode ? 1 : 0;
  gp.raw_out = (gp.mode == GRAM_MIX && b->fkind != 0) ? 1 : 0;
  gp.diag = b->ld_mode ? b->ld_diag : 1.0 + b->params.lambda;   // computeLD.cpp:107 vs dist.cpp:172

  // ---- windows
  b->plan_status.assign((size_t)nw, GB_OK);
  std::vector<int32_t>& h_gather = b->h_gather;
  std::vector<GramTile> h_tiles_tt, h_tiles_ut;
  const double N = (double)pn->n_samples;
  for (int64_t w = 0; w < nw; w++) {
    const int64_t nt = b->t_off[w + 1] - b->t_off[w];
    const int64_t nu = b->u_off[w + 1] - b->u_off[w];
    if (nt < 0 || nu < 0) {
      ctx->err = "window offsets must be non-decreasing";
      return GB_ERR_BAD_ARG;
    }
    if (!b->counts_mode) {
      if (nt <= b->params.min_num_measured_snp) {            // dist.cpp:146, computeLD.cpp:89
        b->plan_status[(size_t)w] = GB_ERR_TOO_FEW_MEASURED;
        continue;
      }
      if (!b->ld_mode && nu <= b->params.min_num_unmeasured_snp) {
        b->plan_status[(size_t)w] = GB_ERR_TOO_FEW_UNMEASURED;
        continue;
      }
    } else if (nt == 0 || nu == 0) {
      continue;
    }
    SolveWin sw{};
    sw.n_t = (int)nt;
    sw.n_u = (int)nu;
    sw.ld_t = round_up((int)nt, 8);
    sw.ld_u = round_up((int)std::max<int64_t>(nu, 1), 8);
    sw.off_t = b->t_off[w];
    sw.off_u = b->u_off[w];
    sw.off_tt = b->tt_elems;
    sw.off_ut = b->ut_elems;
    sw.off_dinv = b->dinv_elems;
    sw.flags = 1;
    long long counts_off = b->counts_elems;
    if (b->counts_mode) {
      b->counts_elems += (long long)nu * nt;
    } else {
      b->tt_elems += (long long)nt * sw.ld_t;
      if (!b->ld_mode) b->ut_elems += (long long)nt * sw.ld_u;
      b->dinv_elems += (long long)((nt + 63) / 64) * 64 * 64;
    }
    b->active.push_back((int)w);
    b->h_wins.push_back(sw);

    // operand placement: a contiguous ascending run is read straight from the panel by TMA,
    // anything else is gathered into scratch rows first
    auto place = [&](const int32_t* rows, int64_t n, int& src, int& row0) {
      bool contig = true;
      for (int64_t i = 1; i < n && contig; i++) contig = rows[i] == rows[0] + i;
      if (contig) {
        src = 0;
        row0 = rows[0];
      } else {
        src = 1;
        row0 = (int)h_gather.size();
        h_gather.insert(h_gather.end(), rows, rows + n);
      }
    };
    int t_src = 0, t_row0 = 0, u_src = 0, u_row0 = 0;
    place(h_rows_t.data() + b->t_off[w], nt, t_src, t_row0);
    if (nu > 0) place(h_rows_u.data() + b->u_off[w], nu, u_src, u_row0);

    const int nbt = (int)((nt + 127) / 128), nbu = (int)((nu + 127) / 128);
    auto make_tile = [&](bool a_is_u, int bi, int bj) {
      GramTile t{};
      const int64_t na = a_is_u ? nu : nt;
      t.a_src = a_is_u ? u_src : t_src;
      t.a_row0 = (a_is_u ? u_row0 : t_row0) + bi * 128;
      t.a_list0 = (int)((a_is_u ? b->u_off[w] : b->t_off[w]) + bi * 128);
      t.a_valid = (int)std::min<int64_t>(128, na - (int64_t)bi * 128);
      t.b_src = t_src;
      t.b_row0 = t_row0 + bj * 128;
      t.b_list0 = (int)(b->t_off[w] + bj * 128);
      t.b_valid = (int)std::min<int64_t>(128, nt - (int64_t)bj * 128);
      t.a_is_u = a_is_u ? 1 : 0;
      t.i0 = bi * 128;
      t.j0 = bj * 128;
      if (b->counts_mode) {
        t.ld_out = (int)nt;
        t.out_off = counts_off;
      } else {
        t.ld_out = a_is_u ? sw.ld_u : sw.ld_t;
        t.out_off = a_is_u ? sw.off_ut : sw.off_tt;
      }
      return t;
    };
    // Tiles: the B11 lower-triangle tiles of every window go to the FRONT part of the list (h_tiles_tt), the
    // B21 tiles behind them, so the factorisation can start once the front part is done (gb_batch_run).
    // Within a window: A block outer / B block inner keeps concurrently running CTAs on the same A rows.
    if (!b->counts_mode)
      for (int bi = 0; bi < nbt; bi++)
        for (int bj = 0; bj <= bi; bj++) h_tiles_tt.push_back(make_tile(false, bi, bj));
    for (int bi = 0; bi < nbu; bi++)
      for (int bj = 0; bj < nbt; bj++) h_tiles_ut.push_back(make_tile(true, bi, bj));
    if (!b->counts_mode) {
      // algorithmic work, SURVEY.md §8(d)
      const double dnt = (double)nt, dnu = (double)nu;
      b->work_gram_ops += 2.0 * N * (dnu * dnt + dnt * (dnt + 1) / 2);
      b->work_panel_bytes += (dnu + dnt) * N;
      if (!b->ld_mode) b->work_solve_flops += dnt * dnt * dnt / 3 + dnt * dnt * dnu + dnt * dnt + 4 * dnt * dnu;
    }
  }
  b->n_gather = (int64_t)h_gather.size();
  b->n_tiles_tt = (int)h_tiles_tt.size();
  b->h_tiles = std::move(h_tiles_tt);
  b->h_tiles.insert(b->h_tiles.end(), h_tiles_ut.begin(), h_tiles_ut.end());
  {
    // heaviest windows first: the solve kernels map blockIdx.y to this list, and a window's cost
    // grows with n_t^2, so this is longest-processing-time-first scheduling of their CTAs
    std::vector<size_t> perm(b->h_wins.size());
    for (size_t i = 0; i < perm.size(); i++) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(),
                     [&](size_t a, size_t c) { return b->h_wins[a].n_t > b->h_wins[c].n_t; });
    std::vector<SolveWin> hw(perm.size());
    std::vector<int> act(perm.size());
    for (size_t i = 0; i < perm.size(); i++) {
      hw[i] = b->h_wins[perm[i]];
      act[i] = b->active[perm[i]];
    }
    b->h_wins.swap(hw);
    b->active.swap(act);
  }
  for (const SolveWin& sw : b->h_wins) {
    b->max_nt = std::max(b->max_nt, sw.n_t);
    b->max_nu = std::max(b->max_nu, sw.n_u);
  }
  // (mixture weights that are negative or sum well beyond 1 can push |cor| past what the B21 digit planes carry: fp64 then)
  b->ozaki = ctx->solve_ozaki && !b->ld_mode && !b->counts_mode && !b->h_wins.empty() && b->max_nt <= 2048 &&
             (b->mode != GRAM_MIX || b->gneg <= 0.25);
  if (b->ozaki) {
    b->oz_kpad = (b->max_nt + 127) / 128 * 128;
    b->h_oz_wins.resize(b->h_wins.size() * ozaki_win_bytes());
    ozaki_plan(b->h_wins.data(), (int)b->h_wins.size(), b->oz_kpad, ctx->heavy_sms > 0 ? ctx->heavy_sms : ctx->sm_count, b->h_oz_wins.data(), &b->h_oz_tiles, &b->oz_a_rows, &b->oz_b_rows);
    b->oz_n_tiles = (int)(b->h_oz_tiles.size() / ozaki_tile_bytes());
    // B21 tiles learn where their rows sit in the digit planes (a tile belongs to the window whose B21 block it writes)
    for (GramTile& t : b->h_tiles) {
      if (!t.a_is_u) continue;
      for (size_t i = 0; i < b->h_wins.size(); i++)
        if (b->h_wins[i].off_ut == t.out_off && b->h_wins[i].ld_u == t.ld_out) {
          long long a_row0;
          int ra;
          ozaki_tile_rows(b->h_oz_wins.data(), (int)i, &a_row0, &ra);
          t.oz_ra = ra;
          t.oz_row0 = a_row0 + t.i0;
          break;
        }
    }
  }
  return GB_OK;
}

// The transient buffers of phase 2, in allocation order: (bytes, which).  One table serves both the arena sizing and
// the allocation itself so the two cannot drift apart.
namespace {
struct TransientSizes {
  size_t sd_u, pool_u, st_sx_t, st_mean_t, st_sx_u, st_mean_u, counts, tt, ut, dinv, zu, info, scratch, clip, oz_x, oz_pa, oz_pb;
};
TransientSizes transient_sizes(const gb_batch* b) {
  const Panel* pn = b->panel;
  TransientSizes s{};
  const size_t P = (size_t)pn->n_pops;
  s.sd_u = sizeof(double) * (size_t)b->n_u_total;
  s.pool_u = sizeof(int32_t) * (size_t)b->n_u_total;
  if (b->mode == GRAM_MIX && !b->counts_mode) {
    s.st_sx_t = sizeof(int32_t) * P * (size_t)b->n_t_total;
    s.st_mean_t = sizeof(double) * P * (size_t)b->n_t_total;
    s.st_sx_u = sizeof(int32_t) * P * (size_t)b->n_u_total;
    s.st_mean_u = sizeof(double) * P * (size_t)b->n_u_total;
  }
  if (b->counts_mode) {
    s.counts = sizeof(int32_t) * (size_t)b->counts_elems * P;
  } else {
    const bool cert = b->params.check_pd && !b->ld_mode;
    s.tt = sizeof(double) * (size_t)b->tt_elems * (cert ? 2 : 1);
    if (!b->ld_mode) {
      s.ut = sizeof(double) * (size_t)b->ut_elems;
      s.dinv = sizeof(double) * (size_t)b->dinv_elems * (cert ? 2 : 1);
      s.zu = sizeof(double) * (size_t)b->n_u_total;
      s.info = sizeof(double) * (size_t)b->n_u_total;
    }
  }
  s.scratch = (size_t)b->n_gather * (size_t)pn->k_stride;
  if (b->ozaki) {
    s.oz_x = sizeof(double) * (size_t)b->tt_elems;
    s.oz_pa = (size_t)b->oz_a_rows * (size_t)b->oz_kpad;
    s.oz_pb = (size_t)b->oz_b_rows * (size_t)b->oz_kpad;
  }
  return s;
}
}  // namespace

size_t batch_arena_bytes(const gb_batch* b) {
  const TransientSizes s = transient_sizes(b);
  const size_t all[] = {s.sd_u, s.pool_u, s.st_sx_t, s.st_mean_t, s.st_sx_u, s.st_mean_u, s.counts,
                        s.tt, s.ut, s.dinv, s.zu, s.info, s.scratch, s.oz_x, s.oz_pa, s.oz_pb};
  size_t total = 0;
  for (size_t v : all) total += align256(v ? v : 1);
  return total + 4096;
}

int batch_plan_device(gb_batch* b, Arena arena, bool sync) {
  Ctx* ctx = b->ctx;
  Panel* pn = b->panel;
  const int64_t nw = b->n_windows;
  b->arena = arena;
  b->arena_used = 0;
  GramParams& gp = b->gp;
  int rc;
  if ((rc = dev_upload(b, &b->d_rows_t, b->h_rows_t))) return rc;
  if ((rc = dev_upload(b, &b->d_rows_u, b->h_rows_u))) return rc;
  if ((rc = dev_upload(b, &b->d_gather, b->h_gather))) return rc;
  if ((rc = dev_upload(b, &b->d_coef, b->h_coef))) return rc;
  if ((rc = dev_upload(b, &b->d_wgt, b->h_wgt))) return rc;
  {
    // the PD certificate factors B11 - min_abs_eig*I in the same launches: its windows are appended
    // after the real ones and live in the second half of the TT buffer
    std::vector<SolveWin>& all = b->h_wins_all;   // a member: the upload may still be in flight when this returns
    all = b->h_wins;
    if (b->params.check_pd && !b->ld_mode && !b->counts_mode)
      for (SolveWin sw : b->h_wins) {
        sw.off_tt += b->tt_elems;
        sw.off_dinv += b->dinv_elems;
        sw.flags = 0;
        all.push_back(sw);
      }
    b->n_chol_wins = (int)all.size();
    if ((rc = dev_upload(b, &b->d_wins, all))) return rc;
  }
  if ((rc = dev_upload(b, &b->d_tiles, b->h_tiles))) return rc;
  if ((rc = dev_alloc(b, &b->d_sd_t, (size_t)b->n_t_total))) return rc;
  if ((rc = dev_alloc(b, &b->d_rq_t, (size_t)b->n_t_total))) return rc;
  if ((rc = dev_alloc(b, &b->d_skip, 2 * (size_t)nw + 2))) return rc;
  if ((rc = dev_alloc(b, &b->d_pool_t, (size_t)b->n_t_total))) return rc;
  if ((rc = dev_alloc(b, &b->d_status, 2 * (size_t)nw + 2))) return rc;
  if ((rc = dev_alloc(b, &b->d_tile_counter, 1))) return rc;
  const TransientSizes s = transient_sizes(b);
  uint8_t* raw = nullptr;
  auto take = [&](auto** p, size_t bytes) -> int {
    int r = dev_alloc_transient(b, &raw, bytes);
    *p = reinterpret_cast<std::remove_reference_t<decltype(**p)>*>(raw);
    return r;
  };
  if ((rc = take(&b->d_sd_u, s.sd_u))) return rc;
  if ((rc = take(&b->d_pool_u, s.pool_u))) return rc;
  if (s.st_sx_t || s.st_sx_u) {
    if ((rc = take(&b->d_st_sx_t, s.st_sx_t))) return rc;
    if ((rc = take(&b->d_st_mean_t, s.st_mean_t))) return rc;
    if ((rc = take(&b->d_st_sx_u, s.st_sx_u))) return rc;
    if ((rc = take(&b->d_st_mean_u, s.st_mean_u))) return rc;
  }
  if (b->counts_mode) {
    if ((rc = take(&b->d_counts, s.counts))) return rc;
  } else {
    if ((rc = take(&b->d_tt, s.tt))) return rc;
    if (!b->ld_mode) {
      if ((rc = take(&b->d_ut, s.ut))) return rc;
      if ((rc = take(&b->d_dinv, s.dinv))) return rc;
      if ((rc = dev_upload(b, &b->d_zt, b->h_zt))) return rc;
      if ((rc = take(&b->d_zu, s.zu))) return rc;
      if ((rc = take(&b->d_info, s.info))) return rc;
    }
  }
  if (b->ozaki) {
    uint8_t* tmp = nullptr;
    if ((rc = dev_upload(b, &tmp, b->h_oz_wins))) return rc;
    b->d_oz_wins = tmp;
    if ((rc = dev_upload(b, &tmp, b->h_oz_tiles))) return rc;
    b->d_oz_tiles = tmp;
    if ((rc = dev_alloc(b, &b->d_oz_y, (size_t)b->n_t_total))) return rc;
    if ((rc = dev_alloc(b, &b->d_oz_amax, b->h_wins.size()))) return rc;
    if ((rc = dev_alloc(b, &b->d_oz_ex, b->h_wins.size()))) return rc;
    if ((rc = dev_alloc(b, &b->d_oz_nan, (size_t)std::max<int64_t>(b->n_u_total, 1)))) return rc;
    if ((rc = take(&b->d_x, s.oz_x))) return rc;
    if ((rc = take(&b->d_oz_pa, s.oz_pa))) return rc;
    if ((rc = take(&b->d_oz_pb, s.oz_pb))) return rc;
  }
  if (const char* e = getenv("GB_POISON")) {
    // diagnostics: fill chosen transient buffers with a byte pattern before anything writes them ("mask:byte"), to
    // expose reads of memory the batch never initialised
    unsigned mask = 0, byte = 0x55;
    sscanf(e, "%u:%u", &mask, &byte);
    const struct { void* p; size_t n; } bufs[] = {
        {b->d_x, s.oz_x}, {b->d_oz_pa, s.oz_pa}, {b->d_oz_pb, s.oz_pb}, {b->d_oz_y, sizeof(double) * (size_t)b->n_t_total},
        {nullptr, 0}, {b->d_ut, s.ut}, {b->d_tt, s.tt}, {b->d_dinv, s.dinv},
        {b->d_zu, s.zu}, {b->d_info, s.info}, {b->d_scratch, 0}};
    for (int i = 0; i < 10; i++)
      if ((mask >> i & 1) && bufs[i].p && bufs[i].n) cudaMemsetAsync(bufs[i].p, (int)byte, bufs[i].n, ctx->stream);
  }
  if (b->clip_mode) {
    if ((rc = dev_alloc(b, &b->d_eig_G, (size_t)b->tt_elems))) return rc;
    if ((rc = dev_alloc(b, &b->d_eig_V, (size_t)b->tt_elems))) return rc;
    if ((rc = dev_alloc(b, &b->d_evals, (size_t)b->n_t_total))) return rc;
  }
  if (b->n_gather > 0) {
    if ((rc = take(&b->d_scratch, s.scratch))) return rc;
    if ((rc = make_row_tensor_maps(ctx, &b->tmaps_scratch, b->d_scratch, b->n_gather, pn->k_elems, pn->k_stride,
                                   b->fkind == 7 ? MAP_E2M1_PACKED : b->fkind == 6 ? MAP_E2M1_EXPAND : MAP_INT8)))
      return rc;
  } else {
    b->tmaps_scratch = b->fkind == 7 ? pn->tmaps_packed : pn->tmaps;
  }

  gp.tiles = b->d_tiles;
  gp.n_tiles = (int)b->h_tiles.size();
  gp.st_sx_t = b->d_st_sx_t;
  gp.st_sx_u = b->d_st_sx_u;
  gp.st_mean_t = b->d_st_mean_t;
  gp.st_mean_u = b->d_st_mean_u;
  gp.st_ld_t = b->n_t_total;
  gp.st_ld_u = b->n_u_total;
  gp.sd_t = b->d_sd_t;
  gp.sd_u = b->d_sd_u;
  gp.pool_t = b->d_pool_t;
  gp.pool_u = b->d_pool_u;
  gp.out_tt = b->d_tt;
  gp.out_ut = b->d_ut;
  gp.out_counts = b->d_counts;
  gp.counts_seg_stride = b->counts_elems;
  gp.oz_pa = (b->ozaki && gp.raw_out) ? b->d_oz_pa : nullptr;   // the finish pass writes B21's digit planes itself
  gp.oz_kpad = b->oz_kpad;
  gp.oz_nan = b->d_oz_nan;
  if (!sync) return GB_OK;
  // the same sync makes the pack kernels' representability flag readable
  int h_flags = 0;
  if (!b->defer_flag_check)
    GB_CUDA(cudaMemcpyAsync(&h_flags, pn->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_flags & 1) return unrepresentable(ctx);
  if (b->fkind == 0 && !(h_flags & 2) && !b->defer_flag_check) {
    // every packed byte is a dosage in {0, 1, 2}: |m*sumxy - sumx*sumy| <= 4 m^2
    const long long m = (long long)*std::max_element(pn->pop_sizes.begin(), pn->pop_sizes.end());
    gp.wide_fold = 4 * m * m > 2147483647ll ? 1 : 0;
  }
  return GB_OK;
}

// pack5 host rows (five base-3 digits per byte): population p starts at byte boff[p], blocks padded to 4 bytes, the row
// to 16.  Returns the row size in bytes (-1: bad sizes).
int pack5_layout(int n_pops, const int* pop_sizes, std::vector<int>* boff) {
  long long b = 0;
  if (boff) boff->clear();
  for (int i = 0; i < n_pops; i++) {
    if (pop_sizes[i] < 1) return -1;
    if (boff) boff->push_back((int)b);
    b += ((pop_sizes[i] + 4) / 5 + 3) / 4 * 4;
    if (b > (1ll << 30)) return -1;
  }
  return (int)((b + 15) / 16 * 16);
}


// Results leave in two halves so the pipelined path can enqueue the copies at submit time and
// interpret them after its own event wait.
int batch_fetch_enqueue(gb_batch* b, double* z_u, double* info_u, int* status_staging) {
  Ctx* ctx = b->ctx;
  if (b->n_u_total) {
    if (z_u) GB_CUDA(cudaMemcpyAsync(z_u, b->d_zu, sizeof(double) * (size_t)b->n_u_total, cudaMemcpyDeviceToHost, ctx->stream));
    if (info_u) GB_CUDA(cudaMemcpyAsync(info_u, b->d_info, sizeof(double) * (size_t)b->n_u_total, cudaMemcpyDeviceToHost, ctx->stream));
  }
  // d_status is indexed by position in the factorisation list: [real windows | certificate copies]
  const size_t n_st = 2 * (size_t)b->n_windows + 2;
  GB_CUDA(cudaMemcpyAsync(status_staging, b->d_status, sizeof(int) * n_st, cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaMemcpyAsync(status_staging + n_st, b->panel->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  return GB_OK;
}

int batch_fetch_finish(gb_batch* b, const int* st, double* z_u, double* info_u, int* window_status_out) {
  const size_t nreal = b->h_wins.size();
  const size_t n_st = 2 * (size_t)b->n_windows + 2;
  if (st[n_st] & 1) return unrepresentable(b->ctx);
  std::vector<int> st_w((size_t)b->n_windows, 0), st_pd_w((size_t)b->n_windows, 0);
  for (size_t a = 0; a < b->active.size(); a++) {
    st_w[(size_t)b->active[a]] = st[a];
    if (b->params.check_pd) st_pd_w[(size_t)b->active[a]] = st[nreal + a];
  }
  int worst = GB_OK;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (int64_t w = 0; w < b->n_windows; w++) {
    int s = b->plan_status[(size_t)w];
    bool no_result = s != GB_OK;                      // skipped window: no result exists
    if (s == GB_OK && st_w[(size_t)w]) {              // the factorisation of B11 itself broke down: what the solve wrote is garbage
      s = GB_ERR_BREAKDOWN;
      no_result = true;
    } else if (s == GB_OK && st_pd_w[(size_t)w]) {
      s = GB_ERR_NOT_PD;                              // results exist (computed without the clip), not certified
    }
    if (window_status_out) window_status_out[w] = s;
    if (s != GB_OK && worst == GB_OK) worst = s;
    if (no_result)
      for (int64_t i = b->u_off[w]; i < b->u_off[w + 1]; i++) {
        if (z_u) z_u[i] = nan;
        if (info_u) info_u[i] = nan;
      }
  }
  return window_status_out ? GB_OK : worst;
}


gb_batch* batch_new(gb_ctx* ctx, gb_panel* panel, int64_t n_windows, const double* pop_wgt, const gb_params* params,
                    bool ld_mode, bool counts_mode, bool defer_flag_check, double ld_diag) {
  gb_batch* b = new (std::nothrow) gb_batch();
  if (!b) return nullptr;
  b->ctx = ctx;
  b->panel = panel;
  b->mode = (pop_wgt || counts_mode) ? GRAM_MIX : GRAM_POOLED;
  b->ld_mode = ld_mode;
  b->ld_diag = ld_diag;
  b->counts_mode = counts_mode;
  if (params) b->params = *params;
  else gb_params_default(&b->params);
  if (getenv("GB_NO_CERT")) b->params.check_pd = 0;   // diagnostics (timing only): no certificate windows in the factorisation launches
  b->n_windows = n_windows;
  b->defer_flag_check = defer_flag_check;
  return b;
}

}  // namespace gb

namespace {

int check_device(Ctx* ctx) {
  GB_CUDA(cudaSetDevice(ctx->device));
  return GB_OK;
}

int run_stage_impl(gb_batch* b, int stage);

// NVTX range per stage: shows up in nsys / ncu timelines, costs nothing without a profiler attached
int run_stage(gb_batch* b, int stage) {
  static const char* const names[] = {"gb:row_stats", "gb:gram", "gb:cholesky", "gb:solve"};
  nvtxRangePushA(stage >= 0 && stage < 4 ? names[stage] : stage == 10 ? "gb:gram_mma" : stage == 11 ? "gb:gram_finish" : stage == 20 ? "gb:cholesky_only" : "gb:trtri");
  const int rc = run_stage_impl(b, stage);
  nvtxRangePop();
  return rc;
}

int run_stage_impl(gb_batch* b, int stage) {
  Ctx* ctx = b->ctx;
  Panel* pn = b->panel;
  int rc = check_device(ctx);
  if (rc) return rc;
  switch (stage) {
    case 0: {
      if (b->ozaki) GB_CUDA(cudaMemsetAsync(b->d_oz_nan, 0, (size_t)std::max<int64_t>(b->n_u_total, 1), ctx->stream));
      if (b->n_gather > 0)
        if ((rc = launch_gather_rows(ctx, pn, b->d_gather, b->n_gather, b->d_scratch))) return rc;
      if (b->counts_mode) return GB_OK;
      if ((rc = launch_row_prep(ctx, pn, b->d_rows_t, b->n_t_total, b->mode, b->d_coef, b->d_wgt, b->d_sd_t,
                                b->d_pool_t, b->d_rq_t, b->d_st_sx_t, b->d_st_mean_t)))
        return rc;
      return launch_row_prep(ctx, pn, b->d_rows_u, b->n_u_total, b->mode, b->d_coef, b->d_wgt, b->d_sd_u,
                             b->d_pool_u, nullptr, b->d_st_sx_u, b->d_st_mean_u);
    }
    case 1:
      if ((rc = launch_gram(ctx, b->fkind == 7 ? pn->tmaps_packed : pn->tmaps, b->tmaps_scratch, b->gp, 0)))
        return rc;
      return b->gp.raw_out ? launch_gram_finalize(ctx, b->gp, (int)b->h_tiles.size()) : GB_OK;
    case 10:  // profiling: the tensor-core kernel alone
      return launch_gram(ctx, b->fkind == 7 ? pn->tmaps_packed : pn->tmaps, b->tmaps_scratch, b->gp, 0);
    case 11:  // profiling: the finish pass alone (a no-op for panels whose finish is fused)
      return b->gp.raw_out ? launch_gram_finalize(ctx, b->gp, (int)b->h_tiles.size()) : GB_OK;
    case 2:
    case 20: {
      if (b->ld_mode || b->counts_mode) return GB_OK;
      const int nreal = (int)b->h_wins.size();
      GB_CUDA(cudaMemsetAsync(b->d_status, 0, sizeof(int) * (2 * (size_t)b->n_windows + 2), ctx->stream));
      const int* skip = nullptr;
      if (b->clip_mode) {
        // MakePosDef proper (util.cpp:302-318): eigendecomposition, spectrum clipped from below, in place
        if ((rc = launch_eig_jacobi(ctx, b->d_wins, nreal, b->max_nt, b->d_tt, b->d_eig_G, b->d_eig_V, b->d_evals,
                                    b->params.min_abs_eig, 1, nullptr)))
          return rc;
      }
      if (b->params.check_pd) {
        // certificate that MakePosDef is a no-op: the analytic lower bound on lambda_min(B11) when it
        // applies, else a Cholesky of B11 - min_abs_eig*I (succeeds <=> lambda_min > min_abs_eig)
        if ((rc = launch_pd_bound(ctx, b->d_wins, nreal, b->d_rq_t, b->params.lambda, b->gneg,
                                  b->params.min_abs_eig, b->d_skip)))
          return rc;
        if ((rc = launch_copy_shift(ctx, b->d_wins, nreal, b->d_tt, b->d_tt + b->tt_elems, b->params.min_abs_eig,
                                    b->d_skip)))
          return rc;
        skip = b->d_skip;
      }
      // int8-split solve: X = L^-1 and y = L^-1 z_t grow row block by row block on an auxiliary stream while the
      // factorisation takes its own steps (linv_row_kernel); stage 20 (profiling) factors only, stage 21 inverts only
      const bool with_linv = stage == 2 && b->ozaki && !b->d_y;
      LinvArgs la{b->d_wins, nreal, b->d_tt, b->d_dinv, b->d_x, b->d_zt, b->d_oz_y, b->d_oz_amax};
      if (with_linv) GB_CUDA(cudaMemsetAsync(b->d_oz_amax, 0, sizeof(unsigned long long) * (size_t)nreal, ctx->stream));
      if ((rc = launch_cholesky(ctx, b->d_wins, b->n_chol_wins, b->max_nt, b->d_tt, b->d_dinv, b->d_status, skip,
                                with_linv ? &la : nullptr)))
        return rc;
      return GB_OK;
    }
    case 21: {   // profiling: the explicit inverse alone, after stage 20
      if (b->ld_mode || b->counts_mode || !(b->ozaki && !b->d_y)) return GB_OK;
      const int nw = (int)b->h_wins.size();
      LinvArgs la{b->d_wins, nw, b->d_tt, b->d_dinv, b->d_x, b->d_zt, b->d_oz_y, b->d_oz_amax};
      GB_CUDA(cudaMemsetAsync(b->d_oz_amax, 0, sizeof(unsigned long long) * (size_t)nw, ctx->stream));
      return launch_linv_rows(ctx, la, b->max_nt);
    }
    case 3:
      if (b->ld_mode || b->counts_mode) return GB_OK;
      if (b->ozaki && !b->d_y)
        // W^T = B21 X^T as an exact digit-split GEMM on tcgen05 (gb_ozaki.cu); X and y come from stage 2
        return launch_ozaki_solve(ctx, b->d_wins, b->d_oz_wins, b->h_oz_wins.data(), (int)b->h_wins.size(), b->d_oz_tiles,
                                  b->oz_n_tiles, b->oz_kpad, b->d_x, b->d_ut, b->gp.oz_pa == nullptr, b->d_oz_pa, b->oz_a_rows,
                                  b->d_oz_pb, b->oz_b_rows, b->d_oz_amax, b->d_oz_ex, b->d_oz_y, b->d_oz_nan, b->d_zu, b->d_info);
      return launch_trsm_finalize(ctx, b->d_wins, (int)b->h_wins.size(), b->max_nt, b->max_nu, b->d_tt, b->d_dinv,
                                  b->d_ut, b->d_zt, b->d_zu, b->d_info, b->d_y);
    default:
      ctx->err = "unknown stage";
      return GB_ERR_BAD_ARG;
  }
}

int create_batch_internal(gb_ctx* ctx, gb_panel* panel, int64_t n_windows, const int64_t* t_off,
                          const int64_t* rows_t, const int64_t* u_off, const int64_t* rows_u, const double* z_t,
                          const double* pop_wgt, const gb_params* params, bool ld_mode, bool counts_mode,
                          gb_batch** out, bool defer_flag_check = false, double ld_diag = 1.0) {
  if (!ctx || !panel || !out || n_windows < 0 || !t_off || (!rows_t && t_off[n_windows] > 0)) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  if (panel->ctx != ctx) {
    ctx->err = "panel belongs to another context";
    return GB_ERR_BAD_ARG;
  }
  int rc = check_device(ctx);
  if (rc) return rc;
  gb_batch* b = gb::batch_new(ctx, panel, n_windows, pop_wgt, params, ld_mode, counts_mode, defer_flag_check, ld_diag);
  if (!b) return GB_ERR_OOM;
  rc = gb::batch_plan_host(b, t_off, rows_t, u_off, rows_u, z_t, pop_wgt);
  if (!rc) rc = gb::batch_plan_device(b, gb::Arena{}, /*sync=*/true);
  if (rc) {
    batch_free_device(b);
    delete b;
    return rc;
  }
  *out = b;
  return GB_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

void gb_params_default(gb_params* p) {  // gauss.cpp:18-35
  if (!p) return;
  p->lambda = 0.1;
  p->min_abs_eig = 1e-5;
  p->min_num_measured_snp = 10;
  p->min_num_unmeasured_snp = 10;
  p->check_pd = 1;
  p->reserved = 0;
}

int gb_version(void) { return GB_VERSION; }

const char* gb_status_string(int s) {
  switch (s) {
    case GB_OK: return "ok";
    case GB_ERR_BAD_ARG: return "bad argument";
    case GB_ERR_CUDA: return "CUDA error";
    case GB_ERR_NO_DEVICE: return "no CUDA device (gauss_b200 has no CPU fallback)";
    case GB_ERR_OOM: return "out of memory";
    case GB_ERR_TOO_FEW_MEASURED:
    case GB_ERR_TOO_FEW_UNMEASURED: return "Not enough number of SNPs loaded";
    case GB_ERR_NOT_PD: return "B11 not certified positive definite above min_abs_eig";
    case GB_ERR_UNSUPPORTED: return "unsupported configuration";
    case GB_ERR_BREAKDOWN: return "Cholesky of B11 broke down (not positive definite): no result";
    default: return "unknown status";
  }
}

int gb_ctx_create(int device, gb_ctx** out) {
  if (!out) return GB_ERR_BAD_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return GB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    g_create_err = "device index out of range";
    return GB_ERR_BAD_ARG;
  }
  gb_ctx* ctx = new (std::nothrow) gb_ctx();
  if (!ctx) return GB_ERR_OOM;
  ctx->device = device;
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    delete ctx;
    return GB_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_err = "gauss_b200 kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                   std::to_string(prop.minor);
    delete ctx;
    return GB_ERR_UNSUPPORTED;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    delete ctx;
    return GB_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  if (const char* e = getenv("GB_PANEL_FORMAT")) {
    if (!strcmp(e, "int8")) ctx->panel_format = GB_PANEL_INT8;
    else if (!strcmp(e, "e2m1")) ctx->panel_format = GB_PANEL_E2M1;
  }
  if (const char* e = getenv("GB_SEG_ORDER")) ctx->seg_order = atoi(e);   // tuning knob: 0 panel order, 1 descending, 2 alternating
  if (const char* e = getenv("GB_CHOL_SMS")) ctx->chol_sms = atoi(e);   // tuning knob; 0 = no overlap
  if (const char* e = getenv("GB_GRAM_KIND")) ctx->e2m1_mxf4 = strcmp(e, "f8f6f4") != 0;
  if (const char* e = getenv("GB_SOLVE")) ctx->solve_ozaki = strcmp(e, "fp64") != 0;   // "fp64": the DMMA triangular solve
  {  // keep freed blocks cached in the device's default pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  *out = ctx;
  return GB_OK;
}

void gb_ctx_destroy(gb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->win_panel) {
    gb_panel_destroy(static_cast<gb_panel*>(ctx->win_panel));
    ctx->win_panel = nullptr;
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  for (cudaEvent_t e : ctx->ev_stage)
    if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2; i++) {
    if (ctx->chrom_streams[i]) cudaStreamDestroy(ctx->chrom_streams[i]);
    if (ctx->chrom_sides[i]) cudaStreamDestroy(ctx->chrom_sides[i]);
  }
  delete ctx;
}

int gb_ctx_set_stream(gb_ctx* ctx, void* s) {
  if (!ctx) return GB_ERR_BAD_ARG;
  ctx->stream = s ? static_cast<cudaStream_t>(s) : ctx->own_stream;
  return GB_OK;
}

int gb_ctx_synchronize(gb_ctx* ctx) {
  if (!ctx) return GB_ERR_BAD_ARG;
  GB_CUDA(cudaSetDevice(ctx->device));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  return GB_OK;
}

const char* gb_last_error(const gb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
int64_t gb_ctx_launch_count(const gb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- panel ----------------------------------------------------------------------------------
int gb_panel_create(gb_ctx* ctx, int n_pops, const int* pop_sizes, int64_t capacity_rows, gb_panel** out) {
  return gb_panel_create_fmt(ctx, n_pops, pop_sizes, capacity_rows, ctx ? ctx->panel_format : GB_PANEL_E2M1, out);
}

int gb_panel_format_of(const gb_panel* p) { return p ? p->format : -1; }

int gb_panel_create_fmt(gb_ctx* ctx, int n_pops, const int* pop_sizes, int64_t capacity_rows, int format,
                        gb_panel** out) {
  if (!ctx || !out || !pop_sizes || n_pops < 1 || capacity_rows < 1 ||
      (format != GB_PANEL_INT8 && format != GB_PANEL_E2M1)) {
    if (ctx) ctx->err = "bad panel description";
    return GB_ERR_BAD_ARG;
  }
  if (n_pops > P_MAX) {
    ctx->err = "more than " + std::to_string(P_MAX) + " flagged populations";
    return GB_ERR_UNSUPPORTED;
  }
  if (capacity_rows > (int64_t)std::numeric_limits<int32_t>::max() - 256) return GB_ERR_BAD_ARG;
  int rc = check_device(ctx);
  if (rc) return rc;
  gb_panel* p = new (std::nothrow) gb_panel();
  if (!p) return GB_ERR_OOM;
  p->ctx = ctx;
  p->n_pops = n_pops;
  p->format = format;
  // population blocks start on a K-atom boundary (32 columns); E2M1 rows on a 128-column boundary
  // because the nibble-expanding TMA type addresses global memory in units of 128 elements
  p->seg_align = format == GB_PANEL_E2M1 ? K_BLOCK : K_ATOM;
  int k = 0;
  for (int i = 0; i < n_pops; i++) {
    if (pop_sizes[i] < 1) {
      delete p;
      ctx->err = "population size must be >= 1";
      return GB_ERR_BAD_ARG;
    }
    p->pop_sizes.push_back(pop_sizes[i]);
    p->koff.push_back(k);
    k += round_up(pop_sizes[i], p->seg_align);
    p->n_samples += pop_sizes[i];
  }
  p->k_elems = round_up(k, K_BLOCK);
  p->k_stride = format == GB_PANEL_E2M1 ? p->k_elems / 2 : p->k_elems;
  p->capacity = capacity_rows;
  auto fail = [&](int code) {
    gb_panel_destroy(p);
    return code;
  };
  auto pmalloc = [&](void** ptr, size_t bytes) -> int {
    cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
      ctx->err = std::string("cudaMalloc(panel): ") + cudaGetErrorString(e);
      return e == cudaErrorMemoryAllocation ? GB_ERR_OOM : GB_ERR_CUDA;
    }
    return GB_OK;
  };
  if ((rc = pmalloc((void**)&p->d_rows, (size_t)capacity_rows * p->k_stride))) return fail(rc);
  if ((rc = pmalloc((void**)&p->d_sx, sizeof(int32_t) * (size_t)capacity_rows * n_pops))) return fail(rc);
  if ((rc = pmalloc((void**)&p->d_sxx, sizeof(int32_t) * (size_t)capacity_rows * n_pops))) return fail(rc);
  if ((rc = pmalloc((void**)&p->d_pop_sizes, sizeof(int) * (size_t)n_pops))) return fail(rc);
  if ((rc = pmalloc((void**)&p->d_koff, sizeof(int) * (size_t)n_pops))) return fail(rc);
  if ((rc = pmalloc((void**)&p->d_boff5, sizeof(int) * (size_t)n_pops))) return fail(rc);
  std::vector<int> boff5;
  p->pack5_row_bytes = gb::pack5_layout(n_pops, pop_sizes, &boff5);
  if ((rc = pmalloc((void**)&p->d_flags, sizeof(int)))) return fail(rc);
  cudaMemsetAsync(p->d_flags, 0, sizeof(int), ctx->stream);
  cudaMemcpyAsync(p->d_pop_sizes, p->pop_sizes.data(), sizeof(int) * (size_t)n_pops, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(p->d_koff, p->koff.data(), sizeof(int) * (size_t)n_pops, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(p->d_boff5, boff5.data(), sizeof(int) * (size_t)n_pops, cudaMemcpyHostToDevice, ctx->stream);
  if ((rc = make_row_tensor_maps(ctx, &p->tmaps, p->d_rows, capacity_rows, p->k_elems, p->k_stride,
                                 format == GB_PANEL_E2M1 ? MAP_E2M1_EXPAND : MAP_INT8)))
    return fail(rc);
  if (format == GB_PANEL_E2M1 && (rc = make_row_tensor_maps(ctx, &p->tmaps_packed, p->d_rows, capacity_rows, p->k_elems,
                                                            p->k_stride, MAP_E2M1_PACKED)))
    return fail(rc);
  cudaStreamSynchronize(ctx->stream);
  *out = p;
  return GB_OK;
}

void gb_panel_destroy(gb_panel* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  if (p->d_rows) cudaFree(p->d_rows);
  if (p->d_sx) cudaFree(p->d_sx);
  if (p->d_sxx) cudaFree(p->d_sxx);
  if (p->d_pop_sizes) cudaFree(p->d_pop_sizes);
  if (p->d_koff) cudaFree(p->d_koff);
  if (p->d_boff5) cudaFree(p->d_boff5);
  if (p->d_flags) cudaFree(p->d_flags);
  delete p;
}

int gb_panel_clear(gb_panel* p) {
  if (!p) return GB_ERR_BAD_ARG;
  p->n_rows = 0;
  cudaSetDevice(p->ctx->device);
  cudaMemsetAsync(p->d_flags, 0, sizeof(int), p->ctx->stream);
  return GB_OK;
}
int64_t gb_panel_num_rows(const gb_panel* p) { return p ? p->n_rows : -1; }
int64_t gb_panel_num_samples(const gb_panel* p) { return p ? p->n_samples : -1; }

int gb_panel_append_device(gb_panel* p, int64_t n_rows, const void* dev_rows, int64_t row_stride, int is_ascii) {
  if (!p || n_rows < 0 || (!dev_rows && n_rows > 0) || row_stride < p->n_samples) {
    if (p) p->ctx->err = "bad append arguments";
    return GB_ERR_BAD_ARG;
  }
  Ctx* ctx = p->ctx;
  if (p->n_rows + n_rows > p->capacity) {
    ctx->err = "panel capacity exceeded";
    return GB_ERR_BAD_ARG;
  }
  int rc = check_device(ctx);
  if (rc) return rc;
  if ((rc = launch_pack(ctx, p, dev_rows, row_stride, is_ascii, p->n_rows, n_rows))) return rc;
  p->n_rows += n_rows;
  return GB_OK;
}

int gb_panel_append_host(gb_panel* p, int64_t n_rows, const void* rows, int64_t row_stride, int is_ascii) {
  if (!p || n_rows < 0 || (!rows && n_rows > 0) || row_stride < p->n_samples) {
    if (p) p->ctx->err = "bad append arguments";
    return GB_ERR_BAD_ARG;
  }
  if (n_rows == 0) return GB_OK;
  Ctx* ctx = p->ctx;
  int rc = check_device(ctx);
  if (rc) return rc;
  void* d = nullptr;
  const size_t bytes = (size_t)n_rows * (size_t)row_stride;
  GB_CUDA(cudaMallocAsync(&d, bytes, ctx->stream));
  cudaError_t e = cudaMemcpyAsync(d, rows, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    rc = gb_panel_append_device(p, n_rows, d, row_stride, is_ascii);
  } else {
    ctx->err = cudaGetErrorString(e);
    rc = GB_ERR_CUDA;
  }
  cudaFreeAsync(d, ctx->stream);
  return rc;
}

int gb_panel_append_strings(gb_panel* p, int64_t n_rows, const char* const* pop_strings) {
  if (!p || n_rows < 0 || (!pop_strings && n_rows > 0)) return GB_ERR_BAD_ARG;
  if (n_rows == 0) return GB_OK;
  Ctx* ctx = p->ctx;
  int rc = check_device(ctx);
  if (rc) return rc;
  // The per-population strings of each SNP are concatenated into pinned staging rows, copied and packed.  This is the
  // drop-in seam (a window of run_distmix arrives as ~4,400 SNPs x 21 strings, 140 MB), so it is built for rate: the
  // staging buffer lives in the context (pinning 140 MB per call cost more than everything else), host threads gather
  // disjoint row ranges, and the rows go in chunks so that the copy + pack of one chunk runs while the next is gathered.
  const size_t N = (size_t)p->n_samples;
  const int n_pops = p->n_pops;
  constexpr int64_t CHUNK = 1024;                 // rows per chunk; two chunk buffers
  const size_t need = (size_t)std::min<int64_t>(n_rows, 2 * CHUNK) * N;
  if (ctx->h_stage_cap < need) {
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_cap = 0;
    GB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_stage), need));
    ctx->h_stage_cap = need;
  }
  if (!ctx->ev_stage[0]) {
    GB_CUDA(cudaEventCreateWithFlags(&ctx->ev_stage[0], cudaEventDisableTiming));
    GB_CUDA(cudaEventCreateWithFlags(&ctx->ev_stage[1], cudaEventDisableTiming));
  }
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<long long> bad_row{-1};
  std::atomic<int> bad_pop{-1}, bad_kind{0};   // kind 1: null string, 2: wrong length
  auto gather = [&](char* stage, int64_t r0, int64_t r1) {
    for (int64_t r = r0; r < r1 && bad_row.load(std::memory_order_relaxed) < 0; r++) {
      char* dst = stage + (size_t)(r - r0) * N;
      for (int k = 0; k < n_pops; k++) {
        const char* s = pop_strings[r * n_pops + k];
        // the string must hold exactly the population's individuals: a shorter one would be over-read here, and the
        // reference's CalCor walks x[i].length() characters (util.cpp:55), so a longer one cannot be reproduced by
        // truncating it
        const size_t m = (size_t)p->pop_sizes[(size_t)k];
        const int kind = !s ? 1 : (strnlen(s, m + 1) != m ? 2 : 0);
        if (kind) {
          long long expect = -1;
          if (bad_row.compare_exchange_strong(expect, (long long)r)) {
            bad_pop = k;
            bad_kind = kind;
          }
          return;
        }
        std::memcpy(dst, s, m);
        dst += m;
      }
    }
  };
  int n_chunk = 0;
  for (int64_t c0 = 0; c0 < n_rows && !rc; c0 += CHUNK, n_chunk++) {
    const int64_t c1 = std::min(n_rows, c0 + CHUNK);
    const int buf = n_chunk & 1;
    char* stage = ctx->h_stage + (size_t)buf * (size_t)CHUNK * N;
    if (n_chunk >= 2) GB_CUDA(cudaEventSynchronize(ctx->ev_stage[buf]));   // the copy that last read this buffer is done
    const int n_thr = (int)std::min<int64_t>(std::min<unsigned>(hw, 8u), std::max<int64_t>(1, (c1 - c0) / 64));
    if (n_thr <= 1) {
      gather(stage, c0, c1);
    } else {
      std::vector<std::thread> th;
      const int64_t per = (c1 - c0 + n_thr - 1) / n_thr;
      for (int t = 0; t < n_thr; t++) {
        const int64_t a = c0 + t * per, b = std::min(c1, a + per);
        if (a >= b) continue;
        try {
          th.emplace_back([&, a, b, stage, c0] { gather(stage + (size_t)(a - c0) * N, a, b); });
        } catch (...) {   // no thread to be had: this range is gathered here (nothing may be thrown across the C-ABI)
          gather(stage + (size_t)(a - c0) * N, a, b);
        }
      }
      for (auto& t : th) t.join();
    }
    if (bad_row.load() >= 0) break;
    rc = gb_panel_append_host(p, c1 - c0, stage, (int64_t)N, 1);
    if (!rc) GB_CUDA(cudaEventRecord(ctx->ev_stage[buf], ctx->stream));
  }
  cudaStreamSynchronize(ctx->stream);   // the staging buffers are free for the next call; the panel rows are packed
  if (bad_row.load() >= 0) {
    // rows of earlier chunks were appended: take them back so that a failed call leaves the panel as it found it
    p->n_rows -= std::min<int64_t>(p->n_rows, (int64_t)n_chunk * CHUNK);
    const int k = bad_pop.load();
    if (bad_kind.load() == 1) {
      ctx->err = "null genotype string";
    } else {
      ctx->err = "genotype string of SNP " + std::to_string(bad_row.load()) + ", population " + std::to_string(k) + " does not hold " +
                 std::to_string(p->pop_sizes[(size_t)k]) + " characters";
    }
    return GB_ERR_BAD_ARG;
  }
  return rc;
}

// ---- batches ---------------------------------------------------------------------------------
int gb_batch_create(gb_ctx* ctx, gb_panel* panel, int64_t n_windows, const int64_t* t_off, const int64_t* rows_t,
                    const int64_t* u_off, const int64_t* rows_u, const double* z_t, const double* pop_wgt,
                    const gb_params* params, gb_batch** out) {
  if (!u_off || !z_t) {
    if (ctx) ctx->err = "null argument";
    return GB_ERR_BAD_ARG;
  }
  return create_batch_internal(ctx, panel, n_windows, t_off, rows_t, u_off, rows_u, z_t, pop_wgt, params, false,
                               false, out);
}

// computeLD() blocks as a resident batch (one n x n correlation matrix per "window", diagonal forced to `diag`,
// computeLD.cpp:95-116): lets a caller run / time the Gram + epilogue stages alone (BASELINE config 3).
int gb_batch_create_ld(gb_ctx* ctx, gb_panel* panel, int64_t n_windows, const int64_t* t_off, const int64_t* rows_t,
                       const double* pop_wgt, double diag, gb_batch** out) {
  gb_params p;
  gb_params_default(&p);
  return create_batch_internal(ctx, panel, n_windows, t_off, rows_t, nullptr, nullptr, nullptr, pop_wgt, &p, true, false, out,
                               false, diag);
}

void gb_batch_destroy(gb_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  // a run that failed half-way may have left work on the forked streams: nothing may touch the buffers freed below
  cudaStreamSynchronize(b->ctx->stream);
  if (b->ctx->side_stream) cudaStreamSynchronize(b->ctx->side_stream);
  if (b->ctx->aux_stream) cudaStreamSynchronize(b->ctx->aux_stream);
  batch_free_device(b);
  delete b;
}

// Gram kernel + finish pass over tiles [first, first + count) of the batch's list, on at most max_ctas SMs
// (0 = all of them).
static int run_gram_range(gb_batch* b, int first, int count, int max_ctas, int* tile_counter = nullptr,
                          bool with_kernel = true, bool with_finish = true) {
  if (count <= 0) return GB_OK;
  Ctx* ctx = b->ctx;
  Panel* pn = b->panel;
  GramParams gp = b->gp;
  gp.tiles = b->gp.tiles + first;
  gp.n_tiles = count;
  gp.tile_counter = tile_counter;
  int rc = GB_OK;
  if (with_kernel)
    rc = launch_gram(ctx, b->fkind == 7 ? pn->tmaps_packed : pn->tmaps, b->tmaps_scratch, gp, max_ctas);
  if (rc || !with_finish) return rc;
  return gp.raw_out ? launch_gram_finalize(ctx, gp, count) : GB_OK;
}

// Whole pipeline of the batch.  The Cholesky chain of stage 2 is a latency-bound sequence of small launches
// (three per 64-column block step) that needs B11 only, and B11 is a tenth of the Gram work: the B11 tiles run
// first on every SM, then the factorisation runs on a side stream next to the B21 tiles, which leave
// `chol_sms` SMs free for it (the Gram kernel is persistent with one CTA per SM).  Both join before the solve.
int gb_batch_run(gb_batch* b) {
  if (!b) return GB_ERR_BAD_ARG;
  Ctx* ctx = b->ctx;
  const int n_all = (int)b->h_tiles.size(), n_tt = b->n_tiles_tt;
  const bool overlap = !b->ld_mode && !b->counts_mode && ctx->chol_sms > 0 && n_tt > 0 && n_all - n_tt >= ctx->sm_count &&
                       ctx->sm_count > 2 * ctx->chol_sms;
  if (!overlap) {
    for (int s = 0; s < 4; s++) {
      int rc = run_stage(b, s);
      if (rc) return rc;
    }
    return GB_OK;
  }
  int rc = check_device(ctx);
  if (rc) return rc;
  if (!ctx->side_stream) GB_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  if (!b->ev_fork) {
    GB_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    GB_CUDA(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
  }
  if ((rc = run_stage(b, 0))) return rc;
  if ((rc = run_gram_range(b, 0, n_tt, 0))) return rc;
  // B21 tiles: the main launch owns sm_count - chol_sms SMs (launched first, so its persistent CTAs are placed before
  // the factorisation's many small CTAs arrive) and draws tile ids from a counter; when the factorisation is done, a
  // helper launch of the same kernel takes the SMs it leaves and draws from the same counter, so nothing idles while
  // the main launch finishes.  The finish pass follows the join.
  GB_CUDA(cudaMemsetAsync(b->d_tile_counter, 0, sizeof(int), ctx->stream));
  GB_CUDA(cudaEventRecord(b->ev_fork, ctx->stream));      // B11 is final and the counter is zero: the side stream may start
  GB_CUDA(cudaStreamWaitEvent(ctx->side_stream, b->ev_fork, 0));
  if ((rc = run_gram_range(b, n_tt, n_all - n_tt, ctx->sm_count - ctx->chol_sms, b->d_tile_counter, true, false))) return rc;
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->side_stream;
  rc = run_stage(b, 2);
  if (!rc) rc = run_gram_range(b, n_tt, n_all - n_tt, ctx->chol_sms, b->d_tile_counter, true, false);
  ctx->stream = main_stream;
  if (rc) return rc;
  GB_CUDA(cudaEventRecord(b->ev_join, ctx->side_stream));
  GB_CUDA(cudaStreamWaitEvent(ctx->stream, b->ev_join, 0));
  if ((rc = run_gram_range(b, n_tt, n_all - n_tt, 0, nullptr, false, true))) return rc;
  return run_stage(b, 3);
}

}  // extern "C"

namespace gb {
int batch_run_front(gb_batch* b, int max_ctas) {
  int rc = run_stage(b, 0);
  if (rc || b->counts_mode) return rc;
  return run_gram_range(b, 0, b->n_tiles_tt, max_ctas);
}
int batch_run_chain(gb_batch* b) { return run_stage(b, 2); }
int batch_run_b21(gb_batch* b, int max_ctas) {
  return run_gram_range(b, b->n_tiles_tt, (int)b->h_tiles.size() - b->n_tiles_tt, max_ctas);
}
int batch_run_solve(gb_batch* b) { return run_stage(b, 3); }
}  // namespace gb

extern "C" {

int gb_batch_run_stage(gb_batch* b, int stage) {
  if (!b) return GB_ERR_BAD_ARG;
  return run_stage(b, stage);
}

int gb_batch_fetch(gb_batch* b, double* z_u, double* info_u, int* window_status_out) {
  if (!b || b->ld_mode || b->counts_mode) return GB_ERR_BAD_ARG;
  Ctx* ctx = b->ctx;
  int rc = check_device(ctx);
  if (rc) return rc;
  b->h_status.assign(2 * (size_t)b->n_windows + 3, 0);
  if ((rc = gb::batch_fetch_enqueue(b, z_u, info_u, b->h_status.data()))) return rc;
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  return gb::batch_fetch_finish_repair(b, b->h_status.data(), z_u, info_u, window_status_out);
}

int gb_batch_work(const gb_batch* b, double* gram_ops, double* solve_flops, double* panel_bytes) {
  if (!b) return GB_ERR_BAD_ARG;
  if (gram_ops) *gram_ops = b->work_gram_ops;
  if (solve_flops) *solve_flops = b->work_solve_flops;
  if (panel_bytes) *panel_bytes = b->work_panel_bytes;
  return GB_OK;
}

// ---- single windows -----------------------------------------------------------------------------
static int window_impute(gb_ctx* ctx, gb_panel* panel, int64_t n_t, const int64_t* rows_t, int64_t n_u,
                         const int64_t* rows_u, const double* z_t, const double* pop_wgt, const gb_params* params,
                         double* z_u, double* info_u) {
  if (!ctx || !panel || n_t < 0 || n_u < 0 || !z_u || !info_u || (!z_t && n_t > 0)) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  const int64_t t_off[2] = {0, n_t}, u_off[2] = {0, n_u};
  gb_batch* b = nullptr;
  double dummy = 0.0;
  int rc = gb_batch_create(ctx, panel, 1, t_off, rows_t, u_off, rows_u, z_t ? z_t : &dummy, pop_wgt, params, &b);
  if (rc) return rc;
  if ((rc = gb_batch_run(b)) == GB_OK) rc = gb_batch_fetch(b, z_u, info_u, nullptr);
  gb_batch_destroy(b);
  return rc;
}

int gb_window_dist(gb_ctx* ctx, gb_panel* panel, int64_t n_t, const int64_t* rows_t, int64_t n_u,
                   const int64_t* rows_u, const double* z_t, const gb_params* params, double* z_u, double* info_u) {
  return window_impute(ctx, panel, n_t, rows_t, n_u, rows_u, z_t, nullptr, params, z_u, info_u);
}

int gb_window_distmix(gb_ctx* ctx, gb_panel* panel, int64_t n_t, const int64_t* rows_t, int64_t n_u,
                      const int64_t* rows_u, const double* z_t, const double* pop_wgt, const gb_params* params,
                      double* z_u, double* info_u) {
  if (!pop_wgt) {
    if (ctx) ctx->err = "distmix needs population weights";
    return GB_ERR_BAD_ARG;
  }
  return window_impute(ctx, panel, n_t, rows_t, n_u, rows_u, z_t, pop_wgt, params, z_u, info_u);
}

int gb_window_ld(gb_ctx* ctx, gb_panel* panel, int64_t n, const int64_t* rows, const double* pop_wgt,
                 double* cormat) {
  if (!ctx || !panel || n < 0 || !cormat || !pop_wgt) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  const int64_t t_off[2] = {0, n};
  gb_batch* b = nullptr;
  int rc = create_batch_internal(ctx, panel, 1, t_off, rows, nullptr, nullptr, nullptr, pop_wgt, nullptr, true,
                                 false, &b);
  if (rc) return rc;
  if (b->plan_status[0] != GB_OK) {
    rc = b->plan_status[0];
  } else {
    rc = run_stage(b, 0);
    if (!rc) rc = run_stage(b, 1);
    if (!rc) {
      const SolveWin& w = b->h_wins[0];
      cudaError_t e = cudaMemcpy2DAsync(cormat, sizeof(double) * (size_t)n, b->d_tt + w.off_tt,
                                        sizeof(double) * (size_t)w.ld_t, sizeof(double) * (size_t)n, (size_t)n,
                                        cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) {
        ctx->err = cudaGetErrorString(e);
        rc = GB_ERR_CUDA;
      }
    }
  }
  gb_batch_destroy(b);
  return rc;
}

// ---- per-gene LD blocks (BASELINE config 5: jepeg / jepegmix, gene.cpp:300-316 and 569-587) ---------------
// CorG of every gene in ONE batch: gene g owns rows[g_off[g] .. g_off[g+1]) and gets its n_g x n_g correlation
// matrix (symmetric, so column-major == row-major) with `diag` forced on the diagonal (1 + lambda in
// Gene::CalJepegPval / CalJepegmixPval, 1.0 in computeLD).  pop_wgt == NULL -> pooled CalCor (jepeg), else the
// CalWgtCov correlation (jepegmix).  Blocks are written back to back in gene order.  The <= 6 x 6 category
// algebra and the p-values that follow stay in the Rcpp caller.
int gb_genes_ld(gb_ctx* ctx, gb_panel* panel, int64_t n_genes, const int64_t* g_off, const int64_t* rows,
                const double* pop_wgt, double diag, double* out) {
  if (!ctx || !panel || n_genes < 0 || !g_off || (!rows && g_off[n_genes] > 0) || !out) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  if (n_genes == 0) return GB_OK;
  gb_params p;
  gb_params_default(&p);
  p.min_num_measured_snp = 0;   // a gene may hold a single SNP
  gb_batch* b = nullptr;
  int rc = create_batch_internal(ctx, panel, n_genes, g_off, rows, nullptr, nullptr, nullptr, pop_wgt, &p, true, false, &b,
                                 false, diag);
  if (rc) return rc;
  rc = run_stage(b, 0);
  if (!rc) rc = run_stage(b, 1);
  if (!rc) {
    std::vector<double> tt((size_t)std::max<long long>(b->tt_elems, 1));
    cudaError_t e = cudaMemcpyAsync(tt.data(), b->d_tt, sizeof(double) * (size_t)b->tt_elems, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = cudaGetErrorString(e);
      rc = GB_ERR_CUDA;
    } else {
      std::vector<int64_t> out_off((size_t)n_genes + 1, 0);
      for (int64_t g = 0; g < n_genes; g++) {
        const int64_t n = g_off[g + 1] - g_off[g];
        out_off[(size_t)g + 1] = out_off[(size_t)g] + n * n;
      }
      for (size_t a = 0; a < b->h_wins.size(); a++) {   // h_wins is sorted by size; active[] maps back to the gene
        const SolveWin& w = b->h_wins[a];
        double* dst = out + out_off[(size_t)b->active[a]];
        for (int c = 0; c < w.n_t; c++)
          std::memcpy(dst + (size_t)c * w.n_t, tt.data() + w.off_tt + (size_t)c * w.ld_t, sizeof(double) * (size_t)w.n_t);
      }
    }
  }
  gb_batch_destroy(b);
  return rc;
}

// ---- jepeg() / jepegmix() statistics of every gene in one batch (gene.cpp:288-547, 553-822; jepegmix.cpp:115-139) ------
// CorG of all genes through the Gram path (as gb_genes_ld), then jepeg_gene_kernel: one thread per gene for the
// <= 6-category algebra.  out: [n_genes][16] doubles, layout in include/gauss_b200.h.
int gb_genes_jepeg(gb_ctx* ctx, gb_panel* panel, int64_t n_genes, const int64_t* g_off, const int64_t* rows,
                   const double* pop_wgt, const double* z, const double* info, const double* categ_wgt, double lambda,
                   double min_abs_eig, double categ_cor_cutoff, int denorm_norm_w, double* out) {
  if (!ctx || !panel || n_genes < 0 || !g_off || (!rows && g_off[n_genes] > 0) || !z || !info || !categ_wgt || !out ||
      denorm_norm_w == 0) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  if (n_genes == 0) return GB_OK;
  gb_params p;
  gb_params_default(&p);
  p.min_num_measured_snp = 0;   // a gene may hold a single SNP
  gb_batch* b = nullptr;
  int rc = create_batch_internal(ctx, panel, n_genes, g_off, rows, nullptr, nullptr, nullptr, pop_wgt, &p, true, false, &b,
                                 false, 1.0 + lambda);
  if (rc) return rc;
  const int64_t n_snps = g_off[n_genes];
  const size_t db = jepeg_gene_desc_bytes();
  std::vector<uint8_t> h_desc((size_t)n_genes * db);
  std::vector<char> seen((size_t)n_genes, 0);
  for (size_t a = 0; a < b->h_wins.size(); a++) {   // h_wins is sorted by size; active[] maps back to the gene
    const SolveWin& w = b->h_wins[a];
    const int g = b->active[a];
    jepeg_gene_desc_fill(h_desc.data() + (size_t)g * db, w.off_tt, g_off[g], w.ld_t, w.n_t);
    seen[(size_t)g] = 1;
  }
  for (int64_t g = 0; g < n_genes; g++)
    if (!seen[(size_t)g]) jepeg_gene_desc_fill(h_desc.data() + (size_t)g * db, 0, g_off[g], 0, 0);   // empty gene
  void* d_desc = nullptr;
  double *d_z = nullptr, *d_info = nullptr, *d_cw = nullptr, *d_out = nullptr;
  auto done = [&](int code) {
    for (void* q : {d_desc, (void*)d_z, (void*)d_info, (void*)d_cw, (void*)d_out})
      if (q) cudaFreeAsync(q, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    gb_batch_destroy(b);
    return code;
  };
  const size_t ns = (size_t)std::max<int64_t>(n_snps, 1);
  if (cudaMallocAsync(&d_desc, h_desc.size(), ctx->stream) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void**>(&d_z), sizeof(double) * ns, ctx->stream) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void**>(&d_info), sizeof(double) * ns, ctx->stream) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void**>(&d_cw), sizeof(double) * 6 * ns, ctx->stream) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void**>(&d_out), sizeof(double) * 16 * (size_t)n_genes, ctx->stream) != cudaSuccess) {
    ctx->err = "cudaMallocAsync(jepeg) failed";
    cudaGetLastError();
    return done(GB_ERR_OOM);
  }
  cudaMemcpyAsync(d_desc, h_desc.data(), h_desc.size(), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_z, z, sizeof(double) * (size_t)n_snps, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_info, info, sizeof(double) * (size_t)n_snps, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_cw, categ_wgt, sizeof(double) * 6 * (size_t)n_snps, cudaMemcpyHostToDevice, ctx->stream);
  if ((rc = run_stage(b, 0)) || (rc = run_stage(b, 1))) return done(rc);
  if ((rc = launch_jepeg_genes(ctx, d_desc, (int)n_genes, b->d_tt, d_z, d_info, d_cw, min_abs_eig, categ_cor_cutoff,
                               denorm_norm_w, d_out)))
    return done(rc);
  if (cudaMemcpyAsync(out, d_out, sizeof(double) * 16 * (size_t)n_genes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "jepeg results copy failed";
    return done(GB_ERR_CUDA);
  }
  return done(GB_OK);
}

int gb_window_cor(gb_ctx* ctx, gb_panel* panel, int64_t n_t, const int64_t* rows_t, int64_t n_u,
                  const int64_t* rows_u, const double* pop_wgt, const gb_params* params, double* B11, double* B21) {
  if (!ctx || !panel || n_t < 1 || n_u < 0) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  gb_params p;
  if (params) p = *params;
  else gb_params_default(&p);
  p.min_num_measured_snp = 0;  // debug surface: no thresholds
  p.min_num_unmeasured_snp = -1;
  const int64_t t_off[2] = {0, n_t}, u_off[2] = {0, n_u};
  std::vector<double> zt((size_t)n_t, 0.0);
  gb_batch* b = nullptr;
  const int keep_solver = ctx->solve_ozaki;   // this surface returns B21 as doubles: not the digit planes of the int8-split solve
  ctx->solve_ozaki = 0;
  int rc = gb_batch_create(ctx, panel, 1, t_off, rows_t, u_off, rows_u, zt.data(), pop_wgt, &p, &b);
  ctx->solve_ozaki = keep_solver;
  if (rc) return rc;
  rc = run_stage(b, 0);
  if (!rc) rc = run_stage(b, 1);
  if (!rc) {
    const SolveWin& w = b->h_wins[0];
    std::vector<double> tt((size_t)n_t * w.ld_t), ut((size_t)n_t * w.ld_u);
    cudaError_t e = cudaMemcpyAsync(tt.data(), b->d_tt + w.off_tt, tt.size() * sizeof(double),
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && n_u > 0)
      e = cudaMemcpyAsync(ut.data(), b->d_ut + w.off_ut, ut.size() * sizeof(double), cudaMemcpyDeviceToHost,
                          ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = cudaGetErrorString(e);
      rc = GB_ERR_CUDA;
    } else {
      if (B11)
        for (int64_t i = 0; i < n_t; i++)
          for (int64_t j = 0; j <= i; j++) {
            const double v = tt[(size_t)j * w.ld_t + i];  // column-major lower
            B11[i * n_t + j] = v;
            B11[j * n_t + i] = v;
          }
      if (B21)
        for (int64_t u = 0; u < n_u; u++)
          for (int64_t t = 0; t < n_t; t++) B21[u * n_t + t] = ut[(size_t)t * w.ld_u + u];
    }
  }
  gb_batch_destroy(b);
  return rc;
}

int gb_gram_counts(gb_ctx* ctx, gb_panel* panel, int64_t n_a, const int64_t* rows_a, int64_t n_b,
                   const int64_t* rows_b, int32_t* out_sxy, int32_t* out_sx, int32_t* out_sxx) {
  if (!ctx || !panel || n_a < 1 || n_b < 1 || !rows_a || !rows_b) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  const int64_t t_off[2] = {0, n_b}, u_off[2] = {0, n_a};
  gb_batch* b = nullptr;
  int rc = create_batch_internal(ctx, panel, 1, t_off, rows_b, u_off, rows_a, nullptr, nullptr, nullptr, false,
                                 true, &b);
  if (rc) return rc;
  rc = run_stage(b, 0);
  if (!rc) rc = run_stage(b, 1);
  if (!rc) {
    cudaError_t e = cudaSuccess;
    if (out_sxy)
      e = cudaMemcpyAsync(out_sxy, b->d_counts, sizeof(int32_t) * (size_t)(n_a * n_b * panel->n_pops),
                          cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && (out_sx || out_sxx)) {
      std::vector<int32_t> col((size_t)panel->n_rows);
      for (int p = 0; p < panel->n_pops && e == cudaSuccess; p++) {
        for (int which = 0; which < 2 && e == cudaSuccess; which++) {
          int32_t* dst = which ? out_sxx : out_sx;
          if (!dst) continue;
          e = cudaMemcpy(col.data(), (which ? panel->d_sxx : panel->d_sx) + (size_t)p * panel->capacity,
                         sizeof(int32_t) * (size_t)panel->n_rows, cudaMemcpyDeviceToHost);
          for (int64_t i = 0; i < n_a; i++) dst[(size_t)p * n_a + i] = col[(size_t)rows_a[i]];
        }
      }
    }
    if (e != cudaSuccess) {
      ctx->err = cudaGetErrorString(e);
      rc = GB_ERR_CUDA;
    }
  }
  gb_batch_destroy(b);
  return rc;
}


// ---- prep_zmix5 pair correlations -------------------------------------------
// zmix.cpp:151-170 walks every SNP pair and calls CalCor(std::string&, std::string&) once per population.  Here the
// Gram kernel's GRAM_COUNTS mode gives the exact per-population counts of all pairs in one launch and
// zmix_pair_kernel turns them into the reference's output matrix.
int gb_zmix_pair_cor(gb_ctx* ctx, gb_panel* panel, int64_t n, const int64_t* rows, const double* z, double* out) {
  if (!ctx || !panel || n < 2 || !rows || !z || !out || n > 46340) {
    if (ctx) ctx->err = "null argument or n outside [2, 46340]";
    return GB_ERR_BAD_ARG;
  }
  const int64_t t_off[2] = {0, n}, u_off[2] = {0, n};
  gb_batch* b = nullptr;
  int rc = create_batch_internal(ctx, panel, 1, t_off, rows, u_off, rows, nullptr, nullptr, nullptr, false, true, &b);
  if (rc) return rc;
  const size_t n_out = (size_t)(n * (n - 1) / 2) * (size_t)(1 + panel->n_pops);
  double *d_out = nullptr, *d_z = nullptr;
  auto done = [&](int code) {
    if (d_out) cudaFreeAsync(d_out, ctx->stream);
    if (d_z) cudaFreeAsync(d_z, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    gb_batch_destroy(b);
    return code;
  };
  if (cudaMallocAsync(reinterpret_cast<void**>(&d_out), sizeof(double) * n_out, ctx->stream) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void**>(&d_z), sizeof(double) * (size_t)n, ctx->stream) != cudaSuccess) {
    ctx->err = "cudaMallocAsync(zmix pair matrix) failed";
    cudaGetLastError();
    return done(GB_ERR_OOM);
  }
  if (cudaMemcpyAsync(d_z, z, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return done(GB_ERR_CUDA);
  if ((rc = run_stage(b, 0)) || (rc = run_stage(b, 1))) return done(rc);
  if ((rc = launch_zmix_pairs(ctx, panel, b->d_counts, (int)n, b->d_rows_t, d_z, d_out))) return done(rc);
  if (cudaMemcpyAsync(out, d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "zmix pair matrix copy failed";
    return done(GB_ERR_CUDA);
  }
  return done(GB_OK);
}

// ---- 2-bit host panel format ("pack2") --------------------------------------------------------------
// A dosage in {0,1,2} needs two bits; a char/int8 host row spends eight, and PCIe (~50 GB/s) is the
// longest leg of any host-buffer path.  pack2 rows hold 4 dosages per byte at the column positions of the
// E2M1 device row (population blocks on 128-dosage boundaries, zero padded), so the device side is a pure
// bit expansion (expand2_rows_kernel).  This is the format a cached packed panel would be stored in.
static int pack2_k_elems(int n_pops, const int* pop_sizes) {
  long long k = 0;
  for (int i = 0; i < n_pops; i++) {
    if (pop_sizes[i] < 1) return -1;
    k += round_up(pop_sizes[i], K_BLOCK);
  }
  return k > (1ll << 30) ? -1 : round_up((int)k, K_BLOCK);
}

int64_t gb_pack2_row_bytes(int n_pops, const int* pop_sizes) {
  if (n_pops < 1 || !pop_sizes) return -1;
  const int k = pack2_k_elems(n_pops, pop_sizes);
  return k < 0 ? -1 : k / 4;
}

int gb_pack2_rows_host(int n_pops, const int* pop_sizes, int64_t n_rows, const void* rows, int64_t row_stride,
                       int is_ascii, void* out, int64_t out_stride) {
  if (n_pops < 1 || !pop_sizes || n_rows < 0 || (n_rows && (!rows || !out))) return GB_ERR_BAD_ARG;
  const int64_t rb = gb_pack2_row_bytes(n_pops, pop_sizes);
  int64_t n_samples = 0;
  for (int i = 0; i < n_pops; i++) n_samples += pop_sizes[i];
  if (rb < 0 || out_stride < rb || row_stride < n_samples) return GB_ERR_BAD_ARG;
  const int sub = is_ascii ? 48 : 0;
  std::atomic<int> bad{0};
  auto work = [&](int64_t r0, int64_t r1) {
    for (int64_t r = r0; r < r1; r++) {
      const uint8_t* src = static_cast<const uint8_t*>(rows) + r * row_stride;
      uint8_t* dst = static_cast<uint8_t*>(out) + r * out_stride;
      std::memset(dst, 0, (size_t)rb);
      int64_t col = 0;
      for (int p = 0; p < n_pops; p++) {
        const int m = pop_sizes[p];
        uint8_t* d = dst + col / 4;
        int j = 0;
        for (; j + 4 <= m; j += 4) {
          const unsigned a = (unsigned)(uint8_t)(src[j] - sub), b = (unsigned)(uint8_t)(src[j + 1] - sub),
                         c = (unsigned)(uint8_t)(src[j + 2] - sub), e = (unsigned)(uint8_t)(src[j + 3] - sub);
          if ((a | b | c | e) > 3u || a == 3u || b == 3u || c == 3u || e == 3u) bad.store(1, std::memory_order_relaxed);
          d[j >> 2] = (uint8_t)((a & 3u) | ((b & 3u) << 2) | ((c & 3u) << 4) | ((e & 3u) << 6));
        }
        for (; j < m; j++) {
          const unsigned a = (unsigned)(uint8_t)(src[j] - sub);
          if (a > 2u) bad.store(1, std::memory_order_relaxed);
          d[j >> 2] |= (uint8_t)((a & 3u) << (2 * (j & 3)));
        }
        src += m;
        col += round_up(m, K_BLOCK);
      }
    }
  };
  const unsigned hw = std::thread::hardware_concurrency();
  const int nth = (int)std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 1, n_rows / 256));
  if (nth <= 1) {
    work(0, n_rows);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nth; t++) {
      try {
        th.emplace_back(work, n_rows * t / nth, n_rows * (t + 1) / nth);
      } catch (...) {   // no thread to be had: this range runs here (nothing may be thrown across the C-ABI)
        work(n_rows * t / nth, n_rows * (t + 1) / nth);
      }
    }
    for (auto& t : th) t.join();
  }
  return bad.load() ? GB_ERR_UNSUPPORTED : GB_OK;
}

// ---- ternary host rows ("pack5") ---------------------------------------------------------------------
int64_t gb_pack5_row_bytes(int n_pops, const int* pop_sizes) {
  if (n_pops < 1 || !pop_sizes) return -1;
  return gb::pack5_layout(n_pops, pop_sizes, nullptr);
}

int gb_pack5_rows_host(int n_pops, const int* pop_sizes, int64_t n_rows, const void* rows, int64_t row_stride,
                       int is_ascii, void* out, int64_t out_stride) {
  if (n_pops < 1 || !pop_sizes || n_rows < 0 || (n_rows && (!rows || !out))) return GB_ERR_BAD_ARG;
  std::vector<int> boff;
  const int rb = gb::pack5_layout(n_pops, pop_sizes, &boff);
  int64_t n_samples = 0;
  for (int i = 0; i < n_pops; i++) n_samples += pop_sizes[i];
  if (rb < 0 || out_stride < rb || row_stride < n_samples) return GB_ERR_BAD_ARG;
  const int sub = is_ascii ? 48 : 0;
  std::atomic<int> bad{0};
  auto work = [&](int64_t r0, int64_t r1) {
    for (int64_t r = r0; r < r1; r++) {
      const uint8_t* src = static_cast<const uint8_t*>(rows) + r * row_stride;
      uint8_t* dst = static_cast<uint8_t*>(out) + r * out_stride;
      std::memset(dst, 0, (size_t)rb);
      for (int p = 0; p < n_pops; p++) {
        const int m = pop_sizes[p];
        uint8_t* d = dst + boff[(size_t)p];
        int j = 0;
        for (; j + 5 <= m; j += 5) {
          const unsigned a = (uint8_t)(src[j] - sub), b = (uint8_t)(src[j + 1] - sub), c = (uint8_t)(src[j + 2] - sub),
                         e = (uint8_t)(src[j + 3] - sub), f = (uint8_t)(src[j + 4] - sub);
          if (a > 2u || b > 2u || c > 2u || e > 2u || f > 2u) bad.store(1, std::memory_order_relaxed);
          d[j / 5] = (uint8_t)(a + 3u * b + 9u * c + 27u * e + 81u * f);
        }
        unsigned v = 0, mul = 1;
        for (int k = j; k < m; k++, mul *= 3) {
          const unsigned a = (uint8_t)(src[k] - sub);
          if (a > 2u) bad.store(1, std::memory_order_relaxed);
          v += mul * (a % 3u);
        }
        if (j < m) d[j / 5] = (uint8_t)v;
        src += m;
      }
    }
  };
  const unsigned hw = std::thread::hardware_concurrency();
  const int nth = (int)std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 1, n_rows / 256));
  if (nth <= 1) {
    work(0, n_rows);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nth; t++) {
      try {
        th.emplace_back(work, n_rows * t / nth, n_rows * (t + 1) / nth);
      } catch (...) {   // no thread to be had: this range runs here (nothing may be thrown across the C-ABI)
        work(n_rows * t / nth, n_rows * (t + 1) / nth);
      }
    }
    for (auto& t : th) t.join();
  }
  return bad.load() ? GB_ERR_UNSUPPORTED : GB_OK;
}

static int append_packed_host(gb_panel* p, int64_t n_rows, const void* rows_p, int64_t row_stride, int host_format) {
  if (!p || n_rows < 0 || (!rows_p && n_rows > 0)) {
    if (p) p->ctx->err = "bad append arguments";
    return GB_ERR_BAD_ARG;
  }
  Ctx* ctx = p->ctx;
  const int64_t need = host_format == 5 ? p->pack5_row_bytes : p->k_elems / 4;
  if (p->format != GB_PANEL_E2M1 || row_stride < need) {
    ctx->err = "packed host rows need an E2M1 panel and a row stride of at least gb_pack2/5_row_bytes()";
    return GB_ERR_BAD_ARG;
  }
  if (p->n_rows + n_rows > p->capacity) {
    ctx->err = "panel capacity exceeded";
    return GB_ERR_BAD_ARG;
  }
  if (n_rows == 0) return GB_OK;
  int rc = check_device(ctx);
  if (rc) return rc;
  void* d = nullptr;
  const size_t bytes = (size_t)n_rows * (size_t)row_stride;
  GB_CUDA(cudaMallocAsync(&d, bytes, ctx->stream));
  cudaError_t e = cudaMemcpyAsync(d, rows_p, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    rc = host_format == 5 ? launch_expand5(ctx, p, d, row_stride, p->n_rows, n_rows)
                          : launch_expand2(ctx, p, d, row_stride, p->n_rows, n_rows);
    if (!rc) p->n_rows += n_rows;
  } else {
    ctx->err = cudaGetErrorString(e);
    rc = GB_ERR_CUDA;
  }
  cudaFreeAsync(d, ctx->stream);
  return rc;
}

int gb_panel_append_pack5_host(gb_panel* p, int64_t n_rows, const void* rows5, int64_t row_stride) {
  return append_packed_host(p, n_rows, rows5, row_stride, 5);
}

// Same rows already in DEVICE memory (a resident ternary panel, gb_synth_pack5_rows): the expansion kernel alone.
int gb_panel_append_pack5_device(gb_panel* p, int64_t n_rows, const void* dev_rows5, int64_t row_stride) {
  if (!p || n_rows < 0 || (!dev_rows5 && n_rows > 0)) {
    if (p) p->ctx->err = "bad append arguments";
    return GB_ERR_BAD_ARG;
  }
  Ctx* ctx = p->ctx;
  if (p->format != GB_PANEL_E2M1 || row_stride < p->pack5_row_bytes || p->n_rows + n_rows > p->capacity) {
    ctx->err = "pack5 rows need an E2M1 panel with room for them and a row stride of at least gb_pack5_row_bytes()";
    return GB_ERR_BAD_ARG;
  }
  int rc = check_device(ctx);
  if (rc) return rc;
  if ((rc = launch_expand5(ctx, p, dev_rows5, row_stride, p->n_rows, n_rows))) return rc;
  p->n_rows += n_rows;
  return GB_OK;
}

int gb_panel_append_pack2_host(gb_panel* p, int64_t n_rows, const void* rows2, int64_t row_stride) {
  return append_packed_host(p, n_rows, rows2, row_stride, 2);
}

// ---- chromosome driver on pack2 HOST rows -----------------------------------------------------------
// One call = one chromosome (or any bp-sorted run of windows) of dist()/distmix(): the pack2 rows are copied
// to the GPU in n_groups contiguous chunks on a copy stream; the windows are cut into n_groups contiguous,
// cost-balanced batches, and batch g starts as soon as the rows its windows touch have landed and been
// expanded -- so all but the last batch's kernels hide behind the PCIe copy.  Results and statuses are
// written to HOST buffers before the call returns.
static int chrom_run_packed(gb_ctx* ctx, gb_panel* panel, int host_format, int64_t n_rows, const void* host_rows2,
                            int64_t row_stride, int64_t n_windows, const int64_t* t_off, const int64_t* rows_t,
                            const int64_t* u_off, const int64_t* rows_u, const double* z_t, const double* pop_wgt,
                            const gb_params* params, int n_groups, double* z_u, double* info_u, int* window_status) {
  if (!ctx || !panel || n_rows < 0 || n_windows < 0 || !t_off || !u_off || (n_rows && !host_rows2) || !z_u || !info_u ||
      n_groups < 1) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  if (panel->ctx != ctx || panel->format != GB_PANEL_E2M1 || n_rows > panel->capacity ||
      row_stride < (host_format == 5 ? panel->pack5_row_bytes : panel->k_elems / 4)) {
    ctx->err = "chromosome driver needs an E2M1 panel of this context with capacity >= n_rows";
    return GB_ERR_BAD_ARG;
  }
  int rc = check_device(ctx);
  if (rc) return rc;
  if (!ctx->copy_stream) GB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    if (!ctx->chrom_streams[i]) GB_CUDA(cudaStreamCreateWithFlags(&ctx->chrom_streams[i], cudaStreamNonBlocking));
    if (!ctx->chrom_sides[i]) GB_CUDA(cudaStreamCreateWithFlags(&ctx->chrom_sides[i], cudaStreamNonBlocking));
  }
  if (n_groups > n_windows) n_groups = (int)std::max<int64_t>(1, n_windows);
  gb_panel_clear(panel);
  panel->n_rows = n_rows;   // the batches are planned against the full row range before the rows arrive

  // contiguous groups of ~equal cost (Gram ~ n_u n_t + n_t^2/2 samples, solve ~ n_t^2 n_u + n_t^3/3)
  std::vector<double> cost((size_t)n_windows);
  double total = 0;
  const double N = (double)panel->n_samples;
  for (int64_t w = 0; w < n_windows; w++) {
    const double nt = (double)(t_off[w + 1] - t_off[w]), nu = (double)(u_off[w + 1] - u_off[w]);
    cost[(size_t)w] = N * (nu * nt + nt * nt / 2) / 16 + 100.0 * (nt * nt * nu + nt * nt * nt / 3) + 1.0;
    total += cost[(size_t)w];
  }
  std::vector<int64_t> g_lo((size_t)n_groups + 1, n_windows);
  g_lo[0] = 0;
  {
    double acc = 0;
    int g = 1;
    for (int64_t w = 0; w < n_windows && g < n_groups; w++) {
      acc += cost[(size_t)w];
      // even cost shares: tapering the groups so that the last batch is the smallest, or giving it one or two
      // windows only, was measured and does not help (a batch's latency is set by its Cholesky chain and its
      // single wave of solve CTAs, not by its size)
      if (acc >= total * g / n_groups) g_lo[(size_t)g++] = w + 1;
    }
  }
  for (int g = 1; g <= n_groups; g++) g_lo[(size_t)g] = std::max(g_lo[(size_t)g], g_lo[(size_t)g - 1]);

  std::vector<gb_batch*> batches((size_t)n_groups, nullptr);
  std::vector<int64_t> need((size_t)n_groups, 0);
  std::vector<cudaEvent_t> landed((size_t)n_groups, nullptr);
  uint8_t* d_stage = nullptr;
  int* h_status = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->chrom_streams[0]);
    cudaStreamSynchronize(ctx->chrom_streams[1]);
    cudaStreamSynchronize(ctx->stream);
    for (auto b : batches)
      if (b) {
        batch_free_device(b);
        delete b;
      }
    for (auto e : landed)
      if (e) cudaEventDestroy(e);
    if (d_stage) cudaFreeAsync(d_stage, ctx->stream);
    if (h_status) cudaFreeHost(h_status);
  };
  auto fail = [&](int code) {
    cleanup();
    return code;
  };
  // plan every group first (host work + small uploads: 0.65 ms for 6 groups), then let copies and kernels stream
  const bool trace = getenv("GB_CHROM_TRACE") != nullptr;   // diagnostics: when each chunk landed / each batch finished
  std::vector<cudaEvent_t> done_ev;
  cudaEvent_t ev_t0 = nullptr;
  const auto host_t0 = std::chrono::steady_clock::now();
  int64_t need_run = 0;
  for (int g = 0; g < n_groups; g++) {
    const int64_t w0 = g_lo[(size_t)g], w1 = g_lo[(size_t)g + 1], nw = w1 - w0;
    std::vector<int64_t> to((size_t)nw + 1), uo((size_t)nw + 1);
    for (int64_t i = 0; i <= nw; i++) {
      to[(size_t)i] = t_off[w0 + i] - t_off[w0];
      uo[(size_t)i] = u_off[w0 + i] - u_off[w0];
    }
    for (int64_t i = t_off[w0]; i < t_off[w1]; i++) need_run = std::max(need_run, rows_t[i] + 1);
    for (int64_t i = u_off[w0]; i < u_off[w1]; i++) need_run = std::max(need_run, rows_u[i] + 1);
    need[(size_t)g] = g == n_groups - 1 ? n_rows : std::min(need_run, n_rows);
    double dummy = 0.0;
    rc = create_batch_internal(ctx, panel, nw, to.data(), rows_t ? rows_t + t_off[w0] : nullptr, uo.data(),
                               rows_u ? rows_u + u_off[w0] : nullptr, z_t ? z_t + t_off[w0] : &dummy, pop_wgt, params,
                               false, false, &batches[(size_t)g], /*defer_flag_check=*/true);
    if (rc) return fail(rc);
    if (cudaEventCreateWithFlags(&landed[(size_t)g], trace ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess)
      return fail(GB_ERR_CUDA);
  }
  size_t st_total = 0;
  for (int g = 0; g < n_groups; g++) st_total += 2 * (size_t)(g_lo[(size_t)g + 1] - g_lo[(size_t)g]) + 3;
  if (cudaMallocHost(reinterpret_cast<void**>(&h_status), sizeof(int) * st_total) != cudaSuccess) return fail(GB_ERR_CUDA);
  if (n_rows) {
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&d_stage), (size_t)n_rows * (size_t)row_stride, ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = std::string("cudaMallocAsync(pack2 staging): ") + cudaGetErrorString(e);
      d_stage = nullptr;
      return fail(e == cudaErrorMemoryAllocation ? GB_ERR_OOM : GB_ERR_CUDA);
    }
    // the copy stream may not touch the staging buffer before the allocation is ordered
    cudaEvent_t alloc_done;
    if (cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming) != cudaSuccess) return fail(GB_ERR_CUDA);
    cudaEventRecord(alloc_done, ctx->stream);
    cudaStreamWaitEvent(ctx->copy_stream, alloc_done, 0);
    cudaEventDestroy(alloc_done);
  }
  if (trace) {
    cudaEventCreate(&ev_t0);
    cudaEventRecord(ev_t0, ctx->copy_stream);
  }
  const double plan_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
  int64_t have = 0;
  size_t st_off = 0;
  std::vector<size_t> st_offs((size_t)n_groups);
  for (int g = 0; g < n_groups; g++) {
    const int64_t lo = have, hi = need[(size_t)g];
    if (hi > lo) {
      cudaError_t e = cudaMemcpyAsync(d_stage + (size_t)lo * (size_t)row_stride,
                                      static_cast<const uint8_t*>(host_rows2) + (size_t)lo * (size_t)row_stride,
                                      (size_t)(hi - lo) * (size_t)row_stride, cudaMemcpyHostToDevice, ctx->copy_stream);
      if (e != cudaSuccess) {
        ctx->err = cudaGetErrorString(e);
        return fail(GB_ERR_CUDA);
      }
      have = hi;
    }
    cudaEventRecord(landed[(size_t)g], ctx->copy_stream);
    cudaStreamWaitEvent(ctx->stream, landed[(size_t)g], 0);
    if (hi > lo) {
      const uint8_t* chunk = d_stage + (size_t)lo * (size_t)row_stride;
      rc = host_format == 5 ? launch_expand5(ctx, panel, chunk, row_stride, lo, hi - lo)
                            : launch_expand2(ctx, panel, chunk, row_stride, lo, hi - lo);
      if (rc) return fail(rc);
    }
    // The batches alternate between two compute streams: a quarter-chromosome batch cannot fill the GPU by itself
    // (its Cholesky chain is latency-bound, its solve is a single wave of CTAs), so the next batch's Gram kernel
    // runs beside them.  Expansion stays on the context stream; `landed[g]` is re-recorded there as "rows expanded".
    cudaEventRecord(landed[(size_t)g], ctx->stream);
    cudaStream_t cs = ctx->chrom_streams[g & 1];
    cudaStreamWaitEvent(cs, landed[(size_t)g], 0);
    gb_batch* b = batches[(size_t)g];
    cudaStream_t main_stream = ctx->stream, main_side = ctx->side_stream;
    ctx->stream = cs;
    ctx->side_stream = ctx->chrom_sides[g & 1];   // each compute stream forks its factorisation onto its own side stream
    rc = gb_batch_run(b);
    const int64_t w0 = g_lo[(size_t)g];
    st_offs[(size_t)g] = st_off;
    if (!rc) rc = gb::batch_fetch_enqueue(b, z_u + u_off[w0], info_u + u_off[w0], h_status + st_off);
    ctx->stream = main_stream;
    ctx->side_stream = main_side;
    if (rc) return fail(rc);
    st_off += 2 * (size_t)b->n_windows + 3;
    if (trace) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, cs);
      done_ev.push_back(e);
    }
  }
  if (cudaStreamSynchronize(ctx->chrom_streams[0]) != cudaSuccess || cudaStreamSynchronize(ctx->chrom_streams[1]) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "chromosome driver: stream synchronisation failed";
    return fail(GB_ERR_CUDA);
  }
  if (trace) {
    fprintf(stderr, "[chrom trace] plan %.2f ms |", plan_ms);
    for (int g = 0; g < n_groups; g++) {
      float a = 0, c = 0;
      cudaEventElapsedTime(&a, ev_t0, landed[(size_t)g]);
      cudaEventElapsedTime(&c, ev_t0, done_ev[(size_t)g]);
      fprintf(stderr, " group %d: rows expanded %.2f ms, batch done %.2f ms |", g, a, c);
    }
    fprintf(stderr, " total host %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
    for (auto e : done_ev) cudaEventDestroy(e);
    cudaEventDestroy(ev_t0);
  }
  int worst = GB_OK;
  for (int g = 0; g < n_groups; g++) {
    const int64_t w0 = g_lo[(size_t)g];
    gb_batch* b = batches[(size_t)g];
    rc = gb::batch_fetch_finish_repair(b, h_status + st_offs[(size_t)g], z_u + u_off[w0], info_u + u_off[w0],
                                       window_status ? window_status + w0 : nullptr);
    if (rc != GB_OK && worst == GB_OK) worst = rc;
  }
  cleanup();
  return worst;
}


// ---- qcat() / qcatmix() window -------------------------------------------
// Same B11 / B21 as dist() / distmix(); the tail is y = L^-1 Z1, w_s = L^-1 b_s and a Pearson correlation per
// tested SNP s (qcat.cpp:203-250).  The tested MEASURED SNPs (rows_t[core_first .. core_first + n_core)) are
// appended to the unmeasured list as extra right-hand-side columns -- their b_s is a row of B11 -- and the one
// entry where such a column meets itself is patched to the forced diagonal 1 + lambda.  CountPC (util.cpp:355-388)
// returns n_t unless an eigenvalue of B11 lies below eig_cutoff; like MakePosDef in dist() that is certified, not
// computed: analytic bound or a Cholesky of B11 - eig_cutoff*I.  GB_ERR_NOT_PD = not certified (no eigen-count path).
int gb_window_qcat(gb_ctx* ctx, gb_panel* panel, int64_t n_t, const int64_t* rows_t, const double* z_t,
                   int64_t core_first, int64_t n_core, int64_t n_u, const int64_t* rows_u, const double* pop_wgt,
                   const gb_params* params, double eig_cutoff, int* num_eig, double* t_m, double* chisq_m,
                   double* t_u, double* chisq_u) {
  if (!ctx || !panel || n_t < 0 || n_u < 0 || n_core < 0 || core_first < 0 || core_first + n_core > n_t ||
      (n_t && (!rows_t || !z_t)) || (n_u && (!rows_u || !t_u || !chisq_u)) || (n_core && (!t_m || !chisq_m))) {
    if (ctx) ctx->err = "null or inconsistent argument";
    return GB_ERR_BAD_ARG;
  }
  gb_params p;
  if (params) p = *params;
  else gb_params_default(&p);
  p.check_pd = 1;
  p.min_abs_eig = eig_cutoff;         // the certificate threshold plays CountPC's cut-off
  // run_qcat only checks the measured count (qcat.cpp:157); run_qcatmix also refuses a window with too few unmeasured
  // SNPs (qcatmix.cpp:168-169)
  if (pop_wgt && n_u <= p.min_num_unmeasured_snp) {
    ctx->err = "too few unmeasured SNPs in the window";
    return GB_ERR_TOO_FEW_UNMEASURED;
  }
  p.min_num_unmeasured_snp = -1;
  const int64_t n_test = n_u + n_core;
  std::vector<int64_t> ru((size_t)n_test);
  for (int64_t i = 0; i < n_u; i++) ru[(size_t)i] = rows_u[i];
  for (int64_t i = 0; i < n_core; i++) ru[(size_t)(n_u + i)] = rows_t[core_first + i];
  const int64_t t_off[2] = {0, n_t}, u_off[2] = {0, n_test};
  double dummy = 0.0;
  gb_batch* b = nullptr;
  // qcat correlates y with every column of W: it keeps the fp64 triangular solve, which stores W
  const int keep_solver = ctx->solve_ozaki;
  ctx->solve_ozaki = 0;
  int rc = create_batch_internal(ctx, panel, 1, t_off, rows_t, u_off, ru.data(), z_t ? z_t : &dummy, pop_wgt, &p, false,
                                 false, &b);
  ctx->solve_ozaki = keep_solver;
  if (rc) return rc;
  auto done = [&](int code) {
    cudaStreamSynchronize(ctx->stream);
    batch_free_device(b);
    delete b;
    return code;
  };
  if (b->plan_status[0] != GB_OK) return done(b->plan_status[0]);
  double *d_qt = nullptr, *d_qc = nullptr;
  const size_t nb_test = sizeof(double) * (size_t)std::max<int64_t>(n_test, 1);
  if (cudaMallocAsync(reinterpret_cast<void**>(&b->d_y), sizeof(double) * (size_t)std::max<int64_t>(n_t, 1), ctx->stream) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void**>(&d_qt), nb_test, ctx->stream) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void**>(&d_qc), nb_test, ctx->stream) != cudaSuccess) {
    ctx->err = "cudaMallocAsync(qcat) failed";
    return done(GB_ERR_OOM);
  }
  auto done2 = [&](int code) {
    cudaFreeAsync(d_qt, ctx->stream);
    cudaFreeAsync(d_qc, ctx->stream);
    return done(code);
  };
  if ((rc = run_stage(b, 0)) || (rc = run_stage(b, 1))) return done2(rc);
  if ((rc = launch_qcat_patch(ctx, b->d_wins, b->d_ut, (int)n_u, (int)core_first, (int)n_core, 1.0 + p.lambda))) return done2(rc);
  if ((rc = run_stage(b, 2)) || (rc = run_stage(b, 3))) return done2(rc);
  if ((rc = launch_qcat_finalize(ctx, b->d_wins, b->d_ut, b->d_y, (int)n_test, (int)n_t, d_qt, d_qc))) return done2(rc);
  std::vector<double> h_t((size_t)n_test), h_c((size_t)n_test);
  if (n_test) {
    cudaMemcpyAsync(h_t.data(), d_qt, sizeof(double) * (size_t)n_test, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(h_c.data(), d_qc, sizeof(double) * (size_t)n_test, cudaMemcpyDeviceToHost, ctx->stream);
  }
  rc = gb_batch_fetch(b, nullptr, nullptr, nullptr);   // synchronises; window status (breakdown / not certified / format)
  int n_eig = (int)n_t;
  if (rc == GB_ERR_NOT_PD) {
    // CountPC proper (util.cpp:355-388): the certificate could not show that every eigenvalue of B11 lies above
    // eig_cutoff, so count them.  B11 has been overwritten by its factor: rebuild the blocks, eigendecompose B11 on the
    // device, then factor / solve / test again with the right number of components.
    double *d_G = nullptr, *d_V = nullptr, *d_ev = nullptr;
    const size_t tt = (size_t)std::max<long long>(b->tt_elems, 1);
    if (cudaMallocAsync(reinterpret_cast<void**>(&d_G), sizeof(double) * tt, ctx->stream) != cudaSuccess ||
        cudaMallocAsync(reinterpret_cast<void**>(&d_V), sizeof(double) * tt, ctx->stream) != cudaSuccess ||
        cudaMallocAsync(reinterpret_cast<void**>(&d_ev), sizeof(double) * (size_t)std::max<int64_t>(n_t, 1), ctx->stream) != cudaSuccess) {
      ctx->err = "cudaMallocAsync(qcat eigenvalues) failed";
      cudaGetLastError();
      return done2(GB_ERR_OOM);
    }
    std::vector<double> ev((size_t)n_t);
    rc = run_stage(b, 1);
    if (!rc) rc = launch_qcat_patch(ctx, b->d_wins, b->d_ut, (int)n_u, (int)core_first, (int)n_core, 1.0 + p.lambda);
    if (!rc) rc = launch_eig_jacobi(ctx, b->d_wins, 1, b->max_nt, b->d_tt, d_G, d_V, d_ev, eig_cutoff, /*clip=*/0, nullptr);
    if (!rc && cudaMemcpyAsync(ev.data(), d_ev, sizeof(double) * (size_t)n_t, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = GB_ERR_CUDA;
    if (!rc) rc = run_stage(b, 2);
    if (!rc) rc = run_stage(b, 3);
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GB_ERR_CUDA;
    if (!rc) {
      for (double v : ev) n_eig -= v < eig_cutoff;     // util.cpp:380-386
      rc = launch_qcat_finalize(ctx, b->d_wins, b->d_ut, b->d_y, (int)n_test, n_eig, d_qt, d_qc);
    }
    if (!rc && n_test) {
      cudaMemcpyAsync(h_t.data(), d_qt, sizeof(double) * (size_t)n_test, cudaMemcpyDeviceToHost, ctx->stream);
      cudaMemcpyAsync(h_c.data(), d_qc, sizeof(double) * (size_t)n_test, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (!rc) {
      rc = gb_batch_fetch(b, nullptr, nullptr, nullptr);
      if (rc == GB_ERR_NOT_PD) rc = GB_OK;             // counted, not certified: that is the answer now
    }
    cudaFreeAsync(d_G, ctx->stream);
    cudaFreeAsync(d_V, ctx->stream);
    cudaFreeAsync(d_ev, ctx->stream);
  }
  if (rc == GB_OK) {
    for (int64_t i = 0; i < n_u; i++) {
      t_u[i] = h_t[(size_t)i];
      chisq_u[i] = h_c[(size_t)i];
    }
    for (int64_t i = 0; i < n_core; i++) {
      t_m[i] = h_t[(size_t)(n_u + i)];
      chisq_m[i] = h_c[(size_t)(n_u + i)];
    }
    if (num_eig) *num_eig = n_eig;
  }
  return done2(rc);
}

int gb_chrom_run_pack2(gb_ctx* ctx, gb_panel* panel, int64_t n_rows, const void* host_rows2, int64_t row_stride,
                       int64_t n_windows, const int64_t* t_off, const int64_t* rows_t, const int64_t* u_off,
                       const int64_t* rows_u, const double* z_t, const double* pop_wgt, const gb_params* params,
                       int n_groups, double* z_u, double* info_u, int* window_status) {
  return chrom_run_packed(ctx, panel, 2, n_rows, host_rows2, row_stride, n_windows, t_off, rows_t, u_off, rows_u, z_t,
                          pop_wgt, params, n_groups, z_u, info_u, window_status);
}

int gb_chrom_run_pack5(gb_ctx* ctx, gb_panel* panel, int64_t n_rows, const void* host_rows5, int64_t row_stride,
                       int64_t n_windows, const int64_t* t_off, const int64_t* rows_t, const int64_t* u_off,
                       const int64_t* rows_u, const double* z_t, const double* pop_wgt, const gb_params* params,
                       int n_groups, double* z_u, double* info_u, int* window_status) {
  return chrom_run_packed(ctx, panel, 5, n_rows, host_rows5, row_stride, n_windows, t_off, rows_t, u_off, rows_u, z_t,
                          pop_wgt, params, n_groups, z_u, info_u, window_status);
}

// ---- pipelined single windows on HOST buffers ---------------------------------------------------
// dist()/distmix() are called once per window with genotypes that live in host memory.  A gb_pipe
// keeps `depth` device slots (raw staging rows + packed panel); submitting window w+1 starts its
// host->device copy on a dedicated copy stream while window w is still packing / multiplying /
// solving on the ctx stream, so the PCIe copy -- the longest leg of a window -- never waits for compute.
struct gb_pipe {
  Ctx* ctx = nullptr;
  int depth = 0;
  int64_t max_rows = 0;
  int64_t n_samples = 0;
  cudaStream_t copy_stream = nullptr;
  struct Slot {
    gb_panel* panel = nullptr;
    uint8_t* d_stage = nullptr;   // [max_rows][n_samples] raw host rows
    cudaEvent_t h2d_done = nullptr, done = nullptr;
    gb_batch* batch = nullptr;    // in flight
    int* h_status = nullptr;      // pinned
    double *h_z = nullptr, *h_info = nullptr;  // pinned result staging [max_rows] (caller buffers may be pageable,
                                               // and a device->pageable copy would block the submitting thread)
    double *z_u = nullptr, *info_u = nullptr;  // caller's result buffers, filled at wait time
    int64_t n_u = 0;
    int64_t ticket = -1;
    int early_status = GB_OK;     // window rejected at plan time (too few SNPs)
  };
  std::vector<Slot> slots;
  int64_t next_ticket = 0;
};

static int pipe_retire(gb_pipe* pp, gb_pipe::Slot& sl, int* status_out) {
  Ctx* ctx = pp->ctx;
  int rc = sl.early_status;
  if (sl.batch) {
    GB_CUDA(cudaEventSynchronize(sl.done));
    std::memcpy(sl.z_u, sl.h_z, sizeof(double) * (size_t)sl.n_u);
    std::memcpy(sl.info_u, sl.h_info, sizeof(double) * (size_t)sl.n_u);
    rc = gb::batch_fetch_finish_repair(sl.batch, sl.h_status, sl.z_u, sl.info_u, nullptr);
    batch_free_device(sl.batch);
    delete sl.batch;
    sl.batch = nullptr;
  }
  sl.ticket = -1;
  sl.early_status = GB_OK;
  if (status_out) *status_out = rc;
  return GB_OK;
}

int gb_pipe_create(gb_ctx* ctx, int n_pops, const int* pop_sizes, int64_t max_rows_per_window, int depth, int format,
                   gb_pipe** out) {
  if (!ctx || !out || depth < 1 || depth > 8 || max_rows_per_window < 1) {
    if (ctx) ctx->err = "bad pipe description";
    return GB_ERR_BAD_ARG;
  }
  int rc = check_device(ctx);
  if (rc) return rc;
  gb_pipe* pp = new (std::nothrow) gb_pipe();
  if (!pp) return GB_ERR_OOM;
  pp->ctx = ctx;
  pp->depth = depth;
  pp->max_rows = max_rows_per_window;
  pp->slots.resize((size_t)depth);
  auto fail = [&](int code) {
    gb_pipe_destroy(pp);
    return code;
  };
  if (cudaStreamCreateWithFlags(&pp->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(GB_ERR_CUDA);
  for (auto& sl : pp->slots) {
    if ((rc = gb_panel_create_fmt(ctx, n_pops, pop_sizes, max_rows_per_window, format < 0 ? ctx->panel_format : format,
                                  &sl.panel)))
      return fail(rc);
    pp->n_samples = sl.panel->n_samples;
    if (cudaMalloc(reinterpret_cast<void**>(&sl.d_stage), (size_t)max_rows_per_window * (size_t)pp->n_samples) != cudaSuccess)
      return fail(GB_ERR_OOM);
    if (cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&sl.h_status), sizeof(int) * 8) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&sl.h_z), sizeof(double) * (size_t)max_rows_per_window) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&sl.h_info), sizeof(double) * (size_t)max_rows_per_window) != cudaSuccess)
      return fail(GB_ERR_CUDA);
  }
  *out = pp;
  return GB_OK;
}

void gb_pipe_destroy(gb_pipe* pp) {
  if (!pp) return;
  cudaSetDevice(pp->ctx->device);
  for (auto& sl : pp->slots) {
    if (sl.batch) {
      cudaEventSynchronize(sl.done);
      batch_free_device(sl.batch);
      delete sl.batch;
    }
    if (sl.h2d_done) cudaEventDestroy(sl.h2d_done);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.h_status) cudaFreeHost(sl.h_status);
    if (sl.h_z) cudaFreeHost(sl.h_z);
    if (sl.h_info) cudaFreeHost(sl.h_info);
    if (sl.d_stage) cudaFree(sl.d_stage);
    if (sl.panel) gb_panel_destroy(sl.panel);
  }
  if (pp->copy_stream) cudaStreamDestroy(pp->copy_stream);
  delete pp;
}

int gb_pipe_submit(gb_pipe* pp, int64_t n_t, const void* host_rows_t, int64_t n_u, const void* host_rows_u,
                   int64_t row_stride, int is_ascii, const double* z_t, const double* pop_wgt, const gb_params* params,
                   double* z_u, double* info_u, int64_t* ticket) {
  if (!pp || !ticket || n_t < 0 || n_u < 0 || (n_t && !host_rows_t) || (n_u && !host_rows_u) || !z_u || !info_u ||
      (!z_t && n_t > 0) || row_stride < pp->n_samples || n_t + n_u > pp->max_rows) {
    if (pp) pp->ctx->err = "bad pipe submit arguments";
    return GB_ERR_BAD_ARG;
  }
  Ctx* ctx = pp->ctx;
  int rc = check_device(ctx);
  if (rc) return rc;
  gb_pipe::Slot& sl = pp->slots[(size_t)(pp->next_ticket % pp->depth)];
  if (sl.ticket >= 0 && (rc = pipe_retire(pp, sl, nullptr))) return rc;  // slot still owned by an un-waited ticket
  const size_t N = (size_t)pp->n_samples;
  // 1. raw rows -> device staging on the copy stream (pinned host memory makes this asynchronous)
  if (n_t) GB_CUDA(cudaMemcpy2DAsync(sl.d_stage, N, host_rows_t, (size_t)row_stride, N, (size_t)n_t,
                                     cudaMemcpyHostToDevice, pp->copy_stream));
  if (n_u) GB_CUDA(cudaMemcpy2DAsync(sl.d_stage + (size_t)n_t * N, N, host_rows_u, (size_t)row_stride, N, (size_t)n_u,
                                     cudaMemcpyHostToDevice, pp->copy_stream));
  GB_CUDA(cudaEventRecord(sl.h2d_done, pp->copy_stream));
  // 2. plan the window (host work + descriptor uploads; synchronises the ctx stream, i.e. at most
  //    the previous window's kernels -- the copy above keeps running meanwhile)
  gb_panel_clear(sl.panel);
  sl.panel->n_rows = n_t + n_u;
  std::vector<int64_t> rt((size_t)n_t), ru((size_t)n_u);
  for (int64_t i = 0; i < n_t; i++) rt[(size_t)i] = i;
  for (int64_t i = 0; i < n_u; i++) ru[(size_t)i] = n_t + i;
  const int64_t t_off[2] = {0, n_t}, u_off[2] = {0, n_u};
  double dummy = 0.0;
  gb_batch* b = nullptr;
  sl.z_u = z_u;
  sl.info_u = info_u;
  sl.n_u = n_u;
  sl.ticket = pp->next_ticket;
  *ticket = pp->next_ticket++;
  rc = create_batch_internal(static_cast<gb_ctx*>(ctx), sl.panel, 1, t_off, rt.data(), u_off, ru.data(),
                             z_t ? z_t : &dummy, pop_wgt, params, false, false, &b, /*defer_flag_check=*/true);
  if (rc) {
    sl.early_status = rc;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int64_t i = 0; i < n_u; i++) z_u[i] = info_u[i] = nan;
    return GB_OK;  // reported by gb_pipe_wait, like a window the reference refuses
  }
  if (b->plan_status[0] != GB_OK) {
    sl.early_status = b->plan_status[0];
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int64_t i = 0; i < n_u; i++) z_u[i] = info_u[i] = nan;
    batch_free_device(b);
    delete b;
    return GB_OK;
  }
  sl.batch = b;
  // 3. pack + window kernels + result copies on the ctx stream, after the rows have landed
  cudaError_t ce = cudaStreamWaitEvent(ctx->stream, sl.h2d_done, 0);
  if (ce != cudaSuccess) {
    ctx->err = cudaGetErrorString(ce);
    rc = GB_ERR_CUDA;
  }
  if (!rc) rc = launch_pack(ctx, sl.panel, sl.d_stage, (int64_t)N, is_ascii, 0, n_t + n_u);
  for (int st = 0; st < 4 && !rc; st++) rc = run_stage(b, st);
  if (!rc) rc = gb::batch_fetch_enqueue(b, sl.h_z, sl.h_info, sl.h_status);
  if (!rc && (ce = cudaEventRecord(sl.done, ctx->stream)) != cudaSuccess) {
    ctx->err = cudaGetErrorString(ce);
    rc = GB_ERR_CUDA;
  }
  if (rc) {
    // something failed after kernels may have been enqueued: nothing may still touch the batch's buffers when they
    // are freed, and gb_pipe_wait must not wait on an event that was never recorded
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(pp->copy_stream);
    batch_free_device(b);
    delete b;
    sl.batch = nullptr;
    sl.early_status = rc;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int64_t i = 0; i < n_u; i++) z_u[i] = info_u[i] = nan;
    return GB_OK;   // reported by gb_pipe_wait
  }
  return GB_OK;
}

int gb_pipe_wait(gb_pipe* pp, int64_t ticket, int* window_status) {
  if (!pp || ticket < 0 || ticket >= pp->next_ticket) return GB_ERR_BAD_ARG;
  int rc = check_device(pp->ctx);
  if (rc) return rc;
  gb_pipe::Slot& sl = pp->slots[(size_t)(ticket % pp->depth)];
  if (sl.ticket != ticket) {  // already retired when its slot was reused
    pp->ctx->err = "ticket was already retired (its slot has been reused): wait within `depth` submissions";
    return GB_ERR_BAD_ARG;
  }
  return pipe_retire(pp, sl, window_status);
}

// ---- host-side mirror of run_dist / run_distmix ---------------------------------------------------
// The context's working panel for the per-window string entry points: reused while the population layout and the
// format stay the same, regrown (x 1.25) when a window needs more rows.  Freed with the context.
static int window_panel(gb_ctx* ctx, int n_pops, const int* pop_sizes, int64_t rows, int format, gb_panel** out) {
  gb_panel* w = static_cast<gb_panel*>(ctx->win_panel);
  if (w) {
    bool same = w->format == format && w->n_pops == n_pops && w->capacity >= rows;
    for (int k = 0; same && k < n_pops; k++) same = w->pop_sizes[(size_t)k] == pop_sizes[k];
    if (same) {
      gb_panel_clear(w);
      *out = w;
      return GB_OK;
    }
    ctx->win_panel = nullptr;
    gb_panel_destroy(w);
  }
  int rc = gb_panel_create_fmt(ctx, n_pops, pop_sizes, rows + rows / 4, format, &w);
  if (rc) return rc;
  ctx->win_panel = w;
  *out = w;
  return GB_OK;
}

int gb_run_window_strings(gb_ctx* ctx, int64_t n_snps, const int* type, const long long* bp, double* z,
                          double* info, const char* const* pop_strings, int n_pops, const int* pop_sizes,
                          const double* pop_wgt, long long start_bp, long long end_bp, const gb_params* params,
                          int* n_measured, int* n_unmeasured) {
  if (!ctx || n_snps < 0 || !type || !bp || !z || !info || !pop_strings || !pop_sizes) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  gb_params p;
  if (params) p = *params;
  else gb_params_default(&p);
  // dist.cpp:132-141 -- type 0 inside the prediction window -> unmeasured; type 1 anywhere in the
  // extended window -> measured; type 2 ignored
  std::vector<int64_t> meas, unme;
  for (int64_t i = 0; i < n_snps; i++) {
    if (type[i] == 0 && bp[i] >= start_bp && bp[i] <= end_bp) unme.push_back(i);
    else if (type[i] == 1) meas.push_back(i);
  }
  if (n_measured) *n_measured = (int)meas.size();
  if (n_unmeasured) *n_unmeasured = (int)unme.size();
  if ((int64_t)meas.size() <= p.min_num_measured_snp) return GB_ERR_TOO_FEW_MEASURED;      // dist.cpp:146
  if ((int64_t)unme.size() <= p.min_num_unmeasured_snp) return GB_ERR_TOO_FEW_UNMEASURED;  // dist.cpp:147
  const int64_t nt = (int64_t)meas.size(), nu = (int64_t)unme.size();
  std::vector<const char*> strs((size_t)(nt + nu) * n_pops);
  for (int64_t i = 0; i < nt + nu; i++) {
    const int64_t s = i < nt ? meas[(size_t)i] : unme[(size_t)(i - nt)];
    for (int k = 0; k < n_pops; k++) strs[(size_t)(i * n_pops + k)] = pop_strings[s * n_pops + k];
  }
  int rc = GB_OK;
  gb_panel* panel = nullptr;
  // Genotype strings of a real panel hold only '0','1','2' -> 4-bit operands.  The reference's
  // (c - '0') arithmetic accepts any byte; strings with other characters are repacked as int8,
  // which reproduces it for every 7-bit char.
  const bool trace = getenv("GB_STRINGS_TRACE") != nullptr;   // diagnostics: wall-clock phases of a call on stderr
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(now() - t0).count(); };
  for (int format = ctx->panel_format;; format = GB_PANEL_INT8) {
    // the working panel is kept in the context between calls (a genome is ~2,900 calls of the same shape): creating
    // and destroying 70 MB of device rows, statistics and tensor maps per window cost as much as the window itself
    auto t0 = now();
    rc = window_panel(ctx, n_pops, pop_sizes, nt + nu, format, &panel);
    if (rc) return rc;
    const double t_panel = ms_since(t0);
    t0 = now();
    rc = gb_panel_append_strings(panel, nt + nu, strs.data());
    const double t_append = ms_since(t0);
    double t_impute = 0.0;
    if (!rc) {
      t0 = now();
      std::vector<int64_t> rt((size_t)nt), ru((size_t)nu);
      std::vector<double> zt((size_t)nt), zu((size_t)nu), iu((size_t)nu);
      for (int64_t i = 0; i < nt; i++) rt[(size_t)i] = i, zt[(size_t)i] = z[meas[(size_t)i]];
      for (int64_t i = 0; i < nu; i++) ru[(size_t)i] = nt + i;
      rc = window_impute(ctx, panel, nt, rt.data(), nu, ru.data(), zt.data(), pop_wgt, &p, zu.data(), iu.data());
      if (rc == GB_OK || rc == GB_ERR_NOT_PD)
        for (int64_t i = 0; i < nu; i++) {  // SetZ / SetInfo, dist.cpp:200-202
          z[unme[(size_t)i]] = zu[(size_t)i];
          info[unme[(size_t)i]] = iu[(size_t)i];
        }
      t_impute = ms_since(t0);
    }
    if (trace)
      fprintf(stderr, "[strings trace] n_t %lld n_u %lld | panel %.2f ms | gather + copy + pack %.2f ms | plan + run + fetch %.2f ms\n",
              (long long)nt, (long long)nu, t_panel, t_append, t_impute);
    panel = nullptr;
    if (rc != GB_ERR_UNSUPPORTED || format == GB_PANEL_INT8) break;
  }
  return rc;
}

// ---- host-side mirror of run_qcat / run_qcatmix (qcat.cpp:134-262, qcatmix.cpp:145-286) -----------------
int gb_run_qcat_strings(gb_ctx* ctx, int64_t n_snps, const int* type, const long long* bp, const double* z,
                        const char* const* pop_strings, int n_pops, const int* pop_sizes, const double* pop_wgt,
                        long long start_bp, long long end_bp, const gb_params* params, double eig_cutoff, double* qcat_m,
                        double* qcat_t, double* qcat_chisq, int* n_measured, int* n_unmeasured) {
  if (!ctx || n_snps < 0 || !type || !bp || !z || !pop_strings || !pop_sizes || !qcat_m || !qcat_t || !qcat_chisq) {
    if (ctx) ctx->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  gb_params p;
  if (params) p = *params;
  else gb_params_default(&p);
  // qcat.cpp:139-152 -- type 0 inside the prediction window is tested as unmeasured; every type 1 SNP of the extended
  // window is measured, those before start_bp are the head wing, those inside the window are tested as measured
  std::vector<int64_t> meas, unme;
  int64_t headwing = 0, n_pred = 0;
  for (int64_t i = 0; i < n_snps; i++) {
    if (type[i] == 0 && bp[i] >= start_bp && bp[i] <= end_bp) {
      unme.push_back(i);
    } else if (type[i] == 1) {
      meas.push_back(i);
      if (bp[i] < start_bp) headwing++;
      else if (bp[i] <= end_bp) n_pred++;
    }
  }
  if (n_measured) *n_measured = (int)meas.size();
  if (n_unmeasured) *n_unmeasured = (int)unme.size();
  if ((int64_t)meas.size() <= p.min_num_measured_snp) return GB_ERR_TOO_FEW_MEASURED;                      // qcat.cpp:157
  if (pop_wgt && (int64_t)unme.size() <= p.min_num_unmeasured_snp) return GB_ERR_TOO_FEW_UNMEASURED;      // qcatmix.cpp:168
  const int64_t nt = (int64_t)meas.size(), nu = (int64_t)unme.size();
  std::vector<const char*> strs((size_t)(nt + nu) * n_pops);
  for (int64_t i = 0; i < nt + nu; i++) {
    const int64_t s = i < nt ? meas[(size_t)i] : unme[(size_t)(i - nt)];
    for (int k = 0; k < n_pops; k++) strs[(size_t)(i * n_pops + k)] = pop_strings[s * n_pops + k];
  }
  int rc = GB_OK;
  for (int format = ctx->panel_format;; format = GB_PANEL_INT8) {   // strings with other characters are repacked as int8
    gb_panel* panel = nullptr;
    rc = window_panel(ctx, n_pops, pop_sizes, nt + nu, format, &panel);   // the context's working panel, kept between calls
    if (rc) return rc;
    rc = gb_panel_append_strings(panel, nt + nu, strs.data());
    if (!rc) {
      std::vector<int64_t> rt((size_t)nt), ru((size_t)nu);
      std::vector<double> zt((size_t)nt), tm((size_t)n_pred), cm((size_t)n_pred), tu((size_t)nu), cu((size_t)nu);
      for (int64_t i = 0; i < nt; i++) rt[(size_t)i] = i, zt[(size_t)i] = z[meas[(size_t)i]];
      for (int64_t i = 0; i < nu; i++) ru[(size_t)i] = nt + i;
      int num_eig = 0;
      rc = gb_window_qcat(ctx, panel, nt, rt.data(), zt.data(), headwing, n_pred, nu, ru.data(), pop_wgt, &p, eig_cutoff,
                          &num_eig, tm.data(), cm.data(), tu.data(), cu.data());
      if (rc == GB_OK) {   // SetQcatM / SetQcatT / SetQcatChisq, qcat.cpp:230-232, 247-249
        for (int64_t i = 0; i < n_pred; i++) {
          const int64_t s = meas[(size_t)(headwing + i)];
          qcat_m[s] = num_eig, qcat_t[s] = tm[(size_t)i], qcat_chisq[s] = cm[(size_t)i];
        }
        for (int64_t i = 0; i < nu; i++) {
          const int64_t s = unme[(size_t)i];
          qcat_m[s] = num_eig, qcat_t[s] = tu[(size_t)i], qcat_chisq[s] = cu[(size_t)i];
        }
      }
    }
    if (rc != GB_ERR_UNSUPPORTED || format == GB_PANEL_INT8) break;
  }
  return rc;
}

}  // extern "C"

namespace gb {

int batch_fetch_finish_repair(gb_batch* b, const int* st, double* z_u, double* info_u, int* window_status_out) {
  std::vector<int> wst((size_t)b->n_windows, GB_OK);
  int rc = batch_fetch_finish(b, st, z_u, info_u, wst.data());
  if (rc) return rc;
  std::vector<int64_t> bad;
  for (int64_t w = 0; w < b->n_windows; w++)
    if (wst[(size_t)w] == GB_ERR_NOT_PD || wst[(size_t)w] == GB_ERR_BREAKDOWN) bad.push_back(w);
  if (!bad.empty() && !b->clip_mode && !b->ld_mode && !b->counts_mode && !b->d_y) {
    Ctx* ctx = b->ctx;
    // sub-batch of the failed windows on the same panel rows; every one of them takes the eigen-clip path
    std::vector<int64_t> to(1, 0), uo(1, 0), rt, ru;
    std::vector<double> zt;
    for (int64_t w : bad) {
      for (int64_t i = b->t_off[(size_t)w]; i < b->t_off[(size_t)w + 1]; i++) {
        rt.push_back(b->h_rows_t[(size_t)i]);
        zt.push_back(b->h_zt.empty() ? 0.0 : b->h_zt[(size_t)i]);
      }
      for (int64_t i = b->u_off[(size_t)w]; i < b->u_off[(size_t)w + 1]; i++) ru.push_back(b->h_rows_u[(size_t)i]);
      to.push_back((int64_t)rt.size());
      uo.push_back((int64_t)ru.size());
    }
    gb_params p = b->params;
    p.check_pd = 0;
    gb_batch* r = batch_new(static_cast<gb_ctx*>(ctx), static_cast<gb_panel*>(b->panel), (int64_t)bad.size(),
                            b->mode == GRAM_MIX ? b->h_wgt.data() : nullptr, &p, false, false, /*defer_flag_check=*/true, 1.0);
    if (!r) return GB_ERR_OOM;
    r->clip_mode = true;
    rc = batch_plan_host(r, to.data(), rt.data(), uo.data(), ru.data(), zt.data(), b->mode == GRAM_MIX ? b->h_wgt.data() : nullptr);
    if (!rc) rc = batch_plan_device(r, Arena{}, true);
    for (int stg = 0; stg < 4 && !rc; stg++) rc = run_stage(r, stg);
    std::vector<double> rz((size_t)uo.back()), ri((size_t)uo.back());
    std::vector<int> rst(2 * bad.size() + 3, 0), rw(bad.size(), GB_OK);
    if (!rc) rc = batch_fetch_enqueue(r, rz.data(), ri.data(), rst.data());
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GB_ERR_CUDA;
    if (!rc) rc = batch_fetch_finish(r, rst.data(), rz.data(), ri.data(), rw.data());
    if (!rc)
      for (size_t k = 0; k < bad.size(); k++) {
        const int64_t w = bad[k], n = b->u_off[(size_t)w + 1] - b->u_off[(size_t)w];
        if (z_u) std::memcpy(z_u + b->u_off[(size_t)w], rz.data() + uo[k], sizeof(double) * (size_t)n);
        if (info_u) std::memcpy(info_u + b->u_off[(size_t)w], ri.data() + uo[k], sizeof(double) * (size_t)n);
        wst[(size_t)w] = rw[k];
      }
    cudaStreamSynchronize(ctx->stream);
    batch_free_device(r);
    delete r;
    if (rc) return rc;
  }
  int worst = GB_OK;
  for (int64_t w = 0; w < b->n_windows; w++) {
    if (window_status_out) window_status_out[w] = wst[(size_t)w];
    if (wst[(size_t)w] != GB_OK && worst == GB_OK) worst = wst[(size_t)w];
  }
  return window_status_out ? GB_OK : worst;
}

}  // namespace gb
