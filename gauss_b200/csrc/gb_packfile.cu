// gb_packfile.cu -- native converter of the reference's panel data file into a cached packed panel (".gbpack", SURVEY.md
// section 8f row 2).  Pure host code (it lives in the library so that every consumer of the C-ABI has it).
//
// The reference re-inflates and re-parses its BGZF text panel on every call: MakeSnpVecMix and ReadGenotype each
// bgzf_seek to a SNP's line and push ~33 KB of text through an istringstream (gauss.cpp:631-693, 720-785; one line per
// SNP: P genotype strings of '0'/'1'/'2' and P allele frequencies, gauss.cpp:572-585).  This converter does that work
// ONCE: BGZF blocks (bgzf.c:486-536: gzip members with a 'BC' extra field holding the block size) are inflated in
// parallel on host threads, lines are parsed in parallel, and every SNP becomes
//     * one ternary row ("pack5", 5 dosages per byte, gb_pack5_rows_host) over ALL populations,
//     * its P allele frequencies as doubles (the AF filter of MakeSnpVecMix needs them),
//     * the BGZF virtual offset of its line (the `fpos` column of the reference's index file, bgzf.h:108), so an index
//       entry finds its row.
// File layout (little endian): 64-byte header {magic "GBPACK5\n", version, n_rows, n_pops, row_bytes, off_sizes,
// off_rows, off_fpos, off_af1}, int32 sizes[n_pops], rows at a 4096-byte boundary, then fpos int64[n_rows] and
// af1 float64[n_rows][n_pops].
#include <zlib.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gb_batch.cuh"

using namespace gb;

namespace {

struct PackHeader {
  char magic[8];
  uint64_t version, n_rows, n_pops, row_bytes, off_sizes, off_rows, off_fpos, off_af1;
};
static_assert(sizeof(PackHeader) == 72, "header layout");

struct Block {
  int64_t coffset;            // file offset of the BGZF block
  std::vector<uint8_t> comp;  // raw deflate payload
  uint32_t isize;             // uncompressed size
  int64_t pos;                // position of the block's first byte in the batch text buffer
};

bool read_block(FILE* f, Block* b, std::string* err) {
  uint8_t head[12];
  b->coffset = ftello(f);
  const size_t got = fread(head, 1, 12, f);
  if (got == 0) return false;
  if (got < 12 || head[0] != 0x1f || head[1] != 0x8b || head[2] != 8 || !(head[3] & 4)) {
    *err = "not a BGZF block header at offset " + std::to_string(b->coffset);
    return false;
  }
  const int xlen = head[10] | (head[11] << 8);
  std::vector<uint8_t> extra((size_t)xlen);
  if (fread(extra.data(), 1, (size_t)xlen, f) != (size_t)xlen) {
    *err = "truncated BGZF extra field";
    return false;
  }
  int bsize = -1;
  for (int i = 0; i + 4 <= xlen;) {
    const int slen = extra[(size_t)i + 2] | (extra[(size_t)i + 3] << 8);
    if (extra[(size_t)i] == 'B' && extra[(size_t)i + 1] == 'C' && slen == 2) bsize = extra[(size_t)i + 4] | (extra[(size_t)i + 5] << 8);
    i += 4 + slen;
  }
  if (bsize < 0) {
    *err = "gzip member without the BGZF 'BC' field (plain gzip is not supported by the native converter)";
    return false;
  }
  const int payload = bsize + 1 - 12 - xlen - 8;
  if (payload < 0) {
    *err = "bad BGZF block size";
    return false;
  }
  b->comp.resize((size_t)payload);
  uint8_t tail[8];
  if (fread(b->comp.data(), 1, (size_t)payload, f) != (size_t)payload || fread(tail, 1, 8, f) != 8) {
    *err = "truncated BGZF block";
    return false;
  }
  b->isize = tail[4] | (tail[5] << 8) | (tail[6] << 16) | ((uint32_t)tail[7] << 24);
  return true;
}

bool inflate_block(const Block& b, uint8_t* dst) {
  if (b.isize == 0) return true;
  z_stream zs;
  std::memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -15) != Z_OK) return false;
  zs.next_in = const_cast<uint8_t*>(b.comp.data());
  zs.avail_in = (uInt)b.comp.size();
  zs.next_out = dst;
  zs.avail_out = b.isize;
  const int rc = inflate(&zs, Z_FINISH);
  inflateEnd(&zs);
  return rc == Z_STREAM_END && zs.total_out == b.isize;
}

template <class F>
void parallel_for(int n_threads, int64_t n, F&& fn) {
  if (n_threads <= 1 || n < 2) {
    for (int64_t i = 0; i < n; i++) fn(i);
    return;
  }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> th;
  auto drain = [&] {
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n) return;
      fn(i);
    }
  };
  for (int t = 0; t < n_threads; t++) {
    try {
      th.emplace_back(drain);
    } catch (...) {   // no thread to be had: fewer workers (nothing may be thrown across the C-ABI)
      break;
    }
  }
  if (th.empty()) drain();
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" int gb_packfile_convert(const char* geno_path, int n_pops, const int* pop_sizes, const char* out_path,
                                   int n_threads, int64_t* n_rows_out, double* text_bytes_out, double* seconds_out,
                                   char* err_out, int err_cap) {
  auto fail = [&](const std::string& msg, int code) {
    if (err_out && err_cap > 0) snprintf(err_out, (size_t)err_cap, "%s", msg.c_str());
    return code;
  };
  if (!geno_path || !out_path || n_pops < 1 || !pop_sizes) return fail("null argument", GB_ERR_BAD_ARG);
  std::vector<int> boff;
  const int row_bytes = pack5_layout(n_pops, pop_sizes, &boff);
  if (row_bytes < 0) return fail("bad population sizes", GB_ERR_BAD_ARG);
  if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  FILE* in = fopen(geno_path, "rb");
  if (!in) return fail(std::string("can't open reference data file '") + geno_path + "'", GB_ERR_BAD_ARG);
  FILE* out = fopen(out_path, "wb");
  if (!out) {
    fclose(in);
    return fail(std::string("can't create '") + out_path + "'", GB_ERR_BAD_ARG);
  }
  const auto t0 = std::chrono::steady_clock::now();
  PackHeader h{};
  std::memcpy(h.magic, "GBPACK5\n", 8);
  h.version = 1;
  h.n_pops = (uint64_t)n_pops;
  h.row_bytes = (uint64_t)row_bytes;
  h.off_sizes = sizeof(PackHeader);
  h.off_rows = (sizeof(PackHeader) + sizeof(int32_t) * (size_t)n_pops + 4095) / 4096 * 4096;
  {
    std::vector<uint8_t> pad((size_t)h.off_rows, 0);
    std::memcpy(pad.data() + h.off_sizes, pop_sizes, sizeof(int32_t) * (size_t)n_pops);
    fwrite(pad.data(), 1, pad.size(), out);
  }
  std::vector<int64_t> fpos_all;
  std::vector<double> af_all;
  std::string err;
  const int BATCH = 64 * n_threads;             // BGZF blocks per batch (<= 64 KB of text each)
  std::vector<uint8_t> text;                    // [carry of the previous batch | this batch]
  std::vector<Block> carry_blocks;              // blocks the carried bytes came from (for their virtual offsets)
  std::vector<Block> blocks;
  int64_t carry_len = 0;
  double text_bytes = 0;
  bool eof = false;
  std::atomic<int> bad{0};
  std::string bad_msg;
  std::vector<uint8_t> rows;
  while (!eof) {
    // ---- read a batch of blocks (sequential I/O), inflate them in parallel behind the carried bytes
    blocks.clear();
    int64_t pos = carry_len;
    while ((int)blocks.size() < BATCH) {
      Block b;
      if (!read_block(in, &b, &err)) {
        if (!err.empty()) {
          fclose(in);
          fclose(out);
          return fail(err, GB_ERR_BAD_ARG);
        }
        eof = true;
        break;
      }
      b.pos = pos;
      pos += b.isize;
      blocks.push_back(std::move(b));
    }
    text.resize((size_t)pos);
    parallel_for(n_threads, (int64_t)blocks.size(), [&](int64_t i) {
      if (!inflate_block(blocks[(size_t)i], text.data() + blocks[(size_t)i].pos)) bad.store(1);
    });
    if (bad.load()) {
      fclose(in);
      fclose(out);
      return fail("BGZF block failed to inflate", GB_ERR_BAD_ARG);
    }
    text_bytes += (double)(pos - carry_len);
    // ---- line starts; the bytes behind the last newline are carried into the next batch
    std::vector<int64_t> starts, ends;
    int64_t p = 0;
    const int64_t end = pos;
    while (p < end) {
      const void* nl = memchr(text.data() + p, '\n', (size_t)(end - p));
      if (!nl) break;
      starts.push_back(p);
      ends.push_back((const uint8_t*)nl - text.data());
      p = ends.back() + 1;
    }
    int64_t consumed = p;
    if (eof && p < end) {   // last line without a newline
      starts.push_back(p);
      ends.push_back(end);
      consumed = end;
    }
    const int64_t n = (int64_t)starts.size();
    // all blocks that cover this buffer, in position order
    std::vector<const Block*> cover;
    for (const Block& b : carry_blocks) cover.push_back(&b);
    for (const Block& b : blocks) cover.push_back(&b);
    const size_t row0 = fpos_all.size();
    fpos_all.resize(row0 + (size_t)n);
    af_all.resize((row0 + (size_t)n) * (size_t)n_pops);
    rows.assign((size_t)n * (size_t)row_bytes, 0);
    parallel_for(n_threads, n, [&](int64_t i) {
      const int64_t s = starts[(size_t)i];
      const int64_t e = ends[(size_t)i];
      // virtual offset of the line's first byte: the block holding it (a position on a block boundary belongs to the
      // NEXT block at offset 0, as bgzf_write's tell reports it) -- skip empty blocks
      size_t lo = 0, hi = cover.size();
      while (hi - lo > 1) {
        const size_t mid = (lo + hi) / 2;
        if (cover[mid]->pos <= s) lo = mid;
        else hi = mid;
      }
      while (lo + 1 < cover.size() && cover[lo]->isize == 0) lo++;
      fpos_all[row0 + (size_t)i] = (cover[lo]->coffset << 16) | (s - cover[lo]->pos);
      // parse: P genotype strings, then P allele frequencies (whitespace separated, gauss.cpp:660-674)
      const char* c = (const char*)text.data() + s;
      const char* ce = (const char*)text.data() + e;
      uint8_t* dst = rows.data() + (size_t)i * (size_t)row_bytes;
      for (int k = 0; k < n_pops; k++) {
        while (c < ce && (*c == ' ' || *c == '\t' || *c == '\r')) c++;
        const char* t = c;
        while (c < ce && !(*c == ' ' || *c == '\t' || *c == '\r')) c++;
        const int m = pop_sizes[k];
        if (c - t != m) {
          if (!bad.exchange(2)) bad_msg = "line " + std::to_string(row0 + (size_t)i + 1) + ": population " + std::to_string(k) + " has " +
                                          std::to_string(c - t) + " genotypes, expected " + std::to_string(m);
          return;
        }
        uint8_t* d = dst + boff[(size_t)k];
        int j = 0;
        for (; j + 5 <= m; j += 5) {
          const unsigned a = (uint8_t)(t[j] - '0'), b = (uint8_t)(t[j + 1] - '0'), cc = (uint8_t)(t[j + 2] - '0'),
                         dd = (uint8_t)(t[j + 3] - '0'), ee = (uint8_t)(t[j + 4] - '0');
          if (a > 2u || b > 2u || cc > 2u || dd > 2u || ee > 2u) bad.store(3);
          d[j / 5] = (uint8_t)(a + 3u * b + 9u * cc + 27u * dd + 81u * ee);
        }
        unsigned v = 0, mul = 1;
        for (int q = j; q < m; q++, mul *= 3) {
          const unsigned a = (uint8_t)(t[q] - '0');
          if (a > 2u) bad.store(3);
          v += mul * (a % 3u);
        }
        if (j < m) d[j / 5] = (uint8_t)v;
      }
      std::string tail(c, ce);   // strtod needs a terminated buffer
      const char* q = tail.c_str();
      for (int k = 0; k < n_pops; k++) {
        char* endp = nullptr;
        const double af = strtod(q, &endp);
        af_all[(row0 + (size_t)i) * (size_t)n_pops + (size_t)k] = endp == q ? 0.0 : af;   // a missing field reads as 0 (istringstream failure)
        q = endp;
      }
    });
    if (bad.load()) {
      fclose(in);
      fclose(out);
      return fail(bad.load() == 3 ? "a genotype character outside '0','1','2' (keep such a panel as text)" : bad_msg, GB_ERR_UNSUPPORTED);
    }
    if (n && fwrite(rows.data(), 1, rows.size(), out) != rows.size()) {
      fclose(in);
      fclose(out);
      return fail("write failed", GB_ERR_BAD_ARG);
    }
    // ---- carry: bytes behind the last complete line, and the blocks they came from
    carry_len = end - consumed;
    std::vector<Block> next_carry;
    for (const Block* b : cover)
      if (b->pos + (int64_t)b->isize > consumed) {
        Block nb;
        nb.coffset = b->coffset;
        nb.isize = b->isize;
        nb.pos = b->pos - consumed;   // may be negative: the block started before the carried bytes
        next_carry.push_back(std::move(nb));
      }
    if (carry_len > 0) std::memmove(text.data(), text.data() + consumed, (size_t)carry_len);
    carry_blocks.swap(next_carry);
    if (carry_len == 0) carry_blocks.clear();
  }
  fclose(in);
  h.n_rows = (uint64_t)fpos_all.size();
  h.off_fpos = h.off_rows + h.n_rows * h.row_bytes;
  h.off_af1 = h.off_fpos + sizeof(int64_t) * h.n_rows;
  fwrite(fpos_all.data(), sizeof(int64_t), fpos_all.size(), out);
  fwrite(af_all.data(), sizeof(double), af_all.size(), out);
  fseeko(out, 0, SEEK_SET);
  fwrite(&h, sizeof(h), 1, out);
  if (fclose(out) != 0) return fail("close failed", GB_ERR_BAD_ARG);
  if (n_rows_out) *n_rows_out = (int64_t)h.n_rows;
  if (text_bytes_out) *text_bytes_out = text_bytes;
  if (seconds_out) *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return GB_OK;
}
