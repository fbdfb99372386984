// Binary block dump of the factor and its streaming conversion to the reference's text format
// (SURVEY 8(f)-4).  write_matrix (mmat.rg:102-147) prints one "%d %d %0.8g" line per entry, which is the
// wire format verify.py reads but is ~25 bytes and one fprintf per nonzero: 128^3 has 3.4e9 entries.
// The dump keeps the dense filled clusters as they sit in HBM.
//
//   header  : char magic[8] = "CHOLFAC1"; int32 n, ncols; char typecode[4]; int32 reserved;
//             int64 nrecords; int64 nnz (entries != 0, what write_matrix counts, mmat.rg:114-127)
//   record  : int32 row0, col0, nrows, ncols (global permuted, 0-based); double v[nrows * ncols], column-major
//
// The converter never holds more than one record: pass 1 is the header's count, pass 2 prints.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/chol_mmio.h"
#include "../../include/cholesky.h"
#include "chol_internal.h"

namespace chb {

static const char kMagic[8] = {'C', 'H', 'O', 'L', 'F', 'A', 'C', '1'};
struct FactorFileHeader {
  char magic[8];
  int32_t n, ncols;
  char typecode[4];
  int32_t reserved;
  int64_t nrecords, nnz;
};
static_assert(sizeof(FactorFileHeader) == 40, "header layout");

int write_factor_binary(const Problem &P, const Symbolic &S, const double *fac, int rank, int world, int depth, const char *path,
                        std::string &err) {
  FILE *f = fopen(path, "wb");
  if (!f) return err = std::string("cannot write ") + path, -1;
  FactorFileHeader h;
  memcpy(h.magic, kMagic, 8);
  h.n = P.n, h.ncols = P.ncols, h.reserved = 0, h.nrecords = 0, h.nnz = 0;
  memcpy(h.typecode, P.typecode, 4);
  fwrite(&h, sizeof h, 1, f);
  std::vector<double> buf;
  for (int hc = 1; hc <= P.N; hc++) {
    if (world > 1) {  // a rank dumps its own subtree; rank 0 also the shared top panels
      int lv = P.level_of(hc), own = lv < depth ? -1 : (hc >> (lv - depth)) - (1 << depth);
      if (!(own == rank || (own < 0 && rank == 0))) continue;
    }
    const int nc = P.sz[hc], ld = S.ld[hc];
    if (nc == 0) continue;
    const double *pan = fac + S.poff[hc];
    for (int64_t s = S.seg_ptr[hc]; s < S.seg_ptr[hc + 1]; s++) {
      const Seg &sg = S.segs[s];
      const int nr = sg.hi - sg.lo;
      if (nr == 0) continue;
      int32_t rec[4] = {P.start[sg.anc] + sg.lo, P.start[hc], nr, nc};
      buf.resize((size_t)nr * nc);
      for (int col = 0; col < nc; col++)
        for (int r = 0; r < nr; r++) {
          const double v = pan[sg.off + r + (size_t)col * ld];
          buf[(size_t)col * nr + r] = v;
          h.nnz += (v != 0);
        }
      fwrite(rec, sizeof rec, 1, f);
      fwrite(buf.data(), sizeof(double), buf.size(), f);
      h.nrecords++;
    }
  }
  fseek(f, 0, SEEK_SET);
  fwrite(&h, sizeof h, 1, f);
  const bool bad = ferror(f) != 0;
  if (fclose(f) != 0 || bad) return err = std::string("write error on ") + path, -1;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// The symbolic structure as one binary file.  With one process per GPU every rank of a node needs the same
// Symbolic; one of them analyses (all host threads) and the others read the result instead of repeating the
// analysis side by side on a share of the cores (8 ranks: 7.6 s of analysis each).  A stamp of the problem
// (sizes, separator sizes, number of entries) guards against reading the analysis of a different problem.
namespace {
const char kSymMagic[8] = {'C', 'H', 'O', 'L', 'S', 'Y', 'M', '1'};
template <typename T>
bool put(FILE *f, const std::vector<T> &v) {
  const uint64_t n = v.size();
  return fwrite(&n, sizeof n, 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == n);
}
template <typename T>
bool get(FILE *f, std::vector<T> &v) {
  uint64_t n = 0;
  if (fread(&n, sizeof n, 1, f) != 1 || n > (1ULL << 36)) return false;
  v.resize(n);
  return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}
uint64_t problem_stamp(const Problem &P) {
  uint64_t h = mix64((uint64_t)P.n) ^ mix64((uint64_t)P.nz + 17) ^ mix64((uint64_t)P.levels + 31);
  for (int i = 1; i <= P.N; i++) h = mix64(h ^ (uint64_t)P.sz[i]);
  for (size_t e = 0; e < P.ei.size(); e += std::max<size_t>(1, P.ei.size() / 4096)) h = mix64(h ^ ((uint64_t)P.ei[e] << 32) ^ (uint64_t)P.ej[e]);
  return h;
}
}  // namespace

int save_symbolic(const Problem &P, const Symbolic &S, const char *path, std::string &err) {
  const std::string tmp = std::string(path) + ".tmp";
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return err = std::string("cannot write ") + tmp, -1;
  const uint64_t stamp = problem_stamp(P);
  std::vector<int> flat, ptr;  // cb[h][k][j], flattened
  ptr.push_back(0);
  std::vector<int> nk((size_t)P.N + 2, 0);
  for (int h = 1; h <= P.N; h++) {
    nk[h] = (int)S.cb[h].size();
    for (const auto &lst : S.cb[h]) {
      flat.insert(flat.end(), lst.begin(), lst.end());
      ptr.push_back((int)flat.size());
    }
  }
  int64_t scal[7] = {S.total_doubles, S.nblocks, S.nclusters0, S.calls[0], S.calls[1], S.calls[2], S.calls[3]};
  bool ok = fwrite(kSymMagic, 8, 1, f) == 1 && fwrite(&stamp, 8, 1, f) == 1 && fwrite(scal, sizeof scal, 1, f) == 1 && put(f, nk) && put(f, ptr) &&
            put(f, flat) && put(f, S.seg_ptr) && put(f, S.segs) && put(f, S.rows) && put(f, S.ld) && put(f, S.poff) && put(f, S.nfilled) &&
            put(f, S.checksum) && put(f, S.f_potrf) && put(f, S.f_trsm) && put(f, S.f_syrk) && put(f, S.f_gemm);
  const uint64_t nrec = S.records.size();
  ok = ok && fwrite(&nrec, 8, 1, f) == 1;
  for (const auto &r : S.records) ok = ok && put(f, r);
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) return err = std::string("write error on ") + path, -1;  // readers never see a partial file
  return 0;
}

int load_symbolic(const Problem &P, Symbolic &S, const char *path, std::string &err) {
  FILE *f = fopen(path, "rb");
  if (!f) return err = std::string("cannot read ") + path, -1;
  S = Symbolic();
  char magic[8];
  uint64_t stamp = 0, nrec = 0;
  int64_t scal[7];
  std::vector<int> flat, ptr, nk;
  bool ok = fread(magic, 8, 1, f) == 1 && !memcmp(magic, kSymMagic, 8) && fread(&stamp, 8, 1, f) == 1 && fread(scal, sizeof scal, 1, f) == 1 &&
            get(f, nk) && get(f, ptr) && get(f, flat) && get(f, S.seg_ptr) && get(f, S.segs) && get(f, S.rows) && get(f, S.ld) && get(f, S.poff) &&
            get(f, S.nfilled) && get(f, S.checksum) && get(f, S.f_potrf) && get(f, S.f_trsm) && get(f, S.f_syrk) && get(f, S.f_gemm) &&
            fread(&nrec, 8, 1, f) == 1 && nrec <= 64;
  if (ok) {
    S.records.resize(nrec);
    for (auto &r : S.records) ok = ok && get(f, r);
  }
  fclose(f);
  if (!ok) return err = std::string(path) + " is not a symbolic-analysis file (or is truncated)", -1;
  if (stamp != problem_stamp(P) || (int)nk.size() != P.N + 2 || (int)S.rows.size() != P.N + 2)
    return err = std::string(path) + " holds the analysis of a different problem", -1;
  S.total_doubles = scal[0], S.nblocks = scal[1], S.nclusters0 = scal[2];
  for (int i = 0; i < 4; i++) S.calls[i] = scal[3 + i];
  S.cb.assign((size_t)P.N + 2, {});
  size_t q = 0;
  for (int h = 1; h <= P.N; h++) {
    S.cb[h].resize(nk[h]);
    for (int k = 0; k < nk[h]; k++, q++) {
      if (q + 1 >= ptr.size() || ptr[q] > ptr[q + 1] || (size_t)ptr[q + 1] > flat.size()) return err = std::string(path) + " is corrupt", -1;
      S.cb[h][k].assign(flat.begin() + ptr[q], flat.begin() + ptr[q + 1]);
    }
  }
  return 0;
}

}  // namespace chb

extern "C" int chol_factor_binary_to_mtx(const char *bin_path, const char *mtx_path, int full_precision) {
  using namespace chb;
  FILE *in = fopen(bin_path, "rb");
  if (!in) return -1;
  FactorFileHeader h;
  if (fread(&h, sizeof h, 1, in) != 1 || memcmp(h.magic, kMagic, 8) != 0) return fclose(in), -2;
  FILE *out = fopen(mtx_path, "w");
  if (!out) return fclose(in), -3;
  MM_typecode tc;
  memcpy(tc, h.typecode, 4);
  mm_write_banner(out, tc);
  fprintf(out, "%d %d %lld\n", h.n, h.ncols, (long long)h.nnz);
  std::vector<double> buf;
  int64_t seen = 0;
  int rc = 0;
  for (int64_t k = 0; k < h.nrecords && !rc; k++) {
    int32_t rec[4];
    if (fread(rec, sizeof rec, 1, in) != 1 || rec[2] < 0 || rec[3] < 0 || rec[0] < 0 || rec[1] < 0 || (int64_t)rec[0] + rec[2] > h.n ||
        (int64_t)rec[1] + rec[3] > h.ncols) {
      rc = -4;
      break;
    }
    buf.resize((size_t)rec[2] * rec[3]);
    if (!buf.empty() && fread(buf.data(), sizeof(double), buf.size(), in) != buf.size()) {
      rc = -4;
      break;
    }
    // row-major inside the record, as write_matrix walks a block (mmat.rg:131-141)
    for (int r = 0; r < rec[2]; r++)
      for (int col = 0; col < rec[3]; col++) {
        const double v = buf[(size_t)col * rec[2] + r];
        if (v == 0) continue;
        fprintf(out, full_precision ? "%d %d %.17g\n" : "%d %d %0.8g\n", rec[0] + r + 1, rec[1] + col + 1, v);
        seen++;
      }
  }
  if (!rc && seen != h.nnz) rc = -5;
  fclose(in);
  if (fclose(out) != 0 && !rc) rc = -3;
  return rc;
}
