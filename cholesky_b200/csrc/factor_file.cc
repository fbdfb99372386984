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
