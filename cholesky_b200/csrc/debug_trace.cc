// The reference's `-d` debug path, host side: the log verify.debug_factor (verify.py:216-275) replays.
//
// Grammar (one Python dict literal per line, parsed with eval by verify.py:26-29):
//   Block:   partition_matrix        mmat.rg:331, 352
//   Cluster: partition_separator     mmat.rg:396 ("Partitioning" header), 432
//   Fill:    compute_filled_clusters mmat.rg:1010
//   POTRF:   fused_dpotrf            blas.rg:308
//   TRSM:    fused_dtrsm             blas.rg:340
//   GEMM:    fused_dsyrk/fused_dgemm blas.rg:405, 422, 490
// Everything is derived from the symbolic structure (block bounds, composed cluster boundaries and
// the `Filled` records per interval label), so the log needs no GPU; the numeric snapshots that go
// with it (write_blocks, mmat.rg:174-218) come from chol_factor_debug in engine.cu.
// Records of one block are visited in ascending cluster index (the reference walks the Legion
// region in allocation order; the order inside one fused task does not change any result).
#include <algorithm>
#include <map>

#include "chol_internal.h"

namespace chb {

std::vector<DebugStep> debug_steps(const Problem &P) {
  std::vector<DebugStep> out;
  char name[256];
  for (int lvl = P.levels - 1; lvl >= 0; lvl--) {
    const int first = 1 << lvl, last = (1 << (lvl + 1)) - 1;
    for (int hs = first; hs <= last; hs++) {  // mmat.rg:1240-1257
      snprintf(name, sizeof name, "potrf_lvl%d_a%d%d", lvl, P.label_of(hs), P.label_of(hs));
      out.push_back(DebugStep{0, lvl, hs, 0, 0, name});
    }
    for (int hs = first; hs <= last; hs++)  // mmat.rg:1259-1290
      for (int hp = hs >> 1; hp >= 1; hp >>= 1) {
        snprintf(name, sizeof name, "trsm_lvl%d_a%d%d_b%d%d", lvl, P.label_of(hs), P.label_of(hs), P.label_of(hp), P.label_of(hs));
        out.push_back(DebugStep{1, lvl, hs, hp, 0, name});
      }
    for (int hs = first; hs <= last; hs++)  // mmat.rg:1292-1346
      for (int hp = hs >> 1; hp >= 1; hp >>= 1)
        for (int hg = hp; hg >= 1; hg >>= 1) {
          snprintf(name, sizeof name, "gemm_lvl%d_a%d%d_b%d%d_c%d%d", lvl, P.label_of(hg), P.label_of(hs), P.label_of(hp), P.label_of(hs),
                   P.label_of(hg), P.label_of(hp));
          out.push_back(DebugStep{2, lvl, hs, hp, hg, name});
        }
  }
  return out;
}

namespace {

struct LogWriter {
  const Problem &P;
  const Symbolic &S;
  FILE *f;
  int t = 0, lvl = 0;
  // records of interval label t by block (row label, col label)
  std::map<std::pair<int64_t, int64_t>, std::pair<size_t, size_t>> range;
  void index(int tt) {
    t = tt, lvl = P.levels - 1 - tt;
    range.clear();
    const auto &r = S.records[t];
    for (size_t i = 0; i < r.size();) {
      size_t j = i;
      while (j < r.size() && r[j].sep_x == r[i].sep_x && r[j].sep_y == r[i].sep_y) j++;
      range[{r[i].sep_x, r[i].sep_y}] = {i, j};
      i = j;
    }
  }
  std::pair<size_t, size_t> block(int hr, int hc) const {
    auto it = range.find({P.label_of(hr), P.label_of(hc)});
    return it == range.end() ? std::pair<size_t, size_t>{0, 0} : it->second;
  }
  void operand(const char *name, const char *size_key, const FilledRec &r) {
    fprintf(f, "'%s': (%lld, %lld, %lld), '%s_Lo': (%lld, %lld), '%s_Hi': (%lld, %lld), '%s%s': (%lld, %lld), ", name, (long long)r.sep_x,
            (long long)r.sep_y, (long long)r.cluster, name, (long long)r.lo_x, (long long)r.lo_y, name, (long long)r.hi_x, (long long)r.hi_y,
            size_key, name, (long long)(r.hi_x - r.lo_x + 1), (long long)(r.hi_y - r.lo_y + 1));
  }
  void tail(const FilledRec &blk) {
    fprintf(f, "'Block': (%lld, %lld), 'Level': %d, 'Interval': %d}\n", (long long)blk.sep_x, (long long)blk.sep_y, lvl, t);
  }
  void clusters_of(int hr, int hc, int k) {
    const int rows = k < (int)S.cb[hr].size() ? (int)S.cb[hr][k].size() - 1 : -1;
    const int cols = k < (int)S.cb[hc].size() ? (int)S.cb[hc][k].size() - 1 : -1;
    fprintf(f, "\t\tPartitioning (%d, %d) Cluster: %d Rows: %d Cols: %d\n", P.label_of(hr), P.label_of(hc), k, rows, cols);
    for (int row = 0; row < rows; row++) {
      for (int col = 0; col < cols; col++) {
        const long long lox = P.start[hr] + S.cb[hr][k][row], hix = P.start[hr] + S.cb[hr][k][row + 1] - 1;
        const long long loy = P.start[hc] + S.cb[hc][k][col], hiy = P.start[hc] + S.cb[hc][k][col + 1] - 1;
        const long long sx = hix - lox + 1, sy = hiy - loy + 1;
        fprintf(f,
                "\t\tCluster: {'Block': (%d, %d), 'color': (%d, %d, %d), 'Lo': (%lld, %lld), 'Hi': (%lld, %lld), 'size': (%lld, %lld), "
                "'vol': %lld, 'Interval': %d}\n",
                P.label_of(hr), P.label_of(hc), P.label_of(hr), P.label_of(hc), row * cols + col, lox, loy, hix, hiy, sx, sy,
                (sx > 0 && sy > 0) ? sx * sy : 0LL, t);
      }
      fprintf(f, "\n");
    }
  }
};

}  // namespace

int write_debug_log(const Problem &P, const Symbolic &S, FILE *f, std::string &err) {
  const int L = P.levels;
  if ((int)S.records.size() != L) return err = "the debug log needs the Filled records: analyze with keep_records", -1;
  LogWriter W{P, S, f};
  // partition_matrix (mmat.rg:316-360): root first, every separator followed by its ancestor blocks
  for (int lvl = 0; lvl < L; lvl++)
    for (int h = 1 << lvl; h < (1 << (lvl + 1)); h++)
      for (int hr = h; hr >= 1; hr >>= 1)
        fprintf(f, "Block: {'Block': (%d, %d), 'Lo': (%d, %d), 'Hi': (%d, %d)}\n", P.label_of(hr), P.label_of(h), P.start[hr], P.start[h],
                P.start[hr] + P.sz[hr] - 1, P.start[h] + P.sz[h] - 1);
  // compute_filled_clusters (mmat.rg:918-1026)
  for (int t = 0; t < L; t++) {
    W.index(t);
    const int lvl = L - 1 - t, k = std::max(0, L - 2 - lvl);  // cluster partition index, mmat.rg:1018-1022
    for (int l2 = 0; l2 <= lvl; l2++)
      for (int hr = 1 << l2; hr < (1 << (l2 + 1)); hr++)
        for (int cl = l2; cl <= lvl; cl++)
          for (int hc = hr << (cl - l2); hc < ((hr + 1) << (cl - l2)); hc++) W.clusters_of(hr, hc, k);
    for (const FilledRec &r : S.records[t])
      fprintf(f,
              "Fill: {'Level': %d, 'Interval': %d, 'Block': (%lld, %lld), 'Cluster': (%lld, %lld, %lld), 'Filled': 0, 'Lo': (%lld, %lld), "
              "'Hi': (%lld, %lld), 'Size': (%lld, %lld)}\n",
              lvl, t, (long long)r.sep_x, (long long)r.sep_y, (long long)r.sep_x, (long long)r.sep_y, (long long)r.cluster, (long long)r.lo_x,
              (long long)r.lo_y, (long long)r.hi_x, (long long)r.hi_y, (long long)(r.hi_x - r.lo_x + 1), (long long)(r.hi_y - r.lo_y + 1));
  }
  // the level loop (mmat.rg:1227-1355)
  for (int lvl = L - 1; lvl >= 0; lvl--) {
    const int t = L - 1 - lvl, k = std::max(0, L - 2 - lvl);
    W.index(t);
    const auto &R = S.records[t];
    fprintf(f, "Factoring Level: %d Interval: %d Iteration: %d\n", lvl, k, 0);
    const int first = 1 << lvl, last = (1 << (lvl + 1)) - 1;
    for (int hs = first; hs <= last; hs++) {
      auto a = W.block(hs, hs);
      for (size_t i = a.first; i < a.second; i++) {
        fprintf(f, "POTRF: {");
        W.operand("A", "Size", R[i]);
        W.tail(R[i]);
      }
    }
    for (int hs = first; hs <= last; hs++)
      for (int hp = hs >> 1; hp >= 1; hp >>= 1) {
        auto a = W.block(hs, hs), b = W.block(hp, hs);
        for (size_t i = a.first; i < a.second; i++)
          for (size_t j = b.first; j < b.second; j++) {
            fprintf(f, "TRSM: {");
            W.operand("A", "Size", R[i]);
            W.operand("B", "Size", R[j]);
            W.tail(R[j]);
          }
      }
    for (int hs = first; hs <= last; hs++)
      for (int hp = hs >> 1; hp >= 1; hp >>= 1)
        for (int hg = hp; hg >= 1; hg >>= 1) {
          auto a = W.block(hg, hs), b = W.block(hp, hs), cc = W.block(hg, hp);
          const int64_t col_cluster_size = (int64_t)S.cb[hp][k].size() - 1;  // mmat.rg:1315
          for (size_t i = a.first; i < a.second; i++)
            for (size_t j = b.first; j < b.second; j++) {
              const int64_t row = R[i].cluster, col = R[j].cluster, z = row * col_cluster_size + col;
              // destination lookup (blas.rg:385-392, 466-473): absent or empty => the pair is skipped
              size_t c = cc.first;
              while (c < cc.second && R[c].cluster != z) c++;
              if (c == cc.second) continue;
              if (hg == hp && col > row) continue;  // fused_dsyrk touches the lower cluster triangle only
              fprintf(f, "GEMM: {");
              W.operand("A", "size", R[i]);
              W.operand("B", "size", R[j]);
              W.operand("C", "size", R[c]);
              W.tail(R[c]);
            }
        }
  }
  fprintf(f, "Done factoring Iteration: %d.\n", 0);
  return ferror(f) ? (err = "write error on the debug log", -1) : 0;
}

}  // namespace chb
