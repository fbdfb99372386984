// Schedule compiler: turns the symbolic structure into the launch list the GPU executes.
//
// Per tree level, leaves first (mmat.rg:1227), three phases that mirror the reference's three
// __demand(__parallel) loops (mmat.rg:1240, 1259, 1293):
//   (a) fused_dpotrf  -> blocked Cholesky of every pivot block of the level, in lock step:
//                        left-looking over NBO-wide block columns (one grouped GEMM with K = all
//                        columns to the left), right-looking over NB-wide tiles inside one
//                        (tile POTRF, slab TRSM, small trailing GEMM).
//   (b) fused_dtrsm   -> the same blocking applied to the filled off-diagonal rows of the panels.
//   (c) fused_dsyrk / fused_dgemm -> one grouped GEMM over DESTINATION clusters: every filled
//                        cluster (g, p, ia, jb) owns the ordered list of its contributors
//                        (s ascending), so accumulation is atomic-free and deterministic; the
//                        extend-add index map is the precomputed destination offset.
#include <algorithm>
#include <cstring>

#include "chol_internal.h"

namespace chb {

namespace {

struct Builder {
  const Problem &P;
  const Symbolic &S;
  Schedule &D;
  Builder(const Problem &p, const Symbolic &s, Schedule &d) : P(p), S(s), D(d) {}

  // tiles of the launch being built, per tile configuration (0: 64x64, 1: 128x128)
  std::vector<TileRef> cur[2];
  double cur_flops[2] = {0, 0};
  void begin_gemm() {
    cur[0].clear(), cur[1].clear();
    cur_flops[0] = cur_flops[1] = 0;
  }
  int cfg_of(const GemmProblem &g) const { return (g.M >= D.big_m && g.N >= D.big_n) ? 1 : 0; }
  // one problem with a single contributor (in-panel updates)
  void add_problem(int64_t c_off, int ldc, int M, int N, int tri, int64_t a_off, int64_t b_off, int lda, int ldb, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return;
    GemmProblem g;
    g.c_off = c_off, g.ldc = ldc, g.M = M, g.N = N, g.tri = tri;
    g.contrib_begin = (int)D.contribs.size(), g.contrib_count = 1;
    D.contribs.push_back(GemmContrib{a_off, b_off, lda, ldb, K, 0});
    D.probs.push_back(g);
    add_tiles((int)D.probs.size() - 1, (tri ? 1.0 : 2.0) * M * N * K);
  }
  void add_tiles(int prob, double flops) {
    const GemmProblem &g = D.probs[prob];
    const int cfg = cfg_of(g), bm = cfg ? 128 : 64, bn = bm;
    int tr_n = (g.M + bm - 1) / bm, tc_n = (g.N + bn - 1) / bn;
    for (int tc = 0; tc < tc_n; tc++)
      for (int tr = 0; tr < tr_n; tr++) {
        if (g.tri && (tr + 1) * bm - 1 < tc * bn) continue;  // wholly above the diagonal
        cur[cfg].push_back(TileRef{prob, (uint16_t)tr, (uint16_t)tc});
      }
    cur_flops[cfg] += flops;
  }
  void end_gemm(int level, int phase) {
    for (int cfg = 1; cfg >= 0; cfg--) {
      if (cur[cfg].empty()) continue;
      int64_t b = (int64_t)D.tiles.size();
      D.tiles.insert(D.tiles.end(), cur[cfg].begin(), cur[cfg].end());
      D.launches.push_back(Launch{K_GEMM, level, phase, b, (int64_t)cur[cfg].size(), cur_flops[cfg], cfg});
    }
  }
};

struct Pair {  // one (A cluster, B cluster) contribution of separator hs
  int p;       // destination panel
  int crow;    // destination stored row in that panel
  int ccol;    // destination column in that panel
  int hs;
  int M, N;
  int tri;
  int64_t a_off, b_off;
  int ld, K;
};

}  // namespace

int build_schedule(const Problem &P, const Symbolic &S, Schedule &D, std::string &err) {
  D = Schedule();
  const int L = P.levels, N = P.N;
  const int NB = D.nb, NBO = D.nbo, SLAB = D.slab;
  Builder B(P, S, D);

  // global permuted row of every segment start, per panel, for destination lookups
  std::vector<int> seg_grow(S.segs.size());
  for (int h = 1; h <= N; h++)
    for (int64_t i = S.seg_ptr[h]; i < S.seg_ptr[h + 1]; i++) seg_grow[i] = P.start[S.segs[i].anc] + S.segs[i].lo;
  auto locate = [&](int p, int grow) -> int {  // stored row of global row `grow` in panel p, -1 if absent
    int64_t lo = S.seg_ptr[p], hi = S.seg_ptr[p + 1];
    int64_t it = std::upper_bound(seg_grow.begin() + lo, seg_grow.begin() + hi, grow) - seg_grow.begin() - 1;
    if (it < lo) return -1;
    const Seg &s = S.segs[it];
    int local = grow - seg_grow[it];
    if (local >= s.hi - s.lo) return -1;
    return s.off + local;
  };

  // ---- assembly map (fill_block, mmat.rg:529-633, as a scatter)
  {
    std::vector<int> iperm(P.n), rowheap(P.n);
    for (int p = 0; p < P.n; p++) iperm[P.perm[p]] = p;
    for (int h = 1; h <= N; h++)
      for (int i = 0; i < P.sz[h]; i++) rowheap[P.start[h] + i] = h;
    D.a_off.assign((size_t)P.nz, -1);
    for (int64_t e = 0; e < P.nz; e++) {
      if (P.ev[e] == 0.0) continue;
      int pi = iperm[P.ei[e]], pj = iperm[P.ej[e]];
      if (pi < pj) std::swap(pi, pj);
      int hr = rowheap[pi], hc = rowheap[pj];
      int d = P.level_of(hc) - P.level_of(hr);
      if (d < 0 || (hc >> d) != hr) continue;
      int r = locate(hc, pi);
      if (r < 0) return err = "internal: nonzero outside the filled pattern", -1;
      D.a_off[e] = S.poff[hc] + r + (int64_t)(pj - P.start[hc]) * S.ld[hc];
    }
  }

  std::vector<Pair> pairs;
  for (int lvl = L - 1; lvl >= 0; lvl--) {
    const int h0 = 1 << lvl, h1 = 1 << (lvl + 1);
    int maxn = 0;
    for (int h = h0; h < h1; h++) maxn = std::max(maxn, P.sz[h]);
    const int nouter = (maxn + NBO - 1) / NBO;

    // which == 0: pivot blocks (rows [0, n));  which == 1: off-diagonal rows [r0, R)
    for (int which = 0; which < 2; which++) {
      const int phase = which == 0 ? PH_POTRF : PH_TRSM;
      for (int J = 0; J < nouter; J++) {
        const int c0 = J * NBO;
        // left-looking update of block column J with everything to its left
        if (J > 0) {
          B.begin_gemm();
          for (int h = h0; h < h1; h++) {
            int n = P.sz[h], ld = S.ld[h];
            if (n <= c0) continue;
            int cw = std::min(NBO, n - c0);
            int64_t base = S.poff[h];
            if (which == 0) B.add_problem(base + c0 + (int64_t)c0 * ld, ld, n - c0, cw, 1, base + c0, base + c0, ld, ld, c0);
            else {
              int r0 = (n + 1) / 2 * 2, m = S.rows[h] - r0;
              B.add_problem(base + r0 + (int64_t)c0 * ld, ld, m, cw, 0, base + r0, base + c0, ld, ld, c0);
            }
          }
          B.end_gemm(lvl, phase);
        }
        for (int jj = 0; jj < NBO / NB; jj++) {
          const int d0 = c0 + jj * NB;
          if (d0 >= maxn) break;
          if (which == 0) {
            int64_t b = (int64_t)D.potrf.size();
            for (int h = h0; h < h1; h++) {
              int n = P.sz[h];
              if (n <= d0) continue;
              D.potrf.push_back(PotrfDesc{S.poff[h] + d0 + (int64_t)d0 * S.ld[h], S.ld[h], std::min(NB, n - d0), P.start[h] + d0, 0});
            }
            if ((int64_t)D.potrf.size() > b) D.launches.push_back(Launch{K_POTRF, lvl, phase, b, (int64_t)D.potrf.size() - b, 0, 0});
          }
          {
            int64_t b = (int64_t)D.trsm_tiles.size();
            for (int h = h0; h < h1; h++) {
              int n = P.sz[h], ld = S.ld[h];
              if (n <= d0) continue;
              int dw = std::min(NB, n - d0);
              int rbeg = which == 0 ? d0 + dw : (n + 1) / 2 * 2;
              int rend = which == 0 ? n : S.rows[h];
              if (rend <= rbeg) continue;
              D.trsm.push_back(TrsmDesc{S.poff[h] + d0 + (int64_t)d0 * ld, S.poff[h] + rbeg + (int64_t)d0 * ld, ld, dw, rend - rbeg, 0});
              int ns = (rend - rbeg + SLAB - 1) / SLAB;
              for (int s = 0; s < ns; s++) D.trsm_tiles.push_back(TileRef{(int)D.trsm.size() - 1, (uint16_t)(s & 0xffff), (uint16_t)(s >> 16)});
            }
            if ((int64_t)D.trsm_tiles.size() > b) D.launches.push_back(Launch{K_TRSM, lvl, phase, b, (int64_t)D.trsm_tiles.size() - b, 0, 0});
          }
          // right-looking update of the rest of this outer block column
          B.begin_gemm();
          for (int h = h0; h < h1; h++) {
            int n = P.sz[h], ld = S.ld[h];
            if (n <= d0) continue;
            int dw = std::min(NB, n - d0), e0 = d0 + dw, cend = std::min(c0 + NBO, n);
            if (e0 >= cend) continue;
            int64_t base = S.poff[h];
            if (which == 0)
              B.add_problem(base + e0 + (int64_t)e0 * ld, ld, n - e0, cend - e0, 1, base + e0 + (int64_t)d0 * ld, base + e0 + (int64_t)d0 * ld, ld, ld, dw);
            else {
              int r0 = (n + 1) / 2 * 2, m = S.rows[h] - r0;
              B.add_problem(base + r0 + (int64_t)e0 * ld, ld, m, cend - e0, 0, base + r0 + (int64_t)d0 * ld, base + e0 + (int64_t)d0 * ld, ld, ld, dw);
            }
          }
          B.end_gemm(lvl, phase);
        }
      }
    }

    // ---- (c) Schur updates of the level, grouped by destination cluster
    pairs.clear();
    for (int hs = h0; hs < h1; hs++) {
      int64_t s0 = S.seg_ptr[hs] + 1, s1 = S.seg_ptr[hs + 1];  // off-diagonal segments
      int ld = S.ld[hs], K = P.sz[hs];
      int64_t base = S.poff[hs];
      for (int64_t j = s0; j < s1; j++) {
        const Seg &b = S.segs[j];
        int p = b.anc;
        for (int64_t i = j; i < s1; i++) {
          const Seg &a = S.segs[i];
          int crow = locate(p, P.start[a.anc] + a.lo);
          if (crow < 0) return err = "internal: update destination outside the filled pattern", -1;
          Pair q;
          q.p = p, q.crow = crow, q.ccol = b.lo, q.hs = hs;
          q.M = a.hi - a.lo, q.N = b.hi - b.lo, q.tri = (i == j);
          q.a_off = base + a.off, q.b_off = base + b.off, q.ld = ld, q.K = K;
          pairs.push_back(q);
        }
      }
    }
    std::sort(pairs.begin(), pairs.end(), [](const Pair &x, const Pair &y) {
      if (x.p != y.p) return x.p < y.p;
      if (x.ccol != y.ccol) return x.ccol < y.ccol;
      if (x.crow != y.crow) return x.crow < y.crow;
      return x.hs < y.hs;
    });
    B.begin_gemm();
    for (size_t i = 0; i < pairs.size();) {
      size_t j = i;
      while (j < pairs.size() && pairs[j].p == pairs[i].p && pairs[j].ccol == pairs[i].ccol && pairs[j].crow == pairs[i].crow) j++;
      const Pair &q = pairs[i];
      GemmProblem g;
      g.c_off = S.poff[q.p] + q.crow + (int64_t)q.ccol * S.ld[q.p];
      g.ldc = S.ld[q.p], g.M = q.M, g.N = q.N, g.tri = q.tri;
      g.contrib_begin = (int)D.contribs.size(), g.contrib_count = (int)(j - i);
      double pf = 0;
      for (size_t c = i; c < j; c++) {
        if (pairs[c].M != q.M || pairs[c].N != q.N || pairs[c].tri != q.tri) return err = "internal: contributors of one destination cluster disagree on its shape", -1;
        D.contribs.push_back(GemmContrib{pairs[c].a_off, pairs[c].b_off, pairs[c].ld, pairs[c].ld, pairs[c].K, 0});
        pf += (q.tri ? 1.0 : 2.0) * q.M * q.N * pairs[c].K;
      }
      D.probs.push_back(g);
      B.add_tiles((int)D.probs.size() - 1, pf);
      i = j;
    }
    B.end_gemm(lvl, PH_UPDATE);
  }
  if (D.contribs.size() > 0x7fffffffULL || D.probs.size() > 0x7fffffffULL) return err = "schedule too large", -1;
  return 0;
}

}  // namespace chb
