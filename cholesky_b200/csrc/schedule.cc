// Schedule compiler: turns the symbolic structure into the launch list the GPU executes.
//
// Per tree level, leaves first (mmat.rg:1227), three phases that mirror the reference's three
// __demand(__parallel) loops (mmat.rg:1240, 1259, 1293):
//   (a) fused_dpotrf  -> blocked Cholesky of every pivot block of the level, in lock step:
//                        NBO-wide block columns; inside one, NB-wide tiles (tile POTRF, slab TRSM,
//                        small trailing GEMM); after one, a right-looking grouped GEMM (K = NBO) over
//                        the whole trailing panel, which keeps the GPU full even for one front.
//   (b) fused_dtrsm   -> the same blocking applied to the filled off-diagonal rows of the panels
//                        (by default fused with (a): every row below a pivot tile advances together).
//   (c) fused_dsyrk / fused_dgemm -> one grouped GEMM over DESTINATION clusters: every filled
//                        cluster (g, p, ia, jb) owns the ordered list of its contributors
//                        (s ascending), so accumulation is atomic-free and deterministic; the
//                        extend-add index map is the precomputed destination offset.
//
// Look-ahead: the latency-bound chain of a block column (tile POTRF, slab TRSM, K = NB GEMM) runs on
// a second stream.  The trailing update of block column J is split into the tiles of block column
// J + 1 (part A) and the rest (part B); the chain of J + 1 only waits for A, so it overlaps B.
//
// Multi-GPU (world = 2^d ranks, one process per GPU): rank r owns the subtree under heap index
// 2^d + r and schedules only its separators on levels >= d.  Its Schur contributions to the top
// d levels land in its own copy of the top panels; one K_ALLREDUCE sums the copies over NVLink.
// On the top levels every rank runs the chain redundantly (bit-identical).  Trailing updates use a
// static ownership of tile rows (row tile index mod world): part A is stored into every rank's
// copy (SHARED kernel) and followed by a K_BARRIER, part B stays local until its block column's turn.
// Schur updates of top levels are split by contiguous tile slices and stored into every copy.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "chol_internal.h"

namespace chb {

namespace {

struct Builder {
  const Problem &P;
  const Symbolic &S;
  Schedule &D;
  Builder(const Problem &p, const Symbolic &s, Schedule &d) : P(p), S(s), D(d) {}

  // ---- launches, streams and events
  int last[2] = {-1, -1};   // index of the last launch pushed on each stream
  int pendw[2] = {-1, -1};  // event the next launch of the stream has to wait for
  static Launch mk(int kind, int level, int phase, int64_t begin, int64_t count, double flops, int cfg, int shared) {
    Launch l;
    l.kind = kind, l.level = level, l.phase = phase, l.begin = begin, l.count = count, l.flops = flops, l.cfg = cfg, l.shared = shared;
    l.stream = 0, l.wait_ev = -1, l.rec_ev = -1;
    return l;
  }
  void push(Launch l, int stream) {
    if (!D.lookahead) stream = 0;
    l.stream = stream;
    l.wait_ev = pendw[stream];
    pendw[stream] = -1;
    D.launches.push_back(l);
    last[stream] = (int)D.launches.size() - 1;
  }
  // everything pushed so far on stream `src` happens before whatever is pushed next on stream `dst`
  void depend(int dst, int src) {
    if (!D.lookahead || dst == src || last[src] < 0) return;
    if (D.launches[last[src]].rec_ev < 0) D.launches[last[src]].rec_ev = D.num_events++;
    const int ev = D.launches[last[src]].rec_ev;
    if (pendw[dst] >= 0 && pendw[dst] != ev) push(mk(K_NOP, D.launches[last[src]].level, 0, 0, 0, 0, 0, 0), dst);
    pendw[dst] = ev;
  }

  // ---- grouped GEMM launches
  struct Pending {
    int prob;
    double flops;
    // trailing update of a panel: tile rows are owned by rank (row_tile0 + tr) % world for the whole
    // panel factorization; tile columns < bcast_tc belong to the next block column (part A)
    int row_tile0 = -1, bcast_tc = 0;
  };
  std::vector<Pending> pend;
  int mode = 0;  // 0: Schur update; 1: trailing update of a block column (parts A / B); 2: chain GEMM (K = NB)
  void begin_gemm(int m) {
    pend.clear();
    mode = m;
  }
  bool is_big(const GemmProblem &g) const { return g.M >= D.big_m && g.N >= D.big_n; }
  static int cfg_bm(int cfg) { return cfg == 3 ? 32 : cfg == 0 ? 64 : 128; }
  static int cfg_bn(int cfg) { return cfg == 3 ? 32 : cfg == 1 ? 128 : 64; }
  static int64_t ntiles(const GemmProblem &g, int cfg) {
    const int bm = cfg_bm(cfg), bn = cfg_bn(cfg);
    int64_t tr_n = (g.M + bm - 1) / bm, tc_n = (g.N + bn - 1) / bn;
    if (!g.tri) return tr_n * tc_n;
    int64_t t = 0;
    for (int64_t tc = 0; tc < tc_n; tc++)
      for (int64_t tr = 0; tr < tr_n; tr++) t += !((tr + 1) * bm - 1 < tc * bn);
    return t;
  }
  // one problem with a single contributor (in-panel updates)
  void add_problem(int64_t c_off, int ldc, int M, int N, int tri, int64_t a_off, int64_t b_off, int lda, int ldb, int K,
                   int row_tile0 = -1, int bcast_tc = 0) {
    if (M <= 0 || N <= 0 || K <= 0) return;
    GemmProblem g;
    g.c_off = c_off, g.ldc = ldc, g.M = M, g.N = N, g.tri = tri;
    g.contrib_begin = (int)D.contribs.size(), g.contrib_count = 1;
    D.contribs.push_back(GemmContrib{a_off, b_off, lda, ldb, K, 0});
    D.probs.push_back(g);
    Pending pd;
    pd.prob = (int)D.probs.size() - 1;
    // executed flops: the strict upper triangle of the leading N x N part is skipped when tri
    pd.flops = 2.0 * K * ((double)M * N - (tri ? 0.5 * N * (N - 1.0) : 0.0));
    pd.row_tile0 = row_tile0, pd.bcast_tc = bcast_tc;
    pend.push_back(pd);
  }
  void push_gemm(int level, int phase, int cfg, int64_t begin, int64_t count, double flops, int shared, int stream) {
    if (count > 0 || shared == 1) push(mk(K_GEMM, level, phase, begin, count, flops, cfg, shared), stream);
    if (shared == 1) push(mk(K_BARRIER, level, phase, 0, 0, 0, 0, 0), stream);
  }
  void emit(int level, int phase, int cfg, const std::vector<Pending> &list, bool top) {
    if (list.empty()) return;
    const int bm = cfg_bm(cfg), bn = cfg_bn(cfg);
    const bool multi = top && D.world > 1;
    if (mode == 1 && cfg == 0) {
      // part A: the next block column's tiles (in multi-GPU mode stored into every rank's copy);
      // part B: the rest of the trailing panel (local).  Tile rows have a static owner.
      double all_tiles = 0, flops = 0;
      std::vector<TileRef> pa, pb;
      for (const Pending &pd : list) {
        const GemmProblem &g = D.probs[pd.prob];
        int tr_n = (g.M + bm - 1) / bm, tc_n = (g.N + bn - 1) / bn;
        for (int tc = 0; tc < tc_n; tc++)
          for (int tr = 0; tr < tr_n; tr++) {
            if (g.tri && (tr + 1) * bm - 1 < tc * bn) continue;
            all_tiles += 1;
            if (multi && (pd.row_tile0 + tr) % D.world != D.rank) continue;
            (tc < pd.bcast_tc ? pa : pb).push_back(TileRef{pd.prob, (uint16_t)tr, (uint16_t)tc});
          }
        flops += pd.flops;
      }
      const double per_tile = flops / std::max(1.0, all_tiles);
      depend(0, 1);  // the chain of this block column is done
      int64_t b0 = (int64_t)D.tiles.size();
      D.tiles.insert(D.tiles.end(), pa.begin(), pa.end());
      push_gemm(level, phase, cfg, b0, (int64_t)pa.size(), per_tile * (double)pa.size(), multi ? 1 : 0, 0);
      depend(1, 0);  // the next chain may start as soon as part A is in place
      int64_t b1 = (int64_t)D.tiles.size();
      D.tiles.insert(D.tiles.end(), pb.begin(), pb.end());
      push_gemm(level, phase, cfg, b1, (int64_t)pb.size(), per_tile * (double)pb.size(), multi ? 2 : 0, 0);
      return;
    }
    int64_t begin = (int64_t)D.tiles.size();
    double flops = 0;
    for (const Pending &pd : list) {
      const GemmProblem &g = D.probs[pd.prob];
      int tr_n = (g.M + bm - 1) / bm, tc_n = (g.N + bn - 1) / bn;
      for (int tc = 0; tc < tc_n; tc++)
        for (int tr = 0; tr < tr_n; tr++) {
          if (g.tri && (tr + 1) * bm - 1 < tc * bn) continue;  // wholly above the diagonal
          D.tiles.push_back(TileRef{pd.prob, (uint16_t)tr, (uint16_t)tc});
        }
      flops += pd.flops;
    }
    int64_t count = (int64_t)D.tiles.size() - begin;
    int shared = 0;
    if (multi && mode != 2 && flops >= D.shared_min_flops) {
      // split the tile list across the ranks; every rank builds the same list, keeps its slice
      int64_t lo = count * D.rank / D.world, hi = count * (D.rank + 1) / D.world;
      flops *= (double)(hi - lo) / (double)std::max<int64_t>(1, count);
      begin += lo, count = hi - lo;
      shared = 1;
    }
    const int stream = mode == 2 ? 1 : 0;
    if (stream == 0) depend(0, 1);
    push_gemm(level, phase, cfg, begin, count, flops, shared, stream);
  }
  // tile configuration is decided per launch: larger tiles only pay when they fill the GPU
  void end_gemm(int level, int phase, bool top) {
    int64_t n128 = 0;
    for (const Pending &pd : pend)
      if (is_big(D.probs[pd.prob])) n128 += ntiles(D.probs[pd.prob], D.big_cfg);
    int64_t share = (top && D.world > 1) ? D.world : 1;
    bool use128 = n128 >= (int64_t)D.min_tiles_128 * share;
    // small fronts (bottom of the tree): one warp per 32x32 tile, operands straight from global memory
    const bool small_ok = D.small_front && mode == 0 && !(top && D.world > 1);
    std::vector<Pending> l128, l64, lsmall;
    for (const Pending &pd : pend) {
      const GemmProblem &g = D.probs[pd.prob];
      int ksum = 0;
      for (int c = 0; c < g.contrib_count; c++) ksum += D.contribs[g.contrib_begin + c].K;
      if (small_ok && g.M <= D.small_mn && g.N <= D.small_mn && ksum <= D.small_k) lsmall.push_back(pd);
      else (use128 && is_big(g) ? l128 : l64).push_back(pd);
    }
    emit(level, phase, D.big_cfg, l128, top);
    emit(level, phase, 0, l64, top);
    emit(level, phase, 3, lsmall, top);
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

int build_schedule(const Problem &P, const Symbolic &S, Schedule &D, int rank, int world, bool split_phases, std::string &err, int only_heap) {
  D = Schedule();
  D.rank = rank, D.world = world;
  D.split_phases = split_phases;
  int depth = 0;
  while ((1 << depth) < world) depth++;
  if ((1 << depth) != world || rank < 0 || rank >= world) return err = "world size must be a power of two and 0 <= rank < world", -1;
  if (depth >= P.levels) return err = "more ranks than subtrees", -1;
  if (only_heap && (world != 1 || only_heap < 1 || only_heap > P.N)) return err = "a single-separator schedule needs a single-GPU handle and a valid separator", -1;
  D.depth = depth;
  if (const char *e = getenv("CHOL_BIG_CFG")) D.big_cfg = atoi(e);                    // tuning knob: 1 = 128x128, 2 = 128x64
  if (const char *e = getenv("CHOL_MIN_TILES_128")) D.min_tiles_128 = atoi(e);        // tuning knob
  if (const char *e = getenv("CHOL_SHARED_MIN_FLOPS")) D.shared_min_flops = atof(e);  // tests lower it to split small grids
  if (const char *e = getenv("CHOL_LOOKAHEAD")) D.lookahead = atoi(e) != 0;
  if (const char *e = getenv("CHOL_SMALL_FRONT")) D.small_front = atoi(e) != 0;
  if (const char *e = getenv("CHOL_SMALL_MN")) D.small_mn = atoi(e);
  if (const char *e = getenv("CHOL_SMALL_K")) D.small_k = atoi(e);
  if (const char *e = getenv("CHOL_NBO")) D.nbo = std::max(64, atoi(e) / 64 * 64);  // tuning knob: block-column width
  if (const char *e = getenv("CHOL_NBO_SMALL")) D.nbo_small = atoi(e) > 0 ? std::max(64, atoi(e) / 64 * 64) : 0;
  if (const char *e = getenv("CHOL_NBO_SMALL_MAXN")) D.nbo_small_maxn = atoi(e);
  if (split_phases) D.lookahead = false;  // the piecewise entry points run one phase of one level at a time
  const int L = P.levels, N = P.N;
  const int NB = D.nb, SLAB = D.slab;
  Builder B(P, S, D);
  auto owner_of = [&](int h) -> int {  // -1: shared top separator
    int lv = P.level_of(h);
    return lv < depth ? -1 : (h >> (lv - depth)) - (1 << depth);
  };
  D.top_doubles = world > 1 ? S.poff[1 << depth] : 0;  // panels are laid out in heap order: the top ones come first

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

  // ---- assembly map (fill_block, mmat.rg:529-633, as a scatter).  With several ranks an entry is
  // assembled by the owner of its column separator; top entries by rank 0 only (the copies are summed).
  if (!only_heap) {
    std::vector<int> iperm(P.n), rowheap(P.n);
    for (int p = 0; p < P.n; p++) iperm[P.perm[p]] = p;
    for (int h = 1; h <= N; h++)
      for (int i = 0; i < P.sz[h]; i++) rowheap[P.start[h] + i] = h;
    D.a_off.assign((size_t)P.nz, -1);
    std::vector<int> bad(64, 0);
    parallel_chunks(P.nz, [&](int64_t e0, int64_t e1, int w) {  // every entry is independent
      for (int64_t e = e0; e < e1; e++) {
        if (P.ev[e] == 0.0) continue;
        int pi = iperm[P.ei[e]], pj = iperm[P.ej[e]];
        if (pi < pj) std::swap(pi, pj);
        int hr = rowheap[pi], hc = rowheap[pj];
        int d = P.level_of(hc) - P.level_of(hr);
        if (d < 0 || (hc >> d) != hr) continue;
        int own = owner_of(hc);
        if (world > 1 && !(own == rank || (own < 0 && rank == 0))) continue;
        int r = locate(hc, pi);
        if (r < 0) {
          bad[w] = 1;
          return;
        }
        D.a_off[e] = S.poff[hc] + r + (int64_t)(pj - P.start[hc]) * S.ld[hc];
      }
    });
    for (int b : bad)
      if (b) return err = "internal: nonzero outside the filled pattern", -1;
  }

  std::vector<Pair> pairs;
  for (int lvl = L - 1; lvl >= 0; lvl--) {
    const bool top = lvl < depth;
    if (only_heap && P.level_of(only_heap) != lvl) continue;
    // separators of this level this rank works on
    int h0 = 1 << lvl, h1 = 1 << (lvl + 1);
    if (only_heap) h0 = only_heap, h1 = only_heap + 1;  // debug trace: one fused task group at a time
    if (!top && world > 1) {
      h0 = ((1 << depth) + rank) << (lvl - depth);
      h1 = h0 + (1 << (lvl - depth));
    }
    int maxn = 0;
    for (int h = h0; h < h1; h++) maxn = std::max(maxn, P.sz[h]);
    // block-column width of the level: narrower columns shorten the chain of small fronts (64^3, root 4 096:
    // 28.3 ms at 128 against 29.5 at 256), wide ones keep the trailing updates of the big fronts efficient
    // (128^3: 1 062 ms at 128 against 976 at 256).  The per-level choice is an experiment, off by default.
    const int NBO = (D.nbo_small > 0 && maxn <= D.nbo_small_maxn) ? D.nbo_small : D.nbo;
    const int nouter = (maxn + NBO - 1) / NBO;
    B.depend(1, 0);  // the chain of this level starts after the previous level's updates

    // which == 0: pivot blocks (rows [0, n));  which == 1: off-diagonal rows [r0, R);  which == 2: both at
    // once (rows [0, R)): the default, it halves the number of dependent small launches.  The split form
    // serves the piecewise fused_dpotrf / fused_dtrsm entry points.
    for (int which = (D.split_phases ? 0 : 2); which < (D.split_phases ? 2 : 3); which++) {
      const int phase = which == 0 ? PH_POTRF : which == 1 ? PH_TRSM : (PH_POTRF | PH_TRSM);
      for (int J = 0; J < nouter; J++) {
        const int c0 = J * NBO;
        for (int jj = 0; jj < NBO / NB; jj++) {
          const int d0 = c0 + jj * NB;
          if (d0 >= maxn) break;
          if (which != 1) {
            int64_t b = (int64_t)D.potrf.size();
            for (int h = h0; h < h1; h++) {
              int n = P.sz[h];
              if (n <= d0) continue;
              D.potrf.push_back(PotrfDesc{S.poff[h] + d0 + (int64_t)d0 * S.ld[h], S.ld[h], std::min(NB, n - d0), P.start[h] + d0, 0});
            }
            if ((int64_t)D.potrf.size() > b) B.push(Builder::mk(K_POTRF, lvl, phase, b, (int64_t)D.potrf.size() - b, 0, 0, 0), 1);
          }
          {
            int64_t b = (int64_t)D.trsm_tiles.size();
            for (int h = h0; h < h1; h++) {
              int n = P.sz[h], ld = S.ld[h];
              if (n <= d0) continue;
              int dw = std::min(NB, n - d0);
              int rbeg = which == 1 ? (n + 1) / 2 * 2 : d0 + dw;
              int rend = which == 0 ? n : S.rows[h];
              if (rend <= rbeg) continue;
              D.trsm.push_back(TrsmDesc{S.poff[h] + d0 + (int64_t)d0 * ld, S.poff[h] + rbeg + (int64_t)d0 * ld, ld, dw, rend - rbeg, 0});
              int ns = (rend - rbeg + SLAB - 1) / SLAB;
              for (int s = 0; s < ns; s++) D.trsm_tiles.push_back(TileRef{(int)D.trsm.size() - 1, (uint16_t)(s & 0xffff), (uint16_t)(s >> 16)});
            }
            if ((int64_t)D.trsm_tiles.size() > b) B.push(Builder::mk(K_TRSM, lvl, phase, b, (int64_t)D.trsm_tiles.size() - b, 0, 0, 0), 1);
          }
          // right-looking update of the rest of this block column (K = NB); replicated on a shared top panel
          B.begin_gemm(2);
          for (int h = h0; h < h1; h++) {
            int n = P.sz[h], ld = S.ld[h];
            if (n <= d0) continue;
            int dw = std::min(NB, n - d0), e0 = d0 + dw, cend = std::min(c0 + NBO, n);
            if (e0 >= cend) continue;
            int64_t base = S.poff[h];
            if (which != 1)
              B.add_problem(base + e0 + (int64_t)e0 * ld, ld, (which == 0 ? n : S.rows[h]) - e0, cend - e0, 1, base + e0 + (int64_t)d0 * ld,
                            base + e0 + (int64_t)d0 * ld, ld, ld, dw);
            else {
              int r0 = (n + 1) / 2 * 2, m = S.rows[h] - r0;
              B.add_problem(base + r0 + (int64_t)e0 * ld, ld, m, cend - e0, 0, base + r0 + (int64_t)d0 * ld, base + e0 + (int64_t)d0 * ld, ld, ld, dw);
            }
          }
          B.end_gemm(lvl, phase, top);
        }
        // right-looking update of everything to the right of block column J (K = NBO)
        B.begin_gemm(1);
        for (int h = h0; h < h1; h++) {
          int n = P.sz[h], ld = S.ld[h];
          int c1 = c0 + NBO;
          if (n <= c1) continue;
          int64_t base = S.poff[h];
          const int next_tc = (std::min(NBO, n - c1) + 63) / 64;  // tile columns of the next block column
          if (which != 1)
            B.add_problem(base + c1 + (int64_t)c1 * ld, ld, (which == 0 ? n : S.rows[h]) - c1, n - c1, 1, base + c1 + (int64_t)c0 * ld,
                          base + c1 + (int64_t)c0 * ld, ld, ld, NBO, c1 / 64, next_tc);
          else {
            int r0 = (n + 1) / 2 * 2, m = S.rows[h] - r0;
            B.add_problem(base + r0 + (int64_t)c1 * ld, ld, m, n - c1, 0, base + r0 + (int64_t)c0 * ld, base + c1 + (int64_t)c0 * ld, ld, ld, NBO,
                          r0 / 64, next_tc);
          }
        }
        B.end_gemm(lvl, phase, top);
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
    B.begin_gemm(0);
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
        pf += 2.0 * pairs[c].K * ((double)q.M * q.N - (q.tri ? 0.5 * q.N * (q.N - 1.0) : 0.0));
      }
      D.probs.push_back(g);
      Builder::Pending pd;
      pd.prob = (int)D.probs.size() - 1, pd.flops = pf;
      B.pend.push_back(pd);
      i = j;
    }
    B.end_gemm(lvl, PH_UPDATE, top);

    // the subtrees are done: sum every rank's copy of the top panels before the top is factored
    // One reduction per top panel: only the ranks under that separator (and rank 0, which assembled A's
    // entries) hold contributions, and the strictly upper part of the pivot block is never touched.
    if (world > 1 && lvl == depth) {
      B.depend(0, 1);
      B.push(Builder::mk(K_BARRIER, lvl, PH_UPDATE, 0, 0, 0, 0, 0), 0);
      for (int h = 1; h < (1 << depth); h++) {
        const int lv = P.level_of(h);
        unsigned mask = 1u;  // rank 0
        for (int r = 0; r < world; r++)
          if ((((1 << depth) + r) >> (depth - lv)) == h) mask |= 1u << r;
        // begin = panel offset, count = panel doubles, cfg = contributor mask, shared = heap index of the panel
        B.push(Builder::mk(K_ALLREDUCE, lvl, PH_UPDATE, S.poff[h], S.poff[h + 1] - S.poff[h], 0, (int)mask, h), 0);
      }
      B.push(Builder::mk(K_BARRIER, lvl, PH_UPDATE, 0, 0, 0, 0, 0), 0);
    }
  }
  // the step ends on stream 0
  B.depend(0, 1);
  if (B.pendw[0] >= 0) B.push(Builder::mk(K_NOP, 0, 0, 0, 0, 0, 0, 0), 0);
  if (D.contribs.size() > 0x7fffffffULL || D.probs.size() > 0x7fffffffULL) return err = "schedule too large", -1;
  return 0;
}

}  // namespace chb
