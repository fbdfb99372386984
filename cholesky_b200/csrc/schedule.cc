// Schedule compiler: turns the symbolic structure into the launch list the GPU executes.
//
// Per tree level, leaves first (mmat.rg:1227), three phases that mirror the reference's three
// __demand(__parallel) loops (mmat.rg:1240, 1259, 1293):
//   (a) fused_dpotrf  -> blocked Cholesky of every pivot block of the level, in lock step:
//                        NBO-wide block columns; one panel_kernel launch factors the diagonal block of
//                        a block column and solves every row below it (64-row slabs, csrc/kernels.cuh);
//                        after one, a right-looking grouped GEMM (K = NBO) over the whole trailing
//                        panel, which keeps the GPU full even for one front.
//   (b) fused_dtrsm   -> the same launch applied to the filled off-diagonal rows of the panels
//                        (by default fused with (a): every row below a diagonal block advances together).
//   (c) fused_dsyrk / fused_dgemm -> one grouped GEMM over DESTINATION clusters: every filled
//                        cluster (g, p, ia, jb) owns the ordered list of its contributors
//                        (s ascending), so accumulation is atomic-free and deterministic; the
//                        extend-add index map is the precomputed destination offset.
//
// Look-ahead: the latency-bound chain of a block column (panel_kernel) runs on a second stream.  The trailing update of block column J is split into the tiles of block column
// J + 1 (part A) and the rest (part B); the chain of J + 1 only waits for A, so it overlaps B.
//
// Multi-GPU (world = 2^d ranks, one per GPU): rank r owns the subtree under heap index 2^d + r and
// schedules only its separators on levels >= d.  Its Schur contributions to the top d levels land in
// its own copy of the top panels (partial sums).  Every top panel belongs to the group of ranks under
// it and its stored rows are dealt to them in blocks of 256 (TopGroup, chol_internal.h); from then on
// the owner of a row computes it and nobody else does:
//   * after level d one K_REDUCE per rank sums the group's partial sums of the rows it owns (peer loads);
//   * per 256-wide block column J of a top panel: the owner of the diagonal block factors it and pushes
//     it to every rank (peer stores + flag SLOT_DIAG); every rank of the group solves its own rows
//     below it (stream 3), pushes the pivot-block part of them to the group (flag SLOT_GROUP, the trailing
//     update needs them as its B operand) and, on a background stream, everything else to everybody; then
//     each rank updates its own rows of the trailing matrix (parts A / B as above, stream 0).  The chain
//     of the diagonal blocks (stream 1) runs one block column ahead: the owner of diagonal block J + 1
//     solves its rows of block column J first, applies them to its diagonal block, factors and sends it;
//   * after a top level a world barrier (all pushed rows have landed: every rank now holds the complete
//     factored panels of the level), then its Schur updates, each destination row block computed by
//     its owner from its local copies.
// Every entry of the factor is computed by exactly one rank and copied, so all copies of the top
// panels end bit-identical.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "chol_internal.h"

namespace chb {

namespace {

constexpr int kNoTri = 1 << 30;
// Order of the CTAs of one GEMM problem: groups of 16 tile rows, and inside a group column by column.  The CTAs in
// flight at any time (three per SM) then share 16 row tiles of the A operand and a couple of dozen column tiles of
// the B operand, which fit the 126 MB L2 even at K = 8 128, instead of streaming the whole A operand once per
// tile column.  ncu on the level-3 Schur update of 128^3 (K = 4 009, 123 922 tiles, 136 ms at 96.5 % DMMA pipe):
// DRAM reads 148 GB with plain column-major order against ~10 GB of distinct operands and destinations.
constexpr int kTileRowGroup = 16;

struct Builder {
  const Problem &P;
  const Symbolic &S;
  Schedule &D;
  Builder(const Problem &p, const Symbolic &s, Schedule &d) : P(p), S(s), D(d) {}

  // ---- launches, streams and events
  int last[4] = {-1, -1, -1, -1};   // index of the last launch pushed on each stream
  int pendw[4] = {-1, -1, -1, -1};  // event the next launch of the stream has to wait for
  int chain = 1;                    // stream of the panel launches and chain GEMMs being emitted (1, or 3: the rows stream)
  static Launch mk(int kind, int level, int phase, int64_t begin = 0, int64_t count = 0, double flops = 0, int cfg = 0) {
    Launch l;
    l.kind = kind, l.level = level, l.phase = phase, l.begin = begin, l.count = count, l.flops = flops, l.cfg = cfg;
    l.stream = 0, l.wait_ev = -1, l.rec_ev = -1;
    l.mask = l.sig_mask = l.wait_mask = 0, l.slot = 0, l.seq = 0;
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
    if (pendw[dst] >= 0 && pendw[dst] != ev) push(mk(K_NOP, D.launches[last[src]].level, 0), dst);
    pendw[dst] = ev;
  }
  void sync(int level, int phase, int slot, int64_t seq, unsigned sig, unsigned wait, int stream) {
    Launch l = mk(K_SYNC, level, phase);
    l.slot = slot, l.seq = seq, l.sig_mask = sig, l.wait_mask = wait;
    push(l, stream);
  }
  // rectangles [begin, D.rects.size()) go to the ranks in `mask`; then flag (slot, me) of `sig` rises to seq
  void push_rects(int level, int phase, int64_t begin, unsigned mask, int slot, int64_t seq, unsigned sig, int stream) {
    Launch l = mk(K_PUSH, level, phase, begin, (int64_t)D.rects.size() - begin);
    l.mask = mask, l.slot = slot, l.seq = seq, l.sig_mask = sig;
    if (l.count > 0 || sig) push(l, stream);
  }

  // ---- panel launches (one block column of every panel of the level)
  std::vector<PanelSlab> pdiag, prows;
  int pwmax = 0;
  int add_panel_desc(int64_t off, int ld, int c0, int w, int col0, int ready) {
    D.pdesc.push_back(PanelDesc{off, ld, c0, w, col0, 4 * (int)D.pdesc.size(), ready});
    pwmax = std::max(pwmax, w);
    return (int)D.pdesc.size() - 1;
  }
  void add_diag_slabs(int desc) {
    const PanelDesc &d = D.pdesc[desc];
    for (int t = 0; t * 64 < d.w; t++) pdiag.push_back(PanelSlab{desc, d.c0 + 64 * t, std::min(64, d.w - 64 * t), t});
  }
  void add_row_slabs(int desc, int rb, int re) {
    for (int r = rb; r < re; r += 64) prows.push_back(PanelSlab{desc, r, std::min(64, re - r), -1});
  }
  // diagonal slabs first, in tile order: a slab only ever waits for CTAs earlier in the grid
  void end_panel(int level, int phase) {
    std::stable_sort(pdiag.begin(), pdiag.end(), [](const PanelSlab &a, const PanelSlab &b) { return a.t < b.t; });
    const int64_t b = (int64_t)D.pslabs.size();
    D.pslabs.insert(D.pslabs.end(), pdiag.begin(), pdiag.end());
    D.pslabs.insert(D.pslabs.end(), prows.begin(), prows.end());
    if ((int64_t)D.pslabs.size() > b) push(mk(K_PANEL, level, phase, b, (int64_t)D.pslabs.size() - b, 0, pwmax), chain);
    pdiag.clear(), prows.clear(), pwmax = 0;
  }

  // ---- rows below already factored diagonal blocks, the throughput way: per 64-column tile step one trsm_tile
  // launch and one grouped GEMM (K = 64) that updates the rest of the block column, over all row ranges at once
  struct RowRange {
    int64_t base;  // panel offset
    int ld, c0, w, rb, re;
  };
  std::vector<RowRange> bulk;
  int64_t bulk_slabs() const {
    int64_t s = 0;
    for (const RowRange &r : bulk) s += (r.re - r.rb + 63) / 64;
    return s;
  }
  void emit_bulk_rows(int level, int phase) {
    int wmax = 0;
    for (const RowRange &r : bulk) wmax = std::max(wmax, r.w);
    for (int d0 = 0; d0 < wmax; d0 += 64) {
      const int64_t b = (int64_t)D.trsm_tiles.size();
      for (const RowRange &r : bulk) {
        if (r.w <= d0) continue;
        const int dw = std::min(64, r.w - d0), cd = r.c0 + d0;
        D.trsm.push_back(TrsmDesc{r.base + cd + (int64_t)cd * r.ld, r.base + r.rb + (int64_t)cd * r.ld, r.ld, dw, r.re - r.rb, 0});
        for (int sl = 0; sl < (r.re - r.rb + 127) / 128; sl++) D.trsm_tiles.push_back(TileRef{(int)D.trsm.size() - 1, (uint16_t)(sl & 0xffff), (uint16_t)(sl >> 16)});
      }
      if ((int64_t)D.trsm_tiles.size() > b) push(mk(K_TRSM, level, phase, b, (int64_t)D.trsm_tiles.size() - b), chain);
      begin_gemm(2);
      for (const RowRange &r : bulk) {
        const int dw = std::min(64, r.w - d0), e0 = d0 + dw;
        if (r.w <= e0) continue;
        add_problem(r.base + r.rb + (int64_t)(r.c0 + e0) * r.ld, r.ld, r.re - r.rb, r.w - e0, 0, r.base + r.rb + (int64_t)(r.c0 + d0) * r.ld,
                    r.base + r.c0 + e0 + (int64_t)(r.c0 + d0) * r.ld, r.ld, r.ld, dw);
      }
      end_gemm(level, phase, false);
    }
    bulk.clear();
  }

  // ---- grouped GEMM launches
  struct Pending {
    int prob;
    double flops;
    // trailing update of a panel (mode 1): tile row tr of the problem is tile row row_tile0 + tr of the
    // panel; tile columns < bcast_tc belong to the next block column (part A)
    int row_tile0 = -1, bcast_tc = 0;
  };
  // deep look-ahead on a top panel (mode 1): q tiles per row block; the tiles of the diagonal block of the next block
  // column were already updated by its owner (skip them), those of the diagonal block after that join part A
  int deep_q = 0;
  std::vector<Pending> pend;
  int mode = 0;                   // 0: Schur update; 1: trailing update of a block column (parts A / B); 2: chain GEMM (K = 64)
  const TopGroup *own = nullptr;  // mode 1 on a top panel: only the tile rows this rank owns
  void begin_gemm(int m) {
    pend.clear();
    mode = m;
  }
  static double gemm_flops(int M, int N, int K, int tri) { return 2.0 * K * ((double)M * N - ((tri & 1) ? 0.5 * N * (N - 1.0) : 0.0)); }
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
    pd.flops = gemm_flops(M, N, K, tri);  // executed flops: the strict upper triangle of the leading N x N part is skipped when tri
    pd.row_tile0 = row_tile0, pd.bcast_tc = bcast_tc;
    pend.push_back(pd);
  }
  void push_gemm(int level, int phase, int cfg, int64_t begin, int64_t count, double flops, int stream) {
    if (count > 0) push(mk(K_GEMM, level, phase, begin, count, flops, cfg), stream);
  }
  void emit(int level, int phase, int cfg, const std::vector<Pending> &list) {
    if (list.empty()) return;
    const int bm = cfg == 3 ? 32 : 64, bn = bm;
    if (mode == 1) {
      // part A: the next block column's tiles; part B: the rest of the trailing panel
      double all_tiles = 0, flops = 0;
      std::vector<TileRef> pa, pb;
      for (const Pending &pd : list) {
        const GemmProblem &g = D.probs[pd.prob];
        int tr_n = (g.M + bm - 1) / bm, tc_n = (g.N + bn - 1) / bn;
        for (int tb = 0; tb < tr_n; tb += kTileRowGroup)  // L2-friendly order, see kTileRowGroup
          for (int tc = 0; tc < tc_n; tc++)
            for (int tr = tb; tr < std::min(tr_n, tb + kTileRowGroup); tr++) {
              if ((g.tri & 1) && (tr + 1) * bm - 1 < tc * bn) continue;
              all_tiles += 1;
              if (own && own->owner((pd.row_tile0 + tr) * bm / own->rb) != D.rank) continue;
              bool a = tc < pd.bcast_tc;
              if (deep_q) {
                if (tr < deep_q && tc < deep_q) continue;                                           // D(J+1): done early
                if (tr >= deep_q && tr < 2 * deep_q && tc >= deep_q && tc < 2 * deep_q) a = true;   // D(J+2): needed early
              }
              (a ? pa : pb).push_back(TileRef{pd.prob, (uint16_t)tr, (uint16_t)tc});
            }
        flops += pd.flops;
      }
      const double per_tile = flops / std::max(1.0, all_tiles);
      depend(0, 1);  // the chain of this block column is done
      int64_t b0 = (int64_t)D.tiles.size();
      D.tiles.insert(D.tiles.end(), pa.begin(), pa.end());
      push_gemm(level, phase, cfg, b0, (int64_t)pa.size(), per_tile * (double)pa.size(), 0);
      depend(1, 0);  // the next chain may start as soon as part A is in place
      if (deep_q) depend(3, 0);
      int64_t b1 = (int64_t)D.tiles.size();
      D.tiles.insert(D.tiles.end(), pb.begin(), pb.end());
      push_gemm(level, phase, cfg, b1, (int64_t)pb.size(), per_tile * (double)pb.size(), 0);
      return;
    }
    int64_t begin = (int64_t)D.tiles.size();
    double flops = 0;
    for (const Pending &pd : list) {
      const GemmProblem &g = D.probs[pd.prob];
      int tr_n = (g.M + bm - 1) / bm, tc_n = (g.N + bn - 1) / bn;
      for (int tb = 0; tb < tr_n; tb += kTileRowGroup)
        for (int tc = 0; tc < tc_n; tc++)
          for (int tr = tb; tr < std::min(tr_n, tb + kTileRowGroup); tr++) {
            if ((g.tri & 1) && (tr + 1) * bm - 1 < tc * bn) continue;  // wholly above the diagonal
            D.tiles.push_back(TileRef{pd.prob, (uint16_t)tr, (uint16_t)tc});
          }
      flops += pd.flops;
    }
    const int stream = mode == 2 ? chain : 0;
    if (stream == 0) depend(0, 1);
    push_gemm(level, phase, cfg, begin, (int64_t)D.tiles.size() - begin, flops, stream);
  }
  void end_gemm(int level, int phase, bool small_ok) {
    // small fronts (bottom of the tree): one warp per 32x32 tile, operands straight from global memory
    small_ok = small_ok && D.small_front && mode == 0;
    std::vector<Pending> l64, lsmall;
    for (const Pending &pd : pend) {
      const GemmProblem &g = D.probs[pd.prob];
      int ksum = 0;
      for (int c = 0; c < g.contrib_count; c++) ksum += D.contribs[g.contrib_begin + c].K;
      (small_ok && g.M <= D.small_mn && g.N <= D.small_mn && ksum <= D.small_k ? lsmall : l64).push_back(pd);
    }
    emit(level, phase, 0, l64);
    emit(level, phase, 3, lsmall);
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
  if (split_phases && world != 1) return err = "the piecewise fused tasks run on a single-GPU handle", -1;
  D.depth = depth;
  if (const char *e = getenv("CHOL_LOOKAHEAD")) D.lookahead = atoi(e) != 0;
  if (const char *e = getenv("CHOL_SMALL_FRONT")) D.small_front = atoi(e) != 0;
  if (const char *e = getenv("CHOL_SMALL_MN")) D.small_mn = atoi(e);
  if (const char *e = getenv("CHOL_SMALL_K")) D.small_k = atoi(e);
  if (const char *e = getenv("CHOL_NBO")) D.nbo = std::min(256, std::max(64, atoi(e) / 64 * 64));  // tuning knob: block-column width (single GPU)
  if (const char *e = getenv("CHOL_FUSED_ROWS_MAX")) D.fused_rows_max = D.fused_rows_max_top = atoi(e);
  if (const char *e = getenv("CHOL_FUSED_ROWS_MAX_TOP")) D.fused_rows_max_top = atoi(e);
  if (const char *e = getenv("CHOL_ROW_BLOCK")) D.row_block = std::min(kRowBlock, std::max(64, atoi(e) / 64 * 64));
  if (split_phases) D.lookahead = false;  // the piecewise entry points run one phase of one level at a time
  const int L = P.levels, N = P.N;
  const int NBO = D.nbo, RB = D.row_block;
  const unsigned wmask = (1u << world) - 1u, me = 1u << rank;
  Builder B(P, S, D);
  auto owner_of = [&](int h) -> int {  // -1: top separator (rows dealt to its group)
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
  // assembled by the owner of its column separator; entries of the top panels by the owner of their row.
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
        if (world > 1 && own >= 0 && own != rank) continue;
        int r = locate(hc, pi);
        if (r < 0) {
          bad[w] = 1;
          return;
        }
        if (world > 1 && own < 0 && top_group(hc, P.level_of(hc), depth, RB).owner(r / RB) != rank) continue;
        D.a_off[e] = S.poff[hc] + r + (int64_t)(pj - P.start[hc]) * S.ld[hc];
      }
    });
    for (int b : bad)
      if (b) return err = "internal: nonzero outside the filled pattern", -1;
  }

  // Schur destination (panel p, stored row crow, column ccol, M x N, tri) with contributors cs[0 .. cnt).
  // `split`: the destination is a top panel updated from a top level -- every 256-row block of it is
  // computed by its owner, this rank keeps its own blocks.  Operand tiles are fetched with 16-byte bulk
  // copies, so a sub-problem starts on an even operand row; when the ownership boundary falls on an odd
  // one the sub-problem starts one row early and that row is masked in the epilogue (tri bit 1).
  auto add_dest = [&](const Pair *cs, size_t cnt, bool split) -> int {
    const Pair &q = cs[0];
    for (size_t c = 0; c < cnt; c++)
      if (cs[c].M != q.M || cs[c].N != q.N || cs[c].tri != q.tri) return err = "internal: contributors of one destination cluster disagree on its shape", -1;
    const int ldc = S.ld[q.p];
    const int64_t c_off = S.poff[q.p] + q.crow + (int64_t)q.ccol * ldc;
    auto sub = [&](int m0, int M, int n0, int N, int tri) {  // rows [m0, m0 + M) x cols [n0, n0 + N) of the destination
      if (M <= 0 || N <= 0) return;
      GemmProblem g;
      g.c_off = c_off + m0 + (int64_t)n0 * ldc, g.ldc = ldc, g.M = M, g.N = N, g.tri = tri;
      g.contrib_begin = (int)D.contribs.size(), g.contrib_count = (int)cnt;
      double pf = 0;
      for (size_t c = 0; c < cnt; c++) {
        D.contribs.push_back(GemmContrib{cs[c].a_off + m0, cs[c].b_off + n0, cs[c].ld, cs[c].ld, cs[c].K, 0});
        pf += Builder::gemm_flops(M, N, cs[c].K, tri);
      }
      D.probs.push_back(g);
      Builder::Pending pd;
      pd.prob = (int)D.probs.size() - 1, pd.flops = pf;
      B.pend.push_back(pd);
    };
    if (!split) {
      sub(0, q.M, 0, q.N, q.tri);
      return 0;
    }
    if (q.tri && q.M != q.N) return err = "internal: a diagonal destination cluster is not square", -1;
    const TopGroup grp = top_group(q.p, P.level_of(q.p), depth, RB);
    for (int m0 = 0; m0 < q.M;) {
      const int blk = (q.crow + m0) / RB;
      const int m1 = std::min(q.M, (blk + 1) * RB - q.crow);
      if (grp.owner(blk) == rank) {
        const int e0 = m0 & ~1, skip = (m0 & 1) ? 2 : 0;
        if (!q.tri) sub(e0, m1 - e0, 0, q.N, skip);
        else {
          sub(e0, m1 - e0, 0, e0, skip);        // columns left of the diagonal block of these rows
          sub(e0, m1 - e0, e0, m1 - e0, 1 | skip);  // the diagonal block itself
        }
      }
      m0 = m1;
    }
    return 0;
  };

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
    B.depend(1, 0);  // the chain of this level starts after the previous level's updates
    B.depend(3, 0);

    if (!top) {
      int maxn = 0;
      for (int h = h0; h < h1; h++) maxn = std::max(maxn, P.sz[h]);
      const int nouter = (maxn + NBO - 1) / NBO;
      // which == 0: pivot blocks (rows [0, n));  which == 1: off-diagonal rows [r0, R);  which == 2: both at
      // once (rows [0, R)): the default, it halves the number of dependent small launches.  The split form
      // serves the piecewise fused_dpotrf / fused_dtrsm entry points.
      for (int which = (D.split_phases ? 0 : 2); which < (D.split_phases ? 2 : 3); which++) {
        const int phase = which == 0 ? PH_POTRF : which == 1 ? PH_TRSM : (PH_POTRF | PH_TRSM);
        for (int J = 0; J < nouter; J++) {
          const int c0 = J * NBO;
          // the block column of every panel of the level: diagonal blocks and the rows below them
          for (int h = h0; h < h1; h++) {
            const int n = P.sz[h];
            if (n <= c0) continue;
            const int w = std::min(NBO, n - c0), c1 = c0 + w, r0 = (n + 1) / 2 * 2;
            const int rb = which == 1 ? r0 : (c1 < n ? c1 : r0), re = which == 0 ? n : S.rows[h];
            if (re > rb) B.bulk.push_back(Builder::RowRange{S.poff[h], S.ld[h], c0, w, rb, re});
          }
          const bool fused = B.bulk_slabs() <= D.fused_rows_max;
          for (int h = h0; h < h1; h++) {
            const int n = P.sz[h];
            if (n <= c0) continue;
            const int w = std::min(NBO, n - c0), c1 = c0 + w, r0 = (n + 1) / 2 * 2;
            const int rb = which == 1 ? r0 : (c1 < n ? c1 : r0), re = which == 0 ? n : S.rows[h];
            if (which == 1 && !fused) continue;
            const int desc = B.add_panel_desc(S.poff[h], S.ld[h], c0, w, P.start[h] + c0, which == 1 ? 1 : 0);
            if (which != 1) B.add_diag_slabs(desc);
            if (fused && re > rb) B.add_row_slabs(desc, rb, re);
          }
          B.end_panel(lvl, phase);
          if (fused) B.bulk.clear();
          else B.emit_bulk_rows(lvl, phase);
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
          B.end_gemm(lvl, phase, false);
        }
      }
    } else {
      // ---- a top level: this rank works on the one separator above its subtree, with the rest of its group
      const int p = ((1 << depth) + rank) >> (depth - lvl);
      const TopGroup grp = top_group(p, lvl, depth, RB);
      const unsigned gmask = grp.mask();
      const int n = P.sz[p], R = S.rows[p], ld = S.ld[p], r0 = (n + 1) / 2 * 2;
      const int64_t base = S.poff[p];
      const int phase = PH_POTRF | PH_TRSM;
      const int nblk = (R + RB - 1) / RB;
      const int64_t lvl_seq = (int64_t)(depth - lvl) << 24;
      B.own = &grp;
      for (int J = 0; J * RB < n; J++) {
        const int c0 = J * RB, w = std::min(RB, n - c0), c1 = c0 + w;
        const int diag_owner = grp.owner(J);
        const int below0 = c1 == n ? r0 : c1;  // first stored row below the diagonal block
        // this rank's rows below the diagonal block, block by block
        std::vector<std::pair<int, int>> mine;  // [begin, end)
        for (int b = J; b < nblk; b++) {
          if (grp.owner(b) != rank) continue;
          const int rb = std::max(b * RB, below0), re = std::min((b + 1) * RB, R);
          if (re > rb) mine.push_back({rb, re});
        }
        // ---- the chain of the diagonal blocks runs ahead of the group's row exchange (deep look-ahead; the version
        // that factored diagonal block J+1 after the trailing update of J: 158.3 ms on 8 GPUs against 154.1).
        // Stream 1: the owner of diagonal block J+1 solves ITS rows of block column J as soon as L(J,J) is there,
        // applies them to its diagonal block (the last update it is missing), factors it and sends it out.
        // Stream 3: every rank solves the rest of its rows of block column J and pushes them to the group.
        // Stream 0: the trailing update of J waits for the group's rows, not for the diagonal chain.
        const bool has_next = c1 < n;
        const int next_owner = has_next ? grp.owner(J + 1) : -1;
        const int w2 = has_next ? std::min(RB, n - c1) : 0;
        auto diag_and_push = [&](int JJ, int cc0, int ww) {
          B.chain = 1;
          B.add_diag_slabs(B.add_panel_desc(base, ld, cc0, ww, P.start[p] + cc0, 0));
          B.end_panel(lvl, phase);
          int64_t rb = (int64_t)D.rects.size();
          D.rects.push_back(RectDesc{base + cc0 + (int64_t)cc0 * ld, ld, ww, ww, 0, 0u, 0});
          B.push_rects(lvl, phase, rb, wmask & ~me, SLOT_DIAG, lvl_seq + JJ + 1, gmask & ~me, 1);
        };
        if (J == 0 && rank == diag_owner) diag_and_push(0, c0, w);  // (later diagonal blocks were factored one step ahead)
        std::vector<std::pair<int, int>> early, rest;
        for (auto &rg : mine) (rank == next_owner && rg.first / RB == J + 1 ? early : rest).push_back(rg);
        if (rank == next_owner) {
          if (rank != diag_owner) B.sync(lvl, phase, SLOT_DIAG, lvl_seq + J + 1, 0, 1u << diag_owner, 1);
          B.chain = 1;
          if (!early.empty()) {
            const int desc = B.add_panel_desc(base, ld, c0, w, P.start[p] + c0, 1);
            for (auto &rg : early) B.add_row_slabs(desc, rg.first, rg.second);
            B.end_panel(lvl, phase);
          }
          B.depend(3, 1);  // the rows stream pushes these rows along with the rest
          B.begin_gemm(2);
          // (all rows of row block J+1: below a narrower last diagonal block they are off-diagonal rows, which the
          // trailing update skips along with the diagonal block's)
          B.add_problem(base + c1 + (int64_t)c1 * ld, ld, std::min(RB, R - c1), w2, 1, base + c1 + (int64_t)c0 * ld, base + c1 + (int64_t)c0 * ld, ld,
                        ld, w);
          B.end_gemm(lvl, phase, false);
          diag_and_push(J + 1, c1, w2);
        }
        B.chain = 3;
        if (rank == diag_owner) B.depend(3, 1);  // (its own diagonal block: stream order of stream 1)
        else B.sync(lvl, phase, SLOT_DIAG, lvl_seq + J + 1, 0, 1u << diag_owner, 3);
        for (auto &rg : rest) B.bulk.push_back(Builder::RowRange{base, ld, c0, w, rg.first, rg.second});
        if (B.bulk_slabs() <= D.fused_rows_max_top) {
          if (!rest.empty()) {
            const int desc = B.add_panel_desc(base, ld, c0, w, P.start[p] + c0, 1);
            for (auto &rg : rest) B.add_row_slabs(desc, rg.first, rg.second);
            B.end_panel(lvl, phase);
          }
          B.bulk.clear();
        } else
          B.emit_bulk_rows(lvl, phase);
        B.chain = 1;
        {
          int64_t rb = (int64_t)D.rects.size();
          for (auto &rg : mine)
            if (rg.first < n) D.rects.push_back(RectDesc{base + rg.first + (int64_t)c0 * ld, ld, std::min(rg.second, n) - rg.first, w, kNoTri, 0u, 0});
          B.push_rects(lvl, phase, rb, gmask & ~me, SLOT_GROUP, lvl_seq + J + 1, gmask & ~me, 3);
        }
        {
          B.depend(2, 3);
          int64_t rb = (int64_t)D.rects.size();
          for (auto &rg : mine)
            if (rg.first < n) D.rects.push_back(RectDesc{base + rg.first + (int64_t)c0 * ld, ld, std::min(rg.second, n) - rg.first, w, kNoTri, 0u, 0});
          B.push_rects(lvl, phase, rb, wmask & ~gmask, 0, 0, 0, 2);
          rb = (int64_t)D.rects.size();
          for (auto &rg : mine)
            if (rg.second > r0) D.rects.push_back(RectDesc{base + std::max(rg.first, r0) + (int64_t)c0 * ld, ld, rg.second - std::max(rg.first, r0), w, kNoTri, 0u, 0});
          B.push_rects(lvl, phase, rb, wmask & ~me, 0, 0, 0, 2);
        }
        B.depend(0, 3);
        B.sync(lvl, phase, SLOT_GROUP, lvl_seq + J + 1, 0, gmask & ~me, 0);
        B.begin_gemm(1);
        if (has_next) {
          const int next_tc = (w2 + 63) / 64;
          B.add_problem(base + c1 + (int64_t)c1 * ld, ld, R - c1, n - c1, 1, base + c1 + (int64_t)c0 * ld, base + c1 + (int64_t)c0 * ld, ld, ld, w,
                        c1 / 64, next_tc);
        }
        B.deep_q = RB / 64;
        B.end_gemm(lvl, phase, false);
        B.deep_q = 0;
      }
      B.own = nullptr;
      // every rank holds the complete panels of the level once all pushes have landed
      B.depend(0, 1);
      B.depend(0, 2);
      B.depend(0, 3);
      B.sync(lvl, PH_UPDATE, SLOT_WORLD, lvl_seq + 1, wmask & ~me, wmask & ~me, 0);
      h0 = 1 << lvl, h1 = 1 << (lvl + 1);  // the Schur updates below take contributions from every panel of the level
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
      if (add_dest(&pairs[i], j - i, top)) return -1;
      i = j;
    }
    B.end_gemm(lvl, PH_UPDATE, !top);

    // The subtrees are done: every rank sums, for the rows it owns in each top panel above its subtree, the
    // partial sums of that panel's group.  A subtree touches only the part of the top panels that borders its
    // box, so the sum is taken block by block (row block x column block of the panel) over the ranks whose
    // subtrees have a Schur destination in that block, and blocks nobody else touched are left alone; the
    // strictly upper part of a pivot block is never touched at all.
    if (world > 1 && lvl == depth) {
      B.depend(0, 1);
      B.sync(lvl, PH_UPDATE, SLOT_WORLD, 1, wmask & ~me, wmask & ~me, 0);
      // touched[lv][rb * ncb + cb]: ranks with a contribution to block (rb, cb) of this rank's level-lv panel
      std::vector<std::vector<unsigned>> touched(depth);
      std::vector<int> ncbs(depth, 0);
      for (int lv = 0; lv < depth; lv++) {
        const int p = ((1 << depth) + rank) >> (depth - lv);
        ncbs[lv] = (P.sz[p] + RB - 1) / RB;
        touched[lv].assign((size_t)((S.rows[p] + RB - 1) / RB) * std::max(1, ncbs[lv]), 0u);
      }
      std::vector<std::vector<std::vector<unsigned>>> part(host_threads(), touched);
      parallel_chunks((int64_t)N - (1 << depth) + 1, [&](int64_t i0, int64_t i1, int wk) {
        auto &mine = part[wk];
        for (int64_t hi = i0; hi < i1; hi++) {
          const int hs = (1 << depth) + (int)hi;  // every separator below the top levels
          const int q = owner_of(hs);
          if (q == rank) continue;  // this rank's own partial sums are already in its copy
          const int64_t s0 = S.seg_ptr[hs] + 1, s1 = S.seg_ptr[hs + 1];
          for (int64_t j = s0; j < s1; j++) {
            const Seg &b = S.segs[j];
            const int p = b.anc, lv = P.level_of(p);
            if (lv >= depth || p != (((1 << depth) + rank) >> (depth - lv))) continue;  // not a top panel above this rank
            const TopGroup grp = top_group(p, lv, depth, RB);
            for (int64_t i = j; i < s1; i++) {
              const Seg &a = S.segs[i];
              const int crow = locate(p, P.start[a.anc] + a.lo);
              if (crow < 0) continue;  // (reported by the Schur pass of the owning rank)
              for (int rb = crow / RB; rb <= (crow + a.hi - a.lo - 1) / RB; rb++) {
                if (grp.owner(rb) != rank) continue;
                for (int cb = b.lo / RB; cb <= (b.hi - 1) / RB; cb++) mine[lv][(size_t)rb * ncbs[lv] + cb] |= 1u << q;
              }
            }
          }
        }
      });
      for (auto &pt : part)
        for (int lv = 0; lv < depth; lv++)
          for (size_t i = 0; i < touched[lv].size(); i++) touched[lv][i] |= pt[lv][i];
      for (int lv = depth - 1; lv >= 0; lv--) {
        const int p = ((1 << depth) + rank) >> (depth - lv);
        const TopGroup grp = top_group(p, lv, depth, RB);
        const int R = S.rows[p], n = P.sz[p];
        int64_t rb0 = (int64_t)D.rects.size();
        for (int rb = 0; rb * RB < R; rb++) {
          if (grp.owner(rb) != rank) continue;
          for (int cb = 0; cb * RB < n; cb++) {
            const unsigned m = touched[lv][(size_t)rb * ncbs[lv] + cb];
            const int r0b = rb * RB, c0b = cb * RB, rows = std::min(RB, R - r0b);
            if (!m || (r0b < n && r0b + rows - 1 < c0b)) continue;  // nobody else contributed / wholly above the diagonal
            D.rects.push_back(RectDesc{S.poff[p] + r0b + (int64_t)c0b * S.ld[p], S.ld[p], rows, std::min(RB, n - c0b), r0b - c0b, m | me, 0});
          }
        }
        Launch l = Builder::mk(K_REDUCE, lvl, PH_UPDATE, rb0, (int64_t)D.rects.size() - rb0);
        l.mask = grp.mask();
        if (l.count > 0) B.push(l, 0);
      }
      B.sync(lvl, PH_UPDATE, SLOT_WORLD, 2, wmask & ~me, wmask & ~me, 0);
    }
  }
  // the step ends on stream 0
  B.depend(0, 1);
  B.depend(0, 2);
  B.depend(0, 3);
  if (B.pendw[0] >= 0) B.push(Builder::mk(K_NOP, 0, 0), 0);
  if (D.contribs.size() > 0x7fffffffULL || D.probs.size() > 0x7fffffffULL) return err = "schedule too large", -1;
  return 0;
}

}  // namespace chb
