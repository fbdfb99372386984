// Solve schedule: forward substitution leaves -> root, backward root -> leaves (mmat.rg:1394-1479),
// blocked by 256 columns inside a pivot block (one solve_block launch for the diagonal blocks, one GEMV launch for
// the rest of the pivot block), all separators of a tree level in lock step.
#include "solve.h"

#include <algorithm>

namespace chb {

namespace {
constexpr int NB = 256, SLAB = 128, COLG = 8;  // NB: block-column width of the sweeps (kSolveWB)
constexpr int KSPLIT = 512;  // a pull tile reads at most this many contributor columns

struct SegRef {
  int grow;  // global permuted row of the segment start
  int rows;
  int h;
  int64_t p_off;
  int ld, K, x0;
};
}  // namespace

int build_solve(const Problem &P, const Symbolic &S, SolveSchedule &V, int rank, int world, std::string &err) {
  V = SolveSchedule();
  const int L = P.levels;
  // Partitioned handle (world = 2^depth ranks, schedule.cc): a rank sweeps the separators of its own
  // subtree on tree levels >= depth and every separator of the shared top levels (all ranks hold
  // identical top panels after the factorization).  Between the two parts of the forward sweep the top
  // rows of the right-hand side are summed over the ranks (SK_EXCHANGE): the pulls of a subtree only
  // carry that subtree's contributions to them.
  int depth = 0;
  while ((1 << depth) < world) depth++;
  if ((1 << depth) != world || rank < 0 || rank >= world || depth >= L) return err = "bad partition for the solve schedule", -1;
  V.rank = rank, V.world = world, V.depth = depth;
  V.top_row0 = world > 1 ? P.start[(1 << depth) - 1] : P.n;  // top separators carry the highest labels: the last rows
  auto level_range = [&](int lvl, int &h0, int &h1) {
    h0 = 1 << lvl, h1 = 1 << (lvl + 1);
    if (world > 1 && lvl >= depth) {
      h0 = ((1 << depth) + rank) << (lvl - depth);
      h1 = h0 + (1 << (lvl - depth));
    }
  };
  // row map of the off-diagonal part of every panel
  std::vector<int64_t> map_off(P.N + 2, 0);
  for (int h = 1; h <= P.N; h++) {
    const int n = P.sz[h], r0 = (n + 1) / 2 * 2;
    map_off[h] = (int64_t)V.rowmap.size();
    V.rowmap.resize(V.rowmap.size() + (size_t)std::max(0, S.rows[h] - r0), -1);
    for (int64_t s = S.seg_ptr[h] + 1; s < S.seg_ptr[h + 1]; s++) {
      const Seg &sg = S.segs[s];
      for (int r = 0; r < sg.hi - sg.lo; r++) V.rowmap[map_off[h] + (sg.off - r0) + r] = P.start[sg.anc] + sg.lo + r;
    }
  }
  auto add_gemv = [&](const SolveGemv &g) {
    V.gemv.push_back(g);
    return (int)V.gemv.size() - 1;
  };

  // ---- forward: leaves to root
  std::vector<SegRef> refs;
  for (int lvl = L - 1; lvl >= 0; lvl--) {
    int h0, h1;
    level_range(lvl, h0, h1);
    if (world > 1 && lvl == depth - 1) V.launches.push_back(SolveLaunch{SK_EXCHANGE, 0, 0, lvl, 0});
    int maxn = 0;
    for (int h = h0; h < h1; h++) maxn = std::max(maxn, P.sz[h]);
    for (int d0 = 0; d0 < maxn; d0 += NB) {
      int64_t b = (int64_t)V.tiles.size();
      for (int h = h0; h < h1; h++) {
        int n = P.sz[h];
        if (n <= d0) continue;
        V.tiles.push_back(SolveTile{S.poff[h] + d0 + (int64_t)d0 * S.ld[h], S.ld[h], std::min(NB, n - d0), P.start[h] + d0, 0});
      }
      if ((int64_t)V.tiles.size() > b) V.launches.push_back(SolveLaunch{SK_TILE_F, b, (int64_t)V.tiles.size() - b, lvl, std::min(NB, maxn - d0)});
      int64_t gb = (int64_t)V.gemv_tiles.size();
      for (int h = h0; h < h1; h++) {
        int n = P.sz[h], ld = S.ld[h];
        if (n <= d0) continue;
        int dw = std::min(NB, n - d0), rows = n - (d0 + dw);
        if (rows <= 0) continue;
        int g = add_gemv(SolveGemv{S.poff[h] + (d0 + dw) + (int64_t)d0 * ld, ld, rows, dw, P.start[h] + d0, P.start[h] + d0 + dw, 0});
        for (int s = 0; s < (rows + SLAB - 1) / SLAB; s++) V.gemv_tiles.push_back(TileRef{g, (uint16_t)(s & 0xffff), (uint16_t)(s >> 16)});
      }
      if ((int64_t)V.gemv_tiles.size() > gb) V.launches.push_back(SolveLaunch{SK_GEMV_F, gb, (int64_t)V.gemv_tiles.size() - gb, lvl, 0});
    }
    // ancestors pull the level's contributions: all segments that hit one ancestor row cluster
    refs.clear();
    for (int h = h0; h < h1; h++)
      for (int64_t s = S.seg_ptr[h] + 1; s < S.seg_ptr[h + 1]; s++) {
        const Seg &sg = S.segs[s];
        refs.push_back(SegRef{P.start[sg.anc] + sg.lo, sg.hi - sg.lo, h, S.poff[h] + sg.off, S.ld[h], P.sz[h], P.start[h]});
      }
    std::sort(refs.begin(), refs.end(), [](const SegRef &a, const SegRef &b) { return a.grow != b.grow ? a.grow < b.grow : a.h < b.h; });
    int64_t pb = (int64_t)V.pull_tiles.size(), sb = (int64_t)V.pull_sums.size(), slots = 0;
    for (size_t i = 0; i < refs.size();) {
      size_t j = i;
      while (j < refs.size() && refs[j].grow == refs[i].grow) j++;
      PullDest d{refs[i].grow, refs[i].rows, (int)V.pull_contrib.size(), (int)(j - i)};
      for (size_t c = i; c < j; c++) {
        if (refs[c].rows != refs[i].rows) return err = "internal: solve contributors disagree on a cluster's size", -1;
        V.pull_contrib.push_back(PullContrib{refs[c].p_off, refs[c].ld, refs[c].K, refs[c].x0, 0});
      }
      V.pull.push_back(d);
      int ktot = 0;
      for (size_t c = i; c < j; c++) ktot += refs[c].K;
      const int parts = (ktot + KSPLIT - 1) / KSPLIT;
      for (int s = 0; s < (d.rows + SLAB - 1) / SLAB; s++) {
        if (parts <= 1) {
          V.pull_tiles.push_back(PullTile{(int)V.pull.size() - 1, s, 0, ktot, -1});
          continue;
        }
        V.pull_sums.push_back(PullSum{d.y0 + s * SLAB, std::min(SLAB, d.rows - s * SLAB), (int)slots, parts});
        for (int q = 0; q < parts; q++)
          V.pull_tiles.push_back(PullTile{(int)V.pull.size() - 1, s, (int)((int64_t)ktot * q / parts), (int)((int64_t)ktot * (q + 1) / parts), (int)slots++});
      }
      i = j;
    }
    V.pull_slots = std::max(V.pull_slots, slots);
    if ((int64_t)V.pull_tiles.size() > pb) V.launches.push_back(SolveLaunch{SK_PULL, pb, (int64_t)V.pull_tiles.size() - pb, lvl, 0});
    if ((int64_t)V.pull_sums.size() > sb) V.launches.push_back(SolveLaunch{SK_PULL_SUM, sb, (int64_t)V.pull_sums.size() - sb, lvl, 0});
  }

  // ---- backward: root to leaves
  for (int lvl = 0; lvl < L; lvl++) {
    int h0, h1;
    level_range(lvl, h0, h1);
    int maxn = 0;
    for (int h = h0; h < h1; h++) maxn = std::max(maxn, P.sz[h]);
    int64_t gb = (int64_t)V.gather_tiles.size();
    for (int h = h0; h < h1; h++) {
      const int n = P.sz[h], r0 = (n + 1) / 2 * 2, nrows = S.rows[h] - r0;
      if (nrows <= 0 || n <= 0) continue;
      V.gather.push_back(GatherDesc{S.poff[h] + r0, map_off[h], S.ld[h], nrows, n, P.start[h]});
      for (int g = 0; g < (n + COLG - 1) / COLG; g++) V.gather_tiles.push_back(TileRef{(int)V.gather.size() - 1, (uint16_t)(g & 0xffff), (uint16_t)(g >> 16)});
    }
    if ((int64_t)V.gather_tiles.size() > gb) V.launches.push_back(SolveLaunch{SK_GATHER, gb, (int64_t)V.gather_tiles.size() - gb, lvl, 0});
    const int nblk = (maxn + NB - 1) / NB;
    for (int blk = nblk - 1; blk >= 0; blk--) {
      const int d0 = blk * NB;
      int64_t b = (int64_t)V.tiles.size();
      for (int h = h0; h < h1; h++) {
        int n = P.sz[h];
        if (n <= d0) continue;
        V.tiles.push_back(SolveTile{S.poff[h] + d0 + (int64_t)d0 * S.ld[h], S.ld[h], std::min(NB, n - d0), P.start[h] + d0, 0});
      }
      if ((int64_t)V.tiles.size() > b) V.launches.push_back(SolveLaunch{SK_TILE_B, b, (int64_t)V.tiles.size() - b, lvl, std::min(NB, maxn - d0)});
      if (d0 == 0) continue;
      int64_t tb = (int64_t)V.gemv_tiles.size();
      for (int h = h0; h < h1; h++) {
        int n = P.sz[h], ld = S.ld[h];
        if (n <= d0) continue;
        int dw = std::min(NB, n - d0);
        int g = add_gemv(SolveGemv{S.poff[h] + d0, ld, d0, dw, P.start[h] + d0, P.start[h], 0});
        for (int s = 0; s < (d0 + COLG - 1) / COLG; s++) V.gemv_tiles.push_back(TileRef{g, (uint16_t)(s & 0xffff), (uint16_t)(s >> 16)});
      }
      if ((int64_t)V.gemv_tiles.size() > tb) V.launches.push_back(SolveLaunch{SK_GEMV_B, tb, (int64_t)V.gemv_tiles.size() - tb, lvl, 0});
    }
  }
  return 0;
}

}  // namespace chb
