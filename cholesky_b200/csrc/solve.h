// Triangular solve on the panel storage (reference: mmat.rg:1364-1495, dtrsv/dgemv blas.rg:217-290).
// Level-scheduled like the factorization: forward leaves -> root, backward root -> leaves.  Updates of
// ancestor entries are destination-owned (every ancestor row cluster pulls from its contributors in a
// fixed order), so the solve is atomic-free and deterministic.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "chol_internal.h"

namespace chb {

struct SolveTile {  // nb x nb pivot tile: forward L y = b, backward L^T x = y, in place on x[x0 .. x0+nb)
  int64_t l_off;
  int ld, nb, x0, pad;
};
struct SolveGemv {
  // forward : y[y0 + r] -= sum_{c < nb} P[r, c] x[x0 + c],   r < rows   (P at p_off, column-major, ld)
  // backward: y[y0 + c] -= sum_{r < nb} P[r, c] x[x0 + r],   c < rows
  int64_t p_off;
  int ld, rows, nb, x0, y0, pad;
};
struct PullDest {  // x[y0 + r] -= sum over contributors of P_s[seg rows, :] x_s
  int y0, rows, cbeg, ccnt;
};
struct PullContrib {
  int64_t p_off;
  int ld, K, x0, pad;
};
// One CTA of a pull: 128 rows of a destination, columns [kbeg, kend) of the destination's contributors laid end to
// end.  Destinations with many contributor columns (the top of the tree: a few thousand rows against K of several
// thousand) are split over CTAs so that the sweep has enough loads in flight; a split tile parks its partial sums
// in scratch slot `slot` and a second launch adds the slots up in order (deterministic), slot < 0: the tile holds
// all columns and subtracts from x directly.
struct PullTile {
  int dest, slab, kbeg, kend, slot;
};
struct PullSum {  // x[y0 + r] -= sum_{p < nparts} scratch[(slot0 + p) * 128 + r], r < rows
  int y0, rows, slot0, nparts;
};
struct GatherDesc {  // x[x0 + c] -= sum_{r < nrows} P[r, c] x[rowmap[map_off + r]]
  int64_t p_off, map_off;
  int ld, nrows, n, x0;
};
enum SolveKind { SK_TILE_F = 0, SK_GEMV_F = 1, SK_PULL = 2, SK_GATHER = 3, SK_TILE_B = 4, SK_GEMV_B = 5, SK_EXCHANGE = 6, SK_PULL_SUM = 7 };
struct SolveLaunch {
  int kind;
  int64_t begin, count;  // range in tiles_f / gemv_tiles / pull_tiles / gather_tiles / tiles_b
  int level;             // tree level the launch works on
  int width;             // SK_TILE_*: the widest diagonal block of the launch (<= 64: the light tile kernel)
};
struct SolveSchedule {
  std::vector<SolveTile> tiles;       // indexed directly by SK_TILE_* launches
  std::vector<SolveGemv> gemv;
  std::vector<TileRef> gemv_tiles;    // (gemv desc, slab)
  std::vector<PullDest> pull;
  std::vector<PullContrib> pull_contrib;
  std::vector<PullTile> pull_tiles;   // (dest, slab, column range)
  std::vector<PullSum> pull_sums;
  int64_t pull_slots = 0;             // scratch slots (128 doubles each) the largest pull launch needs
  std::vector<GatherDesc> gather;
  std::vector<TileRef> gather_tiles;  // (gather desc, column group)
  std::vector<int> rowmap;            // global permuted row of every stored off-diagonal panel row (-1: padding)
  std::vector<SolveLaunch> launches;  // partitioned handles: one SK_EXCHANGE splits the forward sweep (solve.cc)
  int rank = 0, world = 1, depth = 0;
  int top_row0 = 0;  // first permuted row of the shared top separators (== n on a single-GPU handle)
};

int build_solve(const Problem &P, const Symbolic &S, SolveSchedule &V, int rank, int world, std::string &err);

}  // namespace chb
