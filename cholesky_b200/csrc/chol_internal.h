// Internal data model of the engine: problem -> symbolic structure -> device schedule.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace chb {

// ----------------------------------------------------------------------------- inputs
// What the reference reads from its three files (mnd.c:22-199), as plain arrays.
struct Problem {
  int n = 0, ncols = 0;
  int64_t nz = 0;
  char typecode[4] = {'M', 'C', 'R', 'H'};
  std::vector<int32_t> ei, ej;  // 0-based, as given (row >= col in the reference's files)
  std::vector<double> ev;
  int levels = 0, N = 0;  // N = 2^levels - 1 separators
  // separators by 1-based heap index h (root = 1, label = N + 1 - h, file id = label - 1)
  std::vector<int> perm;       // permuted row -> original dof (ascending label == file order)
  std::vector<int> sz, start;  // [N + 2]
  std::vector<std::vector<std::vector<int>>> iv;  // iv[h][k]: raw interval list k of separator h
  int max_int_size = -1;       // as mnd.c:71-150 would return it
  int level_of(int h) const {
    int l = 0;
    while (h > 1) h >>= 1, l++;
    return l;
  }
  int label_of(int h) const { return N + 1 - h; }
  int heap_of(int label) const { return N + 1 - label; }
};

int read_problem(Problem &P, const char *mtx, const char *ord, const char *clust, std::string &err);
int finish_problem(Problem &P, std::string &err);  // start[], validation
int write_problem(const Problem &P, const char *mtx, const char *ord, const char *clust, std::string &err);
int generate_problem(Problem &P, int nx, int ny, int nz, int stencil, int levels, std::string &err);

// ----------------------------------------------------------------------------- symbolic
struct FilledRec {  // reference `Filled` (blas.rg:55-61)
  int64_t filled, sep_x, sep_y, interval, cluster, lo_x, lo_y, hi_x, hi_y;
};

// One filled row cluster of block (anc, s) at the time s is eliminated: rows [lo, hi) of
// separator `anc` (local dof positions), stored at row `off` of s's panel.
struct Seg {
  int anc;  // heap index of the row separator (== s for the diagonal block)
  int lo, hi;
  int off;
  int cluster;  // row-cluster index at the elimination interval
};

struct Symbolic {
  std::vector<std::vector<std::vector<int>>> cb;  // composed cluster boundaries cb[h][k][j]
  // supernodal panel of every separator: segs[seg_ptr[h] .. seg_ptr[h+1]); first = diagonal block
  std::vector<int64_t> seg_ptr;
  std::vector<Seg> segs;
  std::vector<int> rows;        // stored rows of the panel (even-aligned segment starts)
  std::vector<int> ld;          // leading dimension (>= rows, even)
  std::vector<int64_t> poff;    // panel offset in the factor buffer (doubles)
  int64_t total_doubles = 0;
  // pattern evidence
  std::vector<int64_t> nfilled;
  std::vector<uint64_t> checksum;
  std::vector<std::vector<FilledRec>> records;  // only with keep_records
  int64_t nblocks = 0, nclusters0 = 0;
  // algorithmic flops / call counts of the reference BLAS call list, per tree level
  std::vector<double> f_potrf, f_trsm, f_syrk, f_gemm;
  int64_t calls[4] = {0, 0, 0, 0};
  double flops() const {
    double s = 0;
    for (size_t i = 0; i < f_potrf.size(); i++) s += f_potrf[i] + f_trsm[i] + f_syrk[i] + f_gemm[i];
    return s;
  }
};

int analyze(const Problem &P, Symbolic &S, bool keep_records, std::string &err);
// the symbolic structure as one binary file (factor_file.cc): the ranks of a node analyse once and share the result
int save_symbolic(const Problem &P, const Symbolic &S, const char *path, std::string &err);
int load_symbolic(const Problem &P, Symbolic &S, const char *path, std::string &err);

// ----------------------------------------------------------------------------- schedule
// Everything the device executes is one of three grouped kernels over descriptor arrays.
// Offsets are in doubles from the factor buffer base, so the schedule is relocatable.
struct GemmProblem {  // C[M x N] -= sum_c A_c[M x K_c] * B_c[N x K_c]^T  (column-major)
  int64_t c_off;
  int ldc, M, N;
  int tri;  // 1: only i >= j is stored (SYRK on a diagonal cluster / in-panel trailing update)
  int contrib_begin, contrib_count;
};
struct GemmContrib {
  int64_t a_off, b_off;
  int lda, ldb, K, pad;
};
struct TileRef {  // one CTA
  int prob;
  uint16_t tr, tc;
};
// One block column (w <= 256 columns starting at column c0) of one panel, factored by ONE launch of panel_kernel:
// its w x w diagonal block (rows c0 .. c0 + w of the panel) by Cholesky, the rows below by X <- X L^-T.
// A CTA owns a slab of up to 64 rows for the whole block column; the CTAs that own the diagonal tiles publish
// them through flag words flag0 .. flag0 + 3 (one per 64-column tile step), the others wait for them.
// ready != 0: the diagonal block is already factored in memory (rows of a top panel owned by another rank, or
// the fused_dtrsm phase on its own), nobody waits.
struct PanelDesc {
  int64_t off;  // panel base in the factor buffer
  int ld, c0, w;
  int col0;     // permuted column of the block column's first column (for info reporting)
  int flag0, ready;
};
struct TrsmDesc {  // rows x nb slab below a factored tile: B <- B * L^-T (trsm_tile, 128-row slabs)
  int64_t l_off, b_off;
  int ld, nb, rows, pad;
};
struct PanelSlab {
  int desc;
  int row0, rows;  // stored rows [row0, row0 + rows) of the panel, rows <= 64
  int t;           // >= 0: this slab is rows c0 + 64 t .. of the diagonal block and factors diagonal tile t; -1: rows below
};

// One rectangle of a top panel: rows [r0, r0 + rows) x cols [c0, c0 + cols), first entry at factor offset `off`.
// K_PUSH copies it into the peers' copies of the factor (NVLink peer stores); K_REDUCE sums the peers' partial
// sums of it into this rank's copy (peer loads); `tri0` >= 0: rows and columns are the pivot block's, entries
// with column > tri0 + row (strictly above the diagonal) are skipped.
struct RectDesc {
  int64_t off;
  int ld, rows, cols, tri0;
  unsigned mask;  // K_REDUCE: the ranks whose partial sums of this rectangle are not identically zero (plus the owner)
  int pad;
};

enum LaunchKind { K_PANEL = 0, K_TRSM = 1, K_GEMM = 2, K_SYNC = 3, K_REDUCE = 4, K_NOP = 5, K_PUSH = 6 };
enum Phase { PH_POTRF = 1, PH_TRSM = 2, PH_UPDATE = 4 };  // which reference fused task the launch belongs to
enum FlagSlot { SLOT_WORLD = 0, SLOT_GROUP = 1, SLOT_DIAG = 2, kFlagSlots = 4 };
struct Launch {
  int kind;
  int level;
  int phase;
  int64_t begin, count;  // range in pslabs[] / trsm_tiles[] / tiles[] / rects[]
  double flops;          // executed flops (for per-kernel accounting), GEMM only
  int cfg;               // GEMM: 0 = 64x64 CTA tiles on shared-memory operand rings, 3 = one warp per 32x32 tile;
                         // K_PANEL: the widest block column of the launch (sizes its shared memory)
  int stream;            // 0 = update stream, 1 = chain stream (look-ahead), 2 = background pushes, 3 = rows stream (top panels)
  int wait_ev, rec_ev;   // event to wait for before / to record after the launch (-1: none)
  // multi-GPU (K_PUSH / K_SYNC / K_REDUCE): `mask` = ranks the rectangles are pushed to / reduced from;
  // after a push (or as the first half of a sync) flag word (slot, this rank) of every rank in `sig_mask`
  // is raised to `seq`; a sync then waits until its own words (slot, r) have reached `seq` for every r in
  // `wait_mask`.  seq values grow in program order on every (slot, source) pair.
  unsigned mask, sig_mask, wait_mask;
  int slot;
  int64_t seq;
};

constexpr int kRowBlock = 256;  // default and largest row block (CHOL_ROW_BLOCK: 64 / 128 / 256, so that tests deal small panels too)
struct Schedule {
  std::vector<GemmProblem> probs;
  std::vector<GemmContrib> contribs;
  std::vector<TileRef> tiles;
  std::vector<PanelDesc> pdesc;
  std::vector<PanelSlab> pslabs;
  std::vector<TrsmDesc> trsm;
  std::vector<TileRef> trsm_tiles;  // prob = index into trsm[], tr = slab index
  // The rows below the diagonal blocks of a launch go through panel_kernel's slabs (one launch, lowest latency)
  // when there are at most this many 64-row slabs, and through trsm_tile + grouped GEMM launches per 64-column
  // tile step (highest throughput) when there are more.  Measured on one B200 (128^3 / 64^3): always slabs
  // 958 / 24.2 ms, <= 296 slabs 906 / 23.2 ms, <= 64 slabs 903 / 23.0 ms.  On the top levels of a partition the
  // chain of a block column is the critical path of the whole group, so slabs are used up to two waves of CTAs.
  // (CHOL_FUSED_ROWS_MAX / CHOL_FUSED_ROWS_MAX_TOP)
  int fused_rows_max = 64, fused_rows_max_top = 296;
  std::vector<RectDesc> rects;      // rectangles of K_PUSH / K_REDUCE launches
  std::vector<Launch> launches;
  // assembly: value e of the input goes to factor[a_off[e]] (-1: dropped, mmat.rg:1191)
  std::vector<int64_t> a_off;
  int nb = 64, nbo = 256;
  // small fronts: Schur problems with M, N <= small_mn and total K <= small_k run as one warp per 32x32
  // tile (cfg 3, gemm_small_warp): no shared memory, no barriers (CHOL_SMALL_FRONT=0: off)
  bool small_front = true;
  int small_mn = 64, small_k = 256;  // (48..96, 128..512 measured within noise of each other on 128^3)
  // multi-GPU partition (world = 2^depth ranks)
  int rank = 0, world = 1, depth = 0;
  bool split_phases = false;  // true: fused_dpotrf and fused_dtrsm as separate launch sequences (piecewise API)
  bool lookahead = true;      // chain kernels on a second stream, overlapping the trailing updates (CHOL_LOOKAHEAD=0: off)
  int num_events = 0;         // cross-stream events the launch list refers to
  int64_t top_doubles = 0;    // leading part of the factor buffer that holds the top panels (one copy per rank)
  int row_block = kRowBlock;  // rows of a top panel are dealt to its group in blocks of this many (= block-column width there)
};

// Ownership of the rows of a top panel (multi-GPU): separator h on tree level lv < depth belongs to the
// 2^(depth - lv) ranks whose subtrees hang under it; its stored rows are dealt to them in blocks of 256,
// boustrophedon (0 1 .. G-1 G-1 .. 1 0), which balances both the triangular Schur updates and the shrinking
// trailing matrix.
struct TopGroup {
  int base, size, rb;
  unsigned mask() const { return ((1u << size) - 1u) << base; }
  int owner(int block) const {
    const int m = block % (2 * size);
    return base + (m < size ? m : 2 * size - 1 - m);
  }
};
inline TopGroup top_group(int h, int lv, int depth, int rb) {
  TopGroup g;
  g.rb = rb;
  g.size = 1 << (depth - lv);
  g.base = (h << (depth - lv)) - (1 << depth);
  return g;
}

// only_heap != 0: the launches of that one separator alone (no assembly map) -- the stepwise debug trace
int build_schedule(const Problem &P, const Symbolic &S, Schedule &D, int rank, int world, bool split_phases, std::string &err,
                   int only_heap = 0);

// ----------------------------------------------------------------------------- debug trace (`-d`)
// One fused task of the reference's level loop (mmat.rg:1240-1343) and the snapshot it is followed by.
struct DebugStep {
  int op;  // 0 POTRF, 1 TRSM, 2 GEMM (fused_dsyrk / fused_dgemm)
  int lvl;
  int hs, hp, hg;    // heap indices: eliminated separator, parent-side and grandparent-side ancestors
  std::string name;  // gen_filename (mmat.rg:149-172) without directory and extension
};
std::vector<DebugStep> debug_steps(const Problem &P);
// the whole `-d` log of one factorization: Block / Cluster / Fill lines of the symbolic phase and the
// POTRF / TRSM / GEMM lines of the level loop, in program order (needs Symbolic::records)
int write_debug_log(const Problem &P, const Symbolic &S, FILE *f, std::string &err);

// binary block dump of the factor (factor_file.cc); the converter back to text is chol_factor_binary_to_mtx
int write_factor_binary(const Problem &P, const Symbolic &S, const double *fac, int rank, int world, int depth, const char *path,
                        std::string &err);

// host-side parallel loop over [0, n) in contiguous chunks, one per worker (CHOL_HOST_THREADS, default: hardware
// concurrency capped at 16); fn(begin, end, worker).  Used where every item is independent and the result does
// not depend on the split (assembly map, interval-0 flags, Filled-record evidence).
void parallel_chunks(int64_t n, const std::function<void(int64_t, int64_t, int)> &fn, int *workers_out = nullptr);
int host_threads();

uint64_t mix64(uint64_t x);
uint64_t filled_hash(const FilledRec &r);

}  // namespace chb
