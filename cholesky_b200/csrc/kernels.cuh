// sm_100a kernels of the numeric factorization.  A handful of grouped kernels execute the whole level
// schedule (see schedule.cc); one scatter kernel assembles A; three small kernels carry the multi-GPU
// exchange over NVLink peer memory.
//
//   gemm_grouped_ws  C -= sum_c A_c B_c^T on FP64 tensor cores (mma.sync m8n8k4 -> SASS DMMA.8x8x4),
//                  warp-specialised: a producer warp stages operand tiles with the TMA bulk-copy engine
//                  (cp.async.bulk + mbarrier expect_tx), consumer warps issue the MMAs; one CTA per
//                  destination tile, contributors accumulated in registers in a fixed order
//                  (atomic-free, deterministic), lower-triangle masking for SYRK destinations.
//                  Replaces cblas_dgemm / cblas_dsyrk as called at blas.rg:139-142, 187-189.
//   gemm_small_warp  the same update for small fronts: one warp per 32x32 tile, no shared memory.
//   potrf_tile     Cholesky of one NB x NB pivot tile, register-resident and right-looking
//                  (LAPACKE_dpotrf, blas.rg:71).
//   trsm_tile      one 128-row slab times L^-T by forward substitution, one row per thread
//                  (cblas_dtrsm Right/Lower/Trans/NonUnit, blas.rg:99-100).
//   assemble       factor[a_off[e]] = value[e]   (fill_block, mmat.rg:529-633).
//   push_rects / reduce_rects / peer_sync   rows of the top panels pushed to the peers' copies, partial
//                  sums of the rows a rank owns pulled from its group, flag words in peer memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chol_internal.h"

namespace chb {

constexpr int kMaxPeers = 8;
struct Peers {
  double *fac[kMaxPeers];
  unsigned long long *flags[kMaxPeers];  // kFlagSlots x kMaxPeers words per rank: word (slot, source rank)
  int n, rank;
};

// D(8x8) += A(8x4, row) * B(4x8, col); lane = 4*g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// gemm_grouped_ws: the same grouped update, warp-specialised.  One producer warp stages operand
// tiles with the TMA bulk-copy engine (cp.async.bulk global -> shared, SASS UBLKCP; every column of
// a column-major tile is one contiguous 16-byte aligned run, so no tensor map is needed) and signals
// the consumers through mbarriers (expect_tx / complete_tx); WM x WN consumer warps issue FP64 DMMA
// and hand stages back through a second set of mbarriers.  No block-wide barrier in the main loop,
// the stage ring keeps running across the contributors of a destination tile (K-concatenation).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES>
struct GemmWsCfg {
  static constexpr int kWarpsM = BM / WM, kWarpsN = BN / WN;
  static constexpr int kConsumers = kWarpsM * kWarpsN;
  static constexpr int kThreads = (kConsumers + 1) * 32;
  static constexpr int kPad = 4;
  static constexpr int kLdA = BM + kPad, kLdB = BN + kPad;
  static constexpr int kStageDoubles = BK * (kLdA + kLdB);
  static constexpr int kSmemBytes = STAGES * kStageDoubles * 8 + 2 * STAGES * 8 + STAGES * 4 + 64;
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB>
__global__ void __launch_bounds__(GemmWsCfg<BM, BN, BK, WM, WN, STAGES>::kThreads, MINB)
    gemm_grouped_ws(const GemmProblem *__restrict__ probs, const GemmContrib *__restrict__ contribs,
                    const TileRef *__restrict__ tiles, double *__restrict__ fac) {
  using Cfg = GemmWsCfg<BM, BN, BK, WM, WN, STAGES>;
  static_assert(BK * 2 <= 32 || BK == 32, "one bulk copy per producer lane");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *smem = reinterpret_cast<double *>(smem_raw);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + STAGES * Cfg::kStageDoubles * 8);
  unsigned long long *empty = full + STAGES;
  int *kvalid = reinterpret_cast<int *>(empty + STAGES);

  const TileRef tile = tiles[blockIdx.x];
  const GemmProblem pr = probs[tile.prob];
  const int row0 = tile.tr * BM, col0 = tile.tc * BN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mrem = pr.M - row0, nrem = pr.N - col0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], Cfg::kConsumers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == Cfg::kConsumers) {
    // ===== producer warp: lane l < BK copies column l of the A tile, lane BK + l column l of the B tile
    const unsigned a_bytes = (unsigned)(min(BM, (mrem + 1) & ~1) * 8), b_bytes = (unsigned)(min(BN, (nrem + 1) & ~1) * 8);
    // The epilogue is a read-modify-write of a destination tile that sits in HBM (the trailing matrix is
    // far larger than L2): the producer pulls the tile's lines into L2 kPrefetchLead stages before the
    // consumers get there.
    constexpr int kPrefetchLead = 48;
    int total = 0;
    for (int c = 0; c < pr.contrib_count; c++) total += (contribs[pr.contrib_begin + c].K + BK - 1) / BK;
    const int pf_at = max(0, total - kPrefetchLead);
    int it = 0;
    for (int c = 0; c < pr.contrib_count; c++) {
      const GemmContrib cb = contribs[pr.contrib_begin + c];
      const double *__restrict__ A = fac + cb.a_off + row0;
      const double *__restrict__ Bp = fac + cb.b_off + col0;
      for (int k0 = 0; k0 < cb.K; k0 += BK, it++) {
        if (it == pf_at) {
          const int rows = min(BM, mrem);
          for (int j = lane; j < min(BN, nrem); j += 32) {
            const char *p0 = reinterpret_cast<const char *>(fac + pr.c_off + row0 + (size_t)(col0 + j) * pr.ldc);
            const char *p1 = p0 + (size_t)rows * 8;
            for (const char *q = reinterpret_cast<const char *>(reinterpret_cast<uintptr_t>(p0) & ~(uintptr_t)127); q < p1; q += 128)
              asm volatile("prefetch.global.L2 [%0];\n" ::"l"(q));
          }
        }
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
        const int kv = min(BK, cb.K - k0);
        double *As = smem + s * Cfg::kStageDoubles;
        double *Bs = As + BK * Cfg::kLdA;
        if (lane == 0) {
          kvalid[s] = kv;
          mbar_arrive_expect_tx(&full[s], (unsigned)kv * (a_bytes + b_bytes));
        }
        __syncwarp();
        if (BK <= 16) {
          const int kk = lane & (BK - 1);
          if (kk < kv) {
            if (lane < BK) bulk_g2s(As + kk * Cfg::kLdA, A + (size_t)(k0 + kk) * cb.lda, a_bytes, &full[s]);
            else if (lane < 2 * BK) bulk_g2s(Bs + kk * Cfg::kLdB, Bp + (size_t)(k0 + kk) * cb.ldb, b_bytes, &full[s]);
          }
        } else {
          if (lane < kv) {
            bulk_g2s(As + lane * Cfg::kLdA, A + (size_t)(k0 + lane) * cb.lda, a_bytes, &full[s]);
            bulk_g2s(Bs + lane * Cfg::kLdB, Bp + (size_t)(k0 + lane) * cb.ldb, b_bytes, &full[s]);
          }
        }
      }
    }
    return;
  }

  // ===== consumer warps
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % Cfg::kWarpsM) * WM, wn0 = (warp / Cfg::kWarpsM) * WN;
  constexpr int MB = WM / 8, NBk = WN / 8;
  double acc[MB][NBk][2];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NBk; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  int total = 0;
  for (int c = 0; c < pr.contrib_count; c++) total += (contribs[pr.contrib_begin + c].K + BK - 1) / BK;
  for (int it = 0; it < total; it++) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    const int kv = kvalid[s];
    const double *As = smem + s * Cfg::kStageDoubles;
    const double *Bs = As + BK * Cfg::kLdA;
    if (kv == BK) {
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; k4++) {
        double a[MB], b[NBk];
#pragma unroll
        for (int i = 0; i < MB; i++) a[i] = As[(k4 * 4 + t) * Cfg::kLdA + wm0 + i * 8 + g];
#pragma unroll
        for (int j = 0; j < NBk; j++) b[j] = Bs[(k4 * 4 + t) * Cfg::kLdB + wn0 + j * 8 + g];
#pragma unroll
        for (int i = 0; i < MB; i++)
#pragma unroll
          for (int j = 0; j < NBk; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    } else {  // K tail of a contributor: columns >= kv of the stage are stale, feed zeros instead
      for (int k4 = 0; k4 * 4 < kv; k4++) {
        const bool ok = k4 * 4 + t < kv;
        double a[MB], b[NBk];
#pragma unroll
        for (int i = 0; i < MB; i++) a[i] = ok ? As[(k4 * 4 + t) * Cfg::kLdA + wm0 + i * 8 + g] : 0.0;
#pragma unroll
        for (int j = 0; j < NBk; j++) b[j] = ok ? Bs[(k4 * 4 + t) * Cfg::kLdB + wn0 + j * 8 + g] : 0.0;
#pragma unroll
        for (int i = 0; i < MB; i++)
#pragma unroll
          for (int j = 0; j < NBk; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  double *__restrict__ C = fac + pr.c_off;
  // read-modify-write of the destination straight from the accumulator fragments.  All loads of one row
  // block are issued before the first store: written as `C[..] -= acc` the compiler has to keep every store
  // ahead of the next load (same array), i.e. 32 dependent round trips to L2/HBM per thread -- ncu's source
  // view had 40 % of the warp samples of the K = 256 launches on those DADDs, and the DMMA pipe at 82 %.
  const int rlo = (pr.tri & 2) ? 1 : 0;  // an ownership boundary on an odd row: row 0 belongs to the neighbour
  const bool tri = pr.tri & 1;
#pragma unroll
  for (int i = 0; i < MB; i++) {
    const int r = row0 + wm0 + i * 8 + g;
    double cv[NBk][2];
    bool ok[NBk][2];
#pragma unroll
    for (int j = 0; j < NBk; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int cc = col0 + wn0 + j * 8 + 2 * t + e;
        ok[j][e] = r >= rlo && r < pr.M && cc < pr.N && (!tri || r >= cc);
        cv[j][e] = ok[j][e] ? C[r + (size_t)cc * pr.ldc] : 0.0;
      }
#pragma unroll
    for (int j = 0; j < NBk; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int cc = col0 + wn0 + j * 8 + 2 * t + e;
        if (ok[j][e]) C[r + (size_t)cc * pr.ldc] = cv[j][e] - acc[i][j][e];
      }
  }
}

// ---------------------------------------------------------------------------------------------
// gemm_small_warp: the same grouped update for the small fronts at the bottom of the tree.  One warp
// per 32x32 destination tile, register-blocked (16 DMMA accumulator blocks), operands read straight
// from global memory in the DMMA fragment pattern (eight consecutive rows per column: 64-byte runs,
// L2-resident), no shared memory and no barriers, so thousands of tiny problems run per launch without
// per-CTA pipeline set-up.  Contributors are still accumulated in their fixed order.
constexpr int kSmallWarps = 4;
__global__ void __launch_bounds__(kSmallWarps * 32) gemm_small_warp(const GemmProblem *__restrict__ probs,
                                                                    const GemmContrib *__restrict__ contribs,
                                                                    const TileRef *__restrict__ tiles, int64_t ntiles,
                                                                    double *__restrict__ fac) {
  const int64_t ti = (int64_t)blockIdx.x * kSmallWarps + (threadIdx.x >> 5);
  if (ti >= ntiles) return;
  const TileRef tile = tiles[ti];
  const GemmProblem pr = probs[tile.prob];
  const int row0 = tile.tr * 32, col0 = tile.tc * 32;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  bool rok[4], cok[4];
#pragma unroll
  for (int i = 0; i < 4; i++) rok[i] = row0 + i * 8 + g < pr.M, cok[i] = col0 + i * 8 + g < pr.N;
  for (int c = 0; c < pr.contrib_count; c++) {
    const GemmContrib cb = contribs[pr.contrib_begin + c];
    const double *__restrict__ A = fac + cb.a_off + row0 + g;
    const double *__restrict__ Bp = fac + cb.b_off + col0 + g;
    const int K = cb.K;
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 4) {
      const bool kok = k0 + t < K;
      const size_t ka = (size_t)(k0 + t) * cb.lda, kb = (size_t)(k0 + t) * cb.ldb;
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = (kok && rok[i]) ? A[ka + i * 8] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = (kok && cok[j]) ? Bp[kb + j * 8] : 0.0;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  // read-modify-write of the destination: the eight loads of a row block go out together, then the stores
  // (`C[..] -= acc` would chain 32 load/store round trips, see gemm_grouped_ws)
  double *__restrict__ C = fac + pr.c_off;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = row0 + i * 8 + g;
    double cv[4][2];
    bool ok[4][2];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int cc = col0 + j * 8 + 2 * t + e;
        ok[j][e] = r >= ((pr.tri >> 1) & 1) && r < pr.M && cc < pr.N && (!(pr.tri & 1) || r >= cc);
        cv[j][e] = ok[j][e] ? C[r + (size_t)cc * pr.ldc] : 0.0;
      }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int cc = col0 + j * 8 + 2 * t + e;
        if (ok[j][e]) C[r + (size_t)cc * pr.ldc] = cv[j][e] - acc[i][j][e];
      }
  }
}

// ---------------------------------------------------------------------------------------------
// panel_kernel: one 256-wide block column of every panel of a tree level in ONE launch -- LAPACKE_dpotrf of its
// diagonal block (blas.rg:71) and cblas_dtrsm Right/Lower/Trans/NonUnit of the rows below (blas.rg:99-100).
//
// A CTA owns a slab of up to 64 rows for the whole block column and keeps it in shared memory (column-major, row
// stride 68: conflict-free FP64 MMA fragment loads).  Tile step j = columns 64 j .. 64 j + 63, left-looking:
//   1. columns j of the slab -= sum_{i<j} (columns i of the slab) x L(j, i)^T   -- 64 x 64 x 64 products on the
//      tensor cores (DMMA), accumulated in registers over i;
//   2. the slab that holds diagonal tile j factors it in shared memory (potrf_smem), stores its rows and raises flag
//      j of the block column; every other slab waits for that flag, loads L(j, j) and solves its 64 x 64 tile
//      (one row per thread, eight independent FMA chains).
// So the dependent chain of a block column -- four tile factorizations with a solve and a rank-64 update of the
// next diagonal tile in between -- runs inside one kernel through flag words in L2 instead of through eleven
// stream-ordered launches (measured at the root of 128^3: potrf_tile 26 us + trsm_tile 17 us + K = 64 GEMM 18 us per
// tile step, launch latencies included).  Diagonal slabs come first in the grid, in tile order, so a waiting CTA
// only ever waits for CTAs that were scheduled before it.
constexpr int kNB = 64;
constexpr int kPanelThreads = 256;
constexpr int kLdx = kNB + 4;
constexpr long long kSpinLimit = 16000000000LL;  // ~8 s at 2 GHz: a wait gives up and raises an error code instead of hanging the GPU
inline size_t panel_smem_bytes(int wmax) { return ((size_t)((wmax + kNB - 1) / kNB * kNB) * kLdx + (size_t)kNB * kLdx + kNB) * sizeof(double); }

// Cholesky of the 64 x 64 tile T (entry (i, c) at T[c * kLdx + i]) in shared memory, right-looking in panels of
// eight columns, with look-ahead inside the tile.  The 64 dependent column steps are what a tile costs, so the
// warp that does them never waits for anything else:
//   warp 0       holds the current panel in registers (two rows per lane); a column step is a shuffle of the
//                pivot and of the panel's raw pivot-row entries, rsqrt, and the rank-1 update of the remaining
//                panel columns (B200, tools/lat_bench.cu: rsqrt(double) 74 cycles, dependent DFMA 8).  After its
//                eight steps it stores the panel, applies it to the NEXT panel's columns itself (registers) and
//                goes on;
//   warps 1 - 7  apply every finished panel to the columns two panels and more to the right, one panel behind
//                warp 0 (named barriers A_p: panel p is in shared memory, B_p: its trailing update is done).
// Measured per tile in tools/potrf_bench.cu.  Columns >= dw are an identity block.  Ends with a block barrier.
// named barriers 1 .. 4 over the whole CTA (immediate ids: a register id would make ptxas reserve all sixteen
// barriers of the SM for one CTA); odd selects the second of a pair
template <int ID>
__device__ __forceinline__ void bar_sync_pair(bool odd) {
  if (odd) asm volatile("bar.sync %0, %1;\n" ::"n"(ID + 1), "n"(kPanelThreads) : "memory");
  else asm volatile("bar.sync %0, %1;\n" ::"n"(ID), "n"(kPanelThreads) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_arrive_pair(bool odd) {
  if (odd) asm volatile("bar.arrive %0, %1;\n" ::"n"(ID + 1), "n"(kPanelThreads) : "memory");
  else asm volatile("bar.arrive %0, %1;\n" ::"n"(ID), "n"(kPanelThreads) : "memory");
}
__device__ __forceinline__ void potrf_smem(double *__restrict__ T, int dw, int col0, int *__restrict__ info) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = (dw + 7) >> 3;  // panels; panel p has trailing columns (two panels and more to the right) iff p + 2 < np
  if (warp == 0) {
    double x0[8], x1[8];  // rows lane and lane + 32 of the current panel
#pragma unroll
    for (int j = 0; j < 8; j++) x0[j] = T[j * kLdx + lane], x1[j] = T[j * kLdx + lane + 32];
    for (int p = 0; p < np; p++) {
      const int k0 = 8 * p;
      const bool hi = k0 >= 32;  // the panel's own rows k0 .. k0 + 7 sit in x1 (of lanes k0 - 32 ..) or in x0
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
        double dk = __shfl_sync(0xffffffffu, hi ? x1[kk] : x0[kk], (k0 + kk) & 31);
        double raw[8];  // column kk of the panel's own rows below the pivot, before scaling: independent of the rsqrt
#pragma unroll
        for (int j = kk + 1; j < 8; j++) raw[j] = __shfl_sync(0xffffffffu, hi ? x1[kk] : x0[kk], (k0 + j) & 31);
        if (!(dk > 0.0)) {
          if (lane == 0) atomicMin(info, col0 + k0 + kk + 1);  // 1-based permuted column of the first bad pivot
          dk = 1.0;
        }
        const double r = rsqrt(dk);
        x0[kk] *= r, x1[kk] *= r;  // the pivot row's own entry becomes dk / sqrt(dk); rows above it hold garbage that is never stored
#pragma unroll
        for (int j = kk + 1; j < 8; j++) {
          const double ljk = raw[j] * r;
          x0[j] = fma(-x0[kk], ljk, x0[j]);
          x1[j] = fma(-x1[kk], ljk, x1[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (lane >= k0 + j) T[(k0 + j) * kLdx + lane] = x0[j];
        if (lane + 32 >= k0 + j) T[(k0 + j) * kLdx + lane + 32] = x1[j];
      }
      __syncwarp();
      if (p + 2 < np) bar_arrive_pair<1>(p & 1);  // A_p
      if (p + 1 < np) {
        if (p >= 1) bar_sync_pair<3>((p - 1) & 1);  // B_{p-1}: the next panel's columns have every update but this panel's
        double y0[8], y1[8];
#pragma unroll
        for (int j = 0; j < 8; j++) y0[j] = T[(k0 + 8 + j) * kLdx + lane], y1[j] = T[(k0 + 8 + j) * kLdx + lane + 32];
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const double ljk = T[(k0 + k) * kLdx + k0 + 8 + j];  // entry (row k0 + 8 + j, column k0 + k) of the panel just stored
            y0[j] = fma(-x0[k], ljk, y0[j]);
            y1[j] = fma(-x1[k], ljk, y1[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) x0[j] = y0[j], x1[j] = y1[j];
      }
    }
  } else {
    for (int p = 0; p + 2 < np; p++) {
      const int k0 = 8 * p, e0 = k0 + 16;
      bar_sync_pair<1>(p & 1);  // A_p
      double r0[8], r1[8];
#pragma unroll
      for (int k = 0; k < 8; k++) r0[k] = T[(k0 + k) * kLdx + lane], r1[k] = T[(k0 + k) * kLdx + lane + 32];
      for (int j = e0 + warp - 1; j < 8 * np; j += 7) {  // T(i, j) -= sum_k T(i, k) T(j, k), i >= j
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const double ljk = T[(k0 + k) * kLdx + j];
          s0 = fma(r0[k], ljk, s0);
          s1 = fma(r1[k], ljk, s1);
        }
        if (lane >= j) T[j * kLdx + lane] -= s0;
        if (lane + 32 >= j) T[j * kLdx + lane + 32] -= s1;
      }
      __syncwarp();
      bar_arrive_pair<3>(p & 1);  // B_p
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kPanelThreads) panel_kernel(const PanelDesc *__restrict__ descs, const PanelSlab *__restrict__ slabs,
                                                              double *__restrict__ fac, int *__restrict__ pflags, int wcols, int *__restrict__ info) {
  extern __shared__ __align__(16) double psm[];
  double *__restrict__ Xs = psm;                              // the slab: entry (r, c) at Xs[c * kLdx + r]
  double *__restrict__ Ls = psm + (size_t)wcols * kLdx;       // one tile of L, k-major: Ls[k * kLdx + n] = L(n, k)
  double *__restrict__ rdiag = Ls + kNB * kLdx;               // 1 / L(n, n) of the diagonal tile in Ls
  const PanelSlab sl = slabs[blockIdx.x];
  const PanelDesc d = descs[sl.desc];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncol = sl.t >= 0 ? min(d.w, kNB * (sl.t + 1)) : d.w;  // a diagonal slab ends with its own tile
  const int nt = (ncol + kNB - 1) / kNB;
  double *__restrict__ G = fac + d.off + sl.row0 + (size_t)d.c0 * d.ld;         // entry (r, c) of the slab
  const double *__restrict__ Gd = fac + d.off + d.c0 + (size_t)d.c0 * d.ld;     // entry (n, k) of the diagonal block
  {  // load: thread = row, every fourth column, eight loads in flight; rows past the slab are zeros and the rows and
     // columns past a partial diagonal tile an identity block
    const int r = tid & (kNB - 1), cq = tid >> 6, cload = nt * kNB;
    for (int cb = cq; cb < cload; cb += 32) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int c = cb + 4 * u;
        v[u] = (c < ncol && r < sl.rows) ? G[r + (size_t)c * d.ld] : ((sl.t >= 0 && c >= ncol && r == c - kNB * sl.t) ? 1.0 : 0.0);
      }
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (cb + 4 * u < cload) Xs[(cb + 4 * u) * kLdx + r] = v[u];
    }
  }
  __syncthreads();
  const int wm = warp & 1, wn = warp >> 1;  // warp tile of a 64 x 64 update: rows 32 wm .., columns 16 wn ..
  const int g = lane >> 2, t4 = lane & 3;
  for (int j = 0; j < nt; j++) {
    const int d0 = j * kNB, dw = min(kNB, d.w - d0);
    const bool mine = sl.t == j;
    if (!mine && !d.ready) {  // the tiles L(j, 0 .. j) are final once the slab of diagonal tile j has raised its flag
      if (tid == 0) {
        const int *f = pflags + d.flag0 + j;
        const long long t0 = clock64();
        int v;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(f) : "memory");
          if (v == 0 && clock64() - t0 > kSpinLimit) {
            atomicExch(info + 1, 2);
            break;
          }
        } while (v == 0);
      }
      __syncthreads();
    }
    if (j > 0) {
      double acc[4][2][2];
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 2; ni++)
#pragma unroll
          for (int e = 0; e < 2; e++) acc[mi][ni][e] = Xs[(d0 + 16 * wn + 8 * ni + 2 * t4 + e) * kLdx + 32 * wm + 8 * mi + g];
      for (int i = 0; i < j; i++) {
        const double *__restrict__ Bp = Xs + (size_t)(i * kNB) * kLdx;  // the slab of a diagonal tile: L(j, i) is its own rows
        if (!mine) {
          __syncthreads();  // the previous tile in Ls has been consumed
          for (int idx = tid; idx < kNB * kNB; idx += kPanelThreads) {
            const int n = idx & (kNB - 1), k = idx >> 6;
            Ls[k * kLdx + n] = n < dw ? Gd[(d0 + n) + (size_t)(i * kNB + k) * d.ld] : 0.0;
          }
          __syncthreads();
          Bp = Ls;
        }
        const double *__restrict__ Ap = Xs + (size_t)(i * kNB) * kLdx;
#pragma unroll 4
        for (int k4 = 0; k4 < kNB / 4; k4++) {
          double a[4], b[2];
#pragma unroll
          for (int mi = 0; mi < 4; mi++) a[mi] = -Ap[(4 * k4 + t4) * kLdx + 32 * wm + 8 * mi + g];
#pragma unroll
          for (int ni = 0; ni < 2; ni++) b[ni] = Bp[(4 * k4 + t4) * kLdx + 16 * wn + 8 * ni + g];
#pragma unroll
          for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
      }
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 2; ni++)
#pragma unroll
          for (int e = 0; e < 2; e++) Xs[(d0 + 16 * wn + 8 * ni + 2 * t4 + e) * kLdx + 32 * wm + 8 * mi + g] = acc[mi][ni][e];
      __syncthreads();
    }
    if (mine) {
      potrf_smem(Xs + (size_t)d0 * kLdx, dw, d.col0 + d0, info);
      // this slab is done: its tiles left of the diagonal in full, the lower part of the diagonal tile
      const int r = tid & (kNB - 1);
      if (r < sl.rows)
        for (int c = tid >> 6; c < ncol; c += kPanelThreads / kNB)
          if (c < d0 || r >= c - d0) G[r + (size_t)c * d.ld] = Xs[c * kLdx + r];
      __threadfence();
      __syncthreads();
      if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(pflags + d.flag0 + j), "r"(1) : "memory");
      return;
    }
    // ---- rows below diagonal tile j: X <- X L(j, j)^-T
    for (int idx = tid; idx < kNB * kNB; idx += kPanelThreads) {
      const int n = idx & (kNB - 1), k = idx >> 6;
      const double v = (n < dw && k <= n) ? Gd[(d0 + n) + (size_t)(d0 + k) * d.ld] : ((n == k) ? 1.0 : 0.0);
      Ls[k * kLdx + n] = v;
      if (n == k) rdiag[n] = 1.0 / v;
    }
    __syncthreads();
    if (tid < kNB) {  // one row per thread, eight columns at a time with eight independent FMA chains
      double *__restrict__ xr = Xs + (size_t)d0 * kLdx + tid;
      for (int cb = 0; cb < dw; cb += 8) {
        double sv[8];
#pragma unroll
        for (int q = 0; q < 8; q++) sv[q] = xr[(cb + q) * kLdx];
#pragma unroll 4
        for (int k = 0; k < cb; k++) {
          const double xk = xr[k * kLdx];
          const double2 *__restrict__ l2 = reinterpret_cast<const double2 *>(&Ls[k * kLdx + cb]);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const double2 l = l2[q];
            sv[2 * q] = fma(-xk, l.x, sv[2 * q]);
            sv[2 * q + 1] = fma(-xk, l.y, sv[2 * q + 1]);
          }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
#pragma unroll
          for (int qq = 0; qq < q; qq++) sv[q] = fma(-sv[qq], Ls[(cb + qq) * kLdx + cb + q], sv[q]);
          sv[q] *= rdiag[cb + q];
        }
#pragma unroll
        for (int q = 0; q < 8; q++) xr[(cb + q) * kLdx] = sv[q];
      }
    }
    __syncthreads();
  }
  {  // rows below the diagonal block: store the whole slab
    const int r = tid & (kNB - 1);
    if (r < sl.rows)
      for (int c = tid >> 6; c < d.w; c += kPanelThreads / kNB) G[r + (size_t)c * d.ld] = Xs[c * kLdx + r];
  }
}

// trsm_tile: the rows below a factored diagonal tile when there are MANY of them (more slabs than a wave or two of
// CTAs): then throughput counts, not latency, and the rank-64 updates between tile steps run on the tuned grouped
// GEMM instead of inside panel_kernel's slabs (schedule.cc decides per launch).
// 128-row slab per CTA, one row per thread.  The slab (k-major, so a warp reads consecutive words) and
// L^T live in shared memory; columns are solved eight at a time with eight independent FMA chains, the
// eight multipliers of one k come as four broadcast vector loads.  Loops are deliberately not fully
// unrolled: the straight-line version was instruction-fetch bound.
constexpr int kSlab = 128;
constexpr int kTrsmSmemBytes = (kNB * kNB + kNB + kNB * kSlab) * 8;
// All 64 column loads of a slab row are in flight at once (measured: trsm time of 64^3 4.04 -> 3.50 ms against
// eight rounds of eight).
__global__ void __launch_bounds__(kSlab) trsm_tile(const TrsmDesc *__restrict__ descs, const TileRef *__restrict__ tiles,
                                                   double *__restrict__ fac) {
  extern __shared__ __align__(16) double tsm[];
  double(*Lt)[kNB] = reinterpret_cast<double(*)[kNB]>(tsm);             // Lt[k][c] = L[c][k]
  double *rdiag = tsm + kNB * kNB;                                       // 1 / L[c][c]
  double(*xs)[kSlab] = reinterpret_cast<double(*)[kSlab]>(rdiag + kNB);  // xs[c][row in slab]
  const TileRef tl = tiles[blockIdx.x];
  const TrsmDesc d = descs[tl.prob];
  const int slab = (int)tl.tr | ((int)tl.tc << 16);
  const int tid = threadIdx.x, nb = d.nb, nb8 = (nb + 7) & ~7;
  const double *__restrict__ Lg = fac + d.l_off;
  const int row = slab * kSlab + tid;
  const bool live = row < d.rows;
  double *__restrict__ Bp = fac + d.b_off + (live ? row : 0);
  {
    double v[kNB];
#pragma unroll
    for (int c = 0; c < kNB; c++) v[c] = (live && c < nb) ? Bp[(size_t)c * d.ld] : 0.0;
#pragma unroll
    for (int c = 0; c < kNB; c++) xs[c][tid] = v[c];
  }
  {
    constexpr int PER = kNB * kNB / kSlab;
    double v[PER];
#pragma unroll
    for (int u = 0; u < PER; u++) {
      int i = tid + u * kSlab, r = i % kNB, c = i / kNB;
      v[u] = (r < nb && c < nb && r >= c) ? Lg[r + (size_t)c * d.ld] : ((r == c) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < PER; u++) {
      int i = tid + u * kSlab, r = i % kNB, c = i / kNB;
      Lt[c][r] = v[u];
      if (r == c) rdiag[r] = 1.0 / v[u];
    }
  }
  __syncthreads();
  if (!live) return;
  for (int cb = 0; cb < nb8; cb += 8) {
    double s[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = xs[cb + j][tid];
#pragma unroll 4
    for (int k = 0; k < cb; k++) {
      const double xk = xs[k][tid];
      const double2 *l2 = reinterpret_cast<const double2 *>(&Lt[k][cb]);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double2 l = l2[j];
        s[2 * j] -= xk * l.x;
        s[2 * j + 1] -= xk * l.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
      for (int jj = 0; jj < j; jj++) s[j] -= s[jj] * Lt[cb + jj][cb + j];
      s[j] *= rdiag[cb + j];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) xs[cb + j][tid] = s[j];
  }
  for (int c0 = 0; c0 < nb; c0 += 8) {
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (c0 + u < nb) Bp[(size_t)(c0 + u) * d.ld] = xs[c0 + u][tid];
  }
}


__global__ void assemble_kernel(const double *__restrict__ vals, const int64_t *__restrict__ offs, int64_t nz, double *__restrict__ fac) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nz) {
    int64_t o = offs[e];
    if (o >= 0) fac[o] = vals[e];
  }
}

__global__ void gather_diag_kernel(const int64_t *__restrict__ diag_off, int n, const double *__restrict__ fac, double *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fac[diag_off[i]];
}


// ---------------------------------------------------------------------------------------------
// Multi-GPU exchange over NVLink peer memory.  Every rank maps every peer's factor buffer and flag words
// (CUDA IPC between processes, plain peer access inside one process).  Flag word (slot, s) of rank r is
// written only by rank s and only with growing values; a waiter spins with acquire loads on its own words.
// A wait gives up after kSpinLimit cycles (or as soon as another wait of this GPU has given up) and raises
// *err, so a lost peer turns into an error code instead of a hung GPU.
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// lane p: raise word (slot, me) of rank p if p is in sig, then wait for own word (slot, p) if p is in wait
__device__ __forceinline__ void signal_and_wait(const Peers &peers, int slot, unsigned long long value, unsigned sig, unsigned wait, int *err) {
  const int p = threadIdx.x;
  if (p < peers.n && ((sig >> p) & 1u)) st_release_sys(peers.flags[p] + slot * kMaxPeers + peers.rank, value);
  if (p < peers.n && ((wait >> p) & 1u)) {
    const unsigned long long *src = peers.flags[peers.rank] + slot * kMaxPeers + p;
    const long long t0 = clock64();
    while (ld_acquire_sys(src) < value) {
      if (*reinterpret_cast<volatile int *>(err) != 0 || clock64() - t0 > kSpinLimit) {
        atomicExch(err, 1);
        break;
      }
    }
  }
}
__global__ void __launch_bounds__(32) peer_sync(Peers peers, int slot, unsigned long long value, unsigned sig, unsigned wait, int *err) {
  if (sig) __threadfence_system();  // what this rank wrote before the sync is visible to whoever sees the flag
  signal_and_wait(peers, slot, value, sig, wait, err);
}

// push_rects: rectangle d (rows x cols, column-major, leading dimension ld, first entry at fac + off) is stored
// at the same offset into the factor buffer of every rank in `mask`.  One CTA per (rectangle, group of
// columns); rows move as 16-byte pairs (offsets and leading dimensions are even, a trailing odd row
// takes its padding neighbour along; col_groups CTAs share a rectangle, chosen by the host so that even a single
// 256 x 256 block spreads over enough SMs to fill the links).  tri0 < kNoTriDev: entries with column > tri0 + row are skipped (the
// strictly upper part of a diagonal block).  The last CTA to finish raises flag (slot, me) of the ranks in sig.
// One CTA per (rectangle, group of columns).
constexpr int kPushThreads = 256;
constexpr int kNoTriDev = 1 << 29;
__global__ void __launch_bounds__(kPushThreads) push_rects(const RectDesc *__restrict__ rects, double *__restrict__ fac, Peers peers, unsigned mask,
                                                           int col_groups, int slot, unsigned long long value, unsigned sig,
                                                           unsigned *__restrict__ counter) {
  const RectDesc d = rects[blockIdx.x / col_groups];
  const int cg = blockIdx.x % col_groups;
  const int rows2 = (d.rows + 1) / 2;
  const int cpg = (d.cols + col_groups - 1) / col_groups;
  for (int c = cg * cpg + (int)threadIdx.x / 128; c < min(d.cols, (cg + 1) * cpg); c += kPushThreads / 128) {
    for (int r2 = threadIdx.x % 128; r2 < rows2; r2 += 128) {
      if (d.tri0 < kNoTriDev && c > d.tri0 + 2 * r2 + 1) continue;
      const int64_t o = d.off + 2 * r2 + (int64_t)c * d.ld;
      const double2 v = *reinterpret_cast<const double2 *>(fac + o);
#pragma unroll
      for (int p = 0; p < kMaxPeers; p++)
        if (p < peers.n && ((mask >> p) & 1u)) *reinterpret_cast<double2 *>(peers.fac[p] + o) = v;
    }
  }
  if (!sig) return;
  // Every CTA orders its stores at GPU scope before it is counted; the last one fences at system scope and raises
  // the flags (measured on 8 B200s, tools/push_bench.cu: 3.5 us less per push than a system fence in every CTA,
  // and no stale element in 300 x 12 MB of pushes checked by the receiver right after the flag).
  __threadfence();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *counter = 0;
  __threadfence_system();
  if (threadIdx.x < 32) signal_and_wait(peers, slot, value, sig, 0u, nullptr);
}

// reduce_rects: for every rectangle (a block of rows this rank owns of a top panel) the partial sums held by the
// ranks in the rectangle's mask (those whose subtrees touch it, the owner included) are added in ascending rank
// order and stored into this rank's copy.
__global__ void __launch_bounds__(kPushThreads) reduce_rects(const RectDesc *__restrict__ rects, double *__restrict__ fac, Peers peers, unsigned gmask,
                                                             int col_groups) {
  const RectDesc d = rects[blockIdx.x / col_groups];
  const unsigned mask = d.mask & gmask;
  const int cg = blockIdx.x % col_groups;
  const int rows2 = (d.rows + 1) / 2;
  const int cpg = (d.cols + col_groups - 1) / col_groups;
  for (int c = cg * cpg + (int)threadIdx.x / 128; c < min(d.cols, (cg + 1) * cpg); c += kPushThreads / 128) {
    for (int r2 = threadIdx.x % 128; r2 < rows2; r2 += 128) {
      if (d.tri0 < kNoTriDev && c > d.tri0 + 2 * r2 + 1) continue;
      const int64_t o = d.off + 2 * r2 + (int64_t)c * d.ld;
      double2 s = make_double2(0.0, 0.0);
#pragma unroll
      for (int p = 0; p < kMaxPeers; p++)
        if (p < peers.n && ((mask >> p) & 1u)) {
          const double2 v = *reinterpret_cast<const double2 *>(peers.fac[p] + o);
          s.x += v.x, s.y += v.y;
        }
      *reinterpret_cast<double2 *>(fac + o) = s;
    }
  }
}

// compare_rects: largest |difference| between this rank's copy of the rectangles and the copy of rank `peer`
// (verification of the claim that every rank ends with identical top panels); the result is the bit pattern of a
// non-negative double, so an integer atomicMax orders it.
__global__ void __launch_bounds__(kPushThreads) compare_rects(const RectDesc *__restrict__ rects, const double *__restrict__ fac, Peers peers, int peer,
                                                              int col_groups, unsigned long long *__restrict__ worst) {
  const RectDesc d = rects[blockIdx.x / col_groups];
  const int cg = blockIdx.x % col_groups;
  const int cpg = (d.cols + col_groups - 1) / col_groups;
  const double *__restrict__ other = peers.fac[peer];
  double w = 0.0;
  for (int c = cg * cpg + (int)threadIdx.x / 128; c < min(d.cols, (cg + 1) * cpg); c += kPushThreads / 128)
    for (int r = threadIdx.x % 128; r < d.rows; r += 128) {
      if (d.tri0 < kNoTriDev && c > d.tri0 + r) continue;
      const int64_t o = d.off + r + (int64_t)c * d.ld;
      const double df = fabs(fac[o] - other[o]);
      w = fmax(w, (df == df) ? df : 1e300);  // a NaN on either side counts as a difference
    }
  if (w > 0.0) atomicMax(worst, (unsigned long long)__double_as_longlong(w));
}

}  // namespace chb
