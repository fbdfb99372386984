// sm_100a kernels of the numeric factorization.  A handful of grouped kernels execute the whole level
// schedule (see schedule.cc); one scatter kernel assembles A; two small kernels carry the multi-GPU
// exchange over NVLink peer memory.
//
//   gemm_grouped_ws  C -= sum_c A_c B_c^T on FP64 tensor cores (mma.sync m8n8k4 -> SASS DMMA.8x8x4),
//                  warp-specialised: a producer warp stages operand tiles with the TMA bulk-copy engine
//                  (cp.async.bulk + mbarrier expect_tx), consumer warps issue the MMAs; one CTA per
//                  destination tile, contributors accumulated in registers in a fixed order
//                  (atomic-free, deterministic), lower-triangle masking for SYRK destinations.
//                  SHARED variant: the tile is also stored into every peer's copy of the factor
//                  (coalesced P2P stores), fusing the update with its broadcast.
//                  Replaces cblas_dgemm / cblas_dsyrk as called at blas.rg:139-142, 187-189.
//   gemm_grouped   the earlier cp.async version of the same kernel (CHOL_GEMM_WS=0).
//   gemm_small_warp  the same update for small fronts: one warp per 32x32 tile, no shared memory.
//   potrf_tile     Cholesky of one NB x NB pivot tile in shared memory, blocked by 16 columns
//                  (LAPACKE_dpotrf, blas.rg:71).
//   trsm_tile      one 128-row slab times L^-T by forward substitution, one row per thread
//                  (cblas_dtrsm Right/Lower/Trans/NonUnit, blas.rg:99-100).
//   assemble       factor[a_off[e]] = value[e]   (fill_block, mmat.rg:529-633).
//   peer_barrier / allreduce_top   cross-GPU barrier and sum of the ranks' top-panel copies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chol_internal.h"

namespace chb {

constexpr int kMaxPeers = 8;
struct Peers {
  double *fac[kMaxPeers];
  unsigned long long *flags[kMaxPeers];
  int n, rank;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
// D(8x8) += A(8x4, row) * B(4x8, col); lane = 4*g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES>
struct GemmCfg {
  static constexpr int kWarpsM = BM / WM, kWarpsN = BN / WN;
  static constexpr int kThreads = kWarpsM * kWarpsN * 32;
  static constexpr int kPad = 4;  // row stride = 4 mod 16 doubles: conflict-free DMMA fragment loads
  static constexpr int kLdA = BM + kPad, kLdB = BN + kPad;
  static constexpr int kStageDoubles = BK * (kLdA + kLdB);
  static constexpr int kSmemBytes = STAGES * kStageDoubles * 8;
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool SHARED>
__global__ void __launch_bounds__(GemmCfg<BM, BN, BK, WM, WN, STAGES>::kThreads)
    gemm_grouped(const GemmProblem *__restrict__ probs, const GemmContrib *__restrict__ contribs,
                 const TileRef *__restrict__ tiles, double *__restrict__ fac, Peers peers) {
  using Cfg = GemmCfg<BM, BN, BK, WM, WN, STAGES>;
  extern __shared__ __align__(16) double smem[];
  const TileRef tile = tiles[blockIdx.x];
  const GemmProblem pr = probs[tile.prob];
  const int row0 = tile.tr * BM, col0 = tile.tc * BN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % Cfg::kWarpsM) * WM, wn0 = (warp / Cfg::kWarpsM) * WN;
  constexpr int MB = WM / 8, NBk = WN / 8;
  double acc[MB][NBk][2];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NBk; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int mrem = pr.M - row0, nrem = pr.N - col0;  // valid rows of A / B in this tile (may exceed BM/BN)

  for (int c = 0; c < pr.contrib_count; c++) {
    const GemmContrib cb = contribs[pr.contrib_begin + c];
    const double *__restrict__ A = fac + cb.a_off + row0;
    const double *__restrict__ Bp = fac + cb.b_off + col0;
    const int K = cb.K, lda = cb.lda, ldb = cb.ldb;
    const int nk = (K + BK - 1) / BK;

    auto load_stage = [&](int stage, int kt) {
      double *As = smem + stage * Cfg::kStageDoubles;
      double *Bs = As + BK * Cfg::kLdA;
      const int k0 = kt * BK;
#pragma unroll
      for (int i = tid; i < BK * (BM / 2); i += Cfg::kThreads) {
        int kk = i / (BM / 2), m2 = (i % (BM / 2)) * 2;
        bool ok = (m2 < mrem) && (k0 + kk < K);
        cp_async16(As + kk * Cfg::kLdA + m2, ok ? (A + m2 + (size_t)(k0 + kk) * lda) : A, ok);
      }
#pragma unroll
      for (int i = tid; i < BK * (BN / 2); i += Cfg::kThreads) {
        int kk = i / (BN / 2), n2 = (i % (BN / 2)) * 2;
        bool ok = (n2 < nrem) && (k0 + kk < K);
        cp_async16(Bs + kk * Cfg::kLdB + n2, ok ? (Bp + n2 + (size_t)(k0 + kk) * ldb) : Bp, ok);
      }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < nk) load_stage(s, s);
      cp_async_commit();
    }
    for (int kt = 0; kt < nk; kt++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      int nxt = kt + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
      const double *As = smem + (kt % STAGES) * Cfg::kStageDoubles;
      const double *Bs = As + BK * Cfg::kLdA;
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
    }
    cp_async_wait<0>();
    __syncthreads();
  }

  double *__restrict__ C = fac + pr.c_off;
#pragma unroll
  for (int i = 0; i < MB; i++) {
    const int r = row0 + wm0 + i * 8 + g;
    if (r >= pr.M) continue;
#pragma unroll
    for (int j = 0; j < NBk; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int cc = col0 + wn0 + j * 8 + 2 * t + e;
        if (cc < pr.N && (!pr.tri || r >= cc)) {
          const size_t o = r + (size_t)cc * pr.ldc;
          const double v = C[o] - acc[i][j][e];
          if (SHARED) {
#pragma unroll
            for (int p = 0; p < kMaxPeers; p++)
              if (p < peers.n) peers.fac[p][pr.c_off + o] = v;
          } else
            C[o] = v;
        }
      }
    }
  }
  if (SHARED) __threadfence_system();
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

template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB, bool SHARED>
__global__ void __launch_bounds__(GemmWsCfg<BM, BN, BK, WM, WN, STAGES>::kThreads, MINB)
    gemm_grouped_ws(const GemmProblem *__restrict__ probs, const GemmContrib *__restrict__ contribs,
                    const TileRef *__restrict__ tiles, double *__restrict__ fac, Peers peers) {
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
  if (!SHARED) {
    // local destination: read-modify-write straight from the accumulator fragments (measured 2 % faster
    // on one GPU than staging through shared memory).  All loads of one row block are issued before the
    // first store: written as `C[..] -= acc` the compiler has to keep every store ahead of the next load
    // (same array), i.e. 32 dependent round trips to L2/HBM per thread -- ncu's source view had 40 % of
    // the warp samples of the K = 256 launches on those DADDs, and the DMMA pipe at 82 %.
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
          ok[j][e] = r < pr.M && cc < pr.N && (!pr.tri || r >= cc);
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
    return;
  }
  // SHARED: the tile also goes to every peer's copy.  The accumulators are parked column-major in shared
  // memory (stride BM + 2 keeps the fragment stores conflict-free), then every warp walks whole columns,
  // lane = row, so the peer stores over NVLink are coalesced runs instead of 8-byte scatters.
  constexpr int kLdC = BM + 2;
  static_assert(BN * kLdC <= STAGES * Cfg::kStageDoubles, "C staging tile must fit in the stage ring");
  asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::kConsumers * 32) : "memory");  // all consumers are done with the ring
  double *Cs = smem;
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NBk; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) Cs[(wn0 + j * 8 + 2 * t + e) * kLdC + wm0 + i * 8 + g] = acc[i][j][e];
  asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::kConsumers * 32) : "memory");
  // 16-byte peer stores when every column of the tile starts on a 16-byte boundary (always true for the
  // in-panel updates; Schur destinations start wherever their cluster starts)
  const bool vec = (((pr.c_off + row0) | pr.ldc) & 1) == 0;
  // Four columns per step: their loads of C go out together, then the peer stores.  One column at a time,
  // every load had to wait for the previous column's stores (one of the peers is this rank's own copy).
  constexpr int kColBatch = 4;
  for (int jb = warp; jb < BN; jb += Cfg::kConsumers * kColBatch) {
    if (col0 + jb >= pr.N) break;
    if (vec) {
#pragma unroll
      for (int rr0 = 0; rr0 < BM; rr0 += 64) {
        const int rr = rr0 + 2 * lane, r = row0 + rr;
        double2 cv[kColBatch];
        int okm[kColBatch];  // bit 0: row r is written, bit 1: row r + 1
#pragma unroll
        for (int u = 0; u < kColBatch; u++) {
          const int j = jb + u * Cfg::kConsumers, cc = col0 + j;
          const bool in = j < BN && cc < pr.N;
          const bool ok0 = in && r < pr.M && (!pr.tri || r >= cc), ok1 = in && r + 1 < pr.M && (!pr.tri || r + 1 >= cc);
          okm[u] = (ok0 ? 1 : 0) | (ok1 ? 2 : 0);
          const size_t o = r + (size_t)cc * pr.ldc;
          cv[u] = make_double2(0.0, 0.0);
          if (okm[u] == 3) cv[u] = *reinterpret_cast<const double2 *>(C + o);
          else if (okm[u] == 1) cv[u].x = C[o];
          else if (okm[u] == 2) cv[u].y = C[o + 1];
        }
#pragma unroll
        for (int u = 0; u < kColBatch; u++) {
          if (!okm[u]) continue;
          const int j = jb + u * Cfg::kConsumers, cc = col0 + j;
          const size_t o = r + (size_t)cc * pr.ldc;
          const double2 v = make_double2(cv[u].x - Cs[j * kLdC + rr], cv[u].y - Cs[j * kLdC + rr + 1]);
          if (okm[u] == 3) {
#pragma unroll
            for (int p = 0; p < kMaxPeers; p++)
              if (p < peers.n) *reinterpret_cast<double2 *>(peers.fac[p] + pr.c_off + o) = v;
          } else {
            const int q = okm[u] == 1 ? 0 : 1;
            const double w = q ? v.y : v.x;
#pragma unroll
            for (int p = 0; p < kMaxPeers; p++)
              if (p < peers.n) peers.fac[p][pr.c_off + o + q] = w;
          }
        }
      }
    } else {
#pragma unroll
      for (int rr0 = 0; rr0 < BM; rr0 += 32) {
        const int rr = rr0 + lane, r = row0 + rr;
        double cv[kColBatch];
        bool ok[kColBatch];
#pragma unroll
        for (int u = 0; u < kColBatch; u++) {
          const int j = jb + u * Cfg::kConsumers, cc = col0 + j;
          ok[u] = j < BN && cc < pr.N && r < pr.M && (!pr.tri || r >= cc);
          cv[u] = ok[u] ? C[r + (size_t)cc * pr.ldc] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kColBatch; u++) {
          if (!ok[u]) continue;
          const int j = jb + u * Cfg::kConsumers, cc = col0 + j;
          const size_t o = r + (size_t)cc * pr.ldc;
          const double v = cv[u] - Cs[j * kLdC + rr];
#pragma unroll
          for (int p = 0; p < kMaxPeers; p++)
            if (p < peers.n) peers.fac[p][pr.c_off + o] = v;
        }
      }
    }
  }
  __threadfence_system();
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
        ok[j][e] = r < pr.M && cc < pr.N && (!pr.tri || r >= cc);
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
constexpr int kNB = 64;

// Pivot tile in shared memory (row stride 65: conflict-free), blocked by 16 columns.  Inside a block:
// left-looking by column with one thread per row (64 threads, named barrier) -- all rows at or below
// the diagonal take their short dot product with row k at once, thread k turns its result into
// 1/sqrt, the others scale.  After a block: rank-16 update of the trailing lower triangle by all 256
// threads.  Compact loops on purpose: a fully unrolled version is instruction-fetch bound.
constexpr int kPotrfThreads = 256;
constexpr int kPB = 16;
__global__ void __launch_bounds__(kPotrfThreads) potrf_tile(const PotrfDesc *__restrict__ descs, double *__restrict__ fac,
                                                            int *__restrict__ info) {
  __shared__ double T[kNB][kNB + 1];
  __shared__ double rk;  // 1 / L[k][k]
  const PotrfDesc d = descs[blockIdx.x];
  double *__restrict__ A = fac + d.off;
  const int nb = d.nb, tid = threadIdx.x;
  {  // thread (row i, column group cg) loads 16 columns, eight loads in flight
    const int i = tid & (kNB - 1), cg = tid / kNB;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int c = cg * 16 + h * 8 + u;
        v[u] = (i < nb && c <= i) ? A[i + (size_t)c * d.ld] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) T[i][cg * 16 + h * 8 + u] = v[u];
    }
  }
  __syncthreads();
  for (int kb = 0; kb < nb; kb += kPB) {
    const int kend = min(kb + kPB, nb);
    if (tid < kNB) {
      const int i = tid;
      for (int k = kb; k < kend; k++) {
        double sk = 0.0;
        if (i >= k && i < nb) {
          double s0 = T[i][k], s1 = 0.0;
          int j = kb;
          for (; j + 3 < k; j += 4) {
            const double a0 = T[i][j], a1 = T[i][j + 1], a2 = T[i][j + 2], a3 = T[i][j + 3];
            const double b0 = T[k][j], b1 = T[k][j + 1], b2 = T[k][j + 2], b3 = T[k][j + 3];
            s0 -= a0 * b0, s1 -= a1 * b1, s0 -= a2 * b2, s1 -= a3 * b3;
          }
          for (; j < k; j++) s0 -= T[i][j] * T[k][j];
          sk = s0 + s1;
        }
        if (i == k) {
          if (!(sk > 0.0)) {
            atomicMin(info, d.col0 + k + 1);  // 1-based permuted column of the first bad pivot
            sk = 1.0;
          }
          const double r = rsqrt(sk);
          T[k][k] = sk * r;
          rk = r;
        }
        asm volatile("bar.sync 1, 64;\n" ::: "memory");
        if (i > k && i < nb) T[i][k] = sk * rk;
        asm volatile("bar.sync 1, 64;\n" ::: "memory");
      }
    }
    __syncthreads();
    const int rem = nb - kend;
    if (rem > 0) {  // T[i][j] -= sum_k T[i][k] T[j][k] over the block just factored, i >= j >= kend
      for (int idx = tid; idx < rem * rem; idx += kPotrfThreads) {
        const int i = kend + idx % rem, j = kend + idx / rem;
        if (i < j) continue;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int k = kb; k < kend; k += 2) {
          s0 += T[i][k] * T[j][k];
          if (k + 1 < kend) s1 += T[i][k + 1] * T[j][k + 1];
        }
        T[i][j] -= s0 + s1;
      }
    }
    __syncthreads();
  }
  {
    const int i = tid & (kNB - 1), cg = tid / kNB;
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int c = cg * 16 + u;
      if (i < nb && c <= i && c < nb) A[i + (size_t)c * d.ld] = T[i][c];
    }
  }
}

// potrf_tile_w: same tile, but the column steps are done by ONE warp (two rows per lane) so that the 64
// dependent steps synchronise with __syncwarp and a shuffle instead of block barriers and a round trip
// through shared memory; every lane computes the pivot's 1/sqrt itself.  Blocks of 8 columns keep the
// in-block dot products short; the rank-8 trailing update uses all 256 threads.
constexpr int kPB2 = 8;
__global__ void __launch_bounds__(kPotrfThreads) potrf_tile_w(const PotrfDesc *__restrict__ descs, double *__restrict__ fac,
                                                              int *__restrict__ info) {
  __shared__ double T[kNB][kNB + 1];
  const PotrfDesc d = descs[blockIdx.x];
  double *__restrict__ A = fac + d.off;
  const int nb = d.nb, tid = threadIdx.x, lane = tid & 31;
  {
    const int i = tid & (kNB - 1), cg = tid / kNB;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int c = cg * 16 + h * 8 + u;
        v[u] = (i < nb && c <= i) ? A[i + (size_t)c * d.ld] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) T[i][cg * 16 + h * 8 + u] = v[u];
    }
  }
  __syncthreads();
  for (int kb = 0; kb < nb; kb += kPB2) {
    const int kend = min(kb + kPB2, nb);
    if (tid < 32) {
      const int r0 = lane, r1 = lane + 32;
      for (int k = kb; k < kend; k++) {
        double s0 = 0.0, s1 = 0.0;
        if (r0 >= k && r0 < nb) s0 = T[r0][k];
        if (r1 >= k && r1 < nb) s1 = T[r1][k];
        for (int j = kb; j < k; j++) {
          const double b = T[k][j];
          s0 -= T[r0][j] * b;  // rows above the diagonal or beyond nb compute garbage that is never stored
          s1 -= T[r1][j] * b;
        }
        double sk = __shfl_sync(0xffffffffu, k < 32 ? s0 : s1, k & 31);
        if (!(sk > 0.0)) {
          if (lane == 0) atomicMin(info, d.col0 + k + 1);  // 1-based permuted column of the first bad pivot
          sk = 1.0;
        }
        // 1/sqrt from the single-precision seed and two Newton steps in double (error ~2 ulp); the
        // library routine's longer dependent chain sits on the critical path of all 64 columns
        double r;
        if (sk > 1e-30 && sk < 1e30) {
          r = (double)rsqrtf((float)sk);
          const double hs = 0.5 * sk;
          r = r * (1.5 - hs * r * r);
          r = r * (1.5 - hs * r * r);
        } else
          r = rsqrt(sk);
        if (r0 == k || r1 == k) T[k][k] = sk * r;
        if (r0 > k && r0 < nb) T[r0][k] = s0 * r;
        if (r1 > k && r1 < nb) T[r1][k] = s1 * r;
        __syncwarp();
      }
    }
    __syncthreads();
    const int rem = nb - kend;
    if (rem > 0) {
      for (int idx = tid; idx < rem * rem; idx += kPotrfThreads) {
        const int i = kend + idx % rem, j = kend + idx / rem;
        if (i < j) continue;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int k = kb; k < kend; k += 2) {
          s0 += T[i][k] * T[j][k];
          if (k + 1 < kend) s1 += T[i][k + 1] * T[j][k + 1];
        }
        T[i][j] -= s0 + s1;
      }
    }
    __syncthreads();
  }
  {
    const int i = tid & (kNB - 1), cg = tid / kNB;
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int c = cg * 16 + u;
      if (i < nb && c <= i && c < nb) A[i + (size_t)c * d.ld] = T[i][c];
    }
  }
}

// potrf_tile_r (EXPERIMENTAL, CHOL_POTRF_R=1; parity-checked on one fixture, 2 % faster only: it spills): the same tile, right-looking and
// register-resident.  64 threads, thread i holds row i of the tile (64 doubles).  Step k: every thread
// publishes its entry of column k (unscaled) in shared memory, one barrier, then a_ij -= (a_ik / a_kk) a_jk
// for the rest of its row and a_ik *= 1/sqrt(a_kk).  One barrier and one rsqrt per column step instead of the
// blocked dot products of potrf_tile_w (1 140 cycles per column): the step is ~60 dependent cycles of
// arithmetic plus the barrier.  Entries above the diagonal carry finite garbage that is never stored.
constexpr int kPotrfRThreads = kNB;
__device__ __forceinline__ double rsqrt_newton(double s) {
  // 1/sqrt from the single-precision seed and two Newton steps in double (~2 ulp), as in potrf_tile_w
  if (s > 1e-30 && s < 1e30) {
    double r = (double)rsqrtf((float)s);
    const double hs = 0.5 * s;
    r = r * (1.5 - hs * r * r);
    return r * (1.5 - hs * r * r);
  }
  return rsqrt(s);
}
__global__ void __launch_bounds__(kPotrfRThreads) potrf_tile_r(const PotrfDesc *__restrict__ descs, double *__restrict__ fac,
                                                               int *__restrict__ info) {
  __shared__ __align__(16) double colbuf[2][kNB];
  const PotrfDesc d = descs[blockIdx.x];
  double *__restrict__ A = fac + d.off;
  const int i = threadIdx.x, nb = d.nb;
  double a[kNB];
  // rows and columns beyond nb are an identity block: their steps are no-ops
#pragma unroll
  for (int c = 0; c < kNB; c++) a[c] = (i < nb && c <= i) ? A[i + (size_t)c * d.ld] : ((c == i && i >= nb) ? 1.0 : 0.0);
#pragma unroll
  for (int k = 0; k < kNB; k++) {
    double *cb = colbuf[k & 1];  // double-buffered: a thread is at most one step ahead of the slowest one
    cb[i] = a[k];
    __syncthreads();
    double p = cb[k];
    if (!(p > 0.0)) {
      if (i == k) atomicMin(info, d.col0 + k + 1);  // 1-based permuted column of the first bad pivot
      p = 1.0;
    }
    const double r = rsqrt_newton(p);
    const double t = a[k] * (r * r);
    a[k] *= r;
    int j = k + 1;
    if (j & 1) {
      if (j < kNB) a[j] -= t * cb[j];
      j++;
    }
#pragma unroll
    for (; j < kNB; j += 2) {
      const double2 c2 = *reinterpret_cast<const double2 *>(&cb[j]);
      a[j] -= t * c2.x;
      a[j + 1] -= t * c2.y;
    }
  }
  if (i < nb) {
#pragma unroll
    for (int c = 0; c < kNB; c++)
      if (c <= i) A[i + (size_t)c * d.ld] = a[c];
  }
}

// potrf_tile_r2 (EXPERIMENTAL, CHOL_POTRF_R=2; parity-checked on one fixture, pivot-tile time of 64^3 5.97 -> 4.90 ms,
// profiles/experimental_variants_r01.md; off by default until the whole GPU suite has run with it): potrf_tile_r with each row split over
// two threads (columns 0-31 in warps 0-1, columns 32-63 in warps 2-3): 32 doubles per thread instead of 64,
// no register spills.  The two halves run different (warp-uniform) code and meet at one named barrier per step.
__device__ __forceinline__ void bar_sync_128() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }
// MINB = 1: 254 registers (the measured build); MINB = 3: 168 registers, no spills, three CTAs per SM for the
// launches with thousands of tiles at the bottom of the tree (CHOL_POTRF_R=3, not measured yet).
template <int MINB>
__global__ void __launch_bounds__(2 * kNB, MINB) potrf_tile_r2(const PotrfDesc *__restrict__ descs, double *__restrict__ fac,
                                                         int *__restrict__ info) {
  __shared__ __align__(16) double colbuf[2][kNB];
  const PotrfDesc d = descs[blockIdx.x];
  double *__restrict__ A = fac + d.off;
  const int i = threadIdx.x & (kNB - 1), h = threadIdx.x >> 6, nb = d.nb;
  constexpr int H = kNB / 2;
  double a[H];
#pragma unroll
  for (int u = 0; u < H; u++) {
    const int c = h * H + u;
    a[u] = (i < nb && c <= i) ? A[i + (size_t)c * d.ld] : ((c == i && i >= nb) ? 1.0 : 0.0);
  }
  if (h == 0) {
#pragma unroll
    for (int k = 0; k < H; k++) {
      double *cb = colbuf[k & 1];
      cb[i] = a[k];
      bar_sync_128();
      double p = cb[k];
      if (!(p > 0.0)) {
        if (i == k) atomicMin(info, d.col0 + k + 1);
        p = 1.0;
      }
      const double r = rsqrt_newton(p);
      const double t = a[k] * (r * r);
      a[k] *= r;
#pragma unroll
      for (int j = k + 1; j < H; j++) a[j] -= t * cb[j];
    }
#pragma unroll
    for (int k = H; k < kNB; k++) bar_sync_128();  // columns 0-31 are final: keep the barrier count of the other half
  } else {
#pragma unroll
    for (int k = 0; k < H; k++) {
      const double *cb = colbuf[k & 1];
      bar_sync_128();
      double p = cb[k];
      if (!(p > 0.0)) p = 1.0;  // reported by the thread that owns the pivot
      const double r = rsqrt_newton(p);
      const double t = cb[i] * (r * r);
#pragma unroll
      for (int u = 0; u < H; u += 2) {
        const double2 c2 = *reinterpret_cast<const double2 *>(&cb[H + u]);
        a[u] -= t * c2.x;
        a[u + 1] -= t * c2.y;
      }
    }
#pragma unroll
    for (int k = H; k < kNB; k++) {
      double *cb = colbuf[k & 1];
      cb[i] = a[k - H];
      bar_sync_128();
      double p = cb[k];
      if (!(p > 0.0)) {
        if (i == k) atomicMin(info, d.col0 + k + 1);
        p = 1.0;
      }
      const double r = rsqrt_newton(p);
      const double t = a[k - H] * (r * r);
      a[k - H] *= r;
#pragma unroll
      for (int j = k + 1; j < kNB; j++) a[j - H] -= t * cb[j];
    }
  }
  if (i < nb) {
#pragma unroll
    for (int u = 0; u < H; u++)
      if (h * H + u <= i) A[i + (size_t)(h * H + u) * d.ld] = a[u];
  }
}

// 128-row slab per CTA, one row per thread.  The slab (k-major, so a warp reads consecutive words) and
// L^T live in shared memory; columns are solved eight at a time with eight independent FMA chains, the
// eight multipliers of one k come as four broadcast vector loads.  Loops are deliberately not fully
// unrolled: the straight-line version was instruction-fetch bound.
constexpr int kSlab = 128;
constexpr int kTrsmSmemBytes = (kNB * kNB + kNB + kNB * kSlab) * 8;
// BATCH (EXPERIMENTAL, CHOL_TRSM_BATCH=1; parity-checked on one fixture, trsm time of 64^3 4.06 -> 3.47 ms,
// profiles/experimental_variants_r01.md; off by default until the whole GPU suite has run with it): all 64
// column loads of the slab row in flight at once instead of eight rounds of eight.
template <bool BATCH>
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
  if (BATCH) {
    double v[kNB];
#pragma unroll
    for (int c = 0; c < kNB; c++) v[c] = (live && c < nb) ? Bp[(size_t)c * d.ld] : 0.0;
#pragma unroll
    for (int c = 0; c < kNB; c++) xs[c][tid] = v[c];
  } else {
    for (int c0 = 0; c0 < nb8; c0 += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = (live && c0 + u < nb) ? Bp[(size_t)(c0 + u) * d.ld] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++) xs[c0 + u][tid] = v[u];
    }
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
// Cross-GPU barrier over peer-mapped flag words: every rank publishes `epoch` into its slot of every
// peer's flag array, then waits until all peers have published theirs.  One kernel per GPU, each
// GPU runs its own process's kernel, so the waits cannot starve each other.
__global__ void peer_barrier(Peers peers, unsigned long long epoch) {
  const int p = threadIdx.x;
  __threadfence_system();
  if (p < peers.n) {
    unsigned long long *dst = peers.flags[p] + peers.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(dst), "l"(epoch) : "memory");
    const unsigned long long *src = peers.flags[peers.rank] + p;
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(src) : "memory");
    } while (v < epoch);
  }
  __syncthreads();
  __threadfence_system();
}

// Sum of all ranks' copies of the top panels: rank r reduces slice r (peer loads over NVLink, fixed
// summation order, so every rank ends with bit-identical values) and stores the sums into every
// copy (peer stores).  Bracketed by peer_barrier on both sides.
// One panel per launch (off2 / n2 in double2 units, ld2 = ld / 2, npiv = pivot-block columns): chunks of
// 4096 double2 go round-robin to the ranks; only the ranks in `mask` hold contributions (the others'
// copies are still zero), and elements strictly above the pivot block's diagonal are never touched by
// the factorization, so they are skipped.
constexpr int kArChunk = 4096;
__global__ void __launch_bounds__(256) allreduce_top(Peers peers, int64_t off2, int64_t n2, int ld2, int npiv, unsigned mask) {
  const int64_t nchunks = (n2 + kArChunk - 1) / kArChunk;
  for (int64_t ch = (int64_t)blockIdx.x * peers.n + peers.rank; ch < nchunks; ch += (int64_t)gridDim.x * peers.n) {
    for (int64_t i = ch * kArChunk + threadIdx.x; i < min(n2, (ch + 1) * kArChunk); i += blockDim.x) {
      const int64_t col = i / ld2;
      const int64_t row = (i - col * ld2) * 2;
      if (col < npiv && row + 1 < col) continue;
      const int64_t gi = off2 + i;
      double2 s = make_double2(0.0, 0.0);
#pragma unroll
      for (int p = 0; p < kMaxPeers; p++)
        if (p < peers.n && ((mask >> p) & 1u)) {
          const double2 v = reinterpret_cast<const double2 *>(peers.fac[p])[gi];
          s.x += v.x, s.y += v.y;
        }
#pragma unroll
      for (int p = 0; p < kMaxPeers; p++)
        if (p < peers.n) reinterpret_cast<double2 *>(peers.fac[p])[gi] = s;
    }
  }
  __threadfence_system();
}

}  // namespace chb
