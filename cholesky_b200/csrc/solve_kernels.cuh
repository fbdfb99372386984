// Triangular-solve kernels on the panel storage (reference: dtrsv / dgemv tasks, blas.rg:217-290, driven
// by mmat.rg:1394-1479).  HBM-bound work: every factor entry is read once per sweep, coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "solve.h"

namespace chb {

constexpr int kSolveNB = 64, kSolveSlab = 128, kSolveColG = 8;
constexpr int kSolveWB = 256;  // block-column width of the sweeps inside a pivot block (one solve_block + one GEMV launch per block column)

__global__ void permute_in_kernel(const double *__restrict__ b, const int *__restrict__ perm, int n, double *__restrict__ x) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) x[p] = b[perm[p]];  // fill_b, mmat.rg:769-783
}
__global__ void permute_out_kernel(const double *__restrict__ x, const int *__restrict__ perm, int n, double *__restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) out[perm[p]] = x[p];  // mmat.rg:1483-1491
}

// Small pivot blocks (every block of the launch at most 64 wide -- the bottom of the tree, thousands per launch): the
// same tile solve with a light CTA.  All 128 threads bring the tile into shared memory, one warp does the steps.
constexpr int kSolveTileThreads = 128;
template <bool BWD>
__global__ void __launch_bounds__(kSolveTileThreads) solve_tile(const SolveTile *__restrict__ descs, const double *__restrict__ fac,
                                                                 double *__restrict__ x) {
  __shared__ double Ls[kSolveNB][kSolveNB + 1];  // Ls[c][r] = L(r, c), r >= c
  const SolveTile d = descs[blockIdx.x];
  const int tid = threadIdx.x, nb = d.nb;
  const double *__restrict__ Lg = fac + d.l_off;
  {
    const int r = tid & (kSolveNB - 1), cq = tid >> 6;  // thread = row, every other column
    double v[kSolveNB / 2];
#pragma unroll
    for (int u = 0; u < kSolveNB / 2; u++) {
      const int c = 2 * u + cq;
      v[u] = (r < nb && c <= r) ? Lg[r + (size_t)c * d.ld] : (r == c ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < kSolveNB / 2; u++) Ls[2 * u + cq][r] = v[u];
  }
  __syncthreads();
  if (tid >= 32) return;
  const int lane = tid;
  double x0 = lane < nb ? x[d.x0 + lane] : 0.0, x1 = lane + 32 < nb ? x[d.x0 + lane + 32] : 0.0;
  const double rd0 = 1.0 / Ls[lane][lane], rd1 = 1.0 / Ls[lane + 32][lane + 32];
  if (!BWD) {
#pragma unroll 8
    for (int k = 0; k < kSolveNB; k++) {
      const double xk = __shfl_sync(0xffffffffu, k < 32 ? x0 * rd0 : x1 * rd1, k & 31);
      if (lane == k) x0 = xk;
      if (lane + 32 == k) x1 = xk;
      if (lane > k) x0 = fma(-Ls[k][lane], xk, x0);
      if (lane + 32 > k) x1 = fma(-Ls[k][lane + 32], xk, x1);
    }
  } else {
#pragma unroll 8
    for (int k = kSolveNB - 1; k >= 0; k--) {
      const double xk = __shfl_sync(0xffffffffu, k < 32 ? x0 * rd0 : x1 * rd1, k & 31);
      if (lane == k) x0 = xk;
      if (lane + 32 == k) x1 = xk;
      if (lane < k) x0 = fma(-Ls[lane][k], xk, x0);  // L(k, i), i < k
      if (lane + 32 < k) x1 = fma(-Ls[lane + 32][k], xk, x1);
    }
  }
  if (lane < nb) x[d.x0 + lane] = x0;
  if (lane + 32 < nb) x[d.x0 + lane + 32] = x1;
}

// Diagonal block of a block column (w <= 256): forward L y = b (BWD = false) or backward L^T x = y (BWD = true), one
// CTA, the block's part of x in shared memory.  Per 64-column tile: all threads bring the tile into shared memory,
// ONE warp does the 64 dependent steps with two rows per lane (pivot value by shuffle, reciprocals of the diagonal
// taken up front: a step is shuffle + multiply + FMA), and the rest of the block is updated by all threads -- forward:
// one row per thread, its 64 loads in flight together; backward: one warp per eight columns, lanes over the rows, all
// loads of the warp in flight together.  (Round 1's 64-thread tile kernel with two block barriers per step took 21 us per
// tile, and a 64-wide step was two dependent launches: 45 of the 63 ms of a 128^3 solve, CHOL_SOLVE_TIMES.)
constexpr int kSolveBlockThreads = 256;
template <bool BWD>
__global__ void __launch_bounds__(kSolveBlockThreads) solve_block(const SolveTile *__restrict__ descs, const double *__restrict__ fac,
                                                                   double *__restrict__ x) {
  __shared__ double Ls[kSolveNB][kSolveNB + 1];  // Ls[c][r] = L(r, c) of the current tile, r >= c
  __shared__ double xb[kSolveWB];
  const SolveTile d = descs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, w = d.nb;
  const double *__restrict__ Lg = fac + d.l_off;
  xb[tid] = tid < w ? x[d.x0 + tid] : 0.0;
  const int nt = (w + kSolveNB - 1) / kSolveNB;
  for (int step = 0; step < nt; step++) {
    const int t = BWD ? nt - 1 - step : step;
    const int d0 = t * kSolveNB, dw = min(kSolveNB, w - d0), below = w - d0 - kSolveNB;  // rows of the block below this tile
    {
      const int r = tid & (kSolveNB - 1), cq = tid >> 6;  // thread = row, every fourth column
      double tv[kSolveNB / 4];
#pragma unroll
      for (int u = 0; u < kSolveNB / 4; u++) {
        const int c = 4 * u + cq;
        tv[u] = (r < dw && c <= r) ? Lg[d0 + r + (size_t)(d0 + c) * d.ld] : (r == c ? 1.0 : 0.0);
      }
#pragma unroll
      for (int u = 0; u < kSolveNB / 4; u++) Ls[4 * u + cq][r] = tv[u];
    }
    if (BWD && below > 0) {  // xb[c] -= sum_{r below} L(r, c) xb[r], c in the tile: one warp per column, eight columns per warp
      __syncthreads();       // (xb of the previous step is final)
      constexpr int kQ = kSolveNB / (kSolveBlockThreads / 32), kI = (kSolveWB - kSolveNB) / 32;  // columns per warp, rows per lane
      double lv[kQ][kI];     // all loads of the warp's columns in flight before the first is used
#pragma unroll
      for (int q = 0; q < kQ; q++) {
        const int c = warp + q * (kSolveBlockThreads / 32);
        const double *__restrict__ col = Lg + d0 + kSolveNB + (size_t)(d0 + c) * d.ld;
#pragma unroll
        for (int i = 0; i < kI; i++) lv[q][i] = (c < dw && lane + 32 * i < below) ? col[lane + 32 * i] : 0.0;
      }
      double xr[kI];
#pragma unroll
      for (int i = 0; i < kI; i++) xr[i] = lane + 32 * i < below ? xb[d0 + kSolveNB + lane + 32 * i] : 0.0;
#pragma unroll
      for (int q = 0; q < kQ; q++) {
        double a = 0;
#pragma unroll
        for (int i = 0; i < kI; i++) a = fma(lv[q][i], xr[i], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        const int c = warp + q * (kSolveBlockThreads / 32);
        if (lane == 0 && c < dw) xb[d0 + c] -= a;
      }
    }
    __syncthreads();
    if (warp == 0) {
      double x0 = xb[d0 + lane], x1 = xb[d0 + lane + 32];
      const double rd0 = 1.0 / Ls[lane][lane], rd1 = 1.0 / Ls[lane + 32][lane + 32];
      if (!BWD) {
#pragma unroll 8
        for (int k = 0; k < kSolveNB; k++) {
          const double xk = __shfl_sync(0xffffffffu, k < 32 ? x0 * rd0 : x1 * rd1, k & 31);
          if (lane == k) x0 = xk;
          if (lane + 32 == k) x1 = xk;
          if (lane > k) x0 = fma(-Ls[k][lane], xk, x0);
          if (lane + 32 > k) x1 = fma(-Ls[k][lane + 32], xk, x1);
        }
      } else {
#pragma unroll 8
        for (int k = kSolveNB - 1; k >= 0; k--) {
          const double xk = __shfl_sync(0xffffffffu, k < 32 ? x0 * rd0 : x1 * rd1, k & 31);
          if (lane == k) x0 = xk;
          if (lane + 32 == k) x1 = xk;
          if (lane < k) x0 = fma(-Ls[lane][k], xk, x0);  // L(k, i), i < k
          if (lane + 32 < k) x1 = fma(-Ls[lane + 32][k], xk, x1);
        }
      }
      if (lane < dw) xb[d0 + lane] = x0;
      if (lane + 32 < dw) xb[d0 + lane + 32] = x1;
    }
    __syncthreads();
    if (!BWD && tid < below) {  // row d0 + 64 + tid of the tile column, all 64 loads in flight (issuing them before the tile
                                // solve to overlap it was slower: 8.5 vs 4.7 ms of a 128^3 sweep, the registers they pin)
      const double *__restrict__ row = Lg + d0 + kSolveNB + tid + (size_t)d0 * d.ld;
      double v[kSolveNB];
#pragma unroll
      for (int c = 0; c < kSolveNB; c++) v[c] = c < dw ? row[(size_t)c * d.ld] : 0.0;
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int c = 0; c < kSolveNB; c += 2) a0 = fma(v[c], xb[d0 + c], a0), a1 = fma(v[c + 1], xb[d0 + c + 1], a1);
      xb[d0 + kSolveNB + tid] -= a0 + a1;
    }
    __syncthreads();
  }
  if (tid < w) x[d.x0 + tid] = xb[tid];
}

// forward: y[y0 + r] -= sum_c P[r, c] x[x0 + c], c < nb <= 256; one row per thread, 128-row slabs, 64 loads of a row in
// flight at a time
__global__ void __launch_bounds__(kSolveSlab) solve_gemv_fwd(const SolveGemv *__restrict__ descs, const TileRef *__restrict__ tiles,
                                                             const double *__restrict__ fac, double *__restrict__ x) {
  __shared__ double xs[kSolveWB];
  const TileRef tl = tiles[blockIdx.x];
  const SolveGemv d = descs[tl.prob];
  const int slab = (int)tl.tr | ((int)tl.tc << 16), tid = threadIdx.x;
  for (int i = tid; i < kSolveWB; i += kSolveSlab) xs[i] = i < d.nb ? x[d.x0 + i] : 0.0;
  const int r = slab * kSolveSlab + tid;
  const bool live = r < d.rows;
  const double *__restrict__ Pp = fac + d.p_off + (live ? r : 0);
  __syncthreads();
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int cb = 0; cb < d.nb; cb += kSolveNB) {
    double v[kSolveNB];
#pragma unroll
    for (int c = 0; c < kSolveNB; c++) v[c] = (live && cb + c < d.nb) ? Pp[(size_t)(cb + c) * d.ld] : 0.0;
#pragma unroll
    for (int c = 0; c < kSolveNB; c += 4) {
      a0 = fma(v[c], xs[cb + c], a0), a1 = fma(v[c + 1], xs[cb + c + 1], a1);
      a2 = fma(v[c + 2], xs[cb + c + 2], a2), a3 = fma(v[c + 3], xs[cb + c + 3], a3);
    }
  }
  if (live) x[d.y0 + r] -= (a0 + a1) + (a2 + a3);
}

// backward: y[y0 + c] -= sum_{r < nb} P[r, c] x[x0 + r]; one warp per column, eight columns per CTA
__global__ void __launch_bounds__(kSolveColG * 32) solve_gemv_bwd(const SolveGemv *__restrict__ descs, const TileRef *__restrict__ tiles,
                                                                  const double *__restrict__ fac, double *__restrict__ x) {
  __shared__ double xs[kSolveWB];
  const TileRef tl = tiles[blockIdx.x];
  const SolveGemv d = descs[tl.prob];
  const int grp = (int)tl.tr | ((int)tl.tc << 16), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < d.nb) xs[tid] = x[d.x0 + tid];
  __syncthreads();
  const int c = grp * kSolveColG + warp;
  if (c >= d.rows) return;
  const double *__restrict__ Pp = fac + d.p_off + (size_t)c * d.ld;
  double a = 0;
  for (int r = lane; r < d.nb; r += 32) a += Pp[r] * xs[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if (lane == 0) x[d.y0 + c] -= a;
}

// ancestors pull: x[y0 + r] -= sum over contributors (fixed order) of P_s[seg rows, :] x_s.  A tile is 128 rows of a
// destination and a range of the contributors' columns laid end to end; eight columns are loaded at a time.
__global__ void __launch_bounds__(kSolveSlab) solve_pull(const PullDest *__restrict__ dests, const PullContrib *__restrict__ contribs,
                                                         const PullTile *__restrict__ tiles, const double *__restrict__ fac,
                                                         double *__restrict__ x, double *__restrict__ scratch) {
  __shared__ double xs[kSolveSlab];
  const PullTile tl = tiles[blockIdx.x];
  const PullDest d = dests[tl.dest];
  const int tid = threadIdx.x;
  const int r = tl.slab * kSolveSlab + tid;
  const bool live = r < d.rows;
  double acc = 0;
  int kpos = 0;  // first column of the current contributor in the destination's column range
  for (int ci = 0; ci < d.ccnt && kpos < tl.kend; ci++) {
    const PullContrib cb = contribs[d.cbeg + ci];
    const int c0 = max(tl.kbeg - kpos, 0), c1 = min(tl.kend - kpos, cb.K);  // this tile's columns of the contributor
    kpos += cb.K;
    if (c1 <= c0) continue;
    const double *__restrict__ Pp = fac + cb.p_off + (live ? r : 0);
    for (int k0 = c0; k0 < c1; k0 += kSolveSlab) {
      const int kn = min(kSolveSlab, c1 - k0);
      __syncthreads();
      if (tid < kn) xs[tid] = x[cb.x0 + k0 + tid];
      __syncthreads();
      if (live) {
        double a0 = 0, a1 = 0;
        int k = 0;
        for (; k + 7 < kn; k += 8) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = Pp[(size_t)(k0 + k + u) * cb.ld];
#pragma unroll
          for (int u = 0; u < 8; u += 2) a0 = fma(v[u], xs[k + u], a0), a1 = fma(v[u + 1], xs[k + u + 1], a1);
        }
        for (; k < kn; k++) a0 = fma(Pp[(size_t)(k0 + k) * cb.ld], xs[k], a0);
        acc += a0 + a1;
      }
    }
  }
  if (!live) return;
  if (tl.slot < 0) x[d.y0 + r] -= acc;
  else scratch[(size_t)tl.slot * kSolveSlab + tid] = acc;
}
// the partial sums of the split tiles, added up in slot order
__global__ void __launch_bounds__(kSolveSlab) solve_pull_sum(const PullSum *__restrict__ sums, const double *__restrict__ scratch, double *__restrict__ x) {
  const PullSum d = sums[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid >= d.rows) return;
  double acc = 0;
  for (int p = 0; p < d.nparts; p++) acc += scratch[(size_t)(d.slot0 + p) * kSolveSlab + tid];
  x[d.y0 + tid] -= acc;
}

// backward gather: x[x0 + c] -= sum_r P[r, c] x[rowmap[r]] over the off-diagonal rows of the panel
__global__ void __launch_bounds__(kSolveColG * 32) solve_gather(const GatherDesc *__restrict__ descs, const TileRef *__restrict__ tiles,
                                                                const int *__restrict__ rowmap, const double *__restrict__ fac,
                                                                double *__restrict__ x) {
  const TileRef tl = tiles[blockIdx.x];
  const GatherDesc d = descs[tl.prob];
  const int grp = (int)tl.tr | ((int)tl.tc << 16), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = grp * kSolveColG + warp;
  if (c >= d.n) return;
  const double *__restrict__ Pp = fac + d.p_off + (size_t)c * d.ld;
  const int *__restrict__ mp = rowmap + d.map_off;
  double a0 = 0, a1 = 0;
  int r = lane;
  for (; r + 32 < d.nrows; r += 64) {
    const int m0 = mp[r], m1 = mp[r + 32];
    a0 += m0 >= 0 ? Pp[r] * x[m0] : 0.0;
    a1 += m1 >= 0 ? Pp[r + 32] * x[m1] : 0.0;
  }
  if (r < d.nrows) {
    const int m0 = mp[r];
    a0 += m0 >= 0 ? Pp[r] * x[m0] : 0.0;
  }
  double a = a0 + a1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if (lane == 0) x[d.x0 + c] -= a;
}

}  // namespace chb
