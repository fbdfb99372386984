// Triangular-solve kernels on the panel storage (reference: dtrsv / dgemv tasks, blas.rg:217-290, driven
// by mmat.rg:1394-1479).  HBM-bound work: every factor entry is read once per sweep, coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "solve.h"

namespace chb {

constexpr int kSolveNB = 64, kSolveSlab = 128, kSolveColG = 8;

__global__ void permute_in_kernel(const double *__restrict__ b, const int *__restrict__ perm, int n, double *__restrict__ x) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) x[p] = b[perm[p]];  // fill_b, mmat.rg:769-783
}
__global__ void permute_out_kernel(const double *__restrict__ x, const int *__restrict__ perm, int n, double *__restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) out[perm[p]] = x[p];  // mmat.rg:1483-1491
}

// pivot tile: forward L y = b (BWD = false) or backward L^T x = y (BWD = true), 64 threads
template <bool BWD>
__global__ void __launch_bounds__(kSolveNB) solve_tile(const SolveTile *__restrict__ descs, const double *__restrict__ fac, double *__restrict__ x) {
  __shared__ double Ls[kSolveNB][kSolveNB + 1];
  __shared__ double xb[kSolveNB];
  const SolveTile d = descs[blockIdx.x];
  const int i = threadIdx.x, nb = d.nb;
  const double *__restrict__ Lg = fac + d.l_off;
  for (int c0 = 0; c0 < nb; c0 += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = (i < nb && c0 + u <= i) ? Lg[i + (size_t)(c0 + u) * d.ld] : 0.0;
#pragma unroll
    for (int u = 0; u < 8; u++) Ls[i][c0 + u] = v[u];
  }
  xb[i] = i < nb ? x[d.x0 + i] : 0.0;
  __syncthreads();
  if (!BWD) {
    for (int k = 0; k < nb; k++) {
      if (i == k) xb[k] /= Ls[k][k];
      __syncthreads();
      if (i > k && i < nb) xb[i] -= Ls[i][k] * xb[k];
      __syncthreads();
    }
  } else {
    for (int k = nb - 1; k >= 0; k--) {
      if (i == k) xb[k] /= Ls[k][k];
      __syncthreads();
      if (i < k) xb[i] -= Ls[k][i] * xb[k];
      __syncthreads();
    }
  }
  if (i < nb) x[d.x0 + i] = xb[i];
}

// forward: y[y0 + r] -= sum_c P[r, c] x[x0 + c]; one row per thread, 128-row slabs
__global__ void __launch_bounds__(kSolveSlab) solve_gemv_fwd(const SolveGemv *__restrict__ descs, const TileRef *__restrict__ tiles,
                                                             const double *__restrict__ fac, double *__restrict__ x) {
  __shared__ double xs[kSolveNB];
  const TileRef tl = tiles[blockIdx.x];
  const SolveGemv d = descs[tl.prob];
  const int slab = (int)tl.tr | ((int)tl.tc << 16), tid = threadIdx.x;
  if (tid < d.nb) xs[tid] = x[d.x0 + tid];
  __syncthreads();
  const int r = slab * kSolveSlab + tid;
  if (r >= d.rows) return;
  const double *__restrict__ Pp = fac + d.p_off + r;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int c = 0;
  for (; c + 3 < d.nb; c += 4) {
    a0 += Pp[(size_t)c * d.ld] * xs[c], a1 += Pp[(size_t)(c + 1) * d.ld] * xs[c + 1];
    a2 += Pp[(size_t)(c + 2) * d.ld] * xs[c + 2], a3 += Pp[(size_t)(c + 3) * d.ld] * xs[c + 3];
  }
  for (; c < d.nb; c++) a0 += Pp[(size_t)c * d.ld] * xs[c];
  x[d.y0 + r] -= (a0 + a1) + (a2 + a3);
}

// backward: y[y0 + c] -= sum_{r < nb} P[r, c] x[x0 + r]; one warp per column, eight columns per CTA
__global__ void __launch_bounds__(kSolveColG * 32) solve_gemv_bwd(const SolveGemv *__restrict__ descs, const TileRef *__restrict__ tiles,
                                                                  const double *__restrict__ fac, double *__restrict__ x) {
  __shared__ double xs[kSolveNB];
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

// ancestors pull: x[y0 + r] -= sum over contributors (fixed order) of P_s[seg rows, :] x_s
__global__ void __launch_bounds__(kSolveSlab) solve_pull(const PullDest *__restrict__ dests, const PullContrib *__restrict__ contribs,
                                                         const TileRef *__restrict__ tiles, const double *__restrict__ fac,
                                                         double *__restrict__ x) {
  __shared__ double xs[kSolveSlab];
  const TileRef tl = tiles[blockIdx.x];
  const PullDest d = dests[tl.prob];
  const int slab = (int)tl.tr | ((int)tl.tc << 16), tid = threadIdx.x;
  const int r = slab * kSolveSlab + tid;
  const bool live = r < d.rows;
  double acc = 0;
  for (int ci = 0; ci < d.ccnt; ci++) {
    const PullContrib cb = contribs[d.cbeg + ci];
    const double *__restrict__ Pp = fac + cb.p_off + (live ? r : 0);
    for (int k0 = 0; k0 < cb.K; k0 += kSolveSlab) {
      const int kn = min(kSolveSlab, cb.K - k0);
      __syncthreads();
      if (tid < kn) xs[tid] = x[cb.x0 + k0 + tid];
      __syncthreads();
      if (live) {
        double a0 = 0, a1 = 0;
        int k = 0;
        for (; k + 1 < kn; k += 2) a0 += Pp[(size_t)(k0 + k) * cb.ld] * xs[k], a1 += Pp[(size_t)(k0 + k + 1) * cb.ld] * xs[k + 1];
        if (k < kn) a0 += Pp[(size_t)(k0 + k) * cb.ld] * xs[k];
        acc += a0 + a1;
      }
    }
  }
  if (live) x[d.y0 + r] -= acc;
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
