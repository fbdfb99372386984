// Residual check on the factor as it sits in HBM: Z = L (L^T W) for k Rademacher columns W, so that the host
// only has to compare Z with A W (the randomized estimator ||(A - L L^T) W||_F / ||A W||_F; the
// reference's verify.check_matrix is dense O(n^2), verify.py:277-300).  Every stored entry of the panels a rank
// reports is read twice, coalesced; nothing is copied to the host but the n x k result.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace chb {

constexpr int kResK = 4;                         // at most this many probe columns
constexpr int kResSlab = 128, kResChunk = 256;   // rows per CTA / columns per CTA of res_ly
constexpr int kResColG = 8;                      // columns (warps) per CTA of res_ltw

struct ResPanel {
  int64_t off, map_off;  // panel offset in the factor, offset of its off-diagonal rows in rowmap
  int ld, n, rows, r0, start, pad;
};
struct ResTile {
  int panel, a, b, pad;  // res_ltw: (panel, first column); res_ly: (panel, row slab, column chunk)
};

__host__ __device__ inline uint64_t res_mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
// entry (permuted row i, probe column q) of W
__host__ __device__ inline double res_w(uint64_t seed, int64_t i, int q) { return (res_mix64(seed ^ res_mix64((uint64_t)i * kResK + q)) & 1) ? 1.0 : -1.0; }

// Y[start + c, :] = sum_r L[r, c] W[grow(r), :], one warp per column
__global__ void __launch_bounds__(kResColG * 32) res_ltw(const ResPanel *__restrict__ panels, const ResTile *__restrict__ tiles,
                                                         const int *__restrict__ rowmap, const double *__restrict__ fac, int k, uint64_t seed,
                                                         double *__restrict__ y) {
  const ResTile tl = tiles[blockIdx.x];
  const ResPanel p = panels[tl.panel];
  const int lane = threadIdx.x & 31, c = tl.a + (threadIdx.x >> 5);
  if (c >= p.n) return;
  const double *__restrict__ col = fac + p.off + (size_t)c * p.ld;
  double acc[kResK] = {0, 0, 0, 0};
  for (int r = c + lane; r < p.n; r += 32) {
    const double v = col[r];
#pragma unroll
    for (int q = 0; q < kResK; q++)
      if (q < k) acc[q] += v * res_w(seed, p.start + r, q);
  }
  const int *__restrict__ mp = rowmap + p.map_off;
  for (int r = p.r0 + lane; r < p.rows; r += 32) {
    const int g = mp[r - p.r0];
    if (g < 0) continue;
    const double v = col[r];
#pragma unroll
    for (int q = 0; q < kResK; q++)
      if (q < k) acc[q] += v * res_w(seed, g, q);
  }
#pragma unroll
  for (int q = 0; q < kResK; q++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_down_sync(0xffffffffu, acc[q], o);
    if (lane == 0 && q < k) y[(size_t)(p.start + c) * kResK + q] = acc[q];
  }
}

// Z[grow(r), :] += sum_{c in chunk} L[r, c] Y[start + c, :], one row per thread
__global__ void __launch_bounds__(kResSlab) res_ly(const ResPanel *__restrict__ panels, const ResTile *__restrict__ tiles,
                                                   const int *__restrict__ rowmap, const double *__restrict__ fac, int k, const double *__restrict__ y,
                                                   double *__restrict__ z) {
  __shared__ double ys[kResChunk][kResK];
  const ResTile tl = tiles[blockIdx.x];
  const ResPanel p = panels[tl.panel];
  const int c0 = tl.b * kResChunk, c1 = min(p.n, c0 + kResChunk);
  for (int i = threadIdx.x; i < (c1 - c0) * kResK; i += kResSlab) ys[i / kResK][i % kResK] = y[(size_t)(p.start + c0) * kResK + i];
  __syncthreads();
  const int r = tl.a * kResSlab + threadIdx.x;
  if (r >= p.rows || (r >= p.n && r < p.r0)) return;
  const int g = r < p.n ? p.start + r : rowmap[p.map_off + r - p.r0];
  if (g < 0) return;
  const int cend = r < p.n ? min(c1, r + 1) : c1;  // pivot block: lower triangle only
  const double *__restrict__ row = fac + p.off + r;
  double acc[kResK] = {0, 0, 0, 0};
  for (int c = c0; c < cend; c++) {
    const double v = row[(size_t)c * p.ld];
#pragma unroll
    for (int q = 0; q < kResK; q++) acc[q] += v * ys[c - c0][q];
  }
  if (cend > c0)
    for (int q = 0; q < k; q++) atomicAdd(&z[(size_t)g * kResK + q], acc[q]);
}

}  // namespace chb
