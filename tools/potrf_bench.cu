// Timing harness for pivot-tile kernels: one 64 x 64 tile per launch (the root-of-the-tree situation), hot (same
// kernel back to back) and cold (a large unrelated kernel in between).  Variants are compared against a kernel that
// only loads and stores the tile.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/potrf_bench tools/potrf_bench.cu && tools/potrf_bench
#include <cstdio>
#include <vector>
#include <cmath>
#include "../cholesky_b200/csrc/kernels.cuh"
using namespace chb;

__global__ void __launch_bounds__(256) tile_copy_only(const PotrfDesc *descs, double *fac, int *info) {
  __shared__ double T[kNB][kNB + 1];
  const PotrfDesc d = descs[blockIdx.x];
  double *A = fac + d.off;
  const int nb = d.nb, tid = threadIdx.x;
  const int i = tid & 63, cg = tid / 64;
  for (int h = 0; h < 2; h++) {
    double v[8];
    for (int u = 0; u < 8; u++) { const int c = cg * 16 + h * 8 + u; v[u] = (i < nb && c <= i) ? A[i + (size_t)c * d.ld] : 0.0; }
    for (int u = 0; u < 8; u++) T[i][cg * 16 + h * 8 + u] = v[u];
  }
  __syncthreads();
  for (int u = 0; u < 16; u++) { const int c = cg * 16 + u; if (i < nb && c <= i) A[i + (size_t)c * d.ld] = T[i][c] * 1.0000001; }
}

// rolled variant: the panel column being eliminated is always register 0, registers rotate after every step
__global__ void __launch_bounds__(256) potrf_tile_rot(const PotrfDesc *__restrict__ descs, double *__restrict__ fac, int *__restrict__ info) {
  __shared__ double T[kNB][kNB + 1];
  const PotrfDesc d = descs[blockIdx.x];
  double *__restrict__ A = fac + d.off;
  const int nb = d.nb, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const int i = tid & (kNB - 1), cg = tid / kNB;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { const int c = cg * 16 + h * 8 + u; v[u] = (i < nb && c <= i) ? A[i + (size_t)c * d.ld] : ((c == i && i >= nb) ? 1.0 : 0.0); }
#pragma unroll
      for (int u = 0; u < 8; u++) T[i][cg * 16 + h * 8 + u] = v[u];
    }
  }
  __syncthreads();
  for (int k0 = 0; k0 < nb; k0 += 8) {
    if (warp == 0) {
      double x0[8], x1[8];
#pragma unroll
      for (int j = 0; j < 8; j++) x0[j] = T[lane][k0 + j], x1[j] = T[lane + 32][k0 + j];
      const bool hi = k0 >= 32;
#pragma unroll 1
      for (int kk = 0; kk < 8; kk++) {
        const int k = k0 + kk;
        double dk = __shfl_sync(0xffffffffu, hi ? x1[0] : x0[0], k & 31);
        if (!(dk > 0.0)) { if (lane == 0) atomicMin(info, d.col0 + k + 1); dk = 1.0; }
        const double r = rsqrt(dk);
        const double c0 = x0[0] * r, c1 = x1[0] * r;
        if (lane >= k) T[lane][k] = c0;
        if (lane + 32 >= k) T[lane + 32][k] = c1;
#pragma unroll
        for (int j = 1; j < 8; j++) {   // columns past the panel hold zeros rotated in: their updates are no-ops
          const double ljk = __shfl_sync(0xffffffffu, hi ? c1 : c0, (k + j) & 31);
          x0[j - 1] = fma(-c0, ljk, x0[j]);
          x1[j - 1] = fma(-c1, ljk, x1[j]);
        }
        x0[7] = 0.0, x1[7] = 0.0;
      }
    }
    __syncthreads();
    const int e0 = k0 + 8, rem = nb - e0;
    if (rem > 0) {
      for (int idx = tid; idx < rem * rem; idx += 256) {
        const int i = e0 + idx % rem, j = e0 + idx / rem;
        if (i < j) continue;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) { s0 = fma(T[i][k0 + k], T[j][k0 + k], s0); s1 = fma(T[i][k0 + k + 1], T[j][k0 + k + 1], s1); }
        T[i][j] -= s0 + s1;
      }
    }
    __syncthreads();
  }
  {
    const int i = tid & (kNB - 1), cg = tid / kNB;
#pragma unroll
    for (int u = 0; u < 16; u++) { const int c = cg * 16 + u; if (i < nb && c <= i && c < nb) A[i + (size_t)c * d.ld] = T[i][c]; }
  }
}

__global__ void thrash(double *buf, size_t n) {  // something else runs on every SM in between
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double x = 0;
  for (int k = 0; k < 64; k++) x = fma(x, 1.0000001, buf[(i + (size_t)k * 4096) % n]);
  buf[i % n] = x;
}

typedef void (*kern_t)(const PotrfDesc *, double *, int *);
int main() {
  const int ld = 4096, nb = 64;
  std::vector<double> h((size_t)ld * nb), ref;
  for (int c = 0; c < nb; c++) for (int r = 0; r < nb; r++) h[r + (size_t)c * ld] = (r == c) ? 70.0 : 1.0 / (1 + abs(r - c));
  double *d, *big; int *info; PotrfDesc *desc;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&big, (size_t)1 << 28); cudaMemset(big, 0, (size_t)1 << 28);
  cudaMalloc(&info, 8); cudaMalloc(&desc, sizeof(PotrfDesc));
  PotrfDesc pd{0, ld, nb, 0, 0};
  cudaMemcpy(desc, &pd, sizeof pd, cudaMemcpyHostToDevice);
  const char *names[3] = {"copy only", "potrf_tile (panel steps unrolled)", "potrf_tile_rot (rolled, rotating registers)"};
  kern_t ks[3] = {tile_copy_only, potrf_tile, potrf_tile_rot};
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  std::vector<double> out0(h.size());
  for (int v = 0; v < 3; v++) {
    float hot = 0, cold = 0;
    for (int mode = 0; mode < 2; mode++) {
      float tot = 0; const int reps = 20;
      for (int it = 0; it < reps + 2; it++) {
        cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        if (mode == 1) thrash<<<148 * 8, 256>>>(big, ((size_t)1 << 25));
        cudaEventRecord(a); ks[v]<<<1, 256>>>(desc, d, info); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it >= 2) tot += ms;
      }
      (mode ? cold : hot) = tot / reps * 1e3f;
    }
    std::vector<double> out(h.size());
    cudaMemcpy(out.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    double diff = 0;
    if (v == 1) out0 = out;
    if (v == 2) for (int c = 0; c < nb; c++) for (int r = c; r < nb; r++) diff = fmax(diff, fabs(out[r + (size_t)c * ld] - out0[r + (size_t)c * ld]));
    printf("%-46s hot %.2f us   cold %.2f us   L[63][63] %.15g  max diff vs unrolled %.2e\n", names[v], hot, cold, out[63 + (size_t)63 * ld], diff);
  }
  return 0;
}
