// Timing harness for the pivot-tile factorization: one 64 x 64 tile per launch (the root-of-the-tree situation),
// back to back.  Compares potrf_smem (look-ahead inside the tile, csrc/kernels.cuh) with a kernel that only loads
// and stores the tile, and checks the factor against a host Cholesky.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/potrf_bench tools/potrf_bench.cu && tools/potrf_bench
#include <cmath>
#include <cstdio>
#include <vector>
#include "../cholesky_b200/csrc/kernels.cuh"
using namespace chb;

template <bool FACTOR>
__global__ void __launch_bounds__(kPanelThreads) tile_kernel(double *A, int ld, int nb, int *info) {
  __shared__ double T[kNB * kLdx];
  const int tid = threadIdx.x, r = tid & 63;
  for (int c = tid >> 6; c < 64; c += 4) T[c * kLdx + r] = (r < nb && c < nb) ? A[r + (size_t)c * ld] : (r == c ? 1.0 : 0.0);
  __syncthreads();
  if (FACTOR) potrf_smem(T, nb, 0, info);
  for (int c = tid >> 6; c < nb; c += 4)
    if (r < nb && r >= c) A[r + (size_t)c * ld] = T[c * kLdx + r];
}

int main() {
  const int ld = 4096;
  for (int nb : {64, 61, 40, 8}) {
    std::vector<double> h((size_t)ld * 64, 0.0), ref(64 * 64, 0.0);
    for (int c = 0; c < nb; c++)
      for (int r = 0; r < nb; r++) h[r + (size_t)c * ld] = ref[r + 64 * c] = (r == c) ? 70.0 : 1.0 / (1 + abs(r - c));
    for (int k = 0; k < nb; k++) {  // host Cholesky (lower, column-major)
      ref[k + 64 * k] = sqrt(ref[k + 64 * k]);
      for (int i = k + 1; i < nb; i++) ref[i + 64 * k] /= ref[k + 64 * k];
      for (int j = k + 1; j < nb; j++)
        for (int i = j; i < nb; i++) ref[i + 64 * j] -= ref[i + 64 * k] * ref[j + 64 * k];
    }
    double *d;
    int *info;
    cudaMalloc(&d, h.size() * 8);
    cudaMalloc(&info, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float t[2] = {0, 0};
    double worst = 0;
    for (int v = 0; v < 2; v++) {
      const int reps = 20;
      for (int it = 0; it < reps + 2; it++) {
        cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(a);
        if (v) tile_kernel<true><<<1, kPanelThreads>>>(d, ld, nb, info);
        else tile_kernel<false><<<1, kPanelThreads>>>(d, ld, nb, info);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (it >= 2) t[v] += ms * 1e3f / reps;
      }
      if (v) {
        std::vector<double> out(h.size());
        cudaMemcpy(out.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
        for (int c = 0; c < nb; c++)
          for (int r = c; r < nb; r++) worst = fmax(worst, fabs(out[r + (size_t)c * ld] - ref[r + 64 * c]));
      }
    }
    printf("nb %2d: load + store only %.2f us, with potrf_smem %.2f us (factorization %.2f us), max |error| vs host %.2e\n", nb, t[0], t[1], t[1] - t[0], worst);
    cudaFree(d), cudaFree(info);
  }
  return 0;
}
