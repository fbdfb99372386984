// Multi-GPU microbenchmark of the row push (push_rects): how the flag protocol should be fenced.  Device 0 pushes a
// rectangle of its buffer into the same offset of every peer's buffer and raises a flag word there.
//   v0  every CTA: stores, __threadfence_system, count; last CTA: __threadfence_system, flags     (round-2 first version)
//   v1  every CTA: stores, __threadfence (gpu scope), count; last CTA: __threadfence_system, flags
//   v2  stores only; a second one-warp kernel fences (system) and raises the flags
// Timing: CUDA events on device 0 around each variant.  Visibility: device 1 runs a kernel per iteration that waits
// for its flag and compares every element with the iteration's value; a mismatch is a stale read.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/push_bench tools/push_bench.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
constexpr int kMaxPeers = 8;
struct Peers { double *fac[kMaxPeers]; unsigned long long *flags[kMaxPeers]; int n, rank; };
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory"); return v; }

template <int V>
__global__ void __launch_bounds__(256) push(double *fac, Peers peers, unsigned mask, int rows, int cols, int ld, int col_groups, unsigned long long value, unsigned *counter) {
  const int cg = blockIdx.x % col_groups, cpg = (cols + col_groups - 1) / col_groups, rows2 = rows / 2;
  for (int c = cg * cpg + (int)threadIdx.x / 128; c < min(cols, (cg + 1) * cpg); c += 2)
    for (int r2 = threadIdx.x % 128 + 128 * (blockIdx.x / col_groups); r2 < rows2; r2 += 128 * (gridDim.x / col_groups)) {
      const size_t o = 2 * r2 + (size_t)c * ld;
      const double2 v = *reinterpret_cast<const double2 *>(fac + o);
#pragma unroll
      for (int p = 0; p < kMaxPeers; p++)
        if (p < peers.n && ((mask >> p) & 1u)) *reinterpret_cast<double2 *>(peers.fac[p] + o) = v;
    }
  if (V == 2) return;
  if (V == 0) __threadfence_system(); else __threadfence();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *counter = 0;
  __threadfence_system();
  const int p = threadIdx.x;
  if (p < peers.n && ((mask >> p) & 1u)) st_release_sys(peers.flags[p] + peers.rank, value);
}
__global__ void signal_only(Peers peers, unsigned mask, unsigned long long value) {
  __threadfence_system();
  const int p = threadIdx.x;
  if (p < peers.n && ((mask >> p) & 1u)) st_release_sys(peers.flags[p] + peers.rank, value);
}
__global__ void fill(double *fac, int rows, int cols, int ld, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)rows * cols; i += (size_t)gridDim.x * blockDim.x) fac[i % rows + (i / rows) * ld] = v;
}
__global__ void check(const double *fac, const unsigned long long *flag, unsigned long long value, int rows, int cols, int ld, double want, unsigned long long *bad) {
  if (threadIdx.x == 0) while (ld_acquire_sys(flag) < value) {}
  __syncthreads();
  unsigned long long b = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)rows * cols; i += (size_t)gridDim.x * blockDim.x)
    b += (fac[i % rows + (i / rows) * ld] != want);
  if (b) atomicAdd(bad, b);
}
template <int V>
float run(int ndev, std::vector<double *> &buf, std::vector<unsigned long long *> &flg, Peers peers, unsigned mask, int rows, int cols, int ld, int ctas, int cg,
          unsigned *counter, unsigned long long *bad, unsigned long long &seq, int iters, bool verify) {
  cudaSetDevice(0);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  float tot = 0;
  for (int it = 0; it < iters; it++) {
    seq++;
    cudaSetDevice(0);
    fill<<<148, 256>>>(buf[0], rows, cols, ld, (double)seq);
    if (verify) { cudaSetDevice(1); check<<<64, 256>>>(buf[1], flg[1] + 0, seq, rows, cols, ld, (double)seq, bad); cudaSetDevice(0); }
    cudaEventRecord(a);
    push<V><<<ctas, 256>>>(buf[0], peers, mask, rows, cols, ld, cg, seq, counter);
    if (V == 2) signal_only<<<1, 32>>>(peers, mask, seq);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it >= 2) tot += ms;
    if (verify) { cudaSetDevice(1); cudaDeviceSynchronize(); }
  }
  return tot / (iters - 2) * 1e3f;
}
int main() {
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (ndev < 2) { printf("needs >= 2 GPUs\n"); return 0; }
  if (ndev > kMaxPeers) ndev = kMaxPeers;
  const int ld = 16384; const size_t doubles = (size_t)ld * 256;
  std::vector<double *> buf(ndev); std::vector<unsigned long long *> flg(ndev);
  for (int d = 0; d < ndev; d++) {
    cudaSetDevice(d);
    for (int e = 0; e < ndev; e++) if (e != d) cudaDeviceEnablePeerAccess(e, 0);
    cudaMalloc(&buf[d], doubles * 8); cudaMemset(buf[d], 0, doubles * 8);
    cudaMalloc(&flg[d], 64 * 8); cudaMemset(flg[d], 0, 64 * 8);
  }
  cudaSetDevice(0);
  unsigned *counter; cudaMalloc(&counter, 4); cudaMemset(counter, 0, 4);
  unsigned long long *bad; cudaSetDevice(1); cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8); cudaSetDevice(0);
  Peers peers; peers.n = ndev; peers.rank = 0;
  for (int d = 0; d < ndev; d++) peers.fac[d] = buf[d], peers.flags[d] = flg[d];
  unsigned long long seq = 0;
  struct Case { const char *name; int rows, cols, ctas, cg; } cases[] = {{"diag block 256 x 256", 256, 256, 64, 64}, {"rows 1024 x 256", 1024, 256, 256, 64}, {"rows 6144 x 256", 6144, 256, 296, 8}};
  for (int npeer : {1, 3, 7}) {
    if (npeer >= ndev) continue;
    unsigned mask = 0; for (int p = 1; p <= npeer; p++) mask |= 1u << p;
    for (auto &cs : cases) {
      float t0 = run<0>(ndev, buf, flg, peers, mask, cs.rows, cs.cols, ld, cs.ctas, cs.cg, counter, bad, seq, 30, false);
      float t1 = run<1>(ndev, buf, flg, peers, mask, cs.rows, cs.cols, ld, cs.ctas, cs.cg, counter, bad, seq, 30, false);
      float t2 = run<2>(ndev, buf, flg, peers, mask, cs.rows, cs.cols, ld, cs.ctas, cs.cg, counter, bad, seq, 30, false);
      printf("%d peer(s), %-22s %8.1f KB each:  v0 %.1f us   v1 %.1f us   v2 %.1f us\n", npeer, cs.name, cs.rows * cs.cols * 8 / 1024.0, t0, t1, t2);
    }
  }
  // visibility stress of v1 and v2 against device 1
  for (int v = 1; v <= 2; v++) {
    cudaSetDevice(1); cudaMemset(bad, 0, 8); cudaSetDevice(0);
    unsigned mask = 2;
    if (v == 1) run<1>(ndev, buf, flg, peers, mask, 6144, 256, ld, 296, 8, counter, bad, seq, 300, true);
    else run<2>(ndev, buf, flg, peers, mask, 6144, 256, ld, 296, 8, counter, bad, seq, 300, true);
    unsigned long long hb = 0; cudaSetDevice(1); cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost); cudaSetDevice(0);
    printf("visibility stress v%d: %llu stale elements in 300 pushes of 12 MB\n", v, hb);
  }
  return 0;
}
