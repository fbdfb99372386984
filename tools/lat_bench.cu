// Latency microbenchmarks that size the pivot-tile kernel (potrf_tile): dependent DFMA chain, the rsqrt chain of a
// column step, warp shuffle, shared-memory publish + named barrier, dependent DMMA.  One warp / one CTA, clock64.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/lat_bench tools/lat_bench.cu && tools/lat_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsqrt_newton(double s) {
  double r = (double)rsqrtf((float)s);
  const double hs = 0.5 * s;
  r = r * (1.5 - hs * r * r);
  return r * (1.5 - hs * r * r);
}
__global__ void k(double *out, long long *cyc, double seed) {
  __shared__ double buf[2][128];
  const int N = 256;
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  // 1. dependent DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, y, 1e-12);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / N;
  // 2. rsqrt_newton chain (each feeds the next)
  double s = 2.0 + x * 1e-20;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; i++) s = rsqrt_newton(s) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 64;
  // 2b. library rsqrt(double)
  double s2 = 2.0 + x * 1e-20;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; i++) s2 = rsqrt(s2) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 64;
  // 2c. sqrt(double) and division
  double s3 = 2.0 + x * 1e-20;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; i++) s3 = 1.0 / sqrt(s3) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 64;
  // 3. shuffle chain
  double z = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) z = __shfl_sync(0xffffffffu, z, (i * 7) & 31) + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0) / N;
  // 4. publish to shared memory, named barrier over 128 threads, read another thread's value
  double w = x;
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < N; i++) {
    buf[i & 1][threadIdx.x] = w;
    asm volatile("bar.sync 1, 128;\n" ::: "memory");
    w = buf[i & 1][(threadIdx.x + i) & 127] + 1.0;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0) / N;
  // 5. dependent DMMA m8n8k4
  double d0 = x, d1 = y;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(y), "d"(y));
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = (t1 - t0) / N;
  // 6. 32 independent DFMA per thread, 4 warps (throughput of the row update of a column step)
  double a[32];
#pragma unroll
  for (int j = 0; j < 32; j++) a[j] = x + j;
  t0 = clock64();
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int j = 0; j < 32; j++) a[j] = fma(a[j], y, z);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = (t1 - t0) / 64;
  double acc = x + s + s2 + s3 + z + w + d0 + d1;
#pragma unroll
  for (int j = 0; j < 32; j++) acc += a[j];
  out[threadIdx.x] = acc;
}
int main() {
  double *out;
  long long *cyc, h[8];
  cudaMalloc(&out, 128 * 8);
  cudaMalloc(&cyc, 64);
  for (int rep = 0; rep < 2; rep++) k<<<1, 128>>>(out, cyc, 1.0);
  cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  const char *n[8] = {"dependent DFMA", "rsqrt_newton (f32 seed + 2 Newton) + add", "rsqrt(double) + add", "1/sqrt(double) + add", "shfl + add",
                      "smem publish + bar.sync(128) + load + add", "dependent DMMA m8n8k4", "32 independent DFMA x 4 warps"};
  for (int i = 0; i < 8; i++) printf("%-48s %lld cycles\n", n[i], h[i]);
  return 0;
}
