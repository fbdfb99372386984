// FP64 peak microbenchmark for the roofline denominator (MEASURED_PEAKS.json has no FP64 figure):
// issue-rate of each DMMA shape, DFMA, and cuBLAS DGEMM 8192^3 as a yardstick.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -lcublas -o tools/fp64_peak
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#define ITERS 4096
template <int SHAPE>
__global__ void dmma_loop(double *out) {
  double c[8][4];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 4; j++) c[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; i++) b[i] = 1.0 - threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
      else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      else if (SHAPE == 3)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      else {
#pragma unroll
        for (int j = 0; j < 4; j++) c[i][j] = fma(a[j], b[j], c[i][j]);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
double run(const char *name, double flop_per_warp_instr, int warps_per_block, double *d) {
  int sms = 148, blocks = sms * 4;
  dim3 g(blocks), b(32 * warps_per_block);
  dmma_loop<SHAPE><<<g, b>>>(d);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    dmma_loop<SHAPE><<<g, b>>>(d);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double flops = (double)blocks * warps_per_block * ITERS * 8 * flop_per_warp_instr;
  double tf = flops / (best * 1e-3) * 1e-12;
  printf("%-12s warps/block %d : %8.3f ms  %7.2f TFLOP/s\n", name, warps_per_block, best, tf);
  return tf;
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  double *d;
  cudaMalloc(&d, 148 * 4 * 1024 * sizeof(double));
  for (int w : {4, 8, 16}) {
    run<0>("dmma m8n8k4", 2.0 * 8 * 8 * 4, w, d);
    run<1>("dmma m16n8k4", 2.0 * 16 * 8 * 4, w, d);
    run<2>("dmma m16n8k8", 2.0 * 16 * 8 * 8, w, d);
    run<3>("dmma m16n8k16", 2.0 * 16 * 8 * 16, w, d);
    run<4>("dfma", 2.0 * 32 * 4, w, d);
  }
  // cuBLAS DGEMM yardstick
  cublasHandle_t h;
  if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) { printf("cublasCreate failed\n"); return 1; }
  for (int n : {4096, 8192}) {
    double *A, *B, *C;
    size_t bytes = (size_t)n * n * sizeof(double);
    cudaMalloc(&A, bytes), cudaMalloc(&B, bytes), cudaMalloc(&C, bytes);
    cudaMemset(A, 0, bytes), cudaMemset(B, 0, bytes), cudaMemset(C, 0, bytes);
    double one = 1.0, zero = 0.0;
    cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
      cudaEventRecord(e0);
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("cublasDgemm NT n=%d : %8.3f ms  %7.2f TFLOP/s\n", n, best, 2.0 * n * n * n / (best * 1e-3) * 1e-12);
    // sustained: back to back for ~2 s
    cudaEventRecord(e0);
    int reps = 0;
    for (; reps < 40; reps++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cublasDgemm NT n=%d sustained x%d : %7.2f TFLOP/s\n", n, reps, 2.0 * n * n * n * reps / (ms * 1e-3) * 1e-12);
    cudaFree(A), cudaFree(B), cudaFree(C);
  }
  return 0;
}
