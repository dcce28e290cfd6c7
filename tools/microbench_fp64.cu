// FP64 peak probes for B200 (sm_100a): the denominators of the cell-GEMM
// roofline.  MEASURED_PEAKS.json carries only bf16/HBM figures, so this tool
// measures (1) cuBLAS DGEMM 8192^3 burst + sustained, (2) cuBLAS
// DgemmStridedBatched in the reference's cell-GEMM shape (m=B=256, n=k=343,
// batch=4913; matrixVectorProductImplementationsDevice.cc:51-79), (3) raw
// DMMA.8x8x4 issue rate from registers, (4) raw DFMA issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_fp64 microbench_fp64.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__,      \
             __LINE__);                                                        \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_rate(double *out, int iters) {
  double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[i][0]), "+d"(c[i][1])
        : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_rate(double *out, int iters) {
  double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, nsm,
         prop.clockRate);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  double *out;
  CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));

  // ---- raw DMMA / DFMA issue rates
  for (int warps : {4, 8, 16, 32}) {
    int iters = 20000;
    dmma_rate<8><<<nsm, warps * 32>>>(out, 100);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dmma_rate<8><<<nsm, warps * 32>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1);
    double flops = 2.0 * 256 * 8 * (double)iters * warps * nsm;
    printf("{\"probe\": \"dmma_8x8x4\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n",
           warps, ms, flops / ms * 1e-9);
  }
  for (int warps : {4, 8, 16, 32}) {
    int iters = 20000;
    dfma_rate<16><<<nsm, warps * 32>>>(out, 100);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dfma_rate<16><<<nsm, warps * 32>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1);
    double flops = 2.0 * 32 * 16 * (double)iters * warps * nsm;
    printf("{\"probe\": \"dfma\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n",
           warps, ms, flops / ms * 1e-9);
  }

  // ---- cuBLAS DGEMM
  cublasHandle_t h;
  cublasCreate(&h);
  {
    const int n = 8192;
    double *A, *B, *C;
    CK(cudaMalloc(&A, sizeof(double) * n * n));
    CK(cudaMalloc(&B, sizeof(double) * n * n));
    CK(cudaMalloc(&C, sizeof(double) * n * n));
    CK(cudaMemset(A, 0, sizeof(double) * n * n));
    CK(cudaMemset(B, 0, sizeof(double) * n * n));
    std::vector<double> hA((size_t)n * n);
    for (size_t i = 0; i < hA.size(); ++i) hA[i] = (double)rand() / RAND_MAX - 0.5;
    CK(cudaMemcpy(A, hA.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B, hA.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    double al = 1, be = 0;
    for (int i = 0; i < 3; ++i)
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 10; ++i) {
      CK(cudaEventRecord(e0));
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = time_ms(e0, e1);
      if (ms < best) best = ms;
    }
    printf("{\"probe\": \"cublas_dgemm_8192_burst\", \"ms\": %.3f, \"tflops\": %.2f}\n", best,
           2.0 * n * n * n / best * 1e-9);
    int reps = 0;
    CK(cudaEventRecord(e0));
    for (; reps < 100; ++reps)
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1);
    printf("{\"probe\": \"cublas_dgemm_8192_sustained\", \"reps\": %d, \"ms_total\": %.1f, \"tflops\": %.2f}\n",
           reps, ms, 2.0 * n * n * n * reps / ms * 1e-9);
    cudaFree(A);
    cudaFree(B);
    cudaFree(C);
  }
  // ---- the reference's cell GEMM shape through cuBLAS (K2)
  {
    const int Bw = 256, nn = 343, batch = 4913;
    double *X, *H, *Y;
    CK(cudaMalloc(&X, sizeof(double) * (size_t)Bw * nn * batch));
    CK(cudaMalloc(&Y, sizeof(double) * (size_t)Bw * nn * batch));
    CK(cudaMalloc(&H, sizeof(double) * (size_t)nn * nn * batch));
    CK(cudaMemset(X, 0, sizeof(double) * (size_t)Bw * nn * batch));
    CK(cudaMemset(H, 0, sizeof(double) * (size_t)nn * nn * batch));
    double al = 1, be = 0;
    for (int i = 0; i < 3; ++i)
      cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, Bw, nn, nn, &al, X, Bw,
                                (long long)nn * Bw, H, nn, (long long)nn * nn, &be, Y, Bw,
                                (long long)nn * Bw, batch);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 5; ++i) {
      CK(cudaEventRecord(e0));
      cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, Bw, nn, nn, &al, X, Bw,
                                (long long)nn * Bw, H, nn, (long long)nn * nn, &be, Y, Bw,
                                (long long)nn * Bw, batch);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = time_ms(e0, e1);
      if (ms < best) best = ms;
    }
    printf("{\"probe\": \"cublas_cell_gemm_256x343x343_b4913\", \"ms\": %.3f, \"tflops\": %.2f}\n",
           best, 2.0 * Bw * nn * nn * batch / best * 1e-9);
    // a long-k projection-like GEMM: (2048 x 256) = X^T(2048 x M) * Y(M x 256), M=1061208
    cudaFree(X);
    cudaFree(Y);
    cudaFree(H);
  }
  {
    const int N = 2048, Bv = 256;
    const size_t M = 1061208;
    double *X, *Y, *S;
    CK(cudaMalloc(&X, sizeof(double) * M * N));
    CK(cudaMalloc(&Y, sizeof(double) * M * Bv));
    CK(cudaMalloc(&S, sizeof(double) * N * Bv));
    CK(cudaMemset(X, 0, sizeof(double) * M * N));
    CK(cudaMemset(Y, 0, sizeof(double) * M * Bv));
    double al = 1, be = 0;
    for (int i = 0; i < 2; ++i)
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, N, Bv, (int)M, &al, X, N, Y, Bv, &be, S, N);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, N, Bv, (int)M, &al, X, N, Y, Bv, &be, S, N);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1);
    printf("{\"probe\": \"cublas_proj_gemm_2048x256xM\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms,
           2.0 * N * Bv * (double)M / ms * 1e-9);
  }
  return 0;
}
