// What DMMA rate can a 4x4-tile warp loop reach with distinct operand registers?
// (ceiling probe for the k-loop of cell_matvec_persistent_kernel)
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while(0)
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MT, int NT, bool SMEMB>
__global__ void __launch_bounds__(416, 1) k(double *out, const double *in, int iters) {
  __shared__ double xs[128 * 36];
  for (int i = threadIdx.x; i < 128 * 36; i += blockDim.x) xs[i] = in[i % 1024];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc[MT][NT][2];
#pragma unroll
  for (int t = 0; t < MT; ++t)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[t][n][0] = acc[t][n][1] = 0;
  double a[MT], b[NT];
#pragma unroll
  for (int t = 0; t < MT; ++t) a[t] = in[threadIdx.x + t * 7];
#pragma unroll
  for (int n = 0; n < NT; ++n) b[n] = in[threadIdx.x + 100 + n * 3];
  const double *xb = xs + (lane & 3) * 36 + (lane >> 2);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ks = 0; ks < 86; ++ks) {
      if (SMEMB) {
#pragma unroll
        for (int n = 0; n < NT; ++n) b[n] = xb[(ks % 30) * 4 * 36 + n * 8];
      }
#pragma unroll
      for (int t = 0; t < MT; ++t)
#pragma unroll
        for (int n = 0; n < NT; ++n) dmma(acc[t][n][0], acc[t][n][1], a[t], b[n]);
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < MT; ++t)
#pragma unroll
    for (int n = 0; n < NT; ++n) s += acc[t][n][0] + acc[t][n][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MT, int NT, bool SMEMB>
int run(const char *name, int warps, double *out, double *in) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 40;
  k<MT, NT, SMEMB><<<148, warps * 32>>>(out, in, 2);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k<MT, NT, SMEMB><<<148, warps * 32>>>(out, in, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * 256 * MT * NT * 86.0 * iters * warps * 148;
  double cyc_per_dmma = ms * 1e-3 * 1.965e9 / (MT * NT * 86.0 * iters * (warps / 4.0));
  printf("{\"probe\": \"%s\", \"warps\": %d, \"mt\": %d, \"nt\": %d, \"smem_b\": %d, \"tflops\": %.2f, \"cycles_per_dmma_per_smsp\": %.2f}\n", name, warps, MT, NT, (int)SMEMB, flops / ms * 1e-9, cyc_per_dmma);
  return 0;
}
int main() {
  double *out, *in;
  CK(cudaMalloc(&out, 148 * 1024 * 8)); CK(cudaMalloc(&in, 4096 * 8)); CK(cudaMemset(in, 0, 4096 * 8));
  run<4, 4, false>("regs", 12, out, in);
  run<4, 4, false>("regs", 8, out, in);
  run<4, 4, false>("regs", 4, out, in);
  run<4, 4, true>("smemB", 12, out, in);
  run<4, 4, true>("smemB", 8, out, in);
  run<4, 4, true>("smemB", 4, out, in);
  run<2, 4, true>("smemB", 12, out, in);
  run<4, 2, true>("smemB", 12, out, in);
  run<2, 8, true>("smemB", 12, out, in);
  run<8, 2, true>("smemB", 8, out, in);
  run<4, 8, true>("smemB", 4, out, in);
  return 0;
}
