// In-run measurement of the FP64 tensor (DMMA.8x8x4) issue rate: the denominator of the cell-kernel and
// projection rooflines.  MEASURED_PEAKS.json carries bf16 / HBM figures only, so bench.py measures the FP64 peak
// on the device it is running on (same probe as tools/microbench_fp64.cu: independent accumulator chains fed from
// registers, no memory traffic).
#include "common.cuh"

namespace dftfe_b200 {
namespace {

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_rate_kernel(double *out, int iters) {
  double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace
}  // namespace dftfe_b200

using namespace dftfe_b200;

extern "C" int dftfe_b200_measure_fp64_tensor_peak(dftfe_b200_ctx *ctx, double *tflops_out) {
  if (!ctx || !tflops_out) {
    set_error("measure_fp64_tensor_peak: null argument");
    return DFTFE_B200_ERR_INVALID;
  }
  DB_CUDA(cudaSetDevice(ctx->desc.device));
  const int warps = 32, iters = 20000;
  DevBuf<double> out;
  DB_TRY(out.alloc((size_t)ctx->num_sms * warps * 32));
  cudaEvent_t e0, e1;
  DB_CUDA(cudaEventCreate(&e0));
  DB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition is the warm-up
    DB_CUDA(cudaEventRecord(e0, ctx->stream));
    dmma_rate_kernel<8><<<ctx->num_sms, warps * 32, 0, ctx->stream>>>(out.p, rep == 0 ? 200 : iters);
    DB_CUDA(cudaEventRecord(e1, ctx->stream));
    DB_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    DB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::max(best, 2.0 * 256 * 8 * (double)iters * warps * ctx->num_sms / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return 0;
}
