// Mixed-precision variants of the subspace projections and rotations.  Real build: as described below.  Complex build:
// the same routines run on the REAL VIEW of the interleaved storage (2N real columns, block width 2 Bw) and solver.cu
// recombines the real 2N x 2N result into the complex N x N matrix - a complex FP32 block of the reference (cgemm on
// cuFloatComplex copies) becomes the four real FP32 blocks of its real / imaginary parts.
//
// Reference (all under src/linAlg/linearAlgebraOperationsDevice.cc unless stated):
//   fillParallelOverlapMatMixedPrecScalapack      :3543-3798  S = X^T X, diagonal Bw x Bw blocks FP64, the blocks
//                                                             below them FP32 from an FP32 copy of X, FP32 all-reduce
//   XtHXMixedPrecOverlapComputeCommun (kohnShamDFTOperatorDevice.cc:4550-5080)
//                                                             Hp = X^T H~ X, column blocks that end inside the
//                                                             first Noc states entirely FP32, the rest FP64
//   subspaceRotationCGSMixedPrecScalapack         :2243-2658  X <- X U (U = L^-T): diagonal blocks FP64,
//                                                             off-diagonal FP32
//   subspaceRotationRRMixedPrecScalapack          :2660-3076  X <- X diag(Q) (FP64) + X_sp (Q - diag Q)_sp (FP32)
// The reference issues cuBLAS Sgemm for every FP32 block.  Here they run on the tcgen05 tensor cores (tf32_gemm.cu:
// kind::tf32 MMAs with TMEM accumulators, 3xTF32 operand split for FP32-class accuracy, TMA-fed stages); the FP64
// diagonal blocks go through the hand-written DMMA kernel of projection.cu.  Option "cublas_projections" = 1 keeps the
// cuBLAS Sgemm / Dgemm calls for A/B comparisons.  The FP32 partial sums are all-reduced as floats, like
// DeviceCCLWrapper::deviceDirectAllReduceMixedPrecGroupWrapper (utils/DeviceDirectCCLWrapper.cc:196-261).
#include <algorithm>

#include "common.cuh"

namespace dftfe_b200 {

namespace {

__global__ void to_float_rows_kernel(const double *__restrict__ in, int64_t ldi, float *__restrict__ out, int64_t ldo,
                                     int ncols, int64_t rows) {
  // one warp per row, float2/double2 when the shapes allow
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * (int64_t)blockDim.x) >> 5;
  const bool vec = (ncols % 2 == 0) && (ldi % 2 == 0) && (ldo % 2 == 0) &&
                   ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
  for (int64_t r = w0; r < rows; r += nw) {
    const double *pi = in + r * ldi;
    float *po = out + r * ldo;
    if (vec) {
      for (int c = lane * 2; c < ncols; c += 64) {
        const double2 v = *reinterpret_cast<const double2 *>(pi + c);
        *reinterpret_cast<float2 *>(po + c) = make_float2((float)v.x, (float)v.y);
      }
    } else {
      for (int c = lane; c < ncols; c += 32) po[c] = (float)pi[c];
    }
  }
}

// Qsp (row-major N x N float) = off-diagonal part of Q; mode 1: the Bw x Bw diagonal blocks are dropped,
// mode 2: only the diagonal entries, mode 3: the 2 x 2 diagonal blocks (= the complex diagonal entries of the real
// embedding of a complex Q; diag[4 c + 2 r + s] = Q(2c + r, 2c + s)).  qColMajor: Q(k,j) at k + j*N, else k*N + j.
// Qthi / Qtlo (optional): the same off-diagonal part TRANSPOSED ([j][k], pitch ldt) and split in TF32-exact hi / lo
// parts - the K-major B operand of the tcgen05 rotation GEMM.
__global__ void split_rotation_kernel(const double *__restrict__ Q, int N, int qColMajor, int mode, int Bw,
                                      float *__restrict__ Qsp, double *__restrict__ diag, float *__restrict__ Qthi,
                                      float *__restrict__ Qtlo, int64_t ldt) {
  const int64_t total = (int64_t)N * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = idx / N, j = idx % N;
    const double v = Q[qColMajor ? ((int64_t)k + (int64_t)j * N) : idx];
    const bool onDiag = mode == 1 ? (k / Bw == j / Bw) : (mode == 3 ? (k / 2 == j / 2) : (k == j));
    const float f = onDiag ? 0.0f : (float)v;
    if (Qsp) Qsp[idx] = f;
    if (Qthi) {
      const float hi = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
      Qthi[(int64_t)j * ldt + k] = hi;
      Qtlo[(int64_t)j * ldt + k] = __uint_as_float(__float_as_uint(f - hi) & 0xffffe000u);
    }
    if (mode == 2 && k == j) diag[k] = v;
    if (mode == 3 && onDiag) diag[4 * (k / 2) + 2 * (k & 1) + (j & 1)] = v;
  }
}

// X[r, :] = X[r, :] * diag[:] + T[r, :]
__global__ void combine_diag_kernel(double *__restrict__ X, int N, int64_t rows, const double *__restrict__ diag,
                                    const float *__restrict__ T) {
  const int64_t total = rows * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    X[idx] = X[idx] * diag[idx % N] + (double)T[idx];
}

// complex diagonal through the real embedding: (X[r, 2c], X[r, 2c+1]) <- (X[r, 2c], X[r, 2c+1]) D_c + T, D_c 2 x 2
// (computeDiagQTimesXKernel, complex overload, linearAlgebraOperationsDevice.cc:220-237)
__global__ void combine_diag2_kernel(double *__restrict__ X, int N, int64_t rows, const double *__restrict__ diag,
                                     const float *__restrict__ T) {
  const int64_t total = rows * (N / 2);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = idx % (N / 2);
    const int64_t at = (idx / (N / 2)) * N + 2 * c;
    const double x0 = X[at], x1 = X[at + 1];
    const double *d = diag + 4 * c;
    X[at] = x0 * d[0] + x1 * d[2] + (double)T[at];
    X[at + 1] = x0 * d[1] + x1 * d[3] + (double)T[at + 1];
  }
}

// X[r, :] = D[r, :] + T[r, :]
__global__ void combine_block_kernel(double *__restrict__ X, const double *__restrict__ D,
                                     const float *__restrict__ T, int64_t total) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    X[idx] = D[idx] + (double)T[idx];
}

// full symmetric S (column- == row-major, double) from FP64 diagonal blocks (dp: Bw x N, block of column j at
// dp[r + j*Bw]) and the FP32 strictly-lower blocks (sp: column-major N x N)
__global__ void merge_overlap_kernel(const double *__restrict__ dp, const float *__restrict__ sp, int N, int Bw,
                                     double *__restrict__ S) {
  const int64_t total = (int64_t)N * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int a = idx % N, b = idx / N;
    const int i = max(a, b), j = min(a, b);  // lower-triangular source element
    const int bj = j / Bw;
    S[idx] = (i / Bw == bj) ? dp[(i - bj * Bw) + (int64_t)j * Bw] : (double)sp[(int64_t)i + (int64_t)j * N];
  }
}

// full symmetric Hp from the FP64 lower blocks (G, column-major N x N) and the FP32 column blocks (sp) that end
// inside the first Noc states.  mirror = 0 (real view of a complex build, not symmetric): element-wise copy of the
// block-lower part, the rest is never read.
__global__ void merge_projham_kernel(const double *__restrict__ G, const float *__restrict__ sp, int N, int Bw,
                                     int Noc, double *__restrict__ Hp, int mirror) {
  const int64_t total = (int64_t)N * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int a = idx % N, b = idx / N;
    const int i = mirror ? max(a, b) : a, j = mirror ? min(a, b) : b;
    const int j0 = (j / Bw) * Bw;
    const bool single = (j0 + min(Bw, N - j0)) <= Noc;
    const int64_t src = (int64_t)i + (int64_t)j * N;
    Hp[idx] = single ? (double)sp[src] : G[src];
  }
}

// FP64 column block [j, j+Bc) of the lower triangle (G, column-major N x N) -> FP32 copy of its rows below the
// diagonal block (sp) and the diagonal block itself (dp: Bw x N as in merge_overlap_kernel); skipDiag = 0 sends the
// whole block to sp (copyFromOverlapMatBlockToDPSPBlocks, linearAlgebraOperationsDevice.cc:4435-4447)
__global__ void split_dp_sp_kernel(const double *__restrict__ G, int N, int j, int Bc, int Bw, int skipDiag,
                                   double *__restrict__ dp, float *__restrict__ sp) {
  const int D = N - j;
  const int64_t total = (int64_t)D * Bc;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = idx % D, c = idx / D;  // row j + r, column j + c
    const double v = G[(int64_t)(j + r) + (int64_t)(j + c) * N];
    if (skipDiag && r < Bc)
      dp[r + (int64_t)(j + c) * Bw] = v;
    else
      sp[(int64_t)(j + r) + (int64_t)(j + c) * N] = (float)v;
  }
}

inline int grid_for(const dftfe_b200_ctx *ctx, int64_t work) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)ctx->num_sms * 8));
}

int to_float(dftfe_b200_ctx *ctx, const double *in, int64_t ldi, float *out, int64_t ldo, int ncols, int64_t rows) {
  if (rows == 0) return 0;
  ctx->launches += 1;
  to_float_rows_kernel<<<grid_for(ctx, rows * 32), 256, 0, ctx->stream>>>(in, ldi, out, ldo, ncols, rows);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// S = X^T X with FP64 diagonal blocks and FP32 off-diagonal blocks; full symmetric result, all-reduced.
// N, Bw: REAL columns / block width (complex build: twice the wavefunction counts, S = the real 2N x 2N Gram matrix)
int xtx_mixed_impl(dftfe_b200_ctx *ctx, const double *X, int N, int Bw, double *S, bool commOnly) {
  const int64_t M = ctx->M;
  if (commOnly) {
    // FP64 arithmetic, FP32 only on the wire (fillParallelOverlapMatMixedPrecCommunScalapackAsyncComputeCommun,
    // linearAlgebraOperationsDevice.cc:4233-4608): the off-diagonal blocks are rounded to FP32 for the all-reduce
    DB_TRY(ctx->mpSp.alloc((size_t)N * N));
    DB_TRY(ctx->mpDp.alloc((size_t)N * Bw));
    DB_TRY(ctx->denseW.alloc((size_t)N * N));
    double *G = ctx->denseW.p;
    DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
    DB_CUDA(cudaMemsetAsync(G, 0, (size_t)N * N * sizeof(double), ctx->stream));
    DB_CUDA(cudaMemsetAsync(ctx->mpSp.p, 0, (size_t)N * N * sizeof(float), ctx->stream));
    DB_CUDA(cudaMemsetAsync(ctx->mpDp.p, 0, (size_t)N * Bw * sizeof(double), ctx->stream));
    if (M > 0) {
      if (!ctx->use_cublas_dense && al16(X) && dmma_projection_usable(ctx, N, N, N, 0, 0, N, N)) {
        DB_TRY(launch_xty(ctx, X, N, 0, X, N, 0, N, N, 0, 0, true, G, N));
      } else {
        const double one = 1.0, zero = 0.0;
        for (int j = 0; j < N; j += Bw) {
          const int Bc = std::min(Bw, N - j);
          ProfScope ps(ctx, "projection");
          DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, N - j, Bc, (int)M, &one, X + j, N, X + j, N,
                                &zero, G + j + (size_t)j * N, N));
        }
      }
    }
    for (int j = 0; j < N; j += Bw) {
      const int Bc = std::min(Bw, N - j);
      ctx->launches += 1;
      split_dp_sp_kernel<<<grid_for(ctx, (int64_t)(N - j) * Bc), 256, 0, ctx->stream>>>(G, N, j, Bc, Bw, 1, ctx->mpDp.p,
                                                                                         ctx->mpSp.p);
    }
    DB_CUDA(cudaGetLastError());
    DB_TRY(allreduce_sum(ctx, ctx->mpDp.p, (size_t)N * Bw));
    DB_TRY(allreduce_sum_f32(ctx, ctx->mpSp.p, (size_t)N * N));
    ctx->launches += 1;
    merge_overlap_kernel<<<grid_for(ctx, (int64_t)N * N), 256, 0, ctx->stream>>>(ctx->mpDp.p, ctx->mpSp.p, N, Bw, S);
    DB_CUDA(cudaGetLastError());
    return 0;
  }
  const bool tc = !ctx->use_cublas_dense;           // tcgen05 3xTF32 kernel for the FP32 blocks
  const int64_t Mp = (std::max<int64_t>(M, 1) + 3) / 4 * 4;  // pitch of the K-major (transposed) FP32 copies
  if (tc) {
    DB_TRY(ctx->mpThi.alloc((size_t)N * Mp));
    DB_TRY(ctx->mpTlo.alloc((size_t)N * Mp));
  } else {
    DB_TRY(ctx->mpXsp.alloc((size_t)std::max<int64_t>(M, 1) * N));
  }
  DB_TRY(ctx->mpSp.alloc((size_t)N * N));
  DB_TRY(ctx->mpDp.alloc((size_t)N * Bw));
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  DB_CUDA(cudaMemsetAsync(ctx->mpSp.p, 0, (size_t)N * N * sizeof(float), ctx->stream));
  DB_CUDA(cudaMemsetAsync(ctx->mpDp.p, 0, (size_t)N * Bw * sizeof(double), ctx->stream));
  if (tc)
    DB_TRY(launch_split_transpose(ctx, X, N, 0, N, M, ctx->mpThi.p, ctx->mpTlo.p, Mp));
  else
    DB_TRY(to_float(ctx, X, N, ctx->mpXsp.p, N, N, M));
  const double one = 1.0, zero = 0.0;
  const float onef = 1.0f, zerof = 0.0f;
  for (int j = 0; j < N && M > 0; j += Bw) {
    const int Bc = std::min(Bw, N - j), DRem = N - j - Bc;
    double *Cd = ctx->mpDp.p + (size_t)j * Bw;
    if (!ctx->use_cublas_dense && al16(X) && dmma_projection_usable(ctx, N, N, N, j, j, Bc, Bc)) {
      DB_TRY(launch_xty(ctx, X, N, j, X, N, j, Bc, Bc, j, j, true, Cd, Bw));
    } else {
      ProfScope ps(ctx, "projection");
      DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, Bc, Bc, (int)M, &one, X + j, N, X + j, N, &zero, Cd,
                            Bw));
    }
    if (DRem > 0 && tc) {
      // S(j+Bc.., j..j+Bc) = X[:, j+Bc:]^T X[:, j:j+Bc], k = local DoFs (split-k)
      DB_TRY(launch_tf32x3_gemm(ctx, ctx->mpThi.p, ctx->mpTlo.p, N, Mp, j + Bc, DRem, ctx->mpThi.p, ctx->mpTlo.p, N, Mp,
                                j, Bc, M, nullptr, 0, ctx->mpSp.p + (j + Bc) + (size_t)j * N, N));
    } else if (DRem > 0) {
      ProfScope ps(ctx, "projection_fp32");
      DB_CUBLAS(cublasSgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, DRem, Bc, (int)M, &onef, ctx->mpXsp.p + j + Bc, N,
                            ctx->mpXsp.p + j, N, &zerof, ctx->mpSp.p + (j + Bc) + (size_t)j * N, N));
    }
  }
  DB_TRY(allreduce_sum(ctx, ctx->mpDp.p, (size_t)N * Bw));
  DB_TRY(allreduce_sum_f32(ctx, ctx->mpSp.p, (size_t)N * N));
  ctx->launches += 1;
  merge_overlap_kernel<<<grid_for(ctx, (int64_t)N * N), 256, 0, ctx->stream>>>(ctx->mpDp.p, ctx->mpSp.p, N, Bw, S);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// Hp = X^H (H~ X): column blocks [j, j+Bc) with j + Bc <= Noc entirely in FP32, the others FP64.  N, Noc, j count
// wavefunctions; all matrix indices below are REAL columns (x cm).  Real build: Hout = full symmetric Hp.  Complex
// build: Hout = the block-lower part of the real 2N x 2N matrix G(a,b) = sum_m Xr[m,a] (H~X)r[m,b] (not symmetric;
// the caller combines it into the complex lower triangle).
int xthx_mixed_impl(dftfe_b200_ctx *ctx, const double *X, int N, int Noc, double *Hout, bool commOnly) {
  const int cm = ctx->cm, Nr = N * cm, Nocr = Noc * cm;
  const int Bw0 = std::min(ctx->B, N), Bw = Bw0 * cm;
  const int64_t M = ctx->M;
  const bool tc = !ctx->use_cublas_dense && !commOnly;
  const int64_t Mp = (std::max<int64_t>(M, 1) + 3) / 4 * 4;
  if (tc) {
    DB_TRY(ctx->mpThi.alloc((size_t)Nr * Mp));
    DB_TRY(ctx->mpTlo.alloc((size_t)Nr * Mp));
    DB_TRY(ctx->mpBhi.alloc((size_t)Bw * Mp));
    DB_TRY(ctx->mpBlo.alloc((size_t)Bw * Mp));
  } else {
    DB_TRY(ctx->mpXsp.alloc(commOnly ? 1 : (size_t)std::max<int64_t>(M, 1) * Nr));
    DB_TRY(ctx->mpBlockSp.alloc((size_t)std::max<int64_t>(M, 1) * Bw));
  }
  DB_TRY(ctx->mpSp.alloc((size_t)Nr * Nr));
  DB_TRY(ctx->denseW.alloc((size_t)Nr * Nr));
  double *G = ctx->denseW.p;
  DB_CUDA(cudaMemsetAsync(ctx->mpSp.p, 0, (size_t)Nr * Nr * sizeof(float), ctx->stream));
  DB_CUDA(cudaMemsetAsync(G, 0, (size_t)Nr * Nr * sizeof(double), ctx->stream));
  if (tc && Noc >= Bw0)  // only the columns an FP32 block touches: rows j.. of X^T for blocks ending inside Noc
    DB_TRY(launch_split_transpose(ctx, X, Nr, 0, Nr, M, ctx->mpThi.p, ctx->mpTlo.p, Mp));
  else if (!commOnly && !tc)
    DB_TRY(to_float(ctx, X, Nr, ctx->mpXsp.p, Nr, Nr, M));
  const double one = 1.0, zero = 0.0;
  const float onef = 1.0f, zerof = 0.0f;
  for (int j0 = 0; j0 < N; j0 += Bw0) {
    const int Bc0 = std::min(Bw0, N - j0);
    const int j = j0 * cm, Bc = Bc0 * cm, D = Nr - j;
    DB_TRY(apply_H_to_columns(ctx, X, N, j0, Bc0));  // blockY = H~ X[:, j0:j0+Bc0]  (M x Bc real columns)
    if (M == 0) continue;
    DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
    if (j0 + Bc0 <= Noc && tc) {
      // Hp(j.., j..j+Bc) = X[:, j:]^T (H~ X)[:, j:j+Bc] in FP32 on the tcgen05 tensor cores
      DB_TRY(launch_split_transpose(ctx, ctx->blockY.p, Bc, 0, Bc, M, ctx->mpBhi.p, ctx->mpBlo.p, Mp));
      DB_TRY(launch_tf32x3_gemm(ctx, ctx->mpThi.p, ctx->mpTlo.p, Nr, Mp, j, D, ctx->mpBhi.p, ctx->mpBlo.p, Bc, Mp, 0,
                                Bc, M, nullptr, 0, ctx->mpSp.p + j + (size_t)j * Nr, Nr));
    } else if (j0 + Bc0 <= Noc && !commOnly) {
      DB_TRY(to_float(ctx, ctx->blockY.p, Bc, ctx->mpBlockSp.p, Bc, Bc, M));
      ProfScope ps(ctx, "projection_fp32");
      DB_CUBLAS(cublasSgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, D, Bc, (int)M, &onef, ctx->mpXsp.p + j, Nr,
                            ctx->mpBlockSp.p, Bc, &zerof, ctx->mpSp.p + j + (size_t)j * Nr, Nr));
    } else if (!ctx->use_cublas_dense && al16(X) && dmma_projection_usable(ctx, Nr, Nr, Bc, j, 0, D, Bc)) {
      DB_TRY(launch_xty(ctx, X, Nr, j, ctx->blockY.p, Bc, 0, D, Bc, j, j, true, G + j + (size_t)j * Nr, Nr));
    } else {
      ProfScope ps(ctx, "projection");
      DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, D, Bc, (int)M, &one, X + j, Nr, ctx->blockY.p, Bc,
                            &zero, G + j + (size_t)j * Nr, Nr));
    }
  }
  if (commOnly) {
    // FP64 blocks that end inside the core states travel as FP32 (XtHXMixedPrecCommunOverlapComputeCommun,
    // kohnShamDFTOperatorDevice.cc:5082-5536: copyValueType1ArrToValueType2Arr before the all-reduce)
    for (int j = 0; j < Nr; j += Bw) {
      const int Bc = std::min(Bw, Nr - j);
      if (j + Bc > Nocr) break;
      ctx->launches += 2;
      split_dp_sp_kernel<<<grid_for(ctx, (int64_t)(Nr - j) * Bc), 256, 0, ctx->stream>>>(G, Nr, j, Bc, Bw, 0, nullptr,
                                                                                          ctx->mpSp.p);
      DB_CUDA(cudaMemset2DAsync(G + j + (size_t)j * Nr, (size_t)Nr * sizeof(double), 0,
                                (size_t)(Nr - j) * sizeof(double), (size_t)Bc, ctx->stream));
    }
    DB_CUDA(cudaGetLastError());
  }
  DB_TRY(allreduce_sum(ctx, G, (size_t)Nr * Nr));
  DB_TRY(allreduce_sum_f32(ctx, ctx->mpSp.p, (size_t)Nr * Nr));
  ctx->launches += 1;
  merge_projham_kernel<<<grid_for(ctx, (int64_t)Nr * Nr), 256, 0, ctx->stream>>>(G, ctx->mpSp.p, Nr, Bw, Nocr, Hout,
                                                                                 cm == 1 ? 1 : 0);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// X <- X Q in mixed precision.  mode 1 (CGS): FP64 diagonal Bw x Bw blocks + FP32 off-diagonal;
// mode 2 (RR): FP64 diag(Q) + FP32 (Q - diag Q); mode 3: mode 2 for the real embedding of a complex Q (2 x 2
// diagonal blocks).  N, Bw: REAL columns / block width.  Row chunks, in place.
int rotate_mixed_impl(dftfe_b200_ctx *ctx, double *X, int N, int Bw, const double *Q, bool qColMajor, int mode) {
  DB_CHECK(mode >= 1 && mode <= 3, "rotate: unknown mixed-precision mode %d", mode);
  DB_CHECK(mode != 3 || N % 2 == 0, "rotate: mode 3 needs an even (embedded complex) column count");
  const int64_t M = ctx->M;
  if (M == 0) return 0;
  const int64_t chunk = std::min<int64_t>(M, 148 * 128);
  const bool tc = !ctx->use_cublas_dense;
  const int64_t Np = ((int64_t)N + 3) / 4 * 4;       // pitch of the K-major FP32 operands (k = wavefunction index)
  DB_TRY(ctx->mpSp.alloc((size_t)N * N));            // Q off-diagonal part, FP32 row-major
  DB_TRY(ctx->mpDp.alloc((size_t)N * std::max(Bw, 2)));  // diag(Q) (mode 2), 2 x 2 diagonal blocks (mode 3)
  DB_TRY(ctx->mpXsp.alloc((size_t)chunk * N * 2));   // [chunk x N] FP32 copy of X, then the FP32 product
  float *Xsp = ctx->mpXsp.p, *Tsp = ctx->mpXsp.p + (size_t)chunk * N;
  if (tc) {
    DB_TRY(ctx->mpQhi.alloc((size_t)N * Np));
    DB_TRY(ctx->mpQlo.alloc((size_t)N * Np));
    DB_TRY(ctx->mpThi.alloc((size_t)chunk * Np));
    DB_TRY(ctx->mpTlo.alloc((size_t)chunk * Np));
    if (Np != N) {  // pad columns of Q^T must read as zero
      DB_CUDA(cudaMemsetAsync(ctx->mpQhi.p, 0, (size_t)N * Np * sizeof(float), ctx->stream));
      DB_CUDA(cudaMemsetAsync(ctx->mpQlo.p, 0, (size_t)N * Np * sizeof(float), ctx->stream));
    }
  }
  if (mode == 1) DB_TRY(ctx->rotScratch.alloc((size_t)chunk * N));
  ctx->launches += 1;
  split_rotation_kernel<<<grid_for(ctx, (int64_t)N * N), 256, 0, ctx->stream>>>(
      Q, N, qColMajor ? 1 : 0, mode, Bw, ctx->mpSp.p, ctx->mpDp.p, tc ? ctx->mpQhi.p : nullptr,
      tc ? ctx->mpQlo.p : nullptr, Np);
  DB_CUDA(cudaGetLastError());
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  const float onef = 1.0f, zerof = 0.0f;
  const double one = 1.0, zero = 0.0;
  for (int64_t r0 = 0; r0 < M; r0 += chunk) {
    const int mc = (int)std::min<int64_t>(chunk, M - r0);
    double *Xc = X + (size_t)r0 * N;
    if (tc) {
      // T (row-major mc x N) = X (Q - diagonal part): A = X rows (k = wavefunction index), B = Q^T, no split-k
      DB_TRY(launch_split_rows(ctx, Xc, N, N, mc, ctx->mpThi.p, ctx->mpTlo.p, Np));
      DB_TRY(launch_tf32x3_gemm(ctx, ctx->mpThi.p, ctx->mpTlo.p, mc, Np, 0, mc, ctx->mpQhi.p, ctx->mpQlo.p, N, Np, 0, N,
                                N, Tsp, N, nullptr, 0));
    } else {
      DB_TRY(to_float(ctx, Xc, N, Xsp, N, N, mc));
      ProfScope ps(ctx, "rotation_fp32");
      // T (row-major mc x N) = Xsp * Qsp  <=>  T_cm (N x mc) = Qsp_rm-as-cm (N x N) * Xsp_cm (N x mc)
      DB_CUBLAS(cublasSgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_N, N, mc, N, &onef, ctx->mpSp.p, N, Xsp, N, &zerof,
                            Tsp, N));
    }
    if (mode == 2) {
      ctx->launches += 1;
      combine_diag_kernel<<<grid_for(ctx, (int64_t)mc * N), 256, 0, ctx->stream>>>(Xc, N, mc, ctx->mpDp.p, Tsp);
    } else if (mode == 3) {
      ctx->launches += 1;
      combine_diag2_kernel<<<grid_for(ctx, (int64_t)mc * N / 2), 256, 0, ctx->stream>>>(Xc, N, mc, ctx->mpDp.p, Tsp);
    } else {
      ProfScope ps(ctx, "rotation", N / Bw + 1);
      for (int j = 0; j < N; j += Bw) {
        const int Bc = std::min(Bw, N - j);
        // D[:, j:j+Bc] (row-major) = Xc[:, j:j+Bc] * Q[j:j+Bc, j:j+Bc]
        if (qColMajor)
          DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_T, CUBLAS_OP_N, Bc, mc, Bc, &one, Q + j + (size_t)j * N, N,
                                Xc + j, N, &zero, ctx->rotScratch.p + j, N));
        else
          DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_N, Bc, mc, Bc, &one, Q + j + (size_t)j * N, N,
                                Xc + j, N, &zero, ctx->rotScratch.p + j, N));
      }
      combine_block_kernel<<<grid_for(ctx, (int64_t)mc * N), 256, 0, ctx->stream>>>(Xc, ctx->rotScratch.p, Tsp,
                                                                                   (int64_t)mc * N);
    }
    DB_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace dftfe_b200
