// Operator application, Chebyshev filter, projections, Rayleigh-Ritz and the
// solve() driver.  Host C++ orchestration over the kernels of this library.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace dftfe_b200 {

// ---------------------------------------------------------------------------
// operator
// ---------------------------------------------------------------------------

// dst = live*(a*src + b*dst) + s * M^-1/2 H M^-1/2 src   (Loewdin basis, fused)
// One pass: ghost update -> distribute -> coloured fused cell kernel -> slave->master
// -> ghost accumulate.  Constrained and ghost rows of dst end at 0.
static int fused_apply_impl(dftfe_b200_ctx *ctx, double *src, double *dst, int ncols, double a, double b,
                            double s, const double *rowB, bool fp32Comm = false) {
  DB_CHECK(ctx->have_mass, "set_mass must be called before applying the operator");
  ncols *= ctx->cm;  // real columns from here on (complex vectors are interleaved re/im)
  const int ldx = ncols;
  fp32Comm = fp32Comm && (ncols % 2 == 0);
  DB_TRY(ghost_update(ctx, src, ncols, ldx, fp32Comm));
  DB_TRY(launch_distribute(ctx, src, ncols, ldx, ctx->invSqrtM.p));
  EpilogueParams ep;
  ep.a = a;
  ep.b = b;
  ep.s = s;
  // rowIn / rowOut (M^-1/2 on free rows) are folded into the tiled cell matrices
  ep.rowIn = nullptr;
  ep.rowOut = nullptr;
  ep.rowA = nullptr;  // liveness travels in the row words of the index map
  ep.rowB = rowB;
  DB_TRY(nonlocal_project(ctx, src, ncols, ldx, ctx->rowIn.p));  // C^T (M^-1/2 x), all-reduced
  DB_TRY(launch_cell_matvec(ctx, src, dst, ncols, ldx, ep));
  DB_TRY(launch_orphan_first_touch(ctx, src, dst, ncols, ldx, ep));
  DB_TRY(nonlocal_apply(ctx, dst, ncols, ldx, ctx->rowOut.p, s));  // += s M^-1/2 C V (C^T ...)
  DB_TRY(launch_slave_to_master(ctx, dst, ncols, ldx, ctx->rowOut.p));
  DB_TRY(ghost_accumulate(ctx, dst, ncols, ldx, ctx->rowOut.p, fp32Comm));
  DB_TRY(ghost_zero(ctx, dst, ncols, ldx));
  DB_TRY(ghost_zero(ctx, src, ncols, ldx));
  return 0;
}

int op_fused_apply(dftfe_b200_ctx *ctx, double *src, double *dst, int ncols, double a, double b, double s,
                   bool fp32Comm) {
  return fused_apply_impl(ctx, src, dst, ncols, a, b, s, nullptr, fp32Comm);
}

// operatorDFTDeviceClass::HX net effect (kohnShamDFTOperatorDevice.cc:3765-3860)
int op_hx(dftfe_b200_ctx *ctx, double *src, double *dst, int ncols, int scaleFlag, double scalar, int doUnscale,
          bool fp32Comm) {
  // dst <- (scaleFlag ? dst : M^-1/2 dst) + scalar * H~ src on owned free rows, 0 on constrained rows
  DB_TRY(fused_apply_impl(ctx, src, dst, ncols, 0.0, 1.0, scalar,
                          scaleFlag ? nullptr : ctx->invSqrtM.p, fp32Comm));
  // src side effects of the reference: constrained rows end at 0 (x M^1/2 = 0), ghosts zeroed;
  // without unscaling the caller sees scalar * M^-1/2 * src.
  const int nr = ncols * ctx->cm;
  if (doUnscale) {
    DB_TRY(launch_set_zero_rows(ctx, src, nr, nr));
  } else {
    // (constrained rows keep the distributed, scaled values exactly as the reference leaves them)
    DB_TRY(launch_row_scale(ctx, src, ctx->M, nr, nr, scalar, ctx->rowIn.p));
  }
  return 0;
}

// operatorDFTDeviceClass::HXCheby, FP64 (kohnShamDFTOperatorDevice.cc:3874-3997): dst += H src
int op_hx_cheby(dftfe_b200_ctx *ctx, double *src, double *dst, int ncols, bool mixedPrec) {
  ncols *= ctx->cm;
  const int ldx = ncols;
  mixedPrec = mixedPrec && (ncols % 2 == 0);
  DB_TRY(ghost_update(ctx, src, ncols, ldx, mixedPrec));
  DB_TRY(launch_distribute(ctx, src, ncols, ldx, nullptr));
  EpilogueParams ep;  // a=0, b=1, s=1: pure accumulate; undo the scales folded into the tiled H
  ep.rowIn = ctx->rowInInv.p;
  ep.rowOut = ctx->rowOutInv.p;
  ep.allLive = 1;
  DB_TRY(nonlocal_project(ctx, src, ncols, ldx, nullptr));
  DB_TRY(launch_cell_matvec(ctx, src, dst, ncols, ldx, ep));
  DB_TRY(nonlocal_apply(ctx, dst, ncols, ldx, nullptr, 1.0));
  DB_TRY(launch_slave_to_master(ctx, dst, ncols, ldx, nullptr));
  DB_TRY(ghost_zero(ctx, src, ncols, ldx));
  DB_TRY(ghost_accumulate(ctx, dst, ncols, ldx, nullptr, mixedPrec));
  DB_TRY(ghost_zero(ctx, dst, ncols, ldx));
  return 0;
}

// ---------------------------------------------------------------------------
// Chebyshev filter (linearAlgebraOperationsDevice.cc:531-727), Loewdin basis,
// recurrence fused into the cell kernel epilogue: one HBM pass per degree.
// ---------------------------------------------------------------------------
// The recurrence state of one block: which of (x, y) holds the newest iterate, and the running sigma.
struct ChebState {
  double *X, *Y;  // Y = newest after step 1
  double e, c, sigma, sigma1, gamma;
  int degree = 0;
};

static void cheb_begin(ChebState &st, double *x_d, double *y_d, double a, double b, double a0) {
  st.e = (b - a) / 2.0;
  st.c = (b + a) / 2.0;
  st.sigma = st.e / (a0 - st.c);
  st.sigma1 = st.sigma;
  st.gamma = 2.0 / st.sigma1;
  st.X = x_d;
  st.Y = y_d;
  st.degree = 0;
}

// one more degree (1-based).  mixedPrec: FP32 ghost payloads for degrees 2..m-1, exactly where the reference
// passes mixedPrecOverall && useMixedPrecCheby to HXCheby (linearAlgebraOperationsDevice.cc:612-622, 700-708);
// degrees 1 and m go through the FP64 HX (:560-566, 650-656).
static int cheb_step(dftfe_b200_ctx *ctx, ChebState &st, int ncols, int m, bool mixedPrec) {
  const int degree = ++st.degree;
  if (degree == 1) {
    // Y = (sigma1/e) (H~ X - c X)
    return op_fused_apply(ctx, st.X, st.Y, ncols, -st.c * st.sigma1 / st.e, 0.0, st.sigma1 / st.e, false);
  }
  const double sigma2 = 1.0 / (st.gamma - st.sigma);
  const double alpha1 = 2.0 * sigma2 / st.e, alpha2 = -(st.sigma * sigma2);
  // X <- alpha1 (H~ - c) Y + alpha2 X
  DB_TRY(op_fused_apply(ctx, st.Y, st.X, ncols, -st.c * alpha1, alpha2, alpha1, mixedPrec && degree < m));
  std::swap(st.X, st.Y);
  st.sigma = sigma2;
  return 0;
}

static int cheb_end(dftfe_b200_ctx *ctx, ChebState &st, double *x_d, int ncols) {
  if (st.Y != x_d) {
    ctx->launches += 1;
    DB_CUDA(cudaMemcpyAsync(x_d, st.Y, (size_t)(ctx->M + ctx->G) * ncols * ctx->cm * sizeof(double),
                            cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return 0;
}

static int cheb_filter_impl(dftfe_b200_ctx *ctx, double *x_d, double *y_d, int ncols, int m, double a, double b,
                            double a0, bool mixedPrec = false) {
  DB_CHECK(m >= 1, "Chebyshev degree must be >= 1");
  ChebState st;
  cheb_begin(st, x_d, y_d, a, b, a0);
  for (int degree = 1; degree <= m; ++degree) DB_TRY(cheb_step(ctx, st, ncols, m, mixedPrec));
  return cheb_end(ctx, st, x_d, ncols);
}

// ---------------------------------------------------------------------------
// projections / rotation
// ---------------------------------------------------------------------------
namespace {

__global__ void symmetrise_from_lower_kernel(double *S, int N) {
  // column-major lower -> full
  const int64_t total = (int64_t)N * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = idx % N, j = idx / N;  // element (i,j) col-major
    if (i < j) S[idx] = S[(int64_t)j + (int64_t)i * N];
  }
}

// partial[blockIdx.x][col] = sum over this block's rows of (hx - lambda*x)^2
__global__ void residual_partial_kernel(const double *__restrict__ X, int N, int j0, const double *__restrict__ HXb,
                                        int ncols, int64_t rows, const double *__restrict__ eig, int colsPerEig,
                                        double *__restrict__ partial) {
  __shared__ double red[8][33];
  const int col = blockIdx.y * 32 + threadIdx.x;
  double acc = 0.0;
  if (col < ncols) {
    const double lam = eig[(j0 + col) / colsPerEig];
    for (int64_t r = blockIdx.x * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.x * 8) {
      const double d = HXb[(size_t)r * ncols + col] - lam * X[(size_t)r * N + j0 + col];
      acc += d * d;
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < ncols) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    partial[(size_t)blockIdx.x * ncols + col] = s;
  }
}

__global__ void residual_final_kernel(const double *__restrict__ partial, int nParts, int ncols,
                                      double *__restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col < ncols) {
    double s = 0.0;
    for (int p = 0; p < nParts; ++p) s += partial[(size_t)p * ncols + col];
    out[col] = s;
  }
}

__global__ void column_dot_partial_kernel(const double *__restrict__ x, const double *__restrict__ y, int64_t n,
                                          double *__restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += x[i] * y[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// complex N x N (column-major, interleaved) from the real 2N x 2N embedding G(a,b) = sum_m Xr[m,a] Yr[m,b]:
//   S(i,j) = sum_m conj(x_i) y_j = (G(2i,2j) + G(2i+1,2j+1)) + i (G(2i,2j+1) - G(2i+1,2j))
// lowerOnly: only i >= j is formed from G's lower tiles, the rest by Hermitian symmetry.
__global__ void cplx_combine_kernel(const double *__restrict__ G, int N, double *__restrict__ S, int lowerOnly) {
  const int64_t total = (int64_t)N * N;
  const int64_t ld = 2 * (int64_t)N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int i = idx % N, j = idx / N;
    const bool flip = lowerOnly && i < j;
    if (flip) {
      const int t = i;
      i = j;
      j = t;
    }
    const double re = G[2 * i + (2 * j) * ld] + G[2 * i + 1 + (2 * j + 1) * ld];
    const double im = G[2 * i + (2 * j + 1) * ld] - G[2 * i + 1 + (2 * j) * ld];
    S[2 * idx] = re;
    S[2 * idx + 1] = flip ? -im : im;
  }
}

// real row-major 2N x 2N embedding of a complex N x N matrix R for X <- X R on interleaved storage:
//   Rt[2k][2j] = Re R(k,j), Rt[2k][2j+1] = Im R(k,j), Rt[2k+1][2j] = -Im R(k,j), Rt[2k+1][2j+1] = Re R(k,j)
__global__ void cplx_embed_kernel(const double *__restrict__ R, int N, int colMajor, double *__restrict__ Rt) {
  const int64_t total = (int64_t)N * N;
  const int64_t ld = 2 * (int64_t)N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = idx / N, j = idx % N;
    const int64_t src = colMajor ? ((int64_t)k + (int64_t)j * N) : idx;
    const double re = R[2 * src], im = R[2 * src + 1];
    Rt[(2 * k) * ld + 2 * j] = re;
    Rt[(2 * k) * ld + 2 * j + 1] = im;
    Rt[(2 * k + 1) * ld + 2 * j] = -im;
    Rt[(2 * k + 1) * ld + 2 * j + 1] = re;
  }
}

// out(row-major complex) <- in(column-major complex)
__global__ void cplx_transpose_kernel(const double *__restrict__ in, double *__restrict__ out, int N) {
  const int64_t total = (int64_t)N * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = idx / N, j = idx % N;  // out[i][j] = in(i,j)
    const int64_t src = (int64_t)i + (int64_t)j * N;
    out[2 * idx] = in[2 * src];
    out[2 * idx + 1] = in[2 * src + 1];
  }
}

__global__ void pair_sum_kernel(const double *__restrict__ in, double *__restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[2 * i] + in[2 * i + 1];
}

__global__ void axpy_kernel(double *__restrict__ y, const double *__restrict__ x, double alpha, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += alpha * x[i];
}

__global__ void scale_copy_kernel(double *__restrict__ y, const double *__restrict__ x, double alpha, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = alpha * x[i];
}

// densityMatrixEigenBasisFirstOrderResponse, the N x N step (src/linAlg/rayleighRitzDevice.cc:1559-1668): the recursive
// Fermi-operator expansion applied to -c H' is an element-wise factor, because every ScaLAPACK operation of the
// reference (row / column scalings by the x0 recurrences and sums of those) acts on element (i,j) alone:
//   per level:  D(i,j) <- Y0_i [ (x0_i + x0_j)(1 - 2 x0'_j) + 2 x0'_j ] D(i,j),  Y0 = 1/(2 x0 (x0 - 1) + 1),  x0' = Y0 x0^2
// on the lower triangle XtHX filled (i >= j), followed by D <- D + D^H with the diagonal halved (:1634-1668).
// Hp, D: column-major N x N (complex: interleaved); x0: the N start values 0.5 - c (eps_i - mu).
__global__ void fermi_response_kernel(const double *__restrict__ Hp, int N, int cm, const double *__restrict__ x0,
                                      double c, int levels, double *__restrict__ D) {
  const int64_t total = (int64_t)N * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int a = idx % N, b = idx / N;           // element (a, b) of the output
    const int i = max(a, b), j = min(a, b);        // its lower-triangle source (i, j)
    double xi = x0[i], xj = x0[j], f = -c;
    for (int l = 0; l < levels; ++l) {
      const double yi = 1.0 / (2.0 * xi * (xi - 1.0) + 1.0), yj = 1.0 / (2.0 * xj * (xj - 1.0) + 1.0);
      const double xin = yi * xi * xi, xjn = yj * xj * xj;
      f *= yi * ((xi + xj) * (1.0 - 2.0 * xjn) + 2.0 * xjn);
      xi = xin;
      xj = xjn;
    }
    const int64_t src = (int64_t)i + (int64_t)j * N;
    if (cm == 1) {
      D[idx] = f * Hp[src];
    } else {
      const double re = f * Hp[2 * src], im = f * Hp[2 * src + 1];
      D[2 * idx] = re;
      D[2 * idx + 1] = (a == b) ? 0.0 : (a > b ? im : -im);  // upper = conjugate of lower, real diagonal
    }
  }
}

}  // namespace

static int ensure_block_scratch(dftfe_b200_ctx *ctx) {
  DB_TRY(ctx->blockX.alloc((size_t)(ctx->M + ctx->G) * ctx->B * ctx->cm));
  DB_TRY(ctx->blockY.alloc((size_t)(ctx->M + ctx->G) * ctx->B * ctx->cm));
  return 0;
}

// Gram matrix of the local rows: real: S = X^T X (N x N, symmetric, column- == row-major);
// complex: S(i,j) = sum_m conj(X[m,i]) X[m,j], column-major interleaved, via the real 2N x 2N embedding.
// Lower-triangular tiles/blocks only, mirrored afterwards; all-reduced over the ranks.
static int xtx_impl(dftfe_b200_ctx *ctx, const double *X, int N, double *S) {
  const int cm = ctx->cm, Nr = N * cm;
  const int Bw = std::min(ctx->B, N) * cm;
  const double one = 1.0, zero = 0.0;
  double *G = S;
  if (ctx->cplx) {
    DB_TRY(ctx->denseG.alloc((size_t)Nr * Nr));
    G = ctx->denseG.p;
  }
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  DB_CUDA(cudaMemsetAsync(G, 0, (size_t)Nr * Nr * sizeof(double), ctx->stream));
  if (ctx->M > 0 && !ctx->use_cublas_dense && al16(X) && dmma_projection_usable(ctx, Nr, Nr, Nr, 0, 0, Nr, Nr)) {
    // hand-written DMMA kernel, lower-triangular 128x128 tiles of the whole matrix in one pass
    DB_TRY(launch_xty(ctx, X, Nr, 0, X, Nr, 0, Nr, Nr, 0, 0, true, G, Nr));
  } else if (ctx->M > 0) {
    for (int j = 0; j < Nr; j += Bw) {
      const int Bc = std::min(Bw, Nr - j), D = Nr - j;
      ProfScope ps(ctx, "projection");
      // block(D x Bc) = X_cm[j:, :] (D x M) * X_cm[j:j+Bc, :]^T
      DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, D, Bc, (int)ctx->M, &one, X + j, Nr, X + j, Nr,
                            &zero, G + j + (size_t)j * Nr, Nr));
    }
  }
  ctx->launches += 1;
  symmetrise_from_lower_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(G, Nr);
  if (ctx->cplx) {
    ctx->launches += 1;
    cplx_combine_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(G, N, S, 0);
  }
  DB_CUDA(cudaGetLastError());
  DB_TRY(allreduce_sum(ctx, S, (size_t)N * N * cm));
  return 0;
}

// HXb(M x ncols, dense) = H~ * X[:, j0:j0+ncols]
int apply_H_to_columns(dftfe_b200_ctx *ctx, const double *X, int N, int j0, int ncols) {
  const int cm = ctx->cm;
  DB_TRY(ensure_block_scratch(ctx));
  DB_TRY(launch_block_copy_from_full(ctx, X, N * cm, j0 * cm, ctx->blockX.p, ncols * cm, ctx->M, nullptr));
  DB_TRY(ghost_zero(ctx, ctx->blockX.p, ncols * cm, ncols * cm));
  // dst = H~ src (b = 0: dst is never read)
  DB_TRY(op_fused_apply(ctx, ctx->blockX.p, ctx->blockY.p, ncols, 0.0, 0.0, 1.0));
  return 0;
}

// Hp = X^H (H~ X), same storage conventions as xtx_impl
static int xthx_impl(dftfe_b200_ctx *ctx, const double *X, int N, double *Hp) {
  const int cm = ctx->cm, Nr = N * cm;
  const int Bc0 = std::min(ctx->B, N);
  const double one = 1.0, zero = 0.0;
  double *G = Hp;
  if (ctx->cplx) {
    DB_TRY(ctx->denseG.alloc((size_t)Nr * Nr));
    G = ctx->denseG.p;
  }
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  DB_CUDA(cudaMemsetAsync(G, 0, (size_t)Nr * Nr * sizeof(double), ctx->stream));
  for (int j = 0; j < N; j += Bc0) {
    const int Bc = std::min(Bc0, N - j);
    const int jr = j * cm, Bcr = Bc * cm, D = Nr - jr;
    DB_TRY(apply_H_to_columns(ctx, X, N, j, Bc));
    if (ctx->M > 0 && !ctx->use_cublas_dense && al16(X) && dmma_projection_usable(ctx, Nr, Nr, Bcr, jr, 0, D, Bcr)) {
      DB_TRY(launch_xty(ctx, X, Nr, jr, ctx->blockY.p, Bcr, 0, D, Bcr, jr, jr, true, G + jr + (size_t)jr * Nr, Nr));
    } else if (ctx->M > 0) {
      ProfScope ps(ctx, "projection");
      // block(D x Bc) = X_cm[j:, :] (D x M) * HXb_cm (Bc x M)^T
      DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, D, Bcr, (int)ctx->M, &one, X + jr, Nr,
                            ctx->blockY.p, Bcr, &zero, G + jr + (size_t)jr * Nr, Nr));
    }
  }
  ctx->launches += 1;
  if (ctx->cplx)
    cplx_combine_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(G, N, Hp, 1);
  else
    symmetrise_from_lower_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(G, Nr);
  DB_CUDA(cudaGetLastError());
  DB_TRY(allreduce_sum(ctx, Hp, (size_t)N * N * cm));
  return 0;
}

// Out(M x Nout, ld ldo) = X(M x Nr real) * Q[:, c0:c0+Nout].  Out == nullptr: in place (c0 = 0, Nout = Nr) through a
// row-chunk scratch.  qColMajor: Q memory holds Q(i,j) at i + j*Nr (cuSOLVER output), else row-major.
static int rotate_real(dftfe_b200_ctx *ctx, double *X, int N, const double *Q, bool qColMajor, int c0 = 0,
                       int Nout = -1, double *Out = nullptr, int ldo = 0) {
  if (ctx->M == 0) return 0;
  if (Nout < 0) Nout = N;
  const bool inPlace = Out == nullptr;
  DB_CHECK(!inPlace || (c0 == 0 && Nout == N), "rotate: an in-place rotation needs the full square Q");
  if (inPlace) ldo = N;
  const int64_t chunk = std::min<int64_t>(ctx->M, 148 * 128);
  if (inPlace) DB_TRY(ctx->rotScratch.alloc((size_t)chunk * N));
  const double one = 1.0, zero = 0.0;
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  const bool dmma = !ctx->use_cublas_dense && al16(X) && al16(Q) && (Out == nullptr || al16(Out)) &&
                    dmma_rotation_usable(N, Nout, N, ldo) && (c0 % 2 == 0);
  const double *Qrm = Q;
  if (dmma && qColMajor) {  // the kernel wants Q(k, j) with j fastest
    DB_TRY(ctx->denseC.alloc((size_t)N * N));
    DB_TRY(launch_transpose_square(ctx, Q, ctx->denseC.p, N));
    Qrm = ctx->denseC.p;
  }
  for (int64_t r0 = 0; r0 < ctx->M; r0 += chunk) {
    const int mc = (int)std::min<int64_t>(chunk, ctx->M - r0);
    double *dst = inPlace ? ctx->rotScratch.p : Out + (size_t)r0 * ldo;
    if (dmma) {
      DB_TRY(launch_xq(ctx, X + (size_t)r0 * N, N, mc, Qrm + c0, N, Nout, dst, ldo));
    } else {
      ProfScope ps(ctx, "rotation");
      // dst_cm (Nout x mc) = Qsub^T * X_cm ; row-major Q memory is col-major Q^T
      if (qColMajor)
        DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_T, CUBLAS_OP_N, Nout, mc, N, &one, Q + (size_t)c0 * N, N,
                              X + (size_t)r0 * N, N, &zero, dst, ldo));
      else
        DB_CUBLAS(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_N, Nout, mc, N, &one, Q + c0, N, X + (size_t)r0 * N,
                              N, &zero, dst, ldo));
    }
    if (inPlace) {
      ctx->launches += 1;
      DB_CUDA(cudaMemcpyAsync(X + (size_t)r0 * N, ctx->rotScratch.p, (size_t)mc * N * sizeof(double),
                              cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  return 0;
}

// X <- X * Q for N wavefunction columns (complex: through the real 2N x 2N embedding of Q).
// mixedMode 0: FP64; 1 / 2: the reference's CGS / RR mixed-precision rotations (complex: block width 2 Bw for the
// CGS variant, the complex diagonal = 2 x 2 diagonal blocks of the embedding for the RR variant).
static int rotate_impl(dftfe_b200_ctx *ctx, double *X, int N, const double *Q, bool qColMajor, int mixedMode = 0) {
  const int Bw = std::min(ctx->B, N);
  if (!ctx->cplx) {
    if (mixedMode != 0) return rotate_mixed_impl(ctx, X, N, Bw, Q, qColMajor, mixedMode);
    return rotate_real(ctx, X, N, Q, qColMajor);
  }
  const int Nr = 2 * N;
  DB_TRY(ctx->denseG.alloc((size_t)Nr * Nr));
  ctx->launches += 1;
  cplx_embed_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(Q, N, qColMajor ? 1 : 0, ctx->denseG.p);
  DB_CUDA(cudaGetLastError());
  if (mixedMode == 1) return rotate_mixed_impl(ctx, X, Nr, 2 * Bw, ctx->denseG.p, false, 1);
  if (mixedMode == 2) return rotate_mixed_impl(ctx, X, Nr, 2, ctx->denseG.p, false, 3);
  return rotate_real(ctx, X, Nr, ctx->denseG.p, false);
}

// XFrac(M x Nfr) = X * Q[:, c0:c0+Nfr]  (subspaceRotationSpectrumSplitScalapack, linearAlgebraOperationsDevice.cc:1446-1830)
static int rotate_into(dftfe_b200_ctx *ctx, double *X, int N, const double *Qcm, int c0, int Nfr, double *XFrac) {
  if (!ctx->cplx) return rotate_real(ctx, X, N, Qcm, true, c0, Nfr, XFrac, Nfr);
  // complex: embed the N x Nfr column slice of Q as real 2N x 2Nfr (row-major, ld 2N of the full embedding)
  const int Nr = 2 * N;
  DB_TRY(ctx->denseG.alloc((size_t)Nr * Nr));
  ctx->launches += 1;
  cplx_embed_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(Qcm, N, 1, ctx->denseG.p);
  DB_CUDA(cudaGetLastError());
  return rotate_real(ctx, X, Nr, ctx->denseG.p, false, 2 * c0, 2 * Nfr, XFrac, 2 * Nfr);
}

static int residual_impl(dftfe_b200_ctx *ctx, const double *X, int N, const double *eig_h, double *res_h) {
  const int cm = ctx->cm;
  DB_TRY(ctx->eigDev.alloc(N));
  DB_TRY(ctx->resDev.alloc((size_t)N * cm + N));
  DB_CUDA(cudaMemcpyAsync(ctx->eigDev.p, eig_h, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const int Bc0 = std::min(ctx->B, N);
  const int nParts = ctx->num_sms * 2;
  DB_TRY(ctx->partials.alloc((size_t)nParts * Bc0 * cm));
  double *perCol = ctx->resDev.p;              // N*cm per-real-column sums
  double *perState = ctx->resDev.p + (size_t)N * cm;
  for (int j = 0; j < N; j += Bc0) {
    const int Bc = std::min(Bc0, N - j), Bcr = Bc * cm;
    DB_TRY(apply_H_to_columns(ctx, X, N, j, Bc));
    ProfScope ps(ctx, "residual", 2);
    dim3 grid(nParts, (Bcr + 31) / 32), block(32, 8);
    residual_partial_kernel<<<grid, block, 0, ctx->stream>>>(X, N * cm, j * cm, ctx->blockY.p, Bcr, ctx->M,
                                                             ctx->eigDev.p, cm, ctx->partials.p);
    residual_final_kernel<<<(Bcr + 127) / 128, 128, 0, ctx->stream>>>(ctx->partials.p, nParts, Bcr,
                                                                      perCol + (size_t)j * cm);
    DB_CUDA(cudaGetLastError());
  }
  if (ctx->cplx) {
    ctx->launches += 1;
    pair_sum_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(perCol, perState, N);
  } else {
    perState = perCol;
  }
  DB_TRY(allreduce_sum(ctx, perState, N));
  DB_CUDA(cudaMemcpyAsync(res_h, perState, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < N; ++i) res_h[i] = std::sqrt(res_h[i]);
  return 0;
}

// ---------------------------------------------------------------------------
// Lanczos (linearAlgebraOperationsDevice.cc:340-527) - vectors stay on device
// ---------------------------------------------------------------------------
static int global_dot(dftfe_b200_ctx *ctx, const double *x, const double *y, double *out) {
  const int nParts = 256;
  DB_TRY(ctx->partials.alloc(nParts + 8));
  ctx->launches += 2;
  column_dot_partial_kernel<<<nParts, 256, 0, ctx->stream>>>(x, y, ctx->M * ctx->cm, ctx->partials.p);
  residual_final_kernel<<<1, 32, 0, ctx->stream>>>(ctx->partials.p, nParts, 1, ctx->partials.p + nParts);
  DB_CUDA(cudaGetLastError());
  DB_TRY(allreduce_sum(ctx, ctx->partials.p + nParts, 1));
  DB_CUDA(cudaMemcpyAsync(out, ctx->partials.p + nParts, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// symmetric eigenvalues by cyclic Jacobi (n <= 40)
static void jacobi_eigenvalues(std::vector<double> &A, int n, std::vector<double> &ev) {
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-30) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = cs * akp - sn * akq;
          A[k * n + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = cs * apk - sn * aqk;
          A[q * n + k] = sn * apk + cs * aqk;
        }
      }
  }
  ev.resize(n);
  for (int i = 0; i < n; ++i) ev[i] = A[i * n + i];
  std::sort(ev.begin(), ev.end());
}

static int lanczos_impl(dftfe_b200_ctx *ctx, int reproducible, double out[2]) {
  const int iters = reproducible ? 40 : 20;
  // single-column vectors; complex build: (re, im) pairs, so every length below is in doubles and the
  // real dot products over 2M doubles are Re<x,y> (H~ is Hermitian: alpha is real up to rounding)
  const int cm = ctx->cm;
  const int64_t rows = (ctx->M + ctx->G) * cm;
  const int64_t owned = ctx->M * cm;
  DB_TRY(ensure_block_scratch(ctx));
  DB_TRY(ctx->HXfull.alloc((size_t)rows * 4));
  double *v = ctx->HXfull.p, *f = v + rows, *v0 = f + rows, *src = v0 + rows;
  std::vector<double> hv(rows, 0.0);
  std::srand(ctx->rank);
  for (int64_t i = 0; i < ctx->M; ++i) hv[i * cm] = ((double)std::rand()) / ((double)RAND_MAX);
  for (uint32_t r : ctx->conRows_h)
    if (r < ctx->M) hv[(int64_t)r * cm] = 0.0;
  DB_CUDA(cudaMemcpyAsync(v, hv.data(), rows * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  double nrm2 = 0;
  DB_TRY(global_dot(ctx, v, v, &nrm2));
  const int g = ctx->num_sms * 4;
  ctx->launches += 1;
  scale_copy_kernel<<<g, 256, 0, ctx->stream>>>(v, v, 1.0 / std::sqrt(nrm2), rows);

  auto applyH = [&](const double *in, double *outv) -> int {
    ctx->launches += 1;
    DB_CUDA(cudaMemcpyAsync(src, in, rows * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return op_fused_apply(ctx, src, outv, 1, 0.0, 0.0, 1.0);
  };
  std::vector<double> T((size_t)iters * iters, 0.0);
  double alpha = 0, beta = 0;
  DB_TRY(applyH(v, f));
  DB_TRY(global_dot(ctx, f, v, &alpha));
  ctx->launches += 1;
  axpy_kernel<<<g, 256, 0, ctx->stream>>>(f, v, -alpha, owned);
  T[0] = alpha;
  for (int j = 1; j < iters; ++j) {
    double ff = 0;
    DB_TRY(global_dot(ctx, f, f, &ff));
    beta = std::sqrt(ff);
    ctx->launches += 3;
    DB_CUDA(cudaMemcpyAsync(v0, v, rows * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    scale_copy_kernel<<<g, 256, 0, ctx->stream>>>(v, f, 1.0 / beta, owned);
    DB_TRY(applyH(v, f));
    axpy_kernel<<<g, 256, 0, ctx->stream>>>(f, v0, -beta, owned);
    DB_TRY(global_dot(ctx, f, v, &alpha));
    ctx->launches += 1;
    axpy_kernel<<<g, 256, 0, ctx->stream>>>(f, v, -alpha, owned);
    T[(size_t)j * iters + j - 1] = beta;
    T[(size_t)(j - 1) * iters + j] = beta;
    T[(size_t)j * iters + j] = alpha;
  }
  DB_CUDA(cudaGetLastError());
  std::vector<double> ev;
  jacobi_eigenvalues(T, iters, ev);
  double ff = 0;
  DB_TRY(global_dot(ctx, f, f, &ff));
  const double fnorm = std::sqrt(ff);
  out[0] = std::floor(ev[0]);
  out[1] = std::ceil(ev[iters - 1] + (reproducible ? fnorm : fnorm / 10.0));
  return 0;
}

// ---------------------------------------------------------------------------
// dense N x N step on device (reference: host ScaLAPACK/ELPA,
// src/linAlg/rayleighRitzDevice.cc:485-782)
// ---------------------------------------------------------------------------
static int dense_check_info(dftfe_b200_ctx *ctx, const char *what) {
  int info = 0;
  DB_CUDA(cudaMemcpyAsync(&info, ctx->devInfo.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (info != 0) {
    set_error("%s failed: info = %d", what, info);
    return DFTFE_B200_ERR_NUMERIC;
  }
  return 0;
}

static int dense_cholesky(dftfe_b200_ctx *ctx, double *S, int N) {
  DB_CUSOLVER(cusolverDnSetStream(ctx->cusolver, ctx->stream));
  DB_TRY(ctx->devInfo.alloc(1));
  int lwork = 0;
  DB_CUSOLVER(cusolverDnDpotrf_bufferSize(ctx->cusolver, CUBLAS_FILL_MODE_LOWER, N, S, N, &lwork));
  DB_TRY(ctx->cusolverWork.alloc(lwork));
  ctx->launches += 1;
  DB_CUSOLVER(cusolverDnDpotrf(ctx->cusolver, CUBLAS_FILL_MODE_LOWER, N, S, N, ctx->cusolverWork.p, lwork,
                               ctx->devInfo.p));
  return dense_check_info(ctx, "Cholesky factorisation of X^T X (cusolverDnDpotrf)");
}

static int dense_eigh(dftfe_b200_ctx *ctx, double *A, int N, double *W) {
  DB_CUSOLVER(cusolverDnSetStream(ctx->cusolver, ctx->stream));
  DB_TRY(ctx->devInfo.alloc(1));
  int lwork = 0;
  DB_CUSOLVER(cusolverDnDsyevd_bufferSize(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, A, N,
                                          W, &lwork));
  DB_TRY(ctx->cusolverWork.alloc(lwork));
  ctx->launches += 1;
  DB_CUSOLVER(cusolverDnDsyevd(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, A, N, W,
                               ctx->cusolverWork.p, lwork, ctx->devInfo.p));
  return dense_check_info(ctx, "eigendecomposition of the projected Hamiltonian (cusolverDnDsyevd)");
}

static int dense_cholesky_cplx(dftfe_b200_ctx *ctx, double *S, int N) {
  DB_CUSOLVER(cusolverDnSetStream(ctx->cusolver, ctx->stream));
  DB_TRY(ctx->devInfo.alloc(1));
  int lwork = 0;
  cuDoubleComplex *Sz = reinterpret_cast<cuDoubleComplex *>(S);
  DB_CUSOLVER(cusolverDnZpotrf_bufferSize(ctx->cusolver, CUBLAS_FILL_MODE_LOWER, N, Sz, N, &lwork));
  DB_TRY(ctx->cusolverWork.alloc((size_t)lwork * 2));
  ctx->launches += 1;
  DB_CUSOLVER(cusolverDnZpotrf(ctx->cusolver, CUBLAS_FILL_MODE_LOWER, N, Sz, N,
                               reinterpret_cast<cuDoubleComplex *>(ctx->cusolverWork.p), lwork, ctx->devInfo.p));
  return dense_check_info(ctx, "Cholesky factorisation of X^H X (cusolverDnZpotrf)");
}

static int dense_eigh_cplx(dftfe_b200_ctx *ctx, double *A, int N, double *W) {
  DB_CUSOLVER(cusolverDnSetStream(ctx->cusolver, ctx->stream));
  DB_TRY(ctx->devInfo.alloc(1));
  int lwork = 0;
  cuDoubleComplex *Az = reinterpret_cast<cuDoubleComplex *>(A);
  DB_CUSOLVER(cusolverDnZheevd_bufferSize(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, Az, N,
                                          W, &lwork));
  DB_TRY(ctx->cusolverWork.alloc((size_t)lwork * 2));
  ctx->launches += 1;
  DB_CUSOLVER(cusolverDnZheevd(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, Az, N, W,
                               reinterpret_cast<cuDoubleComplex *>(ctx->cusolverWork.p), lwork, ctx->devInfo.p));
  return dense_check_info(ctx, "eigendecomposition of the projected Hamiltonian (cusolverDnZheevd)");
}

// Which pieces run in mixed precision (dftParameters::useMixedPrec*, all gated by useMixedPrecOverall) and the
// number of "core" states whose XtHX blocks may be FP32 (numCoreWfcXtHX / Noc).
struct RRFlags {
  bool commOnly = false;    // useMixedPrecCommunOnlyXTHXCGSO: FP64 arithmetic, FP32 all-reduce payloads
  bool mpOverlap = false;   // useMixedPrecCGS_O
  bool mpCgsRot = false;    // useMixedPrecCGS_SR
  bool mpXtHX = false;      // useMixedPrecXTHXSpectrumSplit
  bool mpRRRot = false;     // useMixedPrecSubspaceRotRR
  int nCore = 0;
};

// complex results: column-major interleaved, like xtx_impl / xthx_impl
static int overlap_any(dftfe_b200_ctx *ctx, const double *X, int N, double *S, bool mixed, bool commOnly = false) {
  if (!mixed) return xtx_impl(ctx, X, N, S);
  const int Bw = std::min(ctx->B, N);
  if (!ctx->cplx) return xtx_mixed_impl(ctx, X, N, Bw, S, commOnly);
  // real 2N x 2N Gram matrix of the interleaved storage in mixed precision, then S = its complex combination
  DB_TRY(ctx->denseG.alloc((size_t)4 * N * N));
  DB_TRY(xtx_mixed_impl(ctx, X, 2 * N, 2 * Bw, ctx->denseG.p, commOnly));
  ctx->launches += 1;
  cplx_combine_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->denseG.p, N, S, 0);
  DB_CUDA(cudaGetLastError());
  return 0;
}
static int projham_any(dftfe_b200_ctx *ctx, const double *X, int N, double *Hp, bool mixed, int nCore,
                       bool commOnly = false) {
  if (!mixed || nCore <= 0) return xthx_impl(ctx, X, N, Hp);
  if (!ctx->cplx) return xthx_mixed_impl(ctx, X, N, nCore, Hp, commOnly);
  DB_TRY(ctx->denseG.alloc((size_t)4 * N * N));
  DB_TRY(xthx_mixed_impl(ctx, X, N, nCore, ctx->denseG.p, commOnly));
  ctx->launches += 1;
  cplx_combine_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->denseG.p, N, Hp, 1);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// complex build of rayleighRitzGEP / CGS+RR: S = X^H X = L L^H, Hs = L^-1 Hp L^-H = Q' D Q'^H, X <- X L^-H Q'
static int rr_cplx(dftfe_b200_ctx *ctx, double *X, int N, double *eig_h, bool cgsFirst, const RRFlags &f) {
  const size_t nn = (size_t)N * N * 2;
  DB_TRY(ctx->denseA.alloc(nn));
  DB_TRY(ctx->denseB.alloc(nn));
  DB_TRY(ctx->eigDev.alloc(N));
  double *S = ctx->denseA.p, *Hp = ctx->denseB.p;
  cuDoubleComplex *Sz = reinterpret_cast<cuDoubleComplex *>(S), *Hz = reinterpret_cast<cuDoubleComplex *>(Hp);
  const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0);
  DB_TRY(overlap_any(ctx, X, N, S, f.mpOverlap, f.commOnly));
  DB_TRY(dense_cholesky_cplx(ctx, S, N));
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  if (cgsFirst) {
    // X <- X L^-H, then plain RR
    std::vector<double> eye(nn, 0.0);
    for (int i = 0; i < N; ++i) eye[((size_t)i * N + i) * 2] = 1.0;
    DB_CUDA(cudaMemcpyAsync(Hp, eye.data(), nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 1;
    DB_CUBLAS(cublasZtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT, N,
                          N, &one, Sz, N, Hz, N));
    DB_TRY(rotate_impl(ctx, X, N, Hp, true, f.mpCgsRot ? 1 : 0));
    DB_TRY(projham_any(ctx, X, N, Hp, f.mpXtHX, f.nCore, f.commOnly));
    DB_TRY(dense_eigh_cplx(ctx, Hp, N, ctx->eigDev.p));
    DB_TRY(rotate_impl(ctx, X, N, Hp, true, f.mpRRRot ? 2 : 0));
  } else {
    DB_TRY(xthx_impl(ctx, X, N, Hp));
    DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
    ctx->launches += 3;
    DB_CUBLAS(cublasZtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, N,
                          N, &one, Sz, N, Hz, N));
    DB_CUBLAS(cublasZtrsm(ctx->cublas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT,
                          N, N, &one, Sz, N, Hz, N));
    DB_TRY(dense_eigh_cplx(ctx, Hp, N, ctx->eigDev.p));
    DB_CUBLAS(cublasZtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT, N,
                          N, &one, Sz, N, Hz, N));
    DB_TRY(rotate_impl(ctx, X, N, Hp, true, f.mpRRRot ? 2 : 0));
  }
  DB_CUDA(cudaMemcpyAsync(eig_h, ctx->eigDev.p, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static int identity_to(dftfe_b200_ctx *ctx, double *A, int N) {
  std::vector<double> eye((size_t)N * N, 0.0);
  for (int i = 0; i < N; ++i) eye[(size_t)i * N + i] = 1.0;
  ctx->launches += 1;
  DB_CUDA(cudaMemcpyAsync(A, eye.data(), eye.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// rayleighRitzGEP (src/linAlg/rayleighRitzDevice.cc:355-819)
static int rr_gep(dftfe_b200_ctx *ctx, double *X, int N, double *eig_h, const RRFlags &f = RRFlags()) {
  if (ctx->cplx) return rr_cplx(ctx, X, N, eig_h, false, f);
  const size_t nn = (size_t)N * N;
  DB_TRY(ctx->denseA.alloc(nn));
  DB_TRY(ctx->denseB.alloc(nn));
  DB_TRY(ctx->eigDev.alloc(N));
  double *S = ctx->denseA.p, *Hp = ctx->denseB.p;
  const double one = 1.0;
  DB_TRY(overlap_any(ctx, X, N, S, f.mpOverlap, f.commOnly));
  DB_TRY(dense_cholesky(ctx, S, N));  // S lower <- L
  DB_TRY(xthx_impl(ctx, X, N, Hp));
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  ctx->launches += 2;
  // Hs = L^-1 Hp L^-T
  DB_CUBLAS(cublasDtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, N,
                        N, &one, S, N, Hp, N));
  DB_CUBLAS(cublasDtrsm(ctx->cublas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT,
                        N, N, &one, S, N, Hp, N));
  DB_TRY(dense_eigh(ctx, Hp, N, ctx->eigDev.p));  // Hp <- Q' (columns)
  // R = L^-T Q'
  ctx->launches += 1;
  DB_CUBLAS(cublasDtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, N,
                        N, &one, S, N, Hp, N));
  DB_TRY(rotate_impl(ctx, X, N, Hp, true, f.mpRRRot ? 2 : 0));
  DB_CUDA(cudaMemcpyAsync(eig_h, ctx->eigDev.p, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// pseudoGramSchmidtOrthogonalization (src/linAlg/pseudoGSDevice.cc:81-463): X <- X L^-T with S = X^T X = L L^T
static int cgs_orthogonalise(dftfe_b200_ctx *ctx, double *X, int N, const RRFlags &f) {
  const size_t nn = (size_t)N * N;
  DB_TRY(ctx->denseA.alloc(nn));
  DB_TRY(ctx->denseB.alloc(nn));
  double *S = ctx->denseA.p, *U = ctx->denseB.p;
  const double one = 1.0;
  DB_TRY(overlap_any(ctx, X, N, S, f.mpOverlap, f.commOnly));
  DB_TRY(dense_cholesky(ctx, S, N));
  DB_TRY(identity_to(ctx, U, N));  // L^-T as a column-major matrix: solve L^T U = I
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  ctx->launches += 1;
  DB_CUBLAS(cublasDtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, N,
                        N, &one, S, N, U, N));
  return rotate_impl(ctx, X, N, U, true, f.mpCgsRot ? 1 : 0);
}

// pseudoGramSchmidtOrthogonalization + rayleighRitz
// (src/linAlg/pseudoGSDevice.cc:81-463, src/linAlg/rayleighRitzDevice.cc:81-353)
static int cgs_rr(dftfe_b200_ctx *ctx, double *X, int N, double *eig_h, const RRFlags &f = RRFlags()) {
  if (ctx->cplx) return rr_cplx(ctx, X, N, eig_h, true, f);
  DB_TRY(ctx->eigDev.alloc(N));
  DB_TRY(cgs_orthogonalise(ctx, X, N, f));
  double *Hp = ctx->denseB.p;
  DB_TRY(projham_any(ctx, X, N, Hp, f.mpXtHX, f.nCore, f.commOnly));
  DB_TRY(dense_eigh(ctx, Hp, N, ctx->eigDev.p));
  DB_TRY(rotate_impl(ctx, X, N, Hp, true, f.mpRRRot ? 2 : 0));
  DB_CUDA(cudaMemcpyAsync(eig_h, ctx->eigDev.p, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// rayleighRitzGEPSpectrumSplitDirect (src/linAlg/rayleighRitzDevice.cc:821-1454): X is orthonormalised
// (X <- X L^-T) but NOT rotated; only the top Nfr = N - Noc eigenpairs of the projected Hamiltonian are kept and
// XFrac (M x Nfr) = X Q[:, Noc:N] receives the rotated fractionally-occupied states.  eig_h: Nfr values.
static int rr_spectrum_split(dftfe_b200_ctx *ctx, double *X, double *XFrac, int N, int Noc, double *eig_h,
                             const RRFlags &f) {
  DB_CHECK(Noc > 0 && Noc < N && XFrac, "spectrum split needs 0 < n_core_states < N and an XFrac buffer");
  DB_TRY(ctx->eigDev.alloc(N));
  double *Hp = nullptr;
  if (ctx->cplx) {
    const size_t nn = (size_t)N * N * 2;
    DB_TRY(ctx->denseA.alloc(nn));
    DB_TRY(ctx->denseB.alloc(nn));
    double *S = ctx->denseA.p;
    Hp = ctx->denseB.p;
    cuDoubleComplex *Sz = reinterpret_cast<cuDoubleComplex *>(S), *Hz = reinterpret_cast<cuDoubleComplex *>(Hp);
    const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0);
    DB_TRY(overlap_any(ctx, X, N, S, f.mpOverlap, f.commOnly));
    DB_TRY(dense_cholesky_cplx(ctx, S, N));
    std::vector<double> eye(nn, 0.0);
    for (int i = 0; i < N; ++i) eye[((size_t)i * N + i) * 2] = 1.0;
    DB_CUDA(cudaMemcpyAsync(Hp, eye.data(), nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DB_CUDA(cudaStreamSynchronize(ctx->stream));
    DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
    ctx->launches += 1;
    DB_CUBLAS(cublasZtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT, N,
                          N, &one, Sz, N, Hz, N));
    DB_TRY(rotate_impl(ctx, X, N, Hp, true, f.mpCgsRot ? 1 : 0));
    DB_TRY(projham_any(ctx, X, N, Hp, f.mpXtHX, Noc, f.commOnly));
    DB_TRY(dense_eigh_cplx(ctx, Hp, N, ctx->eigDev.p));
  } else {
    DB_TRY(cgs_orthogonalise(ctx, X, N, f));
    Hp = ctx->denseB.p;
    DB_TRY(projham_any(ctx, X, N, Hp, f.mpXtHX, Noc, f.commOnly));
    DB_TRY(dense_eigh(ctx, Hp, N, ctx->eigDev.p));
  }
  DB_TRY(rotate_into(ctx, X, N, Hp, Noc, N - Noc, XFrac));
  DB_CUDA(cudaMemcpyAsync(eig_h, ctx->eigDev.p + Noc, (N - Noc) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// chebyshevOrthogonalizedSubspaceIterationSolverDevice::densityMatrixEigenBasisFirstOrderResponse (solver .cc:1084-1196
// -> src/linAlg/rayleighRitzDevice.cc:1456-1737): X <- M^1/2 X;  H'p = X^H H' X with the caller's H' cell matrices and
// without the non-local term (onlyHPrime);  D = first-order response of the Fermi operator in the eigenbasis
// (recursive expansion, m = 10 levels);  X <- X D;  X <- M^-1/2 X.  densityMatDerFermiEnergy = beta x0 (1 - x0).
static int first_order_response_impl(dftfe_b200_ctx *ctx, double *X, int N, const double *eig_h, double fermiEnergy,
                                     double temperature, bool singlePrec, double *dmDer_h) {
  const int cm = ctx->cm;
  const size_t nn = (size_t)N * N * cm;
  DB_TRY(ctx->denseA.alloc(nn));
  DB_TRY(ctx->denseB.alloc(nn));
  DB_TRY(ctx->eigDev.alloc(N));
  const int levels = 10;
  const double kb = 3.166811429e-06;  // C_kb, include/constants.h:30 (Ha / K)
  const double beta = 1.0 / kb / temperature;
  const double c = std::pow(2.0, -2.0 - levels) * beta;
  std::vector<double> x0(N);
  for (int i = 0; i < N; ++i) x0[i] = 0.5 - c * (eig_h[i] - fermiEnergy);
  DB_CUDA(cudaMemcpyAsync(ctx->eigDev.p, x0.data(), N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  for (int l = 0; l < levels; ++l)
    for (int i = 0; i < N; ++i) x0[i] = x0[i] * x0[i] / (2.0 * x0[i] * (x0[i] - 1.0) + 1.0);
  if (dmDer_h)
    for (int i = 0; i < N; ++i) dmDer_h[i] = beta * x0[i] * (1.0 - x0[i]);
  DB_TRY(launch_row_scale(ctx, X, ctx->M, N * cm, N * cm, 1.0, ctx->sqrtM.p));
  const bool skipWas = ctx->skip_nonlocal;
  ctx->skip_nonlocal = true;
  // singlePrecLRD: XtHXMixedPrecOverlapComputeCommun with Noc = N, i.e. every block in FP32 (rayleighRitzDevice.cc:1509-1523)
  int rc = projham_any(ctx, X, N, ctx->denseB.p, singlePrec, N);
  ctx->skip_nonlocal = skipWas;
  if (rc != 0) return rc;
  ctx->launches += 1;
  fermi_response_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->denseB.p, N, cm, ctx->eigDev.p, c, levels,
                                                                   ctx->denseA.p);
  DB_CUDA(cudaGetLastError());
  // subspaceRotation[RRMixedPrec]Scalapack with D^H (:1670-1700): X <- X D
  DB_TRY(rotate_impl(ctx, X, N, ctx->denseA.p, true, singlePrec ? 2 : 0));
  DB_TRY(launch_row_scale(ctx, X, ctx->M, N * cm, N * cm, 1.0, ctx->invSqrtM.p));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// HOT LOOP 1 of solve() (solver .cc:376-526): slice block -> filter -> write back, for every block of B
// columns of X (row-major M x N), X resident in HBM (Xd) or in host memory (Xh, pinned recommended).
//
// Two blocks are kept in flight on two lanes (stream + scratch + exchange buffers each), their degrees
// enqueued alternately: while one block's cell kernels own the SMs the other block's pack / NCCL send-recv /
// unpack run beside them - the reference's overlapComputeCommunCheby two-block schedule
// (linearAlgebraOperationsDevice.cc:734-1443) expressed with streams instead of a hand-interleaved 20-step
// sequence - and the tail of every colour launch (SMs idle while the last items finish) is filled by the other
// lane's kernel, which is worth 2.6 % even on one GPU.  The arithmetic per block is unchanged, so results are
// bit-identical to the single-lane loop.  Host-resident X: the next pair's H2D copies and the previous pair's
// D2H copies run on two copy streams under the current pair's kernels (two sets of lane buffers).
static int ensure_lanes(dftfe_b200_ctx *ctx, bool hostSets) {
  for (int l = 0; l < 2; ++l) {
    if (!ctx->laneStream[l]) DB_CUDA(cudaStreamCreateWithFlags(&ctx->laneStream[l], cudaStreamNonBlocking));
    if (!ctx->laneEvent[l]) DB_CUDA(cudaEventCreateWithFlags(&ctx->laneEvent[l], cudaEventDisableTiming));
  }
  if (!ctx->forkEvent) DB_CUDA(cudaEventCreateWithFlags(&ctx->forkEvent, cudaEventDisableTiming));
  const size_t blk = (size_t)(ctx->M + ctx->G) * ctx->B * ctx->cm;
  DB_TRY(ctx->blockX2.alloc(blk));
  DB_TRY(ctx->blockY2.alloc(blk));
  if (hostSets) {
    DB_TRY(ctx->blockX3.alloc(blk));
    DB_TRY(ctx->blockX4.alloc(blk));
    if (!ctx->copyIn) {
      DB_CUDA(cudaStreamCreateWithFlags(&ctx->copyIn, cudaStreamNonBlocking));
      DB_CUDA(cudaStreamCreateWithFlags(&ctx->copyOut, cudaStreamNonBlocking));
    }
  }
  return 0;
}

static int filter_blocks_impl(dftfe_b200_ctx *ctx, double *Xd, double *Xh, int N, int m, double a, double b,
                              double a0, const double *inScale, bool mixedPrec) {
  const int B = std::min(ctx->B, N);
  DB_CHECK(N % B == 0, "number of wavefunctions (%d) must be a multiple of the Chebyshev block size (%d)", N, B);
  DB_CHECK(m >= 1, "Chebyshev degree must be >= 1");
  DB_TRY(ensure_block_scratch(ctx));
  const int cm = ctx->cm;
  // band parallelisation: this band group filters the blocks that end inside its column range (solver .cc:394-400)
  std::vector<int> myBlocks;
  {
    int lo = 0, hi = N;
    if (ctx->nBandGroups > 1) band_group_range(ctx->nBandGroups, N, ctx->bandId, lo, hi);
    for (int j = 0; j < N; j += B)
      if (j + B <= hi && j + B > lo) myBlocks.push_back(j / B);
  }
  const int nb = (int)myBlocks.size();
  const bool host = Xh != nullptr;
  const bool lanes = ctx->overlap_lanes != 0 && nb >= 2;
  const int nl = lanes ? 2 : 1;
  DB_TRY(ensure_lanes(ctx, host));
  cudaStream_t mainStream = ctx->stream;
  // lane buffers: [set][lane]; the device-resident loop needs one set only
  double *bx[2][2] = {{ctx->blockX.p, ctx->blockX2.p}, {ctx->blockX3.p, ctx->blockX4.p}};
  double *by[2] = {ctx->blockY.p, ctx->blockY2.p};
  const size_t rowBytes = (size_t)B * cm * sizeof(double), pitchBytes = (size_t)N * cm * sizeof(double);
  const int nGroups = (nb + nl - 1) / nl;
  // events of the host-resident loop live in the context (created once, reused by every call)
  cudaEvent_t *evIn = nullptr, *evComp = nullptr, *evOut = nullptr;
  if (host) {
    while ((int)ctx->hostLoopEvents.size() < 3 * nb) {
      cudaEvent_t e;
      DB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->hostLoopEvents.push_back(e);
    }
    evIn = ctx->hostLoopEvents.data();
    evComp = evIn + nb;
    evOut = evComp + nb;
  }
  DB_CUDA(cudaEventRecord(ctx->forkEvent, mainStream));
  for (int l = 0; l < nl; ++l) DB_CUDA(cudaStreamWaitEvent(ctx->laneStream[l], ctx->forkEvent, 0));
  if (host) DB_CUDA(cudaStreamWaitEvent(ctx->copyIn, ctx->forkEvent, 0));
  auto on_lane = [&](int l) {
    ctx->lane = l;
    ctx->stream = ctx->laneStream[l];
  };
  auto body = [&]() -> int {
    for (int g = 0; g < nGroups; ++g) {
      const int set = host ? (g & 1) : 0;
      const int nCur = std::min(nl, nb - g * nl);
      ChebState st[2];
      for (int l = 0; l < nCur; ++l) {
        const int blkIdx = myBlocks[g * nl + l];
        const int ev = g * nl + l;  // event slot
        double *buf = bx[set][l];
        on_lane(l);
        if (host) {
          // the buffer is free once the block that used it two groups ago has been copied out
          if (g >= 2) DB_CUDA(cudaStreamWaitEvent(ctx->copyIn, evOut[(g - 2) * nl + l], 0));
          DB_CUDA(cudaMemcpy2DAsync(buf, rowBytes, Xh + (size_t)blkIdx * B * cm, pitchBytes, rowBytes, (size_t)ctx->M,
                                    cudaMemcpyHostToDevice, ctx->copyIn));
          DB_CUDA(cudaEventRecord(evIn[ev], ctx->copyIn));
          DB_CUDA(cudaStreamWaitEvent(ctx->stream, evIn[ev], 0));
          if (inScale) DB_TRY(launch_row_scale(ctx, buf, ctx->M, B * cm, B * cm, 1.0, inScale));
        } else {
          DB_TRY(launch_block_copy_from_full(ctx, Xd, N * cm, blkIdx * B * cm, buf, B * cm, ctx->M, inScale));
        }
        DB_TRY(ghost_zero(ctx, buf, B * cm, B * cm));
        cheb_begin(st[l], buf, by[l], a, b, a0);
      }
      for (int degree = 1; degree <= m; ++degree)
        for (int l = 0; l < nCur; ++l) {
          on_lane(l);
          DB_TRY(cheb_step(ctx, st[l], B, m, mixedPrec));
        }
      for (int l = 0; l < nCur; ++l) {
        const int blkIdx = myBlocks[g * nl + l];
        const int ev = g * nl + l;
        double *buf = bx[set][l];
        on_lane(l);
        DB_TRY(cheb_end(ctx, st[l], buf, B));
        if (host) {
          DB_CUDA(cudaEventRecord(evComp[ev], ctx->stream));
          DB_CUDA(cudaStreamWaitEvent(ctx->copyOut, evComp[ev], 0));
          DB_CUDA(cudaMemcpy2DAsync(Xh + (size_t)blkIdx * B * cm, pitchBytes, buf, rowBytes, rowBytes, (size_t)ctx->M,
                                    cudaMemcpyDeviceToHost, ctx->copyOut));
          DB_CUDA(cudaEventRecord(evOut[ev], ctx->copyOut));
        } else {
          DB_TRY(launch_block_copy_to_full(ctx, Xd, N * cm, blkIdx * B * cm, buf, B * cm, ctx->M, nullptr));
        }
      }
    }
    return 0;
  };
  int rc = body();
  ctx->lane = 0;
  ctx->stream = mainStream;
  if (rc == 0) {
    for (int l = 0; l < nl && rc == 0; ++l) {
      if (cudaEventRecord(ctx->laneEvent[l], ctx->laneStream[l]) != cudaSuccess ||
          cudaStreamWaitEvent(mainStream, ctx->laneEvent[l], 0) != cudaSuccess)
        rc = DFTFE_B200_ERR_CUDA;
    }
    if (host) {
      // the call returns with the result in X_h: wait for every copy-out
      for (int i = std::max(0, nb - 2 * nl); i < nb; ++i) cudaStreamWaitEvent(mainStream, evOut[i], 0);
      if (cudaStreamSynchronize(mainStream) != cudaSuccess) rc = DFTFE_B200_ERR_CUDA;
      if (rc == DFTFE_B200_ERR_CUDA) set_error("host-resident filter loop: stream synchronisation failed");
    }
  } else {
    cudaDeviceSynchronize();
  }
  return rc;
}

static int filter_all_impl(dftfe_b200_ctx *ctx, double *X, int N, int m, double a, double b, double a0,
                           const double *inScale, bool mixedPrec = false) {
  DB_TRY(filter_blocks_impl(ctx, X, nullptr, N, m, a, b, a0, inScale, mixedPrec));
  // band groups: the columns of the other groups (untouched here, not even by the M^1/2 scaling folded into the
  // copy-in) are replaced by their owners' filtered columns; a no-op for one band group
  return band_group_merge(ctx, X, N);
}

static int filter_all_host_impl(dftfe_b200_ctx *ctx, double *X_h, int N, int m, double a, double b, double a0,
                                bool mixedPrec) {
  DB_CHECK(ctx->nBandGroups == 1, "cheb_filter_all_host: band groups need the device-resident X (band_group_merge)");
  return filter_blocks_impl(ctx, nullptr, X_h, N, m, a, b, a0, nullptr, mixedPrec);
}

static const unsigned int order_lookup[][2] = {{500, 24},     {750, 30},     {1000, 39},    {1500, 50},
                                               {2000, 53},    {3000, 57},    {4000, 62},    {5000, 69},
                                               {9000, 77},    {14000, 104},  {20000, 119},  {30000, 162},
                                               {50000, 300},  {80000, 450},  {100000, 550}, {200000, 700},
                                               {500000, 1000}};

static unsigned int set_chebyshev_order(double upperBound) {
  for (const auto &row : order_lookup)
    if (upperBound <= row[0]) return row[1];
  return 1250;
}

static int solve_impl(dftfe_b200_ctx *ctx, double *X, double *XFrac, int N, const dftfe_b200_solve_params *p,
                      double *eig_h, double *res_h, double *upper_h) {
  const int B = std::min(ctx->B, N);
  DB_CHECK(N % B == 0, "number of wavefunctions (%d) must be a multiple of the Chebyshev block size (%d)", N, B);
  const int Noc = p->n_core_states;
  DB_CHECK(Noc >= 0 && Noc < N, "solve: n_core_states (%d) must be in [0, N)", Noc);
  DB_CHECK(Noc == 0 || XFrac, "solve: spectrum splitting (n_core_states > 0) needs the X_frac_d buffer");
  DB_TRY(ensure_block_scratch(ctx));
  // spectrum bounds (solver .cc:241-297)
  if (p->is_first_filtering_call || !ctx->bounds_valid) {
    double bounds[2];
    DB_TRY(lanczos_impl(ctx, p->reproducible_output, bounds));
    ctx->a0 = bounds[0];
    ctx->bUp = bounds[1];
    ctx->bLow = ctx->a0 + (ctx->bUp - ctx->a0) * (double)N / (double)ctx->desc.n_global_dofs *
                              (p->reproducible_output ? 10.0 : 200.0);
    ctx->bounds_valid = true;
  } else if (!p->reuse_lanczos_upper_bound) {
    double bounds[2];
    DB_TRY(lanczos_impl(ctx, p->reproducible_output, bounds));
    ctx->bUp = bounds[1];
  }
  unsigned int order = p->chebyshev_order;
  if (order == 0) {
    order = set_chebyshev_order(ctx->bUp);
    if (!p->is_pseudopotential) {
      // orthogType is always "CGS" on the device path (Auto -> CGS, GS throws) whichever projection is used:
      // all-electron runs filter at half the tabulated degree (solver .cc:314-316)
      order = (unsigned int)(order * 0.5);
    }
  }
  if (p->is_first_scf && p->is_pseudopotential) order = (unsigned int)(order * p->first_scf_scaling);
  if (order < 1) order = 1;

  const bool mp = p->use_mixed_prec_overall != 0;
  RRFlags f;
  f.mpOverlap = mp && p->use_mixed_prec_cgs_o;
  f.mpCgsRot = mp && p->use_mixed_prec_cgs_sr;
  f.mpXtHX = mp && p->use_mixed_prec_xthx_spectrum_split;
  f.mpRRRot = mp && p->use_mixed_prec_subspace_rot_rr;
  f.nCore = p->num_core_wfc_xthx;
  f.commOnly = p->use_mixed_prec_commun_only_xthx_cgs_o != 0;

  // X <- M^1/2 X (solver .cc:358-363) fused into the block copy; filter; copy back (:376-526)
  DB_TRY(filter_all_impl(ctx, X, N, (int)order, ctx->bLow, ctx->bUp, ctx->a0, ctx->sqrtM.p,
                         mp && p->use_mixed_prec_cheby));
  const int Nev = N - Noc;  // eigenValues.size() of the reference
  if (Noc > 0) {
    DB_TRY(rr_spectrum_split(ctx, X, XFrac, N, Noc, eig_h, f));  // solver .cc:553-575
  } else if (p->use_cgs_rr) {
    DB_TRY(cgs_rr(ctx, X, N, eig_h, f));
  } else {
    DB_TRY(rr_gep(ctx, X, N, eig_h, f));
  }
  if (p->compute_residual && res_h) DB_TRY(residual_impl(ctx, Noc > 0 ? XFrac : X, Nev, eig_h, res_h));
  // X <- M^-1/2 X (solver .cc:719-733)
  DB_TRY(launch_row_scale(ctx, X, ctx->M, N * ctx->cm, N * ctx->cm, 1.0, ctx->invSqrtM.p));
  if (Noc > 0) DB_TRY(launch_row_scale(ctx, XFrac, ctx->M, Nev * ctx->cm, Nev * ctx->cm, 1.0, ctx->invSqrtM.p));
  if (upper_h) *upper_h = ctx->bUp;
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// chebyshevOrthogonalizedSubspaceIterationSolverDevice::solveNoRR (solver .cc:742-1071): numberPasses x
// (filter every block, Cholesky-Gram-Schmidt), no Rayleigh-Ritz step, no eigenvalues.
static int solve_no_rr_impl(dftfe_b200_ctx *ctx, double *X, int N, const dftfe_b200_solve_params *p, int numberPasses,
                            double *upper_h) {
  const int B = std::min(ctx->B, N);
  DB_CHECK(N % B == 0, "number of wavefunctions (%d) must be a multiple of the Chebyshev block size (%d)", N, B);
  DB_CHECK(numberPasses >= 1, "solveNoRR: numberPasses must be >= 1");
  DB_CHECK(ctx->bounds_valid, "solveNoRR: spectrum bounds are not set (call solve() with is_first_filtering_call "
                              "or reinit_spectrum_bounds + lanczos first)");
  DB_TRY(ensure_block_scratch(ctx));
  if (!p->reuse_lanczos_upper_bound) {  // :813-826
    double bounds[2];
    DB_TRY(lanczos_impl(ctx, p->reproducible_output, bounds));
    ctx->bUp = bounds[1];
  }
  unsigned int order = p->chebyshev_order;
  if (order == 0) order = set_chebyshev_order(ctx->bUp);
  if (p->is_pseudopotential) order = (unsigned int)(order * p->first_scf_scaling);  // :836-840, unconditional here
  if (order < 1) order = 1;
  const bool mp = p->use_mixed_prec_overall != 0;
  RRFlags f;
  f.mpOverlap = mp && p->use_mixed_prec_cgs_o;
  f.mpCgsRot = mp && p->use_mixed_prec_cgs_sr;
  DB_TRY(launch_row_scale(ctx, X, ctx->M, N * ctx->cm, N * ctx->cm, 1.0, ctx->sqrtM.p));
  for (int ipass = 0; ipass < numberPasses; ++ipass) {
    DB_TRY(filter_all_impl(ctx, X, N, (int)order, ctx->bLow, ctx->bUp, ctx->a0, nullptr, mp && p->use_mixed_prec_cheby));
    if (ctx->cplx) {
      // complex CGS: S = X^H X = L L^H, X <- X L^-H
      const size_t nn = (size_t)N * N * 2;
      DB_TRY(ctx->denseA.alloc(nn));
      DB_TRY(ctx->denseB.alloc(nn));
      cuDoubleComplex *Sz = reinterpret_cast<cuDoubleComplex *>(ctx->denseA.p);
      cuDoubleComplex *Uz = reinterpret_cast<cuDoubleComplex *>(ctx->denseB.p);
      const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0);
      DB_TRY(overlap_any(ctx, X, N, ctx->denseA.p, f.mpOverlap, f.commOnly));
      DB_TRY(dense_cholesky_cplx(ctx, ctx->denseA.p, N));
      std::vector<double> eye(nn, 0.0);
      for (int i = 0; i < N; ++i) eye[((size_t)i * N + i) * 2] = 1.0;
      DB_CUDA(cudaMemcpyAsync(ctx->denseB.p, eye.data(), nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      DB_CUDA(cudaStreamSynchronize(ctx->stream));
      DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
      ctx->launches += 1;
      DB_CUBLAS(cublasZtrsm(ctx->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT, N,
                            N, &one, Sz, N, Uz, N));
      DB_TRY(rotate_impl(ctx, X, N, ctx->denseB.p, true, f.mpCgsRot ? 1 : 0));
    } else {
      DB_TRY(cgs_orthogonalise(ctx, X, N, f));
    }
  }
  DB_TRY(launch_row_scale(ctx, X, ctx->M, N * ctx->cm, N * ctx->cm, 1.0, ctx->invSqrtM.p));
  if (upper_h) *upper_h = ctx->bUp;
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // namespace dftfe_b200

// ===========================================================================
// C ABI
// ===========================================================================
using namespace dftfe_b200;

#define DB_CTX(ctx)                                  \
  do {                                               \
    if (!(ctx)) {                                    \
      set_error("null context");                     \
      return DFTFE_B200_ERR_INVALID;                 \
    }                                                \
    cudaError_t e__ = cudaSetDevice((ctx)->desc.device); \
    if (e__ != cudaSuccess) {                        \
      set_error("cudaSetDevice failed: %s", cudaGetErrorString(e__)); \
      return DFTFE_B200_ERR_CUDA;                    \
    }                                                \
  } while (0)

static int check_cols(dftfe_b200_ctx *ctx, int ncols) {
  DB_CHECK(ncols >= 1 && ncols <= ctx->B, "ncols (%d) must be in [1, cheby_block=%d]", ncols, ctx->B);
  return 0;
}

extern "C" {

int dftfe_b200_hx(dftfe_b200_ctx *ctx, double *src_d, double *dst_d, int32_t ncols, int32_t scale_flag,
                  double scalar, int32_t do_unscaling_src, int32_t single_prec_commun) {
  DB_CTX(ctx);
  DB_TRY(check_cols(ctx, ncols));
  DB_CHECK(scalar != 0.0, "HX: scalar must be non-zero");
  return op_hx(ctx, src_d, dst_d, ncols, scale_flag, scalar, do_unscaling_src, single_prec_commun != 0);
}

int dftfe_b200_hx_cheby(dftfe_b200_ctx *ctx, double *src_d, double *dst_d, int32_t ncols, int32_t mixed_prec) {
  DB_CTX(ctx);
  DB_TRY(check_cols(ctx, ncols));
  return op_hx_cheby(ctx, src_d, dst_d, ncols, mixed_prec != 0);
}

int dftfe_b200_cheb_filter(dftfe_b200_ctx *ctx, double *x_d, double *y_d, int32_t ncols, int32_t m, double a,
                           double b, double a0, int32_t mixed_prec) {
  DB_CTX(ctx);
  DB_TRY(check_cols(ctx, ncols));
  return cheb_filter_impl(ctx, x_d, y_d, ncols, m, a, b, a0, mixed_prec != 0);
}

int dftfe_b200_cheb_filter_all(dftfe_b200_ctx *ctx, double *X_d, int32_t N, int32_t m, double a, double b,
                               double a0, int32_t mixed_prec) {
  DB_CTX(ctx);
  return filter_all_impl(ctx, X_d, N, m, a, b, a0, nullptr, mixed_prec != 0);
}

int dftfe_b200_cheb_filter_all_host(dftfe_b200_ctx *ctx, double *X_h, int32_t N, int32_t m, double a, double b,
                                    double a0, int32_t mixed_prec) {
  DB_CTX(ctx);
  DB_CHECK(X_h, "cheb_filter_all_host: null host pointer");
  return filter_all_host_impl(ctx, X_h, N, m, a, b, a0, mixed_prec != 0);
}

static int to_row_major_cplx(dftfe_b200_ctx *ctx, double *S_d, int N) {
  DB_TRY(ctx->denseW.alloc((size_t)N * N * 2));
  ctx->launches += 2;
  cplx_transpose_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(S_d, ctx->denseW.p, N);
  DB_CUDA(cudaGetLastError());
  DB_CUDA(cudaMemcpyAsync(S_d, ctx->denseW.p, (size_t)N * N * 2 * sizeof(double), cudaMemcpyDeviceToDevice,
                          ctx->stream));
  return 0;
}

int dftfe_b200_xtx(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, double *S_d, int32_t mixed_prec) {
  DB_CTX(ctx);
  DB_CHECK(mixed_prec >= 0 && mixed_prec <= 2, "xtx: mixed_prec must be 0, 1 or 2");
  DB_TRY(overlap_any(ctx, X_d, N, S_d, mixed_prec != 0, mixed_prec == 2));
  return ctx->cplx ? to_row_major_cplx(ctx, S_d, N) : 0;
}

int dftfe_b200_xthx(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, int32_t n_core, double *Hp_d,
                    int32_t mixed_prec) {
  DB_CTX(ctx);
  DB_CHECK(n_core >= 0 && n_core <= N, "xthx: n_core (%d) must be in [0, N]", n_core);
  DB_CHECK(mixed_prec >= 0 && mixed_prec <= 2, "xthx: mixed_prec must be 0, 1 or 2");
  DB_TRY(projham_any(ctx, X_d, N, Hp_d, mixed_prec != 0, n_core, mixed_prec == 2));
  return ctx->cplx ? to_row_major_cplx(ctx, Hp_d, N) : 0;
}

int dftfe_b200_rotate(dftfe_b200_ctx *ctx, double *X_d, int32_t N, const double *Q_d, int32_t mixed_mode) {
  DB_CTX(ctx);
  DB_CHECK(mixed_mode >= 0 && mixed_mode <= 2, "rotate: mixed_mode must be 0, 1 or 2");
  return rotate_impl(ctx, X_d, N, Q_d, false, mixed_mode);
}

int dftfe_b200_rotate_spectrum_split(dftfe_b200_ctx *ctx, double *X_d, int32_t N, const double *Q_d,
                                     int32_t n_frac, double *X_frac_d) {
  DB_CTX(ctx);
  DB_CHECK(n_frac > 0 && n_frac <= N && X_frac_d, "rotate_spectrum_split: bad n_frac / null output");
  // Q_d row-major N x N: the kernels want exactly that, so go straight to rotate_real on the column slice
  if (!ctx->cplx) return rotate_real(ctx, X_d, N, Q_d, false, N - n_frac, n_frac, X_frac_d, n_frac);
  const int Nr = 2 * N;
  DB_TRY(ctx->denseG.alloc((size_t)Nr * Nr));
  ctx->launches += 1;
  cplx_embed_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(Q_d, N, 0, ctx->denseG.p);
  DB_CUDA(cudaGetLastError());
  return rotate_real(ctx, X_d, Nr, ctx->denseG.p, false, 2 * (N - n_frac), 2 * n_frac, X_frac_d, 2 * n_frac);
}

int dftfe_b200_lanczos_bounds(dftfe_b200_ctx *ctx, int32_t reproducible, double bounds_out_h[2]) {
  DB_CTX(ctx);
  return lanczos_impl(ctx, reproducible, bounds_out_h);
}

int dftfe_b200_residual_norms(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, const double *eig_h,
                              double *res_out_h) {
  DB_CTX(ctx);
  return residual_impl(ctx, X_d, N, eig_h, res_out_h);
}

int dftfe_b200_reinit_spectrum_bounds(dftfe_b200_ctx *ctx, double lower_wanted, double lower_unwanted) {
  DB_CTX(ctx);
  ctx->a0 = lower_wanted;
  ctx->bLow = lower_unwanted;
  return 0;  // bounds_valid still needs an upper bound: the first solve() / lanczos call provides it
}

int dftfe_b200_solve_no_rr(dftfe_b200_ctx *ctx, double *X_d, int32_t N, const dftfe_b200_solve_params *params,
                           int32_t number_passes, double *upper_bound_out_h) {
  DB_CTX(ctx);
  DB_CHECK(params, "solve_no_rr: params are required");
  return solve_no_rr_impl(ctx, X_d, N, params, number_passes, upper_bound_out_h);
}

int dftfe_b200_density_matrix_first_order_response(dftfe_b200_ctx *ctx, double *X_d, int32_t N, const double *eig_h,
                                                   double fermi_energy, double temperature, int32_t single_prec,
                                                   double *dm_der_fermi_out_h) {
  DB_CTX(ctx);
  DB_CHECK(X_d && eig_h && N > 0 && temperature > 0.0, "first_order_response: bad arguments");
  return first_order_response_impl(ctx, X_d, N, eig_h, fermi_energy, temperature, single_prec != 0, dm_der_fermi_out_h);
}

int dftfe_b200_get_spectrum_bounds(dftfe_b200_ctx *ctx, double out_h[3]) {
  DB_CTX(ctx);
  out_h[0] = ctx->a0;
  out_h[1] = ctx->bLow;
  out_h[2] = ctx->bUp;
  return 0;
}

int dftfe_b200_solve(dftfe_b200_ctx *ctx, double *X_d, double *X_frac_d, int32_t N,
                     const dftfe_b200_solve_params *params, double *eig_out_h, double *res_out_h,
                     double *upper_bound_out_h) {
  DB_CTX(ctx);
  DB_CHECK(params && eig_out_h, "solve: params and eig_out_h are required");
  return solve_impl(ctx, X_d, X_frac_d, N, params, eig_out_h, res_out_h, upper_bound_out_h);
}

}  // extern "C"
