// HBM-bound row kernels: constraints, ghost pack/unpack, block slicing, scalings.
// All are coalesced along the wavefunction index (rows are contiguous runs of
// ncols doubles), grid-stride, grid sized in multiples of the SM count, and
// atomics-free: the two "transpose" operations of the reference that use
// atomicAdd (distributeSlaveToMasterKernelAtomicAdd,
// utils/constraintMatrixInfoDevice.cc:237-328, and accumAddFromRecvBufferDeviceKernel,
// utils/MPICommunicatorP2PKernelsDevice.cc:88-163) are restated as gathers over a
// transposed map built once on the host, which also fixes the summation order.
#include <algorithm>

#include "common.cuh"

namespace dftfe_b200 {

namespace {

inline int grid_for(const dftfe_b200_ctx *ctx, int64_t work, int block = 256) {
  int64_t g = (work + block - 1) / block;
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// K11 distributeKernel: x[row,:] = inhom + sum_j w_j * x[col_j,:]
// Products and sums are rounded separately, in CSR order (the oracle's statement).
__global__ void distribute_kernel(double *__restrict__ x, int ncols, int ldx, int64_t nCon,
                                  const uint32_t *__restrict__ rows, const uint32_t *__restrict__ sizes,
                                  const uint32_t *__restrict__ starts, const uint32_t *__restrict__ cols,
                                  const double *__restrict__ vals, const double *__restrict__ inhom,
                                  const double *__restrict__ colScale) {
  const int64_t total = nCon * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / ncols;
    const int c = idx % ncols;
    const uint32_t s = starts[i], nz = sizes[i];
    double v = inhom[i];
    for (uint32_t j = 0; j < nz; ++j) {
      const uint32_t cj = cols[s + j];
      double xv = x[(size_t)cj * ldx + c];
      if (colScale) xv = __dmul_rn(xv, colScale[cj]);
      v = __dadd_rn(v, __dmul_rn(vals[s + j], xv));
    }
    x[(size_t)rows[i] * ldx + c] = v;
  }
}

// K12 distribute_slave_to_master, gather form: for each master row m,
// x[m,:] += scale[m] * (w_1 x[s_1,:]) ... in ascending constraint order.
__global__ void slave_to_master_kernel(double *__restrict__ x, int ncols, int ldx, int64_t nMasters,
                                       const uint32_t *__restrict__ masters, const uint32_t *__restrict__ mstarts,
                                       const uint32_t *__restrict__ slaves, const double *__restrict__ vals,
                                       const double *__restrict__ masterScale) {
  const int64_t total = nMasters * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / ncols;
    const int c = idx % ncols;
    const uint32_t m = masters[i];
    const double sc = masterScale ? masterScale[m] : 1.0;
    double v = x[(size_t)m * ldx + c];
    for (uint32_t j = mstarts[i]; j < mstarts[i + 1]; ++j) {
      double t = __dmul_rn(vals[j], x[(size_t)slaves[j] * ldx + c]);
      if (masterScale) t = __dmul_rn(t, sc);
      v = __dadd_rn(v, t);
    }
    x[(size_t)m * ldx + c] = v;
  }
}

// K13 setzeroKernel (also the "x[row]=0" tail of K12)
__global__ void zero_rows_kernel(double *__restrict__ x, int ncols, int ldx, int64_t nRows,
                                 const uint32_t *__restrict__ rows) {
  const int64_t total = nRows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    x[(size_t)rows[idx / ncols] * ldx + (idx % ncols)] = 0.0;
}

// K14 gatherSendBufferDeviceKernel: send[k,:] = x[sendRows[k],:]
__global__ void pack_rows_kernel(const double *__restrict__ x, int ncols, int ldx, int64_t nRows,
                                 const uint32_t *__restrict__ rows, double *__restrict__ buf) {
  const int64_t total = nRows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    buf[idx] = x[(size_t)rows[idx / ncols] * ldx + (idx % ncols)];
}

// contiguous rows [row0,row0+nRows) <-> dense buffer (ghost segment when ldx != ncols)
__global__ void copy_rows_kernel(double *__restrict__ x, int ncols, int ldx, int64_t row0, int64_t nRows,
                                 double *__restrict__ buf, int toBuf) {
  const int64_t total = nRows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double *px = x + (size_t)(row0 + idx / ncols) * ldx + (idx % ncols);
    if (toBuf)
      buf[idx] = *px;
    else
      *px = buf[idx];
  }
}

// K15 accumAddFromRecvBuffer, gather form: boundary row r sums its slots in a fixed order
__global__ void unpack_add_kernel(double *__restrict__ x, int ncols, int ldx, int64_t nRows,
                                  const uint32_t *__restrict__ rows, const uint32_t *__restrict__ starts,
                                  const uint32_t *__restrict__ slots, const double *__restrict__ buf,
                                  const double *__restrict__ rowScale) {
  const int64_t total = nRows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / ncols;
    const int c = idx % ncols;
    const uint32_t r = rows[i];
    const double sc = rowScale ? rowScale[r] : 1.0;
    double v = x[(size_t)r * ldx + c];
    for (uint32_t j = starts[i]; j < starts[i + 1]; ++j) {
      double t = buf[(size_t)slots[j] * ncols + c];
      if (rowScale) t *= sc;
      v += t;
    }
    x[(size_t)r * ldx + c] = v;
  }
}

// K5 stridedBlockScale: x[r,:] *= alpha * s[r]
__global__ void row_scale_kernel(double *__restrict__ x, int64_t rows, int ncols, int ldx, double alpha,
                                 const double *__restrict__ s) {
  const int64_t total = rows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / ncols;
    const double f = s ? alpha * s[r] : alpha;
    x[(size_t)r * ldx + (idx % ncols)] *= f;
  }
}

// K8 stridedCopyToBlockConstantStride (+ optional fused row scale)
__global__ void block_from_full_kernel(const double *__restrict__ X, int N, int j0, double *__restrict__ blk,
                                       int ncols, int64_t rows, const double *__restrict__ s) {
  const int64_t total = rows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / ncols;
    double v = X[(size_t)r * N + j0 + (idx % ncols)];
    if (s) v *= s[r];
    blk[idx] = v;
  }
}

__global__ void block_to_full_kernel(double *__restrict__ X, int N, int j0, const double *__restrict__ blk,
                                     int ncols, int64_t rows, const double *__restrict__ s) {
  const int64_t total = rows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / ncols;
    double v = blk[idx];
    if (s) v *= s[r];
    X[(size_t)r * N + j0 + (idx % ncols)] = v;
  }
}

}  // namespace

int launch_distribute(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *colScale) {
  if (ctx->nCon == 0) return 0;
  ProfScope ps(ctx, "distribute");
  distribute_kernel<<<grid_for(ctx, ctx->nCon * ncols), 256, 0, ctx->stream>>>(
      x, ncols, ldx, ctx->nCon, ctx->conRows.p, ctx->conSizes.p, ctx->conStarts.p, ctx->conCols.p, ctx->conVals.p,
      ctx->conInhom.p, colScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_slave_to_master(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *masterScale) {
  if (ctx->nCon == 0) return 0;
  ProfScope ps(ctx, "slave_to_master", 2);
  if (ctx->nMasters > 0)
    slave_to_master_kernel<<<grid_for(ctx, ctx->nMasters * ncols), 256, 0, ctx->stream>>>(
        x, ncols, ldx, ctx->nMasters, ctx->masterRows.p, ctx->masterStarts.p, ctx->masterSlaves.p,
        ctx->masterVals.p, masterScale);
  zero_rows_kernel<<<grid_for(ctx, ctx->nCon * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, ctx->nCon,
                                                                             ctx->conRows.p);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_set_zero_rows(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx) {
  if (ctx->nCon == 0) return 0;
  ProfScope ps(ctx, "set_zero");
  zero_rows_kernel<<<grid_for(ctx, ctx->nCon * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, ctx->nCon,
                                                                             ctx->conRows.p);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_row_scale(dftfe_b200_ctx *ctx, double *x, int64_t rows, int ncols, int ldx, double alpha,
                     const double *rowScale) {
  if (rows == 0) return 0;
  ProfScope ps(ctx, "row_scale");
  row_scale_kernel<<<grid_for(ctx, rows * ncols), 256, 0, ctx->stream>>>(x, rows, ncols, ldx, alpha, rowScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_block_copy_from_full(dftfe_b200_ctx *ctx, const double *X, int N, int j0, double *blk, int ncols,
                                int64_t rows, const double *rowScale) {
  ProfScope ps(ctx, "block_copy");
  block_from_full_kernel<<<grid_for(ctx, rows * ncols), 256, 0, ctx->stream>>>(X, N, j0, blk, ncols, rows,
                                                                              rowScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_block_copy_to_full(dftfe_b200_ctx *ctx, double *X, int N, int j0, const double *blk, int ncols,
                              int64_t rows, const double *rowScale) {
  ProfScope ps(ctx, "block_copy");
  block_to_full_kernel<<<grid_for(ctx, rows * ncols), 256, 0, ctx->stream>>>(X, N, j0, blk, ncols, rows, rowScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// ---- kernels used by comm.cu ------------------------------------------------
int launch_pack_rows(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, int64_t nRows, const uint32_t *rows,
                     double *buf) {
  if (nRows == 0) return 0;
  ProfScope ps(ctx, "ghost_pack");
  pack_rows_kernel<<<grid_for(ctx, nRows * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, nRows, rows, buf);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_copy_rows(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, int64_t row0, int64_t nRows, double *buf,
                     int toBuf) {
  if (nRows == 0) return 0;
  ProfScope ps(ctx, toBuf ? "ghost_pack" : "ghost_unpack");
  copy_rows_kernel<<<grid_for(ctx, nRows * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, row0, nRows, buf, toBuf);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_unpack_add(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *buf,
                      const double *rowScale) {
  if (ctx->nBoundaryRows == 0) return 0;
  ProfScope ps(ctx, "ghost_unpack");
  unpack_add_kernel<<<grid_for(ctx, ctx->nBoundaryRows * ncols), 256, 0, ctx->stream>>>(
      x, ncols, ldx, ctx->nBoundaryRows, ctx->bndRows.p, ctx->bndStarts.p, ctx->bndSlots.p, buf, rowScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dftfe_b200
