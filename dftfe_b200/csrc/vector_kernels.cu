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

// K11 distributeKernel: x[row,:] = inhom + sum_j w_j * x[col_j,:], accumulated in CSR order with ONE rounding per
// term (fused multiply-add): the reference's `xVec[row] += w * xVec[col]` (utils/constraintMatrixInfoDevice.cc:71-74)
// is compiled by nvcc with its default -fmad=true into DFMA, and bit-exact parity is against that build
// (tests/test_gpu_reference_kernels.py runs the reference kernel itself).
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
      v = __fma_rn(vals[s + j], xv, v);
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


// ---------------------------------------------------------------------------
// Vector path (the one that runs for every even column count / leading dimension):
// one warp per row, 16-byte (double2) accesses, four independent row segments in
// flight per lane, no integer division.  A 256-column row is 2 KB = four 512-byte
// warp requests.  Arithmetic (and its order) is identical to the scalar kernels
// above, so the results are bit-identical between the two paths.
// ---------------------------------------------------------------------------
constexpr int VU = 4;  // double2 chunks in flight per lane

struct WarpRows {
  int lane;
  int64_t first, step;
  __device__ WarpRows() {
    lane = threadIdx.x & 31;
    first = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    step = (gridDim.x * (int64_t)blockDim.x) >> 5;
  }
};

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }
// streaming variants for data that is touched once per pass
__device__ __forceinline__ double2 ld2_stream(const double *p) {
  double2 v;
  asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st2_stream(double *p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__global__ void __launch_bounds__(256, 5)  // <= 48 registers: 8 warps within what a resident cell CTA leaves of the SM's registers
distribute_vec_kernel(double *__restrict__ x, int ncols, int ldx, int64_t nCon, const uint32_t *__restrict__ rows,
                      const uint32_t *__restrict__ sizes, const uint32_t *__restrict__ starts,
                      const uint32_t *__restrict__ cols, const double *__restrict__ vals,
                      const double *__restrict__ inhom, const double *__restrict__ colScale) {
  const WarpRows w;
  for (int64_t i = w.first; i < nCon; i += w.step) {
    const uint32_t s = starts[i], nz = sizes[i];
    const double ih = inhom[i];
    double *out = x + (size_t)rows[i] * ldx;
    for (int c0 = w.lane * 2; c0 < ncols; c0 += 64 * VU) {
      double2 v[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) v[u] = make_double2(ih, ih);
      for (uint32_t j = 0; j < nz; ++j) {
        const uint32_t cj = cols[s + j];
        const double wj = vals[s + j];
        const double sc = colScale ? colScale[cj] : 1.0;
        const double *in = x + (size_t)cj * ldx + c0;
        double2 xv[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) xv[u] = (c0 + 64 * u < ncols) ? ld2(in + 64 * u) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < VU; ++u) {
          if (colScale) {
            xv[u].x = __dmul_rn(xv[u].x, sc);
            xv[u].y = __dmul_rn(xv[u].y, sc);
          }
          v[u].x = __fma_rn(wj, xv[u].x, v[u].x);
          v[u].y = __fma_rn(wj, xv[u].y, v[u].y);
        }
      }
#pragma unroll
      for (int u = 0; u < VU; ++u)
        if (c0 + 64 * u < ncols) st2(out + c0 + 64 * u, v[u]);
    }
  }
}

__global__ void __launch_bounds__(256, 5)  // <= 48 registers: 8 warps within what a resident cell CTA leaves of the SM's registers
slave_to_master_vec_kernel(double *__restrict__ x, int ncols, int ldx, int64_t nMasters,
                           const uint32_t *__restrict__ masters, const uint32_t *__restrict__ mstarts,
                           const uint32_t *__restrict__ slaves, const double *__restrict__ vals,
                           const double *__restrict__ masterScale) {
  const WarpRows w;
  for (int64_t i = w.first; i < nMasters; i += w.step) {
    const uint32_t m = masters[i];
    const double sc = masterScale ? masterScale[m] : 1.0;
    const uint32_t j0 = mstarts[i], j1 = mstarts[i + 1];
    double *row = x + (size_t)m * ldx;
    for (int c0 = w.lane * 2; c0 < ncols; c0 += 64 * VU) {
      double2 v[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) v[u] = (c0 + 64 * u < ncols) ? ld2(row + c0 + 64 * u) : make_double2(0.0, 0.0);
      for (uint32_t j = j0; j < j1; ++j) {
        const double wj = vals[j];
        const double *in = x + (size_t)slaves[j] * ldx + c0;
        double2 t[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) t[u] = (c0 + 64 * u < ncols) ? ld2(in + 64 * u) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < VU; ++u) {
          t[u].x = __dmul_rn(wj, t[u].x);
          t[u].y = __dmul_rn(wj, t[u].y);
          if (masterScale) {
            t[u].x = __dmul_rn(t[u].x, sc);
            t[u].y = __dmul_rn(t[u].y, sc);
          }
          v[u].x = __dadd_rn(v[u].x, t[u].x);
          v[u].y = __dadd_rn(v[u].y, t[u].y);
        }
      }
#pragma unroll
      for (int u = 0; u < VU; ++u)
        if (c0 + 64 * u < ncols) st2(row + c0 + 64 * u, v[u]);
    }
  }
}

__global__ void __launch_bounds__(256)
zero_rows_vec_kernel(double *__restrict__ x, int ncols, int ldx, int64_t nRows, const uint32_t *__restrict__ rows) {
  const WarpRows w;
  for (int64_t i = w.first; i < nRows; i += w.step) {
    double *row = x + (size_t)rows[i] * ldx;
    for (int c = w.lane * 2; c < ncols; c += 64) st2(row + c, make_double2(0.0, 0.0));
  }
}

// send[k,:] = x[rows[k],:]   (TOUT = double, or float for the FP32 payload of the mixed-precision filter)
template <typename TOUT>
__global__ void __launch_bounds__(256)
pack_rows_vec_kernel(const double *__restrict__ x, int ncols, int ldx, int64_t nRows,
                     const uint32_t *__restrict__ rows, int64_t row0, TOUT *__restrict__ buf) {
  const WarpRows w;
  for (int64_t i = w.first; i < nRows; i += w.step) {
    const double *in = x + (size_t)(rows ? (int64_t)rows[i] : row0 + i) * ldx;
    TOUT *out = buf + (size_t)i * ncols;
    for (int c0 = w.lane * 2; c0 < ncols; c0 += 64 * VU) {
      double2 v[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) v[u] = (c0 + 64 * u < ncols) ? ld2(in + c0 + 64 * u) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < VU; ++u)
        if (c0 + 64 * u < ncols) {
          if constexpr (sizeof(TOUT) == 8)
            st2(reinterpret_cast<double *>(out) + c0 + 64 * u, v[u]);
          else
            *reinterpret_cast<float2 *>(out + c0 + 64 * u) = make_float2((float)v[u].x, (float)v[u].y);
        }
    }
  }
}

// Peer-memory push of the ghost exchange (comm.cu, "p2p" transport): row i of the outgoing payload (x[rows[i], :] or
// the contiguous row row0 + i) is stored STRAIGHT into the receive buffer of the rank that needs it - a peer-mapped
// pointer over NVLink (cudaIpc) - instead of a local send buffer + ncclSend/ncclRecv.  Rows are grouped in
// segments (one per destination rank, MPIPatternP2P order); segs.dst[s] is the address of segment s's first row in
// the peer's buffer.  16-byte stores, one warp per row: the NVLink writes are fully coalesced 256-byte+ bursts.
template <typename TOUT>
__global__ void __launch_bounds__(256)
push_rows_vec_kernel(const double *__restrict__ x, int ncols, int ldx, int64_t nRows,
                     const uint32_t *__restrict__ rows, int64_t row0, PushSegs segs) {
  const WarpRows w;
  for (int64_t i = w.first; i < nRows; i += w.step) {
    int sgm = 0;
    while (sgm + 1 < segs.n && i >= segs.start[sgm + 1]) ++sgm;
    const double *in = x + (size_t)(rows ? (int64_t)rows[i] : row0 + i) * ldx;
    TOUT *out = reinterpret_cast<TOUT *>(segs.dst[sgm]) + (size_t)(i - segs.start[sgm]) * ncols;
    for (int c0 = w.lane * 2; c0 < ncols; c0 += 64 * VU) {
      double2 v[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) v[u] = (c0 + 64 * u < ncols) ? ld2(in + c0 + 64 * u) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < VU; ++u)
        if (c0 + 64 * u < ncols) {
          if constexpr (sizeof(TOUT) == 8)
            st2(reinterpret_cast<double *>(out) + c0 + 64 * u, v[u]);
          else
            *reinterpret_cast<float2 *>(out + c0 + 64 * u) = make_float2((float)v[u].x, (float)v[u].y);
        }
    }
  }
}

// scalar fallback (odd column counts / unaligned rows), FP64 payload
__global__ void push_rows_kernel(const double *__restrict__ x, int ncols, int ldx, int64_t nRows,
                                 const uint32_t *__restrict__ rows, int64_t row0, PushSegs segs) {
  const int64_t total = nRows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / ncols;
    const int c = idx % ncols;
    int sgm = 0;
    while (sgm + 1 < segs.n && i >= segs.start[sgm + 1]) ++sgm;
    reinterpret_cast<double *>(segs.dst[sgm])[(size_t)(i - segs.start[sgm]) * ncols + c] =
        x[(size_t)(rows ? (int64_t)rows[i] : row0 + i) * ldx + c];
  }
}

// flag[t] = value on up to 32 (peer-mapped) addresses, release semantics at system scope: everything the stream
// wrote before this kernel (the pushed rows) is visible to the peer once it sees the flag
__global__ void signal_flags_kernel(SignalList l, uint32_t value) {
  const int t = threadIdx.x;
  if (t < l.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(l.addr[t]), "r"(value) : "memory");
  }
}

// contiguous rows [row0, row0+nRows) of x <- dense buffer (ghost segment fill)
template <typename TIN>
__global__ void __launch_bounds__(256)
unpack_rows_vec_kernel(double *__restrict__ x, int ncols, int ldx, int64_t row0, int64_t nRows,
                       const TIN *__restrict__ buf) {
  const WarpRows w;
  for (int64_t i = w.first; i < nRows; i += w.step) {
    double *out = x + (size_t)(row0 + i) * ldx;
    const TIN *in = buf + (size_t)i * ncols;
    for (int c = w.lane * 2; c < ncols; c += 64) {
      if constexpr (sizeof(TIN) == 8) {
        st2(out + c, ld2(reinterpret_cast<const double *>(in) + c));
      } else {
        const float2 f = *reinterpret_cast<const float2 *>(in + c);
        st2(out + c, make_double2((double)f.x, (double)f.y));
      }
    }
  }
}

// FP64 payload: x[r,:] += scale[r] * slot_k, slots in ascending order.
// FP32 payload (reference: the whole accumulate runs on a float copy of dst and the processor-boundary rows
// are copied back as doubles, kohnShamDFTOperatorDevice.cc:3953-3990):
//   x[r,:] = double( float(x[r,:]) +f float(scale[r]*slot_1) +f float(scale[r]*slot_2) ... )   (FP32 adds)
template <typename TIN>
__global__ void __launch_bounds__(256, 5)  // <= 48 registers: 8 warps within what a resident cell CTA leaves of the SM's registers
unpack_add_vec_kernel(double *__restrict__ x, int ncols, int ldx, int64_t nRows, const uint32_t *__restrict__ rows,
                      const uint32_t *__restrict__ starts, const uint32_t *__restrict__ slots,
                      const TIN *__restrict__ buf, const double *__restrict__ rowScale) {
  const WarpRows w;
  for (int64_t i = w.first; i < nRows; i += w.step) {
    const uint32_t r = rows[i];
    const double sc = rowScale ? rowScale[r] : 1.0;
    const uint32_t j0 = starts[i], j1 = starts[i + 1];
    double *row = x + (size_t)r * ldx;
    for (int c0 = w.lane * 2; c0 < ncols; c0 += 64 * VU) {
      double2 v[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) v[u] = (c0 + 64 * u < ncols) ? ld2(row + c0 + 64 * u) : make_double2(0.0, 0.0);
      if constexpr (sizeof(TIN) == 8) {
        for (uint32_t j = j0; j < j1; ++j) {
          const double *in = reinterpret_cast<const double *>(buf) + (size_t)slots[j] * ncols + c0;
#pragma unroll
          for (int u = 0; u < VU; ++u)
            if (c0 + 64 * u < ncols) {
              double2 t = ld2(in + 64 * u);
              if (rowScale) {
                t.x *= sc;
                t.y *= sc;
              }
              v[u].x += t.x;
              v[u].y += t.y;
            }
        }
      } else {
        float2 f[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) f[u] = make_float2((float)v[u].x, (float)v[u].y);
        for (uint32_t j = j0; j < j1; ++j) {
          const float *in = reinterpret_cast<const float *>(buf) + (size_t)slots[j] * ncols + c0;
#pragma unroll
          for (int u = 0; u < VU; ++u)
            if (c0 + 64 * u < ncols) {
              const float2 t = *reinterpret_cast<const float2 *>(in + 64 * u);
              f[u].x = __fadd_rn(f[u].x, rowScale ? (float)((double)t.x * sc) : t.x);
              f[u].y = __fadd_rn(f[u].y, rowScale ? (float)((double)t.y * sc) : t.y);
            }
        }
#pragma unroll
        for (int u = 0; u < VU; ++u) v[u] = make_double2((double)f[u].x, (double)f[u].y);
      }
#pragma unroll
      for (int u = 0; u < VU; ++u)
        if (c0 + 64 * u < ncols) st2(row + c0 + 64 * u, v[u]);
    }
  }
}

__global__ void __launch_bounds__(256)
row_scale_vec_kernel(double *__restrict__ x, int64_t rows, int ncols, int ldx, double alpha,
                     const double *__restrict__ s) {
  const WarpRows w;
  for (int64_t r = w.first; r < rows; r += w.step) {
    const double f = s ? alpha * s[r] : alpha;
    double *row = x + (size_t)r * ldx;
    for (int c0 = w.lane * 2; c0 < ncols; c0 += 64 * VU) {
      double2 v[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) v[u] = (c0 + 64 * u < ncols) ? ld2(row + c0 + 64 * u) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < VU; ++u)
        if (c0 + 64 * u < ncols) st2(row + c0 + 64 * u, make_double2(v[u].x * f, v[u].y * f));
    }
  }
}

// dst[r, 0:ncols] = s[r] * src[r, 0:ncols] with independent leading dimensions: the block slice out of /
// back into the full wavefunction matrix (K8).  Both sides are touched once -> streaming loads / stores.
__global__ void __launch_bounds__(256)
strided_copy_vec_kernel(const double *__restrict__ src, int lds, double *__restrict__ dst, int ldd, int ncols,
                        int64_t rows, const double *__restrict__ s) {
  const WarpRows w;
  for (int64_t r = w.first; r < rows; r += w.step) {
    const double f = s ? s[r] : 1.0;
    const double *in = src + (size_t)r * lds;
    double *out = dst + (size_t)r * ldd;
    for (int c0 = w.lane * 2; c0 < ncols; c0 += 64 * VU) {
      double2 v[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) v[u] = (c0 + 64 * u < ncols) ? ld2_stream(in + c0 + 64 * u) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < VU; ++u)
        if (c0 + 64 * u < ncols) {
          if (s) {
            v[u].x *= f;
            v[u].y *= f;
          }
          st2_stream(out + c0 + 64 * u, v[u]);
        }
    }
  }
}

// Streaming row mover, software pipelined: dst[r, :] = f(r) * src[map(r), :] for r < rows, one warp per row.
// The next row's NCH 16-byte loads per lane are issued BEFORE the current row is stored, so a warp always has
// a full row (ncols * 8 bytes) of loads in flight; a pure load-then-store loop leaves the memory system idle
// during each warp's store phase and measured 3.4 TB/s, about half of the copy bandwidth.
// Covers the block slices (K8), stridedBlockScale (K5, src == dst), the ghost pack (K14, row map) and the
// ghost-segment fill.  NCH = ceil(ncols / 64) <= 8.
template <int NCH>
__global__ void __launch_bounds__(256)
stream_rows_kernel(const double *src, int64_t lds, const uint32_t *__restrict__ srcRows, double *dst, int64_t ldd,
                   int ncols, int64_t rows, const double *__restrict__ s, double alpha, int streaming) {
  const WarpRows w;
  const int c = w.lane * 2;
  int64_t r = w.first;
  if (r >= rows) return;
  double2 v[NCH];
  double f;
  auto load = [&](int64_t row, double2(&out)[NCH], double &fac) {
    const double *in = src + (srcRows ? (int64_t)srcRows[row] : row) * lds + c;
#pragma unroll
    for (int u = 0; u < NCH; ++u)
      out[u] = (c + 64 * u < ncols) ? (streaming ? ld2_stream(in + 64 * u) : ld2(in + 64 * u)) : make_double2(0.0, 0.0);
    fac = s ? alpha * s[row] : alpha;
  };
  load(r, v, f);
  while (true) {
    const int64_t rn = r + w.step;
    double2 vn[NCH];
    double fn = 1.0;
    if (rn < rows) load(rn, vn, fn);
    double *out = dst + r * ldd + c;
#pragma unroll
    for (int u = 0; u < NCH; ++u)
      if (c + 64 * u < ncols) {
        const double2 o = make_double2(v[u].x * f, v[u].y * f);
        if (streaming)
          st2_stream(out + 64 * u, o);
        else
          st2(out + 64 * u, o);
      }
    if (rn >= rows) break;
#pragma unroll
    for (int u = 0; u < NCH; ++u) v[u] = vn[u];
    f = fn;
    r = rn;
  }
}

inline bool vec_ok(const void *p, int ncols, int ldx) {
  return (ncols % 2 == 0) && (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

// one warp per row, 8 warps per CTA; the grid is exactly one resident wave (SM count x the CTAs of THIS kernel
// that fit on an SM), so that no CTA queues behind a full machine and leaves a half-empty second wave
template <typename K>
inline int grid_rows(dftfe_b200_ctx *ctx, K kernel, int64_t nRows) {
  int &perSm = ctx->rowKernelCtasPerSm[reinterpret_cast<const void *>(kernel)];
  if (perSm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, 256, 0) != cudaSuccess || perSm < 1) perSm = 4;
  }
  int64_t g = (nRows + 7) / 8;
  const int64_t cap = (int64_t)ctx->num_sms * perSm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// dst[r, 0:ncols] = alpha * s[r] * src[map(r), 0:ncols]; returns false when the shape needs the fallback kernels
inline bool launch_stream_rows(dftfe_b200_ctx *ctx, const double *src, int64_t lds, const uint32_t *srcRows,
                               double *dst, int64_t ldd, int ncols, int64_t rows, const double *s, double alpha,
                               bool streaming) {
  if (ctx->force_scalar_row_kernels || ncols > 512 || ncols % 2 || lds % 2 || ldd % 2 ||
      (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    return false;
  const int nch = (ncols + 63) / 64;
  const int st = streaming ? 1 : 0;
#define DB_STREAM(N)                                                                                      \
  stream_rows_kernel<N><<<grid_rows(ctx, stream_rows_kernel<N>, rows), 256, 0, ctx->stream>>>(src, lds, srcRows, dst, \
                                                                                             ldd, ncols, rows, s,   \
                                                                                             alpha, st)
  if (nch <= 1)
    DB_STREAM(1);
  else if (nch <= 2)
    DB_STREAM(2);
  else if (nch <= 4)
    DB_STREAM(4);
  else
    DB_STREAM(8);
#undef DB_STREAM
  return true;
}

}  // namespace

int launch_distribute(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *colScale) {
  if (ctx->nCon == 0) return 0;
  ProfScope ps(ctx, "distribute");
  if (vec_ok(x, ncols, ldx) && !ctx->force_scalar_row_kernels)
    distribute_vec_kernel<<<grid_rows(ctx, distribute_vec_kernel, ctx->nCon), 256, 0, ctx->stream>>>(
        x, ncols, ldx, ctx->nCon, ctx->conRows.p, ctx->conSizes.p, ctx->conStarts.p, ctx->conCols.p,
        ctx->conVals.p, ctx->conInhom.p, colScale);
  else
    distribute_kernel<<<grid_for(ctx, ctx->nCon * ncols), 256, 0, ctx->stream>>>(
        x, ncols, ldx, ctx->nCon, ctx->conRows.p, ctx->conSizes.p, ctx->conStarts.p, ctx->conCols.p,
        ctx->conVals.p, ctx->conInhom.p, colScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_slave_to_master(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *masterScale) {
  if (ctx->nCon == 0) return 0;
  ProfScope ps(ctx, "slave_to_master", 2);
  const bool vec = vec_ok(x, ncols, ldx) && !ctx->force_scalar_row_kernels;
  if (ctx->nMasters > 0) {
    if (vec)
      slave_to_master_vec_kernel<<<grid_rows(ctx, slave_to_master_vec_kernel, ctx->nMasters), 256, 0, ctx->stream>>>(
          x, ncols, ldx, ctx->nMasters, ctx->masterRows.p, ctx->masterStarts.p, ctx->masterSlaves.p,
          ctx->masterVals.p, masterScale);
    else
      slave_to_master_kernel<<<grid_for(ctx, ctx->nMasters * ncols), 256, 0, ctx->stream>>>(
          x, ncols, ldx, ctx->nMasters, ctx->masterRows.p, ctx->masterStarts.p, ctx->masterSlaves.p,
          ctx->masterVals.p, masterScale);
  }
  if (vec)
    zero_rows_vec_kernel<<<grid_rows(ctx, zero_rows_vec_kernel, ctx->nCon), 256, 0, ctx->stream>>>(x, ncols, ldx, ctx->nCon, ctx->conRows.p);
  else
    zero_rows_kernel<<<grid_for(ctx, ctx->nCon * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, ctx->nCon,
                                                                               ctx->conRows.p);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_set_zero_rows(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx) {
  if (ctx->nCon == 0) return 0;
  ProfScope ps(ctx, "set_zero");
  if (vec_ok(x, ncols, ldx) && !ctx->force_scalar_row_kernels)
    zero_rows_vec_kernel<<<grid_rows(ctx, zero_rows_vec_kernel, ctx->nCon), 256, 0, ctx->stream>>>(x, ncols, ldx, ctx->nCon, ctx->conRows.p);
  else
    zero_rows_kernel<<<grid_for(ctx, ctx->nCon * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, ctx->nCon,
                                                                               ctx->conRows.p);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_row_scale(dftfe_b200_ctx *ctx, double *x, int64_t rows, int ncols, int ldx, double alpha,
                     const double *rowScale) {
  if (rows == 0) return 0;
  ProfScope ps(ctx, "row_scale");
  if (launch_stream_rows(ctx, x, ldx, nullptr, x, ldx, ncols, rows, rowScale, alpha, false)) {
  } else if (vec_ok(x, ncols, ldx) && !ctx->force_scalar_row_kernels)
    row_scale_vec_kernel<<<grid_rows(ctx, row_scale_vec_kernel, rows), 256, 0, ctx->stream>>>(x, rows, ncols, ldx, alpha, rowScale);
  else
    row_scale_kernel<<<grid_for(ctx, rows * ncols), 256, 0, ctx->stream>>>(x, rows, ncols, ldx, alpha, rowScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_block_copy_from_full(dftfe_b200_ctx *ctx, const double *X, int N, int j0, double *blk, int ncols,
                                int64_t rows, const double *rowScale, int ldBlk) {
  if (rows == 0) return 0;
  if (ldBlk <= 0) ldBlk = ncols;
  ProfScope ps(ctx, "block_copy");
  if (launch_stream_rows(ctx, X + j0, N, nullptr, blk, ldBlk, ncols, rows, rowScale, 1.0, true)) {
  } else if (ldBlk == ncols) {
    block_from_full_kernel<<<grid_for(ctx, rows * ncols), 256, 0, ctx->stream>>>(X, N, j0, blk, ncols, rows,
                                                                                rowScale);
  } else {  // odd shapes with a padded block: row by row through the generic 2-D copy
    DB_CHECK(rowScale == nullptr, "block copy: a row scale needs an even column count when the block is padded");
    DB_CUDA(cudaMemcpy2DAsync(blk, (size_t)ldBlk * sizeof(double), X + j0, (size_t)N * sizeof(double),
                              (size_t)ncols * sizeof(double), (size_t)rows, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_block_copy_to_full(dftfe_b200_ctx *ctx, double *X, int N, int j0, const double *blk, int ncols,
                              int64_t rows, const double *rowScale) {
  if (rows == 0) return 0;
  ProfScope ps(ctx, "block_copy");
  if (launch_stream_rows(ctx, blk, ncols, nullptr, X + j0, N, ncols, rows, rowScale, 1.0, true)) {
  } else if (vec_ok(X + j0, ncols, N) && vec_ok(blk, ncols, ncols) && !ctx->force_scalar_row_kernels)
    strided_copy_vec_kernel<<<grid_rows(ctx, strided_copy_vec_kernel, rows), 256, 0, ctx->stream>>>(blk, ncols, X + j0, N, ncols, rows,
                                                                          rowScale);
  else
    block_to_full_kernel<<<grid_for(ctx, rows * ncols), 256, 0, ctx->stream>>>(X, N, j0, blk, ncols, rows, rowScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// ---- kernels used by comm.cu ------------------------------------------------
// rows == nullptr: contiguous rows [row0, row0 + nRows)
int launch_pack_rows(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, int64_t nRows, const uint32_t *rows,
                     int64_t row0, double *buf) {
  if (nRows == 0) return 0;
  ProfScope ps(ctx, "ghost_pack");
  if (launch_stream_rows(ctx, rows ? x : x + (size_t)row0 * ldx, ldx, rows, buf, ncols, ncols, nRows, nullptr, 1.0,
                         false)) {
  } else if (vec_ok(x, ncols, ldx) && vec_ok(buf, ncols, ncols) && !ctx->force_scalar_row_kernels)
    pack_rows_vec_kernel<double><<<grid_rows(ctx, pack_rows_vec_kernel<double>, nRows), 256, 0, ctx->stream>>>(x, ncols, ldx, nRows, rows, row0, buf);
  else if (rows)
    pack_rows_kernel<<<grid_for(ctx, nRows * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, nRows, rows, buf);
  else
    copy_rows_kernel<<<grid_for(ctx, nRows * ncols), 256, 0, ctx->stream>>>(const_cast<double *>(x), ncols, ldx, row0,
                                                                           nRows, buf, 1);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// FP32 payload of the mixed-precision filter (HXCheby with chebMixedPrec, kohnShamDFTOperatorDevice.cc:3899-3915)
int launch_pack_rows_f32(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, int64_t nRows,
                         const uint32_t *rows, int64_t row0, float *buf) {
  if (nRows == 0) return 0;
  DB_CHECK(vec_ok(x, ncols, ldx), "FP32 ghost payload needs an even column count and 16-byte aligned rows");
  ProfScope ps(ctx, "ghost_pack");
  pack_rows_vec_kernel<float><<<grid_rows(ctx, pack_rows_vec_kernel<float>, nRows), 256, 0, ctx->stream>>>(x, ncols, ldx, nRows, rows, row0, buf);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// p2p transport: push rows into the peers' receive buffers (see push_rows_vec_kernel)
int launch_push_rows(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, int64_t nRows, const uint32_t *rows,
                     int64_t row0, const PushSegs &segs, bool fp32) {
  if (nRows == 0) return 0;
  ProfScope ps(ctx, "ghost_pack");
  bool aligned = vec_ok(x, ncols, ldx) && !ctx->force_scalar_row_kernels;
  for (int s = 0; s < segs.n; ++s) aligned = aligned && ((reinterpret_cast<uintptr_t>(segs.dst[s]) & 15) == 0);
  if (fp32) {
    DB_CHECK(vec_ok(x, ncols, ldx), "FP32 ghost payload needs an even column count and 16-byte aligned rows");
    push_rows_vec_kernel<float><<<grid_rows(ctx, push_rows_vec_kernel<float>, nRows), 256, 0, ctx->stream>>>(
        x, ncols, ldx, nRows, rows, row0, segs);
  } else if (aligned) {
    push_rows_vec_kernel<double><<<grid_rows(ctx, push_rows_vec_kernel<double>, nRows), 256, 0, ctx->stream>>>(
        x, ncols, ldx, nRows, rows, row0, segs);
  } else {
    push_rows_kernel<<<grid_for(ctx, nRows * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, nRows, rows, row0, segs);
  }
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_signal_flags(dftfe_b200_ctx *ctx, const SignalList &l, uint32_t value) {
  if (l.n == 0) return 0;
  ctx->launches += 1;
  signal_flags_kernel<<<1, 32, 0, ctx->stream>>>(l, value);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_unpack_rows(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, int64_t row0, int64_t nRows,
                       const double *buf) {
  if (nRows == 0) return 0;
  ProfScope ps(ctx, "ghost_unpack");
  if (launch_stream_rows(ctx, buf, ncols, nullptr, x + (size_t)row0 * ldx, ldx, ncols, nRows, nullptr, 1.0, false)) {
  } else if (vec_ok(x, ncols, ldx) && vec_ok(buf, ncols, ncols) && !ctx->force_scalar_row_kernels)
    unpack_rows_vec_kernel<double><<<grid_rows(ctx, unpack_rows_vec_kernel<double>, nRows), 256, 0, ctx->stream>>>(x, ncols, ldx, row0, nRows, buf);
  else
    copy_rows_kernel<<<grid_for(ctx, nRows * ncols), 256, 0, ctx->stream>>>(x, ncols, ldx, row0, nRows,
                                                                           const_cast<double *>(buf), 0);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_unpack_rows_f32(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, int64_t row0, int64_t nRows,
                           const float *buf) {
  if (nRows == 0) return 0;
  ProfScope ps(ctx, "ghost_unpack");
  unpack_rows_vec_kernel<float><<<grid_rows(ctx, unpack_rows_vec_kernel<float>, nRows), 256, 0, ctx->stream>>>(x, ncols, ldx, row0, nRows, buf);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_unpack_add(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *buf,
                      const double *rowScale) {
  if (ctx->nBoundaryRows == 0) return 0;
  ProfScope ps(ctx, "ghost_unpack");
  if (vec_ok(x, ncols, ldx) && vec_ok(buf, ncols, ncols) && !ctx->force_scalar_row_kernels)
    unpack_add_vec_kernel<double><<<grid_rows(ctx, unpack_add_vec_kernel<double>, ctx->nBoundaryRows), 256, 0, ctx->stream>>>(
        x, ncols, ldx, ctx->nBoundaryRows, ctx->bndRows.p, ctx->bndStarts.p, ctx->bndSlots.p, buf, rowScale);
  else
    unpack_add_kernel<<<grid_for(ctx, ctx->nBoundaryRows * ncols), 256, 0, ctx->stream>>>(
        x, ncols, ldx, ctx->nBoundaryRows, ctx->bndRows.p, ctx->bndStarts.p, ctx->bndSlots.p, buf, rowScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_unpack_add_f32(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const float *buf,
                          const double *rowScale) {
  if (ctx->nBoundaryRows == 0) return 0;
  ProfScope ps(ctx, "ghost_unpack");
  unpack_add_vec_kernel<float><<<grid_rows(ctx, unpack_add_vec_kernel<float>, ctx->nBoundaryRows), 256, 0, ctx->stream>>>(
      x, ncols, ldx, ctx->nBoundaryRows, ctx->bndRows.p, ctx->bndStarts.p, ctx->bndSlots.p, buf, rowScale);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dftfe_b200
