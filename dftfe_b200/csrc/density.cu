// Electron density from the wavefunctions: the step right after solve() in the SCF.
//
// Reference: computeRhoFromPSI (src/dft/densityCalculator.cc:39-560): per block of wavefunctions
//   stridedCopyToBlock -> updateGhostValues -> distribute -> interpolateKernel (gather + batched GEMM with the
//   shape-function values: psi_i(q) = sum_I N_I(q) x_i[row(c,I)]) -> computeRhoGradRhoFromInterpolatedValues
//   (psi^2, densityCalculatorDeviceKernels.cc:35-140) -> GEMV with the partial occupancies -> rho[c][q] +=.
// The interpolated values (nC x nq x B doubles per block) and their squares make two HBM round trips there.
//
// Here one kernel per block does gather + GEMM + square + occupancy-weighted column sum: a persistent CTA owns a
// cell and walks the block's 32-column tiles; the X tile is gathered through the index map by 1-D TMA bulk copies
// into a double-buffered shared-memory tile exactly as in the cell matvec kernel; twelve MMA warps hold the
// quadrature rows (m8 tiles of q, three per warp per pass) as FP64 DMMA accumulators, square them in registers,
// weight by f_i and add into per-thread partial sums that live across the tiles; one shuffle reduction and one
// store per (cell, q) at the end.  psi(q) never touches memory.  Complex build: columns are interleaved (re, im), so
// re^2 + im^2 falls out of the same column sum.
//
// grad rho (GGA functionals; computeRhoGradRhoFromInterpolatedValues with isEvaluateGradRho,
// densityCalculatorDeviceKernels.cc:35-140: gradRho[c][q][d] = sum_i f_i 2 Re(conj(psi_i) d_d psi_i)): a second
// kernel of the same structure that, per (cell, pass over 192 quadrature points, column tile), runs four k-loops -
// psi and the three reference-cell derivatives d_e psi = sum_I (d_e N_I)(q) x_I - keeps f*2*psi in registers,
// accumulates rho and the three reference-coordinate components, and applies the cell's inverse Jacobian once at the
// end (the map to physical derivatives is linear: grad_d = sum_e Jinv[c][d][e] d_e).  The interpolated gradients
// (3 x nC x nq x B doubles per block in the reference) never touch memory either.
#include "common.cuh"

namespace dftfe_b200 {

namespace {

constexpr int BT = 32, NT = BT / 8, LDS = BT + 4;
constexpr int WARPS = 12, TPW = 3, TPWP = 4;  // q tiles per warp per pass (padded to 4 for one 32-byte load)
constexpr int QT_PER_PASS = WARPS * TPW;      // 36 m8 tiles = 288 quadrature points per pass
constexpr int MAX_PASSES = 4;                 // nq <= 1152
constexpr int THREADS = (WARPS + 1) * 32;
constexpr int APF = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void load4(const double *p, double (&a)[TPWP]) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[0]), "=d"(a[1]), "=d"(a[2]), "=d"(a[3]) : "l"(p));
}

// shape values N[I][q] (n x nq) -> fragment-major A(q, k) = N_k(q):
//   Nf[pass][warp][ks][lane][t] = A[(pass*36 + warp + t*12)*8 + lane/4][ks*4 + lane%4], zero padded
__global__ void tile_shape_kernel(const double *__restrict__ N, int n, int nq, int KS, int passes,
                                  double *__restrict__ Nf) {
  const int64_t total = (int64_t)passes * WARPS * KS * 32 * TPWP;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int t = r % TPWP;
    r /= TPWP;
    const int lane = r % 32;
    r /= 32;
    const int ks = r % KS;
    r /= KS;
    const int w = r % WARPS;
    const int pass = (int)(r / WARPS);
    const int q = t < TPW ? (pass * QT_PER_PASS + w + t * WARPS) * 8 + lane / 4 : nq;
    const int k = ks * 4 + lane % 4;
    Nf[idx] = (q < nq && k < n) ? N[(size_t)k * nq + q] : 0.0;
  }
}

template <int NODES>
struct DCfg {
  static constexpr int KS = (NODES + 3) / 4;
  static constexpr int KPAD = KS * 4;
  static constexpr size_t XBUF = (size_t)KPAD * LDS;
  static constexpr size_t SMEM = 2 * XBUF * sizeof(double) + 4 * sizeof(uint64_t);
};

// rho[cell][q] += sum_{cols of this block} f[col / cm] * psi_col(q)^2
template <int NODES>
__global__ void __launch_bounds__(THREADS, 1)
density_kernel(const double *__restrict__ Nf, const uint32_t *__restrict__ cellRows, int64_t nCells,
               const double *__restrict__ x, int ldx, int nColTiles, const double *__restrict__ fcol, int nq,
               int passes, double *__restrict__ rho) {
  using D = DCfg<NODES>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *Xs = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + 2 * D::XBUF * sizeof(double));
  uint64_t *empty = full + 2;
  const int tid = threadIdx.x, lane = tid & 31, pwarp = tid >> 5, warp = pwarp - 1;
  constexpr uint32_t ROW_MASK = 0x3fffffffu;

  for (int i = tid; i < (int)(2 * D::XBUF); i += THREADS) Xs[i] = 0.0;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&empty[0], WARPS);
    mbar_init(&empty[1], WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (pwarp == 0) {
    // ===== producer: X tile of (cell, column tile) through the index map =====
    uint32_t it = 0;
    for (int64_t cell = blockIdx.x; cell < nCells; cell += gridDim.x) {
      uint32_t myRows[(NODES + 31) / 32];
#pragma unroll
      for (int j = 0; j < (NODES + 31) / 32; ++j) {
        const int k = lane + 32 * j;
        myRows[j] = (k < NODES) ? (__ldg(cellRows + (size_t)cell * NODES + k) & ROW_MASK) : 0u;
      }
      for (int tile = 0; tile < nColTiles; ++tile, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&empty[buf], ph ^ 1);
        double *xs = Xs + buf * D::XBUF;
        if (lane == 0) mbar_arrive_expect_tx(&full[buf], (uint32_t)(NODES * BT * sizeof(double)));
        __syncwarp();
#pragma unroll
        for (int j = 0; j < (NODES + 31) / 32; ++j) {
          const int k = lane + 32 * j;
          if (k < NODES)
            tma_bulk_g2s(xs + k * LDS, x + (size_t)myRows[j] * ldx + tile * BT, BT * sizeof(double), &full[buf]);
        }
      }
    }
  } else {
    // ===== MMA warps =====
    const double *xb0 = Xs + (lane & 3) * LDS + (lane >> 2);
    uint32_t it = 0;
    for (int64_t cell = blockIdx.x; cell < nCells; cell += gridDim.x) {
      double part[MAX_PASSES][TPW];
#pragma unroll
      for (int p = 0; p < MAX_PASSES; ++p)
#pragma unroll
        for (int t = 0; t < TPW; ++t) part[p][t] = 0.0;
      for (int tile = 0; tile < nColTiles; ++tile, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const double *xb = xb0 + buf * D::XBUF;
        // occupancy weights of the columns this lane owns in the accumulator fragment
        double fw[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) fw[nt][e] = __ldg(fcol + tile * BT + nt * 8 + (lane & 3) * 2 + e);
        mbar_wait(&full[buf], ph);
#pragma unroll
        for (int p = 0; p < MAX_PASSES; ++p) {
          if (p < passes) {
            const double *Ap = Nf + ((size_t)(p * WARPS + warp) * D::KS) * 32 * TPWP + lane * TPWP;
            double acc[TPW][NT][2];
#pragma unroll
            for (int t = 0; t < TPW; ++t)
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) acc[t][nt][0] = acc[t][nt][1] = 0.0;
            double a[APF][TPWP];
#pragma unroll
            for (int s = 0; s < APF; ++s) load4(Ap + (size_t)min(s, D::KS - 1) * 32 * TPWP, a[s]);
            int ks = 0;
            for (; ks + APF <= D::KS; ks += APF) {
#pragma unroll
              for (int s = 0; s < APF; ++s) {
                double b[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) b[nt] = xb[(ks + s) * 4 * LDS + nt * 8];
#pragma unroll
                for (int t = 0; t < TPW; ++t)
#pragma unroll
                  for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
                if (ks + s + APF < D::KS) load4(Ap + (size_t)(ks + s + APF) * 32 * TPWP, a[s]);
              }
            }
#pragma unroll
            for (int s = 0; s < D::KS % APF; ++s) {
              double b[NT];
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) b[nt] = xb[(ks + s) * 4 * LDS + nt * 8];
#pragma unroll
              for (int t = 0; t < TPW; ++t)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
            }
            // psi^2, weighted, summed over this lane's columns (fixed order)
#pragma unroll
            for (int t = 0; t < TPW; ++t) {
              double s = 0.0;
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                s += fw[nt][0] * acc[t][nt][0] * acc[t][nt][0];
                s += fw[nt][1] * acc[t][nt][1] * acc[t][nt][1];
              }
              part[p][t] += s;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[buf]);
      }
      // ---- one reduction over the four lanes that share a quadrature row, one read-modify-write per (cell, q)
#pragma unroll
      for (int p = 0; p < MAX_PASSES; ++p) {
        if (p < passes) {
#pragma unroll
          for (int t = 0; t < TPW; ++t) {
            double s = part[p][t];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            const int q = (p * QT_PER_PASS + warp + t * WARPS) * 8 + (lane >> 2);
            if ((lane & 3) == 0 && q < nq) rho[(size_t)cell * nq + q] += s;
          }
        }
      }
    }
  }
}

__global__ void expand_weights_kernel(const double *__restrict__ f, int j0, int ncolsValid, int cm, int ncolsPad,
                                      double *__restrict__ fcol) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncolsPad) fcol[c] = c < ncolsValid ? f[j0 + c / cm] : 0.0;
}

template <int NODES>
int launch_density(dftfe_b200_ctx *ctx, const double *x, int ldx, int nColTiles, int nq, int passes, double *rho) {
  using D = DCfg<NODES>;
  DB_DYN_SMEM(ctx, density_kernel<NODES>, D::SMEM);
  ProfScope ps(ctx, "density");
  const int grid = (int)std::min<int64_t>(ctx->nC, ctx->num_sms);
  density_kernel<NODES><<<grid, THREADS, D::SMEM, ctx->stream>>>(ctx->denNf.p, ctx->cellRowsFlagged.p, ctx->nC, x, ldx,
                                                                nColTiles, ctx->denF.p, nq, passes, rho);
  DB_CUDA(cudaGetLastError());
  return 0;
}


// ---------------------------------------------------------------------------
// rho AND grad rho
// ---------------------------------------------------------------------------
constexpr int GTPW = 2;                        // q tiles per warp per pass (two: psi and one derivative stay in registers)
constexpr int GQT_PER_PASS = WARPS * GTPW;     // 24 m8 tiles = 192 quadrature points per pass
constexpr int GMAX_PASSES = 6;                 // nq <= 1152
#ifndef DB_GAPF
#define DB_GAPF 4
#endif
constexpr int GAPF = DB_GAPF;                  // A-fragment prefetch depth of the gradient kernel

__device__ __forceinline__ void load2(const double *p, double (&a)[GTPW]) {
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a[0]), "=d"(a[1]) : "l"(p));
}

// values N[I][q] and reference-cell derivatives dN[e][I][q] -> fragment-major
//   Nf[comp][pass][warp][ks][lane][t] = A_comp[(pass*24 + warp + t*12)*8 + lane/4][ks*4 + lane%4],  comp 0 = N, 1..3 = d_e N
__global__ void tile_shape_grad_kernel(const double *__restrict__ N, const double *__restrict__ dN, int n, int nq,
                                       int KS, int passes, double *__restrict__ Nf) {
  const int64_t per = (int64_t)passes * WARPS * KS * 32 * GTPW;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < 4 * per;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int comp = (int)(idx / per);
    int64_t r = idx % per;
    const int t = r % GTPW;
    r /= GTPW;
    const int lane = r % 32;
    r /= 32;
    const int ks = r % KS;
    r /= KS;
    const int w = r % WARPS;
    const int pass = (int)(r / WARPS);
    const int q = (pass * GQT_PER_PASS + w + t * WARPS) * 8 + lane / 4;
    const int k = ks * 4 + lane % 4;
    double v = 0.0;
    if (q < nq && k < n) v = comp == 0 ? N[(size_t)k * nq + q] : dN[((size_t)(comp - 1) * n + k) * nq + q];
    Nf[idx] = v;
  }
}

template <int NODES>
__global__ void __launch_bounds__(THREADS, 1)
density_grad_kernel(const double *__restrict__ Nf, const uint32_t *__restrict__ cellRows, int64_t nCells,
                    const double *__restrict__ x, int ldx, int nColTiles, const double *__restrict__ fcol, int nq,
                    int passes, const double *__restrict__ invJac, double *__restrict__ rho,
                    double *__restrict__ gradRho) {
  using D = DCfg<NODES>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *Xs = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + 2 * D::XBUF * sizeof(double));
  uint64_t *empty = full + 2;
  const int tid = threadIdx.x, lane = tid & 31, pwarp = tid >> 5, warp = pwarp - 1;
  constexpr uint32_t ROW_MASK = 0x3fffffffu;
  const size_t perComp = (size_t)passes * WARPS * D::KS * 32 * GTPW;

  for (int i = tid; i < (int)(2 * D::XBUF); i += THREADS) Xs[i] = 0.0;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&empty[0], WARPS);
    mbar_init(&empty[1], WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (pwarp == 0) {
    // ===== producer: the block's column tiles are walked once per pass =====
    uint32_t it = 0;
    for (int64_t cell = blockIdx.x; cell < nCells; cell += gridDim.x) {
      uint32_t myRows[(NODES + 31) / 32];
#pragma unroll
      for (int j = 0; j < (NODES + 31) / 32; ++j) {
        const int k = lane + 32 * j;
        myRows[j] = (k < NODES) ? (__ldg(cellRows + (size_t)cell * NODES + k) & ROW_MASK) : 0u;
      }
      for (int p = 0; p < passes; ++p)
        for (int tile = 0; tile < nColTiles; ++tile, ++it) {
          const int buf = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(&empty[buf], ph ^ 1);
          double *xs = Xs + buf * D::XBUF;
          if (lane == 0) mbar_arrive_expect_tx(&full[buf], (uint32_t)(NODES * BT * sizeof(double)));
          __syncwarp();
#pragma unroll
          for (int j = 0; j < (NODES + 31) / 32; ++j) {
            const int k = lane + 32 * j;
            if (k < NODES)
              tma_bulk_g2s(xs + k * LDS, x + (size_t)myRows[j] * ldx + tile * BT, BT * sizeof(double), &full[buf]);
          }
        }
    }
  } else {
    const double *xb0 = Xs + (lane & 3) * LDS + (lane >> 2);
    uint32_t it = 0;
    for (int64_t cell = blockIdx.x; cell < nCells; cell += gridDim.x) {
      for (int p = 0; p < passes; ++p) {
        double part[GTPW], gpart[3][GTPW];
#pragma unroll
        for (int t = 0; t < GTPW; ++t) part[t] = gpart[0][t] = gpart[1][t] = gpart[2][t] = 0.0;
        for (int tile = 0; tile < nColTiles; ++tile, ++it) {
          const int buf = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          const double *xb = xb0 + buf * D::XBUF;
          mbar_wait(&full[buf], ph);
          double w2[GTPW][NT][2];  // f * 2 * psi of this lane's accumulator elements
#pragma unroll
          for (int comp = 0; comp < 4; ++comp) {
            const double *Ap = Nf + comp * perComp + ((size_t)(p * WARPS + warp) * D::KS) * 32 * GTPW + lane * GTPW;
            double acc[GTPW][NT][2];
#pragma unroll
            for (int t = 0; t < GTPW; ++t)
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) acc[t][nt][0] = acc[t][nt][1] = 0.0;
            double a[GAPF][GTPW];
#pragma unroll
            for (int s = 0; s < GAPF; ++s) load2(Ap + (size_t)min(s, D::KS - 1) * 32 * GTPW, a[s]);
            int ks = 0;
            for (; ks + GAPF <= D::KS; ks += GAPF) {
#pragma unroll
              for (int s = 0; s < GAPF; ++s) {
                double b[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) b[nt] = xb[(ks + s) * 4 * LDS + nt * 8];
#pragma unroll
                for (int t = 0; t < GTPW; ++t)
#pragma unroll
                  for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
                if (ks + s + GAPF < D::KS) load2(Ap + (size_t)(ks + s + GAPF) * 32 * GTPW, a[s]);
              }
            }
#pragma unroll
            for (int s = 0; s < D::KS % GAPF; ++s) {
              double b[NT];
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) b[nt] = xb[(ks + s) * 4 * LDS + nt * 8];
#pragma unroll
              for (int t = 0; t < GTPW; ++t)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
            }
            if (comp == 0) {
              // rho contribution; keep f * 2 * psi for the three derivative products
#pragma unroll
              for (int t = 0; t < GTPW; ++t) {
                double s = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    const double f = __ldg(fcol + tile * BT + nt * 8 + (lane & 3) * 2 + e);
                    const double fp = f * acc[t][nt][e];
                    s += fp * acc[t][nt][e];
                    w2[t][nt][e] = 2.0 * fp;
                  }
                part[t] += s;
              }
            } else {
#pragma unroll
              for (int t = 0; t < GTPW; ++t) {
                double s = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) s += w2[t][nt][0] * acc[t][nt][0] + w2[t][nt][1] * acc[t][nt][1];
                gpart[comp - 1][t] += s;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[buf]);
        }
        // ---- reduce over the four lanes of a quadrature row; reference -> physical derivatives; one RMW per (cell, q)
        double J[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) J[e] = invJac ? __ldg(invJac + (size_t)cell * 9 + e) : ((e % 4 == 0) ? 1.0 : 0.0);
#pragma unroll
        for (int t = 0; t < GTPW; ++t) {
          double s = part[t], g0 = gpart[0][t], g1 = gpart[1][t], g2 = gpart[2][t];
#pragma unroll
          for (int m = 1; m <= 2; m <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, m);
            g0 += __shfl_xor_sync(0xffffffffu, g0, m);
            g1 += __shfl_xor_sync(0xffffffffu, g1, m);
            g2 += __shfl_xor_sync(0xffffffffu, g2, m);
          }
          const int q = (p * GQT_PER_PASS + warp + t * WARPS) * 8 + (lane >> 2);
          if ((lane & 3) == 0 && q < nq) {
            rho[(size_t)cell * nq + q] += s;
            double *g = gradRho + ((size_t)cell * nq + q) * 3;
            g[0] += J[0] * g0 + J[1] * g1 + J[2] * g2;
            g[1] += J[3] * g0 + J[4] * g1 + J[5] * g2;
            g[2] += J[6] * g0 + J[7] * g1 + J[8] * g2;
          }
        }
      }
    }
  }
}

template <int NODES>
int launch_density_grad(dftfe_b200_ctx *ctx, const double *x, int ldx, int nColTiles, int nq, int passes,
                        const double *invJac, double *rho, double *gradRho) {
  using D = DCfg<NODES>;
  DB_DYN_SMEM(ctx, density_grad_kernel<NODES>, D::SMEM);
  ProfScope ps(ctx, "density");
  const int grid = (int)std::min<int64_t>(ctx->nC, ctx->num_sms);
  density_grad_kernel<NODES><<<grid, THREADS, D::SMEM, ctx->stream>>>(ctx->denNf.p, ctx->cellRowsFlagged.p, ctx->nC, x,
                                                                     ldx, nColTiles, ctx->denF.p, nq, passes, invJac, rho,
                                                                     gradRho);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// rho_out[c][q] = sum_i f_i |psi_i(x_q)|^2 over the N columns of X (row-major M x N, FE basis), cells owned by this rank
// shapeGradValues [3][n][nq] (reference-cell derivatives), invJac [nC][3][3] (Jinv[c][d][e] = d xi_e / d x_d, the
// reference's inverseJacobianValues layout; nullptr: identity) and gradRho [nC][nq][3] non-null: also grad rho (GGA)
int compute_density(dftfe_b200_ctx *ctx, const double *X, int N, const double *occ_h, int nq, const double *shapeValues,
                    double *rho, const double *shapeGradValues, const double *invJac, double *gradRho) {
  DB_CHECK(ctx->have_map, "compute_density: set_index_map first");
  const bool grad = gradRho != nullptr;
  DB_CHECK(!grad || shapeGradValues, "compute_density: grad rho needs the shape-function derivatives");
  const int maxq = grad ? GMAX_PASSES * GQT_PER_PASS * 8 : MAX_PASSES * QT_PER_PASS * 8;
  DB_CHECK(nq >= 1 && nq <= maxq, "compute_density: n_quad (%d) must be in [1, %d]", nq, maxq);
  DB_CHECK(ctx->n != 512, "compute_density: FE order 7 does not fit the double-buffered tile");
  const int cm = ctx->cm, n = ctx->n;
  const int KS = (n + 3) / 4;
  const int qtPerPass = grad ? GQT_PER_PASS : QT_PER_PASS;
  const int passes = ((nq + 7) / 8 + qtPerPass - 1) / qtPerPass;
  const int B = std::min(ctx->B, N);
  const int Bpad = ((B * cm + BT - 1) / BT) * BT;  // real columns per block, padded to full 32-column tiles
  DB_TRY(ctx->denNf.alloc(grad ? (size_t)4 * passes * WARPS * KS * 32 * GTPW : (size_t)passes * WARPS * KS * 32 * TPWP));
  DB_TRY(ctx->denOcc.upload(occ_h, N, ctx->stream));
  DB_TRY(ctx->denF.alloc(Bpad));
  DB_TRY(ctx->denBlock.alloc((size_t)(ctx->M + ctx->G) * Bpad));
  ctx->launches += 2;
  if (grad) {
    tile_shape_grad_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(shapeValues, shapeGradValues, n, nq, KS, passes,
                                                                     ctx->denNf.p);
    DB_CUDA(cudaMemsetAsync(gradRho, 0, (size_t)ctx->nC * nq * 3 * sizeof(double), ctx->stream));
  } else {
    tile_shape_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(shapeValues, n, nq, KS, passes, ctx->denNf.p);
  }
  DB_CUDA(cudaMemsetAsync(rho, 0, (size_t)ctx->nC * nq * sizeof(double), ctx->stream));
  const bool padded = Bpad != B * cm;
  if (padded)
    DB_CUDA(cudaMemsetAsync(ctx->denBlock.p, 0, (size_t)(ctx->M + ctx->G) * Bpad * sizeof(double), ctx->stream));
  for (int j = 0; j < N; j += B) {
    const int Bc = std::min(B, N - j), Bcr = Bc * cm;
    // block slice (leading dimension Bpad), ghost values, constraints - as computeRhoFromPSI does per block
    DB_TRY(launch_block_copy_from_full(ctx, X + (size_t)j * cm, N * cm, 0, ctx->denBlock.p, Bcr, ctx->M, nullptr, Bpad));
    DB_TRY(ghost_update(ctx, ctx->denBlock.p, Bcr, Bpad));
    DB_TRY(launch_distribute(ctx, ctx->denBlock.p, Bcr, Bpad, nullptr));
    ctx->launches += 1;
    expand_weights_kernel<<<(Bpad + 127) / 128, 128, 0, ctx->stream>>>(ctx->denOcc.p, j, Bcr, cm, Bpad, ctx->denF.p);
    const int nColTiles = (Bcr + BT - 1) / BT;
    if (grad) {
#define DB_DG(NN) case NN: DB_TRY(launch_density_grad<NN>(ctx, ctx->denBlock.p, Bpad, nColTiles, nq, passes, invJac, rho, gradRho)); break;
      switch (n) {
        DB_DG(8) DB_DG(27) DB_DG(64) DB_DG(125) DB_DG(216) DB_DG(343)
        default:
          set_error("compute_density: no kernel for %d nodes per cell", n);
          return DFTFE_B200_ERR_UNSUPPORTED;
      }
#undef DB_DG
      continue;
    }
    switch (n) {
      case 8: DB_TRY(launch_density<8>(ctx, ctx->denBlock.p, Bpad, nColTiles, nq, passes, rho)); break;
      case 27: DB_TRY(launch_density<27>(ctx, ctx->denBlock.p, Bpad, nColTiles, nq, passes, rho)); break;
      case 64: DB_TRY(launch_density<64>(ctx, ctx->denBlock.p, Bpad, nColTiles, nq, passes, rho)); break;
      case 125: DB_TRY(launch_density<125>(ctx, ctx->denBlock.p, Bpad, nColTiles, nq, passes, rho)); break;
      case 216: DB_TRY(launch_density<216>(ctx, ctx->denBlock.p, Bpad, nColTiles, nq, passes, rho)); break;
      case 343: DB_TRY(launch_density<343>(ctx, ctx->denBlock.p, Bpad, nColTiles, nq, passes, rho)); break;
      default:
        set_error("compute_density: no kernel for %d nodes per cell", n);
        return DFTFE_B200_ERR_UNSUPPORTED;
    }
  }
  return 0;
}

}  // namespace dftfe_b200
