// Cell-level Hamiltonian assembly: what runs every SCF right before the ChFSI path.
//
// Reference: hamMatrixKernelLDA (src/dftOperator/hamiltonianMatrixCalculatorFlattenedDevice.cc:63-117), called from
// kohnShamDFTOperatorDeviceClass::computeHamiltonianMatricesAllkpt - one thread per matrix entry (cell, I, J)
// looping over the quadrature points:
//     H_c(I,J) = 1/2 K_c(I,J) + sum_q vEffJxW[c,q] N_I(q) N_J(q)  (+ the external-potential correction matrix)
// i.e. per cell the GEMM  N diag(w_c) N^T  with N (n x nq) shared by every cell, computed there with scalar FMAs
// and three global/L1 loads per FMA.
//
// Here it is a batched FP64 DMMA GEMM: item = (cell, lower 128x128 output tile); eight MMA warps (2 x 4, 64 x 32 each)
// fed from a shared-memory ring that a producer warp fills with 1-D TMA bulk copies of N^T rows (k = quadrature
// point) and of the cell's weights; the weight multiplies the B fragment (one DMUL per four DMMAs); H_c is
// symmetric, so only the 6 lower tiles of the 3 x 3 tile grid (n = 343) are computed and the strictly-lower ones are
// mirrored on the way out.  Output: the reference's own layout mem[c*n*n + I*n + J], ready for
// dftfe_b200_set_cell_hamiltonian.
//
// GGA and k-point terms (hamMatrixKernelGGAMemOpt / the complex kernels, :119-657) are contractions of the same kind
// with the reference-cell derivative tables on one side:
//   GGA      H += sum_q 2 g_d(c,q) (d_d N_I N_J + N_I d_d N_J) = G + G^T,  G(I,J) = sum_q N_I(q) [1/2 w N_J + sum_e u_e d_e N_J](q)
//   k-points Re H_k = H + 1/2 |k|^2 Mc,  Mc = N diag(JxW) N^T;   Im H_k(I,J) = - sum_d k_d sum_q d_d N_I N_J JxW = sum_e kappa_e(c) E_e(I,J)
// with u_e = 2 sum_d g_d Jinv[c][d][e], kappa_e = sum_d k_d Jinv[c][d][e] (affine cells) and
// E_e(I,J) = - sum_q JxW d_e N_I N_J.  `assemble_general_kernel` is the same TMA-fed DMMA pipeline with a B operand
// that is a per-quadrature-point linear combination of up to four tables (values + three derivatives), all nine
// tiles, optional transposed store; the k-point matrices are combined per k-point by one elementwise pass.
#include "common.cuh"

namespace dftfe_b200 {

namespace {

constexpr int TM = 128, TN = 128;
constexpr int PITCH = TN + 4;  // = 4 mod 16: conflict-free fragment LDS.64
constexpr int MMA_WARPS = 8;
constexpr int THREADS = (MMA_WARPS + 1) * 32;
constexpr int KC = 16;  // quadrature points per stage
constexpr int STAGES = 6;
constexpr size_t STAGE_DOUBLES = 2 * (size_t)KC * PITCH + KC;  // A rows, B rows, weights
constexpr size_t SMEM = STAGES * STAGE_DOUBLES * sizeof(double) + 2 * STAGES * sizeof(uint64_t);

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

// Nt[q][I] (row pitch npad, zero padded) = N[I][q]
__global__ void transpose_shape_kernel(const double *__restrict__ N, int n, int nq, int nqPad, int npad,
                                       double *__restrict__ Nt) {
  const int64_t total = (int64_t)nqPad * npad;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int q = idx / npad, i = idx % npad;
    Nt[idx] = (q < nq && i < n) ? N[(size_t)i * nq + q] : 0.0;
  }
}

// wPad[c][q] (row pitch nqPad, zero padded) = w[c][q]
__global__ void pad_weights_kernel(const double *__restrict__ w, int64_t nC, int nq, int nqPad,
                                   double *__restrict__ wPad) {
  const int64_t total = nC * nqPad;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / nqPad;
    const int q = idx % nqPad;
    wPad[idx] = q < nq ? w[c * nq + q] : 0.0;
  }
}

struct HamArgs {
  const double *Nt;        // nqPad x npad
  const double *wPad;      // nC x nqPad
  const double *K;         // n x n shared, or nC x n x n
  const double *kscale;    // nC or nullptr (factor on the shared K)
  const double *extCorr;   // nC x n x n or nullptr
  double *H;               // nC x n x n
  int64_t nC;
  int n, npad, nqPad, kPerCell;
};

__global__ void __launch_bounds__(THREADS, 1) ham_assemble_kernel(HamArgs g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *st = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + STAGES * STAGE_DOUBLES * sizeof(double));
  uint64_t *empty = full + STAGES;
  const int tid = threadIdx.x, lane = tid & 31, pwarp = tid >> 5;
  const int nT = g.npad / TM;                 // tiles per dimension
  const int tilesPerCell = nT * (nT + 1) / 2;  // lower triangle
  const int64_t nItems = g.nC * tilesPerCell;
  const int nChunks = g.nqPad / KC;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MMA_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // lower-triangular tile (ti >= tj) of index t
  auto tile_of = [&](int t, int &ti, int &tj) {
    ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  };

  if (pwarp == MMA_WARPS) {
    // ===== producer: KC rows of N^T for the A and B column ranges + KC weights per stage =====
    uint32_t cnt = 0;
    for (int64_t item = blockIdx.x; item < nItems; item += gridDim.x) {
      const int64_t cell = item / tilesPerCell;
      int ti, tj;
      tile_of((int)(item % tilesPerCell), ti, tj);
      for (int c = 0; c < nChunks; ++c, ++cnt) {
        const int s = cnt % STAGES;
        const uint32_t ph = (cnt / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        double *sA = st + s * STAGE_DOUBLES;
        double *sB = sA + KC * PITCH;
        double *sW = sB + KC * PITCH;
        if (lane == 0)
          mbar_arrive_expect_tx(&full[s], (uint32_t)((KC * (TM + TN) + KC) * sizeof(double)));
        __syncwarp();
        const size_t q0 = (size_t)c * KC;
        if (lane < KC)
          tma_bulk_g2s(sA + lane * PITCH, g.Nt + (q0 + lane) * g.npad + ti * TM, TM * sizeof(double), &full[s]);
        else
          tma_bulk_g2s(sB + (lane - KC) * PITCH, g.Nt + (q0 + lane - KC) * g.npad + tj * TN, TN * sizeof(double),
                       &full[s]);
        if (lane == 0) tma_bulk_g2s(sW, g.wPad + (size_t)cell * g.nqPad + q0, KC * sizeof(double), &full[s]);
      }
    }
  } else {
    // ===== MMA warps: warp (wm, wn) owns rows [wm*64, +64) x cols [wn*32, +32) of the 128 x 128 tile =====
    const int wm = pwarp >> 2, wn = pwarp & 3;
    const int aoff = (lane & 3) * PITCH + wm * 64 + (lane >> 2);
    const int boff = (lane & 3) * PITCH + wn * 32 + (lane >> 2);
    uint32_t cnt = 0;
    for (int64_t item = blockIdx.x; item < nItems; item += gridDim.x) {
      const int64_t cell = item / tilesPerCell;
      int ti, tj;
      tile_of((int)(item % tilesPerCell), ti, tj);
      double acc[8][4][2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      for (int c = 0; c < nChunks; ++c, ++cnt) {
        const int s = cnt % STAGES;
        const uint32_t ph = (cnt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        const double *sA = st + s * STAGE_DOUBLES;
        const double *sB = sA + KC * PITCH;
        const double *sW = sB + KC * PITCH;
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
          double a[8], b[4];
          const double wk = sW[ks * 4 + (lane & 3)];  // weight of this lane's k index
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = sA[ks * 4 * PITCH + aoff + i * 8];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = sB[ks * 4 * PITCH + boff + j * 8] * wk;
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      // ---- epilogue: + 1/2 K (+ correction), write the tile and (off-diagonal tiles) its mirror image
      const int n = g.n;
      const double ks_ = g.kscale ? g.kscale[cell] : 1.0;
      const double *Kc = g.kPerCell ? g.K + (size_t)cell * n * n : g.K;  // nullptr: no stiffness term (mass-type matrix)
      const double *Ec = g.extCorr ? g.extCorr + (size_t)cell * n * n : nullptr;
      double *Hc = g.H + (size_t)cell * n * n;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int I = ti * TM + wm * 64 + i * 8 + (lane >> 2);
        if (I >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int J = tj * TN + wn * 32 + j * 8 + (lane & 3) * 2 + e;
            if (J >= n || J > I) continue;  // lower triangle only (diagonal tiles included): H_c comes out exactly symmetric
            const size_t ij = (size_t)I * n + J;
            double v = acc[i][j][e];
            if (Kc) v += 0.5 * ks_ * Kc[ij];
            if (Ec) v += Ec[ij];
            Hc[ij] = v;
            if (I != J) Hc[(size_t)J * n + I] = v;  // K and the correction matrix are symmetric too
          }
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------
// General (non-symmetric) assembly: Out_c(I,J) = sum_q N_I(q) * sum_t w_t[c][q] Tab_t[q][J]
// ---------------------------------------------------------------------------
constexpr int GKC = 8;        // quadrature points per stage
constexpr int GSTAGES = 4;
constexpr int GMAXT = 4;      // terms: table 0 = N, 1..3 = d_e N
constexpr size_t GSTAGE_DOUBLES = (size_t)(1 + GMAXT) * GKC * PITCH + GMAXT * GKC;
constexpr size_t GSMEM = GSTAGES * GSTAGE_DOUBLES * sizeof(double) + 2 * GSTAGES * sizeof(uint64_t);

struct GenArgs {
  const double *tabs;      // [4][nqPad][npad]: N^T, (d_x N)^T, (d_y N)^T, (d_z N)^T
  const double *w[GMAXT];  // nC x nqPad each
  int tab[GMAXT];
  int nTerms;
  double *Out;             // nC x n x n
  int transposed;          // 1: Out_c(J,I) = acc(I,J)
  int64_t nC;
  int n, npad, nqPad;
};

__global__ void __launch_bounds__(THREADS, 1) assemble_general_kernel(GenArgs g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *st = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + GSTAGES * GSTAGE_DOUBLES * sizeof(double));
  uint64_t *empty = full + GSTAGES;
  const int tid = threadIdx.x, lane = tid & 31, pwarp = tid >> 5;
  const int nT = g.npad / TM;
  const int tilesPerCell = nT * nT;
  const int64_t nItems = g.nC * tilesPerCell;
  const int nChunks = g.nqPad / GKC;
  const size_t tabStride = (size_t)g.nqPad * g.npad;

  if (tid == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MMA_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (pwarp == MMA_WARPS) {
    uint32_t cnt = 0;
    for (int64_t item = blockIdx.x; item < nItems; item += gridDim.x) {
      const int64_t cell = item / tilesPerCell;
      const int ti = (int)(item % tilesPerCell) / nT, tj = (int)(item % tilesPerCell) % nT;
      for (int c = 0; c < nChunks; ++c, ++cnt) {
        const int s = cnt % GSTAGES;
        mbar_wait(&empty[s], ((cnt / GSTAGES) & 1) ^ 1);
        double *sA = st + s * GSTAGE_DOUBLES;
        double *sB = sA + GKC * PITCH;             // [term][GKC][PITCH]
        double *sW = sB + GMAXT * GKC * PITCH;     // [term][GKC]
        if (lane == 0)
          mbar_arrive_expect_tx(&full[s], (uint32_t)((GKC * TM + g.nTerms * (GKC * TN + GKC)) * sizeof(double)));
        __syncwarp();
        const size_t q0 = (size_t)c * GKC;
        if (lane < GKC) tma_bulk_g2s(sA + lane * PITCH, g.tabs + (q0 + lane) * g.npad + ti * TM, TM * sizeof(double), &full[s]);
        for (int t = 0; t < g.nTerms; ++t) {
          if (lane < GKC)
            tma_bulk_g2s(sB + (t * GKC + lane) * PITCH, g.tabs + g.tab[t] * tabStride + (q0 + lane) * g.npad + tj * TN,
                         TN * sizeof(double), &full[s]);
          if (lane == GKC) tma_bulk_g2s(sW + t * GKC, g.w[t] + (size_t)cell * g.nqPad + q0, GKC * sizeof(double), &full[s]);
        }
      }
    }
  } else {
    const int wm = pwarp >> 2, wn = pwarp & 3;
    const int aoff = (lane & 3) * PITCH + wm * 64 + (lane >> 2);
    const int boff = (lane & 3) * PITCH + wn * 32 + (lane >> 2);
    uint32_t cnt = 0;
    for (int64_t item = blockIdx.x; item < nItems; item += gridDim.x) {
      const int64_t cell = item / tilesPerCell;
      const int ti = (int)(item % tilesPerCell) / nT, tj = (int)(item % tilesPerCell) % nT;
      double acc[8][4][2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      for (int c = 0; c < nChunks; ++c, ++cnt) {
        const int s = cnt % GSTAGES;
        mbar_wait(&full[s], (cnt / GSTAGES) & 1);
        const double *sA = st + s * GSTAGE_DOUBLES;
        const double *sB = sA + GKC * PITCH;
        const double *sW = sB + GMAXT * GKC * PITCH;
#pragma unroll
        for (int ks = 0; ks < GKC / 4; ++ks) {
          double a[8], b[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = sA[ks * 4 * PITCH + aoff + i * 8];
          for (int t = 0; t < g.nTerms; ++t) {
            const double wk = sW[t * GKC + ks * 4 + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] += sB[(t * GKC + ks * 4) * PITCH + boff + j * 8] * wk;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      const int n = g.n;
      double *Oc = g.Out + (size_t)cell * n * n;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int I = ti * TM + wm * 64 + i * 8 + (lane >> 2);
        if (I >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int J = tj * TN + wn * 32 + j * 8 + (lane & 3) * 2 + e;
            if (J >= n) continue;
            Oc[g.transposed ? (size_t)J * n + I : (size_t)I * n + J] = acc[i][j][e];
          }
      }
    }
  }
}

// tabs[0] = N^T, tabs[1 + e] = (d_e N)^T, each [nqPad][npad], zero padded
__global__ void transpose_tables_kernel(const double *__restrict__ N, const double *__restrict__ dN, int n, int nq,
                                        int nqPad, int npad, double *__restrict__ tabs) {
  const int64_t per = (int64_t)nqPad * npad;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < 4 * per;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(idx / per);
    const int q = (int)((idx % per) / npad), i = (int)(idx % npad);
    double v = 0.0;
    if (q < nq && i < n) v = t == 0 ? N[(size_t)i * nq + q] : dN[((size_t)(t - 1) * n + i) * nq + q];
    tabs[idx] = v;
  }
}

// GGA weights: w0[c][q] = 1/2 vEffJxW, u_e[c][q] = 2 sum_d g[c][q][d] Jinv[c][d][e]  (padded rows of nqPad)
__global__ void gga_weights_kernel(const double *__restrict__ vEffJxW, const double *__restrict__ g,
                                   const double *__restrict__ invJac, int64_t nC, int nq, int nqPad,
                                   double *__restrict__ w) {
  const int64_t per = nC * nqPad;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < per; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / nqPad;
    const int q = (int)(idx % nqPad);
    double w0 = 0.0, u[3] = {0.0, 0.0, 0.0};
    if (q < nq) {
      w0 = 0.5 * vEffJxW[c * nq + q];
      const double *gq = g + (c * nq + q) * 3;
      for (int e = 0; e < 3; ++e)
        for (int d = 0; d < 3; ++d) {
          const double J = invJac ? invJac[c * 9 + 3 * d + e] : (d == e ? 1.0 : 0.0);
          u[e] += 2.0 * gq[d] * J;
        }
    }
    w[idx] = w0;
    w[per + idx] = u[0];
    w[2 * per + idx] = u[1];
    w[3 * per + idx] = u[2];
  }
}

// H_c = G_c + G_c^T + 1/2 ks K (+ ext), in place over G (each unordered pair handled by one thread)
__global__ void gga_finish_kernel(double *__restrict__ H, const double *__restrict__ K, int kPerCell,
                                  const double *__restrict__ kscale, const double *__restrict__ ext, int64_t nC, int n) {
  const int64_t nn = (int64_t)n * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nC * nn;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / nn;
    const int I = (int)((idx % nn) / n), J = (int)(idx % n);
    if (J > I) continue;
    double *Hc = H + c * nn;
    const int64_t ij = (int64_t)I * n + J, ji = (int64_t)J * n + I;
    const double ks_ = kscale ? kscale[c] : 1.0;
    const double *Kc = kPerCell ? K + c * nn : K;
    double v = Hc[ij] + Hc[ji] + 0.5 * ks_ * Kc[ij];
    if (ext) v += ext[c * nn + ij];
    Hc[ij] = v;
    Hc[ji] = v;
  }
}

// -JxW padded: w[c][q] = -JxW[c][q]
__global__ void neg_pad_weights_kernel(const double *__restrict__ w, int64_t nC, int nq, int nqPad, double sign,
                                       double *__restrict__ wPad) {
  const int64_t total = nC * nqPad;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / nqPad;
    const int q = (int)(idx % nqPad);
    wPad[idx] = q < nq ? sign * w[c * nq + q] : 0.0;
  }
}

// H_k[c](I,J) = (Hreal + 1/2 |k|^2 Mc)(I,J) + i sum_e kappa_e(c) E_e(I,J),  kappa_e = sum_d k_d Jinv[c][d][e]
__global__ void kpoint_combine_kernel(const double *__restrict__ Hreal, const double *__restrict__ Mc,
                                      const double *__restrict__ E, const double *__restrict__ invJac, double kx,
                                      double ky, double kz, int64_t nC, int n, double *__restrict__ Hk) {
  const int64_t nn = (int64_t)n * n, total = nC * nn;
  const double k2h = 0.5 * (kx * kx + ky * ky + kz * kz);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / nn;
    double kap[3];
    for (int e = 0; e < 3; ++e) {
      if (invJac)
        kap[e] = kx * invJac[c * 9 + e] + ky * invJac[c * 9 + 3 + e] + kz * invJac[c * 9 + 6 + e];
      else
        kap[e] = e == 0 ? kx : (e == 1 ? ky : kz);
    }
    Hk[2 * idx] = Hreal[idx] + k2h * Mc[idx];
    Hk[2 * idx + 1] = kap[0] * E[idx] + kap[1] * E[total + idx] + kap[2] * E[2 * total + idx];
  }
}

}  // namespace

int compute_cell_hamiltonian(dftfe_b200_ctx *ctx, int nq, const double *shapeValues, const double *vEffJxW,
                             const double *gradIntegral, int gradPerCell, const double *cellKScale,
                             const double *extPotCorr, double *H) {
  // (real matrices: the k-point build assembles its real part here and adds the k-dependent terms with
  // compute_cell_hamiltonian_kpoints; gradIntegral == nullptr gives the bare N diag(w) N^T)
  DB_CHECK(nq >= 1 && shapeValues && vEffJxW && H, "compute_cell_hamiltonian: null argument");
  if (ctx->nC == 0) return 0;
  DB_DYN_SMEM(ctx, ham_assemble_kernel, SMEM);
  const int n = ctx->n;
  const int npad = ((n + TM - 1) / TM) * TM;
  const int nqPad = ((nq + KC - 1) / KC) * KC;
  DB_TRY(ctx->hamNt.alloc((size_t)nqPad * npad));
  DB_TRY(ctx->hamW.alloc((size_t)ctx->nC * nqPad));
  ctx->launches += 2;
  transpose_shape_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(shapeValues, n, nq, nqPad, npad, ctx->hamNt.p);
  pad_weights_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(vEffJxW, ctx->nC, nq, nqPad, ctx->hamW.p);
  HamArgs g;
  g.Nt = ctx->hamNt.p;
  g.wPad = ctx->hamW.p;
  g.K = gradIntegral;
  g.kscale = cellKScale;
  g.extCorr = extPotCorr;
  g.H = H;
  g.nC = ctx->nC;
  g.n = n;
  g.npad = npad;
  g.nqPad = nqPad;
  g.kPerCell = gradPerCell ? 1 : 0;
  const int nT = npad / TM;
  const int64_t nItems = ctx->nC * (int64_t)(nT * (nT + 1) / 2);
  {
    ProfScope ps(ctx, "ham_assembly");
    ham_assemble_kernel<<<(int)std::min<int64_t>(nItems, ctx->num_sms), THREADS, SMEM, ctx->stream>>>(g);
  }
  DB_CUDA(cudaGetLastError());
  return 0;
}


static int general_setup(dftfe_b200_ctx *ctx, int nq, const double *shapeValues, const double *shapeGradValues,
                         int &npad, int &nqPad) {
  const int n = ctx->n;
  npad = ((n + TM - 1) / TM) * TM;
  nqPad = ((nq + GKC - 1) / GKC) * GKC;
  DB_DYN_SMEM(ctx, assemble_general_kernel, GSMEM);
  DB_TRY(ctx->hamTabs.alloc((size_t)4 * nqPad * npad));
  ctx->launches += 1;
  transpose_tables_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(shapeValues, shapeGradValues, n, nq, nqPad, npad,
                                                                    ctx->hamTabs.p);
  DB_CUDA(cudaGetLastError());
  return 0;
}

static int launch_general(dftfe_b200_ctx *ctx, GenArgs &g) {
  const int nT = g.npad / TM;
  const int64_t nItems = g.nC * (int64_t)(nT * nT);
  ProfScope ps(ctx, "ham_assembly");
  assemble_general_kernel<<<(int)std::min<int64_t>(nItems, ctx->num_sms), THREADS, GSMEM, ctx->stream>>>(g);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// hamMatrixKernelGGAMemOpt (hamiltonianMatrixCalculatorFlattenedDevice.cc:281-440), real part:
//   H_c(I,J) = 1/2 K + sum_q [ vEffJxW N_I N_J + 2 sum_d g_d (d_d N_I N_J + d_d N_J N_I) ]  (+ correction)
// derExcSigmaGradRhoJxW: [nC][nq][3]; shapeGradValues: [3][n][nq] reference-cell derivatives; invJac: [nC][3][3] with
// Jinv[c][d][e] = d xi_e / d x_d (nullptr: identity)
int compute_cell_hamiltonian_gga(dftfe_b200_ctx *ctx, int nq, const double *shapeValues, const double *shapeGradValues,
                                 const double *invJac, const double *vEffJxW, const double *derExcSigmaGradRhoJxW,
                                 const double *gradIntegral, int gradPerCell, const double *cellKScale,
                                 const double *extPotCorr, double *H) {
  DB_CHECK(nq >= 1 && shapeValues && shapeGradValues && vEffJxW && derExcSigmaGradRhoJxW && gradIntegral && H,
           "compute_cell_hamiltonian_gga: null argument");
  if (ctx->nC == 0) return 0;
  int npad, nqPad;
  DB_TRY(general_setup(ctx, nq, shapeValues, shapeGradValues, npad, nqPad));
  const size_t per = (size_t)ctx->nC * nqPad;
  DB_TRY(ctx->hamW.alloc(4 * per));
  ctx->launches += 2;
  gga_weights_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(vEffJxW, derExcSigmaGradRhoJxW, invJac, ctx->nC, nq, nqPad,
                                                               ctx->hamW.p);
  GenArgs g;
  g.tabs = ctx->hamTabs.p;
  g.nTerms = 4;
  for (int t = 0; t < 4; ++t) {
    g.w[t] = ctx->hamW.p + t * per;
    g.tab[t] = t;
  }
  g.Out = H;  // G first, finished in place
  g.transposed = 0;
  g.nC = ctx->nC;
  g.n = ctx->n;
  g.npad = npad;
  g.nqPad = nqPad;
  DB_TRY(launch_general(ctx, g));
  gga_finish_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(H, gradIntegral, gradPerCell ? 1 : 0, cellKScale, extPotCorr,
                                                              ctx->nC, ctx->n);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// The k-point dependent terms of the complex kernels (:119-278, 442-657): for every k-point
//   H_k = (Hreal + 1/2 |k|^2 sum_q JxW N_I N_J) - i sum_d k_d sum_q JxW d_d N_I N_J
// Hreal: [nC][n][n] from compute_cell_hamiltonian / _gga (this spin channel); Hk: [nk][nC][n][n] complex (re, im)
int compute_cell_hamiltonian_kpoints(dftfe_b200_ctx *ctx, int nq, const double *shapeValues,
                                     const double *shapeGradValues, const double *invJac, const double *JxW,
                                     const double *Hreal, int nk, const double *kpoints_h, double *Hk) {
  DB_CHECK(nq >= 1 && shapeValues && shapeGradValues && JxW && Hreal && Hk && nk >= 1 && kpoints_h,
           "compute_cell_hamiltonian_kpoints: null argument");
  if (ctx->nC == 0) return 0;
  const int n = ctx->n;
  const size_t nn = (size_t)ctx->nC * n * n;
  DB_TRY(ctx->hamMc.alloc(nn));
  DB_TRY(ctx->hamE.alloc(3 * nn));
  // Mc = N diag(JxW) N^T through the symmetric kernel (no stiffness term)
  DB_TRY(compute_cell_hamiltonian(ctx, nq, shapeValues, JxW, nullptr, 0, nullptr, nullptr, ctx->hamMc.p));
  // E_e(I,J) = - sum_q JxW d_e N_I N_J = F[N; -JxW d_e N](J,I): general kernel, transposed store
  int npad, nqPad;
  DB_TRY(general_setup(ctx, nq, shapeValues, shapeGradValues, npad, nqPad));
  const size_t per = (size_t)ctx->nC * nqPad;
  DB_TRY(ctx->hamW.alloc(std::max<size_t>(per, ctx->hamW.n)));
  ctx->launches += 1;
  neg_pad_weights_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(JxW, ctx->nC, nq, nqPad, -1.0, ctx->hamW.p);
  DB_CUDA(cudaGetLastError());
  for (int e = 0; e < 3; ++e) {
    GenArgs g;
    g.tabs = ctx->hamTabs.p;
    g.nTerms = 1;
    g.w[0] = ctx->hamW.p;
    g.tab[0] = 1 + e;
    for (int t = 1; t < GMAXT; ++t) {
      g.w[t] = nullptr;
      g.tab[t] = 0;
    }
    g.Out = ctx->hamE.p + e * nn;
    g.transposed = 1;
    g.nC = ctx->nC;
    g.n = n;
    g.npad = npad;
    g.nqPad = nqPad;
    DB_TRY(launch_general(ctx, g));
  }
  for (int k = 0; k < nk; ++k) {
    ctx->launches += 1;
    kpoint_combine_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(Hreal, ctx->hamMc.p, ctx->hamE.p, invJac,
                                                                    kpoints_h[3 * k], kpoints_h[3 * k + 1],
                                                                    kpoints_h[3 * k + 2], ctx->nC, n, Hk + 2 * nn * k);
  }
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dftfe_b200
