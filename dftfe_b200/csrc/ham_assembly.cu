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
// dftfe_b200_set_cell_hamiltonian.  Real (Gamma-point, LDA-type local potential) build; the GGA gradient terms and
// the k-point terms are not provided.
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
      const double *Kc = g.kPerCell ? g.K + (size_t)cell * n * n : g.K;
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
            double v = acc[i][j][e] + 0.5 * ks_ * Kc[ij];
            if (Ec) v += Ec[ij];
            Hc[ij] = v;
            if (I != J) Hc[(size_t)J * n + I] = v;  // K and the correction matrix are symmetric too
          }
        }
      }
    }
  }
}

}  // namespace

int compute_cell_hamiltonian(dftfe_b200_ctx *ctx, int nq, const double *shapeValues, const double *vEffJxW,
                             const double *gradIntegral, int gradPerCell, const double *cellKScale,
                             const double *extPotCorr, double *H) {
  DB_CHECK(!ctx->cplx, "compute_cell_hamiltonian: the k-point (complex) terms are not provided");
  DB_CHECK(nq >= 1 && shapeValues && vEffJxW && gradIntegral && H, "compute_cell_hamiltonian: null argument");
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

}  // namespace dftfe_b200
