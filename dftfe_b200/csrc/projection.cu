// Hand-written FP64 DMMA GEMMs for the subspace projections and the rotation.
//
//   xty:  C(i,j) = sum_m A[m, i] * B[m, j]        A, B row-major with the long dimension m = local DoFs
//         -> S = X^T X (fillParallelOverlapMatScalapack, src/linAlg/linearAlgebraOperationsDevice.cc:3078-3240)
//         -> Hp = X^T (H~ X) per column block (XtHX, src/dftOperator/kohnShamDFTOperatorDevice.cc:4096-4112)
//   xq :  Out[m, j] = sum_k X[m, k] * Q[k, j]       (subspaceRotationScalapack, linearAlgebraOperationsDevice.cc:2196-2210)
//
// The reference issues cuBLAS Dgemm for these (K22/K25 in SURVEY.md 2.4).  Here: 128x128 CTA tiles, eight MMA
// warps (2 x 4, 64x32 each = 8x4 DMMA.8x8x4 accumulators) fed from a multi-stage shared-memory ring that a
// producer warp fills with 1-D TMA bulk copies (row segments, completion on mbarriers); row pitches are = 4 mod
// 16 doubles so that every fragment LDS.64 is bank-conflict free.  xty is split along m into equal segments
// (persistent CTAs, static round-robin); each segment writes its partial tile to a workspace and a second
// kernel sums the partials of a tile in a fixed order - deterministic, no atomics - and only lower-triangular
// tiles are computed, as the reference does.
#include "common.cuh"

namespace dftfe_b200 {

namespace {

constexpr int TM = 128, TN = 128;   // CTA tile
constexpr int PITCH = TN + 4;       // doubles; = 4 mod 16
constexpr int MMA_WARPS = 8;
constexpr int THREADS = (MMA_WARPS + 1) * 32;

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

// ---------------------------------------------------------------------------
// xty: partial tiles of C = A^T B
// ---------------------------------------------------------------------------
constexpr int XTY_KC = 16;      // rows of m per stage
constexpr int XTY_STAGES = 6;
constexpr size_t XTY_STAGE_DOUBLES = 2 * (size_t)XTY_KC * PITCH;
constexpr size_t XTY_SMEM = XTY_STAGES * XTY_STAGE_DOUBLES * sizeof(double) + 2 * XTY_STAGES * sizeof(uint64_t);

struct XtyTile {
  int i0, j0;  // column offsets into A and B
  int wa, wb;  // columns that exist (<= TM / TN, even): a ragged last tile copies only these; the stale (finite)
               // values beyond them only reach rows / columns of the tile that the reduction masks
};

__global__ void __launch_bounds__(THREADS, 1)
xty_partial_kernel(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb, int64_t M,
                   const XtyTile *__restrict__ tiles, int nTiles, int nSeg, int64_t chunksPerSeg,
                   double *__restrict__ ws) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *st = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + XTY_STAGES * XTY_STAGE_DOUBLES * sizeof(double));
  uint64_t *empty = full + XTY_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, pwarp = tid >> 5;
  const int64_t nChunks = (M + XTY_KC - 1) / XTY_KC;
  const int nItems = nTiles * nSeg;

  for (int i = tid; i < (int)(XTY_STAGES * XTY_STAGE_DOUBLES); i += THREADS) st[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < XTY_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MMA_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (pwarp == MMA_WARPS) {
    // ===== producer: one row segment (TM or TN doubles) per lane per stage =====
    uint32_t n = 0;  // global stage counter
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
      const XtyTile t = tiles[item / nSeg];
      const int seg = item % nSeg;
      const int64_t c0 = seg * chunksPerSeg, c1 = min((long long)nChunks, (long long)(c0 + chunksPerSeg));
      for (int64_t c = c0; c < c1; ++c, ++n) {
        const int s = n % XTY_STAGES;
        const uint32_t ph = (n / XTY_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        double *sA = st + s * XTY_STAGE_DOUBLES;
        double *sB = sA + XTY_KC * PITCH;
        const int64_t m0 = c * XTY_KC;
        const int rows = (int)min((long long)XTY_KC, (long long)(M - m0));
        if (rows < XTY_KC) {  // ragged last chunk: zero the missing rows (generic proxy), then fence
          for (int i = lane; i < (XTY_KC - rows) * PITCH; i += 32) {
            sA[rows * PITCH + i] = 0.0;
            sB[rows * PITCH + i] = 0.0;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
        }
        if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)(rows * (t.wa + t.wb) * sizeof(double)));
        __syncwarp();
        if (lane < rows)
          tma_bulk_g2s(sA + lane * PITCH, A + (size_t)(m0 + lane) * lda + t.i0, t.wa * sizeof(double), &full[s]);
        else if (lane >= 16 && lane - 16 < rows)
          tma_bulk_g2s(sB + (lane - 16) * PITCH, B + (size_t)(m0 + lane - 16) * ldb + t.j0, t.wb * sizeof(double),
                       &full[s]);
      }
    }
  } else {
    // ===== MMA warps: warp (wm, wn) owns rows [wm*64, +64) x cols [wn*32, +32) of the tile =====
    const int wm = pwarp >> 2, wn = pwarp & 3;
    const int aoff = (lane & 3) * PITCH + wm * 64 + (lane >> 2);
    const int boff = (lane & 3) * PITCH + wn * 32 + (lane >> 2);
    uint32_t n = 0;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
      const int seg = item % nSeg;
      const int64_t c0 = seg * chunksPerSeg, c1 = min((long long)nChunks, (long long)(c0 + chunksPerSeg));
      double acc[8][4][2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      for (int64_t c = c0; c < c1; ++c, ++n) {
        const int s = n % XTY_STAGES;
        const uint32_t ph = (n / XTY_STAGES) & 1;
        mbar_wait(&full[s], ph);
        const double *sA = st + s * XTY_STAGE_DOUBLES;
        const double *sB = sA + XTY_KC * PITCH;
#pragma unroll
        for (int ks = 0; ks < XTY_KC / 4; ++ks) {
          double a[8], b[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = sA[ks * 4 * PITCH + aoff + i * 8];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = sB[ks * 4 * PITCH + boff + j * 8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      // partial tile -> workspace slot [item][TM][TN] (row i of C = column of A)
      double *w = ws + (size_t)item * TM * TN;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = wm * 64 + i * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = wn * 32 + j * 8 + (lane & 3) * 2;
          *reinterpret_cast<double2 *>(w + (size_t)r * TN + col) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
      }
    }
  }
}

// C(i0+r, j0+c) [column-major, ld N] = sum over segments, in segment order
__global__ void xty_reduce_kernel(const double *__restrict__ ws, const XtyTile *__restrict__ tiles, int nSeg,
                                  double *__restrict__ C, int ldc, int iBase, int jBase, int rowsValid,
                                  int colsValid) {
  const int tile = blockIdx.x;
  const XtyTile t = tiles[tile];
  for (int e = threadIdx.x; e < TM * TN; e += blockDim.x) {
    const int r = e / TN, c = e % TN;
    double s = 0.0;
    for (int k = 0; k < nSeg; ++k) s += ws[((size_t)tile * nSeg + k) * TM * TN + e];
    const int gi = t.i0 - iBase + r, gj = t.j0 - jBase + c;
    if (gi < rowsValid && gj < colsValid) C[(size_t)(gi) + (size_t)(gj)*ldc] = s;
  }
}

// ---------------------------------------------------------------------------
// xq: Out = X * Q  (row-major everywhere), 128 x 128 output tiles, K = N
// ---------------------------------------------------------------------------
constexpr int XQ_KC = 32;
constexpr int XQ_PA = XQ_KC + 4;  // pitch of the A stage (rows of X, k fastest); 36 = 4 mod 16
constexpr int XQ_STAGES = 3;
constexpr size_t XQ_STAGE_DOUBLES = (size_t)TM * XQ_PA + (size_t)XQ_KC * PITCH;
constexpr size_t XQ_SMEM = XQ_STAGES * XQ_STAGE_DOUBLES * sizeof(double) + 2 * XQ_STAGES * sizeof(uint64_t);

__global__ void __launch_bounds__(THREADS, 1)
xq_kernel(const double *__restrict__ X, int N, int64_t rows, const double *__restrict__ Q, int ldq, int Nout,
          double *__restrict__ Out, int ldo) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *st = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + XQ_STAGES * XQ_STAGE_DOUBLES * sizeof(double));
  uint64_t *empty = full + XQ_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, pwarp = tid >> 5;
  const int nColTiles = (Nout + TN - 1) / TN;
  const int64_t nRowBlocks = (rows + TM - 1) / TM;
  const int64_t nItems = nRowBlocks * nColTiles;
  const int nK = (N + XQ_KC - 1) / XQ_KC;

  for (int i = tid; i < (int)(XQ_STAGES * XQ_STAGE_DOUBLES); i += THREADS) st[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < XQ_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MMA_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (pwarp == MMA_WARPS) {
    uint32_t n = 0;
    for (int64_t item = blockIdx.x; item < nItems; item += gridDim.x) {
      const int64_t m0 = (item / nColTiles) * TM;
      const int j0 = (int)(item % nColTiles) * TN;
      const int vrows = (int)min((long long)TM, (long long)(rows - m0));
      const int nw = min(TN, Nout - j0);  // ragged last column tile: stale columns beyond nw are masked at the store
      for (int kc = 0; kc < nK; ++kc, ++n) {
        const int s = n % XQ_STAGES;
        const uint32_t ph = (n / XQ_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        double *sA = st + s * XQ_STAGE_DOUBLES;
        double *sB = sA + TM * XQ_PA;
        const int kw = min(XQ_KC, N - kc * XQ_KC);  // ragged last k chunk
        if (vrows < TM || kw < XQ_KC) {
          // rows of X beyond the end must read as zero; a short k chunk zeroes the Q rows it does not load (the
          // matching stale columns of the X stage are finite and multiply those zeros)
          if (vrows < TM)
            for (int i = lane; i < (TM - vrows) * XQ_PA; i += 32) sA[vrows * XQ_PA + i] = 0.0;
          if (kw < XQ_KC)
            for (int i = lane; i < (XQ_KC - kw) * PITCH; i += 32) sB[kw * PITCH + i] = 0.0;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
        }
        if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)((vrows * kw + kw * nw) * sizeof(double)));
        __syncwarp();
        for (int r = lane; r < vrows; r += 32)
          tma_bulk_g2s(sA + r * XQ_PA, X + (size_t)(m0 + r) * N + kc * XQ_KC, kw * sizeof(double), &full[s]);
        if (lane < kw)
          tma_bulk_g2s(sB + lane * PITCH, Q + (size_t)(kc * XQ_KC + lane) * ldq + j0, nw * sizeof(double), &full[s]);
      }
    }
  } else {
    const int wm = pwarp >> 2, wn = pwarp & 3;
    const int aoff = (wm * 64 + (lane >> 2)) * XQ_PA + (lane & 3);
    const int boff = (lane & 3) * PITCH + wn * 32 + (lane >> 2);
    uint32_t n = 0;
    for (int64_t item = blockIdx.x; item < nItems; item += gridDim.x) {
      const int64_t m0 = (item / nColTiles) * TM;
      const int j0 = (int)(item % nColTiles) * TN;
      double acc[8][4][2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      for (int kc = 0; kc < nK; ++kc, ++n) {
        const int s = n % XQ_STAGES;
        const uint32_t ph = (n / XQ_STAGES) & 1;
        mbar_wait(&full[s], ph);
        const double *sA = st + s * XQ_STAGE_DOUBLES;
        const double *sB = sA + TM * XQ_PA;
#pragma unroll
        for (int ks = 0; ks < XQ_KC / 4; ++ks) {
          double a[8], b[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = sA[aoff + i * 8 * XQ_PA + ks * 4];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = sB[ks * 4 * PITCH + boff + j * 8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + wm * 64 + i * 8 + (lane >> 2);
        if (r < rows) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = j0 + wn * 32 + j * 8 + (lane & 3) * 2;
            if (col < Nout)
              *reinterpret_cast<double2 *>(Out + (size_t)r * ldo + col) = make_double2(acc[i][j][0], acc[i][j][1]);
          }
        }
      }
    }
  }
}

__global__ void transpose_square_kernel(const double *__restrict__ in, double *__restrict__ out, int N) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (by + r < N && bx + threadIdx.x < N) tile[r][threadIdx.x] = in[(size_t)(by + r) * N + bx + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (bx + r < N && by + threadIdx.x < N) out[(size_t)(bx + r) * N + by + threadIdx.x] = tile[threadIdx.x][r];
}

// number of equal m-segments per tile so that tiles*segments fills the persistent grid evenly
int pick_segments(int nTiles, int64_t nChunks, int nSms) {
  int best = 1;
  double bestEff = 0.0;
  const int64_t minChunks = 128;  // keep a segment's k-loop long enough to amortise its prologue/epilogue
  for (int s = 1; s <= 256; ++s) {
    if (nChunks / s < minChunks && s > 1) break;
    const int64_t items = (int64_t)nTiles * s;
    const int64_t waves = (items + nSms - 1) / nSms;
    const double eff = (double)items / (double)(waves * nSms);
    if (eff > bestEff + 1e-9) {
      bestEff = eff;
      best = s;
    }
  }
  return best;
}

}  // namespace

bool dmma_projection_usable(const dftfe_b200_ctx *ctx, int N, int lda, int ldb, int i0, int j0, int nRowsC,
                            int nColsC) {
  (void)ctx;
  // 16-byte aligned row segments of even length (the last tile of a row / column of tiles may be ragged)
  return (nRowsC % 2 == 0) && (nColsC % 2 == 0) && (lda % 2 == 0) && (ldb % 2 == 0) && (i0 % 2 == 0) &&
         (j0 % 2 == 0) && N > 0 && nRowsC > 0 && nColsC > 0;
}

// C[iBase.., jBase..] (column-major, ld = ldc) = A[:, iBase:iBase+nRowsC]^T B[:, jBase:jBase+nColsC], lower tiles
// only when `lowerOnly` (tile row offset >= tile column offset in the global N x N numbering, given by
// iGlobal/jGlobal).
int launch_xty(dftfe_b200_ctx *ctx, const double *A, int lda, int iOff, const double *B, int ldb, int jOff,
               int nRowsC, int nColsC, int iGlobal, int jGlobal, bool lowerOnly, double *C, int ldc) {
  DB_DYN_SMEM(ctx, xty_partial_kernel, XTY_SMEM);
  std::vector<XtyTile> tiles;
  for (int j = 0; j < nColsC; j += TN)
    for (int i = 0; i < nRowsC; i += TM)
      if (!lowerOnly || (iGlobal + std::min(i + TM, nRowsC) > jGlobal + j))
        tiles.push_back(XtyTile{iOff + i, jOff + j, std::min(TM, nRowsC - i), std::min(TN, nColsC - j)});
  const int nTiles = (int)tiles.size();
  if (nTiles == 0 || ctx->M == 0) return 0;
  const int64_t nChunks = (ctx->M + XTY_KC - 1) / XTY_KC;
  const int nSeg = pick_segments(nTiles, nChunks, ctx->num_sms);
  const int64_t chunksPerSeg = (nChunks + nSeg - 1) / nSeg;
  DB_CHECK(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0,
           "DMMA projection: operand base pointers must be 16-byte aligned");
  DB_TRY(ctx->projTiles.upload(reinterpret_cast<const int32_t *>(tiles.data()), (size_t)nTiles * 4, ctx->stream));
  DB_TRY(ctx->projWs.alloc((size_t)nTiles * nSeg * TM * TN));
  {
    ProfScope ps(ctx, "projection", 2);
    const int grid = std::min(nTiles * nSeg, ctx->num_sms);
    xty_partial_kernel<<<grid, THREADS, XTY_SMEM, ctx->stream>>>(
        A, lda, B, ldb, ctx->M, reinterpret_cast<const XtyTile *>(ctx->projTiles.p), nTiles, nSeg, chunksPerSeg,
        ctx->projWs.p);
    xty_reduce_kernel<<<nTiles, 256, 0, ctx->stream>>>(ctx->projWs.p,
                                                       reinterpret_cast<const XtyTile *>(ctx->projTiles.p), nSeg, C,
                                                       ldc, iOff, jOff, nRowsC, nColsC);
  }
  DB_CUDA(cudaGetLastError());
  return 0;
}

bool dmma_rotation_usable(int N, int Nout, int ldq, int ldo) {
  return Nout > 0 && N > 0 && Nout % 2 == 0 && N % 2 == 0 && ldq % 2 == 0 && ldo % 2 == 0;
}

// Out[rows x Nout] (ld ldo) = X[rows x N] (ld N) * Q[N x Nout] (row-major, ld ldq)
int launch_xq(dftfe_b200_ctx *ctx, const double *X, int N, int64_t rows, const double *Qrm, int ldq, int Nout,
              double *Out, int ldo) {
  DB_DYN_SMEM(ctx, xq_kernel, XQ_SMEM);
  if (rows == 0) return 0;
  ProfScope ps(ctx, "rotation");
  DB_CHECK(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Qrm) | reinterpret_cast<uintptr_t>(Out)) & 15) == 0,
           "DMMA rotation: operand base pointers must be 16-byte aligned");
  const int64_t items = ((rows + TM - 1) / TM) * ((Nout + TN - 1) / TN);
  const int grid = (int)std::min<int64_t>(items, ctx->num_sms);
  xq_kernel<<<grid, THREADS, XQ_SMEM, ctx->stream>>>(X, N, rows, Qrm, ldq, Nout, Out, ldo);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_transpose_square(dftfe_b200_ctx *ctx, const double *in, double *out, int N) {
  ctx->launches += 1;
  dim3 grid((N + 31) / 32, (N + 31) / 32), block(32, 8);
  transpose_square_kernel<<<grid, block, 0, ctx->stream>>>(in, out, N);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dftfe_b200
