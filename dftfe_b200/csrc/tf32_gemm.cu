// FP32-accuracy GEMMs of the mixed-precision projections / rotations on the 5th-generation tensor cores:
// tcgen05.mma kind::tf32 with TMEM accumulators, TMA-fed (cp.async.bulk.tensor, 128-byte swizzle) shared-memory
// stages, 3xTF32 operand splitting.
//
// Reference: the FP32 blocks of fillParallelOverlapMatMixedPrecScalapack
// (src/linAlg/linearAlgebraOperationsDevice.cc:3543-3798, cuBLAS Sgemm at :3674-3691), of
// XtHXMixedPrecOverlapComputeCommun (src/dftOperator/kohnShamDFTOperatorDevice.cc:4550-5080) and of
// subspaceRotationCGS/RRMixedPrecScalapack (:2243-3076) - all cuBLAS Sgemm there.
//
// One kernel computes  D[m][n] = sum_k A[m][k] * B[n][k]  (both operands K-major: k contiguous) in 128 x 128 tiles:
//   projections: A = X^T (wavefunction index x DoFs), B = (X or H~X)^T, k = local DoFs (split-k over the CTAs,
//                deterministic two-stage reduction);  rotation: A = X (DoFs x N), B = Q^T, k = wavefunction index.
// TF32 keeps 11 significand bits, so every FP32 operand x is stored as two TF32-exact arrays
//   hi = x with the 13 low mantissa bits cleared,  lo = (x - hi) likewise truncated        (x = hi + lo + O(2^-22 |x|))
// and  A B^T ~= Ahi Bhi^T + Ahi Blo^T + Alo Bhi^T  accumulates in FP32 in tensor memory: three tensor-core passes per
// k-block give FP32-class accuracy (the dropped lo*lo term and the truncation are O(2^-22)) at ~1/3 of the TF32 rate,
// still an order of magnitude above the FP64 DMMA rate.  The split arrays are produced by the FP64 -> FP32
// conversion pass the reference also makes (its XSP copy), so the GEMM kernel itself is a pure
// TMA -> tcgen05.mma -> tcgen05.ld pipeline:
//   warp 0   : TMA producer (one elected lane), 4 tile loads per stage (Ahi, Alo, Bhi, Blo), 3 stages x 64 KB
//   warp 1   : TMEM allocation + the single-thread tcgen05.mma issuer (12 MMAs of 128x128x8 per k-block)
//   warps 2-5: epilogue - tcgen05.ld of the 128 x 128 FP32 accumulator (two accumulator stages in TMEM, so the next
//              tile's MMAs run under this tile's drain), direct row-major store or split-k partial tile
#include <cuda.h>

#include "common.cuh"

namespace dftfe_b200 {

namespace {

constexpr int TILE = 128;            // UMMA M = N = 128
constexpr int BK = 32;               // floats per k-block = one 128-byte swizzle row
constexpr int UK = 8;                // UMMA K for tf32
constexpr int STAGES = 3;
constexpr uint32_t TILE_BYTES = TILE * BK * sizeof(float);  // 16 KB
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;            // Ahi, Alo, Bhi, Blo
constexpr int THREADS = 6 * 32;
constexpr int ACC_STAGES = 2;
constexpr uint32_t TMEM_COLS = ACC_STAGES * TILE;  // 256 columns of 128 lanes x 32 bit
constexpr size_t SMEM_BYTES = 1024 /*alignment slack*/ + STAGES * STAGE_BYTES + 256 /*barriers + tmem ptr*/;

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
// 2-D tiled TMA load: box {BK floats (inner, k), TILE rows}; out-of-range elements arrive as zeros
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int k0, int row0, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(k0), "r"(row0), "r"(smem_u32(bar))
      : "memory");
}
// shared-memory matrix descriptor of a K-major [128 rows][32 floats] tile, 128-byte swizzle (cute::UMMA::SmemDescriptor:
// start address >> 4 in bits 0-13, leading byte offset (unused for swizzled K-major: 1) in 16-29, stride byte offset
// = 8 rows x 128 B = 1024 B >> 4 in 32-45, descriptor version 1 in 46-47, layout type SWIZZLE_128B = 2 in 61-63)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): FP32 accumulate (bits 4-5 = 1), A and B formats TF32 (= 2 in
// bits 7-9 / 10-12), both K-major (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

struct Tf32Params {
  int tilesM, tilesN;        // output tiles
  int nSeg;                  // split-k segments per tile (1: direct store)
  int kBlocks;               // ceil(K / BK)
  int kBlocksPerSeg;
  int rowsA, rowsB;          // valid rows of the A / B operands from the tile origins (ragged masks)
  int rowA0, rowB0;          // first operand row of tile (0, 0) in the tensor maps
  float *out;                // direct mode: row-major [rowsA][ldo], element (m, n) at out[m * ldo + n]
  int64_t ldo;
  float *ws;                 // split-k mode: partial tiles [item][128][128]
};

__global__ void __launch_bounds__(THREADS, 1)
tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                   const Tf32Params P) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte swizzled tiles must sit on 1024-byte boundaries
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *tiles = smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *empty = full + STAGES;
  uint64_t *accFull = empty + STAGES;
  uint64_t *accEmpty = accFull + ACC_STAGES;
  uint32_t *tmemPtr = reinterpret_cast<uint32_t *>(accEmpty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nItems = P.tilesM * P.tilesN * P.nSeg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(&accFull[s], 1);
      mbar_init(&accEmpty[s], 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemPtr)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmemBase = *tmemPtr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        // segment-major item order: the CTAs of a wave work on the same few k-segments of ALL tiles, so every
        // operand panel is fetched from HBM once and shared through L2 (tile-major order re-read them: ncu showed
        // 2x the algorithmic DRAM bytes)
        const int nTiles = P.tilesM * P.tilesN;
        const int tile = item % nTiles, seg = item / nTiles;
        const int tm = tile / P.tilesN, tn = tile % P.tilesN;
        const int kb0 = seg * P.kBlocksPerSeg, kb1 = min(P.kBlocks, kb0 + P.kBlocksPerSeg);
        for (int kb = kb0; kb < kb1; ++kb, ++n) {
          const int s = n % STAGES;
          mbar_wait(&empty[s], ((n / STAGES) & 1) ^ 1);
          unsigned char *st = tiles + (size_t)s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          tma_load_2d(st, &mapAhi, kb * BK, P.rowA0 + tm * TILE, &full[s]);
          tma_load_2d(st + TILE_BYTES, &mapAlo, kb * BK, P.rowA0 + tm * TILE, &full[s]);
          tma_load_2d(st + 2 * TILE_BYTES, &mapBhi, kb * BK, P.rowB0 + tn * TILE, &full[s]);
          tma_load_2d(st + 3 * TILE_BYTES, &mapBlo, kb * BK, P.rowB0 + tn * TILE, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t n = 0, it = 0;
      for (int item = blockIdx.x; item < nItems; item += gridDim.x, ++it) {
        const int seg = item / (P.tilesM * P.tilesN);
        const int kb0 = seg * P.kBlocksPerSeg, kb1 = min(P.kBlocks, kb0 + P.kBlocksPerSeg);
        const int as = it % ACC_STAGES;
        mbar_wait(&accEmpty[as], ((it / ACC_STAGES) & 1) ^ 1);  // epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmemD = tmemBase + (uint32_t)(as * TILE);
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++n) {
          const int s = n % STAGES;
          mbar_wait(&full[s], (n / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(tiles + (size_t)s * STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / UK; ++kk) {
            const uint32_t koff = kk * UK * sizeof(float);  // 32 bytes inside the 128-byte swizzle row
            const uint64_t ahi = umma_desc(base + koff), alo = umma_desc(base + TILE_BYTES + koff);
            const uint64_t bhi = umma_desc(base + 2 * TILE_BYTES + koff), blo = umma_desc(base + 3 * TILE_BYTES + koff);
            umma_tf32(tmemD, alo, bhi, acc);  // small terms first
            umma_tf32(tmemD, ahi, blo, 1);
            umma_tf32(tmemD, ahi, bhi, 1);
            acc = 1;
          }
          umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
        }
        umma_commit(&accFull[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===== epilogue warps: TMEM lanes [32 q, 32 q + 32), q = warp % 4 =====
    const int q = warp & 3;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x, ++it) {
      const int nTiles = P.tilesM * P.tilesN;
      const int tile = item % nTiles, seg = item / nTiles;
      const int tm = tile / P.tilesN, tn = tile % P.tilesN;
      const int as = it % ACC_STAGES;
      mbar_wait(&accFull[as], (it / ACC_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int r = q * 32 + lane;  // tile row = TMEM lane
      const uint32_t taddr = tmemBase + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TILE);
#pragma unroll
      for (int c0 = 0; c0 < TILE; c0 += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (P.nSeg == 1 && P.out) {
          // direct row-major store, ragged rows / columns masked
          const int64_t gm = (int64_t)tm * TILE + r;
          const int gn0 = tn * TILE + c0;
          if (gm < P.rowsA) {
            float *o = P.out + gm * P.ldo + gn0;
            if (gn0 + 32 <= P.rowsB && (P.ldo % 4 == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (gn0 + j < P.rowsB) o[j] = __uint_as_float(v[j]);
            }
          }
        } else {
          float *o = P.ws + (((size_t)tile * P.nSeg + seg) * TILE + r) * TILE + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4 *>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                             __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&accEmpty[as]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "n"(TMEM_COLS));
  }
}

// C(i, j) [column-major, ld ldc, offsets (i0, j0)] = sum over the split-k segments of the partial tiles, in segment
// order (deterministic)
__global__ void tf32_reduce_kernel(const float *__restrict__ ws, int tilesN, int nSeg, int rowsValid, int colsValid,
                                   float *__restrict__ C, int64_t ldc) {
  const int tile = blockIdx.x;
  const int tm = tile / tilesN, tn = tile % tilesN;
  for (int e = threadIdx.x; e < TILE * TILE; e += blockDim.x) {
    const int c = e / TILE, r = e % TILE;  // consecutive threads walk a column of C (contiguous in memory)
    const int gi = tm * TILE + r, gj = tn * TILE + c;
    if (gi >= rowsValid || gj >= colsValid) continue;
    float s = 0.0f;
    for (int k = 0; k < nSeg; ++k) s += ws[(((size_t)tile * nSeg + k) * TILE + r) * TILE + c];
    C[(int64_t)gi + (int64_t)gj * ldc] = s;
  }
}

// hi / lo split of the FP32 rounding of x (see the file header)
__device__ __forceinline__ void split_tf32(double x, float &hi, float &lo) {
  const float f = (float)x;
  hi = __uint_as_float(__float_as_uint(f) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(f - hi) & 0xffffe000u);
}

// Thi / Tlo [c][r] (pitch ldt floats) = split(X[r][c0 + c]) for r < rows, c < ncols: the transposed (K-major) FP32
// copies of a column block of a row-major FP64 matrix; 32 x 32 tiles through shared memory, both sides coalesced
__global__ void split_transpose_kernel(const double *__restrict__ X, int64_t ldx, int c0, int ncols, int64_t rows,
                                       float *__restrict__ Thi, float *__restrict__ Tlo, int64_t ldt) {
  __shared__ float th[32][33], tl[32][33];
  const int64_t nRowTiles = (rows + 31) / 32;
  const int nColTiles = (ncols + 31) / 32;
  for (int64_t t = blockIdx.x; t < nRowTiles * nColTiles; t += gridDim.x) {
    const int64_t r0 = (t / nColTiles) * 32;
    const int cb = (int)(t % nColTiles) * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
      const int64_t r = r0 + rr;
      const int c = cb + threadIdx.x;
      float hi = 0.0f, lo = 0.0f;
      if (r < rows && c < ncols) split_tf32(X[r * ldx + c0 + c], hi, lo);
      th[rr][threadIdx.x] = hi;
      tl[rr][threadIdx.x] = lo;
    }
    __syncthreads();
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
      const int c = cb + cc;
      const int64_t r = r0 + threadIdx.x;
      if (c < ncols && r < rows) {
        Thi[(int64_t)c * ldt + r] = th[threadIdx.x][cc];
        Tlo[(int64_t)c * ldt + r] = tl[threadIdx.x][cc];
      }
    }
    __syncthreads();
  }
}

// Xhi / Xlo [r][c] (pitch ldo) = split(X[r][c]): same orientation (the rotation's A operand)
__global__ void split_rows_kernel(const double *__restrict__ X, int64_t ldx, int ncols, int64_t rows,
                                  float *__restrict__ Xhi, float *__restrict__ Xlo, int64_t ldo) {
  const int64_t total = rows * ldo;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / ldo;
    const int c = (int)(idx % ldo);
    float hi = 0.0f, lo = 0.0f;
    if (c < ncols) split_tf32(X[r * ldx + c], hi, lo);
    Xhi[idx] = hi;
    Xlo[idx] = lo;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// tensor map of a K-major FP32 operand: [rows][K] with row pitch `pitch` floats; box = {BK, TILE}, 128-byte swizzle
int make_operand_map(CUtensorMap *map, const float *base, int64_t rows, int64_t K, int64_t pitch) {
  EncodeTiledFn enc = encode_tiled_fn();
  DB_CHECK(enc, "cuTensorMapEncodeTiled is not available from this driver");
  DB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (pitch * sizeof(float)) % 16 == 0,
           "tf32 GEMM operand: base must be 16-byte aligned and the row pitch a multiple of 16 bytes");
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)(pitch * sizeof(float))};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)TILE};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld operand, pitch %lld", (int)r, (long long)rows,
              (long long)K, (long long)pitch);
    return DFTFE_B200_ERR_CUDA;
  }
  return 0;
}

}  // namespace

int launch_split_transpose(dftfe_b200_ctx *ctx, const double *X, int64_t ldx, int c0, int ncols, int64_t rows,
                           float *Thi, float *Tlo, int64_t ldt) {
  if (rows == 0 || ncols == 0) return 0;
  ctx->launches += 1;
  const int64_t tiles = ((rows + 31) / 32) * ((ncols + 31) / 32);
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)ctx->num_sms * 16);
  split_transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(X, ldx, c0, ncols, rows, Thi, Tlo, ldt);
  DB_CUDA(cudaGetLastError());
  return 0;
}

int launch_split_rows(dftfe_b200_ctx *ctx, const double *X, int64_t ldx, int ncols, int64_t rows, float *Xhi,
                      float *Xlo, int64_t ldo) {
  if (rows == 0) return 0;
  ctx->launches += 1;
  const int grid = (int)std::min<int64_t>((rows * ldo + 255) / 256, (int64_t)ctx->num_sms * 16);
  split_rows_kernel<<<grid, 256, 0, ctx->stream>>>(X, ldx, ncols, rows, Xhi, Xlo, ldo);
  DB_CUDA(cudaGetLastError());
  return 0;
}

// D[m][n] = sum_k A[rowA0 + m][k] B[rowB0 + n][k], m < rowsA, n < rowsB, k < K (operands K-major, split in hi / lo
// TF32-exact parts, row pitches pitchA / pitchB floats; nRowsA / nRowsB = rows the arrays hold).
//   outRowMajor != nullptr : D stored row-major, element (m, n) at outRowMajor[m * ldo + n]   (rotation)
//   else                   : D stored column-major, element (m, n) at outColMajor[m + n * ldc] (projections; split-k)
int launch_tf32x3_gemm(dftfe_b200_ctx *ctx, const float *Ahi, const float *Alo, int64_t nRowsA, int64_t pitchA,
                       int rowA0, int rowsA, const float *Bhi, const float *Blo, int64_t nRowsB, int64_t pitchB,
                       int rowB0, int rowsB, int64_t K, float *outRowMajor, int64_t ldo, float *outColMajor,
                       int64_t ldc) {
  if (rowsA <= 0 || rowsB <= 0 || K <= 0) return 0;
  DB_DYN_SMEM(ctx, tf32x3_gemm_kernel, SMEM_BYTES);
  CUtensorMap mAhi, mAlo, mBhi, mBlo;
  DB_TRY(make_operand_map(&mAhi, Ahi, nRowsA, K, pitchA));
  DB_TRY(make_operand_map(&mAlo, Alo, nRowsA, K, pitchA));
  DB_TRY(make_operand_map(&mBhi, Bhi, nRowsB, K, pitchB));
  DB_TRY(make_operand_map(&mBlo, Blo, nRowsB, K, pitchB));
  Tf32Params P;
  P.tilesM = (rowsA + TILE - 1) / TILE;
  P.tilesN = (rowsB + TILE - 1) / TILE;
  P.kBlocks = (int)((K + BK - 1) / BK);
  P.rowsA = rowsA;
  P.rowsB = rowsB;
  P.rowA0 = rowA0;
  P.rowB0 = rowB0;
  const int nTiles = P.tilesM * P.tilesN;
  int nSeg = 1;
  if (!outRowMajor) {
    // split k so that tiles x segments fills the persistent grid; a segment keeps >= 64 k-blocks
    double bestEff = 0.0;
    for (int s = 1; s <= 64; ++s) {  // (<= 64 with >= 64 k-blocks each: no segment is ever empty)
      if (s > 1 && P.kBlocks / s < 64) break;
      const int64_t items = (int64_t)nTiles * s;
      const int64_t waves = (items + ctx->num_sms - 1) / ctx->num_sms;
      const double eff = (double)items / (double)(waves * ctx->num_sms);
      if (eff > bestEff + 1e-9) {
        bestEff = eff;
        nSeg = s;
      }
    }
  }
  P.nSeg = nSeg;
  P.kBlocksPerSeg = (P.kBlocks + nSeg - 1) / nSeg;
  P.out = outRowMajor;
  P.ldo = ldo;
  P.ws = nullptr;
  if (!outRowMajor) {
    DB_TRY(ctx->tf32Ws.alloc((size_t)nTiles * nSeg * TILE * TILE));
    P.ws = ctx->tf32Ws.p;
  }
  {
    ProfScope ps(ctx, outRowMajor ? "rotation_fp32" : "projection_fp32", outRowMajor ? 1 : 2);
    const int grid = std::min(nTiles * nSeg, ctx->num_sms);
    tf32x3_gemm_kernel<<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(mAhi, mAlo, mBhi, mBlo, P);
    if (!outRowMajor)
      tf32_reduce_kernel<<<nTiles, 256, 0, ctx->stream>>>(ctx->tf32Ws.p, P.tilesN, nSeg, rowsA, rowsB, outColMajor, ldc);
  }
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dftfe_b200
