// Fused cell-level Hamiltonian apply for sm_100a.
//
// Replaces, in ONE kernel per cell colour, the reference's per-degree sequence
//   K1 stridedCopyToBlock gather      (utils/DeviceKernelsGeneric.cc:101-126)
//   K2 cuBLAS gemmStridedBatched       (src/dftOperator/matrixVectorProductImplementationsDevice.cc:51-79)
//   K3 axpyStridedBlockAtomicAdd       (utils/DeviceKernelsGeneric.cc:296-318)
//   K5 stridedBlockScale x4            (src/dftOperator/kohnShamDFTOperatorDevice.cc:3790-3859)
//   K7 combinedDeviceKernel            (src/linAlg/linearAlgebraOperationsDevice.cc:37-64)
// i.e. for every owned cell c and wavefunction column tile:
//   Xc[k,:]   = src[row(c,k), tile]                               (gather through the index map)
//   Yc        = A_c * Xc       A_c[i][k] = rowOut*H_c(i,k)*rowIn   (real: dgemm 'N','N' of the reference)
//                              A_c[i][k] = rowOut*H_c(k,i)*rowIn   (complex: zgemm 'N','T', i.e. H^T)
//   dst[row(c,i), tile] (first touch) = ca*src + cb*dst + s*Yc[i,:]
//                        (otherwise) += s*Yc[i,:]
// Cells of one colour share no row, so the read-modify-write needs no atomics and
// the summation order is fixed (deterministic, unlike the reference's atomicAdd).
//
// Tensor path: Blackwell's tcgen05.mma has no FP64 kind; FP64 tensor work on
// sm_100a is the warp-level DMMA.8x8x4 (all mma.sync f64 shapes lower to it).
// Measured on this pool's B200: 37.0 TFLOP/s raw DMMA issue, 35.7 cuBLAS DGEMM
// (profiles/r01_fp64_peaks.jsonl).
//
// Work decomposition (p=6: n=343): a CTA owns (cell, 32-column tile); the 43
// m8 row tiles are dealt round-robin to 12 warps so each of the SM's four
// tensor pipes gets 11/11/11/10 tiles (97.7 % balance).  A fragments stream
// straight from L2/HBM into registers out of a fragment-major copy of H_c made
// once per set_cell_hamiltonian (every H element is used exactly once per CTA,
// so staging it through shared memory buys nothing); the gathered X tile lives
// in shared memory with a row pitch of 36 doubles (conflict-free B-fragment
// reads).
//
// Complex (k-point) build: vectors are interleaved (re, im), so a tile of 32 real
// columns holds 16 complex ones and  Y = (Ar + i Ai) X  is computed with real DMMAs as
// Y = Ar * X + Ai * X',  X'[k, 2j] = -X[k, 2j+1],  X'[k, 2j+1] = X[k, 2j]:
// the k loop runs over 2*KS "virtual" k-steps, even ones take Ar and the B fragment as
// stored, odd ones take Ai and the partner column with the sign of the real part flipped
// (one integer XOR per fragment).  No split re/im temporaries (the reference needs them
// around its atomics, matrixVectorProductImplementationsDevice.cc:86-108).
#include "common.cuh"

namespace dftfe_b200 {

namespace {

constexpr int BT = 32;       // real wavefunction columns per CTA tile
constexpr int NT = BT / 8;   // n8 tiles per warp
constexpr int LDS = BT + 4;  // shared-memory row pitch in doubles ( = 4 mod 16 )
// row word of the flagged index map: bits 0..29 local row, bit 30 = live (owned and
// unconstrained), bit 31 = first touch (this cell is the first, in colour order, to write the row)
constexpr uint32_t ROW_MASK = 0x3fffffffu, LIVE_BIT = 0x40000000u, FIRST_BIT = 0x80000000u;

// first touch: dst = ca*src + cb*dst + s*acc ; later touches: dst += s*acc
__device__ __forceinline__ void epilogue_coeffs(uint32_t word, const EpilogueParams &ep, double &ca, double &cb) {
  const uint32_t r = word & ROW_MASK;
  const bool live = (word & LIVE_BIT) != 0 || ep.allLive;
  ca = 0.0;
  cb = 1.0;
  if (word & FIRST_BIT) {
    ca = live ? ep.a * (ep.rowA ? __ldg(ep.rowA + r) : 1.0) : 0.0;
    cb = live ? ep.b * (ep.rowB ? __ldg(ep.rowB + r) : 1.0) : 0.0;
  }
}

template <int NODES, bool CPLX = false>
struct CellCfg {
  static constexpr int MT = (NODES + 7) / 8;  // m8 row tiles
  static constexpr int KS = (NODES + 3) / 4;  // k4 steps
  static constexpr int KSV = CPLX ? 2 * KS : KS;  // virtual k-steps (complex: Ar / Ai interleaved)
  static constexpr int KPAD = KS * 4;
#ifndef DB_MMA_WARPS_MAX
#define DB_MMA_WARPS_MAX 12  // A/B hook (tools/build_variants.sh): 11 leaves the CTA at 12 warps incl. the producer
#endif
  static constexpr int WARPS = MT >= 12 ? DB_MMA_WARPS_MAX : (MT >= 8 ? 8 : 4);
  static constexpr int TPW = (MT + WARPS - 1) / WARPS;  // row tiles per warp (max)
  static constexpr int THREADS = WARPS * 32;
  static constexpr size_t SMEM = (size_t)KPAD * LDS * sizeof(double);
  // per-lane fragment vector (padded so that it is loadable with 16/32-byte vector loads)
  static constexpr int TPWP = TPW <= 1 ? 1 : (TPW <= 2 ? 2 : (TPW <= 4 ? 4 : 8));
  static constexpr size_t HT_PER_WARP = (size_t)KSV * 32 * TPWP;      // doubles
  static constexpr size_t HT_PER_CELL = (size_t)WARPS * HT_PER_WARP;  // doubles
};

// one k-step of A fragments for a warp: TPWP consecutive doubles per lane
template <int N>
__device__ __forceinline__ void load_frags(const double *p, double (&a)[N]) {
  if constexpr (N == 1) {
    a[0] = __ldg(p);
  } else if constexpr (N == 2) {
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a[0]), "=d"(a[1]) : "l"(p));
  } else if constexpr (N == 4) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[0]), "=d"(a[1]), "=d"(a[2]), "=d"(a[3]) : "l"(p));
  } else {
    static_assert(N == 8, "unsupported fragment vector");
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[0]), "=d"(a[1]), "=d"(a[2]), "=d"(a[3]) : "l"(p));
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a[4]), "=d"(a[5]), "=d"(a[6]), "=d"(a[7])
                 : "l"(p + 4));
  }
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// B fragments of k-step `ks` for the NT column tiles of a warp.
//   PARTNER = false: xb = tile + (lane&3)*LDS + (lane>>2), as stored
//   PARTNER = true : xb = tile + (lane&3)*LDS + ((lane>>2)^1), the other half of the complex pair, with the
//                    sign bit flipped on lanes that hold a real-part column (sgn mask)
template <bool PARTNER, int NTL = NT>
__device__ __forceinline__ void load_b(const double *xb, unsigned long long sgn, int ks, double (&b)[NTL]) {
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt) {
    const double v = xb[(ks * 4) * LDS + nt * 8];
    b[nt] = PARTNER ? __longlong_as_double(__double_as_longlong(v) ^ (long long)sgn) : v;
  }
}

// H_c (as stored by the reference: mem[c][I][J] = H_c(I,J), complex interleaved for CPLX) -> fragment-major,
// grouped per MMA warp:  Ht[cell][warp][ksv][lane][t] = A_c[(warp + t*WARPS)*8 + lane/4][k*4 + lane%4],
// zero padded, so a lane fetches the A fragments of all its row tiles for one k-step with ONE 32-byte load
// (LDG.E.256) and a warp's stream for a cell is one contiguous run.  The mass scalings of
// H~ = M^-1/2 H M^-1/2 are folded in here, once per set_cell_hamiltonian.
template <int NODES, bool CPLX>
__global__ void retile_H_kernel(const double *__restrict__ H, double *__restrict__ Ht, int64_t nCells,
                                const uint32_t *__restrict__ cellRows, const double *__restrict__ rowIn,
                                const double *__restrict__ rowOut) {
  using C = CellCfg<NODES, CPLX>;
  const int64_t total = nCells * (int64_t)C::HT_PER_CELL;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int t = r % C::TPWP;
    r /= C::TPWP;
    const int lane = r % 32;
    r /= 32;
    const int ksv = r % C::KSV;
    r /= C::KSV;
    const int w = r % C::WARPS;
    const int64_t cell = r / C::WARPS;
    const int ks = CPLX ? (ksv >> 1) : ksv;
    const int mt = w + t * C::WARPS;
    const int i = (t < C::TPW && mt < C::MT) ? mt * 8 + lane / 4 : NODES;
    const int k = ks * 4 + lane % 4;
    double v = 0.0;
    if (i < NODES && k < NODES) {
      if (CPLX)  // A(i,k) = H_c(k,i): the reference's zgemm uses transB = 'T'
        v = H[(cell * (int64_t)NODES * NODES + (int64_t)k * NODES + i) * 2 + (ksv & 1)];
      else
        v = H[cell * (int64_t)NODES * NODES + (int64_t)i * NODES + k];
      const uint32_t ri = cellRows[cell * NODES + i] & ROW_MASK, rk = cellRows[cell * NODES + k] & ROW_MASK;
      v = rowOut[ri] * v * rowIn[rk];
    }
    Ht[idx] = v;
  }
}

// ---------------------------------------------------------------------------
// Generic kernel: one CTA per (cell, column tile); any column count / leading dimension;
// optional extra gather / output scales (the bare HXCheby entry point undoes the folded ones).
// ---------------------------------------------------------------------------
template <int NODES, bool CPLX>
__global__ void __launch_bounds__(CellCfg<NODES, CPLX>::THREADS, 1)
cell_matvec_kernel(const double *__restrict__ Ht, const uint32_t *__restrict__ cellRows,
                   const int32_t *__restrict__ cells, const double *__restrict__ src,
                   double *__restrict__ dst, int ncols, int ldx, int nColTiles, EpilogueParams ep) {
  using C = CellCfg<NODES, CPLX>;
  extern __shared__ __align__(16) double Xs[];  // [KPAD][LDS]
  __shared__ uint32_t rowsS[NODES];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int item = blockIdx.x;
  const int cell = cells[item / nColTiles];
  const int col0 = (item % nColTiles) * BT;

  for (int i = tid; i < NODES; i += C::THREADS) rowsS[i] = cellRows[(size_t)cell * NODES + i];
  __syncthreads();

  // ---- gather: one 256-byte row segment per warp instruction
  const bool colOk = (col0 + lane) < ncols;
#pragma unroll 4
  for (int k = warp; k < C::KPAD; k += C::WARPS) {
    double v = 0.0;
    if (k < NODES && colOk) {
      const uint32_t r = rowsS[k] & ROW_MASK;
      v = __ldg(src + (size_t)r * ldx + col0 + lane);
      if (ep.rowIn) v *= __ldg(ep.rowIn + r);
    }
    Xs[k * LDS + lane] = v;
  }
  __syncthreads();

  // ---- DMMA main loop over pairs of (virtual) k-steps
  double acc[C::TPW][NT][2];
#pragma unroll
  for (int t = 0; t < C::TPW; ++t)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[t][nt][0] = acc[t][nt][1] = 0.0;

  const double *Hc = Ht + (size_t)cell * C::HT_PER_CELL + (size_t)warp * C::HT_PER_WARP + lane * C::TPWP;
  const double *xb = Xs + (lane & 3) * LDS + (lane >> 2);
  const double *xbp = Xs + (lane & 3) * LDS + ((lane >> 2) ^ 1);
  const unsigned long long sgn = ((lane >> 2) & 1) ? 0ull : 0x8000000000000000ull;

  double a0[C::TPWP], a1[C::TPWP];
  load_frags<C::TPWP>(Hc, a0);
  load_frags<C::TPWP>(Hc + (C::KSV > 1 ? 32 * C::TPWP : 0), a1);

  for (int ks = 0; ks < C::KSV; ks += 2) {
    // prefetch A for ks+2, ks+3 (clamped re-reads at the end are harmless)
    double n0[C::TPWP], n1[C::TPWP];
    load_frags<C::TPWP>(Hc + (size_t)min(ks + 2, C::KSV - 1) * 32 * C::TPWP, n0);
    load_frags<C::TPWP>(Hc + (size_t)min(ks + 3, C::KSV - 1) * 32 * C::TPWP, n1);
    {
      double b[NT];
      load_b<false>(xb, sgn, CPLX ? (ks >> 1) : ks, b);  // even virtual step: Ar (or real A), B as stored
#pragma unroll
      for (int t = 0; t < C::TPW; ++t)
        if (warp + t * C::WARPS < C::MT) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a0[t], b[nt]);
        }
    }
    if (ks + 1 < C::KSV) {
      double b[NT];
      if (CPLX)
        load_b<true>(xbp, sgn, ks >> 1, b);  // odd virtual step: Ai with the partner column
      else
        load_b<false>(xb, sgn, ks + 1, b);
#pragma unroll
      for (int t = 0; t < C::TPW; ++t)
        if (warp + t * C::WARPS < C::MT) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a1[t], b[nt]);
        }
    }
#pragma unroll
    for (int t = 0; t < C::TPWP; ++t) {
      a0[t] = n0[t];
      a1[t] = n1[t];
    }
  }

  // ---- epilogue: recurrence + scaling + coloured (atomics-free) assembly
  const bool vec2 = ((ldx & 1) == 0) && (col0 + BT <= ncols);
#pragma unroll
  for (int t = 0; t < C::TPW; ++t) {
    const int mt = warp + t * C::WARPS;
    const int i = mt * 8 + (lane >> 2);
    if (mt < C::MT && i < NODES) {
      const uint32_t fr = rowsS[i];
      const uint32_t r = fr & ROW_MASK;
      const double so = ep.s * (ep.rowOut ? __ldg(ep.rowOut + r) : 1.0);
      double ca, cb;
      epilogue_coeffs(fr, ep, ca, cb);
      double *drow = dst + (size_t)r * ldx + col0 + (lane & 3) * 2;
      const double *srow = src + (size_t)r * ldx + col0 + (lane & 3) * 2;
      if (vec2) {
        double2 d[NT], sv[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          d[nt] = (cb != 0.0) ? *reinterpret_cast<const double2 *>(drow + nt * 8) : make_double2(0.0, 0.0);
          sv[nt] = (ca != 0.0) ? *reinterpret_cast<const double2 *>(srow + nt * 8) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          double2 o;
          o.x = so * acc[t][nt][0] + ca * sv[nt].x + cb * d[nt].x;
          o.y = so * acc[t][nt][1] + ca * sv[nt].y + cb * d[nt].y;
          *reinterpret_cast<double2 *>(drow + nt * 8) = o;
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = col0 + nt * 8 + (lane & 3) * 2 + e;
            if (col < ncols) {
              double o = so * acc[t][nt][e];
              if (ca != 0.0) o += ca * srow[nt * 8 + e];
              if (cb != 0.0) o += cb * drow[nt * 8 + e];
              drow[nt * 8 + e] = o;
            }
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// One wavefunction column (real build): y_c = H~_c x_c is a matrix-vector product bound by the stream of H~_c
// (941 KB per cell at FE order 6).  The Lanczos bounds (linearAlgebraOperationsDevice.cc:340-527: 20 single-vector
// applies per SCF step) come through here.  Same fragment-major H~ as the DMMA kernels: a lane's 32-byte vector of
// k-step ks holds A[(warp + t*WARPS)*8 + lane/4][ks*4 + lane%4] for its TPWP row tiles, so the lane accumulates those
// rows against x[ks*4 + lane%4] and the four lanes of a row are summed by two shuffles.  One CTA per cell of the
// colour, several CTAs per SM, eight 32-byte loads in flight per lane.
// ---------------------------------------------------------------------------
template <int NODES>
__global__ void __launch_bounds__(CellCfg<NODES, false>::THREADS)
cell_gemv_kernel(const double *__restrict__ Ht, const uint32_t *__restrict__ cellRows,
                 const int32_t *__restrict__ cells, const double *__restrict__ src, double *__restrict__ dst, int ldx,
                 EpilogueParams ep) {
  using C = CellCfg<NODES, false>;
  __shared__ double xs[C::KPAD];
  __shared__ uint32_t rowsS[NODES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cell = cells[blockIdx.x];
  for (int i = tid; i < NODES; i += C::THREADS) rowsS[i] = cellRows[(size_t)cell * NODES + i];
  __syncthreads();
  for (int k = tid; k < C::KPAD; k += C::THREADS) {
    double v = 0.0;
    if (k < NODES) {
      const uint32_t r = rowsS[k] & ROW_MASK;
      v = __ldg(src + (size_t)r * ldx);
      if (ep.rowIn) v *= __ldg(ep.rowIn + r);
    }
    xs[k] = v;
  }
  __syncthreads();
  const double *Hc = Ht + (size_t)cell * C::HT_PER_CELL + (size_t)warp * C::HT_PER_WARP + lane * C::TPWP;
  const double *xk = xs + (lane & 3);
  double acc[C::TPWP];
#pragma unroll
  for (int t = 0; t < C::TPWP; ++t) acc[t] = 0.0;
#pragma unroll 8
  for (int ks = 0; ks < C::KS; ++ks) {
    double a[C::TPWP];
    load_frags<C::TPWP>(Hc + (size_t)ks * 32 * C::TPWP, a);
    const double x = xk[ks * 4];
#pragma unroll
    for (int t = 0; t < C::TPWP; ++t) acc[t] = fma(a[t], x, acc[t]);
  }
#pragma unroll
  for (int t = 0; t < C::TPWP; ++t) {
    acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 1);
    acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 2);
  }
  // epilogue (recurrence + scaling + coloured assembly, as in cell_matvec_kernel): lane%4 == t%4 writes row tile t
#pragma unroll
  for (int t = 0; t < C::TPW; ++t) {
    const int mt = warp + t * C::WARPS;
    const int i = mt * 8 + (lane >> 2);
    if ((lane & 3) == (t & 3) && mt < C::MT && i < NODES) {
      const uint32_t fr = rowsS[i];
      const uint32_t r = fr & ROW_MASK;
      const double so = ep.s * (ep.rowOut ? __ldg(ep.rowOut + r) : 1.0);
      double ca, cb;
      epilogue_coeffs(fr, ep, ca, cb);
      double o = so * acc[t];
      if (ca != 0.0) o += ca * src[(size_t)r * ldx];
      if (cb != 0.0) o += cb * dst[(size_t)r * ldx];
      dst[(size_t)r * ldx] = o;
    }
  }
}

// ---------------------------------------------------------------------------
// Fast path: persistent, warp-specialised version of the kernel above.
//   * one CTA per SM loops over the (cell, column-tile) items of a colour;
//   * a producer warp gathers the NEXT item's X tile with 1-D TMA bulk copies
//     (cp.async.bulk, one 256-byte row segment each, completion on an mbarrier)
//     into the other half of a double-buffered shared-memory tile while
//   * twelve MMA warps run the DMMA k-loop of the current item with a 4-deep
//     register prefetch of their A fragments (one LDG.E.256 per k-step) and then do
//     the recurrence / assembly epilogue straight from registers.
// No CTA-wide barrier inside the loop: warps drift, so one warp's epilogue
// overlaps its SMSP neighbours' DMMAs.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
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

template <int NODES, bool CPLX>
struct PersistCfg {
  using C = CellCfg<NODES, CPLX>;
  static constexpr int MMA_WARPS = C::WARPS;
  static constexpr int THREADS = (MMA_WARPS + 1) * 32;  // + one producer warp
  static constexpr size_t XBUF = (size_t)C::KPAD * LDS;  // doubles per buffer
  static constexpr size_t SMEM = 2 * XBUF * sizeof(double) + 2 * NODES * sizeof(uint32_t) + 4 * sizeof(uint64_t);
#ifndef DB_APF
#define DB_APF 4
#endif
// what-if switches for profiling (results are wrong when set): skip the A-fragment refills / the epilogue's global
// traffic to measure what each costs (profiles/r01_cell_kernel_variants.txt)
#ifndef DB_WHATIF_NOA
#define DB_WHATIF_NOA 0
#endif
#ifndef DB_WHATIF_NOEPI
#define DB_WHATIF_NOEPI 0
#endif
  static constexpr int APF = DB_APF;  // A-fragment prefetch depth in (virtual) k-steps
};

// Item loop of an MMA warp for blocks whose column count is a multiple of 32 (every tile full): kept as ONE
// function - splitting it into per-item pieces as the ragged variant below does costs registers (64 B of spills,
// -5 % on the headline configuration).  The row-tile count NTILE (TPW or TPW-1) is a compile-time constant per
// instantiation because a predicated-off DMMA still occupies its tensor-pipe slot.
template <int NODES, bool CPLX, int NTILE>
__device__ __forceinline__ void mma_warp_items_full(const double *__restrict__ Ht, const int32_t *__restrict__ cells,
                                               int nItems, const double *__restrict__ src,
                                               double *__restrict__ dst, int ldx, int nColTiles,
                                               const EpilogueParams &ep, const double *Xs, const uint32_t *rowsS,
                                               uint64_t *full, uint64_t *empty, int warp, int lane) {
  using C = CellCfg<NODES, CPLX>;
  using P = PersistCfg<NODES, CPLX>;
  static_assert(P::APF % 2 == 0, "the parity of a virtual k-step must be that of its ring slot");
  const double *xb0 = Xs + (lane & 3) * LDS + (lane >> 2);
  const double *xbp0 = Xs + (lane & 3) * LDS + ((lane >> 2) ^ 1);
  const unsigned long long sgn = ((lane >> 2) & 1) ? 0ull : 0x8000000000000000ull;
  // A prefetch ring; primed for the first item here, re-primed for the next item before each epilogue
  double a[P::APF][C::TPWP];
  auto prime = [&](int item) {
    const int cell = cells[item / nColTiles];
    const double *Hc = Ht + (size_t)cell * C::HT_PER_CELL + (size_t)warp * C::HT_PER_WARP + lane * C::TPWP;
#pragma unroll
    for (int s = 0; s < P::APF; ++s) load_frags<C::TPWP>(Hc + (size_t)min(s, C::KSV - 1) * 32 * C::TPWP, a[s]);
  };
  if ((int)blockIdx.x < nItems) prime(blockIdx.x);
  int it = 0;
  for (int item = blockIdx.x; item < nItems; item += gridDim.x, ++it) {
    const int buf = it & 1;
    const uint32_t ph = (it >> 1) & 1;
    const int cell = cells[item / nColTiles];
    const int col0 = (item % nColTiles) * BT;
    const double *Hc = Ht + (size_t)cell * C::HT_PER_CELL + (size_t)warp * C::HT_PER_WARP + lane * C::TPWP;
    const double *xb = xb0 + buf * P::XBUF;
    const double *xbp = xbp0 + buf * P::XBUF;

    double acc[NTILE][NT][2];
#pragma unroll
    for (int t = 0; t < NTILE; ++t)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[t][nt][0] = acc[t][nt][1] = 0.0;

    mbar_wait(&full[buf], ph);

    int ks = 0;  // virtual k-step; advances by APF (even), so the parity of ks+s is that of s
    for (; ks + P::APF <= C::KSV; ks += P::APF) {
#pragma unroll
      for (int s = 0; s < P::APF; ++s) {
        double b[NT];
        if (CPLX && (s & 1))
          load_b<true>(xbp, sgn, (ks + s) >> 1, b);
        else
          load_b<false>(xb, sgn, CPLX ? ((ks + s) >> 1) : (ks + s), b);
#pragma unroll
        for (int t = 0; t < NTILE; ++t)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
#if !DB_WHATIF_NOA
        if (ks + s + P::APF < C::KSV) load_frags<C::TPWP>(Hc + (size_t)(ks + s + P::APF) * 32 * C::TPWP, a[s]);
#endif
      }
    }
#pragma unroll
    for (int s = 0; s < C::KSV % P::APF; ++s) {
      double b[NT];
      if (CPLX && (s & 1))
        load_b<true>(xbp, sgn, (ks + s) >> 1, b);
      else
        load_b<false>(xb, sgn, CPLX ? ((ks + s) >> 1) : (ks + s), b);
#pragma unroll
      for (int t = 0; t < NTILE; ++t)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
    }

    // rows of this warp's tiles, then release the buffer to the producer
    uint32_t fr[NTILE];
#pragma unroll
    for (int t = 0; t < NTILE; ++t) {
      const int i = (warp + t * C::WARPS) * 8 + (lane >> 2);
      fr[t] = (i < NODES) ? rowsS[buf * NODES + i] : 0u;
    }
    // next item's first A fragments fly while this item's epilogue runs
    if (item + (int)gridDim.x < nItems) prime(item + gridDim.x);

    // ---- epilogue (full tiles, even ldx: guaranteed by the launcher)
#if DB_WHATIF_NOEPI
    {
      double sacc = 0.0;
#pragma unroll
      for (int t = 0; t < NTILE; ++t)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) sacc += acc[t][nt][0] + acc[t][nt][1];
      if (sacc == 1234.5678) dst[0] = sacc;
    }
#else
#pragma unroll
    for (int t = 0; t < NTILE; ++t) {
      const int i = (warp + t * C::WARPS) * 8 + (lane >> 2);
      if (i < NODES) {
        const uint32_t r = fr[t] & ROW_MASK;
        const double so = ep.s * (ep.rowOut ? __ldg(ep.rowOut + r) : 1.0);
        double ca, cb;
        epilogue_coeffs(fr[t], ep, ca, cb);
        double *drow = dst + (size_t)r * ldx + col0 + (lane & 3) * 2;
        // src[row(c,i), tile] is row i of the X tile still sitting in shared memory (the buffer is released
        // after the epilogue): the a*src term of a first touch costs no global read
        const double *srow = Xs + buf * P::XBUF + i * LDS + (lane & 3) * 2;
        double2 d[NT], sv[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          d[nt] = (cb != 0.0) ? *reinterpret_cast<const double2 *>(drow + nt * 8) : make_double2(0.0, 0.0);
          sv[nt] = (ca != 0.0) ? *reinterpret_cast<const double2 *>(srow + nt * 8) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          double2 o;
          o.x = so * acc[t][nt][0] + ca * sv[nt].x + cb * d[nt].x;
          o.y = so * acc[t][nt][1] + ca * sv[nt].y + cb * d[nt].y;
          *reinterpret_cast<double2 *>(drow + nt * 8) = o;
        }
      }
    }
#endif
    // release the X tile to the producer (its rows were the epilogue's src operand)
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[buf]);
  }
}

// One (cell, column tile) item of an MMA warp.  The row-tile count NTILE (TPW or TPW-1) and the number of n8
// column tiles NTL (4 for a full 32-column tile, fewer for the ragged last tile of a block whose column count is
// not a multiple of 32 - the reference's AUTO block sizes are 90..130 and 180..220, src/dft/dft.cc:554-690) are
// compile-time constants per instantiation because a predicated-off DMMA still occupies its tensor-pipe slot.
template <int NODES, bool CPLX, int NTILE, int NTL, bool RAGGED>
__device__ __forceinline__ void mma_one_item(const double *__restrict__ Ht, const int32_t *__restrict__ cells,
                                             int nItems, int item, int it, int col0,
                                             const double *__restrict__ src,
                                             double *__restrict__ dst, int ncols, int ldx, int nColTiles,
                                             const EpilogueParams &ep, const double *Xs, const uint32_t *rowsS,
                                             uint64_t *full, uint64_t *empty, int warp, int lane,
                                             double (&a)[PersistCfg<NODES, CPLX>::APF][CellCfg<NODES, CPLX>::TPWP]) {
  using C = CellCfg<NODES, CPLX>;
  using P = PersistCfg<NODES, CPLX>;
  const int buf = it & 1;
  const uint32_t ph = (it >> 1) & 1;
  const int cell = cells[item / nColTiles];
  const double *Hc = Ht + (size_t)cell * C::HT_PER_CELL + (size_t)warp * C::HT_PER_WARP + lane * C::TPWP;
  const double *xb = Xs + (lane & 3) * LDS + (lane >> 2) + buf * P::XBUF;
  const double *xbp = Xs + (lane & 3) * LDS + ((lane >> 2) ^ 1) + buf * P::XBUF;
  const unsigned long long sgn = ((lane >> 2) & 1) ? 0ull : 0x8000000000000000ull;

  double acc[NTILE][NTL][2];
#pragma unroll
  for (int t = 0; t < NTILE; ++t)
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) acc[t][nt][0] = acc[t][nt][1] = 0.0;

  mbar_wait(&full[buf], ph);

  int ks = 0;  // virtual k-step; advances by APF (even), so the parity of ks+s is that of s
  for (; ks + P::APF <= C::KSV; ks += P::APF) {
#pragma unroll
    for (int s = 0; s < P::APF; ++s) {
      double b[NTL];
      if (CPLX && (s & 1))
        load_b<true, NTL>(xbp, sgn, (ks + s) >> 1, b);
      else
        load_b<false, NTL>(xb, sgn, CPLX ? ((ks + s) >> 1) : (ks + s), b);
#pragma unroll
      for (int t = 0; t < NTILE; ++t)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
      if (ks + s + P::APF < C::KSV) load_frags<C::TPWP>(Hc + (size_t)(ks + s + P::APF) * 32 * C::TPWP, a[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < C::KSV % P::APF; ++s) {
    double b[NTL];
    if (CPLX && (s & 1))
      load_b<true, NTL>(xbp, sgn, (ks + s) >> 1, b);
    else
      load_b<false, NTL>(xb, sgn, CPLX ? ((ks + s) >> 1) : (ks + s), b);
#pragma unroll
    for (int t = 0; t < NTILE; ++t)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
  }

  // rows of this warp's tiles, then release the buffer to the producer
  uint32_t fr[NTILE];
#pragma unroll
  for (int t = 0; t < NTILE; ++t) {
    const int i = (warp + t * C::WARPS) * 8 + (lane >> 2);
    fr[t] = (i < NODES) ? rowsS[buf * NODES + i] : 0u;
  }
  // next item's first A fragments fly while this item's epilogue runs
  if (item + (int)gridDim.x < nItems) {
    const int ncell = cells[(item + gridDim.x) / nColTiles];
    const double *Hn = Ht + (size_t)ncell * C::HT_PER_CELL + (size_t)warp * C::HT_PER_WARP + lane * C::TPWP;
#pragma unroll
    for (int s = 0; s < P::APF; ++s) load_frags<C::TPWP>(Hn + (size_t)min(s, C::KSV - 1) * 32 * C::TPWP, a[s]);
  }

  // ---- epilogue (even ncols / ldx: guaranteed by the launcher; a ragged tile masks its missing column pairs)
#pragma unroll
  for (int t = 0; t < NTILE; ++t) {
    const int i = (warp + t * C::WARPS) * 8 + (lane >> 2);
    if (i < NODES) {
      const uint32_t r = fr[t] & ROW_MASK;
      const double so = ep.s * (ep.rowOut ? __ldg(ep.rowOut + r) : 1.0);
      double ca, cb;
      epilogue_coeffs(fr[t], ep, ca, cb);
      const int cl = col0 + (lane & 3) * 2;
      double *drow = dst + (size_t)r * ldx + cl;
      const double *srow = Xs + buf * P::XBUF + i * LDS + (lane & 3) * 2;  // src rows = the X tile (see above)
      double2 d[NTL], sv[NTL];
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const bool ok = !RAGGED || (cl + nt * 8 < ncols);
        d[nt] = (cb != 0.0 && ok) ? *reinterpret_cast<const double2 *>(drow + nt * 8) : make_double2(0.0, 0.0);
        sv[nt] = (ca != 0.0 && ok) ? *reinterpret_cast<const double2 *>(srow + nt * 8) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        double2 o;
        o.x = so * acc[t][nt][0] + ca * sv[nt].x + cb * d[nt].x;
        o.y = so * acc[t][nt][1] + ca * sv[nt].y + cb * d[nt].y;
        if (!RAGGED || (cl + nt * 8 < ncols)) *reinterpret_cast<double2 *>(drow + nt * 8) = o;
      }
    }
  }
  // release the X tile to the producer (its rows were the epilogue's src operand)
  __syncwarp();
  if (lane == 0) mbar_arrive(&empty[buf]);
}

// Column tiling of a block of `ncols` columns: nColTiles = ceil(ncols / 32) tiles; the ceil(ncols / 8) n8 tiles
// are dealt as evenly as possible (e.g. 100 columns -> 4,3,3,3), because every item streams the whole H_c and a
// narrow tile would make that stream the bottleneck.
struct ColTiling {
  int base, rem;  // tiles [0, rem) hold base+1 n8 tiles, the others base
  __host__ __device__ ColTiling(int ncols, int nColTiles) {
    const int total8 = (ncols + 7) >> 3;
    base = total8 / nColTiles;
    rem = total8 % nColTiles;
  }
  __host__ __device__ int ntl(int tile) const { return base + (tile < rem ? 1 : 0); }
  __host__ __device__ int col0(int tile) const { return 8 * (tile * base + (tile < rem ? tile : rem)); }
};

template <int NODES, bool CPLX, int NTILE, bool RAGGED>
__device__ __forceinline__ void mma_warp_items(const double *__restrict__ Ht, const int32_t *__restrict__ cells,
                                               int nItems, const double *__restrict__ src,
                                               double *__restrict__ dst, int ncols, int ldx, int nColTiles,
                                               const EpilogueParams &ep, const double *Xs, const uint32_t *rowsS,
                                               uint64_t *full, uint64_t *empty, int warp, int lane) {
  using C = CellCfg<NODES, CPLX>;
  using P = PersistCfg<NODES, CPLX>;
  static_assert(P::APF % 2 == 0, "the parity of a virtual k-step must be that of its ring slot");
  // A prefetch ring; primed for the first item here, re-primed for the next item before each epilogue
  double a[P::APF][C::TPWP];
  if ((int)blockIdx.x < nItems) {
    const int cell = cells[blockIdx.x / nColTiles];
    const double *Hc = Ht + (size_t)cell * C::HT_PER_CELL + (size_t)warp * C::HT_PER_WARP + lane * C::TPWP;
#pragma unroll
    for (int s = 0; s < P::APF; ++s) load_frags<C::TPWP>(Hc + (size_t)min(s, C::KSV - 1) * 32 * C::TPWP, a[s]);
  }
  const ColTiling ct(ncols, nColTiles);
  int it = 0;
  for (int item = blockIdx.x; item < nItems; item += gridDim.x, ++it) {
#define DB_ITEM(NTL, COL0)                                                                                          \
  mma_one_item<NODES, CPLX, NTILE, NTL, RAGGED>(Ht, cells, nItems, item, it, COL0, src, dst, ncols, ldx, nColTiles, ep, Xs, \
                                        rowsS, full, empty, warp, lane, a)
    if (!RAGGED) {
      DB_ITEM(4, (item % nColTiles) * BT);
    } else {
      const int tile = item % nColTiles;
      const int ntl = ct.ntl(tile), c0 = ct.col0(tile);
      if (ntl == 4)
        DB_ITEM(4, c0);
      else if (ntl == 3)
        DB_ITEM(3, c0);
      else if (ntl == 2)
        DB_ITEM(2, c0);
      else
        DB_ITEM(1, c0);
    }
#undef DB_ITEM
  }
}

// (13 warps: SMSP 0 holds four of them, so the register file allows 16384 / (4 x 32) = 128 registers per thread;
// __maxnreg__ 144 / 152 compiles without spills but fails to launch.)
template <int NODES, bool CPLX, bool RAGGED>
__global__ void __launch_bounds__(PersistCfg<NODES, CPLX>::THREADS, 1)
cell_matvec_persistent_kernel(const double *__restrict__ Ht, const uint32_t *__restrict__ cellRows,
                              const int32_t *__restrict__ cells, int nItems, const double *__restrict__ src,
                              double *__restrict__ dst, int ncols, int ldx, int nColTiles, EpilogueParams ep) {
  using C = CellCfg<NODES, CPLX>;
  using P = PersistCfg<NODES, CPLX>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *Xs = reinterpret_cast<double *>(smem_raw);                                  // [2][KPAD][LDS]
  uint32_t *rowsS = reinterpret_cast<uint32_t *>(smem_raw + 2 * P::XBUF * sizeof(double));  // [2][NODES]
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + 2 * P::XBUF * sizeof(double) +
                                                ((2 * NODES * sizeof(uint32_t) + 7) / 8) * 8);
  uint64_t *full = bars, *empty = bars + 2;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  // physical warp 0 is the producer; MMA warp ids 0..11 sit on physical warps 1..12, which puts the
  // three lightest MMA warps (ids 3, 7, 11: 4+3+3 row tiles) on the producer's scheduler (SMSP 0)
  const int pwarp = tid >> 5;
  const int warp = pwarp - 1;

  // zero both tiles once (pad rows k >= NODES and pad columns stay zero forever)
  for (int i = tid; i < (int)(2 * P::XBUF); i += P::THREADS) Xs[i] = 0.0;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    // only warps that own row tiles run the item loop and release the buffers (FE order 1: one of four)
    mbar_init(&empty[0], C::MT < P::MMA_WARPS ? C::MT : P::MMA_WARPS);
    mbar_init(&empty[1], C::MT < P::MMA_WARPS ? C::MT : P::MMA_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // generic-proxy zero fill must be ordered before the async-proxy (TMA) writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (pwarp == 0) {
    // ===== producer warp: gather rows of the next item through the index map =====
    int it = 0;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int cell = cells[item / nColTiles];
      const ColTiling ct(ncols, nColTiles);
      const int tile = item % nColTiles;
      const int col0 = RAGGED ? ct.col0(tile) : tile * BT;
      const int width = RAGGED ? min(8 * ct.ntl(tile), ncols - col0) : BT;
      const uint32_t *cr = cellRows + (size_t)cell * NODES;
      // row words first (global latency overlaps the wait for the buffer)
      uint32_t myRows[(NODES + 31) / 32];
#pragma unroll
      for (int j = 0; j < (NODES + 31) / 32; ++j) {
        const int k = lane + 32 * j;
        myRows[j] = (k < NODES) ? __ldg(cr + k) : 0u;
      }
      mbar_wait(&empty[buf], ph ^ 1);
      uint32_t *rs = rowsS + buf * NODES;
      double *xs = Xs + buf * P::XBUF;
#pragma unroll
      for (int j = 0; j < (NODES + 31) / 32; ++j) {
        const int k = lane + 32 * j;
        if (k < NODES) rs[k] = myRows[j];
      }
      __syncwarp();
      // a ragged last tile copies only the columns that exist (even count -> multiple of 16 bytes); the
      // shared-memory columns beyond them keep finite values of earlier tiles and are masked in the epilogue
      const uint32_t rowBytes = (uint32_t)(width * sizeof(double));
      if (lane == 0) mbar_arrive_expect_tx(&full[buf], (uint32_t)NODES * rowBytes);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < (NODES + 31) / 32; ++j) {
        const int k = lane + 32 * j;
        if (k < NODES)
          tma_bulk_g2s(xs + k * LDS, src + (size_t)(myRows[j] & ROW_MASK) * ldx + col0, rowBytes, &full[buf]);
      }
    }
  } else {
    constexpr int FULL_WARPS = C::MT - (C::TPW - 1) * C::WARPS;  // warps that own TPW tiles
    if (!RAGGED) {
      if (warp < FULL_WARPS)
        mma_warp_items_full<NODES, CPLX, C::TPW>(Ht, cells, nItems, src, dst, ldx, nColTiles, ep, Xs, rowsS, full,
                                                 empty, warp, lane);
      else if (C::TPW > 1)
        mma_warp_items_full<NODES, CPLX, (C::TPW > 1 ? C::TPW - 1 : 1)>(Ht, cells, nItems, src, dst, ldx, nColTiles,
                                                                        ep, Xs, rowsS, full, empty, warp, lane);
    } else {
      if (warp < FULL_WARPS)
        mma_warp_items<NODES, CPLX, C::TPW, true>(Ht, cells, nItems, src, dst, ncols, ldx, nColTiles, ep, Xs, rowsS,
                                                  full, empty, warp, lane);
      else if (C::TPW > 1)
        mma_warp_items<NODES, CPLX, (C::TPW > 1 ? C::TPW - 1 : 1), true>(Ht, cells, nItems, src, dst, ncols, ldx,
                                                                         nColTiles, ep, Xs, rowsS, full, empty, warp,
                                                                         lane);
    }
  }
}

// rows that no owned cell touches still need the first-touch formula (contrib = 0)
__global__ void orphan_first_touch_kernel(const uint32_t *__restrict__ rows, int64_t nRows,
                                          const double *__restrict__ src, double *__restrict__ dst, int ncols,
                                          int ldx, EpilogueParams ep) {
  const int64_t total = nRows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w = rows[idx / ncols] | FIRST_BIT;
    const uint32_t r = w & ROW_MASK;
    const int col = idx % ncols;
    double ca, cb;
    epilogue_coeffs(w, ep, ca, cb);
    double o = 0.0;
    if (ca != 0.0) o += ca * src[(size_t)r * ldx + col];
    if (cb != 0.0) o += cb * dst[(size_t)r * ldx + col];
    dst[(size_t)r * ldx + col] = o;
  }
}

template <int NODES, bool CPLX>
int launch_impl2(dftfe_b200_ctx *ctx, const double *src, double *dst, int ncols, int ldx,
                 const EpilogueParams &ep) {
  using C = CellCfg<NODES, CPLX>;
  using P = PersistCfg<NODES, CPLX>;
  DB_DYN_SMEM(ctx, (cell_matvec_kernel<NODES, CPLX>), C::SMEM);
  if (P::SMEM <= 227 * 1024) {
    DB_DYN_SMEM(ctx, (cell_matvec_persistent_kernel<NODES, CPLX, false>), P::SMEM);
    DB_DYN_SMEM(ctx, (cell_matvec_persistent_kernel<NODES, CPLX, true>), P::SMEM);
  }
  const int nColTiles = (ncols + BT - 1) / BT;
  // fast path needs an even column count (16-byte row segments and column pairs), 16-byte aligned rows and no
  // extra gather scale; the last column tile may be ragged
  const bool fast = (P::SMEM <= 227 * 1024) && (ncols % 2 == 0) && (ldx % 2 == 0) && ep.rowIn == nullptr &&
                    ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                    !ctx->force_generic_cell_kernel;
  for (int k = 0; k < ctx->nColours; ++k) {
    const int nCellsK = ctx->colourStart_h[k + 1] - ctx->colourStart_h[k];
    if (nCellsK == 0) continue;
    ProfScope ps(ctx, "cell_matvec");
    const int nItems = nCellsK * nColTiles;
    if constexpr (!CPLX) {
      if (ncols == 1 && !ctx->force_generic_cell_kernel) {  // single vector: H-stream-bound matrix-vector kernel
        cell_gemv_kernel<NODES><<<nCellsK, C::THREADS, 0, ctx->stream>>>(
            ctx->Hactive, ctx->cellRowsFlagged.p, ctx->colourCells.p + ctx->colourStart_h[k], src, dst, ldx, ep);
        continue;
      }
    }
    if (fast) {
      const int grid = std::min(nItems, std::max(1, ctx->num_sms - ctx->reserved_sms));
      if (ncols % BT == 0)
        cell_matvec_persistent_kernel<NODES, CPLX, false><<<grid, P::THREADS, P::SMEM, ctx->stream>>>(
            ctx->Hactive, ctx->cellRowsFlagged.p, ctx->colourCells.p + ctx->colourStart_h[k], nItems, src, dst, ncols,
            ldx, nColTiles, ep);
      else
        cell_matvec_persistent_kernel<NODES, CPLX, true><<<grid, P::THREADS, P::SMEM, ctx->stream>>>(
            ctx->Hactive, ctx->cellRowsFlagged.p, ctx->colourCells.p + ctx->colourStart_h[k], nItems, src, dst, ncols,
            ldx, nColTiles, ep);
    } else {
      cell_matvec_kernel<NODES, CPLX><<<nItems, C::THREADS, C::SMEM, ctx->stream>>>(
          ctx->Hactive, ctx->cellRowsFlagged.p, ctx->colourCells.p + ctx->colourStart_h[k], src, dst, ncols, ldx,
          nColTiles, ep);
    }
  }
  DB_CUDA(cudaGetLastError());
  return 0;
}

template <int NODES>
int launch_impl(dftfe_b200_ctx *ctx, const double *src, double *dst, int ncols, int ldx,
                const EpilogueParams &ep) {
  return ctx->cplx ? launch_impl2<NODES, true>(ctx, src, dst, ncols, ldx, ep)
                   : launch_impl2<NODES, false>(ctx, src, dst, ncols, ldx, ep);
}

template <int NODES>
int retile_impl(dftfe_b200_ctx *ctx, const double *H_d) {
  const size_t perCell = ctx->cplx ? CellCfg<NODES, true>::HT_PER_CELL : CellCfg<NODES, false>::HT_PER_CELL;
  dftfe_b200::DevBuf<double> &Ht = ctx->Hsets[ctx->activeK];
  DB_TRY(Ht.alloc((size_t)ctx->nC * perCell));
  ctx->Hactive = Ht.p;
  ctx->launches += 1;
  if (ctx->cplx)
    retile_H_kernel<NODES, true><<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(
        H_d, Ht.p, ctx->nC, ctx->cellRowsFlagged.p, ctx->rowIn.p, ctx->rowOut.p);
  else
    retile_H_kernel<NODES, false><<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(
        H_d, Ht.p, ctx->nC, ctx->cellRowsFlagged.p, ctx->rowIn.p, ctx->rowOut.p);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

#define DB_DISPATCH_NODES(n, FN, ...)                                  \
  switch (n) {                                                         \
    case 8: return FN<8>(__VA_ARGS__);                                 \
    case 27: return FN<27>(__VA_ARGS__);                               \
    case 64: return FN<64>(__VA_ARGS__);                               \
    case 125: return FN<125>(__VA_ARGS__);                             \
    case 216: return FN<216>(__VA_ARGS__);                             \
    case 343: return FN<343>(__VA_ARGS__);                             \
    case 512: return FN<512>(__VA_ARGS__);                             \
    default:                                                           \
      set_error("no cell kernel instantiated for %d nodes per cell", n); \
      return DFTFE_B200_ERR_UNSUPPORTED;                               \
  }

int cell_kernel_supported(int n) {
  return n == 8 || n == 27 || n == 64 || n == 125 || n == 216 || n == 343 || n == 512;
}

// H_d: real n x n doubles per cell, or complex interleaved (2 n x n doubles) for a complex context
int retile_cell_hamiltonian(dftfe_b200_ctx *ctx, const double *H_d) {
  DB_CHECK(ctx->have_map && ctx->have_mass,
           "set_cell_hamiltonian needs set_index_map, set_constraints and set_mass first (the M^-1/2 scalings are "
           "folded into the re-tiled cell matrices)");
  DB_DISPATCH_NODES(ctx->n, retile_impl, ctx, H_d);
}

// ncols / ldx in REAL columns (a complex context passes 2 x its complex column count)
int launch_cell_matvec(dftfe_b200_ctx *ctx, const double *src, double *dst, int ncols, int ldx,
                       const EpilogueParams &ep) {
  DB_CHECK(ctx->have_map && ctx->have_H, "cell matvec needs set_index_map and set_cell_hamiltonian first");
  DB_CHECK(src != dst, "cell matvec: src and dst must not alias");
  DB_DISPATCH_NODES(ctx->n, launch_impl, ctx, src, dst, ncols, ldx, ep);
}

int launch_orphan_first_touch(dftfe_b200_ctx *ctx, const double *src, double *dst, int ncols, int ldx,
                              const EpilogueParams &ep) {
  if (ctx->nOrphan == 0) return 0;
  ProfScope ps(ctx, "orphan_rows");
  const int64_t total = ctx->nOrphan * ncols;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, ctx->num_sms * 16);
  orphan_first_touch_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->orphanRows.p, ctx->nOrphan, src, dst, ncols, ldx,
                                                          ep);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dftfe_b200
