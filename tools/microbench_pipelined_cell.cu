// Round-2 design probe (NOT part of the product): the fused cell kernel restructured so that the tensor pipe never
// waits on the epilogue (profiles/r01_cell_kernel_stall_regions.csv: an MMA warp spends 22 % of its time there and
// the three warps of an SMSP get there together).
//
//   * 12 warps per CTA, no dedicated producer warp -> 3 warps per SMSP, register ceiling 168 instead of 128
//     (warp 11, which owns the fewest row tiles, issues the TMA gather of the next item);
//   * every warp splits its row tiles in two halves (2 + 2, or 2 + 1) and software-pipelines them: the dst rows of
//     one half are loaded while the other half's DMMAs run, the read-modify-write completes after that k-loop;
//   * the A fragments are laid out per (warp, half) so that a lane fetches 16 contiguous bytes per k-step.
//
// Synthetic problem of the bench size (4913 cells, n = 343, 256 columns, disjoint rows per cell, every row a live
// first touch).  Prints TF/s and checks a few rows against a host computation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o microbench_pipelined_cell microbench_pipelined_cell.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);       \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

constexpr int NODES = 343, KS = 86, KPAD = 344, MT = 43, WARPS = 12;
constexpr int BT = 32, NT = 4, LDS = 36;
constexpr int THREADS = WARPS * 32;
constexpr size_t XBUF = (size_t)KPAD * LDS;
constexpr size_t SMEM = 2 * XBUF * sizeof(double) + 4 * sizeof(uint64_t);
#ifndef PROBE_APF
#define PROBE_APF 6
#endif
#ifndef PROBE_NOEPI
#define PROBE_NOEPI 0  // 1: skip the epilogue's global traffic (ceiling of the two-half k-loop structure; wrong results)
#endif
constexpr int APF = PROBE_APF;  // A prefetch depth (k-steps), 2 doubles per lane per k-step
constexpr size_t H2_PER_HALF = (size_t)KS * 32 * 2;           // doubles
constexpr size_t H2_PER_CELL = (size_t)WARPS * 2 * H2_PER_HALF;

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
__device__ __forceinline__ void load2(const double *p, double (&a)[2]) {
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a[0]), "=d"(a[1]) : "l"(p));
}

// H (mem[c][I][J]) -> H2[cell][warp][half][ks][lane][u] = H_c[(warp + (2*half+u)*12)*8 + lane/4][ks*4 + lane%4]
__global__ void retile2(const double *__restrict__ H, double *__restrict__ H2, int64_t nCells) {
  const int64_t total = nCells * (int64_t)H2_PER_CELL;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int u = r % 2;
    r /= 2;
    const int lane = r % 32;
    r /= 32;
    const int ks = r % KS;
    r /= KS;
    const int half = r % 2;
    r /= 2;
    const int w = r % WARPS;
    const int64_t cell = r / WARPS;
    const int mt = w + (2 * half + u) * WARPS;
    const int i = mt * 8 + lane / 4, k = ks * 4 + lane % 4;
    H2[idx] = (mt < MT && i < NODES && k < NODES) ? H[cell * (int64_t)NODES * NODES + (int64_t)i * NODES + k] : 0.0;
  }
}

struct Ep {
  double a, b, s;
};

// k-loop of one half: NH row tiles (2 or 1) x 4 column tiles
// the ring must have been primed (prime()) for this half before the call: its first loads are issued as soon as
// the previous k-loop ends, so they fly under the epilogue work in between
__device__ __forceinline__ void prime(const double *__restrict__ Ah, double (&a)[APF][2]) {
#pragma unroll
  for (int s = 0; s < APF; ++s) load2(Ah + (size_t)s * 64, a[s]);
}

template <int NH>
__device__ __forceinline__ void kloop(const double *__restrict__ Ah, const double *xb, double (&acc)[2][NT][2],
                                      double (&a)[APF][2]) {
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[t][nt][0] = acc[t][nt][1] = 0.0;
  int ks = 0;
  for (; ks + APF <= KS; ks += APF) {
#pragma unroll
    for (int s = 0; s < APF; ++s) {
      double b[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b[nt] = xb[(ks + s) * 4 * LDS + nt * 8];
#pragma unroll
      for (int t = 0; t < NH; ++t)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
      if (ks + s + APF < KS) load2(Ah + (size_t)(ks + s + APF) * 64, a[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < KS % APF; ++s) {
    double b[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b[nt] = xb[(ks + s) * 4 * LDS + nt * 8];
#pragma unroll
    for (int t = 0; t < NH; ++t)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(acc[t][nt][0], acc[t][nt][1], a[s][t], b[nt]);
  }
}

// issue the dst loads of one half (NH tiles): d[t][nt]
template <int NH>
__device__ __forceinline__ void issue_dst_loads(const double *dst, int ldx, int col0, const uint32_t (&rows)[2], int lane,
                                                double2 (&d)[2][NT]) {
#if PROBE_NOEPI
  return;
#endif
#pragma unroll
  for (int t = 0; t < NH; ++t) {
    const double *drow = dst + (size_t)rows[t] * ldx + col0 + (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)  // volatile asm: keeps the loads where they are written, ahead of the next k-loop
      asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(d[t][nt].x), "=d"(d[t][nt].y) : "l"(drow + nt * 8) : "memory");
  }
}

template <int NH>
__device__ __forceinline__ void complete_epilogue(double *dst, int ldx, int col0, const uint32_t (&rows)[2],
                                                  const int (&irow)[2], const double *Xbuf, int lane, const Ep &ep,
                                                  const double (&acc)[2][NT][2], const double2 (&d)[2][NT]) {
#if PROBE_NOEPI
  {
    double sacc = 0.0;
#pragma unroll
    for (int t = 0; t < NH; ++t)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) sacc += acc[t][nt][0] + acc[t][nt][1];
    if (sacc == 1234.5678) dst[0] = sacc;
    return;
  }
#endif
#pragma unroll
  for (int t = 0; t < NH; ++t) {
    if (irow[t] < NODES) {
      double *drow = dst + (size_t)rows[t] * ldx + col0 + (lane & 3) * 2;
      const double *srow = Xbuf + irow[t] * LDS + (lane & 3) * 2;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 sv = *reinterpret_cast<const double2 *>(srow + nt * 8);
        double2 o;
        o.x = ep.s * acc[t][nt][0] + ep.a * sv.x + ep.b * d[t][nt].x;
        o.y = ep.s * acc[t][nt][1] + ep.a * sv.y + ep.b * d[t][nt].y;
        *reinterpret_cast<double2 *>(drow + nt * 8) = o;
      }
    }
  }
}

template <int NH1>  // tiles in the second half: 2 (warps 0..6) or 1 (warps 7..11)
__device__ __forceinline__ void warp_items(const double *__restrict__ H2, const uint32_t *__restrict__ cellRows,
                                           int nItems, int nColTiles, const double *__restrict__ src,
                                           double *__restrict__ dst, int ldx, Ep ep, double *Xs, uint64_t *full,
                                           uint64_t *empty, int warp, int lane) {
  const bool issuer = (warp == WARPS - 1);
  // gather of item `item` into buffer `buf` (issuer warp only)
  auto gather = [&](int item, int buf) {
    const int cell = item / nColTiles, col0 = (item % nColTiles) * BT;
    if (lane == 0) mbar_arrive_expect_tx(&full[buf], (uint32_t)(NODES * BT * sizeof(double)));
    __syncwarp();
    for (int k = lane; k < NODES; k += 32) {
      const uint32_t r = __ldg(cellRows + (size_t)cell * NODES + k);
      tma_bulk_g2s(Xs + buf * XBUF + k * LDS, src + (size_t)r * ldx + col0, BT * sizeof(double), &full[buf]);
    }
  };
  if (issuer) {
    if ((int)blockIdx.x < nItems) gather(blockIdx.x, 0);
    if ((int)(blockIdx.x + gridDim.x) < nItems) gather(blockIdx.x + gridDim.x, 1);
  }
  double acc0[2][NT][2], acc1[2][NT][2];
  double2 d0[2][NT], d1[2][NT];
  double ring[APF][2];
  if ((int)blockIdx.x < nItems)
    prime(H2 + (size_t)(blockIdx.x / nColTiles) * H2_PER_CELL + (size_t)warp * 2 * H2_PER_HALF + lane * 2, ring);
  uint32_t rows0[2] = {0, 0}, rows1[2] = {0, 0}, prows1[2] = {0, 0};
  int irow0[2], irow1[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    irow0[t] = (warp + t * WARPS) * 8 + (lane >> 2);
    irow1[t] = (warp + (2 + t) * WARPS) * 8 + (lane >> 2);
  }
  int pcol0 = 0, pbuf = 0;
  bool pending1 = false;
  int it = 0;
  for (int item = blockIdx.x; item < nItems; item += gridDim.x, ++it) {
    const int buf = it & 1;
    const uint32_t ph = (it >> 1) & 1;
    const int cell = item / nColTiles, col0 = (item % nColTiles) * BT;
    const double *Aw = H2 + (size_t)cell * H2_PER_CELL + (size_t)warp * 2 * H2_PER_HALF + lane * 2;
    const double *xb = Xs + buf * XBUF + (lane & 3) * LDS + (lane >> 2);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      rows0[t] = irow0[t] < NODES ? __ldg(cellRows + (size_t)cell * NODES + irow0[t]) : 0u;
      rows1[t] = (t < NH1 && irow1[t] < NODES) ? __ldg(cellRows + (size_t)cell * NODES + irow1[t]) : 0u;
    }
    mbar_wait(&full[buf], ph);
    // ---- half 0 (d1 of the previous item is in flight)
    kloop<2>(Aw, xb, acc0, ring);
    prime(Aw + H2_PER_HALF, ring);  // half 1's first fragments fly under the epilogue work below
    if (pending1) {
      complete_epilogue<NH1>(dst, ldx, pcol0, prows1, irow1, Xs + pbuf * XBUF, lane, ep, acc1, d1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[pbuf]);  // previous item's X tile is free
      if (issuer) {
        // the buffer of the previous item takes the item after this one
        const int nitem = item + gridDim.x;
        if (nitem < nItems) {
          mbar_wait(&empty[pbuf], ((it - 1) >> 1) & 1);
          gather(nitem, pbuf);
        }
      }
    }
    issue_dst_loads<2>(dst, ldx, col0, rows0, lane, d0);
    // ---- half 1 (d0 in flight)
    kloop<NH1>(Aw + H2_PER_HALF, xb, acc1, ring);
    if (item + (int)gridDim.x < nItems)
      prime(H2 + (size_t)((item + gridDim.x) / nColTiles) * H2_PER_CELL + (size_t)warp * 2 * H2_PER_HALF + lane * 2, ring);
    complete_epilogue<2>(dst, ldx, col0, rows0, irow0, Xs + buf * XBUF, lane, ep, acc0, d0);
    issue_dst_loads<NH1>(dst, ldx, col0, rows1, lane, d1);
    prows1[0] = rows1[0];
    prows1[1] = rows1[1];
    pcol0 = col0;
    pbuf = buf;
    pending1 = true;
  }
  if (pending1) complete_epilogue<NH1>(dst, ldx, pcol0, prows1, irow1, Xs + pbuf * XBUF, lane, ep, acc1, d1);
}

__global__ void __launch_bounds__(THREADS, 1)
cell_pipelined_kernel(const double *__restrict__ H2, const uint32_t *__restrict__ cellRows, int nItems, int nColTiles,
                      const double *__restrict__ src, double *__restrict__ dst, int ldx, Ep ep) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *Xs = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + 2 * XBUF * sizeof(double));
  uint64_t *empty = full + 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (int)(2 * XBUF); i += THREADS) Xs[i] = 0.0;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&empty[0], WARPS);
    mbar_init(&empty[1], WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp < MT - 3 * WARPS)  // warps 0..6 own four row tiles
    warp_items<2>(H2, cellRows, nItems, nColTiles, src, dst, ldx, ep, Xs, full, empty, warp, lane);
  else
    warp_items<1>(H2, cellRows, nItems, nColTiles, src, dst, ldx, ep, Xs, full, empty, warp, lane);
}

int main(int argc, char **argv) {
  const int nCells = argc > 1 ? atoi(argv[1]) : 1480;
  const int B = 256, nColTiles = B / BT;
  const int64_t M = (int64_t)nCells * NODES;  // disjoint rows per cell
  printf("cells %d, rows %lld, B %d, smem %zu B\n", nCells, (long long)M, B, SMEM);
  std::vector<double> hH((size_t)nCells * NODES * NODES);
  for (size_t i = 0; i < hH.size(); ++i) hH[i] = (double)((i * 2654435761u) % 2001) / 1000.0 - 1.0;
  std::vector<uint32_t> hRows((size_t)nCells * NODES);
  for (int c = 0; c < nCells; ++c)
    for (int i = 0; i < NODES; ++i) hRows[(size_t)c * NODES + i] = (uint32_t)((size_t)c * NODES + (i * 97) % NODES);
  std::vector<double> hX((size_t)M * B), hD((size_t)M * B);
  for (size_t i = 0; i < hX.size(); ++i) {
    hX[i] = (double)((i * 40503u) % 1999) / 1000.0 - 1.0;
    hD[i] = (double)((i * 69069u) % 1777) / 1000.0 - 0.9;
  }
  double *H, *H2, *X, *D;
  uint32_t *R;
  CK(cudaMalloc(&H, hH.size() * 8));
  CK(cudaMalloc(&H2, (size_t)nCells * H2_PER_CELL * 8));
  CK(cudaMalloc(&X, hX.size() * 8));
  CK(cudaMalloc(&D, hD.size() * 8));
  CK(cudaMalloc(&R, hRows.size() * 4));
  CK(cudaMemcpy(H, hH.data(), hH.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(X, hX.data(), hX.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(D, hD.data(), hD.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(R, hRows.data(), hRows.size() * 4, cudaMemcpyHostToDevice));
  retile2<<<148 * 8, 256>>>(H, H2, nCells);
  CK(cudaDeviceSynchronize());
  CK(cudaFree(H));
  CK(cudaFuncSetAttribute(cell_pipelined_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  const Ep ep{-0.3, 0.7, 0.5};
  const int nItems = nCells * nColTiles;
  cell_pipelined_kernel<<<148, THREADS, SMEM>>>(H2, R, nItems, nColTiles, X, D, B, ep);
  CK(cudaDeviceSynchronize());
  // check a few rows of cell 0 / cell nCells-1 against the host
  std::vector<double> out((size_t)M * B);
  CK(cudaMemcpy(out.data(), D, out.size() * 8, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int c : {0, nCells / 2, nCells - 1})
    for (int i : {0, 7, 100, 191, 342})
      for (int col : {0, 31, 32, 129, 255}) {
        const size_t r = hRows[(size_t)c * NODES + i];
        double s = 0;
        for (int k = 0; k < NODES; ++k)
          s += hH[((size_t)c * NODES + i) * NODES + k] * hX[(size_t)hRows[(size_t)c * NODES + k] * B + col];
        const double ref = ep.s * s + ep.a * hX[r * B + col] + ep.b * hD[r * B + col];
        maxerr = fmax(maxerr, fabs(ref - out[r * B + col]));
        maxref = fmax(maxref, fabs(ref));
      }
  printf("check: max abs err %.3e (max |ref| %.3e)\n", maxerr, maxref);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) cell_pipelined_kernel<<<148, THREADS, SMEM>>>(H2, R, nItems, nColTiles, X, D, B, ep);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * NODES * NODES * B * nCells;
  printf("{\"probe\": \"pipelined_cell_kernel\", \"cells\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"apf\": %d}\n", nCells,
         ms / reps, flops / (ms / reps * 1e-3) / 1e12, APF);
  return 0;
}
