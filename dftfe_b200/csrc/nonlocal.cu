// Non-local (separable pseudopotential) part of H:  y += C V C^T x.
//
// Reference: computeNonLocalHamiltonianTimesX (device:
// src/dftOperator/computeNonLocalHamiltonianTimesXMemoryOptBatchGEMMDevice.cc:27-283; CPU twin
// computeNonLocalHamiltonianTimesXMemoryOpt.cc:266-505): per non-local cell a batched GEMM C_c^T X_c, a GEMM with a
// 0/1 matrix as a segmented sum, three permutation kernels, accumulate/update of a distributed projector
// vector, a per-cell batched GEMM C_c (V C^T X) and ONE kernel launch per atom to add the result into the cell
// scratch (K17-K21 in SURVEY.md 2.4).
//
// Here the per-cell blocks are assembled once (set_nonlocal) into the equivalent row form
//     Chat[row, (a,p)] = sum_{cells c of a containing row} C_c[i(row), p]
// so that  sum_c C_c^T X_c = Chat^T x  exactly when X_c is the gather of x.  Per operator apply:
//   1. nl_project_kernel : proj[(a,p), :] = sum_{rows of a} Chat[row,(a,p)] * in(row) * x[row, :]   (one CTA per
//      atom and 64-column chunk, row groups reduced in a fixed order)
//   2. all-reduce of proj over the ranks (atoms whose support spans several ranks)
//   3. nl_apply_kernel   : y[row, :] += s * out(row) * sum_e Chat[row,e] V[e] proj[e, :]           (row-parallel)
// Both are coalesced along the wavefunction index, atomics-free and deterministic; the projector block
// (totalProj x B doubles) stays L2 resident between 1 and 3.  For even column counts both steps run in
// bandwidth-shaped ATOM-parallel kernels (16-byte accesses, four rows in flight per lane): the apply keeps the
// atom's V-weighted projector rows in registers and streams its y rows once - atoms that share a row carry different
// colours (greedy colouring at set_nonlocal) and each colour is its own launch, in a fixed order.
//
// Complex (k-point) build: C carries the Bloch phase, one set per k-point; vectors are interleaved (re, im) and
//   proj = Chat^H x  (zgemm with d_nonLocalProjectorElementMatricesConjugate, ...MemoryOpt.cc:98-112),
//   y   += Chat V proj  (zgemm with ...MatricesTranspose, :230-246),
// computed on the interleaved real columns: a lane that holds the real (imaginary) part of a column also reads
// its partner lane's value of the same row - the same 16-byte segment, so no extra memory traffic.
#include <map>

#include "common.cuh"

namespace dftfe_b200 {

namespace {

constexpr int NL_MAXP = 32;   // projectors per atom held in registers
constexpr int NL_COLS = 64;   // columns per CTA
constexpr int NL_RG = 2;      // row groups per CTA (static smem: RG*32*65*8 B)

// cm = 1: vals[r][p] real.  cm = 2: vals[r][p] = (re, im); column `col` is the real part of a complex column when
// even and the imaginary part when odd:  (conj(c) x)_re = cr xr + ci xi,  (conj(c) x)_im = cr xi - ci xr.
template <int CM>
__global__ void __launch_bounds__(NL_COLS *NL_RG)
nl_project_kernel(const double *__restrict__ x, int ncols, int ldx, const int32_t *__restrict__ atomRowStart,
                  const uint32_t *__restrict__ atomRows, const int64_t *__restrict__ atomValStart,
                  const double *__restrict__ vals, const int32_t *__restrict__ projOffset,
                  const double *__restrict__ rowScale, double *__restrict__ proj, size_t sliceStride) {
  __shared__ double red[NL_RG][NL_MAXP][NL_COLS + 1];
  proj += blockIdx.z * sliceStride;  // row slice z of every atom -> its own partial block (summed in slice order)
  const int a = blockIdx.x;
  const int col = blockIdx.y * NL_COLS + threadIdx.x;
  const int rg = threadIdx.y;
  const int P = projOffset[a + 1] - projOffset[a];
  const int r0 = atomRowStart[a], r1 = atomRowStart[a + 1];
  const double *va = vals + atomValStart[a] * CM;
  const double sgn = (col & 1) ? -1.0 : 1.0;
  double acc[NL_MAXP];
#pragma unroll
  for (int p = 0; p < NL_MAXP; ++p) acc[p] = 0.0;
  if (col < ncols) {
    for (int r = r0 + blockIdx.z * NL_RG + rg; r < r1; r += NL_RG * gridDim.z) {
      const uint32_t row = atomRows[r];
      double xv = x[(size_t)row * ldx + col];
      double xp = CM == 2 ? x[(size_t)row * ldx + (col ^ 1)] : 0.0;
      if (rowScale) {
        const double sc = rowScale[row];
        xv *= sc;
        xp *= sc;
      }
      const double *v = va + (size_t)(r - r0) * P * CM;
#pragma unroll
      for (int p = 0; p < NL_MAXP; ++p)
        if (p < P) {
          if (CM == 2)
            acc[p] += v[2 * p] * xv + sgn * v[2 * p + 1] * xp;
          else
            acc[p] += v[p] * xv;
        }
    }
  }
#pragma unroll
  for (int p = 0; p < NL_MAXP; ++p)
    if (p < P) red[rg][p][threadIdx.x] = acc[p];
  __syncthreads();
  if (col < ncols)
    for (int p = rg; p < P; p += NL_RG) {
      double s = 0.0;
#pragma unroll
      for (int g = 0; g < NL_RG; ++g) s += red[g][p][threadIdx.x];
      proj[(size_t)(projOffset[a] + p) * ncols + col] = s;
    }
}

// (c q)_re = cr qr - ci qi,  (c q)_im = cr qi + ci qr
template <int CM>
__global__ void nl_apply_kernel(double *__restrict__ y, int ncols, int ldx, int64_t nRows,
                                const uint32_t *__restrict__ rows, const int64_t *__restrict__ rowStart,
                                const int32_t *__restrict__ entProj, const double *__restrict__ entVal,
                                const double *__restrict__ V, const double *__restrict__ proj,
                                const double *__restrict__ rowScale, double s) {
  const int64_t total = nRows * ncols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / ncols;
    const int c = idx % ncols;
    const uint32_t row = rows[i];
    const double sgn = (c & 1) ? 1.0 : -1.0;
    double sum = 0.0;
    for (int64_t e = rowStart[i]; e < rowStart[i + 1]; ++e) {
      const int id = entProj[e];
      if (CM == 2)
        sum += V[id] * (entVal[2 * e] * proj[(size_t)id * ncols + c] +
                        sgn * entVal[2 * e + 1] * proj[(size_t)id * ncols + (c ^ 1)]);
      else
        sum += entVal[e] * V[id] * proj[(size_t)id * ncols + c];
    }
    const double f = rowScale ? s * rowScale[row] : s;
    y[(size_t)row * ldx + c] += f * sum;
  }
}


// ---- bandwidth-shaped variants (even column counts, 16-byte aligned rows) ---------------------------------------
// One CTA per (atom, 64-column chunk[, row slice]), four warps; a lane owns two adjacent real columns (one complex
// column in the k-point build); a warp walks the atom's rows in groups of four (16-byte loads, four rows in flight).
// Nothing on the per-group path waits for a second memory latency: the row ids and the projector values of the NEXT
// group are fetched while this group is processed - the values as one coalesced warp load (the rows of an atom are
// consecutive in `vals`), handed to the other lanes through a per-warp shared-memory slot.  (ncu before this: every row
// stalled on its own broadcast load of the projector values, 1.5-1.8 TB/s; profiles/r02_nonlocal_ncu_summary*.csv.)
constexpr int NLV_WARPS = 4;
#ifndef DB_NLV_UNROLL
#define DB_NLV_UNROLL 4  // A/B hook: rows in flight per warp
#endif
constexpr int NLV_UNROLL = DB_NLV_UNROLL;

template <int CM, int PMAX>
struct NlStage {
  static constexpr int VPI = NLV_UNROLL * PMAX * CM;  // projector values of one group of rows (at most)
  static constexpr int VPL = (VPI + 31) / 32;         // ... per lane
};

// values of the rows [rb, rb + NLV_UNROLL) of the atom: lane l takes elements l, l + 32, ... of the contiguous run
template <int CM, int PMAX>
__device__ __forceinline__ void nl_fetch_vals(const double *__restrict__ va, int rb, int r0, int r1, int P, int lane,
                                              double (&v)[NlStage<CM, PMAX>::VPL]) {
  const int64_t base = (int64_t)(rb - r0) * P * CM;
  const int64_t avail = (int64_t)(r1 - rb) * P * CM;  // <= 0 beyond the atom's last row
  const int n = NLV_UNROLL * P * CM;
#pragma unroll
  for (int k = 0; k < NlStage<CM, PMAX>::VPL; ++k) {
    const int e = lane + 32 * k;
    v[k] = (e < n && e < avail) ? va[base + e] : 0.0;
  }
}

template <int CM, int PMAX>
__global__ void __launch_bounds__(NLV_WARPS * 32)
nl_project_vec_kernel(const double *__restrict__ x, int ncols, int ldx, const int32_t *__restrict__ atomRowStart,
                      const uint32_t *__restrict__ atomRows, const int64_t *__restrict__ atomValStart,
                      const double *__restrict__ vals, const int32_t *__restrict__ projOffset,
                      const double *__restrict__ rowScale, double *__restrict__ proj, size_t sliceStride) {
  using S = NlStage<CM, PMAX>;
  __shared__ double2 red[NLV_WARPS][PMAX][32];
  __shared__ double vsh[NLV_WARPS][S::VPI];
  proj += blockIdx.z * sliceStride;
  const int a = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.y * 64 + lane * 2;
  const bool active = col < ncols;
  const int P = projOffset[a + 1] - projOffset[a];
  const int r0 = atomRowStart[a], r1 = atomRowStart[a + 1];
  const double *va = vals + atomValStart[a] * CM;
  double2 acc[PMAX];
#pragma unroll
  for (int p = 0; p < PMAX; ++p) acc[p] = make_double2(0.0, 0.0);
  const int stride = gridDim.z * NLV_WARPS * NLV_UNROLL;
  int rb = r0 + (blockIdx.z * NLV_WARPS + warp) * NLV_UNROLL;
  uint32_t nextRows[NLV_UNROLL];
  double vnext[S::VPL];
#pragma unroll
  for (int u = 0; u < NLV_UNROLL; ++u) nextRows[u] = (rb + u < r1) ? atomRows[rb + u] : 0u;
  nl_fetch_vals<CM, PMAX>(va, rb, r0, r1, P, lane, vnext);
  for (; rb < r1; rb += stride) {  // warp-uniform trip count: every lane stages values, inactive lanes skip the data
#pragma unroll
    for (int k = 0; k < S::VPL; ++k)
      if (lane + 32 * k < S::VPI) vsh[warp][lane + 32 * k] = vnext[k];
    __syncwarp();
    double2 xv[NLV_UNROLL];
    double sc[NLV_UNROLL];
#pragma unroll
    for (int u = 0; u < NLV_UNROLL; ++u) {
      xv[u] = make_double2(0.0, 0.0);
      sc[u] = 1.0;
      if (active && rb + u < r1) {
        const uint32_t row = nextRows[u];
        xv[u] = *reinterpret_cast<const double2 *>(x + (size_t)row * ldx + col);
        if (rowScale) sc[u] = rowScale[row];
      }
    }
#pragma unroll
    for (int u = 0; u < NLV_UNROLL; ++u) nextRows[u] = (rb + stride + u < r1) ? atomRows[rb + stride + u] : 0u;
    nl_fetch_vals<CM, PMAX>(va, rb + stride, r0, r1, P, lane, vnext);
#pragma unroll
    for (int u = 0; u < NLV_UNROLL; ++u) {
      if (rb + u < r1) {
        const double xr = xv[u].x * sc[u], xi = xv[u].y * sc[u];
        const double *v = vsh[warp] + u * P * CM;
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
          if (p < P) {
            if (CM == 2) {  // conj(c) x: re = cr xr + ci xi, im = cr xi - ci xr
              const double cr = v[2 * p], ci = v[2 * p + 1];
              acc[p].x += cr * xr + ci * xi;
              acc[p].y += cr * xi - ci * xr;
            } else {
              acc[p].x += v[p] * xr;
              acc[p].y += v[p] * xi;
            }
          }
      }
    }
    __syncwarp();  // the slot is rewritten at the top of the next iteration
  }
#pragma unroll
  for (int p = 0; p < PMAX; ++p)
    if (p < P) red[warp][p][lane] = acc[p];
  __syncthreads();
  if (active)
    for (int p = warp; p < P; p += NLV_WARPS) {
      double2 sum = red[0][p][lane];
#pragma unroll
      for (int g = 1; g < NLV_WARPS; ++g) {
        sum.x += red[g][p][lane].x;
        sum.y += red[g][p][lane].y;
      }
      *reinterpret_cast<double2 *>(proj + (size_t)(projOffset[a] + p) * ncols + col) = sum;
    }
}

// proj[i] = sum over the row slices z, in slice order, of part[z][i]
__global__ void nl_sum_slices_kernel(const double *__restrict__ part, size_t stride, int slices,
                                     double *__restrict__ proj, size_t count) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    double s = part[i];
    for (int z = 1; z < slices; ++z) s += part[(size_t)z * stride + i];
    proj[i] = s;
  }
}

// y[row, :] += s * out(row) * sum_p C_a[row][p] V[a,p] proj[(a,p), :] for the atoms of ONE colour (atoms of a colour
// share no row, so the read-modify-write needs no atomics; the colours run in a fixed order -> deterministic).
// The V-weighted projector rows of the atom sit in registers; y rows stream through once (16-byte accesses).
template <int CM, int PMAX>
__global__ void __launch_bounds__(NLV_WARPS * 32)
nl_apply_vec_kernel(double *__restrict__ y, int ncols, int ldx, const int32_t *__restrict__ colourAtoms,
                    const int32_t *__restrict__ atomRowStart, const uint32_t *__restrict__ atomRows,
                    const int64_t *__restrict__ atomValStart, const double *__restrict__ vals,
                    const int32_t *__restrict__ projOffset, const double *__restrict__ V,
                    const double *__restrict__ proj, const double *__restrict__ rowScale, double s) {
  using S = NlStage<CM, PMAX>;
  __shared__ double vsh[NLV_WARPS][S::VPI];
  const int a = colourAtoms[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.y * 64 + lane * 2;
  const bool active = col < ncols;
  const int P = projOffset[a + 1] - projOffset[a];
  const int r0 = atomRowStart[a], r1 = atomRowStart[a + 1];
  const double *va = vals + atomValStart[a] * CM;
  double2 q[PMAX];  // V * proj of this atom, this lane's column pair
#pragma unroll
  for (int p = 0; p < PMAX; ++p) {
    q[p] = make_double2(0.0, 0.0);
    if (p < P && active) {
      const double2 t = *reinterpret_cast<const double2 *>(proj + (size_t)(projOffset[a] + p) * ncols + col);
      const double v = V[projOffset[a] + p];
      q[p] = make_double2(v * t.x, v * t.y);
    }
  }
  // the atom's rows are dealt over gridDim.z CTAs (row slices): a colour holds few atoms, and one CTA per (atom,
  // column chunk) would leave most SMs without work
  const int stride = gridDim.z * NLV_WARPS * NLV_UNROLL;
  int rb = r0 + (blockIdx.z * NLV_WARPS + warp) * NLV_UNROLL;
  uint32_t nextRows[NLV_UNROLL];
  double vnext[S::VPL];
#pragma unroll
  for (int u = 0; u < NLV_UNROLL; ++u) nextRows[u] = (rb + u < r1) ? atomRows[rb + u] : 0u;
  nl_fetch_vals<CM, PMAX>(va, rb, r0, r1, P, lane, vnext);
  for (; rb < r1; rb += stride) {
#pragma unroll
    for (int k = 0; k < S::VPL; ++k)
      if (lane + 32 * k < S::VPI) vsh[warp][lane + 32 * k] = vnext[k];
    __syncwarp();
    double2 yv[NLV_UNROLL];
    uint32_t rows[NLV_UNROLL];
#pragma unroll
    for (int u = 0; u < NLV_UNROLL; ++u) {
      rows[u] = nextRows[u];
      yv[u] = make_double2(0.0, 0.0);
      if (active && rb + u < r1) yv[u] = *reinterpret_cast<const double2 *>(y + (size_t)rows[u] * ldx + col);
    }
#pragma unroll
    for (int u = 0; u < NLV_UNROLL; ++u) nextRows[u] = (rb + stride + u < r1) ? atomRows[rb + stride + u] : 0u;
    nl_fetch_vals<CM, PMAX>(va, rb + stride, r0, r1, P, lane, vnext);
#pragma unroll
    for (int u = 0; u < NLV_UNROLL; ++u) {
      if (active && rb + u < r1) {
        const double *v = vsh[warp] + u * P * CM;
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
          if (p < P) {
            if (CM == 2) {  // c q: re = cr qr - ci qi, im = cr qi + ci qr
              const double cr = v[2 * p], ci = v[2 * p + 1];
              sr += cr * q[p].x - ci * q[p].y;
              si += cr * q[p].y + ci * q[p].x;
            } else {
              sr += v[p] * q[p].x;
              si += v[p] * q[p].y;
            }
          }
        const double f = rowScale ? s * rowScale[rows[u]] : s;
        yv[u].x += f * sr;
        yv[u].y += f * si;
        *reinterpret_cast<double2 *>(y + (size_t)rows[u] * ldx + col) = yv[u];
      }
    }
    __syncwarp();
  }
}

}  // namespace

// C: [nEntries][n][pMax] doubles (real build) or (re, im) pairs (complex build)
int nonlocal_setup(dftfe_b200_ctx *ctx, int kpt, int32_t nAtoms, const int32_t *nProj, const double *V,
                   int64_t nEntries, const int32_t *entryCell, const int32_t *entryAtom, const double *C, int32_t pMax) {
  DB_CHECK(ctx->have_map, "set_nonlocal: set_index_map first");
  const int n = ctx->n, cm = ctx->cm;
  std::vector<int32_t> off(nAtoms + 1, 0);
  for (int a = 0; a < nAtoms; ++a) {
    DB_CHECK(nProj[a] >= 0 && nProj[a] <= NL_MAXP && nProj[a] <= pMax,
             "set_nonlocal: atom %d has %d projectors (supported: <= %d and <= p_max)", a, nProj[a], NL_MAXP);
    off[a + 1] = off[a] + nProj[a];
  }
  const int totalProj = off[nAtoms];
  // assemble Chat per atom: row -> P (complex) values
  std::vector<std::map<uint32_t, std::vector<double>>> perAtom(nAtoms);
  for (int64_t e = 0; e < nEntries; ++e) {
    const int a = entryAtom[e];
    const int64_t c = entryCell[e];
    DB_CHECK(a >= 0 && a < nAtoms && c >= 0 && c < ctx->nC, "set_nonlocal: entry %lld out of range", (long long)e);
    const int P = nProj[a];
    for (int i = 0; i < n; ++i) {
      const uint32_t row = ctx->cellRows_h[c * n + i];
      auto &v = perAtom[a][row];
      if (v.empty()) v.assign((size_t)P * cm, 0.0);
      const double *src = C + ((size_t)e * n + i) * pMax * cm;
      for (int p = 0; p < P * cm; ++p) v[p] += src[p];
    }
  }
  std::vector<int32_t> atomRowStart(nAtoms + 1, 0);
  std::vector<int64_t> atomValStart(nAtoms + 1, 0);
  std::vector<uint32_t> atomRows;
  std::vector<double> vals;
  std::map<uint32_t, std::vector<std::pair<int32_t, std::pair<double, double>>>> byRow;
  for (int a = 0; a < nAtoms; ++a) {
    const int P = nProj[a];
    for (auto &kv : perAtom[a]) {
      bool nz = false;
      for (double x : kv.second) nz = nz || (x != 0.0);
      if (!nz) continue;
      atomRows.push_back(kv.first);
      vals.insert(vals.end(), kv.second.begin(), kv.second.end());
      for (int p = 0; p < P; ++p) {
        const double re = kv.second[(size_t)p * cm], im = cm == 2 ? kv.second[(size_t)p * cm + 1] : 0.0;
        if (re != 0.0 || im != 0.0) byRow[kv.first].push_back({off[a] + p, {re, im}});
      }
    }
    atomRowStart[a + 1] = (int32_t)atomRows.size();
    atomValStart[a + 1] = (int64_t)vals.size() / cm;
  }
  std::vector<uint32_t> rows;
  std::vector<int64_t> rowStart(1, 0);
  std::vector<int32_t> entProj;
  std::vector<double> entVal;
  for (auto &kv : byRow) {
    rows.push_back(kv.first);
    for (auto &pr : kv.second) {
      entProj.push_back(pr.first);
      entVal.push_back(pr.second.first);
      if (cm == 2) entVal.push_back(pr.second.second);
    }
    rowStart.push_back((int64_t)entProj.size());
  }
  // atoms that share a row get different colours (greedy, atom order): the atom-parallel apply kernel runs one
  // launch per colour and needs no atomics
  std::vector<int32_t> atomColour(nAtoms, -1), colourStart, colourAtoms;
  int nAtomColours = 0;
  {
    std::map<uint32_t, std::vector<int32_t>> atomsOfRow;
    for (int a = 0; a < nAtoms; ++a)
      for (int r = atomRowStart[a]; r < atomRowStart[a + 1]; ++r) atomsOfRow[atomRows[r]].push_back(a);
    std::vector<std::vector<int32_t>> nb(nAtoms);
    for (auto &kv : atomsOfRow)
      for (int32_t a : kv.second)
        for (int32_t b : kv.second)
          if (a != b) nb[a].push_back(b);
    for (int a = 0; a < nAtoms; ++a) {
      std::vector<char> used(nAtomColours + 1, 0);
      for (int32_t b : nb[a])
        if (atomColour[b] >= 0) used[atomColour[b]] = 1;
      int c = 0;
      while (c < nAtomColours && used[c]) ++c;
      atomColour[a] = c;
      nAtomColours = std::max(nAtomColours, c + 1);
    }
    colourStart.assign(nAtomColours + 1, 0);
    for (int a = 0; a < nAtoms; ++a)
      if (atomRowStart[a + 1] > atomRowStart[a]) colourStart[atomColour[a] + 1]++;
    for (int c = 0; c < nAtomColours; ++c) colourStart[c + 1] += colourStart[c];
    colourAtoms.assign(colourStart[nAtomColours], 0);
    std::vector<int32_t> fill(colourStart.begin(), colourStart.end() - 1);
    for (int a = 0; a < nAtoms; ++a)
      if (atomRowStart[a + 1] > atomRowStart[a]) colourAtoms[fill[atomColour[a]]++] = a;
  }
  dftfe_b200_ctx::NonlocalSet &ns = ctx->nlSets[kpt];
  ns.nAtomColours = nAtomColours;
  ns.maxProj = 0;
  ns.maxAtomRows = 0;
  for (int a = 0; a < nAtoms; ++a) {
    ns.maxProj = std::max(ns.maxProj, (int)nProj[a]);
    ns.maxAtomRows = std::max(ns.maxAtomRows, atomRowStart[a + 1] - atomRowStart[a]);
  }
  ns.colourStart_h = colourStart;
  DB_TRY(ns.colourAtoms.upload(colourAtoms.data(), colourAtoms.size(), ctx->stream));
  ns.nAtoms = nAtoms;
  ns.totalProj = totalProj;
  ns.nRows = (int64_t)rows.size();
  DB_TRY(ns.projOffset.upload(off.data(), off.size(), ctx->stream));
  DB_TRY(ns.V.upload(V, totalProj, ctx->stream));
  DB_TRY(ns.atomRowStart.upload(atomRowStart.data(), atomRowStart.size(), ctx->stream));
  DB_TRY(ns.atomValStart.upload(atomValStart.data(), atomValStart.size(), ctx->stream));
  DB_TRY(ns.atomRows.upload(atomRows.data(), atomRows.size(), ctx->stream));
  DB_TRY(ns.vals.upload(vals.data(), vals.size(), ctx->stream));
  DB_TRY(ns.rowList.upload(rows.data(), rows.size(), ctx->stream));
  DB_TRY(ns.rowStart.upload(rowStart.data(), rowStart.size(), ctx->stream));
  DB_TRY(ns.entProj.upload(entProj.data(), entProj.size(), ctx->stream));
  DB_TRY(ns.entVal.upload(entVal.data(), entVal.size(), ctx->stream));
  for (int l = 0; l < 2; ++l) DB_TRY(ctx->nlProj[l].alloc((size_t)std::max(totalProj, 1) * ctx->B * cm));
  ctx->nl = totalProj > 0 ? &ns : nullptr;
  return 0;
}

// proj = Chat^H (in o x), all-reduced over ranks
int nonlocal_project(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, const double *rowScaleIn) {
  if (!ctx->nl || ctx->skip_nonlocal) return 0;
  const dftfe_b200_ctx::NonlocalSet &ns = *ctx->nl;
  const bool vec = (ncols % 2 == 0) && (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   !ctx->force_scalar_row_kernels && ns.maxProj <= 16;
  // few atoms (small systems, many ranks): the rows of an atom are dealt over `slices` CTAs whose partial blocks are
  // summed in slice order, so that the launch fills the SMs
  const int chunks = (ncols + 63) / 64;
  const int wantSlices = (4 * ctx->num_sms + ns.nAtoms * chunks - 1) / std::max(1, ns.nAtoms * chunks);
  const int slices = std::max(1, std::min({32, wantSlices, ns.maxAtomRows / 64}));
  const size_t count = (size_t)ns.totalProj * ncols;
  double *out = ctx->nlProj[ctx->lane].p;
  if (slices > 1) {
    DB_TRY(ctx->nlPart[ctx->lane].alloc((size_t)slices * count));
    out = ctx->nlPart[ctx->lane].p;
  }
  {
    ProfScope ps(ctx, "nonlocal", slices > 1 ? 2 : 1);
    if (vec) {
      dim3 grid(ns.nAtoms, chunks, slices);
#define DB_NLP(CM, PM)                                                                                                 \
  nl_project_vec_kernel<CM, PM><<<grid, NLV_WARPS * 32, 0, ctx->stream>>>(x, ncols, ldx, ns.atomRowStart.p, ns.atomRows.p, \
                                                                         ns.atomValStart.p, ns.vals.p, ns.projOffset.p,  \
                                                                         rowScaleIn, out, count)
      if (ctx->cplx) {
        if (ns.maxProj <= 8) DB_NLP(2, 8); else DB_NLP(2, 16);
      } else {
        if (ns.maxProj <= 8) DB_NLP(1, 8); else DB_NLP(1, 16);
      }
#undef DB_NLP
    } else {
      dim3 grid(ns.nAtoms, (ncols + NL_COLS - 1) / NL_COLS, slices), block(NL_COLS, NL_RG);
      if (ctx->cplx)
        nl_project_kernel<2><<<grid, block, 0, ctx->stream>>>(x, ncols, ldx, ns.atomRowStart.p, ns.atomRows.p,
                                                              ns.atomValStart.p, ns.vals.p, ns.projOffset.p,
                                                              rowScaleIn, out, count);
      else
        nl_project_kernel<1><<<grid, block, 0, ctx->stream>>>(x, ncols, ldx, ns.atomRowStart.p, ns.atomRows.p,
                                                              ns.atomValStart.p, ns.vals.p, ns.projOffset.p,
                                                              rowScaleIn, out, count);
    }
    if (slices > 1) {
      const int grid = (int)std::max<size_t>(1, std::min<size_t>((count + 255) / 256, (size_t)ctx->num_sms * 4));
      nl_sum_slices_kernel<<<grid, 256, 0, ctx->stream>>>(out, count, slices, ctx->nlProj[ctx->lane].p, count);
    }
    DB_CUDA(cudaGetLastError());
  }
  return allreduce_sum(ctx, ctx->nlProj[ctx->lane].p, count);
}

// y += s * (out o Chat) V proj
int nonlocal_apply(dftfe_b200_ctx *ctx, double *y, int ncols, int ldx, const double *rowScaleOut, double s) {
  if (!ctx->nl || ctx->nl->nRows == 0 || ctx->skip_nonlocal) return 0;
  const dftfe_b200_ctx::NonlocalSet &ns = *ctx->nl;
  const bool vec = (ncols % 2 == 0) && (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                   !ctx->force_scalar_row_kernels && ns.maxProj <= 16;
  if (vec) {
    // atom-parallel, one launch per atom colour (fixed order)
    for (int c = 0; c < ns.nAtomColours; ++c) {
      const int nA = ns.colourStart_h[c + 1] - ns.colourStart_h[c];
      if (nA == 0) continue;
      ProfScope ps(ctx, "nonlocal");
      const int chunks = (ncols + 63) / 64;
      const int slices = std::max(1, std::min(32, (6 * ctx->num_sms + nA * chunks - 1) / (nA * chunks)));
      dim3 grid(nA, chunks, slices);
#define DB_NLA(CM, PM)                                                                                              \
  nl_apply_vec_kernel<CM, PM><<<grid, NLV_WARPS * 32, 0, ctx->stream>>>(                                              \
      y, ncols, ldx, ns.colourAtoms.p + ns.colourStart_h[c], ns.atomRowStart.p, ns.atomRows.p, ns.atomValStart.p,  \
      ns.vals.p, ns.projOffset.p, ns.V.p, ctx->nlProj[ctx->lane].p, rowScaleOut, s)
      if (ctx->cplx) {
        if (ns.maxProj <= 8) DB_NLA(2, 8); else DB_NLA(2, 16);
      } else {
        if (ns.maxProj <= 8) DB_NLA(1, 8); else DB_NLA(1, 16);
      }
#undef DB_NLA
    }
    DB_CUDA(cudaGetLastError());
    return 0;
  }
  ProfScope ps(ctx, "nonlocal");
  const int64_t total = ns.nRows * ncols;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 16);
  if (ctx->cplx)
    nl_apply_kernel<2><<<grid, 256, 0, ctx->stream>>>(y, ncols, ldx, ns.nRows, ns.rowList.p, ns.rowStart.p,
                                                      ns.entProj.p, ns.entVal.p, ns.V.p, ctx->nlProj[ctx->lane].p, rowScaleOut, s);
  else
    nl_apply_kernel<1><<<grid, 256, 0, ctx->stream>>>(y, ncols, ldx, ns.nRows, ns.rowList.p, ns.rowStart.p,
                                                      ns.entProj.p, ns.entVal.p, ns.V.p, ctx->nlProj[ctx->lane].p, rowScaleOut, s);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dftfe_b200
