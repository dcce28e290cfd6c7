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
// (totalProj x B doubles) stays L2 resident between 1 and 3.
#include <map>

#include "common.cuh"

namespace dftfe_b200 {

namespace {

constexpr int NL_MAXP = 32;   // projectors per atom held in registers
constexpr int NL_COLS = 64;   // columns per CTA
constexpr int NL_RG = 2;      // row groups per CTA (static smem: RG*32*65*8 B)

__global__ void __launch_bounds__(NL_COLS *NL_RG)
nl_project_kernel(const double *__restrict__ x, int ncols, int ldx, const int32_t *__restrict__ atomRowStart,
                  const uint32_t *__restrict__ atomRows, const int64_t *__restrict__ atomValStart,
                  const double *__restrict__ vals, const int32_t *__restrict__ projOffset,
                  const double *__restrict__ rowScale, double *__restrict__ proj) {
  __shared__ double red[NL_RG][NL_MAXP][NL_COLS + 1];
  const int a = blockIdx.x;
  const int col = blockIdx.y * NL_COLS + threadIdx.x;
  const int rg = threadIdx.y;
  const int P = projOffset[a + 1] - projOffset[a];
  const int r0 = atomRowStart[a], r1 = atomRowStart[a + 1];
  const double *va = vals + atomValStart[a];
  double acc[NL_MAXP];
#pragma unroll
  for (int p = 0; p < NL_MAXP; ++p) acc[p] = 0.0;
  if (col < ncols) {
    for (int r = r0 + rg; r < r1; r += NL_RG) {
      const uint32_t row = atomRows[r];
      double xv = x[(size_t)row * ldx + col];
      if (rowScale) xv *= rowScale[row];
      const double *v = va + (size_t)(r - r0) * P;
#pragma unroll
      for (int p = 0; p < NL_MAXP; ++p)
        if (p < P) acc[p] += v[p] * xv;
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
    double sum = 0.0;
    for (int64_t e = rowStart[i]; e < rowStart[i + 1]; ++e) {
      const int id = entProj[e];
      sum += entVal[e] * V[id] * proj[(size_t)id * ncols + c];
    }
    const double f = rowScale ? s * rowScale[row] : s;
    y[(size_t)row * ldx + c] += f * sum;
  }
}

}  // namespace

int nonlocal_setup(dftfe_b200_ctx *ctx, int32_t nAtoms, const int32_t *nProj, const double *V, int64_t nEntries,
                   const int32_t *entryCell, const int32_t *entryAtom, const double *C, int32_t pMax) {
  DB_CHECK(ctx->have_map, "set_nonlocal: set_index_map first");
  const int n = ctx->n;
  std::vector<int32_t> off(nAtoms + 1, 0);
  for (int a = 0; a < nAtoms; ++a) {
    DB_CHECK(nProj[a] >= 0 && nProj[a] <= NL_MAXP && nProj[a] <= pMax,
             "set_nonlocal: atom %d has %d projectors (supported: <= %d and <= p_max)", a, nProj[a], NL_MAXP);
    off[a + 1] = off[a] + nProj[a];
  }
  const int totalProj = off[nAtoms];
  // assemble Chat per atom: row -> P values
  std::vector<std::map<uint32_t, std::vector<double>>> perAtom(nAtoms);
  for (int64_t e = 0; e < nEntries; ++e) {
    const int a = entryAtom[e];
    const int64_t c = entryCell[e];
    DB_CHECK(a >= 0 && a < nAtoms && c >= 0 && c < ctx->nC, "set_nonlocal: entry %lld out of range", (long long)e);
    const int P = nProj[a];
    for (int i = 0; i < n; ++i) {
      const uint32_t row = ctx->cellRows_h[c * n + i];
      auto &v = perAtom[a][row];
      if (v.empty()) v.assign(P, 0.0);
      const double *src = C + ((size_t)e * n + i) * pMax;
      for (int p = 0; p < P; ++p) v[p] += src[p];
    }
  }
  std::vector<int32_t> atomRowStart(nAtoms + 1, 0);
  std::vector<int64_t> atomValStart(nAtoms + 1, 0);
  std::vector<uint32_t> atomRows;
  std::vector<double> vals;
  std::map<uint32_t, std::vector<std::pair<int32_t, double>>> byRow;
  for (int a = 0; a < nAtoms; ++a) {
    const int P = nProj[a];
    for (auto &kv : perAtom[a]) {
      bool nz = false;
      for (double x : kv.second) nz = nz || (x != 0.0);
      if (!nz) continue;
      atomRows.push_back(kv.first);
      vals.insert(vals.end(), kv.second.begin(), kv.second.end());
      for (int p = 0; p < P; ++p)
        if (kv.second[p] != 0.0) byRow[kv.first].push_back({off[a] + p, kv.second[p]});
    }
    atomRowStart[a + 1] = (int32_t)atomRows.size();
    atomValStart[a + 1] = (int64_t)vals.size();
  }
  std::vector<uint32_t> rows;
  std::vector<int64_t> rowStart(1, 0);
  std::vector<int32_t> entProj;
  std::vector<double> entVal;
  for (auto &kv : byRow) {
    rows.push_back(kv.first);
    for (auto &pr : kv.second) {
      entProj.push_back(pr.first);
      entVal.push_back(pr.second);
    }
    rowStart.push_back((int64_t)entProj.size());
  }
  ctx->nlAtoms = nAtoms;
  ctx->nlTotalProj = totalProj;
  ctx->nlRows = (int64_t)rows.size();
  DB_TRY(ctx->nlProjOffset.upload(off.data(), off.size(), ctx->stream));
  DB_TRY(ctx->nlV.upload(V, totalProj, ctx->stream));
  DB_TRY(ctx->nlAtomRowStart.upload(atomRowStart.data(), atomRowStart.size(), ctx->stream));
  DB_TRY(ctx->nlAtomValStart.upload(atomValStart.data(), atomValStart.size(), ctx->stream));
  DB_TRY(ctx->nlAtomRows.upload(atomRows.data(), atomRows.size(), ctx->stream));
  DB_TRY(ctx->nlVals.upload(vals.data(), vals.size(), ctx->stream));
  DB_TRY(ctx->nlRowList.upload(rows.data(), rows.size(), ctx->stream));
  DB_TRY(ctx->nlRowStart.upload(rowStart.data(), rowStart.size(), ctx->stream));
  DB_TRY(ctx->nlEntProj.upload(entProj.data(), entProj.size(), ctx->stream));
  DB_TRY(ctx->nlEntVal.upload(entVal.data(), entVal.size(), ctx->stream));
  DB_TRY(ctx->nlProj.alloc((size_t)std::max(totalProj, 1) * ctx->B));
  ctx->have_nonlocal = totalProj > 0;
  return 0;
}

// proj = Chat^T (in o x), all-reduced over ranks
int nonlocal_project(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, const double *rowScaleIn) {
  if (!ctx->have_nonlocal) return 0;
  {
    ProfScope ps(ctx, "nonlocal");
    dim3 grid(ctx->nlAtoms, (ncols + NL_COLS - 1) / NL_COLS), block(NL_COLS, NL_RG);
    nl_project_kernel<<<grid, block, 0, ctx->stream>>>(x, ncols, ldx, ctx->nlAtomRowStart.p, ctx->nlAtomRows.p,
                                                       ctx->nlAtomValStart.p, ctx->nlVals.p, ctx->nlProjOffset.p,
                                                       rowScaleIn, ctx->nlProj.p);
    DB_CUDA(cudaGetLastError());
  }
  return allreduce_sum(ctx, ctx->nlProj.p, (size_t)ctx->nlTotalProj * ncols);
}

// y += s * (out o Chat) V proj
int nonlocal_apply(dftfe_b200_ctx *ctx, double *y, int ncols, int ldx, const double *rowScaleOut, double s) {
  if (!ctx->have_nonlocal || ctx->nlRows == 0) return 0;
  ProfScope ps(ctx, "nonlocal");
  const int64_t total = ctx->nlRows * ncols;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 16);
  nl_apply_kernel<<<grid, 256, 0, ctx->stream>>>(y, ncols, ldx, ctx->nlRows, ctx->nlRowList.p, ctx->nlRowStart.p,
                                                 ctx->nlEntProj.p, ctx->nlEntVal.p, ctx->nlV.p, ctx->nlProj.p,
                                                 rowScaleOut, s);
  DB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dftfe_b200
