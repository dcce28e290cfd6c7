// Context lifetime and setup: the B200-native equivalent of
// kohnShamDFTOperatorDeviceClass::reinit (src/dftOperator/kohnShamDFTOperatorDevice.cc:492-933).
#include <algorithm>
#include <cstring>
#include <mutex>
#include <numeric>
#include <set>

#include "common.cuh"

namespace dftfe_b200 {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ensure_dyn_smem(const void *kernel, int device, size_t bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void *, int>> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({kernel, device})) return 0;
  DB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  done.insert({kernel, device});
  return 0;
}

int loopback_join(dftfe_b200_ctx *ctx, int group_id, int rank, int nranks);
int band_loopback_join(dftfe_b200_ctx *ctx, int group_id, int band_id, int n_groups);
void loopback_forget(dftfe_b200_ctx *ctx);

// row words of the flagged index map: row | live<<30 | first<<31 (see cell_matvec.cu)
static int rebuild_row_words(dftfe_b200_ctx *ctx) {
  if (!ctx->have_map) return 0;
  const int64_t R = ctx->M + ctx->G;
  std::vector<char> live(R, 0);
  for (int64_t r = 0; r < ctx->M; ++r) live[r] = 1;
  for (uint32_t r : ctx->conRows_h) live[r] = 0;
  std::vector<uint32_t> words(ctx->cellRows_h.size());
  for (size_t k = 0; k < words.size(); ++k) {
    const uint32_t r = ctx->cellRows_h[k];
    words[k] = r | (live[r] ? 0x40000000u : 0u) | (ctx->firstTouch_h[k] ? 0x80000000u : 0u);
  }
  std::vector<uint32_t> orph(ctx->orphanRows_h.size());
  for (size_t k = 0; k < orph.size(); ++k) orph[k] = ctx->orphanRows_h[k] | (live[ctx->orphanRows_h[k]] ? 0x40000000u : 0u);
  DB_TRY(ctx->cellRowsFlagged.upload(words.data(), words.size(), ctx->stream));
  DB_TRY(ctx->orphanRows.upload(orph.data(), orph.size(), ctx->stream));
  return 0;
}

// derived per-row scale vectors (see solver.cu: fused_apply_impl)
static int rebuild_row_vectors(dftfe_b200_ctx *ctx, const double *sqrtM_h, const double *invSqrtM_h) {
  const int64_t R = ctx->M + ctx->G;
  std::vector<char> con(R, 0);
  for (uint32_t r : ctx->conRows_h) con[r] = 1;
  std::vector<double> rowIn(R), rowOut(R), rowLive(R), rowLiveInv(R), rowInInv(R), rowOutInv(R);
  for (int64_t r = 0; r < R; ++r) {
    const bool owned = r < ctx->M;
    rowIn[r] = con[r] ? 1.0 : invSqrtM_h[r];
    rowOut[r] = (owned && !con[r]) ? invSqrtM_h[r] : 1.0;
    rowLive[r] = (owned && !con[r]) ? 1.0 : 0.0;
    rowLiveInv[r] = (owned && !con[r]) ? invSqrtM_h[r] : 0.0;
    rowInInv[r] = rowIn[r] != 0.0 ? 1.0 / rowIn[r] : 0.0;
    rowOutInv[r] = rowOut[r] != 0.0 ? 1.0 / rowOut[r] : 0.0;
  }
  DB_TRY(ctx->rowIn.upload(rowIn.data(), R, ctx->stream));
  DB_TRY(ctx->rowOut.upload(rowOut.data(), R, ctx->stream));
  DB_TRY(ctx->rowLive.upload(rowLive.data(), R, ctx->stream));
  DB_TRY(ctx->rowLiveInvSqrtM.upload(rowLiveInv.data(), R, ctx->stream));
  DB_TRY(ctx->rowInInv.upload(rowInInv.data(), R, ctx->stream));
  DB_TRY(ctx->rowOutInv.upload(rowOutInv.data(), R, ctx->stream));
  ctx->have_H = false;  // scales are folded into the tiled H: it must be set again
  (void)sqrtM_h;
  return 0;
}

}  // namespace dftfe_b200

using namespace dftfe_b200;

#define DB_CTX(ctx)                                                   \
  do {                                                                \
    if (!(ctx)) {                                                     \
      set_error("null context");                                      \
      return DFTFE_B200_ERR_INVALID;                                  \
    }                                                                 \
    cudaError_t e__ = cudaSetDevice((ctx)->desc.device);              \
    if (e__ != cudaSuccess) {                                         \
      set_error("cudaSetDevice failed: %s", cudaGetErrorString(e__)); \
      return DFTFE_B200_ERR_CUDA;                                     \
    }                                                                 \
  } while (0)

extern "C" {

const char *dftfe_b200_version(void) { return "dftfe_b200 0.1.0 (sm_100a)"; }
const char *dftfe_b200_last_error(void) { return g_err; }

int dftfe_b200_create(const dftfe_b200_problem_desc *desc, dftfe_b200_ctx **out) {
  DB_CHECK(desc && out, "create: null argument");
  DB_CHECK(desc->n_cells >= 0 && desc->n_owned >= 0 && desc->n_ghost >= 0, "create: negative size");
  DB_CHECK(desc->cheby_block >= 1, "create: cheby_block must be >= 1");
  DB_CHECK((desc->n_owned + desc->n_ghost) < (int64_t)0x3fffffff, "create: more than 2^30-1 local rows");
  if (!cell_kernel_supported(desc->nodes_per_cell)) {
    set_error("create: no sm_100a cell kernel for %d nodes per cell (supported FE orders 1..7)",
              desc->nodes_per_cell);
    return DFTFE_B200_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("create: no CUDA device available (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return DFTFE_B200_ERR_CUDA;
  }
  DB_CHECK(desc->device >= 0 && desc->device < ndev, "create: device %d out of range [0,%d)", desc->device, ndev);
  DB_CUDA(cudaSetDevice(desc->device));
  cudaDeviceProp prop;
  DB_CUDA(cudaGetDeviceProperties(&prop, desc->device));
  if (prop.major != 10) {
    set_error("create: device %d is sm_%d%d; this library is built for sm_100a only", desc->device, prop.major,
              prop.minor);
    return DFTFE_B200_ERR_UNSUPPORTED;
  }
  auto *ctx = new dftfe_b200_ctx();
  ctx->desc = *desc;
  ctx->n = desc->nodes_per_cell;
  ctx->B = desc->cheby_block;
  ctx->nC = desc->n_cells;
  ctx->M = desc->n_owned;
  ctx->G = desc->n_ghost;
  ctx->cplx = (desc->flags & DFTFE_B200_FLAG_COMPLEX) != 0;
  ctx->cm = ctx->cplx ? 2 : 1;
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cublasCreate(&ctx->cublas) != CUBLAS_STATUS_SUCCESS ||
      cusolverDnCreate(&ctx->cusolver) != CUSOLVER_STATUS_SUCCESS) {
    set_error("create: stream / cuBLAS / cuSOLVER handle creation failed");
    delete ctx;
    return DFTFE_B200_ERR_CUDA;
  }
  ctx->own_stream = true;
  ctx->owned_stream = ctx->stream;
  *out = ctx;
  return 0;
}

void dftfe_b200_destroy(dftfe_b200_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->desc.device);
  cudaDeviceSynchronize();
  loopback_forget(ctx);
  for (auto &kv : ctx->prof)
    for (auto &pr : kv.second.pending) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  for (int l = 0; l < 2; ++l) {
    if (ctx->laneStream[l]) cudaStreamDestroy(ctx->laneStream[l]);
    if (ctx->laneEvent[l]) cudaEventDestroy(ctx->laneEvent[l]);
  }
  if (ctx->forkEvent) cudaEventDestroy(ctx->forkEvent);
  for (cudaEvent_t e : ctx->hostLoopEvents) cudaEventDestroy(e);
  if (ctx->copyIn) cudaStreamDestroy(ctx->copyIn);
  if (ctx->copyOut) cudaStreamDestroy(ctx->copyOut);
  p2p_release(ctx);
  if (ctx->bandNccl && nccl_api()) nccl_api()->CommDestroy(ctx->bandNccl);
  if (ctx->nccl && nccl_api()) nccl_api()->CommDestroy(ctx->nccl);
  if (ctx->cusolver) cusolverDnDestroy(ctx->cusolver);
  if (ctx->cublas) cublasDestroy(ctx->cublas);
  if (ctx->owned_stream) cudaStreamDestroy(ctx->owned_stream);
  delete ctx;
}

int dftfe_b200_set_stream(dftfe_b200_ctx *ctx, void *cuda_stream) {
  DB_CTX(ctx);
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->owned_stream;
  return 0;
}

int dftfe_b200_sync(dftfe_b200_ctx *ctx) {
  DB_CTX(ctx);
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dftfe_b200_build_index_map(const int64_t *cell_global_dofs_h, int64_t n_cells, int32_t nodes_per_cell,
                               int64_t owned_start, int64_t owned_end, const int64_t *ghost_sorted_h,
                               int64_t n_ghost, int32_t block, uint64_t *map_out_h) {
  DB_CHECK(cell_global_dofs_h && map_out_h && (n_ghost == 0 || ghost_sorted_h), "build_index_map: null argument");
  const int64_t M = owned_end - owned_start;
  const int64_t total = n_cells * nodes_per_cell;
  for (int64_t k = 0; k < total; ++k) {
    const int64_t g = cell_global_dofs_h[k];
    int64_t loc;
    if (g >= owned_start && g < owned_end) {
      loc = g - owned_start;
    } else {
      const int64_t *it = std::lower_bound(ghost_sorted_h, ghost_sorted_h + n_ghost, g);
      DB_CHECK(it != ghost_sorted_h + n_ghost && *it == g,
               "build_index_map: global DoF %lld of cell entry %lld is neither owned nor ghost", (long long)g,
               (long long)k);
      loc = M + (it - ghost_sorted_h);
    }
    map_out_h[k] = (uint64_t)loc * (uint64_t)block;
  }
  return 0;
}

int dftfe_b200_set_index_map(dftfe_b200_ctx *ctx, const uint64_t *map_h) {
  DB_CTX(ctx);
  DB_CHECK(map_h || ctx->nC == 0, "set_index_map: null map");
  const int n = ctx->n;
  const int64_t nC = ctx->nC, R = ctx->M + ctx->G;
  const int64_t total = nC * n;
  ctx->cellRows_h.resize(total);
  for (int64_t k = 0; k < total; ++k) {
    DB_CHECK(map_h[k] % (uint64_t)ctx->B == 0, "set_index_map: entry %lld is not a multiple of the block size",
             (long long)k);
    const uint64_t r = map_h[k] / (uint64_t)ctx->B;
    DB_CHECK((int64_t)r < R, "set_index_map: entry %lld points past the local vector (%llu >= %lld)", (long long)k,
             (unsigned long long)r, (long long)R);
    ctx->cellRows_h[k] = (uint32_t)r;
  }
  // row -> cells adjacency (CSR)
  std::vector<int64_t> rowStart(R + 1, 0);
  for (int64_t k = 0; k < total; ++k) rowStart[ctx->cellRows_h[k] + 1]++;
  for (int64_t r = 0; r < R; ++r) rowStart[r + 1] += rowStart[r];
  std::vector<int32_t> rowCells(total);
  {
    std::vector<int64_t> fill(rowStart.begin(), rowStart.end() - 1);
    for (int64_t c = 0; c < nC; ++c)
      for (int i = 0; i < n; ++i) rowCells[fill[ctx->cellRows_h[c * n + i]]++] = (int32_t)c;
  }
  // a cell must not list the same row twice (the assembly would race inside one CTA)
  for (int64_t r = 0; r < R; ++r)
    for (int64_t k = rowStart[r] + 1; k < rowStart[r + 1]; ++k)
      DB_CHECK(rowCells[k] != rowCells[k - 1], "set_index_map: cell %d references local row %lld twice",
               rowCells[k], (long long)r);
  // greedy colouring of the cell adjacency graph (cells sharing a row are adjacent)
  ctx->cellColour_h.assign(nC, -1);
  int nColours = 0;
  {
    std::vector<int64_t> stamp;  // stamp[colour] == c  <=> colour used by a neighbour of c
    for (int64_t c = 0; c < nC; ++c) {
      for (int i = 0; i < n; ++i) {
        const uint32_t r = ctx->cellRows_h[c * n + i];
        for (int64_t k = rowStart[r]; k < rowStart[r + 1]; ++k) {
          const int col = ctx->cellColour_h[rowCells[k]];
          if (col >= 0) {
            if ((int)stamp.size() <= col) stamp.resize(col + 1, -1);
            stamp[col] = c;
          }
        }
      }
      int pick = 0;
      while (pick < (int)stamp.size() && stamp[pick] == c) ++pick;
      if ((int)stamp.size() <= pick) stamp.resize(pick + 1, -1);
      ctx->cellColour_h[c] = pick;
      nColours = std::max(nColours, pick + 1);
    }
  }
  ctx->nColours = nColours;
  ctx->colourStart_h.assign(nColours + 1, 0);
  for (int64_t c = 0; c < nC; ++c) ctx->colourStart_h[ctx->cellColour_h[c] + 1]++;
  for (int k = 0; k < nColours; ++k) ctx->colourStart_h[k + 1] += ctx->colourStart_h[k];
  std::vector<int32_t> colourCells(nC);
  {
    std::vector<int32_t> fill(ctx->colourStart_h.begin(), ctx->colourStart_h.end() - 1);
    for (int64_t c = 0; c < nC; ++c) colourCells[fill[ctx->cellColour_h[c]]++] = (int32_t)c;
  }
  // first-touch flags in colour processing order
  ctx->firstTouch_h.assign(total, 0);
  std::vector<char> touched(R, 0);
  for (int64_t q = 0; q < nC; ++q) {
    const int64_t c = colourCells[q];
    for (int i = 0; i < n; ++i) {
      const uint32_t r = ctx->cellRows_h[c * n + i];
      if (!touched[r]) {
        touched[r] = 1;
        ctx->firstTouch_h[c * n + i] = 1;
      }
    }
  }
  ctx->orphanRows_h.clear();
  for (int64_t r = 0; r < R; ++r)
    if (!touched[r]) ctx->orphanRows_h.push_back((uint32_t)r);
  ctx->nOrphan = (int64_t)ctx->orphanRows_h.size();
  DB_TRY(ctx->colourCells.upload(colourCells.data(), colourCells.size(), ctx->stream));
  ctx->have_map = true;
  return rebuild_row_words(ctx);
}

int dftfe_b200_set_constraints(dftfe_b200_ctx *ctx, int64_t nCon, const uint32_t *rows_h, const uint32_t *sizes_h,
                               const uint32_t *starts_h, const uint32_t *cols_h, const double *vals_h,
                               const double *inhom_h) {
  DB_CTX(ctx);
  DB_CHECK(nCon >= 0, "set_constraints: negative count");
  const int64_t R = ctx->M + ctx->G;
  int64_t nnz = 0;
  std::vector<char> isCon(R, 0);
  for (int64_t i = 0; i < nCon; ++i) {
    DB_CHECK((int64_t)rows_h[i] < R, "set_constraints: row %lld out of range", (long long)i);
    DB_CHECK(!isCon[rows_h[i]], "set_constraints: row %u constrained twice", rows_h[i]);
    isCon[rows_h[i]] = 1;
    DB_CHECK((int64_t)starts_h[i] == nnz, "set_constraints: row_starts must be the exclusive prefix sum of sizes");
    nnz += sizes_h[i];
  }
  for (int64_t k = 0; k < nnz; ++k) {
    DB_CHECK((int64_t)cols_h[k] < R, "set_constraints: column entry %lld out of range", (long long)k);
    DB_CHECK(!isCon[cols_h[k]], "set_constraints: column %u is itself constrained (constraints must be closed)",
             cols_h[k]);
  }
  ctx->nCon = nCon;
  ctx->nnz = nnz;
  ctx->conRows_h.assign(rows_h, rows_h + nCon);
  DB_TRY(ctx->conRows.upload(rows_h, nCon, ctx->stream));
  DB_TRY(ctx->conSizes.upload(sizes_h, nCon, ctx->stream));
  DB_TRY(ctx->conStarts.upload(starts_h, nCon, ctx->stream));
  DB_TRY(ctx->conCols.upload(cols_h, nnz, ctx->stream));
  DB_TRY(ctx->conVals.upload(vals_h, nnz, ctx->stream));
  DB_TRY(ctx->conInhom.upload(inhom_h, nCon, ctx->stream));
  // transposed CSR: master -> (slave, weight), entries in ascending (constraint, column) order
  std::vector<uint32_t> cnt(R, 0);
  for (int64_t k = 0; k < nnz; ++k) cnt[cols_h[k]]++;
  std::vector<uint32_t> masters, mstarts(1, 0);
  std::vector<int64_t> slotOf(R, -1);
  for (int64_t r = 0; r < R; ++r)
    if (cnt[r]) {
      slotOf[r] = (int64_t)masters.size();
      masters.push_back((uint32_t)r);
      mstarts.push_back(mstarts.back() + cnt[r]);
    }
  std::vector<uint32_t> slaves(nnz), fill(mstarts.begin(), mstarts.end() - 1);
  std::vector<double> mvals(nnz);
  for (int64_t i = 0; i < nCon; ++i)
    for (uint32_t j = 0; j < sizes_h[i]; ++j) {
      const int64_t k = (int64_t)starts_h[i] + j;
      const int64_t s = slotOf[cols_h[k]];
      slaves[fill[s]] = rows_h[i];
      mvals[fill[s]] = vals_h[k];
      fill[s]++;
    }
  ctx->nMasters = (int64_t)masters.size();
  DB_TRY(ctx->masterRows.upload(masters.data(), masters.size(), ctx->stream));
  DB_TRY(ctx->masterStarts.upload(mstarts.data(), mstarts.size(), ctx->stream));
  DB_TRY(ctx->masterSlaves.upload(slaves.data(), slaves.size(), ctx->stream));
  DB_TRY(ctx->masterVals.upload(mvals.data(), mvals.size(), ctx->stream));
  if (ctx->have_mass) DB_TRY(rebuild_row_vectors(ctx, ctx->sqrtM_h.data(), ctx->invSqrtM_h.data()));
  ctx->have_H = false;
  return rebuild_row_words(ctx);
}

int dftfe_b200_set_mass(dftfe_b200_ctx *ctx, const double *sqrt_mass_h, const double *inv_sqrt_mass_h) {
  DB_CTX(ctx);
  DB_CHECK(sqrt_mass_h && inv_sqrt_mass_h, "set_mass: null argument");
  const int64_t R = ctx->M + ctx->G;
  ctx->sqrtM_h.assign(sqrt_mass_h, sqrt_mass_h + R);
  ctx->invSqrtM_h.assign(inv_sqrt_mass_h, inv_sqrt_mass_h + R);
  DB_TRY(ctx->sqrtM.upload(sqrt_mass_h, R, ctx->stream));
  DB_TRY(ctx->invSqrtM.upload(inv_sqrt_mass_h, R, ctx->stream));
  ctx->have_mass = true;
  return rebuild_row_vectors(ctx, sqrt_mass_h, inv_sqrt_mass_h);
}

int dftfe_b200_set_ghost_pattern(dftfe_b200_ctx *ctx, int32_t rank, int32_t nranks, int32_t nGhostProcs,
                                 const int32_t *ghostProcs_h, const int32_t *ghostRanges_h, int32_t nTargetProcs,
                                 const int32_t *targetProcs_h, const int32_t *nOwnedForTargets_h,
                                 const uint32_t *ownedIdx_h) {
  DB_CTX(ctx);
  DB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "set_ghost_pattern: bad rank/nranks");
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->ghostProcs_h.assign(ghostProcs_h, ghostProcs_h + nGhostProcs);
  ctx->ghostRanges_h.assign(ghostRanges_h, ghostRanges_h + 2 * nGhostProcs);
  ctx->targetProcs_h.assign(targetProcs_h, targetProcs_h + nTargetProcs);
  ctx->nOwnedForTargets_h.assign(nOwnedForTargets_h, nOwnedForTargets_h + nTargetProcs);
  int64_t covered = 0;
  for (int g = 0; g < nGhostProcs; ++g) {
    DB_CHECK(ghostRanges_h[2 * g] == covered && ghostRanges_h[2 * g + 1] >= ghostRanges_h[2 * g],
             "set_ghost_pattern: ghost ranges must tile the ghost segment contiguously");
    covered = ghostRanges_h[2 * g + 1];
  }
  DB_CHECK(covered == ctx->G, "set_ghost_pattern: ghost ranges cover %lld rows, expected %lld", (long long)covered,
           (long long)ctx->G);
  ctx->targetOffsets_h.assign(nTargetProcs + 1, 0);
  for (int t = 0; t < nTargetProcs; ++t)
    ctx->targetOffsets_h[t + 1] = ctx->targetOffsets_h[t] + nOwnedForTargets_h[t];
  ctx->nSend = ctx->targetOffsets_h[nTargetProcs];
  for (int64_t k = 0; k < ctx->nSend; ++k)
    DB_CHECK((int64_t)ownedIdx_h[k] < ctx->M, "set_ghost_pattern: send index %lld is not an owned row", (long long)k);
  DB_TRY(ctx->sendRows.upload(ownedIdx_h, ctx->nSend, ctx->stream));
  // transposed unpack map: boundary row -> receive slots (ascending = target order)
  std::vector<uint32_t> cnt(ctx->M, 0);
  for (int64_t k = 0; k < ctx->nSend; ++k) cnt[ownedIdx_h[k]]++;
  std::vector<uint32_t> rows, starts(1, 0);
  std::vector<int64_t> slotOf(ctx->M, -1);
  for (int64_t r = 0; r < ctx->M; ++r)
    if (cnt[r]) {
      slotOf[r] = (int64_t)rows.size();
      rows.push_back((uint32_t)r);
      starts.push_back(starts.back() + cnt[r]);
    }
  std::vector<uint32_t> slots(ctx->nSend), fill(starts.begin(), starts.end() - 1);
  for (int64_t k = 0; k < ctx->nSend; ++k) slots[fill[slotOf[ownedIdx_h[k]]]++] = (uint32_t)k;
  ctx->nBoundaryRows = (int64_t)rows.size();
  DB_TRY(ctx->bndRows.upload(rows.data(), rows.size(), ctx->stream));
  DB_TRY(ctx->bndStarts.upload(starts.data(), starts.size(), ctx->stream));
  DB_TRY(ctx->bndSlots.upload(slots.data(), slots.size(), ctx->stream));
  return 0;
}

int dftfe_b200_nccl_unique_id(uint8_t id_out_h[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  if (!nccl_api()) return DFTFE_B200_ERR_NCCL;
  DB_NCCL(nccl_api()->GetUniqueId(&id));
  std::memcpy(id_out_h, &id, 128);
  return 0;
}

int dftfe_b200_comm_init(dftfe_b200_ctx *ctx, const uint8_t id_h[128], int32_t rank, int32_t nranks) {
  DB_CTX(ctx);
  DB_CHECK(!ctx->nccl, "comm_init: communicator already initialised");
  ncclUniqueId id;
  std::memcpy(&id, id_h, 128);
  if (!nccl_api()) return DFTFE_B200_ERR_NCCL;
  DB_NCCL(nccl_api()->CommInitRank(&ctx->nccl, nranks, id, rank));
  ctx->rank = rank;
  ctx->nranks = nranks;
  return 0;
}

int dftfe_b200_comm_init_loopback(dftfe_b200_ctx *ctx, int32_t group_id, int32_t rank, int32_t nranks) {
  DB_CTX(ctx);
  DB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "comm_init_loopback: bad rank/nranks");
  ctx->rank = rank;
  ctx->nranks = nranks;
  return loopback_join(ctx, group_id, rank, nranks);
}

int dftfe_b200_band_comm_init(dftfe_b200_ctx *ctx, const uint8_t id_h[128], int32_t band_group_id,
                              int32_t n_band_groups) {
  DB_CTX(ctx);
  DB_CHECK(!ctx->bandNccl, "band_comm_init: communicator already initialised");
  DB_CHECK(n_band_groups >= 1 && band_group_id >= 0 && band_group_id < n_band_groups, "band_comm_init: bad group id");
  ncclUniqueId id;
  std::memcpy(&id, id_h, 128);
  if (!nccl_api()) return DFTFE_B200_ERR_NCCL;
  DB_NCCL(nccl_api()->CommInitRank(&ctx->bandNccl, n_band_groups, id, band_group_id));
  ctx->bandId = band_group_id;
  ctx->nBandGroups = n_band_groups;
  return 0;
}

int dftfe_b200_band_comm_init_loopback(dftfe_b200_ctx *ctx, int32_t group_id, int32_t band_group_id,
                                       int32_t n_band_groups) {
  DB_CTX(ctx);
  DB_CHECK(n_band_groups >= 1 && band_group_id >= 0 && band_group_id < n_band_groups,
           "band_comm_init_loopback: bad group id");
  ctx->bandId = band_group_id;
  ctx->nBandGroups = n_band_groups;
  return band_loopback_join(ctx, group_id, band_group_id, n_band_groups);
}

int dftfe_b200_band_group_indices(int32_t n_band_groups, int32_t N, int32_t *low_high_plus_one_out_h) {
  DB_CHECK(n_band_groups >= 1 && N / n_band_groups >= 1 && low_high_plus_one_out_h,
           "band_group_indices: NPBAND is more than the number of bands");
  for (int g = 0; g < n_band_groups; ++g) {
    int lo, hi;
    band_group_range(n_band_groups, N, g, lo, hi);
    low_high_plus_one_out_h[2 * g] = lo;
    low_high_plus_one_out_h[2 * g + 1] = hi;
  }
  return 0;
}

int dftfe_b200_band_group_merge(dftfe_b200_ctx *ctx, double *X_d, int32_t N) {
  DB_CTX(ctx);
  DB_CHECK(X_d && N >= 1, "band_group_merge: null argument");
  return band_group_merge(ctx, X_d, N);
}

int dftfe_b200_set_nonlocal_kpt(dftfe_b200_ctx *ctx, int32_t kpoint_index, int32_t n_atoms,
                                const int32_t *n_proj_per_atom_h, const double *V_h, int64_t n_entries,
                                const int32_t *entry_cell_h, const int32_t *entry_atom_h, const double *C_h,
                                int32_t p_max) {
  DB_CTX(ctx);
  DB_CHECK(n_atoms >= 0 && n_entries >= 0 && p_max >= 0 && kpoint_index >= 0, "set_nonlocal: negative size / index");
  return nonlocal_setup(ctx, kpoint_index, n_atoms, n_proj_per_atom_h, V_h, n_entries, entry_cell_h, entry_atom_h, C_h,
                        p_max);
}

int dftfe_b200_set_nonlocal(dftfe_b200_ctx *ctx, int32_t n_atoms, const int32_t *n_proj_per_atom_h, const double *V_h,
                            int64_t n_entries, const int32_t *entry_cell_h, const int32_t *entry_atom_h,
                            const double *C_h, int32_t p_max) {
  return dftfe_b200_set_nonlocal_kpt(ctx, 0, n_atoms, n_proj_per_atom_h, V_h, n_entries, entry_cell_h, entry_atom_h,
                                     C_h, p_max);
}

// flat key of a (k-point, spin) cell-Hamiltonian set
static inline int hkey(int kpt, int spin) { return 2 * kpt + spin; }

int dftfe_b200_set_cell_hamiltonian_kpt(dftfe_b200_ctx *ctx, int32_t kpoint_index, int32_t spin_index,
                                        const double *H_d) {
  DB_CTX(ctx);
  DB_CHECK(H_d || ctx->nC == 0, "set_cell_hamiltonian: null pointer");
  DB_CHECK(kpoint_index >= 0 && (spin_index == 0 || spin_index == 1),
           "set_cell_hamiltonian: bad (k-point, spin) index (%d, %d)", kpoint_index, spin_index);
  if (!ctx->have_H) ctx->Hsets.clear();  // constraints / mass changed: every stored set is stale
  ctx->activeK = hkey(kpoint_index, spin_index);
  if (ctx->nC > 0) DB_TRY(retile_cell_hamiltonian(ctx, H_d));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));  // caller's buffer is free after return
  ctx->have_H = true;
  auto nl = ctx->nlSets.find(kpoint_index);
  if (nl != ctx->nlSets.end()) ctx->nl = nl->second.totalProj > 0 ? &nl->second : nullptr;
  return 0;
}

int dftfe_b200_set_cell_hamiltonian(dftfe_b200_ctx *ctx, const double *H_d) {
  return dftfe_b200_set_cell_hamiltonian_kpt(ctx, 0, 0, H_d);
}

int dftfe_b200_reinit_kpoint_spin_index(dftfe_b200_ctx *ctx, int32_t kpoint_index, int32_t spin_index) {
  DB_CTX(ctx);
  auto it = ctx->Hsets.find(hkey(kpoint_index, spin_index));
  DB_CHECK(ctx->have_H && it != ctx->Hsets.end() && (it->second.p || ctx->nC == 0),
           "reinit_kpoint_spin_index: no cell Hamiltonian stored for (k-point %d, spin %d)", kpoint_index, spin_index);
  ctx->activeK = it->first;
  ctx->Hactive = it->second.p;
  // the projector matrices depend on the k-point only (h_d_B / h_d_C re-pointing, kohnShamDFTOperatorDevice.cc:1043-1056)
  auto nl = ctx->nlSets.find(kpoint_index);
  if (nl != ctx->nlSets.end())
    ctx->nl = nl->second.totalProj > 0 ? &nl->second : nullptr;
  else
    DB_CHECK(ctx->nlSets.empty() || !ctx->cplx,
             "reinit_kpoint_spin_index: non-local projectors were set for other k-points but not for k-point %d",
             kpoint_index);
  return 0;
}

int dftfe_b200_compute_cell_hamiltonian(dftfe_b200_ctx *ctx, int32_t n_quad, const double *shape_values_d,
                                        const double *veff_jxw_d, const double *grad_integral_d,
                                        int32_t grad_integral_per_cell, const double *cell_kscale_d,
                                        const double *ext_pot_corr_d, double *H_out_d) {
  DB_CTX(ctx);
  return compute_cell_hamiltonian(ctx, n_quad, shape_values_d, veff_jxw_d, grad_integral_d, grad_integral_per_cell,
                                  cell_kscale_d, ext_pot_corr_d, H_out_d);
}

int dftfe_b200_compute_cell_hamiltonian_gga(dftfe_b200_ctx *ctx, int32_t n_quad, const double *shape_values_d,
                                            const double *shape_grad_values_d, const double *inv_jacobian_d,
                                            const double *veff_jxw_d, const double *der_exc_sigma_grad_rho_jxw_d,
                                            const double *grad_integral_d, int32_t grad_integral_per_cell,
                                            const double *cell_kscale_d, const double *ext_pot_corr_d,
                                            double *H_out_d) {
  DB_CTX(ctx);
  return compute_cell_hamiltonian_gga(ctx, n_quad, shape_values_d, shape_grad_values_d, inv_jacobian_d, veff_jxw_d,
                                      der_exc_sigma_grad_rho_jxw_d, grad_integral_d, grad_integral_per_cell,
                                      cell_kscale_d, ext_pot_corr_d, H_out_d);
}

int dftfe_b200_compute_cell_hamiltonian_kpoints(dftfe_b200_ctx *ctx, int32_t n_quad, const double *shape_values_d,
                                                const double *shape_grad_values_d, const double *inv_jacobian_d,
                                                const double *jxw_d, const double *H_real_d, int32_t n_kpoints,
                                                const double *kpoint_coords_h, double *H_k_out_d) {
  DB_CTX(ctx);
  return compute_cell_hamiltonian_kpoints(ctx, n_quad, shape_values_d, shape_grad_values_d, inv_jacobian_d, jxw_d,
                                          H_real_d, n_kpoints, kpoint_coords_h, H_k_out_d);
}

int dftfe_b200_compute_density(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, const double *occupations_h,
                               int32_t n_quad, const double *shape_values_d, double *rho_out_d) {
  DB_CTX(ctx);
  DB_CHECK(X_d && occupations_h && shape_values_d && rho_out_d && N >= 1, "compute_density: null argument");
  return compute_density(ctx, X_d, N, occupations_h, n_quad, shape_values_d, rho_out_d);
}

int dftfe_b200_compute_density_grad(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, const double *occupations_h,
                                    int32_t n_quad, const double *shape_values_d, const double *shape_grad_values_d,
                                    const double *inv_jacobian_d, double *rho_out_d, double *grad_rho_out_d) {
  DB_CTX(ctx);
  DB_CHECK(X_d && occupations_h && shape_values_d && shape_grad_values_d && rho_out_d && grad_rho_out_d && N >= 1,
           "compute_density_grad: null argument");
  return compute_density(ctx, X_d, N, occupations_h, n_quad, shape_values_d, rho_out_d, shape_grad_values_d,
                         inv_jacobian_d, grad_rho_out_d);
}

int dftfe_b200_set_cell_hamiltonian_host(dftfe_b200_ctx *ctx, const double *H_h) {
  DB_CTX(ctx);
  const size_t count = (size_t)ctx->nC * ctx->n * ctx->n * ctx->cm;
  DB_TRY(ctx->Hstage.upload(H_h, count, ctx->stream));
  int rc = dftfe_b200_set_cell_hamiltonian(ctx, ctx->Hstage.p);
  ctx->Hstage.release();
  return rc;
}

int dftfe_b200_update_ghost_values(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols) {
  DB_CTX(ctx);
  DB_CHECK(ncols >= 1 && ncols <= ctx->B, "update_ghost_values: ncols (%d) must be in [1, cheby_block=%d]", ncols, ctx->B);
  return ghost_update(ctx, x_d, ncols * ctx->cm, ncols * ctx->cm);
}
int dftfe_b200_accumulate_add_locally_owned(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols) {
  DB_CTX(ctx);
  DB_CHECK(ncols >= 1 && ncols <= ctx->B, "accumulate_add_locally_owned: ncols (%d) must be in [1, cheby_block=%d]", ncols,
           ctx->B);
  return ghost_accumulate(ctx, x_d, ncols * ctx->cm, ncols * ctx->cm, nullptr);
}
int dftfe_b200_zero_out_ghosts(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols) {
  DB_CTX(ctx);
  return ghost_zero(ctx, x_d, ncols * ctx->cm, ncols * ctx->cm);
}
int dftfe_b200_constraints_distribute(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols) {
  DB_CTX(ctx);
  return launch_distribute(ctx, x_d, ncols * ctx->cm, ncols * ctx->cm, nullptr);
}
int dftfe_b200_constraints_distribute_slave_to_master(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols) {
  DB_CTX(ctx);
  return launch_slave_to_master(ctx, x_d, ncols * ctx->cm, ncols * ctx->cm, nullptr);
}
int dftfe_b200_constraints_set_zero(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols) {
  DB_CTX(ctx);
  return launch_set_zero_rows(ctx, x_d, ncols * ctx->cm, ncols * ctx->cm);
}

int dftfe_b200_strided_copy_to_block(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, int32_t j0, double *block_d,
                                     int32_t ncols) {
  DB_CTX(ctx);
  DB_CHECK(j0 >= 0 && ncols >= 1 && j0 + ncols <= N, "strided_copy_to_block: columns [%d, %d) outside [0, %d)", j0,
           j0 + ncols, N);
  const int cm = ctx->cm;
  return launch_block_copy_from_full(ctx, X_d, N * cm, j0 * cm, block_d, ncols * cm, ctx->M, nullptr);
}

int dftfe_b200_strided_copy_from_block(dftfe_b200_ctx *ctx, double *X_d, int32_t N, int32_t j0, const double *block_d,
                                       int32_t ncols) {
  DB_CTX(ctx);
  DB_CHECK(j0 >= 0 && ncols >= 1 && j0 + ncols <= N, "strided_copy_from_block: columns [%d, %d) outside [0, %d)", j0,
           j0 + ncols, N);
  const int cm = ctx->cm;
  return launch_block_copy_to_full(ctx, X_d, N * cm, j0 * cm, block_d, ncols * cm, ctx->M, nullptr);
}

int dftfe_b200_strided_block_scale(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols, double alpha, int32_t which) {
  DB_CTX(ctx);
  DB_CHECK(ctx->have_mass || which == 0, "strided_block_scale: set_mass first");
  const double *s = which == 1 ? ctx->sqrtM.p : (which == 2 ? ctx->invSqrtM.p : nullptr);
  DB_CHECK(which >= 0 && which <= 2, "strided_block_scale: which must be 0 (none), 1 (M^1/2) or 2 (M^-1/2)");
  return launch_row_scale(ctx, x_d, ctx->M, ncols * ctx->cm, ncols * ctx->cm, alpha, s);
}

int dftfe_b200_get_colouring(dftfe_b200_ctx *ctx, int32_t *n_colours_out, int32_t *cell_colour_out_h) {
  DB_CTX(ctx);
  DB_CHECK(ctx->have_map, "get_colouring: set_index_map first");
  if (n_colours_out) *n_colours_out = ctx->nColours;
  if (cell_colour_out_h) std::memcpy(cell_colour_out_h, ctx->cellColour_h.data(), ctx->nC * sizeof(int32_t));
  return 0;
}

int dftfe_b200_set_option(dftfe_b200_ctx *ctx, const char *name, int32_t value) {
  DB_CTX(ctx);
  DB_CHECK(name, "set_option: null name");
  if (std::strcmp(name, "generic_cell_kernel") == 0) {
    ctx->force_generic_cell_kernel = value != 0;
    return 0;
  }
  if (std::strcmp(name, "scalar_row_kernels") == 0) {
    ctx->force_scalar_row_kernels = value != 0;
    return 0;
  }
  if (std::strcmp(name, "reserved_sms") == 0) {
    DB_CHECK(value >= 0 && value < ctx->num_sms, "set_option: reserved_sms out of range");
    ctx->reserved_sms = value;
    return 0;
  }
  if (std::strcmp(name, "p2p_exchange") == 0) {
    DB_CHECK(!ctx->p2p.tried, "set_option: p2p_exchange must be chosen before the first ghost exchange");
    ctx->p2p.requested = value;
    return 0;
  }
  if (std::strcmp(name, "overlap_lanes") == 0) {
    ctx->overlap_lanes = value;
    return 0;
  }
  if (std::strcmp(name, "cublas_projections") == 0) {
    ctx->use_cublas_dense = value != 0;
    return 0;
  }
  if (std::strcmp(name, "only_h_prime") == 0) {
    // onlyHPrimePartForFirstOrderDensityMatResponse of HX / XtHX (kohnShamDFTOperatorDevice.cc:3680-3688): the cell
    // matrices are the caller's H' matrices and the non-local term is skipped
    ctx->skip_nonlocal = value != 0;
    return 0;
  }
  set_error("set_option: unknown option '%s'", name);
  return DFTFE_B200_ERR_INVALID;
}

int dftfe_b200_profile_enable(dftfe_b200_ctx *ctx, int32_t enable) {
  DB_CTX(ctx);
  ctx->profiling = enable != 0;
  return 0;
}

static void profile_drain(dftfe_b200_ctx *ctx) {
  cudaStreamSynchronize(ctx->stream);
  for (auto &kv : ctx->prof) {
    for (auto &pr : kv.second.pending) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) kv.second.total_ms += ms;
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    kv.second.pending.clear();
  }
}

int dftfe_b200_profile_get(dftfe_b200_ctx *ctx, const char *name, double *total_ms_out, int64_t *launches_out) {
  DB_CTX(ctx);
  profile_drain(ctx);
  auto it = ctx->prof.find(name ? name : "");
  if (total_ms_out) *total_ms_out = it == ctx->prof.end() ? 0.0 : it->second.total_ms;
  if (launches_out) *launches_out = it == ctx->prof.end() ? 0 : it->second.launches;
  return 0;
}

int dftfe_b200_profile_reset(dftfe_b200_ctx *ctx) {
  DB_CTX(ctx);
  profile_drain(ctx);
  ctx->prof.clear();
  return 0;
}

int64_t dftfe_b200_launch_count(dftfe_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

const char *dftfe_b200_transport_name(dftfe_b200_ctx *ctx) { return ctx ? transport_name(ctx) : ""; }

}  // extern "C"
