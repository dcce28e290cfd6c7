// Ghost-DoF exchange and projected-matrix all-reduce.
//
// Reference: MPICommunicatorP2P<T,DEVICE>::updateGhostValues[Begin/End] /
// accumulateAddLocallyOwned[Begin/End] (utils/MPICommunicatorP2P.cc:103-418) move
// the packed rows with MPI_Isend/Irecv staged through pinned host memory;
// DeviceCCLWrapper (utils/DeviceDirectCCLWrapper.cc:96-261) all-reduces the
// projected blocks.  Here both go over NCCL (ncclSend/ncclRecv grouped per
// direction, ncclAllReduce) on the context stream - NVLink5/NVSwitch, no host hop.
//
// A second, in-process transport ("loopback") lets several ranks share ONE GPU:
// each rank is a context driven by its own host thread; buffers are exchanged with
// device-to-device copies between two process-wide barriers.  It exists so that
// the multi-rank path (pack -> exchange -> unpack-add, all-reduce) can be
// parity-tested on a single B200.
#include <dlfcn.h>

#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>

#include "common.cuh"

namespace dftfe_b200 {

int launch_pack_rows(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, int64_t nRows, const uint32_t *rows,
                     int64_t row0, double *buf);
int launch_pack_rows_f32(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, int64_t nRows,
                         const uint32_t *rows, int64_t row0, float *buf);
int launch_unpack_rows(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, int64_t row0, int64_t nRows,
                       const double *buf);
int launch_unpack_rows_f32(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, int64_t row0, int64_t nRows,
                           const float *buf);
int launch_unpack_add(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *buf,
                      const double *rowScale);
int launch_unpack_add_f32(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const float *buf,
                          const double *rowScale);

const NcclApi *nccl_api() {
  static NcclApi api;
  static int state = 0;  // 0 = untried, 1 = ok, -1 = failed
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (state == 0) {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    state = -1;
    if (h) {
#define DB_SYM(field, name) *(void **)(&api.field) = dlsym(h, name)
      DB_SYM(GetUniqueId, "ncclGetUniqueId");
      DB_SYM(CommInitRank, "ncclCommInitRank");
      DB_SYM(CommDestroy, "ncclCommDestroy");
      DB_SYM(Send, "ncclSend");
      DB_SYM(Recv, "ncclRecv");
      DB_SYM(AllReduce, "ncclAllReduce");
      DB_SYM(GroupStart, "ncclGroupStart");
      DB_SYM(GroupEnd, "ncclGroupEnd");
      DB_SYM(GetErrorString, "ncclGetErrorString");
#undef DB_SYM
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.AllReduce &&
          api.GroupStart && api.GroupEnd && api.GetErrorString)
        state = 1;
    }
    if (state != 1) set_error("libnccl.so.2 could not be loaded: %s", h ? "missing symbols" : dlerror());
  }
  return state == 1 ? &api : nullptr;
}

struct LoopbackGroup {
  int nranks = 0;
  std::vector<dftfe_b200_ctx *> members;
  std::vector<const double *> pub;  // per-rank published pointer
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  int64_t generation = 0;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const int64_t gen = generation;
    if (++waiting == nranks) {
      waiting = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};

static std::mutex g_groups_mu;
static std::map<int, std::shared_ptr<LoopbackGroup>> g_groups;

struct LoopbackHandle {
  std::shared_ptr<LoopbackGroup> grp;
};
static std::map<dftfe_b200_ctx *, LoopbackHandle> g_handles;

static LoopbackGroup *loopback_of(dftfe_b200_ctx *ctx) {
  std::lock_guard<std::mutex> lk(g_groups_mu);
  auto it = g_handles.find(ctx);
  return it == g_handles.end() ? nullptr : it->second.grp.get();
}

void loopback_forget(dftfe_b200_ctx *ctx) {
  std::lock_guard<std::mutex> lk(g_groups_mu);
  g_handles.erase(ctx);
}

int loopback_join(dftfe_b200_ctx *ctx, int group_id, int rank, int nranks) {
  std::lock_guard<std::mutex> lk(g_groups_mu);
  auto &g = g_groups[group_id];
  if (!g || g->nranks != nranks || (int)g->members.size() != nranks || g->members[rank] != nullptr) {
    g = std::make_shared<LoopbackGroup>();
    g->nranks = nranks;
    g->members.assign(nranks, nullptr);
    g->pub.assign(nranks, nullptr);
  }
  g->members[rank] = ctx;
  g_handles[ctx] = LoopbackHandle{g};
  return 0;
}

static int64_t target_offset(const dftfe_b200_ctx *c, int targetRank) {
  for (size_t t = 0; t < c->targetProcs_h.size(); ++t)
    if (c->targetProcs_h[t] == targetRank) return c->targetOffsets_h[t];
  return -1;
}

// Payload buffers of the current lane (see dftfe_b200_ctx::lane); FP32 payloads use the same storage.
static int ensure_payload(dftfe_b200_ctx *ctx) {
  const size_t cnt = (size_t)std::max<int64_t>(ctx->G, ctx->nSend) * ctx->B * ctx->cm;
  DB_TRY(ctx->sendB[ctx->lane].alloc(cnt));
  DB_TRY(ctx->recvB[ctx->lane].alloc(cnt));
  return 0;
}

// One exchange step.  forward: my packed owned rows (sendB, target order) -> the peers' ghost ranges;
// reverse: my ghost ranges -> the owners' receive slots (recvB, target order).  `esz` = payload element size.
// srcBase / dstBase are byte pointers to the first row of the outgoing / incoming payload.
static int exchange(dftfe_b200_ctx *ctx, bool forward, const char *srcBase, char *dstBase, int ncols, size_t esz) {
  const ncclDataType_t dt = esz == 8 ? ncclDouble : ncclFloat;
  const size_t rowBytes = (size_t)ncols * esz;
  if (ctx->nccl) {
    DB_NCCL(nccl_api()->GroupStart());
    for (size_t t = 0; t < ctx->targetProcs_h.size(); ++t) {
      const size_t off = (size_t)ctx->targetOffsets_h[t] * rowBytes, cnt = (size_t)ctx->nOwnedForTargets_h[t] * ncols;
      if (forward)
        DB_NCCL(nccl_api()->Send(srcBase + off, cnt, dt, ctx->targetProcs_h[t], ctx->nccl, ctx->stream));
      else
        DB_NCCL(nccl_api()->Recv(dstBase + off, cnt, dt, ctx->targetProcs_h[t], ctx->nccl, ctx->stream));
    }
    for (size_t g = 0; g < ctx->ghostProcs_h.size(); ++g) {
      const int64_t s = ctx->ghostRanges_h[2 * g], e = ctx->ghostRanges_h[2 * g + 1];
      if (forward)
        DB_NCCL(nccl_api()->Recv(dstBase + (size_t)s * rowBytes, (size_t)(e - s) * ncols, dt, ctx->ghostProcs_h[g],
                                 ctx->nccl, ctx->stream));
      else
        DB_NCCL(nccl_api()->Send(srcBase + (size_t)s * rowBytes, (size_t)(e - s) * ncols, dt, ctx->ghostProcs_h[g],
                                 ctx->nccl, ctx->stream));
    }
    DB_NCCL(nccl_api()->GroupEnd());
    return 0;
  }
  LoopbackGroup *grp = loopback_of(ctx);
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  grp->pub[ctx->rank] = reinterpret_cast<const double *>(srcBase);
  grp->barrier();
  if (forward) {
    for (size_t g = 0; g < ctx->ghostProcs_h.size(); ++g) {
      const int64_t s = ctx->ghostRanges_h[2 * g], e = ctx->ghostRanges_h[2 * g + 1];
      dftfe_b200_ctx *peer = grp->members[ctx->ghostProcs_h[g]];
      const int64_t off = target_offset(peer, ctx->rank);
      DB_CHECK(off >= 0, "loopback: rank %d is not a target of rank %d", ctx->rank, ctx->ghostProcs_h[g]);
      DB_CUDA(cudaMemcpyAsync(dstBase + (size_t)s * rowBytes,
                              reinterpret_cast<const char *>(grp->pub[peer->rank]) + (size_t)off * rowBytes,
                              (size_t)(e - s) * rowBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
  } else {
    for (size_t t = 0; t < ctx->targetProcs_h.size(); ++t) {
      dftfe_b200_ctx *peer = grp->members[ctx->targetProcs_h[t]];
      int64_t s = -1, e = -1;  // my range inside the peer's ghost segment
      for (size_t g = 0; g < peer->ghostProcs_h.size(); ++g)
        if (peer->ghostProcs_h[g] == ctx->rank) {
          s = peer->ghostRanges_h[2 * g];
          e = peer->ghostRanges_h[2 * g + 1];
        }
      DB_CHECK(s >= 0 && (e - s) == ctx->nOwnedForTargets_h[t], "loopback: inconsistent ghost pattern");
      DB_CUDA(cudaMemcpyAsync(dstBase + (size_t)ctx->targetOffsets_h[t] * rowBytes,
                              reinterpret_cast<const char *>(grp->pub[peer->rank]) + (size_t)s * rowBytes,
                              (size_t)(e - s) * rowBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  grp->barrier();
  return 0;
}

// forward: owned rows needed by my targets -> their ghost rows.  fp32: the payload travels as floats
// (HXCheby with chebMixedPrec, kohnShamDFTOperatorDevice.cc:3899-3915): ghost values arrive rounded to FP32.
int ghost_update(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, bool fp32) {
  if (ctx->nranks == 1) return 0;
  DB_CHECK(ctx->nccl || loopback_of(ctx), "ghost exchange needs comm_init (NCCL) or a loopback group");
  DB_TRY(ensure_payload(ctx));
  double *sendBuf = ctx->sendB[ctx->lane].p, *recvBuf = ctx->recvB[ctx->lane].p;
  if (fp32) {
    DB_TRY(launch_pack_rows_f32(ctx, x, ncols, ldx, ctx->nSend, ctx->sendRows.p, 0, reinterpret_cast<float *>(sendBuf)));
    DB_TRY(exchange(ctx, true, reinterpret_cast<const char *>(sendBuf), reinterpret_cast<char *>(recvBuf), ncols, 4));
    return launch_unpack_rows_f32(ctx, x, ncols, ldx, ctx->M, ctx->G, reinterpret_cast<const float *>(recvBuf));
  }
  DB_TRY(launch_pack_rows(ctx, x, ncols, ldx, ctx->nSend, ctx->sendRows.p, 0, sendBuf));
  const bool direct = (ldx == ncols);  // receives land straight in the ghost segment
  double *ghostBase = direct ? x + (size_t)ctx->M * ldx : recvBuf;
  DB_TRY(exchange(ctx, true, reinterpret_cast<const char *>(sendBuf), reinterpret_cast<char *>(ghostBase), ncols, 8));
  if (!direct) DB_TRY(launch_unpack_rows(ctx, x, ncols, ldx, ctx->M, ctx->G, recvBuf));
  return 0;
}

// reverse: my ghost rows -> added into their owners' rows (optionally scaled by rowScale[owner row]).
// fp32: kohnShamDFTOperatorDevice.cc:3953-3990 (float copy, float accumulate, boundary rows copied back).
int ghost_accumulate(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *rowScale, bool fp32) {
  if (ctx->nranks == 1) return 0;
  DB_CHECK(ctx->nccl || loopback_of(ctx), "ghost exchange needs comm_init (NCCL) or a loopback group");
  DB_TRY(ensure_payload(ctx));
  double *sendBuf = ctx->sendB[ctx->lane].p, *recvBuf = ctx->recvB[ctx->lane].p;
  if (fp32) {
    DB_TRY(launch_pack_rows_f32(ctx, x, ncols, ldx, ctx->G, nullptr, ctx->M, reinterpret_cast<float *>(sendBuf)));
    DB_TRY(exchange(ctx, false, reinterpret_cast<const char *>(sendBuf), reinterpret_cast<char *>(recvBuf), ncols, 4));
    return launch_unpack_add_f32(ctx, x, ncols, ldx, reinterpret_cast<const float *>(recvBuf), rowScale);
  }
  const bool direct = (ldx == ncols);
  const double *ghostBase = direct ? x + (size_t)ctx->M * ldx : sendBuf;
  if (!direct) DB_TRY(launch_pack_rows(ctx, x, ncols, ldx, ctx->G, nullptr, ctx->M, sendBuf));
  DB_TRY(exchange(ctx, false, reinterpret_cast<const char *>(ghostBase), reinterpret_cast<char *>(recvBuf), ncols, 8));
  return launch_unpack_add(ctx, x, ncols, ldx, recvBuf, rowScale);
}

int ghost_zero(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx) {
  if (ctx->G == 0) return 0;
  ctx->launches += 1;
  if (ldx == ncols) {
    DB_CUDA(cudaMemsetAsync(x + (size_t)ctx->M * ldx, 0, (size_t)ctx->G * ldx * sizeof(double), ctx->stream));
  } else {
    DB_CUDA(cudaMemset2DAsync(x + (size_t)ctx->M * ldx, (size_t)ldx * sizeof(double), 0,
                              (size_t)ncols * sizeof(double), (size_t)ctx->G, ctx->stream));
  }
  return 0;
}

int allreduce_sum(dftfe_b200_ctx *ctx, double *buf, size_t count) {
  if (ctx->nranks == 1) return 0;
  if (ctx->nccl) {
    DB_NCCL(nccl_api()->AllReduce(buf, buf, count, ncclDouble, ncclSum, ctx->nccl, ctx->stream));
    return 0;
  }
  LoopbackGroup *grp = loopback_of(ctx);
  DB_CHECK(grp, "allreduce needs comm_init (NCCL) or a loopback group");
  DB_TRY(ctx->arTmp.alloc(count));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  grp->pub[ctx->rank] = buf;
  grp->barrier();
  // rank-ordered sum so every rank gets the identical result
  DB_CUDA(cudaMemcpyAsync(ctx->arTmp.p, grp->pub[0], count * sizeof(double), cudaMemcpyDeviceToDevice,
                          ctx->stream));
  const double one = 1.0;
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  for (int r = 1; r < grp->nranks; ++r)
    DB_CUBLAS(cublasDaxpy(ctx->cublas, (int)count, &one, grp->pub[r], 1, ctx->arTmp.p, 1));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  grp->barrier();
  DB_CUDA(cudaMemcpyAsync(buf, ctx->arTmp.p, count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  grp->barrier();
  return 0;
}

int allreduce_sum_f32(dftfe_b200_ctx *ctx, float *buf, size_t count) {
  if (ctx->nranks == 1) return 0;
  if (ctx->nccl) {
    DB_NCCL(nccl_api()->AllReduce(buf, buf, count, ncclFloat, ncclSum, ctx->nccl, ctx->stream));
    return 0;
  }
  LoopbackGroup *grp = loopback_of(ctx);
  DB_CHECK(grp, "allreduce needs comm_init (NCCL) or a loopback group");
  DB_TRY(ctx->arTmpF.alloc(count));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  grp->pub[ctx->rank] = reinterpret_cast<const double *>(buf);
  grp->barrier();
  DB_CUDA(cudaMemcpyAsync(ctx->arTmpF.p, grp->pub[0], count * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  const float one = 1.0f;
  DB_CUBLAS(cublasSetStream(ctx->cublas, ctx->stream));
  for (int r = 1; r < grp->nranks; ++r)
    DB_CUBLAS(cublasSaxpy(ctx->cublas, (int)count, &one, reinterpret_cast<const float *>(grp->pub[r]), 1,
                          ctx->arTmpF.p, 1));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  grp->barrier();
  DB_CUDA(cudaMemcpyAsync(buf, ctx->arTmpF.p, count * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  grp->barrier();
  return 0;
}

}  // namespace dftfe_b200
