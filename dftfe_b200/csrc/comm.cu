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

#include <time.h>

#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>

#include <cuda.h>

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
int launch_push_rows(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, int64_t nRows, const uint32_t *rows,
                     int64_t row0, const PushSegs &segs, bool fp32);
int launch_signal_flags(dftfe_b200_ctx *ctx, const SignalList &l, uint32_t value);

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
      DB_SYM(Broadcast, "ncclBroadcast");
      DB_SYM(GroupStart, "ncclGroupStart");
      DB_SYM(GroupEnd, "ncclGroupEnd");
      DB_SYM(GetErrorString, "ncclGetErrorString");
#undef DB_SYM
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.AllReduce && api.Broadcast &&
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
static std::map<int, std::shared_ptr<LoopbackGroup>> g_band_groups;       // in-process band-group communicators
static std::map<dftfe_b200_ctx *, LoopbackHandle> g_band_handles;

static LoopbackGroup *loopback_of(dftfe_b200_ctx *ctx) {
  std::lock_guard<std::mutex> lk(g_groups_mu);
  auto it = g_handles.find(ctx);
  return it == g_handles.end() ? nullptr : it->second.grp.get();
}

void loopback_forget(dftfe_b200_ctx *ctx) {
  std::lock_guard<std::mutex> lk(g_groups_mu);
  g_handles.erase(ctx);
  g_band_handles.erase(ctx);
}

static LoopbackGroup *band_loopback_of(dftfe_b200_ctx *ctx) {
  std::lock_guard<std::mutex> lk(g_groups_mu);
  auto it = g_band_handles.find(ctx);
  return it == g_band_handles.end() ? nullptr : it->second.grp.get();
}

int band_loopback_join(dftfe_b200_ctx *ctx, int group_id, int band_id, int n_groups) {
  std::lock_guard<std::mutex> lk(g_groups_mu);
  auto &g = g_band_groups[group_id];
  if (!g || g->nranks != n_groups || (int)g->members.size() != n_groups || g->members[band_id] != nullptr) {
    g = std::make_shared<LoopbackGroup>();
    g->nranks = n_groups;
    g->members.assign(n_groups, nullptr);
    g->pub.assign(n_groups, nullptr);
  }
  g->members[band_id] = ctx;
  g_band_handles[ctx] = LoopbackHandle{g};
  return 0;
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

// ---------------------------------------------------------------------------------------------------------
// "p2p" transport: the exchange without NCCL's send/recv kernels.
//
// NCCL's point-to-point kernels need a whole SM's worth of registers/threads, so they cannot co-reside with the
// persistent cell kernel of the other filter lane: every exchange queued behind a colour launch (0.5 ms of the
// 1.26 ms launch exposed per exchange, 169 ms per filter step at 8 GPUs).  Here the packing kernel itself stores
// the rows into the destination rank's receive buffer through a peer mapping (NVLink / NVSwitch), a one-warp
// kernel publishes a sequence number in the destination's flag array, and the consumer's STREAM waits for it with
// cuStreamWaitValue32 - a stream memory operation that holds no SM.  The small row kernels fit beside a resident
// cell CTA (256 threads, <= 48 registers), so one lane's exchange really runs under the other lane's GEMMs.
// Semantics are those of MPICommunicatorP2P::updateGhostValues / accumulateAddLocallyOwned
// (utils/MPICommunicatorP2P.cc:103-418); the arithmetic (unpack-add order, FP32 payload rounding) is unchanged.
//
// Hand-shake per (lane, direction), sequence number s = 1, 2, ...:
//   producer: wait ACK[dst] >= s-1 (receive buffer free)  -> push rows -> signal DATA[me] = s at each destination
//   consumer: wait DATA[src] >= s from each source        -> unpack    -> signal ACK[me]  = s at each source
// Every rank issues the same sequence of exchanges (the operator apply is collective), so the counters agree.
// ---------------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

WaitValue32Fn wait_value32_fn() {
  static WaitValue32Fn fn = nullptr;
  static bool tried = false;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WaitValue32Fn>(p);
  }
  return fn;
}

enum { P2P_DATA = 0, P2P_ACK = 1 };
enum { P2P_FWD = 0, P2P_REV = 1, P2P_AR = 2 };  // flag banks: forward / reverse exchange, projector all-reduce
constexpr int P2P_NFLAGS = 2 * 3 * 2;              // [lane][bank][DATA | ACK] x nranks

inline uint32_t *p2p_flag(char *slab, size_t offFlags, int nranks, int lane, int dir, int kind, int src) {
  return reinterpret_cast<uint32_t *>(slab + offFlags) + ((size_t)((lane * 3 + dir) * 2 + kind) * nranks + src);
}

int p2p_wait(dftfe_b200_ctx *ctx, int dir, int kind, int src, uint32_t value) {
  uint32_t *f = p2p_flag(ctx->p2p.slab, ctx->p2p.offFlags, ctx->nranks, ctx->lane, dir, kind, src);
  ctx->launches += 0;  // a stream memory operation, not a kernel
  const CUresult r = wait_value32_fn()(ctx->stream, (CUdeviceptr)f, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) {
    set_error("cuStreamWaitValue32 failed (%d)", (int)r);
    return DFTFE_B200_ERR_CUDA;
  }
  return 0;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// record every rank publishes: [0..7] cudaIpc handle (64 bytes), [8] offFlags, [9..10] offRecvF, [11..12] offRecvR,
// [13 .. 13+nranks) first row of rank q's rows in MY ghost segment, [13+nranks .. 13+2 nranks) first row of rank
// q's range in MY reverse receive buffer (-1: q is not a neighbour)
constexpr int P2P_REC_FIXED = 16;  // ... [13..14] offAr, [15] arMax

void p2p_fill_starts(const dftfe_b200_ctx *c, int64_t *ghostStart, int64_t *targetStart) {
  for (int r = 0; r < c->nranks; ++r) ghostStart[r] = targetStart[r] = -1;
  for (size_t g = 0; g < c->ghostProcs_h.size(); ++g) ghostStart[c->ghostProcs_h[g]] = c->ghostRanges_h[2 * g];
  for (size_t t = 0; t < c->targetProcs_h.size(); ++t) targetStart[c->targetProcs_h[t]] = c->targetOffsets_h[t];
}

}  // namespace

void p2p_release(dftfe_b200_ctx *ctx) {
  auto &p = ctx->p2p;
  if (p.active && p.slab) {
    // my neighbours acknowledge my last payloads by storing into THIS slab: wait (bounded) until every expected
    // acknowledgement has landed before the memory goes away
    const int nr = ctx->nranks;
    std::vector<uint32_t> flags((size_t)P2P_NFLAGS * nr);
    for (int spin = 0; spin < 2000; ++spin) {
      if (cudaMemcpy(flags.data(), p.slab + p.offFlags, flags.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost) !=
          cudaSuccess)
        break;
      bool done = true;
      for (int lane = 0; lane < 2; ++lane) {
        for (int dir = 0; dir < 2; ++dir) {
          const std::vector<int32_t> &dst = dir == 0 ? ctx->targetProcs_h : ctx->ghostProcs_h;
          for (int q : dst)
            done = done && (int32_t)(flags[((size_t)((lane * 3 + dir) * 2 + P2P_ACK)) * nr + q] - p.seq[lane][dir]) >= 0;
        }
        if (p.arMax > 0)
          for (int q = 0; q < nr; ++q)
            if (q != ctx->rank)
              done = done && (int32_t)(flags[((size_t)((lane * 3 + P2P_AR) * 2 + P2P_ACK)) * nr + q] - p.seqAr[lane]) >= 0;
      }
      if (done) break;
      struct timespec ts = {0, 1000000};
      nanosleep(&ts, nullptr);
    }
  }
  if (p.ipc)
    for (char *q : p.peerSlab)
      if (q) cudaIpcCloseMemHandle(q);
  p.peerSlab.clear();
  if (p.slab) cudaFree(p.slab);
  p.slab = nullptr;
  p.active = false;
}

static int p2p_setup(dftfe_b200_ctx *ctx) {
  auto &p = ctx->p2p;
  p.tried = true;
  LoopbackGroup *grp = loopback_of(ctx);
  const bool wanted = p.requested == 1 || (p.requested == -1 && ctx->nccl != nullptr);
  const int nr = ctx->nranks;
  // all ranks take the same decision: `requested` is set identically by the caller, the rest is agreed below
  if (!wanted || (!ctx->nccl && !grp)) return 0;
  bool ok = wait_value32_fn() != nullptr && ctx->targetProcs_h.size() <= (size_t)P2P_MAX_PEERS &&
            ctx->ghostProcs_h.size() <= (size_t)P2P_MAX_PEERS;
  // slab layout
  const size_t rowMax = (size_t)ctx->B * ctx->cm * sizeof(double);
  p.offFlags = 0;
  size_t off = align_up((size_t)P2P_NFLAGS * nr * sizeof(uint32_t), 256);
  for (int l = 0; l < 2; ++l) {
    p.offRecvF[l] = off;
    off = align_up(off + (size_t)ctx->G * rowMax, 256);
    p.offRecvR[l] = off;
    off = align_up(off + (size_t)ctx->nSend * rowMax, 256);
  }
  // projector all-reduce slots (only when projectors are already set: the block size is the same on every rank)
  p.arMax = 0;
  for (auto &kv : ctx->nlSets) p.arMax = std::max<size_t>(p.arMax, (size_t)kv.second.totalProj * ctx->B * ctx->cm);
  for (int l = 0; l < 2; ++l) {
    p.offAr[l] = off;
    off = align_up(off + (size_t)nr * p.arMax * sizeof(double), 256);
  }
  p.slabBytes = off;
  DB_CUDA(cudaMalloc(&p.slab, p.slabBytes));
  DB_CUDA(cudaMemsetAsync(p.slab, 0, align_up((size_t)P2P_NFLAGS * nr * sizeof(uint32_t), 256), ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ok) {
    // probe: a stream wait on an already satisfied flag must be accepted by this driver / device (else: NCCL transport)
    const CUresult pr = wait_value32_fn()(ctx->stream, (CUdeviceptr)(p.slab + p.offFlags), 0, CU_STREAM_WAIT_VALUE_GEQ);
    if (pr != CUDA_SUCCESS || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      cudaGetLastError();
      ok = false;
    }
  }
  p.peerSlab.assign(nr, nullptr);
  p.peerOffFlags.assign(nr, 0);
  for (int l = 0; l < 2; ++l) {
    p.peerOffRecvF[l].assign(nr, 0);
    p.peerOffRecvR[l].assign(nr, 0);
    p.peerOffAr[l].assign(nr, 0);
  }
  p.peerGhostStartOfMe.assign(nr, -1);
  p.peerTargetStartOfMe.assign(nr, -1);
  std::vector<char> neighbour(nr, 0);
  for (int q : ctx->targetProcs_h) neighbour[q] = 1;
  for (int q : ctx->ghostProcs_h) neighbour[q] = 1;
  if (p.arMax > 0)  // the all-reduce talks to every rank
    for (int q = 0; q < nr; ++q) neighbour[q] = (q != ctx->rank);
  neighbour[ctx->rank] = 0;

  if (grp && !ctx->nccl) {
    // one process, several ranks on one device: peers' slabs are plain pointers
    grp->barrier();
    for (int q = 0; q < nr; ++q) {
      if (!neighbour[q]) continue;
      const dftfe_b200_ctx *pc = grp->members[q];
      if (!pc->p2p.slab) {
        ok = false;
        continue;
      }
      p.peerSlab[q] = pc->p2p.slab;
      p.peerOffFlags[q] = pc->p2p.offFlags;
      for (int l = 0; l < 2; ++l) {
        p.peerOffRecvF[l][q] = pc->p2p.offRecvF[l];
        p.peerOffRecvR[l][q] = pc->p2p.offRecvR[l];
        p.peerOffAr[l][q] = pc->p2p.offAr[l];
      }
      if (pc->p2p.arMax != p.arMax) ok = false;
      std::vector<int64_t> gs(nr), ts(nr);
      p2p_fill_starts(pc, gs.data(), ts.data());
      p.peerGhostStartOfMe[q] = gs[ctx->rank];
      p.peerTargetStartOfMe[q] = ts[ctx->rank];
    }
    grp->barrier();
    p.active = ok;
    if (!ok) {
      set_error("p2p exchange requested but not available (stream memory operations / peer slabs missing)");
      return DFTFE_B200_ERR_UNSUPPORTED;
    }
    return 0;
  }

  // one process per GPU: publish (IPC handle, offsets, range starts) with an all-reduce over a zero-padded table
  const int W = P2P_REC_FIXED + 2 * nr;
  std::vector<int64_t> table((size_t)nr * W, 0);
  int64_t *mine = table.data() + (size_t)ctx->rank * W;
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  if (cudaIpcGetMemHandle(&h, p.slab) != cudaSuccess) {
    cudaGetLastError();
    ok = false;
    std::memset(&h, 0, sizeof(h));
  }
  std::memcpy(mine, &h, 64);
  mine[8] = (int64_t)p.offFlags;
  mine[9] = (int64_t)p.offRecvF[0];
  mine[10] = (int64_t)p.offRecvF[1];
  mine[11] = (int64_t)p.offRecvR[0];
  mine[12] = (int64_t)p.offRecvR[1];
  mine[13] = (int64_t)p.offAr[0];
  mine[14] = (int64_t)p.offAr[1];
  mine[15] = (int64_t)p.arMax;
  p2p_fill_starts(ctx, mine + P2P_REC_FIXED, mine + P2P_REC_FIXED + nr);
  DevBuf<int64_t> dtab;
  DB_TRY(dtab.upload(table.data(), table.size(), ctx->stream));
  DB_NCCL(nccl_api()->AllReduce(dtab.p, dtab.p, table.size(), ncclInt64, ncclSum, ctx->nccl, ctx->stream));
  DB_CUDA(cudaMemcpyAsync(table.data(), dtab.p, table.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < nr && ok; ++q) {
    if (!neighbour[q]) continue;
    const int64_t *rec = table.data() + (size_t)q * W;
    cudaIpcMemHandle_t hq;
    std::memcpy(&hq, rec, 64);
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = false;
      break;
    }
    p.ipc = true;
    p.peerSlab[q] = static_cast<char *>(ptr);
    p.peerOffFlags[q] = (size_t)rec[8];
    p.peerOffRecvF[0][q] = (size_t)rec[9];
    p.peerOffRecvF[1][q] = (size_t)rec[10];
    p.peerOffRecvR[0][q] = (size_t)rec[11];
    p.peerOffRecvR[1][q] = (size_t)rec[12];
    p.peerOffAr[0][q] = (size_t)rec[13];
    p.peerOffAr[1][q] = (size_t)rec[14];
    if ((size_t)rec[15] != p.arMax) ok = false;  // (projectors set on some ranks only: keep NCCL)
    p.peerGhostStartOfMe[q] = rec[P2P_REC_FIXED + ctx->rank];
    p.peerTargetStartOfMe[q] = rec[P2P_REC_FIXED + nr + ctx->rank];
  }
  // agree: p2p only if EVERY rank could map all of its neighbours
  int64_t bad = ok ? 0 : 1;
  DevBuf<int64_t> dbad;
  DB_TRY(dbad.upload(&bad, 1, ctx->stream));
  DB_NCCL(nccl_api()->AllReduce(dbad.p, dbad.p, 1, ncclInt64, ncclSum, ctx->nccl, ctx->stream));
  DB_CUDA(cudaMemcpyAsync(&bad, dbad.p, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  DB_CUDA(cudaStreamSynchronize(ctx->stream));
  p.active = bad == 0;
  if (!p.active) {
    p2p_release(ctx);
    if (p.requested == 1) {
      set_error("p2p exchange requested but %lld rank(s) could not map their neighbours' buffers (no peer access?)",
                (long long)bad);
      return DFTFE_B200_ERR_UNSUPPORTED;
    }
  }
  return 0;
}

const char *transport_name(dftfe_b200_ctx *ctx) {
  if (ctx->nranks == 1) return "none (single rank)";
  if (ctx->p2p.active)
    return ctx->p2p.ipc ? "p2p: rows stored into cudaIpc-mapped peer buffers over NVLink, stream-memory-op hand-shake"
                        : "p2p (in-process): rows stored into the peer rank's buffers, stream-memory-op hand-shake";
  if (ctx->nccl) return ctx->p2p.tried ? "nccl send/recv (p2p unavailable or disabled)" : "nccl send/recv";
  return "loopback (host-synchronised device copies)";
}

// one exchange over the p2p transport.  forward: owned rows (sendRows) -> the targets' ghost rows;
// reverse: my ghost rows -> added into the owners' rows.
static int exchange_p2p(dftfe_b200_ctx *ctx, bool forward, double *x, int ncols, int ldx, const double *rowScale,
                        bool fp32) {
  auto &p = ctx->p2p;
  const int lane = ctx->lane, dir = forward ? 0 : 1, nr = ctx->nranks;
  const uint32_t seq = ++p.seq[lane][dir];
  const size_t rowBytes = (size_t)ncols * (fp32 ? 4 : 8);
  const std::vector<int32_t> &dstProcs = forward ? ctx->targetProcs_h : ctx->ghostProcs_h;
  const std::vector<int32_t> &srcProcs = forward ? ctx->ghostProcs_h : ctx->targetProcs_h;
  // 1. the destinations have consumed the previous payload of this (lane, direction)
  if (seq > 1)
    for (int q : dstProcs) DB_TRY(p2p_wait(ctx, dir, P2P_ACK, q, seq - 1));
  // 2. push my rows into their receive buffers, then publish the sequence number
  PushSegs segs;
  SignalList sig;
  segs.n = sig.n = (int)dstProcs.size();
  for (int s = 0; s < segs.n; ++s) {
    const int q = dstProcs[s];
    const int64_t at = forward ? p.peerGhostStartOfMe[q] : p.peerTargetStartOfMe[q];
    DB_CHECK(at >= 0 && p.peerSlab[q], "p2p exchange: rank %d does not expect rows from rank %d", q, ctx->rank);
    segs.start[s] = forward ? ctx->targetOffsets_h[s] : ctx->ghostRanges_h[2 * s];
    segs.dst[s] = p.peerSlab[q] + (forward ? p.peerOffRecvF[lane][q] : p.peerOffRecvR[lane][q]) + (size_t)at * rowBytes;
    sig.addr[s] = p2p_flag(p.peerSlab[q], p.peerOffFlags[q], nr, lane, dir, P2P_DATA, ctx->rank);
  }
  segs.start[segs.n] = forward ? ctx->nSend : ctx->G;
  if (forward) {
    DB_TRY(launch_push_rows(ctx, x, ncols, ldx, ctx->nSend, ctx->sendRows.p, 0, segs, fp32));
  } else if (!fp32 && ldx == ncols && ctx->G > 0) {
    // reverse, FP64, dense rows: a neighbour's rows are one contiguous run of my ghost segment - peer copies on the
    // copy engines, no SM (the forward direction gathers scattered owned rows and needs the push kernel)
    ProfScope ps(ctx, "ghost_pack", segs.n);
    for (int s = 0; s < segs.n; ++s) {
      const int64_t r0 = ctx->ghostRanges_h[2 * s], r1 = ctx->ghostRanges_h[2 * s + 1];
      if (r1 > r0)
        DB_CUDA(cudaMemcpyAsync(segs.dst[s], x + (size_t)(ctx->M + r0) * ldx, (size_t)(r1 - r0) * rowBytes,
                                cudaMemcpyDeviceToDevice, ctx->stream));
    }
  } else {
    DB_TRY(launch_push_rows(ctx, x, ncols, ldx, ctx->G, nullptr, ctx->M, segs, fp32));
  }
  DB_TRY(launch_signal_flags(ctx, sig, seq));
  // In-process ranks share ONE device: a device-synchronising call (cudaFree, ...) made by one rank's host thread
  // waits for every rank's streams, so a stream must never wait on work a sibling thread has not submitted yet.
  // The host barrier guarantees every rank's signal is enqueued before anyone enqueues the wait (test transport
  // only - separate processes on separate devices need no host hand-shake).
  LoopbackGroup *inproc = ctx->nccl ? nullptr : loopback_of(ctx);
  if (inproc) inproc->barrier();
  // 3. wait for every source's payload, unpack, acknowledge
  for (int q : srcProcs) DB_TRY(p2p_wait(ctx, dir, P2P_DATA, q, seq));
  char *recv = p.slab + (forward ? p.offRecvF[lane] : p.offRecvR[lane]);
  if (forward) {
    if (fp32)
      DB_TRY(launch_unpack_rows_f32(ctx, x, ncols, ldx, ctx->M, ctx->G, reinterpret_cast<const float *>(recv)));
    else
      DB_TRY(launch_unpack_rows(ctx, x, ncols, ldx, ctx->M, ctx->G, reinterpret_cast<const double *>(recv)));
  } else {
    if (fp32)
      DB_TRY(launch_unpack_add_f32(ctx, x, ncols, ldx, reinterpret_cast<const float *>(recv), rowScale));
    else
      DB_TRY(launch_unpack_add(ctx, x, ncols, ldx, reinterpret_cast<const double *>(recv), rowScale));
  }
  SignalList ack;
  ack.n = (int)srcProcs.size();
  for (int s = 0; s < ack.n; ++s)
    ack.addr[s] = p2p_flag(p.peerSlab[srcProcs[s]], p.peerOffFlags[srcProcs[s]], nr, lane, dir, P2P_ACK, ctx->rank);
  DB_TRY(launch_signal_flags(ctx, ack, seq));
  if (inproc) inproc->barrier();  // the next exchange's ACK wait must find this acknowledgement submitted
  return 0;
}

// forward: owned rows needed by my targets -> their ghost rows.  fp32: the payload travels as floats
// (HXCheby with chebMixedPrec, kohnShamDFTOperatorDevice.cc:3899-3915): ghost values arrive rounded to FP32.
int ghost_update(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, bool fp32) {
  if (ctx->nranks == 1) return 0;
  DB_CHECK(ctx->nccl || loopback_of(ctx), "ghost exchange needs comm_init (NCCL) or a loopback group");
  if (!ctx->p2p.tried) DB_TRY(p2p_setup(ctx));
  if (ctx->p2p.active) return exchange_p2p(ctx, true, x, ncols, ldx, nullptr, fp32);
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
  if (!ctx->p2p.tried) DB_TRY(p2p_setup(ctx));
  if (ctx->p2p.active) return exchange_p2p(ctx, false, x, ncols, ldx, rowScale, fp32);
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

namespace {
// buf[i] = sum over the ranks, in rank order (identical bits on every rank), of the slots; my own contribution is buf
__global__ void p2p_sum_slots_kernel(double *__restrict__ buf, const double *__restrict__ slots, size_t slotStride,
                                     size_t count, int nranks, int me) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < nranks; ++r) s += (r == me) ? buf[i] : slots[(size_t)r * slotStride + i];
    buf[i] = s;
  }
}
}  // namespace

// All-reduce over the peer-mapped slabs: every rank copies its block into its slot of every other rank's slab (peer
// copies on the copy engines, no SM), publishes the sequence number, waits for the others with stream memory
// operations and sums the slots in rank order.  Replaces ncclAllReduce for the per-apply projector block, whose
// kernels - like NCCL's send/recv - cannot co-reside with the persistent cell CTAs of the other filter lane.
static int allreduce_p2p(dftfe_b200_ctx *ctx, double *buf, size_t count) {
  auto &p = ctx->p2p;
  const int lane = ctx->lane, nr = ctx->nranks, me = ctx->rank;
  const uint32_t seq = ++p.seqAr[lane];
  if (seq > 1)
    for (int q = 0; q < nr; ++q)
      if (q != me) DB_TRY(p2p_wait(ctx, P2P_AR, P2P_ACK, q, seq - 1));
  SignalList sig, ack;
  sig.n = ack.n = 0;
  for (int q = 0; q < nr; ++q) {
    if (q == me) continue;
    DB_CHECK(p.peerSlab[q], "p2p all-reduce: rank %d is not mapped", q);
    ctx->launches += 1;
    DB_CUDA(cudaMemcpyAsync(p.peerSlab[q] + p.peerOffAr[lane][q] + (size_t)me * p.arMax * sizeof(double), buf,
                            count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    sig.addr[sig.n++] = p2p_flag(p.peerSlab[q], p.peerOffFlags[q], nr, lane, P2P_AR, P2P_DATA, me);
    ack.addr[ack.n++] = p2p_flag(p.peerSlab[q], p.peerOffFlags[q], nr, lane, P2P_AR, P2P_ACK, me);
  }
  DB_TRY(launch_signal_flags(ctx, sig, seq));
  LoopbackGroup *inproc = ctx->nccl ? nullptr : loopback_of(ctx);
  if (inproc) inproc->barrier();
  for (int q = 0; q < nr; ++q)
    if (q != me) DB_TRY(p2p_wait(ctx, P2P_AR, P2P_DATA, q, seq));
  ctx->launches += 1;
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((count + 255) / 256, (size_t)ctx->num_sms * 4));
  p2p_sum_slots_kernel<<<grid, 256, 0, ctx->stream>>>(buf, reinterpret_cast<const double *>(p.slab + p.offAr[lane]),
                                                     p.arMax, count, nr, me);
  DB_CUDA(cudaGetLastError());
  DB_TRY(launch_signal_flags(ctx, ack, seq));
  if (inproc) inproc->barrier();
  return 0;
}

int allreduce_sum(dftfe_b200_ctx *ctx, double *buf, size_t count) {
  if (ctx->nranks == 1) return 0;
  if (ctx->p2p.active && ctx->p2p.arMax >= count && count > 0 && ctx->nranks <= P2P_MAX_PEERS)
    return allreduce_p2p(ctx, buf, count);
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

// dftUtils::createBandParallelizationIndices (utils/dftUtils.cc:219-240): group g owns the wavefunction columns
// [g * (N / nGroups), (g + 1) * (N / nGroups)), the last group up to N
void band_group_range(int nGroups, int N, int g, int &lo, int &hi) {
  const int w = N / nGroups;
  lo = g * w;
  hi = (g == nGroups - 1) ? N : (g + 1) * w;
}

// Merge of the band groups after the filter (solver .cc:539-567, pseudoGSDevice.cc:370-452): every group holds its own
// filtered columns of X; afterwards all groups hold all columns.  The reference copies the whole zero-padded X to the
// host and MPI_Allreduce-sums it; here each group's column slice is packed (M x w, contiguous) and broadcast from its
// owner over NCCL (one grouped call), i.e. an all-gather that moves every element once instead of summing zeros.
int band_group_merge(dftfe_b200_ctx *ctx, double *X, int N) {
  const int ng = ctx->nBandGroups;
  if (ng <= 1) return 0;
  LoopbackGroup *grp = band_loopback_of(ctx);
  DB_CHECK(ctx->bandNccl || grp, "band_group_merge needs band_comm_init (NCCL) or a band loopback group");
  DB_CHECK(N / ng >= 1, "band_group_merge: more band groups (%d) than wavefunctions (%d)", ng, N);
  const int cm = ctx->cm;
  const int64_t M = ctx->M;
  DB_TRY(ctx->bandBuf.alloc((size_t)std::max<int64_t>(M, 1) * N * cm));
  std::vector<size_t> off(ng + 1, 0);
  for (int g = 0; g < ng; ++g) {
    int lo, hi;
    band_group_range(ng, N, g, lo, hi);
    off[g + 1] = off[g] + (size_t)M * (hi - lo) * cm;
  }
  int lo, hi;
  band_group_range(ng, N, ctx->bandId, lo, hi);
  double *mine = ctx->bandBuf.p + off[ctx->bandId];
  DB_TRY(launch_block_copy_from_full(ctx, X, N * cm, lo * cm, mine, (hi - lo) * cm, M, nullptr));
  if (ctx->bandNccl) {
    DB_NCCL(nccl_api()->GroupStart());
    for (int g = 0; g < ng; ++g)
      DB_NCCL(nccl_api()->Broadcast(ctx->bandBuf.p + off[g], ctx->bandBuf.p + off[g], off[g + 1] - off[g], ncclDouble, g,
                                    ctx->bandNccl, ctx->stream));
    DB_NCCL(nccl_api()->GroupEnd());
  } else {
    DB_CUDA(cudaStreamSynchronize(ctx->stream));
    grp->pub[ctx->bandId] = mine;
    grp->barrier();
    for (int g = 0; g < ng; ++g)
      if (g != ctx->bandId)
        DB_CUDA(cudaMemcpyAsync(ctx->bandBuf.p + off[g], grp->pub[g], (off[g + 1] - off[g]) * sizeof(double),
                                cudaMemcpyDeviceToDevice, ctx->stream));
    DB_CUDA(cudaStreamSynchronize(ctx->stream));
    grp->barrier();
  }
  for (int g = 0; g < ng; ++g) {
    if (g == ctx->bandId) continue;
    int glo, ghi;
    band_group_range(ng, N, g, glo, ghi);
    DB_TRY(launch_block_copy_to_full(ctx, X, N * cm, glo * cm, ctx->bandBuf.p + off[g], (ghi - glo) * cm, M, nullptr));
  }
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
