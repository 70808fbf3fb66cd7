// shard_comm.cu -- the exchange step of the lexicon-sharded mode (mode 2) inside the library.
//
// Every rank holds 1/N of the anagram keys and has scored the WHOLE query batch against its shard (FINISH_SHARD: per
// query its survivors above the score threshold, unranked, with raw frequencies and global gather ids).  Here the N
// exports are exchanged over NCCL (NVLink / NVSwitch) and merged on every rank:
//   1. one 8-byte all-gather: survivor records and longest survivor list per rank (sizes the receive buffers)
//   2. ONE grouped collective: per rank four broadcasts (headers, flags, records, gather ids) with its exact sizes --
//      an all-gather with per-rank counts; nothing is padded on the wire
//   3. merge_kernel ranks the union per query with the GLOBAL max frequency (frequency normalisation is global,
//      src/lib.rs:1460,1521-1525), crop and cut-off follow, then the export stage
// All survivors travel, not a per-shard top-K': the reference's crop keeps a prefix of the ranked list whose length
// depends on score ties across the whole list (src/lib.rs:1536-1589), so no fixed per-shard K' is exact.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: in a torch process that is the copy torch already loaded), so
// the library itself has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "engine.h"
#include "kernel_common.cuh"

namespace anl {

namespace {
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  void* handle = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_m;

bool nccl_load(std::string* err) {
  std::lock_guard<std::mutex> lk(g_nccl_m);
  if (g_nccl.handle) return true;
  void* h = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    *err = std::string("NCCL is not available (dlopen libnccl.so.2: ") + dlerror() + ")";
    return false;
  }
  NcclApi a;
  a.handle = h;
#define ANL_SYM(field, sym)                                                     \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, #sym));                \
  if (!a.field) {                                                               \
    *err = "NCCL symbol " #sym " not found";                                    \
    return false;                                                               \
  }
  ANL_SYM(GetUniqueId, ncclGetUniqueId)
  ANL_SYM(CommInitRank, ncclCommInitRank)
  ANL_SYM(CommDestroy, ncclCommDestroy)
  ANL_SYM(AllGather, ncclAllGather)
  ANL_SYM(Broadcast, ncclBroadcast)
  ANL_SYM(GroupStart, ncclGroupStart)
  ANL_SYM(GroupEnd, ncclGroupEnd)
  ANL_SYM(GetErrorString, ncclGetErrorString)
#undef ANL_SYM
  g_nccl = a;
  return true;
}
}  // namespace

#define NCCL_TRY(expr)                                                                        \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      *err = std::string("NCCL error: ") + g_nccl.GetErrorString(_r) + " at " #expr;          \
      return false;                                                                           \
    }                                                                                         \
  } while (0)
#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      *err = std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr;          \
      return false;                                                                        \
    }                                                                                      \
  } while (0)

bool shard_unique_id(uint8_t* id128, std::string* err) {
  if (!nccl_load(err)) return false;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  static_assert(sizeof id == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, sizeof id);
  return true;
}

struct ShardComm {
  ncclComm_t comm = nullptr;
  int rank = 0, n_ranks = 1;
  // receive side, grow-only
  OutHead* heads_all = nullptr;
  uint32_t* flags_all = nullptr;
  OutRec* recs_all = nullptr;
  uint32_t* gids_all = nullptr;
  size_t cap_queries = 0, cap_records = 0;  // per rank
  uint32_t* d_sizes = nullptr;              // [n_ranks][2] + own [2]
  uint32_t* h_sizes = nullptr;              // pinned
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

void Engine::shard_comm_free() {
  if (!shard_comm_) return;
  ShardComm* c = shard_comm_;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (void* p : {(void*)c->heads_all, (void*)c->flags_all, (void*)c->recs_all, (void*)c->gids_all, (void*)c->d_sizes})
    if (p) cudaFree(p);
  if (c->h_sizes) cudaFreeHost(c->h_sizes);
  for (cudaEvent_t e : c->ev)
    if (e) cudaEventDestroy(e);
  delete c;
  shard_comm_ = nullptr;
}

bool Engine::shard_comm_init(const uint8_t* id128, int rank, int n_ranks, std::string* err) {
  if (!uploaded()) {
    *err = "model has not been built";
    return false;
  }
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || (uint32_t)n_ranks != hm_->index.n_shards || (uint32_t)rank != hm_->index.shard) {
    *err = "communicator rank / size must equal the shard coordinates the index was built with";
    return false;
  }
  if (!nccl_load(err)) return false;
  CUDA_TRY(cudaSetDevice(device_));
  shard_comm_free();
  ShardComm* c = new ShardComm();
  shard_comm_ = c;
  c->rank = rank;
  c->n_ranks = n_ranks;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NCCL_TRY(g_nccl.CommInitRank(&c->comm, n_ranks, id, rank));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_sizes), ((size_t)n_ranks + 1) * 2 * sizeof(uint32_t)));
  CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&c->h_sizes), ((size_t)n_ranks + 1) * 2 * sizeof(uint32_t)));
  for (cudaEvent_t& e : c->ev) CUDA_TRY(cudaEventCreate(&e));
  return true;
}

// longest survivor list of the batch -> sizes[1] (sizes[0] = records, from the pool cursor)
__global__ void shard_sizes_kernel(uint32_t n, const OutHead* __restrict__ head, const unsigned int* __restrict__ pool_cursor,
                                   uint32_t* __restrict__ sizes) {
  uint32_t mx = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mx = max(mx, head[i].count);
  mx = __reduce_max_sync(0xFFFFFFFFu, mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(sizes + 1, mx);
  if (blockIdx.x == 0 && threadIdx.x == 0) sizes[0] = *pool_cursor;
}

bool Engine::shard_step(DeviceBatch* b, ResultSet* out, ShardStepStats* stats, std::string* err, int* status) {
  *status = ANL_ERR_CUDA;
  ShardComm* c = shard_comm_;
  if (!c) {
    *err = "no shard communicator: call anl_shard_comm_init first";
    *status = ANL_ERR_INVALID;
    return false;
  }
  if (!b->sharded) {
    *err = "not a batch of a sharded model";
    *status = ANL_ERR_INVALID;
    return false;
  }
  CUDA_TRY(cudaSetDevice(device_));
  const uint32_t n = b->n;
  const int N = c->n_ranks, me = c->rank;
  cudaStream_t st = b->stream;
  CUDA_TRY(cudaEventRecord(c->ev[0], st));
  if (!run_batch(b, nullptr, err)) return false;
  if (!settle(b, err, status)) return false;  // pool / staged-node queue overflows are repaired here
  *status = ANL_ERR_CUDA;
  CUDA_TRY(cudaEventRecord(c->ev[1], st));
  // ---- sizes: records and longest list per rank ------------------------------------------------------------
  uint32_t* own = c->d_sizes + (size_t)N * 2;
  CUDA_TRY(cudaMemsetAsync(own, 0, 2 * sizeof(uint32_t), st));
  if (n) {
    shard_sizes_kernel<<<148, 256, 0, st>>>(n, b->d_head, b->d_work + 2, own);
    count_launch(1);
  }
  NCCL_TRY(g_nccl.AllGather(own, c->d_sizes, 2, ncclUint32, c->comm, st));
  CUDA_TRY(cudaMemcpyAsync(c->h_sizes, c->d_sizes, (size_t)N * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  uint64_t stride = 1, max_surv = 0, total_all = 0;
  for (int r = 0; r < N; ++r) {
    stride = std::max<uint64_t>(stride, c->h_sizes[2 * r]);
    max_surv += c->h_sizes[2 * r + 1];
    total_all += c->h_sizes[2 * r];
  }
  if (c->h_sizes[2 * me] > b->cap_pool) {
    *err = "internal error: shard export larger than its pool";
    return false;
  }
  // ---- receive buffers (grow-only) ----------------------------------------------------------------------------
  if (n > c->cap_queries) {
    if (c->heads_all) cudaFree(c->heads_all);
    if (c->flags_all) cudaFree(c->flags_all);
    c->heads_all = nullptr;
    c->flags_all = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->heads_all), (size_t)N * n * sizeof(OutHead)));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->flags_all), (size_t)N * n * sizeof(uint32_t)));
    c->cap_queries = n;
  }
  if (stride > c->cap_records) {
    if (c->recs_all) cudaFree(c->recs_all);
    if (c->gids_all) cudaFree(c->gids_all);
    c->recs_all = nullptr;
    c->gids_all = nullptr;
    const size_t cap = (size_t)stride + stride / 4 + 1024;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->recs_all), (size_t)N * cap * sizeof(OutRec)));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->gids_all), (size_t)N * cap * sizeof(uint32_t)));
    c->cap_records = cap;
  }
  const size_t qstride = c->cap_queries, rstride = c->cap_records;
  // ---- the exchange: one grouped collective, exact sizes -------------------------------------------------------
  uint64_t received = 0;
  NCCL_TRY(g_nccl.GroupStart());
  for (int r = 0; r < N; ++r) {
    const size_t recs = c->h_sizes[2 * r];
    if (n) {
      NCCL_TRY(g_nccl.Broadcast(b->d_head, c->heads_all + (size_t)r * qstride, (size_t)n * sizeof(OutHead), ncclUint8, r, c->comm, st));
      NCCL_TRY(g_nccl.Broadcast(b->d_qflags, c->flags_all + (size_t)r * qstride, (size_t)n * sizeof(uint32_t), ncclUint8, r, c->comm, st));
    }
    if (recs) {
      NCCL_TRY(g_nccl.Broadcast(b->d_out, c->recs_all + (size_t)r * rstride, recs * sizeof(OutRec), ncclUint8, r, c->comm, st));
      NCCL_TRY(g_nccl.Broadcast(b->d_gid, c->gids_all + (size_t)r * rstride, recs * sizeof(uint32_t), ncclUint8, r, c->comm, st));
    }
    if (r != me) received += (uint64_t)n * (sizeof(OutHead) + sizeof(uint32_t)) + (uint64_t)recs * (sizeof(OutRec) + sizeof(uint32_t));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  CUDA_TRY(cudaEventRecord(c->ev[2], st));
  // ---- merge + export ---------------------------------------------------------------------------------------------
  // (heads / flags of rank r start at r * qstride: shard_merge takes one stride for both kinds of array, so the
  // per-query arrays are addressed through their own stride argument)
  const bool ok = shard_merge_strided(b, (uint32_t)N, c->heads_all, c->recs_all, c->gids_all, c->flags_all, qstride, rstride,
                                      (uint32_t)std::min<uint64_t>(max_surv, 0xFFFFFFF0u), out, err, status);
  if (!ok) return false;
  *status = ANL_ERR_CUDA;
  CUDA_TRY(cudaEventRecord(c->ev[3], st));
  CUDA_TRY(cudaEventSynchronize(c->ev[3]));
  if (stats) {
    float a = 0, x = 0, m = 0;
    cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&x, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&m, c->ev[2], c->ev[3]);
    stats->score_ms = a;
    stats->exchange_ms = x;
    stats->merge_ms = m;
    stats->bytes_received = received;
    stats->records_local = c->h_sizes[2 * me];
    stats->records_total = total_all;
  }
  *status = ANL_OK;
  return true;
}

}  // namespace anl
