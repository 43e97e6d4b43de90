// am_comm.cu -- the multi-GPU exchange of libam_b200.
//
// The scan shards by contiguous byte ranges with a halo and needs no data-path collective; the only exchange is one
// 64-bit count per rank (-> offsets into the global match list).  It runs over NCCL, resolved at RUN time with
// dlopen("libnccl.so.2"): the drop-in library has no link-time dependency on NCCL, a single-GPU host never loads it,
// and inside a process that already holds an NCCL (torch ships its own) the dynamic loader hands back that copy.
// The header is used for its types only.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "am_api_internal.h"

using namespace am;

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* names[] = {std::getenv("AM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return; }
    auto sym = [&](const char* s) -> void* { void* p = dlsym(api.handle, s); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}

int nccl_fail(ncclResult_t r, const char* what) {
  NcclApi* n = nccl();
  return fail(AM_E_CUDA, std::string(what) + ": NCCL: " + (n->GetErrorString ? n->GetErrorString(r) : "error"));
}
int nccl_ready() {
  NcclApi* n = nccl();
  if (!n->error.empty()) return fail(AM_E_UNSUPPORTED, n->error);
  return AM_OK;
}

}  // namespace

struct am_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = -1;
  unsigned long long* d_tmp = nullptr;    // 8 + 8 * nranks bytes: host-side collectives (am_comm_allreduce_u64)
  unsigned long long* h_tmp = nullptr;    // pinned mirror
};

namespace am {

int comm_check(am_comm* c, int device) {
  if (!c) return fail(AM_E_BADARG, "communicator is null");
  if (c->device != device) return fail(AM_E_BADARG, "communicator and automaton live on different devices");
  return AM_OK;
}
int comm_size(const am_comm* c) { return c->nranks; }

int comm_allgather_u64(am_comm* c, const void* d_send, void* d_recv, cudaStream_t st) {
  if (c->nranks == 1) {
    cudaError_t e = cudaMemcpyAsync(d_recv, d_send, 8, cudaMemcpyDeviceToDevice, st);
    return e == cudaSuccess ? AM_OK : cuda_fail(e, "count copy");
  }
  ncclResult_t r = nccl()->AllGather(d_send, d_recv, 1, ncclUint64, c->comm, st);
  return r == ncclSuccess ? AM_OK : nccl_fail(r, "ncclAllGather");
}

void comm_offsets(const am_comm* c, const uint64_t* counts, am_shard_result* out) {
  uint64_t off = 0, total = 0;
  for (int i = 0; i < c->nranks; i++) { if (i < c->rank) off += counts[i]; total += counts[i]; }
  out->n_local = counts[c->rank]; out->global_offset = off; out->total = total;
}

}  // namespace am

extern "C" {

int am_comm_unique_id(uint8_t* id) {
  if (!id) return fail(AM_E_BADARG, "id is null");
  int rc = nccl_ready(); if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == AM_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  ncclResult_t r = nccl()->GetUniqueId(&u);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  std::memcpy(id, &u, sizeof u);
  return AM_OK;
}

int am_comm_init(int rank, int nranks, const uint8_t* id, int device, am_comm** out) {
  if (!out) return fail(AM_E_BADARG, "out is null");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks || nranks > 64) return fail(AM_E_BADARG, "bad rank / nranks (at most 64 ranks)");
  if (nranks > 1 && !id) return fail(AM_E_BADARG, "id is null");
  if (am_device_count() == 0) return fail(AM_E_NODEVICE, "no sm_100 CUDA device");
  int dev = device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  DeviceGuard g;
  cudaError_t e = g.enter(dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  am_comm* c = new am_comm();
  c->rank = rank; c->nranks = nranks; c->device = dev;
  if (cudaMalloc((void**)&c->d_tmp, 8 + 8 * (size_t)nranks) != cudaSuccess || cudaMallocHost((void**)&c->h_tmp, 8 + 8 * (size_t)nranks) != cudaSuccess) {
    cudaGetLastError(); am_comm_free(c); return fail(AM_E_OOM, "communicator scratch");
  }
  if (nranks > 1) {
    int rc = nccl_ready();
    if (rc) { am_comm_free(c); return rc; }
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    ncclResult_t r = nccl()->CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) { c->comm = nullptr; am_comm_free(c); return nccl_fail(r, "ncclCommInitRank"); }
  }
  *out = c;
  return AM_OK;
}

void am_comm_free(am_comm* c) {
  if (!c) return;
  DeviceGuard g;
  if (c->device >= 0) g.enter(c->device);
  if (c->comm) nccl()->CommDestroy(c->comm);
  if (c->d_tmp) cudaFree(c->d_tmp);
  if (c->h_tmp) cudaFreeHost(c->h_tmp);
  delete c;
}

int am_shard_halo_exchange(am_comm* c, void* dev_buf, uint64_t halo_bytes, uint64_t shard_len, void* stream) {
  if (!c || (!dev_buf && halo_bytes + shard_len > 0)) return fail(AM_E_BADARG, "null argument");
  if (c->nranks == 1 || halo_bytes == 0) return AM_OK;
  if (shard_len < halo_bytes) return fail(AM_E_BADARG, "a shard must hold at least halo_bytes bytes");
  DeviceGuard g;
  cudaError_t e = g.enter(c->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* buf = static_cast<uint8_t*>(dev_buf);
  NcclApi* n = nccl();
  ncclResult_t r = n->GroupStart();
  if (r == ncclSuccess && c->rank + 1 < c->nranks) r = n->Send(buf + shard_len, halo_bytes, ncclUint8, c->rank + 1, c->comm, st);   // last halo_bytes of [halo | shard]
  if (r == ncclSuccess && c->rank > 0) r = n->Recv(buf, halo_bytes, ncclUint8, c->rank - 1, c->comm, st);
  ncclResult_t r2 = n->GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) return nccl_fail(r, "halo exchange");
  e = cudaStreamSynchronize(st);
  return e == cudaSuccess ? AM_OK : cuda_fail(e, "halo exchange");
}

int am_comm_allreduce_u64(am_comm* c, uint64_t* value, int op, void* stream) {
  if (!c || !value) return fail(AM_E_BADARG, "null argument");
  if (op < 0 || op > 2) return fail(AM_E_BADARG, "unknown reduction");
  if (c->nranks == 1) return AM_OK;
  DeviceGuard g;
  cudaError_t e = g.enter(c->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  c->h_tmp[0] = *value;
  e = cudaMemcpyAsync(c->d_tmp, c->h_tmp, 8, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "allreduce");
  const ncclRedOp_t ops[3] = {ncclSum, ncclMax, ncclMin};
  ncclResult_t r = nccl()->AllReduce(c->d_tmp, c->d_tmp, 1, ncclUint64, ops[op], c->comm, st);
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  e = cudaMemcpyAsync(c->h_tmp, c->d_tmp, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "allreduce");
  *value = c->h_tmp[0];
  return AM_OK;
}

}  // extern "C"
