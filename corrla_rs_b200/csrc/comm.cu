#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/corrla_b200.h"

namespace corrla {

namespace {
thread_local char g_last_error[512] = "";

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    // Prefer the copy already mapped into the process (PyTorch bundles its own libnccl.so.2).
    a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (a.lib == nullptr) a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (a.lib == nullptr) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (a.lib == nullptr) return;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.lib, "ncclAllReduce"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.lib, "ncclAllGather"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString;
  });
  return a;
}
}  // namespace

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

const char* last_error_cstr() { return g_last_error; }

namespace {
// Allocate the symmetric region, exchange CUDA IPC handles through NCCL (all-gather of 64-byte blobs) and map the
// peers.  Any failure leaves c->p2p == false.  Disabled with CORRLA_B200_NO_P2P=1 (A/B measurements).
void setup_peer_memory(corrla_comm* c) {
  NcclApi& a = api();
  const char* off = getenv("CORRLA_B200_NO_P2P");
  if ((off != nullptr && off[0] == '1') || c->nranks < 2 || c->nranks > kMaxPeers || a.AllGather == nullptr) return;
  const size_t bytes = kSymFlagBytes + 2 * kSymHalfDoubles * sizeof(double);
  void* local = nullptr;
  unsigned char* dev_handles = nullptr;
  int ok_local = 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&local, bytes) != cudaSuccess || cudaMemset(local, 0, bytes) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine, local) != cudaSuccess)
    ok_local = 0;
  // every rank must take part in the collectives below even if its own allocation failed
  const size_t hb = sizeof(cudaIpcMemHandle_t) + 8;
  std::vector<unsigned char> host((size_t)c->nranks * hb, 0);
  if (cudaMalloc(&dev_handles, (size_t)(c->nranks + 1) * hb) != cudaSuccess) { cudaGetLastError(); if (local) cudaFree(local); return; }
  unsigned char blob[sizeof(cudaIpcMemHandle_t) + 8];
  memset(blob, 0, sizeof(blob));
  memcpy(blob, &mine, sizeof(mine));
  blob[sizeof(mine)] = (unsigned char)ok_local;
  cudaMemcpy(dev_handles + (size_t)c->nranks * hb, blob, hb, cudaMemcpyHostToDevice);
  ncclResult_t r = a.AllGather(dev_handles + (size_t)c->nranks * hb, dev_handles, hb, ncclUint8,
                               static_cast<ncclComm_t>(c->nccl_comm), nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  bool all_ok = (r == ncclSuccess) && (e == cudaSuccess);
  if (all_ok) cudaMemcpy(host.data(), dev_handles, host.size(), cudaMemcpyDeviceToHost);
  cudaFree(dev_handles);
  for (int p = 0; all_ok && p < c->nranks; ++p) all_ok = host[(size_t)p * hb + sizeof(cudaIpcMemHandle_t)] == 1;
  if (all_ok) {
    for (int p = 0; p < c->nranks; ++p) {
      if (p == c->rank) { c->sym_peer[p] = local; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, host.data() + (size_t)p * hb, sizeof(h));
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); all_ok = false; break; }
      c->sym_peer[p] = ptr;
    }
  }
  // agree on the outcome: sum of "ok" flags must equal nranks
  double* agree = nullptr;
  double hv = all_ok ? 1.0 : 0.0;
  if (cudaMalloc(&agree, 8) == cudaSuccess) {
    cudaMemcpy(agree, &hv, 8, cudaMemcpyHostToDevice);
    a.AllReduce(agree, agree, 1, ncclFloat64, ncclSum, static_cast<ncclComm_t>(c->nccl_comm), nullptr);
    cudaDeviceSynchronize();
    cudaMemcpy(&hv, agree, 8, cudaMemcpyDeviceToHost);
    cudaFree(agree);
  } else { hv = 0.0; }
  cudaGetLastError();
  if (hv == (double)c->nranks) {
    unsigned int* ctr = nullptr;
    if (cudaMalloc(&ctr, 64) == cudaSuccess && cudaMemset(ctr, 0, 64) == cudaSuccess) {
      c->sym_local = local; c->block_counter = ctr; c->err_flag = reinterpret_cast<int*>(ctr + 8); c->p2p = true;
      return;
    }
  }
  for (int p = 0; p < c->nranks; ++p)
    if (p != c->rank && c->sym_peer[p] != nullptr) { cudaIpcCloseMemHandle(c->sym_peer[p]); c->sym_peer[p] = nullptr; }
  if (local) cudaFree(local);
  cudaGetLastError();
}
}  // namespace

int comm_unique_id(unsigned char id[128]) {
  NcclApi& a = api();
  if (!a.ok) { set_last_error("libnccl.so.2 could not be loaded"); return CORRLA_ERR_COMM; }
  ncclUniqueId uid;
  ncclResult_t r = a.GetUniqueId(&uid);
  if (r != ncclSuccess) { set_last_error("ncclGetUniqueId: %s", a.GetErrorString(r)); return CORRLA_ERR_COMM; }
  static_assert(sizeof(uid) == 128, "ncclUniqueId size");
  memcpy(id, &uid, 128);
  return CORRLA_OK;
}

int comm_init(const unsigned char id[128], int rank, int nranks, int device, corrla_comm** out) {
  NcclApi& a = api();
  if (!a.ok) { set_last_error("libnccl.so.2 could not be loaded"); return CORRLA_ERR_COMM; }
  if (out == nullptr || nranks < 1 || rank < 0 || rank >= nranks) return CORRLA_ERR_INVALID;
  if (device >= 0) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { set_last_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e)); return CORRLA_ERR_CUDA; }
  } else {
    cudaGetDevice(&device);
  }
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclComm_t c = nullptr;
  ncclResult_t r = a.CommInitRank(&c, nranks, uid, rank);
  if (r != ncclSuccess) { set_last_error("ncclCommInitRank: %s", a.GetErrorString(r)); return CORRLA_ERR_COMM; }
  corrla_comm* cc = new corrla_comm();
  cc->lib = a.lib; cc->nccl_comm = c; cc->rank = rank; cc->nranks = nranks; cc->device = device;
  if (const char* e = getenv("CORRLA_B200_XCHG_TIMEOUT_CYCLES")) { const long long v = atoll(e); if (v > 0) cc->timeout_cycles = v; }
  setup_peer_memory(cc);      // best effort: on failure the communicator stays NCCL-only
  *out = cc;
  return CORRLA_OK;
}

void comm_destroy(corrla_comm* c) {
  if (c == nullptr) return;
  if (c->sym_local != nullptr) {
    cudaDeviceSynchronize();
    for (int r = 0; r < c->nranks && r < kMaxPeers; ++r)
      if (r != c->rank && c->sym_peer[r] != nullptr) cudaIpcCloseMemHandle(c->sym_peer[r]);
    cudaFree(c->sym_local);
    if (c->block_counter) cudaFree(c->block_counter);
    cudaGetLastError();
  }
  NcclApi& a = api();
  if (a.ok && c->nccl_comm != nullptr) a.CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
  delete c;
}

}  // namespace corrla

int corrla_comm::allreduce_f64(double* buf, size_t count, cudaStream_t stream) {
  if (nranks <= 1 || count == 0) return CORRLA_OK;
  auto& a = corrla::api();
  ncclResult_t r = a.AllReduce(buf, buf, count, ncclFloat64, ncclSum, static_cast<ncclComm_t>(nccl_comm), stream);
  if (r != ncclSuccess) { corrla::set_last_error("ncclAllReduce: %s", a.GetErrorString(r)); return CORRLA_ERR_COMM; }
  return CORRLA_OK;
}

bool corrla_comm::next_exchange(size_t count, corrla::PeerExchange* px) {
  using namespace corrla;
  if (!p2p || nranks < 2 || count > kSymHalfDoubles) return false;
  ++epoch;
  const size_t half = (size_t)(epoch & 1ull) * kSymHalfDoubles * sizeof(double);
  for (int r = 0; r < kMaxPeers; ++r) { px->peer[r] = nullptr; px->peer_flags[r] = nullptr; }
  for (int r = 0; r < nranks; ++r) {
    unsigned char* base = static_cast<unsigned char*>(sym_peer[r]);
    px->peer[r] = reinterpret_cast<const double*>(base + kSymFlagBytes + half);
    px->peer_flags[r] = reinterpret_cast<unsigned long long*>(base);
  }
  unsigned char* mybase = static_cast<unsigned char*>(sym_local);
  px->mine = reinterpret_cast<double*>(mybase + kSymFlagBytes + half);
  px->my_flags = reinterpret_cast<unsigned long long*>(mybase);
  px->block_counter = block_counter;
  px->err = err_flag;
  px->rank = rank; px->nranks = nranks; px->epoch = epoch;
  px->timeout_cycles = timeout_cycles;
  return true;
}
