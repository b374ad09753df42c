#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/corrla_b200.h"

namespace corrla {

namespace {
thread_local char g_last_error[512] = "";

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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
  *out = cc;
  return CORRLA_OK;
}

void comm_destroy(corrla_comm* c) {
  if (c == nullptr) return;
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
