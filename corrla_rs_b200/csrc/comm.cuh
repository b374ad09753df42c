// NCCL communicator for the row-sharded RSVD, bound at run time with dlopen so that libcorrla_b200.so
// neither links libnccl nor fights with the copy PyTorch has already loaded into the process.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

struct corrla_comm {
  void* lib = nullptr;
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
  // sum-all-reduce `count` doubles in place on `stream`; returns 0 or a negative corrla_status
  int allreduce_f64(double* buf, size_t count, cudaStream_t stream);
};

namespace corrla {
int comm_unique_id(unsigned char id[128]);
int comm_init(const unsigned char id[128], int rank, int nranks, int device, corrla_comm** out);
void comm_destroy(corrla_comm* c);
void set_last_error(const char* fmt, ...);
const char* last_error_cstr();
}  // namespace corrla
