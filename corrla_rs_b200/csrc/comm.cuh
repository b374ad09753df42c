// Communicator for the row-sharded RSVD (one process per GPU).
//  * NCCL, bound at run time with dlopen so that libcorrla_b200.so neither links libnccl nor fights with the copy
//    PyTorch has already loaded into the process: bootstrap, large or rare all-reduces.
//  * Peer memory over NVLink/NVSwitch: every rank cudaMalloc's one "symmetric" region, exports it with CUDA IPC, and
//    maps everybody else's.  The split-K reduction kernel of the GEMM then finishes the all-reduce itself
//    (reduce_exchange_kernel in skinny_gemm.cu): local partial sums -> my region, one release flag per peer, acquire
//    the peers' flags, sum the P regions in rank order straight out of peer memory.  No NCCL call on that path.
#pragma once
#include <mutex>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace corrla {
constexpr int kMaxPeers = 8;
constexpr size_t kSymFlagBytes = 1024;                 // kMaxPeers epochs (one per source rank), padded
constexpr size_t kSymHalfDoubles = (size_t)1 << 19;    // 4 MiB per half; two halves alternate by epoch parity

// One exchange, by value into the kernel.
struct PeerExchange {
  double* mine;                                        // my half for this epoch
  const double* peer[kMaxPeers];                       // every rank's half for this epoch (peer[rank] == mine)
  unsigned long long* my_flags;                        // [src] = last epoch rank `src` has published, written by src
  unsigned long long* peer_flags[kMaxPeers];           // the same array on every rank
  unsigned int* block_counter;                         // local scratch, zero between launches
  int* err;                                            // local: set to 1 if a peer never showed up (timeout)
  int rank, nranks;
  unsigned long long epoch;
  long long timeout_cycles;                            // bound of the acquire spin (clock64 ticks)
};
}  // namespace corrla

struct corrla_comm {
  void* lib = nullptr;
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
  // peer-memory path
  bool p2p = false;
  void* sym_local = nullptr;
  void* sym_peer[corrla::kMaxPeers] = {};
  unsigned int* block_counter = nullptr;
  int* err_flag = nullptr;
  unsigned long long epoch = 0;
  long long timeout_cycles = 20000000000ll;            // ~10 s; CORRLA_B200_XCHG_TIMEOUT_CYCLES overrides (tests)
  // A communicator is one collective resource (one exchange region, one epoch counter): calls that take it are serialised
  // inside a process by this lock (held for the whole call, taken after the context's).  Across ranks the usual rule of
  // collectives applies: every rank issues its calls on a communicator in the same order.
  std::recursive_mutex call_mu;
  // sum-all-reduce `count` doubles in place on `stream` with NCCL; returns 0 or a negative corrla_status
  int allreduce_f64(double* buf, size_t count, cudaStream_t stream);
  // next exchange descriptor (advances the epoch); false if the peer path is off or `count` does not fit
  bool next_exchange(size_t count, corrla::PeerExchange* px);
};

namespace corrla {
int comm_unique_id(unsigned char id[128]);
int comm_init(const unsigned char id[128], int rank, int nranks, int device, corrla_comm** out);
void comm_destroy(corrla_comm* c);
void set_last_error(const char* fmt, ...);
const char* last_error_cstr();
}  // namespace corrla
