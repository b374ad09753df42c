// Host <-> device copies that keep PCIe busy when the host side is ordinary pageable memory (numpy arrays):
// double-buffered pinned bounce buffers, the pageable side moved by a few CPU threads while the DMA of the
// neighbouring chunk is in flight.  Pinned / registered host memory takes the direct cudaMemcpy2DAsync path.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

#include <thread>
#include <vector>

namespace corrla {

struct BounceBuffers {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  bool in_flight[2] = {false, false};   // a DMA out of this slot may still be running (cleared by waiting on done[])
  size_t bytes = 0;
  int threads = 4;
  cudaError_t ensure(size_t want);
  void release();
};

// rows x row_bytes block, pitches in bytes.  With sync_at_end the call returns when the data is on the device; without
// it the last DMA may still be in flight on `st` (the source must stay valid; bounce slots are tracked across calls).
cudaError_t copy_h2d_2d(BounceBuffers& bb, cudaStream_t st, void* dst_dev, size_t dst_pitch, const void* src_host,
                        size_t src_pitch, size_t row_bytes, size_t rows, bool sync_at_end = true);
cudaError_t copy_d2h_2d(BounceBuffers& bb, cudaStream_t st, void* dst_host, size_t dst_pitch, const void* src_dev,
                        size_t src_pitch, size_t row_bytes, size_t rows);

// Fault in the pages of a pageable host OUTPUT buffer on helper threads while the GPU is still computing, so that the
// device->host copy at the end of the call runs at DMA speed instead of page-fault speed (a fresh 3 GB numpy array
// costs ~60 ms of first-touch faults).  Writes zeros: the buffer is about to be overwritten with the results.
struct PreTouch {
  std::vector<std::thread> th;
  void add(void* p, size_t bytes);     // no-op for small, null or pinned buffers
  void join();
  ~PreTouch() { join(); }
};

}  // namespace corrla
