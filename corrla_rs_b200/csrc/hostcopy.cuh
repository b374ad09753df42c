// Host <-> device copies that keep PCIe busy when the host side is ordinary pageable memory (numpy arrays):
// double-buffered pinned bounce buffers, the pageable side moved by a few CPU threads while the DMA of the
// neighbouring chunk is in flight.  Pinned / registered host memory takes the direct cudaMemcpy2DAsync path.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace corrla {

struct BounceBuffers {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  size_t bytes = 0;
  int threads = 4;
  cudaError_t ensure(size_t want);
  void release();
};

// rows x row_bytes block, pitches in bytes.  Synchronous with respect to the host on return.
cudaError_t copy_h2d_2d(BounceBuffers& bb, cudaStream_t st, void* dst_dev, size_t dst_pitch, const void* src_host,
                        size_t src_pitch, size_t row_bytes, size_t rows);
cudaError_t copy_d2h_2d(BounceBuffers& bb, cudaStream_t st, void* dst_host, size_t dst_pitch, const void* src_dev,
                        size_t src_pitch, size_t row_bytes, size_t rows);

}  // namespace corrla
