#include "hostcopy.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace corrla {

namespace {
constexpr size_t kChunkBytes = (size_t)256 << 20;

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// copy `rows` rows of row_bytes between pitched host buffers using up to `threads` threads
void par_copy_2d(char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t row_bytes, size_t rows,
                 int threads) {
  const size_t total = row_bytes * rows;
  int nt = (int)std::min<size_t>((size_t)std::max(threads, 1), std::max<size_t>(1, total >> 20));
  auto work = [&](size_t r0, size_t r1) {
    if (dst_pitch == row_bytes && src_pitch == row_bytes) {
      memcpy(dst + r0 * row_bytes, src + r0 * row_bytes, (r1 - r0) * row_bytes);
    } else {
      for (size_t r = r0; r < r1; ++r) memcpy(dst + r * dst_pitch, src + r * src_pitch, row_bytes);
    }
  };
  if (nt <= 1 || rows < (size_t)nt) {
    if (rows >= 1 && nt > 1 && dst_pitch == row_bytes && src_pitch == row_bytes) {
      // few long rows: split the flat range instead
      std::vector<std::thread> th;
      const size_t per = (total + nt - 1) / nt;
      for (int t = 0; t < nt; ++t) {
        const size_t b0 = std::min(total, (size_t)t * per), b1 = std::min(total, b0 + per);
        if (b1 > b0) th.emplace_back([=] { memcpy(dst + b0, src + b0, b1 - b0); });
      }
      for (auto& x : th) x.join();
      return;
    }
    work(0, rows);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (rows + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const size_t r0 = std::min(rows, (size_t)t * per), r1 = std::min(rows, r0 + per);
    if (r1 > r0) th.emplace_back(work, r0, r1);
  }
  for (auto& x : th) x.join();
}
}  // namespace

cudaError_t BounceBuffers::ensure(size_t want) {
  if (bytes >= want && buf[0] != nullptr) return cudaSuccess;
  release();
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaMallocHost(&buf[i], want);
    if (e != cudaSuccess) { release(); return e; }
    e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { release(); return e; }
  }
  bytes = want;
  in_flight[0] = in_flight[1] = false;
  // the staging copies share the host cores with the other ranks of the node: up to 16 threads, but no more than this
  // process's share (torchrun exports LOCAL_WORLD_SIZE; CORRLA_B200_COPY_THREADS overrides)
  unsigned hc = std::thread::hardware_concurrency();
  if (hc == 0) hc = 4;
  unsigned share = hc;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int lw = atoi(e); if (lw > 1) share = std::max(1u, hc / (unsigned)lw); }
  threads = (int)std::max(1u, std::min(16u, share));
  if (const char* e = getenv("CORRLA_B200_COPY_THREADS")) { const int v = atoi(e); if (v > 0) threads = std::min(v, 64); }
  return cudaSuccess;
}

void BounceBuffers::release() {
  for (int i = 0; i < 2; ++i) {
    if (in_flight[i] && done[i]) cudaEventSynchronize(done[i]);
    in_flight[i] = false;
    if (buf[i]) cudaFreeHost(buf[i]);
    if (done[i]) cudaEventDestroy(done[i]);
    buf[i] = nullptr; done[i] = nullptr;
  }
  bytes = 0;
}

void PreTouch::add(void* p, size_t bytes) {
  if (p == nullptr || bytes < ((size_t)64 << 20) || is_pinned(p)) return;
  const int nt = 4;
  const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
  for (int t = 0; t < nt; ++t) {
    const size_t b0 = std::min(bytes, (size_t)t * per), b1 = std::min(bytes, b0 + per);
    if (b1 <= b0) continue;
    volatile char* base = static_cast<volatile char*>(p);
    th.emplace_back([base, b0, b1] { for (size_t off = b0; off < b1; off += 4096) base[off] = 0; });
  }
}

void PreTouch::join() {
  for (auto& t : th) if (t.joinable()) t.join();
  th.clear();
}

cudaError_t copy_h2d_2d(BounceBuffers& bb, cudaStream_t st, void* dst_dev, size_t dst_pitch, const void* src_host,
                        size_t src_pitch, size_t row_bytes, size_t rows, bool sync_at_end) {
  if (rows == 0 || row_bytes == 0) return cudaSuccess;
  cudaError_t e;
  if (is_pinned(src_host)) {
    e = cudaMemcpy2DAsync(dst_dev, dst_pitch, src_host, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    return sync_at_end ? cudaStreamSynchronize(st) : cudaSuccess;
  }
  const size_t total = row_bytes * rows;
  const size_t chunk = std::min(kChunkBytes, std::max<size_t>(total, 4096));
  if ((e = bb.ensure(std::max(chunk, row_bytes))) != cudaSuccess) return e;
  const size_t rows_per = std::max<size_t>(1, bb.bytes / row_bytes);
  int slot = 0;
  for (size_t r = 0; r < rows; r += rows_per, slot ^= 1) {
    const size_t nr = std::min(rows_per, rows - r);
    if (bb.in_flight[slot]) {
      if ((e = cudaEventSynchronize(bb.done[slot])) != cudaSuccess) return e;
      bb.in_flight[slot] = false;
    }
    par_copy_2d(static_cast<char*>(bb.buf[slot]), row_bytes, static_cast<const char*>(src_host) + r * src_pitch,
                src_pitch, row_bytes, nr, bb.threads);
    e = cudaMemcpy2DAsync(static_cast<char*>(dst_dev) + r * dst_pitch, dst_pitch, bb.buf[slot], row_bytes, row_bytes, nr,
                          cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventRecord(bb.done[slot], st)) != cudaSuccess) return e;
    bb.in_flight[slot] = true;
  }
  if (!sync_at_end) return cudaSuccess;
  bb.in_flight[0] = bb.in_flight[1] = false;
  return cudaStreamSynchronize(st);
}

cudaError_t copy_d2h_2d(BounceBuffers& bb, cudaStream_t st, void* dst_host, size_t dst_pitch, const void* src_dev,
                        size_t src_pitch, size_t row_bytes, size_t rows) {
  if (rows == 0 || row_bytes == 0) return cudaSuccess;
  cudaError_t e;
  if (is_pinned(dst_host)) {
    e = cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, row_bytes, rows, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
  }
  // long contiguous blocks (one "row" of gigabytes) are cut into pseudo-rows so that they can be chunked
  if (rows == 1 || (dst_pitch == row_bytes && src_pitch == row_bytes)) {
    const size_t total = row_bytes * rows;
    const size_t pr = (size_t)1 << 20;
    if (total > pr && total % pr == 0) { row_bytes = pr; rows = total / pr; dst_pitch = src_pitch = pr; }
    else if (total > pr) {
      const size_t head = total / pr * pr;
      e = copy_d2h_2d(bb, st, dst_host, pr, src_dev, pr, pr, head / pr);
      if (e != cudaSuccess) return e;
      return copy_d2h_2d(bb, st, static_cast<char*>(dst_host) + head, total - head,
                         static_cast<const char*>(src_dev) + head, total - head, total - head, 1);
    }
  }
  const size_t total = row_bytes * rows;
  const size_t chunk = std::min(kChunkBytes, std::max<size_t>(total, 4096));
  if ((e = bb.ensure(std::max(chunk, row_bytes))) != cudaSuccess) return e;
  for (int i = 0; i < 2; ++i)
    if (bb.in_flight[i]) { if ((e = cudaEventSynchronize(bb.done[i])) != cudaSuccess) return e; bb.in_flight[i] = false; }
  const size_t rows_per = std::max<size_t>(1, bb.bytes / row_bytes);
  const size_t nchunks = (rows + rows_per - 1) / rows_per;
  auto issue = [&](size_t ci) -> cudaError_t {
    const size_t r = ci * rows_per, nr = std::min(rows_per, rows - r);
    const int slot = (int)(ci & 1);
    cudaError_t ee = cudaMemcpy2DAsync(bb.buf[slot], row_bytes, static_cast<const char*>(src_dev) + r * src_pitch,
                                       src_pitch, row_bytes, nr, cudaMemcpyDeviceToHost, st);
    if (ee != cudaSuccess) return ee;
    return cudaEventRecord(bb.done[slot], st);
  };
  if ((e = issue(0)) != cudaSuccess) return e;
  for (size_t ci = 0; ci < nchunks; ++ci) {
    const int slot = (int)(ci & 1);
    if ((e = cudaEventSynchronize(bb.done[slot])) != cudaSuccess) return e;
    if (ci + 1 < nchunks && (e = issue(ci + 1)) != cudaSuccess) return e;   // other slot: free since chunk ci-1 was drained
    const size_t r = ci * rows_per, nr = std::min(rows_per, rows - r);
    par_copy_2d(static_cast<char*>(dst_host) + r * dst_pitch, dst_pitch, static_cast<const char*>(bb.buf[slot]), row_bytes,
                row_bytes, nr, bb.threads);
  }
  return cudaSuccess;
}

}  // namespace corrla
