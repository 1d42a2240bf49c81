// Host -> device copies of large PAGEABLE buffers (the drop-in's normal case: a Go slice is ordinary memory).
//
// cudaMemcpyAsync from pageable memory is staged by the driver through one bounce buffer on the calling thread
// (~10 GB/s measured: the 2^24 MSM ran at 287 Mpts/s end to end from pageable scalars against 423 from pinned ones,
// and a 2^24 prove took 359 ms against 171).  h2d_copy() does the staging itself: the transfer is cut into stripes,
// one short-lived host thread per stripe copies 4 MiB chunks into its own pair of pinned slots and issues the DMA
// from there on its own stream — memcpy and DMA of different stripes overlap, and the host side runs at the memory
// system's rate instead of one core's.  Pinned / registered sources and small transfers go straight to
// cudaMemcpyAsync.  Like cudaMemcpyAsync on pageable memory, the call returns once `src` has been read completely.
// d2h_copy() is the mirror image for large results (NTT / computeH / Keccak outputs).
#pragma once
#include <thread>

#include "common.cuh"

namespace b200 {

constexpr size_t H2D_STAGE_MIN = (size_t)16 << 20;   // below this the driver's own staging is as good
constexpr size_t H2D_SLOT = (size_t)4 << 20;

inline bool host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  const bool pinned = cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost;
  cudaGetLastError();   // an unregistered pointer is not an error worth keeping
  return pinned;
}

inline int h2d_stager_init(b200g16_ctx* ctx) {
  H2DStager& s = ctx->stager;
  if (s.pinned) return 0;
  unsigned hw = std::thread::hardware_concurrency();
  s.workers = (int)(hw / 2 < 2 ? 2 : (hw / 2 > H2D_MAX_WORKERS ? H2D_MAX_WORKERS : hw / 2));
  B200_CUDA(cudaHostAlloc(&s.pinned, (size_t)s.workers * 2 * H2D_SLOT, cudaHostAllocDefault));
  for (int w = 0; w < s.workers; w++) {
    B200_CUDA(cudaStreamCreateWithFlags(&s.stream[w], cudaStreamNonBlocking));
    B200_CUDA(cudaEventCreateWithFlags(&s.done[w], cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) B200_CUDA(cudaEventCreateWithFlags(&s.slot_free[w][k], cudaEventDisableTiming));
  }
  B200_CUDA(cudaEventCreateWithFlags(&s.start, cudaEventDisableTiming));
  return 0;
}

inline void h2d_stager_release(b200g16_ctx* ctx) {
  H2DStager& s = ctx->stager;
  if (!s.pinned) return;
  for (int w = 0; w < s.workers; w++) {
    cudaStreamDestroy(s.stream[w]);
    cudaEventDestroy(s.done[w]);
    for (int k = 0; k < 2; k++) cudaEventDestroy(s.slot_free[w][k]);
  }
  cudaEventDestroy(s.start);
  cudaFreeHost(s.pinned);
  s.pinned = nullptr;
}

// dst (device) <- src (host), ordered on `stream` like a cudaMemcpyAsync issued there.
inline int h2d_copy(b200g16_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return 0;
  if (bytes < H2D_STAGE_MIN || host_is_pinned(src) || h2d_stager_init(ctx) != 0) {   // (no pinned slots: driver's staging)
    B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return 0;
  }
  H2DStager& s = ctx->stager;
  B200_CUDA(cudaEventRecord(s.start, stream));   // dst may still be read by work already on `stream`
  const int W = s.workers;
  const size_t chunks = (bytes + H2D_SLOT - 1) / H2D_SLOT;
  cudaError_t err[H2D_MAX_WORKERS];
  std::thread th[H2D_MAX_WORKERS];
  for (int w = 0; w < W; w++) {
    err[w] = cudaSuccess;
    th[w] = std::thread([&, w]() {
      cudaError_t e = cudaSetDevice(ctx->device);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s.stream[w], s.start, 0);
      char* slots = (char*)s.pinned + (size_t)w * 2 * H2D_SLOT;
      const size_t c0 = chunks * (size_t)w / (size_t)W, c1 = chunks * (size_t)(w + 1) / (size_t)W;   // this worker's stripe
      for (size_t c = c0; c < c1 && e == cudaSuccess; c++) {
        const int k = (int)((c - c0) & 1);
        const size_t off = c * H2D_SLOT, len = bytes - off < H2D_SLOT ? bytes - off : H2D_SLOT;
        e = cudaEventSynchronize(s.slot_free[w][k]);   // the DMA that last read this slot (first use: nothing recorded)
        if (e != cudaSuccess) break;
        memcpy(slots + (size_t)k * H2D_SLOT, (const char*)src + off, len);
        e = cudaMemcpyAsync((char*)dst + off, slots + (size_t)k * H2D_SLOT, len, cudaMemcpyHostToDevice, s.stream[w]);
        if (e == cudaSuccess) e = cudaEventRecord(s.slot_free[w][k], s.stream[w]);
      }
      if (e == cudaSuccess) e = cudaEventRecord(s.done[w], s.stream[w]);
      err[w] = e;
    });
  }
  for (int w = 0; w < W; w++) th[w].join();
  for (int w = 0; w < W; w++) {
    if (err[w] != cudaSuccess) return fail(B200G16_ERR_CUDA, "h2d_copy: %s", cudaGetErrorString(err[w]));
    B200_CUDA(cudaStreamWaitEvent(stream, s.done[w], 0));
  }
  return 0;
}

// dst (host) <- src (device), ordered after the work already on `stream`; returns when dst holds the data (like a
// cudaMemcpy to pageable memory, but at the memory system's rate: DMA into pinned slots, host threads copy out).
inline int d2h_copy(b200g16_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return 0;
  if (bytes < H2D_STAGE_MIN || host_is_pinned(dst) || h2d_stager_init(ctx) != 0) {
    B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  H2DStager& s = ctx->stager;
  B200_CUDA(cudaEventRecord(s.start, stream));
  const int W = s.workers;
  const size_t chunks = (bytes + H2D_SLOT - 1) / H2D_SLOT;
  cudaError_t err[H2D_MAX_WORKERS];
  std::thread th[H2D_MAX_WORKERS];
  for (int w = 0; w < W; w++) {
    err[w] = cudaSuccess;
    th[w] = std::thread([&, w]() {
      cudaError_t e = cudaSetDevice(ctx->device);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s.stream[w], s.start, 0);
      char* slots = (char*)s.pinned + (size_t)w * 2 * H2D_SLOT;
      const size_t c0 = chunks * (size_t)w / (size_t)W, c1 = chunks * (size_t)(w + 1) / (size_t)W;
      auto span = [&](size_t c, size_t* off) { *off = c * H2D_SLOT; return bytes - *off < H2D_SLOT ? bytes - *off : H2D_SLOT; };
      // two DMAs in flight: chunk c + 1 lands in the other slot while chunk c is copied out
      size_t off, len;
      if (c0 < c1 && e == cudaSuccess) {
        len = span(c0, &off);
        e = cudaMemcpyAsync(slots, (const char*)src + off, len, cudaMemcpyDeviceToHost, s.stream[w]);
        if (e == cudaSuccess) e = cudaEventRecord(s.slot_free[w][0], s.stream[w]);
      }
      for (size_t c = c0; c < c1 && e == cudaSuccess; c++) {
        const int k = (int)((c - c0) & 1);
        if (c + 1 < c1) {
          size_t off2;
          const size_t len2 = span(c + 1, &off2);
          e = cudaMemcpyAsync(slots + (size_t)(k ^ 1) * H2D_SLOT, (const char*)src + off2, len2, cudaMemcpyDeviceToHost, s.stream[w]);
          if (e == cudaSuccess) e = cudaEventRecord(s.slot_free[w][k ^ 1], s.stream[w]);
          if (e != cudaSuccess) break;
        }
        e = cudaEventSynchronize(s.slot_free[w][k]);   // chunk c has landed
        if (e != cudaSuccess) break;
        len = span(c, &off);
        memcpy((char*)dst + off, slots + (size_t)k * H2D_SLOT, len);
      }
      err[w] = e;
    });
  }
  for (int w = 0; w < W; w++) th[w].join();
  for (int w = 0; w < W; w++)
    if (err[w] != cudaSuccess) return fail(B200G16_ERR_CUDA, "d2h_copy: %s", cudaGetErrorString(err[w]));
  return 0;
}

}  // namespace b200
