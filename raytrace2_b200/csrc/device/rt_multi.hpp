// The object behind an rt2_renderer handle: one wavefront renderer (device/rt_render.hpp) per GPU of the box, all driven by
// the calling host thread ("handle owns CUDA streams, device buffers ... for all GPUs of the box (single process)").
//
// Replaces the only parallel region of the reference, the pixel-parallel std::for_each of RayTracer::Update
// (src/cpu_raytrace/RayTracer.cpp:69), at the granularity that suits 8 GPUs: whole frames.  Global frame k (the k-th Update
// since the last Reset) is traced by replica k mod N; its stratum and Philox counters depend only on k, so the partition
// never changes which samples are drawn.  Kernel launches are asynchronous, so one host thread keeps every GPU busy: a batch
// is ~150 launches (~0.5 ms of host time) for ~100 ms of GPU work.
//
// Read-out (RayTracer::NonConvertedPixels / Pixels, RayTracer.cpp:16-18,105-112): the first replica's stream waits for an
// event on every other replica's stream, then ONE kernel on the first GPU sums the accumulators in replica order with peer
// loads over NVLink (cudaDeviceEnablePeerAccess — plain pointers, no IPC inside one process), divides by the total frame
// count and writes the mean / RGBA8 preview (k_resolve_peers).  Without a peer path the accumulators are first copied to
// the first GPU (cudaMemcpyPeerAsync) and the same kernel runs on the copies.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "rt_render.hpp"

namespace rt2 {

class MultiRenderer {
 public:
  MultiRenderer() = default;
  ~MultiRenderer();
  MultiRenderer(const MultiRenderer&) = delete;
  MultiRenderer& operator=(const MultiRenderer&) = delete;

  int Init(const HostScene& scene, const rt2_config& cfg);
  int UploadScene(const HostScene& scene);
  int Resize(int w, int h);
  int Reset();
  int Update(uint32_t n_frames);
  int Flush();
  int Synchronize();
  int ReadMean(float* dst);
  int ReadRGBA8(uint8_t* dst);
  int ReadAccum(float* sum, float* sumsq);
  int WriteAccum(const float* sum, const float* sumsq, uint64_t frames);
  int GetStats(rt2_stats* out);
  void SetProfiling(bool on);
  uint64_t FrameIdx() const { return frames_; }
  int Width() const { return reps_.empty() ? 0 : reps_[0]->Width(); }
  int Height() const { return reps_.empty() ? 0 : reps_[0]->Height(); }
  size_t Replicas() const { return reps_.size(); }
  // single-GPU handles only (external reduce / IPC plumbing of the one-process-per-GPU harness); first replica otherwise
  Renderer* Single(const char* what);
  Renderer& First() { return *reps_[0]; }
  int SetFrameIdx(uint64_t frames);
  const std::string& Error() const { return err_; }

 private:
  int Fail(size_t g, int rc);
  int Resolve(float* dst_mean, uint8_t* dst_rgba8);
  uint64_t CountOf(size_t g, uint64_t frames) const;  // frames k < `frames` with k mod N == g
  std::vector<std::unique_ptr<Renderer>> reps_;
  rt2_config cfg_{};
  uint64_t frames_{0};  // global frames requested since the last reset
  bool peer_ok_{true};  // the first GPU can address the memory of every other one
  std::vector<void*> staged_;  // no peer path: copies of the peers' accumulators on the first GPU
  size_t staged_bytes_{0};
  std::string err_;
};

}  // namespace rt2
