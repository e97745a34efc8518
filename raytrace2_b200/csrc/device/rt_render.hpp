// Host-side handle of the wavefront renderer (implementation in rt_kernels.cu).  Mirrors the state of
// raytrace2::cpu::RayTracer (src/cpu_raytrace/RayTracer.hpp:15-42): dims, frame index, accumulation buffer.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/rt2.h"
#include "../host/scene_host.hpp"

namespace rt2 {

class Renderer {
 public:
  struct Impl;
  Renderer();
  ~Renderer();
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  int Init(const HostScene& scene, const rt2_config& cfg);
  int UploadScene(const HostScene& scene);
  int Resize(int w, int h);          // RayTracer::OnResize
  int Reset();                       // RayTracer::Reset
  int Update(uint32_t n_frames);     // n x RayTracer::Update (frames are traced when a wavefront batch is full, or by Flush)
  int Flush();                       // trace every frame requested so far (asynchronous on the stream)
  int Synchronize();
  int ReadMean(float* dst);          // RayTracer::NonConvertedPixels
  int ReadRGBA8(uint8_t* dst);       // RayTracer::Pixels
  int ReadAccum(float* sum, float* sumsq);
  int WriteAccum(const float* sum, const float* sumsq, uint64_t frames);
  int AccumDevicePtr(void** ptr, size_t* n_floats);
  int AccumIpcHandle(uint8_t* handle);
  int ResolvePeers(const uint8_t* handles, uint32_t n_ranks, uint32_t self_rank, uint64_t total_frames, float* dst_mean, uint8_t* dst_rgba8);
  int ResolvePointers(const void* const* accums, uint32_t n, uint64_t total_frames, float* dst_mean, uint8_t* dst_rgba8);
  int RecordDone(void** event_out);  // cudaEvent_t after everything queued so far
  int WaitFor(void* event);          // this renderer's stream waits for another renderer's event
  void* AccumPtr();
  void* AccumSqPtr();
  int DebugCounters(uint64_t* out16, int* enabled);
  int QueueSizes(uint32_t* out, uint32_t max_bounces, uint32_t* n_bounces);
  int TextureValue(uint32_t tex_idx, const float* points, const float* uv, size_t n, float* rgb);
  int Intersect(const float* rays, size_t n, float tmin, float tmax, int skip_media, rt2_hit* out);
  int GetStats(rt2_stats* out);
  int ReadBvh(rt2_bvh_node* nodes, size_t max_nodes, uint32_t* prim_refs, size_t max_refs, uint32_t* n_pairs, uint32_t* n_refs,
              uint32_t* tlas_root);
  void* Stream();
  void SetProfiling(bool on) { profiling_ = on; }
  void SetFrameIdx(uint64_t f) {
    frame_idx_ = f;
    pending_frames_ = 0;
  }
  uint64_t FrameIdx() const { return frame_idx_ + pending_frames_; }  // RayTracer::FrameIdx: frames requested so far
  int Device() const { return cfg_.device; }
  int FramesPerBatch() const { return frames_per_batch_; }
  int Width() const { return width_; }
  int Height() const { return height_; }
  size_t SceneBytes() const { return scene_bytes_; }
  const std::string& Error() const { return err_; }

 private:
  int RenderBatch(uint32_t n_frames);
  int TraceFrames(uint32_t n_frames);
  int BuildTreesOnDevice(const HostScene& scene);
  int AllocState();
  int AllocSplitState();
  int CheckOverflow();
  int NoState();
  void FreeState();
  Impl* impl_{nullptr};
  rt2_config cfg_{};
  CameraParams cam_params_{};
  rt2_camera camera_{};
  int width_{0}, height_{0};
  int frames_per_batch_{1};
  int sm_count_{0};
  uint64_t frame_idx_{0};       // frames traced (queued on the stream)
  uint64_t pending_frames_{0};  // frames requested by Update() but not yet traced
  bool state_ok_{false};        // the frame buffers are allocated (false after a failed Resize)
  bool world_tree_ok_{false};   // the scene carries a surfaces-only world TLAS (instance split)
  uint32_t world_root_{0};
  bool unified_tree_ok_{false};  // the scene carries the unified world tree (instances flattened into world-space leaves)
  uint32_t unified_root_{0};
  uint32_t unified_depth_{0};
  uint32_t world_depth_{0};
  uint32_t n_textures_{0};
  std::vector<uint32_t> tree_depths_;  // node-pair depth of the trees on the device: [0] TLAS, [1 + i] BLAS of instance i, ...
  float node_inflation_{0.0f};         // surface-area growth of the quantised node boxes (rt_qnodes.cu)
  uint32_t max_stack_need_{0};         // stack entries the deepest traversal of this scene can need (checked against kStackSize)
  uint64_t launches_{0};
  double gpu_ms_total_{0};
  double prof_ms_[7]{0, 0, 0, 0, 0, 0, 0};
  bool profiling_{false};
  bool timing_pending_{false};
  size_t prof_used_{0};
  size_t scene_bytes_{0};
  uint32_t n_node_pairs_{0};
  uint32_t n_prim_refs_{0};
  double bvh_build_ms_{0};
  std::string err_;
};

int DeviceCount();
int MeasureFp32Peak(int device, double* tflops, std::string* err);
int MeasureL2Bandwidth(int device, double* gbs, std::string* err);

}  // namespace rt2
