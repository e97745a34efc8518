// Multi-GPU renderer behind one handle (see rt_multi.hpp).  Host code only; compiled by nvcc for the CUDA runtime API.
#include "rt_multi.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

namespace rt2 {

MultiRenderer::~MultiRenderer() {
  if (!staged_.empty() && !reps_.empty()) {
    cudaSetDevice(reps_[0]->Device());
    for (void* p : staged_)
      if (p) cudaFree(p);
  }
}

int MultiRenderer::Fail(size_t g, int rc) {
  if (rc != RT2_OK) err_ = reps_.size() > 1 ? "gpu " + std::to_string(reps_[g]->Device()) + ": " + reps_[g]->Error() : reps_[g]->Error();
  return rc;
}

uint64_t MultiRenderer::CountOf(size_t g, uint64_t frames) const {
  const uint64_t n = reps_.size();
  return frames > g ? (frames - g + n - 1) / n : 0;
}

int MultiRenderer::Init(const HostScene& scene, const rt2_config& cfg) {
  cfg_ = cfg;
  int n_dev = DeviceCount();
  if (n_dev < 1) {
    err_ = "no CUDA device available (this backend has no CPU fallback)";
    return RT2_ERR_CUDA;
  }
  int n = cfg.n_gpus;
  if (n == 0) n = 1;
  if (n < 0) n = n_dev - (cfg.device > 0 ? cfg.device : 0);
  if (cfg.device < 0 || n < 1 || cfg.device + n > n_dev) {
    err_ = "devices " + std::to_string(cfg.device) + ".." + std::to_string(cfg.device + n - 1) + " requested, " + std::to_string(n_dev) + " visible";
    return RT2_ERR_CUDA;
  }
  if (n > 16) {
    err_ = "at most 16 GPUs behind one handle";
    return RT2_ERR_INVALID_ARG;
  }
  const int stride0 = cfg.frame_stride < 1 ? 1 : cfg.frame_stride;
  for (int g = 0; g < n; g++) {
    rt2_config c = cfg;
    c.device = cfg.device + g;
    c.n_gpus = 1;
    c.frame_offset = cfg.frame_offset + g * stride0;
    c.frame_stride = stride0 * n;
    reps_.emplace_back(new Renderer);
    int rc = reps_.back()->Init(scene, c);
    if (rc != RT2_OK) return Fail(static_cast<size_t>(g), rc);
  }
  // peer paths from the first GPU to the others (NVLink / NVSwitch on a B200 box)
  peer_ok_ = true;
  if (n > 1) {
    const int d0 = cfg.device;
    if (cudaSetDevice(d0) != cudaSuccess) {
      err_ = "cudaSetDevice failed";
      return RT2_ERR_CUDA;
    }
    for (int g = 1; g < n; g++) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, d0, d0 + g) != cudaSuccess || !can) {
        peer_ok_ = false;
        continue;
      }
      const cudaError_t e = cudaDeviceEnablePeerAccess(d0 + g, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
      } else if (e != cudaSuccess) {
        cudaGetLastError();
        peer_ok_ = false;
      }
    }
  }
  frames_ = 0;
  return RT2_OK;
}

int MultiRenderer::UploadScene(const HostScene& scene) {
  for (size_t g = 0; g < reps_.size(); g++) {
    int rc = reps_[g]->UploadScene(scene);
    if (rc != RT2_OK) return Fail(g, rc);
  }
  return RT2_OK;
}

int MultiRenderer::Resize(int w, int h) {
  frames_ = 0;
  for (size_t g = 0; g < reps_.size(); g++) {
    int rc = reps_[g]->Resize(w, h);
    if (rc != RT2_OK) return Fail(g, rc);
  }
  return RT2_OK;
}

int MultiRenderer::Reset() {
  frames_ = 0;
  for (size_t g = 0; g < reps_.size(); g++) {
    int rc = reps_[g]->Reset();
    if (rc != RT2_OK) return Fail(g, rc);
  }
  return RT2_OK;
}

// Frames frames_ .. frames_ + n - 1 go to replica (k mod N).  The replicas are fed round-robin in batch-sized portions so
// that no GPU waits for the host to finish queueing another GPU's work.
int MultiRenderer::Update(uint32_t n_frames) {
  const size_t n = reps_.size();
  if (n == 1) {
    int rc = reps_[0]->Update(n_frames);
    if (rc == RT2_OK) frames_ += n_frames;
    return Fail(0, rc);
  }
  std::vector<uint64_t> todo(n);
  uint64_t left = 0;
  for (size_t g = 0; g < n; g++) {
    todo[g] = CountOf(g, frames_ + n_frames) - CountOf(g, frames_);
    left += todo[g];
  }
  while (left > 0) {
    for (size_t g = 0; g < n; g++) {
      if (todo[g] == 0) continue;
      const uint64_t portion = std::min<uint64_t>(todo[g], static_cast<uint64_t>(std::max(reps_[g]->FramesPerBatch(), 1)));
      int rc = reps_[g]->Update(static_cast<uint32_t>(portion));
      if (rc != RT2_OK) return Fail(g, rc);
      todo[g] -= portion;
      left -= portion;
    }
  }
  frames_ += n_frames;
  return RT2_OK;
}

int MultiRenderer::Flush() {
  for (size_t g = 0; g < reps_.size(); g++) {
    int rc = reps_[g]->Flush();
    if (rc != RT2_OK) return Fail(g, rc);
  }
  return RT2_OK;
}

int MultiRenderer::Synchronize() {
  int rc = Flush();  // queue everything on every GPU before waiting for any of them
  if (rc != RT2_OK) return rc;
  for (size_t g = 0; g < reps_.size(); g++) {
    rc = reps_[g]->Synchronize();
    if (rc != RT2_OK) return Fail(g, rc);
  }
  return RT2_OK;
}

int MultiRenderer::Resolve(float* dst_mean, uint8_t* dst_rgba8) {
  const size_t n = reps_.size();
  int rc = Flush();
  if (rc != RT2_OK) return rc;
  Renderer& r0 = *reps_[0];
  std::vector<const void*> ptrs(n);
  ptrs[0] = r0.AccumPtr();
  const size_t bytes = static_cast<size_t>(r0.Width()) * r0.Height() * 4 * sizeof(float);
  if (!peer_ok_) {
    // staging copies on the first GPU (PCIe / host bounce chosen by the driver)
    if (cudaSetDevice(r0.Device()) != cudaSuccess) {
      err_ = "cudaSetDevice failed";
      return RT2_ERR_CUDA;
    }
    if (staged_.size() != n || staged_bytes_ < bytes) {
      for (void* p : staged_)
        if (p) cudaFree(p);
      staged_.assign(n, nullptr);
      staged_bytes_ = 0;
      for (size_t g = 1; g < n; g++) {
        if (cudaMalloc(&staged_[g], bytes) != cudaSuccess) {
          err_ = "cudaMalloc(peer staging) failed";
          return RT2_ERR_CUDA;
        }
      }
      staged_bytes_ = bytes;
    }
  }
  for (size_t g = 1; g < n; g++) {
    void* ev = nullptr;
    rc = reps_[g]->RecordDone(&ev);
    if (rc != RT2_OK) return Fail(g, rc);
    rc = r0.WaitFor(ev);
    if (rc != RT2_OK) return Fail(0, rc);
    if (peer_ok_) {
      ptrs[g] = reps_[g]->AccumPtr();
    } else {
      cudaSetDevice(r0.Device());
      if (cudaMemcpyPeerAsync(staged_[g], r0.Device(), reps_[g]->AccumPtr(), reps_[g]->Device(), bytes,
                              static_cast<cudaStream_t>(r0.Stream())) != cudaSuccess) {
        err_ = "cudaMemcpyPeerAsync failed";
        return RT2_ERR_CUDA;
      }
      ptrs[g] = staged_[g];
    }
  }
  rc = r0.ResolvePointers(ptrs.data(), static_cast<uint32_t>(n), frames_, dst_mean, dst_rgba8);
  if (rc != RT2_OK) return Fail(0, rc);
  // the peers' kernels finished before the resolve ran; fold their timing events and check their traversal stacks
  for (size_t g = 1; g < n; g++) {
    rc = reps_[g]->Synchronize();
    if (rc == RT2_OK) {
      rt2_stats st{};
      rc = reps_[g]->GetStats(&st);
      if (rc == RT2_OK && st.stack_overflows) {
        err_ = "gpu " + std::to_string(reps_[g]->Device()) + ": BVH traversal stack overflow, the image is incomplete";
        return RT2_ERR_STATE;
      }
    }
    if (rc != RT2_OK) return Fail(g, rc);
  }
  return RT2_OK;
}

int MultiRenderer::ReadMean(float* dst) {
  if (reps_.size() == 1) return Fail(0, reps_[0]->ReadMean(dst));
  return Resolve(dst, nullptr);
}

int MultiRenderer::ReadRGBA8(uint8_t* dst) {
  if (reps_.size() == 1) return Fail(0, reps_[0]->ReadRGBA8(dst));
  return Resolve(nullptr, dst);
}

// Raw sums over all replicas, added on the host in replica order.
int MultiRenderer::ReadAccum(float* sum, float* sumsq) {
  if (reps_.size() == 1) return Fail(0, reps_[0]->ReadAccum(sum, sumsq));
  int rc = Flush();
  if (rc != RT2_OK) return rc;
  const size_t count = static_cast<size_t>(Width()) * Height() * 3;
  std::vector<float> a(sum ? count : 0), b(sumsq ? count : 0);
  for (size_t g = 0; g < reps_.size(); g++) {
    rc = reps_[g]->ReadAccum(sum ? (g == 0 ? sum : a.data()) : nullptr, sumsq ? (g == 0 ? sumsq : b.data()) : nullptr);
    if (rc != RT2_OK) return Fail(g, rc);
    if (g > 0) {
      if (sum)
        for (size_t i = 0; i < count; i++) sum[i] += a[i];
      if (sumsq)
        for (size_t i = 0; i < count; i++) sumsq[i] += b[i];
    }
  }
  return RT2_OK;
}

// Restore: the first replica takes the sums, the others start from zero; every replica's frame index is set to the number
// of the first `frames` global frames it owns, so the render continues with exactly the frames an uninterrupted one traces.
int MultiRenderer::WriteAccum(const float* sum, const float* sumsq, uint64_t frames) {
  if (reps_.size() == 1) {
    int rc = reps_[0]->WriteAccum(sum, sumsq, frames);
    if (rc == RT2_OK) frames_ = frames;
    return Fail(0, rc);
  }
  if (!sum) {
    err_ = "null sum";
    return RT2_ERR_INVALID_ARG;
  }
  const size_t count = static_cast<size_t>(Width()) * Height() * 3;
  std::vector<float> zero(count, 0.0f);
  for (size_t g = 0; g < reps_.size(); g++) {
    int rc = reps_[g]->WriteAccum(g == 0 ? sum : zero.data(), sumsq ? (g == 0 ? sumsq : zero.data()) : nullptr, CountOf(g, frames));
    if (rc != RT2_OK) return Fail(g, rc);
  }
  frames_ = frames;
  return RT2_OK;
}

int MultiRenderer::SetFrameIdx(uint64_t frames) {
  if (reps_.size() != 1) {
    err_ = "rt2_set_frame_idx: only for single-GPU handles (a multi-GPU handle does its own reduce)";
    return RT2_ERR_UNSUPPORTED;
  }
  reps_[0]->SetFrameIdx(frames);
  frames_ = frames;
  return RT2_OK;
}

Renderer* MultiRenderer::Single(const char* what) {
  if (reps_.size() != 1) {
    err_ = std::string(what) + ": only for single-GPU handles (a multi-GPU handle does its own reduce over peer memory)";
    return nullptr;
  }
  return reps_[0].get();
}

void MultiRenderer::SetProfiling(bool on) {
  for (auto& r : reps_) r->SetProfiling(on);
}

// Work counters add up over the replicas; times are those of the slowest replica (they run concurrently).
int MultiRenderer::GetStats(rt2_stats* out) {
  int rc = Flush();
  if (rc != RT2_OK) return rc;
  std::memset(out, 0, sizeof(*out));
  for (size_t g = 0; g < reps_.size(); g++) {
    rt2_stats s{};
    rc = reps_[g]->GetStats(&s);
    if (rc != RT2_OK) return Fail(g, rc);
    out->rays += s.rays;
    out->paths += s.paths;
    out->launches += s.launches;
    out->box_pair_tests += s.box_pair_tests;
    out->sphere_tests += s.sphere_tests;
    out->quad_tests += s.quad_tests;
    out->instance_visits += s.instance_visits;
    out->stack_overflows += s.stack_overflows;
    out->gpu_ms_total = std::max(out->gpu_ms_total, s.gpu_ms_total);
    out->gpu_ms_extend = std::max(out->gpu_ms_extend, s.gpu_ms_extend);
    out->gpu_ms_shade = std::max(out->gpu_ms_shade, s.gpu_ms_shade);
    out->gpu_ms_other = std::max(out->gpu_ms_other, s.gpu_ms_other);
    out->gpu_ms_finish = std::max(out->gpu_ms_finish, s.gpu_ms_finish);
    out->gpu_ms_sort = std::max(out->gpu_ms_sort, s.gpu_ms_sort);
    out->gpu_ms_extend_inst = std::max(out->gpu_ms_extend_inst, s.gpu_ms_extend_inst);
    out->gpu_ms_bvh_build = std::max(out->gpu_ms_bvh_build, s.gpu_ms_bvh_build);
    out->instance_split = s.instance_split;
    out->instance_mode = s.instance_mode;
    out->compact_nodes = s.compact_nodes;
    out->node_inflation = s.node_inflation;
    out->max_stack_need = std::max(out->max_stack_need, s.max_stack_need);
  }
  out->frames = frames_;
  out->pending_frames = 0;
  out->n_gpus = static_cast<uint32_t>(reps_.size());
  return RT2_OK;
}

}  // namespace rt2
