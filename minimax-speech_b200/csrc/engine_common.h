// Host-side helpers shared by the flow and DAC engines: named-weight lookup, packed device arena, errors.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ls_b200.h"
#include "kernels.h"

namespace ls {

void set_error(const char* fmt, ...);
const char* get_error();

struct EngineError : std::runtime_error {
  int code;
  EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define LS_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw ::ls::EngineError(LS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
  } while (0)

inline void require(bool ok, const std::string& msg, int code = LS_ERR_INVALID) {
  if (!ok) throw EngineError(code, msg);
}

// Read-only view of the caller's state_dict (host fp32).
class Weights {
 public:
  Weights(const ls_tensor* t, int n) {
    for (int i = 0; i < n; ++i) map_[t[i].name] = &t[i];
  }
  bool has(const std::string& name) const { return map_.count(name) != 0; }
  const ls_tensor& get(const std::string& name) const {
    auto it = map_.find(name);
    if (it == map_.end()) throw EngineError(LS_ERR_WEIGHTS, "missing weight '" + name + "'");
    return *it->second;
  }
  const std::map<std::string, const ls_tensor*>& all() const { return map_; }
  const ls_tensor& get(const std::string& name, std::initializer_list<long long> shape) const {
    const ls_tensor& t = get(name);
    bool ok = t.ndim == (int)shape.size();
    int i = 0;
    for (long long s : shape) ok = ok && (i < t.ndim) && t.shape[i++] == s;
    if (!ok) throw EngineError(LS_ERR_WEIGHTS, "unexpected shape for weight '" + name + "'");
    return t;
  }

 private:
  std::map<std::string, const ls_tensor*> map_;
};

// Host staging buffer that becomes one device allocation; offsets are 256-byte aligned.
class Arena {
 public:
  size_t reserve(size_t bytes) {
    size_t off = (host_.size() + 255) & ~size_t(255);
    host_.resize(off + bytes, 0);
    return off;
  }
  size_t put_f32(const float* src, size_t n) {
    size_t off = reserve(n * 4);
    std::memcpy(host_.data() + off, src, n * 4);
    return off;
  }
  uint8_t* host(size_t off) { return host_.data() + off; }
  void upload() {
    LS_CUDA(cudaMalloc(&dev_, host_.size() ? host_.size() : 256));
    LS_CUDA(cudaMemcpy(dev_, host_.data(), host_.size(), cudaMemcpyHostToDevice));
    bytes_ = host_.size();
    host_.clear();
    host_.shrink_to_fit();
  }
  template <typename T>
  T* ptr(size_t off) const { return reinterpret_cast<T*>(dev_ + off); }
  size_t bytes() const { return bytes_; }
  ~Arena() {
    if (dev_) cudaFree(dev_);
  }

 private:
  std::vector<uint8_t> host_;
  uint8_t* dev_ = nullptr;
  size_t bytes_ = 0;
};

// Stream-ordered workspace (re)allocation: growth happens with cudaFreeAsync / cudaMallocAsync on the calling stream --
// no device-wide synchronisation, nothing another handle or stream has to wait for ("no device-wide syncs").
// The C-ABI layer orders calls that reach one handle from different streams (StreamSerial in c_api.cu), so the old
// block's last users are ordered before the free.  Growth cannot be captured into a CUDA graph: run the call once
// outside capture with the largest shapes first.
inline void ws_release(uint8_t*& base, cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  LS_CUDA(cudaStreamIsCapturing(s, &st));
  require(st == cudaStreamCaptureStatusNone,
          "workspace growth during stream capture: run the call once outside capture with these shapes first");
  if (base) LS_CUDA(cudaFreeAsync(base, s));
  base = nullptr;
}
inline void ws_alloc(uint8_t*& base, size_t bytes, cudaStream_t s) {
  void* p = nullptr;
  LS_CUDA(cudaMallocAsync(&p, bytes, s));
  base = reinterpret_cast<uint8_t*>(p);
  LS_CUDA(cudaMemsetAsync(base, 0, bytes, s));
}

// One dense contraction's weights in kernel layout: bf16 [taps*N][K] (K contiguous) + its TMA map.
struct PackedLinear {
  size_t w_off = 0, bias_off = 0;
  bool has_bias = false;
  int N = 0, K = 0, taps = 1, block_n = 0;
  CUtensorMap map;
  const float* bias = nullptr;  // device
};

inline int pick_block_n(int N) {
  if (N <= 256) return N;
  for (int bn = 256; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 16;
}

}  // namespace ls
