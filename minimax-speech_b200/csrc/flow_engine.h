#pragma once
#include <map>
#include <memory>
#include <vector>

#include "engine_common.h"

namespace ls {

class FlowEngine {
 public:
  // fp16: the 16-bit GEMM operands (weights and activations) are fp16 instead of bf16; same kernels, same speed
  FlowEngine(const Weights& w, int device, bool fp16 = false);
  ~FlowEngine();
  void estimator_forward(const float* x, const float* mask, const float* mu, const float* t, const float* spks,
                         const float* cond, float* out, int rows, int T, bool streaming, cudaStream_t s);
  void solve(const float* mu, const float* mask, const float* spks, const float* cond, const float* noise,
             long long noise_stride, const float* t_span, int n_steps, float temperature, float cfg_rate,
             bool streaming, float* out, int B, int T, cudaStream_t s);
  int feat() const { return feat_; }
  int device() const { return device_; }
  // changes whenever the workspace was reallocated (CUDA graphs captured over it must be captured again)
  unsigned long long ws_generation() const { return ws_generation_; }
  // throws if an earlier call saw a non-prefix mask (detected on the device, reported without a synchronisation)
  void check_sticky();

 private:
  struct ResnetW;
  struct TBlockW;
  struct GroupW;
  struct Plan;

  void ensure_workspace(int B2, int T, int nt, cudaStream_t s);
  const Plan& plan_for(int B2, int T);
  void run_estimator(int B2, int T, const float* temb, long long temb_bstride, bool streaming, cudaStream_t s);
  void time_embed(const float* t_dev, const float* t_host, int nt, cudaStream_t s);
  template <typename T>
  T* ws(size_t off) const { return reinterpret_cast<T*>(ws_base_ + off); }

  int device_ = 0, num_sms_ = 148;
  int fp16_ = 0;
  bool causal_ = true;  // false: the non-causal ConditionalDecoder (Conv1d pad 1 + GroupNorm(8) blocks)
  bool fused_blocks_ = false;
  int C_ = 256, in_ch_ = 320, feat_ = 80, heads_ = 8, hid_ = 1024, n_blocks_ = 4, n_mid_ = 12, chunk_ = 50;
  Arena arena_;
  std::vector<GroupW> groups_;
  PackedLinear down_conv_, up_conv_, final_conv_, final_proj_;
  size_t final_lng_ = 0, final_lnb_ = 0;
  size_t freqs_ = 0, w1_ = 0, b1_ = 0, w2_ = 0, b2_ = 0, wr_ = 0, br_ = 0;

  uint8_t* ws_base_ = nullptr;
  size_t ws_bytes_ = 0;
  long long cap_rows_ = 0;
  int cap_nt_ = 0, cap_b2_ = 0;
  size_t o_xin_ = 0, o_hA_ = 0, o_hB_ = 0, o_skip_ = 0, o_nrm_ = 0, o_qkv_ = 0, o_att_ = 0, o_ff_ = 0, o_u_ = 0,
         o_r_ = 0, o_v_ = 0, o_x_ = 0, o_len_ = 0, o_t_ = 0, o_temb_ = 0, o_tscr_ = 0, o_raw_ = 0, o_gn_ = 0;
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> plans_;
  std::vector<float> t_host_, dt_host_;
  unsigned long long ws_generation_ = 0;
  int* bad_mask_host_ = nullptr;  // mapped pinned flag written by mask_to_lengths_kernel
  int* bad_mask_dev_ = nullptr;
};

}  // namespace ls
