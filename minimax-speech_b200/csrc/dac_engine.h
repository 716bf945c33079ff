#pragma once
#include <map>
#include <memory>
#include <vector>

#include "engine_common.h"

namespace ls {

// shared with the encoder engine (dac_enc_engine.cu): weight-norm folding + bf16 [taps][N][K] packing
PackedLinear pack_wn_conv(Arena& a, const Weights& w, const std::string& p, int n_pad_to = 0);
void finalize_linear(const Arena& a, PackedLinear& pl);

// DAC-VAE decoder (dac-vae/model.py:326-379, 485-488) on time-major activations.
class DacEngine {
 public:
  DacEngine(const Weights& w, int device);
  ~DacEngine();
  void decode(const float* z, const int* lengths, float* wav, int B, int L, cudaStream_t s);
  int hop() const { return hop_; }
  int latent_dim() const { return latent_; }
  int device() const { return device_; }
  unsigned long long ws_generation() const { return ws_generation_; }

 private:
  struct UnitW;
  struct StageW;
  struct Plan;
  void ensure_workspace(int B, int L, cudaStream_t s);
  const Plan& plan_for(int B, int L);
  template <typename T>
  T* ws(size_t off) const { return reinterpret_cast<T*>(ws_base_ + off); }

  int device_ = 0, num_sms_ = 148;
  int latent_ = 80, dim_ = 1536, hop_ = 1, out_ch_ = 1;
  std::vector<int> rates_;
  Arena arena_;
  PackedLinear pre_, in_, final_;
  std::vector<StageW> stages_;
  size_t final_alpha_ = 0, final_ialpha_ = 0;

  uint8_t* ws_base_ = nullptr;
  long long cap_frames_ = 0;  // B*L capacity
  int cap_b_ = 0;
  unsigned long long ws_generation_ = 0;
  size_t o_zt_ = 0, o_a0_ = 0, o_x_ = 0, o_sA_[2] = {0, 0}, o_sB_ = 0, o_len_ = 0;
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> plans_;
};

}  // namespace ls
