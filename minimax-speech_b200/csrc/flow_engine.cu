// Flow engine: CausalConditionalDecoder estimator + Euler/CFG solve on time-major activations.
//
// Follows, layer for layer, speech/cosyvoice/flow/decoder.py:405-496 (channels=[256]) and
// speech/cosyvoice/flow/flow_matching.py:74-126,323-348; every dense contraction is a conv_gemm launch, the
// self-attention core is the flash kernel, everything row-local is fused into GEMM epilogues.
#include <cmath>
#include <memory>
#include <mutex>

#include <cuda_fp16.h>

#include "engine_common.h"
#include "flow_engine.h"

namespace ls {
namespace {

// 16-bit operand format of a handle: bf16, or fp16 (ls_flow_create_fp16: same kernels, compiled for fp16 operands)
bool g_pack_fp16 = false;  // set for the duration of a constructor (host-side weight packing only)
inline __nv_bfloat16 to_h16(float x) {
  if (!g_pack_fp16) return __float2bfloat16(x);
  const __half h = __float2half_rn(x);
  __nv_bfloat16 out;
  std::memcpy(&out, &h, 2);
  return out;
}
void to_bf16(const float* src, __nv_bfloat16* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = to_h16(src[i]);
}

// weight [N][K] (linear) or [N][K][taps] (conv1d) -> 16-bit [taps][N][K]
PackedLinear pack_linear(Arena& a, const ls_tensor& w, const ls_tensor* bias) {
  PackedLinear pl;
  pl.N = (int)w.shape[0];
  pl.K = (int)w.shape[1];
  pl.taps = w.ndim == 3 ? (int)w.shape[2] : 1;
  pl.block_n = pick_block_n(pl.N);
  pl.w_off = a.reserve((size_t)pl.taps * pl.N * pl.K * 2);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.host(pl.w_off));
  for (int t = 0; t < pl.taps; ++t)
    for (int n = 0; n < pl.N; ++n)
      for (int k = 0; k < pl.K; ++k)
        dst[((size_t)t * pl.N + n) * pl.K + k] = to_h16(w.data[((size_t)n * pl.K + k) * pl.taps + t]);
  if (bias) {
    require(bias->shape[0] == pl.N, std::string("bias shape mismatch for ") + w.name, LS_ERR_WEIGHTS);
    pl.bias_off = a.put_f32(bias->data, pl.N);
    pl.has_bias = true;
  }
  return pl;
}

void finalize_linear(const Arena& a, PackedLinear& pl) {
  require(make_weight_map(&pl.map, a.ptr<uint8_t>(pl.w_off), pl.K, pl.taps * pl.N, pl.block_n),
          "cuTensorMapEncodeTiled failed for a weight matrix", LS_ERR_CUDA);
  pl.bias = pl.has_bias ? a.ptr<float>(pl.bias_off) : nullptr;
}

bool fused_blocks_check(int C, int inner) { return C == 256 && inner == 512; }

// LS_HEAD_VIA_CONV=1 (development aid): the first block's QKV as a conv_gemm launch over the LayerNorm output of block2's
// epilogue instead of the fused kernel's head launch.  Measured: tblock -4.5 ms, conv_gemm +5.0 ms per step (70.3 -> 71.3 ms,
// profiles/r02_ab_head_via_conv.log), so it is off.
bool head_via_conv_enabled() {
  static const bool on = [] {
    const char* e = getenv("LS_HEAD_VIA_CONV");
    return e && e[0] == '1';
  }();
  return on;
}

int count_prefix(const Weights& w, const char* fmt) {
  int n = 0;
  char buf[160];
  for (;; ++n) {
    snprintf(buf, sizeof buf, fmt, n);
    if (!w.has(buf)) break;
  }
  return n;
}

}  // namespace

struct FlowEngine::ResnetW {
  PackedLinear conv1, conv2, res;
  size_t ln1g, ln1b, ln2g, ln2b;
  int cin;
};
struct FlowEngine::TBlockW {
  PackedLinear qkv, out, ff1, ff2;
  size_t n1g, n1b, n3g, n3b;
  // fused transformer-block kernel (tblock.cu): per-block vector pack + weight maps with 128-row boxes
  size_t vec = 0;
  CUtensorMap m_out, m_ff1, m_ff2, m_qkv;
};
struct FlowEngine::GroupW {
  ResnetW res;
  std::vector<TBlockW> tb;
  size_t head_vec = 0;  // fused-kernel vector pack holding norm1 of the first block (head mode: LayerNorm + QKV)
};

// activation views: k1 = 128-row boxes (linear / 1x1 layers), k3 = boxes with the 2-row halo of a causal k=3 conv
struct ActMaps {
  CUtensorMap k1, k3;
};
struct FlowEngine::Plan {
  ActMaps xin, hA, hB, skip, nrm, qkv, att, ff;
  // flattened [B2*T][C] views for the fused block kernel (TMA loads and stores)
  CUtensorMap att_flat, u_flat, qkv_flat, tail_skip, tail_hB, qkv_attn;
};

FlowEngine::~FlowEngine() {
  if (ws_base_) cudaFree(ws_base_);
  if (bad_mask_host_) cudaFreeHost(bad_mask_host_);
}

// A non-prefix mask is detected on the device (mask_to_lengths_kernel) and reported here, at the next call on the
// handle: validating it on the hot call would cost a host synchronisation per call.
void FlowEngine::check_sticky() {
  if (bad_mask_host_ && *bad_mask_host_) {
    *bad_mask_host_ = 0;
    throw EngineError(LS_ERR_INVALID, "an earlier call on this handle was given a mask that is not a prefix (right-padding) "
                                      "mask; its result is undefined");
  }
}

FlowEngine::FlowEngine(const Weights& w, int device, bool fp16) : device_(device), fp16_(fp16 ? 1 : 0) {
  static std::mutex pack_mu;  // g_pack_fp16 is process-wide: constructors are serialised
  std::lock_guard<std::mutex> pack_lock(pack_mu);
  g_pack_fp16 = fp16;
  LS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LS_CUDA(cudaGetDeviceProperties(&prop, device));
  require(prop.major == 10, "this library only runs on sm_100 (B200) devices", LS_ERR_UNSUPPORTED);
  num_sms_ = prop.multiProcessorCount;
  LS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&bad_mask_host_), sizeof(int), cudaHostAllocMapped));
  *bad_mask_host_ = 0;
  LS_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&bad_mask_dev_), bad_mask_host_, 0));

  const ls_tensor& l1 = w.get("time_mlp.linear_1.weight");
  hid_ = (int)l1.shape[0];
  in_ch_ = (int)l1.shape[1];
  const ls_tensor& fp = w.get("final_proj.weight");
  feat_ = (int)fp.shape[0];
  C_ = (int)fp.shape[1];
  n_blocks_ = count_prefix(w, "down_blocks.0.1.%d.norm1.weight");
  n_mid_ = count_prefix(w, "mid_blocks.%d.0.mlp.1.weight");
  const int inner = (int)w.get("down_blocks.0.1.0.attn1.to_q.weight").shape[0];
  heads_ = inner / 64;
  require(C_ == 256 && inner % 64 == 0 && feat_ % 16 == 0 && in_ch_ == 4 * feat_ && hid_ == 4 * C_ &&
              !w.has("down_blocks.1.0.mlp.1.weight") && n_blocks_ >= 1,
          "estimator configuration not covered: need channels=[256], head_dim 64, in_channels = 4*out_channels",
          LS_ERR_UNSUPPORTED);

  // CausalConditionalDecoder keeps a LayerNorm at block.2 of every block (decoder.py:65-76), the non-causal
  // ConditionalDecoder (decoder.py:88-291; matcha Block1D) a GroupNorm(8) at block.1
  causal_ = w.has("final_block.block.2.weight");
  require(causal_ || w.has("final_block.block.1.weight"), "estimator state dict: neither block.2 (causal) nor block.1 norms",
          LS_ERR_WEIGHTS);
  const std::string nk = causal_ ? ".block.2" : ".block.1";
  auto vec = [&](const std::string& name, int n) { return arena_.put_f32(w.get(name, {n}).data, n); };
  auto resnet = [&](const std::string& p, int cin) {
    ResnetW r;
    r.cin = cin;
    r.conv1 = pack_linear(arena_, w.get(p + ".block1.block.0.weight", {C_, cin, 3}), &w.get(p + ".block1.block.0.bias"));
    r.ln1g = vec(p + ".block1" + nk + ".weight", C_);
    r.ln1b = vec(p + ".block1" + nk + ".bias", C_);
    r.conv2 = pack_linear(arena_, w.get(p + ".block2.block.0.weight", {C_, C_, 3}), &w.get(p + ".block2.block.0.bias"));
    r.ln2g = vec(p + ".block2" + nk + ".weight", C_);
    r.ln2b = vec(p + ".block2" + nk + ".bias", C_);
    r.res = pack_linear(arena_, w.get(p + ".res_conv.weight", {C_, cin, 1}), &w.get(p + ".res_conv.bias"));
    return r;
  };
  auto tblock = [&](const std::string& p) {
    TBlockW t;
    t.n1g = vec(p + ".norm1.weight", C_);
    t.n1b = vec(p + ".norm1.bias", C_);
    // fused QKV weight: rows [to_q ; to_k ; to_v]
    const ls_tensor& q = w.get(p + ".attn1.to_q.weight", {inner, C_});
    const ls_tensor& k = w.get(p + ".attn1.to_k.weight", {inner, C_});
    const ls_tensor& v = w.get(p + ".attn1.to_v.weight", {inner, C_});
    t.qkv.N = 3 * inner, t.qkv.K = C_, t.qkv.taps = 1, t.qkv.block_n = pick_block_n(3 * inner);
    t.qkv.w_off = arena_.reserve((size_t)3 * inner * C_ * 2);
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(arena_.host(t.qkv.w_off));
    to_bf16(q.data, d, (size_t)inner * C_);
    to_bf16(k.data, d + (size_t)inner * C_, (size_t)inner * C_);
    to_bf16(v.data, d + (size_t)2 * inner * C_, (size_t)inner * C_);
    t.out = pack_linear(arena_, w.get(p + ".attn1.to_out.0.weight", {C_, inner}), &w.get(p + ".attn1.to_out.0.bias"));
    t.n3g = vec(p + ".norm3.weight", C_);
    t.n3b = vec(p + ".norm3.bias", C_);
    t.ff1 = pack_linear(arena_, w.get(p + ".ff.net.0.proj.weight", {4 * C_, C_}), &w.get(p + ".ff.net.0.proj.bias"));
    t.ff2 = pack_linear(arena_, w.get(p + ".ff.net.2.weight", {C_, 4 * C_}), &w.get(p + ".ff.net.2.bias"));
    return t;
  };
  auto group = [&](const std::string& p, int cin) {
    GroupW g;
    g.res = resnet(p + ".0", cin);
    for (int j = 0; j < n_blocks_; ++j) g.tb.push_back(tblock(p + ".1." + std::to_string(j)));
    return g;
  };

  groups_.push_back(group("down_blocks.0", in_ch_));
  for (int i = 0; i < n_mid_; ++i) groups_.push_back(group("mid_blocks." + std::to_string(i), C_));
  groups_.push_back(group("up_blocks.0", 2 * C_));
  down_conv_ = pack_linear(arena_, w.get("down_blocks.0.2.weight", {C_, C_, 3}), &w.get("down_blocks.0.2.bias"));
  up_conv_ = pack_linear(arena_, w.get("up_blocks.0.2.weight", {C_, C_, 3}), &w.get("up_blocks.0.2.bias"));
  final_conv_ = pack_linear(arena_, w.get("final_block.block.0.weight", {C_, C_, 3}), &w.get("final_block.block.0.bias"));
  final_lng_ = vec("final_block" + nk + ".weight", C_);
  final_lnb_ = vec("final_block" + nk + ".bias", C_);
  require(causal_ || fused_blocks_check(C_, inner), "the non-causal estimator needs the fused-block geometry (C = 256, 8 x 64)",
          LS_ERR_UNSUPPORTED);
  final_proj_ = pack_linear(arena_, w.get("final_proj.weight", {feat_, C_, 1}), &w.get("final_proj.bias"));

  // timestep conditioning (fp32): sinusoid frequencies, time_mlp, one Linear per resnet
  {
    const int half = in_ch_ / 2;
    std::vector<float> f(half);
    const float step = (float)(-(std::log(10000.0) / (half - 1)));  // matcha decoder.py:24-25
    for (int i = 0; i < half; ++i) f[i] = (float)std::exp((double)((float)i * step));
    freqs_ = arena_.put_f32(f.data(), half);
    w1_ = arena_.put_f32(l1.data, (size_t)hid_ * in_ch_);
    b1_ = vec("time_mlp.linear_1.bias", hid_);
    w2_ = arena_.put_f32(w.get("time_mlp.linear_2.weight", {hid_, hid_}).data, (size_t)hid_ * hid_);
    b2_ = vec("time_mlp.linear_2.bias", hid_);
    const int nres = (int)groups_.size();
    wr_ = arena_.reserve((size_t)nres * C_ * hid_ * 4);
    br_ = arena_.reserve((size_t)nres * C_ * 4);
    for (int r = 0; r < nres; ++r) {
      const std::string p = r == 0 ? "down_blocks.0.0" : (r == nres - 1 ? "up_blocks.0.0" : "mid_blocks." + std::to_string(r - 1) + ".0");
      std::memcpy(arena_.host(wr_) + (size_t)r * C_ * hid_ * 4, w.get(p + ".mlp.1.weight", {C_, hid_}).data, (size_t)C_ * hid_ * 4);
      std::memcpy(arena_.host(br_) + (size_t)r * C_ * 4, w.get(p + ".mlp.1.bias", {C_}).data, (size_t)C_ * 4);
    }
  }
  // fused transformer-block kernel: bo | norm3 | ff bias 1 | ff bias 2 | norm1 of the NEXT block of the group
  fused_blocks_ = C_ == 256 && inner == 512;
  if (fused_blocks_) {
    for (auto& g : groups_) {
      g.head_vec = arena_.reserve(TBLOCK_VEC_FLOATS * 4);
      std::memcpy(arena_.host(g.head_vec) + (size_t)2048 * 4, arena_.host(g.tb[0].n1g), 256 * 4);
      std::memcpy(arena_.host(g.head_vec) + (size_t)2304 * 4, arena_.host(g.tb[0].n1b), 256 * 4);
    }
    for (auto& g : groups_)
      for (size_t j = 0; j < g.tb.size(); ++j) {
        TBlockW& t = g.tb[j];
        t.vec = arena_.reserve(TBLOCK_VEC_FLOATS * 4);
        auto put = [&](int off, size_t src_off, int n) {
          std::memcpy(arena_.host(t.vec) + (size_t)off * 4, arena_.host(src_off), (size_t)n * 4);
        };
        put(0, t.out.bias_off, 256);
        put(256, t.n3g, 256);
        put(512, t.n3b, 256);
        put(768, t.ff1.bias_off, 1024);
        put(1792, t.ff2.bias_off, 256);
        if (j + 1 < g.tb.size()) {
          put(2048, g.tb[j + 1].n1g, 256);
          put(2304, g.tb[j + 1].n1b, 256);
        }
      }
  }
  arena_.upload();
  if (fused_blocks_) {
    for (auto& g : groups_)
      for (auto& t : g.tb) {
        require(make_weight_map(&t.m_out, arena_.ptr<uint8_t>(t.out.w_off), 512, 256, TBLOCK_WIDE_BOX_ROWS) &&
                    make_weight_map(&t.m_ff2, arena_.ptr<uint8_t>(t.ff2.w_off), 1024, 256, TBLOCK_WIDE_BOX_ROWS) &&
                    (TBLOCK_PAIR ? make_weight_map_kb(&t.m_ff1, arena_.ptr<uint8_t>(t.ff1.w_off), 256, 1024, 64, 2) &&
                                       make_weight_map_kb(&t.m_qkv, arena_.ptr<uint8_t>(t.qkv.w_off), 256, 1536, 64, 2)
                                 : make_weight_map(&t.m_ff1, arena_.ptr<uint8_t>(t.ff1.w_off), 256, 1024, TBLOCK_WBOX_ROWS) &&
                                       make_weight_map(&t.m_qkv, arena_.ptr<uint8_t>(t.qkv.w_off), 256, 1536, TBLOCK_WBOX_ROWS)),
                "cuTensorMapEncodeTiled failed for a fused-block weight matrix", LS_ERR_CUDA);
      }
  }
  for (auto& g : groups_) {
    finalize_linear(arena_, g.res.conv1);
    finalize_linear(arena_, g.res.conv2);
    finalize_linear(arena_, g.res.res);
    for (auto& t : g.tb) {
      finalize_linear(arena_, t.qkv);
      finalize_linear(arena_, t.out);
      finalize_linear(arena_, t.ff1);
      finalize_linear(arena_, t.ff2);
    }
  }
  finalize_linear(arena_, down_conv_);
  finalize_linear(arena_, up_conv_);
  finalize_linear(arena_, final_conv_);
  finalize_linear(arena_, final_proj_);
}

// ------------------------------------------------------------------------------------------------
void FlowEngine::ensure_workspace(int B2, int T, int nt, cudaStream_t s) {
  const long long rows = (long long)B2 * T;
  if (rows <= cap_rows_ && nt <= cap_nt_ && B2 <= cap_b2_) return;
  ws_release(ws_base_, s);
  plans_.clear();
  cap_rows_ = std::max(rows, cap_rows_);
  cap_nt_ = std::max(nt, cap_nt_);
  cap_b2_ = std::max(B2, cap_b2_);
  const int inner = heads_ * 64;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = (off + bytes + 1023) & ~size_t(1023);
    return o;
  };
  const size_t R = (size_t)cap_rows_;
  o_xin_ = take(R * in_ch_ * 2);
  o_hA_ = take(R * C_ * 2);
  o_hB_ = take(R * C_ * 2);
  o_skip_ = take(R * C_ * 2);
  o_nrm_ = take(R * C_ * 2);
  o_qkv_ = take(R * 3 * inner * 2);
  o_att_ = take(R * inner * 2);
  o_ff_ = take(R * 4 * C_ * 2);
  o_u_ = take(R * C_ * 4);
  o_r_ = take(R * C_ * 4);
  o_v_ = take(R * feat_ * 4);
  o_x_ = take(R * feat_ * 4);
  o_len_ = take((size_t)cap_b2_ * 4);
  o_t_ = take((size_t)cap_nt_ * 4);
  o_temb_ = take((size_t)cap_nt_ * groups_.size() * C_ * 4);
  o_tscr_ = take((size_t)cap_nt_ * 2 * hid_ * 4);
  if (!causal_) {
    o_raw_ = take(R * C_ * 4);                  // convolution output before GroupNorm (fp32)
    o_gn_ = take((size_t)cap_b2_ * 8 * 8);      // (mean, rstd) per (batch row, group)
  }
  ws_alloc(ws_base_, off, s);
  ws_bytes_ = off;
  ++ws_generation_;
}

const FlowEngine::Plan& FlowEngine::plan_for(int B2, int T) {
  auto key = std::make_pair(B2, T);
  auto it = plans_.find(key);
  if (it != plans_.end()) return *it->second;
  auto pl = std::make_unique<Plan>();
  const int inner = heads_ * 64;
  const int box3 = conv_halo_enabled() ? conv_halo_box_rows(3, 1) : 128;
  auto mk = [&](ActMaps* m, size_t off, int C) {
    require(make_act_map(&m->k1, ws_base_ + off, C, T, B2, C, (long long)T * C, 128) &&
                make_act_map(&m->k3, ws_base_ + off, C, T, B2, C, (long long)T * C, box3),
            "cuTensorMapEncodeTiled failed for an activation buffer", LS_ERR_CUDA);
  };
  mk(&pl->xin, o_xin_, in_ch_);
  mk(&pl->hA, o_hA_, C_);
  mk(&pl->hB, o_hB_, C_);
  mk(&pl->skip, o_skip_, C_);
  mk(&pl->nrm, o_nrm_, C_);
  mk(&pl->qkv, o_qkv_, 3 * inner);
  require(make_act_map(&pl->qkv_attn, ws_base_ + o_qkv_, 3 * inner, T, B2, 3 * inner, (long long)T * 3 * inner, ATTN_KV),
          "cuTensorMapEncodeTiled failed for the attention view of QKV", LS_ERR_CUDA);
  mk(&pl->att, o_att_, inner);
  mk(&pl->ff, o_ff_, 4 * C_);
  const long long R = (long long)B2 * T;
  require(make_tile_map(&pl->att_flat, ws_base_ + o_att_, 2, inner, R, 128) &&
              make_tile_map(&pl->u_flat, ws_base_ + o_u_, 4, C_, R, 128) &&
              make_tile_map(&pl->qkv_flat, ws_base_ + o_qkv_, 2, 3 * inner, R, 128) &&
              make_tile_map(&pl->tail_skip, ws_base_ + o_skip_, 2, C_, R, 128) &&
              make_tile_map(&pl->tail_hB, ws_base_ + o_hB_, 2, C_, R, 128),
          "cuTensorMapEncodeTiled failed for a fused-block activation view", LS_ERR_CUDA);
  const Plan& ref = *pl;
  plans_[key] = std::move(pl);
  return ref;
}

namespace {
struct Epi {
  int act = ACT_NONE;
  const float* ln_g = nullptr;
  const float* ln_b = nullptr;
  const float* temb = nullptr;
  long long temb_bstride = 0;
  const void* addend = nullptr;
  int addend_dtype = OUT_F32;
  void* out0 = nullptr;
  int out0_dtype = OUT_NONE;
  void* out1 = nullptr;
  int out1_mode = OUT1_NONE;
  const float* p1_a = nullptr;
  const float* p1_b = nullptr;
  int zero_skipped = 0;
  // fused residual GEMM (addend_dtype == ADD_GEMM): res_w applied to the activation maps res_a0 (+ res_a1 after res_a0_ch channels)
  const PackedLinear* res_w = nullptr;
  const CUtensorMap* res_a0 = nullptr;
  const CUtensorMap* res_a1 = nullptr;
  int res_a0_ch = 0;
};
}  // namespace

void FlowEngine::run_estimator(int B2, int T, const float* temb, long long temb_bstride, bool streaming,
                               cudaStream_t s) {
  if (!causal_) streaming = false;  // ConditionalDecoder.forward ignores the flag: full attention (decoder.py:241)
  const Plan& pl = plan_for(B2, T);
  const int* lengths = ws<int>(o_len_);
  const int inner = heads_ * 64;

  const bool halo = conv_halo_enabled();
  auto gemm = [&](const ActMaps& am0, const ActMaps* am1, int a0_channels, const PackedLinear& w, const Epi& e) {
    const CUtensorMap& a0 = w.taps > 1 ? am0.k3 : am0.k1;
    const CUtensorMap* a1 = am1 ? (w.taps > 1 ? &am1->k3 : &am1->k1) : nullptr;
    ConvGemmParams p{};
    p.B = B2, p.M = T, p.N = w.N, p.block_n = w.block_n;
    p.taps = w.taps, p.dil = 1;
    p.pad = causal_ ? w.taps - 1 : (w.taps - 1) / 2;  // causal: left context only (decoder.py:59-62); else Conv1d(padding=1)
    p.kb_per_tap = (w.K + 63) / 64;
    p.kb_split = a1 ? a0_channels / 64 : p.kb_per_tap;
    p.lengths = lengths, p.m_len_mul = 1, p.m_len_add = 0, p.skip_halo = 0;
    p.chan_mod = w.N, p.bias = w.bias, p.act = e.act, p.ln_g = e.ln_g, p.ln_b = e.ln_b;
    p.temb = e.temb, p.temb_bstride = e.temb_bstride, p.addend = e.addend, p.addend_dtype = e.addend_dtype;
    p.out0 = e.out0, p.out0_dtype = e.out0_dtype, p.out1 = e.out1, p.out1_mode = e.out1_mode;
    p.p1_a = e.p1_a, p.p1_b = e.p1_b, p.n_store = w.N;
    p.out_ld = w.N, p.out_shift = 0, p.out_bstride = (long long)T * w.N, p.out_alloc = (long long)T * w.N;
    p.out_valid_mul = w.N;
    p.k_true = w.K, p.tag = 0, p.zero_skipped = e.zero_skipped, p.fp16 = fp16_;
    p.halo_mode = halo ? conv_halo_mode() : 0;
    if (e.addend_dtype == ADD_GEMM) {
      p.addend = nullptr;
      p.res_a0 = *e.res_a0, p.res_a1 = e.res_a1 ? *e.res_a1 : *e.res_a0, p.res_w = e.res_w->map;
      p.res_kb = (e.res_w->K + 63) / 64, p.res_split = e.res_a1 ? e.res_a0_ch / 64 : p.res_kb, p.res_k_true = e.res_w->K;
      p.res_bias = e.res_w->bias;
    }
    LS_CUDA(launch_conv_gemm(a0, a1 ? *a1 : a0, w.map, p, num_sms_, s));
  };
  auto f32 = [&](size_t off) { return arena_.ptr<float>(off); };

  float* u = ws<float>(o_u_);
  float* r = ws<float>(o_r_);

  // One resnet + n_blocks transformer blocks (decoder.py:437-452 / 459-473 / 475-491).
  // `a0`(+`a1`) = masked bf16 input; `tail` = bf16 buffer that receives the masked group output.
  auto group = [&](int gi, const ActMaps& a0, const ActMaps* a1, int a0_ch, void* tail) {
    const GroupW& g = groups_[gi];
    const bool head_via_conv = fused_blocks_ && causal_ && head_via_conv_enabled();
    if (!causal_) {
      // matcha Block1D (decoder.py:32-43): conv3 (pad 1) -> GroupNorm(8) over the utterance's frames -> Mish -> mask.  The
      // statistics span the whole utterance, so the convolution writes its raw fp32 output and two bandwidth kernels
      // (statistics, apply) follow; the time-embedding add and the residual add ride in the apply pass.
      float* raw = ws<float>(o_raw_);
      float2* gst = ws<float2>(o_gn_);
      {
        Epi e;
        e.out0 = raw, e.out0_dtype = OUT_F32, e.zero_skipped = 1;
        gemm(a0, a1, a0_ch, g.res.conv1, e);
      }
      LS_CUDA(launch_groupnorm_mish(raw, gst, f32(g.res.ln1g), f32(g.res.ln1b), temb + (long long)gi * C_, temb_bstride, nullptr,
                                    nullptr, ws<__nv_bfloat16>(o_hA_), B2, C_, T, 8, lengths, fp16_, s));
      {  // res_conv(x*mask)
        Epi e;
        e.out0 = r, e.out0_dtype = OUT_F32, e.zero_skipped = 1;
        gemm(a0, a1, a0_ch, g.res.res, e);
      }
      {
        Epi e;
        e.out0 = raw, e.out0_dtype = OUT_F32, e.zero_skipped = 1;
        gemm(pl.hA, nullptr, 0, g.res.conv2, e);
      }
      LS_CUDA(launch_groupnorm_mish(raw, gst, f32(g.res.ln2g), f32(g.res.ln2b), nullptr, 0, r, u, nullptr, B2, C_, T, 8, lengths,
                                    fp16_, s));
    } else {
    {  // block1: conv3 -> LN -> Mish -> mask, then + Linear(Mish(temb))   (matcha decoder.py:57-58)
      Epi e;
      e.act = ACT_LN_MISH, e.ln_g = f32(g.res.ln1g), e.ln_b = f32(g.res.ln1b);
      e.temb = temb + (long long)gi * C_, e.temb_bstride = temb_bstride;
      e.out1 = ws<void>(o_hA_), e.out1_mode = OUT1_COPY;
      gemm(a0, a1, a0_ch, g.res.conv1, e);
    }
    // res_conv(x*mask): when every CTA has at most one tile (the headline shape: 125 tiles on 148 SMs) it rides in
    // block2's launch as a second GEMM into the idle second accumulator (no launch, no fp32 round trip of r through
    // global memory); otherwise its own launch
    const bool fuse_res = conv_fuse_res_enabled() && (long long)B2 * ((T + 127) / 128) <= num_sms_ && g.res.res.bias != nullptr &&
                          g.res.res.block_n == g.res.conv2.block_n && g.res.conv2.N == g.res.conv2.block_n &&
                          (a1 == nullptr || a0_ch % 64 == 0);
    if (!fuse_res) {
      Epi e;
      e.out0 = r, e.out0_dtype = OUT_F32;
      gemm(a0, a1, a0_ch, g.res.res, e);
    }
    {  // block2 + residual -> residual stream u (fp32) and LayerNorm(norm1 of first block) in bf16
      Epi e;
      e.act = ACT_LN_MISH, e.ln_g = f32(g.res.ln2g), e.ln_b = f32(g.res.ln2b);
      if (fuse_res) {
        e.addend_dtype = ADD_GEMM, e.res_w = &g.res.res, e.res_a0 = &a0.k1, e.res_a1 = a1 ? &a1->k1 : nullptr, e.res_a0_ch = a0_ch;
      } else {
        e.addend = r, e.addend_dtype = OUT_F32;
      }
      e.out0 = u, e.out0_dtype = OUT_F32;
      // LayerNorm(norm1 of the first block) rides in this epilogue unless the fused kernel's head launch computes it from u
      if (!fused_blocks_ || head_via_conv)
        e.out1 = ws<void>(o_nrm_), e.out1_mode = OUT1_LN, e.p1_a = f32(g.tb[0].n1g), e.p1_b = f32(g.tb[0].n1b);
      gemm(pl.hA, nullptr, 0, g.res.conv2, e);
    }
    }  // causal
    auto attention = [&]() {
      AttnParams ap{};
      ap.B = B2, ap.T = T, ap.H = heads_, ap.lengths = lengths, ap.chunk = streaming ? chunk_ : 0;
      ap.scale_log2e = 0.125f * 1.4426950408889634f, ap.fp16 = fp16_;
      ap.out = ws<__nv_bfloat16>(o_att_);
      LS_CUDA(launch_attention(pl.qkv_attn, ap, s));
    };
    if (fused_blocks_) {
      // QKV of the first block: LayerNorm(norm1) + projection straight from u (head mode of the fused kernel), or -- causal
      // estimator -- the projection alone as a conv_gemm launch over the LayerNorm output block2's epilogue wrote (750
      // tiles pipelined over persistent CTAs instead of 125 single-tile CTAs that each stream the whole weight);
      // every later QKV comes out of the previous block's launch
      if (head_via_conv) {
        Epi e;
        e.out1 = ws<void>(o_qkv_), e.out1_mode = OUT1_COPY;
        gemm(pl.nrm, nullptr, 0, g.tb[0].qkv, e);
      } else {
        TBlockParams tp{};
        tp.R = B2 * T, tp.T = T, tp.lengths = lengths, tp.vec = arena_.ptr<float>(g.head_vec), tp.tail_mode = 2, tp.fp16 = fp16_, tp.no_skip = causal_ ? 0 : 1;
        TBlockMaps tm;
        tm.att = pl.att_flat, tm.wo = g.tb[0].m_out, tm.w1 = g.tb[0].m_ff1, tm.w2 = g.tb[0].m_ff2;  // unused in head mode
        tm.wqkv = g.tb[0].m_qkv, tm.u = pl.u_flat, tm.qkv_out = pl.qkv_flat, tm.tail_out = pl.tail_hB;
        LS_CUDA(launch_tblock(tm, tp, num_sms_, s));
      }
      for (int j = 0; j < n_blocks_; ++j) {
        const TBlockW& t = g.tb[j];
        attention();
        const bool last = j + 1 == n_blocks_;
        TBlockParams tp{};
        tp.R = B2 * T, tp.T = T, tp.lengths = lengths, tp.vec = arena_.ptr<float>(t.vec), tp.tail_mode = last ? 1 : 0, tp.fp16 = fp16_, tp.no_skip = causal_ ? 0 : 1;
        TBlockMaps tm;
        tm.att = pl.att_flat, tm.wo = t.m_out, tm.w1 = t.m_ff1, tm.w2 = t.m_ff2;
        tm.wqkv = last ? t.m_qkv : g.tb[j + 1].m_qkv;
        tm.u = pl.u_flat, tm.qkv_out = pl.qkv_flat;
        tm.tail_out = tail == ws<void>(o_skip_) ? pl.tail_skip : pl.tail_hB;
        LS_CUDA(launch_tblock(tm, tp, num_sms_, s));
      }
      return;
    }
    for (int j = 0; j < n_blocks_; ++j) {
      const TBlockW& t = g.tb[j];
      {
        Epi e;
        e.out1 = ws<void>(o_qkv_), e.out1_mode = OUT1_COPY;
        gemm(pl.nrm, nullptr, 0, t.qkv, e);
      }
      {
        AttnParams ap{};
        ap.B = B2, ap.T = T, ap.H = heads_, ap.lengths = lengths, ap.chunk = streaming ? chunk_ : 0;
        ap.scale_log2e = 0.125f * 1.4426950408889634f, ap.fp16 = fp16_;
        ap.out = ws<__nv_bfloat16>(o_att_);
        LS_CUDA(launch_attention(pl.qkv_attn, ap, s));
      }
      {  // to_out + residual, then LayerNorm(norm3)
        Epi e;
        e.addend = u, e.addend_dtype = OUT_F32, e.out0 = u, e.out0_dtype = OUT_F32;
        e.out1 = ws<void>(o_nrm_), e.out1_mode = OUT1_LN, e.p1_a = f32(t.n3g), e.p1_b = f32(t.n3b);
        gemm(pl.att, nullptr, 0, t.out, e);
      }
      {  // FF in: Linear + exact GELU
        Epi e;
        e.act = ACT_GELU, e.out1 = ws<void>(o_ff_), e.out1_mode = OUT1_COPY;
        gemm(pl.nrm, nullptr, 0, t.ff1, e);
      }
      {  // FF out + residual; feeds either the next block's norm1 or (masked copy) the conv after the group
        Epi e;
        e.addend = u, e.addend_dtype = OUT_F32;
        if (j + 1 < n_blocks_) {
          e.out0 = u, e.out0_dtype = OUT_F32;
          e.out1 = ws<void>(o_nrm_), e.out1_mode = OUT1_LN, e.p1_a = f32(g.tb[j + 1].n1g), e.p1_b = f32(g.tb[j + 1].n1b);
        } else {
          e.out1 = tail, e.out1_mode = OUT1_COPY;
        }
        gemm(pl.ff, nullptr, 0, t.ff2, e);
      }
    }
  };

  // down (decoder.py:437-455): group output is both the skip connection and the input of the conv
  group(0, pl.xin, nullptr, 0, ws<void>(o_skip_));
  {
    Epi e;
    e.out1 = ws<void>(o_hB_), e.out1_mode = OUT1_COPY, e.zero_skipped = causal_ ? 0 : 1;  // (non-causal: the next conv reads row `len`)
    gemm(pl.skip, nullptr, 0, down_conv_, e);
  }
  for (int i = 0; i < n_mid_; ++i) group(1 + i, pl.hB, nullptr, 0, ws<void>(o_hB_));
  // up (decoder.py:475-493): channel concat [x, skip] becomes a K split over two tensor maps
  group(1 + n_mid_, pl.hB, &pl.skip, C_, ws<void>(o_hB_));
  {
    Epi e;
    e.out1 = ws<void>(o_hA_), e.out1_mode = OUT1_COPY, e.zero_skipped = causal_ ? 0 : 1;
    gemm(pl.hB, nullptr, 0, up_conv_, e);
  }
  if (!causal_) {  // final_block = Block1D (decoder.py:195,290)
    Epi e;
    e.out0 = ws<float>(o_raw_), e.out0_dtype = OUT_F32, e.zero_skipped = 1;
    gemm(pl.hA, nullptr, 0, final_conv_, e);
    LS_CUDA(launch_groupnorm_mish(ws<float>(o_raw_), ws<float2>(o_gn_), f32(final_lng_), f32(final_lnb_), nullptr, 0, nullptr,
                                  nullptr, ws<__nv_bfloat16>(o_hB_), B2, C_, T, 8, lengths, fp16_, s));
  } else {  // final_block (decoder.py:494)
    Epi e;
    e.act = ACT_LN_MISH, e.ln_g = f32(final_lng_), e.ln_b = f32(final_lnb_);
    e.out1 = ws<void>(o_hB_), e.out1_mode = OUT1_COPY;
    gemm(pl.hA, nullptr, 0, final_conv_, e);
  }
  {  // final_proj (decoder.py:495-496) -> v fp32 [B2][T][80], masked
    Epi e;
    e.out0 = ws<void>(o_v_), e.out0_dtype = OUT_F32, e.zero_skipped = 1;
    gemm(pl.hB, nullptr, 0, final_proj_, e);
  }
  (void)inner;
}

void FlowEngine::time_embed(const float* t_dev, const float* t_host, int nt, cudaStream_t s) {
  TimeEmbedParams tp{};
  tp.t = t_dev, tp.t_host = t_host, tp.scratch = ws<float>(o_tscr_), tp.freqs = arena_.ptr<float>(freqs_);
  tp.w1 = arena_.ptr<float>(w1_), tp.b1 = arena_.ptr<float>(b1_);
  tp.w2 = arena_.ptr<float>(w2_), tp.b2 = arena_.ptr<float>(b2_);
  tp.wr = arena_.ptr<float>(wr_), tp.br = arena_.ptr<float>(br_);
  tp.out = ws<float>(o_temb_);
  tp.nt = nt, tp.in_dim = in_ch_, tp.hid = hid_, tp.n_res = (int)groups_.size(), tp.out_dim = C_;
  LS_CUDA(launch_time_embed(tp, s));
}

void FlowEngine::estimator_forward(const float* x, const float* mask, const float* mu, const float* t,
                                   const float* spks, const float* cond, float* out, int rows, int T, bool streaming,
                                   cudaStream_t s) {
  require(rows > 0 && T > 0, "rows and T must be positive");
  LS_CUDA(cudaSetDevice(device_));
  check_sticky();
  ensure_workspace(rows, T, rows, s);
  int* lengths = ws<int>(o_len_);
  __nv_bfloat16* xin = ws<__nv_bfloat16>(o_xin_);
  LS_CUDA(launch_mask_to_lengths(mask, lengths, rows, T, 1, s, bad_mask_dev_));
  const long long bs = (long long)feat_ * T;
  LS_CUDA(launch_pack_nct(x, xin, rows, feat_, T, bs, in_ch_, 0, lengths, s, fp16_));
  LS_CUDA(launch_pack_nct(mu, xin, rows, feat_, T, bs, in_ch_, feat_, lengths, s, fp16_));
  LS_CUDA(launch_pack_bcast(spks, xin, rows, feat_, T, in_ch_, 2 * feat_, lengths, s, fp16_));
  LS_CUDA(launch_pack_nct(cond, xin, rows, feat_, T, bs, in_ch_, 3 * feat_, lengths, s, fp16_));
  time_embed(t, nullptr, rows, s);
  run_estimator(rows, T, ws<float>(o_temb_), (long long)groups_.size() * C_, streaming, s);
  LS_CUDA(launch_unpack_nct(ws<float>(o_v_), out, rows, feat_, T, lengths, s));
}

void FlowEngine::solve(const float* mu, const float* mask, const float* spks, const float* cond, const float* noise,
                       long long noise_stride, const float* t_span, int n_steps, float temperature, float cfg_rate,
                       bool streaming, float* out, int B, int T, cudaStream_t s) {
  require(B > 0 && T > 0 && n_steps > 0, "B, T and n_timesteps must be positive");
  require(noise_stride >= T, "noise buffer shorter than T (reference: rand_noise holds 15000 frames)");
  LS_CUDA(cudaSetDevice(device_));
  check_sticky();
  const int B2 = 2 * B;
  ensure_workspace(B2, T, n_steps, s);
  int* lengths = ws<int>(o_len_);
  __nv_bfloat16* xin = ws<__nv_bfloat16>(o_xin_);
  float* x_state = ws<float>(o_x_);

  // Euler time grid exactly as flow_matching.py:88,120-124 accumulates it (fp32)
  t_host_.resize(n_steps);
  dt_host_.resize(n_steps);
  {
    float t = t_span[0], dt = t_span[1] - t_span[0];
    for (int step = 1; step <= n_steps; ++step) {
      t_host_[step - 1] = t;
      dt_host_[step - 1] = dt;
      t = t + dt;
      if (step < n_steps) dt = t_span[step + 1] - t;
    }
  }
  // up to 64 time values travel inside the kernel parameters (no host -> device copy: the solve stays capturable)
  const bool t_inline = n_steps <= 64;
  if (!t_inline) LS_CUDA(cudaMemcpyAsync(ws<float>(o_t_), t_host_.data(), (size_t)n_steps * 4, cudaMemcpyHostToDevice, s));

  LS_CUDA(launch_mask_to_lengths(mask, lengths, B, T, 2, s, bad_mask_dev_));
  const long long bs = (long long)feat_ * T;
  // conditional half rows [0,B): [x | mu | spks | cond]; unconditional half rows [B,2B): [x | 0 | 0 | 0]
  LS_CUDA(launch_pack_nct(mu, xin, B, feat_, T, bs, in_ch_, feat_, lengths, s, fp16_));
  LS_CUDA(launch_pack_bcast(spks, xin, B, feat_, T, in_ch_, 2 * feat_, lengths, s, fp16_));
  LS_CUDA(launch_pack_nct(cond, xin, B, feat_, T, bs, in_ch_, 3 * feat_, lengths, s, fp16_));
  LS_CUDA(launch_pack_zero(xin + (long long)B * T * in_ch_, B, 3 * feat_, T, in_ch_, feat_, s));
  LS_CUDA(launch_init_state(noise, (int)noise_stride, temperature, x_state, xin, B, feat_, T, in_ch_, lengths, s, fp16_));
  time_embed(t_inline ? nullptr : ws<float>(o_t_), t_host_.data(), n_steps, s);
  const long long per_t = (long long)groups_.size() * C_;
  for (int k = 0; k < n_steps; ++k) {
    run_estimator(B2, T, ws<float>(o_temb_) + k * per_t, 0, streaming, s);
    LS_CUDA(launch_cfg_euler(ws<float>(o_v_), x_state, xin, B, feat_, T, in_ch_, dt_host_[k], cfg_rate, s, fp16_));
  }
  LS_CUDA(launch_unpack_nct(x_state, out, B, feat_, T, lengths, s));
}

}  // namespace ls
