// Score-network engine: builds the DDPM 3-D U-Net (reference lib/diffusion/models/ddpm_res64.py:41-199,
// ddpm_res128.py:43-215, layers.py:573-689) as a static plan of tcgen05 GEMM ops + bandwidth kernels over a
// liveness-packed HBM arena, and replays it per denoising step.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "gemm_host.h"
#include "elementwise.cuh"

namespace mdb {

struct UNetConfig {
  int image_size = 64;
  int nf = 128;
  int n_levels = 5;
  int ch_mult[8] = {1, 1, 2, 4, 4, 0, 0, 0};
  int num_res_blocks = 3;
  int level0_blocks = -1;  // ddpm_res128.py:98 uses 2 blocks at level 0; -1 = num_res_blocks
  int n_attn = 1;
  int attn_resolutions[4] = {16, 0, 0, 0};
  int num_channels = 4;
  int stem_ksize = 3;   // 3 (res64) or 5 (res128)
  int use_pos_bias = 1; // ddpm_res64.py:148 adds pos_layer(coords*0) == its bias; res128 does not
  int max_batch = 1;
  int precision = 0;    // 0 = bf16 operands, 1 = tf32 operands (fp32 accumulate in TMEM either way)
};

struct ParamInfo {
  std::string name;
  std::vector<long long> shape;
  long long numel = 0;
  float* d = nullptr;
  bool external = false;  // storage is a slice of a larger buffer
};

class Arena {
 public:
  size_t alloc(size_t bytes);
  void release(size_t off);
  size_t peak() const { return peak_; }
  void reset() { blocks_.clear(); end_ = 0; }
 private:
  struct Block { size_t off, size; bool free; };
  std::vector<Block> blocks_;
  size_t end_ = 0, peak_ = 0;
};

struct Tens {
  size_t off = 0, bytes = 0;
  int C = 0, R = 0;
  long long* stats = nullptr;
  void* ptr = nullptr;
};
typedef std::shared_ptr<Tens> TensP;

class UNet {
 public:
  // dry_only: size the plan and enumerate parameters without touching the GPU (host-side tests)
  explicit UNet(const UNetConfig& cfg, bool dry_only = false);
  ~UNet();
  const std::vector<ParamInfo>& params() const { return params_; }
  // copies `numel` floats into the named parameter (src on host or device)
  void set_param(const std::string& name, const float* src, long long numel, bool src_device, cudaStream_t s);
  void get_param(const std::string& name, float* dst, long long numel, bool dst_device, cudaStream_t s);
  // after (re)loading parameters: derived vectors, constant stem field, weight packing
  void commit(cudaStream_t s);
  // x: fp32 NCDHW [B][Cin][R^3]; labels: fp32 [B]; out: fp32 NCDHW [B][Cin][R^3]
  void forward(const float* x, const float* labels, float* out, int B, cudaStream_t s);
  double flops_per_sample() const { return flops_ / cfg_.max_batch; }
  size_t arena_bytes() const { return arena_bytes_; }
  int num_gemm_launches() const { return (int)gemms_.size(); }
  int num_steps() const { return (int)steps_.size(); }
  const UNetConfig& cfg() const { return cfg_; }
  // per-GEMM timing breakdown of one forward (ms), for profiling
  std::vector<std::pair<std::string, float>> profile(const float* x, const float* labels, float* out, int B, cudaStream_t s);

 private:
  UNetConfig cfg_;
  Precision prec_;
  bool dry_ = true;
  std::vector<ParamInfo> params_;
  std::map<std::string, int> pindex_;
  std::vector<void*> owned_;  // cudaMalloc'd buffers
  Arena arena_;
  char* arena_base_ = nullptr;
  size_t arena_bytes_ = 0;
  long long* stats_base_ = nullptr;
  size_t stats_doubles_ = 0, stats_cursor_ = 0;
  std::vector<std::unique_ptr<GemmOp>> gemms_;
  std::vector<std::unique_ptr<GemmOp>> commit_gemms_;
  struct Step { std::string name; std::function<void(cudaStream_t, int)> fn; };
  std::vector<Step> steps_, commit_steps_;
  bool committed_ = false;
  double flops_ = 0;
  // runtime pointers
  const float* rt_x_ = nullptr;
  const float* rt_labels_ = nullptr;
  float* rt_out_ = nullptr;
  // temb
  float* temb_act_ = nullptr;
  float* dense_w_ = nullptr; float* dense_b_ = nullptr; float* dense_out_ = nullptr;
  int dense_total_ = 0, dense_cursor_ = 0;

  void build();
  float* P(const std::string& name, std::vector<long long> shape, float* external = nullptr);
  void* dmalloc(size_t bytes, bool zero = true);
  TensP new_act(int C, int R, bool stats);
  void release(TensP& t);
  Act act_of(const TensP& t) const;
  GemmOp* new_gemm(const std::string& name, bool commit_time = false);
  void add_step(const std::string& name, std::function<void(cudaStream_t, int)> fn) { if (!dry_) steps_.push_back({name, fn}); }
  struct Scratch { int S = 1; size_t off = 0; float* ptr = nullptr; bool active = false; };
  Scratch split_begin(int R, int N, int cin_total, int taps);
  void split_end(Scratch& s);
  TensP gn(const std::string& pname, const std::vector<TensP>& ins, bool silu);
  TensP resblock(const std::vector<TensP>& ins, int out_ch, int midx);
  TensP attn(const TensP& x, int midx);
  TensP downsample(const TensP& x, int midx);
  TensP upsample(const TensP& x, int midx);
};

}  // namespace mdb
