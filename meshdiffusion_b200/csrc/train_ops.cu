// Bandwidth-bound training kernels around the score network (reference: lib/diffusion/losses.py:38-85,
// lib/diffusion/models/ema.py:43-64): the masked DDPM loss with its gradient, global-norm gradient clipping, and a
// single multi-tensor pass that applies clipping + Adam + the EMA update (the reference makes ~5 separate passes over
// the 364 M parameters per step: clip_grad_norm_, Adam's foreach ops, EMA).
// These are the optimiser-side building blocks of the training step (the network backward is unet_train.cu).
#include "../../include/meshdiff_b200.h"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdexcept>
#include <string>

namespace mdb { void set_last_error(const std::string& msg); }

#define TR_API_BEGIN try {
#define TR_API_END                                                              \
  }                                                                             \
  catch (const std::exception& e) { mdb::set_last_error(e.what()); return 1; }  \
  return 0;
#define TR_CHECK(expr)                                                                                         \
  do {                                                                                                         \
    cudaError_t _e = (expr);                                                                                   \
    if (_e != cudaSuccess) throw std::runtime_error(std::string("mdb train: ") + cudaGetErrorString(_e) + " (" #expr ")"); \
  } while (0)

namespace {

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

// losses = (pred - noise)^2 * mask; loss = mean_b(mean_{c,v}(losses)) * V / sum(mask)   (losses.py:69-78)
// d loss / d pred = 2 (pred - noise) mask * coef,  coef = 1 / (B C V) * V / sum(mask)
__global__ void ddpm_loss_kernel(const float* __restrict__ pred, const float* __restrict__ noise, const float* __restrict__ mask,
                                 double coef, double* __restrict__ acc, float* __restrict__ grad, long long V, long long total) {
  __shared__ double sh[32];
  double s = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float m = __ldg(mask + (i % V));
    const float d = pred[i] - noise[i];
    s += (double)(d * d * m);
    if (grad) grad[i] = (float)(2.0 * coef) * d * m;
  }
  const double t = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(acc, t * coef);
}
__global__ void store_scalar_kernel(const double* acc, float* out) { *out = (float)*acc; }

__global__ void sq_norm_kernel(const float* const* __restrict__ grads, const long long* __restrict__ numels, double* __restrict__ acc) {
  __shared__ double sh[32];
  const float* g = grads[blockIdx.y];
  const long long n = numels[blockIdx.y];
  double s = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    s += (double)v * v;
  }
  const double t = block_sum(s, sh);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(acc, t);
}
// clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6))
__global__ void clip_coef_kernel(const double* acc, float max_norm, float* coef, float* total_norm) {
  const float tn = (float)sqrt(*acc);
  if (total_norm) *total_norm = tn;
  const float c = max_norm / (tn + 1e-6f);
  *coef = c < 1.f ? c : 1.f;
}

struct AdamArgs {
  float* const* params; const float* const* grads; float* const* exp_avg; float* const* exp_avg_sq; float* const* ema;
  const long long* numels;
  float lr, beta1, beta2, eps, step_size, inv_bc2_sqrt, one_minus_decay;
  const float* clip_coef;
};
// torch.optim.Adam (no amsgrad, weight_decay 0) followed by ExponentialMovingAverage.update, one read-modify-write pass
__global__ void adam_ema_kernel(AdamArgs a) {
  const int t = blockIdx.y;
  float* p = a.params[t]; const float* g = a.grads[t]; float* m = a.exp_avg[t]; float* v = a.exp_avg_sq[t];
  float* e = a.ema ? a.ema[t] : nullptr;
  const long long n = a.numels[t];
  const float cc = a.clip_coef ? *a.clip_coef : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * cc;
    float mi = m[i], vi = v[i], pi = p[i];
    mi = mi + (gi - mi) * (1.f - a.beta1);                 // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * a.beta2 + (1.f - a.beta2) * gi * gi;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) * a.inv_bc2_sqrt + a.eps;
    pi = pi - a.step_size * (mi / denom);                  // param.addcdiv_(exp_avg, denom, value=-step_size)
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (e) { const float s = e[i]; e[i] = s - a.one_minus_decay * (s - pi); }
  }
}

}  // namespace

extern "C" {

int mdb_ddpm_loss(const float* pred, const float* noise, const float* mask, double mask_sum, float* loss_out,
                  float* grad_pred, double* scratch, int B, int C, long long V, void* stream) {
  TR_API_BEGIN
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)B * C * V;
  const double coef = 1.0 / ((double)B * C * V) * ((double)V / mask_sum);
  TR_CHECK(cudaMemsetAsync(scratch, 0, sizeof(double), s));
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ddpm_loss_kernel<<<(unsigned)blocks, 256, 0, s>>>(pred, noise, mask, coef, scratch, grad_pred, V, total);
  store_scalar_kernel<<<1, 1, 0, s>>>(scratch, loss_out);
  TR_CHECK(cudaGetLastError());
  TR_API_END
}

int mdb_grad_clip_coef(const float* const* grads_dev, const long long* numels_dev, int n, float max_norm, float* coef_out,
                       float* total_norm_out, double* scratch, void* stream) {
  TR_API_BEGIN
  cudaStream_t s = (cudaStream_t)stream;
  TR_CHECK(cudaMemsetAsync(scratch, 0, sizeof(double), s));
  sq_norm_kernel<<<dim3(32, n), 256, 0, s>>>(grads_dev, numels_dev, scratch);
  clip_coef_kernel<<<1, 1, 0, s>>>(scratch, max_norm, coef_out, total_norm_out);
  TR_CHECK(cudaGetLastError());
  TR_API_END
}

int mdb_adam_ema_step(float* const* params_dev, const float* const* grads_dev, float* const* exp_avg_dev,
                      float* const* exp_avg_sq_dev, float* const* ema_dev, const long long* numels_dev, int n, float lr,
                      float beta1, float beta2, float eps, int step, const float* clip_coef_dev, float ema_decay,
                      void* stream) {
  TR_API_BEGIN
  if (step < 1) throw std::runtime_error("mdb: Adam step counter starts at 1");
  AdamArgs a{};
  a.params = params_dev; a.grads = grads_dev; a.exp_avg = exp_avg_dev; a.exp_avg_sq = exp_avg_sq_dev; a.ema = ema_dev;
  a.numels = numels_dev; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  a.step_size = (float)(lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  a.one_minus_decay = 1.f - ema_decay;
  a.clip_coef = clip_coef_dev;
  adam_ema_kernel<<<dim3(64, n), 256, 0, (cudaStream_t)stream>>>(a);
  TR_CHECK(cudaGetLastError());
  TR_API_END
}

// ---- data-parallel gradient exchange over a caller-owned NCCL communicator. NCCL is resolved at call time from the
// library already loaded in the process (the host created the communicator with it), so this shared object carries no
// link-time dependency on a particular libnccl.
__global__ void scale_kernel(float* p, float f, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] *= f;
}

int mdb_allreduce_grads(void* nccl_comm, float* grads, long long numel, int world_size, void* stream) {
  TR_API_BEGIN
  if (!nccl_comm || world_size < 1) throw std::runtime_error("mdb: mdb_allreduce_grads needs a communicator and its size");
  typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static allreduce_fn fn = nullptr;
  if (!fn) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error("mdb: libnccl.so.2 is not loadable");
    fn = reinterpret_cast<allreduce_fn>(dlsym(h, "ncclAllReduce"));
    if (!fn) throw std::runtime_error("mdb: ncclAllReduce not found");
  }
  const int ncclFloat32 = 7, ncclSum = 0;  // nccl.h enums (stable since NCCL 2.0)
  const int rc = fn(grads, grads, (size_t)numel, ncclFloat32, ncclSum, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) throw std::runtime_error("mdb: ncclAllReduce failed with code " + std::to_string(rc));
  if (world_size > 1) {
    scale_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(grads, 1.f / (float)world_size, numel);
    TR_CHECK(cudaGetLastError());
  }
  TR_API_END
}

}  // extern "C"
