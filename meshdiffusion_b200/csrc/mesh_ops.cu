// Mesh post-ops that follow marching tetrahedra in the reference's getMesh (nvdiffrec/lib/geometry/dmtet.py:283-289):
// smooth vertex normals (nvdiffrec/lib/render/mesh.py:200-227) and MikkTSpace-style tangents (mesh.py:233-277).
// Both are scatter-adds of per-face quantities onto vertices. The sums are accumulated as 2^-40 fixed-point integers
// (integer atomics commute), so the result does not depend on the order the faces arrive in: bitwise reproducible,
// unlike the reference's float scatter_add_ on the GPU.
#include "../../include/meshdiff_b200.h"
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace mdb { void set_last_error(const std::string& msg); }

namespace {

constexpr double kFix = 1099511627776.0;  // 2^40

__device__ __forceinline__ void fix_add(long long* dst, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(dst), static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v) * kFix)));
}

__global__ void face_normals_kernel(const float* __restrict__ v, const long long* __restrict__ f, int F, long long* acc, float* f_nrm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  const long long i0 = f[3 * i], i1 = f[3 * i + 1], i2 = f[3 * i + 2];
  const float ax = v[3 * i1] - v[3 * i0], ay = v[3 * i1 + 1] - v[3 * i0 + 1], az = v[3 * i1 + 2] - v[3 * i0 + 2];
  const float bx = v[3 * i2] - v[3 * i0], by = v[3 * i2 + 1] - v[3 * i0 + 1], bz = v[3 * i2 + 2] - v[3 * i0 + 2];
  // torch.cross: no fused multiply-add across the subtraction
  const float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
  const float ny = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
  const float nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
  if (f_nrm) { f_nrm[3 * i] = nx; f_nrm[3 * i + 1] = ny; f_nrm[3 * i + 2] = nz; }
  const long long idx[3] = {i0, i1, i2};
  for (int k = 0; k < 3; ++k) { fix_add(acc + 3 * idx[k], nx); fix_add(acc + 3 * idx[k] + 1, ny); fix_add(acc + 3 * idx[k] + 2, nz); }
}

__global__ void finish_normals_kernel(const long long* __restrict__ acc, int N, float* v_nrm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float x = (float)((double)acc[3 * i] / kFix), y = (float)((double)acc[3 * i + 1] / kFix), z = (float)((double)acc[3 * i + 2] / kFix);
  float d = x * x + y * y + z * z;
  if (!(d > 1e-20f)) { x = 0.f; y = 0.f; z = 1.f; d = 1.f; }
  const float len = sqrtf(fmaxf(d, 1e-20f));
  v_nrm[3 * i] = x / len; v_nrm[3 * i + 1] = y / len; v_nrm[3 * i + 2] = z / len;
}

__global__ void face_tangents_kernel(const float* __restrict__ v, const long long* __restrict__ tp, const float* __restrict__ uv,
                                     const long long* __restrict__ tt, const long long* __restrict__ tn, int F, long long* acc, int* cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  const long long p0 = tp[3 * i], p1 = tp[3 * i + 1], p2 = tp[3 * i + 2];
  const long long t0 = tt[3 * i], t1 = tt[3 * i + 1], t2 = tt[3 * i + 2];
  const float u1x = uv[2 * t1] - uv[2 * t0], u1y = uv[2 * t1 + 1] - uv[2 * t0 + 1];
  const float u2x = uv[2 * t2] - uv[2 * t0], u2y = uv[2 * t2 + 1] - uv[2 * t0 + 1];
  float denom = __fsub_rn(__fmul_rn(u1x, u2y), __fmul_rn(u1y, u2x));
  denom = denom > 0.f ? fmaxf(denom, 1e-6f) : fminf(denom, -1e-6f);
  float tang[3];
  for (int c = 0; c < 3; ++c) {
    const float e1 = v[3 * p1 + c] - v[3 * p0 + c], e2 = v[3 * p2 + c] - v[3 * p0 + c];
    tang[c] = __fdiv_rn(__fsub_rn(__fmul_rn(e1, u2y), __fmul_rn(e2, u1y)), denom);
  }
  for (int k = 0; k < 3; ++k) {
    const long long n = tn[3 * i + k];
    for (int c = 0; c < 3; ++c) fix_add(acc + 3 * n + c, tang[c]);
    atomicAdd(cnt + n, 1);
  }
}

__global__ void finish_tangents_kernel(const long long* __restrict__ acc, const int* __restrict__ cnt, const float* __restrict__ nrm, int N, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float c = (float)cnt[i];
  float t[3], n[3];
  for (int k = 0; k < 3; ++k) { t[k] = (float)((double)acc[3 * i + k] / kFix) / c; n[k] = nrm[3 * i + k]; }
  float len = sqrtf(fmaxf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2], 1e-20f));
  for (int k = 0; k < 3; ++k) t[k] /= len;
  const float d = t[0] * n[0] + t[1] * n[1] + t[2] * n[2];
  for (int k = 0; k < 3; ++k) t[k] = t[k] - d * n[k];
  len = sqrtf(fmaxf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2], 1e-20f));
  for (int k = 0; k < 3; ++k) out[3 * i + k] = t[k] / len;
}

int fail(const std::string& m) { mdb::set_last_error(m); return 1; }
int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail(std::string("mdb mesh op launch: ") + cudaGetErrorString(e));
}

}  // namespace

extern "C" {

int mdb_mesh_auto_normals(const float* v_pos, const long long* faces, int n_verts, int n_faces, float* v_nrm, float* f_nrm,
                          long long* scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_verts <= 0) return 0;
  if (cudaMemsetAsync(scratch, 0, (size_t)n_verts * 3 * sizeof(long long), s) != cudaSuccess) return fail("mdb: memset failed");
  if (n_faces > 0) face_normals_kernel<<<(n_faces + 255) / 256, 256, 0, s>>>(v_pos, faces, n_faces, scratch, f_nrm);
  finish_normals_kernel<<<(n_verts + 255) / 256, 256, 0, s>>>(scratch, n_verts, v_nrm);
  return check_launch();
}

int mdb_mesh_compute_tangents(const float* v_pos, const long long* t_pos_idx, const float* v_tex, const long long* t_tex_idx,
                              const float* v_nrm, const long long* t_nrm_idx, int n_nrm, int n_faces, float* v_tng,
                              long long* scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_nrm <= 0) return 0;
  // scratch: [n_nrm][3] int64 sums followed by [n_nrm] int32 counts
  int* cnt = reinterpret_cast<int*>(scratch + (size_t)n_nrm * 3);
  if (cudaMemsetAsync(scratch, 0, (size_t)n_nrm * 3 * sizeof(long long) + (size_t)n_nrm * sizeof(int), s) != cudaSuccess) return fail("mdb: memset failed");
  if (n_faces > 0) face_tangents_kernel<<<(n_faces + 255) / 256, 256, 0, s>>>(v_pos, t_pos_idx, v_tex, t_tex_idx, t_nrm_idx, n_faces, scratch, cnt);
  finish_tangents_kernel<<<(n_nrm + 255) / 256, 256, 0, s>>>(scratch, cnt, v_nrm, n_nrm, v_tng);
  return check_launch();
}

}  // extern "C"
