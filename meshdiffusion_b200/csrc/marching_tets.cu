// Batched marching tetrahedra with the reference's exact output ordering (nvdiffrec/lib/geometry/dmtet.py:105-163).
//
// The reference builds, per call, the lexicographically sorted unique edge list of the *valid* tets with
// torch.unique and numbers the crossing edges by their rank in that list. The set of crossing edges of valid tets
// equals the set of crossing edges of the whole grid (a tet owning a crossing edge has mixed occupancy, hence is
// valid), and ranks inside a sorted subsequence are preserved, so the same numbering is obtained from a STATIC
// sorted edge table of the tet grid built once in mdb_marching_tets_prepare:
//     vertex id of edge e  =  exclusive prefix sum of crossing flags over the globally sorted edge table.
// Faces follow the reference's order: all 1-triangle tets in tet order, then all 2-triangle tets in tet order.
// Per call the work is: flag kernels + three exclusive scans + emit kernels; one thread per edge / tet with
// coalesced table reads, all samples of a batch in one launch (blockIdx.y = sample). HBM-bound: ~7 MB / sample
// at R=64 (tets 2.5 MB + tet->edge table 3.8 MB + sdf/pos 0.5 MB).
#include "../../include/meshdiff_b200.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace mdbmt {

#define MT_CHECK(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) throw std::runtime_error(std::string("mdb MT: ") + cudaGetErrorString(_e) + " (" #expr ")"); \
  } while (0)

// dmtet.py:34-53
__constant__ int8_t c_tri_table[16][6] = {
    {-1, -1, -1, -1, -1, -1}, {1, 0, 2, -1, -1, -1}, {4, 0, 3, -1, -1, -1}, {1, 4, 2, 1, 3, 4},
    {3, 1, 5, -1, -1, -1},    {2, 3, 0, 2, 5, 3},    {1, 4, 0, 1, 5, 4},    {4, 2, 5, -1, -1, -1},
    {4, 5, 2, -1, -1, -1},    {4, 1, 0, 4, 5, 1},    {3, 2, 0, 3, 5, 2},    {1, 3, 5, -1, -1, -1},
    {4, 1, 2, 4, 3, 1},       {3, 0, 4, -1, -1, -1}, {2, 0, 1, -1, -1, -1}, {-1, -1, -1, -1, -1, -1}};
__constant__ int8_t c_num_tri[16] = {0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0};

struct Handle {
  int F = 0, Nv = 0, E = 0, max_batch = 0;
  int* d_tets = nullptr;      // [F][4]
  int2* d_edges = nullptr;    // [E] sorted (a<b), lexicographic
  int* d_tet_edges = nullptr; // [F][6] edge ids in base_tet_edges order
  // per-call workspace (sized for max_batch)
  uint32_t *d_eflag = nullptr, *d_escan = nullptr;   // [B][E]
  uint32_t *d_t1 = nullptr, *d_t1scan = nullptr;     // [B][F] (ntri==1)
  uint32_t *d_t2 = nullptr, *d_t2scan = nullptr;     // [B][F] (ntri==2)
  uint32_t *d_vflag = nullptr, *d_vscan = nullptr;   // [B][Nv]
  uint8_t* d_tetidx = nullptr;                       // [B][F]
  uint32_t* d_partials = nullptr;                    // scan scratch
  int* d_counts = nullptr;                           // [B][4]: nverts, n1, n2, nvalidverts
  int* h_counts = nullptr;                           // pinned
  int last_batch = 0;
};

constexpr int SCAN_TILE = 2048;   // elements per block (256 threads x 8)

// ---------------------------------------------------------------- exclusive scan of uint32 rows (3 kernels)
__global__ void scan_tile_sums(const uint32_t* __restrict__ in, uint32_t* __restrict__ partials, int n, int tiles) {
  __shared__ uint32_t red[8];
  const int row = blockIdx.y, tile = blockIdx.x;
  const uint32_t* p = in + (size_t)row * n;
  uint32_t s = 0;
  const int base = tile * SCAN_TILE;
  for (int i = threadIdx.x; i < SCAN_TILE; i += 256) { const int k = base + i; if (k < n) s += p[k]; }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += red[w]; partials[(size_t)row * tiles + tile] = t; }
}
__global__ void scan_partials(uint32_t* partials, int tiles, int* totals, int total_slot, int slots) {
  // one block per row; sequential chunks of 1024 with a running carry
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  const int row = blockIdx.x;
  uint32_t* p = partials + (size_t)row * tiles;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += 1024) {
    const int k = base + threadIdx.x;
    const uint32_t v = k < tiles ? p[k] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const uint32_t incl = sh[threadIdx.x];
    if (k < tiles) p[k] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && totals) totals[row * slots + total_slot] = (int)carry;
}
__global__ void scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ partials,
                           int n, int tiles) {
  __shared__ uint32_t sh[256];
  const int row = blockIdx.y, tile = blockIdx.x;
  const uint32_t* p = in + (size_t)row * n;
  uint32_t* q = out + (size_t)row * n;
  const int base = tile * SCAN_TILE + threadIdx.x * 8;
  uint32_t v[8], s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = (base + i < n) ? p[base + i] : 0; s += v[i]; }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = partials[(size_t)row * tiles + tile] + sh[threadIdx.x] - s;
#pragma unroll
  for (int i = 0; i < 8; ++i) { if (base + i < n) q[base + i] = run; run += v[i]; }
}

static void exclusive_scan_rows(const uint32_t* in, uint32_t* out, uint32_t* partials, int n, int rows, int* totals,
                                int total_slot, int slots, cudaStream_t s) {
  const int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_tile_sums<<<dim3(tiles, rows), 256, 0, s>>>(in, partials, n, tiles);
  scan_partials<<<rows, 1024, 0, s>>>(partials, tiles, totals, total_slot, slots);
  scan_apply<<<dim3(tiles, rows), 256, 0, s>>>(in, out, partials, n, tiles);
}

// ---------------------------------------------------------------- per-sample kernels
__global__ void edge_flags_kernel(const int2* __restrict__ edges, const float* __restrict__ sdf, uint32_t* __restrict__ eflag,
                                  int E, int Nv) {
  const int b = blockIdx.y;
  const float* s = sdf + (size_t)b * Nv;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int2 ab = edges[e];
    eflag[(size_t)b * E + e] = ((s[ab.x] > 0.f) != (s[ab.y] > 0.f)) ? 1u : 0u;  // occ = sdf > 0 (dmtet.py:107)
  }
}

__global__ void tet_flags_kernel(const int* __restrict__ tets, const float* __restrict__ sdf, uint8_t* __restrict__ tetidx,
                                 uint32_t* __restrict__ t1, uint32_t* __restrict__ t2, uint32_t* __restrict__ vflag, int F, int Nv) {
  const int b = blockIdx.y;
  const float* s = sdf + (size_t)b * Nv;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < F; t += gridDim.x * blockDim.x) {
    const int4 v = reinterpret_cast<const int4*>(tets)[t];
    const int idx = (s[v.x] > 0.f ? 1 : 0) | (s[v.y] > 0.f ? 2 : 0) | (s[v.z] > 0.f ? 4 : 0) | (s[v.w] > 0.f ? 8 : 0);
    const int nt = c_num_tri[idx];
    tetidx[(size_t)b * F + t] = (uint8_t)idx;
    t1[(size_t)b * F + t] = nt == 1;
    t2[(size_t)b * F + t] = nt == 2;
    if (nt > 0) {  // valid_vert_idx = unique(tets with >= 1 triangle) (dmtet.py:161)
      uint32_t* vf = vflag + (size_t)b * Nv;
      vf[v.x] = 1; vf[v.y] = 1; vf[v.z] = 1; vf[v.w] = 1;
    }
  }
}

// verts[vid] = (p_a * (-s_b) + p_b * s_a) / (s_a - s_b) in the reference's operation order (dmtet.py:125-132)
__global__ void emit_verts_kernel(const int2* __restrict__ edges, const uint32_t* __restrict__ eflag,
                                  const uint32_t* __restrict__ escan, const float* __restrict__ pos, long long pos_bstride,
                                  const float* __restrict__ sdf, float* __restrict__ verts, const long long* __restrict__ vert_off,
                                  int E, int Nv) {
  const int b = blockIdx.y;
  const float* s = sdf + (size_t)b * Nv;
  const float* p = pos + (size_t)b * pos_bstride;
  float* out = verts + vert_off[b] * 3;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    if (!eflag[(size_t)b * E + e]) continue;
    const int2 ab = edges[e];
    const float sa = s[ab.x], sb = -s[ab.y];
    const float den = __fadd_rn(sa, sb);
    const float wa = __fdiv_rn(sb, den), wb = __fdiv_rn(sa, den);  // flip: weight of a is (-s_b)/den
    const uint32_t vid = escan[(size_t)b * E + e];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      out[(size_t)vid * 3 + k] = __fadd_rn(__fmul_rn(p[(size_t)ab.x * 3 + k], wa), __fmul_rn(p[(size_t)ab.y * 3 + k], wb));
  }
}

__global__ void emit_faces_kernel(const int* __restrict__ tet_edges, const uint8_t* __restrict__ tetidx,
                                  const uint32_t* __restrict__ t1scan, const uint32_t* __restrict__ t2scan,
                                  const uint32_t* __restrict__ escan, const int* __restrict__ counts,
                                  long long* __restrict__ faces, long long* __restrict__ uv_idx, long long* __restrict__ f2t,
                                  const long long* __restrict__ face_off, int F, int E) {
  const int b = blockIdx.y;
  const int n1 = counts[b * 4 + 1];
  long long* fo = faces + face_off[b] * 3;
  long long* uo = uv_idx + face_off[b] * 3;
  long long* to = f2t + face_off[b];
  const uint32_t* es = escan + (size_t)b * E;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < F; t += gridDim.x * blockDim.x) {
    const int idx = tetidx[(size_t)b * F + t];
    const int nt = c_num_tri[idx];
    if (nt == 0) continue;
    const long long f0 = nt == 1 ? (long long)t1scan[(size_t)b * F + t] : (long long)n1 + 2LL * t2scan[(size_t)b * F + t];
    int eid[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) eid[k] = tet_edges[(size_t)t * 6 + k];
    for (int j = 0; j < nt; ++j) {
      const long long f = f0 + j;
#pragma unroll
      for (int k = 0; k < 3; ++k) fo[f * 3 + k] = (long long)es[eid[c_tri_table[idx][j * 3 + k]]];
      // map_uv (dmtet.py:70-99): face_gidx = 2*t + j -> tet cell t, triangle j
      uo[f * 3 + 0] = 4LL * t;
      uo[f * 3 + 1] = 4LL * t + j + 1;
      uo[f * 3 + 2] = 4LL * t + j + 2;
      to[f] = t;
    }
  }
}

__global__ void emit_valid_verts_kernel(const uint32_t* __restrict__ vflag, const uint32_t* __restrict__ vscan,
                                        long long* __restrict__ out, const long long* __restrict__ vv_off, int Nv) {
  const int b = blockIdx.y;
  long long* o = out + vv_off[b];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < Nv; v += gridDim.x * blockDim.x)
    if (vflag[(size_t)b * Nv + v]) o[vscan[(size_t)b * Nv + v]] = v;
}

static int grid1d(int n) { int g = (n + 255) / 256; return g > 148 * 4 ? 148 * 4 : (g < 1 ? 1 : g); }

}  // namespace mdbmt

using namespace mdbmt;

namespace mdb { void set_last_error(const std::string& msg); }  // api.cu

#define MT_API_BEGIN try {
#define MT_API_END                                                              \
  }                                                                             \
  catch (const std::exception& e) { mdb::set_last_error(e.what()); return 1; }  \
  return 0;

// map_uv's static atlas (dmtet.py:70-88): N = ceil(sqrt(F)) cells, 4 corners per cell.
__global__ void uvs_kernel(float* __restrict__ uvs, int N) {
  const float start = 0.f, end = 1.f - (1.f / (float)N);
  const float step = (end - start) / (float)(N - 1);
  const float pad = 0.9f / (float)N;
  const long long total = (long long)N * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int yi = (int)(i / N), xi = (int)(i % N);
    // torch.linspace evaluates symmetrically from both ends
    const float tx = xi < N / 2 ? start + step * (float)xi : end - step * (float)(N - 1 - xi);
    const float ty = yi < N / 2 ? start + step * (float)yi : end - step * (float)(N - 1 - yi);
    float* o = uvs + i * 8;
    o[0] = tx;       o[1] = ty;
    o[2] = tx + pad; o[3] = ty;
    o[4] = tx + pad; o[5] = ty + pad;
    o[6] = tx;       o[7] = ty + pad;
  }
}

extern "C" {

int mdb_marching_tets_prepare(const int* tets_host, int F, int Nv, int max_batch, void** handle) {
  MT_API_BEGIN
  auto* h = new Handle;
  h->F = F; h->Nv = Nv; h->max_batch = max_batch;
  // static sorted unique edge table + tet -> edge ids (base_tet_edges = [0,1, 0,2, 0,3, 1,2, 1,3, 2,3], dmtet.py:54)
  static const int be[12] = {0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3};
  std::vector<uint64_t> keys((size_t)F * 6);
  for (int t = 0; t < F; ++t)
    for (int k = 0; k < 6; ++k) {
      uint32_t a = (uint32_t)tets_host[t * 4 + be[2 * k]], b = (uint32_t)tets_host[t * 4 + be[2 * k + 1]];
      if (a > b) std::swap(a, b);
      keys[(size_t)t * 6 + k] = ((uint64_t)a << 32) | b;
    }
  std::vector<uint64_t> uniq(keys);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  h->E = (int)uniq.size();
  std::vector<int2> edges(h->E);
  for (int e = 0; e < h->E; ++e) edges[e] = make_int2((int)(uniq[e] >> 32), (int)(uniq[e] & 0xffffffffu));
  std::vector<int> te((size_t)F * 6);
  for (size_t i = 0; i < keys.size(); ++i)
    te[i] = (int)(std::lower_bound(uniq.begin(), uniq.end(), keys[i]) - uniq.begin());
  const size_t B = max_batch;
  MT_CHECK(cudaMalloc(&h->d_tets, (size_t)F * 16));
  MT_CHECK(cudaMalloc(&h->d_edges, (size_t)h->E * 8));
  MT_CHECK(cudaMalloc(&h->d_tet_edges, (size_t)F * 24));
  MT_CHECK(cudaMemcpy(h->d_tets, tets_host, (size_t)F * 16, cudaMemcpyHostToDevice));
  MT_CHECK(cudaMemcpy(h->d_edges, edges.data(), (size_t)h->E * 8, cudaMemcpyHostToDevice));
  MT_CHECK(cudaMemcpy(h->d_tet_edges, te.data(), (size_t)F * 24, cudaMemcpyHostToDevice));
  MT_CHECK(cudaMalloc(&h->d_eflag, B * h->E * 4)); MT_CHECK(cudaMalloc(&h->d_escan, B * h->E * 4));
  MT_CHECK(cudaMalloc(&h->d_t1, B * F * 4)); MT_CHECK(cudaMalloc(&h->d_t1scan, B * F * 4));
  MT_CHECK(cudaMalloc(&h->d_t2, B * F * 4)); MT_CHECK(cudaMalloc(&h->d_t2scan, B * F * 4));
  MT_CHECK(cudaMalloc(&h->d_vflag, B * Nv * 4)); MT_CHECK(cudaMalloc(&h->d_vscan, B * Nv * 4));
  MT_CHECK(cudaMalloc(&h->d_tetidx, B * F));
  const int maxn = std::max(std::max(h->E, F), Nv);
  MT_CHECK(cudaMalloc(&h->d_partials, B * ((maxn + SCAN_TILE - 1) / SCAN_TILE + 1) * 4));
  MT_CHECK(cudaMalloc(&h->d_counts, B * 4 * sizeof(int)));
  MT_CHECK(cudaMallocHost(&h->h_counts, B * 4 * sizeof(int)));
  *handle = h;
  MT_API_END
}

void mdb_marching_tets_destroy(void* handle) {
  auto* h = static_cast<Handle*>(handle);
  if (!h) return;
  cudaFree(h->d_tets); cudaFree(h->d_edges); cudaFree(h->d_tet_edges);
  cudaFree(h->d_eflag); cudaFree(h->d_escan); cudaFree(h->d_t1); cudaFree(h->d_t1scan);
  cudaFree(h->d_t2); cudaFree(h->d_t2scan); cudaFree(h->d_vflag); cudaFree(h->d_vscan);
  cudaFree(h->d_tetidx); cudaFree(h->d_partials); cudaFree(h->d_counts); cudaFreeHost(h->h_counts);
  delete h;
}

int mdb_marching_tets_info(void* handle, int* n_edges, int* uv_grid_n) {
  auto* h = static_cast<Handle*>(handle);
  if (n_edges) *n_edges = h->E;
  if (uv_grid_n) *uv_grid_n = (int)std::ceil(std::sqrt((double)((2LL * h->F + 1) / 2)));
  return 0;
}

/* uvs fp32 [N*N*4][2] (N from mdb_marching_tets_info) */
int mdb_marching_tets_uvs(void* handle, float* uvs, void* stream) {
  MT_API_BEGIN
  auto* h = static_cast<Handle*>(handle);
  const int N = (int)std::ceil(std::sqrt((double)((2LL * h->F + 1) / 2)));
  uvs_kernel<<<grid1d(N * N), 256, 0, (cudaStream_t)stream>>>(uvs, N);
  MT_CHECK(cudaGetLastError());
  MT_API_END
}

/* Phase 1: occupancy flags + scans for `batch` samples (sdf [B][Nv]); counts_host[b] = {nverts, nfaces, nvalidverts}.
 * Synchronises the stream (the caller needs the counts to size the outputs). */
int mdb_marching_tets_count(void* handle, const float* sdf, int batch, int* counts_host, void* stream) {
  MT_API_BEGIN
  auto* h = static_cast<Handle*>(handle);
  cudaStream_t s = (cudaStream_t)stream;
  if (batch < 1 || batch > h->max_batch) throw std::runtime_error("mdb MT: batch out of range");
  const int E = h->E, F = h->F, Nv = h->Nv;
  MT_CHECK(cudaMemsetAsync(h->d_vflag, 0, (size_t)batch * Nv * 4, s));
  edge_flags_kernel<<<dim3(grid1d(E), batch), 256, 0, s>>>(h->d_edges, sdf, h->d_eflag, E, Nv);
  tet_flags_kernel<<<dim3(grid1d(F), batch), 256, 0, s>>>(h->d_tets, sdf, h->d_tetidx, h->d_t1, h->d_t2, h->d_vflag, F, Nv);
  exclusive_scan_rows(h->d_eflag, h->d_escan, h->d_partials, E, batch, h->d_counts, 0, 4, s);
  exclusive_scan_rows(h->d_t1, h->d_t1scan, h->d_partials, F, batch, h->d_counts, 1, 4, s);
  exclusive_scan_rows(h->d_t2, h->d_t2scan, h->d_partials, F, batch, h->d_counts, 2, 4, s);
  exclusive_scan_rows(h->d_vflag, h->d_vscan, h->d_partials, Nv, batch, h->d_counts, 3, 4, s);
  MT_CHECK(cudaGetLastError());
  MT_CHECK(cudaMemcpyAsync(h->h_counts, h->d_counts, (size_t)batch * 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
  MT_CHECK(cudaStreamSynchronize(s));
  for (int b = 0; b < batch; ++b) {
    counts_host[b * 3 + 0] = h->h_counts[b * 4 + 0];
    counts_host[b * 3 + 1] = h->h_counts[b * 4 + 1] + 2 * h->h_counts[b * 4 + 2];
    counts_host[b * 3 + 2] = h->h_counts[b * 4 + 3];
  }
  h->last_batch = batch;
  MT_API_END
}

/* Phase 2: emit. pos [B][Nv][3] (pos_batch_stride = 0 shares one vertex array), sdf as in phase 1.
 * Outputs are packed per sample at the element offsets given (device arrays of `batch` int64):
 *   verts fp32 [sum nverts][3], faces / uv_idx int64 [sum nfaces][3], face_to_tet int64 [sum nfaces],
 *   valid_vert_idx int64 [sum nvalidverts]. */
int mdb_marching_tets_extract(void* handle, const float* pos, long long pos_batch_stride, const float* sdf, int batch,
                              float* verts, long long* faces, long long* uv_idx, long long* face_to_tet,
                              long long* valid_vert_idx, const long long* vert_off, const long long* face_off,
                              const long long* vv_off, void* stream) {
  MT_API_BEGIN
  auto* h = static_cast<Handle*>(handle);
  cudaStream_t s = (cudaStream_t)stream;
  if (batch != h->last_batch) throw std::runtime_error("mdb MT: call mdb_marching_tets_count first with the same batch");
  const int E = h->E, F = h->F, Nv = h->Nv;
  emit_verts_kernel<<<dim3(grid1d(E), batch), 256, 0, s>>>(h->d_edges, h->d_eflag, h->d_escan, pos, pos_batch_stride, sdf,
                                                           verts, vert_off, E, Nv);
  emit_faces_kernel<<<dim3(grid1d(F), batch), 256, 0, s>>>(h->d_tet_edges, h->d_tetidx, h->d_t1scan, h->d_t2scan, h->d_escan,
                                                           h->d_counts, faces, uv_idx, face_to_tet, face_off, F, E);
  emit_valid_verts_kernel<<<dim3(grid1d(Nv), batch), 256, 0, s>>>(h->d_vflag, h->d_vscan, valid_vert_idx, vv_off, Nv);
  MT_CHECK(cudaGetLastError());
  MT_API_END
}

}  // extern "C"
