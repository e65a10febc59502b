// Marching cubes of a 0/1 occupancy volume at level 0.5 -- the surface extraction behind meshify_colored_voxel_grid
// (reference utils/voxel_utils.py:53-95, which calls skimage.measure.marching_cubes at :69-72).  sm_100a only.
//
// Canonical output order of this project (scikit-image's Lewiner order cannot be reproduced without scikit-image; see
// DESIGN.md 4.7 and oracle/mesh_oracle.py, whose NumPy restatement this file matches bit for bit):
//   vertices : one per grid edge whose two voxels differ in occupancy, the edge's midpoint, ordered by (flat index of the
//              edge's lower voxel, axis);
//   faces    : per cell in flat order of its origin voxel, triangles from the 256-case table p3d_mc_table.inc (generated
//              by tools/gen_mc_table.py: ambiguous faces separate the occupied corners, counter-clockwise seen from the
//              empty side);
//   normals  : -(g(p) + g(p + e)) normalised, g = central differences of the volume with a replicated border; a vanishing
//              gradient falls back to the edge direction from the occupied to the empty voxel.
// Three passes: count (per-voxel code byte + per-CTA sums), one-CTA scan of the sums, emit (vertices, then faces).
#include "p3d_common.cuh"
#include "p3d_mc_table.inc"

namespace {

constexpr int kMeshThreads = 256;

struct MeshWs {
  uint8_t* code;        // per voxel: bits 0..2 = owns a vertex on the edge along axis 0/1/2, bits 3..5 = triangles of its cell
  int32_t* voff;        // per voxel: index of its first vertex
  int32_t* block_sums;  // per CTA: vertices | triangles << 16, then (after the scan) two arrays of exclusive offsets
  int32_t* block_v;
  int32_t* block_t;
  int64_t n, nb;
};

__host__ __device__ inline size_t mesh_align(size_t x) { return (x + 255) / 256 * 256; }

inline MeshWs mesh_layout(void* ws, int B0, int B1, int B2) {
  MeshWs w;
  w.n = (int64_t)B0 * B1 * B2;
  w.nb = (w.n + kMeshThreads - 1) / kMeshThreads;
  unsigned char* p = static_cast<unsigned char*>(ws);
  w.code = p;
  p += mesh_align((size_t)w.n);
  w.voff = reinterpret_cast<int32_t*>(p);
  p += mesh_align((size_t)w.n * 4);
  w.block_sums = reinterpret_cast<int32_t*>(p);
  p += mesh_align((size_t)w.nb * 4);
  w.block_v = reinterpret_cast<int32_t*>(p);
  p += mesh_align((size_t)w.nb * 4);
  w.block_t = reinterpret_cast<int32_t*>(p);
  return w;
}

// exclusive scan of one packed value per thread over the CTA (vertices in the low 16 bits, triangles in the high 16:
// at most 3 * 256 and 5 * 256, no carry between the halves); returns the CTA total through `total`
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& total) {
  __shared__ uint32_t warp_sums[kMeshThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t base = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kMeshThreads / 32; ++w) {
    const uint32_t s = warp_sums[w];
    if (w < warp) base += s;
    all += s;
  }
  __syncthreads();
  total = all;
  return base + incl - v;
}

__global__ void __launch_bounds__(kMeshThreads)
mesh_count_kernel(const uint8_t* __restrict__ m, int B0, int B1, int B2, uint8_t* __restrict__ code,
                  int32_t* __restrict__ block_sums) {
  const int64_t n = (int64_t)B0 * B1 * B2;
  const int64_t v = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  uint32_t c = 0;
  if (v < n) {
    const int k = (int)(v % B2);
    const int64_t r = v / B2;
    const int j = (int)(r % B1), i = (int)(r / B1);
    const int64_t s0 = (int64_t)B1 * B2, s1 = B2;
    const bool o = m[v] != 0;
    if (i + 1 < B0 && (m[v + s0] != 0) != o) c |= 1u;
    if (j + 1 < B1 && (m[v + s1] != 0) != o) c |= 2u;
    if (k + 1 < B2 && (m[v + 1] != 0) != o) c |= 4u;
    if (i + 1 < B0 && j + 1 < B1 && k + 1 < B2) {
      uint32_t cs = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (m[v + ((q >> 2) & 1) * s0 + ((q >> 1) & 1) * s1 + (q & 1)] != 0) cs |= 1u << q;
      c |= (uint32_t)kMcTriCount[cs] << 3;
    }
    code[v] = (uint8_t)c;
  }
  uint32_t total;
  block_exclusive_scan((uint32_t)__popc(c & 7u) | ((c >> 3) << 16), total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = (int32_t)total;
}

// one CTA: exclusive scans of the per-CTA vertex and triangle counts, totals[0] = vertices, totals[1] = triangles.
// Every thread owns kScanItems consecutive sums per round (the first form, one sum per thread and round, took 1.5 ms for
// the 524 288 CTAs of a 512^3 volume).
constexpr int kScanItems = 8;
__global__ void __launch_bounds__(1024)
mesh_scan_kernel(const int32_t* __restrict__ block_sums, int64_t nb, int32_t* __restrict__ block_v, int32_t* __restrict__ block_t,
                 int64_t* __restrict__ totals) {
  __shared__ long long warp_v[32], warp_t[32];
  __shared__ long long carry_v, carry_t;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_v = 0; carry_t = 0; }
  __syncthreads();
  for (int64_t base = 0; base < nb; base += 1024 * kScanItems) {
    const int64_t b0 = base + (int64_t)threadIdx.x * kScanItems;
    uint32_t s[kScanItems];
    long long iv = 0, it = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
      s[q] = b0 + q < nb ? (uint32_t)block_sums[b0 + q] : 0u;
      iv += s[q] & 0xffffu;
      it += s[q] >> 16;
    }
    const long long ov = iv, ot = it;                        // this thread's totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long tv = __shfl_up_sync(0xffffffffu, iv, d), tt = __shfl_up_sync(0xffffffffu, it, d);
      if (lane >= d) { iv += tv; it += tt; }
    }
    if (lane == 31) { warp_v[warp] = iv; warp_t[warp] = it; }
    __syncthreads();
    long long bv = carry_v, bt = carry_t, av = 0, at = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) { bv += warp_v[w]; bt += warp_t[w]; }
      av += warp_v[w]; at += warp_t[w];
    }
    long long ev = bv + iv - ov, et = bt + it - ot;           // exclusive prefix of this thread's first sum
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
      if (b0 + q < nb) {
        block_v[b0 + q] = (int32_t)ev;                       // the host refuses meshes with 2^31 or more vertices / triangles
        block_t[b0 + q] = (int32_t)et;
      }
      ev += s[q] & 0xffffu;
      et += s[q] >> 16;
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry_v += av; carry_t += at; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = carry_v; totals[1] = carry_t; }
}

__device__ __forceinline__ float vox(const uint8_t* __restrict__ m, int i, int j, int k, int B0, int B1, int B2) {
  i = min(max(i, 0), B0 - 1);
  j = min(max(j, 0), B1 - 1);
  k = min(max(k, 0), B2 - 1);
  return m[((int64_t)i * B1 + j) * B2 + k] != 0 ? 1.f : 0.f;
}
// central differences with a replicated border, float32, one rounding-free step each (differences of 0/1 times 0.5)
__device__ __forceinline__ void gradient(const uint8_t* __restrict__ m, int i, int j, int k, int B0, int B1, int B2, float (&g)[3]) {
  g[0] = __fmul_rn(__fsub_rn(vox(m, i + 1, j, k, B0, B1, B2), vox(m, i - 1, j, k, B0, B1, B2)), 0.5f);
  g[1] = __fmul_rn(__fsub_rn(vox(m, i, j + 1, k, B0, B1, B2), vox(m, i, j - 1, k, B0, B1, B2)), 0.5f);
  g[2] = __fmul_rn(__fsub_rn(vox(m, i, j, k + 1, B0, B1, B2), vox(m, i, j, k - 1, B0, B1, B2)), 0.5f);
}

__global__ void __launch_bounds__(kMeshThreads)
mesh_vertices_kernel(const uint8_t* __restrict__ m, int B0, int B1, int B2, const uint8_t* __restrict__ code,
                     const int32_t* __restrict__ block_sums, const int32_t* __restrict__ block_v, int32_t* __restrict__ voff,
                     float* __restrict__ verts, float* __restrict__ normals) {
  // a CTA without vertices has nothing to write: its voxels' offsets are never read either (a face only refers to voxels
  // that own a vertex) -- most of a monument grid is empty space or solid interior
  if (((uint32_t)block_sums[blockIdx.x] & 0xffffu) == 0u) return;
  const int64_t n = (int64_t)B0 * B1 * B2;
  const int64_t v = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  const uint32_t c = v < n ? code[v] : 0u;
  uint32_t total;
  const uint32_t ex = block_exclusive_scan((uint32_t)__popc(c & 7u), total);
  if (v >= n) return;
  int32_t idx = block_v[blockIdx.x] + (int32_t)ex;
  voff[v] = idx;
  if ((c & 7u) == 0u) return;
  const int k = (int)(v % B2);
  const int64_t r = v / B2;
  const int j = (int)(r % B1), i = (int)(r / B1);
  const bool occ = m[v] != 0;
  float g0[3];
  gradient(m, i, j, k, B0, B1, B2, g0);
#pragma unroll
  for (int axis = 0; axis < 3; ++axis) {
    if (!((c >> axis) & 1u)) continue;
    float g1[3];
    gradient(m, i + (axis == 0), j + (axis == 1), k + (axis == 2), B0, B1, B2, g1);
    float nv[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) nv[q] = -__fadd_rn(g0[q], g1[q]);
    const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nv[0], nv[0]), __fmul_rn(nv[1], nv[1])), __fmul_rn(nv[2], nv[2])));
    if (norm > 0.f) {
#pragma unroll
      for (int q = 0; q < 3; ++q) nv[q] = __fdiv_rn(nv[q], norm);
    } else {                                               // the edge direction from the occupied to the empty voxel
#pragma unroll
      for (int q = 0; q < 3; ++q) nv[q] = q == axis ? (occ ? 1.f : -1.f) : 0.f;
    }
    float* pv = verts + (size_t)idx * 3;
    pv[0] = (float)i + (axis == 0 ? 0.5f : 0.f);
    pv[1] = (float)j + (axis == 1 ? 0.5f : 0.f);
    pv[2] = (float)k + (axis == 2 ? 0.5f : 0.f);
    float* pn = normals + (size_t)idx * 3;
    pn[0] = nv[0]; pn[1] = nv[1]; pn[2] = nv[2];
    ++idx;
  }
}

__global__ void __launch_bounds__(kMeshThreads)
mesh_faces_kernel(const uint8_t* __restrict__ m, int B0, int B1, int B2, const uint8_t* __restrict__ code,
                  const int32_t* __restrict__ block_sums, const int32_t* __restrict__ block_t, const int32_t* __restrict__ voff,
                  int32_t* __restrict__ faces) {
  if (((uint32_t)block_sums[blockIdx.x] >> 16) == 0u) return;
  const int64_t n = (int64_t)B0 * B1 * B2;
  const int64_t v = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  const uint32_t c = v < n ? code[v] : 0u;
  const uint32_t nt = c >> 3;
  uint32_t total;
  const uint32_t ex = block_exclusive_scan(nt, total);
  if (nt == 0u) return;
  const int64_t s0 = (int64_t)B1 * B2, s1 = B2;
  uint32_t cs = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (m[v + ((q >> 2) & 1) * s0 + ((q >> 1) & 1) * s1 + (q & 1)] != 0) cs |= 1u << q;
  int32_t* out = faces + ((size_t)block_t[blockIdx.x] + ex) * 3;
  for (uint32_t t = 0; t < nt; ++t) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = kMcTris[cs][3 * t + q];
      const int64_t u = v + kMcEdgeInfo[e][0] * s0 + kMcEdgeInfo[e][1] * s1 + kMcEdgeInfo[e][2];
      const uint32_t axis = kMcEdgeInfo[e][3];
      out[3 * t + q] = voff[u] + __popc((uint32_t)code[u] & ((1u << axis) - 1u));
    }
  }
}

}  // namespace

P3D_API size_t p3d_mesh_workspace_bytes(int B0, int B1, int B2) {
  if (B0 <= 0 || B1 <= 0 || B2 <= 0) return 0;
  const size_t n = (size_t)B0 * B1 * B2, nb = (n + kMeshThreads - 1) / kMeshThreads;
  return mesh_align(n) + mesh_align(n * 4) + 3 * mesh_align(nb * 4);
}

P3D_API int p3d_mesh_count(const uint8_t* mask, int B0, int B1, int B2, void* workspace, size_t workspace_bytes,
                           int64_t* totals, p3d_stream_t stream) {
  P3D_REQUIRE(B0 > 0 && B1 > 0 && B2 > 0, "mesh_count: bad shape");
  P3D_REQUIRE(mask && workspace && totals, "mesh_count: null pointer");
  if (workspace_bytes < p3d_mesh_workspace_bytes(B0, B1, B2)) {
    p3d::set_error("mesh_count: workspace too small");
    return P3D_E_WORKSPACE;
  }
  const MeshWs w = mesh_layout(workspace, B0, B1, B2);
  P3D_REQUIRE(w.nb < (1ll << 31), "mesh_count: volume too large");
  cudaStream_t st = p3d::as_stream(stream);
  mesh_count_kernel<<<(unsigned)w.nb, kMeshThreads, 0, st>>>(mask, B0, B1, B2, w.code, w.block_sums);
  mesh_scan_kernel<<<1, 1024, 0, st>>>(w.block_sums, w.nb, w.block_v, w.block_t, totals);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_mesh_emit(const uint8_t* mask, int B0, int B1, int B2, void* workspace, size_t workspace_bytes,
                          int64_t n_vertices, int64_t n_faces, float* verts, float* normals, int32_t* faces,
                          p3d_stream_t stream) {
  P3D_REQUIRE(B0 > 0 && B1 > 0 && B2 > 0, "mesh_emit: bad shape");
  P3D_REQUIRE(n_vertices >= 0 && n_faces >= 0 && n_vertices < (1ll << 31) && n_faces < (1ll << 31),
              "mesh_emit: vertex / face counts must fit 32-bit indices");
  P3D_REQUIRE(mask && workspace, "mesh_emit: null pointer");
  P3D_REQUIRE((n_vertices == 0 || (verts && normals)) && (n_faces == 0 || faces), "mesh_emit: null output");
  if (workspace_bytes < p3d_mesh_workspace_bytes(B0, B1, B2)) {
    p3d::set_error("mesh_emit: workspace too small");
    return P3D_E_WORKSPACE;
  }
  const MeshWs w = mesh_layout(workspace, B0, B1, B2);
  cudaStream_t st = p3d::as_stream(stream);
  mesh_vertices_kernel<<<(unsigned)w.nb, kMeshThreads, 0, st>>>(mask, B0, B1, B2, w.code, w.block_sums, w.block_v, w.voff, verts, normals);
  if (n_faces > 0)
    mesh_faces_kernel<<<(unsigned)w.nb, kMeshThreads, 0, st>>>(mask, B0, B1, B2, w.code, w.block_sums, w.block_t, w.voff, faces);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}
