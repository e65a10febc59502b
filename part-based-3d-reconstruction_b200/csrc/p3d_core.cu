// libp3d_b200: error plumbing, colour<->label conversion, ordered point compaction.
//
// Reference call sites replaced here (paths under the reference root):
//   np.all(x == colour, axis=-1) scans   utils/voxel_utils.py:12-15, utils/mask_utils.py:92-95
//   get_voxel_points_by_parts            utils/voxel_utils.py:7-21
#include <stdarg.h>
#include <string.h>

#include "p3d_common.cuh"

namespace p3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {                       // SMs of the CURRENT device (cached per device, read-mostly)
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev] = n;                   // benign race: every thread writes the same value
  }
  return n;
}

}  // namespace p3d

P3D_API int p3d_version(void) { return 100; }
P3D_API const char* p3d_last_error(void) { return p3d::g_err; }

P3D_API int p3d_device_info(int* sm, int* major, int* minor, int64_t* l2_bytes) {
  int dev = 0, v = 0;
  P3D_CUDA(cudaGetDevice(&dev));
  if (sm) { P3D_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm = v; }
  if (major) { P3D_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *major = v; }
  if (minor) { P3D_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *minor = v; }
  if (l2_bytes) { P3D_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev)); *l2_bytes = v; }
  return P3D_OK;
}

// ---------------------------------------------------------------------------------------------
// RGB -> label.  One thread converts 4 pixels: three aligned 32-bit loads, one 32-bit store.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxPalette = 255;

__device__ __forceinline__ uint32_t match_palette(uint32_t c, const uint32_t* pal, int n) {
  for (int k = 0; k < n; ++k)
    if (pal[k] == c) return (uint32_t)(k + 1);
  return 0u;
}

__global__ void __launch_bounds__(256) rgb_to_labels_kernel(const uint8_t* __restrict__ rgb, int64_t n,
                                                            const uint8_t* __restrict__ palette, int n_colors,
                                                            uint8_t* __restrict__ labels, int vec_ok) {
  __shared__ uint32_t pal[kMaxPalette + 1];
  __shared__ int has_black;
  if (threadIdx.x == 0) has_black = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < n_colors; k += blockDim.x) {
    uint32_t c = palette[3 * k] | (palette[3 * k + 1] << 8) | (palette[3 * k + 2] << 16);
    pal[k] = c;
    if (c == 0) has_black = 1;
  }
  __syncthreads();
  const bool black = has_black != 0;
  const int64_t ngroups = (n + 3) >> 2;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = g << 2;
    if (vec_ok && i0 + 4 <= n) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(rgb + 3 * i0);
      uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
      uint32_t out = 0;
      if ((w0 | w1 | w2) != 0 || black) {
        uint32_t c0 = w0 & 0xffffffu;
        uint32_t c1 = (w0 >> 24) | ((w1 & 0xffffu) << 8);
        uint32_t c2 = (w1 >> 16) | ((w2 & 0xffu) << 16);
        uint32_t c3 = w2 >> 8;
        out = ((c0 | black) ? match_palette(c0, pal, n_colors) : 0u) |
              (((c1 | black) ? match_palette(c1, pal, n_colors) : 0u) << 8) |
              (((c2 | black) ? match_palette(c2, pal, n_colors) : 0u) << 16) |
              (((c3 | black) ? match_palette(c3, pal, n_colors) : 0u) << 24);
      }
      *reinterpret_cast<uint32_t*>(labels + i0) = out;
    } else {
      for (int64_t i = i0; i < n && i < i0 + 4; ++i) {
        uint32_t c = rgb[3 * i] | (rgb[3 * i + 1] << 8) | (rgb[3 * i + 2] << 16);
        labels[i] = (uint8_t)match_palette(c, pal, n_colors);
      }
    }
  }
}

__global__ void __launch_bounds__(256) labels_to_rgb_kernel(const uint8_t* __restrict__ labels, int64_t n,
                                                            const uint8_t* __restrict__ lut,
                                                            uint8_t* __restrict__ rgb, int vec_ok) {
  __shared__ uint32_t s_lut[256];
  for (int k = threadIdx.x; k < 256; k += blockDim.x)
    s_lut[k] = lut[3 * k] | (lut[3 * k + 1] << 8) | (lut[3 * k + 2] << 16);
  __syncthreads();
  const int64_t ngroups = (n + 3) >> 2;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = g << 2;
    if (vec_ok && i0 + 4 <= n) {
      uint32_t l = __ldg(reinterpret_cast<const uint32_t*>(labels + i0));
      uint32_t c0 = s_lut[l & 0xff], c1 = s_lut[(l >> 8) & 0xff], c2 = s_lut[(l >> 16) & 0xff],
               c3 = s_lut[l >> 24];
      uint32_t* o = reinterpret_cast<uint32_t*>(rgb + 3 * i0);
      o[0] = c0 | (c1 << 24);
      o[1] = (c1 >> 8) | (c2 << 16);
      o[2] = (c2 >> 16) | (c3 << 8);
    } else {
      for (int64_t i = i0; i < n && i < i0 + 4; ++i) {
        uint32_t c = s_lut[labels[i]];
        rgb[3 * i] = c & 0xff; rgb[3 * i + 1] = (c >> 8) & 0xff; rgb[3 * i + 2] = (c >> 16) & 0xff;
      }
    }
  }
}

inline int stream_grid(int64_t work_items, int threads, int waves = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)p3d::sm_count() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

P3D_API int p3d_rgb_to_labels(const uint8_t* rgb, int64_t n, const uint8_t* palette_rgb, int n_colors,
                              uint8_t* labels, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && n_colors >= 0 && n_colors <= kMaxPalette, "rgb_to_labels: n=%lld n_colors=%d",
              (long long)n, n_colors);
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(rgb && labels && (palette_rgb || n_colors == 0), "rgb_to_labels: null pointer");
  int vec_ok = ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(labels)) & 3) == 0;
  rgb_to_labels_kernel<<<stream_grid((n + 3) / 4, 256), 256, 0, p3d::as_stream(stream)>>>(
      rgb, n, palette_rgb, n_colors, labels, vec_ok);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_labels_to_rgb(const uint8_t* labels, int64_t n, const uint8_t* lut_rgb, uint8_t* rgb,
                              p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0, "labels_to_rgb: n=%lld", (long long)n);
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(labels && lut_rgb && rgb, "labels_to_rgb: null pointer");
  int vec_ok = ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(labels)) & 3) == 0;
  labels_to_rgb_kernel<<<stream_grid((n + 3) / 4, 256), 256, 0, p3d::as_stream(stream)>>>(
      labels, n, lut_rgb, rgb, vec_ok);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// ---------------------------------------------------------------------------------------------
// Ordered compaction of non-zero labels.  Tile = 256 threads x 16 labels.  Pass 1 counts per tile,
// pass 2 scans the tile counts (one CTA), pass 3 rewrites each tile in index order.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kTileThreads = 256;
constexpr int kPerThread = 16;
constexpr int kTile = kTileThreads * kPerThread;

__device__ __forceinline__ uint4 load16(const uint8_t* __restrict__ labels, int64_t i0, int64_t n, bool vec_ok) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (i0 >= n) return v;
  if (vec_ok && i0 + 16 <= n) return __ldg(reinterpret_cast<const uint4*>(labels + i0));
  uint32_t w[4] = {0, 0, 0, 0};
  for (int j = 0; j < 16 && i0 + j < n; ++j) w[j >> 2] |= (uint32_t)labels[i0 + j] << (8 * (j & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// number of non-zero bytes in a 32-bit word
__device__ __forceinline__ int nz_bytes(uint32_t w) {
  uint32_t m = (w | (w >> 1) | (w >> 2) | (w >> 3) | (w >> 4) | (w >> 5) | (w >> 6) | (w >> 7)) & 0x01010101u;
  return __popc(m);
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[kTileThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int base = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kTileThreads / 32; ++w) {
    int s = warp_sums[w];
    if (w < warp) base += s;
    all += s;
  }
  *total = all;
  return base + incl - v;
}

__global__ void __launch_bounds__(kTileThreads) points_count_kernel(const uint8_t* __restrict__ labels, int64_t n,
                                                                    int64_t* __restrict__ tile_counts, int vec_ok) {
  const int64_t i0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kPerThread;
  uint4 v = load16(labels, i0, n, vec_ok);
  int c = nz_bytes(v.x) + nz_bytes(v.y) + nz_bytes(v.z) + nz_bytes(v.w);
  int total;
  block_exclusive_scan(c, &total);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// In-place exclusive scan of tile_counts[0..m) with the grand total written to tile_counts[m] and n_out.
__global__ void __launch_bounds__(1024) points_scan_kernel(int64_t* __restrict__ tile_counts, int64_t m,
                                                           int64_t* __restrict__ n_out) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < m; base += 1024) {
    int64_t i = base + threadIdx.x;
    int64_t v = i < m ? tile_counts[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int64_t wbase = 0, all = 0;
    for (int w = 0; w < 32; ++w) {
      int64_t s = warp_sums[w];
      if (w < warp) wbase += s;
      all += s;
    }
    const int64_t c = carry;
    if (i < m) tile_counts[i] = c + wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + all;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tile_counts[m] = carry;
    n_out[0] = carry;
  }
}

__global__ void __launch_bounds__(kTileThreads) points_fill_kernel(const uint8_t* __restrict__ labels, int64_t n,
                                                                   int A1, int A2,
                                                                   const int64_t* __restrict__ tile_offsets,
                                                                   float* __restrict__ pts,
                                                                   uint8_t* __restrict__ pt_label,
                                                                   int64_t capacity, int vec_ok) {
  const int64_t i0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kPerThread;
  uint4 v = load16(labels, i0, n, vec_ok);
  int c = nz_bytes(v.x) + nz_bytes(v.y) + nz_bytes(v.z) + nz_bytes(v.w);
  int total;
  int excl = block_exclusive_scan(c, &total);
  if (c == 0) return;
  int64_t out = tile_offsets[blockIdx.x] + excl;
  // first voxel of this thread in (a0, a1, a2); then walk a2 with carries
  const int64_t plane = (int64_t)A1 * A2;
  int a0 = (int)(i0 / plane);
  int64_t rem = i0 - (int64_t)a0 * plane;
  int a1 = (int)(rem / A2);
  int a2 = (int)(rem - (int64_t)a1 * A2);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    uint32_t lab = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
    if (lab && out < capacity) {
      pts[3 * out + 0] = (float)a2;
      pts[3 * out + 1] = (float)a1;
      pts[3 * out + 2] = (float)a0;
      pt_label[out] = (uint8_t)lab;
      ++out;
    }
    if (++a2 == A2) { a2 = 0; if (++a1 == A1) { a1 = 0; ++a0; } }
  }
}

}  // namespace

P3D_API size_t p3d_points_workspace_bytes(int64_t n_voxels) {
  if (n_voxels < 0) return 0;
  int64_t tiles = (n_voxels + kTile - 1) / kTile;
  return (size_t)(tiles + 1) * sizeof(int64_t);
}

P3D_API int p3d_points_count(const uint8_t* labels, int64_t n_voxels, int64_t* n_out, void* workspace,
                             size_t workspace_bytes, p3d_stream_t stream) {
  P3D_REQUIRE(n_voxels >= 0 && n_out && workspace, "points_count: bad arguments");
  P3D_REQUIRE(labels || n_voxels == 0, "points_count: null labels");
  if (workspace_bytes < p3d_points_workspace_bytes(n_voxels)) {
    p3d::set_error("points_count: workspace %zu < %zu", workspace_bytes, p3d_points_workspace_bytes(n_voxels));
    return P3D_E_WORKSPACE;
  }
  int64_t tiles = (n_voxels + kTile - 1) / kTile;
  P3D_REQUIRE(tiles < (1ll << 31), "points_count: grid too large");
  int64_t* tc = static_cast<int64_t*>(workspace);
  int vec_ok = (reinterpret_cast<uintptr_t>(labels) & 15) == 0;
  cudaStream_t st = p3d::as_stream(stream);
  if (tiles > 0) {
    points_count_kernel<<<(unsigned)tiles, kTileThreads, 0, st>>>(labels, n_voxels, tc, vec_ok);
    P3D_LAUNCH_CHECK();
  }
  points_scan_kernel<<<1, 1024, 0, st>>>(tc, tiles, n_out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_points_fill(const uint8_t* labels, int A0, int A1, int A2, const void* workspace, float* pts,
                            uint8_t* pt_label, int64_t capacity, p3d_stream_t stream) {
  P3D_REQUIRE(A0 >= 0 && A1 >= 0 && A2 >= 0 && capacity >= 0 && workspace, "points_fill: bad arguments");
  int64_t n = (int64_t)A0 * A1 * A2;
  if (n == 0 || capacity == 0) return P3D_OK;
  P3D_REQUIRE(labels && pts && pt_label, "points_fill: null pointer");
  int64_t tiles = (n + kTile - 1) / kTile;
  int vec_ok = (reinterpret_cast<uintptr_t>(labels) & 15) == 0;
  points_fill_kernel<<<(unsigned)tiles, kTileThreads, 0, p3d::as_stream(stream)>>>(
      labels, n, A1, A2, static_cast<const int64_t*>(workspace), pts, pt_label, capacity, vec_ok);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// ---------------------------------------------------------------------------------------------
// x-run segments of a point list (input of the segment splat, p3d_camera.cu).  get_voxel_points_by_parts
// (utils/voxel_utils.py:17-19) emits points in ascending flat index, so the voxels of one (z, y) row are consecutive in
// the list and in x.  A chunk is a run of <= 32 L list-consecutive points with the same label, the same (y, z),
// x increasing by exactly 1, that does not cross a multiple of 32 L in x; point i starts a chunk iff any of these breaks
// against point i-1 (a purely local test on the list, so any point list works; a list without such runs degenerates to
// one segment per point).  A chunk of n points is dealt out column-wise to T = ceil(n / L) segments (threads of the
// splat): segment r = 0..T-1 owns the points r, r + T, r + 2T, ... (<= L of them), so that neighbouring threads hold
// neighbouring voxels (their z-buffer accesses coalesce) while every thread still walks an arithmetic progression in x.
// Record = uint4 { x_first | y << 16, z | (count-1) << 16 | (T-1) << 20 | label << 26, index of the first point, 0 }.
// Points that are not integer-valued in [0, 65535]^3 or whose label is outside 1..32 are counted in n_out[1]; the
// caller must then not use the segment path.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kSegTile = 1024;

__device__ __forceinline__ bool seg_point_ok(const float* __restrict__ p, uint32_t lab) {
  bool ok = lab >= 1u && lab <= 32u;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float v = p[k];
    ok = ok && v >= 0.f && v <= 65535.f && v == truncf(v);
  }
  return ok;
}

__device__ __forceinline__ bool seg_is_start(const float* __restrict__ pts, const uint8_t* __restrict__ lab, int64_t i,
                                             int chunk) {
  if (i == 0) return true;
  const float* a = pts + 3 * (i - 1);
  const float* b = pts + 3 * i;
  const float bx = __ldg(b);
  return __ldg(lab + i) != __ldg(lab + i - 1) || __ldg(b + 1) != __ldg(a + 1) || __ldg(b + 2) != __ldg(a + 2) ||
         bx != __ldg(a) + 1.f || ((int)bx % chunk) == 0;
}

// Points in the chunk that starts at point i = tile_base + t.  s_start holds the tile's start flags as a bitmask (bit t
// = point tile_base + t starts a chunk), so the length is the distance to the next set bit; only a chunk that runs past
// the tile's end looks at the following points one by one.
__device__ __forceinline__ int seg_chunk_points(const float* __restrict__ pts, const uint8_t* __restrict__ lab, int64_t i,
                                                int64_t n, int chunk, const uint32_t* s_start, int t) {
  int len = 1, pos = t + 1;
  while (pos < kSegTile && len < chunk) {
    const uint32_t w = s_start[pos >> 5] >> (pos & 31);      // flags of pos .. end of its word
    if (w) {
      len += __ffs(w) - 1;
      return len < chunk ? len : chunk;
    }
    const int step = 32 - (pos & 31);
    len += step;
    pos += step;
  }
  if (len >= chunk) return chunk;
  while (len < chunk && i + len < n && !seg_is_start(pts, lab, i + len, chunk)) ++len;   // beyond the tile
  return len;
}

__device__ __forceinline__ int seg_block_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int base = 0, all = 0;
  for (int w = 0; w < kSegTile / 32; ++w) {
    const int sv = warp_sums[w];
    if (w < warp) base += sv;
    all += sv;
  }
  *total = all;
  return base + incl - v;
}

__global__ void __launch_bounds__(kSegTile) segments_count_kernel(const float* __restrict__ pts,
                                                                  const uint8_t* __restrict__ lab, int64_t n, int L,
                                                                  int64_t* __restrict__ tile_counts,
                                                                  unsigned long long* __restrict__ bad) {
  __shared__ int warp_sums[kSegTile / 32];
  __shared__ uint32_t s_start[kSegTile / 32];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  const int64_t i = (int64_t)blockIdx.x * kSegTile + threadIdx.x;
  const bool start = i < n && seg_is_start(pts, lab, i, 32 * L);
  const uint32_t ms = __ballot_sync(0xffffffffu, start || i >= n);          // the end of the list ends a chunk, too
  if ((threadIdx.x & 31) == 0) s_start[threadIdx.x >> 5] = ms;
  __syncthreads();
  int mine = 0;
  bool isbad = false;
  if (i < n) {
    if (start) mine = (seg_chunk_points(pts, lab, i, n, 32 * L, s_start, threadIdx.x) + L - 1) / L;
    isbad = !seg_point_ok(pts + 3 * i, lab[i]);
  }
  int total;
  seg_block_scan(mine, warp_sums, &total);                   // contains the barrier that publishes s_bad = 0
  const uint32_t mb = __ballot_sync(0xffffffffu, isbad);
  if ((threadIdx.x & 31) == 0 && mb) atomicAdd(&s_bad, __popc(mb));
  __syncthreads();
  if (threadIdx.x == 0) {
    tile_counts[blockIdx.x] = total;
    if (s_bad) atomicAdd(bad, (unsigned long long)s_bad);
  }
}

__global__ void __launch_bounds__(kSegTile) segments_fill_kernel(const float* __restrict__ pts,
                                                                 const uint8_t* __restrict__ lab, int64_t n, int L,
                                                                 const int64_t* __restrict__ tile_offsets,
                                                                 uint4* __restrict__ segs, int64_t capacity) {
  __shared__ int warp_sums[kSegTile / 32];
  __shared__ uint32_t s_start[kSegTile / 32];
  const int64_t i = (int64_t)blockIdx.x * kSegTile + threadIdx.x;
  const bool start = i < n && seg_is_start(pts, lab, i, 32 * L);
  const uint32_t ms = __ballot_sync(0xffffffffu, start || i >= n);
  if ((threadIdx.x & 31) == 0) s_start[threadIdx.x >> 5] = ms;
  __syncthreads();
  int len = 0, T = 0;
  if (start) {
    len = seg_chunk_points(pts, lab, i, n, 32 * L, s_start, threadIdx.x);
    T = (len + L - 1) / L;
  }
  int total;
  const int before = seg_block_scan(T, warp_sums, &total);
  if (T == 0) return;
  const int64_t out = tile_offsets[blockIdx.x] + before;
  const float* p = pts + 3 * i;
  const uint32_t x0 = (uint32_t)(int)p[0], y = (uint32_t)(int)p[1] & 0xffffu, z = (uint32_t)(int)p[2] & 0xffffu;
  const uint32_t label = lab[i];
  for (int r = 0; r < T; ++r) {
    if (out + r >= capacity) break;
    const int cnt = (len - r + T - 1) / T;                   // points r, r + T, ... < len
    segs[out + r] = make_uint4(((x0 + (uint32_t)r) & 0xffffu) | (y << 16),
                               z | ((uint32_t)(cnt - 1) << 16) | ((uint32_t)(T - 1) << 20) | (label << 26),
                               (uint32_t)(i + r), 0u);
  }
}

}  // namespace

P3D_API size_t p3d_segments_workspace_bytes(int64_t n_points) {
  if (n_points < 0) return 0;
  const int64_t tiles = (n_points + kSegTile - 1) / kSegTile;
  return (size_t)(tiles + 1) * sizeof(int64_t);
}

P3D_API int p3d_segments_count(const float* pts, const uint8_t* pt_label, int64_t n, int seg_len, int64_t* n_out,
                               void* workspace, size_t workspace_bytes, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && n < 0xffffffffll && n_out && workspace, "segments_count: bad arguments");
  P3D_REQUIRE(seg_len >= 1 && seg_len <= 16, "segments_count: seg_len=%d", seg_len);
  P3D_REQUIRE((pts && pt_label) || n == 0, "segments_count: null points");
  if (workspace_bytes < p3d_segments_workspace_bytes(n)) {
    p3d::set_error("segments_count: workspace %zu < %zu", workspace_bytes, p3d_segments_workspace_bytes(n));
    return P3D_E_WORKSPACE;
  }
  const int64_t tiles = (n + kSegTile - 1) / kSegTile;
  int64_t* tc = static_cast<int64_t*>(workspace);
  cudaStream_t st = p3d::as_stream(stream);
  P3D_CUDA(cudaMemsetAsync(n_out, 0, 2 * sizeof(int64_t), st));
  if (tiles > 0) {
    segments_count_kernel<<<(unsigned)tiles, kSegTile, 0, st>>>(pts, pt_label, n, seg_len, tc,
                                                                reinterpret_cast<unsigned long long*>(n_out + 1));
    P3D_LAUNCH_CHECK();
  }
  points_scan_kernel<<<1, 1024, 0, st>>>(tc, tiles, n_out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_segments_fill(const float* pts, const uint8_t* pt_label, int64_t n, int seg_len, const void* workspace,
                              uint32_t* segs, int64_t capacity, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && capacity >= 0 && workspace && seg_len >= 1 && seg_len <= 16, "segments_fill: bad arguments");
  if (n == 0 || capacity == 0) return P3D_OK;
  P3D_REQUIRE(pts && pt_label && segs, "segments_fill: null pointer");
  P3D_REQUIRE((reinterpret_cast<uintptr_t>(segs) & 15) == 0, "segments_fill: segs must be 16-byte aligned");
  const int64_t tiles = (n + kSegTile - 1) / kSegTile;
  segments_fill_kernel<<<(unsigned)tiles, kSegTile, 0, p3d::as_stream(stream)>>>(
      pts, pt_label, n, seg_len, static_cast<const int64_t*>(workspace), reinterpret_cast<uint4*>(segs), capacity);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// ---------------------------------------------------------------------------------------------
// voxel_grid_to_points (utils/voxel_utils.py:35-51): strided occupancy of an RGB grid and the colours of the kept
// voxels.  mask[(a,b,c)] = any(grid[a*s, b*s, c*s, :]) on the sub-sampled lattice; colours are gathered at the
// compacted points afterwards.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) strided_occupancy_kernel(const uint8_t* __restrict__ grid, int A1, int A2, int stride,
                                                                int B0, int B1, int B2, uint8_t* __restrict__ mask) {
  const int64_t n = (int64_t)B0 * B1 * B2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % B2);
    const int64_t r = i / B2;
    const int b = (int)(r % B1), a = (int)(r / B1);
    const uint8_t* p = grid + ((((size_t)a * stride) * A1 + (size_t)b * stride) * A2 + (size_t)c * stride) * 3;
    mask[i] = (p[0] | p[1] | p[2]) != 0;
  }
}

// rgb[i] = grid[z*s][y*s][x*s] for pts[i] = (x, y, z) on the sub-sampled lattice; pts are scaled by s in place
__global__ void __launch_bounds__(256) gather_scale_points_kernel(const uint8_t* __restrict__ grid, int A1, int A2, int stride,
                                                                  float* __restrict__ pts, int64_t n,
                                                                  uint8_t* __restrict__ rgb) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)pts[3 * i], y = (int)pts[3 * i + 1], z = (int)pts[3 * i + 2];
    const uint8_t* p = grid + ((((size_t)z * stride) * A1 + (size_t)y * stride) * A2 + (size_t)x * stride) * 3;
    rgb[3 * i] = p[0]; rgb[3 * i + 1] = p[1]; rgb[3 * i + 2] = p[2];
    pts[3 * i] = __fmul_rn((float)x, (float)stride);
    pts[3 * i + 1] = __fmul_rn((float)y, (float)stride);
    pts[3 * i + 2] = __fmul_rn((float)z, (float)stride);
  }
}

}  // namespace

P3D_API int p3d_strided_occupancy(const uint8_t* grid_rgb, int A0, int A1, int A2, int stride, uint8_t* mask,
                                  p3d_stream_t stream) {
  P3D_REQUIRE(A0 >= 0 && A1 >= 0 && A2 >= 0 && stride >= 1, "strided_occupancy: bad arguments");
  const int B0 = (A0 + stride - 1) / stride, B1 = (A1 + stride - 1) / stride, B2 = (A2 + stride - 1) / stride;
  const int64_t n = (int64_t)B0 * B1 * B2;
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(grid_rgb && mask, "strided_occupancy: null pointer");
  strided_occupancy_kernel<<<stream_grid(n, 256), 256, 0, p3d::as_stream(stream)>>>(grid_rgb, A1, A2, stride, B0, B1, B2, mask);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_gather_scale_points(const uint8_t* grid_rgb, int A0, int A1, int A2, int stride, float* pts, int64_t n,
                                    uint8_t* rgb, p3d_stream_t stream) {
  P3D_REQUIRE(A0 >= 0 && A1 >= 0 && A2 >= 0 && stride >= 1 && n >= 0, "gather_scale_points: bad arguments");
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(grid_rgb && pts && rgb, "gather_scale_points: null pointer");
  gather_scale_points_kernel<<<stream_grid(n, 256), 256, 0, p3d::as_stream(stream)>>>(grid_rgb, A1, A2, stride, pts, n, rgb);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// ---------------------------------------------------------------------------------------------
// compute_binary_gt (utils/eval_helpers_intra.py:274-285): pixels whose colour occurs (as a non-black voxel colour)
// in the grid.  Pass 1 marks the 24-bit colours present in the grid in a 2 MiB bitmap, pass 2 looks every pixel up.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) colour_presence_kernel(const uint8_t* __restrict__ rgb, int64_t n,
                                                              uint32_t* __restrict__ present) {
  uint32_t last = 0;                                       // runs of one colour set their bit once
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t c = rgb[3 * i] | (rgb[3 * i + 1] << 8) | (rgb[3 * i + 2] << 16);
    if (c == 0 || c == last) continue;
    last = c;
    const uint32_t bit = 1u << (c & 31u);
    if ((present[c >> 5] & bit) == 0u) atomicOr(present + (c >> 5), bit);
  }
}

__global__ void __launch_bounds__(256) colour_lookup_kernel(const uint8_t* __restrict__ rgb, int64_t n,
                                                            const uint32_t* __restrict__ present,
                                                            uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t c = rgb[3 * i] | (rgb[3 * i + 1] << 8) | (rgb[3 * i + 2] << 16);
    mask[i] = c != 0 && ((present[c >> 5] >> (c & 31u)) & 1u);
  }
}

}  // namespace

P3D_API size_t p3d_colour_presence_bytes(void) { return (size_t)1 << 21; }

P3D_API int p3d_colour_presence(const uint8_t* grid_rgb, int64_t n, uint32_t* present, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && present, "colour_presence: bad arguments");
  cudaStream_t st = p3d::as_stream(stream);
  P3D_CUDA(cudaMemsetAsync(present, 0, p3d_colour_presence_bytes(), st));
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(grid_rgb, "colour_presence: null grid");
  colour_presence_kernel<<<stream_grid(n, 256), 256, 0, st>>>(grid_rgb, n, present);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_colour_lookup(const uint8_t* image_rgb, int64_t n, const uint32_t* present, uint8_t* mask,
                              p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0, "colour_lookup: bad arguments");
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(image_rgb && present && mask, "colour_lookup: null pointer");
  colour_lookup_kernel<<<stream_grid(n, 256), 256, 0, p3d::as_stream(stream)>>>(image_rgb, n, present, mask);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}
