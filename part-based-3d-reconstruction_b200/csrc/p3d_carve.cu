// libp3d_b200: orthographic semantic voxel carving (stage 1 of the reference).
//
// Reference call sites replaced here (paths under the reference root, utils/voxel_carving_utils.py):
//   process_voxel_grid :104-126  (scipy.ndimage.affine_transform order=1 + carve_voxel_grid_with_masks :76-87)
//   apply_colored_mask_to_voxel_grid :128-136,  global_carve :269-298
//   part_carve :139-160,  left_right_guided_carve :163-210 (crop / paste around scipy.ndimage.label)
//   extrude_from_surface :213-248,  recolor_backward_components :252-266,  partwise_carve :302-400
//
// Grids are (W,H,D[,3]) uint8, C order.  2-D masks arrive already oriented as (W,H) ("mask_wh", the result
// of the reference's _mask_to_wh) and, where a kernel gathers along x, additionally as (H,W) ("mask_hw").
//
// Exactness: the resample follows scipy's NI_GeometricTransform for order=1, mode="constant" --
//   cc[h] = off[h]; cc[h] += M[h][0]*o0; += M[h][1]*o1; += M[h][2]*o2   (FP64, separately rounded)
//   outside if cc < 0 or cc > dim-1; fl = floor(cc); t = cc - fl; i1 = min(i0+1, dim-1)
//   acc = sum over the 8 corners (axis0, axis1, axis2 nesting) of v * w0 * w1 * w2 ; out = (uint8)(acc + 0.5)
// When every in-range coordinate of a pass is within 1e-9 of an integer (angle 0, and angle 90 on a grid with
// D == W) the blend collapses to a nearest-index gather; that case is detected at run time from the actual
// host-computed matrix/offset ("fold table") and served by streaming kernels.
#include "p3d_common.cuh"

namespace {

struct Affine {
  double M[9];
  double off[3];
};

// n / d as (n * magic) >> 40 with magic = ceil(2^40 / d): exact for n <= n_max while n_max * d < 2^40; 0 = the kernel
// divides (div_magic)
inline unsigned long long magic_for(uint64_t n_max, uint64_t d) {
  return n_max * d < (1ull << 40) ? ((1ull << 40) + d - 1) / d : 0ull;
}

inline int grid_for(int64_t items, int threads, int waves) {
  int64_t blocks = (items + threads - 1) / threads;
  int64_t cap = (int64_t)p3d::sm_count() * waves;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

__device__ __forceinline__ bool rgb_nonzero(const uint8_t* p) { return (p[0] | p[1] | p[2]) != 0; }

// ------------------------------------------------------------------------------------------
// General trilinear resample + mask carve (one thread per output voxel).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double coord(const Affine& A, int h, int o0, int o1, int o2) {
  double c = A.off[h];
  c = __dadd_rn(c, __dmul_rn(A.M[3 * h + 0], (double)o0));
  c = __dadd_rn(c, __dmul_rn(A.M[3 * h + 1], (double)o1));
  c = __dadd_rn(c, __dmul_rn(A.M[3 * h + 2], (double)o2));
  return c;
}

// scipy.ndimage.affine_transform(order=1, mode="constant", cval=0) at output voxel (o0, o1, o2) of an (n0, n1, n2) uint8
// volume: FP64 coordinates accumulated in scipy's order, bounds rule, 8 corners, round half up.
__device__ __forceinline__ uint8_t resample_voxel(const uint8_t* __restrict__ vin, int n0, int n1, int n2, const Affine& A,
                                                  int o0, int o1, int o2) {
  uint8_t res = 0;
  const int dim[3] = {n0, n1, n2};
  double cc[3];
  bool inside = true;
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    cc[h] = coord(A, h, o0, o1, o2);
    if (cc[h] < 0.0 || cc[h] > (double)(dim[h] - 1)) inside = false;
  }
  if (inside) {
    int i0[3], i1[3];
    double t[3];
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      const double fl = floor(cc[h]);
      t[h] = __dsub_rn(cc[h], fl);
      i0[h] = (int)fl;
      i1[h] = i0[h] + 1 < dim[h] ? i0[h] + 1 : dim[h] - 1;
    }
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int ix = a ? i1[0] : i0[0], iy = b ? i1[1] : i0[1], iz = c ? i1[2] : i0[2];
          double w = (double)vin[((size_t)ix * n1 + iy) * n2 + iz];
          w = __dmul_rn(w, a ? t[0] : __dsub_rn(1.0, t[0]));
          w = __dmul_rn(w, b ? t[1] : __dsub_rn(1.0, t[1]));
          w = __dmul_rn(w, c ? t[2] : __dsub_rn(1.0, t[2]));
          acc = __dadd_rn(acc, w);
        }
    if (acc > 0.0) {
      const double rr = __dadd_rn(acc, 0.5);
      res = rr >= 255.0 ? 255 : (uint8_t)rr;
    }
  }
  return res;
}

// ------------------------------------------------------------------------------------------
// left_right_guided_carve (:163-210) for ALL components of one colour in a fixed number of launches: blockIdx.y = the
// component.  comps[c] = { x0, y0, z0, w, h, d, scratch offset low, high } (bounding-box crop and where it lives in the
// two scratch buffers); the crop's 2-D mask is read in place, m[x][y] = mask_hw[y0 + y][x0 + x] (the reference's
// _mask_to_wh of the (h, w) crop is always its transpose).  One pass = one launch over every crop: Ms[pass] is the
// pass's inverse rotation, offs[c][pass] the offset for crop c's shape.
// ------------------------------------------------------------------------------------------
struct CompBox { int x0, y0, z0, w, h, d; long long off; };
__device__ __forceinline__ CompBox load_comp(const int32_t* __restrict__ comps, int c) {
  const int32_t* p = comps + (size_t)c * 8;
  CompBox b;
  b.x0 = p[0]; b.y0 = p[1]; b.z0 = p[2]; b.w = p[3]; b.h = p[4]; b.d = p[5];
  b.off = (long long)(uint32_t)p[6] | ((long long)p[7] << 32);
  return b;
}

__global__ void __launch_bounds__(256)
lr_crop_kernel(const uint8_t* __restrict__ grid, int H, int D, const int32_t* __restrict__ comps, uint8_t* __restrict__ buf) {
  const CompBox b = load_comp(comps, blockIdx.y);
  const int64_t n = (int64_t)b.w * b.h * b.d;
  uint8_t* occ = buf + b.off;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % b.d);
    const int64_t q = i / b.d;
    const int y = (int)(q % b.h), x = (int)(q / b.h);
    occ[i] = rgb_nonzero(grid + 3 * ((((size_t)(b.x0 + x)) * H + (b.y0 + y)) * D + (b.z0 + z))) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256)
lr_resample_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int32_t* __restrict__ comps,
                   const double* __restrict__ M, const double* __restrict__ offs, int pass, int n_pass,
                   const uint8_t* __restrict__ mask_hw, int Wfull) {
  const CompBox b = load_comp(comps, blockIdx.y);
  Affine A;
#pragma unroll
  for (int k = 0; k < 9; ++k) A.M[k] = M[(size_t)pass * 9 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) A.off[k] = offs[((size_t)blockIdx.y * n_pass + pass) * 3 + k];
  const int64_t n = (int64_t)b.w * b.h * b.d;
  const uint8_t* vin = src + b.off;
  uint8_t* vout = dst + b.off;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int o2 = (int)(i % b.d);
    const int64_t q = i / b.d;
    const int o1 = (int)(q % b.h), o0 = (int)(q / b.h);
    uint8_t res = 0;
    if (mask_hw[(size_t)(b.y0 + o1) * Wfull + (b.x0 + o0)]) res = resample_voxel(vin, b.w, b.h, b.d, A, o0, o1, o2);
    vout[i] = res;
  }
}

// paste of :199-201 over each crop (src = the call's INPUT grid) + the component's "carved voxels" count.  With
// comp_first >= 0 only that component is pasted (sequential launches when bounding boxes overlap: the reference's
// loop order decides there).
__global__ void __launch_bounds__(256)
lr_paste_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ labels, const uint8_t* __restrict__ buf,
                const int32_t* __restrict__ comps, int comp_first, int H, int D, uint8_t* __restrict__ out,
                unsigned long long* __restrict__ counts) {
  const int c = comp_first >= 0 ? comp_first : (int)blockIdx.y;
  const CompBox b = load_comp(comps, c);
  const int64_t n = (int64_t)b.w * b.h * b.d;
  const uint8_t* kept = buf + b.off;
  unsigned int mine = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % b.d);
    const int64_t q = i / b.d;
    const int y = (int)(q % b.h), x = (int)(q / b.h);
    const size_t v = (((size_t)(b.x0 + x)) * H + (b.y0 + y)) * D + (b.z0 + z);
    const uint8_t* p = src + 3 * v;
    const bool k = kept[i] != 0;
    mine += k;
    if (k && rgb_nonzero(p)) {
      out[3 * v] = p[0]; out[3 * v + 1] = p[1]; out[3 * v + 2] = p[2];
    } else if (labels[v] == c + 1) {
      out[3 * v] = 0; out[3 * v + 1] = 0; out[3 * v + 2] = 0;
    }
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(counts + c, (unsigned long long)mine);
}

// part_carve's group image on the device: gm[y][x] bit g = pixel colour is one of group g's colours -- and, for a
// square image, also at the transposed pixel (the reference's _mask_to_wh quirk: m & m.T).  keys[k] = r | g << 8 |
// b << 16, group_of[k] = its group.
__global__ void __launch_bounds__(256)
group_image_kernel(const uint8_t* __restrict__ mask_rgb, int H, int W, const uint32_t* __restrict__ keys,
                   const int32_t* __restrict__ group_of, int n_keys, uint32_t* __restrict__ gm) {
  __shared__ uint32_t s_key[128];
  __shared__ int32_t s_grp[128];
  for (int k = threadIdx.x; k < n_keys; k += blockDim.x) { s_key[k] = keys[k]; s_grp[k] = group_of[k]; }
  __syncthreads();
  const int n = H * W;
  auto bits_at = [&](int pix) {
    const uint8_t* p = mask_rgb + 3 * (size_t)pix;
    const uint32_t c = p[0] | (p[1] << 8) | (p[2] << 16);
    uint32_t bits = 0;
    for (int k = 0; k < n_keys; ++k)
      if (s_key[k] == c) bits |= 1u << s_grp[k];
    return bits;
  };
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t bits = bits_at(i);
    if (W == H && bits) {
      const int y = i / W, x = i - y * W;
      bits &= bits_at(x * W + y);
    }
    gm[i] = bits;
  }
}

__global__ void __launch_bounds__(256)
resample_carve_kernel(const uint8_t* __restrict__ vin, int n0, int n1, int n2, Affine A,
                      const uint8_t* __restrict__ mask_wh, uint8_t* __restrict__ vout) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int o2 = (int)(i % n2);
    const int64_t r = i / n2;
    const int o1 = (int)(r % n1);
    const int o0 = (int)(r / n1);
    uint8_t res = 0;
    if (mask_wh == nullptr || mask_wh[(size_t)o0 * n1 + o1]) res = resample_voxel(vin, n0, n1, n2, A, o0, o1, o2);
    vout[i] = res;
  }
}

// ------------------------------------------------------------------------------------------
// Fold table: for a pass whose y axis is decoupled (M row/col 1 = identity, off[1] = 0), the source index of
// output (o0, o2) as (src0 << 16 | src2), or -1 when scipy treats the point as outside.  flag[0] is set
// when some in-range coordinate is NOT within 1e-9 of an integer (then the pass needs the general kernel).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fold_table_kernel(int n0, int n2, Affine A, int32_t* __restrict__ table, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0 * n2) return;
  const int o0 = i / n2, o2 = i - o0 * n2;
  const double c0 = coord(A, 0, o0, 0, o2), c2 = coord(A, 2, o0, 0, o2);
  int32_t e = -1;
  if (!(c0 < 0.0 || c0 > (double)(n0 - 1) || c2 < 0.0 || c2 > (double)(n2 - 1))) {
    const double f0 = floor(c0), f2 = floor(c2);
    const double t0 = __dsub_rn(c0, f0), t2 = __dsub_rn(c2, f2);
    const bool near0 = t0 < 1e-9 || t0 > 1.0 - 1e-9, near2 = t2 < 1e-9 || t2 > 1.0 - 1e-9;
    if (!(near0 && near2)) atomicOr(flag, 1);
    int s0 = (int)f0, s2 = (int)f2;
    if (t0 > 0.5) s0 = s0 + 1 < n0 ? s0 + 1 : n0 - 1;
    if (t2 > 0.5) s2 = s2 + 1 < n2 ? s2 + 1 : n2 - 1;
    e = (s0 << 16) | s2;
  }
  table[i] = e;
}

// Gather pass for a collapsible transform: out[x,y,z] = mask[x,y] ? in[src0(x,z), y, src2(x,z)] : 0.
__global__ void __launch_bounds__(256)
fold_gather_kernel(const uint8_t* __restrict__ vin, int n0, int n1, int n2, const int32_t* __restrict__ table,
                   const uint8_t* __restrict__ mask_wh, uint8_t* __restrict__ vout) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int o2 = (int)(i % n2);
    const int64_t r = i / n2;
    const int o1 = (int)(r % n1);
    const int o0 = (int)(r / n1);
    uint8_t res = 0;
    if (mask_wh == nullptr || mask_wh[(size_t)o0 * n1 + o1]) {
      const int32_t e = table[o0 * n2 + o2];
      if (e >= 0) res = vin[((size_t)(e >> 16) * n1 + o1) * n2 + (e & 0xffff)];
    }
    vout[i] = res;
  }
}

// ------------------------------------------------------------------------------------------
// global_carve, 90-degree closed form, fused with the colouring:
//   out[x,y,z,:] = colour[y,x,:]  if  m[x,y] && table[x,z] >= 0 && m[src0(x,z), y]   else 0
// (the angle-0 pass is the identity; carved = rot90(ones & m) & m).
// Vector path (D % 16 == 0): a thread owns 16 voxels of one z-row = 48 output bytes, staged through shared
// memory so that every store instruction writes 512 contiguous bytes per warp.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void expand4(uint32_t bits, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t* o) {
  const uint32_t b0 = bits & 1u ? 0xffffffffu : 0u, b1 = bits & 2u ? 0xffffffffu : 0u;
  const uint32_t b2 = bits & 4u ? 0xffffffffu : 0u, b3 = bits & 8u ? 0xffffffffu : 0u;
  o[0] = w0 & ((b0 & 0x00ffffffu) | (b1 & 0xff000000u));
  o[1] = w1 & ((b1 & 0x0000ffffu) | (b2 & 0xffff0000u));
  o[2] = w2 & ((b2 & 0x000000ffu) | (b3 & 0xffffff00u));
}

template <bool RGB>
__global__ void __launch_bounds__(256)
global_fold_kernel(int W, int H, int D, const int32_t* __restrict__ table, const uint8_t* __restrict__ mask_hw,
                   const uint8_t* __restrict__ colour_hw /* (H,W,3) if RGB else (H,W) labels */,
                   uint8_t* __restrict__ out) {
  __shared__ uint4 stage[8][96];                       // 8 warps x 1536 bytes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t groups = (int64_t)W * H * D / 16;
  const int64_t warp_groups = (groups + 31) / 32;
  for (int64_t wg = (int64_t)blockIdx.x * 8 + warp; wg < warp_groups; wg += (int64_t)gridDim.x * 8) {
    const int64_t g = wg * 32 + lane;
    uint32_t bits = 0;
    uint32_t col = 0;
    if (g < groups) {
      const int64_t v0 = g * 16;
      const int z0 = (int)(v0 % D);
      const int64_t r = v0 / D;
      const int y = (int)(r % H), x = (int)(r / H);
      if (mask_hw[(size_t)y * W + x]) {
        const int32_t* trow = table + (size_t)x * D + z0;
        const uint8_t* mrow = mask_hw + (size_t)y * W;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int32_t e = __ldg(trow + j);
          if (e >= 0 && __ldg(mrow + (e >> 16))) bits |= 1u << j;
        }
        if (RGB) {
          const uint8_t* c = colour_hw + ((size_t)y * W + x) * 3;
          col = c[0] | (c[1] << 8) | (c[2] << 16);
        } else {
          col = colour_hw[(size_t)y * W + x];
        }
      }
    }
    if (RGB) {
      const uint32_t r8 = col & 0xff, g8 = (col >> 8) & 0xff, b8 = (col >> 16) & 0xff;
      const uint32_t w0 = r8 | (g8 << 8) | (b8 << 16) | (r8 << 24);
      const uint32_t w1 = g8 | (b8 << 8) | (r8 << 16) | (g8 << 24);
      const uint32_t w2 = b8 | (r8 << 8) | (g8 << 16) | (b8 << 24);
      uint32_t o[12];
#pragma unroll
      for (int q = 0; q < 4; ++q) expand4((bits >> (4 * q)) & 0xfu, w0, w1, w2, o + 3 * q);
      uint4* mine = &stage[warp][lane * 3];
      mine[0] = make_uint4(o[0], o[1], o[2], o[3]);
      mine[1] = make_uint4(o[4], o[5], o[6], o[7]);
      mine[2] = make_uint4(o[8], o[9], o[10], o[11]);
      __syncwarp();
      uint4* dst = reinterpret_cast<uint4*>(out) + wg * 96;
      const int64_t limit = groups * 3;                // uint4 count of the whole output
#pragma unroll
      for (int p = 0; p < 3; ++p)
        if (wg * 96 + p * 32 + lane < limit) dst[p * 32 + lane] = stage[warp][p * 32 + lane];
      __syncwarp();
    } else if (g < groups) {
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t b = (bits >> (4 * q)) & 0xfu;
        o[q] = (col * 0x01010101u) & ((b & 1u ? 0xffu : 0u) | (b & 2u ? 0xff00u : 0u) | (b & 4u ? 0xff0000u : 0u) |
                                     (b & 8u ? 0xff000000u : 0u));
      }
      reinterpret_cast<uint4*>(out)[g] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Bit-level fast path of the fused global_carve kernel.  When the fold table is "z-separable" -- every in-range
// entry satisfies src0(x,z) = c - z for one constant c (checked by fold_analyse_kernel on the actual table, true
// for the 90-degree pass) -- the 16 mask lookups of a thread are 16 consecutive bits of the bit-packed mask row,
// read in reverse, and the 16 table lookups collapse to 16 bits of an "inside" bit matrix.
//   inside_bits : (W, ceil(D/32)) uint32, bit z%32 of word [x][z/32] = table[x][z] >= 0
//   mask_bits   : (H, wpr) uint32 with one zero word of padding on each side: pixel x is bit (x+32)%32 of word
//                 [y][(x+32)/32]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fold_analyse_kernel(const int32_t* __restrict__ table, int W, int D, uint32_t* __restrict__ inside_bits,
                    int* __restrict__ info /* over in-range entries: [0] = max(src0+z), [1] = min(src0+z),
                                              [2] = max(src2-x), [3] = min(src2-x) */) {
  const int words = (D + 31) >> 5;                              // a ragged last word keeps zero bits beyond D
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // one thread per (x, word)
  if (i >= W * words) return;
  const int x = i / words, w = i - x * words;
  uint32_t bits = 0;
  int mx = -0x7fffffff, mn = 0x7fffffff, mx2 = -0x7fffffff, mn2 = 0x7fffffff;
  for (int j = 0; j < 32; ++j) {
    const int z = w * 32 + j;
    if (z >= D) break;
    const int32_t e = table[(size_t)x * D + z];
    if (e >= 0) {
      bits |= 1u << j;
      const int c = (e >> 16) + z, c2 = (e & 0xffff) - x;
      mx = max(mx, c);
      mn = min(mn, c);
      mx2 = max(mx2, c2);
      mn2 = min(mn2, c2);
    }
  }
  inside_bits[i] = bits;
  if (bits) { atomicMax(info, mx); atomicMin(info + 1, mn); atomicMax(info + 2, mx2); atomicMin(info + 3, mn2); }
}

__global__ void __launch_bounds__(256)
pack_mask_bits_kernel(const uint8_t* __restrict__ mask_hw, int H, int W, int wpr, uint32_t* __restrict__ bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // one thread per output word
  if (i >= H * wpr) return;
  const int y = i / wpr, w = i - y * wpr;
  uint32_t v = 0;
  const int x0 = (w - 1) * 32;                                  // one word of left padding
  for (int j = 0; j < 32; ++j) {
    const int x = x0 + j;
    if (x >= 0 && x < W && mask_hw[(size_t)y * W + x]) v |= 1u << j;
  }
  bits[i] = v;
}

// Indexing: warps walk the slab's thread groups in flat order (consecutive warps write consecutive 1536-byte
// chunks, which is what the DRAM pages like) in 32-bit arithmetic; row = group / (D/16) and x = row / H are each one
// multiply-high with a host-made reciprocal (ceil(2^40 / d), exact while dividend_max * d < 2^40; 0 = plain division).
// The earlier 64-bit flat index cost two emulated 64-bit divisions per thread and made this write-only kernel
// issue-bound.
__device__ __forceinline__ uint32_t div_magic(uint32_t n, uint32_t d, unsigned long long magic) {
  return magic ? (uint32_t)(((unsigned long long)n * magic) >> 40) : n / d;
}

template <bool RGB>
__global__ void __launch_bounds__(256)
global_fold_bits_kernel(int W, int H, int D, int x_begin, int x_count, const uint32_t* __restrict__ inside_bits, int c,
                        const uint32_t* __restrict__ mask_bits, int wpr, const uint8_t* __restrict__ colour_hw,
                        uint8_t* __restrict__ out, unsigned long long magic_gpr, unsigned long long magic_h) {
  __shared__ uint4 stage[8][96];
#ifdef P3D_GFB_TMA
  __shared__ __align__(128) uint4 stage2[8][2][96];
  int it = 0;
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t gpr = (uint32_t)D >> 4;                      // thread groups (16 voxels) per z-row
  const uint32_t groups = (uint32_t)x_count * (uint32_t)H * gpr;     // < 2^31 (the host splits larger slabs)
  const uint32_t warp_groups = (groups + 31u) >> 5;
  const int words = D >> 5;
  for (uint32_t wg = blockIdx.x * 8u + warp; wg < warp_groups; wg += gridDim.x * 8u) {
    const uint32_t g = wg * 32u + lane;
    uint32_t bits = 0, col = 0;
    if (g < groups) {
      const uint32_t row = div_magic(g, gpr, magic_gpr);
      const int z0 = (int)(g - row * gpr) << 4;
      const uint32_t xi = div_magic(row, (uint32_t)H, magic_h);
      const uint32_t y = row - xi * (uint32_t)H;
      const int x = x_begin + (int)xi, xb = x + 32;
      const uint32_t* mrow = mask_bits + (size_t)y * wpr;
      if ((__ldg(mrow + (xb >> 5)) >> (xb & 31)) & 1u) {
        const uint32_t in16 = (__ldg(inside_bits + (size_t)x * words + (z0 >> 5)) >> (z0 & 31)) & 0xffffu;
        // mask pixels c - z for z = z0 .. z0+15  ==  bits [lo, lo+16) of the row read backwards, lo = c - z0 - 15
        const int lo = c - z0 - 15 + 32;                      // + padding; in [0, 32*wpr - 16] when in16 != 0
        uint32_t m16 = 0;
        if (in16 && lo >= 0 && (lo >> 5) + 1 < wpr) {
          const uint32_t w0 = __ldg(mrow + (lo >> 5)), w1 = __ldg(mrow + (lo >> 5) + 1);
          m16 = __funnelshift_r(w0, w1, lo & 31) & 0xffffu;
        }
        bits = in16 & (__brev(m16) >> 16);
        if (bits) {
          if (RGB) {
            const uint8_t* cc = colour_hw + ((size_t)y * W + x) * 3;
            col = cc[0] | (cc[1] << 8) | (cc[2] << 16);
          } else {
            col = colour_hw[(size_t)y * W + x];
          }
        }
      }
    }
    if (RGB) {
      const uint32_t r8 = col & 0xff, g8 = (col >> 8) & 0xff, b8 = (col >> 16) & 0xff;
      const uint32_t w0 = r8 | (g8 << 8) | (b8 << 16) | (r8 << 24);
      const uint32_t w1 = g8 | (b8 << 8) | (r8 << 16) | (g8 << 24);
      const uint32_t w2 = b8 | (r8 << 8) | (g8 << 16) | (b8 << 24);
      uint32_t o[12];
#pragma unroll
      for (int q = 0; q < 4; ++q) expand4((bits >> (4 * q)) & 0xfu, w0, w1, w2, o + 3 * q);
#ifdef P3D_GFB_TMA
      // experiment: the warp's 1536 bytes leave shared memory as ONE bulk async copy (TMA, 1-D) issued by lane 0;
      // two staging buffers per warp, at most one copy still reading when the other buffer is rewritten
      {
        uint4* sbuf = stage2[warp][it & 1];
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        uint4* mine2 = &sbuf[lane * 3];
        mine2[0] = make_uint4(o[0], o[1], o[2], o[3]);
        mine2[1] = make_uint4(o[4], o[5], o[6], o[7]);
        mine2[2] = make_uint4(o[8], o[9], o[10], o[11]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        uint4* dst2 = reinterpret_cast<uint4*>(out) + (size_t)wg * 96;
        ++it;
        if ((wg + 1u) * 32u <= groups) {
          if (lane == 0) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sbuf);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst2), "r"(sa), "r"(1536u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          const uint32_t left2 = (groups - wg * 32u) * 3u;
#pragma unroll
          for (int p = 0; p < 3; ++p)
            if ((uint32_t)(p * 32 + lane) < left2) __stcs(dst2 + p * 32 + lane, sbuf[p * 32 + lane]);
        }
        __syncwarp();
        continue;
      }
#endif
      uint4* mine = &stage[warp][lane * 3];
      mine[0] = make_uint4(o[0], o[1], o[2], o[3]);
      mine[1] = make_uint4(o[4], o[5], o[6], o[7]);
      mine[2] = make_uint4(o[8], o[9], o[10], o[11]);
      __syncwarp();
      uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)wg * 96;
      if ((wg + 1u) * 32u <= groups) {                         // full warp: 3 x 512 contiguous bytes
#pragma unroll
        for (int p = 0; p < 3; ++p) __stcs(dst + p * 32 + lane, stage[warp][p * 32 + lane]);
      } else {
        const uint32_t left = (groups - wg * 32u) * 3u;         // uint4 still inside the slab
#pragma unroll
        for (int p = 0; p < 3; ++p)
          if ((uint32_t)(p * 32 + lane) < left) __stcs(dst + p * 32 + lane, stage[warp][p * 32 + lane]);
      }
      __syncwarp();
    } else if (g < groups) {
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t b = (bits >> (4 * q)) & 0xfu;
        o[q] = (col * 0x01010101u) & ((b & 1u ? 0xffu : 0u) | (b & 2u ? 0xff00u : 0u) | (b & 4u ? 0xff0000u : 0u) |
                                     (b & 8u ? 0xff000000u : 0u));
      }
      __stcs(reinterpret_cast<uint4*>(out) + g, make_uint4(o[0], o[1], o[2], o[3]));
    }
  }
#ifdef P3D_GFB_TMA
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}

// ------------------------------------------------------------------------------------------
// The same bit-level kernels for RAGGED rows (D % 32 != 0, e.g. the 88 / 177 / 246 wide grids of a portrait mask): a
// z-row is then neither a whole number of 16-voxel groups nor 16-byte aligned, so the groups are cut from the FLAT
// voxel order instead -- group g = voxels [16 g, 16 g + 16), 48 aligned bytes wherever the rows fall -- and a group
// spans at most two z-rows (D >= 16).  Each piece (row, z, n voxels) gets its bits exactly like an aligned group: the
// inside bits by a funnel shift of two words, the mask pixels c - z .. c - z - 15 as 16 reversed bits of the padded
// mask row; the two pieces are merged by position.  Loads, stores and the per-warp staging stay as in the aligned
// kernels; only a grid whose byte count is no multiple of 16 ends in a byte-wise tail.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t inside16_ragged(const uint32_t* __restrict__ inside_bits, int x, int words, int z, int n) {
  const uint32_t* ir = inside_bits + (size_t)x * words;
  const int wi = z >> 5;
  const uint32_t w0 = __ldg(ir + wi), w1 = wi + 1 < words ? __ldg(ir + wi + 1) : 0u;
  return __funnelshift_r(w0, w1, z & 31) & ((1u << n) - 1u);  // n in 1..16
}

// OR the RGB bytes of `col` into the 48-byte image `o` of a group at every voxel position whose bit is set
__device__ __forceinline__ void paint16(uint32_t bits, uint32_t col, uint32_t* o) {
  const uint32_t r8 = col & 0xff, g8 = (col >> 8) & 0xff, b8 = (col >> 16) & 0xff;
  const uint32_t w0 = r8 | (g8 << 8) | (b8 << 16) | (r8 << 24);
  const uint32_t w1 = g8 | (b8 << 8) | (r8 << 16) | (g8 << 24);
  const uint32_t w2 = b8 | (r8 << 8) | (g8 << 16) | (b8 << 24);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t t[3];
    expand4((bits >> (4 * q)) & 0xfu, w0, w1, w2, t);
    o[3 * q] |= t[0]; o[3 * q + 1] |= t[1]; o[3 * q + 2] |= t[2];
  }
}

__global__ void __launch_bounds__(256)
global_fold_bits_ragged_kernel(int W, int H, int D, int x_begin, int x_count, const uint32_t* __restrict__ inside_bits,
                               int c, const uint32_t* __restrict__ mask_bits, int wpr, const uint8_t* __restrict__ colour_hw,
                               uint8_t* __restrict__ out, unsigned long long magic_d, unsigned long long magic_h) {
  __shared__ uint4 stage[8][96];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rows = (uint32_t)x_count * (uint32_t)H;
  const uint32_t nvox = rows * (uint32_t)D;                    // < 2^31 (the host splits larger slabs)
  const uint32_t groups = (nvox + 15u) >> 4;
  const uint32_t warp_groups = (groups + 31u) >> 5;
  const uint32_t full16 = (uint32_t)(((uint64_t)nvox * 3u) >> 4), tail = (uint32_t)(((uint64_t)nvox * 3u) & 15u);
  const int words = (D + 31) >> 5;
  for (uint32_t wg = blockIdx.x * 8u + warp; wg < warp_groups; wg += gridDim.x * 8u) {
    const uint32_t g = wg * 32u + lane;
    uint32_t o[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) o[k] = 0u;
    if (g < groups) {
      const uint32_t v0 = g << 4;
      uint32_t row = div_magic(v0, (uint32_t)D, magic_d);
      int z = (int)(v0 - row * (uint32_t)D), done = 0;
#pragma unroll
      for (int piece = 0; piece < 2; ++piece) {
        if (done < 16 && row < rows) {
          const int n = min(16 - done, D - z);
          const uint32_t xi = div_magic(row, (uint32_t)H, magic_h);
          const uint32_t y = row - xi * (uint32_t)H;
          const int x = x_begin + (int)xi, xb = x + 32;
          const uint32_t* mrow = mask_bits + (size_t)y * wpr;
          if ((__ldg(mrow + (xb >> 5)) >> (xb & 31)) & 1u) {
            const uint32_t in16 = inside16_ragged(inside_bits, x, words, z, n);
            const int lo = c - z - 15 + 32;                  // >= 17 and inside the padded row whenever in16 != 0
            uint32_t m16 = 0;
            if (in16 && lo >= 0 && (lo >> 5) + 1 < wpr) {
              const uint32_t w0 = __ldg(mrow + (lo >> 5)), w1 = __ldg(mrow + (lo >> 5) + 1);
              m16 = __funnelshift_r(w0, w1, lo & 31) & 0xffffu;
            }
            const uint32_t bits = in16 & (__brev(m16) >> 16);
            if (bits) {
              const uint8_t* cc = colour_hw + ((size_t)y * W + x) * 3;
              paint16(bits << done, (uint32_t)cc[0] | ((uint32_t)cc[1] << 8) | ((uint32_t)cc[2] << 16), o);
            }
          }
          done += n;
          ++row;
          z = 0;
        }
      }
    }
    uint4* mine = &stage[warp][lane * 3];
    mine[0] = make_uint4(o[0], o[1], o[2], o[3]);
    mine[1] = make_uint4(o[4], o[5], o[6], o[7]);
    mine[2] = make_uint4(o[8], o[9], o[10], o[11]);
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)wg * 96;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const uint32_t q = wg * 96u + (uint32_t)(p * 32 + lane);  // uint4 index within the slab
      if (q < full16) {
        __stcs(dst + p * 32 + lane, stage[warp][p * 32 + lane]);
      } else if (q == full16 && tail) {                        // the slab's last bytes: fewer than 16
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(&stage[warp][p * 32 + lane]);
        uint8_t* db = reinterpret_cast<uint8_t*>(dst + p * 32 + lane);
        for (uint32_t b = 0; b < tail; ++b) db[b] = sb[b];
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Bit-level fast path of the fused part_carve kernel (all groups at 90 degrees, z-separable table with
// src0 = c - z and src2 = x + c2):
//   keep(x,y,z) = grid[x,y,z] != 0 && inside(x,z) && occ[c-z, y, x+c2] && OR_g (g in gm[y][x] && g in gm[y][c-z])
//   pack_group_bits_kernel : gbits[g][y][1 + x/32] bit x%32 = bit g of the group image gm[y][x]  (one zero word of
//                            padding on each side)
//   part_copy_bits_kernel / part_clear_kernel : the copy-then-clear pair described below
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rgb16_occupancy(const uint4& a, const uint4& b, const uint4& c) {
  const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {                         // 4 voxels = 3 words
    const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
    if (w0 & 0x00ffffffu) bits |= 1u << (4 * q);
    if ((w0 & 0xff000000u) | (w1 & 0x0000ffffu)) bits |= 2u << (4 * q);
    if ((w1 & 0xffff0000u) | (w2 & 0x000000ffu)) bits |= 4u << (4 * q);
    if (w2 & 0xffffff00u) bits |= 8u << (4 * q);
  }
  return bits;
}

__global__ void __launch_bounds__(256)
pack_group_bits_kernel(const uint32_t* __restrict__ gm_hw, int H, int W, int n_groups, int xwp, uint32_t* __restrict__ gbits) {
  const int lane = threadIdx.x & 31;
  const int64_t tasks = (int64_t)H * xwp;                   // one warp per (y, word): 32 pixels, one ballot per group
  for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < tasks; t += (int64_t)gridDim.x * 8) {
    const int w = (int)(t % xwp), y = (int)(t / xwp);
    const int x = (w - 1) * 32 + lane;
    const uint32_t m = (x >= 0 && x < W) ? __ldg(gm_hw + (size_t)y * W + x) : 0u;
    for (int g = 0; g < n_groups; ++g) {
      const uint32_t v = __ballot_sync(0xffffffffu, (m >> g) & 1u);
      if (lane == 0) gbits[((size_t)g * H + y) * xwp + w] = v;
    }
  }
}

// 16 bits [lo, lo+16) of a padded bit row, returned in REVERSE order (bit j = row bit lo+15-j); 0 when out of range
__device__ __forceinline__ uint32_t rev16_bits(const uint32_t* __restrict__ row, int lo, int xwp) {
  if (lo < 0 || (lo >> 5) + 1 >= xwp) return 0u;
  const uint32_t w0 = __ldg(row + (lo >> 5)), w1 = __ldg(row + (lo >> 5) + 1);
  return __brev(__funnelshift_r(w0, w1, lo & 31) & 0xffffu) >> 16;
}

// ------------------------------------------------------------------------------------------
// Copy-then-clear (about 6.3 bytes of traffic per voxel; a separate occupancy pre-pass made it 9).  The source
// occupancy occ[c-z, y, x+c2] is a transposed access, so it needs a pass over the grid before it can be used; instead
// of spending that pass on occupancy bits alone, pass A already writes the output with every term that is local to
// the voxel (own occupancy, inside bits, group masks) and records the z-packed occupancy bits of the input and the
// z-packed "alive" bits of what it wrote.  Pass B then only reads bits: alive & ~source-occupancy tells it which
// voxels still have to be cleared (none at all for a grid that is already 4-way symmetric, e.g. the output of
// global_carve) and it rewrites just those 96-byte runs.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mask96(uint4 (&v)[6], uint32_t keep) {
  uint32_t* w = reinterpret_cast<uint32_t*>(v);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t o[3];
    expand4((keep >> (4 * q)) & 0xfu, w[3 * q], w[3 * q + 1], w[3 * q + 2], o);
    w[3 * q] = o[0]; w[3 * q + 1] = o[1]; w[3 * q + 2] = o[2];
  }
}

// Slab form (multi-GPU): the clear pass of an output x slab [x0, x1) needs occ[c - z, y, x + c2], i.e. the z range
// [x0 + c2, x1 + c2) of EVERY row of the (replicated) input.  One thread per (row, word): 32 voxels = 96 contiguous bytes.
__global__ void __launch_bounds__(256)
occ_zrange_bits_kernel(const uint8_t* __restrict__ grid, int64_t rows, int D, int w_begin, int w_count,
                       uint32_t* __restrict__ occz) {
  const int words = D >> 5;
  const int64_t n = rows * w_count;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / w_count;
    const int w = w_begin + (int)(t - row * w_count);
    const uint4* src = reinterpret_cast<const uint4*>(grid + (row * D + (int64_t)w * 32) * 3);
    const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3), e = __ldg(src + 4), f = __ldg(src + 5);
    occz[row * words + w] = rgb16_occupancy(a, b, c) | (rgb16_occupancy(d, e, f) << 16);
  }
}

// pass A: thread = 16 voxels of one z-row (coalesced 48-byte loads and stores, like part_fold_bits_kernel); lane pairs
// merge their 16 occupancy / alive bits into z-packed words occz / alive [x][y][z/32].
__global__ void __launch_bounds__(256)
part_copy_bits_kernel(const uint8_t* __restrict__ grid, int W, int H, int D, const uint32_t* __restrict__ inside_bits,
                      int c, const uint32_t* __restrict__ gm_hw, const uint32_t* __restrict__ gbits, int xwp,
                      uint32_t* __restrict__ occz, uint32_t* __restrict__ alive, uint8_t* __restrict__ out,
                      unsigned long long magic_gpr, unsigned long long magic_h, uint32_t g_begin, uint32_t g_end,
                      uint32_t grid_g0 /* first group held by `grid`: 0 = whole grid, g_begin = the slab only */) {
  // groups [g_begin, g_end) of the full grid (an x slab, multiples of 32 groups since D % 32 == 0 ... H * D/16 even);
  // `out` starts at the slab's first voxel; occz == nullptr: the caller already has the input's occupancy bits
  const uint32_t gpr = (uint32_t)D >> 4;                    // thread groups per z-row
  const uint32_t groups = g_end;                            // < 2^31 (checked by the host); even: D % 32 == 0
  const int words = D >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  out -= (size_t)g_begin * 48;                              // 16 voxels x 3 bytes per group
  grid -= (size_t)grid_g0 * 48;
  for (uint32_t wb = g_begin + blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wb < groups; wb += stride) {
    const uint32_t g = wb + lane;
    const bool in = g < groups;
    uint32_t occ = 0, keep = 0;
    if (in) {
      const size_t v0 = (size_t)g * 16;
      const uint32_t r = div_magic(g, gpr, magic_gpr);      // 32-bit index arithmetic, see global_fold_bits_kernel
      const int z0 = (int)(g - r * gpr) << 4;
      const uint32_t xq = div_magic(r, (uint32_t)H, magic_h);
      const int y = (int)(r - xq * (uint32_t)H), x = (int)xq;
      const uint4* src = reinterpret_cast<const uint4*>(grid + v0 * 3);
      const uint4 a = __ldg(src), b = __ldg(src + 1), cc = __ldg(src + 2);
      occ = keep = rgb16_occupancy(a, b, cc);
      const uint32_t self = keep ? __ldg(gm_hw + (size_t)y * W + x) : 0u;
      if (self == 0u) keep = 0u;
      if (keep) keep &= (__ldg(inside_bits + (size_t)x * words + (z0 >> 5)) >> (z0 & 31)) & 0xffffu;
      if (keep) {
        const int lo = c - z0 - 15 + 32;                    // padded bit index of source x' = c - z0 - 15
        uint32_t grp = 0, rem = self;
        while (rem) {
          const int gi = __ffs(rem) - 1;
          rem &= rem - 1;
          grp |= rev16_bits(gbits + ((size_t)gi * H + y) * xwp, lo, xwp);
        }
        keep &= grp;
      }
      uint32_t m[12];
#pragma unroll
      for (int q = 0; q < 4; ++q) expand4((keep >> (4 * q)) & 0xfu, 0xffffffffu, 0xffffffffu, 0xffffffffu, m + 3 * q);
      uint4* dst = reinterpret_cast<uint4*>(out + v0 * 3);
      dst[0] = make_uint4(a.x & m[0], a.y & m[1], a.z & m[2], a.w & m[3]);
      dst[1] = make_uint4(b.x & m[4], b.y & m[5], b.z & m[6], b.w & m[7]);
      dst[2] = make_uint4(cc.x & m[8], cc.y & m[9], cc.z & m[10], cc.w & m[11]);
    }
    const uint32_t occ_hi = __shfl_xor_sync(0xffffffffu, occ, 1), keep_hi = __shfl_xor_sync(0xffffffffu, keep, 1);
    if (in && !(lane & 1)) {                                // even lane: z0 is a multiple of 32
      if (occz) occz[g >> 1] = occ | (occ_hi << 16);        // (g * 16) / 32 == ((x * H + y) * D + z0) / 32
      alive[g >> 1] = keep | (keep_hi << 16);
    }
  }
}

// pass A for ragged rows (see global_fold_bits_ragged_kernel): thread = 16 voxels in FLAT order, at most two z-row
// pieces.  The z-packed occupancy / alive words of a row no longer belong to one lane pair, so every piece ORs its bits
// into the (zeroed) arrays with at most two atomics each -- only occupied / surviving pieces issue any.
__global__ void __launch_bounds__(256)
part_copy_bits_ragged_kernel(const uint8_t* __restrict__ grid, int W, int H, int D, const uint32_t* __restrict__ inside_bits,
                             int c, const uint32_t* __restrict__ gm_hw, const uint32_t* __restrict__ gbits, int xwp,
                             uint32_t* __restrict__ occz, uint32_t* __restrict__ alive, uint8_t* __restrict__ out,
                             unsigned long long magic_d, unsigned long long magic_h, uint32_t nvox) {
  const uint32_t groups = (nvox + 15u) >> 4;
  const uint32_t rows = (uint32_t)W * (uint32_t)H;
  const int words = (D + 31) >> 5;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const uint32_t v0 = g << 4;
    const bool whole = v0 + 16u <= nvox;                      // false only for the last group of a grid with nvox % 16 != 0
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a, cc = a;
    if (whole) {
      const uint4* src = reinterpret_cast<const uint4*>(grid + (size_t)v0 * 3);
      a = __ldg(src); b = __ldg(src + 1); cc = __ldg(src + 2);
    } else {
      uint32_t w[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) w[k] = 0u;
      const uint32_t nb = (nvox - v0) * 3u;
      for (uint32_t k = 0; k < nb; ++k) w[k >> 2] |= (uint32_t)grid[(size_t)v0 * 3 + k] << (8 * (k & 3u));
      a = make_uint4(w[0], w[1], w[2], w[3]); b = make_uint4(w[4], w[5], w[6], w[7]); cc = make_uint4(w[8], w[9], w[10], w[11]);
    }
    const uint32_t occ = rgb16_occupancy(a, b, cc);
    uint32_t keep = 0;
    if (occ) {
      uint32_t row = div_magic(v0, (uint32_t)D, magic_d);
      int z = (int)(v0 - row * (uint32_t)D), done = 0;
#pragma unroll
      for (int piece = 0; piece < 2; ++piece) {
        if (done < 16 && row < rows) {
          const int n = min(16 - done, D - z);
          const uint32_t o = (occ >> done) & ((1u << n) - 1u);
          if (o) {
            const size_t wbase = (size_t)row * words + (z >> 5);
            const int sh = z & 31;
            atomicOr(occz + wbase, o << sh);
            if (sh + n > 32) atomicOr(occz + wbase + 1, o >> (32 - sh));
            const uint32_t xq = div_magic(row, (uint32_t)H, magic_h);
            const int y = (int)(row - xq * (uint32_t)H), x = (int)xq;
            const uint32_t self = __ldg(gm_hw + (size_t)y * W + x);
            uint32_t k = self ? o & inside16_ragged(inside_bits, x, words, z, n) : 0u;
            if (k) {
              const int lo = c - z - 15 + 32;
              uint32_t grp = 0, rem = self;
              while (rem) {
                const int gi = __ffs(rem) - 1;
                rem &= rem - 1;
                grp |= rev16_bits(gbits + ((size_t)gi * H + y) * xwp, lo, xwp);
              }
              k &= grp;
            }
            if (k) {
              atomicOr(alive + wbase, k << sh);
              if (sh + n > 32) atomicOr(alive + wbase + 1, k >> (32 - sh));
              keep |= k << done;
            }
          }
          done += n;
          ++row;
          z = 0;
        }
      }
    }
    uint32_t m[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) expand4((keep >> (4 * q)) & 0xfu, 0xffffffffu, 0xffffffffu, 0xffffffffu, m + 3 * q);
    const uint4 r0 = make_uint4(a.x & m[0], a.y & m[1], a.z & m[2], a.w & m[3]);
    const uint4 r1 = make_uint4(b.x & m[4], b.y & m[5], b.z & m[6], b.w & m[7]);
    const uint4 r2 = make_uint4(cc.x & m[8], cc.y & m[9], cc.z & m[10], cc.w & m[11]);
    if (whole) {
      uint4* dst = reinterpret_cast<uint4*>(out + (size_t)v0 * 3);
      dst[0] = r0; dst[1] = r1; dst[2] = r2;
    } else {
      const uint32_t w[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
      const uint32_t nb = (nvox - v0) * 3u;
      for (uint32_t k = 0; k < nb; ++k) out[(size_t)v0 * 3 + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3u)));
    }
  }
}

// 32x32 bit-matrix transpose across a warp: lane l enters with row l, lane b leaves with column b (bit j = row j's
// bit b).  Five butterfly stages (swap the off-diagonal s x s blocks), one shuffle each.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t v, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t lo = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t p = __shfl_xor_sync(0xffffffffu, v, s);
    v = (lane & s) ? ((v & ~lo) | ((p >> s) & lo)) : ((v & lo) | ((p << s) & ~lo));
  }
  return v;
}

constexpr int kClearX = 256;     // x per CTA task: one thread each
constexpr int kClearZW = 8;      // z words per CTA task (256 voxels along z = one 32-byte sector of a z-packed bit row)
constexpr int kClearRowW = kClearZW + 1;   // staged words per source row (the c2 shift may straddle a word)

// pass B: CTA = (y, 256 x, 256 z), thread = one x.  The source occupancy of output (x, y, z) is occ[c - z, y, x + c2]:
// a transposed access.  Every thread first reads its own 256 alive bits (one full 32-byte sector); tiles without alive
// voxels (empty space, most of a monument grid) end there.  Otherwise thread i stages the 256 + 32 bits
// [xb*256 + c2, ...) of source row x' = c - z0 - i in shared memory (again whole sectors), and each warp turns the
// 32 rows x 32 bits it needs per z word into per-lane source words with a butterfly bit transpose.  Only runs with
// alive voxels whose source is empty are rewritten.
__global__ void __launch_bounds__(kClearX)
part_clear_kernel(int W, int H, int D, int c, int c2, const uint32_t* __restrict__ occz,
                  const uint32_t* __restrict__ alive, uint8_t* __restrict__ out, int x_begin, int x_count,
                  const uint32_t* const* __restrict__ peer_occ, int rows_per_rank) {
  // peer_occ != nullptr (multi-GPU, sharded input): source row sx lives in the occupancy array of rank sx / rows_per_rank,
  // peer_occ[rank] = that rank's array (same full-grid layout, mapped over NVLink); the row is read straight from there
  // output x in [x_begin, x_begin + x_count); `out` starts at the slab's first voxel; bit arrays are indexed by the
  // full grid
  __shared__ uint32_t s_occ[kClearX * kClearRowW];          // row stride 9 words: conflict-free both ways
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int words = (D + 31) >> 5;                            // ragged rows: the last word's bits beyond D are never alive
  const bool rows16 = (D & 15) == 0;                          // 32-voxel runs start on 16-byte boundaries
  const int xb_n = (x_count + kClearX - 1) / kClearX, zb_n = (words + kClearZW - 1) / kClearZW;
  const int x_end = x_begin + x_count;
  out -= (size_t)x_begin * H * D * 3;
  const int64_t tasks = (int64_t)H * zb_n * xb_n;
  const bool vec = (words % kClearZW) == 0;                  // rows of 8 words start on 32-byte boundaries
  for (int64_t t = blockIdx.x; t < tasks; t += gridDim.x) {
    const int xb = (int)(t % xb_n);
    const int64_t r = t / xb_n;
    const int zb = (int)(r % zb_n), y = (int)(r / zb_n);
    const int x = x_begin + xb * kClearX + (int)threadIdx.x, zw0 = zb * kClearZW;
    uint32_t a[kClearZW];
#pragma unroll
    for (int k = 0; k < kClearZW; ++k) a[k] = 0u;
    if (x < x_end) {
      const uint32_t* row = alive + ((size_t)x * H + y) * words + zw0;
      if (vec) {
        const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(row)), v1 = __ldg(reinterpret_cast<const uint4*>(row) + 1);
        a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w; a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
      } else {
#pragma unroll
        for (int k = 0; k < kClearZW; ++k)
          if (zw0 + k < words) a[k] = __ldg(row + k);
      }
    }
    uint32_t any = 0u;
#pragma unroll
    for (int k = 0; k < kClearZW; ++k) any |= a[k];
    if (!__syncthreads_or(any != 0u)) continue;              // also orders the previous task's s_occ reads before the writes below
    {
      const int sx = c - zw0 * 32 - (int)threadIdx.x;        // source row of z = z0 + threadIdx.x
      const int wb = (x_begin + xb * kClearX + c2) >> 5;      // arithmetic shift; bits outside [0, D) read 0
      const bool row_ok = sx >= 0 && sx < W;
      const uint32_t* base = occz;
      if (peer_occ != nullptr && row_ok) base = peer_occ[sx / rows_per_rank];
      const uint32_t* row = base + ((size_t)(row_ok ? sx : 0) * H + y) * words;
      uint32_t* so = s_occ + threadIdx.x * kClearRowW;
#pragma unroll
      for (int k = 0; k < kClearRowW; ++k) so[k] = 0u;
      if (row_ok) {
        // only the words this tile's x range consumes: bits [xs, xs + nb) of the row.  A 128-wide slab of an 8-GPU run
        // needs ONE 16-byte load per source row (these are the loads that cross NVLink in the peer form); the first
        // version always fetched 32 bytes, and nine scalar words per row whenever the slab did not start on a multiple
        // of 256 (every odd rank).
        const int xs = x_begin + xb * kClearX + c2;
        const int nb = min(kClearX, x_end - (x_begin + xb * kClearX));
        const int wlo = max(wb, 0), whi = min(wb + (((xs & 31) + nb + 31) >> 5), words);
        if ((words & 3) == 0) {                              // every bit row starts on a 16-byte boundary
          for (int q = wlo & ~3; q < whi; q += 4) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(row + q));
            const uint32_t t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k = q + e - wb;
              if (k >= 0 && k < kClearRowW) so[k] = t[e];
            }
          }
        } else {
          for (int w = wlo; w < whi; ++w) so[w - wb] = __ldcg(row + w);
        }
      }
    }
    __syncthreads();
    const int sh = (x_begin + xb * kClearX + c2) & 31;
#pragma unroll
    for (int k = 0; k < kClearZW; ++k) {
      if (!__any_sync(0xffffffffu, a[k] != 0u)) continue;
      // lane l: 32 source bits of row z = z0 + 32k + l starting at this warp's x tile; after the transpose lane b holds
      // bit j = occ[c - (z0 + 32k + j), y, x + c2]
      const uint32_t* rp = s_occ + (k * 32 + lane) * kClearRowW + warp;
      const uint32_t src = warp_transpose32(__funnelshift_r(rp[0], rp[1], sh), lane);
      uint32_t clear = a[k] & ~src;
      if (clear == 0u) continue;
      uint8_t* run = out + (((size_t)x * H + y) * D + (size_t)(zw0 + k) * 32) * 3;
      if (rows16 && (zw0 + k) * 32 + 32 <= D) {
        uint4* dst = reinterpret_cast<uint4*>(run);
        uint4 v[6] = {dst[0], dst[1], dst[2], dst[3], dst[4], dst[5]};
        mask96(v, a[k] & src);
#pragma unroll
        for (int q = 0; q < 6; ++q) dst[q] = v[q];
      } else {                                                 // ragged row or its last partial word: voxel by voxel
        while (clear) {
          const int j = __ffs(clear) - 1;
          clear &= clear - 1u;
          run[3 * j] = 0; run[3 * j + 1] = 0; run[3 * j + 2] = 0;
        }
      }
    }
  }
}

// Scalar path for any D.
template <bool RGB>
__global__ void __launch_bounds__(256)
global_fold_scalar_kernel(int W, int H, int D, const int32_t* __restrict__ table,
                          const uint8_t* __restrict__ mask_hw, const uint8_t* __restrict__ colour_hw,
                          uint8_t* __restrict__ out) {
  const int64_t n = (int64_t)W * H * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % D);
    const int64_t r = i / D;
    const int y = (int)(r % H), x = (int)(r / H);
    bool keep = false;
    if (mask_hw[(size_t)y * W + x]) {
      const int32_t e = table[(size_t)x * D + z];
      keep = e >= 0 && mask_hw[(size_t)y * W + (e >> 16)];
    }
    if (RGB) {
      const uint8_t* c = colour_hw + ((size_t)y * W + x) * 3;
      out[3 * i + 0] = keep ? c[0] : 0;
      out[3 * i + 1] = keep ? c[1] : 0;
      out[3 * i + 2] = keep ? c[2] : 0;
    } else {
      out[i] = keep ? colour_hw[(size_t)y * W + x] : 0;
    }
  }
}

// apply_colored_mask_to_voxel_grid: out[x,y,z,:] = colour[y,x,:] if carved[x,y,z] == 1 else 0
__global__ void __launch_bounds__(256)
colourise_kernel(const uint8_t* __restrict__ carved, int W, int H, int D, const uint8_t* __restrict__ colour_hw,
                 uint8_t* __restrict__ out) {
  const int64_t n = (int64_t)W * H * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D;
    const int y = (int)(r % H), x = (int)(r / H);
    const bool keep = carved[i] == 1;
    const uint8_t* c = colour_hw + ((size_t)y * W + x) * 3;
    out[3 * i + 0] = keep ? c[0] : 0;
    out[3 * i + 1] = keep ? c[1] : 0;
    out[3 * i + 2] = keep ? c[2] : 0;
  }
}

// ------------------------------------------------------------------------------------------
// part_carve, all groups at 90 degrees, fused:
//   keep(x,y,z) = grid[x,y,z] != 0 && table[x,z] >= 0 && grid[s0,y,s2] != 0 &&
//                 (G[x,y] & MM[x,y] & G[s0,y] & MM[s0,y]) != 0
// G  = bit g set when pixel (x,y) belongs to group g's 2-D mask; MM = the same masks after the reference's
// _mask_to_wh (transposed when W == H).  Both are (H,W) uint32 images.  Output = grid where keep else 0.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
part_fold_kernel(const uint8_t* __restrict__ grid, int W, int H, int D, const int32_t* __restrict__ table,
                 const uint32_t* __restrict__ gm_hw, uint8_t* __restrict__ out) {
  const int64_t n = (int64_t)W * H * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t* p = grid + 3 * i;
    uint8_t r = p[0], g = p[1], b = p[2];
    bool keep = false;
    if (r | g | b) {
      const int z = (int)(i % D);
      const int64_t q = i / D;
      const int y = (int)(q % H), x = (int)(q / H);
      const uint32_t self = gm_hw[(size_t)y * W + x];
      const int32_t e = table[(size_t)x * D + z];
      if (self && e >= 0) {
        const int s0 = e >> 16, s2 = e & 0xffff;
        const uint32_t src = gm_hw[(size_t)y * W + s0];
        if ((self & src) && rgb_nonzero(grid + 3 * (((size_t)s0 * H + y) * D + s2))) keep = true;
      }
    }
    out[3 * i + 0] = keep ? r : 0;
    out[3 * i + 1] = keep ? g : 0;
    out[3 * i + 2] = keep ? b : 0;
  }
}

// Building blocks of the general (any angle) part_carve / left_right_guided_carve path -------------------
// occ[x,y,z] = any(grid[x0+x, y0+y, z0+z, :] > 0) && (sel == null || sel[x,y])  over a crop
__global__ void __launch_bounds__(256)
crop_occupancy_kernel(const uint8_t* __restrict__ grid, int H, int D, int x0, int y0, int z0, int w, int h, int d,
                      const uint8_t* __restrict__ sel_wh, uint8_t* __restrict__ occ) {
  const int64_t n = (int64_t)w * h * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % d);
    const int64_t q = i / d;
    const int y = (int)(q % h), x = (int)(q / h);
    const uint8_t* p = grid + 3 * ((((size_t)(x0 + x)) * H + (y0 + y)) * D + (z0 + z));
    occ[i] = (rgb_nonzero(p) && (sel_wh == nullptr || sel_wh[(size_t)x * h + y])) ? 1 : 0;
  }
}

// part_carve accumulate: final[v] = grid[v] where sel[x,y] && grid[v] != 0 && carved[v] != 0
__global__ void __launch_bounds__(256)
accumulate_part_kernel(const uint8_t* __restrict__ grid, const uint8_t* __restrict__ carved, int W, int H, int D,
                       const uint8_t* __restrict__ sel_wh, uint8_t* __restrict__ final_grid) {
  const int64_t n = (int64_t)W * H * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!carved[i]) continue;
    const int64_t q = i / D;
    const int y = (int)(q % H), x = (int)(q / H);
    const uint8_t* p = grid + 3 * i;
    if (sel_wh[(size_t)x * H + y] && rgb_nonzero(p)) {
      final_grid[3 * i] = p[0]; final_grid[3 * i + 1] = p[1]; final_grid[3 * i + 2] = p[2];
    }
  }
}

// left_right_guided_carve paste (:199-201) over the bbox of component `comp`:
//   if labels == comp: out = 0 ; if kept && src != 0: out = src      (src = the call's INPUT grid)
__global__ void __launch_bounds__(256)
paste_component_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ labels, int comp,
                       const uint8_t* __restrict__ kept, int H, int D, int x0, int y0, int z0, int w, int h, int d,
                       uint8_t* __restrict__ out) {
  const int64_t n = (int64_t)w * h * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % d);
    const int64_t q = i / d;
    const int y = (int)(q % h), x = (int)(q / h);
    const size_t v = (((size_t)(x0 + x)) * H + (y0 + y)) * D + (z0 + z);
    const uint8_t* p = src + 3 * v;
    if (kept[i] && rgb_nonzero(p)) {
      out[3 * v] = p[0]; out[3 * v + 1] = p[1]; out[3 * v + 2] = p[2];
    } else if (labels[v] == comp) {
      out[3 * v] = 0; out[3 * v + 1] = 0; out[3 * v + 2] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Connected components, 6-connectivity, ids in raster order of each component's first voxel
// (scipy.ndimage.label default).  Union-find on flat indices with the smaller index as root.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colour_mask_kernel(const uint8_t* __restrict__ grid, int64_t n, uint32_t colour, uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t* p = grid + 3 * i;
    mask[i] = (p[0] | (p[1] << 8) | (p[2] << 16)) == colour;
  }
}

__device__ __forceinline__ int32_t uf_find(volatile int32_t* parent, int32_t i) {
  int32_t p = parent[i];
  while (p != i) { i = p; p = parent[i]; }
  return i;
}

__device__ __forceinline__ void uf_union(int32_t* parent, int32_t a, int32_t b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) { int32_t t = a; a = b; b = t; }
    const int32_t old = atomicMin(parent + b, a);       // hook the larger root under the smaller
    if (old == b) return;
    b = old;
  }
}

__global__ void __launch_bounds__(256)
ccl_init_kernel(const uint8_t* __restrict__ mask, int64_t n, int n2, int32_t* __restrict__ parent) {
  // Every voxel starts linked to the first voxel of its run inside the warp's 32 consecutive voxels (same row, contiguous
  // along the last axis: connected under every connectivity used here), found from two ballots; the merge kernels then
  // only join runs, not voxels.
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (n + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t i = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n;
    const bool on = in && mask[i];
    const uint32_t m = __ballot_sync(0xffffffffu, on);
    const uint32_t rowstart = __ballot_sync(0xffffffffu, in && (i % n2) == 0);
    const uint32_t starts = m & (~(m << 1) | rowstart | 1u);
    if (!in) continue;
    const uint32_t below = starts & (0xffffffffu >> (31 - lane));
    parent[i] = on ? (int32_t)(i - lane + (31 - __clz(below))) : -1;
  }
}

__global__ void __launch_bounds__(256)
ccl_merge_kernel(const uint8_t* __restrict__ mask, int n0, int n1, int n2, int32_t* __restrict__ parent) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!mask[i]) continue;
    const int c = (int)(i % n2);
    const int64_t r = i / n2;
    const int b = (int)(r % n1);
    const int a = (int)(r / n1);
    const int64_t plane = (int64_t)n1 * n2;
    // along the row only run heads that continue a run of the previous warp need a join (ccl_init linked the rest);
    // across rows / planes a join is needed only where the pair (i-1, its neighbour) did not already make it
    const bool prev = c > 0 && mask[i - 1];
    if (prev && (i & 31) == 0) uf_union(parent, (int32_t)i, (int32_t)(i - 1));   // i % 32 = its lane in ccl_init_kernel
    if (b > 0 && mask[i - n2] && !(prev && mask[i - n2 - 1])) uf_union(parent, (int32_t)i, (int32_t)(i - n2));
    if (a > 0 && mask[i - plane] && !(prev && mask[i - plane - 1])) uf_union(parent, (int32_t)i, (int32_t)(i - plane));
  }
}

// 26-connected 3-D variant (scipy.ndimage.label with structure = ones((3,3,3)), voxel_utils.py:24): the 13 neighbours
// that precede a voxel in raster order.
__global__ void __launch_bounds__(256)
ccl_merge26_kernel(const uint8_t* __restrict__ mask, int n0, int n1, int n2, int32_t* __restrict__ parent) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  const int64_t plane = (int64_t)n1 * n2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!mask[i]) continue;
    const int c = (int)(i % n2);
    const int64_t r = i / n2;
    const int b = (int)(r % n1);
    const int a = (int)(r / n1);
    for (int da = -1; da <= 0; ++da)
      for (int db = -1; db <= 1; ++db)
        for (int dc = -1; dc <= 1; ++dc) {
          if (da == 0 && (db > 0 || (db == 0 && dc >= 0))) continue;       // only raster-order predecessors
          const int aa = a + da, bb = b + db, cc = c + dc;
          if (aa < 0 || bb < 0 || bb >= n1 || cc < 0 || cc >= n2) continue;
          const int64_t j = i + da * plane + (int64_t)db * n2 + dc;
          if (mask[j]) uf_union(parent, (int32_t)i, (int32_t)j);
        }
  }
}

// 8-connected 2-D variant (skimage.measure.label's default for images, camera_estimation.py:263): W, N, NW, NE.
__global__ void __launch_bounds__(256)
ccl_merge8_kernel(const uint8_t* __restrict__ mask, int H, int W, int32_t* __restrict__ parent) {
  const int64_t n = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!mask[i]) continue;
    const int x = (int)(i % W), y = (int)(i / W);
    if (x > 0 && mask[i - 1]) uf_union(parent, (int32_t)i, (int32_t)(i - 1));
    if (y > 0) {
      if (mask[i - W]) uf_union(parent, (int32_t)i, (int32_t)(i - W));
      if (x > 0 && mask[i - W - 1]) uf_union(parent, (int32_t)i, (int32_t)(i - W - 1));
      if (x + 1 < W && mask[i - W + 1]) uf_union(parent, (int32_t)i, (int32_t)(i - W + 1));
    }
  }
}

// parent[i] <- root(i); is_root[i] = (root == i)
__global__ void __launch_bounds__(256)
ccl_flatten_kernel(int64_t n, int32_t* __restrict__ parent, uint8_t* __restrict__ is_root) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = parent[i];
    uint8_t root = 0;
    if (p >= 0) {
      const int32_t r = uf_find(parent, (int32_t)i);
      root = r == (int32_t)i;
      if (!root) parent[i] = r;                         // roots keep parent == self
    }
    is_root[i] = root;
  }
}

// rank[root] = 1-based position of the root among all roots in raster order (ordered compaction offsets)
constexpr int kRankThreads = 256, kRankPer = 16, kRankTile = kRankThreads * kRankPer;

__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_sums[kRankThreads / 32];
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
  for (int w = 0; w < kRankThreads / 32; ++w) {
    int s = warp_sums[w];
    if (w < warp) base += s;
    all += s;
  }
  *total = all;
  return base + incl - v;
}

__global__ void __launch_bounds__(kRankThreads)
root_count_kernel(const uint8_t* __restrict__ is_root, int64_t n, int32_t* __restrict__ tile_counts) {
  const int64_t i0 = (int64_t)blockIdx.x * kRankTile + (int64_t)threadIdx.x * kRankPer;
  int c = 0;
  for (int j = 0; j < kRankPer; ++j)
    if (i0 + j < n) c += is_root[i0 + j];
  int total;
  block_excl_scan(c, &total);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
root_scan_kernel(int32_t* __restrict__ tile_counts, int m, int32_t* __restrict__ n_components) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < m; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < m ? tile_counts[i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int wbase = 0, all = 0;
    for (int w = 0; w < 32; ++w) {
      int s = warp_sums[w];
      if (w < warp) wbase += s;
      all += s;
    }
    const int c = carry;
    if (i < m) tile_counts[i] = c + wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + all;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_components[0] = carry;
}

// labels[root] = rank (1-based), written in place of parent[root]; non-roots resolved in the next kernel
__global__ void __launch_bounds__(kRankThreads)
root_rank_kernel(const uint8_t* __restrict__ is_root, int64_t n, const int32_t* __restrict__ tile_offsets,
                 int32_t* __restrict__ rank_of) {
  const int64_t i0 = (int64_t)blockIdx.x * kRankTile + (int64_t)threadIdx.x * kRankPer;
  int c = 0;
  for (int j = 0; j < kRankPer; ++j)
    if (i0 + j < n) c += is_root[i0 + j];
  int total;
  int excl = block_excl_scan(c, &total);
  int next = tile_offsets[blockIdx.x] + excl + 1;
  for (int j = 0; j < kRankPer; ++j)
    if (i0 + j < n && is_root[i0 + j]) rank_of[i0 + j] = next++;
}

// labels[i] = rank_of[root(i)] (0 for background).  parent holds root indices after ccl_flatten.
__global__ void __launch_bounds__(256)
ccl_relabel_kernel(const int32_t* parent, const int32_t* __restrict__ rank_of, int64_t n, int32_t* labels) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = parent[i];
    labels[i] = p < 0 ? 0 : rank_of[p];
  }
}

// per component: bbox (min0,min1,min2,max0,max1,max2), voxel count and coordinate sums along each axis.  Lanes of a
// warp that hold the same label are combined first (match.any + warp reductions over the matching lanes), so a run of
// one component costs 10 atomics per warp instead of 10 per voxel (a minaret's 60 k voxels on six addresses made the
// per-voxel form the slowest kernel of a guided carve: 0.50 ms at 256^3).
__global__ void __launch_bounds__(256)
component_stats_kernel(const int32_t* __restrict__ labels, int n0, int n1, int n2, int capacity,
                       int32_t* __restrict__ bbox,
                       unsigned long long* __restrict__ sums /* [comp][4] = count, sum0, sum1, sum2 */) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (n + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t i = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t l = i < n ? labels[i] : 0;
    const bool valid = l > 0 && l <= capacity;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (!valid) continue;                                    // (no warp-wide call below this line includes these lanes)
    const int c = (int)(i % n2);
    const int64_t r = i / n2;
    const int b = (int)(r % n1);
    const int a = (int)(r / n1);
    const unsigned m = __match_any_sync(act, l);
    const int lo0 = __reduce_min_sync(m, a), lo1 = __reduce_min_sync(m, b), lo2 = __reduce_min_sync(m, c);
    const int hi0 = __reduce_max_sync(m, a), hi1 = __reduce_max_sync(m, b), hi2 = __reduce_max_sync(m, c);
    const unsigned s0 = __reduce_add_sync(m, (unsigned)a), s1 = __reduce_add_sync(m, (unsigned)b), s2 = __reduce_add_sync(m, (unsigned)c);
    if (lane == __ffs(m) - 1) {
      int32_t* bb = bbox + (size_t)(l - 1) * 6;
      atomicMin(bb + 0, lo0); atomicMin(bb + 1, lo1); atomicMin(bb + 2, lo2);
      atomicMax(bb + 3, hi0); atomicMax(bb + 4, hi1); atomicMax(bb + 5, hi2);
      unsigned long long* sp = sums + (size_t)(l - 1) * 4;
      atomicAdd(sp + 0, (unsigned long long)__popc(m)); atomicAdd(sp + 1, (unsigned long long)s0);
      atomicAdd(sp + 2, (unsigned long long)s1); atomicAdd(sp + 3, (unsigned long long)s2);
    }
  }
}

// per component and end (0 = minimum, 1 = maximum of the coordinate along `axis`, taken from bbox): voxel count and
// coordinate sums of the voxels AT that extreme -- the top/bottom keypoints of camera_estimation.py:329-344.
__global__ void __launch_bounds__(256)
component_extremes_kernel(const int32_t* __restrict__ labels, int n0, int n1, int n2, int axis,
                          const int32_t* __restrict__ bbox, unsigned long long* __restrict__ sums /* [comp][2][4] */) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t l = labels[i];
    if (l <= 0) continue;
    int c[3];
    c[2] = (int)(i % n2);
    const int64_t r = i / n2;
    c[1] = (int)(r % n1);
    c[0] = (int)(r / n1);
    const int32_t* bb = bbox + (size_t)(l - 1) * 6;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (c[axis] != bb[3 * e + axis]) continue;
      unsigned long long* s = sums + ((size_t)(l - 1) * 2 + e) * 4;
      atomicAdd(s + 0, 1ull); atomicAdd(s + 1, (unsigned long long)c[0]);
      atomicAdd(s + 2, (unsigned long long)c[1]); atomicAdd(s + 3, (unsigned long long)c[2]);
    }
  }
}

// mask[i] = (labels[i] == id): one component as a 0/1 volume (np.argwhere(labeled == cid), camera_estimation.py:184)
__global__ void __launch_bounds__(256)
label_equals_kernel(const int32_t* __restrict__ labels, int64_t n, int32_t id, uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = labels[i] == id;
}

// coordinate lists (n,3) int32: range of one column, then count / column sums of the rows at either end of that range
__global__ void __launch_bounds__(256)
coords_minmax_kernel(const int32_t* __restrict__ coords, int64_t n, int axis, int32_t* __restrict__ mm) {
  int lo = 0x7fffffff, hi = (int)0x80000000;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = coords[3 * i + axis];
    lo = min(lo, v); hi = max(hi, v);
  }
  for (int d = 16; d > 0; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(mm, lo); atomicMax(mm + 1, hi); }
}

__global__ void __launch_bounds__(256)
coords_extreme_sums_kernel(const int32_t* __restrict__ coords, int64_t n, int axis, const int32_t* __restrict__ mm,
                           long long* __restrict__ sums /* [2][4] = count, sum of columns 0..2 */) {
  const int lo = mm[0], hi = mm[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = coords[3 * i + axis];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (v != (e ? hi : lo)) continue;
      unsigned long long* s = reinterpret_cast<unsigned long long*>(sums) + e * 4;
      atomicAdd(s + 0, 1ull);
      for (int k = 0; k < 3; ++k) atomicAdd(s + 1 + k, (unsigned long long)(long long)coords[3 * i + k]);
    }
  }
}

__global__ void fold_info_init_kernel(int* info) {           // (max, min, max2, min2) seeds of fold_analyse_kernel
  if (threadIdx.x < 4) info[threadIdx.x] = (threadIdx.x & 1) ? 0x7fffffff : -0x7fffffff;
}

__global__ void coords_mm_init_kernel(int32_t* mm) {
  if (threadIdx.x == 0) { mm[0] = 0x7fffffff; mm[1] = (int)0x80000000; }
}

__global__ void __launch_bounds__(256) bbox_init_kernel(int32_t* __restrict__ bbox, int ncomp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncomp * 6) bbox[i] = (i % 6) < 3 ? 0x7fffffff : -1;
}

// recolor_backward_components: voxels of components flagged in `recolour[comp-1]` get `colour`
__global__ void __launch_bounds__(256)
recolour_kernel(const int32_t* __restrict__ labels, const uint8_t* __restrict__ recolour, int64_t n, uint32_t colour,
                uint8_t* __restrict__ grid) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t l = labels[i];
    if (l > 0 && recolour[l - 1]) {
      grid[3 * i] = colour & 0xff; grid[3 * i + 1] = (colour >> 8) & 0xff; grid[3 * i + 2] = (colour >> 16) & 0xff;
    }
  }
}

// ------------------------------------------------------------------------------------------
// extrude_from_surface (:213-248), in place.  One thread per column that the 2-D mask selects: find the first
// occupied voxel from the chosen end (index 0 / last when the column is empty -- np.argmax of all-False), then
// paint `depth` voxels from there in the chosen direction.
//   axis 2: columns (x,y) along z, valid = mask[y][x];   axis 0: columns (y,z) along x, valid = mask[y][z]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
extrude_kernel(uint8_t* __restrict__ grid, int W, int H, int D, const uint8_t* __restrict__ mask_hw, int mask_w,
               int axis, int sign, int depth, uint32_t colour) {
  const int ncol = axis == 2 ? W * H : H * D;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncol) return;
  int x = 0, y, z = 0, len;
  size_t stride;
  if (axis == 2) { x = t / H; y = t - x * H; if (!mask_hw[(size_t)y * mask_w + x]) return; len = D; stride = 1; }
  else { y = t / D; z = t - y * D; if (!mask_hw[(size_t)y * mask_w + z]) return; len = W; stride = (size_t)H * D; }
  uint8_t* col = grid + 3 * (axis == 2 ? ((size_t)x * H + y) * D : (size_t)y * D + z);
  int start = sign > 0 ? 0 : len - 1;
  for (int k = 0; k < len; ++k) {
    const int idx = sign > 0 ? k : len - 1 - k;
    if (rgb_nonzero(col + 3 * stride * idx)) { start = idx; break; }
  }
  const uint8_t r = colour & 0xff, g = (colour >> 8) & 0xff, b = (colour >> 16) & 0xff;
  for (int dd = 0; dd < depth; ++dd) {
    const int idx = start + sign * dd;
    if (idx < 0 || idx >= len) continue;
    uint8_t* p = col + 3 * stride * idx;
    p[0] = r; p[1] = g; p[2] = b;
  }
}

// ------------------------------------------------------------------------------------------
// partwise_carve's re-orientation (:384-385): out[z][H-1-y][x][:] = in[x][y][z][:]  ((W,H,D,3) -> (D,H,W,3)),
// 32x32 (x,z) tiles through shared memory so that reads run along z and writes along x.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reorient_kernel(const uint8_t* __restrict__ in, int W, int H, int D, uint8_t* __restrict__ out) {
  __shared__ uint8_t tile[32][32 * 3 + 4];
  const int y = blockIdx.y;
  const int tiles_z = (D + 31) / 32;
  const int x0 = (blockIdx.x / tiles_z) * 32, z0 = (blockIdx.x % tiles_z) * 32;
  for (int k = threadIdx.x; k < 32 * 96; k += 256) {
    const int xr = k / 96, bz = k - xr * 96;              // byte bz of row xr (z-major, 3 bytes per voxel)
    const int x = x0 + xr, z = z0 + bz / 3;
    tile[xr][bz] = (x < W && z < D) ? in[(((size_t)x * H + y) * D + z0) * 3 + bz] : 0;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 32 * 96; k += 256) {
    const int zr = k / 96, bx = k - zr * 96;
    const int z = z0 + zr, x = x0 + bx / 3;
    if (z < D && x < W) out[(((size_t)z * H + (H - 1 - y)) * W + x0) * 3 + bx] = tile[bx / 3][zr * 3 + bx % 3];
  }
}

// carve_voxel_grid_with_masks :76-97: out[x,y,z,c] = mask[x,y,(c)] ? grid[x,y,z,c] : 0
__global__ void __launch_bounds__(256)
mask_carve_kernel(const uint8_t* __restrict__ grid, int64_t n_vox, int D, int C, const uint8_t* __restrict__ mask_wh,
                  int MC, uint8_t* __restrict__ out) {
  const int64_t n = n_vox * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t xy = (i / C) / D;
    out[i] = mask_wh[xy * MC + (MC == 1 ? 0 : c)] ? grid[i] : 0;
  }
}

// Vector form for W % 32 == 0 and D % 32 == 0: 16-byte loads along z and 16-byte stores along x.
__global__ void __launch_bounds__(192)
reorient_vec_kernel(const uint8_t* __restrict__ in, int W, int H, int D, uint8_t* __restrict__ out) {
  __shared__ __align__(16) uint8_t tile[32][96 + 16];    // [x][z*3 + c], padded rows
  const int y = blockIdx.y;
  const int tiles_z = D >> 5;
  const int x0 = (blockIdx.x / tiles_z) << 5, z0 = (blockIdx.x % tiles_z) << 5;
  const int t = threadIdx.x;
  {
    const int xr = t / 6, part = t - xr * 6;              // 32 rows x 6 uint4
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)(x0 + xr) * H + y) * D + z0) * 3) + part);
    *reinterpret_cast<uint4*>(&tile[xr][part * 16]) = v;
  }
  __syncthreads();
  {
    const int zr = t / 6, part = t - zr * 6;              // output row z0+zr: 96 bytes = x-major RGB
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t acc = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ob = part * 16 + q * 4 + b;             // output byte within the 96-byte row
        const int xr = ob / 3, c = ob - xr * 3;
        acc |= (uint32_t)tile[xr][zr * 3 + c] << (8 * b);
      }
      w[q] = acc;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (((size_t)(z0 + zr) * H + (H - 1 - y)) * W + x0) * 3) + part;
    __stcs(dst, make_uint4(w[0], w[1], w[2], w[3]));
  }
}

// axis-2 extrusion, one WARP per masked column: the z-row is scanned 32 voxels (96 contiguous bytes) at a time.
__global__ void __launch_bounds__(256)
extrude_z_kernel(uint8_t* __restrict__ grid, int W, int H, int D, const uint8_t* __restrict__ mask_hw, int sign, int depth,
                 uint32_t colour) {
  const int lane = threadIdx.x & 31;
  const int ncol = W * H;
  for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < ncol; t += gridDim.x * 8) {
    const int x = t / H, y = t - x * H;
    if (!mask_hw[(size_t)y * W + x]) continue;
    uint8_t* col = grid + 3 * ((size_t)x * H + y) * D;
    int start = sign > 0 ? 0 : D - 1;                     // argmax of an all-False column
    for (int base = 0; base < D; base += 32) {
      const int k = base + lane;                          // k-th voxel from the scanning end
      const int idx = sign > 0 ? k : D - 1 - k;
      const bool occ = k < D && rgb_nonzero(col + 3 * (size_t)idx);
      const uint32_t m = __ballot_sync(0xffffffffu, occ);
      if (m) {
        const int first = base + __ffs(m) - 1;
        start = sign > 0 ? first : D - 1 - first;
        break;
      }
    }
    const uint8_t r = colour & 0xff, g = (colour >> 8) & 0xff, b = (colour >> 16) & 0xff;
    for (int dd = lane; dd < depth; dd += 32) {
      const int idx = start + sign * dd;
      if (idx < 0 || idx >= D) continue;
      uint8_t* p = col + 3 * (size_t)idx;
      p[0] = r; p[1] = g; p[2] = b;
    }
  }
}

int check_affine(const double* M, const double* off, Affine* A) {
  P3D_REQUIRE(M && off, "affine: null matrix/offset (host pointers)");
  for (int i = 0; i < 9; ++i) A->M[i] = M[i];
  for (int i = 0; i < 3; ++i) A->off[i] = off[i];
  return P3D_OK;
}

}  // namespace

// ==============================================================================================
// C ABI
// ==============================================================================================
P3D_API int p3d_resample_carve(const uint8_t* vol_in, int n0, int n1, int n2, const double* M, const double* off,
                               const uint8_t* mask_wh, uint8_t* vol_out, p3d_stream_t stream) {
  P3D_REQUIRE(n0 >= 0 && n1 >= 0 && n2 >= 0, "resample_carve: bad shape");
  const int64_t n = (int64_t)n0 * n1 * n2;
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(vol_in && vol_out && vol_in != vol_out, "resample_carve: null or aliased volumes");
  Affine A;
  int rc = check_affine(M, off, &A);
  if (rc) return rc;
  resample_carve_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(vol_in, n0, n1, n2, A, mask_wh, vol_out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_resample_carve_passes(uint8_t* buf_a, uint8_t* buf_b, int n0, int n1, int n2, const double* Ms,
                                      const double* offs, int n_passes, const uint8_t* mask_wh, int* result_in_b,
                                      p3d_stream_t stream) {
  P3D_REQUIRE(n0 >= 0 && n1 >= 0 && n2 >= 0 && n_passes >= 0 && result_in_b, "resample_carve_passes: bad arguments");
  *result_in_b = 0;
  const int64_t n = (int64_t)n0 * n1 * n2;
  if (n == 0 || n_passes == 0) return P3D_OK;
  P3D_REQUIRE(buf_a && buf_b && buf_a != buf_b && Ms && offs, "resample_carve_passes: null or aliased volumes");
  cudaStream_t st = p3d::as_stream(stream);
  uint8_t* src = buf_a;
  uint8_t* dst = buf_b;
  for (int p = 0; p < n_passes; ++p) {
    Affine A;
    int rc = check_affine(Ms + 9 * p, offs + 3 * p, &A);
    if (rc) return rc;
    resample_carve_kernel<<<grid_for(n, 256, 16), 256, 0, st>>>(src, n0, n1, n2, A, mask_wh, dst);
    uint8_t* t = src; src = dst; dst = t;
  }
  P3D_LAUNCH_CHECK();
  *result_in_b = src == buf_b;
  return P3D_OK;
}

P3D_API int p3d_fold_table(int n0, int n2, const double* M, const double* off, int32_t* table, int* flag,
                           p3d_stream_t stream) {
  P3D_REQUIRE(n0 > 0 && n2 > 0 && n0 < 32768 && n2 < 65536 && (int64_t)n0 * n2 < (1ll << 31), "fold_table: bad shape");
  P3D_REQUIRE(table && flag, "fold_table: null pointer");
  Affine A;
  int rc = check_affine(M, off, &A);
  if (rc) return rc;
  // the y axis must be decoupled for a per-(x,z) table to describe the pass
  P3D_REQUIRE(A.M[1] == 0.0 && A.M[3] == 0.0 && A.M[4] == 1.0 && A.M[5] == 0.0 && A.M[7] == 0.0 && A.off[1] == 0.0,
              "fold_table: transform couples the y axis");
  cudaStream_t st = p3d::as_stream(stream);
  P3D_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  fold_table_kernel<<<(n0 * n2 + 255) / 256, 256, 0, st>>>(n0, n2, A, table, flag);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_fold_gather(const uint8_t* vol_in, int n0, int n1, int n2, const int32_t* table,
                            const uint8_t* mask_wh, uint8_t* vol_out, p3d_stream_t stream) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  P3D_REQUIRE(n0 >= 0 && n1 >= 0 && n2 >= 0, "fold_gather: bad shape");
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(vol_in && vol_out && table && vol_in != vol_out, "fold_gather: null or aliased pointers");
  fold_gather_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(vol_in, n0, n1, n2, table, mask_wh, vol_out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_global_carve_fold(int W, int H, int D, const int32_t* table, const uint8_t* mask_hw,
                                  const uint8_t* colour_hw, int rgb, uint8_t* out, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0, "global_carve_fold: bad shape");
  P3D_REQUIRE(table && mask_hw && colour_hw && out, "global_carve_fold: null pointer");
  const int64_t n = (int64_t)W * H * D;
  cudaStream_t st = p3d::as_stream(stream);
  const bool vec = (D % 16 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec) {
    const int64_t warp_groups = (n / 16 + 31) / 32;
    const int blocks = grid_for(warp_groups, 8, 32);
    if (rgb) global_fold_kernel<true><<<blocks, 256, 0, st>>>(W, H, D, table, mask_hw, colour_hw, out);
    else global_fold_kernel<false><<<blocks, 256, 0, st>>>(W, H, D, table, mask_hw, colour_hw, out);
  } else {
    const int blocks = grid_for(n, 256, 16);
    if (rgb) global_fold_scalar_kernel<true><<<blocks, 256, 0, st>>>(W, H, D, table, mask_hw, colour_hw, out);
    else global_fold_scalar_kernel<false><<<blocks, 256, 0, st>>>(W, H, D, table, mask_hw, colour_hw, out);
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_mask_carve(const uint8_t* grid, int W, int H, int D, int channels, const uint8_t* mask_wh,
                           int mask_channels, uint8_t* out, p3d_stream_t stream) {
  P3D_REQUIRE(W >= 0 && H >= 0 && D >= 0 && (channels == 1 || channels == 3) &&
              (mask_channels == 1 || (mask_channels == 3 && channels == 3)), "mask_carve: bad arguments");
  const int64_t n = (int64_t)W * H * D;
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(grid && mask_wh && out, "mask_carve: null pointer");
  mask_carve_kernel<<<grid_for(n * channels, 256, 16), 256, 0, p3d::as_stream(stream)>>>(grid, n, D, channels, mask_wh,
                                                                                      mask_channels, out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_fold_analyse(const int32_t* table, int W, int D, uint32_t* inside_bits, int* info, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && D > 0, "fold_analyse: bad shape");
  P3D_REQUIRE(table && inside_bits && info, "fold_analyse: null pointer");
  cudaStream_t st = p3d::as_stream(stream);
  fold_info_init_kernel<<<1, 32, 0, st>>>(info);             // device-side seed: no pageable copy, legal under stream capture
  const int n = W * ((D + 31) / 32);
  fold_analyse_kernel<<<(n + 255) / 256, 256, 0, st>>>(table, W, D, inside_bits, info);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_pack_mask_bits(const uint8_t* mask_hw, int H, int W, uint32_t* bits, int words_per_row,
                               p3d_stream_t stream) {
  P3D_REQUIRE(H > 0 && W > 0 && words_per_row >= (W + 31) / 32 + 2, "pack_mask_bits: words_per_row too small");
  P3D_REQUIRE(mask_hw && bits, "pack_mask_bits: null pointer");
  const int n = H * words_per_row;
  pack_mask_bits_kernel<<<(n + 255) / 256, 256, 0, p3d::as_stream(stream)>>>(mask_hw, H, W, words_per_row, bits);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_global_carve_fold_bits(int W, int H, int D, int x_begin, int x_count, const uint32_t* inside_bits, int c,
                                       const uint32_t* mask_bits, int words_per_row, const uint8_t* colour_hw, int rgb,
                                       uint8_t* out, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && (D % 32 == 0 || (D >= 16 && rgb)),
              "global_carve_fold_bits: D must be a multiple of 32, or at least 16 with an RGB colour image");
  P3D_REQUIRE(x_begin >= 0 && x_count >= 0 && x_begin + x_count <= W, "global_carve_fold_bits: bad x slab");
  if (x_count == 0) return P3D_OK;
  P3D_REQUIRE(words_per_row >= (W + 31) / 32 + 2, "global_carve_fold_bits: words_per_row too small");
  P3D_REQUIRE(inside_bits && mask_bits && colour_hw && out, "global_carve_fold_bits: null pointer");
  P3D_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "global_carve_fold_bits: out must be 16-byte aligned");
  if (D % 32 != 0) {                                           // ragged rows: flat 16-voxel groups
    const uint64_t plane = (uint64_t)H * (uint64_t)D;
    int planes_max = (int)((((1ull << 31) - 1) / plane) & ~15ull);   // multiples of 16 planes keep every launch 16-byte aligned
    P3D_REQUIRE(planes_max >= 16 || (uint64_t)x_count * plane < (1ull << 31), "global_carve_fold_bits: x plane too large");
    if (planes_max < 16 || planes_max > x_count) planes_max = x_count;
    const unsigned long long magic_d = magic_for((uint64_t)planes_max * plane, (uint64_t)D);
    const unsigned long long magic_h = magic_for((uint64_t)planes_max * H, (uint64_t)H);
    cudaStream_t st = p3d::as_stream(stream);
    for (int xs = 0; xs < x_count; xs += planes_max) {
      const int xc = x_count - xs < planes_max ? x_count - xs : planes_max;
      const int64_t warp_groups = (((int64_t)xc * (int64_t)plane + 15) / 16 + 31) / 32;
      global_fold_bits_ragged_kernel<<<grid_for(warp_groups, 8, 64), 256, 0, st>>>(
          W, H, D, x_begin + xs, xc, inside_bits, c, mask_bits, words_per_row, colour_hw, out + (size_t)xs * plane * 3, magic_d,
          magic_h);
    }
    P3D_LAUNCH_CHECK();
    return P3D_OK;
  }
  const uint64_t gpr = (uint64_t)D / 16, plane_groups = (uint64_t)H * gpr;
  P3D_REQUIRE(plane_groups < (1ull << 31), "global_carve_fold_bits: x plane too large");
  int slab_max = (int)(((1ull << 31) - 1) / plane_groups);      // x planes per launch: 32-bit group indices in the kernel
  if (slab_max > x_count) slab_max = x_count;
  const unsigned long long magic_gpr = magic_for((uint64_t)slab_max * plane_groups, gpr);
  const unsigned long long magic_h = magic_for((uint64_t)slab_max * H, (uint64_t)H);
  cudaStream_t st = p3d::as_stream(stream);
  for (int xs = 0; xs < x_count; xs += slab_max) {
    const int xc = x_count - xs < slab_max ? x_count - xs : slab_max;
    const int64_t warp_groups = ((int64_t)xc * (int64_t)plane_groups + 31) / 32;
    // CTAs per SM over the whole grid (each warp then walks ~3-4 chunks at 512^3); measured 8: 0.069 ms, 16: 0.065,
    // 32: 0.060, 64: 0.059
    static const int waves = [] { const char* e = getenv("P3D_GFB_WAVES"); const int v = e ? atoi(e) : 64; return v > 0 ? v : 64; }();
    const int blocks = grid_for(warp_groups, 8, waves);
    uint8_t* o = out + (size_t)xs * plane_groups * 16 * (rgb ? 3 : 1);
    if (rgb) global_fold_bits_kernel<true><<<blocks, 256, 0, st>>>(W, H, D, x_begin + xs, xc, inside_bits, c, mask_bits, words_per_row, colour_hw, o, magic_gpr, magic_h);
    else global_fold_bits_kernel<false><<<blocks, 256, 0, st>>>(W, H, D, x_begin + xs, xc, inside_bits, c, mask_bits, words_per_row, colour_hw, o, magic_gpr, magic_h);
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_colourise(const uint8_t* carved, int W, int H, int D, const uint8_t* colour_hw, uint8_t* out,
                          p3d_stream_t stream) {
  const int64_t n = (int64_t)W * H * D;
  P3D_REQUIRE(W >= 0 && H >= 0 && D >= 0, "colourise: bad shape");
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(carved && colour_hw && out, "colourise: null pointer");
  colourise_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(carved, W, H, D, colour_hw, out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_part_carve_fold(const uint8_t* grid, int W, int H, int D, const int32_t* table,
                                const uint32_t* group_mask_hw, uint8_t* out, p3d_stream_t stream) {
  const int64_t n = (int64_t)W * H * D;
  P3D_REQUIRE(W > 0 && H > 0 && D > 0, "part_carve_fold: bad shape");
  P3D_REQUIRE(grid && table && group_mask_hw && out && grid != out, "part_carve_fold: null or aliased pointers");
  part_fold_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(grid, W, H, D, table, group_mask_hw, out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API size_t p3d_part_carve_bits_workspace_bytes(int W, int H, int D, int n_groups) {
  if (W <= 0 || H <= 0 || D <= 0 || n_groups < 0) return 0;
  const size_t xwp = (size_t)(W + 31) / 32 + 2;
  const size_t zbits = p3d_align_up((size_t)W * H * ((size_t)(D + 31) / 32) * 4, 256);
  return 2 * zbits + p3d_align_up((size_t)n_groups * H * xwp * 4, 256);
}

P3D_API int p3d_part_carve_fold_bits(const uint8_t* grid, int W, int H, int D, const uint32_t* inside_bits, int c, int c2,
                                     const uint32_t* group_mask_hw, int n_groups, uint8_t* out, void* workspace,
                                     size_t workspace_bytes, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && (D % 32 == 0 || D >= 16) && n_groups >= 1 && n_groups <= 32,
              "part_carve_fold_bits: bad shape (D must be a multiple of 32 or at least 16)");
  P3D_REQUIRE(grid && inside_bits && group_mask_hw && out && workspace && grid != out, "part_carve_fold_bits: null/aliased");
  P3D_REQUIRE(((reinterpret_cast<uintptr_t>(grid) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "part_carve_fold_bits: grids must be 16-byte aligned");
  if (workspace_bytes < p3d_part_carve_bits_workspace_bytes(W, H, D, n_groups)) {
    p3d::set_error("part_carve_fold_bits: workspace too small");
    return P3D_E_WORKSPACE;
  }
  const int xwp = (W + 31) / 32 + 2, words = (D + 31) / 32;
  const size_t zbits = p3d_align_up((size_t)W * H * (size_t)words * 4, 256);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  uint32_t* occz = reinterpret_cast<uint32_t*>(ws);
  uint32_t* alive = reinterpret_cast<uint32_t*>(ws + zbits);
  uint32_t* gbits = reinterpret_cast<uint32_t*>(ws + 2 * zbits);
  cudaStream_t st = p3d::as_stream(stream);
  pack_group_bits_kernel<<<grid_for((int64_t)H * xwp, 8, 32), 256, 0, st>>>(group_mask_hw, H, W, n_groups, xwp, gbits);
  if (D % 32 != 0) {                                           // ragged rows: flat groups, bit words by atomicOr
    const int64_t nvox = (int64_t)W * H * D;
    P3D_REQUIRE(nvox < (1ll << 31), "part_carve_fold_bits: grid too large for 32-bit voxel indices");
    P3D_CUDA(cudaMemsetAsync(ws, 0, 2 * zbits, st));
    const unsigned long long magic_d = magic_for((uint64_t)nvox, (uint64_t)D);
    const unsigned long long magic_hh = magic_for((uint64_t)W * H, (uint64_t)H);
    part_copy_bits_ragged_kernel<<<grid_for((nvox + 15) / 16, 256, 256), 256, 0, st>>>(
        grid, W, H, D, inside_bits, c, group_mask_hw, gbits, xwp, occz, alive, out, magic_d, magic_hh, (uint32_t)nvox);
    const int64_t rtasks = (int64_t)H * ((words + kClearZW - 1) / kClearZW) * ((W + kClearX - 1) / kClearX);
    part_clear_kernel<<<grid_for(rtasks, 1, 16), kClearX, 0, st>>>(W, H, D, c, c2, occz, alive, out, 0, W, nullptr, 1);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
  }
  const int64_t n16 = (int64_t)W * H * D / 16;
  P3D_REQUIRE(n16 < (1ll << 31), "part_carve_fold_bits: grid too large for 32-bit group indices");
  const unsigned long long magic_gpr = magic_for((uint64_t)n16, (uint64_t)D / 16);
  const unsigned long long magic_h = magic_for((uint64_t)W * H, (uint64_t)H);
  // grid: enough CTAs that every thread owns one 16-voxel group (no grid-stride loop) up to 256 CTAs per SM -- measured
  // at 512^3 for the three kernels together, CTAs per SM 8: 0.172 ms, 16: 0.162, 32: 0.146, 64: 0.140, 128: 0.137,
  // 256: 0.135; capping the residency below 8 CTAs per SM costs 10 %
  static const int pcb_waves = [] { const char* e = getenv("P3D_PCB_WAVES"); const int v = e ? atoi(e) : 256; return v > 0 ? v : 256; }();
  part_copy_bits_kernel<<<grid_for(n16, 256, pcb_waves), 256, 0, st>>>(grid, W, H, D, inside_bits, c, group_mask_hw, gbits, xwp,
                                                              occz, alive, out, magic_gpr, magic_h, 0u, (uint32_t)n16, 0u);
  const int64_t tasks = (int64_t)H * ((D / 32 + kClearZW - 1) / kClearZW) * ((W + kClearX - 1) / kClearX);
  part_clear_kernel<<<grid_for(tasks, 1, 16), kClearX, 0, st>>>(W, H, D, c, c2, occz, alive, out, 0, W, nullptr, 1);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// Output x slab [x_begin, x_begin + x_count) of the same carve from the whole (replicated) input grid: the unit of
// multi-GPU sharding.  No exchange: the slab reads its own rows (pass A) and the z range [x_begin + c2, ...) of every
// row for the source occupancy (occ_zrange_bits_kernel), 2 x 3 bytes per slab voxel in all.
P3D_API int p3d_part_carve_fold_bits_slab(const uint8_t* grid, int W, int H, int D, int x_begin, int x_count,
                                          const uint32_t* inside_bits, int c, int c2, const uint32_t* group_mask_hw,
                                          int n_groups, uint8_t* out_slab, void* workspace, size_t workspace_bytes,
                                          p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && D % 32 == 0 && n_groups >= 1 && n_groups <= 32, "part_carve_fold_bits_slab: bad shape");
  P3D_REQUIRE(x_begin >= 0 && x_count >= 0 && x_begin + x_count <= W, "part_carve_fold_bits_slab: bad x slab");
  if (x_count == 0) return P3D_OK;
  P3D_REQUIRE(grid && inside_bits && group_mask_hw && out_slab && workspace, "part_carve_fold_bits_slab: null pointer");
  P3D_REQUIRE(((reinterpret_cast<uintptr_t>(grid) | reinterpret_cast<uintptr_t>(out_slab)) & 15) == 0,
              "part_carve_fold_bits_slab: grids must be 16-byte aligned");
  if (workspace_bytes < p3d_part_carve_bits_workspace_bytes(W, H, D, n_groups)) {
    p3d::set_error("part_carve_fold_bits_slab: workspace too small");
    return P3D_E_WORKSPACE;
  }
  const int xwp = (W + 31) / 32 + 2, words = D / 32;
  const size_t zbits = p3d_align_up((size_t)W * H * (size_t)words * 4, 256);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  uint32_t* occz = reinterpret_cast<uint32_t*>(ws);
  uint32_t* alive = reinterpret_cast<uint32_t*>(ws + zbits);
  uint32_t* gbits = reinterpret_cast<uint32_t*>(ws + 2 * zbits);
  cudaStream_t st = p3d::as_stream(stream);
  pack_group_bits_kernel<<<grid_for((int64_t)H * xwp, 8, 32), 256, 0, st>>>(group_mask_hw, H, W, n_groups, xwp, gbits);
  // source occupancy: z range [x_begin + c2, x_begin + x_count + c2) clipped to the grid, whole words, every row
  int z_lo = x_begin + c2, z_hi = x_begin + x_count + c2;
  if (z_lo < 0) z_lo = 0;
  if (z_hi > D) z_hi = D;
  if (z_hi > z_lo) {
    const int w_begin = z_lo >> 5, w_count = ((z_hi + 31) >> 5) - w_begin;
    const int64_t rows = (int64_t)W * H;
    occ_zrange_bits_kernel<<<grid_for(rows * w_count, 256, 64), 256, 0, st>>>(grid, rows, D, w_begin, w_count, occz);
  }
  const int64_t n16 = (int64_t)W * H * D / 16;
  P3D_REQUIRE(n16 < (1ll << 31), "part_carve_fold_bits_slab: grid too large for 32-bit group indices");
  const unsigned long long magic_gpr = magic_for((uint64_t)n16, (uint64_t)D / 16);
  const unsigned long long magic_h = magic_for((uint64_t)W * H, (uint64_t)H);
  const int64_t gps = (int64_t)H * (D / 16);                   // groups per x plane
  const uint32_t g_begin = (uint32_t)(x_begin * gps), g_end = (uint32_t)((x_begin + x_count) * gps);
  part_copy_bits_kernel<<<grid_for((int64_t)x_count * gps, 256, 256), 256, 0, st>>>(
      grid, W, H, D, inside_bits, c, group_mask_hw, gbits, xwp, nullptr, alive, out_slab, magic_gpr, magic_h, g_begin, g_end, 0u);
  const int64_t tasks = (int64_t)H * ((words + kClearZW - 1) / kClearZW) * ((x_count + kClearX - 1) / kClearX);
  part_clear_kernel<<<grid_for(tasks, 1, 16), kClearX, 0, st>>>(W, H, D, c, c2, occz, alive, out_slab, x_begin, x_count, nullptr, 1);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// The same slab carve with a SHARDED input: every rank holds only its x slab of the grid.  Pass A writes the slab's
// output and its rows of the z-packed occupancy / alive bits; the ranks then exchange the occupancy rows (the first
// W*H*(D/32) uint32 of the workspace, [x][y][word], slab rows contiguous: one all-gather, 1/24 of the grid bytes) and pass
// B clears the runs whose rotated source is empty.
static int slab_pass_a(const uint8_t* grid_slab, int W, int H, int D, int x_begin, int x_count, const uint32_t* inside_bits,
                       int c, const uint32_t* group_mask_hw, int n_groups, uint8_t* out_slab, void* workspace,
                       size_t workspace_bytes, p3d_stream_t stream, bool pack_groups);

P3D_API int p3d_part_carve_slab_pass_a(const uint8_t* grid_slab, int W, int H, int D, int x_begin, int x_count,
                                       const uint32_t* inside_bits, int c, const uint32_t* group_mask_hw, int n_groups,
                                       uint8_t* out_slab, void* workspace, size_t workspace_bytes, p3d_stream_t stream) {
  return slab_pass_a(grid_slab, W, H, D, x_begin, x_count, inside_bits, c, group_mask_hw, n_groups, out_slab, workspace,
                     workspace_bytes, stream, true);
}

// The same pass with the per-group mask bits already in the workspace (p3d_part_carve_pack_groups): a caller that carves
// many grids with one mask and one job list packs them once.
P3D_API int p3d_part_carve_slab_pass_a_packed(const uint8_t* grid_slab, int W, int H, int D, int x_begin, int x_count,
                                              const uint32_t* inside_bits, int c, const uint32_t* group_mask_hw,
                                              int n_groups, uint8_t* out_slab, void* workspace, size_t workspace_bytes,
                                              p3d_stream_t stream) {
  return slab_pass_a(grid_slab, W, H, D, x_begin, x_count, inside_bits, c, group_mask_hw, n_groups, out_slab, workspace,
                     workspace_bytes, stream, false);
}

P3D_API int p3d_part_carve_pack_groups(const uint32_t* group_mask_hw, int W, int H, int D, int n_groups, void* workspace,
                                       size_t workspace_bytes, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && n_groups >= 1 && n_groups <= 32, "part_carve_pack_groups: bad shape");
  P3D_REQUIRE(group_mask_hw && workspace, "part_carve_pack_groups: null pointer");
  if (workspace_bytes < p3d_part_carve_bits_workspace_bytes(W, H, D, n_groups)) {
    p3d::set_error("part_carve_pack_groups: workspace too small");
    return P3D_E_WORKSPACE;
  }
  const int xwp = (W + 31) / 32 + 2;
  const size_t zbits = p3d_align_up((size_t)W * H * (size_t)((D + 31) / 32) * 4, 256);
  uint32_t* gbits = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(workspace) + 2 * zbits);
  pack_group_bits_kernel<<<grid_for((int64_t)H * xwp, 8, 32), 256, 0, p3d::as_stream(stream)>>>(group_mask_hw, H, W, n_groups,
                                                                                              xwp, gbits);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

static int slab_pass_a(const uint8_t* grid_slab, int W, int H, int D, int x_begin, int x_count, const uint32_t* inside_bits,
                       int c, const uint32_t* group_mask_hw, int n_groups, uint8_t* out_slab, void* workspace,
                       size_t workspace_bytes, p3d_stream_t stream, bool pack_groups) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && D % 32 == 0 && n_groups >= 1 && n_groups <= 32, "part_carve_slab_pass_a: bad shape");
  P3D_REQUIRE(x_begin >= 0 && x_count >= 0 && x_begin + x_count <= W, "part_carve_slab_pass_a: bad x slab");
  if (x_count == 0) return P3D_OK;
  P3D_REQUIRE(grid_slab && inside_bits && group_mask_hw && out_slab && workspace && grid_slab != out_slab,
              "part_carve_slab_pass_a: null/aliased");
  P3D_REQUIRE(((reinterpret_cast<uintptr_t>(grid_slab) | reinterpret_cast<uintptr_t>(out_slab)) & 15) == 0,
              "part_carve_slab_pass_a: grids must be 16-byte aligned");
  if (workspace_bytes < p3d_part_carve_bits_workspace_bytes(W, H, D, n_groups)) {
    p3d::set_error("part_carve_slab_pass_a: workspace too small");
    return P3D_E_WORKSPACE;
  }
  const int xwp = (W + 31) / 32 + 2;
  const size_t zbits = p3d_align_up((size_t)W * H * (size_t)(D / 32) * 4, 256);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  uint32_t* occz = reinterpret_cast<uint32_t*>(ws);
  uint32_t* alive = reinterpret_cast<uint32_t*>(ws + zbits);
  uint32_t* gbits = reinterpret_cast<uint32_t*>(ws + 2 * zbits);
  cudaStream_t st = p3d::as_stream(stream);
  if (pack_groups)
    pack_group_bits_kernel<<<grid_for((int64_t)H * xwp, 8, 32), 256, 0, st>>>(group_mask_hw, H, W, n_groups, xwp, gbits);
  const int64_t n16 = (int64_t)W * H * D / 16;
  P3D_REQUIRE(n16 < (1ll << 31), "part_carve_slab_pass_a: grid too large for 32-bit group indices");
  const unsigned long long magic_gpr = magic_for((uint64_t)n16, (uint64_t)D / 16);
  const unsigned long long magic_h = magic_for((uint64_t)W * H, (uint64_t)H);
  const int64_t gps = (int64_t)H * (D / 16);
  const uint32_t g_begin = (uint32_t)(x_begin * gps), g_end = (uint32_t)((x_begin + x_count) * gps);
  part_copy_bits_kernel<<<grid_for((int64_t)x_count * gps, 256, 256), 256, 0, st>>>(
      grid_slab, W, H, D, inside_bits, c, group_mask_hw, gbits, xwp, occz, alive, out_slab, magic_gpr, magic_h, g_begin, g_end,
      g_begin);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_part_carve_slab_pass_b(int W, int H, int D, int x_begin, int x_count, int c, int c2, uint8_t* out_slab,
                                       void* workspace, size_t workspace_bytes, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && D % 32 == 0, "part_carve_slab_pass_b: bad shape");
  P3D_REQUIRE(x_begin >= 0 && x_count >= 0 && x_begin + x_count <= W, "part_carve_slab_pass_b: bad x slab");
  if (x_count == 0) return P3D_OK;
  P3D_REQUIRE(out_slab && workspace, "part_carve_slab_pass_b: null pointer");
  const int words = D / 32;
  const size_t zbits = p3d_align_up((size_t)W * H * (size_t)words * 4, 256);
  if (workspace_bytes < 2 * zbits) {
    p3d::set_error("part_carve_slab_pass_b: workspace too small");
    return P3D_E_WORKSPACE;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const uint32_t* occz = reinterpret_cast<const uint32_t*>(ws);
  const uint32_t* alive = reinterpret_cast<const uint32_t*>(ws + zbits);
  const int64_t tasks = (int64_t)H * ((words + kClearZW - 1) / kClearZW) * ((x_count + kClearX - 1) / kClearX);
  part_clear_kernel<<<grid_for(tasks, 1, 16), kClearX, 0, p3d::as_stream(stream)>>>(W, H, D, c, c2, occz, alive, out_slab,
                                                                                  x_begin, x_count, nullptr, 1);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_part_carve_slab_pass_b_peers(int W, int H, int D, int x_begin, int x_count, int c, int c2,
                                             uint8_t* out_slab, void* workspace, size_t workspace_bytes,
                                             const void* const* peer_workspaces, int n_ranks, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && D % 32 == 0, "part_carve_slab_pass_b_peers: bad shape");
  P3D_REQUIRE(x_begin >= 0 && x_count >= 0 && x_begin + x_count <= W, "part_carve_slab_pass_b_peers: bad x slab");
  P3D_REQUIRE(n_ranks >= 1 && W % n_ranks == 0, "part_carve_slab_pass_b_peers: %d rows do not divide over %d ranks", W, n_ranks);
  if (x_count == 0) return P3D_OK;
  P3D_REQUIRE(out_slab && workspace && peer_workspaces, "part_carve_slab_pass_b_peers: null pointer");
  const int words = D / 32;
  const size_t zbits = p3d_align_up((size_t)W * H * (size_t)words * 4, 256);
  if (workspace_bytes < 2 * zbits) {
    p3d::set_error("part_carve_slab_pass_b_peers: workspace too small");
    return P3D_E_WORKSPACE;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const uint32_t* occz = reinterpret_cast<const uint32_t*>(ws);
  const uint32_t* alive = reinterpret_cast<const uint32_t*>(ws + zbits);
  const int64_t tasks = (int64_t)H * ((words + kClearZW - 1) / kClearZW) * ((x_count + kClearX - 1) / kClearX);
  part_clear_kernel<<<grid_for(tasks, 1, 16), kClearX, 0, p3d::as_stream(stream)>>>(
      W, H, D, c, c2, occz, alive, out_slab, x_begin, x_count, reinterpret_cast<const uint32_t* const*>(peer_workspaces),
      W / n_ranks);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_crop_occupancy(const uint8_t* grid, int W, int H, int D, int x0, int y0, int z0, int w, int h, int d,
                               const uint8_t* sel_wh, uint8_t* occ, p3d_stream_t stream) {
  P3D_REQUIRE(x0 >= 0 && y0 >= 0 && z0 >= 0 && w >= 0 && h >= 0 && d >= 0 && x0 + w <= W && y0 + h <= H && z0 + d <= D,
              "crop_occupancy: crop outside the grid");
  const int64_t n = (int64_t)w * h * d;
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(grid && occ, "crop_occupancy: null pointer");
  crop_occupancy_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(grid, H, D, x0, y0, z0, w, h, d, sel_wh, occ);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_accumulate_part(const uint8_t* grid, const uint8_t* carved, int W, int H, int D, const uint8_t* sel_wh,
                                uint8_t* final_grid, p3d_stream_t stream) {
  const int64_t n = (int64_t)W * H * D;
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(grid && carved && sel_wh && final_grid, "accumulate_part: null pointer");
  accumulate_part_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(grid, carved, W, H, D, sel_wh, final_grid);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_paste_component(const uint8_t* src, const int32_t* labels, int comp, const uint8_t* kept, int W, int H,
                                int D, int x0, int y0, int z0, int w, int h, int d, uint8_t* out, p3d_stream_t stream) {
  P3D_REQUIRE(x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + w <= W && y0 + h <= H && z0 + d <= D, "paste_component: bad crop");
  const int64_t n = (int64_t)w * h * d;
  if (n <= 0) return P3D_OK;
  P3D_REQUIRE(src && labels && kept && out, "paste_component: null pointer");
  paste_component_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(src, labels, comp, kept, H, D, x0, y0,
                                                                                 z0, w, h, d, out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_lr_carve_components(const uint8_t* grid, const int32_t* labels, int W, int H, int D,
                                    const uint8_t* mask_hw, const int32_t* comps, int n_comp, int64_t max_crop_voxels,
                                    const double* Ms, const double* offs, int n_pass, uint8_t* buf_a, uint8_t* buf_b,
                                    int sequential_paste, uint8_t* out, int64_t* counts, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && n_comp >= 0 && n_pass >= 0 && max_crop_voxels >= 0, "lr_carve_components: bad arguments");
  if (n_comp == 0) return P3D_OK;
  P3D_REQUIRE(grid && labels && mask_hw && comps && buf_a && buf_b && buf_a != buf_b && out && counts &&
              (n_pass == 0 || (Ms && offs)), "lr_carve_components: null or aliased pointer");
  P3D_REQUIRE(n_comp <= 65535, "lr_carve_components: %d components exceed one launch", n_comp);
  cudaStream_t st = p3d::as_stream(stream);
  P3D_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_comp * sizeof(int64_t), st));
  int per = grid_for(max_crop_voxels, 256, 16);
  const int cap = (p3d::sm_count() * 16 + n_comp - 1) / n_comp;         // about 16 CTAs per SM over all components
  if (per > cap) per = cap < 1 ? 1 : cap;
  dim3 g((unsigned)per, (unsigned)n_comp);
  lr_crop_kernel<<<g, 256, 0, st>>>(grid, H, D, comps, buf_a);
  uint8_t* src = buf_a;
  uint8_t* dst = buf_b;
  for (int p = 0; p < n_pass; ++p) {
    lr_resample_kernel<<<g, 256, 0, st>>>(src, dst, comps, Ms, offs, p, n_pass, mask_hw, W);
    uint8_t* t = src; src = dst; dst = t;
  }
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts);
  if (sequential_paste) {
    for (int c = 0; c < n_comp; ++c) lr_paste_kernel<<<dim3((unsigned)per, 1), 256, 0, st>>>(grid, labels, src, comps, c, H, D, out, cnt);
  } else {
    lr_paste_kernel<<<g, 256, 0, st>>>(grid, labels, src, comps, -1, H, D, out, cnt);
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_group_image(const uint8_t* mask_rgb, int H, int W, const uint32_t* keys, const int32_t* group_of, int n_keys,
                            uint32_t* gm, p3d_stream_t stream) {
  P3D_REQUIRE(H >= 0 && W >= 0 && n_keys >= 0 && n_keys <= 128, "group_image: bad arguments (at most 128 colours)");
  if ((int64_t)H * W == 0) return P3D_OK;
  P3D_REQUIRE((int64_t)H * W < (1ll << 31), "group_image: image too large");
  P3D_REQUIRE(mask_rgb && gm && (n_keys == 0 || (keys && group_of)), "group_image: null pointer");
  group_image_kernel<<<grid_for((int64_t)H * W, 256, 8), 256, 0, p3d::as_stream(stream)>>>(mask_rgb, H, W, keys, group_of,
                                                                                          n_keys, gm);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_colour_mask(const uint8_t* grid_rgb, int64_t n, int r, int g, int b, uint8_t* mask, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0, "colour_mask: n < 0");
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(grid_rgb && mask, "colour_mask: null pointer");
  const bool representable = r >= 0 && r < 256 && g >= 0 && g < 256 && b >= 0 && b < 256;
  const uint32_t colour = representable ? (uint32_t)(r | (g << 8) | (b << 16)) : 0xffffffffu;
  colour_mask_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(grid_rgb, n, colour, mask);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API size_t p3d_label6_workspace_bytes(int64_t n) {
  if (n <= 0) return 0;
  const int64_t tiles = (n + kRankTile - 1) / kRankTile;
  return p3d_align_up((size_t)n * 4, 256) + p3d_align_up((size_t)n, 256) + p3d_align_up((size_t)(tiles + 1) * 4, 256);
}

static int label_components(const uint8_t* mask, int n0, int n1, int n2, int conn /* 6, 26 (3-D) or 8 (2-D) */, int32_t* labels,
                            int32_t* n_components, void* workspace, size_t workspace_bytes, p3d_stream_t stream) {
  P3D_REQUIRE(n0 >= 0 && n1 >= 0 && n2 >= 0 && n_components, "label6: bad arguments");
  const int64_t n = (int64_t)n0 * n1 * n2;
  cudaStream_t st = p3d::as_stream(stream);
  if (n == 0) { P3D_CUDA(cudaMemsetAsync(n_components, 0, 4, st)); return P3D_OK; }
  P3D_REQUIRE(n < (1ll << 31), "label6: volume too large for 32-bit labels");
  P3D_REQUIRE(mask && labels && workspace, "label6: null pointer");
  if (workspace_bytes < p3d_label6_workspace_bytes(n)) {
    p3d::set_error("label6: workspace %zu < %zu", workspace_bytes, p3d_label6_workspace_bytes(n));
    return P3D_E_WORKSPACE;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  int32_t* rank_of = reinterpret_cast<int32_t*>(ws);
  uint8_t* is_root = ws + p3d_align_up((size_t)n * 4, 256);
  int32_t* tiles_buf = reinterpret_cast<int32_t*>(is_root + p3d_align_up((size_t)n, 256));
  const int tiles = (int)((n + kRankTile - 1) / kRankTile);
  const int blocks = grid_for(n, 256, 16);
  int32_t* parent = labels;                              // labels doubles as the union-find forest
  ccl_init_kernel<<<blocks, 256, 0, st>>>(mask, n, n2, parent);
  if (conn == 8) ccl_merge8_kernel<<<blocks, 256, 0, st>>>(mask, n1, n2, parent);
  else if (conn == 26) ccl_merge26_kernel<<<blocks, 256, 0, st>>>(mask, n0, n1, n2, parent);
  else ccl_merge_kernel<<<blocks, 256, 0, st>>>(mask, n0, n1, n2, parent);
  ccl_flatten_kernel<<<blocks, 256, 0, st>>>(n, parent, is_root);
  root_count_kernel<<<tiles, kRankThreads, 0, st>>>(is_root, n, tiles_buf);
  root_scan_kernel<<<1, 1024, 0, st>>>(tiles_buf, tiles, n_components);
  root_rank_kernel<<<tiles, kRankThreads, 0, st>>>(is_root, n, tiles_buf, rank_of);
  // labels are rewritten from the forest: out of place via rank_of, then in place
  ccl_relabel_kernel<<<blocks, 256, 0, st>>>(parent, rank_of, n, labels);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_label6(const uint8_t* mask, int n0, int n1, int n2, int32_t* labels, int32_t* n_components,
                       void* workspace, size_t workspace_bytes, p3d_stream_t stream) {
  return label_components(mask, n0, n1, n2, 6, labels, n_components, workspace, workspace_bytes, stream);
}

P3D_API int p3d_label26(const uint8_t* mask, int n0, int n1, int n2, int32_t* labels, int32_t* n_components,
                        void* workspace, size_t workspace_bytes, p3d_stream_t stream) {
  return label_components(mask, n0, n1, n2, 26, labels, n_components, workspace, workspace_bytes, stream);
}

P3D_API int p3d_label8_2d(const uint8_t* mask, int H, int W, int32_t* labels, int32_t* n_components, void* workspace,
                          size_t workspace_bytes, p3d_stream_t stream) {
  return label_components(mask, 1, H, W, 8, labels, n_components, workspace, workspace_bytes, stream);
}

P3D_API int p3d_component_stats(const int32_t* labels, int n0, int n1, int n2, int n_components, int32_t* bbox,
                                int64_t* sums, p3d_stream_t stream) {
  P3D_REQUIRE(n_components >= 0, "component_stats: n_components < 0");
  const int64_t n = (int64_t)n0 * n1 * n2;
  if (n_components == 0 || n == 0) return P3D_OK;
  P3D_REQUIRE(labels && bbox && sums, "component_stats: null pointer");
  cudaStream_t st = p3d::as_stream(stream);
  bbox_init_kernel<<<(n_components * 6 + 255) / 256, 256, 0, st>>>(bbox, n_components);
  P3D_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_components * 4 * sizeof(int64_t), st));
  component_stats_kernel<<<grid_for(n, 256, 16), 256, 0, st>>>(labels, n0, n1, n2, n_components, bbox,
                                                             reinterpret_cast<unsigned long long*>(sums));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_component_extremes(const int32_t* labels, int n0, int n1, int n2, int n_components, int axis,
                                   const int32_t* bbox, int64_t* sums, p3d_stream_t stream) {
  P3D_REQUIRE(n_components >= 0 && axis >= 0 && axis < 3, "component_extremes: bad arguments");
  const int64_t n = (int64_t)n0 * n1 * n2;
  if (n_components == 0 || n == 0) return P3D_OK;
  P3D_REQUIRE(labels && bbox && sums, "component_extremes: null pointer");
  cudaStream_t st = p3d::as_stream(stream);
  P3D_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_components * 8 * sizeof(int64_t), st));
  component_extremes_kernel<<<grid_for(n, 256, 16), 256, 0, st>>>(labels, n0, n1, n2, axis, bbox,
                                                                reinterpret_cast<unsigned long long*>(sums));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_label_equals(const int32_t* labels, int64_t n, int32_t id, uint8_t* mask, p3d_stream_t stream) {
  if (n <= 0) return P3D_OK;
  P3D_REQUIRE(labels && mask, "label_equals: null pointer");
  label_equals_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(labels, n, id, mask);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_coords_extremes(const int32_t* coords, int64_t n, int axis, int32_t* minmax, int64_t* sums,
                                p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && axis >= 0 && axis < 3 && minmax && sums, "coords_extremes: bad arguments");
  cudaStream_t st = p3d::as_stream(stream);
  coords_mm_init_kernel<<<1, 32, 0, st>>>(minmax);
  P3D_CUDA(cudaMemsetAsync(sums, 0, 8 * sizeof(int64_t), st));
  if (n > 0) {
    P3D_REQUIRE(coords, "coords_extremes: null coordinates");
    coords_minmax_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(coords, n, axis, minmax);
    coords_extreme_sums_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(coords, n, axis, minmax,
                                                                   reinterpret_cast<long long*>(sums));
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_recolour_components(const int32_t* labels, const uint8_t* recolour, int64_t n, int r, int g, int b,
                                    uint8_t* grid_rgb, p3d_stream_t stream) {
  if (n <= 0) return P3D_OK;
  P3D_REQUIRE(labels && recolour && grid_rgb, "recolour_components: null pointer");
  const uint32_t colour = (uint32_t)((r & 0xff) | ((g & 0xff) << 8) | ((b & 0xff) << 16));
  recolour_kernel<<<grid_for(n, 256, 16), 256, 0, p3d::as_stream(stream)>>>(labels, recolour, n, colour, grid_rgb);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_extrude(uint8_t* grid_rgb, int W, int H, int D, const uint8_t* mask_hw, int mask_h, int mask_w, int axis,
                        int sign, int depth, int r, int g, int b, p3d_stream_t stream) {
  P3D_REQUIRE(W > 0 && H > 0 && D > 0 && (axis == 0 || axis == 2) && (sign == 1 || sign == -1) && depth >= 0,
              "extrude: bad arguments");
  P3D_REQUIRE(grid_rgb && mask_hw, "extrude: null pointer");
  if (axis == 2) P3D_REQUIRE(mask_h == H && mask_w == W, "extrude: mask (%d,%d) does not match (H,W)=(%d,%d)", mask_h, mask_w, H, W);
  else P3D_REQUIRE(mask_h == H && mask_w == D, "extrude: mask (%d,%d) does not match (H,D)=(%d,%d)", mask_h, mask_w, H, D);
  const int ncol = axis == 2 ? W * H : H * D;
  const uint32_t colour = (uint32_t)((r & 0xff) | ((g & 0xff) << 8) | ((b & 0xff) << 16));
  if (axis == 2)
    extrude_z_kernel<<<grid_for(ncol, 8, 16), 256, 0, p3d::as_stream(stream)>>>(grid_rgb, W, H, D, mask_hw, sign, depth, colour);
  else
    extrude_kernel<<<(ncol + 127) / 128, 128, 0, p3d::as_stream(stream)>>>(grid_rgb, W, H, D, mask_hw, mask_w, axis, sign,
                                                                        depth, colour);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_reorient(const uint8_t* in, int W, int H, int D, uint8_t* out, p3d_stream_t stream) {
  P3D_REQUIRE(W >= 0 && H >= 0 && D >= 0, "reorient: bad shape");
  if ((int64_t)W * H * D == 0) return P3D_OK;
  P3D_REQUIRE(in && out && in != out, "reorient: null or aliased pointers");
  P3D_REQUIRE(H <= 65535, "reorient: H too large");
  dim3 grid((unsigned)(((W + 31) / 32) * ((D + 31) / 32)), (unsigned)H);
  const bool vec = W % 32 == 0 && D % 32 == 0 &&
                   ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) reorient_vec_kernel<<<grid, 192, 0, p3d::as_stream(stream)>>>(in, W, H, D, out);
  else reorient_kernel<<<grid, 256, 0, p3d::as_stream(stream)>>>(in, W, H, D, out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}
