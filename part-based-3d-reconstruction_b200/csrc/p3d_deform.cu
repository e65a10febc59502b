// libp3d_b200: part-wise deformation with a fixed camera (stage 3 of the reference).
//
// Reference call sites replaced here (paths under the reference root):
//   deform_coords                utils/deformation_estimation.py:70-103   (7 jittered copies, scale/shift about the
//                                                                          copy's own mean, round half-even)
//   save_params / update         utils/deformation_estimation.py:105-145, 263-284  (bounds test, projection, one-part IoU)
//   save_deformed_grid           utils/deformation_estimation.py:288-311
//   run_auto_align (commented)   utils/deformation_estimation.py:148-258  (grid of deformations -> batched here)
//
// Exactness: the deformation is FP64 in the reference (coords float32 + float64 offsets -> float64), with separate
// multiply and add ufuncs (no contraction):
//   c   = (p + off) - centre                      centre = mean over the (sub-sampled) part of p + off
//   x'  = c0*scale_xz + (shift_xz*pix2vox_x)*sign(c0)      y' = c1*scale_y - shift_y*pix2vox_y      z' like x'
//   out = rint(x' + centre)  (half-even)
// Voxel coordinates are integers, so sum(p + off) is exact in any order and the mean is ONE correctly rounded division
// of an exact sum; the kernels take the exact int64 coordinate sums and reproduce every later operation with explicit
// round-to-nearest intrinsics.  The IoU of a single-colour part does not depend on duplicate or re-ordered points, so
// the sweep never sorts: it sets bits in a per-candidate coverage bitmap.
#include "p3d_common.cuh"
#include "p3d_project.cuh"

namespace {

constexpr int kJitters = 7;
// deformation_estimation.py:87-92: jitter j moves axis kJitAxis[j] by kJitSign[j] * 0.25
__constant__ int c_jit_axis[kJitters] = {-1, 0, 0, 1, 1, 2, 2};
__constant__ double c_jit_off[kJitters] = {0.0, 0.25, -0.25, 0.25, -0.25, 0.25, -0.25};

// ---------------------------------------------------------------------------------------------
// exact coordinate sums of the sub-sampled part: sums[k] = sum_i p[i*stride][k], sums[3] = number of coordinates
// that are not integers below 2^24 (the deformation kernels refuse such inputs).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) deform_sums_kernel(const float* __restrict__ pts, int64_t m, int64_t stride,
                                                          unsigned long long* __restrict__ sums) {
  long long s[3] = {0, 0, 0};
  unsigned long long bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float v = __ldg(pts + 3 * i * stride + k);
      const bool ok = fabsf(v) < 16777216.f && rintf(v) == v;
      if (ok) s[k] += (long long)v; else ++bad;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    for (int d = 16; d > 0; d >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], d);
  for (int d = 16; d > 0; d >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, d);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(sums + k, (unsigned long long)s[k]);
    if (bad) atomicAdd(sums + 3, bad);
  }
}

// centres[k][0..2] = mean of axis k for jitter offset 0, +0.25, -0.25:  fl((S_k + m*off) / m)
__global__ void deform_centres_kernel(const unsigned long long* __restrict__ sums, int64_t m, double* __restrict__ centres) {
  const int t = threadIdx.x;
  if (t >= 9) return;
  const int k = t / 3, o = t % 3;
  const double off = o == 0 ? 0.0 : (o == 1 ? 0.25 : -0.25);
  const double S = (double)(long long)sums[k];
  const double dm = (double)m;
  centres[t] = __ddiv_rn(__dadd_rn(S, __dmul_rn(dm, off)), dm);          // both inner operations are exact
}

struct DeformParams {            // one candidate, pre-multiplied shifts (Python evaluates shift*pix2vox first)
  double sy, sxz, ky, kx, kz;
};

__device__ __forceinline__ DeformParams load_deform(const double* __restrict__ d, const double* __restrict__ pix2vox) {
  DeformParams q;
  q.sy = d[0];
  q.sxz = d[2];
  q.ky = __dmul_rn(d[1], pix2vox[1]);
  q.kx = __dmul_rn(d[3], pix2vox[0]);
  q.kz = __dmul_rn(d[3], pix2vox[2]);
  return q;
}

__device__ __forceinline__ double np_sign(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : v); }   // NaN / 0 -> itself

// One coordinate of one jittered copy: axis k (0 = x, 1 = y, 2 = z) of a point whose coordinate is v, jitter slot o
// (0: offset 0, 1: +0.25, 2: -0.25).  A jitter only moves ONE axis, so the other two coordinates of that copy equal
// those of the un-jittered copy -- the sweep evaluates 3 + 6 axis values per point instead of 7 x 3.
__device__ __forceinline__ double deform_axis(double v, int k, int o, const double* __restrict__ centres,
                                              const DeformParams& q) {
  const double off = o == 0 ? 0.0 : (o == 1 ? 0.25 : -0.25);
  const double ctr = centres[3 * k + o];
  const double c = __dsub_rn(__dadd_rn(v, off), ctr);                     // v + off is exact
  double t;
  if (k == 1) t = __dsub_rn(__dmul_rn(c, q.sy), q.ky);
  else t = __dadd_rn(__dmul_rn(c, q.sxz), __dmul_rn(k == 0 ? q.kx : q.kz, np_sign(c)));
  return rint(__dadd_rn(t, ctr));
}

// One jittered copy of one point -> deformed coordinates as doubles holding integers (before the int cast).
__device__ __forceinline__ void deform_one(const double p[3], int j, const double* __restrict__ centres,
                                           const DeformParams& q, double out[3]) {
  double c[3], ctr[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const bool mine = c_jit_axis[j] == k;
    const double off = mine ? c_jit_off[j] : 0.0;
    ctr[k] = centres[3 * k + (mine ? (c_jit_off[j] > 0.0 ? 1 : 2) : 0)];
    c[k] = __dsub_rn(__dadd_rn(p[k], off), ctr[k]);                      // p + off is exact
  }
  const double x = __dadd_rn(__dmul_rn(c[0], q.sxz), __dmul_rn(q.kx, np_sign(c[0])));
  const double y = __dsub_rn(__dmul_rn(c[1], q.sy), q.ky);
  const double z = __dadd_rn(__dmul_rn(c[2], q.sxz), __dmul_rn(q.kz, np_sign(c[2])));
  out[0] = rint(__dadd_rn(x, ctr[0]));
  out[1] = rint(__dadd_rn(y, ctr[1]));
  out[2] = rint(__dadd_rn(z, ctr[2]));
}

__device__ __forceinline__ long long to_int64_np(double v) {              // NumPy's float64 -> int64 cast on x86-64
  if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return (long long)0x8000000000000000ull;
  return (long long)v;
}

// deform_coords before np.unique: out (7, m, 3) int64, jitter-major like np.vstack.
__global__ void __launch_bounds__(256) deform_points_kernel(const float* __restrict__ pts, int64_t m, int64_t stride,
                                                            const double* __restrict__ centres,
                                                            const double* __restrict__ deform,
                                                            const double* __restrict__ pix2vox,
                                                            long long* __restrict__ out) {
  const DeformParams q = load_deform(deform, pix2vox);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = (double)__ldg(pts + 3 * i * stride + k);
    for (int j = 0; j < kJitters; ++j) {
      double o[3];
      deform_one(p, j, centres, q, o);
      long long* dst = out + ((int64_t)j * m + i) * 3;
      dst[0] = to_int64_np(o[0]); dst[1] = to_int64_np(o[1]); dst[2] = to_int64_np(o[2]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Sweep: candidates x points x jitters -> coverage bits of the deformed part seen through ONE camera.
// grid = (point tiles, candidate groups).  cov: (D, words) uint32, bit p of a candidate = pixel p covered.
// nvalid[d] = number of (point, jitter) pairs inside the grid (before de-duplication).
// ---------------------------------------------------------------------------------------------
constexpr int kDeformThreads = 128;
constexpr int kDeformPerBlock = 8;       // candidates handled by one CTA

// T = working dtype of the reference's projection: double for float64 camera arrays (notebook 2 JSON -> float64),
// float for float32 camera arrays (notebook 3 converts them with to_numpy(dtype=float32)).
template <typename T, bool kFilter>
__global__ void __launch_bounds__(kDeformThreads) deform_splat_kernel(
    const float* __restrict__ pts, int64_t m, int64_t stride, const double* __restrict__ centres,
    const double* __restrict__ deforms, int D, const double* __restrict__ pix2vox, int A0, int A1, int A2,
    const T* __restrict__ cam, const float* __restrict__ fast, const float* __restrict__ bbox, int H, int W,
    uint32_t* __restrict__ cov, int64_t words, unsigned long long* __restrict__ nvalid) {
  __shared__ T s_cam[16];
  __shared__ double s_ctr[9];
  __shared__ FastCam s_fast;
  __shared__ DeformParams s_def[kDeformPerBlock];
  if (threadIdx.x < 16) s_cam[threadIdx.x] = cam[threadIdx.x];
  if (threadIdx.x < 9) s_ctr[threadIdx.x] = centres[threadIdx.x];
  if (kFilter && threadIdx.x < 16) reinterpret_cast<float*>(&s_fast)[threadIdx.x] = fast[threadIdx.x];
  const int d0 = blockIdx.y * kDeformPerBlock;
  const int nd = min(kDeformPerBlock, D - d0);
  if (threadIdx.x < nd) s_def[threadIdx.x] = load_deform(deforms + (size_t)(d0 + threadIdx.x) * 4, pix2vox);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kDeformThreads + threadIdx.x;
  const bool live = i < m;
  double p[3] = {0.0, 0.0, 0.0};
  if (live)
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = (double)__ldg(pts + 3 * i * stride + k);
  float ctr0 = 0.f, ctr1 = 0.f, ctr2 = 0.f;
  FastCam fc;
  if (kFilter) {
    ctr0 = bbox_centre(bbox, 0); ctr1 = bbox_centre(bbox, 1); ctr2 = bbox_centre(bbox, 2);
    fc = s_fast;
  }
  const T tW = (T)W, tH = (T)H;
  const double fA0 = (double)A0, fA1 = (double)A1, fA2 = (double)A2;
  const float kMagic = 12582912.f;
  for (int dd = 0; dd < nd; ++dd) {
    const DeformParams q = s_def[dd];
    uint32_t* cv = cov + (size_t)(d0 + dd) * words;
    int count = 0;
    if (live) {
      // splat one deformed voxel (already known to lie inside the grid)
      auto splat_voxel = [&](double ox, double oy, double oz) {
        bool hit, decided = false;
        uint32_t pix = 0;
        if (kFilter) {   // FP32 filter (same arithmetic and thresholds as splat_filtered_kernel)
          const float qx = __fsub_rn((float)ox, ctr0), qy = __fsub_rn((float)oy, ctr1), qz = __fsub_rn((float)oz, ctr2);
          const float X = __fmaf_rn(qz, fc.A[2], __fmaf_rn(qy, fc.A[1], __fmaf_rn(qx, fc.A[0], fc.TA)));
          const float Y = __fmaf_rn(qz, fc.B[2], __fmaf_rn(qy, fc.B[1], __fmaf_rn(qx, fc.B[0], fc.TB)));
          const float Z = __fmaf_rn(qz, fc.C[2], __fmaf_rn(qy, fc.C[1], __fmaf_rn(qx, fc.C[0], fc.TC)));
          float r;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(Z));
          const float u = __fmaf_rn(X, r, fc.cx), v = __fmaf_rn(Y, r, fc.cy);
          const float su = __fadd_rn(u, kMagic), sv = __fadd_rn(v, kMagic);
          decided = fabsf(__fsub_rn(u, __fsub_rn(su, kMagic))) < fc.thr_u &&
                    fabsf(__fsub_rn(v, __fsub_rn(sv, kMagic))) < fc.thr_v;
          const uint32_t iu = (uint32_t)(__float_as_int(su) - 0x4B400000), iv = (uint32_t)(__float_as_int(sv) - 0x4B400000);
          hit = decided && iu < (uint32_t)W && iv < (uint32_t)H;
          pix = iv * (uint32_t)W + iu;
        }
        if (!decided) hit = exact_pixel<T>((T)ox, (T)oy, (T)oz, s_cam, W, tW, tH, pix);   // the reference's sequence
        if (hit) {
          const uint32_t bit = 1u << (pix & 31u);
          uint32_t* w = cv + (pix >> 5);
          if ((__ldcg(w) & bit) == 0u) atomicOr(w, bit);
        }
      };
      // bounds in the grid: x < A2, y < A1, z < A0 (deformation_estimation.py:111-115); NaN fails every test
      const double lim[3] = {fA2, fA1, fA0};
      double o0[3];
      bool in0[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        o0[k] = deform_axis(p[k], k, 0, s_ctr, q);
        in0[k] = o0[k] >= 0.0 && o0[k] < lim[k];
      }
      if (in0[0] && in0[1] && in0[2]) {                        // jitter 0
        ++count;
        splat_voxel(o0[0], o0[1], o0[2]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {                            // jitters 1..6: axis k moved by +0.25 / -0.25
        const bool others = in0[(k + 1) % 3] && in0[(k + 2) % 3];
        if (!others) continue;
        double prev = o0[k];
        bool prev_done = in0[k];                               // a voxel equal to an already splatted one adds no pixel
#pragma unroll
        for (int o = 1; o <= 2; ++o) {
          const double v = deform_axis(p[k], k, o, s_ctr, q);
          if (!(v >= 0.0 && v < lim[k])) continue;
          ++count;
          if ((v == o0[k] && in0[k]) || (v == prev && prev_done)) continue;
          double c3[3] = {o0[0], o0[1], o0[2]};
          c3[k] = v;
          splat_voxel(c3[0], c3[1], c3[2]);
          prev = v;
          prev_done = true;
        }
      }
    }
    for (int s = 16; s > 0; s >>= 1) count += __shfl_xor_sync(0xffffffffu, count, s);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd(nvalid + d0 + dd, (unsigned long long)count);
  }
}

// counts[d] = (|cov & gt|, |cov | gt|); the coverage words are cleared on the way out.
__global__ void __launch_bounds__(256) deform_score_kernel(uint32_t* __restrict__ cov, const uint32_t* __restrict__ gt_bits,
                                                           int64_t words, long long* __restrict__ counts) {
  const int d = blockIdx.y;
  uint32_t* cv = cov + (size_t)d * words;
  unsigned int inter = 0, uni = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t c = cv[w], g = __ldg(gt_bits + w);
    if (c) cv[w] = 0u;
    inter += __popc(c & g);
    uni += __popc(c | g);
  }
  for (int s = 16; s > 0; s >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, s);
    uni += __shfl_xor_sync(0xffffffffu, uni, s);
  }
  if ((threadIdx.x & 31) == 0) {
    if (inter) atomicAdd(reinterpret_cast<unsigned long long*>(counts + 2 * d), (unsigned long long)inter);
    if (uni) atomicAdd(reinterpret_cast<unsigned long long*>(counts + 2 * d + 1), (unsigned long long)uni);
  }
}

// gt_bits[w] bit b = (label image pixel 32w+b == label)
__global__ void __launch_bounds__(256) pack_label_bits_kernel(const uint8_t* __restrict__ labels, int64_t n, int label,
                                                              uint32_t* __restrict__ bits) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = p < n && labels[p] == (uint8_t)label;
  const uint32_t m = __ballot_sync(0xffffffffu, on);
  if ((threadIdx.x & 31) == 0 && (p >> 5) < ((n + 31) >> 5)) bits[p >> 5] = m;
}

// save_deformed_grid for one part: grid[z][y][x] = colour at every valid deformed coordinate.
__global__ void __launch_bounds__(256) deform_scatter_kernel(const float* __restrict__ pts, int64_t m, int64_t stride,
                                                             const double* __restrict__ centres,
                                                             const double* __restrict__ deform,
                                                             const double* __restrict__ pix2vox, int A0, int A1, int A2,
                                                             int r, int g, int b, uint8_t* __restrict__ grid,
                                                             unsigned long long* __restrict__ nvalid) {
  const DeformParams q = load_deform(deform, pix2vox);
  const double fA0 = (double)A0, fA1 = (double)A1, fA2 = (double)A2;
  int count = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = (double)__ldg(pts + 3 * i * stride + k);
    for (int j = 0; j < kJitters; ++j) {
      double o[3];
      deform_one(p, j, centres, q, o);
      if (!(o[0] >= 0.0 && o[0] < fA2 && o[1] >= 0.0 && o[1] < fA1 && o[2] >= 0.0 && o[2] < fA0)) continue;
      ++count;
      uint8_t* dst = grid + (((size_t)(int)o[2] * A1 + (size_t)(int)o[1]) * A2 + (size_t)(int)o[0]) * 3;
      dst[0] = (uint8_t)r; dst[1] = (uint8_t)g; dst[2] = (uint8_t)b;
    }
  }
  for (int s = 16; s > 0; s >>= 1) count += __shfl_xor_sync(0xffffffffu, count, s);
  if (nvalid && (threadIdx.x & 31) == 0 && count) atomicAdd(nvalid, (unsigned long long)count);
}

inline unsigned grid_for(int64_t n, int threads, int per_sm) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)p3d::sm_count() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
P3D_API int p3d_deform_centres(const float* pts, int64_t n, int64_t stride, int64_t* sums, double* centres,
                               p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && stride >= 1 && sums && centres, "deform_centres: bad arguments");
  cudaStream_t st = p3d::as_stream(stream);
  P3D_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(int64_t), st));
  const int64_t m = (n + stride - 1) / stride;
  if (m > 0) {
    P3D_REQUIRE(pts, "deform_centres: null points");
    deform_sums_kernel<<<grid_for(m, 256, 8), 256, 0, st>>>(pts, m, stride, reinterpret_cast<unsigned long long*>(sums));
    P3D_LAUNCH_CHECK();
  }
  deform_centres_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const unsigned long long*>(sums), m, centres);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_deform_points(const float* pts, int64_t n, int64_t stride, const double* centres, const double* deform,
                              const double* pix2vox, int64_t* out, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && stride >= 1 && centres && deform && pix2vox, "deform_points: bad arguments");
  const int64_t m = (n + stride - 1) / stride;
  if (m == 0) return P3D_OK;
  P3D_REQUIRE(pts && out, "deform_points: null pointer");
  deform_points_kernel<<<grid_for(m, 256, 8), 256, 0, p3d::as_stream(stream)>>>(pts, m, stride, centres, deform, pix2vox,
                                                                                reinterpret_cast<long long*>(out));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_pack_label_bits(const uint8_t* labels, int64_t n, int label, uint32_t* bits, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && label >= 0 && label < 256, "pack_label_bits: bad arguments");
  if (n == 0) return P3D_OK;
  P3D_REQUIRE(labels && bits, "pack_label_bits: null pointer");
  const int64_t padded = (n + 31) / 32 * 32;
  pack_label_bits_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, p3d::as_stream(stream)>>>(labels, n, label, bits);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

namespace {
template <typename T>
int deform_sweep(const float* pts, int64_t n, int64_t stride, const double* centres, const double* deforms, int D,
                 const double* pix2vox, int A0, int A1, int A2, const T* cam, const float* fast, const float* bbox,
                 const uint32_t* gt_bits, int H, int W, uint32_t* cov, int64_t* counts, int64_t* nvalid,
                 p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && stride >= 1 && D >= 0 && H > 0 && W > 0 && A0 > 0 && A1 > 0 && A2 > 0,
              "deform_sweep: n=%lld stride=%lld D=%d H=%d W=%d", (long long)n, (long long)stride, D, H, W);
  P3D_REQUIRE((int64_t)H * W < (1ll << 31) && W < (1 << 21) && H < (1 << 21), "deform_sweep: image too large");
  if (D == 0) return P3D_OK;
  P3D_REQUIRE(centres && deforms && pix2vox && cam && gt_bits && cov && counts && nvalid, "deform_sweep: null pointer");
  P3D_REQUIRE((fast == nullptr) == (bbox == nullptr), "deform_sweep: fast and bbox go together");
  cudaStream_t st = p3d::as_stream(stream);
  const int64_t m = (n + stride - 1) / stride;
  const int64_t words = ((int64_t)H * W + 31) / 32;
  P3D_CUDA(cudaMemsetAsync(counts, 0, (size_t)D * 2 * sizeof(int64_t), st));
  P3D_CUDA(cudaMemsetAsync(nvalid, 0, (size_t)D * sizeof(int64_t), st));
  if (m > 0) {
    P3D_REQUIRE(pts, "deform_sweep: null points");
    const int64_t tiles = (m + kDeformThreads - 1) / kDeformThreads;
    P3D_REQUIRE(tiles < (1ll << 31), "deform_sweep: too many tiles");
    dim3 grid((unsigned)tiles, (unsigned)((D + kDeformPerBlock - 1) / kDeformPerBlock));
    if (fast)                                              // FP32 filter block present (either exact dtype)
      deform_splat_kernel<T, true><<<grid, kDeformThreads, 0, st>>>(pts, m, stride, centres, deforms, D, pix2vox, A0, A1, A2,
                                                                    cam, fast, bbox, H, W, cov, words,
                                                                    reinterpret_cast<unsigned long long*>(nvalid));
    else
      deform_splat_kernel<T, false><<<grid, kDeformThreads, 0, st>>>(pts, m, stride, centres, deforms, D, pix2vox, A0, A1, A2,
                                                                     cam, fast, bbox, H, W, cov, words,
                                                                     reinterpret_cast<unsigned long long*>(nvalid));
    P3D_LAUNCH_CHECK();
  }
  dim3 sgrid(grid_for(words, 256, 2), (unsigned)D);
  deform_score_kernel<<<sgrid, 256, 0, st>>>(cov, gt_bits, words, reinterpret_cast<long long*>(counts));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}
}  // namespace

P3D_API int p3d_deform_sweep_f64(const float* pts, int64_t n, int64_t stride, const double* centres,
                                 const double* deforms, int D, const double* pix2vox, int A0, int A1, int A2,
                                 const double* cam, const float* fast, const float* bbox, const uint32_t* gt_bits, int H,
                                 int W, uint32_t* cov, int64_t* counts, int64_t* nvalid, p3d_stream_t stream) {
  return deform_sweep<double>(pts, n, stride, centres, deforms, D, pix2vox, A0, A1, A2, cam, fast, bbox, gt_bits, H, W, cov,
                              counts, nvalid, stream);
}
P3D_API int p3d_deform_sweep_f32(const float* pts, int64_t n, int64_t stride, const double* centres,
                                 const double* deforms, int D, const double* pix2vox, int A0, int A1, int A2,
                                 const float* cam, const float* fast, const float* bbox, const uint32_t* gt_bits, int H,
                                 int W, uint32_t* cov, int64_t* counts, int64_t* nvalid, p3d_stream_t stream) {
  return deform_sweep<float>(pts, n, stride, centres, deforms, D, pix2vox, A0, A1, A2, cam, fast, bbox, gt_bits, H, W, cov,
                             counts, nvalid, stream);
}

P3D_API int p3d_deform_scatter(const float* pts, int64_t n, int64_t stride, const double* centres, const double* deform,
                               const double* pix2vox, int A0, int A1, int A2, int r, int g, int b, uint8_t* grid_rgb,
                               int64_t* nvalid, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && stride >= 1 && A0 > 0 && A1 > 0 && A2 > 0 && centres && deform && pix2vox && grid_rgb,
              "deform_scatter: bad arguments");
  const int64_t m = (n + stride - 1) / stride;
  cudaStream_t st = p3d::as_stream(stream);
  if (nvalid) P3D_CUDA(cudaMemsetAsync(nvalid, 0, sizeof(int64_t), st));
  if (m == 0) return P3D_OK;
  P3D_REQUIRE(pts, "deform_scatter: null points");
  deform_scatter_kernel<<<grid_for(m, 256, 8), 256, 0, st>>>(pts, m, stride, centres, deform, pix2vox, A0, A1, A2, r, g, b,
                                                            grid_rgb, reinterpret_cast<unsigned long long*>(nvalid));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}
