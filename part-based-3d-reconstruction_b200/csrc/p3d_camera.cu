// libp3d_b200: perspective camera-candidate scoring (stage 2 of the reference).
//
// Reference call sites replaced here (paths under the reference root):
//   look_at_rotation            utils/camera_geometry.py:3-14
//   project_colored_voxels      utils/projection_utils.py:5-23
//   compute_partwise_iou        utils/camera_estimation.py:770-787
//   evaluate / run_random loop  utils/camera_estimation.py:597-603, 606-650
//
// Exactness contract (DESIGN.md "Projection arithmetic"): every FP operation is an explicit
// round-to-nearest intrinsic in the order NumPy/OpenBLAS use on the reference host --
//   d = p - c ; (X,Y,Z)[j] = fma(d2,R[j][2], fma(d1,R[j][1], d0*R[j][0])) ; Z<1e-8 -> 1e-8 ;
//   u = (X/Z)*f + cx ; v = (-(Y/Z))*f + cy ; rint half-even ; bounds test in floating point.
// Visibility is "largest point index wins" (NumPy fancy-assignment order), resolved with a
// 32-bit atomicMax on index+1; a plain load in front of the atomic skips pixels that already hold
// a larger index (valid because the buffer only grows), and tiles are walked from the high-index
// end so that this filter hits for most points.
#include <stdlib.h>

#include <new>
#include <vector>

#include "p3d_common.cuh"
#include "p3d_project.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// look_at_rotation for K candidates (one thread each).  cams[k] = cam_pos[3], R[9], f, cx, cy, 0.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) setup_cameras_kernel(const T* __restrict__ cand, int K, T* __restrict__ cams) {
  using F = Fp<T>;
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const T* c = cand + (size_t)k * 9;
  T eye[3] = {c[0], c[1], c[2]};
  T z[3], x[3], y[3], up[3] = {(T)0, (T)1, (T)0};
#pragma unroll
  for (int i = 0; i < 3; ++i) z[i] = F::sub(c[3 + i], eye[i]);
  T n = F::sqrt(F::dot3(z, z));
#pragma unroll
  for (int i = 0; i < 3; ++i) z[i] = F::div(z[i], n);
  T dzu = F::dot3(z, up);
  double a = fabs((double)dzu);
  if (fabs(__dsub_rn(a, 1.0)) <= 1e-8 + 1e-5 * 1.0) { up[0] = (T)0; up[1] = (T)0; up[2] = (T)1; }
  x[0] = F::sub(F::mul(up[1], z[2]), F::mul(up[2], z[1]));
  x[1] = F::sub(F::mul(up[2], z[0]), F::mul(up[0], z[2]));
  x[2] = F::sub(F::mul(up[0], z[1]), F::mul(up[1], z[0]));
  T nx = F::sqrt(F::dot3(x, x));
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = F::div(x[i], nx);
  y[0] = F::sub(F::mul(z[1], x[2]), F::mul(z[2], x[1]));
  y[1] = F::sub(F::mul(z[2], x[0]), F::mul(z[0], x[2]));
  y[2] = F::sub(F::mul(z[0], x[1]), F::mul(z[1], x[0]));
  T* o = cams + (size_t)k * 16;
  o[0] = eye[0]; o[1] = eye[1]; o[2] = eye[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[3 + i] = x[i]; o[6 + i] = y[i]; o[9 + i] = z[i]; }
  o[12] = c[6]; o[13] = c[7]; o[14] = c[8]; o[15] = (T)0;
}

// ------------------------------------------------------------------------------------------
// Splat.  CTA = 256 threads x kPpt points; blockIdx.y selects a group of `cams_per_block` cameras.
// ------------------------------------------------------------------------------------------
#ifndef P3D_SPLAT_THREADS
#define P3D_SPLAT_THREADS 128
#endif
#ifndef P3D_PPTF
#define P3D_PPTF 4
#endif
#ifndef P3D_MINBLOCKS
#define P3D_MINBLOCKS 8
#endif
constexpr int kSplatThreads = P3D_SPLAT_THREADS;
// Internal third mode of the sweep: joint visibility with the point's label carried in the key,
// key = (index + 1) << 5 | (label - 1).  The order of keys is the order of indices, so the winner is unchanged, and
// the score kernel reads the label out of the z-buffer instead of gathering pt_label[index] (n < 2^27, labels <= 32).
constexpr int kModeJointPacked = 2;
constexpr int kLabelBits = 5;
template <int MODE>
__device__ __forceinline__ uint32_t make_key(int64_t i, const uint8_t* __restrict__ pt_label) {
  if (MODE == P3D_MODE_JOINT) return (uint32_t)(i + 1);
  const uint32_t lab = (uint32_t)__ldg(pt_label + i) - 1u;
  if (MODE == kModeJointPacked) return ((uint32_t)(i + 1) << kLabelBits) | lab;
  return 1u << lab;
}
constexpr int kPpt = 2;

template <typename T, int MODE>
__global__ void __launch_bounds__(kSplatThreads)
splat_kernel(const float* __restrict__ pts, const uint8_t* __restrict__ pt_label, int64_t n,
             const T* __restrict__ cams, int K, int cams_per_block, int H, int W,
             uint32_t* __restrict__ zbuf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_cam = reinterpret_cast<T*>(smem_raw);
  const int c0 = blockIdx.y * cams_per_block;
  const int nc = min(cams_per_block, K - c0);
  for (int i = threadIdx.x; i < nc * 16; i += kSplatThreads) s_cam[i] = cams[(size_t)c0 * 16 + i];
  __syncthreads();

  // walk the point list from its high-index end: later points own the pixels
  const int64_t tile = (int64_t)gridDim.x - 1 - blockIdx.x;
  const int64_t base = tile * (kSplatThreads * kPpt) + threadIdx.x;
  T px[kPpt], py[kPpt], pz[kPpt];
  uint32_t key[kPpt];
#pragma unroll
  for (int j = 0; j < kPpt; ++j) {
    const int64_t i = base + (int64_t)j * kSplatThreads;
    key[j] = 0;
    px[j] = py[j] = pz[j] = (T)0;
    if (i < n) {
      px[j] = (T)__ldg(pts + 3 * i + 0);
      py[j] = (T)__ldg(pts + 3 * i + 1);
      pz[j] = (T)__ldg(pts + 3 * i + 2);
      key[j] = make_key<MODE>(i, pt_label);
    }
  }
  const T fW = (T)W, fH = (T)H;
  const size_t HW = (size_t)H * W;
  for (int c = 0; c < nc; ++c) {
    uint32_t* zb = zbuf + (size_t)(c0 + c) * HW;
#pragma unroll
    for (int j = 0; j < kPpt; ++j)
      exact_splat<T, MODE>(px[j], py[j], pz[j], key[j], s_cam + c * 16, zb, W, fW, fH);
  }
}

// ------------------------------------------------------------------------------------------
// Filtered FP64 splat.  The pixel a point lands on is decided in FP32 together with a rigorous bound on
// |u32 - u_ref| (u_ref = the reference's FP64 value); only when rounding could go either way within that bound is
// the point re-projected with the exact FP64 sequence.  Undecided (point, camera) pairs are parked in a per-warp
// shared-memory queue and drained 32 at a time, so the FP64 pass runs with full warps instead of diverging.
// The result is bit-identical to splat_kernel<double> (tests/test_camera_gpu.py).
//
// FP32 form.  Points are first re-centred on the middle c of the list's bounding box, q = fl(p - c) (one rounding; exact
// for voxel indices), which halves the magnitudes the roundings act on.  Per camera, computed in FP64 and rounded once:
// A_j = f R_0j, B_j = -f R_1j, C_j = R_2j, T_A = f (c - e).R_0, T_B = -f (c - e).R_1, T_C = (c - e).R_2:
//   X' = fma(q2,A2, fma(q1,A1, fma(q0,A0, T_A)))   (= f X),   Y' likewise (= -f Y),   Z = fma chain with C, T_C
//   r = rcp.approx(Z) ;  u = fma(X', r, cx) ;  v = fma(Y', r, cy)
// Error model (eps = 2^-24, |R| <= 1, S = sum_k max|p_k - c_k| over the box, Zmin = min of Z over the box in FP64).
// Roundings: q (1), each coefficient (1), T (1), three fma results (3 x the running magnitude <= f S + |T_A|):
//   |X'32 - f X| <= eps DX,  DX = 5.05 f S + 4.04 |T_A|            |Z32 - Z| <= eps DZ,  DZ = 5.05 S + 4.04 |T_C|
// rcp.approx is within 1 ulp (2 eps); the final fma rounds once (eps |u|); cx is rounded to FP32 (eps |cx|):
//   |u32 - u_ref| <= eps [ DX/Zmin + |u - cx| (DZ/Zmin + 2) + |cx| + |u| ]
// evaluated over the padded image, |u| <= W + 2 and |u - cx| <= Ud = max(|cx + 2|, |W + 2 - cx|) + 1; the kernel uses
// Bmax_u = 1.25 x that (second-order terms, FP64 noise of the reference itself) + 1e-7 and thr_u = 0.5 - Bmax_u.  A
// coordinate is decided when |u32 - rint(u32)| < thr_u; decided + in range -> pixel; decided + out of range -> dropped;
// else exact path.  (Out-of-range decisions stay valid beyond the padded image because the bound grows with slope
// b1 = 1.25 eps (DZ/Zmin + 3) per pixel of |u32|, and the camera is only "fast" when b1 * 2 max(W,H) <= 0.25 and
// Bmax <= 0.25: the true coordinate cannot come back across the image border.)
// Cameras whose box comes within max(1e-3, 2^-17 DZ) of the camera plane, or with non-finite / non-positive-f
// parameters, or whose bound exceeds 0.25 px, get thr = -1: every point takes the exact path.
// ------------------------------------------------------------------------------------------
constexpr int kPptF = P3D_PPTF;
constexpr int kQueueCap = 64;        // per warp: at most 32 new entries on top of < 32 pending

template <typename T>
__global__ void __launch_bounds__(64) fast_cams_kernel(const T* __restrict__ cams, int K,
                                                        const float* __restrict__ bbox, int H, int W,
                                                        float* __restrict__ fast, int4* __restrict__ rect,
                                                        uint32_t* __restrict__ flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double cam[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) cam[i] = (double)cams[(size_t)k * 16 + i];       // float32 blocks widen exactly
  FastCam* fc = reinterpret_cast<FastCam*>(fast) + k;
  make_fast_cam(cam, bbox, H, W, fc, sizeof(T) == 4);
  if (rect || flags) {
    const bool inside = footprint_rect(cam, bbox, H, W, fc->thr_u >= 0.f && fc->thr_v >= 0.f, rect ? rect + k : nullptr);
    if (flags) flags[k] = inside ? 1u : 0u;
  }
}

#ifndef P3D_SCALAR_FILTER
#define P3D_PACKED 1
#endif
#ifndef P3D_CAM_UNROLL
#define P3D_CAM_UNROLL 1
#endif
constexpr int kCamUnroll = P3D_CAM_UNROLL;   // camera-loop unroll factor (tuning knob; 1 measured best)
constexpr int kPackCamFloats = 16;   // per camera in shared memory: one FastCam (FFMA2 broadcasts a scalar operand)
constexpr int kFlushEvery = 64 / kPptF; // cameras between queue flushes: kPptF bits per camera in a 64-bit mask

template <typename T, int MODE>
__global__ void __launch_bounds__(kSplatThreads, P3D_MINBLOCKS)
splat_filtered_kernel(const float* __restrict__ pts, const uint8_t* __restrict__ pt_label, int64_t n,
                      const T* __restrict__ cams, int K, int cams_per_block, int H, int W,
                      uint32_t* __restrict__ zbuf, const float* __restrict__ fast, const float* __restrict__ bbox) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_cam = reinterpret_cast<T*>(smem_raw);                                             // nc x 16 camera scalars
  float* s_fast = reinterpret_cast<float*>(s_cam + (size_t)cams_per_block * 16);          // nc x kPackCamFloats
  uint2* s_queue = reinterpret_cast<uint2*>(s_fast + (size_t)cams_per_block * kPackCamFloats);   // 8 warps x kQueueCap

  const int c0 = blockIdx.y * cams_per_block;
  const int nc = min(cams_per_block, K - c0);
  for (int i = threadIdx.x; i < nc * 16; i += kSplatThreads) {
    s_cam[i] = cams[(size_t)c0 * 16 + i];
    s_fast[i] = fast[(size_t)c0 * 16 + i];                 // FastCam: A[3],TA, B[3],TB, C[3],TC, cx, cy, thr_u, thr_v
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint2* q = s_queue + warp * kQueueCap;
  int qn = 0;                                              // entries pending in the warp's queue (warp-uniform)
  const float ctr0 = bbox_centre(bbox, 0), ctr1 = bbox_centre(bbox, 1), ctr2 = bbox_centre(bbox, 2);

  const int64_t tile = (int64_t)gridDim.x - 1 - blockIdx.x;
  const int64_t base = tile * (kSplatThreads * kPptF) + threadIdx.x;
  float px[kPptF], py[kPptF], pz[kPptF];
  uint32_t key[kPptF];
  unsigned long long live = 0ull;                          // parked-mask bits of the points that exist (index < n)
#pragma unroll
  for (int j = 0; j < kPptF; ++j) {
    const int64_t i = base + (int64_t)j * kSplatThreads;
    key[j] = 0;
    px[j] = py[j] = pz[j] = __int_as_float(0x7fc00000);     // NaN: dead lanes never decide (and are never parked)
    if (i < n) {
      px[j] = __fsub_rn(__ldg(pts + 3 * i + 0), ctr0);      // re-centred coordinates of the FP32 filter
      py[j] = __fsub_rn(__ldg(pts + 3 * i + 1), ctr1);
      pz[j] = __fsub_rn(__ldg(pts + 3 * i + 2), ctr2);
      live |= 1ull << j;
      key[j] = make_key<MODE>(i, pt_label);
    }
  }
#pragma unroll
  for (int g = kPptF; g < 64; g <<= 1) live |= live << g;
  const T dW = (T)W, dH = (T)H;
  const uint32_t HW = (uint32_t)H * (uint32_t)W;
  const float kMagic = 12582912.f;                         // 1.5 * 2^23: (x + m) - m == rint(x) for |x| < 2^22
#ifdef P3D_PACKED
  static_assert(kPptF % 2 == 0, "packed path pairs points");
  f32x2 PX[kPptF / 2], PY[kPptF / 2], PZ[kPptF / 2];
#pragma unroll
  for (int jp = 0; jp < kPptF / 2; ++jp) {
    PX[jp] = pack2(px[2 * jp], px[2 * jp + 1]);
    PY[jp] = pack2(py[2 * jp], py[2 * jp + 1]);
    PZ[jp] = pack2(pz[2 * jp], pz[2 * jp + 1]);
  }
#endif

  // exact FP64 projection of the newest 32 (or all remaining) queue entries, one per lane
  auto drain32 = [&]() {
    const int take = qn < 32 ? qn : 32;
    if (lane < take) {
      const uint2 e = q[qn - take + lane];                  // (point index, camera)
      const float* pp = pts + 3 * (size_t)e.x;
      const uint32_t k = make_key<MODE>((int64_t)e.x, pt_label);
      exact_splat<T, MODE>((T)__ldg(pp), (T)__ldg(pp + 1), (T)__ldg(pp + 2), k, s_cam + e.y * 16,
                           zbuf + (size_t)(c0 + e.y) * HW, W, dW, dH);
    }
    qn -= take;
    __syncwarp();
  };

  // push the undecided (point, camera) pairs recorded in `mask`: group g of kPptF bits (from the low end) belongs to
  // camera clast - g, bit j of the group to point j.  Every round each lane pushes at most one entry -- slots come
  // from a ballot, the count stays in a register -- then a full group of 32 is drained: the queue never exceeds 63.
  auto flush = [&](unsigned long long mask, int clast) {
    mask &= live;
    const uint32_t lt = (1u << lane) - 1u;
    for (;;) {
      const bool has = mask != 0ull;
      const uint32_t m = __ballot_sync(0xffffffffu, has);
      if (m == 0u) break;
      if (has) {
        const uint32_t b = (uint32_t)(__ffsll((long long)mask) - 1);
        mask &= mask - 1ull;
        const uint32_t idx = (uint32_t)base + (b % (uint32_t)kPptF) * (uint32_t)kSplatThreads;
        q[qn + __popc(m & lt)] = make_uint2(idx, (uint32_t)clast - b / (uint32_t)kPptF);
      }
      qn += __popc(m);
      __syncwarp();
      if (qn >= 32) drain32();
    }
  };

  unsigned long long parked = 0ull;
  uint32_t* zb = zbuf + (size_t)c0 * HW;                   // this camera's z-buffer: advanced once per pass
#pragma unroll kCamUnroll
  for (int c = 0; c < nc; ++c, zb += HW) {
    uint32_t* addr[kPptF];
    bool hit[kPptF];
    uint32_t und = 0;
    const float4* fc4 = reinterpret_cast<const float4*>(s_fast + c * kPackCamFloats);
    const float4 vA = fc4[0], vB = fc4[1], vC = fc4[2], vD = fc4[3];
    const float thr_u = vD.z, thr_v = vD.w;
#ifdef P3D_PACKED
    // pack2(a, a) costs nothing: FFMA2 takes a scalar register as a broadcast operand
    const f32x2 kM2 = pack2(kMagic, kMagic), kNegM2 = pack2(-kMagic, -kMagic), kNeg1 = pack2(-1.f, -1.f);
    const f32x2 A0 = pack2(vA.x, vA.x), A1 = pack2(vA.y, vA.y), A2 = pack2(vA.z, vA.z), TA = pack2(vA.w, vA.w);
    const f32x2 B0 = pack2(vB.x, vB.x), B1 = pack2(vB.y, vB.y), B2 = pack2(vB.z, vB.z), TB = pack2(vB.w, vB.w);
    const f32x2 C0 = pack2(vC.x, vC.x), C1 = pack2(vC.y, vC.y), C2 = pack2(vC.z, vC.z), TC = pack2(vC.w, vC.w);
    const f32x2 CX = pack2(vD.x, vD.x), CY = pack2(vD.y, vD.y);
#pragma unroll
    for (int jp = 0; jp < kPptF / 2; ++jp) {              // points 2jp and 2jp+1 share each FFMA2 / FADD2
      const f32x2 X = fma2(PZ[jp], A2, fma2(PY[jp], A1, fma2(PX[jp], A0, TA)));
      const f32x2 Y = fma2(PZ[jp], B2, fma2(PY[jp], B1, fma2(PX[jp], B0, TB)));
      const f32x2 Z = fma2(PZ[jp], C2, fma2(PY[jp], C1, fma2(PX[jp], C0, TC)));
      float z0, z1, r0, r1;
      unpack2(Z, z0, z1);
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(z0));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(z1));
      const f32x2 R = pack2(r0, r1);
      const f32x2 U = fma2(X, R, CX), V = fma2(Y, R, CY);
      const f32x2 SU = add2(U, kM2), SV = add2(V, kM2);
      const f32x2 DU = fma2(add2(SU, kNegM2), kNeg1, U), DV = fma2(add2(SV, kNegM2), kNeg1, V);   // u - rint(u)
      float du[2], dv[2], su[2], sv[2];
      unpack2(DU, du[0], du[1]); unpack2(DV, dv[0], dv[1]); unpack2(SU, su[0], su[1]); unpack2(SV, sv[0], sv[1]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * jp + h;
        const bool decided = fabsf(du[h]) < thr_u && fabsf(dv[h]) < thr_v;                        // false for NaN
        const uint32_t iu = (uint32_t)(__float_as_int(su[h]) - 0x4B400000), iv = (uint32_t)(__float_as_int(sv[h]) - 0x4B400000);
        hit[j] = decided && iu < (uint32_t)W && iv < (uint32_t)H;
        addr[j] = zb + (iv * (uint32_t)W + iu);          // only dereferenced when hit
        if (!decided) und |= 1u << j;
      }
    }
#else
#pragma unroll
    for (int j = 0; j < kPptF; ++j) {
      const float X = fmaf(pz[j], vA.z, fmaf(py[j], vA.y, fmaf(px[j], vA.x, vA.w)));
      const float Y = fmaf(pz[j], vB.z, fmaf(py[j], vB.y, fmaf(px[j], vB.x, vB.w)));
      const float Z = fmaf(pz[j], vC.z, fmaf(py[j], vC.y, fmaf(px[j], vC.x, vC.w)));
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(Z));
      const float u = fmaf(X, r, vD.x);
      const float v = fmaf(Y, r, vD.y);
      const float su = u + kMagic, sv = v + kMagic;
      const bool decided = fabsf(u - (su - kMagic)) < vD.z && fabsf(v - (sv - kMagic)) < vD.w;   // false for NaN
      const uint32_t iu = (uint32_t)(__float_as_int(su) - 0x4B400000), iv = (uint32_t)(__float_as_int(sv) - 0x4B400000);
      hit[j] = decided && iu < (uint32_t)W && iv < (uint32_t)H;
      addr[j] = zb + (iv * (uint32_t)W + iu);            // only dereferenced when hit
      if (!decided) und |= 1u << j;
    }
#endif
    parked = (parked << kPptF) | (unsigned long long)und;
    // all early-out loads first (memory-level parallelism), then the reductions
    uint32_t cur[kPptF];                                   // only meaningful (and only read) where hit[j]
#pragma unroll
    for (int j = 0; j < kPptF; ++j) {
#ifdef P3D_EARLY_CA
      if (hit[j]) cur[j] = __ldca(addr[j]);
#else
      if (hit[j]) cur[j] = __ldcg(addr[j]);
#endif
    }
#pragma unroll
    for (int j = 0; j < kPptF; ++j) {
      if (MODE != P3D_MODE_PER_PART) {
        if (hit[j] && cur[j] < key[j]) atomicMax(addr[j], key[j]);
      } else {
        if (hit[j] && (cur[j] & key[j]) != key[j]) atomicOr(addr[j], key[j]);
      }
    }
    if ((c & (kFlushEvery - 1)) == kFlushEvery - 1) {
      flush(parked, c);
      parked = 0ull;
    }
  }
  flush(parked, nc - 1);
  while (qn > 0) drain32();
}

// ------------------------------------------------------------------------------------------
// Segment splat: the same filter, one thread per segment of an x-run (p3d_segments_*, p3d_core.cu) instead of kPptF
// unrelated points.  A chunk of n <= 32 L consecutive voxels of one (y, z) row is dealt out column-wise to T = ceil(n/L)
// neighbouring threads: thread r owns x0 + r + T j, j < L.  What the run structure buys per (point, camera):
//   * the (y, z) part of the three fma chains is evaluated once per segment:  Xrow = fma(qz,A2, fma(qy,A1, TA)) (X and Y
//     rows share packed instructions), then X_j = fma(qx_j, A0, Xrow) for the points of the segment -- the same three
//     roundings per coordinate as the per-point chain in another association order, and the error model above bounds
//     every partial sum by f S + |T_A| regardless of order, so the thresholds of make_fast_cam() hold unchanged;
//     qx_j = fl((x_first + T j) - c_x) is the very value the per-point kernel computes (x_first + T j is an exact
//     integer);
//   * a thread carries 3 coordinates + key and key step for kSegLen points (key_j = key0 + j * step), so kSegLen = 8
//     points per thread fit the register budget and the per-camera loop overhead (camera block loads, parked-mask
//     upkeep) is shared by twice as many points;
//   * the pixel index comes out of the FP pipe: lin = fma(rint(v), W, u + magic) holds row * W + col in its mantissa
//     (exact while H W <= 2^22) and indexes the camera's z-buffer directly (base biased by bits(magic));
//   * cameras whose footprint rectangle (footprint_rect, unclamped, 2.5 px margin) lies inside the image need no bounds
//     test at all: a decided point of such a camera is in range by construction (kCamInView);
//   * neighbouring threads still hold neighbouring voxels, so a warp's early-out loads touch about as many 32-byte
//     sectors as the per-point kernel's (a thread owning 8 ADJACENT voxels measured 23 k cand/s against 35 k: every
//     lane its own sector, L2-sector bound).
// Points beyond a segment's count carry NaN coordinates: never decided, never parked (`live`).
// Bit-identical to splat_kernel<T> like the per-point filter (tests/test_camera_gpu.py, tests/test_segments_gpu.py).
// ------------------------------------------------------------------------------------------
#ifndef P3D_SEG_LEN
#define P3D_SEG_LEN 8
#endif
#ifndef P3D_SEG_GROUP
#define P3D_SEG_GROUP 4
#endif
#ifndef P3D_SEG_MINBLOCKS
#define P3D_SEG_MINBLOCKS 8
#endif
constexpr int kSegLen = P3D_SEG_LEN;          // points per segment (p3d_segment_length())
constexpr int kSegGroup = P3D_SEG_GROUP;      // points whose early-out loads are in flight together
#ifndef P3D_SEG_THREADS
#define P3D_SEG_THREADS 128
#endif
constexpr int kSegThreads = P3D_SEG_THREADS;
constexpr int kSegFlushEvery = 64 / kSegLen;  // cameras between queue flushes: kSegLen bits per camera in a 64-bit mask
constexpr uint32_t kCamInView = 1u;           // cam_flags bit: every decided point of this camera is inside the image
static_assert(kSegLen % kSegGroup == 0 && kSegGroup % 2 == 0 && 64 % kSegLen == 0, "segment shape");

template <int MODE>
__device__ __forceinline__ uint32_t seg_key0(uint32_t idx0, uint32_t lab) {
  if (MODE == P3D_MODE_JOINT) return idx0 + 1u;
  if (MODE == kModeJointPacked) return ((idx0 + 1u) << kLabelBits) | (lab - 1u);
  return 1u << (lab - 1u);
}
// key step between two points of a segment whose list indices are T apart
template <int MODE>
__device__ __forceinline__ uint32_t seg_key_step(uint32_t T) {
  if (MODE == P3D_MODE_JOINT) return T;
  if (MODE == kModeJointPacked) return T << kLabelBits;
  return 0u;
}

// 16-byte shared-memory load from a 32-bit shared address (kept in a register and stepped per camera: the compiler's own
// address arithmetic re-derived the CTA's shared window every iteration)
__device__ __forceinline__ float4 lds128(uint32_t saddr, int byte_off) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr + byte_off));
  return v;
}

// One camera for one segment: returns the mask of undecided points (bit j).  zb = this camera's z-buffer minus
// bits(magic) elements, so that the raw bits of `lin` index it.  c0..c3 = the camera's 16 floats (layout below).
// kmax = the greatest key of the segment: `cur >= kmax` proves that no point of the segment can win the pixel; a
// pixel that fails this coarser test gets the reduction with the point's exact key (atomicMax is idempotent), which
// only costs extra reductions for pixels currently owned by another voxel of the same chunk.
template <bool INVIEW, int MODE>
__device__ __forceinline__ uint32_t seg_camera_pass(const f32x2 (&QX)[kSegLen / 2], float qy, float qz, uint32_t key0,
                                                    uint32_t kstep, uint32_t kmax, const float4 c0, const float4 c1,
                                                    const float4 c2, const float4 c3, uint32_t* __restrict__ zb,
                                                    uint32_t W, uint32_t H, float Wf) {
  const float kMagic = 12582912.f;                          // 1.5 * 2^23
  const uint32_t kMagicBits = 0x4B400000u;
  // c0 = A0 B0 C0 thr | c1 = A1 B1 A2 B2 | c2 = TA TB C1 C2 | c3 = TC cx cy flags
  const f32x2 XYr = fma2(pack2(qz, qz), pack2(c1.z, c1.w), fma2(pack2(qy, qy), pack2(c1.x, c1.y), pack2(c2.x, c2.y)));
  float xr, yr;
  unpack2(XYr, xr, yr);
  const float zr = __fmaf_rn(qz, c2.w, __fmaf_rn(qy, c2.z, c3.x));
  const f32x2 A0 = pack2(c0.x, c0.x), B0 = pack2(c0.y, c0.y), C0 = pack2(c0.z, c0.z);
  const f32x2 XR = pack2(xr, xr), YR = pack2(yr, yr), ZR = pack2(zr, zr);
  const f32x2 CX = pack2(c3.y, c3.y), CY = pack2(c3.z, c3.z), WF = pack2(Wf, Wf);
  const f32x2 kM2 = pack2(kMagic, kMagic), kNegM2 = pack2(-kMagic, -kMagic), kNeg1 = pack2(-1.f, -1.f);
  const float thr = c0.w;                                   // min(thr_u, thr_v)
  uint32_t und = 0;
#pragma unroll
  for (int g = 0; g < kSegLen / kSegGroup; ++g) {
    uint32_t* addr[kSegGroup];
    bool hit[kSegGroup];
#pragma unroll
    for (int mp = 0; mp < kSegGroup / 2; ++mp) {
      const int m = g * (kSegGroup / 2) + mp;
      const f32x2 X = fma2(QX[m], A0, XR), Y = fma2(QX[m], B0, YR), Z = fma2(QX[m], C0, ZR);
      float z0, z1, r0, r1;
      unpack2(Z, z0, z1);
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(z0));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(z1));
      const f32x2 R = pack2(r0, r1);
      const f32x2 U = fma2(X, R, CX), V = fma2(Y, R, CY);
      const f32x2 SU = add2(U, kM2), SV = add2(V, kM2);
      const f32x2 RV = add2(SV, kNegM2);
      const f32x2 DU = fma2(add2(SU, kNegM2), kNeg1, U), DV = fma2(RV, kNeg1, V);          // u - rint(u)
      const f32x2 LIN = fma2(RV, WF, SU);                                                   // row * W + col + magic
      float du[2], dv[2], su[2], sv[2], lin[2];
      unpack2(DU, du[0], du[1]); unpack2(DV, dv[0], dv[1]); unpack2(LIN, lin[0], lin[1]);
      unpack2(SU, su[0], su[1]); unpack2(SV, sv[0], sv[1]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int jj = 2 * mp + h;
        const bool decided = fabsf(du[h]) < thr && fabsf(dv[h]) < thr;                      // false for NaN
        if (INVIEW) {
          hit[jj] = decided;
        } else {
          const uint32_t iu = (uint32_t)__float_as_int(su[h]) - kMagicBits, iv = (uint32_t)__float_as_int(sv[h]) - kMagicBits;
          hit[jj] = decided && iu < W && iv < H;
        }
        addr[jj] = zb + (uint32_t)__float_as_int(lin[h]);                                   // only dereferenced when hit
        if (!decided) und |= 1u << (g * kSegGroup + jj);
      }
    }
    uint32_t cur[kSegGroup];                                 // only meaningful (and only read) where hit
#pragma unroll
    for (int jj = 0; jj < kSegGroup; ++jj)
      if (hit[jj]) cur[jj] = __ldcg(addr[jj]);
#pragma unroll
    for (int jj = 0; jj < kSegGroup; ++jj) {
      if (MODE != P3D_MODE_PER_PART) {
        if (hit[jj] && cur[jj] < kmax) atomicMax(addr[jj], key0 + (uint32_t)(g * kSegGroup + jj) * kstep);
      } else {
        if (hit[jj] && (cur[jj] & key0) != key0) atomicOr(addr[jj], key0);
      }
    }
  }
  return und;
}

#ifndef P3D_SEG_QUEUE
#define P3D_SEG_QUEUE 96
#endif
constexpr int kSegQueueCap = P3D_SEG_QUEUE;   // per warp: < 32 pending + one bulk push of <= kSegQueueCap - 32 entries

template <typename T, int MODE>
__global__ void __launch_bounds__(kSegThreads, P3D_SEG_MINBLOCKS)
splat_seg_kernel(const uint4* __restrict__ segs, int64_t n_seg, const float* __restrict__ pts,
                 const uint8_t* __restrict__ pt_label, const T* __restrict__ cams, int K, int cams_per_block, int H,
                 int W, uint32_t* __restrict__ zbuf, const float* __restrict__ fast, const float* __restrict__ bbox,
                 const uint32_t* __restrict__ cam_flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_cam = reinterpret_cast<T*>(smem_raw);                                             // nc x 16 camera scalars (exact path)
  float* s_fast = reinterpret_cast<float*>(s_cam + (size_t)cams_per_block * 16);          // nc x 16, permuted FastCam
  uint2* s_queue = reinterpret_cast<uint2*>(s_fast + (size_t)cams_per_block * 16);        // warps x kSegQueueCap
  // (reading the exact path's camera blocks from global memory instead measured 15 % slower: 32 scattered 120-byte
  // reads per drain)

  const int c0 = blockIdx.y * cams_per_block;
  const int nc = min(cams_per_block, K - c0);
  // FastCam = A[3],TA, B[3],TB, C[3],TC, cx, cy, thr_u, thr_v  ->  A0 B0 C0 thr | A1 B1 A2 B2 | TA TB C1 C2 | TC cx cy flags
  // with thr = min(thr_u, thr_v) (one threshold for both coordinates) and flags = the camera's cam_flags word
  const unsigned long long kPerm = 0xFDCBA9736251E840ull;   // nibble i = source slot of destination slot i
  for (int i = threadIdx.x; i < nc * 16; i += kSegThreads) {
    s_cam[i] = cams[(size_t)c0 * 16 + i];
    const float* f = fast + (size_t)(c0 + (i >> 4)) * 16;
    float v = f[(int)((kPerm >> (4 * (i & 15))) & 15ull)];
    if ((i & 15) == 3) v = fminf(f[14], f[15]);
    if ((i & 15) == 15) v = __uint_as_float(cam_flags ? cam_flags[c0 + (i >> 4)] : 0u);
    s_fast[i] = v;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint2* q = s_queue + warp * kSegQueueCap;
  int qn = 0;                                              // entries pending in the warp's queue (warp-uniform)
  const float ctr0 = bbox_centre(bbox, 0), ctr1 = bbox_centre(bbox, 1), ctr2 = bbox_centre(bbox, 2);

  // walk the segment list from its high-index end: later points own the pixels
  const int64_t tile = (int64_t)gridDim.x - 1 - blockIdx.x;
  const int64_t si = tile * kSegThreads + threadIdx.x;
  uint4 sg = make_uint4(0, 0, 0, 0);
  int len = 0;                                             // points of this segment: x_first + nT j, j < len
  if (si < n_seg) {
    sg = __ldg(segs + si);
    len = (int)((sg.y >> 16) & 0xfu) + 1;
  }
  const uint32_t idx0 = sg.z, nT = ((sg.y >> 20) & 0x3fu) + 1u;   // nT = segments of this chunk = x step
  const uint32_t key0 = seg_key0<MODE>(idx0, len ? (sg.y >> 26) : 1u), kstep = seg_key_step<MODE>(nT);
  const uint32_t kmax = key0 + (uint32_t)(kSegLen - 1) * kstep;   // >= every key of the segment (n < 2^27: no wrap)
  const float qy = __fsub_rn((float)(sg.x >> 16), ctr1), qz = __fsub_rn((float)(sg.y & 0xffffu), ctr2);
  const float nan = __int_as_float(0x7fc00000);
  f32x2 QX[kSegLen / 2];
#pragma unroll
  for (int m = 0; m < kSegLen / 2; ++m) {
    const float xa = __fsub_rn((float)((sg.x & 0xffffu) + nT * (2 * m)), ctr0);
    const float xb = __fsub_rn((float)((sg.x & 0xffffu) + nT * (2 * m + 1)), ctr0);
    QX[m] = pack2(2 * m < len ? xa : nan, 2 * m + 1 < len ? xb : nan);
  }
  unsigned long long live = len >= 64 ? ~0ull : ((1ull << len) - 1ull);   // parked-mask bits of the points that exist
#pragma unroll
  for (int g = kSegLen; g < 64; g <<= 1) live |= live << g;
  const T dW = (T)W, dH = (T)H;
  const uint32_t HW = (uint32_t)H * (uint32_t)W;
  const float Wf = (float)W;

  auto drain32 = [&]() {
    const int take = qn < 32 ? qn : 32;
    if (lane < take) {
      const uint2 e = q[qn - take + lane];                  // (point index, camera)
      const float* pp = pts + 3 * (size_t)e.x;
      const uint32_t k = make_key<MODE>((int64_t)e.x, pt_label);
      exact_splat<T, MODE>((T)__ldg(pp), (T)__ldg(pp + 1), (T)__ldg(pp + 2), k, s_cam + e.y * 16,
                           zbuf + (size_t)(c0 + e.y) * HW, W, dW, dH);
    }
    qn -= take;
    __syncwarp();
  };
  // Push the undecided (point, camera) pairs of `mask`: group g of kSegLen bits (from the low end) belongs to camera
  // clast - g, bit j of the group to point idx0 + nT j.  Usual case: the warp's entries fit the queue, every lane writes
  // its own at an offset from a warp scan of the counts.  Otherwise (cameras without a valid filter: everything is
  // undecided) rounds of at most one entry per lane, draining as the queue fills.
  auto flush = [&](unsigned long long mask, int clast) {
    mask &= live;
    const int cnt = __popcll(mask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    if (qn + total <= kSegQueueCap) {
      uint2* w = q + qn + incl - cnt;
      while (mask) {
        const uint32_t b = (uint32_t)(__ffsll((long long)mask) - 1);
        mask &= mask - 1ull;
        *w++ = make_uint2(idx0 + nT * (b % (uint32_t)kSegLen), (uint32_t)clast - b / (uint32_t)kSegLen);
      }
      qn += total;
      __syncwarp();
      while (qn >= 32) drain32();
      return;
    }
    const uint32_t lt = (1u << lane) - 1u;
    for (;;) {
      const bool has = mask != 0ull;
      const uint32_t m = __ballot_sync(0xffffffffu, has);
      if (m == 0u) break;
      if (has) {
        const uint32_t b = (uint32_t)(__ffsll((long long)mask) - 1);
        mask &= mask - 1ull;
        q[qn + __popc(m & lt)] = make_uint2(idx0 + nT * (b % (uint32_t)kSegLen), (uint32_t)clast - b / (uint32_t)kSegLen);
      }
      qn += __popc(m);
      __syncwarp();
      if (qn >= 32) drain32();
    }
  };

  unsigned long long parked = 0ull;
  uint32_t* zb = zbuf + (size_t)c0 * HW - 0x4B400000ll;    // biased by bits(magic): indexed by the raw bits of `lin`
  uint32_t sf = (uint32_t)__cvta_generic_to_shared(s_fast);
  for (int c = 0; c < nc; ++c, zb += HW, sf += 64) {
    const float4 k0 = lds128(sf, 0), k1 = lds128(sf, 16), k2 = lds128(sf, 32), k3 = lds128(sf, 48);
    uint32_t und;
    if (__float_as_uint(k3.w) & kCamInView)
      und = seg_camera_pass<true, MODE>(QX, qy, qz, key0, kstep, kmax, k0, k1, k2, k3, zb, (uint32_t)W, (uint32_t)H, Wf);
    else
      und = seg_camera_pass<false, MODE>(QX, qy, qz, key0, kstep, kmax, k0, k1, k2, k3, zb, (uint32_t)W, (uint32_t)H, Wf);
    parked = (parked << kSegLen) | (unsigned long long)und;
    if ((c & (kSegFlushEvery - 1)) == kSegFlushEvery - 1) {
      flush(parked, c);
      parked = 0ull;
    }
  }
  flush(parked, nc - 1);
  while (qn > 0) drain32();
}

// Bounding box of a point list: bbox = (min x, min y, min z, max x, max y, max z); NaNs are ignored.
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  int old = __float_as_int(*addr);
  while (v < __int_as_float(old)) {
    const int seen = atomicCAS(reinterpret_cast<int*>(addr), old, __float_as_int(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  int old = __float_as_int(*addr);
  while (v > __int_as_float(old)) {
    const int seen = atomicCAS(reinterpret_cast<int*>(addr), old, __float_as_int(v));
    if (seen == old) break;
    old = seen;
  }
}

__global__ void bbox_init_kernel(float* __restrict__ bbox) {
  if (threadIdx.x < 6) bbox[threadIdx.x] = threadIdx.x < 3 ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
}

__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ pts, int64_t n, float* __restrict__ bbox) {
  float lo[3], hi[3];
  for (int k = 0; k < 3; ++k) { lo[k] = __int_as_float(0x7f800000); hi[k] = __int_as_float(0xff800000); }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; ++k) {
      const float v = __ldg(pts + 3 * i + k);
      lo[k] = fminf(lo[k], v);
      hi[k] = fmaxf(hi[k], v);
    }
  for (int k = 0; k < 3; ++k)
    for (int d = 16; d > 0; d >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
    }
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 3; ++k) { atomic_min_float(bbox + k, lo[k]); atomic_max_float(bbox + 3 + k, hi[k]); }
}

// ------------------------------------------------------------------------------------------
// Score: z-buffer (+ point labels) vs ground-truth labels -> per-part (area, inter) counts.
// Warp ballots over the distinct labels present in the warp, popc, per-warp shared counters, one
// global atomic per (block, part).  The z-buffer is cleared on the way (only touched pixels).
// raw layout: [camera][P+1][2] = (area, inter); row P = combined binary (per-part mode).
// ------------------------------------------------------------------------------------------
constexpr int kScoreThreads = 256;
constexpr int kMaxParts = 32;

// accumulate one pixel per lane into the warp's counters (all 32 lanes must call)
template <int MODE>
__device__ __forceinline__ void score_accumulate(uint32_t key, uint32_t g, bool ga, const uint8_t* __restrict__ pt_label,
                                                 int P, unsigned int* acc, int lane) {
  if (MODE != P3D_MODE_PER_PART) {
    const uint32_t lab = MODE == kModeJointPacked ? (key ? (key & ((1u << kLabelBits) - 1u)) + 1u : 0u)
                                                  : (key ? (uint32_t)__ldg(pt_label + (key - 1)) : 0u);
    uint32_t rem = __ballot_sync(0xffffffffu, lab != 0);
    while (rem) {
      const int leader = __ffs(rem) - 1;
      const uint32_t l = __shfl_sync(0xffffffffu, lab, leader);
      const uint32_t m = __ballot_sync(0xffffffffu, lab == l);
      const uint32_t mi = __ballot_sync(0xffffffffu, lab == l && g == l);
      if (lane == 0 && l <= (uint32_t)P) {
        acc[(l - 1) * 2] += __popc(m);
        acc[(l - 1) * 2 + 1] += __popc(mi);
      }
      rem &= ~m;
    }
  } else {
    uint32_t present = __reduce_or_sync(0xffffffffu, key);
    while (present) {
      const int b = __ffs(present) - 1;
      present &= present - 1;
      const bool has = (key >> b) & 1u;
      const uint32_t m = __ballot_sync(0xffffffffu, has);
      const uint32_t mi = __ballot_sync(0xffffffffu, has && g == (uint32_t)(b + 1));
      if (lane == 0 && b < P) {
        acc[b * 2] += __popc(m);
        acc[b * 2 + 1] += __popc(mi);
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, key != 0);
    const uint32_t mi = __ballot_sync(0xffffffffu, key != 0 && ga);
    if (lane == 0) {
      acc[P * 2] += __popc(m);
      acc[P * 2 + 1] += __popc(mi);
    }
  }
}

// One camera per blockIdx.y.  A thread reads 4 consecutive pixels (uint4 of keys, 4 GT bytes); warps whose 128 pixels
// are all untouched skip after one vote; touched quads are cleared with one 16-byte store.  `vec` = HW % 4 == 0.
template <int MODE>
__global__ void __launch_bounds__(kScoreThreads)
score_kernel(uint32_t* __restrict__ zbuf, const uint8_t* __restrict__ pt_label,
             const uint8_t* __restrict__ gt_label, const uint8_t* __restrict__ gt_any, int HW, int P,
             unsigned long long* __restrict__ raw, int vec, int W, const int4* __restrict__ rect) {
  __shared__ unsigned int s_acc[kScoreThreads / 32][(kMaxParts + 1) * 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = (P + 1) * 2;
  for (int i = lane; i < rows; i += 32) s_acc[warp][i] = 0;
  __syncwarp();
  uint32_t* zb = zbuf + (size_t)blockIdx.y * HW;
  unsigned int* acc = s_acc[warp];
  // footprint rectangle of this camera (footprint_rect): nothing outside it was written, so only it is read and cleared
  int4 rc = make_int4(0, HW / W - 1, 0, W - 1);
  if (rect) rc = rect[blockIdx.y];
  if (rc.y < rc.x) return;                                   // nothing of this camera can be in the image
  if (vec) {
    // quads: with W % 4 == 0 a rectangle of whole quads, else the contiguous quad range covering rows rc.x .. rc.y
    const bool rows4 = (W & 3) == 0;
    const int wq_full = W >> 2;
    const int q0 = rows4 ? rc.z >> 2 : 0, wq = rows4 ? (rc.w >> 2) - q0 + 1 : 0;
    const int first = rows4 ? 0 : (int)(((int64_t)rc.x * W) >> 2);
    const int nq = rows4 ? (rc.y - rc.x + 1) * wq : (int)((((int64_t)(rc.y + 1) * W + 3) >> 2) - first);
    const int stride = gridDim.x * kScoreThreads;
    const int iters = (nq + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
      const int idx = it * stride + blockIdx.x * kScoreThreads + threadIdx.x;
      const bool in = idx < nq;
      int qd = first + idx;
      if (rows4) {
        const int r = idx / wq;
        qd = (rc.x + r) * wq_full + q0 + (idx - r * wq);
      }
      uint4 k4 = make_uint4(0, 0, 0, 0);
      if (in) k4 = __ldcg(reinterpret_cast<const uint4*>(zb) + qd);
      const bool touched = (k4.x | k4.y | k4.z | k4.w) != 0;
      if (!__any_sync(0xffffffffu, touched)) continue;
      uint32_t g4 = 0, a4 = 0;
      if (touched) {
        reinterpret_cast<uint4*>(zb)[qd] = make_uint4(0, 0, 0, 0);
        g4 = __ldg(reinterpret_cast<const uint32_t*>(gt_label) + qd);
        if (MODE == P3D_MODE_PER_PART && gt_any) a4 = __ldg(reinterpret_cast<const uint32_t*>(gt_any) + qd);
      }
      const uint32_t ks[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (!__any_sync(0xffffffffu, ks[s] != 0)) continue;
        score_accumulate<MODE>(ks[s], (g4 >> (8 * s)) & 0xffu, ((a4 >> (8 * s)) & 0xffu) != 0, pt_label, P, acc, lane);
      }
    }
  } else {
    const int first = rc.x * W, npx = (rc.y - rc.x + 1) * W;     // whole rows rc.x .. rc.y
    const int stride = gridDim.x * kScoreThreads;
    const int iters = (npx + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
      const int pix = first + it * stride + blockIdx.x * kScoreThreads + threadIdx.x;
      const bool in = pix < first + npx;
      const uint32_t key = in ? __ldcg(zb + pix) : 0u;
      if (!__any_sync(0xffffffffu, key != 0)) continue;
      if (key) zb[pix] = 0u;
      const uint32_t g = in ? (uint32_t)__ldg(gt_label + pix) : 0u;
      const bool ga = MODE == P3D_MODE_PER_PART && in && gt_any != nullptr && __ldg(gt_any + pix) != 0;
      score_accumulate<MODE>(key, g, ga, pt_label, P, acc, lane);
    }
  }
  __syncthreads();
  unsigned long long* out = raw + (size_t)blockIdx.y * rows;
  for (int i = threadIdx.x; i < rows; i += kScoreThreads) {
    unsigned int s = 0;
#pragma unroll
    for (int w = 0; w < kScoreThreads / 32; ++w) s += s_acc[w][i];
    if (s) atomicAdd(out + i, (unsigned long long)s);
  }
}

// Histogram of the ground-truth labels: gt_area[p-1] = |gt == p| for p = 1..P, gt_area[P] = |gt_any != 0|.
__global__ void __launch_bounds__(256) gt_area_kernel(const uint8_t* __restrict__ gt_label,
                                                      const uint8_t* __restrict__ gt_any, int HW, int P,
                                                      unsigned long long* __restrict__ gt_area) {
  __shared__ unsigned int s_hist[kMaxParts + 2];
  for (int i = threadIdx.x; i < kMaxParts + 2; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    const uint32_t g = gt_label[pix];
    if (g >= 1 && g <= (uint32_t)P) atomicAdd(&s_hist[g - 1], 1u);
    if (gt_any && gt_any[pix]) atomicAdd(&s_hist[P], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= P; i += blockDim.x)
    if (s_hist[i]) atomicAdd(gt_area + i, (unsigned long long)s_hist[i]);
}

// np.mean of P doubles exactly as NumPy reduces a contiguous 1-D float64 array (add.reduce starts from
// the identity 0.0 and hands all P values to DOUBLE_pairwise_sum): pairwise(n<8) = sequential from -0.0;
// pairwise(8<=n<=128) = 8 interleaved accumulators combined as ((0+1)+(2+3))+((4+5)+(6+7)), then the
// tail; finally / P.  (P <= 32 here, so the recursive n > 128 branch never runs.)
__device__ double numpy_mean(const double* a, int P) {
  if (P <= 0) return __longlong_as_double(0x7ff8000000000000ll);
  const int n = P;
  double s;
  if (n < 8) {
    s = -0.0;
    for (int i = 0; i < n; ++i) s = __dadd_rn(s, a[i]);
  } else {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                  __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) s = __dadd_rn(s, a[i]);
  }
  s = __dadd_rn(0.0, s);
  return __ddiv_rn(s, (double)P);
}

__global__ void __launch_bounds__(128) finalize_kernel(const unsigned long long* __restrict__ raw,
                                                       const unsigned long long* __restrict__ gt_area, int K,
                                                       int P, int rows_out, int64_t* __restrict__ counts,
                                                       double* __restrict__ scores) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double iou[kMaxParts];
  const unsigned long long* r = raw + (size_t)k * (P + 1) * 2;
  int64_t* c = counts + (size_t)k * rows_out * 2;
  for (int p = 0; p < rows_out; ++p) {
    const long long area = (long long)r[2 * p], inter = (long long)r[2 * p + 1];
    const long long uni = area + (long long)gt_area[p] - inter;
    c[2 * p] = inter;
    c[2 * p + 1] = uni;
    if (p < P) iou[p] = uni > 0 ? __ddiv_rn((double)inter, (double)uni) : 0.0;
  }
  scores[k] = numpy_mean(iou, P);
}

// First index with the greatest score (strict '>' in the reference loop keeps the earliest).
__global__ void __launch_bounds__(1024) argmax_kernel(const double* __restrict__ scores, int K,
                                                      int64_t* __restrict__ best) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  double bv = -1.0;
  int bi = K;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const double v = scores[i];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
  auto better = [](double v, int i, double w, int j) { return v > w || (v == w && i < j); };
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, bv, d);
    const int oi = __shfl_down_sync(0xffffffffu, bi, d);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    bv = s_val[lane]; bi = s_idx[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, bv, d);
      const int oi = __shfl_down_sync(0xffffffffu, bi, d);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { best[0] = bi < K ? bi : -1; best[1] = 0; }
  }
}

// (score bits, global index) of the local best; index -1 when the block is empty.
__global__ void best_pack_kernel(const double* __restrict__ scores, const int64_t* __restrict__ best, int64_t offset,
                                 int64_t* __restrict__ pair) {
  const int64_t b = best[0];
  pair[0] = b >= 0 ? __double_as_longlong(scores[b]) : __double_as_longlong(-1.0);
  pair[1] = b >= 0 ? b + offset : -1;
}

// Greatest score, ties -> lowest global index, over n gathered pairs (one warp).
__global__ void best_select_kernel(const int64_t* __restrict__ pairs, int n, int64_t* __restrict__ out) {
  double bv = -1.0;
  long long bi = -1;
  for (int i = threadIdx.x; i < n; i += 32) {
    const double v = __longlong_as_double(pairs[2 * i]);
    const long long idx = pairs[2 * i + 1];
    if (idx >= 0 && (bi < 0 || v > bv || (v == bv && idx < bi))) { bv = v; bi = idx; }
  }
  for (int d = 16; d > 0; d >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, bv, d);
    const long long oi = __shfl_down_sync(0xffffffffu, bi, d);
    if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
  }
  if (threadIdx.x == 0) { out[0] = __double_as_longlong(bv); out[1] = bi; }
}

__global__ void __launch_bounds__(256) resolve_rgb_kernel(const uint32_t* __restrict__ zbuf,
                                                          const uint8_t* __restrict__ pt_rgb, int64_t n_pixels,
                                                          uint8_t* __restrict__ img) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pixels;
       p += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = zbuf[p];
    uint8_t r = 0, g = 0, b = 0;
    if (key) {
      const uint8_t* c = pt_rgb + 3 * (size_t)(key - 1);
      r = c[0]; g = c[1]; b = c[2];
    }
    img[3 * p] = r; img[3 * p + 1] = g; img[3 * p + 2] = b;
  }
}

// compute_partwise_iou on two RGB images (camera_estimation.py:770-787); counts[p] = (inter, union).
__global__ void __launch_bounds__(256) partwise_counts_rgb_kernel(const uint8_t* __restrict__ proj,
                                                                  const uint8_t* __restrict__ gt, int64_t n_pixels,
                                                                  const uint8_t* __restrict__ part_rgb, int P,
                                                                  unsigned long long* __restrict__ counts) {
  __shared__ uint32_t s_col[kMaxParts];
  __shared__ unsigned int s_acc[kMaxParts * 2];
  for (int i = threadIdx.x; i < P; i += blockDim.x)
    s_col[i] = part_rgb[3 * i] | (part_rgb[3 * i + 1] << 8) | (part_rgb[3 * i + 2] << 16);
  for (int i = threadIdx.x; i < 2 * P; i += blockDim.x) s_acc[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (n_pixels + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t p = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = 0xffffffffu, b = 0xffffffffu;  // sentinel: matches no 24-bit colour
    if (p < n_pixels) {
      a = proj[3 * p] | (proj[3 * p + 1] << 8) | (proj[3 * p + 2] << 16);
      b = gt[3 * p] | (gt[3 * p + 1] << 8) | (gt[3 * p + 2] << 16);
    }
    for (int q = 0; q < P; ++q) {
      const uint32_t c = s_col[q];
      const uint32_t ma = __ballot_sync(0xffffffffu, a == c);
      const uint32_t mb = __ballot_sync(0xffffffffu, b == c);
      if (lane == 0 && (ma | mb)) {
        atomicAdd(&s_acc[2 * q], __popc(ma & mb));
        atomicAdd(&s_acc[2 * q + 1], __popc(ma | mb));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * P; i += blockDim.x)
    if (s_acc[i]) atomicAdd(counts + i, (unsigned long long)s_acc[i]);
}

// ------------------------------------------------------------------------------------------
// Depth-buffer visibility evaluator (utils/eval_helpers_intra.py:134-190).
//   depth_buffer_kernel : zbuf[pixel] = min Z over points with Z > 1e-6 that round into the image.  Positive IEEE
//                         floats order like their bit patterns, so the min is an integer atomicMin on the bits
//                         (32-bit for float cameras, 64-bit for double cameras) -- order independent.
//   part_visible_kernel : mask[pixel] = 1 when some point there has |Z - zbuf[pixel]| < eps.
// Same operation order as exact_splat, but points behind the camera are culled, not clamped.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int depth_project(const float* __restrict__ p, const T* __restrict__ cam, int W, T fW, T fH,
                                             T* zout) {
  using F = Fp<T>;
  const T d0 = F::sub((T)p[0], cam[0]), d1 = F::sub((T)p[1], cam[1]), d2 = F::sub((T)p[2], cam[2]);
  const T X = F::fma(d2, cam[5], F::fma(d1, cam[4], F::mul(d0, cam[3])));
  const T Y = F::fma(d2, cam[8], F::fma(d1, cam[7], F::mul(d0, cam[6])));
  const T Z = F::fma(d2, cam[11], F::fma(d1, cam[10], F::mul(d0, cam[9])));
  if (!(Z > (T)1e-6)) return -1;
  const T u = F::add(F::mul(F::div(X, Z), cam[12]), cam[13]);
  const T v = F::add(F::mul(-F::div(Y, Z), cam[12]), cam[14]);
  const T ur = F::rint(u), vr = F::rint(v);
  if (!(ur >= (T)0 && ur < fW && vr >= (T)0 && vr < fH)) return -1;
  *zout = Z;
  return (int)vr * W + (int)ur;
}

template <typename T, typename B>
__global__ void __launch_bounds__(256) depth_buffer_kernel(const float* __restrict__ pts, int64_t n,
                                                           const T* __restrict__ cam_g, int H, int W, B* __restrict__ zbits) {
  __shared__ T cam[16];
  if (threadIdx.x < 16) cam[threadIdx.x] = cam_g[threadIdx.x];
  __syncthreads();
  const T fW = (T)W, fH = (T)H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T z;
    const int p = depth_project<T>(pts + 3 * i, cam, W, fW, fH, &z);
    if (p < 0) continue;
    B bits;
    if (sizeof(T) == 4) bits = (B)__float_as_uint((float)z); else bits = (B)__double_as_longlong((double)z);
    if (zbits[p] > bits) atomicMin(zbits + p, bits);
  }
}

__global__ void __launch_bounds__(256) depth_fill_kernel(unsigned long long* __restrict__ z64, uint32_t* __restrict__ z32, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (z64) z64[i] = 0x7ff0000000000000ull;                  // +inf
  else z32[i] = 0x7f800000u;
}

__global__ void __launch_bounds__(256) depth_narrow_kernel(const unsigned long long* __restrict__ z64, float* __restrict__ zbuf, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) zbuf[i] = (float)__longlong_as_double((long long)z64[i]);   // float32(min Z), round to nearest
}

template <typename T>
__global__ void __launch_bounds__(256) part_visible_kernel(const float* __restrict__ pts, int64_t n,
                                                           const T* __restrict__ cam_g, const float* __restrict__ zbuf, T eps,
                                                           int H, int W, uint8_t* __restrict__ mask) {
  __shared__ T cam[16];
  if (threadIdx.x < 16) cam[threadIdx.x] = cam_g[threadIdx.x];
  __syncthreads();
  const T fW = (T)W, fH = (T)H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T z;
    const int p = depth_project<T>(pts + 3 * i, cam, W, fW, fH, &z);
    if (p < 0) continue;
    const T d = Fp<T>::sub(z, (T)zbuf[p]);
    if ((d < (T)0 ? -d : d) < eps) mask[p] = 1;
  }
}

inline int grid_for(int64_t items, int threads, int waves) {
  int64_t blocks = (items + threads - 1) / threads;
  int64_t cap = (int64_t)p3d::sm_count() * waves;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// P3D_SPLAT_EXACT=1 forces the unfiltered FP64 kernel (A/B testing of the FP32 filter); P3D_SPLAT_POINTS=1 keeps the
// per-point filtered kernel even when segments are supplied (A/B testing of the segment splat).
inline bool splat_exact_only() {
  const char* e = getenv("P3D_SPLAT_EXACT");
  return e && e[0] == '1';
}
inline bool splat_points_only() {
  static const bool on = [] { const char* e = getenv("P3D_SPLAT_POINTS"); return e && e[0] == '1'; }();
  return on;
}

}  // namespace

// Caller-owned host-side state of p3d_sweep_* (p3d_sweep_ctx_create): the helper stream and fork/join events of the
// double-buffered batches, the launch counter and the optional per-launch timing events.  One context per concurrent
// caller; the library keeps no global mutable state.
struct p3d_sweep_ctx {
  cudaStream_t helper = nullptr;
  cudaEvent_t splatted[2] = {nullptr, nullptr}, scored[2] = {nullptr, nullptr};
  int device = -1;
  int last_launches = 0;
  bool timing = false;
  std::vector<cudaEvent_t> pool;     // start/stop pairs around splat launches
  size_t used = 0;
};

namespace {

inline cudaEvent_t timing_event(p3d_sweep_ctx* ctx) {
  if (ctx->used == ctx->pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    ctx->pool.push_back(e);
  }
  return ctx->pool[ctx->used++];
}

inline int ctx_open_fork(p3d_sweep_ctx* ctx) {
  if (ctx->helper) return P3D_OK;
  P3D_CUDA(cudaGetDevice(&ctx->device));
  P3D_CUDA(cudaStreamCreateWithFlags(&ctx->helper, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    P3D_CUDA(cudaEventCreateWithFlags(&ctx->splatted[i], cudaEventDisableTiming));
    P3D_CUDA(cudaEventCreateWithFlags(&ctx->scored[i], cudaEventDisableTiming));
  }
  return P3D_OK;
}

template <typename T>
int setup_cameras(const T* cand, int K, T* cams, p3d_stream_t stream) {
  P3D_REQUIRE(K >= 0, "setup_cameras: K=%d", K);
  if (K == 0) return P3D_OK;
  P3D_REQUIRE(cand && cams, "setup_cameras: null pointer");
  setup_cameras_kernel<T><<<(K + 127) / 128, 128, 0, p3d::as_stream(stream)>>>(cand, K, cams);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// cameras handled by one CTA: enough CTAs for ~8 waves when the point list is short
#ifndef P3D_SEG_MAXCAMS
#define P3D_SEG_MAXCAMS 48
#endif
constexpr int kSegMaxCams = P3D_SEG_MAXCAMS;   // cameras per CTA of the segment splat (8: 33.9, 16: 37.3, 24: 37.9, 32: 38.7, 48: 38.9, 64: 38.3, 96: 36.8, 128: 34.5 k cand/s; 64 / 256 threads per CTA: 38.1 / 38.4)
inline int pick_cams_per_block(int64_t tiles, int K, int max_cams = 64) {
  const int64_t want = (int64_t)p3d::sm_count() * 24;
  int groups = (int)((want + tiles - 1) / (tiles > 0 ? tiles : 1));
  if (groups < 1) groups = 1;
  if (groups > K) groups = K;
  int cpb = (K + groups - 1) / groups;
  if (cpb < 4) cpb = K < 4 ? K : 4;
  if (cpb > max_cams) cpb = max_cams;                      // 192 B of shared memory per camera: keep 8 CTAs per SM
  return cpb;
}

int points_bbox(const float* pts, int64_t n, float* bbox, p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && bbox, "points_bbox: bad arguments");
  cudaStream_t st = p3d::as_stream(stream);
  bbox_init_kernel<<<1, 32, 0, st>>>(bbox);
  if (n > 0) {
    P3D_REQUIRE(pts, "points_bbox: null points");
    bbox_kernel<<<grid_for(n, 256, 4), 256, 0, st>>>(pts, n, bbox);
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

template <typename T>
int fast_cameras(const T* cams, int K, const float* bbox, int H, int W, float* fast, p3d_stream_t stream,
                 int4* rect = nullptr, uint32_t* flags = nullptr) {
  P3D_REQUIRE(K >= 0 && H > 0 && W > 0, "fast_cameras: bad arguments");
  if (K == 0) return P3D_OK;
  P3D_REQUIRE(cams && bbox && fast, "fast_cameras: null pointer");
  fast_cams_kernel<T><<<(K + 63) / 64, 64, 0, p3d::as_stream(stream)>>>(cams, K, bbox, H, W, fast, rect, flags);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// the segment splat applies when segments are supplied, the FP32 filter is on, the pixel index fits the FP32 mantissa
// trick (H W <= 2^22) and the launch's z-buffer set stays below 4 GiB (32-bit byte offsets)
inline bool use_segments(const void* segs, int64_t n_seg, bool filtered, int K, int H, int W) {
  return segs != nullptr && n_seg > 0 && filtered && !splat_points_only() && (int64_t)H * W <= (1ll << 22) &&
         (int64_t)K * H * W * 4 < (1ll << 32);   // (keys: the caller checks n < 2^32 - 2^10, so key0 + 7 step cannot wrap)
}

template <typename T>
int splat(const float* pts, const uint8_t* pt_label, int64_t n, const T* cams, int K, int H, int W, int mode,
          uint32_t* zbuf, const float* fast, const float* bbox, p3d_stream_t stream, const uint4* segs = nullptr,
          int64_t n_seg = 0, const uint32_t* cam_flags = nullptr) {
  P3D_REQUIRE(n >= 0 && K >= 0 && H > 0 && W > 0, "splat: n=%lld K=%d H=%d W=%d", (long long)n, K, H, W);
  P3D_REQUIRE(mode == P3D_MODE_JOINT || mode == P3D_MODE_PER_PART || mode == kModeJointPacked, "splat: mode=%d", mode);
  P3D_REQUIRE(n < 0xffffffffll, "splat: n=%lld does not fit 32-bit keys", (long long)n);
  P3D_REQUIRE(mode != kModeJointPacked || n < (1ll << (32 - kLabelBits)) - 1, "splat: n=%lld does not fit packed keys", (long long)n);
  P3D_REQUIRE((int64_t)H * W < (1ll << 31), "splat: image too large");
  if (n == 0 || K == 0) return P3D_OK;
  P3D_REQUIRE(pts && cams && zbuf, "splat: null pointer");
  P3D_REQUIRE(mode == P3D_MODE_JOINT || pt_label, "splat: this mode needs pt_label");
  const bool filtered = fast != nullptr && bbox != nullptr && !splat_exact_only();
  cudaStream_t st = p3d::as_stream(stream);
  if (use_segments(segs, n_seg, filtered, K, H, W) && n < 0xfffffc00ll) {
    P3D_REQUIRE(pt_label, "splat: the segment path needs pt_label");
    const int64_t tiles = (n_seg + kSegThreads - 1) / kSegThreads;
    P3D_REQUIRE(tiles < (1ll << 31), "splat: too many tiles");
    const int cpb = pick_cams_per_block(tiles, K, kSegMaxCams);
    dim3 grid((unsigned)tiles, (unsigned)((K + cpb - 1) / cpb));
    const size_t smem = (size_t)cpb * 16 * (sizeof(T) + sizeof(float)) + (size_t)(kSegThreads / 32) * (kSegQueueCap * sizeof(uint2));
    if (mode == P3D_MODE_JOINT)
      splat_seg_kernel<T, P3D_MODE_JOINT><<<grid, kSegThreads, smem, st>>>(segs, n_seg, pts, pt_label, cams, K, cpb, H, W, zbuf, fast, bbox, cam_flags);
    else if (mode == kModeJointPacked)
      splat_seg_kernel<T, kModeJointPacked><<<grid, kSegThreads, smem, st>>>(segs, n_seg, pts, pt_label, cams, K, cpb, H, W, zbuf, fast, bbox, cam_flags);
    else
      splat_seg_kernel<T, P3D_MODE_PER_PART><<<grid, kSegThreads, smem, st>>>(segs, n_seg, pts, pt_label, cams, K, cpb, H, W, zbuf, fast, bbox, cam_flags);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
  }
  const int ppt = filtered ? kPptF : kPpt;
  const int64_t tiles = (n + kSplatThreads * ppt - 1) / (kSplatThreads * ppt);
  P3D_REQUIRE(tiles < (1ll << 31), "splat: too many tiles");
  const int cpb = pick_cams_per_block(tiles, K);
  dim3 grid((unsigned)tiles, (unsigned)((K + cpb - 1) / cpb));
  if (filtered) {
    const size_t smem = (size_t)cpb * (16 * sizeof(T) + kPackCamFloats * sizeof(float)) +
                        (size_t)(kSplatThreads / 32) * (kQueueCap * sizeof(uint2));
    const T* dc = cams;
    if (mode == P3D_MODE_JOINT)
      splat_filtered_kernel<T, P3D_MODE_JOINT><<<grid, kSplatThreads, smem, st>>>(pts, pt_label, n, dc, K, cpb, H, W, zbuf, fast, bbox);
    else if (mode == kModeJointPacked)
      splat_filtered_kernel<T, kModeJointPacked><<<grid, kSplatThreads, smem, st>>>(pts, pt_label, n, dc, K, cpb, H, W, zbuf, fast, bbox);
    else
      splat_filtered_kernel<T, P3D_MODE_PER_PART><<<grid, kSplatThreads, smem, st>>>(pts, pt_label, n, dc, K, cpb, H, W, zbuf, fast, bbox);
  } else {
    const size_t smem = (size_t)cpb * 16 * sizeof(T);
    if (mode == P3D_MODE_JOINT)
      splat_kernel<T, P3D_MODE_JOINT><<<grid, kSplatThreads, smem, st>>>(pts, pt_label, n, cams, K, cpb, H, W, zbuf);
    else if (mode == kModeJointPacked)
      splat_kernel<T, kModeJointPacked><<<grid, kSplatThreads, smem, st>>>(pts, pt_label, n, cams, K, cpb, H, W, zbuf);
    else
      splat_kernel<T, P3D_MODE_PER_PART><<<grid, kSplatThreads, smem, st>>>(pts, pt_label, n, cams, K, cpb, H, W, zbuf);
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// Cameras per splat launch.  Lower bound `batch_floor`: as many as fit a 128 MB z-buffer budget (P3D_ZBUF_BUDGET_MB), but
// never fewer than 16 while that stays under 1 GiB -- a launch streams the whole point list once, so few cameras per
// launch cost more than an L2-overflowing z-buffer does (1024^3 / 2048^2: 8 cameras 3.5 k cand/s, 16: 4.0 k, 32: 3.8 k).
// Since the batches are double-buffered (the z-buffers alternate and are not L2-resident across launches any more)
// longer launches win: the batch grows until a launch covers about 2.8e9 point-cameras (3-4 ms), up to a 512 MB
// z-buffer set (P3D_ZBUF_CAP_MB).  Measured with this rule at 1024^2: 512^3 32 cameras 34.1 k cand/s, 64: 34.9 k,
// 128: 35.2 k; 256^3 32: 154.9 k, 64: 163.1 k, 128: 172.3 k; at 1024^3 / 2048^2 the rule keeps 16 (32: -4 %, 64: -9 %).
inline int clamp_batch(int64_t c, int K) {
  static const int max_batch = [] { const char* e = getenv("P3D_MAX_BATCH"); int v = e ? atoi(e) : 256; return v > 0 ? v : 256; }();
  if (c > max_batch) c = max_batch;
  if (c > K) c = K;
  if (c < 1) c = 1;
  return (int)c;
}

inline size_t env_mb(const char* name, long dflt) {
  const char* e = getenv(name);
  long mb = e ? atol(e) : dflt;
  if (mb < 1) mb = dflt;
  return (size_t)mb << 20;
}

inline int batch_floor(int K, int H, int W) {
  static const size_t budget = env_mb("P3D_ZBUF_BUDGET_MB", 128);
  const size_t per = (size_t)H * W * sizeof(uint32_t);
  int64_t c = (int64_t)(budget / per);
  int64_t floor16 = (int64_t)(((size_t)1 << 30) / per);
  if (floor16 > 16) floor16 = 16;
  if (c < floor16) c = floor16;
  return clamp_batch(c, K);
}

// capacity of one z-buffer set in the workspace (independent of the point count, which the workspace query does not know)
inline int batch_capacity(int K, int H, int W) {
  static const size_t cap = env_mb("P3D_ZBUF_CAP_MB", 512);
  const size_t per = (size_t)H * W * sizeof(uint32_t);
  const int lo = batch_floor(K, H, W);
  const int hi = clamp_batch((int64_t)(cap / per), K);
  return hi > lo ? hi : lo;
}

// cameras per launch of one sweep over n points
inline int batch_cameras(int K, int H, int W, int64_t n) {
  const int lo = batch_floor(K, H, W), hi = batch_capacity(K, H, W);
  const int64_t want = n > 0 ? (int64_t)((2.8e9 + (double)n - 1.0) / (double)n) : hi;
  int64_t c = want < lo ? lo : want;
  if (c > hi) c = hi;
  return (int)c;
}

struct SweepLayout {
  size_t cams, raw, gt_area, bbox, fast, rect, flags, zbuf, total;
  int batch;      // cameras one z-buffer set can hold
  int floor;      // fewest cameras per launch the sweep will choose
  int zbufs;      // 2 = double-buffered: the score pass of one batch runs beside the splat of the next
};

// P3D_OVERLAP=0 keeps splat and score of every batch in sequence on the caller's stream (A/B runs)
inline bool overlap_enabled() {
  static const bool on = [] { const char* e = getenv("P3D_OVERLAP"); return e == nullptr || atoi(e) != 0; }();
  return on;
}

inline SweepLayout sweep_layout(int K, int H, int W, int P, int elem_bytes) {
  SweepLayout L;
  L.batch = batch_capacity(K, H, W);                       // capacity of a set; the sweep may use fewer per launch
  L.floor = batch_floor(K, H, W);
  size_t off = 0;
  L.cams = off; off = p3d_align_up(off + (size_t)K * 16 * elem_bytes, 256);
  L.raw = off; off = p3d_align_up(off + (size_t)K * (P + 1) * 2 * sizeof(unsigned long long), 256);
  L.gt_area = off; off = p3d_align_up(off + (size_t)(P + 1) * sizeof(unsigned long long), 256);
  L.bbox = off; off = p3d_align_up(off + 8 * sizeof(float), 256);
  L.fast = off; off = p3d_align_up(off + (size_t)K * sizeof(FastCam), 256);
  L.rect = off; off = p3d_align_up(off + (size_t)K * sizeof(int4), 256);
  L.flags = off; off = p3d_align_up(off + (size_t)K * sizeof(uint32_t), 256);
  L.zbufs = (overlap_enabled() && K > L.floor) ? 2 : 1;
  L.zbuf = off; off = p3d_align_up(off + (size_t)L.zbufs * L.batch * H * W * sizeof(uint32_t), 256);
  L.total = off;
  return L;
}

template <typename T>
int sweep(const float* pts, const uint8_t* pt_label, int64_t n, const uint4* segs, int64_t n_seg, const T* cand, int K,
          const uint8_t* gt_label, const uint8_t* gt_any, int H, int W, int P, int mode, int64_t* counts, double* scores,
          int64_t* best, void* workspace, size_t workspace_bytes, p3d_sweep_ctx* ctx, p3d_stream_t stream) {
  int launches = 0;
  if (ctx) ctx->last_launches = 0;
  P3D_REQUIRE(K > 0 && H > 0 && W > 0 && n >= 0, "sweep: K=%d H=%d W=%d n=%lld", K, H, W, (long long)n);
  P3D_REQUIRE(P >= 1 && P <= kMaxParts, "sweep: P=%d (1..%d)", P, kMaxParts);
  P3D_REQUIRE(mode == P3D_MODE_JOINT || mode == P3D_MODE_PER_PART, "sweep: mode=%d", mode);
  P3D_REQUIRE((int64_t)H * W < (1ll << 31), "sweep: image too large");
  P3D_REQUIRE(cand && gt_label && counts && scores && workspace, "sweep: null pointer");
  P3D_REQUIRE(n == 0 || (pts && pt_label), "sweep: null points");
  P3D_REQUIRE(n_seg >= 0 && (segs == nullptr || (reinterpret_cast<uintptr_t>(segs) & 15) == 0), "sweep: bad segments");
  P3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "sweep: workspace must be 256-byte aligned");
  const SweepLayout L = sweep_layout(K, H, W, P, (int)sizeof(T));
  if (workspace_bytes < L.total) {
    p3d::set_error("sweep: workspace %zu < %zu", workspace_bytes, L.total);
    return P3D_E_WORKSPACE;
  }
  cudaStream_t st = p3d::as_stream(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  T* cams = reinterpret_cast<T*>(ws + L.cams);
  unsigned long long* raw = reinterpret_cast<unsigned long long*>(ws + L.raw);
  unsigned long long* gt_area = reinterpret_cast<unsigned long long*>(ws + L.gt_area);
  uint32_t* zbuf = reinterpret_cast<uint32_t*>(ws + L.zbuf);
  const int HW = H * W;
  const int rows = mode == P3D_MODE_PER_PART ? P + 1 : P;

  float* bbox = reinterpret_cast<float*>(ws + L.bbox);
  float* fast = reinterpret_cast<float*>(ws + L.fast);
  uint32_t* flags = reinterpret_cast<uint32_t*>(ws + L.flags);
  static const bool use_rect = [] { const char* e = getenv("P3D_SCORE_RECT"); return e == nullptr || atoi(e) != 0; }();
  int4* rect = use_rect ? reinterpret_cast<int4*>(ws + L.rect) : nullptr;   // P3D_SCORE_RECT=0: score the whole image
  P3D_CUDA(cudaMemsetAsync(ws + L.raw, 0, L.bbox - L.raw, st));   // raw + gt_area
  const int batch = batch_cameras(K, H, W, n);              // cameras per launch (<= L.batch)
  for (int sidx = 0; sidx < (K > batch ? L.zbufs : 1); ++sidx)
    P3D_CUDA(cudaMemsetAsync(zbuf + (size_t)sidx * L.batch * HW, 0, (size_t)batch * HW * sizeof(uint32_t), st));
  int rc = setup_cameras<T>(cand, K, cams, stream);
  if (rc) return rc;
  gt_area_kernel<<<grid_for(HW, 256, 4), 256, 0, st>>>(gt_label, mode == P3D_MODE_PER_PART ? gt_any : nullptr, HW, P,
                                                      gt_area);
  P3D_LAUNCH_CHECK();
  launches += 2;
  if (n > 0) {
    rc = points_bbox(pts, n, bbox, stream);
    if (rc) return rc;
    rc = fast_cameras<T>(cams, K, bbox, H, W, fast, stream, rect, flags);
    if (rc) return rc;
    launches += 3;
  }
  // joint mode: carry the label in the key whenever it fits (P3D_NO_PACKED_KEYS=1 keeps the gather, for A/B runs)
  static const bool no_packed = getenv("P3D_NO_PACKED_KEYS") != nullptr;
  const int smode = (mode == P3D_MODE_JOINT && !no_packed && n < (1ll << (32 - kLabelBits)) - 1) ? kModeJointPacked : mode;
  // Double-buffered batches: splat(b) on the caller's stream, score(b) on the context's helper stream forked and joined
  // with events, so the bandwidth-bound score/clear pass of one batch runs beside the issue-bound splat of the next.
  // Needs a context (the helper stream is caller-owned state); not used while the caller's stream is being captured
  // into a graph, for single-batch sweeps, or with P3D_OVERLAP=0.
  bool forked = false;
  if (ctx && L.zbufs == 2 && K > batch && n > 0) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
      rc = ctx_open_fork(ctx);
      if (rc) return rc;
      forked = true;
    } else {
      (void)cudaGetLastError();
    }
  }
  bool pending[2] = {false, false};                         // score pass of that buffer not yet joined into `st`
  // error path: the helper stream may still be reading the caller-owned workspace; order it before anything the
  // caller enqueues next (or frees) by joining it into `st`
  auto fail = [&](int code) {
    if (forked)
      for (int i = 0; i < 2; ++i)
        if (pending[i]) (void)cudaStreamWaitEvent(st, ctx->scored[i], 0);
    return code;
  };
#define P3D_SWEEP_CUDA(expr)                                                                                        \
  do {                                                                                                              \
    cudaError_t _e = (expr);                                                                                        \
    if (_e != cudaSuccess) {                                                                                        \
      p3d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                   \
      return fail(P3D_E_CUDA);                                                                                      \
    }                                                                                                               \
  } while (0)
  int b = 0;
  for (int k0 = 0; k0 < K; k0 += batch, ++b) {
    const int kb = K - k0 < batch ? K - k0 : batch;
    if (n > 0) {
      const int buf = forked ? (b & 1) : 0;
      uint32_t* zb = zbuf + (size_t)buf * L.batch * HW;
      if (forked && b >= 2) {                              // score(b-2) has cleared zb
        P3D_SWEEP_CUDA(cudaStreamWaitEvent(st, ctx->scored[buf], 0));
        pending[buf] = false;
      }
      cudaEvent_t ev0 = nullptr, ev1 = nullptr;
      if (ctx && ctx->timing && (ev0 = timing_event(ctx)) && (ev1 = timing_event(ctx))) P3D_SWEEP_CUDA(cudaEventRecord(ev0, st));
      rc = splat<T>(pts, pt_label, n, cams + (size_t)k0 * 16, kb, H, W, smode, zb, fast + (size_t)k0 * 16, bbox, stream,
                    segs, n_seg, flags + k0);
      if (rc) return fail(rc);
      if (ev1) P3D_SWEEP_CUDA(cudaEventRecord(ev1, st));
      cudaStream_t ss = st;
      if (forked) {
        P3D_SWEEP_CUDA(cudaEventRecord(ctx->splatted[buf], st));
        P3D_SWEEP_CUDA(cudaStreamWaitEvent(ctx->helper, ctx->splatted[buf], 0));
        ss = ctx->helper;
      }
      const int vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(gt_label) & 3) == 0) &&
                      (gt_any == nullptr || (reinterpret_cast<uintptr_t>(gt_any) & 3) == 0);
      dim3 grid((unsigned)grid_for(vec ? HW / 4 : HW, kScoreThreads, 2), (unsigned)kb);
      // spread one camera's pixels over about 32 CTAs per SM / kb (8: 33.2 k cand/s, 16: 33.5, 32: 33.5, 64: 33.5)
      static const int score_waves = [] { const char* e = getenv("P3D_SCORE_WAVES"); const int v = e ? atoi(e) : 32; return v > 0 ? v : 32; }();
      int per_cam = (p3d::sm_count() * score_waves + kb - 1) / kb;
      if ((int)grid.x > per_cam) grid.x = per_cam < 1 ? 1 : per_cam;
      const int4* rk = rect ? rect + k0 : nullptr;
      if (smode == kModeJointPacked)
        score_kernel<kModeJointPacked><<<grid, kScoreThreads, 0, ss>>>(zb, pt_label, gt_label, nullptr, HW, P,
                                                                       raw + (size_t)k0 * (P + 1) * 2, vec, W, rk);
      else if (mode == P3D_MODE_JOINT)
        score_kernel<P3D_MODE_JOINT><<<grid, kScoreThreads, 0, ss>>>(zb, pt_label, gt_label, nullptr, HW, P,
                                                                     raw + (size_t)k0 * (P + 1) * 2, vec, W, rk);
      else
        score_kernel<P3D_MODE_PER_PART><<<grid, kScoreThreads, 0, ss>>>(zb, pt_label, gt_label, gt_any, HW, P,
                                                                        raw + (size_t)k0 * (P + 1) * 2, vec, W, rk);
      P3D_SWEEP_CUDA(cudaGetLastError());
      if (forked) {
        P3D_SWEEP_CUDA(cudaEventRecord(ctx->scored[buf], ctx->helper));
        pending[buf] = true;
      }
      launches += 2;
    }
  }
  if (forked) {                                          // join: everything after this sees all score passes
    for (int buf = 0; buf < 2; ++buf)
      if (pending[buf]) {
        P3D_SWEEP_CUDA(cudaStreamWaitEvent(st, ctx->scored[buf], 0));
        pending[buf] = false;
      }
  }
#undef P3D_SWEEP_CUDA
  finalize_kernel<<<(K + 127) / 128, 128, 0, st>>>(raw, gt_area, K, P, rows, counts, scores);
  P3D_LAUNCH_CHECK();
  ++launches;
  if (best) {
    argmax_kernel<<<1, 1024, 0, st>>>(scores, K, best);
    P3D_LAUNCH_CHECK();
    ++launches;
  }
  if (ctx) ctx->last_launches = launches;
  return P3D_OK;
}

}  // namespace

P3D_API int p3d_setup_cameras_f64(const double* cand, int K, double* cams, p3d_stream_t stream) {
  return setup_cameras<double>(cand, K, cams, stream);
}
P3D_API int p3d_setup_cameras_f32(const float* cand, int K, float* cams, p3d_stream_t stream) {
  return setup_cameras<float>(cand, K, cams, stream);
}

P3D_API int p3d_points_bbox(const float* pts, int64_t n, float* bbox, p3d_stream_t stream) {
  return points_bbox(pts, n, bbox, stream);
}

P3D_API int p3d_fast_cameras_f64(const double* cams, int K, const float* bbox, int H, int W, float* fast,
                                 p3d_stream_t stream) {
  return fast_cameras<double>(cams, K, bbox, H, W, fast, stream);
}
P3D_API int p3d_fast_cameras_f32(const float* cams, int K, const float* bbox, int H, int W, float* fast,
                                 p3d_stream_t stream) {
  return fast_cameras<float>(cams, K, bbox, H, W, fast, stream);
}

P3D_API int p3d_splat_f64(const float* pts, const uint8_t* pt_label, int64_t n, const double* cams, int K, int H,
                          int W, int mode, uint32_t* zbuf, const float* fast, const float* bbox,
                          p3d_stream_t stream) {
  P3D_REQUIRE(mode == P3D_MODE_JOINT || mode == P3D_MODE_PER_PART, "splat: mode=%d", mode);
  return splat<double>(pts, pt_label, n, cams, K, H, W, mode, zbuf, fast, bbox, stream);
}
P3D_API int p3d_splat_f32(const float* pts, const uint8_t* pt_label, int64_t n, const float* cams, int K, int H,
                          int W, int mode, uint32_t* zbuf, const float* fast, const float* bbox,
                          p3d_stream_t stream) {
  P3D_REQUIRE(mode == P3D_MODE_JOINT || mode == P3D_MODE_PER_PART, "splat: mode=%d", mode);
  return splat<float>(pts, pt_label, n, cams, K, H, W, mode, zbuf, fast, bbox, stream);
}

P3D_API int p3d_resolve_rgb(const uint32_t* zbuf, const uint8_t* pt_rgb, int64_t n_pixels, uint8_t* img,
                            p3d_stream_t stream) {
  P3D_REQUIRE(n_pixels >= 0, "resolve_rgb: n_pixels=%lld", (long long)n_pixels);
  if (n_pixels == 0) return P3D_OK;
  P3D_REQUIRE(zbuf && img, "resolve_rgb: null pointer");
  resolve_rgb_kernel<<<grid_for(n_pixels, 256, 8), 256, 0, p3d::as_stream(stream)>>>(zbuf, pt_rgb, n_pixels, img);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_partwise_counts_rgb(const uint8_t* proj_rgb, const uint8_t* gt_rgb, int64_t n_pixels,
                                    const uint8_t* part_rgb, int P, int64_t* counts, p3d_stream_t stream) {
  P3D_REQUIRE(n_pixels >= 0 && P >= 0 && P <= kMaxParts, "partwise_counts_rgb: n=%lld P=%d", (long long)n_pixels, P);
  if (P == 0) return P3D_OK;
  P3D_REQUIRE(counts && part_rgb, "partwise_counts_rgb: null pointer");
  cudaStream_t st = p3d::as_stream(stream);
  P3D_CUDA(cudaMemsetAsync(counts, 0, (size_t)P * 2 * sizeof(int64_t), st));
  if (n_pixels == 0) return P3D_OK;
  P3D_REQUIRE(proj_rgb && gt_rgb, "partwise_counts_rgb: null image");
  partwise_counts_rgb_kernel<<<grid_for(n_pixels, 256, 4), 256, 0, st>>>(
      proj_rgb, gt_rgb, n_pixels, part_rgb, P, reinterpret_cast<unsigned long long*>(counts));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API size_t p3d_sweep_workspace_bytes(int K, int H, int W, int P, int elem_bytes) {
  if (K <= 0 || H <= 0 || W <= 0 || P <= 0 || (elem_bytes != 4 && elem_bytes != 8)) return 0;
  return sweep_layout(K, H, W, P, elem_bytes).total;
}

P3D_API int p3d_sweep_f64(const float* pts, const uint8_t* pt_label, int64_t n, const uint32_t* segs, int64_t n_seg,
                          const double* cand, int K, const uint8_t* gt_label, const uint8_t* gt_any, int H, int W, int P,
                          int mode, int64_t* counts, double* scores, int64_t* best, void* workspace,
                          size_t workspace_bytes, p3d_sweep_ctx* ctx, p3d_stream_t stream) {
  return sweep<double>(pts, pt_label, n, reinterpret_cast<const uint4*>(segs), n_seg, cand, K, gt_label, gt_any, H, W, P,
                       mode, counts, scores, best, workspace, workspace_bytes, ctx, stream);
}
P3D_API int p3d_sweep_f32(const float* pts, const uint8_t* pt_label, int64_t n, const uint32_t* segs, int64_t n_seg,
                          const float* cand, int K, const uint8_t* gt_label, const uint8_t* gt_any, int H, int W, int P,
                          int mode, int64_t* counts, double* scores, int64_t* best, void* workspace,
                          size_t workspace_bytes, p3d_sweep_ctx* ctx, p3d_stream_t stream) {
  return sweep<float>(pts, pt_label, n, reinterpret_cast<const uint4*>(segs), n_seg, cand, K, gt_label, gt_any, H, W, P,
                      mode, counts, scores, best, workspace, workspace_bytes, ctx, stream);
}

P3D_API int p3d_segment_length(void) { return kSegLen; }

P3D_API p3d_sweep_ctx* p3d_sweep_ctx_create(void) { return new (std::nothrow) p3d_sweep_ctx(); }

P3D_API void p3d_sweep_ctx_destroy(p3d_sweep_ctx* ctx) {
  if (!ctx) return;
  for (cudaEvent_t e : ctx->pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (ctx->splatted[i]) cudaEventDestroy(ctx->splatted[i]);
    if (ctx->scored[i]) cudaEventDestroy(ctx->scored[i]);
  }
  if (ctx->helper) cudaStreamDestroy(ctx->helper);
  delete ctx;
}

P3D_API int p3d_sweep_ctx_launches(const p3d_sweep_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

P3D_API int p3d_sweep_ctx_timing(p3d_sweep_ctx* ctx, int on) {
  P3D_REQUIRE(ctx, "sweep_ctx_timing: null context");
  ctx->timing = on != 0;
  ctx->used = 0;
  return P3D_OK;
}

P3D_API int p3d_sweep_ctx_timing_read(p3d_sweep_ctx* ctx, double* splat_ms, int* n_launches) {
  P3D_REQUIRE(ctx, "sweep_ctx_timing_read: null context");
  double total = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < ctx->used; i += 2) {
    P3D_CUDA(cudaEventSynchronize(ctx->pool[i + 1]));
    float ms = 0.f;
    P3D_CUDA(cudaEventElapsedTime(&ms, ctx->pool[i], ctx->pool[i + 1]));
    total += ms;
    ++n;
  }
  ctx->used = 0;
  if (splat_ms) *splat_ms = total;
  if (n_launches) *n_launches = n;
  return P3D_OK;
}

P3D_API size_t p3d_depth_workspace_bytes(int H, int W, int elem_bytes) {
  if (H <= 0 || W <= 0) return 0;
  return elem_bytes == 8 ? (size_t)H * W * sizeof(unsigned long long) : 0;
}

template <typename T>
static int depth_buffer(const float* pts, int64_t n, const T* cam, int H, int W, float* zbuf, void* ws, size_t ws_bytes,
                        p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && H > 0 && W > 0 && (int64_t)H * W < (1ll << 31), "depth_buffer: bad arguments");
  P3D_REQUIRE(cam && zbuf && (pts || n == 0), "depth_buffer: null pointer");
  cudaStream_t st = p3d::as_stream(stream);
  const int HW = H * W;
  if (sizeof(T) == 8) {
    if (ws == nullptr || ws_bytes < (size_t)HW * 8) { p3d::set_error("depth_buffer: workspace too small"); return P3D_E_WORKSPACE; }
    unsigned long long* z64 = static_cast<unsigned long long*>(ws);
    depth_fill_kernel<<<(HW + 255) / 256, 256, 0, st>>>(z64, nullptr, HW);
    if (n) depth_buffer_kernel<T, unsigned long long><<<grid_for(n, 256, 8), 256, 0, st>>>(pts, n, cam, H, W, z64);
    depth_narrow_kernel<<<(HW + 255) / 256, 256, 0, st>>>(z64, zbuf, HW);
  } else {
    uint32_t* z32 = reinterpret_cast<uint32_t*>(zbuf);
    depth_fill_kernel<<<(HW + 255) / 256, 256, 0, st>>>(nullptr, z32, HW);
    if (n) depth_buffer_kernel<T, uint32_t><<<grid_for(n, 256, 8), 256, 0, st>>>(pts, n, cam, H, W, z32);
  }
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

template <typename T>
static int part_visible(const float* pts, int64_t n, const T* cam, const float* zbuf, T eps, int H, int W, uint8_t* mask,
                        p3d_stream_t stream) {
  P3D_REQUIRE(n >= 0 && H > 0 && W > 0 && (int64_t)H * W < (1ll << 31), "part_visible: bad arguments");
  P3D_REQUIRE(cam && zbuf && mask && (pts || n == 0), "part_visible: null pointer");
  cudaStream_t st = p3d::as_stream(stream);
  P3D_CUDA(cudaMemsetAsync(mask, 0, (size_t)H * W, st));
  if (n) part_visible_kernel<T><<<grid_for(n, 256, 8), 256, 0, st>>>(pts, n, cam, zbuf, eps, H, W, mask);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_depth_buffer_f32(const float* pts, int64_t n, const float* cam, int H, int W, float* zbuf, void* ws,
                                 size_t ws_bytes, p3d_stream_t stream) {
  return depth_buffer<float>(pts, n, cam, H, W, zbuf, ws, ws_bytes, stream);
}
P3D_API int p3d_depth_buffer_f64(const float* pts, int64_t n, const double* cam, int H, int W, float* zbuf, void* ws,
                                 size_t ws_bytes, p3d_stream_t stream) {
  return depth_buffer<double>(pts, n, cam, H, W, zbuf, ws, ws_bytes, stream);
}
P3D_API int p3d_part_visible_f32(const float* pts, int64_t n, const float* cam, const float* zbuf, float eps, int H, int W,
                                 uint8_t* mask, p3d_stream_t stream) {
  return part_visible<float>(pts, n, cam, zbuf, eps, H, W, mask, stream);
}
P3D_API int p3d_part_visible_f64(const float* pts, int64_t n, const double* cam, const float* zbuf, double eps, int H, int W,
                                 uint8_t* mask, p3d_stream_t stream) {
  return part_visible<double>(pts, n, cam, zbuf, eps, H, W, mask, stream);
}

P3D_API int p3d_best_pack(const double* scores, const int64_t* best, int64_t offset, int64_t* pair,
                          p3d_stream_t stream) {
  P3D_REQUIRE(scores && best && pair, "best_pack: null pointer");
  best_pack_kernel<<<1, 1, 0, p3d::as_stream(stream)>>>(scores, best, offset, pair);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

P3D_API int p3d_best_select(const int64_t* pairs, int n, int64_t* out, p3d_stream_t stream) {
  P3D_REQUIRE(pairs && out && n >= 1, "best_select: bad arguments");
  best_select_kernel<<<1, 32, 0, p3d::as_stream(stream)>>>(pairs, n, out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}
