// Device-side projection helpers shared by p3d_camera.cu and p3d_deform.cu (internal header, sm_100a only).
//
//   Fp<T>            explicit round-to-nearest arithmetic in the reference's operation order
//   exact_splat      project_colored_voxels for one point/camera, operation by operation (projection_utils.py:5-21)
//   FastCam          FP32 companion of an f64 camera block + proven rounding thresholds (see p3d_camera.cu)
//   f32x2 helpers    packed FP32x2 arithmetic (FFMA2 / FADD2)
#pragma once
#include "p3d_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// FP helpers: explicit rounding, no contraction.
// ------------------------------------------------------------------------------------------
template <typename T> struct Fp;
template <> struct Fp<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
  // round half to even for |x| < 2^51; larger magnitudes stay far outside any image
  static __device__ __forceinline__ double rint(double x) {
    const double m = 6755399441055744.0;  // 1.5 * 2^52
    return __dsub_rn(__dadd_rn(x, m), m);
  }
  // OpenBLAS ddot tail loop, contracted by the compiler: FMA chain
  static __device__ __forceinline__ double dot3(const double* a, const double* b) {
    return __fma_rn(a[2], b[2], __fma_rn(a[1], b[1], __dmul_rn(a[0], b[0])));
  }
  static __device__ __forceinline__ double eps() { return 1e-8; }
};
template <> struct Fp<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ float rint(float x) { return rintf(x); }
  // OpenBLAS sdot tail loop: float products, double accumulator, one final rounding
  static __device__ __forceinline__ float dot3(const float* a, const float* b) {
    float p0 = __fmul_rn(a[0], b[0]), p1 = __fmul_rn(a[1], b[1]), p2 = __fmul_rn(a[2], b[2]);
    double s = __dadd_rn((double)p0, (double)p1);
    s = __dadd_rn(s, (double)p2);
    return (float)s;
  }
  static __device__ __forceinline__ float eps() { return (float)1e-8; }
};

template <int MODE>
__device__ __forceinline__ void zbuf_update(uint32_t* p, uint32_t key) {
  const uint32_t cur = __ldcg(p);
  if (MODE != P3D_MODE_PER_PART) {
    if (cur < key) atomicMax(p, key);
  } else {
    if ((cur & key) == 0) atomicOr(p, key);
  }
}

// The reference's projection of one point through one camera, operation by operation (see file header).
template <typename T, int MODE>
__device__ __forceinline__ void exact_splat(T px, T py, T pz, uint32_t key, const T* __restrict__ cam,
                                            uint32_t* __restrict__ zb, int W, T fW, T fH) {
  using F = Fp<T>;
  const T d0 = F::sub(px, cam[0]), d1 = F::sub(py, cam[1]), d2 = F::sub(pz, cam[2]);
  const T X = F::fma(d2, cam[5], F::fma(d1, cam[4], F::mul(d0, cam[3])));
  const T Y = F::fma(d2, cam[8], F::fma(d1, cam[7], F::mul(d0, cam[6])));
  T Z = F::fma(d2, cam[11], F::fma(d1, cam[10], F::mul(d0, cam[9])));
  if (Z < F::eps()) Z = F::eps();
  const T u = F::add(F::mul(F::div(X, Z), cam[12]), cam[13]);
  const T v = F::add(F::mul(-F::div(Y, Z), cam[12]), cam[14]);
  const T ur = F::rint(u), vr = F::rint(v);
  if (key != 0 && ur >= (T)0 && ur < fW && vr >= (T)0 && vr < fH)
    zbuf_update<MODE>(zb + ((size_t)(int)vr * W + (size_t)(int)ur), key);
}

// The same sequence without the store: returns whether the point lands inside the image and its pixel index.
template <typename T>
__device__ __forceinline__ bool exact_pixel(T px, T py, T pz, const T* __restrict__ cam, int W, T fW, T fH, uint32_t& pix) {
  using F = Fp<T>;
  const T d0 = F::sub(px, cam[0]), d1 = F::sub(py, cam[1]), d2 = F::sub(pz, cam[2]);
  const T X = F::fma(d2, cam[5], F::fma(d1, cam[4], F::mul(d0, cam[3])));
  const T Y = F::fma(d2, cam[8], F::fma(d1, cam[7], F::mul(d0, cam[6])));
  T Z = F::fma(d2, cam[11], F::fma(d1, cam[10], F::mul(d0, cam[9])));
  if (Z < F::eps()) Z = F::eps();
  const T u = F::add(F::mul(F::div(X, Z), cam[12]), cam[13]);
  const T v = F::add(F::mul(-F::div(Y, Z), cam[12]), cam[14]);
  const T ur = F::rint(u), vr = F::rint(v);
  const bool hit = ur >= (T)0 && ur < fW && vr >= (T)0 && vr < fH;
  pix = hit ? (uint32_t)(int)vr * (uint32_t)W + (uint32_t)(int)ur : 0u;
  return hit;
}

struct FastCam {                     // 16 floats per camera
  float A[3], TA, B[3], TB, C[3], TC, cx, cy, thr_u, thr_v;
};

// centre of the bounding box used by the FP32 filter; the same expression in fast_cams_kernel and the splat
__device__ __forceinline__ float bbox_centre(const float* __restrict__ bbox, int k) {
  return __fmul_rn(0.5f, __fadd_rn(bbox[k], bbox[3 + k]));
}

// ref_f32: the exact path is the reference's FLOAT32 sequence (float32 camera arrays) instead of the float64 one.  Its
// own rounding error against the real-arithmetic value of the same (float32-valued) camera block is added to the bound:
//   d = fl(p - e) (1 rounding), X = fma(d2,R2, fma(d1,R1, fl(d0 R0))) (3 roundings)  =>  |X32 - X| <= 4 eps Md,
//   Md = sum_k max|p_k - e_k| over the box; the same for Z;  u = fl(fl(fl(X/Z) f) + cx)  =>
//   |u_ref32 - u| <= eps [ 4 f Md / Zmin + |u - cx| (4 Md / Zmin + 2) + |u| ]            (Zmin lowered by 4 eps Md)
__device__ __forceinline__ void make_fast_cam(const double* __restrict__ cam, const float* __restrict__ bbox, int H,
                                              int W, FastCam* out, bool ref_f32 = false) {
  FastCam fc;
  bool ok = true;
  for (int k = 0; k < 15; ++k) ok = ok && (fabs(cam[k]) < 1e30);                 // false for NaN / Inf
  for (int k = 3; k < 12; ++k) ok = ok && (fabs(cam[k]) <= 1.0001);
  const double f = cam[12];
  double sp = 0.0, zmin = 0.0, zabs = 0.0, ta = 0.0, tb = 0.0, tc = 0.0, md = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double lo = (double)bbox[k], hi = (double)bbox[3 + k], c = (double)bbox_centre(bbox, k);
    ok = ok && (fabs(lo) < 1e30) && (fabs(hi) < 1e30) && lo <= hi;
    sp += fmax(fabs(lo - c), fabs(hi - c));
    md += fmax(fabs(lo - cam[k]), fabs(hi - cam[k]));
    const double a = (lo - cam[k]) * cam[9 + k], b = (hi - cam[k]) * cam[9 + k];
    zmin += fmin(a, b);
    zabs += fmax(fabs(a), fabs(b));
    fc.A[k] = (float)(f * cam[3 + k]);
    fc.B[k] = (float)(-f * cam[6 + k]);
    fc.C[k] = (float)cam[9 + k];
    ta += f * (c - cam[k]) * cam[3 + k];
    tb -= f * (c - cam[k]) * cam[6 + k];
    tc += (c - cam[k]) * cam[9 + k];
  }
  fc.TA = (float)ta; fc.TB = (float)tb; fc.TC = (float)tc;
  fc.cx = (float)cam[13]; fc.cy = (float)cam[14];
  const double eps = 5.9604644775390625e-08;                                    // 2^-24
  sp *= 1.0000002;                                                              // |q| <= (1 + eps) |p - c|
  const double dz = 5.05 * sp + 4.04 * fabs(tc);
  const double dxu = 5.05 * f * sp + 4.04 * fabs(ta), dxv = 5.05 * f * sp + 4.04 * fabs(tb);
  zmin -= 1e-9 * zabs;
  const double refx = ref_f32 ? 4.05 * md : 0.0;                                 // rounding of the float32 reference chain
  if (ref_f32) zmin -= eps * refx;
  ok = ok && f > 1e-3 && zmin > fmax(1e-3, (dz + refx) * 7.62939453125e-06);     // 2^-17
  double bu = 1.0, bv = 1.0, b1 = 1.0;
  if (ok) {
    const double cx = cam[13], cy = cam[14], dW = (double)W, dH = (double)H;
    const double ud = fmax(fabs(cx + 2.0), fabs(dW + 2.0 - cx)) + 1.0, vd = fmax(fabs(cy + 2.0), fabs(dH + 2.0 - cy)) + 1.0;
    b1 = 1.25 * eps * (dz / zmin + 3.0);
    bu = 1.25 * eps * (dxu / zmin + ud * (dz / zmin + 2.0) + fabs(cx) + dW + 2.0) + 1e-7;
    bv = 1.25 * eps * (dxv / zmin + vd * (dz / zmin + 2.0) + fabs(cy) + dH + 2.0) + 1e-7;
    if (ref_f32) {
      b1 += 1.25 * eps * (refx / zmin + 3.0);
      bu += 1.25 * eps * (f * refx / zmin + ud * (refx / zmin + 2.02) + dW + 2.0);
      bv += 1.25 * eps * (f * refx / zmin + vd * (refx / zmin + 2.02) + dH + 2.0);
    }
  }
  ok = ok && bu <= 0.25 && bv <= 0.25 && b1 * 2.0 * (double)(W > H ? W : H) <= 0.25;
  ok = ok && W < (1 << 21) && H < (1 << 21);                                     // range of the magic-number rounding
  fc.thr_u = ok ? __double2float_rd(0.5 - bu) : -1.f;
  fc.thr_v = ok ? __double2float_rd(0.5 - bv) : -1.f;
  *out = fc;
}

// Image rectangle that contains every pixel a camera can touch: rect = (row0, row1, col0, col1), inclusive; row1 < row0
// when nothing can land in the image.  For a camera that passed the filter's test (the whole bounding box lies in front
// of the camera plane, Z >= zmin > 0, and |u32 - u_ref| <= 0.25) the perspective image of the box is the convex hull of
// its 8 projected corners, so the reference's u, v of every point lie inside the corners' min/max; 2.5 px of margin
// cover the reference's own rounding and the half-even pixel rounding.  Any other camera gets the full image.  The
// score pass reads and clears only this rectangle of the camera's z-buffer (nothing outside it is ever written).
// Returns true when the unclamped rectangle lies inside the image: then every decided point of the camera is in range
// and the segment splat skips the bounds test (kCamInView).
__device__ __forceinline__ bool footprint_rect(const double* __restrict__ cam, const float* __restrict__ bbox, int H, int W,
                                               bool fast_ok, int4* out) {
  int4 r = make_int4(0, H - 1, 0, W - 1);
  bool inside = false;
  if (fast_ok) {
    double umin = 1e300, umax = -1e300, vmin = 1e300, vmax = -1e300;
    bool fin = true;
    for (int corner = 0; corner < 8; ++corner) {
      double d[3];
      for (int k = 0; k < 3; ++k) d[k] = (double)bbox[((corner >> k) & 1) ? 3 + k : k] - cam[k];
      const double X = d[0] * cam[3] + d[1] * cam[4] + d[2] * cam[5];
      const double Y = d[0] * cam[6] + d[1] * cam[7] + d[2] * cam[8];
      const double Z = d[0] * cam[9] + d[1] * cam[10] + d[2] * cam[11];
      const double u = X / Z * cam[12] + cam[13], v = -(Y / Z) * cam[12] + cam[14];
      fin = fin && Z > 0.0 && fabs(u) < 1e15 && fabs(v) < 1e15;                 // false for NaN
      umin = fmin(umin, u); umax = fmax(umax, u);
      vmin = fmin(vmin, v); vmax = fmax(vmax, v);
    }
    if (fin) {
      const double m = 2.5;
      const double r0 = fmax(0.0, floor(vmin - m)), r1 = fmin((double)(H - 1), ceil(vmax + m));
      const double c0 = fmax(0.0, floor(umin - m)), c1 = fmin((double)(W - 1), ceil(umax + m));
      if (r0 > r1 || c0 > c1) r = make_int4(0, -1, 0, -1);
      else r = make_int4((int)r0, (int)r1, (int)c0, (int)c1);
      inside = floor(vmin - m) >= 0.0 && ceil(vmax + m) <= (double)(H - 1) && floor(umin - m) >= 0.0 &&
               ceil(umax + m) <= (double)(W - 1);
    }
  }
  if (out) *out = r;
  return inside;
}

// Packed FP32x2 arithmetic (Blackwell FFMA2 / FADD2): one issue slot for two lanes' worth of IEEE-rn FP32 operations.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

}  // namespace
