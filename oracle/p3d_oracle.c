/*
 * p3d_oracle.c -- CPU restatement of the reference's geometry hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * kernels under part-based-3d-reconstruction_b200/csrc/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path never calls into it.
 *
 * The reference (BarnitaSharma/Part-based-3D-Reconstruction) is pure
 * Python/NumPy/SciPy; the arithmetic it relies on lives in numpy 2.3.5
 * (OpenBLAS 0.3.30 matmul/dot), scipy 1.18.1 (ndimage.affine_transform,
 * ndimage.label).  Each function below restates one of those steps as a plain
 * scalar loop with a fixed operation order and cites the reference call site.
 * Pinning: tests/test_oracle_golden.py checks every function here against
 * fixtures produced by the live reference (tests/golden/make_golden.py).
 *
 * Build:  gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off matters: every multiply/add below is separately rounded
 * unless it is written as an explicit fma()/fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- *
 * look_at_rotation            utils/camera_geometry.py:3-14
 *   z = target - eye; z /= norm(z)         norm = sqrt(BLAS dot), see orc_dot3_*
 *   if allclose(|dot(z, up)|, 1): up = (0,0,1)
 *   x = cross(up, z) / norm(.) ; y = cross(z, x) ; R = rows [x; y; z]
 * np.cross is mul, mul, sub (three separately rounded ufuncs).
 * ------------------------------------------------------------------------- */
/* 3-term dot product as OpenBLAS computes it for n = 3 (scalar tail loop with a
 * double accumulator, kernel/x86_64/{d,s}dot.c): in f64 the compiler contracts
 * `dot += y*x` into an FMA chain; in f32 each product is rounded to float first,
 * summed in double and rounded to float once at the end. */
static inline double orc_dot3_f64(const double* a, const double* b) {
  return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0]));
}
static inline float orc_dot3_f32(const float* a, const float* b) {
  float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2];
  double s = (double)p0 + (double)p1;
  s = s + (double)p2;
  return (float)s;
}

#define DEFINE_LOOK_AT(NAME, T, DOT3, SQRT, FABS)                                  \
  ORC_API void NAME(const T* eye, const T* target, T* R) {                         \
    T z[3], x[3], y[3], up[3] = {0, 1, 0};                                         \
    for (int i = 0; i < 3; ++i) z[i] = target[i] - eye[i];                         \
    T n = SQRT(DOT3(z, z));                                                        \
    for (int i = 0; i < 3; ++i) z[i] = z[i] / n;                                   \
    /* np.dot(z, up): same BLAS dot */                                             \
    T dzu = DOT3(z, up);                                                           \
    /* np.allclose(a, 1.0): |a - 1| <= atol + rtol*|1|, evaluated in double */     \
    double a = (double)FABS(dzu);                                                  \
    if (fabs(a - 1.0) <= 1e-8 + 1e-5 * 1.0) { up[0] = 0; up[1] = 0; up[2] = 1; }   \
    T p, q;                                                                        \
    p = up[1] * z[2]; q = up[2] * z[1]; x[0] = p - q;                              \
    p = up[2] * z[0]; q = up[0] * z[2]; x[1] = p - q;                              \
    p = up[0] * z[1]; q = up[1] * z[0]; x[2] = p - q;                              \
    T nx = SQRT(DOT3(x, x));                                                       \
    for (int i = 0; i < 3; ++i) x[i] = x[i] / nx;                                  \
    p = z[1] * x[2]; q = z[2] * x[1]; y[0] = p - q;                                \
    p = z[2] * x[0]; q = z[0] * x[2]; y[1] = p - q;                                \
    p = z[0] * x[1]; q = z[1] * x[0]; y[2] = p - q;                                \
    for (int i = 0; i < 3; ++i) { R[i] = x[i]; R[3 + i] = y[i]; R[6 + i] = z[i]; } \
  }

DEFINE_LOOK_AT(orc_look_at_f64, double, orc_dot3_f64, sqrt, fabs)
DEFINE_LOOK_AT(orc_look_at_f32, float, orc_dot3_f32, sqrtf, fabsf)

/* ------------------------------------------------------------------------- *
 * project_colored_voxels       utils/projection_utils.py:5-23
 *   pts_cam = (pts - cam_pos) @ R.T     -> per row FMA chain k = 0,1,2 (gemm)
 *   Z = where(Z < 1e-8, 1e-8, Z)        no culling of points behind the camera
 *   u = (X/Z)*f + cx ; v = -(Y/Z)*f + cy   div, mul, add separately rounded
 *   ui, vi = rint (half-even) ; valid = inside the image
 *   img[vi, ui] = colour                last write wins, in point order
 * `pix` receives, per pixel, 1 + index of the winning point (0 = untouched),
 * which is what the CUDA z-buffer holds; img (H,W,3) may be NULL.
 * Working type T follows numpy promotion: f32 points with f64 camera -> f64;
 * all-f32 inputs -> f32 (notebooks 3/4).
 * ------------------------------------------------------------------------- */
#define DEFINE_PROJECT(NAME, T, LOOKAT, FMA, RINT, EPS)                           \
  ORC_API void NAME(const float* pts, const uint8_t* colors, int64_t n,           \
                    const T* cam_pos, const T* target, T f, T cx, T cy,            \
                    int H, int W, uint8_t* img, uint32_t* pix) {                   \
    T R[9];                                                                        \
    LOOKAT(cam_pos, target, R);                                                    \
    if (img) memset(img, 0, (size_t)H * W * 3);                                    \
    if (pix) memset(pix, 0, (size_t)H * W * sizeof(uint32_t));                     \
    for (int64_t i = 0; i < n; ++i) {                                              \
      T d0 = (T)pts[3 * i + 0] - cam_pos[0];                                       \
      T d1 = (T)pts[3 * i + 1] - cam_pos[1];                                       \
      T d2 = (T)pts[3 * i + 2] - cam_pos[2];                                       \
      T X = FMA(d2, R[2], FMA(d1, R[1], d0 * R[0]));                               \
      T Y = FMA(d2, R[5], FMA(d1, R[4], d0 * R[3]));                               \
      T Z = FMA(d2, R[8], FMA(d1, R[7], d0 * R[6]));                               \
      if (Z < (T)EPS) Z = (T)EPS;                                                  \
      T q = X / Z; T u = q * f; u = u + cx;                                        \
      T r = Y / Z; r = -r; T v = r * f; v = v + cy;                                \
      T ur = RINT(u), vr = RINT(v);                                                \
      if (!(ur >= 0 && ur < (T)W && vr >= 0 && vr < (T)H)) continue;               \
      size_t p = (size_t)((int64_t)vr * W + (int64_t)ur);                          \
      if (img) { img[3 * p] = colors[3 * i]; img[3 * p + 1] = colors[3 * i + 1];   \
                 img[3 * p + 2] = colors[3 * i + 2]; }                             \
      if (pix) pix[p] = (uint32_t)(i + 1);                                         \
    }                                                                              \
  }

DEFINE_PROJECT(orc_project_f64, double, orc_look_at_f64, fma, rint, 1e-8)
DEFINE_PROJECT(orc_project_f32, float, orc_look_at_f32, fmaf, rintf, 1e-8)

/* ------------------------------------------------------------------------- *
 * compute_partwise_iou         utils/camera_estimation.py:770-787
 *   per part colour c: inter = |proj==c & gt==c| ; union = |proj==c | gt==c|
 * Integer counts only; the float division / mean stays in the Python wrapper.
 * ------------------------------------------------------------------------- */
ORC_API void orc_partwise_counts(const uint8_t* proj, const uint8_t* gt, int64_t npix,
                                 const uint8_t* part_rgb, int P, int64_t* inter,
                                 int64_t* uni) {
  for (int p = 0; p < P; ++p) {
    const uint8_t* c = part_rgb + 3 * p;
    int64_t ni = 0, nu = 0;
    for (int64_t i = 0; i < npix; ++i) {
      int a = proj[3 * i] == c[0] && proj[3 * i + 1] == c[1] && proj[3 * i + 2] == c[2];
      int b = gt[3 * i] == c[0] && gt[3 * i + 1] == c[1] && gt[3 * i + 2] == c[2];
      ni += a & b;
      nu += a | b;
    }
    inter[p] = ni;
    uni[p] = nu;
  }
}

/* ------------------------------------------------------------------------- *
 * scipy.ndimage.affine_transform(vol_u8, M, offset, order=1, mode="constant",
 * cval=0) as called at utils/voxel_carving_utils.py:116-123 (scipy 1.18.1,
 * NI_GeometricTransform, linear spline, no prefilter).  Coordinates and
 * weights are FP64 with no contraction; accumulation order is the 8-corner
 * nest (axis0, axis1, axis2); result rounded half-up into uint8.
 * ------------------------------------------------------------------------- */
ORC_API void orc_affine_order1_u8(const uint8_t* vol, int n0, int n1, int n2,
                                  const double* M, const double* off, uint8_t* out) {
  const int dim[3] = {n0, n1, n2};
  for (int o0 = 0; o0 < n0; ++o0)
    for (int o1 = 0; o1 < n1; ++o1)
      for (int o2 = 0; o2 < n2; ++o2) {
        double cc[3];
        int inside = 1, i0[3], i1[3];
        double t[3];
        for (int h = 0; h < 3; ++h) {
          double c = off[h];
          c += M[3 * h + 0] * (double)o0;
          c += M[3 * h + 1] * (double)o1;
          c += M[3 * h + 2] * (double)o2;
          cc[h] = c;
          if (c < 0.0 || c > (double)(dim[h] - 1)) inside = 0;
        }
        uint8_t res = 0;
        if (inside) {
          for (int h = 0; h < 3; ++h) {
            double fl = floor(cc[h]);
            t[h] = cc[h] - fl;
            i0[h] = (int)fl;
            i1[h] = i0[h] + 1 < dim[h] ? i0[h] + 1 : dim[h] - 1;
          }
          double acc = 0.0;
          for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
              for (int c = 0; c < 2; ++c) {
                int ix = a ? i1[0] : i0[0], iy = b ? i1[1] : i0[1], iz = c ? i1[2] : i0[2];
                double w = (double)vol[((size_t)ix * n1 + iy) * n2 + iz];
                w *= a ? t[0] : 1.0 - t[0];
                w *= b ? t[1] : 1.0 - t[1];
                w *= c ? t[2] : 1.0 - t[2];
                acc += w;
              }
          if (acc > 0.0) {
            double r = acc + 0.5;
            res = r >= 255.0 ? 255 : (uint8_t)r;
          }
        }
        out[((size_t)o0 * n1 + o1) * n2 + o2] = res;
      }
}

/* ------------------------------------------------------------------------- *
 * scipy.ndimage.label(mask) with the default structure (6-connectivity in
 * 3-D) as called at utils/voxel_carving_utils.py:175,254.  Component ids are
 * 1..n in raster (C) order of each component's first voxel.  Union-find with
 * the smaller flat index as root, then a raster relabel.
 * ------------------------------------------------------------------------- */
static int64_t uf_find(int64_t* parent, int64_t i) {
  while (parent[i] != i) {
    parent[i] = parent[parent[i]];
    i = parent[i];
  }
  return i;
}
static void uf_union(int64_t* parent, int64_t a, int64_t b) {
  a = uf_find(parent, a);
  b = uf_find(parent, b);
  if (a < b) parent[b] = a; else if (b < a) parent[a] = b;
}

ORC_API int orc_label6(const uint8_t* mask, int n0, int n1, int n2, int32_t* labels) {
  size_t n = (size_t)n0 * n1 * n2;
  int64_t* parent = (int64_t*)malloc(n * sizeof(int64_t));
  if (!parent) return -1;
  for (size_t i = 0; i < n; ++i) parent[i] = (int64_t)i;
  for (int a = 0; a < n0; ++a)
    for (int b = 0; b < n1; ++b)
      for (int c = 0; c < n2; ++c) {
        size_t i = ((size_t)a * n1 + b) * n2 + c;
        if (!mask[i]) continue;
        if (c > 0 && mask[i - 1]) uf_union(parent, (int64_t)i, (int64_t)(i - 1));
        if (b > 0 && mask[i - n2]) uf_union(parent, (int64_t)i, (int64_t)(i - n2));
        if (a > 0 && mask[i - (size_t)n1 * n2])
          uf_union(parent, (int64_t)i, (int64_t)(i - (size_t)n1 * n2));
      }
  int32_t next = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!mask[i]) { labels[i] = 0; continue; }
    int64_t r = uf_find(parent, (int64_t)i);
    if ((size_t)r == i) labels[i] = ++next;      /* root == first voxel in raster order */
    else labels[i] = labels[r];                  /* r < i, already numbered */
  }
  free(parent);
  return next;
}

/* ------------------------------------------------------------------------- *
 * Depth-buffer visibility evaluator        utils/eval_helpers_intra.py:134-190
 *   compute_global_depth_buffer: zbuf[v,u] = min Z over the points with Z > 1e-6 that round into the image
 *   project_part_visible:        mask[v,u] = 1 if some point there has |Z - zbuf[v,u]| < eps
 * Projection arithmetic as project_colored_voxels (same NumPy expressions), but points with Z <= 1e-6 are culled
 * instead of clamped.  zbuf is float32 in the reference; with float64 cameras each accepted Z is rounded to float32
 * on store, and the final value is float32(min Z) whatever the order (rounding is monotone).
 * ------------------------------------------------------------------------- */
#define DEFINE_DEPTH(NAME_Z, NAME_V, T, LOOKAT, FMA, RINT, FABS)                                     \
  static int NAME_Z##_pix(const float* p, const T* cam_pos, const T* R, T f, T cx, T cy, int H,      \
                          int W, T* zout) {                                                          \
    T d0 = (T)p[0] - cam_pos[0], d1 = (T)p[1] - cam_pos[1], d2 = (T)p[2] - cam_pos[2];               \
    T X = FMA(d2, R[2], FMA(d1, R[1], d0 * R[0]));                                                   \
    T Y = FMA(d2, R[5], FMA(d1, R[4], d0 * R[3]));                                                   \
    T Z = FMA(d2, R[8], FMA(d1, R[7], d0 * R[6]));                                                   \
    if (!(Z > (T)1e-6)) return -1;                                                                   \
    T q = X / Z; T u = q * f; u = u + cx;                                                            \
    T r = Y / Z; r = -r; T v = r * f; v = v + cy;                                                    \
    T ur = RINT(u), vr = RINT(v);                                                                    \
    if (!(ur >= 0 && ur < (T)W && vr >= 0 && vr < (T)H)) return -1;                                  \
    *zout = Z;                                                                                       \
    return (int)vr * W + (int)ur;                                                                    \
  }                                                                                                  \
  ORC_API void NAME_Z(const float* pts, int64_t n, const T* cam_pos, const T* target, T f, T cx,     \
                      T cy, int H, int W, float* zbuf) {                                             \
    T R[9];                                                                                          \
    LOOKAT(cam_pos, target, R);                                                                      \
    for (int i = 0; i < H * W; ++i) zbuf[i] = INFINITY;                                              \
    for (int64_t i = 0; i < n; ++i) {                                                                \
      T z;                                                                                           \
      int p = NAME_Z##_pix(pts + 3 * i, cam_pos, R, f, cx, cy, H, W, &z);                            \
      if (p >= 0 && z < (T)zbuf[p]) zbuf[p] = (float)z;                                              \
    }                                                                                                \
  }                                                                                                  \
  ORC_API void NAME_V(const float* pts, int64_t n, const T* cam_pos, const T* target, T f, T cx,     \
                      T cy, const float* zbuf, T eps, int H, int W, uint8_t* mask) {                 \
    T R[9];                                                                                          \
    LOOKAT(cam_pos, target, R);                                                                      \
    memset(mask, 0, (size_t)H * W);                                                                  \
    for (int64_t i = 0; i < n; ++i) {                                                                \
      T z;                                                                                           \
      int p = NAME_Z##_pix(pts + 3 * i, cam_pos, R, f, cx, cy, H, W, &z);                            \
      if (p >= 0) { T d = z - (T)zbuf[p]; if (FABS(d) < eps) mask[p] = 1; }                          \
    }                                                                                                \
  }

DEFINE_DEPTH(orc_depth_buffer_f32, orc_part_visible_f32, float, orc_look_at_f32, fmaf, rintf, fabsf)
DEFINE_DEPTH(orc_depth_buffer_f64, orc_part_visible_f64, double, orc_look_at_f64, fma, rint, fabs)
